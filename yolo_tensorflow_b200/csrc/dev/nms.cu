// nms.cu — class-wise non-maximum suppression (do_nms_sort, box.c:58-89) and do_nms_obj (box.c:21-55) on device.
//
// Reference algorithm, per class k: order the live detections by prob[k] descending (qsort with nms_comparator,
// box.c:6-19), then greedily, for each i with prob[k] != 0, zero prob[k] of every later j whose
// box_iou(i,j) > thresh (strict, fp32, box.c:152-182).  Classes are independent (class k only reads/writes
// prob[k] and the boxes), so every (image, class) pair is one unit of work for one CTA (a 64-thread CTA with
// everything in shared memory when the class has <= 128 survivors, the general kernel below otherwise):
//   1. ordered gather of the detections with prob[k] != 0 (warp-ballot compaction keeps the original order),
//   2. rank sort, descending, ties broken by original index (a stable order; the reference's qsort leaves tie
//      order unspecified, and entries with prob 0 can never suppress or be suppressed so they are not sorted),
//   3. bitmask-IoU: bit j of row i says "i suppresses j" (j > i), 32 IoUs per thread per word, rows kept in
//      shared memory when the class has <= NMS_SMEM_M survivors and in a per-CTA HBM slab otherwise,
//   4. one warp scans the rows in order, OR-ing the row of every still-alive box into the removed set,
//   5. removed boxes get prob[k] = 0.
// The IoU uses explicit round-to-nearest intrinsics in the reference's operation order so no FMA contraction can
// change a comparison: keep-lists are bit-exact against the CPU path when fed the same boxes.
#include "kernels.h"

#define NMS_THREADS 1024                     // the general kernel serves the few very heavy classes (thousands of survivors): one unit per CTA, so CTA width is what parallelises it
#define NMS_SMEM_M 512                       // survivors per class whose IoU bit-matrix lives in shared memory
#define host_smem_work_bytes (512 * 32 + 512 * 16 * 4 + 64)   // == host_work_bytes(NMS_SMEM_M): size of the smem work area

// box_iou(a,b) > thresh (box.c:152-182), bit-exact, on PRECOMPUTED per-box values: corners (l, r, t, b) = (x - w/2, x + w/2,
// y - h/2, y + h/2) and area = w*h, each rounded exactly as the reference rounds them inside overlap() / box_union() (w/2 ==
// w*0.5f exactly in binary floating point), so the pair test costs ~20 instructions.  Boxes that do not intersect are decided
// without a division: the reference computes 0/union = 0 (or 0/0 = NaN) there, and neither exceeds a non-negative threshold.
// Otherwise box_iou > thresh <=> inter > thresh * union for a positive finite union: the product decides every pair that is
// not within 1e-5 of the threshold, the rest (and every non-finite case: the NMS stress configuration has boxes of infinite
// size) take the reference's own IEEE division.
struct NmsBox { float l, r, t, b; };

// the same decision straight from (x, y, w, h) records: what the small-unit kernel uses (one 16-byte shared-memory read per
// box; measured there, the extra read of a precomputed area costs more than the arithmetic it saves)
__device__ __forceinline__ bool suppresses_xywh(float4 a, float4 b, float thresh)
{
    const float ahw = __fmul_rn(a.z, 0.5f), bhw = __fmul_rn(b.z, 0.5f), ahh = __fmul_rn(a.w, 0.5f), bhh = __fmul_rn(b.w, 0.5f);
    const float l1 = __fsub_rn(a.x, ahw), l2 = __fsub_rn(b.x, bhw), r1 = __fadd_rn(a.x, ahw), r2 = __fadd_rn(b.x, bhw);
    const float t1 = __fsub_rn(a.y, ahh), t2 = __fsub_rn(b.y, bhh), b1 = __fadd_rn(a.y, ahh), b2 = __fadd_rn(b.y, bhh);
    const float w = __fsub_rn(r1 < r2 ? r1 : r2, l1 > l2 ? l1 : l2), h = __fsub_rn(b1 < b2 ? b1 : b2, t1 > t2 ? t1 : t2);
    if (thresh >= 0.f && (w < 0 || h < 0 || w == 0.f || h == 0.f)) return false;
    const float inter = (w < 0 || h < 0) ? 0.f : __fmul_rn(w, h);
    const float uni = __fsub_rn(__fadd_rn(__fmul_rn(a.z, a.w), __fmul_rn(b.z, b.w)), inter);
    // the fast approximate quotient (<= 2 ulp off) decides every pair that is not within a hair of the threshold;
    // only those few take the IEEE division the reference performs, so the decision is still bit-exact
    const float q = __fdividef(inter, uni);
    if (q > thresh * 1.00001f + 1e-30f && q < 3.0e38f) return true;
    if (q < thresh * 0.99999f - 1e-30f) return false;
    return __fdiv_rn(inter, uni) > thresh;
}

__device__ __forceinline__ NmsBox nms_corners(float4 v)          // v = (x, y, w, h)
{
    const float hw = __fmul_rn(v.z, 0.5f), hh = __fmul_rn(v.w, 0.5f);
    NmsBox c;
    c.l = __fsub_rn(v.x, hw); c.r = __fadd_rn(v.x, hw);
    c.t = __fsub_rn(v.y, hh); c.b = __fadd_rn(v.y, hh);
    return c;
}

__device__ __forceinline__ bool suppresses_c(NmsBox a, float area_a, NmsBox b, float area_b, float thresh)
{
    const float left = a.l > b.l ? a.l : b.l, right = a.r < b.r ? a.r : b.r;       // the reference's ternaries (NaN falls to b)
    const float top = a.t > b.t ? a.t : b.t, bottom = a.b < b.b ? a.b : b.b;
    const float w = __fsub_rn(right, left), h = __fsub_rn(bottom, top);
    if (thresh >= 0.f && (w < 0 || h < 0 || w == 0.f || h == 0.f)) return false;   // 0 / union (or 0 / 0) never exceeds thresh >= 0
    const float inter = (w < 0 || h < 0) ? 0.f : __fmul_rn(w, h);
    const float uni = __fsub_rn(__fadd_rn(area_a, area_b), inter);
    if (uni > 0.f && uni < 3.0e38f && inter < 3.0e38f) {
        const float tu = __fmul_rn(thresh, uni);
        if (inter > __fmaf_rn(fabsf(tu), 1e-5f, tu) + 1e-30f) return true;
        if (inter < __fmaf_rn(fabsf(tu), -1e-5f, tu) - 1e-30f) return false;
    }
    return __fdiv_rn(inter, uni) > thresh;
}

// layout of the per-unit work area (either dynamic shared memory or the CTA's HBM slab)
struct NmsWork {
    float *score;      // [m] gathered scores, original order
    int *src;          // [m] original detection index, original order
    int *order;        // [m] original detection index, sorted order
    float4 *sbox;      // [m] box corners (l, r, t, b), sorted order
    float *area;       // [m] w*h, sorted order
    unsigned *mask;    // [m][words]
};

__device__ __forceinline__ NmsWork carve(unsigned char *base, int m)
{
    NmsWork w;
    size_t off = 0;
    w.sbox = (float4 *)(base + off); off += (size_t)m * 16;
    w.score = (float *)(base + off); off += (size_t)m * 4;
    w.src = (int *)(base + off); off += (size_t)m * 4;
    w.order = (int *)(base + off); off += (size_t)m * 4;
    w.area = (float *)(base + off); off += (size_t)m * 4;
    off = (off + 15) & ~(size_t)15;
    w.mask = (unsigned *)(base + off);
    return w;
}

// score(i) = base[i*stride + k]; live(i) = obj[i] != 0 (NULL obj => all live)
__global__ void __launch_bounds__(NMS_THREADS)
nms_kernel(const float *__restrict__ box, float *__restrict__ score, const float *__restrict__ obj,
           const int *__restrict__ count, int images, int cap, int classes, int score_stride, float thresh,
           unsigned char *__restrict__ slab, size_t slab_bytes, unsigned char *__restrict__ suppressed_out,
           const int *__restrict__ cls_count, int small_limit)
{
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int warp_totals[NMS_THREADS / 32];
    unsigned *removed = (unsigned *)smem;       // bitset of suppressed survivors (see below)
    unsigned char *smem_work = smem + 4096;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    __shared__ int s_unit;
    int *next_unit = const_cast<int *>(cls_count) + images * classes;      // work counter behind the class counts (zeroed by the launcher)
    for (;;) {
        // dynamic claiming: heavy classes cluster on a few class ids, a static stride would pile them on the same CTAs
        __syncthreads();
        if (threadIdx.x == 0) {
            int u;
            do { u = atomicAdd(next_unit, 1); } while (u < images * classes && cls_count[u] <= small_limit);
            s_unit = u;
        }
        __syncthreads();
        const int unit = s_unit;
        if (unit >= images * classes) break;
        const int img = unit / classes, k = unit % classes;
        const int n = count ? count[img] : cap;
        const float *ubox = box + (size_t)img * cap * 4;
        float *uscore = score + (size_t)img * cap * score_stride + k;
        const float *uobj = obj ? obj + (size_t)img * cap : nullptr;

        // ---- survivors of this class (counted by class_count_kernel); units the small kernel handles are skipped
        int m = cls_count[unit];
        if (m <= small_limit) continue;
        if (m <= 1) continue;                                   // nothing can be suppressed
        // m <= NMS_SMEM_M: lists AND the whole IoU bit-matrix live in shared memory.  Larger units keep only the lists — in shared
        // memory when they fit (32 bytes per survivor), else in this CTA's HBM slab — and never materialise the m x m matrix:
        // its rows are produced a chunk at a time into a double-buffered shared-memory stage (below).  The removed bitset sits
        // at the front of dynamic smem (4 KB = 32768 survivors) and moves behind the lists in the slab for anything larger.
        const int words = (m + 31) / 32;
        unsigned char *slab_cta = slab + (size_t)blockIdx.x * slab_bytes;
        size_t lists_bytes = 0;                                  // bytes of the smem work area taken by the lists (chunked path)
        NmsWork wk;
        if (m <= NMS_SMEM_M) wk = carve(smem_work, m);
        else if ((size_t)m * 32 + 64 + 16384 <= (size_t)host_smem_work_bytes) {
            wk = carve(smem_work, m);
            lists_bytes = ((size_t)m * 32 + 64 + 15) & ~(size_t)15;
        } else {
            if ((size_t)m * 32 + 64 + (size_t)words * 4 > slab_bytes) {     // cannot happen: host sizes the slab from max_count
                if (threadIdx.x == 0) printf("b200-darknet: nms slab too small (m=%d)\n", m);
                __trap();
            }
            wk = carve(slab_cta, m);
        }
        if (m > 32768) removed = (unsigned *)(slab_cta + (((size_t)m * 32 + 64 + 15) & ~(size_t)15));
        else removed = (unsigned *)smem;

        // ---- pass B: ordered gather
        int written = 0;
        for (int i0 = 0; i0 < n; i0 += NMS_THREADS) {
            int i = i0 + threadIdx.x;
            float sc = i < n ? uscore[(size_t)i * score_stride] : 0.f;
            bool f = i < n && sc != 0.f && (!uobj || uobj[i] != 0.f);
            unsigned b = __ballot_sync(0xffffffffu, f);
            if (lane == 0) warp_totals[warp] = __popc(b);
            __syncthreads();
            int before = 0, tot = 0;
            for (int w = 0; w < NMS_THREADS / 32; ++w) { int c = warp_totals[w]; if (w < warp) before += c; tot += c; }
            if (f) { int d = written + before + __popc(b & ((1u << lane) - 1)); wk.score[d] = sc; wk.src[d] = i; }
            written += tot;
            __syncthreads();
        }

        // ---- rank sort (descending score, ties by original order)
        for (int i = threadIdx.x; i < m; i += NMS_THREADS) {
            float si = wk.score[i];
            int rank = 0;
            for (int j = 0; j < m; ++j) {
                float sj = wk.score[j];
                rank += (sj > si) || (sj == si && j < i);
            }
            int d = wk.src[i];
            wk.order[rank] = d;
            const float4 v = *reinterpret_cast<const float4 *>(ubox + (size_t)d * 4);
            const NmsBox c = nms_corners(v);
            wk.sbox[rank] = make_float4(c.l, c.r, c.t, c.b);
            wk.area[rank] = __fmul_rn(v.z, v.w);
        }
        for (int w = threadIdx.x; w < words; w += NMS_THREADS) removed[w] = 0u;
        __syncthreads();

        // ---- IoU bit-matrix + greedy scan.  Word (i, wj) of the matrix covers j in [32*wj, 32*wj+32); only j > i matters.  A word is
        // produced by one WARP: lane b tests the pair (i, 32*wj + b) — box i is a broadcast read, boxes j are 32 consecutive
        // entries (no bank conflicts) — and a ballot assembles it.  The scan walks the rows in order, OR-ing the row of every
        // still-alive box into the removed set.
        auto pair_word = [&](int i, int wj) -> unsigned {
            const int j = wj * 32 + lane;
            bool sup = false;
            if (j > i && j < m) {
                const float4 av = wk.sbox[i], bv = wk.sbox[j];
                sup = suppresses_c(NmsBox{av.x, av.y, av.z, av.w}, wk.area[i], NmsBox{bv.x, bv.y, bv.z, bv.w}, wk.area[j], thresh);
            }
            return __ballot_sync(0xffffffffu, sup);
        };
        if (m <= NMS_SMEM_M) {
            // whole matrix in shared memory (a warp takes 32 consecutive words at a time so that the store is one line), then the
            // scan with the removed set in one register per lane (words <= 16) and row i+1 fetched while row i is decided
            const int total_words = m * words;
            for (int t0 = warp * 32; t0 < total_words; t0 += (NMS_THREADS / 32) * 32) {
                unsigned mine = 0u;
                const int kmax = (total_words - t0) < 32 ? (total_words - t0) : 32;
                for (int k = 0; k < kmax; ++k) {
                    const int t = t0 + k, i = t / words, wj = t % words;
                    const unsigned bits = (wj * 32 + 31 > i) ? pair_word(i, wj) : 0u;       // warp-uniform condition
                    if (lane == k) mine = bits;
                }
                if (lane < kmax) wk.mask[t0 + lane] = mine;
            }
            __syncthreads();
            if (warp == 0) {
                unsigned rem = 0u;
                unsigned nxt = lane < words ? wk.mask[lane] : 0u;
                for (int i = 0; i < m; ++i) {
                    const unsigned cur = nxt;
                    if (i + 1 < m) nxt = lane < words ? wk.mask[(size_t)(i + 1) * words + lane] : 0u;
                    const unsigned r = __shfl_sync(0xffffffffu, rem, i >> 5);
                    if (!((r >> (i & 31)) & 1u)) rem |= cur;
                }
                if (lane < words) removed[lane] = rem;
            }
            __syncthreads();
        } else {
            // chunked: R rows of the matrix at a time, in one of two shared-memory buffers; while warp 0 scans chunk c, warps 1-31
            // already produce chunk c+1 (the rows do not depend on the scan), so the serial scan hides behind the pair tests and
            // no m x m slab exists (round 1 sized one from the anchor count: 19 GB at 608 x 608)
            unsigned *stage = (unsigned *)(smem_work + lists_bytes);
            const int half_words = (int)(((size_t)host_smem_work_bytes - lists_bytes) / 8);     // words per buffer
            int R = half_words / words;
            if (R > 256) R = 256;
            if (R < 1) {                                            // a single row does not fit (m > ~300k): not reachable with cap checks on the host
                if (threadIdx.x == 0) printf("b200-darknet: nms unit too large (m=%d)\n", m);
                __trap();
            }
            const int nchunks = (m + R - 1) / R;
            auto produce = [&](int c, int first_warp) {             // rows [c*R, c*R+rows) x words [w_lo, words) -> buffer c & 1
                const int i0 = c * R, rows = (m - i0) < R ? (m - i0) : R;
                const int w_lo = i0 >> 5, span = words - w_lo;
                unsigned *buf = stage + (size_t)(c & 1) * half_words;
                const int nw = NMS_THREADS / 32 - first_warp;
                for (int t = warp - first_warp; t < rows * span; t += nw) {
                    const int rr = t / span, wj = w_lo + t % span, i = i0 + rr;
                    const unsigned bits = (wj * 32 + 31 > i) ? pair_word(i, wj) : 0u;
                    if (lane == 0) buf[rr * span + (wj - w_lo)] = bits;
                }
            };
            for (int w = threadIdx.x; w < words; w += NMS_THREADS) removed[w] = 0u;
            produce(0, 0);
            __syncthreads();
            unsigned rem = 0u;                                      // words <= 32: removed set in registers of warp 0
            for (int c = 0; c < nchunks; ++c) {
                if (warp > 0) {
                    if (c + 1 < nchunks) produce(c + 1, 1);
                } else {
                    const int i0 = c * R, rows = (m - i0) < R ? (m - i0) : R;
                    const int w_lo = i0 >> 5, span = words - w_lo;
                    const unsigned *buf = stage + (size_t)(c & 1) * half_words;
                    if (words <= 32) {
                        unsigned nxt = (lane >= w_lo && lane < words) ? buf[lane - w_lo] : 0u;
                        for (int rr = 0; rr < rows; ++rr) {
                            const int i = i0 + rr;
                            const unsigned cur = nxt;
                            if (rr + 1 < rows) nxt = (lane >= w_lo && lane < words) ? buf[(rr + 1) * span + (lane - w_lo)] : 0u;
                            const unsigned r = __shfl_sync(0xffffffffu, rem, i >> 5);
                            if (!((r >> (i & 31)) & 1u)) rem |= cur;
                        }
                    } else {
                        for (int rr = 0; rr < rows; ++rr) {
                            const int i = i0 + rr;
                            const unsigned r = removed[i >> 5];
                            if (!((r >> (i & 31)) & 1u))
                                for (int w = (i >> 5) + lane; w < words; w += 32) removed[w] |= buf[rr * span + (w - w_lo)];
                            __syncwarp();
                        }
                    }
                }
                __syncthreads();
            }
            if (words <= 32 && warp == 0 && lane < words) removed[lane] = rem;
            __syncthreads();
        }

        // ---- write back
        for (int i = threadIdx.x; i < m; i += NMS_THREADS) {
            if ((removed[i >> 5] >> (i & 31)) & 1u) {
                int d = wk.order[i];
                uscore[(size_t)d * score_stride] = 0.f;
                if (suppressed_out) suppressed_out[(size_t)img * cap + d] = 1;
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------
// scheduling pre-pass: live detections with a non-zero score, per (image, class)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
class_count_kernel(const float *__restrict__ score, const float *__restrict__ obj, const int *__restrict__ count, int cap,
                   int classes, int score_stride, int *__restrict__ cls_count)
{
    // grid = (slices, images); cls_count was zeroed by the launcher
    extern __shared__ int sm_counts[];
    const int img = blockIdx.y;
    const int n = count ? count[img] : cap;
    for (int k = threadIdx.x; k < classes; k += blockDim.x) sm_counts[k] = 0;
    __syncthreads();
    const long long total = (long long)n * classes;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        int i = (int)(e / classes), k = (int)(e % classes);
        float sc = score[((size_t)img * cap + i) * score_stride + k];
        if (sc != 0.f && (!obj || obj[(size_t)img * cap + i] != 0.f)) atomicAdd(&sm_counts[k], 1);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < classes; k += blockDim.x)
        if (sm_counts[k]) atomicAdd(&cls_count[(size_t)img * classes + k], sm_counts[k]);
}

// ---------------------------------------------------------------------------------------------------
// small units (2 <= m <= NMS_SMALL_M survivors in the class): one 128-thread CTA per (image, class), everything in
// ~15 KB of shared memory, so thousands of units are in flight at once (the work is latency-bound, not throughput-bound).
// Same algorithm and the same bit-exact IoU as nms_kernel.
// ---------------------------------------------------------------------------------------------------
#define NMS_SMALL_M 256
#define NMS_SMALL_THREADS 128

__global__ void __launch_bounds__(NMS_SMALL_THREADS)
nms_small_kernel(const float *__restrict__ box, float *__restrict__ score, const float *__restrict__ obj,
                 const int *__restrict__ count, int cap, int classes, int score_stride, float thresh,
                 const int *__restrict__ cls_count, unsigned char *__restrict__ suppressed_out)
{
    const int unit = blockIdx.x;
    const int m = cls_count[unit];
    if (m <= 1 || m > NMS_SMALL_M) return;
    __shared__ float s_score[NMS_SMALL_M];
    __shared__ int s_src[NMS_SMALL_M], s_order[NMS_SMALL_M];
    __shared__ float4 s_box[NMS_SMALL_M];
    __shared__ unsigned s_mask[NMS_SMALL_M][NMS_SMALL_M / 32];
    __shared__ unsigned s_removed[NMS_SMALL_M / 32];
    __shared__ int s_warp_tot[NMS_SMALL_THREADS / 32];
    const int img = unit / classes, k = unit % classes;
    const int n = count ? count[img] : cap;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float *ubox = box + (size_t)img * cap * 4;
    float *uscore = score + (size_t)img * cap * score_stride + k;
    const float *uobj = obj ? obj + (size_t)img * cap : nullptr;

    int written = 0;                                            // ordered gather
    for (int i0 = 0; i0 < n; i0 += NMS_SMALL_THREADS) {
        int i = i0 + threadIdx.x;
        float sc = i < n ? uscore[(size_t)i * score_stride] : 0.f;
        bool f = i < n && sc != 0.f && (!uobj || uobj[i] != 0.f);
        unsigned b = __ballot_sync(0xffffffffu, f);
        if (lane == 0) s_warp_tot[warp] = __popc(b);
        __syncthreads();
        int before = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < NMS_SMALL_THREADS / 32; ++w) { int c = s_warp_tot[w]; if (w < warp) before += c; tot += c; }
        if (f) { int d = written + before + __popc(b & ((1u << lane) - 1)); if (d < NMS_SMALL_M) { s_score[d] = sc; s_src[d] = i; } }
        written += tot;
        __syncthreads();
    }
    for (int i = threadIdx.x; i < m; i += NMS_SMALL_THREADS) {  // rank sort: descending score, ties by original order
        float si = s_score[i];
        int rank = 0;
        for (int j = 0; j < m; ++j) { float sj = s_score[j]; rank += (sj > si) || (sj == si && j < i); }
        int d = s_src[i];
        s_order[rank] = d;
        s_box[rank] = *reinterpret_cast<const float4 *>(ubox + (size_t)d * 4);
    }
    if (threadIdx.x < NMS_SMALL_M / 32) s_removed[threadIdx.x] = 0u;
    __syncthreads();
    const int words = (m + 31) / 32;
    // one thread per word: its 32 pair tests are independent, which is what hides the shared-memory latency in these small
    // CTAs (measured: a warp per word with a ballot, the mapping of the general kernel, is 30 % slower here)
    for (int t = threadIdx.x; t < m * words; t += NMS_SMALL_THREADS) {
        const int i = t / words, wj = t % words;
        unsigned bits = 0u;
        if (wj * 32 + 31 > i) {
            const float4 a = s_box[i];
            for (int b = 0; b < 32; ++b) {
                const int j = wj * 32 + b;
                if (j > i && j < m && suppresses_xywh(a, s_box[j], thresh)) bits |= 1u << b;
            }
        }
        s_mask[i][wj] = bits;
    }
    __syncthreads();
    if (warp == 0) {                                            // greedy scan, removed set in one register per lane < words
        unsigned rem = 0u;
        for (int i = 0; i < m; ++i) {
            unsigned r = __shfl_sync(0xffffffffu, rem, i >> 5);
            if (!((r >> (i & 31)) & 1u) && lane < words) rem |= s_mask[i][lane];
        }
        if (lane < words) s_removed[lane] = rem;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < m; i += NMS_SMALL_THREADS) {
        if ((s_removed[i >> 5] >> (i & 31)) & 1u) {
            int d = s_order[i];
            uscore[(size_t)d * score_stride] = 0.f;
            if (suppressed_out) suppressed_out[(size_t)img * cap + d] = 1;
        }
    }
}

static size_t host_work_bytes(int m)
{
    size_t words = (m + 31) / 32;
    return (size_t)m * (4 + 4 + 4 + 4 + 16) + (size_t)m * words * 4 + 64;
}

// per-CTA HBM slab of the general kernel: the LISTS of a unit too large for shared memory (32 bytes per survivor, > 1024
// survivors) and, beyond 32768 survivors, its removed bitset.  The IoU matrix is never materialised (chunked in shared memory),
// so the slab is linear in the anchor count: 0.73 MB per CTA at 608 x 608 (22 743 anchors), 216 MB for 296 CTAs.
static size_t slab_bytes_for(int max_count)
{
    if (max_count <= 1024) return 0;
    const size_t lists = ((size_t)max_count * 32 + 64 + 15) & ~(size_t)15;
    const size_t bitset = ((size_t)(max_count + 31) / 32) * 4;
    return (lists + bitset + 255) / 256 * 256;
}

static void ensure_scratch(NmsScratch *sc, int max_count, int ctas)
{
    const size_t need = slab_bytes_for(max_count);
    if (need * ctas > sc->words_per_cta * (size_t)sc->ctas || (need && !sc->mask)) {
        if (sc->mask) B200_CHECK(cudaFree(sc->mask));
        B200_CHECK(cudaMalloc((void **)&sc->mask, need * ctas));
    }
    sc->words_per_cta = need;   // bytes per CTA (name kept for the header)
    sc->ctas = ctas;
}

static void run_nms(const float *box, float *score, const float *obj, const int *count, int images, int cap, int classes,
                    int stride, float thresh, int max_count, NmsScratch *scratch, unsigned char *supp, int *cls_count, cudaStream_t s)
{
    int units = images * classes;
    if (units < 1) return;
    B200_CHECK(cudaMemsetAsync(cls_count, 0, ((size_t)units + 1) * sizeof(int), s));      // class counts + the work counter
    class_count_kernel<<<dim3(8, images), 256, classes * sizeof(int), s>>>(score, obj, count, cap, classes, stride, cls_count);
    B200_LAUNCHED();
    nms_small_kernel<<<units, NMS_SMALL_THREADS, 0, s>>>(box, score, obj, count, cap, classes, stride, thresh, cls_count, supp);
    B200_LAUNCHED();
    if (max_count <= NMS_SMALL_M) return;                       // no class can exceed the small kernel's capacity
    int ctas = units < 148 * 2 ? units : 148 * 2;
    ensure_scratch(scratch, max_count, ctas);
    size_t smem = 4096 + host_work_bytes(NMS_SMEM_M);
    // per launch, not once per process: the attribute belongs to the CURRENT device, and a process may hold networks on several
    B200_CHECK(cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    nms_kernel<<<ctas, NMS_THREADS, smem, s>>>(box, score, obj, count, images, cap, classes, stride, thresh,
                                               (unsigned char *)scratch->mask, scratch->words_per_cta, supp, cls_count, NMS_SMALL_M);
    B200_LAUNCHED();
}

void launch_nms_sort(const float *box, float *prob, const float *obj, const int *count, int images, int cap,
                     int classes, float thresh, int max_count, NmsScratch *scratch, int *cls_count, cudaStream_t s)
{
    run_nms(box, prob, obj, count, images, cap, classes, classes, thresh, max_count, scratch, nullptr, cls_count, s);
}

// do_nms_obj: one class-agnostic pass ordered by objectness; a suppressed detection loses objectness and all probs
__global__ void zero_suppressed_kernel(const unsigned char *__restrict__ supp, float *__restrict__ prob,
                                       const int *__restrict__ count, int images, int cap, int classes)
{
    long long total = (long long)images * cap * classes;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        long long d = t / classes;
        int img = (int)(d / cap), i = (int)(d % cap);
        if (i < (count ? count[img] : cap) && supp[d]) prob[t] = 0.f;
    }
}

void launch_nms_obj(const float *box, float *obj, float *prob, const int *count, int images, int cap, int classes,
                    float thresh, int max_count, NmsScratch *scratch, cudaStream_t s)
{
    unsigned char *supp = nullptr;
    int *cls_count = nullptr;
    B200_CHECK(cudaMallocAsync((void **)&supp, (size_t)images * cap, s));
    B200_CHECK(cudaMallocAsync((void **)&cls_count, ((size_t)images + 1) * sizeof(int), s));
    B200_CHECK(cudaMemsetAsync(supp, 0, (size_t)images * cap, s));
    run_nms(box, obj, nullptr, count, images, cap, 1, 1, thresh, max_count, scratch, supp, cls_count, s);
    if (prob) {
        long long total = (long long)images * cap * classes;
        int grid = (int)((total + 255) / 256);
        if (grid > 148 * 8) grid = 148 * 8;
        zero_suppressed_kernel<<<grid, 256, 0, s>>>(supp, prob, count, images, cap, classes);
        B200_LAUNCHED();
    }
    B200_CHECK(cudaFreeAsync(supp, s));
    B200_CHECK(cudaFreeAsync(cls_count, s));
}

// ---------------------------------------------------------------------------------------------------
// collect: surviving (detection, class) pairs -> compact records (one atomic slot per warp-ballot group)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
collect_kernel(const float *__restrict__ box, const float *__restrict__ prob, const float *__restrict__ obj,
               const int *__restrict__ id, const int *__restrict__ count, int images, int cap, int classes,
               DetRecord *__restrict__ out, int max_out, int *__restrict__ out_count, int image_base)
{
    // grid = (slices, images): only the count[img] live candidates of an image are visited
    const int img = blockIdx.y, lane = threadIdx.x & 31;
    const long long total = (long long)count[img] * classes;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long t0 = (long long)blockIdx.x * blockDim.x; t0 < total; t0 += step) {
        long long t = t0 + threadIdx.x;
        bool f = false;
        int i = 0, k = 0;
        float p = 0.f;
        if (t < total) {
            i = (int)(t / classes); k = (int)(t % classes);
            size_t d = (size_t)img * cap + i;
            if (obj[d] != 0.f) { p = prob[d * classes + k]; f = p != 0.f; }
        }
        unsigned b = __ballot_sync(0xffffffffu, f);
        int base = 0;
        if (lane == 0 && b) base = atomicAdd(out_count, __popc(b));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (f) {
            int slot = base + __popc(b & ((1u << lane) - 1));
            if (slot < max_out) {
                size_t d = (size_t)img * cap + i;
                DetRecord r;
                r.image = img + image_base; r.cls = k; r.box_id = id[d]; r.prob = p; r.objectness = obj[d];
                r.x = box[d * 4 + 0]; r.y = box[d * 4 + 1]; r.w = box[d * 4 + 2]; r.h = box[d * 4 + 3];
                out[slot] = r;
            }
        }
    }
}

void launch_collect(const float *box, const float *prob, const float *obj, const int *id, const int *count, int images,
                    int cap, int classes, DetRecord *out, int max_out, int *out_count, cudaStream_t s, int image_base)
{
    if (images < 1) return;
    dim3 grid(8, images);
    collect_kernel<<<grid, 256, 0, s>>>(box, prob, obj, id, count, images, cap, classes, out, max_out, out_count, image_base);
    B200_LAUNCHED();
}
