// conv_stem_tc.cu — the network's first convolution (3 input channels, 3x3, stride 1, 16 or 32 filters) on tcgen05.
//
// What it replaces: forward_convolutional_layer for layer 0 (convolutional_layer.c:445-485: im2col_cpu + gemm_nn +
// batch-norm + leaky) fused with the input's layout / precision conversion.  The CUDA-core stem spends 27 x C_out FMAs
// per pixel and is FMA-bound (0.74 ms for 64 x 416 x 416, 26 TFLOP/s); the layer's traffic (133 MB fp32 in, 709 MB bf16
// out) is worth ~0.13 ms of HBM time.  Here the 27-tap window of every pixel is gathered straight from the fp32 NCHW input
// into a K-major, 64B-swizzled [128 pixels x 32] bf16 tile in shared memory (K = 27 padded to 32: an im2col that never
// leaves the SM), ONE pair of tcgen05.mma (M = 128, N = C_out, K = 2 x 16) per tile does the arithmetic, and the epilogue
// applies folded batch-norm + leaky and writes NHWC bf16.
//
// Warp roles (416 threads, persistent over 32 x 4 pixel tiles): warp 0 = MMA issuer + TMEM owner, warps 1-8 = two
// gather groups taking alternate tiles (hides the global-load latency of the window gather), warps 9-12 = epilogue.
// Everything is handed over through mbarriers: gather -> afull -> MMA -> (aempty, tfull) -> epilogue -> tempty.
//
// Roofline: HBM.  Algorithmic bytes per pixel = 3 x 4 (fp32 in) + C_out x 2 (bf16 out).
#include "kernels.h"
#include "tc_ptx.cuh"

namespace {

constexpr int kTW = 32, kTH = 4;            // pixel tile (128 GEMM rows); a warp = 32 consecutive pixels of one row: the window gather
                                            // reads consecutive shared-memory words (16 x 8 tiles gave 2-way bank conflicts)
constexpr int kSlots = 8;                   // A-tile ring (TMA-fed kernel)
constexpr int kAcc = 8;                     // TMEM accumulator ring (TMA-fed kernel)
constexpr int kSlots1 = 4, kAcc1 = 4;       // the same for the LDG-gather fallback (static shared memory)
constexpr int kThreads = 13 * 32;

struct StemTcArgs {
    const float *in;                        // fp32 NCHW
    bf16 *out;                              // bf16 NHWC, row pitch ldo
    const bf16 *w;                          // [NOUT][27], K order (ky, kx, c)
    const float *scale, *shift;
    int N, H, W, ldo, act;
    int tiles_x, tiles_y, num_tiles;
    int pack;                               // TMA kernel: pixels per 128-byte output row (dense outputs), 1 = one pixel per row
    int pool;                               // TMA kernel: 1 = the store warp 2x2-max-pools every tile and stores only the pooled tensor
};

struct Walk {                               // tile = first, first + step, ... -> (tx, ty, n) without a division per tile
    int tx, ty, tn, sx, sy, sn;
    __device__ __forceinline__ void init(int first, int step, int tiles_x, int tiles_y)
    {
        tx = first % tiles_x; ty = (first / tiles_x) % tiles_y; tn = first / (tiles_x * tiles_y);
        sx = step % tiles_x;  sy = (step / tiles_x) % tiles_y;  sn = step / (tiles_x * tiles_y);
    }
    __device__ __forceinline__ void next(int tiles_x, int tiles_y)
    {
        tx += sx; if (tx >= tiles_x) { tx -= tiles_x; ++ty; }
        ty += sy; if (ty >= tiles_y) { ty -= tiles_y; ++tn; }
        tn += sn;
    }
};

#define STEM_WAIT(bar, parity)                                                                                          \
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"        \
                 "@p bra WAIT_DONE;\n\tbra WAIT_LOOP;\n\tWAIT_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"((uint32_t)(parity)) : "memory")

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi)
{
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&v);
}

template <int NOUT>
__global__ void __launch_bounds__(kThreads, 1)
conv_stem_tc_kernel(const StemTcArgs a)
{
    __shared__ __align__(1024) uint8_t sA[kSlots1][128 * 64];        // [128 rows][32 K] bf16, 64B-swizzled
    __shared__ __align__(1024) uint8_t sB[NOUT * 64];               // [NOUT rows][32 K] bf16, 64B-swizzled
    __shared__ uint64_t afull[kSlots1], aempty[kSlots1], tfull[kAcc1], tempty[kAcc1];
    __shared__ uint32_t tmem_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_tiles = a.num_tiles;

    // weights -> swizzled B tile (K padded 27 -> 32 with zeros)
    for (int idx = threadIdx.x; idx < NOUT * 32; idx += kThreads) {
        const int co = idx >> 5, k = idx & 31;
        const bf16 v = k < 27 ? a.w[co * 27 + k] : __float2bfloat16(0.f);
        *reinterpret_cast<bf16 *>(sB + co * 64 + ((((k >> 3) ^ ((co >> 1) & 3))) << 4) + (k & 7) * 2) = v;
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < kSlots1; ++i) { mbar_init(&afull[i], 4); mbar_init(&aempty[i], 1); }
        for (int i = 0; i < kAcc1; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // the generic-proxy writes of sB -> visible to the tensor core
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(kAcc1 * 32) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    pdl_launch_dependents();

    if (warp == 0) {
        // ===================================== MMA issuer (whole warp, one elected lane issues) =====================
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NOUT >> 3) << 17) | ((128u >> 4) << 24);
        const uint64_t bdesc = make_desc<32>(smem_u32(sB));
        int i = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++i) {
            const int slot = i % kSlots1, acc = i % kAcc1;
            STEM_WAIT(&tempty[acc], ((i / kAcc1) & 1) ^ 1);
            STEM_WAIT(&afull[slot], (i / kSlots1) & 1);
            tc_fence_after();
            const uint64_t adesc = make_desc<32>(smem_u32(sA[slot]));
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 32);
            tc_mma_bf16_elect(d_tmem, adesc, bdesc, idesc, 0u);
            tc_mma_bf16_elect(d_tmem, adesc + 2, bdesc + 2, idesc, 1u);
            tc_commit_elect(&aempty[slot]);
            tc_commit_elect(&tfull[acc]);
        }
    } else if (warp <= 8) {
        // ===================================== window gather (two groups of 4 warps) ================================
        const int g = (warp - 1) >> 2;
        const int r = ((warp - 1) & 3) * 32 + lane;              // GEMM row = pixel of the tile
        const int px = r % kTW, py = r / kTW;
        const size_t plane = (size_t)a.H * a.W;
        pdl_wait();
        Walk t; t.init(blockIdx.x + g * gridDim.x, 2 * gridDim.x, a.tiles_x, a.tiles_y);
        int i = g;
        for (int tile = blockIdx.x + g * gridDim.x; tile < num_tiles; tile += 2 * gridDim.x, i += 2, t.next(a.tiles_x, a.tiles_y)) {
            const int slot = i % kSlots1;
            const int x = t.tx * kTW + px, y = t.ty * kTH + py;
            const bool valid = x < a.W && y < a.H;
            const float *img = a.in + (size_t)t.tn * 3 * plane;
            float v[28];
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const int yy = y + ky - 1;
                const bool yok = valid && yy >= 0 && yy < a.H;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int xx = x + kx - 1;
                    const bool ok = yok && xx >= 0 && xx < a.W;
                    const float *p = img + (size_t)(ok ? yy : 0) * a.W + (ok ? xx : 0);
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float val = __ldg(p + c * plane);
                        v[(ky * 3 + kx) * 3 + c] = ok ? val : 0.f;
                    }
                }
            }
            v[27] = 0.f;
            STEM_WAIT(&aempty[slot], ((i / kSlots1) & 1) ^ 1);     // loads are already in flight while we wait for the slot
            const uint32_t row_addr = smem_u32(sA[slot]) + (uint32_t)r * 64u;
            const uint32_t sw = (uint32_t)(r >> 1) & 3u;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint4 q;
                if (j < 3) {
                    q.x = pack_bf16(v[8 * j + 0], v[8 * j + 1]); q.y = pack_bf16(v[8 * j + 2], v[8 * j + 3]);
                    q.z = pack_bf16(v[8 * j + 4], v[8 * j + 5]); q.w = pack_bf16(v[8 * j + 6], v[8 * j + 7]);
                } else {
                    q.x = pack_bf16(v[24], v[25]); q.y = pack_bf16(v[26], v[27]); q.z = 0u; q.w = 0u;
                }
                sts128(row_addr + (((uint32_t)j ^ sw) << 4), q);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&afull[slot]);
        }
    } else {
        // ===================================== epilogue (warps 9..12) ===============================================
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;
        const int px = r % kTW, py = r / kTW;
        float sc[NOUT], sh[NOUT];
#pragma unroll
        for (int c = 0; c < NOUT; ++c) { sc[c] = a.scale[c]; sh[c] = a.shift[c]; }
        const bool leaky = a.act == ACT_LEAKY;
        pdl_wait();
        Walk t; t.init(blockIdx.x, gridDim.x, a.tiles_x, a.tiles_y);
        int i = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++i, t.next(a.tiles_x, a.tiles_y)) {
            const int acc = i % kAcc1;
            STEM_WAIT(&tfull[acc], (i / kAcc1) & 1);
            tc_fence_after();
            uint32_t d[NOUT];
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 32);
            if constexpr (NOUT == 32) tmem_ld32(taddr, d); else tmem_ld16(taddr, d);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            const int x = t.tx * kTW + px, y = t.ty * kTH + py;
            if (x < a.W && y < a.H) {
                bf16 *dst = a.out + (((size_t)t.tn * a.H + y) * a.W + x) * a.ldo;
#pragma unroll
                for (int c = 0; c < NOUT; c += 8) {
                    float o[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        o[q] = fmaf(__uint_as_float(d[c + q]), sc[c + q], sh[c + q]);
                        if (leaky) o[q] = o[q] > 0.f ? o[q] : 0.1f * o[q];
                    }
                    uint4 pk;
                    pk.x = pack_bf16(o[0], o[1]); pk.y = pack_bf16(o[2], o[3]); pk.z = pack_bf16(o[4], o[5]); pk.w = pack_bf16(o[6], o[7]);
                    *reinterpret_cast<uint4 *>(dst + c) = pk;
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kAcc1 * 32) : "memory");
    }
}


// ---------------------------------------------------------------------------------------------------
// TMA-fed variant (image width a multiple of 4): the tile's fp32 window source arrives as ONE 4-D TMA box
// [3 channels][6 rows][40 columns] (out-of-bounds = zero = the convolution's padding, so the gather needs no predicates
// and reads shared memory instead of global), and the output tile leaves as ONE TMA store from a swizzled staging tile.
// Warp roles (608 threads): 0 = MMA, 1 = TMA loads, 2 = TMA stores, 3-10 = two gather groups, 11-18 = two epilogue groups
// (gather and epilogue groups take alternate tiles: one warp per scheduler cannot issue a tile's epilogue fast enough).
// ---------------------------------------------------------------------------------------------------
// patch box: rows y0-1 .. y0+4, columns x0-4 .. x0+35.  An un-swizzled TMA box must START on a 16-byte boundary in
// dimension 0 (scripts/tma_f32_probe.cu: x0-1 raises an illegal-instruction error), hence the 4-column left margin.
constexpr int kPW = 40, kPH = 6, kPX = 4;
constexpr int kPatchBytes = 3 * kPH * kPW * 4;    // 2880
constexpr int kPatchPitch = 2944;                 // ring pitch (128-byte aligned TMA destinations)
constexpr int kPatches = 16;
constexpr int kOutSlots = 4;
constexpr int kThreads2 = 19 * 32;

struct alignas(64) StemTcMaps { CUtensorMap in, out; };

template <int NOUT, bool POOL>                 // POOL: fused [maxpool] 2/2 (compile-time: the plain stem pays nothing for it)
__global__ void __launch_bounds__(kThreads2, 1)
conv_stem_tc_tma_kernel(const __grid_constant__ StemTcMaps maps, const StemTcArgs a)
{
    constexpr int ROWB = NOUT * 2;                                   // bytes per pixel row of the output tile (= its swizzle span)
    extern __shared__ uint8_t stem_smem_raw[];
    uint8_t *base = (uint8_t *)(((uintptr_t)stem_smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t (*sA)[128 * 64] = reinterpret_cast<uint8_t (*)[128 * 64]>(base);                               // kSlots x 8 KB
    uint8_t (*sC)[128 * ROWB] = reinterpret_cast<uint8_t (*)[128 * ROWB]>(base + kSlots * 8192);           // kOutSlots tiles
    uint8_t *sB = base + kSlots * 8192 + kOutSlots * 8192;                                                 // 2 KB, 1024-aligned
    uint8_t (*sP)[kPatchPitch] = reinterpret_cast<uint8_t (*)[kPatchPitch]>(base + kSlots * 8192 + kOutSlots * 8192 + 2048);
    __shared__ uint64_t pfull[kPatches], pempty[kPatches], afull[kSlots], aempty[kSlots], tfull[kAcc], tempty[kAcc],
                        cwritten[kOutSlots], cempty[kOutSlots];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) float s_scale[NOUT], s_shift[NOUT];
    if (threadIdx.x < NOUT) { s_scale[threadIdx.x] = a.scale[threadIdx.x]; s_shift[threadIdx.x] = a.shift[threadIdx.x]; }

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_tiles = a.num_tiles;

    for (int idx = threadIdx.x; idx < NOUT * 32; idx += kThreads2) {
        const int co = idx >> 5, k = idx & 31;
        const bf16 v = k < 27 ? a.w[co * 27 + k] : __float2bfloat16(0.f);
        *reinterpret_cast<bf16 *>(sB + co * 64 + ((((k >> 3) ^ ((co >> 1) & 3))) << 4) + (k & 7) * 2) = v;
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < kPatches; ++i) { mbar_init(&pfull[i], 1); mbar_init(&pempty[i], 4); }
        for (int i = 0; i < kSlots; ++i) { mbar_init(&afull[i], 4); mbar_init(&aempty[i], 1); }
        for (int i = 0; i < kAcc; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
        for (int i = 0; i < kOutSlots; ++i) { mbar_init(&cwritten[i], 4); mbar_init(&cempty[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(kAcc * 32) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    pdl_launch_dependents();

    if (warp == 0) {
        // ===================================== MMA issuer ===========================================================
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NOUT >> 3) << 17) | ((128u >> 4) << 24);
        const uint64_t bdesc = make_desc<32>(smem_u32(sB));
        int i = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++i) {
            const int slot = i % kSlots, acc = i % kAcc;
            STEM_WAIT(&tempty[acc], ((i / kAcc) & 1) ^ 1);
            STEM_WAIT(&afull[slot], (i / kSlots) & 1);
            tc_fence_after();
            const uint64_t adesc = make_desc<32>(smem_u32(sA[slot]));
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 32);
            tc_mma_bf16_elect(d_tmem, adesc, bdesc, idesc, 0u);
            tc_mma_bf16_elect(d_tmem, adesc + 2, bdesc + 2, idesc, 1u);
            tc_commit_elect(&aempty[slot]);
            tc_commit_elect(&tfull[acc]);
        }
    } else if (warp == 1) {
        // ===================================== TMA loads: one fp32 window-source box per tile =======================
        if (lane == 0) {
            pdl_wait();
            Walk t; t.init(blockIdx.x, gridDim.x, a.tiles_x, a.tiles_y);
            int i = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++i, t.next(a.tiles_x, a.tiles_y)) {
                const int ps = i % kPatches;
                STEM_WAIT(&pempty[ps], ((i / kPatches) & 1) ^ 1);
                mbar_expect_tx(&pfull[ps], kPatchBytes);
                tma_load_4d(&maps.in, sP[ps], &pfull[ps], t.tx * kTW - kPX, t.ty * kTH - 1, 0, t.tn);
            }
        }
    } else if (warp == 2) {
        // ===================================== TMA stores ============================================================
        if (lane == 0) {
            pdl_wait();
            Walk t; t.init(blockIdx.x, gridDim.x, a.tiles_x, a.tiles_y);
            int i = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++i, t.next(a.tiles_x, a.tiles_y)) {
                const int cs = i % kOutSlots;
                STEM_WAIT(&cwritten[cs], (i / kOutSlots) & 1);
                if constexpr (POOL) tma_store_4d(&maps.out, sC[cs], 0, t.tx * (kTW / 2), t.ty * (kTH / 2), t.tn);      // 16 x 2 pooled pixels
                else tma_store_4d(&maps.out, sC[cs], 0, t.tx * kTW / a.pack, t.ty * kTH, t.tn);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                // one store stays in flight (two measured the same): waiting for THIS tile's read before issuing the next serialised the warp at the
                // store engine's read latency (~0.4 us per tile = the kernel's whole tile period).  The previous tile's slot is
                // free once all groups but the newest have been read.
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                if (i > 0) mbar_arrive(&cempty[(i - 1) % kOutSlots]);
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    } else if (warp <= 10) {
        // ===================================== window gather (two groups of 4 warps) ================================
        const int g = (warp - 3) >> 2;
        const int r = ((warp - 3) & 3) * 32 + lane;
        // GEMM row -> pixel of the 32 x 4 tile.  Plain: row-major.  Fused max-pool: quad-major, so that the four pixels of every
        // 2x2 window sit in four neighbouring TMEM lanes = four neighbouring lanes of ONE epilogue warp (shuffles do the max)
        int px = r % kTW, py = r / kTW;
        if constexpr (POOL) { const int w4 = r >> 5, l = r & 31; px = 2 * ((l >> 2) + 8 * (w4 & 1)) + (l & 1); py = 2 * (w4 >> 1) + ((l >> 1) & 1); }
        const uint32_t sw = (uint32_t)(r >> 1) & 3u;
        int i = g;
        for (int tile = blockIdx.x + g * gridDim.x; tile < num_tiles; tile += 2 * gridDim.x, i += 2) {
            const int slot = i % kSlots, ps = i % kPatches;
            STEM_WAIT(&pfull[ps], (i / kPatches) & 1);
            const float *P = reinterpret_cast<const float *>(sP[ps]) + py * kPW + px + (kPX - 1);
            float v[28];
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                    for (int c = 0; c < 3; ++c) v[(ky * 3 + kx) * 3 + c] = P[c * (kPH * kPW) + ky * kPW + kx];
            v[27] = 0.f;
            __syncwarp();
            if (lane == 0) mbar_arrive(&pempty[ps]);              // values are in registers: the patch slot can be refilled
            STEM_WAIT(&aempty[slot], ((i / kSlots) & 1) ^ 1);
            const uint32_t row_addr = smem_u32(sA[slot]) + (uint32_t)r * 64u;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint4 q;
                if (j < 3) {
                    q.x = pack_bf16(v[8 * j + 0], v[8 * j + 1]); q.y = pack_bf16(v[8 * j + 2], v[8 * j + 3]);
                    q.z = pack_bf16(v[8 * j + 4], v[8 * j + 5]); q.w = pack_bf16(v[8 * j + 6], v[8 * j + 7]);
                } else {
                    q.x = pack_bf16(v[24], v[25]); q.y = pack_bf16(v[26], v[27]); q.z = 0u; q.w = 0u;
                }
                sts128(row_addr + (((uint32_t)j ^ sw) << 4), q);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&afull[slot]);
        }
    } else {
        // ===================================== epilogue (two groups of 4 warps) =====================================
        const int g = (warp - 11) >> 2;
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;
        const bool leaky = a.act == ACT_LEAKY;
        const uint32_t sc_addr = smem_u32(s_scale), sh_addr = smem_u32(s_shift);
        // unpacked: output tile rows are NOUT*2 bytes = the swizzle span: 16-byte chunk j of row r sits at chunk j ^ f(r).
        // packed (dense output, the usual case): `pack` neighbouring pixels share one 128-byte row (TMA moves rows, not
        // bytes: ~4 cycles a row, so 64 rows per tile instead of 128 is what makes the stem's store side cheap)
        const uint32_t swz = NOUT == 32 ? ((uint32_t)(r >> 1) & 3u) : ((uint32_t)(r >> 2) & 1u);
        const int px = r % kTW, py = r / kTW;
        const uint32_t prow = (uint32_t)(py * (kTW / a.pack) + px / a.pack);
        const uint32_t pchunk0 = (uint32_t)(px % a.pack) * (NOUT / 8);
        // fused [maxpool] 2/2 (maxpool_layer.c:79-114): rows are quad-major (see the gather), lanes 4q..4q+3 hold one 2x2 window;
        // the max of the four activations is taken with two shuffles and lane 4q stores pooled pixel (qx, qy) of the 16 x 2 tile.
        // max and bf16 rounding commute (rounding is monotonic): bit-identical to pooling the stored bf16 tensor.
        const uint32_t qrow = (uint32_t)((quarter >> 1) * (kTW / 2) + (lane >> 2) + 8 * (quarter & 1));
        const uint32_t qswz = NOUT == 32 ? ((qrow >> 1) & 3u) : ((qrow >> 2) & 1u);
        int i = g;
        for (int tile = blockIdx.x + g * gridDim.x; tile < num_tiles; tile += 2 * gridDim.x, i += 2) {
            const int acc = i % kAcc, cs = i % kOutSlots;
            STEM_WAIT(&tfull[acc], (i / kAcc) & 1);
            tc_fence_after();
            uint32_t d[NOUT];
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 32);
            if constexpr (NOUT == 32) tmem_ld32(taddr, d); else tmem_ld16(taddr, d);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            STEM_WAIT(&cempty[cs], ((i / kOutSlots) & 1) ^ 1);
            const uint32_t row_addr = smem_u32(sC[cs]) + (POOL ? qrow * ROWB : (a.pack > 1 ? prow * 128u : (uint32_t)r * ROWB));
#pragma unroll
            for (int c = 0; c < NOUT; c += 8) {
                float o[8];
#pragma unroll
                for (int q = 0; q < 8; q += 4) {
                    const uint4 s4 = lds128(sc_addr + (uint32_t)(c + q) * 4u), h4 = lds128(sh_addr + (uint32_t)(c + q) * 4u);
                    o[q + 0] = fmaf(__uint_as_float(d[c + q + 0]), __uint_as_float(s4.x), __uint_as_float(h4.x));
                    o[q + 1] = fmaf(__uint_as_float(d[c + q + 1]), __uint_as_float(s4.y), __uint_as_float(h4.y));
                    o[q + 2] = fmaf(__uint_as_float(d[c + q + 2]), __uint_as_float(s4.z), __uint_as_float(h4.z));
                    o[q + 3] = fmaf(__uint_as_float(d[c + q + 3]), __uint_as_float(s4.w), __uint_as_float(h4.w));
                }
                if (leaky) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) o[q] = fmaxf(o[q], 0.1f * o[q]);
                }
                uint4 pk;
                pk.x = pack_bf16(o[0], o[1]); pk.y = pack_bf16(o[2], o[3]); pk.z = pack_bf16(o[4], o[5]); pk.w = pack_bf16(o[6], o[7]);
                if constexpr (POOL) {                      // max over the 2x2 window on the packed pairs: 2 shuffles per word
                    uint32_t *pw = reinterpret_cast<uint32_t *>(&pk);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint32_t o1 = __shfl_xor_sync(0xffffffffu, pw[q], 1);
                        __nv_bfloat162 m = __hmax2(*reinterpret_cast<__nv_bfloat162 *>(&pw[q]), *reinterpret_cast<__nv_bfloat162 *>(&o1));
                        pw[q] = *reinterpret_cast<uint32_t *>(&m);
                        uint32_t o2 = __shfl_xor_sync(0xffffffffu, pw[q], 2);
                        m = __hmax2(*reinterpret_cast<__nv_bfloat162 *>(&pw[q]), *reinterpret_cast<__nv_bfloat162 *>(&o2));
                        pw[q] = *reinterpret_cast<uint32_t *>(&m);
                    }
                    if ((lane & 3) == 0) sts128(row_addr + ((((uint32_t)(c >> 3)) ^ qswz) << 4), pk);
                } else sts128(row_addr + (a.pack > 1 ? (((pchunk0 + (uint32_t)(c >> 3)) ^ (prow & 7u)) << 4) : (((uint32_t)(c >> 3) ^ swz) << 4)), pk);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&cwritten[cs]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kAcc * 32) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------
// Generalised TMA-fed stem for large windows (YOLOv1: 7x7, stride 2, pad 3, 64 filters).  Same pipeline as above; the im2col
// row of a pixel is KS*KS*3 taps padded to a multiple of 32 and laid out as NKB K-blocks of [128 pixels x 32] (64B-swizzled),
// built block by block so that only 32 taps are live in registers at a time; NKB * 2 MMAs per tile.
// Tile = 32 x 4 output pixels; patch = [3][STRIDE*3 + KS rows][STRIDE*31 + KS + margin columns] of the fp32 input.
// ---------------------------------------------------------------------------------------------------
template <int KS, int STRIDE> struct StemGeom {
    static constexpr int PAD = (KS - 1) / 2;
    static constexpr int TAPS = KS * KS * 3;
    static constexpr int NKB = (TAPS + 31) / 32;                      // K-blocks of 32
    static constexpr int MARGIN = 4;                                  // patch starts 4 columns left of STRIDE*ox0 (16-byte aligned start)
    static constexpr int PW = (STRIDE * (kTW - 1) + KS + (MARGIN - PAD) + 3) / 4 * 4;
    static constexpr int PH = STRIDE * (kTH - 1) + KS;
    static constexpr int PATCH_BYTES = 3 * PH * PW * 4;
    static constexpr int PATCH_PITCH = (PATCH_BYTES + 127) / 128 * 128;
};
constexpr int kGenSlots = 2, kGenPatches = 4, kGenOut = 2, kGenAcc = 4;

template <int NOUT, int KS, int STRIDE>
__global__ void __launch_bounds__(kThreads2, 1)
conv_stem_tc_gen_kernel(const __grid_constant__ StemTcMaps maps, const StemTcArgs a)
{
    using G = StemGeom<KS, STRIDE>;
    static_assert(NOUT == 64, "one 128-byte output row per pixel");
    constexpr int A_SLOT = G::NKB * 8192, B_BYTES = G::NKB * NOUT * 64, C_SLOT = 128 * NOUT * 2;
    extern __shared__ uint8_t stem_smem_raw[];
    uint8_t *base = (uint8_t *)(((uintptr_t)stem_smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = base;                                              // kGenSlots x NKB x 8 KB
    uint8_t *sC = sA + kGenSlots * A_SLOT;                           // kGenOut x 16 KB
    uint8_t *sB = sC + kGenOut * C_SLOT;                             // NKB x [NOUT x 32]
    uint8_t *sP = sB + (B_BYTES + 1023) / 1024 * 1024;               // patch ring
    __shared__ uint64_t pfull[kGenPatches], pempty[kGenPatches], afull[kGenSlots], aempty[kGenSlots], tfull[kGenAcc], tempty[kGenAcc],
                        cwritten[kGenOut], cempty[kGenOut];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) float s_scale[NOUT], s_shift[NOUT];
    if (threadIdx.x < NOUT) { s_scale[threadIdx.x] = a.scale[threadIdx.x]; s_shift[threadIdx.x] = a.shift[threadIdx.x]; }

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_tiles = a.num_tiles;

    for (int idx = threadIdx.x; idx < NOUT * G::NKB * 32; idx += kThreads2) {
        const int co = idx / (G::NKB * 32), k = idx % (G::NKB * 32), kb = k >> 5, kk = k & 31;
        const bf16 v = k < G::TAPS ? a.w[co * G::TAPS + k] : __float2bfloat16(0.f);
        *reinterpret_cast<bf16 *>(sB + kb * (NOUT * 64) + co * 64 + ((((kk >> 3) ^ ((co >> 1) & 3))) << 4) + (kk & 7) * 2) = v;
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < kGenPatches; ++i) { mbar_init(&pfull[i], 1); mbar_init(&pempty[i], 4); }
        for (int i = 0; i < kGenSlots; ++i) { mbar_init(&afull[i], 4); mbar_init(&aempty[i], 1); }
        for (int i = 0; i < kGenAcc; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
        for (int i = 0; i < kGenOut; ++i) { mbar_init(&cwritten[i], 4); mbar_init(&cempty[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(kGenAcc * NOUT) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    pdl_launch_dependents();

    if (warp == 0) {
        // ===================================== MMA issuer ===========================================================
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NOUT >> 3) << 17) | ((128u >> 4) << 24);
        const uint64_t bdesc0 = make_desc<32>(smem_u32(sB)), adesc0 = make_desc<32>(smem_u32(sA));
        int i = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++i) {
            const int slot = i % kGenSlots, acc = i % kGenAcc;
            STEM_WAIT(&tempty[acc], ((i / kGenAcc) & 1) ^ 1);
            STEM_WAIT(&afull[slot], (i / kGenSlots) & 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * NOUT);
#pragma unroll
            for (int kb = 0; kb < G::NKB; ++kb) {
                const uint64_t adesc = adesc0 + (uint64_t)((uint32_t)(slot * A_SLOT + kb * 8192) >> 4);
                const uint64_t bdesc = bdesc0 + (uint64_t)((uint32_t)(kb * NOUT * 64) >> 4);
                tc_mma_bf16_elect(d_tmem, adesc, bdesc, idesc, kb != 0 ? 1u : 0u);
                tc_mma_bf16_elect(d_tmem, adesc + 2, bdesc + 2, idesc, 1u);
            }
            tc_commit_elect(&aempty[slot]);
            tc_commit_elect(&tfull[acc]);
        }
    } else if (warp == 1) {
        // ===================================== TMA loads ============================================================
        if (lane == 0) {
            pdl_wait();
            Walk t; t.init(blockIdx.x, gridDim.x, a.tiles_x, a.tiles_y);
            int i = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++i, t.next(a.tiles_x, a.tiles_y)) {
                const int ps = i % kGenPatches;
                STEM_WAIT(&pempty[ps], ((i / kGenPatches) & 1) ^ 1);
                mbar_expect_tx(&pfull[ps], G::PATCH_BYTES);
                tma_load_4d(&maps.in, sP + ps * G::PATCH_PITCH, &pfull[ps], STRIDE * t.tx * kTW - G::MARGIN, STRIDE * t.ty * kTH - G::PAD, 0, t.tn);
            }
        }
    } else if (warp == 2) {
        // ===================================== TMA stores ===========================================================
        if (lane == 0) {
            pdl_wait();
            Walk t; t.init(blockIdx.x, gridDim.x, a.tiles_x, a.tiles_y);
            int i = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++i, t.next(a.tiles_x, a.tiles_y)) {
                const int cs = i % kGenOut;
                STEM_WAIT(&cwritten[cs], (i / kGenOut) & 1);
                tma_store_4d(&maps.out, sC + cs * C_SLOT, 0, t.tx * kTW, t.ty * kTH, t.tn);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                mbar_arrive(&cempty[cs]);
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    } else if (warp <= 10) {
        // ===================================== window gather (two groups of 4 warps) ================================
        const int g = (warp - 3) >> 2;
        const int r = ((warp - 3) & 3) * 32 + lane;
        const int px = r % kTW, py = r / kTW;
        const uint32_t sw = (uint32_t)(r >> 1) & 3u;
        int i = g;
        for (int tile = blockIdx.x + g * gridDim.x; tile < num_tiles; tile += 2 * gridDim.x, i += 2) {
            const int slot = i % kGenSlots, ps = i % kGenPatches;
            STEM_WAIT(&pfull[ps], (i / kGenPatches) & 1);
            STEM_WAIT(&aempty[slot], ((i / kGenSlots) & 1) ^ 1);
            const float *P = reinterpret_cast<const float *>(sP + ps * G::PATCH_PITCH) + (STRIDE * py) * G::PW + STRIDE * px + (G::MARGIN - G::PAD);
            const uint32_t row_addr = smem_u32(sA) + (uint32_t)(slot * A_SLOT) + (uint32_t)r * 64u;
#pragma unroll
            for (int kb = 0; kb < G::NKB; ++kb) {
                float v[32];
#pragma unroll
                for (int kk = 0; kk < 32; ++kk) {
                    constexpr int dummy = 0; (void)dummy;
                    const int t = kb * 32 + kk;                       // tap index (ky, kx, c), compile-time after unrolling
                    if (t < G::TAPS) {
                        const int c = t % 3, kx = (t / 3) % KS, ky = t / (3 * KS);
                        v[kk] = P[c * (G::PH * G::PW) + ky * G::PW + kx];
                    } else v[kk] = 0.f;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint4 q;
                    q.x = pack_bf16(v[8 * j + 0], v[8 * j + 1]); q.y = pack_bf16(v[8 * j + 2], v[8 * j + 3]);
                    q.z = pack_bf16(v[8 * j + 4], v[8 * j + 5]); q.w = pack_bf16(v[8 * j + 6], v[8 * j + 7]);
                    sts128(row_addr + (uint32_t)(kb * 8192) + (((uint32_t)j ^ sw) << 4), q);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&pempty[ps]);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&afull[slot]);
        }
    } else {
        // ===================================== epilogue (two groups of 4 warps) =====================================
        const int g = (warp - 11) >> 2;
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;
        const bool leaky = a.act == ACT_LEAKY;
        const uint32_t sc_addr = smem_u32(s_scale), sh_addr = smem_u32(s_shift);
        const uint32_t swz = (uint32_t)(r & 7);                    // 128-byte rows
        int i = g;
        for (int tile = blockIdx.x + g * gridDim.x; tile < num_tiles; tile += 2 * gridDim.x, i += 2) {
            const int acc = i % kGenAcc, cs = i % kGenOut;
            STEM_WAIT(&tfull[acc], (i / kGenAcc) & 1);
            tc_fence_after();
            uint32_t d[NOUT];
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * NOUT);
            tmem_ld32(taddr, d);
            tmem_ld32(taddr + 32, d + 32);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            STEM_WAIT(&cempty[cs], ((i / kGenOut) & 1) ^ 1);
            const uint32_t row_addr = smem_u32(sC) + (uint32_t)(cs * C_SLOT) + (uint32_t)r * 128u;
#pragma unroll
            for (int c = 0; c < NOUT; c += 8) {
                float o[8];
#pragma unroll
                for (int q = 0; q < 8; q += 4) {
                    const uint4 s4 = lds128(sc_addr + (uint32_t)(c + q) * 4u), h4 = lds128(sh_addr + (uint32_t)(c + q) * 4u);
                    o[q + 0] = fmaf(__uint_as_float(d[c + q + 0]), __uint_as_float(s4.x), __uint_as_float(h4.x));
                    o[q + 1] = fmaf(__uint_as_float(d[c + q + 1]), __uint_as_float(s4.y), __uint_as_float(h4.y));
                    o[q + 2] = fmaf(__uint_as_float(d[c + q + 2]), __uint_as_float(s4.z), __uint_as_float(h4.z));
                    o[q + 3] = fmaf(__uint_as_float(d[c + q + 3]), __uint_as_float(s4.w), __uint_as_float(h4.w));
                }
                if (leaky) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) o[q] = fmaxf(o[q], 0.1f * o[q]);
                }
                uint4 pk;
                pk.x = pack_bf16(o[0], o[1]); pk.y = pack_bf16(o[2], o[3]); pk.z = pack_bf16(o[4], o[5]); pk.w = pack_bf16(o[6], o[7]);
                sts128(row_addr + (((uint32_t)(c >> 3) ^ swz) << 4), pk);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&cwritten[cs]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kGenAcc * NOUT) : "memory");
    }
}

// tensor maps depend on the buffers only: cache the last few (the chunked H2D path launches the stem per chunk)
struct MapCacheEntry { const void *in; void *out; int n, h, w, c, ld, pack, pw, ph, pooled; StemTcMaps maps; };
static MapCacheEntry g_maps[16];
static int g_maps_used = 0, g_maps_next = 0;

static const StemTcMaps *stem_maps(const float *in, int n, int h, int w, TView out, int pack, int pw = kPW, int ph = kPH, int pooled = 0)
{
    for (int i = 0; i < g_maps_used; ++i) {
        const MapCacheEntry &e = g_maps[i];
        if (e.in == in && e.out == out.p && e.n == n && e.h == h && e.w == w && e.c == out.c && e.ld == out.ld && e.pack == pack && e.pw == pw && e.ph == ph && e.pooled == pooled) return &e.maps;
    }
    MapCacheEntry &e = g_maps[g_maps_next];
    g_maps_next = (g_maps_next + 1) % 16;
    if (g_maps_used < 16) ++g_maps_used;
    e.in = in; e.out = out.p; e.n = n; e.h = h; e.w = w; e.c = out.c; e.ld = out.ld; e.pack = pack; e.pw = pw; e.ph = ph; e.pooled = pooled;
    {
        unsigned long long dims[4] = {(unsigned long long)w, (unsigned long long)h, 3ull, (unsigned long long)n};
        unsigned long long strides[3] = {(unsigned long long)w * 4, (unsigned long long)h * w * 4, (unsigned long long)3 * h * w * 4};
        unsigned box[4] = {(unsigned)pw, (unsigned)ph, 3, 1};
        tc_encode_tiled(&e.maps.in, 1, 4, (void *)in, dims, strides, box, 0);
    }
    {
        // packed: the dense output seen as rows of `pack` pixels (128 bytes)
        const unsigned long long ow = (unsigned long long)out.w, oh = (unsigned long long)out.h;      // = w, h for the stride-1 stems
        unsigned long long dims[4] = {(unsigned long long)out.c * pack, ow / pack, oh, (unsigned long long)n};
        unsigned long long strides[3] = {(unsigned long long)out.ld * 2 * pack, ow * out.ld * 2, oh * ow * out.ld * 2};
        unsigned box[4] = {(unsigned)out.c * pack, (unsigned)(kTW / pack), kTH, 1};
        if (pooled) { box[1] = kTW / 2; box[2] = kTH / 2; }      // `out` is the pooled tensor: a tile is 16 x 2 of its pixels
        tc_encode_tiled(&e.maps.out, 0, 4, out.p, dims, strides, box, out.c * 2 * pack);
    }
    return &e.maps;
}

}  // namespace

// true when the layer was launched here; false -> caller uses the CUDA-core stem
// which stems the tcgen05 kernels take with the [maxpool] 2/2 that follows fused in (checked by the planner before it drops the
// conv's own output buffer from the schedule)
bool conv_stem_tc_pool_supported(int h, int w, int c, TView out, ConvParams p)
{
    return !getenv("B200_STEM_SIMT") && !getenv("B200_STEM_LDG") && !getenv("B200_NO_POOL_FUSION") && out.dtype == DT_BF16 && c == 3 &&
           p.size == 3 && p.stride == 1 && p.pad == 1 && out.c == p.cout_pad && (out.c == 16 || out.c == 32) && out.h == h && out.w == w &&
           w % 4 == 0 && h % 2 == 0 && (p.act == ACT_LEAKY || p.act == ACT_LINEAR);
}

bool launch_conv_stem_tc(const float *in_nchw, int n, int h, int w, int c, TView out, ConvParams p, cudaStream_t s, const TView *pool_out)
{
    if (getenv("B200_STEM_SIMT")) return false;
    if (out.dtype == DT_BF16 && c == 3 && p.size == 7 && p.stride == 2 && p.pad == 3 && out.c == 64 && p.cout_pad == 64 &&
        out.h == (h + 6 - 7) / 2 + 1 && out.w == (w + 6 - 7) / 2 + 1 && w % 4 == 0 && ((uintptr_t)in_nchw & 15) == 0 &&
        out.ld % 8 == 0 && ((uintptr_t)out.p & 15) == 0 && (p.act == ACT_LEAKY || p.act == ACT_LINEAR) && !getenv("B200_STEM7_SIMT")) {
        using G = StemGeom<7, 2>;                                   // YOLOv1's first layer
        StemTcArgs a;
        a.in = in_nchw; a.out = (bf16 *)out.p; a.w = (const bf16 *)p.w; a.scale = p.scale; a.shift = p.shift;
        a.N = n; a.H = h; a.W = w; a.ldo = out.ld; a.act = p.act; a.pack = 1; a.pool = 0;
        a.tiles_x = div_up(out.w, kTW); a.tiles_y = div_up(out.h, kTH);
        a.num_tiles = a.tiles_x * a.tiles_y * n;
        const StemTcMaps *maps = stem_maps(in_nchw, n, h, w, out, 1, G::PW, G::PH);
        const size_t smem = 1024 + (size_t)kGenSlots * G::NKB * 8192 + kGenOut * 16384 + (G::NKB * 64 * 64 + 1023) / 1024 * 1024 + kGenPatches * G::PATCH_PITCH;
        static bool configured7 = false;
        if (!configured7) {
            B200_CHECK(cudaFuncSetAttribute(conv_stem_tc_gen_kernel<64, 7, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            configured7 = true;
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(a.num_tiles < 148 ? a.num_tiles : 148); cfg.blockDim = dim3(kThreads2); cfg.dynamicSmemBytes = smem; cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = getenv("B200_NO_PDL") ? 0 : 1;
        B200_CHECK(cudaLaunchKernelEx(&cfg, conv_stem_tc_gen_kernel<64, 7, 2>, *maps, a));
        return true;
    }
    if (out.dtype != DT_BF16 || c != 3 || p.size != 3 || p.stride != 1 || p.pad != 1) return false;
    if (out.c != p.cout_pad || (out.c != 16 && out.c != 32) || out.h != h || out.w != w) return false;
    if (out.ld % 8 != 0 || ((uintptr_t)out.p & 15)) return false;
    if (p.act != ACT_LEAKY && p.act != ACT_LINEAR) return false;
    StemTcArgs a;
    a.in = in_nchw; a.out = (bf16 *)out.p; a.w = (const bf16 *)p.w; a.scale = p.scale; a.shift = p.shift;
    a.N = n; a.H = h; a.W = w; a.ldo = out.ld; a.act = p.act;
    a.tiles_x = div_up(w, kTW); a.tiles_y = div_up(h, kTH);
    const long long tiles = (long long)a.tiles_x * a.tiles_y * n;
    if (tiles > 0x7fffffff) return false;
    a.num_tiles = (int)tiles;
    a.pack = 1; a.pool = 0;
    const int grid = a.num_tiles < 148 ? a.num_tiles : 148;
    if (w % 4 == 0 && ((uintptr_t)in_nchw & 15) == 0 && !getenv("B200_STEM_LDG")) {
        a.pack = (out.ld == out.c && w % (64 / out.c) == 0 && !getenv("B200_STEM_NOPACK")) ? 64 / out.c : 1;
        const StemTcMaps *maps;
        if (pool_out) {                                       // fused [maxpool]: only the pooled tensor is written
            a.pool = 1; a.pack = 1;
            TView pv = *pool_out;
            maps = stem_maps(in_nchw, n, h, w, pv, 1, kPW, kPH, 1);
        } else maps = stem_maps(in_nchw, n, h, w, out, a.pack);
        const size_t smem = 1024 + kSlots * 8192 + kOutSlots * 8192 + 2048 + (kPatches * kPatchPitch + 1023) / 1024 * 1024;
        static bool configured = false;
        if (!configured) {
            B200_CHECK(cudaFuncSetAttribute(conv_stem_tc_tma_kernel<32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            B200_CHECK(cudaFuncSetAttribute(conv_stem_tc_tma_kernel<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            B200_CHECK(cudaFuncSetAttribute(conv_stem_tc_tma_kernel<32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            B200_CHECK(cudaFuncSetAttribute(conv_stem_tc_tma_kernel<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            configured = true;
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads2); cfg.dynamicSmemBytes = smem; cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = getenv("B200_NO_PDL") ? 0 : 1;
        if (a.pool) {
            if (out.c == 32) B200_CHECK(cudaLaunchKernelEx(&cfg, conv_stem_tc_tma_kernel<32, true>, *maps, a));
            else B200_CHECK(cudaLaunchKernelEx(&cfg, conv_stem_tc_tma_kernel<16, true>, *maps, a));
        } else if (out.c == 32) B200_CHECK(cudaLaunchKernelEx(&cfg, conv_stem_tc_tma_kernel<32, false>, *maps, a));
        else B200_CHECK(cudaLaunchKernelEx(&cfg, conv_stem_tc_tma_kernel<16, false>, *maps, a));
        return true;
    }
    if (pool_out) return false;                               // the LDG fallback has no pooled store: the caller runs the layers separately
    if (out.c == 32) conv_stem_tc_kernel<32><<<grid, kThreads, 0, s>>>(a);
    else conv_stem_tc_kernel<16><<<grid, kThreads, 0, s>>>(a);
    return true;
}
