// decode.cu — get_network_boxes on the device: anchor decode + threshold + ordered stream compaction.
//
// Restates, per head layer, get_yolo_detections/get_yolo_box/correct_yolo_boxes (yolo_layer.c:316-343,83-91,247-273),
// get_region_detections/get_region_box/correct_region_boxes (region_layer.c:364-437,76-84,336-362) and
// get_detection_detections (detection_layer.c:225-252), including the reference's float/double promotion at
// every step, for EVERY image of the batch (the reference only reads batch item 0).
//
// Boxes are numbered per image in the reference's enumeration order (heads in layer order; inside a head the
// order of get_*_detections).  A warp ballot per 32 boxes builds a keep bitmap, a per-image prefix sum of the
// popcounts gives every survivor its slot, so the compacted list has exactly the order fill_network_boxes
// produces (deterministic, no atomics).  Only survivors touch HBM on the write side: (4 + 1 + classes) floats each.
#include "kernels.h"
#include "darknet.h"

#define DEC_THREADS 1024

struct BoxF { float x, y, w, h; };

__device__ __forceinline__ BoxF correct_box(BoxF b, int w, int h, int netw, int neth, int relative)
{
    int new_w, new_h;
    if (((float)netw / w) < ((float)neth / h)) { new_w = netw; new_h = (h * netw) / w; }
    else { new_h = neth; new_w = (w * neth) / h; }
    b.x = (float)(((double)b.x - (netw - new_w) / 2. / netw) / (double)((float)new_w / netw));
    b.y = (float)(((double)b.y - (neth - new_h) / 2. / neth) / (double)((float)new_h / neth));
    b.w = __fmul_rn(b.w, (float)netw / new_w);
    b.h = __fmul_rn(b.h, (float)neth / new_h);
    if (!relative) {
        b.x = __fmul_rn(b.x, (float)w); b.w = __fmul_rn(b.w, (float)w);
        b.y = __fmul_rn(b.y, (float)h); b.h = __fmul_rn(b.h, (float)h);
    }
    return b;
}

// block-wide exclusive prefix of a 0/1 flag; returns this thread's slot and (via total) the block count
__device__ __forceinline__ int block_rank(bool flag, int *warp_totals, int &total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned ballot = __ballot_sync(0xffffffffu, flag);
    int within = __popc(ballot & ((1u << lane) - 1));
    if (lane == 0) warp_totals[warp] = __popc(ballot);
    __syncthreads();
    int before = 0, sum = 0;
    for (int i = 0; i < DEC_THREADS / 32; ++i) {
        int c = warp_totals[i];
        if (i < warp) before += c;
        sum += c;
    }
    __syncthreads();
    total = sum;
    return before + within;
}

// ---------------------------------------------------------------------------------------------------
// three small, fully parallel kernels instead of one serial CTA per image:
//   flags : one thread per anchor box evaluates the threshold test; a warp ballot becomes one bitmap word
//   scan  : one CTA per image turns the word popcounts into exclusive offsets (=> the reference's output order)
//   emit  : one thread per surviving box decodes it and writes it to slot offset[word] + rank-in-word
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ int find_head(const HeadDesc *heads, int nheads, int t, int &local)
{
    int hi = 0;
    for (int i = 1; i < nheads; ++i) if (t >= heads[i].box_base) hi = i;
    local = t - heads[hi].box_base;
    return hi;
}

// entry e of anchor a at `cell` of a YOLO head, either from l.output (activations already applied by forward_yolo_layer)
// or straight from the head convolution's logits with the same double-precision logistic (activations.h:32) on the fly
__device__ __forceinline__ float yolo_entry(const HeadDesc &hd, const float *pred, int img, int a, int e, int cell, int use_raw)
{
    const int wh = hd.w * hd.h;
    if (!use_raw) return pred[(size_t)a * wh * (hd.classes + 5) + (size_t)e * wh + cell];
    float v = hd.raw[((size_t)img * wh + cell) * hd.raw_ld + a * (hd.classes + 5) + e];
    return (e == 2 || e == 3) ? v : (float)(1. / (1. + exp(-(double)v)));
}

__device__ __forceinline__ bool box_keep(const HeadDesc &hd, const float *pred, int t, float thresh, int mode, float &objectness,
                                         float &scale, int &cell, int &a, int img = 0, int use_raw = 0)
{
    const int wh = hd.w * hd.h;
    if (hd.type == YOLO) {
        cell = t / hd.n; a = t % hd.n;
        objectness = yolo_entry(hd, pred, img, a, 4, cell, use_raw);
        scale = objectness;
        return objectness > thresh;
    }
    if (hd.type == REGION) {
        a = t / wh; cell = t % wh;
        scale = pred[(size_t)a * wh * (hd.coords + hd.classes + 1) + hd.coords * wh + cell];
        objectness = scale > thresh ? scale : 0.f;
        return mode == 0 ? true : (objectness != 0.f);
    }
    cell = t / hd.n; a = t % hd.n;                                  // DETECTION
    scale = pred[hd.side * hd.side * hd.classes + cell * hd.n + a];
    objectness = scale;
    return mode == 0 ? true : (objectness != 0.f);
}

__global__ void __launch_bounds__(256)
decode_flags_kernel(const HeadDesc *__restrict__ heads, int nheads, int first_image, int boxes, int words, float thresh, int mode,
                    int use_raw, unsigned *__restrict__ flags)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int img = first_image + blockIdx.y;
    bool keep = false;
    if (t < boxes) {
        int local;
        const HeadDesc &hd = heads[find_head(heads, nheads, t, local)];
        float o, s; int cell, a;
        keep = box_keep(hd, hd.out + (size_t)img * hd.outputs, local, thresh, mode, o, s, cell, a, img, use_raw);
    }
    unsigned b = __ballot_sync(0xffffffffu, keep);
    if ((threadIdx.x & 31) == 0 && (t >> 5) < words) flags[(size_t)blockIdx.y * words + (t >> 5)] = b;
}

__global__ void __launch_bounds__(1024)
decode_scan_kernel(const unsigned *__restrict__ flags, int words, int cap, int *__restrict__ offsets, int *__restrict__ count)
{
    __shared__ int warp_sums[32];
    __shared__ int chunk_total;
    const int slot = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int running = 0;
    for (int w0 = 0; w0 < words; w0 += 1024) {
        const int w = w0 + threadIdx.x;
        const int c = w < words ? __popc(flags[(size_t)slot * words + w]) : 0;
        int inc = c;                                               // inclusive scan inside the warp
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
        if (lane == 31) warp_sums[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            const int v = warp_sums[lane];
            int incw = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, incw, o); if (lane >= o) incw += u; }
            warp_sums[lane] = incw - v;                            // exclusive prefix of the warp totals
            if (lane == 31) chunk_total = incw;
        }
        __syncthreads();
        if (w < words) offsets[(size_t)slot * words + w] = running + warp_sums[warp] + inc - c;
        running += chunk_total;
        __syncthreads();
    }
    if (threadIdx.x == 0) count[slot] = running < cap ? running : cap;
}

__global__ void __launch_bounds__(256)
decode_emit_kernel(const HeadDesc *__restrict__ heads, int nheads, int first_image, int boxes, int words, int netw, int neth,
                   int imw, int imh, const int *__restrict__ im_dims, float thresh, int relative, int mode, int use_raw,
                   const unsigned *__restrict__ flags, const int *__restrict__ offsets, CandBuffers cb)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= boxes) return;
    const int slot = blockIdx.y, img = first_image + blockIdx.y, lane = threadIdx.x & 31;
    if (im_dims) { imw = im_dims[2 * img]; imh = im_dims[2 * img + 1]; }      // per-image original sizes (b200_letterbox_batch)
    const unsigned bits = flags[(size_t)slot * words + (t >> 5)];
    if (!((bits >> lane) & 1u)) return;
    const int dst = offsets[(size_t)slot * words + (t >> 5)] + __popc(bits & ((1u << lane) - 1));
    if (dst >= cb.cap) return;
    int local;
    const HeadDesc &hd = heads[find_head(heads, nheads, t, local)];
    const float *pred = hd.out + (size_t)img * hd.outputs;
    float objectness, scale; int cell, a;
    box_keep(hd, pred, local, thresh, mode, objectness, scale, cell, a, img, use_raw);
    const int wh = hd.w * hd.h;
    const int row = cell / hd.w, col = cell % hd.w;
    BoxF b;
    if (hd.type == YOLO) {
        const float e0 = yolo_entry(hd, pred, img, a, 0, cell, use_raw), e1 = yolo_entry(hd, pred, img, a, 1, cell, use_raw);
        const float e2 = yolo_entry(hd, pred, img, a, 2, cell, use_raw), e3 = yolo_entry(hd, pred, img, a, 3, cell, use_raw);
        b.x = __fdiv_rn(__fadd_rn((float)col, e0), (float)hd.w);
        b.y = __fdiv_rn(__fadd_rn((float)row, e1), (float)hd.h);
        b.w = (float)(exp((double)e2) * (double)hd.anchors[2 * a] / netw);
        b.h = (float)(exp((double)e3) * (double)hd.anchors[2 * a + 1] / neth);
        b = correct_box(b, imw, imh, netw, neth, relative);
    } else if (hd.type == REGION) {
        const float *e = pred + (size_t)a * wh * (hd.coords + hd.classes + 1) + cell;
        b.x = __fdiv_rn(__fadd_rn((float)col, e[0]), (float)hd.w);
        b.y = __fdiv_rn(__fadd_rn((float)row, e[wh]), (float)hd.h);
        b.w = (float)(exp((double)e[2 * wh]) * (double)hd.anchors[2 * a] / hd.w);
        b.h = (float)(exp((double)e[3 * wh]) * (double)hd.anchors[2 * a + 1] / hd.h);
        b = correct_box(b, imw, imh, netw, neth, relative);
    } else {
        const float *bx = pred + hd.side * hd.side * (hd.classes + hd.n) + (cell * hd.n + a) * 4;
        b.x = __fmul_rn(__fdiv_rn(__fadd_rn(bx[0], (float)col), (float)hd.side), (float)imw);
        b.y = __fmul_rn(__fdiv_rn(__fadd_rn(bx[1], (float)row), (float)hd.side), (float)imh);
        b.w = (float)(pow((double)bx[2], (double)(hd.sqrt_ ? 2 : 1)) * imw);
        b.h = (float)(pow((double)bx[3], (double)(hd.sqrt_ ? 2 : 1)) * imh);
    }
    float *cbox = cb.box + ((size_t)slot * cb.cap + dst) * 4;
    cbox[0] = b.x; cbox[1] = b.y; cbox[2] = b.w; cbox[3] = b.h;
    cb.obj[(size_t)slot * cb.cap + dst] = objectness;
    cb.id[(size_t)slot * cb.cap + dst] = t;
}

// class probabilities of the survivors: one warp per survivor, lanes over classes (coalesced writes, 32 loads in flight)
// hierarchy_top_prediction (tree.c:53-81) on absolute probabilities strided by `stride`
__device__ int tree_top_prediction(const float *pred, const HeadDesc &hd, float thresh, int stride)
{
    float p = 1.f;
    int group = 0;
    for (;;) {
        float best = 0.f;
        int best_i = 0;
        const int off = hd.tree_goff[group], sz = hd.tree_gsize[group];
        for (int i = 0; i < sz; ++i) {
            const float val = pred[(size_t)(off + i) * stride];
            if (val > best) { best_i = off + i; best = val; }
        }
        if (__fmul_rn(p, best) > thresh) {
            p = __fmul_rn(p, best);
            group = hd.tree_child[best_i];
            if (group < 0) return best_i;
        } else if (group == 0) {
            return best_i;
        } else {
            return hd.tree_parent[hd.tree_goff[group]];
        }
    }
}

__global__ void __launch_bounds__(256)
decode_probs_kernel(const HeadDesc *__restrict__ heads, int nheads, int first_image, float thresh, int mode, int use_raw, CandBuffers cb,
                    float tree_thresh, const int *__restrict__ map)
{
    const int slot = blockIdx.y, img = first_image + blockIdx.y, lane = threadIdx.x & 31;
    const int n = cb.count[slot];
    for (int d = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); d < n; d += gridDim.x * (blockDim.x >> 5)) {
        int local;
        const HeadDesc &hd = heads[find_head(heads, nheads, cb.id[(size_t)slot * cb.cap + d], local)];
        const float *pred = hd.out + (size_t)img * hd.outputs;
        float objectness, scale; int cell, a;
        box_keep(hd, pred, local, thresh, mode, objectness, scale, cell, a, img, use_raw);
        const int wh = hd.w * hd.h;
        float *pr = cb.prob + ((size_t)slot * cb.cap + d) * cb.classes;
        if (hd.type == REGION && hd.tree_parent) {
            // region_layer.c:412-424: the class entries already hold absolute probabilities (hierarchy_predictions ran); with
            // a map the 200 mapped classes are scored like plain classes, without one the single most specific class whose
            // path stays over tree_thresh gets the box's objectness
            const float *cls = pred + ((size_t)a * (hd.coords + hd.classes + 1) + hd.coords + 1) * wh + cell;
            for (int j = lane; j < hd.classes; j += 32) {
                float p = 0.f;
                if (map && j < 200) { p = __fmul_rn(scale, cls[(size_t)map[j] * wh]); p = p > thresh ? p : 0.f; }
                pr[j] = p;
            }
            __syncwarp();
            if (!map && lane == 0) pr[tree_top_prediction(cls, hd, tree_thresh, wh)] = scale > thresh ? scale : 0.f;
            continue;
        }
        for (int j = lane; j < hd.classes; j += 32) {
            float p;
            if (hd.type == YOLO) p = __fmul_rn(objectness, yolo_entry(hd, pred, img, a, 5 + j, cell, use_raw));
            else if (hd.type == REGION) p = objectness != 0.f ? __fmul_rn(scale, pred[(size_t)a * wh * (hd.coords + hd.classes + 1) + (size_t)(hd.coords + 1 + j) * wh + cell]) : 0.f;
            else p = __fmul_rn(scale, pred[cell * hd.classes + j]);
            pr[j] = p > thresh ? p : 0.f;
        }
    }
}

void launch_decode(const HeadDesc *heads_dev, int nheads, int first_image, int nimages, int netw, int neth,
                   int imw, int imh, float thresh, int relative, int mode, int use_raw, CandBuffers cb, cudaStream_t s, const int *im_dims,
                   float tree_thresh, const int *map_dev)
{
    const int boxes = cb.cap;                                       // cap == anchor boxes per image
    const int words = (boxes + 31) / 32;
    dim3 grid(div_up(boxes, 256), nimages);
    decode_flags_kernel<<<grid, 256, 0, s>>>(heads_dev, nheads, first_image, boxes, words, thresh, mode, use_raw, cb.flags);
    B200_LAUNCHED();
    decode_scan_kernel<<<nimages, 1024, 0, s>>>(cb.flags, words, cb.cap, cb.offsets, cb.count);
    B200_LAUNCHED();
    decode_emit_kernel<<<grid, 256, 0, s>>>(heads_dev, nheads, first_image, boxes, words, netw, neth, imw, imh, im_dims, thresh, relative, mode,
                                            use_raw, cb.flags, cb.offsets, cb);
    B200_LAUNCHED();
    decode_probs_kernel<<<dim3(8, nimages), 256, 0, s>>>(heads_dev, nheads, first_image, thresh, mode, use_raw, cb, tree_thresh, map_dev);
    B200_LAUNCHED();
}

// num_detections() for one image (network.c:510-524): yolo heads count obj > thresh, the others every box
__global__ void __launch_bounds__(DEC_THREADS)
count_kernel(const HeadDesc *__restrict__ heads, int nheads, int image, float thresh, int *count)
{
    __shared__ int warp_totals[DEC_THREADS / 32];
    int sum = 0;
    for (int hi = 0; hi < nheads; ++hi) {
        const HeadDesc hd = heads[hi];
        const int wh = hd.w * hd.h, nboxes = wh * hd.n;
        if (hd.type != YOLO) { sum += nboxes; continue; }
        const float *pred = hd.out + (size_t)image * hd.outputs;
        for (int t0 = 0; t0 < nboxes; t0 += DEC_THREADS) {
            int t = t0 + threadIdx.x;
            bool keep = false;
            if (t < nboxes) {
                int cell = t / hd.n, a = t % hd.n;
                keep = pred[(size_t)a * wh * (hd.classes + 5) + 4 * wh + cell] > thresh;
            }
            int total;
            block_rank(keep, warp_totals, total);
            sum += total;
        }
    }
    if (threadIdx.x == 0) *count = sum;
}

void launch_count_yolo(const HeadDesc *heads_dev, int nheads, int image, float thresh, int *count_dev, cudaStream_t s)
{
    count_kernel<<<1, DEC_THREADS, 0, s>>>(heads_dev, nheads, image, thresh, count_dev);
    B200_LAUNCHED();
}
