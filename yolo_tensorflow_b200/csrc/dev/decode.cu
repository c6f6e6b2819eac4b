// decode.cu — get_network_boxes on the device: anchor decode + threshold + ordered stream compaction.
//
// Restates, per head layer, get_yolo_detections/get_yolo_box/correct_yolo_boxes (yolo_layer.c:316-343,83-91,247-273),
// get_region_detections/get_region_box/correct_region_boxes (region_layer.c:364-437,76-84,336-362) and
// get_detection_detections (detection_layer.c:225-252), including the reference's float/double promotion at
// every step, for EVERY image of the batch (the reference only reads batch item 0).
//
// One 1024-thread CTA per image walks the heads in layer order and each head's boxes in the reference's
// enumeration order; survivors are appended with a warp-ballot + block prefix sum, so the compacted list has
// exactly the order fill_network_boxes produces (deterministic, no atomics).  Only survivors touch HBM on the
// write side: (4 + 1 + classes) floats each.
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

__global__ void __launch_bounds__(DEC_THREADS)
decode_kernel(const HeadDesc *__restrict__ heads, int nheads, int first_image, int netw, int neth, int imw, int imh,
              float thresh, int relative, int mode, CandBuffers cb)
{
    __shared__ int warp_totals[DEC_THREADS / 32];
    const int img = first_image + blockIdx.x;      // image inside the batch
    const int slot = blockIdx.x;                   // slot inside the candidate buffers
    float *cbox = cb.box + (size_t)slot * cb.cap * 4;
    float *cobj = cb.obj + (size_t)slot * cb.cap;
    float *cprob = cb.prob + (size_t)slot * cb.cap * cb.classes;
    int *cid = cb.id + (size_t)slot * cb.cap;
    int written = 0;

    for (int hi = 0; hi < nheads; ++hi) {
        const HeadDesc hd = heads[hi];
        const float *pred = hd.out + (size_t)img * hd.outputs;
        const int wh = hd.w * hd.h;
        const int nboxes = wh * hd.n;
        for (int t0 = 0; t0 < nboxes; t0 += DEC_THREADS) {
            const int t = t0 + threadIdx.x;
            bool keep = false;
            float objectness = 0.f, scale = 0.f;
            int cell = 0, a = 0;
            if (t < nboxes) {
                if (hd.type == YOLO) {
                    cell = t / hd.n; a = t % hd.n;
                    objectness = pred[(size_t)a * wh * (hd.classes + 5) + 4 * wh + cell];
                    scale = objectness;
                    keep = objectness > thresh;
                } else if (hd.type == REGION) {
                    a = t / wh; cell = t % wh;
                    scale = pred[(size_t)a * wh * (hd.coords + hd.classes + 1) + hd.coords * wh + cell];
                    objectness = scale > thresh ? scale : 0.f;
                    keep = mode == 0 ? true : (objectness != 0.f);
                } else {                                          // DETECTION
                    cell = t / hd.n; a = t % hd.n;
                    scale = pred[hd.side * hd.side * hd.classes + cell * hd.n + a];
                    objectness = scale;
                    keep = mode == 0 ? true : (objectness != 0.f);
                }
            }
            int total;
            int rank = block_rank(keep, warp_totals, total);
            if (keep && written + rank < cb.cap) {
                const int dst = written + rank;
                const int row = cell / hd.w, col = cell % hd.w;
                BoxF b;
                float *pr = cprob + (size_t)dst * cb.classes;
                if (hd.type == YOLO) {
                    const float *e = pred + (size_t)a * wh * (hd.classes + 5) + cell;
                    b.x = __fdiv_rn(__fadd_rn((float)col, e[0]), (float)hd.w);
                    b.y = __fdiv_rn(__fadd_rn((float)row, e[wh]), (float)hd.h);
                    b.w = (float)(exp((double)e[2 * wh]) * (double)hd.anchors[2 * a] / netw);
                    b.h = (float)(exp((double)e[3 * wh]) * (double)hd.anchors[2 * a + 1] / neth);
                    for (int j = 0; j < hd.classes; ++j) {
                        float p = __fmul_rn(objectness, e[(size_t)(5 + j) * wh]);
                        pr[j] = p > thresh ? p : 0.f;
                    }
                    b = correct_box(b, imw, imh, netw, neth, relative);
                } else if (hd.type == REGION) {
                    const float *e = pred + (size_t)a * wh * (hd.coords + hd.classes + 1) + cell;
                    b.x = __fdiv_rn(__fadd_rn((float)col, e[0]), (float)hd.w);
                    b.y = __fdiv_rn(__fadd_rn((float)row, e[wh]), (float)hd.h);
                    b.w = (float)(exp((double)e[2 * wh]) * (double)hd.anchors[2 * a] / hd.w);
                    b.h = (float)(exp((double)e[3 * wh]) * (double)hd.anchors[2 * a + 1] / hd.h);
                    for (int j = 0; j < hd.classes; ++j) {
                        float p = 0.f;
                        if (objectness != 0.f) {
                            p = __fmul_rn(scale, e[(size_t)(hd.coords + 1 + j) * wh]);
                            p = p > thresh ? p : 0.f;
                        }
                        pr[j] = p;
                    }
                    b = correct_box(b, imw, imh, netw, neth, relative);
                } else {
                    const float *bx = pred + hd.side * hd.side * (hd.classes + hd.n) + (cell * hd.n + a) * 4;
                    b.x = __fmul_rn(__fdiv_rn(__fadd_rn(bx[0], (float)col), (float)hd.side), (float)imw);
                    b.y = __fmul_rn(__fdiv_rn(__fadd_rn(bx[1], (float)row), (float)hd.side), (float)imh);
                    b.w = (float)(pow((double)bx[2], (double)(hd.sqrt_ ? 2 : 1)) * imw);
                    b.h = (float)(pow((double)bx[3], (double)(hd.sqrt_ ? 2 : 1)) * imh);
                    const float *cls = pred + cell * hd.classes;
                    for (int j = 0; j < hd.classes; ++j) {
                        float p = __fmul_rn(scale, cls[j]);
                        pr[j] = p > thresh ? p : 0.f;
                    }
                }
                cbox[dst * 4 + 0] = b.x; cbox[dst * 4 + 1] = b.y; cbox[dst * 4 + 2] = b.w; cbox[dst * 4 + 3] = b.h;
                cobj[dst] = objectness;
                cid[dst] = hd.box_base + t;
            }
            written += total;
        }
    }
    if (threadIdx.x == 0) cb.count[slot] = written < cb.cap ? written : cb.cap;
}

void launch_decode(const HeadDesc *heads_dev, int nheads, int first_image, int nimages, int netw, int neth,
                   int imw, int imh, float thresh, int relative, int mode, CandBuffers cb, cudaStream_t s)
{
    decode_kernel<<<nimages, DEC_THREADS, 0, s>>>(heads_dev, nheads, first_image, netw, neth, imw, imh, thresh, relative, mode, cb);
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
