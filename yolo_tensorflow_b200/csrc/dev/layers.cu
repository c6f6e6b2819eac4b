// layers.cu — the HBM-bandwidth-bound layers of the YOLO inference path as coalesced, 128-bit vectorised
// NHWC kernels: maxpool, upsample, shortcut, route (channel-slice copy), reorg, and the NCHW<->NHWC
// boundary transforms.  Reference semantics: maxpool_layer.c:79-114, blas.c:334-349 (upsample_cpu),
// blas.c:68-92 + shortcut_layer.c:62-67, route_layer.c:74-87, blas.c:9-30 + reorg_layer.c:107-109.
//
// Grid sizing: every kernel is a grid-stride loop launched with a multiple of the SM count (148) so a
// launch is an integral number of waves whatever the tensor size.
#include "kernels.h"
#include <cfloat>

unsigned long long g_b200_launches = 0;

static const int kThreads = 256;
static inline int grid_for(long long work_items)
{
    long long blocks = (work_items + kThreads - 1) / kThreads;
    const long long wave = 148 * 8;              // 8 resident CTAs of 256 threads per SM
    if (blocks > wave) blocks = wave * ((blocks + wave - 1) / wave > 4 ? 4 : (blocks + wave - 1) / wave);
    return (int)(blocks < 1 ? 1 : blocks);
}

// ---------------------------------------------------------------------------------------------------
// NCHW fp32 (darknet host layout) <-> NHWC T (device layout).  Tiled through shared memory so that both
// the global read and the global write are coalesced.
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float *__restrict__ src, T *__restrict__ dst, int C, int HW, int ld)
{
    __shared__ float tile[32][33];
    const int n = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const float *s = src + (size_t)n * C * HW;
    T *d = dst + (size_t)n * HW * ld;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int c = c0 + i, p = p0 + threadIdx.x;
        tile[i][threadIdx.x] = (c < C && p < HW) ? s[(size_t)c * HW + p] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int p = p0 + i, c = c0 + threadIdx.x;
        if (p < HW && c < C) Elem<T>::store(d + (size_t)p * ld + c, tile[threadIdx.x][i]);
    }
}

template <typename T>
__global__ void nhwc_to_nchw_kernel(const T *__restrict__ src, float *__restrict__ dst, int C, int HW, int ld)
{
    __shared__ float tile[32][33];
    const int n = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const T *s = src + (size_t)n * HW * ld;
    float *d = dst + (size_t)n * C * HW;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int p = p0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (p < HW && c < C) ? Elem<T>::load(s + (size_t)p * ld + c) : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int c = c0 + i, p = p0 + threadIdx.x;
        if (c < C && p < HW) d[(size_t)c * HW + p] = tile[threadIdx.x][i];
    }
}

void launch_nchw_f32_to_view(const float *src, TView dst, cudaStream_t s)
{
    int HW = dst.h * dst.w;
    dim3 grid(div_up(HW, 32), div_up(dst.c, 32), dst.n), block(32, 8);
    if (dst.dtype == DT_F32) nchw_to_nhwc_kernel<float><<<grid, block, 0, s>>>(src, (float *)dst.p, dst.c, HW, dst.ld);
    else nchw_to_nhwc_kernel<bf16><<<grid, block, 0, s>>>(src, (bf16 *)dst.p, dst.c, HW, dst.ld);
    B200_LAUNCHED();
}

void launch_view_to_nchw_f32(TView src, float *dst, cudaStream_t s)
{
    int HW = src.h * src.w;
    dim3 grid(div_up(HW, 32), div_up(src.c, 32), src.n), block(32, 8);
    if (src.dtype == DT_F32) nhwc_to_nchw_kernel<float><<<grid, block, 0, s>>>((const float *)src.p, dst, src.c, HW, src.ld);
    else nhwc_to_nchw_kernel<bf16><<<grid, block, 0, s>>>((const bf16 *)src.p, dst, src.c, HW, src.ld);
    B200_LAUNCHED();
}

// ---------------------------------------------------------------------------------------------------
// maxpool: one thread per (output pixel, 16-byte channel vector).  Window starts at -pad, out-of-bounds
// taps are skipped (they are -FLT_MAX in the reference and can never win: every window holds >= 1 valid tap).
// ---------------------------------------------------------------------------------------------------
template <typename T, int VEC>
__global__ void maxpool_kernel(const T *__restrict__ in, T *__restrict__ out, int N, int H, int W, int C, int ldi,
                               int OH, int OW, int ldo, int size, int stride, int pad)
{
    const int cv = C / VEC;
    const long long total = (long long)N * OH * OW * cv;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        int v = (int)(t % cv);
        long long pix = t / cv;
        int ox = (int)(pix % OW), oy = (int)((pix / OW) % OH), n = (int)(pix / ((long long)OW * OH));
        float best[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) best[i] = -FLT_MAX;
        for (int ky = 0; ky < size; ++ky) {
            int y = oy * stride + ky - pad;
            if (y < 0 || y >= H) continue;
            for (int kx = 0; kx < size; ++kx) {
                int x = ox * stride + kx - pad;
                if (x < 0 || x >= W) continue;
                float val[VEC];
                const T *src = in + (((size_t)n * H + y) * W + x) * ldi + v * VEC;
                if (VEC == 1) val[0] = Elem<T>::load(src);
                else load_vec<T>(src, val);
#pragma unroll
                for (int i = 0; i < VEC; ++i) best[i] = val[i] > best[i] ? val[i] : best[i];
            }
        }
        T *dst = out + (size_t)pix * ldo + v * VEC;
        if (VEC == 1) Elem<T>::store(dst, best[0]);
        else store_vec<T>(dst, best);
    }
}

// the common case — size 2, stride 2, no padding, every window inside the image (YOLOv2's five pools, YOLOv3-tiny's first
// five): 32-bit index arithmetic, the four taps loaded unconditionally (four 16-byte loads in flight per thread), no branches.
// ncu (round 2, YOLOv2 416 b64): the general kernel above reached 56 % of the HBM copy bandwidth on these layers, bound by
// its 64-bit divisions and per-tap bounds tests rather than by memory.
template <typename T, int VEC>
__global__ void __launch_bounds__(256)
maxpool2x2_kernel(const T *__restrict__ in, T *__restrict__ out, unsigned total, int W, int cv, int ldi, int OH, int OW, int ldo)
{
    for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const unsigned v = t % cv, pix = t / cv;
        const unsigned ox = pix % OW, row = pix / OW;            // row = n * OH + oy; input row = 2 * row (H = 2 * OH)
        const T *src = in + ((size_t)(2 * row) * W + 2 * ox) * ldi + v * VEC;
        float a[VEC], b[VEC], c[VEC], d[VEC];
        load_vec<T>(src, a); load_vec<T>(src + ldi, b);
        load_vec<T>(src + (size_t)W * ldi, c); load_vec<T>(src + (size_t)W * ldi + ldi, d);
        float best[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            // the reference's scan: max starts at -FLT_MAX and a tap replaces it when strictly greater (maxpool_layer.c:96-106),
            // so a NaN tap never wins — same here
            float m = -FLT_MAX;
            m = a[i] > m ? a[i] : m; m = b[i] > m ? b[i] : m; m = c[i] > m ? c[i] : m; m = d[i] > m ? d[i] : m;
            best[i] = m;
        }
        store_vec<T>(out + (size_t)pix * ldo + v * VEC, best);
    }
}

template <typename T>
static void maxpool_dispatch(TView in, TView out, int size, int stride, int pad, cudaStream_t s)
{
    constexpr int V = Elem<T>::VEC;
    bool vec_ok = (in.c % V == 0) && (in.ld % V == 0) && (out.ld % V == 0) &&
                  ((uintptr_t)in.p % 16 == 0) && ((uintptr_t)out.p % 16 == 0);
    long long pixels = (long long)out.n * out.h * out.w;
    if (vec_ok && V > 1 && size == 2 && stride == 2 && pad == 0 && in.h == 2 * out.h && in.w == 2 * out.w && pixels * (in.c / V) < (1ll << 31)) {
        const unsigned total = (unsigned)(pixels * (in.c / V));
        maxpool2x2_kernel<T, V><<<grid_for(total), kThreads, 0, s>>>((const T *)in.p, (T *)out.p, total, in.w, in.c / V, in.ld, out.h, out.w, out.ld);
        B200_LAUNCHED();
        return;
    }
    if (vec_ok) {
        maxpool_kernel<T, V><<<grid_for(pixels * (in.c / V)), kThreads, 0, s>>>(
            (const T *)in.p, (T *)out.p, in.n, in.h, in.w, in.c, in.ld, out.h, out.w, out.ld, size, stride, pad);
    } else {
        maxpool_kernel<T, 1><<<grid_for(pixels * in.c), kThreads, 0, s>>>(
            (const T *)in.p, (T *)out.p, in.n, in.h, in.w, in.c, in.ld, out.h, out.w, out.ld, size, stride, pad);
    }
    B200_LAUNCHED();
}

void launch_maxpool(TView in, TView out, int size, int stride, int pad, cudaStream_t s)
{
    if (in.dtype == DT_F32) maxpool_dispatch<float>(in, out, size, stride, pad, s);
    else maxpool_dispatch<bf16>(in, out, size, stride, pad, s);
}

// ---------------------------------------------------------------------------------------------------
// upsample (nearest, x stride, times scale): one thread per (output pixel, channel vector)
// ---------------------------------------------------------------------------------------------------
template <typename T, int VEC>
__global__ void upsample_kernel(const T *__restrict__ in, T *__restrict__ out, int N, int H, int W, int C, int ldi,
                                int ldo, int stride, float scale)
{
    const int cv = C / VEC, OW = W * stride, OH = H * stride;
    const long long total = (long long)N * OH * OW * cv;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        int v = (int)(t % cv);
        long long pix = t / cv;
        int ox = (int)(pix % OW), oy = (int)((pix / OW) % OH), n = (int)(pix / ((long long)OW * OH));
        const T *src = in + (((size_t)n * H + oy / stride) * W + ox / stride) * ldi + v * VEC;
        T *dst = out + (size_t)pix * ldo + v * VEC;
        float val[VEC];
        if (VEC == 1) val[0] = Elem<T>::load(src); else load_vec<T>(src, val);
#pragma unroll
        for (int i = 0; i < VEC; ++i) val[i] = scale * val[i];
        if (VEC == 1) Elem<T>::store(dst, val[0]); else store_vec<T>(dst, val);
    }
}

template <typename T>
static void upsample_dispatch(TView in, TView out, int stride, float scale, cudaStream_t s)
{
    constexpr int V = Elem<T>::VEC;
    bool vec_ok = (in.c % V == 0) && (in.ld % V == 0) && (out.ld % V == 0) &&
                  ((uintptr_t)in.p % 16 == 0) && ((uintptr_t)out.p % 16 == 0);
    long long pixels = (long long)out.n * out.h * out.w;
    if (vec_ok)
        upsample_kernel<T, V><<<grid_for(pixels * (in.c / V)), kThreads, 0, s>>>((const T *)in.p, (T *)out.p, in.n, in.h, in.w, in.c, in.ld, out.ld, stride, scale);
    else
        upsample_kernel<T, 1><<<grid_for(pixels * in.c), kThreads, 0, s>>>((const T *)in.p, (T *)out.p, in.n, in.h, in.w, in.c, in.ld, out.ld, stride, scale);
    B200_LAUNCHED();
}

void launch_upsample(TView in, TView out, int stride, float scale, cudaStream_t s)
{
    if (in.dtype == DT_F32) upsample_dispatch<float>(in, out, stride, scale, s);
    else upsample_dispatch<bf16>(in, out, stride, scale, s);
}

// ---------------------------------------------------------------------------------------------------
// shortcut: out = act(alpha*in + beta*add) on the overlapping region, act(in) elsewhere.
// Fast path (identical shapes, the only case in YOLOv3) is a pure 128-bit streaming kernel; the general
// path reproduces shortcut_cpu's stride/sample indexing for mismatched shapes.
// ---------------------------------------------------------------------------------------------------
template <typename T, int VEC, bool EXACT>
__global__ void shortcut_same_kernel(const T *__restrict__ in, const T *__restrict__ add, T *__restrict__ out, long long pixels,
                                     int C, int ldi, int lda, int ldo, float alpha, float beta, int act)
{
    const int cv = C / VEC;
    const long long total = pixels * cv;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        int v = (int)(t % cv);
        long long pix = t / cv;
        float a[VEC], b[VEC];
        if (VEC == 1) { a[0] = Elem<T>::load(in + pix * ldi + v); b[0] = Elem<T>::load(add + pix * lda + v); }
        else { load_vec<T>(in + pix * ldi + v * VEC, a); load_vec<T>(add + pix * lda + v * VEC, b); }
#pragma unroll
        for (int i = 0; i < VEC; ++i) a[i] = apply_act<EXACT>(__fadd_rn(__fmul_rn(alpha, a[i]), __fmul_rn(beta, b[i])), act);
        if (VEC == 1) Elem<T>::store(out + pix * ldo + v, a[0]);
        else store_vec<T>(out + pix * ldo + v * VEC, a);
    }
}

template <typename T, bool EXACT>
__global__ void shortcut_general_kernel(const T *__restrict__ in, const T *__restrict__ add, T *__restrict__ out, int N,
                                        int w1, int h1, int c1, int lda, int w2, int h2, int c2, int ldi, int ldo,
                                        float alpha, float beta, int act)
{
    // (w1,h1,c1) = added tensor, (w2,h2,c2) = tensor flowing through; see blas.c:68-92
    int stride = w1 / w2, sample = w2 / w1;
    if (stride < 1) stride = 1;
    if (sample < 1) sample = 1;
    const int minw = w1 < w2 ? w1 : w2, minh = h1 < h2 ? h1 : h2, minc = c1 < c2 ? c1 : c2;
    const long long total = (long long)N * h2 * w2 * c2;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        int k = (int)(t % c2);
        long long pix = t / c2;
        int x = (int)(pix % w2), y = (int)((pix / w2) % h2), n = (int)(pix / ((long long)w2 * h2));
        float v = Elem<T>::load(in + pix * ldi + k);
        if (k < minc && x % sample == 0 && y % sample == 0 && x / sample < minw && y / sample < minh) {
            int i = x / sample, j = y / sample;
            float a = Elem<T>::load(add + (((size_t)n * h1 + j * stride) * w1 + i * stride) * lda + k);
            v = __fadd_rn(__fmul_rn(alpha, v), __fmul_rn(beta, a));
        }
        Elem<T>::store(out + pix * ldo + k, apply_act<EXACT>(v, act));
    }
}

template <typename T, bool EXACT>
static void shortcut_dispatch(TView in, TView add, TView out, float alpha, float beta, int act, cudaStream_t s)
{
    constexpr int V = Elem<T>::VEC;
    long long pixels = (long long)out.n * out.h * out.w;
    bool same = add.w == out.w && add.h == out.h && add.c == out.c;
    if (same) {
        bool vec_ok = (out.c % V == 0) && (in.ld % V == 0) && (add.ld % V == 0) && (out.ld % V == 0) &&
                      ((uintptr_t)in.p % 16 == 0) && ((uintptr_t)add.p % 16 == 0) && ((uintptr_t)out.p % 16 == 0);
        if (vec_ok)
            shortcut_same_kernel<T, V, EXACT><<<grid_for(pixels * (out.c / V)), kThreads, 0, s>>>(
                (const T *)in.p, (const T *)add.p, (T *)out.p, pixels, out.c, in.ld, add.ld, out.ld, alpha, beta, act);
        else
            shortcut_same_kernel<T, 1, EXACT><<<grid_for(pixels * out.c), kThreads, 0, s>>>(
                (const T *)in.p, (const T *)add.p, (T *)out.p, pixels, out.c, in.ld, add.ld, out.ld, alpha, beta, act);
    } else {
        shortcut_general_kernel<T, EXACT><<<grid_for(pixels * out.c), kThreads, 0, s>>>(
            (const T *)in.p, (const T *)add.p, (T *)out.p, out.n, add.w, add.h, add.c, add.ld, out.w, out.h, out.c,
            in.ld, out.ld, alpha, beta, act);
    }
    B200_LAUNCHED();
}

void launch_shortcut(TView in, TView add, TView out, float alpha, float beta, int act, cudaStream_t s)
{
    if (in.dtype == DT_F32) shortcut_dispatch<float, true>(in, add, out, alpha, beta, act, s);
    else shortcut_dispatch<bf16, false>(in, add, out, alpha, beta, act, s);
}

// ---------------------------------------------------------------------------------------------------
// route: copy one input into its channel slice of the concat buffer (16-byte vectors when aligned)
// ---------------------------------------------------------------------------------------------------
template <typename T, int VEC>
__global__ void copy_channels_kernel(const T *__restrict__ in, T *__restrict__ out, long long pixels, int C, int ldi, int ldo)
{
    const int cv = C / VEC;
    const long long total = pixels * cv;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        int v = (int)(t % cv);
        long long pix = t / cv;
        if (VEC == 1) out[pix * ldo + v] = in[pix * ldi + v];
        else *reinterpret_cast<uint4 *>(out + pix * ldo + v * VEC) = *reinterpret_cast<const uint4 *>(in + pix * ldi + v * VEC);
    }
}

template <typename T>
static void copy_dispatch(TView in, TView out, cudaStream_t s)
{
    constexpr int V = Elem<T>::VEC;
    long long pixels = (long long)in.n * in.h * in.w;
    bool vec_ok = (in.c % V == 0) && (in.ld % V == 0) && (out.ld % V == 0) &&
                  ((uintptr_t)in.p % 16 == 0) && ((uintptr_t)out.p % 16 == 0);
    if (vec_ok) copy_channels_kernel<T, V><<<grid_for(pixels * (in.c / V)), kThreads, 0, s>>>((const T *)in.p, (T *)out.p, pixels, in.c, in.ld, out.ld);
    else copy_channels_kernel<T, 1><<<grid_for(pixels * in.c), kThreads, 0, s>>>((const T *)in.p, (T *)out.p, pixels, in.c, in.ld, out.ld);
    B200_LAUNCHED();
}

void launch_copy_channels(TView in, TView out, cudaStream_t s)
{
    if (in.dtype != out.dtype) { fprintf(stderr, "b200-darknet: route dtype mismatch\n"); abort(); }
    if (in.dtype == DT_F32) copy_dispatch<float>(in, out, s);
    else copy_dispatch<bf16>(in, out, s);
}

// ---------------------------------------------------------------------------------------------------
// reorg (YOLOv2 passthrough).  The reference calls reorg_cpu(..., forward=0): with (w,h,c) the INPUT dims,
// flat NCHW position p of the output takes flat NCHW position q(p) of the input, where p is decomposed as
// (k,j,i) over [c][h][w] and q = w2 + w*s*(h2 + h*s*c2) with c2 = k % (c/s^2), off = k / (c/s^2),
// w2 = i*s + off % s, h2 = j*s + off / s.  This is a permutation but NOT space_to_depth; it is reproduced
// index for index.  One thread per output element, output-channel fastest so writes coalesce in NHWC.
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void reorg_kernel(const T *__restrict__ in, T *__restrict__ out, int N, int h, int w, int c, int ldi,
                             int oh, int ow, int oc, int ldo, int s)
{
    const long long total = (long long)N * oh * ow * oc;
    const int small_c = c / (s * s);
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        int co = (int)(t % oc);
        long long pix = t / oc;
        int xo = (int)(pix % ow), yo = (int)((pix / ow) % oh), n = (int)(pix / ((long long)ow * oh));
        int p = xo + ow * (yo + oh * co);                       // flat NCHW offset in the output image
        int i = p % w, j = (p / w) % h, k = p / (w * h);        // same offset seen through the INPUT shape
        int c2 = k % small_c, off = k / small_c;
        int w2 = i * s + off % s, h2 = j * s + off / s;
        int q = w2 + w * s * (h2 + h * s * c2);                 // flat NCHW offset in the input image
        int xi = q % w, yi = (q / w) % h, ci = q / (w * h);
        out[pix * ldo + co] = in[(((size_t)n * h + yi) * w + xi) * ldi + ci];
    }
}

// Table-driven variant: the permutation is the same for every image, so the engine computes it once per [reorg] layer
// (reorg_build_table, same index formula) and the kernel becomes: stage one image in shared memory with coalesced 16-byte
// loads, then emit the output in NHWC order — table read coalesced, image read from shared memory, stores coalesced.  One CTA
// per image.  (ncu, round 2: the formula kernel above is bound by its eight integer divisions per element: 38 us for 11 MB.)
template <typename T>
__global__ void __launch_bounds__(1024)
reorg_table_kernel(const T *__restrict__ in, T *__restrict__ out, int in_pixels, int c, int ldi, int per_image, size_t out_image_elems,
                   const int2 *__restrict__ table)
{
    extern __shared__ __align__(16) unsigned char reorg_smem[];
    T *img = (T *)reorg_smem;
    const int n = blockIdx.x;
    const T *src = in + (size_t)n * in_pixels * ldi;
    constexpr int V = 16 / sizeof(T);
    if (ldi == c && c % V == 0) {
        const uint4 *s4 = (const uint4 *)src;
        uint4 *d4 = (uint4 *)img;
        for (int i = threadIdx.x; i < in_pixels * c / V; i += blockDim.x) d4[i] = s4[i];
    } else {
        for (int i = threadIdx.x; i < in_pixels * c; i += blockDim.x) img[i] = src[(size_t)(i / c) * ldi + i % c];
    }
    __syncthreads();
    T *dst = out + (size_t)n * out_image_elems;
    for (int e = threadIdx.x; e < per_image; e += blockDim.x) {
        const int2 o = table[e];
        dst[o.y] = img[o.x];
    }
}

// host: entry e (output NHWC order: pixel-major, channel fastest) -> x = offset in the dense [pixel][c] input image, y = offset in
// the output image (pixel * ldo + channel).  Index formula of reorg_kernel / blas.c:9-30 with forward = 0.
void reorg_build_table(int h, int w, int c, int oh, int ow, int oc, int ldo, int s, int *table_xy)
{
    const int small_c = c / (s * s);
    for (int pix = 0; pix < oh * ow; ++pix)
        for (int co = 0; co < oc; ++co) {
            const int xo = pix % ow, yo = pix / ow;
            const int p = xo + ow * (yo + oh * co);
            const int i = p % w, j = (p / w) % h, k = p / (w * h);
            const int c2 = k % small_c, off = k / small_c;
            const int w2 = i * s + off % s, h2 = j * s + off / s;
            const int q = w2 + w * s * (h2 + h * s * c2);
            const int xi = q % w, yi = (q / w) % h, ci = q / (w * h);
            const size_t e = (size_t)pix * oc + co;
            table_xy[2 * e] = (yi * w + xi) * c + ci;
            table_xy[2 * e + 1] = pix * ldo + co;
        }
}

void launch_reorg(TView in, TView out, int stride, cudaStream_t s, const int *table_dev)
{
    const size_t image_bytes = (size_t)in.h * in.w * in.c * dt_size(in.dtype);
    if (table_dev && image_bytes <= 200 * 1024 && in.dtype == out.dtype) {
        const int per_image = out.h * out.w * out.c;
        const size_t out_image = (size_t)out.h * out.w * out.ld;
        if (in.dtype == DT_F32) {
            static bool configured[64];
            int dev = 0; cudaGetDevice(&dev);
            if (dev >= 0 && dev < 64 && !configured[dev]) { B200_CHECK(cudaFuncSetAttribute(reorg_table_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); configured[dev] = true; }
            reorg_table_kernel<float><<<in.n, 1024, image_bytes, s>>>((const float *)in.p, (float *)out.p, in.h * in.w, in.c, in.ld, per_image, out_image, (const int2 *)table_dev);
        } else {
            static bool configured[64];
            int dev = 0; cudaGetDevice(&dev);
            if (dev >= 0 && dev < 64 && !configured[dev]) { B200_CHECK(cudaFuncSetAttribute(reorg_table_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); configured[dev] = true; }
            reorg_table_kernel<bf16><<<in.n, 1024, image_bytes, s>>>((const bf16 *)in.p, (bf16 *)out.p, in.h * in.w, in.c, in.ld, per_image, out_image, (const int2 *)table_dev);
        }
        B200_LAUNCHED();
        return;
    }
    long long total = (long long)out.n * out.h * out.w * out.c;
    if (in.dtype == DT_F32)
        reorg_kernel<float><<<grid_for(total), kThreads, 0, s>>>((const float *)in.p, (float *)out.p, in.n, in.h, in.w, in.c, in.ld, out.h, out.w, out.c, out.ld, stride);
    else
        reorg_kernel<bf16><<<grid_for(total), kThreads, 0, s>>>((const bf16 *)in.p, (bf16 *)out.p, in.n, in.h, in.w, in.c, in.ld, out.h, out.w, out.c, out.ld, stride);
    B200_LAUNCHED();
}


// ---------------------------------------------------------------------------------------------------
// Device-side preprocessing: letterbox_image (image.c:960-979) = aspect-preserving resize_image
// (image.c:1347-1390: horizontal pass into `part`, then vertical pass, scale (src-1)/(dst-1), last row/column copied),
// 0.5 fill, centred embed; for uint8 sources also load_image_stb's HWC uint8 -> CHW float / 255 (image.c:1442-1464).
// Every product and sum is rounded separately, in the reference's order, so the result is bit-identical to the C code.
// HBM-bound: one thread per output pixel, 4 source taps per channel.
// ---------------------------------------------------------------------------------------------------
template <bool U8>
__device__ __forceinline__ float lb_src(const unsigned char *base, int sw, int sh, int x, int y, int k)
{
    if (U8) return (float)((double)base[((size_t)y * sw + x) * 3 + k] / 255.);
    return reinterpret_cast<const float *>(base)[((size_t)k * sh + y) * sw + x];
}
template <bool U8>
__device__ __forceinline__ float lb_part(const unsigned char *base, const LetterboxItem &it, int c, int r, int k)   // horizontal pass
{
    if (c == it.nw - 1 || it.sw == 1) return lb_src<U8>(base, it.sw, it.sh, it.sw - 1, r, k);
    const float sx = __fmul_rn((float)c, it.w_scale);
    const int ix = (int)sx;
    const float dx = __fsub_rn(sx, (float)ix);
    return __fadd_rn(__fmul_rn(__fsub_rn(1.f, dx), lb_src<U8>(base, it.sw, it.sh, ix, r, k)),
                     __fmul_rn(dx, lb_src<U8>(base, it.sw, it.sh, ix + 1, r, k)));
}
template <bool U8>
__global__ void __launch_bounds__(256)
letterbox_kernel(const unsigned char *__restrict__ raw, const LetterboxItem *__restrict__ items, float *__restrict__ dst, int w, int h)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, img = blockIdx.z;
    if (x >= w) return;
    const LetterboxItem it = items[img];
    const unsigned char *base = raw + it.src_off;
    float *out = dst + (size_t)img * 3 * h * w + (size_t)y * w + x;
    const int c = x - it.ox, r = y - it.oy;
    if (c < 0 || c >= it.nw || r < 0 || r >= it.nh) {
        out[0] = .5f; out[(size_t)h * w] = .5f; out[(size_t)2 * h * w] = .5f;
        return;
    }
    const float sy = __fmul_rn((float)r, it.h_scale);
    const int iy = (int)sy;
    const float dy = __fsub_rn(sy, (float)iy);
    const bool last = r == it.nh - 1 || it.sh == 1;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float v = __fmul_rn(__fsub_rn(1.f, dy), lb_part<U8>(base, it, c, iy, k));
        if (!last) v = __fadd_rn(v, __fmul_rn(dy, lb_part<U8>(base, it, c, iy + 1, k)));
        out[(size_t)k * h * w] = v;
    }
}

void launch_letterbox(const unsigned char *raw, int src_is_u8_hwc, const LetterboxItem *items_dev, int n, float *dst_nchw, int w, int h, cudaStream_t s)
{
    dim3 grid(div_up(w, 256), h, n);
    if (src_is_u8_hwc) letterbox_kernel<true><<<grid, 256, 0, s>>>(raw, items_dev, dst_nchw, w, h);
    else letterbox_kernel<false><<<grid, 256, 0, s>>>(raw, items_dev, dst_nchw, w, h);
    B200_LAUNCHED();
}
