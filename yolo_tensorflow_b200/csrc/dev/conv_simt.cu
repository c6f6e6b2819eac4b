// conv_simt.cu — CUDA-core convolutions.
//
//  * conv_stem_kernel: the network's first layer (C_in <= 4).  Reads the caller's fp32 NCHW image directly,
//    so the NCHW->NHWC / fp32->bf16 boundary conversion costs no extra pass, and writes NHWC.
//  * conv_simt_kernel: generic implicit GEMM (64 pixels x 64 filters x 16 K per step, fp32 accumulate).
//    It is the arithmetic of the fp32-exact mode (B200_PREC_FP32, rtol 1e-4 against the reference) and the
//    fallback of the bf16 mode for shapes the tcgen05 kernel does not take.
//
// Both restate forward_convolutional_layer (convolutional_layer.c:445-485): out[m][n] = sum_k W[m][k]*col[k][n]
// with zero padding (im2col.c:3-12), then the inference batch-norm as a per-filter scale/shift
// ((x-mean)/(sqrt(var)+1e-6)*gamma+beta, blas.c:147-158, batchnorm_layer.c:150-154) or the plain bias, then the
// activation.  K is walked in (ky,kx,c) order here instead of the reference's (c,ky,kx): the weights are
// repacked to match at load time, only the fp32 summation order differs.
#include "kernels.h"

// ---------------------------------------------------------------------------------------------------
template <typename T, bool EXACT>
__global__ void __launch_bounds__(256)
conv_stem_kernel(const float *__restrict__ in, int N, int H, int W, int C, T *__restrict__ out, int OH, int OW, int ldo,
                 int Cout, int size, int stride, int pad, const T *__restrict__ wt, const float *__restrict__ scale,
                 const float *__restrict__ shift, int act)
{
    extern __shared__ float ws[];                    // [K][Cout] fp32
    const int K = size * size * C;
    for (int i = threadIdx.x; i < K * Cout; i += blockDim.x) {
        int k = i / Cout, co = i % Cout;
        ws[i] = Elem<T>::load(wt + (size_t)co * K + k);
    }
    __syncthreads();
    const long long total = (long long)N * OH * OW;
    for (long long pix = blockIdx.x * (long long)blockDim.x + threadIdx.x; pix < total; pix += (long long)gridDim.x * blockDim.x) {
        int ox = (int)(pix % OW), oy = (int)((pix / OW) % OH), n = (int)(pix / ((long long)OW * OH));
        const float *img = in + (size_t)n * C * H * W;
        for (int c0 = 0; c0 < Cout; c0 += 32) {
            float acc[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] = 0.f;
            for (int ky = 0; ky < size; ++ky) {
                int y = oy * stride + ky - pad;
                if (y < 0 || y >= H) continue;
                for (int kx = 0; kx < size; ++kx) {
                    int x = ox * stride + kx - pad;
                    if (x < 0 || x >= W) continue;
                    for (int c = 0; c < C; ++c) {
                        float v = __ldg(img + ((size_t)c * H + y) * W + x);
                        const float *wrow = ws + ((ky * size + kx) * C + c) * Cout + c0;
#pragma unroll
                        for (int j = 0; j < 32; ++j) acc[j] = fmaf(v, wrow[j], acc[j]);   // reads past Cout stay inside ws? guarded below
                    }
                }
            }
            T *dst = out + (size_t)pix * ldo + c0;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                int co = c0 + j;
                if (co < Cout) Elem<T>::store(dst + j, apply_act<EXACT>(fmaf(acc[j], scale[co], shift[co]), act));
            }
        }
    }
}

// Specialised first-layer kernel (SIZE x SIZE taps, CIN input channels, 32 filters per thread):
//  * one thread = one output pixel x 32 filters; consecutive threads = consecutive pixels, so the fp32 NCHW
//    image reads are coalesced per (channel, tap) and served from L1 (each input value is reused SIZE^2 times),
//  * the [K][32] weight slab lives in shared memory and is read as broadcast float4 (one wavefront per 4 FMAs),
//  * all loops are compile-time unrolled, borders are handled once while loading the patch row into registers,
//  * the 32 results leave as 16-byte vector stores (64 contiguous bytes per pixel in bf16 NHWC).
template <typename T, bool EXACT, int SIZE, int CIN>
__global__ void __launch_bounds__(256)
conv_stem_fast_kernel(const float *__restrict__ in, int N, int H, int W, T *__restrict__ out, int OH, int OW, int ldo,
                      int Cout, int stride, int pad, const T *__restrict__ wt, const float *__restrict__ scale,
                      const float *__restrict__ shift, int act)
{
    constexpr int K = SIZE * SIZE * CIN;
    __shared__ __align__(16) float ws[K * 32];
    __shared__ float s_scale[32], s_shift[32];
    const int c0 = blockIdx.y * 32;
    for (int i = threadIdx.x; i < K * 32; i += blockDim.x) {
        int k = i / 32, j = i % 32;
        ws[i] = (c0 + j < Cout) ? Elem<T>::load(wt + (size_t)(c0 + j) * K + k) : 0.f;
    }
    if (threadIdx.x < 32) {
        s_scale[threadIdx.x] = c0 + threadIdx.x < Cout ? scale[c0 + threadIdx.x] : 0.f;
        s_shift[threadIdx.x] = c0 + threadIdx.x < Cout ? shift[c0 + threadIdx.x] : 0.f;
    }
    __syncthreads();
    const long long total = (long long)N * OH * OW;
    const size_t plane = (size_t)H * W;
    for (long long pix = blockIdx.x * (long long)blockDim.x + threadIdx.x; pix < total; pix += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(pix % OW), oy = (int)((pix / OW) % OH), n = (int)(pix / ((long long)OW * OH));
        const float *img = in + (size_t)n * CIN * plane;
        float acc[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = 0.f;
#pragma unroll
        for (int ky = 0; ky < SIZE; ++ky) {
            const int y = oy * stride + ky - pad;
            const bool yok = y >= 0 && y < H;
            float v[SIZE * CIN];
#pragma unroll
            for (int kx = 0; kx < SIZE; ++kx) {
                const int x = ox * stride + kx - pad;
                const bool ok = yok && x >= 0 && x < W;
#pragma unroll
                for (int c = 0; c < CIN; ++c) v[kx * CIN + c] = ok ? __ldg(img + c * plane + (size_t)y * W + x) : 0.f;
            }
#pragma unroll
            for (int t = 0; t < SIZE * CIN; ++t) {
                const float4 *wrow = reinterpret_cast<const float4 *>(ws + (ky * SIZE * CIN + t) * 32);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    float4 w4 = wrow[q];
                    acc[4 * q + 0] = fmaf(v[t], w4.x, acc[4 * q + 0]);
                    acc[4 * q + 1] = fmaf(v[t], w4.y, acc[4 * q + 1]);
                    acc[4 * q + 2] = fmaf(v[t], w4.z, acc[4 * q + 2]);
                    acc[4 * q + 3] = fmaf(v[t], w4.w, acc[4 * q + 3]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = apply_act<EXACT>(fmaf(acc[j], s_scale[j], s_shift[j]), act);
        T *dst = out + (size_t)pix * ldo + c0;
        constexpr int V = Elem<T>::VEC;
#pragma unroll
        for (int j = 0; j < 32; j += V)
            if (c0 + j < Cout) store_vec<T>(dst + j, acc + j);
    }
}

// ---------------------------------------------------------------------------------------------------
// Constant-bank variant of the first-layer kernel.  The shared-memory version above is bound by the LSU: one broadcast
// LDS.128 returns 512 bytes to a warp (4 cycles of the 128 B/clk shared-memory pipe) per 4 FMAs, and the four SM
// sub-partitions share that pipe.  Here the [K][COUT] fp32 weights sit in __constant__ memory, so every FMA takes its
// weight as a constant-bank operand (FFMA R, R, c[bank][imm], R): no load instructions at all in the inner loop.
// The bank holds ONE stem at a time; it is refilled (3.4 KB for YOLOv3) whenever a different weight pointer is launched.
// ---------------------------------------------------------------------------------------------------
#define STEM_CONST_MAX (147 * 64)
__constant__ float c_stem_w[STEM_CONST_MAX];
__constant__ float c_stem_scale[64];
__constant__ float c_stem_shift[64];

template <typename T, bool EXACT, int SIZE, int CIN, int COUT, int C0, int PIX>
__device__ __forceinline__ void conv_stem_const_body(const float *__restrict__ in, int N, int H, int W, T *__restrict__ out, int OH,
                                                     int OW, int ldo, int stride, int pad, int act)
{
    // one thread = PIX horizontally adjacent output pixels x CH filters: each constant-bank weight feeds PIX FMAs and the
    // input columns the pixels share are loaded once
    constexpr int CH = COUT < 32 ? COUT : 32;               // filters per thread
    constexpr int c0 = C0;                                  // compile-time, so every weight is an immediate constant-bank operand
    constexpr int MAXS = 2;                                 // strides 1 and 2 are supported
    constexpr int COLS = SIZE + (PIX - 1) * MAXS;
    const int groups_x = (OW + PIX - 1) / PIX;
    const long long total = (long long)N * OH * groups_x;
    const size_t plane = (size_t)H * W;
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
        const int gx = (int)(g % groups_x), oy = (int)((g / groups_x) % OH), n = (int)(g / ((long long)groups_x * OH));
        const int ox0 = gx * PIX;
        const float *img = in + (size_t)n * CIN * plane;
        float acc[PIX][CH];
#pragma unroll
        for (int p = 0; p < PIX; ++p)
#pragma unroll
            for (int j = 0; j < CH; ++j) acc[p][j] = 0.f;
#pragma unroll
        for (int ky = 0; ky < SIZE; ++ky) {
            const int y = oy * stride + ky - pad;
            const bool yok = y >= 0 && y < H;
            float v[COLS * CIN];
#pragma unroll
            for (int cx = 0; cx < COLS; ++cx) {
                const int x = ox0 * stride + cx - pad;
                const bool ok = yok && x >= 0 && x < W && cx < SIZE + (PIX - 1) * stride;
#pragma unroll
                for (int c = 0; c < CIN; ++c) v[cx * CIN + c] = ok ? __ldg(img + c * plane + (size_t)y * W + x) : 0.f;
            }
#pragma unroll
            for (int kx = 0; kx < SIZE; ++kx)
#pragma unroll
                for (int c = 0; c < CIN; ++c) {
                    const float *wrow = c_stem_w + ((ky * SIZE + kx) * CIN + c) * COUT + c0;
                    float xin[PIX];
#pragma unroll
                    for (int p = 0; p < PIX; ++p) xin[p] = stride == 1 ? v[(kx + p) * CIN + c] : v[(kx + 2 * p) * CIN + c];
#pragma unroll
                    for (int j = 0; j < CH; ++j)
#pragma unroll
                        for (int p = 0; p < PIX; ++p) acc[p][j] = fmaf(xin[p], wrow[j], acc[p][j]);
                }
        }
        constexpr int V = Elem<T>::VEC;
#pragma unroll
        for (int p = 0; p < PIX; ++p) {
            if (ox0 + p >= OW) break;
#pragma unroll
            for (int j = 0; j < CH; ++j) acc[p][j] = apply_act<EXACT>(fmaf(acc[p][j], c_stem_scale[c0 + j], c_stem_shift[c0 + j]), act);
            T *dst = out + ((size_t)(n * OH + oy) * OW + ox0 + p) * ldo + c0;
#pragma unroll
            for (int j = 0; j < CH; j += V) store_vec<T>(dst + j, acc[p] + j);
        }
    }
}

template <typename T, bool EXACT, int SIZE, int CIN, int COUT>
__global__ void __launch_bounds__(256)
conv_stem_const_kernel(const float *__restrict__ in, int N, int H, int W, T *__restrict__ out, int OH, int OW, int ldo,
                       int stride, int pad, int act)
{
    constexpr int PIX = 1;      // 2 pixels per thread measured slower (0.76 vs 0.66 ms on YOLOv3-416 b64): register pressure
    if (COUT <= 32 || blockIdx.y == 0) conv_stem_const_body<T, EXACT, SIZE, CIN, COUT, 0, PIX>(in, N, H, W, out, OH, OW, ldo, stride, pad, act);
    else conv_stem_const_body<T, EXACT, SIZE, CIN, COUT, (COUT > 32 ? 32 : 0), PIX>(in, N, H, W, out, OH, OW, ldo, stride, pad, act);
}

template <typename T>
__global__ void stem_weights_to_f32_kernel(const T *__restrict__ wt, int K, int Cout, float *__restrict__ dst)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < K * Cout; i += gridDim.x * blockDim.x) {
        int k = i / Cout, co = i % Cout;
        dst[i] = Elem<T>::load(wt + (size_t)co * K + k);       // [Cout][K] -> [K][Cout]
    }
}

static const void *g_stem_bank_owner = nullptr;
static float *g_stem_scratch = nullptr;
void conv_stem_invalidate_bank() { g_stem_bank_owner = nullptr; }

template <typename T, bool EXACT>
static bool try_stem_const(const float *in_nchw, int n, int h, int w, int c, TView out, ConvParams p, cudaStream_t s)
{
    if (getenv("B200_STEM_SMEM")) return false;
    const int K = p.size * p.size * c;
    // 3x3 stems only: the 7x7 YOLOv1 stem measured slower from the constant bank (3.6 vs 2.9 ms) than from shared memory
    if (c != 3 || p.size != 3 || (p.stride != 1 && p.stride != 2) || (out.c != 16 && out.c != 32 && out.c != 64)) return false;
    if (out.ld % Elem<T>::VEC != 0 || ((uintptr_t)out.p & 15)) return false;
    if (g_stem_bank_owner != p.w) {                          // (re)fill the constant bank for this network's stem
        if (!g_stem_scratch) B200_CHECK(cudaMalloc((void **)&g_stem_scratch, STEM_CONST_MAX * sizeof(float)));
        stem_weights_to_f32_kernel<T><<<8, 256, 0, s>>>((const T *)p.w, K, out.c, g_stem_scratch);
        B200_LAUNCHED();
        B200_CHECK(cudaMemcpyToSymbolAsync(c_stem_w, g_stem_scratch, (size_t)K * out.c * sizeof(float), 0, cudaMemcpyDeviceToDevice, s));
        B200_CHECK(cudaMemcpyToSymbolAsync(c_stem_scale, p.scale, out.c * sizeof(float), 0, cudaMemcpyDeviceToDevice, s));
        B200_CHECK(cudaMemcpyToSymbolAsync(c_stem_shift, p.shift, out.c * sizeof(float), 0, cudaMemcpyDeviceToDevice, s));
        g_stem_bank_owner = p.w;
    }
    long long groups = (long long)n * out.h * out.w;
    int gx = (int)((groups + 255) / 256);
    if (gx > 148 * 8) gx = 148 * 8;
    dim3 grid(gx, out.c > 32 ? out.c / 32 : 1);
#define STEM_CONST_LAUNCH(SZ, CO) conv_stem_const_kernel<T, EXACT, SZ, 3, CO><<<grid, 256, 0, s>>>(in_nchw, n, h, w, (T *)out.p, out.h, out.w, out.ld, p.stride, p.pad, p.act)
    if (out.c == 16) STEM_CONST_LAUNCH(3, 16); else if (out.c == 32) STEM_CONST_LAUNCH(3, 32); else STEM_CONST_LAUNCH(3, 64);
#undef STEM_CONST_LAUNCH
    return true;
}

template <typename T, bool EXACT>
static bool try_stem_fast(const float *in_nchw, int n, int h, int w, int c, TView out, ConvParams p, cudaStream_t s)
{
    if (c != 3 || (p.size != 3 && p.size != 7)) return false;
    if (out.c % Elem<T>::VEC != 0 || out.ld % Elem<T>::VEC != 0 || ((uintptr_t)out.p & 15)) return false;
    long long pixels = (long long)n * out.h * out.w;
    int gx = (int)((pixels + 255) / 256);
    if (gx > 148 * 8) gx = 148 * 8;
    dim3 grid(gx, div_up(out.c, 32));
    if (p.size == 3)
        conv_stem_fast_kernel<T, EXACT, 3, 3><<<grid, 256, 0, s>>>(in_nchw, n, h, w, (T *)out.p, out.h, out.w, out.ld, out.c, p.stride, p.pad,
                                                                     (const T *)p.w, p.scale, p.shift, p.act);
    else
        conv_stem_fast_kernel<T, EXACT, 7, 3><<<grid, 256, 0, s>>>(in_nchw, n, h, w, (T *)out.p, out.h, out.w, out.ld, out.c, p.stride, p.pad,
                                                                     (const T *)p.w, p.scale, p.shift, p.act);
    return true;
}

void launch_conv_stem(const float *in_nchw, int n, int h, int w, int c, TView out, ConvParams p, cudaStream_t s, const TView *pool_out)
{
    if (launch_conv_stem_tc(in_nchw, n, h, w, c, out, p, s, pool_out)) {     // bf16 3x3 stride-1 stems: tcgen05 (conv_stem_tc.cu)
        B200_LAUNCHED();
        return;
    }
    if (pool_out) { fprintf(stderr, "b200-darknet: internal error: fused stem + maxpool plan without a tcgen05 stem\n"); abort(); }
    if (out.dtype == DT_F32 ? try_stem_const<float, true>(in_nchw, n, h, w, c, out, p, s)
                            : try_stem_const<bf16, false>(in_nchw, n, h, w, c, out, p, s)) {
        B200_LAUNCHED();
        return;
    }
    if (out.dtype == DT_F32 ? try_stem_fast<float, true>(in_nchw, n, h, w, c, out, p, s)
                            : try_stem_fast<bf16, false>(in_nchw, n, h, w, c, out, p, s)) {
        B200_LAUNCHED();
        return;
    }
    int K = p.size * p.size * c;
    // ws is read 32 filters at a time: pad the allocation so the tail chunk stays in bounds
    size_t smem = ((size_t)K * out.c + 32) * sizeof(float);
    long long pixels = (long long)n * out.h * out.w;
    int grid = (int)((pixels + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    if (out.dtype == DT_F32) {
        B200_CHECK(cudaFuncSetAttribute(conv_stem_kernel<float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        conv_stem_kernel<float, true><<<grid, 256, smem, s>>>(in_nchw, n, h, w, c, (float *)out.p, out.h, out.w, out.ld, out.c,
                                                                  p.size, p.stride, p.pad, (const float *)p.w, p.scale, p.shift, p.act);
    } else {
        B200_CHECK(cudaFuncSetAttribute(conv_stem_kernel<bf16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        conv_stem_kernel<bf16, false><<<grid, 256, smem, s>>>(in_nchw, n, h, w, c, (bf16 *)out.p, out.h, out.w, out.ld, out.c,
                                                                  p.size, p.stride, p.pad, (const bf16 *)p.w, p.scale, p.shift, p.act);
    }
    B200_LAUNCHED();
}

// ---------------------------------------------------------------------------------------------------
// generic implicit GEMM: block tile 64 pixels x 64 filters, K step 16, 256 threads, 4x4 outputs per thread
// ---------------------------------------------------------------------------------------------------
#define SB_M 64
#define SB_N 64
#define SB_K 16

template <typename T, typename TO, bool EXACT>
__global__ void __launch_bounds__(256)
conv_simt_kernel(const T *__restrict__ in, int N, int H, int W, int C, int ldi, TO *__restrict__ out, int OH, int OW, int ldo,
                 int Cout, int size, int stride, int pad, const T *__restrict__ wt, const float *__restrict__ scale,
                 const float *__restrict__ shift, int act)
{
    __shared__ float As[SB_K][SB_M + 4];
    __shared__ float Bs[SB_K][SB_N + 4];
    const int K = size * size * C;
    const long long M = (long long)N * OH * OW;
    const long long m0 = (long long)blockIdx.x * SB_M;
    const int n0 = blockIdx.y * SB_N;
    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;          // tx -> filters, ty -> pixels

    // loader roles: thread loads A[pixel = tid/4][k = (tid%4)*4 .. +4) and B[filter = tid/4][same k range]
    const int lrow = tid / 4, lk = (tid % 4) * 4;
    const long long apix = m0 + lrow;
    int aox = 0, aoy = 0, an = 0;
    const bool apix_ok = apix < M;
    if (apix_ok) { aox = (int)(apix % OW); aoy = (int)((apix / OW) % OH); an = (int)(apix / ((long long)OW * OH)); }
    const int bco = n0 + lrow;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < K; k0 += SB_K) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int k = k0 + lk + j;
            float a = 0.f, b = 0.f;
            if (k < K) {
                int tap = k / C, c = k - tap * C;
                int ky = tap / size, kx = tap - ky * size;
                int y = aoy * stride + ky - pad, x = aox * stride + kx - pad;
                if (apix_ok && y >= 0 && y < H && x >= 0 && x < W)
                    a = Elem<T>::load(in + (((size_t)an * H + y) * W + x) * ldi + c);
                if (bco < Cout) b = Elem<T>::load(wt + (size_t)bco * K + k);
            }
            As[lk + j][lrow] = a;
            Bs[lk + j][lrow] = b;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < SB_K; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        long long pix = m0 + ty * 4 + i;
        if (pix >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int co = n0 + tx * 4 + j;
            if (co < Cout) Elem<TO>::store(out + pix * ldo + co, apply_act<EXACT>(fmaf(acc[i][j], scale[co], shift[co]), act));
        }
    }
}

void launch_conv_simt(TView in, TView out, ConvParams p, cudaStream_t s)
{
    long long M = (long long)out.n * out.h * out.w;
    dim3 grid(div_up(M, SB_M), div_up(out.c, SB_N));
#define SIMT_ARGS(TI, TO) (const TI *)in.p, in.n, in.h, in.w, in.c, in.ld, (TO *)out.p, out.h, out.w, out.ld, out.c, p.size, p.stride, \
                          p.pad, (const TI *)p.w, p.scale, p.shift, p.act
    if (in.dtype == DT_F32 && out.dtype == DT_F32) conv_simt_kernel<float, float, true><<<grid, 256, 0, s>>>(SIMT_ARGS(float, float));
    else if (in.dtype == DT_BF16 && out.dtype == DT_BF16) conv_simt_kernel<bf16, bf16, false><<<grid, 256, 0, s>>>(SIMT_ARGS(bf16, bf16));
    else if (in.dtype == DT_BF16 && out.dtype == DT_F32) conv_simt_kernel<bf16, float, false><<<grid, 256, 0, s>>>(SIMT_ARGS(bf16, float));
    else { fprintf(stderr, "b200-darknet: conv_simt: unsupported dtype pair\n"); abort(); }
#undef SIMT_ARGS
    B200_LAUNCHED();
}
