// common.cuh — shared device-side helpers for the sm_100a kernels of the YOLO inference engine.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#define B200_CHECK(expr)                                                                          \
    do {                                                                                          \
        cudaError_t err__ = (expr);                                                               \
        if (err__ != cudaSuccess) {                                                               \
            fprintf(stderr, "b200-darknet: CUDA error %s at %s:%d: %s\n", #expr, __FILE__,        \
                    __LINE__, cudaGetErrorString(err__));                                         \
            abort();                                                                              \
        }                                                                                         \
    } while (0)

// every kernel launch of this library goes through this counter (bench.py reports it as gpu_launches)
extern unsigned long long g_b200_launches;
#define B200_LAUNCHED()                                                                           \
    do {                                                                                          \
        ++g_b200_launches;                                                                        \
        B200_CHECK(cudaPeekAtLastError());                                                        \
    } while (0)

typedef __nv_bfloat16 bf16;

// activation ids follow darknet.h ACTIVATION (only the ones the YOLO cfgs use are implemented on device)
#define ACT_LOGISTIC 0
#define ACT_RELU 1
#define ACT_LINEAR 3
#define ACT_LEAKY 7

template <typename T> struct Elem;
template <> struct Elem<float> {
    static __device__ __forceinline__ float load(const float *p) { return *p; }
    static __device__ __forceinline__ void store(float *p, float v) { *p = v; }
    static constexpr int VEC = 4;     // elements per 16-byte vector
};
template <> struct Elem<bf16> {
    static __device__ __forceinline__ float load(const bf16 *p) { return __bfloat162float(*p); }
    static __device__ __forceinline__ void store(bf16 *p, float v) { *p = __float2bfloat16_rn(v); }
    static constexpr int VEC = 8;
};

// 16-byte vector of T unpacked to floats and back
template <typename T> __device__ __forceinline__ void load_vec(const T *p, float *v);
template <> __device__ __forceinline__ void load_vec<float>(const float *p, float *v)
{
    float4 q = *reinterpret_cast<const float4 *>(p);
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
}
template <> __device__ __forceinline__ void load_vec<bf16>(const bf16 *p, float *v)
{
    uint4 q = *reinterpret_cast<const uint4 *>(p);
    const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
template <typename T> __device__ __forceinline__ void store_vec(T *p, const float *v);
template <> __device__ __forceinline__ void store_vec<float>(float *p, const float *v)
{
    *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void store_vec<bf16>(bf16 *p, const float *v)
{
    uint4 q;
    __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4 *>(p) = q;
}

// darknet activations (activations.h:32,38).  `exact` reproduces the reference's double-precision constants.
template <bool EXACT> __device__ __forceinline__ float apply_act(float x, int act)
{
    if (act == ACT_LEAKY) {
        if (EXACT) return (x > 0) ? x : (float)(.1 * (double)x);
        return (x > 0) ? x : 0.1f * x;
    }
    if (act == ACT_LOGISTIC) return (float)(1. / (1. + exp(-(double)x)));
    if (act == ACT_RELU) return x > 0 ? x : 0.f;
    return x;
}

static inline int div_up(long long a, long long b) { return (int)((a + b - 1) / b); }
