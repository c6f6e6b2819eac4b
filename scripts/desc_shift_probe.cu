// Probe: does a tcgen05 K-major SWIZZLE_128B shared-memory descriptor whose start address is shifted by r rows (r*128 B,
// not 1024-aligned) read rows r..r+127 of a TMA-written tile, with or without the base_offset field set to (addr>>7)&7 ?
// A[r][c] = r*8 + c/8 (fp16, exact), B[n][c] = 1 iff c == 8n  =>  D[m][n] should equal (m+shift)*8 + n.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o desc_shift_probe desc_shift_probe.cu -lcuda
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t su32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mwait(uint64_t *bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"(su32(bar)), "r"(parity) : "memory");
}
constexpr int ROWS = 160, NSHIFT = 12;
__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, float *out)
{
    extern __shared__ uint8_t raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem, *sB = smem + 32768;
    uint64_t *bar = (uint64_t *)(smem + 40960); uint64_t *mbar = bar + 1; uint32_t *slot = (uint32_t *)(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(su32(slot)), "r"(32) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(bar)), "r"(ROWS * 128 + 16 * 128) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(su32(sA)), "l"(&mapA), "r"(su32(bar)), "r"(0), "r"(0) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(su32(sB)), "l"(&mapB), "r"(su32(bar)), "r"(0), "r"(0) : "memory");
    }
    mwait(bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t mphase = 0;
    for (int variant = 0; variant < 2; ++variant)
        for (int shift = 0; shift < NSHIFT; ++shift) {
            if (threadIdx.x == 0) {
                const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);   // f16 x f16 -> f32, N=16, M=128
                uint32_t a_addr = su32(sA) + shift * 128, b_addr = su32(sB);
                uint64_t base = (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
                uint64_t adesc = base | ((a_addr & 0x3FFFF) >> 4);
                if (variant == 1) adesc |= (uint64_t)((a_addr >> 7) & 7) << 49;
                uint64_t bdesc = base | ((b_addr & 0x3FFFF) >> 4);
                for (int k = 0; k < 4; ++k) {
                    uint32_t accum = k ? 1u : 0u;
                    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem), "l"(adesc + k * 2), "l"(bdesc + k * 2), "r"(idesc), "r"(accum) : "memory");
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(su32(mbar)) : "memory");
            }
            mwait(mbar, mphase); mphase ^= 1;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t r[16];
            uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            float *o = out + ((size_t)(variant * NSHIFT + shift) * 128 + warp * 32 + lane) * 16;
            for (int j = 0; j < 16; ++j) o[j] = __uint_as_float(r[j]);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
        }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32) : "memory");
}
typedef CUresult (*EncFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main()
{
    std::vector<__half> hA(ROWS * 64), hB(16 * 64);
    for (int r = 0; r < ROWS; ++r) for (int c = 0; c < 64; ++c) hA[r * 64 + c] = __float2half((float)(r * 8 + c / 8));
    for (int n = 0; n < 16; ++n) for (int c = 0; c < 64; ++c) hB[n * 64 + c] = __float2half((n < 8 && c == 8 * n) ? 1.f : 0.f);
    __half *dA, *dB; float *dO;
    CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dO, 2 * NSHIFT * 128 * 16 * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    void *fp = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    EncFn enc = (EncFn)fp;
    CUtensorMap mA, mB; cuuint32_t ones[2] = {1, 1};
    { cuuint64_t d[2] = {64, ROWS}, st[1] = {128}; cuuint32_t box[2] = {64, ROWS};
      if (enc(&mA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dA, d, st, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("encode A failed\n"); return 1; } }
    { cuuint64_t d[2] = {64, 16}, st[1] = {128}; cuuint32_t box[2] = {64, 16};
      if (enc(&mB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dB, d, st, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("encode B failed\n"); return 1; } }
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024));
    probe<<<1, 128, 44 * 1024>>>(mA, mB, dO);
    CK(cudaDeviceSynchronize());
    std::vector<float> o(2 * NSHIFT * 128 * 16);
    CK(cudaMemcpy(o.data(), dO, o.size() * 4, cudaMemcpyDeviceToHost));
    for (int v = 0; v < 2; ++v)
        for (int s = 0; s < NSHIFT; ++s) {
            int bad = 0, first = -1;
            for (int m = 0; m < 128; ++m) for (int n = 0; n < 8; ++n) {
                float want = (float)((m + s) * 8 + n), got = o[((size_t)(v * NSHIFT + s) * 128 + m) * 16 + n];
                if (want != got) { if (first < 0) first = m * 8 + n; ++bad; }
            }
            printf("base_offset=%s shift=%2d rows: mismatches %4d / 1024", v ? "set" : "0  ", s, bad);
            if (bad) { int m = first / 8, n = first % 8; printf("  first at m=%d n=%d want %.0f got %.0f ; row0: ", m, n, (float)((m + s) * 8 + n), o[((size_t)(v * NSHIFT + s) * 128 + m) * 16 + n]);
                       for (int n2 = 0; n2 < 8; ++n2) printf("%.0f ", o[((size_t)(v * NSHIFT + s) * 128 + 0) * 16 + n2]); }
            printf("\n");
        }
    return 0;
}
