// probe: how fast can every SM pull its OWN activation boxes out of L2 with TMA, as a function of the tensor layout?
// The CTA-pair convolution kernel loses ~22 % to its A-operand loads while its B-operand loads (weights: the same lines for
// every SM) are free (scripts/operand_traffic_probe.py).  This kernel does only the loads: 148 CTAs, a 5-slot ring of
// [128 pixel rows x 64 channels] bf16 boxes (16 KB), the access pattern of a 3x3 layer at 26x26x256, batch 64 (22 MB tensor,
// L2-resident): per 128-pixel tile 9 taps x 4 channel blocks, each tap the tile shifted by up to a row.
//   layout 0  NHWC:            element (p, ch) at p*C + ch                     -> box rows 512 bytes apart
//   layout 1  channel-blocked: element (p, ch) at ((ch/64)*npix + p)*64 + ch%64 -> box = 16 contiguous KB
//   layout 2  NHWC, every SM reads the SAME tiles (the weights' situation)
//   layout 3  NHWC through a 4-D (C, W, H, N) map, box 64 x 2 x 2 x 32 pixels (the tiled 3x3 A load of round 1)
//   layout 4  NHWC through an im2col-mode map, 128 consecutive output pixels per load (the 3x3 A load of round 2)
// and, in rounds of 12 boxes per CTA, clusters of 2 / 4 CTAs that want the same boxes: each loading them itself, or each loading a
// part and multicasting it to the cluster (does sharing a box cut the time, i.e. is the L2 -> SM fabric what limits the loads?)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_layout_probe tma_layout_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1)
loads_only(const __grid_constant__ CUtensorMap map, int layout, int npix, int tiles_per_cta, int same, int cblocks, unsigned *sink)
{
    extern __shared__ uint8_t raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t full[5];
    constexpr int ST = 5;
    if (threadIdx.x == 0) {
        for (int i = 0; i < ST; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int total = tiles_per_cta * 9 * cblocks;
        const int tile0 = same ? 0 : blockIdx.x * tiles_per_cta;
        auto issue = [&](int i) {
            const int tile = tile0 + i / (9 * cblocks), r = i % (9 * cblocks), tap = r / cblocks, cb = r % cblocks;
            int p = tile * 128 + (tap / 3 - 1) * 26 + (tap % 3 - 1);
            if (p < 0) p = 0;
            if (p > npix - 128) p = npix - 128;
            const int st = i % ST;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[st])), "r"(16384) : "memory");
            const int c0 = layout == 1 ? 0 : cb * 64, c1 = layout == 1 ? cb * npix + p : p;
            if (layout == 3) {          // tile = 2 x 2 pixels of 32 consecutive images, shifted by the tap
                const int t13 = tile % 169, n0 = (tile / 169) * 32 % 64;
                asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                             ::"r"(s32(smem + st * 16384)), "l"(&map), "r"(s32(&full[st])), "r"(cb * 64), "r"((t13 % 13) * 2 + tap % 3 - 1), "r"((t13 / 13) * 2 + tap / 3 - 1), "r"(n0) : "memory");
            } else if (layout == 4) {
                const int p0 = tile * 128, n0 = p0 / 676, r = p0 % 676;
                asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
                             ::"r"(s32(smem + st * 16384)), "l"(&map), "r"(s32(&full[st])), "r"(cb * 64), "r"(r % 26 - 1), "r"(r / 26 - 1), "r"(n0),
                               "h"((uint16_t)(tap % 3)), "h"((uint16_t)(tap / 3)) : "memory");
            } else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         ::"r"(s32(smem + st * 16384)), "l"(&map), "r"(s32(&full[st])), "r"(c0), "r"(c1) : "memory");
        };
        for (int i = 0; i < ST && i < total; ++i) issue(i);
        for (int i = 0; i < total; ++i) {
            const int st = i % ST;
            const uint32_t parity = (uint32_t)(i / ST) & 1u;
            asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(s32(&full[st])), "r"(parity) : "memory");
            if (i + ST < total) issue(i + ST);        // the slot is free as soon as its box has landed: nobody consumes it
        }
        sink[blockIdx.x] = ((volatile unsigned *)smem)[0];
    }
}


// rounds of RS loads (all in flight, then a cluster barrier): cluster of CL CTAs that all want the SAME boxes (two CTA pairs computing
// different filter tiles of the same pixels).  mc = 0: every CTA loads the whole box itself; mc = 1: every CTA loads 128 / CL
// rows of it and multicasts them to all CTAs of the cluster, so each box crosses the L2 -> SM fabric once per cluster.
template <int CL>
__global__ void __launch_bounds__(128, 1)
loads_rounds(const __grid_constant__ CUtensorMap map_full, const __grid_constant__ CUtensorMap map_part, int mc, int npix, int tiles_per_cluster, int cblocks, unsigned *sink)
{
    extern __shared__ uint8_t raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    constexpr int RS = 12;
    __shared__ uint64_t full[RS];
    uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    if (threadIdx.x == 0) {
        for (int i = 0; i < RS; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    const int total = tiles_per_cluster * 9 * cblocks;
    const int tile0 = (blockIdx.x / CL) * tiles_per_cluster;
    const int part_rows = 128 / CL;
    for (int base = 0, round = 0; base < total; base += RS, ++round) {
        if (threadIdx.x == 0) {
            const int n = total - base < RS ? total - base : RS;
            for (int j = 0; j < n; ++j) {
                const int i = base + j;
                const int tile = tile0 + i / (9 * cblocks), r = i % (9 * cblocks), tap = r / cblocks, cb = r % cblocks;
                int p = tile * 128 + (tap / 3 - 1) * 26 + (tap % 3 - 1);
                if (p < 0) p = 0;
                if (p > npix - 128) p = npix - 128;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[j])), "r"(16384) : "memory");
                if (mc) {
                    const uint16_t mask = (uint16_t)((1u << CL) - 1u);
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
                                 ::"r"(s32(smem + j * 16384 + rank * part_rows * 128)), "l"(&map_part), "r"(s32(&full[j])), "r"(cb * 64), "r"(p + (int)rank * part_rows), "h"(mask) : "memory");
                } else {
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                                 ::"r"(s32(smem + j * 16384)), "l"(&map_full), "r"(s32(&full[j])), "r"(cb * 64), "r"(p) : "memory");
                }
            }
            for (int j = 0; j < n; ++j)
                asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(s32(&full[j])), "r"((uint32_t)round & 1u) : "memory");
        }
        // nobody may send into a slot of a CTA that has not finished the round
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) sink[blockIdx.x] = ((volatile unsigned *)smem)[0];
}

template <int CL> static void run_rounds(void *d, int C, int npix, int cblocks, unsigned *sink)
{
    CUtensorMap full, part;
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)npix}, strides[1] = {(cuuint64_t)C * 2};
    cuuint32_t box[2] = {64, 128}, pbox[2] = {64, (cuuint32_t)(128 / CL)}, ones[2] = {1, 1};
    cuTensorMapEncodeTiled(&full, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cuTensorMapEncodeTiled(&part, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, pbox, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const int smem = 12 * 16384 + 2048;
    cudaFuncSetAttribute(loads_rounds<CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int clusters = 148 / CL, grid = clusters * CL, tiles = npix / 128, per = tiles / clusters;
    for (int mc = 0; mc < 2; ++mc) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int it = 0; it < 3; ++it) cudaLaunchKernelEx(&cfg, loads_rounds<CL>, full, part, mc, npix, per, cblocks, sink);
        cudaEventRecord(e0);
        const int reps = 20;
        for (int it = 0; it < reps; ++it) cudaLaunchKernelEx(&cfg, loads_rounds<CL>, full, part, mc, npix, per, cblocks, sink);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        const double bytes = (double)grid * per * 9 * cblocks * 16384.0 * reps;     // bytes DELIVERED into shared memory
        printf("C %d cluster of %d, %s: %s  %.1f us per launch, %.2f TB/s delivered in aggregate, %.1f GB/s per SM\n", C, CL,
               mc ? "each CTA loads 1/CL of the box and multicasts it" : "every CTA loads the whole box", cudaGetErrorString(e), ms / reps * 1e3,
               bytes / (ms * 1e-3) / 1e12, bytes / (ms * 1e-3) / grid / 1e9);
    }
}

int main(int argc, char **argv)
{
    const int C = argc > 1 ? atoi(argv[1]) : 256, npix = 64 * 26 * 26, cblocks = C / 64;
    const int tiles = npix / 128, per = tiles / 148;
    void *d; unsigned *sink;
    cudaMalloc(&d, (size_t)npix * C * 2); cudaMemset(d, 1, (size_t)npix * C * 2);
    cudaMalloc(&sink, 4096);
    cudaFuncSetAttribute(loads_only, cudaFuncAttributeMaxDynamicSharedMemorySize, 5 * 16384 + 2048);
    for (int layout = 0; layout < 5; ++layout) {
        CUtensorMap map;
        cuuint32_t box[2] = {64, 128}, ones[2] = {1, 1};
        CUresult r;
        if (layout == 3) {
            cuuint64_t dims[4] = {(cuuint64_t)C, 26, 26, 64}, strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * 26, (cuuint64_t)C * 2 * 676};
            cuuint32_t box4[4] = {64, 2, 2, 32}, ones4[4] = {1, 1, 1, 1};
            r = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d, dims, strides, box4, ones4, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        } else if (layout == 4) {
            cuuint64_t dims[4] = {(cuuint64_t)C, 26, 26, 64}, strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * 26, (cuuint64_t)C * 2 * 676};
            int lower[2] = {-1, -1}, upper[2] = {-1, -1};
            cuuint32_t ones4[4] = {1, 1, 1, 1};
            r = cuTensorMapEncodeIm2col(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d, dims, strides, lower, upper, 64, 128, ones4, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        } else if (layout == 1) {
            cuuint64_t dims[2] = {64, (cuuint64_t)cblocks * npix}, strides[1] = {128};
            r = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        } else {
            cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)npix}, strides[1] = {(cuuint64_t)C * 2};
            r = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        }
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int it = 0; it < 3; ++it) loads_only<<<148, 128, 5 * 16384 + 2048>>>(map, layout, npix, per, layout == 2, cblocks, sink);
        cudaEventRecord(e0);
        const int reps = 20;
        for (int it = 0; it < reps; ++it) loads_only<<<148, 128, 5 * 16384 + 2048>>>(map, layout, npix, per, layout == 2, cblocks, sink);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        const double bytes = 148.0 * per * 9 * cblocks * 16384.0 * reps;
        printf("C %d layout %d (%s): %s  %.1f us per launch, %.2f TB/s aggregate, %.1f GB/s per SM (the conv kernel needs 43.7 per SM at the tensor rate)\n", C, layout,
               layout == 0 ? "NHWC, own tiles" : layout == 1 ? "channel-blocked, own tiles" : layout == 2 ? "NHWC, same tiles for every SM" : layout == 3 ? "4-D tiled box 2x2x32" : "im2col 128 pixels", cudaGetErrorString(e), ms / reps * 1e3,
               bytes / (ms * 1e-3) / 1e12, bytes / (ms * 1e-3) / 148 / 1e9);
    }
    run_rounds<2>(d, C, npix, cblocks, sink);
    run_rounds<4>(d, C, npix, cblocks, sink);
    return 0;
}
