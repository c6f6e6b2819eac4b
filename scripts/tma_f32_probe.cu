// probe: 4-D fp32 TMA box loads (the stem's window-source box) — which box shapes / coordinates does the hardware accept?
// FINDING (B200): with SWIZZLE_NONE the box's start in dimension 0 must be 16-byte aligned (x0 = 15 -> illegal instruction,
// x0 = 0 / 12 / -4 fine); negative (out-of-bounds) starts are zero-filled.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o tma_f32_probe tma_f32_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
__global__ void probe(const __grid_constant__ CUtensorMap map, float *out, int x, int y, int c, int n, int bytes, int count)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), d = (uint32_t)__cvta_generic_to_shared(smem);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                     ::"r"(d), "l"(&map), "r"(b), "r"(x), "r"(y), "r"(c), "r"(n) : "memory");
        asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(b) : "memory");
    }
    __syncthreads();
    for (int i = threadIdx.x; i < count; i += blockDim.x) out[i] = ((float *)smem)[i];
}
int main()
{
    const int W = 416, H = 416, C = 3, N = 2;
    std::vector<float> h((size_t)W * H * C * N);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 100003);
    float *d, *o; cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, 65536);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    int boxes[][2] = {{24, 10}, {20, 10}};
    for (auto &bx : boxes) {
        CUtensorMap map;
        cuuint64_t dims[4] = {W, H, C, N}, strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * C * 4};
        cuuint32_t box[4] = {(cuuint32_t)bx[0], (cuuint32_t)bx[1], 3, 1}, ones[4] = {1, 1, 1, 1};
        CUresult r = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        int count = bx[0] * bx[1] * 3;
        for (int x0 : {0, 12, -4, 412}) {
            probe<<<1, 128, 32768>>>(map, o, x0, -1, 0, 1, count * 4, count);
            cudaError_t e = cudaDeviceSynchronize();
            std::vector<float> got(count);
            if (e == cudaSuccess) cudaMemcpy(got.data(), o, count * 4, cudaMemcpyDeviceToHost);
            int bad = 0;
            if (e == cudaSuccess)
                for (int c = 0; c < 3; ++c) for (int yy = 0; yy < bx[1]; ++yy) for (int xx = 0; xx < bx[0]; ++xx) {
                    int gx = x0 + xx, gy = -1 + yy;
                    float want = (gx < 0 || gx >= W || gy < 0 || gy >= H) ? 0.f : h[(((size_t)1 * C + c) * H + gy) * W + gx];
                    if (got[(c * bx[1] + yy) * bx[0] + xx] != want) ++bad;
                }
            printf("box %dx%d x0 %d: encode %d run %s mismatches %d\n", bx[0], bx[1], x0, (int)r, cudaGetErrorString(e), bad);
            if (e != cudaSuccess) return 0;
        }
    }
    return 0;
}
