// conv_tc_plan.h — host-side plan of one tcgen05 convolution launch and the launch helper shared by conv_tc.cu,
// conv_tc_patch.cu (kernels + their launchers) and conv_tc_plan.cu (tensor maps, planner, dispatch).
#pragma once
#include "conv_tc_common.cuh"
#include <vector>

struct ConvTcPlan {
    ConvTcMaps maps;
    ConvTcArgs args;
    int block_k, out_dtype, grid;
    size_t smem_bytes;
    double flops;
    std::string desc;
    // what a flow (conv_tc_flow.cu) needs to re-encode the layer for its own tiling
    TView in, out, res;      // res.p == nullptr: no fused shortcut
    const void *w;           // repacked weights [cout_pad][K]
    int K;
    bool flow_ok;            // the layer can be a member of a flow (see conv_tc_flow_create)
    // split-K plans (args.ksplit > 1): the tile kernel writes raw fp32 partial sums into `ws`, splitk_finalize_kernel produces the layer's output
    struct SplitK {
        float *ws;                       // [ksplit][slab_rows][cout_pad] fp32, owned by the plan
        float *ones;                     // the tile kernel's epilogue constants: scale 1 (cout_pad floats) followed by shift 0
        const float *scale, *shift;      // the layer's folded batch-norm, applied by the finalize kernel
        int act;
        const bf16 *res; int ldr; float res_alpha, res_beta;
        bf16 *out; int ldo;
        long long npix; int cout_pad;
    } sk;
};

// ---------------------------------------------------------------------------------------------------
// flows: a run of layers executed by ONE persistent kernel with tile-level dependencies (conv_tc_flow.cu)
// ---------------------------------------------------------------------------------------------------
struct FlowLayerArgs {
    int item0, items;            // this layer's slice of the flow's global item sequence (item = one 256-pixel x block_n pair tile)
    int m_tiles, n_tiles, block_n, num_kblocks, cin_blocks;
    int im2col, size, stride, pad, OW, OH, npix;
    int in_W, in_H;
    int act, has_res;
    float res_alpha, res_beta;
    int dep, res_dep;            // flow-local index of the layer that produces the A input / the residual; -1: written before the launch
    int dep_off, dep_unit;       // the producer's completion counters (one per 256 of its output pixels) and increments per finished counter
    int res_off, res_unit;
    int done_off;                // this layer's own counters
    int pad_;
    const float *scale, *shift;
};
struct FlowParams {
    const CUtensorMap *maps;     // [layers][4]: A (dense 2-D or im2col), B (block_n / 2 filters per box), C store, R residual load
    const FlowLayerArgs *layers;
    unsigned *done;
    const unsigned *sched;       // the pairs' item lists, back to back: entry = layer << 24 | tile number inside the layer
    const int *sched_off;        // [pairs + 1]
    unsigned long long *trace;   // profiling (B200_FLOW_TRACE=1 at plan time, else null): per item 4 globaltimer stamps of the leader CTA —
                                 // inputs complete, first k-block landed, last MMA issued, tile stored
    unsigned long long *stats;   // [0] ns the TMA producers spent waiting for dependencies, [1] the residual loaders, [2] waits that blocked
    int nl, total_items;
    unsigned epoch;              // launch number: counters only ever grow, the wait target is epoch * unit
};
// counters of the producing layer that cover the input pixels of pair tile `mp` (shared by the kernel and the host scheduler)
__host__ __device__ __forceinline__ void flow_dep_range(const FlowLayerArgs &a, int mp, int &jlo, int &jhi)
{
    if (!a.im2col) { jlo = jhi = mp; return; }                 // 1x1: the same 256 pixels
    const int per = a.OH * a.OW;
    const int p0 = mp * 256;
    int p1 = p0 + 255; if (p1 >= a.npix) p1 = a.npix - 1;
    const int n0 = p0 / per, oy0 = (p0 - n0 * per) / a.OW;
    const int n1 = p1 / per, oy1 = (p1 - n1 * per) / a.OW;
    int rlo = oy0 * a.stride - a.pad; if (rlo < 0) rlo = 0;
    int rhi = oy1 * a.stride - a.pad + a.size - 1; if (rhi > a.in_H - 1) rhi = a.in_H - 1;
    jlo = ((n0 * a.in_H + rlo) * a.in_W) >> 8;                  // whole input rows: a superset of the taps' footprint
    jhi = ((n1 * a.in_H + rhi) * a.in_W + a.in_W - 1) >> 8;
}

struct ConvTcFlow {
    FlowParams fp;
    void *blob;                  // device memory behind fp.maps / fp.layers / fp.done
    size_t smem_bytes;
    double flops;
    std::string desc;
    std::vector<int> item0;      // first item of every member layer (+ the total)
};
static constexpr int kFlowStages = 5, kFlowRing = 3, kFlowMaxLayers = 56;
void conv_tc_flow_launch(ConvTcFlow *f, cudaStream_t s);               // conv_tc_flow.cu

// cudaFuncSetAttribute belongs to the CURRENT device: remember per device, not per process (a process may hold networks on several)
static inline bool first_use_on_this_device(bool (&seen)[64])
{
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || seen[dev]) return false;
    seen[dev] = true;
    return true;
}

// every tcgen05 convolution is launched with programmatic stream serialization (see pdl_wait in tc_ptx.cuh)
template <typename Kernel> static void launch_pdl(Kernel kernel, int grid, int threads, size_t smem, cudaStream_t s, int cluster,
                                                  const ConvTcMaps &maps, const ConvTcArgs &args)
{
    static const bool no_pdl = getenv("B200_NO_PDL") != nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    int n = 0;
    if (cluster > 1) {
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = cluster; attr[n].val.clusterDim.y = 1; attr[n].val.clusterDim.z = 1;
        ++n;
    }
    if (!no_pdl) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    cfg.attrs = attr; cfg.numAttrs = n;
    B200_CHECK(cudaLaunchKernelEx(&cfg, kernel, maps, args));
}

// launchers defined next to their kernels
void conv_tc_launch_tap(ConvTcPlan *p, cudaStream_t s);        // conv_tc.cu: conv_tc_kernel / conv_tc_pair_kernel
void conv_tc_launch_patch(ConvTcPlan *p, cudaStream_t s);      // conv_tc_patch.cu: conv_tc_patch_kernel
void conv_tc_launch_block(ConvTcPlan *p, cudaStream_t s);      // conv_tc_patch.cu: conv_tc_block_kernel
void conv_tc_launch_splitk_finalize(ConvTcPlan *p, cudaStream_t s);   // conv_tc.cu: splitk_finalize_kernel (after the tile kernel of a split-K plan)
