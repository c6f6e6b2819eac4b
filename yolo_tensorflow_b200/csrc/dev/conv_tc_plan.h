// conv_tc_plan.h — host-side plan of one tcgen05 convolution launch and the launch helper shared by conv_tc.cu,
// conv_tc_patch.cu (kernels + their launchers) and conv_tc_plan.cu (tensor maps, planner, dispatch).
#pragma once
#include "conv_tc_common.cuh"

struct ConvTcPlan {
    ConvTcMaps maps;
    ConvTcArgs args;
    int block_k, out_dtype, grid;
    size_t smem_bytes;
    double flops;
    std::string desc;
};

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
