// conv_tc_flow.cu — a RUN of convolution layers as ONE persistent cta_group::2 kernel with tile-level dependencies.
//
// Why: at batch 64 every YOLOv3 convolution launch pays a batch-independent 10-25 us (pipeline fill, last-tile drain, wave
// quantisation: 170 pair tiles on 74 CTA pairs = 3 waves for 2.3 waves of work) — a third of the launch, while the kernels'
// marginal rate is at the tensor peak (profiles/r2_launch_anatomy.txt).  A kernel boundary is a barrier over ALL tiles; the
// data dependency between two layers is local: a 3x3 tile needs the rows of its halo, a 1x1 tile its own pixels.
//
// How: the flow's layers are cut into items (one 256-pixel x block_n tile of one layer, computed by a CTA pair exactly like
// conv_tc_pair_kernel does) numbered layer by layer; every pair walks its own item list across layer borders.  The lists
// come from the planner (conv_tc_flow_create): list scheduling of all items on the 74 pairs in simulated time, a free pair
// taking the oldest item whose inputs are complete — so pairs that finish a layer early work ahead in the following layers
// instead of idling through the last wave.
// Every layer tiles its output in the same way — 256 CONSECUTIVE pixels in (n, y, x) order, which is what the TMA im2col
// mode makes possible for 3x3 layers — and owns one completion counter per tile row in HBM.  The store warp bumps the
// counter once the tile's TMA stores have completed (release, gpu scope); before an item's operands are loaded, the counters
// covering the pixel range it reads and its shortcut operand are polled (acquire) — by a scout warp that runs ahead of the
// TMA producers, so the L2 round trips of the polls stay off their critical path.
// The schedule is a valid execution in simulated time and all 74 pairs are co-resident, so the unfinished item with the
// earliest simulated start can always run (its inputs and its pair's earlier items started before it): no deadlock.  The smem ring, the TMEM double buffer and the epilogue ring never drain between layers: the epilogue of
// layer n's last tile overlaps the mainloop of layer n+1's first.
//
// Replaces (per member layer): convolutional_layer.c:445-485 (+ shortcut_layer.c:62-67 where a shortcut is fused).
// Roofline: tensor pipe; FLOPs = the members' sum.
#include "conv_tc_plan.h"

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu(unsigned *p, unsigned v)
{
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// the scout publishes "the inputs of this pair's first n items are complete" in a shared-memory word of both CTAs
__device__ __forceinline__ void st_release_cluster(uint32_t smem_addr, uint32_t cta, unsigned v)
{
    asm volatile(
        "{\n\t"
        ".reg .b32 remote;\n\t"
        "mapa.shared::cluster.u32 remote, %0, %1;\n\t"
        "st.release.cluster.shared::cluster.u32 [remote], %2;\n\t"
        "}" ::"r"(smem_addr), "r"(cta), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_cluster_smem(uint32_t smem_addr)
{
    unsigned v;
    asm volatile("ld.acquire.cluster.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_addr) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// spin until the counter has reached `target` (counters only grow; wrap-safe signed distance).  A dependency that does not
// arrive within 4 s is a scheduling bug: trap (the launch fails loudly) rather than hang the device.  stat: where the
// nanoseconds of a wait that had to spin are added (profiling: how far the schedule is from never stalling).
__device__ __forceinline__ void flow_wait_counter(const unsigned *ctr, unsigned target, int layer, int item, unsigned long long *stat, unsigned long long *nblocked)
{
    if ((int)(ld_acquire_gpu(ctr) - target) >= 0) return;
    const unsigned long long t0 = global_ns();
    unsigned polls = 0;
    while ((int)(ld_acquire_gpu(ctr) - target) < 0) {
        if ((++polls & 1023u) == 0 && global_ns() - t0 > 4000000000ull) {
            printf("b200-darknet: flow kernel: dependency of layer %d item %d did not arrive (counter %u, target %u)\n", layer, item, *ctr, target);
            __trap();
        }
    }
    atomicAdd(stat, global_ns() - t0);
    atomicAdd(nblocked, 1ull);
}

__global__ void __launch_bounds__(kTcRingThreads, 1)
conv_tc_flow_kernel(const __grid_constant__ FlowParams fp)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;                                        // [stages][128 pixels x 64 channels]
    uint8_t *sB = sA + kFlowStages * 16384;                    // [stages][block_n / 2 filters x 64 channels] in 16 KB slots
    uint8_t *sC = sB + kFlowStages * 16384;                    // ring of swizzled 64-filter output sub-tiles
    uint8_t *aux = sC + kFlowRing * 16384;
    uint64_t *full = (uint64_t *)aux;                          // [stages]  (leader's copy collects both CTAs' bytes)
    uint64_t *empty = full + 8;                                // [stages]
    uint64_t *tfull = empty + 8;                               // [2]
    uint64_t *tempty = tfull + 2;                              // [2]  (leader's copy collects the 16 epilogue warps)
    uint64_t *cfull = tempty + 2, *cempty = cfull + 4, *cwritten = cempty + 4;     // output ring [kFlowRing]
    uint32_t *tmem_slot = (uint32_t *)(cwritten + 4);
    uint32_t *ready = tmem_slot + 1;                           // items of this pair's list whose inputs are complete (written by the scout)
    FlowLayerArgs *L = (FlowLayerArgs *)(aux + 512);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair_id = blockIdx.x >> 1;
    const int k_begin = __ldg(fp.sched_off + pair_id), k_end = __ldg(fp.sched_off + pair_id + 1);    // this pair's slice of the schedule

    for (int i = threadIdx.x; i < fp.nl * (int)(sizeof(FlowLayerArgs) / 4); i += blockDim.x)
        ((int *)L)[i] = ((const int *)fp.layers)[i];           // written by the host at plan time: no kernel ever modifies it
    if (threadIdx.x == 0) {
        *ready = 0;
        for (int i = 0; i < kFlowStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 16); }
        for (int i = 0; i < kFlowRing; ++i) { mbar_init(&cfull[i], 1); mbar_init(&cempty[i], 1); mbar_init(&cwritten[i], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();
    long long clk0 = 0; unsigned long long ns0 = 0;
    if (fp.trace && blockIdx.x == 0 && threadIdx.x == 0) { clk0 = clock64(); ns0 = global_ns(); }

    if (warp == 0) {
        // ===================================== TMA producer (both CTAs) =================================================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            const uint32_t ready_addr = smem_u32(ready);
            pdl_wait();
            unsigned ent = k_begin < k_end ? __ldg(fp.sched + k_begin) : 0u;
            for (int k = k_begin; k < k_end; ++k) {
                const int l = (int)(ent >> 24), t = (int)(ent & 0xffffffu);
                if (k + 1 < k_end) ent = __ldg(fp.sched + k + 1);          // next entry: in flight while this item loads
                const FlowLayerArgs &a = L[l];
                const int n_tile = t % a.n_tiles, mp = t / a.n_tiles;
                const CUtensorMap *amap = fp.maps + 4 * l, *bmap = amap + 1;
                const int m_tile = 2 * mp + (int)rank;
                const int half_n = a.block_n >> 1;
                const uint32_t tx_bytes = 2u * (uint32_t)(16384 + half_n * 128);
                int w0 = 0, h0 = 0, n0 = 0;
                if (a.im2col) {
                    int mt = m_tile < a.m_tiles ? m_tile : a.m_tiles - 1;  // phantom second tile of an odd count: re-read the last real one
                    const int per = a.OH * a.OW, p0 = mt * 128;
                    n0 = p0 / per;
                    const int r = p0 - n0 * per, oy = r / a.OW;
                    h0 = oy * a.stride - a.pad;
                    w0 = (r - oy * a.OW) * a.stride - a.pad;
                }
                // the scout (peer CTA's warp 1) has seen the completion counters of everything this item reads
                while (ld_acquire_cluster_smem(ready_addr) < (unsigned)(k - k_begin + 1)) { }
                if (fp.trace && leader) {
                    const size_t it = (size_t)a.item0 + t;
                    fp.trace[5 * it + 0] = global_ns(); fp.trace[5 * it + 4] = (unsigned long long)pair_id | ((unsigned long long)(k - k_begin) << 16);
                }
                // (no proxy fence here: with operand loads in flight it waits for them — measured 1.5 us per item.  The scout
                // fences between its acquire and the publication; the loads below are issued after the acquire above returned.)
                for (int kb = 0; kb < a.num_kblocks; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    if (leader) mbar_expect_tx(&full[stage], tx_bytes);
                    const int tap = kb / a.cin_blocks, cb = kb - tap * a.cin_blocks;
                    void *dstA = sA + (size_t)stage * 16384;
                    if (a.im2col) tma2_load_im2col_4d(amap, dstA, &full[stage], cb * 64, w0, h0, n0, tap % a.size, tap / a.size);
                    else tma2_load_2d(amap, dstA, &full[stage], cb * 64, m_tile * 128);
                    tma2_load_2d(bmap, sB + (size_t)stage * 16384, &full[stage], kb * 64, n_tile * a.block_n + (int)rank * half_n);
                    if (++stage == kFlowStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer (leader CTA; whole warp, elected lane) ========================
        if (leader) {
            const uint32_t idesc_base = (1u << 4) | (1u << 7) | (1u << 10) | ((256u >> 4) << 24);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            const uint64_t adesc0 = make_desc<64>(smem_u32(sA)), bdesc0 = make_desc<64>(smem_u32(sB));
            unsigned next_ent = k_begin < k_end ? __ldg(fp.sched + k_begin) : 0u;
            for (int k = k_begin; k < k_end; ++k) {
                const unsigned ent = next_ent;
                if (k + 1 < k_end) next_ent = __ldg(fp.sched + k + 1);
                const int l = (int)(ent >> 24);
                const size_t it = (size_t)L[l].item0 + (ent & 0xffffffu);
                const int block_n = L[l].block_n, nkb = L[l].num_kblocks;
                const uint32_t idesc = idesc_base | ((uint32_t)(block_n >> 3) << 17);
                mbar_wait(&tempty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    if (fp.trace && kb == 0 && lane == 0) fp.trace[5 * it + 1] = global_ns();
                    const uint64_t adesc = adesc0 + (uint64_t)((uint32_t)stage * (16384u >> 4));
                    const uint64_t bdesc = bdesc0 + (uint64_t)((uint32_t)stage * (16384u >> 4));
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        tc2_mma_bf16_elect(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
                    tc2_commit_both_elect(&empty[stage]);
                    if (++stage == kFlowStages) { stage = 0; phase ^= 1; }
                }
                tc2_commit_both_elect(&tfull[acc]);
                if (fp.trace && lane == 0) fp.trace[5 * it + 2] = global_ns();
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
        else {
            // ================================= scout (peer CTA's otherwise idle warp 1) ==================================
            // walks the pair's list AHEAD of the producers: polls the completion counters of everything item k reads (its A
            // input rows, one counter per lane, and its shortcut operand) and then publishes k + 1 in both CTAs.  The L2
            // round trips of the polls are off the producers' critical path.
            const uint32_t ready_addr = smem_u32(ready);
            for (int k = k_begin; k < k_end; ++k) {
                const unsigned ent = __ldg(fp.sched + k);
                const int l = (int)(ent >> 24), t = (int)(ent & 0xffffffu);
                const FlowLayerArgs &a = L[l];
                const int mp = t / a.n_tiles, it = a.item0 + t;
                if (a.dep >= 0) {
                    int jlo, jhi;
                    flow_dep_range(a, mp, jlo, jhi);
                    const unsigned target = fp.epoch * (unsigned)a.dep_unit;
                    for (int j = jlo + lane; j <= jhi; j += 32) flow_wait_counter(fp.done + a.dep_off + j, target, l, it, fp.stats, fp.stats + 2);
                }
                if (a.res_dep >= 0 && lane == 31) flow_wait_counter(fp.done + a.res_off + mp, fp.epoch * (unsigned)a.res_unit, l, it, fp.stats + 1, fp.stats + 2);
                asm volatile("fence.proxy.async;" ::: "memory");      // the acquired stores were made by other SMs' TMA engines
                __syncwarp();
                if (lane < 2) st_release_cluster(ready_addr, (uint32_t)lane, (unsigned)(k - k_begin + 1));
                __syncwarp();
            }
        }
    } else if (warp == 2) {
        // ===================================== store warp: slot -> TMA store; tile complete -> bump its counter =========
        if (lane == 0) {
            pdl_wait();
            int j = 0;
            for (int k = k_begin; k < k_end; ++k) {
                const unsigned ent = __ldg(fp.sched + k);
                const int l = (int)(ent >> 24), t = (int)(ent & 0xffffffu);
                const FlowLayerArgs &a = L[l];
                const int n_tile = t % a.n_tiles, mp = t / a.n_tiles;
                const int m_tile = 2 * mp + (int)rank;
                if (m_tile < a.m_tiles) {
                    const CUtensorMap *cmap = fp.maps + 4 * l + 2;
                    const int nsub = a.block_n >> 6, col0 = n_tile * a.block_n;
                    for (int q = 0; q < nsub; ++q, ++j) {
                        const int slot = j % kFlowRing;
                        MBAR_WAIT_HERE(&cwritten[slot], (j / kFlowRing) & 1);
                        tma_store_2d(cmap, sC + (size_t)slot * 16384, col0 + 64 * q, m_tile * 128);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        bulk_wait_read<0>();
                        mbar_arrive(&cempty[slot]);
                    }
                    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // the tile's bytes are in global memory
                    asm volatile("fence.proxy.async;" ::: "memory");
                }
                red_release_gpu(fp.done + a.done_off + mp, 1u);
                if (fp.trace && leader) fp.trace[5 * ((size_t)a.item0 + t) + 3] = global_ns();
            }
        }
    } else if (warp == 3) {
        // ===================================== residual loader / slot recycler ==========================================
        // every slot passes through here: a fused shortcut's sub-tile is TMA-loaded into it, otherwise it is handed on as is
        if (lane == 0) {
            pdl_wait();
            int j = 0;
            const uint32_t ready_addr = smem_u32(ready);
            for (int k = k_begin; k < k_end; ++k) {
                const unsigned ent = __ldg(fp.sched + k);
                const int l = (int)(ent >> 24), t = (int)(ent & 0xffffffu);
                const FlowLayerArgs &a = L[l];
                const int n_tile = t % a.n_tiles, mp = t / a.n_tiles;
                const int m_tile = 2 * mp + (int)rank;
                if (m_tile >= a.m_tiles) continue;
                const int nsub = a.block_n >> 6, col0 = n_tile * a.block_n;
                if (a.has_res && a.res_dep >= 0) {
                    while (ld_acquire_cluster_smem(ready_addr) < (unsigned)(k - k_begin + 1)) { }
                }
                const CUtensorMap *rmap = fp.maps + 4 * l + 3;
                for (int q = 0; q < nsub; ++q, ++j) {
                    const int slot = j % kFlowRing;
                    MBAR_WAIT_HERE(&cempty[slot], ((j / kFlowRing) & 1) ^ 1);
                    if (a.has_res) {
                        mbar_expect_tx(&cfull[slot], 16384u);
                        tma_load_2d(rmap, sC + (size_t)slot * 16384, &cfull[slot], col0 + 64 * q, m_tile * 128);
                    } else mbar_arrive(&cfull[slot]);
                }
            }
        }
    } else {
        // ===================================== epilogue groups (warps 4-11): even / odd sub-tiles =======================
        const int h = (warp - 4) >> 2;
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t sC_addr = smem_u32(sC);
        int acc = 0; uint32_t acc_phase = 0;
        int jbase = 0;
        for (int k = k_begin; k < k_end; ++k) {
            const unsigned ent = __ldg(fp.sched + k);
            const int l = (int)(ent >> 24), t = (int)(ent & 0xffffffu);
            const FlowLayerArgs &a = L[l];
            const int n_tile = t % a.n_tiles, mp = t / a.n_tiles;
            const int m_tile = 2 * mp + (int)rank;
            const bool real = m_tile < a.m_tiles;
            const int NSUB = a.block_n >> 6, col0 = n_tile * a.block_n;
            MBAR_WAIT_HERE(&tfull[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 256);
            if (!real || h >= NSUB) {
                ring_release<true>(tempty, acc, lane);
            } else {
                const bool leaky = a.act == ACT_LEAKY, has_res = a.has_res != 0;
                for (int q = h; q < NSUB; q += 2) {
                    const int j = jbase + q, slot = j % kFlowRing;
                    const uint32_t sphase = (uint32_t)(j / kFlowRing) & 1u;
                    const uint32_t slot_addr = sC_addr + (uint32_t)slot * 16384u;
                    uint32_t r[64];
                    tmem_ld32(taddr + 64 * q, r);
                    tmem_ld32(taddr + 64 * q + 32, r + 32);
                    tmem_ld_wait();
                    if (q + 2 >= NSUB) ring_release<true>(tempty, acc, lane);
                    MBAR_WAIT_HERE(&cfull[slot], sphase);
                    ring_emit<64, false>(r, a.scale + col0 + 64 * q, a.shift + col0 + 64 * q, slot_addr, row, leaky, has_res, a.res_alpha, a.res_beta);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&cwritten[slot]);
                }
            }
            if (real) jbase += NSUB;
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    cluster_sync_all();                      // the leader's MMAs read the peer's shared memory: nobody leaves early
    if (fp.trace && blockIdx.x == 0 && threadIdx.x == 0) { fp.stats[3] = (unsigned long long)(clock64() - clk0); fp.stats[4] = global_ns() - ns0; }
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

void conv_tc_flow_launch(ConvTcFlow *f, cudaStream_t s)
{
    static bool configured[64];
    if (first_use_on_this_device(configured))
        B200_CHECK(cudaFuncSetAttribute(conv_tc_flow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    static const bool no_pdl = getenv("B200_NO_PDL") != nullptr;
    f->fp.epoch += 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148);
    cfg.blockDim = dim3(kTcRingThreads);
    cfg.dynamicSmemBytes = f->smem_bytes;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    int n = 0;
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = 2; attr[n].val.clusterDim.y = 1; attr[n].val.clusterDim.z = 1;
    ++n;
    if (!no_pdl) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    cfg.attrs = attr; cfg.numAttrs = n;
    B200_CHECK(cudaLaunchKernelEx(&cfg, conv_tc_flow_kernel, f->fp));
}

void launch_conv_tc_flow(ConvTcFlow *f, cudaStream_t s)
{
    conv_tc_flow_launch(f, s);
    B200_LAUNCHED();
}
