// conv_tc*.{cu,cuh,h} — the fused implicit-GEMM convolution on 5th-generation tensor cores (sm_100a only).
//
// What it replaces: forward_convolutional_layer's im2col_cpu + gemm_nn + batch-norm + activate_array
// (convolutional_layer.c:445-485, im2col.c:16-39, gemm.c:74-89, batchnorm_layer.c:150-154) as ONE kernel:
//
//   D[pixel, filter] = sum over (ky,kx,c) of  X[n, oy*s+ky-pad, ox*s+kx-pad, c] * W[filter, ky, kx, c]
//   out = act(D * scale[filter] + shift[filter])            (folded inference BN / bias, leaky, cast)
//
//   * GEMM M = output pixels (128 per tile = the 128 TMEM lanes), N = filters (<= 256 per tile), K = taps x C_in.
//   * A operand: never materialised.  For every filter tap the TMA engine gathers a [pixels x BLOCK_K channels]
//     box of the NHWC activation straight into 128B/64B/32B-swizzled shared memory; image borders are the TMA
//     out-of-bounds zero fill (= im2col's zero padding), stride-2 layers read one of four parity-phase views of
//     the tensor.  1x1 layers use a dense 2-D [pixels][C] view, 3x3 layers a 4-D (C,W,H,N) view with a
//     rectangular (TW x TH x TN) pixel tile chosen per layer to fill the 128 rows.
//   * B operand: weights repacked [C_out][ky][kx][C_in] (K-major), 2-D TMA box [BLOCK_N x BLOCK_K].
//   * tcgen05.mma (kind::f16, bf16 x bf16 -> fp32) issued by one elected lane of a convergent warp, accumulators in TMEM,
//     multi-buffered so the epilogue of tile t overlaps the mainloops of the following tiles.
//   * Persistent: grid = min(tiles, #SM); tiles are walked filter-tile-fastest so co-running CTAs share A in L2.
//   * Launched with programmatic stream serialization: prologue and weight loads overlap the previous layer's tail.
//
// Files (one plan per layer, chosen by conv_tc_plan_create / conv_tc_block_plan_create in conv_tc_plan.cu):
//   conv_tc_common.cuh    kernel arguments, epilogue math, the ring-epilogue roles shared by the kernels
//   conv_tc.cu            conv_tc_kernel       1 CTA per tile, weights streamed or resident (small layers, 1x1 layers with <= 128 filters)
//                         conv_tc_pair_kernel  cta_group::2 CTA pair per 256-pixel tile, half of the weight tile per CTA (compute-heavy layers)
//   conv_tc_patch.cu      conv_tc_patch_kernel 3x3 layers with 16/32/64 input channels: halo patch per tile, ALL weights resident, HBM-bound
//                         conv_tc_block_kernel fused residual block 1x1 (64->32) + 3x3 (32->64) + shortcut: the intermediate never leaves the SM
//   conv_tc_plan.{h,cu}   tensor maps, the per-layer planner, launch dispatch
// Epilogues: ring_roles (store warp + residual loader + two epilogue groups over a ring of swizzled 64-filter sub-tile
// slots, TMA stores; also writes a fused 2x upsample) for everything that can be staged, run_epilogue (serial, direct
// 16-byte stores) for odd filter counts.
//
// Roofline: tensor pipe.  FLOPs per launch = 2 * pixels * C_out * K (darknet's own BFLOPs formula,
// convolutional_layer.c:325).
#include "conv_tc_plan.h"

// im2col-mode A tiles: tile m covers output pixels [128 m, 128 m + 128) in (n, oy, ox) order; the load starts at the input
// position of the first pixel's tap origin.  A phantom tile (second CTA of an odd pair count) re-reads the last real one.
__device__ __forceinline__ void im2col_origin(const ConvTcArgs &args, int m_tile, int &w0, int &h0, int &n0)
{
    if (m_tile >= args.m_tiles) m_tile = args.m_tiles - 1;
    const int per_image = args.OH * args.OW;
    const int p0 = m_tile * 128;
    n0 = p0 / per_image;
    const int r = p0 - n0 * per_image;
    const int oy = r / args.OW;
    h0 = oy * args.stride - args.pad;
    w0 = (r - oy * args.OW) * args.stride - args.pad;
}

template <int BLOCK_K, typename OutT, bool RING>      // RING: ring epilogue roles (384 threads); else the serial epilogue (192)
__global__ void __launch_bounds__(RING ? kTcRingThreads : kTcThreads, 1)
conv_tc_kernel(const __grid_constant__ ConvTcMaps maps, const ConvTcArgs args)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    constexpr int A_BYTES = 128 * BLOCK_K * 2;
    const int stages = args.stages;
    uint8_t *sA = smem;
    uint8_t *sB = smem + (size_t)stages * A_BYTES;
    const int b_slots = args.resident_b ? args.num_kblocks : stages;
    uint8_t *sC = sB + (size_t)b_slots * args.b_stage_bytes;       // staged-epilogue tile: (block_n/64) x 16 KB, 1024-aligned
    uint8_t *aux = sC + (args.ring ? (size_t)args.c_bufs * 16384 : (args.staged ? (size_t)(args.block_n >> 6) * 16384 : 0));
    uint64_t *ring_bars = (uint64_t *)(aux + 384);     // ring epilogue: cfull[4], cempty[4], cwritten[4]
    uint64_t *full = (uint64_t *)aux;                 // [stages]
    uint64_t *empty = full + 8;                       // [stages]
    uint64_t *tfull = empty + 8;                      // [acc_stages <= 8]
    uint64_t *tempty = tfull + 8;                     // [acc_stages <= 8]
    uint64_t *wfull = tempty + 8;                     // resident weights landed
    uint64_t *rfull = wfull + 1;                      // residual tile landed (staged epilogue)
    uint32_t *tmem_slot = (uint32_t *)(rfull + 1);
    float *s_scale = (float *)(aux + 512);            // [2][256]
    float *s_shift = s_scale + 512;                   // [2][256]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_tiles = args.m_tiles * args.n_tiles;

    if (threadIdx.x == 0) {
        for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        const int ep_warps = args.ring ? 8 : 4;             // both ring groups hand the accumulator back, even an idle one
        for (int i = 0; i < args.acc_stages; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], ep_warps); }
        if (args.ring)
            for (int i = 0; i < 4; ++i) { mbar_init(&ring_bars[i], 1); mbar_init(&ring_bars[4 + i], 1); mbar_init(&ring_bars[8 + i], 4); }
        mbar_init(wfull, 1);
        mbar_init(rfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(args.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();

    if (warp == 0) {
        // ===================================== TMA producer =====================================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            const uint32_t b_bytes = (uint32_t)(args.block_n * BLOCK_K * 2);
            const uint32_t tx_bytes = (uint32_t)(args.a_rows * BLOCK_K * 2) + (args.resident_b ? 0u : b_bytes);
            if (args.resident_b && (int)blockIdx.x < num_tiles) {          // n_tiles == 1: one slab serves every tile
                mbar_expect_tx(wfull, b_bytes * (uint32_t)args.num_kblocks);
                for (int kb = 0; kb < args.num_kblocks; ++kb)
                    tma_load_2d(&maps.b, sB + (size_t)kb * args.b_stage_bytes, wfull, kb * BLOCK_K, 0);
            }
            pdl_wait();                                   // weights may load early; activations only after the previous layer is done
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int n_tile = tile % args.n_tiles, m_tile = tile / args.n_tiles;
                int ox0 = 0, oy0 = 0, n0 = 0;
                if (args.mode == 1) {
                    int tx = m_tile % args.tiles_x, ty = (m_tile / args.tiles_x) % args.tiles_y, tn = m_tile / (args.tiles_x * args.tiles_y);
                    ox0 = tx * args.TW; oy0 = ty * args.TH; n0 = tn * args.TN;
                } else if (args.im2col) im2col_origin(args, m_tile, ox0, oy0, n0);
                for (int kb = 0; kb < args.num_kblocks; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_expect_tx(&full[stage], tx_bytes);
                    const int tap = kb / args.cin_blocks, cb = kb - tap * args.cin_blocks;
                    void *dstA = sA + (size_t)stage * A_BYTES;
                    if (args.im2col) {
                        tma_load_im2col_4d(&maps.a[0], dstA, &full[stage], cb * BLOCK_K, ox0, oy0, n0, tap % args.size, tap / args.size);
                    } else if (args.mode == 0) {
                        tma_load_2d(&maps.a[0], dstA, &full[stage], cb * BLOCK_K, m_tile * 128);
                    } else {
                        const int ky = tap / args.size, kx = tap - ky * args.size;
                        int dy = ky - args.pad, dx = kx - args.pad;
                        if (args.stride == 1) {
                            tma_load_4d(&maps.a[0], dstA, &full[stage], cb * BLOCK_K, ox0 + dx, oy0 + dy, n0);
                        } else {                  // stride 2: input x = 2*ox + dx  ->  phase dx&1, offset floor(dx/2)
                            int px = dx & 1, py = dy & 1;
                            int xoff = (dx - px) / 2, yoff = (dy - py) / 2;
                            tma_load_4d(&maps.a[py * 2 + px], dstA, &full[stage], cb * BLOCK_K, ox0 + xoff, oy0 + yoff, n0);
                        }
                    }
                    if (args.local)              // unshared convolution: this location's own weight slab
                        tma_load_3d(&maps.b, sB + (size_t)stage * args.b_stage_bytes, &full[stage], kb * BLOCK_K, n_tile * args.block_n,
                                    m_tile % (args.tiles_x * args.tiles_y));
                    else if (!args.resident_b)
                        tma_load_2d(&maps.b, sB + (size_t)stage * args.b_stage_bytes, &full[stage], kb * BLOCK_K, n_tile * args.block_n);
                    if (++stage == stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer =======================================
        // the whole warp walks the loop and one elected lane issues (see tc_mma_bf16_elect): with N <= 128 an MMA is
        // 32-64 cycles of tensor time, less than the ~80 cycles a lane-0-only issue sequence costs
        {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(args.block_n >> 3) << 17) | ((128u >> 4) << 24);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            const uint64_t adesc0 = make_desc<BLOCK_K>(smem_u32(sA)), bdesc0 = make_desc<BLOCK_K>(smem_u32(sB));
            const uint32_t a_step = (uint32_t)A_BYTES >> 4, b_step = (uint32_t)args.b_stage_bytes >> 4;
            if (args.resident_b && (int)blockIdx.x < num_tiles) mbar_wait(wfull, 0);
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait(&tempty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * args.block_n);
                for (int kb = 0; kb < args.num_kblocks; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint64_t adesc = adesc0 + (uint64_t)((uint32_t)stage * a_step);
                    const uint64_t bdesc = bdesc0 + (uint64_t)((uint32_t)(args.resident_b ? kb : stage) * b_step);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / 16; ++k)
                        tc_mma_bf16_elect(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
                    tc_commit_elect(&empty[stage]);           // frees the smem slot once these MMAs have read it
                    if (++stage == stages) { stage = 0; phase ^= 1; }
                }
                tc_commit_elect(&tfull[acc]);                 // accumulator complete -> epilogue
                if (++acc == args.acc_stages) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===================================== epilogue (warps 2..5) ============================
        if constexpr (RING) ring_roles<false>(maps, args, tfull, tempty, ring_bars, sC, tmem_base, blockIdx.x, gridDim.x, num_tiles, 0);
        else run_epilogue<OutT, false>(maps, args, tfull, tempty, rfull, s_scale, s_shift, sC, tmem_base, blockIdx.x, gridDim.x, num_tiles, 0);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(args.tmem_cols) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------
// cta_group::2 variant: a CTA pair (two SMs of one TPC) computes a 256-pixel x BLOCK_N tile.  Each CTA loads its own
// 128 pixel rows of A and HALF of the weight tile; tcgen05.mma.cta_group::2 (issued by the leader only) reads both
// halves, so the weight bytes pulled through L2 per FLOP are halved — the measured bottleneck of the 1-CTA kernel.
// Barriers: full[s] lives in the leader and collects the TMA bytes of both CTAs; empty[s] / tfull[a] are signalled
// in both CTAs by a multicast tcgen05.commit; tempty[a] lives in the leader and collects the 8 epilogue warps.
// ---------------------------------------------------------------------------------------------------
template <int BLOCK_K, typename OutT, bool RING>
__global__ void __launch_bounds__(RING ? kTcRingThreads : kTcThreads, 1)
conv_tc_pair_kernel(const __grid_constant__ ConvTcMaps maps, const ConvTcArgs args)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    constexpr int A_BYTES = 128 * BLOCK_K * 2;
    const int stages = args.stages;
    uint8_t *sA = smem;
    uint8_t *sB = smem + (size_t)stages * A_BYTES;
    uint8_t *sC = sB + (size_t)stages * args.b_stage_bytes;        // b_stage_bytes = half tile here
    uint8_t *aux = sC + (args.ring ? (size_t)args.c_bufs * 16384 : (args.staged ? (size_t)(args.block_n >> 6) * 16384 : 0));
    uint64_t *ring_bars = (uint64_t *)(aux + 384);     // ring epilogue: cfull[4], cempty[4], cwritten[4]
    uint64_t *full = (uint64_t *)aux;
    uint64_t *empty = full + 8;
    uint64_t *tfull = empty + 8;
    uint64_t *tempty = tfull + 8;
    uint64_t *rfull = tempty + 9;
    uint32_t *tmem_slot = (uint32_t *)(rfull + 1);
    float *s_scale = (float *)(aux + 512);
    float *s_shift = s_scale + 512;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    const int m_pairs = (args.m_tiles + 1) / 2;
    const int num_tiles = RING ? args.vtiles : m_pairs * args.n_tiles;   // (virtual) pair tiles: the last wave may be filter-sliced

    if (threadIdx.x == 0) {
        for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        const int ep_warps = args.ring ? 16 : 8;            // epilogue warps of both CTAs
        for (int i = 0; i < args.acc_stages; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], ep_warps); }
        if (args.ring)
            for (int i = 0; i < 4; ++i) { mbar_init(&ring_bars[i], 1); mbar_init(&ring_bars[4 + i], 1); mbar_init(&ring_bars[8 + i], 4); }
        mbar_init(rfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(args.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();                                             // peer barriers are initialised before anyone signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();
    const int half_n = args.block_n / 2;

    if (warp == 0) {
        // ===================================== TMA producer (both CTAs) =========================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            pdl_wait();
            for (int vt = pair_id; vt < num_tiles; vt += num_pairs) {
                VTile v; v.tile = vt; v.col_off = 0; v.width = args.block_n; v.k0 = 0; v.k1 = args.num_kblocks; v.row_off = 0;
                if (RING) v = vtile_of(args, vt);
                const int tile = v.tile, half_w = v.width / 2;
                const uint32_t tx_bytes = 2u * (uint32_t)(args.a_rows * BLOCK_K * 2 + half_w * BLOCK_K * 2);
                const CUtensorMap *bmap = v.width == args.block_n ? &maps.b : &maps.b2;
                const int n_tile = tile % args.n_tiles, m_tile = 2 * (tile / args.n_tiles) + (int)rank;
                int ox0 = 0, oy0 = 0, n0 = 0;
                if (args.mode == 1) {
                    int tx = m_tile % args.tiles_x, ty = (m_tile / args.tiles_x) % args.tiles_y, tn = m_tile / (args.tiles_x * args.tiles_y);
                    ox0 = tx * args.TW; oy0 = ty * args.TH; n0 = tn * args.TN;      // a phantom last tile lands past the batch: TMA zero-fills
                } else if (args.im2col) im2col_origin(args, m_tile, ox0, oy0, n0);
                for (int kb = v.k0; kb < v.k1; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
#ifdef B200_EXPERIMENTS                            // make EXPERIMENTS=1: scripts/operand_traffic_probe.py
                    if (args.exp) {               // timing experiment: skip operand loads once every stage holds something
                        const bool warm = vt != pair_id || kb >= stages;
                        const bool skip_b = warm && (args.exp & 1), skip_a = warm && (args.exp & 2);
                        const uint32_t bytes = 2u * (uint32_t)((skip_a ? 0 : args.a_rows * BLOCK_K * 2) + (skip_b ? 0 : half_w * BLOCK_K * 2));
                        if (leader) { if (bytes) mbar_expect_tx(&full[stage], bytes); else mbar_arrive(&full[stage]); }
                        const int tap = kb / args.cin_blocks, cb = kb - tap * args.cin_blocks;
                        // bit 3 (8): A lands in a buffer the tensor core never reads (a ring-epilogue slot); bit 4 (16): B is issued before A
                        uint8_t *dstA_x = (args.exp & 8) ? sC + 2 * 16384 : sA + (size_t)stage * A_BYTES;
                        if ((args.exp & 16) && !skip_b) tma2_load_2d(bmap, sB + (size_t)stage * args.b_stage_bytes, &full[stage], kb * BLOCK_K, n_tile * args.block_n + v.col_off + (int)rank * half_w);
                        if (!skip_a) {
                            // bit 2: every CTA loads the SAME A rows (pixel tile 0 / 1): private versus shared lines, same bytes per SM
                            if (args.im2col) tma2_load_im2col_4d(&maps.a[0], dstA_x, &full[stage], cb * BLOCK_K, (args.exp & 4) ? -args.pad : ox0,
                                                                 (args.exp & 4) ? -args.pad : oy0, (args.exp & 4) ? (int)rank : n0, tap % args.size, tap / args.size);
                            else if (args.mode == 0) tma2_load_2d(&maps.a[0], dstA_x, &full[stage], cb * BLOCK_K, ((args.exp & 4) ? (int)rank : m_tile) * 128);
                            else tma2_load_4d(&maps.a[0], sA + (size_t)stage * A_BYTES, &full[stage], cb * BLOCK_K, ox0 + tap % args.size - args.pad, oy0 + tap / args.size - args.pad, n0);
                        }
                        if (!(args.exp & 16) && !skip_b) tma2_load_2d(bmap, sB + (size_t)stage * args.b_stage_bytes, &full[stage], kb * BLOCK_K, n_tile * args.block_n + v.col_off + (int)rank * half_w);
                        if (++stage == stages) { stage = 0; phase ^= 1; }
                        continue;
                    }
#endif
                    if (leader) mbar_expect_tx(&full[stage], tx_bytes);
                    const int tap = kb / args.cin_blocks, cb = kb - tap * args.cin_blocks;
                    void *dstA = sA + (size_t)stage * A_BYTES;
                    if (args.im2col) {
                        tma2_load_im2col_4d(&maps.a[0], dstA, &full[stage], cb * BLOCK_K, ox0, oy0, n0, tap % args.size, tap / args.size);
                    } else if (args.mode == 0) {
                        tma2_load_2d(&maps.a[0], dstA, &full[stage], cb * BLOCK_K, m_tile * 128);
                    } else {
                        const int ky = tap / args.size, kx = tap - ky * args.size;
                        int dy = ky - args.pad, dx = kx - args.pad;
                        if (args.stride == 1) {
                            tma2_load_4d(&maps.a[0], dstA, &full[stage], cb * BLOCK_K, ox0 + dx, oy0 + dy, n0);
                        } else {
                            int px = dx & 1, py = dy & 1;
                            int xoff = (dx - px) / 2, yoff = (dy - py) / 2;
                            tma2_load_4d(&maps.a[py * 2 + px], dstA, &full[stage], cb * BLOCK_K, ox0 + xoff, oy0 + yoff, n0);
                        }
                    }
                    tma2_load_2d(bmap, sB + (size_t)stage * args.b_stage_bytes, &full[stage], kb * BLOCK_K,
                                 n_tile * args.block_n + v.col_off + (int)rank * half_w);
                    if (++stage == stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer (leader CTA only; whole warp, elected lane) =====
        if (leader) {
            const uint32_t idesc_base = (1u << 4) | (1u << 7) | (1u << 10) | ((256u >> 4) << 24);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            const uint64_t adesc0 = make_desc<BLOCK_K>(smem_u32(sA)), bdesc0 = make_desc<BLOCK_K>(smem_u32(sB));
            const uint32_t a_step = (uint32_t)A_BYTES >> 4, b_step = (uint32_t)args.b_stage_bytes >> 4;
            for (int vt = pair_id; vt < num_tiles; vt += num_pairs) {
                VTile v; v.width = args.block_n; v.k0 = 0; v.k1 = args.num_kblocks;
                if (RING) v = vtile_of(args, vt);
                const uint32_t idesc = idesc_base | ((uint32_t)(v.width >> 3) << 17);
                mbar_wait(&tempty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * args.block_n);
                for (int kb = v.k0; kb < v.k1; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint64_t adesc = adesc0 + (uint64_t)((uint32_t)stage * a_step);
                    const uint64_t bdesc = bdesc0 + (uint64_t)((uint32_t)stage * b_step);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / 16; ++k)
                        tc2_mma_bf16_elect(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, ((kb - v.k0) | k) != 0 ? 1u : 0u);
                    tc2_commit_both_elect(&empty[stage]);
                    if (++stage == stages) { stage = 0; phase ^= 1; }
                }
                tc2_commit_both_elect(&tfull[acc]);
                if (++acc == args.acc_stages) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===================================== epilogue (warps 2..5, both CTAs) =================
        if constexpr (RING) ring_roles<true>(maps, args, tfull, tempty, ring_bars, sC, tmem_base, pair_id, num_pairs, num_tiles, (int)rank);
        else run_epilogue<OutT, true>(maps, args, tfull, tempty, rfull, s_scale, s_shift, sC, tmem_base, pair_id, num_pairs, num_tiles, (int)rank);
    }

    tc_fence_before();
    cluster_sync_all();                      // the leader's MMAs read the peer's shared memory: nobody leaves early
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(args.tmem_cols) : "memory");
    }
}

template <int BLOCK_K, typename OutT> static void launch_variant(ConvTcPlan *p, cudaStream_t s)
{
    static bool configured[64];
    if (first_use_on_this_device(configured)) {
        B200_CHECK(cudaFuncSetAttribute(conv_tc_kernel<BLOCK_K, OutT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        B200_CHECK(cudaFuncSetAttribute(conv_tc_kernel<BLOCK_K, bf16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    if (p->args.ring) launch_pdl(conv_tc_kernel<BLOCK_K, bf16, true>, p->grid, kTcRingThreads, p->smem_bytes, s, 1, p->maps, p->args);
    else launch_pdl(conv_tc_kernel<BLOCK_K, OutT, false>, p->grid, kTcThreads, p->smem_bytes, s, 1, p->maps, p->args);
}

template <int BLOCK_K, typename OutT> static void launch_pair_variant(ConvTcPlan *p, cudaStream_t s)
{
    static bool configured[64];
    if (first_use_on_this_device(configured)) {
        B200_CHECK(cudaFuncSetAttribute(conv_tc_pair_kernel<BLOCK_K, OutT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        B200_CHECK(cudaFuncSetAttribute(conv_tc_pair_kernel<BLOCK_K, bf16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    if (p->args.ring) launch_pdl(conv_tc_pair_kernel<BLOCK_K, bf16, true>, p->grid, kTcRingThreads, p->smem_bytes, s, 2, p->maps, p->args);
    else launch_pdl(conv_tc_pair_kernel<BLOCK_K, OutT, false>, p->grid, kTcThreads, p->smem_bytes, s, 2, p->maps, p->args);
}

// ---------------------------------------------------------------------------------------------------
// split-K finalize: out[p][c] = act(scale[c] * sum_s ws[s][p][c] + shift[c]) (+ residual), bf16.  The slabs are summed in a
// fixed order (s = 0, 1, ...): deterministic, unlike atomics.  One thread per 8 channels of one pixel; the workspace is
// L2-resident (a few MB).  Replaces the epilogue of convolutional_layer.c:445-485 for launches with too few tiles to fill
// the chip (small batches), where cutting K shortens the layer's critical path (13x13 3x3 512->1024 at batch 1: 72 k-blocks).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
splitk_finalize_kernel(const float *__restrict__ ws, int ksplit, long long slab_elems, long long npix, int cout_pad,
                       const float *__restrict__ scale, const float *__restrict__ shift, int leaky,
                       const bf16 *__restrict__ res, int ldr, float alpha, float beta, bf16 *__restrict__ out, int ldo)
{
    pdl_launch_dependents();
    pdl_wait();
    const int groups = cout_pad >> 3;
    const long long total = npix * groups;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / groups;
        const int c = (int)(i - p * groups) << 3;
        const float *src = ws + p * cout_pad + c;
        float4 a = *reinterpret_cast<const float4 *>(src), b = *reinterpret_cast<const float4 *>(src + 4);
        for (int s = 1; s < ksplit; ++s) {
            const float4 a2 = *reinterpret_cast<const float4 *>(src + s * slab_elems), b2 = *reinterpret_cast<const float4 *>(src + s * slab_elems + 4);
            a.x += a2.x; a.y += a2.y; a.z += a2.z; a.w += a2.w; b.x += b2.x; b.y += b2.y; b.z += b2.z; b.w += b2.w;
        }
        float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        const float4 s0 = *reinterpret_cast<const float4 *>(scale + c), s1 = *reinterpret_cast<const float4 *>(scale + c + 4);
        const float4 h0 = *reinterpret_cast<const float4 *>(shift + c), h1 = *reinterpret_cast<const float4 *>(shift + c + 4);
        const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w}, sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            v[q] = fmaf(v[q], sc[q], sh[q]);
            if (leaky) v[q] = v[q] > 0.f ? v[q] : 0.1f * v[q];
        }
        if (res) {
            float r[8];
            load_vec<bf16>(res + p * ldr + c, r);
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = fmaf(alpha, v[q], beta * r[q]);
        }
        store_vec<bf16>(out + p * ldo + c, v);
    }
}

void conv_tc_launch_splitk_finalize(ConvTcPlan *p, cudaStream_t s)
{
    const ConvTcPlan::SplitK &k = p->sk;
    const long long total = k.npix * (k.cout_pad >> 3);
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    static const bool no_pdl = getenv("B200_NO_PDL") != nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(256); cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = no_pdl ? 0 : 1;
    B200_CHECK(cudaLaunchKernelEx(&cfg, splitk_finalize_kernel, (const float *)k.ws, p->args.ksplit, (long long)p->args.slab_rows * k.cout_pad, k.npix, k.cout_pad,
                                  k.scale, k.shift, k.act == ACT_LEAKY ? 1 : 0, k.res, k.ldr, k.res_alpha, k.res_beta, k.out, k.ldo));
}

void conv_tc_launch_tap(ConvTcPlan *p, cudaStream_t s)
{
    if (p->args.pair) {
        if (p->out_dtype == DT_BF16) {
            if (p->block_k == 64) launch_pair_variant<64, bf16>(p, s);
            else if (p->block_k == 32) launch_pair_variant<32, bf16>(p, s);
            else launch_pair_variant<16, bf16>(p, s);
        } else {
            if (p->block_k == 64) launch_pair_variant<64, float>(p, s);
            else if (p->block_k == 32) launch_pair_variant<32, float>(p, s);
            else launch_pair_variant<16, float>(p, s);
        }
        return;
    }
    if (p->out_dtype == DT_BF16) {
        if (p->block_k == 64) launch_variant<64, bf16>(p, s);
        else if (p->block_k == 32) launch_variant<32, bf16>(p, s);
        else launch_variant<16, bf16>(p, s);
    } else {
        if (p->block_k == 64) launch_variant<64, float>(p, s);
        else if (p->block_k == 32) launch_variant<32, float>(p, s);
        else launch_variant<16, float>(p, s);
    }
}
