// conv_tc_patch.cu — the halo-patch kernels of the tcgen05 convolution family (overview in conv_tc.cu): conv_tc_patch_kernel for
// 3x3 layers with few input channels and conv_tc_block_kernel, the fused residual block of the widest feature maps.
#include "conv_tc_plan.h"

// ---------------------------------------------------------------------------------------------------
// single-CTA patch kernel for the layers with few input channels (Cin = 32 / 64, 3x3).  The tap-per-box kernels above
// pull size^2 activation rows per output pixel through TMA (64-byte rows when Cin = 32) and are bound by the TMA row
// rate (~3 cycles a row), not by the tensor pipe.  Here a tile's input arrives ONCE as a halo patch, every tap's A
// operand is a row-shifted descriptor into it (see the CTA-pair patch kernel), and ALL weights stay resident, so a
// tile costs ~1.5 activation rows per output pixel and the layer becomes HBM-bound.
//   stride 1: one patch [(TH+2) x (TW+2)] pixels, 9 K-segments of Cin.
//   stride 2, Cin = 32: the input is viewed as rows of PIXEL PAIRS (2 x 32 channels = 128 bytes) of one row parity;
//   taps kx = 1,2 are one K = 64 segment of pair ox, tap kx = 0 is the upper half (K = 32) of pair ox - 1.
// The segment table is built by the host (ConvTcArgs::seg_*).
// ---------------------------------------------------------------------------------------------------
// persistent-tile walker: tile = first, first + step, ... decoded into (tx, ty, tn) without a division per tile
struct TileWalk {
    int tx, ty, tn, sx, sy, sn;
    __device__ __forceinline__ void init(int first, int step, int tiles_x, int tiles_y)
    {
        tx = first % tiles_x; ty = (first / tiles_x) % tiles_y; tn = first / (tiles_x * tiles_y);
        sx = step % tiles_x;  sy = (step / tiles_x) % tiles_y;  sn = step / (tiles_x * tiles_y);
    }
    __device__ __forceinline__ void next(int tiles_x, int tiles_y)
    {
        tx += sx; if (tx >= tiles_x) { tx -= tiles_x; ++ty; }
        ty += sy; if (ty >= tiles_y) { ty -= tiles_y; ++tn; }
        tn += sn;
    }
};

__device__ __forceinline__ void group_sync(int group)            // the 128 threads of one epilogue group
{
    asm volatile("bar.sync %0, 128;" ::"r"(group + 1) : "memory");
}

// Warp roles: warp 0 = TMA producer (patches + residual tiles), warp 1 = MMA issuer, warp 2 = TMA store warp, then
// `ep_groups` groups of four epilogue warps that take alternate tiles.  The output leaves through a ring of `c_bufs`
// swizzled tiles: a tile's residual is TMA-loaded into its ring slot tiles ahead of time, the epilogue adds the activation
// in place, the store warp TMA-stores the slot and frees it once the store engine has read it.  With a K pass this short
// (18 MMAs) everything else on a tile's path has to be off the critical path: no role waits on a DRAM round trip or on
// another role's bookkeeping, and the roles talk through mbarriers only.
template <int NSUB, int NSEG, int KS0, int KS1>       // K segment s issues (s odd ? KS1 : KS0) K=16 steps
__global__ void __launch_bounds__(352, 1)
conv_tc_patch_kernel(const __grid_constant__ ConvTcMaps maps, const ConvTcArgs args)
{
    // NSUB = 64-filter sub-tiles per tile; NSUB = 0 stands for a single 32-filter sub-tile (64-byte output rows, 64B swizzle)
    constexpr int N = NSUB ? NSUB * 64 : 32;
    constexpr int SUBS = NSUB ? NSUB : 1;                            // sub-tiles per ring slot
    constexpr int SUBC = NSUB ? 64 : 32;                             // filters per sub-tile
    constexpr int RB = SUBC * 2;                                     // bytes per pixel row of a sub-tile = its swizzle span
    constexpr int SUBT = 128 * RB;                                   // bytes per sub-tile
    constexpr int SLOT = (SUBS * SUBT + 1023) / 1024 * 1024;         // ring slot pitch
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int stages = args.stages;
    uint8_t *sA = smem;                                              // patch ring
    uint8_t *sB = sA + (size_t)stages * args.a_stage_bytes;          // nb resident weight tiles
    uint8_t *sC = sB + (size_t)args.nb * args.b_stage_bytes;         // output ring: c_bufs slots of SUBS sub-tiles
    uint8_t *aux = sC + (size_t)args.c_bufs * ((SUBS * SUBT + 1023) / 1024 * 1024);
    uint64_t *full = (uint64_t *)aux;                                // [8]
    uint64_t *empty = full + 8;                                      // [8]
    uint64_t *tfull = empty + 8;                                     // [8]
    uint64_t *tempty = tfull + 8;                                    // [8]
    uint64_t *cfull = tempty + 8;                                    // [8] residual landed in ring slot
    uint64_t *cempty = cfull + 8;                                    // [8] ring slot read out by its store
    uint64_t *cwritten = cempty + 8;                                 // [8] ring slot written by the four epilogue warps
    uint64_t *wfull = cwritten + 8;
    uint32_t *tmem_slot = (uint32_t *)(wfull + 1);
    float *s_scale = (float *)(aux + 512);
    float *s_shift = s_scale + 256;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_tiles = args.m_tiles;
    const int G = args.ep_groups, NBUF = args.c_bufs;
    // filter split: CTA b computes filters [ch0, ch0 + N) of pixel tiles vb, vb + vgrid, ...; neighbouring CTAs take the
    // filter slices of the SAME pixel tile at the same time, so the second read of its patch is an L2 hit
    const int vb = (int)blockIdx.x / args.n_split, vgrid = (int)gridDim.x / args.n_split;
    const int ch0 = ((int)blockIdx.x % args.n_split) * N;
    const bool has_res = args.res != nullptr;

    if (threadIdx.x == 0) {
        for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < args.acc_stages; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
        for (int i = 0; i < NBUF; ++i) { mbar_init(&cfull[i], 1); mbar_init(&cempty[i], 1); mbar_init(&cwritten[i], 4); }
        mbar_init(wfull, 1);
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
        if (lane == 0 && vb < num_tiles) {
            mbar_expect_tx(wfull, (uint32_t)(args.nb * N * args.b_k * 2));
            for (int i = 0; i < args.nb; ++i) tma_load_2d(&maps.b, sB + (size_t)i * args.b_stage_bytes, wfull, args.b_koff[i], ch0);
            pdl_wait();                                   // weights load early; activations only after the previous layer is done
            int stage = 0; uint32_t phase = 0;
            int cb = 0; uint32_t cphase = 0;
            TileWalk t; t.init(vb, vgrid, args.tiles_x, args.tiles_y);
            for (int tile = vb; tile < num_tiles; tile += vgrid, t.next(args.tiles_x, args.tiles_y)) {
                const int ox0 = t.tx * args.halo_TWv, oy0 = t.ty * args.halo_THv;
                MBAR_WAIT_HERE(&empty[stage], phase ^ 1);
                mbar_expect_tx(&full[stage], (uint32_t)args.stage_tx);
                uint8_t *dst = sA + (size_t)stage * args.a_stage_bytes;
                for (int q = 0; q < args.np; ++q)
                    tma_load_4d(&maps.a[args.patch_map[q]], dst + args.patch_off[q], &full[stage], 0, ox0 + args.patch_dx[q], oy0 + args.patch_dy[q], t.tn);
                if (++stage == stages) { stage = 0; phase ^= 1; }
                if (has_res) {
                    int rows_here = args.OH - oy0; if (rows_here > args.halo_THv) rows_here = args.halo_THv;
                    MBAR_WAIT_HERE(&cempty[cb], cphase ^ 1);
                    mbar_expect_tx(&cfull[cb], (uint32_t)(rows_here * args.halo_TWv * RB * SUBS));
                    uint8_t *cdst = sC + (size_t)cb * SLOT;
                    for (int q = 0; q < SUBS; ++q)
                        for (int yy = 0; yy < rows_here; ++yy)
                            tma_load_4d(&maps.r, cdst + q * SUBT + yy * args.halo_P * RB, &cfull[cb], ch0 + SUBC * q, ox0, oy0 + yy, t.tn);
                    if (++cb == NBUF) { cb = 0; cphase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer =======================================
        if (vb < num_tiles) {                 // all 32 lanes walk the loop; one elected lane issues
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            // every descriptor is a constant offset from the stage's base descriptor: keep the offsets in registers and
            // unroll the whole K pass, so issuing a tile is 18 back-to-back MMAs with no loads or address math between them
            uint32_t a_off[NSEG]; uint64_t bdesc[NSEG];
#pragma unroll
            for (int sgm = 0; sgm < NSEG; ++sgm) {
                a_off[sgm] = (uint32_t)args.seg_a[sgm] >> 4;
                bdesc[sgm] = make_desc_rt(smem_u32(sB) + (uint32_t)args.seg_b[sgm], args.b_k);
            }
            const uint64_t adesc0 = make_desc_rt(smem_u32(sA), args.a_k);
            const uint32_t stage_step = (uint32_t)args.a_stage_bytes >> 4;
            MBAR_WAIT_HERE(wfull, 0);
            for (int tile = vb; tile < num_tiles; tile += vgrid) {
                MBAR_WAIT_HERE(&tempty[acc], acc_phase ^ 1);
                MBAR_WAIT_HERE(&full[stage], phase);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * N);
                const uint64_t adesc = adesc0 + (uint64_t)((uint32_t)stage * stage_step);
#pragma unroll
                for (int sgm = 0; sgm < NSEG; ++sgm) {
#pragma unroll
                    for (int k = 0; k < ((sgm & 1) ? KS1 : KS0); ++k)
                        tc_mma_bf16_elect(d_tmem, adesc + (uint64_t)(a_off[sgm] + 2 * k), bdesc[sgm] + (uint64_t)(2 * k), idesc, (sgm | k) != 0 ? 1u : 0u);
                }
                tc_commit_elect(&empty[stage]);
                tc_commit_elect(&tfull[acc]);
                if (++stage == stages) { stage = 0; phase ^= 1; }
                if (++acc == args.acc_stages) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp == 2) {
        // ===================================== store warp =======================================
        // takes a ring slot once the four epilogue warps have written it, TMA-stores it (one box per image row: the tile
        // is patch-pitched in smem) and frees the slot when the store engine has read it
        if (lane == 0) {
            pdl_wait();
            TileWalk t; t.init(vb, vgrid, args.tiles_x, args.tiles_y);
            int cb = 0; uint32_t cphase = 0;
            for (int tile = vb; tile < num_tiles; tile += vgrid, t.next(args.tiles_x, args.tiles_y)) {
                const int ox0 = t.tx * args.halo_TWv, oy0 = t.ty * args.halo_THv;
                int rows_here = args.OH - oy0; if (rows_here > args.halo_THv) rows_here = args.halo_THv;
                const uint8_t *src = sC + (size_t)cb * SLOT;
                MBAR_WAIT_HERE(&cwritten[cb], cphase);
                for (int q = 0; q < SUBS; ++q)
                    for (int yy = 0; yy < rows_here; ++yy)
                        tma_store_4d(&maps.c, src + q * SUBT + yy * args.halo_P * RB, ch0 + SUBC * q, ox0, oy0 + yy, t.tn);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                bulk_wait_read<0>();
                mbar_arrive(&cempty[cb]);
                if (++cb == NBUF) { cb = 0; cphase ^= 1; }
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    } else {
        // ===================================== epilogue groups ==================================
        const int g = (warp - 3) >> 2;
        const int quarter = warp & 3;                      // TMEM lane quarter this warp may touch
        const int row = quarter * 32 + lane;
        const int ep_tid = threadIdx.x - 96 - 128 * g;
        for (int c = ep_tid; c < N; c += 128) { s_scale[c] = args.scale[ch0 + c]; s_shift[c] = args.shift[ch0 + c]; }   // every group writes the same values
        group_sync(g);
        const bool leaky = args.act == ACT_LEAKY;
        const float alpha = args.res_alpha, beta = args.res_beta;
        const uint32_t scale_addr = smem_u32(s_scale), shift_addr = smem_u32(s_shift);
        const uint32_t row_off = (uint32_t)row * RB, row_x = RB == 128 ? (uint32_t)(row & 7) : ((uint32_t)(row >> 1) & 3u);
        int i = g;                                         // CTA-local tile counter
        for (int tile = vb + g * vgrid; tile < num_tiles; tile += G * vgrid, i += G) {
            const int acc = i % args.acc_stages;
            const uint32_t acc_phase = (uint32_t)(i / args.acc_stages) & 1u;
            const int cb = i % NBUF;
            const uint32_t cphase = (uint32_t)(i / NBUF) & 1u;
            const uint32_t slot = smem_u32(sC + (size_t)cb * SLOT);
            if (has_res) MBAR_WAIT_HERE(&cfull[cb], cphase);          // residual landed (the producer waited for the slot)
            else MBAR_WAIT_HERE(&cempty[cb], cphase ^ 1);             // slot free
            MBAR_WAIT_HERE(&tfull[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * N);
#pragma unroll
            for (int c0 = 0; c0 < N; c0 += SUBC) {
                uint32_t r[SUBC];
                tmem_ld32(taddr + c0, r);
                if constexpr (SUBC == 64) tmem_ld32(taddr + c0 + 32, r + 32);
                tmem_ld_wait();
                if (c0 + SUBC == N) {                      // accumulator fully read: hand it back before the math
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty[acc]);
                }
#pragma unroll
                for (int j = 0; j < SUBC; j += 8) {
                    const uint32_t addr = slot + (uint32_t)(c0 / SUBC) * SUBT + row_off + ((((uint32_t)j >> 3) ^ row_x) << 4);
                    float v[8];
#pragma unroll
                    for (int q = 0; q < 8; q += 4) {
                        const uint4 s4 = lds128(scale_addr + (uint32_t)(c0 + j + q) * 4u);
                        const uint4 h4 = lds128(shift_addr + (uint32_t)(c0 + j + q) * 4u);
                        v[q + 0] = fmaf(__uint_as_float(r[j + q + 0]), __uint_as_float(s4.x), __uint_as_float(h4.x));
                        v[q + 1] = fmaf(__uint_as_float(r[j + q + 1]), __uint_as_float(s4.y), __uint_as_float(h4.y));
                        v[q + 2] = fmaf(__uint_as_float(r[j + q + 2]), __uint_as_float(s4.z), __uint_as_float(h4.z));
                        v[q + 3] = fmaf(__uint_as_float(r[j + q + 3]), __uint_as_float(s4.w), __uint_as_float(h4.w));
                    }
                    if (leaky) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) v[q] = v[q] > 0.f ? v[q] : 0.1f * v[q];
                    }
                    if (has_res) {
                        const uint4 rr = lds128(addr);
                        const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&rr);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float2 f = __bfloat1622float2(h[q]);
                            v[2 * q] = fmaf(alpha, v[2 * q], beta * f.x);
                            v[2 * q + 1] = fmaf(alpha, v[2 * q + 1], beta * f.y);
                        }
                    }
                    uint4 o;
                    __nv_bfloat162 *oh = reinterpret_cast<__nv_bfloat162 *>(&o);
#pragma unroll
                    for (int q = 0; q < 4; ++q) oh[q] = __floats2bfloat162_rn(v[2 * q], v[2 * q + 1]);
                    sts128(addr, o);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");            // generic-proxy writes -> visible to the TMA engine
            __syncwarp();
            if (lane == 0) mbar_arrive(&cwritten[cb]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(args.tmem_cols) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------
// Fused residual block for the widest feature maps:  out = x + leaky(BN2(conv3x3(leaky(BN1(conv1x1(x))))))   (64 -> 32 -> 64)
// YOLOv3 layers 2-4 at 208 x 208 are pure HBM traffic: the 1x1 reads x and writes y (32 channels), the 3x3 reads y, reads x
// again as the residual and writes out.  Here y never leaves the SM: per tile the x halo patch arrives by TMA, a first pair
// of MMAs (M = 2 x 128 patch pixels, N = 32, K = 64) computes y for the WHOLE patch into TMEM, the "middle" epilogue applies
// BN1 + leaky, zeroes the pixels outside the image (the 3x3's padding is zero in y, not leaky(BN1(0))) and writes y as the
// 64-byte-row swizzled patch the patch kernel would have loaded; the 3x3 then runs exactly as in conv_tc_patch_kernel
// (18 MMAs on row-shifted descriptors), and the last epilogue adds the residual — the interior of the x patch, still in
// shared memory — and feeds the output ring.  MEASURED: 0.245 ms against 0.127 + 0.165 ms for the two separate kernels.  DRAM
// traffic is x once + out once (ncu: 354 MB read, 312 MB written), but the SM-side work of both layers now shares one SM's
// shared-memory bandwidth (~240 KB of operand/staging traffic per 120-pixel tile), which is what bounds it: neither more x
// stages, L2 prefetch of the patches, a second middle-epilogue group nor dropping the residual's TMA fetch moved it.
// Roles (480 threads): warp 0 TMA producer (x patches, weights), warp 1 MMA issuer (software-pipelined:
// MMA1 of tile i+1 is issued before MMA2 of tile i), warp 2 store warp, warps 3-10 two middle-epilogue groups, warps 11-14
// final epilogue.  Traffic per block: x once (+ halo, mostly L2) and out once, instead of 2 x + 2 y + out.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(480, 1)
conv_tc_block_kernel(const __grid_constant__ ConvTcMaps maps, const ConvTcArgs args)
{
    constexpr int YS_BYTES = 256 * 64, W2_TILE = 64 * 64, W1_BYTES = 32 * 128;
    const int XS_BYTES = args.a_stage_bytes;                         // patch pixels * 128, 1024-aligned; the first GEMM reads 256 rows,
                                                                     // so the ring is followed by padding up to a full 32 KB
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int XS = args.stages;                                      // x patch stages
    uint8_t *sX = smem;
    uint8_t *sY = sX + (size_t)(XS - 1) * XS_BYTES + 256 * 128;      // 2 y patches
    uint8_t *sW2 = sY + 2 * YS_BYTES;                                // 9 tap tiles [64 filters x 32 ch], 64B-swizzled
    uint8_t *sW1 = sW2 + 9 * W2_TILE;                                // [32 filters x 64 ch], 128B-swizzled
    uint8_t *sC = sW1 + W1_BYTES;                                    // output ring
    uint8_t *aux = sC + (size_t)args.c_bufs * 16384;
    uint64_t *xfull = (uint64_t *)aux;        // [4]
    uint64_t *xempty = xfull + 4;             // [4]
    uint64_t *a1full = xempty + 4;            // [2]
    uint64_t *a1empty = a1full + 2;           // [2]
    uint64_t *yfull = a1empty + 2;            // [2]
    uint64_t *yempty = yfull + 2;             // [2]
    uint64_t *a2full = yempty + 2;            // [4]
    uint64_t *a2empty = a2full + 4;           // [4]
    uint64_t *cfull = a2empty + 4;            // [4]
    uint64_t *cempty = cfull + 4;             // [4]
    uint64_t *cwritten = cempty + 4;          // [4]
    uint64_t *wfull = cwritten + 4;
    uint32_t *tmem_slot = (uint32_t *)(wfull + 1);
    float *s_sc2 = (float *)(aux + 512), *s_sh2 = s_sc2 + 64, *s_sc1 = s_sh2 + 64, *s_sh1 = s_sc1 + 32;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_tiles = args.m_tiles, NBUF = args.c_bufs;
    const int P = args.halo_P, TWv = args.halo_TWv, THv = args.halo_THv;
    const int PR = (THv + 2) * P;                                    // patch pixels (<= 256)
    const int nhalf = PR > 128 ? 2 : 1;
    const int my_first = blockIdx.x, step = gridDim.x;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) {
            mbar_init(&xfull[i], 1); mbar_init(&xempty[i], 4); mbar_init(&a2full[i], 1); mbar_init(&a2empty[i], 4);
            mbar_init(&cfull[i], 1); mbar_init(&cempty[i], 1); mbar_init(&cwritten[i], 4);
        }
        for (int i = 0; i < 2; ++i) { mbar_init(&a1full[i], 1); mbar_init(&a1empty[i], 4); mbar_init(&yfull[i], 4); mbar_init(&yempty[i], 1); }
        mbar_init(wfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (threadIdx.x < 64) { s_sc2[threadIdx.x] = args.scale[threadIdx.x]; s_sh2[threadIdx.x] = args.shift[threadIdx.x]; }
    if (threadIdx.x < 32) { s_sc1[threadIdx.x] = args.scale1[threadIdx.x]; s_sh1[threadIdx.x] = args.shift1[threadIdx.x]; }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();
    // TMEM columns: y accumulators [t][half] at t*64 + half*32 (t = 0,1); output accumulators at 128 + v*64 (v = 0..3)

    if (warp == 0) {
        // ===================================== TMA producer =====================================
        if (lane == 0 && my_first < num_tiles) {
            mbar_expect_tx(wfull, (uint32_t)(9 * W2_TILE + W1_BYTES));
            for (int i = 0; i < 9; ++i) tma_load_2d(&maps.b, sW2 + (size_t)i * W2_TILE, wfull, i * 32, 0);
            tma_load_2d(&maps.a[1], sW1, wfull, 0, 0);
            pdl_wait();
            TileWalk t; t.init(my_first, step, args.tiles_x, args.tiles_y);
            // the x stages are 25 KB and live until the final epilogue has read the residual, so only 4 fit: too few to
            // cover the DRAM latency by themselves.  The patches of the tiles further ahead are pulled into L2 instead.
            const int ahead = args.b_stages;                       // prefetch distance in tiles (0 = off)
            TileWalk tp; tp.init(my_first, step, args.tiles_x, args.tiles_y);
            int pf = 0;
            for (; pf < ahead && my_first + pf * step < num_tiles; ++pf, tp.next(args.tiles_x, args.tiles_y))
                tma_prefetch_4d(&maps.a[0], 0, tp.tx * TWv - 1, tp.ty * THv - 1, tp.tn);
            int i = 0;
            for (int tile = my_first; tile < num_tiles; tile += step, ++i, t.next(args.tiles_x, args.tiles_y)) {
                const int xs = i % XS;
                if (ahead > 0 && my_first + pf * step < num_tiles) {
                    tma_prefetch_4d(&maps.a[0], 0, tp.tx * TWv - 1, tp.ty * THv - 1, tp.tn);
                    ++pf; tp.next(args.tiles_x, args.tiles_y);
                }
                MBAR_WAIT_HERE(&xempty[xs], ((i / XS) & 1) ^ 1);
                mbar_expect_tx(&xfull[xs], (uint32_t)(PR * 128));
                tma_load_4d(&maps.a[0], sX + (size_t)xs * XS_BYTES, &xfull[xs], 0, t.tx * TWv - 1, t.ty * THv - 1, t.tn);
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer (whole warp, elected lane) ============
        if (my_first < num_tiles) {
            const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(32 >> 3) << 17) | ((128u >> 4) << 24);
            const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((128u >> 4) << 24);
            uint32_t a_off[9]; uint64_t b2desc[9];
#pragma unroll
            for (int sgm = 0; sgm < 9; ++sgm) {
                a_off[sgm] = (uint32_t)(((sgm / 3) * P + (sgm % 3)) * 64) >> 4;
                b2desc[sgm] = make_desc<32>(smem_u32(sW2) + (uint32_t)sgm * W2_TILE);
            }
            const uint64_t x0desc = make_desc<64>(smem_u32(sX)), y0desc = make_desc<32>(smem_u32(sY)), w1desc = make_desc<64>(smem_u32(sW1));
            int n_my = 0;
            for (int tile = my_first; tile < num_tiles; tile += step) ++n_my;
            MBAR_WAIT_HERE(wfull, 0);
            for (int it = 0; it <= n_my; ++it) {
                if (it < n_my) {                           // first GEMM of tile `it`: y = x_patch * W1^T
                    const int xs = it % XS, t = it & 1;
                    MBAR_WAIT_HERE(&a1empty[t], ((it >> 1) & 1) ^ 1);
                    MBAR_WAIT_HERE(&xfull[xs], (it / XS) & 1);
                    tc_fence_after();
                    const uint64_t xdesc = x0desc + (uint64_t)((uint32_t)xs * ((uint32_t)XS_BYTES >> 4));
                    for (int half = 0; half < nhalf; ++half) {
                        const uint32_t d1 = tmem_base + (uint32_t)(t * 64 + half * 32);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            tc_mma_bf16_elect(d1, xdesc + (uint64_t)(half * (128 * 128 >> 4) + 2 * k), w1desc + (uint64_t)(2 * k), idesc1, k != 0 ? 1u : 0u);
                    }
                    tc_commit_elect(&a1full[t]);       // the x stage also holds the residual: the final epilogue releases it
                }
                if (it >= 1) {                             // second GEMM of tile `it - 1`: the 3x3 over the y patch
                    const int j = it - 1, u = j & 1, v = j & 3;
                    MBAR_WAIT_HERE(&a2empty[v], ((j >> 2) & 1) ^ 1);
                    MBAR_WAIT_HERE(&yfull[u], (j >> 1) & 1);
                    tc_fence_after();
                    const uint32_t d2 = tmem_base + (uint32_t)(128 + v * 64);
                    const uint64_t ydesc = y0desc + (uint64_t)((uint32_t)u * (YS_BYTES >> 4));
#pragma unroll
                    for (int sgm = 0; sgm < 9; ++sgm) {
#pragma unroll
                        for (int k = 0; k < 2; ++k)
                            tc_mma_bf16_elect(d2, ydesc + (uint64_t)(a_off[sgm] + 2 * k), b2desc[sgm] + (uint64_t)(2 * k), idesc2, (sgm | k) != 0 ? 1u : 0u);
                    }
                    tc_commit_elect(&yempty[u]);
                    tc_commit_elect(&a2full[v]);
                }
            }
        }
    } else if (warp == 2) {
        // ===================================== store warp =======================================
        if (lane == 0) {
            pdl_wait();
            TileWalk t; t.init(my_first, step, args.tiles_x, args.tiles_y);
            int i = 0;
            for (int tile = my_first; tile < num_tiles; tile += step, ++i, t.next(args.tiles_x, args.tiles_y)) {
                const int ox0 = t.tx * TWv, oy0 = t.ty * THv, cb = i % NBUF;
                int rows_here = args.OH - oy0; if (rows_here > THv) rows_here = THv;
                const uint8_t *src = sC + (size_t)cb * 16384;
                MBAR_WAIT_HERE(&cwritten[cb], (i / NBUF) & 1);
                for (int yy = 0; yy < rows_here; ++yy)
                    tma_store_4d(&maps.c, src + yy * P * 128, 0, ox0, oy0 + yy, t.tn);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                bulk_wait_read<0>();
                mbar_arrive(&cempty[cb]);
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    } else if (warp <= 10) {
        // ===================================== middle epilogue: y accumulators -> y patch =======
        // two groups of four warps take alternate tiles (group g owns y accumulator / y patch buffer g): this stage sits
        // between the two GEMMs of a tile, so its latency is the block's critical path
        const int g1 = (warp - 3) >> 2;
        const int quarter = warp & 3;
        const bool leaky1 = args.act1 == ACT_LEAKY;
        const uint32_t sc1 = smem_u32(s_sc1), sh1 = smem_u32(s_sh1);
        TileWalk t; t.init(my_first + g1 * step, 2 * step, args.tiles_x, args.tiles_y);
        int i = g1;
        for (int tile = my_first + g1 * step; tile < num_tiles; tile += 2 * step, i += 2, t.next(args.tiles_x, args.tiles_y)) {
            const int tb = i & 1;
            const uint32_t ph = (uint32_t)(i >> 1) & 1u;
            MBAR_WAIT_HERE(&a1full[tb], ph);
            tc_fence_after();
            uint32_t r0[32], r1[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(tb * 64);
            tmem_ld32(taddr, r0);
            if (nhalf == 2) tmem_ld32(taddr + 32, r1);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a1empty[tb]);
            MBAR_WAIT_HERE(&yempty[tb], ph ^ 1u);          // the 3x3 that last read this y patch has finished
            const uint32_t ybase = smem_u32(sY) + (uint32_t)tb * YS_BYTES;
            const int y0 = t.ty * THv - 1, x0 = t.tx * TWv - 1;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                if (half < nhalf) {
                    const uint32_t *r = half ? r1 : r0;
                    const int pp = half * 128 + quarter * 32 + lane;
                    if (pp < PR) {
                        const int pr = pp / P, pc = pp - pr * P;
                        const bool inside = (unsigned)(y0 + pr) < (unsigned)args.OH && (unsigned)(x0 + pc) < (unsigned)args.OW;
                        const uint32_t row_addr = ybase + (uint32_t)pp * 64u;
                        const uint32_t sw = (uint32_t)(pp >> 1) & 3u;
#pragma unroll
                        for (int c = 0; c < 32; c += 8) {
                            float v[8];
#pragma unroll
                            for (int q = 0; q < 8; q += 4) {
                                const uint4 s4 = lds128(sc1 + (uint32_t)(c + q) * 4u), h4 = lds128(sh1 + (uint32_t)(c + q) * 4u);
                                v[q + 0] = fmaf(__uint_as_float(r[c + q + 0]), __uint_as_float(s4.x), __uint_as_float(h4.x));
                                v[q + 1] = fmaf(__uint_as_float(r[c + q + 1]), __uint_as_float(s4.y), __uint_as_float(h4.y));
                                v[q + 2] = fmaf(__uint_as_float(r[c + q + 2]), __uint_as_float(s4.z), __uint_as_float(h4.z));
                                v[q + 3] = fmaf(__uint_as_float(r[c + q + 3]), __uint_as_float(s4.w), __uint_as_float(h4.w));
                            }
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                if (leaky1) v[q] = v[q] > 0.f ? v[q] : 0.1f * v[q];
                                if (!inside) v[q] = 0.f;
                            }
                            uint4 o;
                            __nv_bfloat162 *oh = reinterpret_cast<__nv_bfloat162 *>(&o);
#pragma unroll
                            for (int q = 0; q < 4; ++q) oh[q] = __floats2bfloat162_rn(v[2 * q], v[2 * q + 1]);
                            sts128(row_addr + ((((uint32_t)c >> 3) ^ sw) << 4), o);
                        }
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&yfull[tb]);
        }
    } else {
        // ===================================== final epilogue (warps 11..14) =====================
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const bool leaky = args.act == ACT_LEAKY;
        const float alpha = args.res_alpha, beta = args.res_beta;
        const uint32_t sc2 = smem_u32(s_sc2), sh2 = smem_u32(s_sh2);
        const uint32_t row_off = (uint32_t)row * 128u, row_x = (uint32_t)(row & 7);
        // the residual of output position (ry, rx) is the patch pixel (ry + 1, rx + 1) of the x stage: no second fetch of x
        const uint32_t res_row = (uint32_t)((row / P + 1) * P + row % P + 1), res_x = res_row & 7u;
        int i = 0;
        for (int tile = my_first; tile < num_tiles; tile += step, ++i) {
            const int v = i & 3, cb = i % NBUF, xs = i % XS;
            const uint32_t slot = smem_u32(sC + (size_t)cb * 16384);
            const uint32_t xres = smem_u32(sX) + (uint32_t)xs * (uint32_t)XS_BYTES + res_row * 128u;     // this row's residual pixel
            MBAR_WAIT_HERE(&cempty[cb], ((i / NBUF) & 1) ^ 1);     // slot free
            MBAR_WAIT_HERE(&a2full[v], (i >> 2) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(128 + v * 64);
            uint32_t r[64];
            tmem_ld32(taddr, r);
            tmem_ld32(taddr + 32, r + 32);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a2empty[v]);
#pragma unroll
            for (int j = 0; j < 64; j += 8) {
                const uint32_t addr = slot + row_off + ((((uint32_t)j >> 3) ^ row_x) << 4);
                float vv[8];
#pragma unroll
                for (int q = 0; q < 8; q += 4) {
                    const uint4 s4 = lds128(sc2 + (uint32_t)(j + q) * 4u), h4 = lds128(sh2 + (uint32_t)(j + q) * 4u);
                    vv[q + 0] = fmaf(__uint_as_float(r[j + q + 0]), __uint_as_float(s4.x), __uint_as_float(h4.x));
                    vv[q + 1] = fmaf(__uint_as_float(r[j + q + 1]), __uint_as_float(s4.y), __uint_as_float(h4.y));
                    vv[q + 2] = fmaf(__uint_as_float(r[j + q + 2]), __uint_as_float(s4.z), __uint_as_float(h4.z));
                    vv[q + 3] = fmaf(__uint_as_float(r[j + q + 3]), __uint_as_float(s4.w), __uint_as_float(h4.w));
                }
                if (leaky) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) vv[q] = vv[q] > 0.f ? vv[q] : 0.1f * vv[q];
                }
                const uint4 rr = lds128(xres + ((((uint32_t)j >> 3) ^ res_x) << 4));
                const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&rr);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float2 f = __bfloat1622float2(h[q]);
                    vv[2 * q] = fmaf(alpha, vv[2 * q], beta * f.x);
                    vv[2 * q + 1] = fmaf(alpha, vv[2 * q + 1], beta * f.y);
                }
                uint4 o;
                __nv_bfloat162 *oh = reinterpret_cast<__nv_bfloat162 *>(&o);
#pragma unroll
                for (int q = 0; q < 4; ++q) oh[q] = __floats2bfloat162_rn(vv[2 * q], vv[2 * q + 1]);
                sts128(addr, o);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) { mbar_arrive(&cwritten[cb]); mbar_arrive(&xempty[xs]); }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

template <int NSUB, int NSEG, int KS0, int KS1> static void launch_patch_variant(ConvTcPlan *p, cudaStream_t s)
{
    static bool configured[64];
    if (first_use_on_this_device(configured)) {
        B200_CHECK(cudaFuncSetAttribute(conv_tc_patch_kernel<NSUB, NSEG, KS0, KS1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    launch_pdl(conv_tc_patch_kernel<NSUB, NSEG, KS0, KS1>, p->grid, 96 + 128 * p->args.ep_groups, p->smem_bytes, s, 1, p->maps, p->args);
}
template <int NSUB> static void launch_patch(ConvTcPlan *p, cudaStream_t s)
{
    const ConvTcArgs &a = p->args;
    if (a.nseg == 6) launch_patch_variant<NSUB, 6, 4, 2>(p, s);                 // stride 2 on pixel-pair rows
    else if (a.seg_k[0] == 1) launch_patch_variant<NSUB, 9, 1, 1>(p, s);        // stride 1, 16 channels
    else if (a.seg_k[0] == 2) launch_patch_variant<NSUB, 9, 2, 2>(p, s);        // stride 1, 32 channels
    else launch_patch_variant<NSUB, 9, 4, 4>(p, s);                             // stride 1, 64 channels
}

void conv_tc_launch_patch(ConvTcPlan *p, cudaStream_t s)
{
    if (p->args.block_n == 32) launch_patch<0>(p, s);
    else if (p->args.block_n == 64) launch_patch<1>(p, s);
    else launch_patch<2>(p, s);
}

void conv_tc_launch_block(ConvTcPlan *p, cudaStream_t s)
{
    static bool configured[64];
    if (first_use_on_this_device(configured))
        B200_CHECK(cudaFuncSetAttribute(conv_tc_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    launch_pdl(conv_tc_block_kernel, p->grid, 480, p->smem_bytes, s, 1, p->maps, p->args);
}
