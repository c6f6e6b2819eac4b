// conv_tc_common.cuh — kernel arguments, epilogue math and the ring-epilogue roles shared by the tcgen05 convolution kernels
// (conv_tc.cu, conv_tc_patch.cu); overview in conv_tc.cu.
#pragma once
#include "kernels.h"
#include <cuda.h>
#include <string>

#include "tc_ptx.cuh"

// ---------------------------------------------------------------------------------------------------
// kernel arguments
// ---------------------------------------------------------------------------------------------------
struct alignas(64) ConvTcMaps {
    CUtensorMap a[4];        // mode 0: a[0] dense [pixels][C] (or the im2col-mode map of the NHWC input); mode 1: a[py*2+px] parity-phase views (stride 1 uses a[0])
    CUtensorMap b;           // weights [cout_pad][K]
    CUtensorMap c;           // output tile store (staged epilogue): same pixel-tile geometry as A, 64-channel boxes
    CUtensorMap r;           // residual tile load (fused shortcut), same geometry
    CUtensorMap b2;          // tail-split tiles (CTA-pair kernel): weight boxes of block_n / split filters
    CUtensorMap cu[3];       // fused 2x upsample: the other three phase views (dy,dx) = (0,1), (1,0), (1,1) of the upsampled tensor
};

struct ConvTcArgs {
    int mode;                // 0 = dense rows (1x1, or any filter size through TMA im2col-mode loads), 1 = spatial tiles
    int exp;                 // TIMING EXPERIMENTS ONLY (env B200_EXP at plan time; results are wrong): bit 0 = the pair kernel re-uses stale weight
                             // stages instead of loading them, bit 1 = stale activation stages (what does operand traffic cost?)
    int im2col;              // mode 0 with size > 1: A tiles = 128 consecutive output pixels gathered per tap by an im2col-mode TMA load
    int batch, OH, OW, cout_pad, ldo;
    int size, stride, pad, cin_blocks;
    int TW, TH, TN, tiles_x, tiles_y;
    int m_tiles, n_tiles, block_n, num_kblocks, stages;
    int a_rows;              // rows the A box really carries (<= 128)
    int b_stage_bytes;       // 1024-aligned
    int tmem_cols;
    int act;
    int resident_b;          // 1: the whole [block_n x K] weight slab stays in shared memory for the CTA's lifetime
    int halo_P, halo_TWv, halo_THv;   // mode 2 (halo patch): row pitch of the patch (TW + size - 1) and the valid tile width / height
    int a_stage_bytes, b_stages;      // mode 2: bytes per patch stage, depth of the separate weight ring
    int acc_stages;          // TMEM accumulator buffers (2..8): small filter tiles let the MMA run many tiles ahead of the epilogue
    int staged;              // 1: epilogue goes TMEM -> registers -> swizzled smem tile -> TMA store (and TMA-loads the residual)
    int pair;                // 1: cta_group::2 kernel (two CTAs share one 256 x block_n accumulator tile and its weights)
    // single-CTA patch kernel (mode 2, pair 0): one K pass per tile over resident weights
    int np;                  // activation patches per tile (1: stride 1; 2: stride 2 on pixel-pair rows)
    int patch_map[4], patch_off[4], patch_dx[4], patch_dy[4];     // tensor map, byte offset in the stage, box origin relative to the tile
    int stage_tx;            // bytes all patch boxes of a tile deliver
    int a_k, b_k;            // elements per smem row of the patches / of the weight tiles (-> swizzle mode)
    int nb, b_koff[12];      // resident weight tiles: K offset of each [block_n x b_k] box
    int nseg, seg_a[12], seg_b[12], seg_k[12];   // K segments: patch byte offset (row shift), weight byte offset, K/16 steps
    int sub_cols, out_f32;   // ring epilogue sub-tile: filters per slot (64 or 32), fp32 output rows
    const float *scale1, *shift1; int act1, block;   // fused residual block (conv_tc_block_kernel): the 1x1's folded BN, block = 1
    int upsample;            // 1: the ring's store warp writes every tile to the four phase views of a 2x upsampled tensor
    // tail splitting (CTA-pair kernel, ring epilogue): the tiles of the last, partly filled wave are cut into `split` filter
    // slices so that every pair works during it.  Virtual tile v < split_from is tile v at full width; the others are slices.
    int split_from, split, vtiles;
    // split-K (CTA-pair kernel, ring epilogue): a launch with few pixel tiles (small batches) cuts every tile's K loop into `ksplit`
    // ranges of `kb_per` k-blocks, one work unit each; unit (tile, s) stores its raw fp32 accumulators into slab s of a workspace
    // ([ksplit][slab_rows][cout_pad], rows = pixels) and splitk_finalize_kernel sums the slabs in a fixed order.
    int ksplit, kb_per, slab_rows;
    int local, ss_stride;    // unshared convolution: weight box and shift row of tile's location (m_tile % locations); floats per shift row
    int ring;                // 1: ring epilogue (ring_roles) with 384 threads; c_bufs = ring depth (<= 4)
    int n_split;             // CTAs per pixel tile, each computing block_n of the cout_pad filters
    int ep_groups, c_bufs;   // epilogue warp groups (1..2) taking alternate tiles; depth of the output/residual tile ring (<= 8)
    long long npix;
    const float *scale, *shift;
    void *out;
    const bf16 *res;         // optional residual (shortcut fused into the epilogue): out = alpha*act(conv) + beta*res
    int ldr;
    float res_alpha, res_beta;
};

static constexpr int kTcThreads = 192;
static constexpr int kTcRingThreads = 384;      // ring epilogue: + store warp, residual loader, second epilogue group

// epilogue math for NC accumulator columns of one pixel row: folded-BN scale/shift, activation, optional
// residual, cast, 16-byte stores.  `sc`/`sh` point at the tile's per-filter constants in shared memory.
template <typename OutT, bool LEAKY, int NC>
__device__ __forceinline__ void emit_columns(const uint32_t *r, const float *sc, const float *sh, OutT *dst, const bf16 *res,
                                             float alpha, float beta, int cols_left)
{
    constexpr int VEC = 16 / (int)sizeof(OutT);
#pragma unroll
    for (int j = 0; j < NC; j += VEC) {
        if (j < cols_left) {
            float v[VEC];
#pragma unroll
            for (int q = 0; q < VEC; q += 4) {
                const float4 s4 = *reinterpret_cast<const float4 *>(sc + j + q);
                const float4 h4 = *reinterpret_cast<const float4 *>(sh + j + q);
                v[q + 0] = fmaf(__uint_as_float(r[j + q + 0]), s4.x, h4.x);
                v[q + 1] = fmaf(__uint_as_float(r[j + q + 1]), s4.y, h4.y);
                v[q + 2] = fmaf(__uint_as_float(r[j + q + 2]), s4.z, h4.z);
                v[q + 3] = fmaf(__uint_as_float(r[j + q + 3]), s4.w, h4.w);
            }
            if (LEAKY) {
#pragma unroll
                for (int q = 0; q < VEC; ++q) v[q] = v[q] > 0.f ? v[q] : 0.1f * v[q];
            }
            if constexpr (sizeof(OutT) == 2) {                 // a residual is only ever fused into a bf16 output
                if (res) {
                    float a[VEC];
                    load_vec<bf16>(res + j, a);
#pragma unroll
                    for (int q = 0; q < VEC; ++q) v[q] = fmaf(alpha, v[q], beta * a[q]);
                }
            }
            store_vec<OutT>(dst + j, v);
        }
    }
}

// staged variant: the thread's pixel row lives in a 128B-swizzled [128 rows x 64 ch] sub-tile per 64 filters (the layout
// TMA expects); 16-byte chunk j of row r sits at chunk j ^ (r & 7), which also makes the per-row accesses of a warp
// bank-conflict-optimal (4 wavefronts per 512-byte request).
template <bool LEAKY, int NC>
__device__ __forceinline__ void emit_staged(const uint32_t *r, const float *sc, const float *sh, uint32_t sC_addr, int row, int c0,
                                            bool has_res, float alpha, float beta)
{
#pragma unroll
    for (int j = 0; j < NC; j += 8) {
        const int c = c0 + j;
        const uint32_t addr = sC_addr + (uint32_t)(c >> 6) * 16384u + (uint32_t)row * 128u + ((uint32_t)(((c & 63) >> 3) ^ (row & 7)) << 4);
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; q += 4) {
            const float4 s4 = *reinterpret_cast<const float4 *>(sc + j + q);
            const float4 h4 = *reinterpret_cast<const float4 *>(sh + j + q);
            v[q + 0] = fmaf(__uint_as_float(r[j + q + 0]), s4.x, h4.x);
            v[q + 1] = fmaf(__uint_as_float(r[j + q + 1]), s4.y, h4.y);
            v[q + 2] = fmaf(__uint_as_float(r[j + q + 2]), s4.z, h4.z);
            v[q + 3] = fmaf(__uint_as_float(r[j + q + 3]), s4.w, h4.w);
        }
        if (LEAKY) {
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = v[q] > 0.f ? v[q] : 0.1f * v[q];
        }
        if (has_res) {
            uint4 rr = lds128(addr);
            const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&rr);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float2 f = __bfloat1622float2(h[q]);
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

// The epilogue loop shared by the 1-CTA and the CTA-pair kernels (warps 2..5 = 128 threads).
template <typename OutT, bool PAIR>
__device__ __forceinline__ void run_epilogue(const ConvTcMaps &maps, const ConvTcArgs &args, uint64_t *tfull, uint64_t *tempty,
                                             uint64_t *rfull, float *s_scale, float *s_shift, uint8_t *sC, uint32_t tmem_base,
                                             int first_tile, int tile_step, int num_tiles, int rank)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int quarter = warp & 3;                          // TMEM lane quarter this warp may touch
    const int row = quarter * 32 + lane;
    const int ep_tid = threadIdx.x - 64;
    int acc = 0; uint32_t acc_phase = 0, rphase = 0;
    int rx = 0, ry = 0, rn = 0;                            // row -> position inside the pixel tile (tile independent)
    if (args.mode == 1) { rx = row % args.TW; ry = (row / args.TW) % args.TH; rn = row / (args.TW * args.TH); }
    if (args.mode == 2) { rx = row % args.halo_P; ry = row / args.halo_P; }      // position inside the patch-pitched tile
    const bool hoist = args.n_tiles == 1;                  // one filter tile: its constants are staged once
    if (hoist) {
        for (int c = ep_tid; c < args.block_n; c += 128) {
            s_scale[c] = c < args.cout_pad ? args.scale[c] : 0.f;
            s_shift[c] = c < args.cout_pad ? args.shift[c] : 0.f;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    pdl_wait();                                            // first residual read / output write comes after this
    const bool leaky = args.act == ACT_LEAKY;
    const bool staged = sizeof(OutT) == 2 && args.staged;
    const bool has_res = args.res != nullptr;
    const int n_sub = args.block_n >> 6;
    const uint32_t sC_addr = smem_u32(sC);
    for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
        const int n_tile = tile % args.n_tiles;
        const int m_tile = PAIR ? 2 * (tile / args.n_tiles) + rank : tile / args.n_tiles;
        const int col0 = n_tile * args.block_n;
        if (!hoist) {
            for (int c = ep_tid; c < args.block_n; c += 128) {
                int co = col0 + c;
                s_scale[acc * 256 + c] = co < args.cout_pad ? args.scale[co] : 0.f;
                s_shift[acc * 256 + c] = co < args.cout_pad ? args.shift[co] : 0.f;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        int tx = 0, ty = 0, tn = 0;
        if (args.mode >= 1) { tx = m_tile % args.tiles_x; ty = (m_tile / args.tiles_x) % args.tiles_y; tn = m_tile / (args.tiles_x * args.tiles_y); }
        // mode 2: image rows of this tile that exist (each is one TMA box of the staged epilogue)
        int rows_here = 0;
        if (args.mode == 2 && m_tile < args.m_tiles) { rows_here = args.OH - ty * args.halo_THv; if (rows_here > args.halo_THv) rows_here = args.halo_THv; }
        OutT *orow = nullptr;
        const bf16 *rrow = nullptr;
        if (staged) {
            if (ep_tid == 0) {
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // the previous tile's stores are done reading sC
                if (has_res) {                                                      // residual tile -> sC while the mainloop runs
                    if (args.mode == 2 && rows_here == 0) {
                        // phantom tile of an odd pair: nothing to load, nothing will be waited for
                    } else if (args.mode == 2) {
                        mbar_expect_tx(rfull, (uint32_t)(rows_here * args.halo_TWv * 128 * n_sub));
                        for (int q = 0; q < n_sub; ++q)
                            for (int yy = 0; yy < rows_here; ++yy)
                                tma_load_4d(&maps.r, sC + q * 16384 + yy * args.halo_P * 128, rfull, col0 + 64 * q, tx * args.halo_TWv,
                                            ty * args.halo_THv + yy, tn);
                    } else {
                        mbar_expect_tx(rfull, (uint32_t)(args.a_rows * args.block_n * 2));
                        for (int q = 0; q < n_sub; ++q) {
                            if (args.mode == 0) tma_load_2d(&maps.r, sC + q * 16384, rfull, col0 + 64 * q, m_tile * 128);
                            else tma_load_4d(&maps.r, sC + q * 16384, rfull, col0 + 64 * q, tx * args.TW, ty * args.TH, tn * args.TN);
                        }
                    }
                }
            }
            if (!has_res) asm volatile("bar.sync 1, 128;" ::: "memory");           // nobody overwrites sC before that wait
        } else {
            long long pix = -1;
            if (m_tile < args.m_tiles) {
                if (args.mode == 0) {
                    long long p = (long long)m_tile * 128 + row;
                    if (p < args.npix) pix = p;
                } else if (args.mode == 1) {
                    int ox = tx * args.TW + rx, oy = ty * args.TH + ry, n = tn * args.TN + rn;
                    if (row < args.a_rows && ox < args.OW && oy < args.OH && n < args.batch) pix = ((long long)n * args.OH + oy) * args.OW + ox;
                } else {
                    int ox = tx * args.halo_TWv + rx, oy = ty * args.halo_THv + ry;
                    if (rx < args.halo_TWv && ry < args.halo_THv && ox < args.OW && oy < args.OH && tn < args.batch)
                        pix = ((long long)tn * args.OH + oy) * args.OW + ox;
                }
            }
            orow = pix >= 0 ? (OutT *)args.out + pix * args.ldo + col0 : nullptr;
            rrow = (has_res && pix >= 0) ? args.res + pix * args.ldr + col0 : nullptr;
        }
        const int cols_valid = args.cout_pad - col0;        // columns of this tile that exist in the output row

        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_after();
        if (staged && has_res && !(args.mode == 2 && rows_here == 0)) { mbar_wait(rfull, rphase); rphase ^= 1; }
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * args.block_n);
        const float *sc = s_scale + (hoist ? 0 : acc * 256), *sh = s_shift + (hoist ? 0 : acc * 256);
        int c0 = 0;
        for (; c0 + 32 <= args.block_n; c0 += 32) {
            uint32_t r[32];
            tmem_ld32(taddr + c0, r);
            tmem_ld_wait();
            if (staged) {
                if (leaky) emit_staged<true, 32>(r, sc + c0, sh + c0, sC_addr, row, c0, has_res, args.res_alpha, args.res_beta);
                else emit_staged<false, 32>(r, sc + c0, sh + c0, sC_addr, row, c0, has_res, args.res_alpha, args.res_beta);
            } else if (orow) {
                if (leaky) emit_columns<OutT, true, 32>(r, sc + c0, sh + c0, orow + c0, rrow ? rrow + c0 : nullptr, args.res_alpha, args.res_beta, cols_valid - c0);
                else emit_columns<OutT, false, 32>(r, sc + c0, sh + c0, orow + c0, rrow ? rrow + c0 : nullptr, args.res_alpha, args.res_beta, cols_valid - c0);
            }
        }
        if (c0 < args.block_n) {                            // 16-column tail (block_n is a multiple of 16; never staged)
            uint32_t r[16];
            tmem_ld16(taddr + c0, r);
            tmem_ld_wait();
            if (orow) {
                if (leaky) emit_columns<OutT, true, 16>(r, sc + c0, sh + c0, orow + c0, rrow ? rrow + c0 : nullptr, args.res_alpha, args.res_beta, cols_valid - c0);
                else emit_columns<OutT, false, 16>(r, sc + c0, sh + c0, orow + c0, rrow ? rrow + c0 : nullptr, args.res_alpha, args.res_beta, cols_valid - c0);
            }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (PAIR) mbar_arrive_leader(&tempty[acc]); else mbar_arrive(&tempty[acc]); }
        if (staged) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");            // generic-proxy writes -> visible to the TMA engine
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (ep_tid == 0) {
                for (int q = 0; q < n_sub; ++q) {
                    if (args.mode == 0) tma_store_2d(&maps.c, sC + q * 16384, col0 + 64 * q, m_tile * 128);
                    else if (args.mode == 1) tma_store_4d(&maps.c, sC + q * 16384, col0 + 64 * q, tx * args.TW, ty * args.TH, tn * args.TN);
                    else
                        for (int yy = 0; yy < rows_here; ++yy)        // one box per image row: the tile is patch-pitched in smem
                            tma_store_4d(&maps.c, sC + q * 16384 + yy * args.halo_P * 128, col0 + 64 * q, tx * args.halo_TWv,
                                         ty * args.halo_THv + yy, tn);
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
        if (++acc == args.acc_stages) { acc = 0; acc_phase ^= 1; }
    }
    if (staged && ep_tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// mbar_wait as a macro: the spin shows up at the CALL SITE's line in profiler source views (which wait is the hot one)
#define MBAR_WAIT_HERE(bar, parity)                                                                                      \
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"        \
                 "@p bra WAIT_DONE;\n\tbra WAIT_LOOP;\n\tWAIT_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"((uint32_t)(parity)) : "memory")

template <int PENDING> __device__ __forceinline__ void bulk_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(PENDING) : "memory");
}

// ---------------------------------------------------------------------------------------------------
// Ring epilogue (args.ring): the tile leaves TMEM in 64-filter sub-tiles through a ring of swizzled 16 KB slots.
//   warp 2      store warp: slot written -> TMA store -> slot free once the store engine has read it
//   warp 3      residual loader (fused shortcut): slot free -> TMA load of the residual sub-tile, tiles ahead of the math
//   warps 4-11  two epilogue groups (even / odd sub-tiles): TMEM -> scale/shift/leaky (+ residual, in place) -> slot
// Nothing in a tile's epilogue waits on a DRAM round trip or on another role's bookkeeping; the roles only meet at
// mbarriers.  With the serial epilogue (run_epilogue) a fused shortcut cost the 52x52 layers 18 % (990 vs 1200 TFLOP/s).
// ---------------------------------------------------------------------------------------------------
struct VTile { int tile, col_off, width, k0, k1, row_off; };
__device__ __forceinline__ VTile vtile_of(const ConvTcArgs &a, int v)
{
    VTile t;
    t.k0 = 0; t.k1 = a.num_kblocks; t.row_off = 0;
    if (a.ksplit > 1) {
        const int s = v % a.ksplit;
        t.tile = v / a.ksplit; t.col_off = 0; t.width = a.block_n;
        t.k0 = s * a.kb_per; t.k1 = t.k0 + a.kb_per < a.num_kblocks ? t.k0 + a.kb_per : a.num_kblocks;
        t.row_off = s * a.slab_rows;
        return t;
    }
    if (v < a.split_from) { t.tile = v; t.col_off = 0; t.width = a.block_n; return t; }
    const int w = v - a.split_from;
    t.width = a.block_n / a.split;
    t.tile = a.split_from + w / a.split;
    t.col_off = (w % a.split) * t.width;
    return t;
}

struct RingTile {
    int n_tile, m_tile, col0, tx, ty, tn, rows_here, nsub, row_off;
    bool real;
};
template <bool PAIR>
__device__ __forceinline__ RingTile ring_tile(const ConvTcArgs &args, int vt, int rank)
{
    RingTile t;
    const VTile v = vtile_of(args, vt);
    const int tile = v.tile;
    t.n_tile = tile % args.n_tiles;
    t.m_tile = PAIR ? 2 * (tile / args.n_tiles) + rank : tile / args.n_tiles;
    t.col0 = t.n_tile * args.block_n + v.col_off;
    t.nsub = v.width / args.sub_cols;
    t.row_off = v.row_off;
    t.tx = t.ty = t.tn = 0; t.rows_here = 0;
    t.real = t.m_tile < args.m_tiles;                       // an odd tile count leaves the pair's second CTA a phantom tile
    if (args.mode >= 1) { t.tx = t.m_tile % args.tiles_x; t.ty = (t.m_tile / args.tiles_x) % args.tiles_y; t.tn = t.m_tile / (args.tiles_x * args.tiles_y); }
    if (args.mode == 2) { t.rows_here = args.OH - t.ty * args.halo_THv; if (t.rows_here > args.halo_THv) t.rows_here = args.halo_THv; }
    return t;
}

// hand a TMEM accumulator back to the MMA issuer (the pair's barrier lives in the leader CTA)
template <bool PAIR> __device__ __forceinline__ void ring_release(uint64_t *tempty, int acc, int lane)
{
    tc_fence_before();
    __syncwarp();
    if (lane == 0) { if (PAIR) mbar_arrive_leader(&tempty[acc]); else mbar_arrive(&tempty[acc]); }
}

// one pixel row of a sub-tile: SUBC accumulator columns -> scale/shift/leaky (+ residual, read from the slot) -> the slot.
// Slot rows are SUBC * esz bytes (128 or 64) = the swizzle span: 16-byte chunk k of row r sits at chunk k ^ f(r).
template <int SUBC, bool F32>
__device__ __forceinline__ void ring_emit(const uint32_t *r, const float *gsc, const float *gsh, uint32_t slot_addr, int row,
                                          bool leaky, bool has_res, float alpha, float beta)
{
    constexpr int RB = SUBC * (F32 ? 4 : 2);
    const uint32_t row_addr = slot_addr + (uint32_t)row * RB;
    const uint32_t swz = RB == 128 ? (uint32_t)(row & 7) : ((uint32_t)(row >> 1) & 3u);
#pragma unroll
    for (int c = 0; c < SUBC; c += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; u += 4) {
            const float4 s4 = __ldg(reinterpret_cast<const float4 *>(gsc + c + u));
            const float4 h4 = __ldg(reinterpret_cast<const float4 *>(gsh + c + u));
            v[u + 0] = fmaf(__uint_as_float(r[c + u + 0]), s4.x, h4.x);
            v[u + 1] = fmaf(__uint_as_float(r[c + u + 1]), s4.y, h4.y);
            v[u + 2] = fmaf(__uint_as_float(r[c + u + 2]), s4.z, h4.z);
            v[u + 3] = fmaf(__uint_as_float(r[c + u + 3]), s4.w, h4.w);
        }
        if (leaky) {
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = v[u] > 0.f ? v[u] : 0.1f * v[u];
        }
        if constexpr (F32) {
            sts128(row_addr + ((((uint32_t)c >> 2) ^ swz) << 4), make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3])));
            sts128(row_addr + (((((uint32_t)c >> 2) + 1) ^ swz) << 4), make_uint4(__float_as_uint(v[4]), __float_as_uint(v[5]), __float_as_uint(v[6]), __float_as_uint(v[7])));
        } else {
            const uint32_t addr = row_addr + ((((uint32_t)c >> 3) ^ swz) << 4);
            if (has_res) {
                const uint4 rr = lds128(addr);
                const __nv_bfloat162 *hh = reinterpret_cast<const __nv_bfloat162 *>(&rr);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float2 f = __bfloat1622float2(hh[u]);
                    v[2 * u] = fmaf(alpha, v[2 * u], beta * f.x);
                    v[2 * u + 1] = fmaf(alpha, v[2 * u + 1], beta * f.y);
                }
            }
            uint4 o;
            __nv_bfloat162 *oh = reinterpret_cast<__nv_bfloat162 *>(&o);
#pragma unroll
            for (int u = 0; u < 4; ++u) oh[u] = __floats2bfloat162_rn(v[2 * u], v[2 * u + 1]);
            sts128(addr, o);
        }
    }
}

template <bool PAIR>
__device__ __forceinline__ void ring_roles(const ConvTcMaps &maps, const ConvTcArgs &args, uint64_t *tfull, uint64_t *tempty,
                                           uint64_t *ring_bars, uint8_t *sC, uint32_t tmem_base,
                                           int first_tile, int tile_step, int num_tiles, int rank)
{
    uint64_t *cfull = ring_bars, *cempty = ring_bars + 4, *cwritten = ring_bars + 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int SUBC = args.sub_cols;                        // filters per sub-tile: 64 (bf16) or 32 (bf16 / fp32 outputs)
    const int NBUF = args.c_bufs;
    const int row_bytes = SUBC * (args.out_f32 ? 4 : 2);   // 128 or 64: also the swizzle span of the slot
    const bool has_res = args.res != nullptr;
    if (warp == 2) {
        // ===================================== store warp =======================================
        if (lane == 0) {
            pdl_wait();
            int j = 0;
            for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
                const RingTile t = ring_tile<PAIR>(args, tile, rank);
                if (!t.real) continue;
                for (int q = 0; q < t.nsub; ++q, ++j) {
                    const int slot = j % NBUF;
                    const uint8_t *src = sC + (size_t)slot * 16384;
                    MBAR_WAIT_HERE(&cwritten[slot], (j / NBUF) & 1);
                    if (args.mode == 0) tma_store_2d(&maps.c, src, t.col0 + SUBC * q, t.m_tile * 128 + t.row_off);
                    else if (args.mode == 1) {
                        tma_store_4d(&maps.c, src, t.col0 + SUBC * q, t.tx * args.TW, t.ty * args.TH, t.tn * args.TN);
                        if (args.upsample)                 // upsample_layer.c:72-96 (nearest, stride 2) fused: same tile, three more phases
                            for (int ph = 0; ph < 3; ++ph)
                                tma_store_4d(&maps.cu[ph], src, t.col0 + SUBC * q, t.tx * args.TW, t.ty * args.TH, t.tn * args.TN);
                    }
                    else
                        for (int yy = 0; yy < t.rows_here; ++yy)
                            tma_store_4d(&maps.c, src + yy * args.halo_P * row_bytes, t.col0 + SUBC * q, t.tx * args.halo_TWv, t.ty * args.halo_THv + yy, t.tn);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    bulk_wait_read<0>();
                    mbar_arrive(&cempty[slot]);
                }
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    } else if (warp == 3) {
        // ===================================== residual loader ==================================
        if (lane == 0 && has_res) {
            pdl_wait();
            int j = 0;
            for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
                const RingTile t = ring_tile<PAIR>(args, tile, rank);
                if (!t.real) continue;
                for (int q = 0; q < t.nsub; ++q, ++j) {
                    const int slot = j % NBUF;
                    uint8_t *dst = sC + (size_t)slot * 16384;
                    MBAR_WAIT_HERE(&cempty[slot], ((j / NBUF) & 1) ^ 1);
                    if (args.mode == 2) {
                        mbar_expect_tx(&cfull[slot], (uint32_t)(t.rows_here * args.halo_TWv * 128));
                        for (int yy = 0; yy < t.rows_here; ++yy)
                            tma_load_4d(&maps.r, dst + yy * args.halo_P * row_bytes, &cfull[slot], t.col0 + SUBC * q, t.tx * args.halo_TWv, t.ty * args.halo_THv + yy, t.tn);
                    } else {
                        mbar_expect_tx(&cfull[slot], (uint32_t)(args.a_rows * 128));
                        if (args.mode == 0) tma_load_2d(&maps.r, dst, &cfull[slot], t.col0 + SUBC * q, t.m_tile * 128);
                        else tma_load_4d(&maps.r, dst, &cfull[slot], t.col0 + SUBC * q, t.tx * args.TW, t.ty * args.TH, t.tn * args.TN);
                    }
                }
            }
        }
    } else {
        // ===================================== epilogue groups ==================================
        const int h = (warp - 4) >> 2;                     // group: sub-tiles q = h, h + 2, ...
        const int quarter = warp & 3;                      // TMEM lane quarter this warp may touch
        const int row = quarter * 32 + lane;
        const bool leaky = args.act == ACT_LEAKY;
        const float alpha = args.res_alpha, beta = args.res_beta;
        const uint32_t sC_addr = smem_u32(sC);
        int acc = 0; uint32_t acc_phase = 0;
        int jbase = 0;
        for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
            const RingTile t = ring_tile<PAIR>(args, tile, rank);
            MBAR_WAIT_HERE(&tfull[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * args.block_n);
            const int NSUB = t.nsub;
            if (!t.real || h >= NSUB) {                    // nothing to emit: just hand the accumulator back
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { if (PAIR) mbar_arrive_leader(&tempty[acc]); else mbar_arrive(&tempty[acc]); }
            } else {
                for (int q = h; q < NSUB; q += 2) {
                    const int j = jbase + q, slot = j % NBUF;
                    const uint32_t sphase = (uint32_t)(j / NBUF) & 1u;
                    const uint32_t slot_addr = sC_addr + (uint32_t)slot * 16384u;
                    const bool last = q + 2 >= NSUB;       // this group's last sub-tile: the accumulator can be reused after the load
                    const float *gsc = args.scale + t.col0 + SUBC * q;
                    const float *gsh = args.shift + t.col0 + SUBC * q + (args.local ? (size_t)(t.m_tile % (args.tiles_x * args.tiles_y)) * args.ss_stride : 0);
                    if (SUBC == 64) {
                        uint32_t r[64];
                        tmem_ld32(taddr + 64 * q, r);
                        tmem_ld32(taddr + 64 * q + 32, r + 32);
                        tmem_ld_wait();
                        if (last) ring_release<PAIR>(tempty, acc, lane);
                        if (has_res) MBAR_WAIT_HERE(&cfull[slot], sphase);
                        else MBAR_WAIT_HERE(&cempty[slot], sphase ^ 1u);
                        ring_emit<64, false>(r, gsc, gsh, slot_addr, row, leaky, has_res, alpha, beta);
                    } else {
                        uint32_t r[32];
                        tmem_ld32(taddr + 32 * q, r);
                        tmem_ld_wait();
                        if (last) ring_release<PAIR>(tempty, acc, lane);
                        MBAR_WAIT_HERE(&cempty[slot], sphase ^ 1u);
                        if (args.out_f32) ring_emit<32, true>(r, gsc, gsh, slot_addr, row, leaky, false, alpha, beta);
                        else ring_emit<32, false>(r, gsc, gsh, slot_addr, row, leaky, false, alpha, beta);
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&cwritten[slot]);
                }
            }
            if (t.real) jbase += NSUB;
            if (++acc == args.acc_stages) { acc = 0; acc_phase ^= 1; }
        }
    }
}

