// conv_tc.cu — the fused implicit-GEMM convolution on 5th-generation tensor cores (sm_100a only).
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
// Kernels in this file (one plan per layer, chosen by conv_tc_plan_create / conv_tc_block_plan_create):
//   conv_tc_kernel        1 CTA per tile, weights streamed or resident (small layers, all 1x1 layers with <= 128 filters)
//   conv_tc_pair_kernel   cta_group::2 CTA pair per 256-pixel tile, half of the weight tile per CTA (compute-heavy layers)
//   conv_tc_halo_pair_kernel  opt-in experiment (B200_HALO): patch loads + row-shifted descriptors for the pair kernel
//   conv_tc_patch_kernel  3x3 layers with 32/64 input channels: halo patch per tile, ALL weights resident, HBM-bound
//   conv_tc_block_kernel  fused residual block 1x1 (64->32) + 3x3 (32->64) + shortcut: the intermediate never leaves the SM
// Epilogues: ring_roles (store warp + residual loader + two epilogue groups over a ring of swizzled 64-filter sub-tile
// slots, TMA stores; also writes a fused 2x upsample) for everything that can be staged, run_epilogue (serial, direct
// 16-byte stores) for odd filter counts.
//
// Roofline: tensor pipe.  FLOPs per launch = 2 * pixels * C_out * K (darknet's own BFLOPs formula,
// convolutional_layer.c:325).
#include "kernels.h"
#include <cuda.h>
#include <string>

#include "tc_ptx.cuh"

// ---------------------------------------------------------------------------------------------------
// kernel arguments
// ---------------------------------------------------------------------------------------------------
struct alignas(64) ConvTcMaps {
    CUtensorMap a[4];        // mode 0: a[0] dense [pixels][C]; mode 1: a[py*2+px] parity-phase views (stride 1 uses a[0])
    CUtensorMap b;           // weights [cout_pad][K]
    CUtensorMap c;           // output tile store (staged epilogue): same pixel-tile geometry as A, 64-channel boxes
    CUtensorMap r;           // residual tile load (fused shortcut), same geometry
    CUtensorMap b2;          // tail-split tiles (CTA-pair kernel): weight boxes of block_n / split filters
    CUtensorMap cu[3];       // fused 2x upsample: the other three phase views (dy,dx) = (0,1), (1,0), (1,1) of the upsampled tensor
};

struct ConvTcArgs {
    int mode;                // 0 = dense rows (1x1), 1 = spatial tiles
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
struct VTile { int tile, col_off, width; };
__device__ __forceinline__ VTile vtile_of(const ConvTcArgs &a, int v)
{
    VTile t;
    if (v < a.split_from) { t.tile = v; t.col_off = 0; t.width = a.block_n; return t; }
    const int w = v - a.split_from;
    t.width = a.block_n / a.split;
    t.tile = a.split_from + w / a.split;
    t.col_off = (w % a.split) * t.width;
    return t;
}

struct RingTile {
    int n_tile, m_tile, col0, tx, ty, tn, rows_here, nsub;
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
                    if (args.mode == 0) tma_store_2d(&maps.c, src, t.col0 + SUBC * q, t.m_tile * 128);
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
                }
                for (int kb = 0; kb < args.num_kblocks; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_expect_tx(&full[stage], tx_bytes);
                    const int tap = kb / args.cin_blocks, cb = kb - tap * args.cin_blocks;
                    void *dstA = sA + (size_t)stage * A_BYTES;
                    if (args.mode == 0) {
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
                VTile v; v.tile = vt; v.col_off = 0; v.width = args.block_n;
                if (RING) v = vtile_of(args, vt);
                const int tile = v.tile, half_w = v.width / 2;
                const uint32_t tx_bytes = 2u * (uint32_t)(args.a_rows * BLOCK_K * 2 + half_w * BLOCK_K * 2);
                const CUtensorMap *bmap = v.width == args.block_n ? &maps.b : &maps.b2;
                const int n_tile = tile % args.n_tiles, m_tile = 2 * (tile / args.n_tiles) + (int)rank;
                int ox0 = 0, oy0 = 0, n0 = 0;
                if (args.mode == 1) {
                    int tx = m_tile % args.tiles_x, ty = (m_tile / args.tiles_x) % args.tiles_y, tn = m_tile / (args.tiles_x * args.tiles_y);
                    ox0 = tx * args.TW; oy0 = ty * args.TH; n0 = tn * args.TN;      // a phantom last tile lands past the batch: TMA zero-fills
                }
                for (int kb = 0; kb < args.num_kblocks; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    if (leader) mbar_expect_tx(&full[stage], tx_bytes);
                    const int tap = kb / args.cin_blocks, cb = kb - tap * args.cin_blocks;
                    void *dstA = sA + (size_t)stage * A_BYTES;
                    if (args.mode == 0) {
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
                const int width = RING ? vtile_of(args, vt).width : args.block_n;
                const uint32_t idesc = idesc_base | ((uint32_t)(width >> 3) << 17);
                mbar_wait(&tempty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * args.block_n);
                for (int kb = 0; kb < args.num_kblocks; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint64_t adesc = adesc0 + (uint64_t)((uint32_t)stage * a_step);
                    const uint64_t bdesc = bdesc0 + (uint64_t)((uint32_t)stage * b_step);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / 16; ++k)
                        tc2_mma_bf16_elect(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
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

// ---------------------------------------------------------------------------------------------------
// Halo-patch variant of the CTA-pair kernel for stride-1 size x size convolutions (mode 2).
// The 1-tap-per-TMA-box scheme above re-reads every activation row size^2 times; measured, the TMA engine retires about
// one <=128-byte row per ~3 SM cycles, and that row rate — not the MMA — bounds the kernel.  Here each CTA loads, once per
// 64-channel block, the (TH+size-1) x (TW+size-1)-pixel input patch of its output tile (ONE 4-D TMA box, zero-filled at
// the borders).  Shared-memory swizzling is a pure function of the address (verified on B200 with
// scripts/desc_shift_probe.cu), so filter tap (ky,kx) is simply the SAME patch read through a descriptor whose start is
// shifted by ky*P + kx rows (P = patch pitch).  GEMM row m then stands for patch position (m / P, m % P); positions in
// the size-1 rightmost columns of every patch row are computed but never stored.  Activation rows per 64-channel block drop
// from size^2 * 128 to (TH+size-1) * P (~216 for a 2 x 52 tile); the weight tiles stream through their own ring.
// ---------------------------------------------------------------------------------------------------
template <typename OutT>
__global__ void __launch_bounds__(kTcThreads, 1)
conv_tc_halo_pair_kernel(const __grid_constant__ ConvTcMaps maps, const ConvTcArgs args)
{
    constexpr int BLOCK_K = 64;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int a_stages = args.stages, b_stages = args.b_stages;
    uint8_t *sA = smem;                                              // a_stages patches
    uint8_t *sB = sA + (size_t)a_stages * args.a_stage_bytes;        // b_stages half weight tiles
    uint8_t *sC = sB + (size_t)b_stages * args.b_stage_bytes;
    uint8_t *aux = sC + (args.staged ? (size_t)(args.block_n >> 6) * 16384 : 0);
    uint64_t *afull = (uint64_t *)aux;                               // [<=4]
    uint64_t *aempty = afull + 4;                                    // [<=4]
    uint64_t *bfull = aempty + 4;                                    // [<=8]
    uint64_t *bempty = bfull + 8;                                    // [<=8]
    uint64_t *tfull = bempty + 8;                                    // [2]
    uint64_t *tempty = tfull + 8;                                    // [2]
    uint64_t *rfull = tempty + 9;
    uint32_t *tmem_slot = (uint32_t *)(rfull + 1);
    float *s_scale = (float *)(aux + 512);
    float *s_shift = s_scale + 512;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    const int m_pairs = (args.m_tiles + 1) / 2;
    const int num_tiles = m_pairs * args.n_tiles;
    const int taps = args.size * args.size;

    if (threadIdx.x == 0) {
        for (int i = 0; i < a_stages; ++i) { mbar_init(&afull[i], 1); mbar_init(&aempty[i], 1); }
        for (int i = 0; i < b_stages; ++i) { mbar_init(&bfull[i], 1); mbar_init(&bempty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 8); }
        mbar_init(rfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(args.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();
    const int half_n = args.block_n / 2;
    const int patch_rows = (args.halo_THv + args.size - 1) * args.halo_P;

    if (warp == 0) {
        // ===================================== TMA producer (both CTAs) =========================
        if (lane == 0) {
            int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
            const uint32_t a_tx = 2u * (uint32_t)(patch_rows * BLOCK_K * 2);
            const uint32_t b_tx = 2u * (uint32_t)(half_n * BLOCK_K * 2);
            pdl_wait();
            for (int tile = pair_id; tile < num_tiles; tile += num_pairs) {
                const int n_tile = tile % args.n_tiles, m_tile = 2 * (tile / args.n_tiles) + (int)rank;
                const int tx = m_tile % args.tiles_x, ty = (m_tile / args.tiles_x) % args.tiles_y, tn = m_tile / (args.tiles_x * args.tiles_y);
                for (int cb = 0; cb < args.cin_blocks; ++cb) {
                    mbar_wait(&aempty[sa], pa ^ 1);
                    if (leader) mbar_expect_tx(&afull[sa], a_tx);
                    tma2_load_4d(&maps.a[0], sA + (size_t)sa * args.a_stage_bytes, &afull[sa], cb * BLOCK_K,
                                 tx * args.halo_TWv - args.pad, ty * args.halo_THv - args.pad, tn);
                    if (++sa == a_stages) { sa = 0; pa ^= 1; }
                    for (int tap = 0; tap < taps; ++tap) {
                        mbar_wait(&bempty[sb], pb ^ 1);
                        if (leader) mbar_expect_tx(&bfull[sb], b_tx);
                        tma2_load_2d(&maps.b, sB + (size_t)sb * args.b_stage_bytes, &bfull[sb], (tap * args.cin_blocks + cb) * BLOCK_K,
                                     n_tile * args.block_n + (int)rank * half_n);
                        if (++sb == b_stages) { sb = 0; pb ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer (leader CTA only) =====================
        if (leader && lane == 0) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(args.block_n >> 3) << 17) | ((256u >> 4) << 24);
            int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int tile = pair_id; tile < num_tiles; tile += num_pairs) {
                mbar_wait(&tempty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * args.block_n);
                for (int cb = 0; cb < args.cin_blocks; ++cb) {
                    mbar_wait(&afull[sa], pa);
                    tc_fence_after();
                    const uint32_t patch = smem_u32(sA + (size_t)sa * args.a_stage_bytes);
                    for (int tap = 0; tap < taps; ++tap) {
                        mbar_wait(&bfull[sb], pb);
                        tc_fence_after();
                        const int ky = tap / args.size, kx = tap - ky * args.size;
                        const uint64_t adesc = make_desc<BLOCK_K>(patch + (uint32_t)(ky * args.halo_P + kx) * 128u);   // row-shifted view
                        const uint64_t bdesc = make_desc<BLOCK_K>(smem_u32(sB + (size_t)sb * args.b_stage_bytes));
#pragma unroll
                        for (int k = 0; k < BLOCK_K / 16; ++k)
                            tc2_mma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (cb | tap | k) != 0 ? 1u : 0u);
                        tc2_commit_both(&bempty[sb]);
                        if (++sb == b_stages) { sb = 0; pb ^= 1; }
                    }
                    tc2_commit_both(&aempty[sa]);
                    if (++sa == a_stages) { sa = 0; pa ^= 1; }
                }
                tc2_commit_both(&tfull[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        run_epilogue<OutT, true>(maps, args, tfull, tempty, rfull, s_scale, s_shift, sC, tmem_base, pair_id, num_pairs, num_tiles, (int)rank);
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(args.tmem_cols) : "memory");
    }
}

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

// ---------------------------------------------------------------------------------------------------
// host side: tensor maps + plan
// ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        B200_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (!p || q != cudaDriverEntryPointSuccess) { fprintf(stderr, "b200-darknet: cuTensorMapEncodeTiled unavailable\n"); abort(); }
        fn = (EncodeTiledFn)p;
    }
    return fn;
}

static CUtensorMapSwizzle swizzle_for(int block_k)
{
    return block_k == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (block_k == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

static void encode(CUtensorMap *map, void *base, int rank, const cuuint64_t *dims, const cuuint64_t *strides_bytes,
                   const cuuint32_t *box, int block_k)
{
    cuuint32_t ones[5] = {1, 1, 1, 1, 1};
    CUresult r = encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, base, dims, strides_bytes, box, ones,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(block_k), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        fprintf(stderr, "b200-darknet: cuTensorMapEncodeTiled failed (%d) rank %d dims %llu %llu box %u %u\n", (int)r, rank,
                (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
        abort();
    }
}

// 64-channel x pixel-tile boxes of an NHWC bf16 tensor, 128B-swizzled: the staged epilogue's store / residual-load view
// generic form for the other tcgen05 translation units (conv_stem_tc.cu): dtype 0 = bf16, 1 = fp32; swizzle_bytes 0/32/64/128
void tc_encode_tiled(void *map, int dtype, int rank, void *base, const unsigned long long *dims, const unsigned long long *strides_bytes,
                     const unsigned *box, int swizzle_bytes)
{
    cuuint32_t ones[5] = {1, 1, 1, 1, 1};
    cuuint64_t d[5], st[5]; cuuint32_t b[5];
    for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; if (i < rank - 1) st[i] = strides_bytes[i]; }
    const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = encode_fn()((CUtensorMap *)map, dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, base,
                             d, st, b, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { fprintf(stderr, "b200-darknet: cuTensorMapEncodeTiled failed (%d)\n", (int)r); abort(); }
}

void tc_encode_tiled(void *map, int dtype, int rank, void *base, const unsigned long long *dims, const unsigned long long *strides_bytes,
                     const unsigned *box, int swizzle_bytes);

// output / residual tile view: boxes of `sub_cols` channels (rows of sub_cols * esz bytes = the swizzle span)
static void encode_tile_view(CUtensorMap *map, const TView &t, int channels, const ConvTcArgs &a, int sub_cols = 64)
{
    const unsigned long long esz = dt_size(t.dtype);
    const int dtype = t.dtype == DT_F32 ? 1 : 0;
    const int swz = (int)(sub_cols * esz);
    if (a.mode == 0) {
        unsigned long long dims[2] = {(unsigned long long)channels, (unsigned long long)a.npix};
        unsigned long long strides[1] = {(unsigned long long)t.ld * esz};
        unsigned box[2] = {(unsigned)sub_cols, 128};
        tc_encode_tiled(map, dtype, 2, t.p, dims, strides, box, swz);
    } else {
        unsigned long long dims[4] = {(unsigned long long)channels, (unsigned long long)t.w, (unsigned long long)t.h, (unsigned long long)t.n};
        unsigned long long strides[3] = {(unsigned long long)t.ld * esz, (unsigned long long)t.w * t.ld * esz, (unsigned long long)t.h * t.w * t.ld * esz};
        unsigned box[4] = {(unsigned)sub_cols, (unsigned)a.TW, (unsigned)a.TH, (unsigned)a.TN};
        if (a.mode == 2) { box[1] = (unsigned)a.halo_TWv; box[2] = 1; box[3] = 1; }      // one image row of the tile per box
        tc_encode_tiled(map, dtype, 4, t.p, dims, strides, box, swz);
    }
}

struct ConvTcPlan {
    ConvTcMaps maps;
    ConvTcArgs args;
    int block_k, out_dtype, grid;
    size_t smem_bytes;
    double flops;
    std::string desc;
};

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

template <int BLOCK_K, typename OutT> static void launch_variant(ConvTcPlan *p, cudaStream_t s)
{
    static bool configured = false;
    if (!configured) {
        B200_CHECK(cudaFuncSetAttribute(conv_tc_kernel<BLOCK_K, OutT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        B200_CHECK(cudaFuncSetAttribute(conv_tc_kernel<BLOCK_K, bf16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured = true;
    }
    if (p->args.ring) launch_pdl(conv_tc_kernel<BLOCK_K, bf16, true>, p->grid, kTcRingThreads, p->smem_bytes, s, 1, p->maps, p->args);
    else launch_pdl(conv_tc_kernel<BLOCK_K, OutT, false>, p->grid, kTcThreads, p->smem_bytes, s, 1, p->maps, p->args);
}

template <int BLOCK_K, typename OutT> static void launch_pair_variant(ConvTcPlan *p, cudaStream_t s)
{
    static bool configured = false;
    if (!configured) {
        B200_CHECK(cudaFuncSetAttribute(conv_tc_pair_kernel<BLOCK_K, OutT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        B200_CHECK(cudaFuncSetAttribute(conv_tc_pair_kernel<BLOCK_K, bf16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured = true;
    }
    if (p->args.ring) launch_pdl(conv_tc_pair_kernel<BLOCK_K, bf16, true>, p->grid, kTcRingThreads, p->smem_bytes, s, 2, p->maps, p->args);
    else launch_pdl(conv_tc_pair_kernel<BLOCK_K, OutT, false>, p->grid, kTcThreads, p->smem_bytes, s, 2, p->maps, p->args);
}

template <typename OutT> static void launch_halo_variant(ConvTcPlan *p, cudaStream_t s)
{
    static bool configured = false;
    if (!configured) {
        B200_CHECK(cudaFuncSetAttribute(conv_tc_halo_pair_kernel<OutT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured = true;
    }
    launch_pdl(conv_tc_halo_pair_kernel<OutT>, p->grid, kTcThreads, p->smem_bytes, s, 2, p->maps, p->args);
}

template <int NSUB, int NSEG, int KS0, int KS1> static void launch_patch_variant(ConvTcPlan *p, cudaStream_t s)
{
    static bool configured = false;
    if (!configured) {
        B200_CHECK(cudaFuncSetAttribute(conv_tc_patch_kernel<NSUB, NSEG, KS0, KS1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured = true;
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

void launch_conv_tc(ConvTcPlan *p, cudaStream_t s)
{
    if (p->args.block) {
        static bool configured = false;
        if (!configured) {
            B200_CHECK(cudaFuncSetAttribute(conv_tc_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            configured = true;
        }
        launch_pdl(conv_tc_block_kernel, p->grid, 480, p->smem_bytes, s, 1, p->maps, p->args);
        B200_LAUNCHED();
        return;
    }
    if (p->args.mode == 2 && !p->args.pair) {
        if (p->args.block_n == 32) launch_patch<0>(p, s);
        else if (p->args.block_n == 64) launch_patch<1>(p, s);
        else launch_patch<2>(p, s);
        B200_LAUNCHED();
        return;
    }
    if (p->args.mode == 2) {
        if (p->out_dtype == DT_BF16) launch_halo_variant<bf16>(p, s);
        else launch_halo_variant<float>(p, s);
        B200_LAUNCHED();
        return;
    }
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
        B200_LAUNCHED();
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
    B200_LAUNCHED();
}

void conv_tc_plan_destroy(ConvTcPlan *p) { delete p; }

// shape test shared with the planner (engine.cu decides about shortcut fusion before buffers exist)
bool conv_tc_shape_supported(int cin, int stride, int act)
{
    if (getenv("B200_DISABLE_TC")) return false;
    return cin % 16 == 0 && (stride == 1 || stride == 2) && (act == ACT_LEAKY || act == ACT_LINEAR);
}
const char *conv_tc_plan_desc(ConvTcPlan *p) { return p->desc.c_str(); }

// Fused residual block x -> 1x1 (64 -> 32) -> 3x3 (32 -> 64) -> + x (conv_tc_block_kernel); nullptr when the shapes do not fit
ConvTcPlan *conv_tc_block_plan_create(TView x, TView out, ConvParams p1, ConvParams p2, float res_alpha, float res_beta)
{
    if (getenv("B200_NO_BLOCK_FUSION") || getenv("B200_DISABLE_TC")) return nullptr;
    if (x.dtype != DT_BF16 || out.dtype != DT_BF16 || x.c != 64 || out.c != 64 || x.n != out.n || x.h != out.h || x.w != out.w) return nullptr;
    if (p1.size != 1 || p1.cout_pad != 32 || p2.size != 3 || p2.stride != 1 || p2.pad != 1 || p2.cout_pad != 64) return nullptr;
    if ((p1.act != ACT_LEAKY && p1.act != ACT_LINEAR) || (p2.act != ACT_LEAKY && p2.act != ACT_LINEAR)) return nullptr;
    if (x.ld % 8 != 0 || out.ld % 8 != 0 || ((uintptr_t)x.p & 15) || ((uintptr_t)out.p & 15)) return nullptr;
    ConvTcPlan *p = new ConvTcPlan();
    memset(&p->maps, 0, sizeof p->maps);
    ConvTcArgs &a = p->args;
    memset(&a, 0, sizeof a);
    p->block_k = 64; p->out_dtype = DT_BF16;
    // tile: P = TW + 2 patch columns, TH = 128 / P output rows, (TH + 2) * P <= 256 patch pixels; same measured TMA cost model
    // as the patch kernel (patch rows + residual rows + store rows, ~38 cycles per TMA instruction)
    double best = 1e30; int bTW = 0;
    for (int tw = 4; tw <= out.w && tw + 2 <= 62; ++tw) {
        int P = tw + 2, th = 128 / P; if (th > out.h) th = out.h;
        if ((th + 2) * P > 256) continue;
        double tiles = (double)div_up(out.w, tw) * div_up(out.h, th);
        double cost = tiles * (4.0 * ((th + 2.0) * P + 2.0 * th * tw) + 38.0 * (1 + 2.0 * th));
        if (cost < best) { best = cost; bTW = tw; }
    }
    if (getenv("B200_BLOCK_TW")) { int f = atoi(getenv("B200_BLOCK_TW")); if (f >= 4 && f <= out.w && f + 2 <= 62 && (128 / (f + 2) + 2) * (f + 2) <= 256) bTW = f; }
    if (!bTW) { delete p; return nullptr; }
    const int TWv = bTW, P = TWv + 2;
    int TH = 128 / P; if (TH > out.h) TH = out.h;
    a.mode = 2; a.block = 1; a.pair = 0; a.resident_b = 1; a.staged = 1;
    a.batch = x.n; a.OH = out.h; a.OW = out.w; a.cout_pad = 64; a.ldo = out.ld;
    a.size = 3; a.stride = 1; a.pad = 1; a.block_n = 64; a.n_tiles = 1;
    a.halo_P = P; a.halo_TWv = TWv; a.halo_THv = TH; a.TW = TWv; a.TH = TH; a.TN = 1;
    a.tiles_x = div_up(out.w, TWv); a.tiles_y = div_up(out.h, TH);
    a.m_tiles = a.tiles_x * a.tiles_y * x.n;
    a.a_rows = TH * P;
    a.npix = (long long)x.n * out.h * out.w;
    a.act = p2.act; a.scale = p2.scale; a.shift = p2.shift;
    a.act1 = p1.act; a.scale1 = p1.scale; a.shift1 = p1.shift;
    a.out = out.p; a.res = (const bf16 *)x.p; a.ldr = x.ld; a.res_alpha = res_alpha; a.res_beta = res_beta;
    a.c_bufs = 2; a.stages = 4;
    a.a_stage_bytes = ((TH + 2) * P * 128 + 1023) / 1024 * 1024;
    if (getenv("B200_BLOCK_XS")) { int f = atoi(getenv("B200_BLOCK_XS")); if (f >= 2 && f <= 4) a.stages = f; }
    if (getenv("B200_BLOCK_RING")) { int f = atoi(getenv("B200_BLOCK_RING")); if (f >= 2 && f <= 4) a.c_bufs = f; }
    a.b_stages = 0;                                                   // L2 prefetch distance of the x patches, in tiles (measured: no effect)
    if (getenv("B200_BLOCK_PREFETCH")) a.b_stages = atoi(getenv("B200_BLOCK_PREFETCH"));
    {
        unsigned long long dims[4] = {64ull, (unsigned long long)x.w, (unsigned long long)x.h, (unsigned long long)x.n};
        unsigned long long strides[3] = {(unsigned long long)x.ld * 2, (unsigned long long)x.w * x.ld * 2, (unsigned long long)x.h * x.w * x.ld * 2};
        unsigned box[4] = {64, (unsigned)P, (unsigned)(TH + 2), 1};
        tc_encode_tiled(&p->maps.a[0], 0, 4, x.p, dims, strides, box, 128);
    }
    {
        unsigned long long dims[2] = {64ull, 32ull}, strides[1] = {64ull * 2};
        unsigned box[2] = {64, 32};
        tc_encode_tiled(&p->maps.a[1], 0, 2, (void *)p1.w, dims, strides, box, 128);
    }
    {
        unsigned long long dims[2] = {288ull, 64ull}, strides[1] = {288ull * 2};
        unsigned box[2] = {32, 64};
        tc_encode_tiled(&p->maps.b, 0, 2, (void *)p2.w, dims, strides, box, 64);
    }
    encode_tile_view(&p->maps.c, out, 64, a);
    encode_tile_view(&p->maps.r, x, 64, a);
    a.tmem_cols = 512;
    p->grid = a.m_tiles < 148 ? a.m_tiles : 148;
    p->smem_bytes = (size_t)(a.stages - 1) * a.a_stage_bytes + 32768 + 2 * 16384 + 9 * 4096 + 4096 + (size_t)a.c_bufs * 16384 + (512 + 4096) + 1024;
    if (p->smem_bytes > 227 * 1024) { delete p; return nullptr; }
    p->flops = 2.0 * (double)a.npix * (32.0 * 64 + 64.0 * 288);
    char buf[256];
    snprintf(buf, sizeof buf, "conv_tc BLOCK 1x1(64->32)+3x3(32->64)+shortcut tile %dx%d (pitch %d) m_tiles %d x-stages %d ring %d smem %zu grid %d",
             TWv, TH, P, a.m_tiles, a.stages, a.c_bufs, p->smem_bytes, p->grid);
    p->desc = buf;
    return p;
}

ConvTcPlan *conv_tc_plan_create(TView in, TView out, ConvParams cp, const TView *residual, float res_alpha, float res_beta, const TView *up_out,
                                int local)
{
    if (local && (residual || up_out || cp.cout_pad % 64 != 0 || out.c != cp.cout_pad || out.dtype != DT_BF16 || in.c % 64 != 0)) return nullptr;
    if (up_out && (residual || up_out->dtype != DT_BF16 || up_out->h != 2 * out.h || up_out->w != 2 * out.w || up_out->c != out.c ||
                   up_out->ld % 8 != 0 || ((uintptr_t)up_out->p & 15))) return nullptr;
    if (in.dtype != DT_BF16) return nullptr;
    if (residual && (residual->dtype != DT_BF16 || out.dtype != DT_BF16 || residual->ld % 8 != 0 || ((uintptr_t)residual->p & 15) ||
                     residual->c != out.c || residual->h != out.h || residual->w != out.w)) return nullptr;
    if (cp.act != ACT_LEAKY && cp.act != ACT_LINEAR) return nullptr;
    if (getenv("B200_DISABLE_TC")) return nullptr;
    const int C = in.c;
    int block_k = C % 64 == 0 ? 64 : (C % 32 == 0 ? 32 : (C % 16 == 0 ? 16 : 0));
    if (!block_k) return nullptr;
    if (cp.stride != 1 && cp.stride != 2) return nullptr;
    if (in.ld % 8 != 0 || ((uintptr_t)in.p & 15) || ((uintptr_t)out.p & 15)) return nullptr;
    if ((out.ld * dt_size(out.dtype)) % 16 != 0 || out.ld < cp.cout_pad) return nullptr;
    if (cp.cout_pad % 16 != 0) return nullptr;

    ConvTcPlan *p = new ConvTcPlan();
    memset(&p->maps, 0, sizeof p->maps);
    ConvTcArgs &a = p->args;
    memset(&a, 0, sizeof a);
    p->block_k = block_k;
    p->out_dtype = out.dtype;
    const int K = cp.size * cp.size * C;
    a.batch = in.n; a.OH = out.h; a.OW = out.w; a.cout_pad = cp.cout_pad; a.ldo = out.ld;
    a.size = cp.size; a.stride = cp.stride; a.pad = cp.pad; a.cin_blocks = C / block_k;
    a.num_kblocks = cp.size * cp.size * a.cin_blocks;
    a.act = cp.act; a.scale = cp.scale; a.shift = cp.shift; a.out = out.p;
    a.res = residual ? (const bf16 *)residual->p : nullptr;
    a.ldr = residual ? residual->ld : 0;
    a.res_alpha = res_alpha; a.res_beta = res_beta;
    a.npix = (long long)in.n * out.h * out.w;
    a.block_n = cp.cout_pad < 256 ? cp.cout_pad : 256;
    // a GEMM with a handful of pixel tiles (connected layer: 64 images = half a tile) is weight-bandwidth bound as well
    if (!local && cp.size == 1 && cp.cout_pad >= 512 && cp.cout_pad % 64 == 0 &&
        ((long long)in.n * out.h * out.w + 127) / 128 * ((cp.cout_pad + 255) / 256) < 37) a.block_n = 64;
    if (local) a.block_n = 64;          // weight-bandwidth bound: many narrow tiles keep every SM streaming its own slab slice
    if (getenv("B200_BLOCK_N") && cp.cout_pad % atoi(getenv("B200_BLOCK_N")) == 0 && atoi(getenv("B200_BLOCK_N")) >= 64) a.block_n = atoi(getenv("B200_BLOCK_N"));
    a.n_tiles = (cp.cout_pad + a.block_n - 1) / a.block_n;
    const size_t esz = 2;

    // ---- A views ----
    if (cp.size == 1 && cp.stride == 1 && cp.pad == 0 && !up_out) {         // (a fused upsample needs spatial tiles)
        a.mode = 0;
        a.a_rows = 128;
        a.m_tiles = (int)((a.npix + 127) / 128);
        cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)a.npix};
        cuuint64_t strides[1] = {(cuuint64_t)in.ld * esz};
        cuuint32_t box[2] = {(cuuint32_t)block_k, 128};
        encode(&p->maps.a[0], in.p, 2, dims, strides, box, block_k);
    } else {
        a.mode = 1;
        // pick the rectangular pixel tile (TW x TH x TN <= 128 rows) that wastes the fewest MMA rows
        double best = -1; int bw = 1, bh = 1, bn = 1;
        for (int tw = 1; tw <= out.w && tw <= 128; ++tw)
            for (int th = 1; th <= out.h && tw * th <= 128; ++th) {
                int tn_max = 128 / (tw * th);
                if (tn_max > in.n) tn_max = in.n;
                for (int tn = 1; tn <= tn_max; ++tn) {
                    double tiles = (double)div_up(out.w, tw) * div_up(out.h, th) * div_up(in.n, tn);
                    double eff = (double)out.w * out.h * in.n / (tiles * 128.0);
                    // prefer wide tiles on ties: longer contiguous runs per TMA box row
                    double score = eff + 1e-6 * tw + 1e-9 * th;
                    if (score > best) { best = score; bw = tw; bh = th; bn = tn; }
                }
            }
        if (local) { bw = 1; bh = 1; bn = in.n < 128 ? in.n : 128; }      // one location per tile: GEMM rows = images
        a.TW = bw; a.TH = bh; a.TN = bn;
        a.tiles_x = div_up(out.w, bw); a.tiles_y = div_up(out.h, bh);
        a.m_tiles = a.tiles_x * a.tiles_y * div_up(in.n, bn);
        a.a_rows = bw * bh * bn;
        cuuint32_t box[4] = {(cuuint32_t)block_k, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
        if (cp.stride == 1) {
            cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)in.w, (cuuint64_t)in.h, (cuuint64_t)in.n};
            cuuint64_t strides[3] = {(cuuint64_t)in.ld * esz, (cuuint64_t)in.w * in.ld * esz, (cuuint64_t)in.h * in.w * in.ld * esz};
            encode(&p->maps.a[0], in.p, 4, dims, strides, box, block_k);
        } else {
            for (int py = 0; py < 2; ++py)
                for (int px = 0; px < 2; ++px) {
                    int pw = (in.w - px + 1) / 2, ph = (in.h - py + 1) / 2;
                    if (pw < 1) pw = 1;
                    if (ph < 1) ph = 1;
                    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)pw, (cuuint64_t)ph, (cuuint64_t)in.n};
                    cuuint64_t strides[3] = {(cuuint64_t)2 * in.ld * esz, (cuuint64_t)2 * in.w * in.ld * esz, (cuuint64_t)in.h * in.w * in.ld * esz};
                    void *base = (unsigned char *)in.p + ((size_t)py * in.w + px) * in.ld * esz;
                    encode(&p->maps.a[py * 2 + px], base, 4, dims, strides, box, block_k);
                }
        }
    }
    // ---- single-CTA patch kernel (mode 2, pair 0): few input channels, all weights resident ----------------------------
    if (a.mode == 1 && !local && !getenv("B200_NO_PATCH") && cp.size == 3 && cp.pad == 1 && out.dtype == DT_BF16 && a.n_tiles == 1 &&
        (cp.cout_pad == 64 || cp.cout_pad == 128 || cp.cout_pad == 256 || (cp.cout_pad == 32 && C == 16 && cp.stride == 1)) &&
        out.c == cp.cout_pad &&
        ((cp.stride == 1 && (C == 16 || C == 32 || C == 64)) || (cp.stride == 2 && C == 32 && in.ld == 32 && in.w % 2 == 0))) {
        const bool s2 = cp.stride == 2;
        const int a_k = s2 ? 64 : C, b_k = a_k, row_bytes = a_k * 2;
        const int halo_x = s2 ? 1 : 2;
        // filters per CTA: all of them while the resident weights leave room for the rings, else 64-filter slices
        // handled by neighbouring CTAs (the patch is then read n_split times, all but the first from L2)
        int N = cp.cout_pad <= 128 ? cp.cout_pad : 64;
        if ((s2 ? 6 : 9) * N * b_k * 2 > 80 * 1024) N = 64;
        if (getenv("B200_PATCH_SPLIT")) N = 64;
        const int n_split = cp.cout_pad / N;
        if (148 % n_split != 0) N = 0;
        const int nb = s2 ? 6 : 9;
        const int b_tile = (N * b_k * 2 + 1023) / 1024 * 1024;
        int groups = 2;
        if (getenv("B200_PATCH_GROUPS")) groups = atoi(getenv("B200_PATCH_GROUPS")) == 1 ? 1 : 2;
        const int aux_bytes = 512 + 2 * 512 * 4, slot_bytes = N >= 64 ? (N / 64) * 16384 : 8192;      // 32 filters: 64-byte rows
        const int np = s2 ? 2 : 1;
        auto stage_bytes_for = [&](int P) {
            const int max_shift = s2 ? P + 1 : 2 * P + 2;
            return np * (((max_shift + 128) * row_bytes + 1023) / 1024 * 1024);
        };
        const int room = 227 * 1024 - 1024 - aux_bytes - nb * b_tile;         // patch ring + output ring share this
        // tile: P = TW + halo_x patch columns, TH = 128 / P rows.  MEASURED (YOLOv3 layers 1 and 3): a tile costs about
        // 4 cycles per TMA row moved (patch + store + residual rows) plus ~38 cycles per TMA instruction.
        double best = 1e30; int bTW = 0;
        for (int tw = 4; tw <= out.w && tw + halo_x <= 128; ++tw) {
            int P = tw + halo_x, th = 128 / P; if (th > out.h) th = out.h;
            if (3 * stage_bytes_for(P) + 2 * slot_bytes > room) continue;
            double tiles = (double)div_up(out.w, tw) * div_up(out.h, th);
            double rows = (s2 ? (2.0 * th + 1) * P : (th + 2.0) * P) + (residual ? 2.0 : 1.0) * th * tw * (N / 64);
            double ops = np + (residual ? 2.0 : 1.0) * th * (N / 64);
            double cost = tiles * (4.0 * rows + 38.0 * ops);
            if (cost < best) { best = cost; bTW = tw; }
        }
        if (getenv("B200_PATCH_TW")) { int f = atoi(getenv("B200_PATCH_TW")); if (f >= 1 && f <= out.w && 3 * stage_bytes_for(f + halo_x) + 2 * slot_bytes <= room) bTW = f; }
        if (N == 0) bTW = 0;
        if (bTW) {
            const int TWv = bTW, P = TWv + halo_x;
            int TH = 128 / P; if (TH > out.h) TH = out.h;
            const int stage_bytes = stage_bytes_for(P), patch_bytes = stage_bytes / np;
            // output ring: a fused residual is prefetched into its slot tiles ahead of the epilogue, so it wants the deeper
            // ring; whatever is left goes to patch stages (3 are enough to cover the load latency, more do not help)
            int c_bufs = residual ? 6 : 3;      // measured on YOLOv3 layer 3: 6 slots + 6 stages beat 8 + 4 and 4 + 8
            while (c_bufs > 2 && 3 * stage_bytes + c_bufs * slot_bytes > room) --c_bufs;
            if (getenv("B200_PATCH_CBUFS")) { int f = atoi(getenv("B200_PATCH_CBUFS")); if (f >= 2 && f <= 8 && 2 * stage_bytes + f * slot_bytes <= room) c_bufs = f; }
            const int sc_bytes = c_bufs * slot_bytes;
            int st = (room - sc_bytes) / stage_bytes; if (st > 8) st = 8;
            if (getenv("B200_PATCH_STAGES")) { int f = atoi(getenv("B200_PATCH_STAGES")); if (f >= 2 && f < st) st = f; }
            if (st >= 2) {
                a.mode = 2; a.pair = 0; a.resident_b = 1; a.staged = 1;
                a.halo_P = P; a.halo_TWv = TWv; a.halo_THv = TH;
                a.TW = TWv; a.TH = TH; a.TN = 1;
                a.tiles_x = div_up(out.w, TWv); a.tiles_y = div_up(out.h, TH);
                a.m_tiles = a.tiles_x * a.tiles_y * in.n;
                a.a_rows = TH * P;
                a.np = np; a.a_k = a_k; a.b_k = b_k; a.nb = nb;
                a.ep_groups = groups; a.c_bufs = c_bufs; a.n_split = n_split; a.block_n = N;
                a.a_stage_bytes = stage_bytes; a.b_stage_bytes = b_tile; a.stages = st;
                if (!s2) {
                    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)in.w, (cuuint64_t)in.h, (cuuint64_t)in.n};
                    cuuint64_t strides[3] = {(cuuint64_t)in.ld * esz, (cuuint64_t)in.w * in.ld * esz, (cuuint64_t)in.h * in.w * in.ld * esz};
                    cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)P, (cuuint32_t)(TH + 2), 1};
                    encode(&p->maps.a[0], in.p, 4, dims, strides, box, a_k);
                    a.patch_map[0] = 0; a.patch_off[0] = 0; a.patch_dx[0] = -1; a.patch_dy[0] = -1;
                    a.stage_tx = (TH + 2) * P * row_bytes;
                    a.nseg = 9;
                    for (int t = 0; t < 9; ++t) {
                        a.b_koff[t] = t * C;
                        a.seg_a[t] = ((t / 3) * P + (t % 3)) * row_bytes;
                        a.seg_b[t] = t * b_tile;
                        a.seg_k[t] = C / 16;
                    }
                } else {
                    // rows of pixel pairs of one row parity: dims {64, W/2, rows of that parity, N}
                    for (int py = 0; py < 2; ++py) {
                        const int ph = (in.h - py + 1) / 2;
                        cuuint64_t dims[4] = {64, (cuuint64_t)(in.w / 2), (cuuint64_t)(ph < 1 ? 1 : ph), (cuuint64_t)in.n};
                        cuuint64_t strides[3] = {(cuuint64_t)128, (cuuint64_t)2 * in.w * 64, (cuuint64_t)in.h * in.w * 64};
                        cuuint32_t box[4] = {64, (cuuint32_t)P, (cuuint32_t)(py ? TH + 1 : TH), 1};
                        encode(&p->maps.a[py], (unsigned char *)in.p + (size_t)py * in.w * 64, 4, dims, strides, box, 64);
                    }
                    // patch 0 = odd input rows (taps ky = 0, 2), patch 1 = even input rows (tap ky = 1)
                    a.patch_map[0] = 1; a.patch_off[0] = 0;           a.patch_dx[0] = -1; a.patch_dy[0] = -1;
                    a.patch_map[1] = 0; a.patch_off[1] = patch_bytes; a.patch_dx[1] = -1; a.patch_dy[1] = 0;
                    a.stage_tx = ((TH + 1) + TH) * P * row_bytes;
                    a.nseg = 6;
                    for (int ky = 0; ky < 3; ++ky) {
                        const int base = (ky == 1 ? patch_bytes : 0) + (ky == 2 ? P : 0) * row_bytes;
                        a.b_koff[2 * ky] = ky * 96 + 32;  a.b_koff[2 * ky + 1] = ky * 96;
                        a.seg_a[2 * ky] = base + row_bytes;   a.seg_b[2 * ky] = (2 * ky) * b_tile;         a.seg_k[2 * ky] = 4;      // kx = 1,2: pair ox
                        a.seg_a[2 * ky + 1] = base + 64;      a.seg_b[2 * ky + 1] = (2 * ky + 1) * b_tile; a.seg_k[2 * ky + 1] = 2;  // kx = 0: upper half of pair ox-1
                    }
                }
                {
                    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)cp.cout_pad};
                    cuuint64_t strides[1] = {(cuuint64_t)K * esz};
                    cuuint32_t box[2] = {(cuuint32_t)b_k, (cuuint32_t)N};
                    encode(&p->maps.b, (void *)cp.w, 2, dims, strides, box, b_k);
                }
                encode_tile_view(&p->maps.c, out, cp.cout_pad, a, N >= 64 ? 64 : 32);
                if (residual) encode_tile_view(&p->maps.r, *residual, cp.cout_pad, a, N >= 64 ? 64 : 32);
                int fit = 512 / N;
                a.acc_stages = fit >= 8 ? 8 : (fit >= 4 ? 4 : 2);
                if (getenv("B200_PATCH_ACC")) { int f = atoi(getenv("B200_PATCH_ACC")); if (f >= 1 && f < a.acc_stages) a.acc_stages = f; }
                int cols = a.acc_stages * N;
                a.tmem_cols = cols <= 32 ? 32 : (cols <= 64 ? 64 : (cols <= 128 ? 128 : (cols <= 256 ? 256 : 512)));
                p->grid = n_split * (a.m_tiles < 148 / n_split ? a.m_tiles : 148 / n_split);
                p->smem_bytes = (size_t)st * stage_bytes + (size_t)nb * b_tile + sc_bytes + aux_bytes + 1024;
                p->flops = 2.0 * (double)a.npix * out.c * K;
                char buf3[320];
                snprintf(buf3, sizeof buf3, "conv_tc PATCH s%d k%d n%d x%d tile %dx%d (pitch %d) m_tiles %d patches %d segs %d stages %d acc %d ring %d groups %d smem %zu grid %d residentB%s stagedEpilogue",
                         cp.stride, a_k, N, n_split, TWv, TH, P, a.m_tiles, np, a.nseg, st, a.acc_stages, c_bufs, groups, p->smem_bytes, p->grid, a.res ? " +residual" : "");
                p->desc = buf3;
                return p;
            }
        }
    }
    // ---- halo-patch mode (mode 2): stride-1 odd-size convolutions, one patch load per 64-channel block -----------------
    // cost model: per 64-channel block a CTA is bound by max(MMA cycles, ~3 cycles per TMA row); compare cycles per VALID
    // output pixel of the tap-per-box pair kernel with those of the patch kernel
    // MEASURED (YOLOv3-416 b64, ncu): the patch kernel keeps the tensor pipe 61 % busy versus 56 % for the tap-per-box
    // kernel, but only ~81 % of its MMA rows are valid outputs, so it ends up 5-20 % slower on every layer; the L2/TMA row
    // rate is therefore NOT the binding limit of the pair kernel (shared-memory bandwidth / MMA issue is).  The mode stays
    // available for experiments (B200_HALO=1 picks it by the cost model, B200_FORCE_HALO=1 always) but is off by default.
    if (a.mode == 1 && !local && (getenv("B200_HALO") || getenv("B200_FORCE_HALO")) && cp.stride == 1 && (cp.size & 1) && cp.size >= 3 && cp.pad == cp.size / 2 &&
        block_k == 64 && a.block_n % 32 == 0 && a.block_n >= 64 && out.dtype == DT_BF16) {
        const int halo = cp.size - 1, taps = cp.size * cp.size;
        int tiles_x = div_up(out.w + halo, 128) > 1 ? div_up(out.w, 128 - halo) : 1;
        int TWv = div_up(out.w, tiles_x);
        int P = TWv + halo;
        int TH = 128 / P; if (TH < 1) TH = 1; if (TH > out.h) TH = out.h;
        int tiles_y = div_up(out.h, TH);
        TH = div_up(out.h, tiles_y);
        const int half_n = a.block_n / 2;
        const int patch_rows = (TH + halo) * P;
        const double mma_cycles = (double)taps * 4.0 * (a.block_n / 2.0);                     // cta_group::2, M=256: N/2 cycles per K=16 step
        const double rows_tap = (double)taps * (128 + half_n), rows_halo = patch_rows + (double)taps * half_n;
        const double eff1 = (double)a.npix / ((double)a.m_tiles * 128.0);
        const double valid_halo = (double)out.w * out.h / ((double)tiles_x * tiles_y);
        const double cost_tap = (mma_cycles > 3.0 * rows_tap ? mma_cycles : 3.0 * rows_tap) / (128.0 * eff1);
        const double cost_halo = (mma_cycles > 3.0 * rows_halo ? mma_cycles : 3.0 * rows_halo) / valid_halo;
        if (P <= 256 && TH + halo <= 256 && (cost_halo < 0.95 * cost_tap || getenv("B200_FORCE_HALO")) &&
            (long long)tiles_x * tiles_y * in.n >= 2) {
            a.mode = 2;
            a.halo_P = P; a.halo_TWv = TWv; a.halo_THv = TH;
            a.TW = TWv; a.TH = TH; a.TN = 1;
            a.tiles_x = tiles_x; a.tiles_y = tiles_y;
            a.m_tiles = tiles_x * tiles_y * in.n;
            a.a_rows = TH * P;
            cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)in.w, (cuuint64_t)in.h, (cuuint64_t)in.n};
            cuuint64_t strides[3] = {(cuuint64_t)in.ld * esz, (cuuint64_t)in.w * in.ld * esz, (cuuint64_t)in.h * in.w * in.ld * esz};
            cuuint32_t box[4] = {64, (cuuint32_t)P, (cuuint32_t)(TH + halo), 1};
            encode(&p->maps.a[0], in.p, 4, dims, strides, box, 64);
            a.a_stage_bytes = ((128 + halo * P + halo) * 128 + 1023) / 1024 * 1024;
        }
    }
    // ---- epilogue staging / weight residency / CTA pairing -------------------------------------------------
    // staged epilogue (TMEM -> registers -> swizzled smem tile -> TMA store, residual TMA-loaded into the same tile) is used
    // where a shortcut is fused: the per-row residual reads of the direct epilogue are what made fused layers slow.
    const int a_bytes_ = 128 * block_k * 2;
    const int budget_all = 227 * 1024 - 1024 - (512 + 4096);
    // staged epilogue = the ring epilogue (ring_roles) for the tap-per-box kernels, the serial staged path of run_epilogue
    // for the (opt-in) halo pair kernel.
    const bool stageable64 = out.dtype == DT_BF16 && a.block_n % 64 == 0 && cp.cout_pad % 64 == 0 && out.c == cp.cout_pad;
    // 32-filter sub-tiles: bf16 layers with 32 (mod 64) filters and the fp32 head convolutions (255 -> 256 padded filters:
    // the pad column lands in the row's own padding, never in a neighbour's slice of a concat buffer)
    const bool stageable32 = !residual && a.block_n % 32 == 0 && cp.cout_pad % 32 == 0 && (out.c == cp.cout_pad || out.ld == cp.cout_pad) &&
                             a.mode != 2 && !getenv("B200_NO_RING32");
    const bool stageable = stageable64 || stageable32;
    // MEASURED (YOLOv3-416 b64): the ring epilogue wins on every stageable layer (1x1 layers -10..-20 %, fused shortcuts
    // -8 %) except the stride-2 3x3 layers without a residual, which lose the pipeline stage the ring's slots cost (+3 %).
    const bool ring_pays = residual || !(cp.size == 3 && cp.stride == 2 && a.block_n == 256);
    const bool want_staged = stageable && !getenv("B200_NO_STAGED") && (getenv("B200_RING_RESIDUAL_ONLY") ? residual != nullptr : ring_pays);
    const bool use_ring = want_staged && a.mode != 2 && !getenv("B200_NO_RING");
    int ring_slots = 3;                                  // measured: 3 slots beat 2 and 4 (a 4th costs a pipeline stage)
    if (getenv("B200_RING_SLOTS")) { int f = atoi(getenv("B200_RING_SLOTS")); if (f >= 2 && f <= 4) ring_slots = f; }
    const int sc_bytes = use_ring ? ring_slots * 16384 : (a.block_n / 64) * 16384;
    const long long slab_ = (long long)a.num_kblocks * ((a.block_n * block_k * 2 + 1023) / 1024 * 1024);
    const bool could_reside = a.n_tiles == 1 && !local && !getenv("B200_NO_RESIDENT_B") &&
                              slab_ + (want_staged ? 3LL * a_bytes_ + sc_bytes : 4LL * a_bytes_) <= budget_all;
    if (a.mode == 2) {
        // the patch kernel always runs as a CTA pair with its own smem budget (patch ring + weight ring + staging tile)
        a.pair = 1; a.resident_b = 0;
        a.b_stage_bytes = ((a.block_n / 2) * block_k * 2 + 1023) / 1024 * 1024;
        int budget2 = 227 * 1024 - 1024 - (512 + 4096);
        a.staged = (want_staged && stageable64 && sc_bytes + 2 * a.a_stage_bytes + 3 * a.b_stage_bytes <= budget2) ? 1 : 0;
        if (a.staged) budget2 -= sc_bytes;
        int a_st = 2;
        int b_st = (budget2 - a_st * a.a_stage_bytes) / a.b_stage_bytes;
        if (b_st > 8) { b_st = 8; int extra = (budget2 - b_st * a.b_stage_bytes) / a.a_stage_bytes; a_st = extra > 4 ? 4 : extra; }
        if (b_st < 3) { delete p; return nullptr; }
        a.stages = a_st; a.b_stages = b_st;
        p->smem_bytes = (size_t)a_st * a.a_stage_bytes + (size_t)b_st * a.b_stage_bytes + (a.staged ? sc_bytes : 0) + (512 + 4096) + 1024;
        {
            cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)cp.cout_pad};
            cuuint64_t strides[1] = {(cuuint64_t)K * esz};
            cuuint32_t box[2] = {(cuuint32_t)block_k, (cuuint32_t)(a.block_n / 2)};
            encode(&p->maps.b, (void *)cp.w, 2, dims, strides, box, block_k);
        }
        if (a.staged) {
            encode_tile_view(&p->maps.c, out, cp.cout_pad, a);
            if (residual) encode_tile_view(&p->maps.r, *residual, cp.cout_pad, a);
        }
        a.acc_stages = 2;
        int cols2 = 2 * a.block_n;
        a.tmem_cols = cols2 <= 32 ? 32 : (cols2 <= 64 ? 64 : (cols2 <= 128 ? 128 : (cols2 <= 256 ? 256 : 512)));
        int pair_tiles = ((a.m_tiles + 1) / 2) * a.n_tiles;
        p->grid = 2 * (pair_tiles < 74 ? pair_tiles : 74);
        p->flops = 2.0 * (double)a.npix * out.c * K;
        char buf2[320];
        snprintf(buf2, sizeof buf2, "conv_tc HALO k64 n%d tile %dx%d (pitch %d) m_tiles %d n_tiles %d patch-stages %d weight-stages %d smem %zu grid %d PAIR(cta_group::2)%s%s",
                 a.block_n, a.halo_TWv, a.halo_THv, a.halo_P, a.m_tiles, a.n_tiles, a_st, b_st, p->smem_bytes, p->grid,
                 a.res ? " +residual" : "", a.staged ? " stagedEpilogue" : "");
        p->desc = buf2;
        return p;
    }
    a.pair = (!getenv("B200_NO_PAIR") && !local && !could_reside && a.block_n % 32 == 0 && a.block_n >= 64 && a.m_tiles >= 2 &&
              (long long)a.num_kblocks * a.block_n >= 4 * 256) ? 1 : 0;
    // ---- B view ----
    {
        cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)cp.cout_pad};
        cuuint64_t strides[1] = {(cuuint64_t)K * esz};
        cuuint32_t box[2] = {(cuuint32_t)block_k, (cuuint32_t)(a.pair ? a.block_n / 2 : a.block_n)};
        if (!local) encode(&p->maps.b, (void *)cp.w, 2, dims, strides, box, block_k);
        else {                               // [location][filters][K]: the tile's location picks the slab
            cuuint64_t dims3[3] = {(cuuint64_t)K, (cuuint64_t)cp.cout_pad, (cuuint64_t)out.h * out.w};
            cuuint64_t strides3[2] = {(cuuint64_t)K * esz, (cuuint64_t)K * cp.cout_pad * esz};
            cuuint32_t box3[3] = {(cuuint32_t)block_k, (cuuint32_t)a.block_n, 1};
            encode(&p->maps.b, (void *)cp.w, 3, dims3, strides3, box3, block_k);
            a.local = 1; a.ss_stride = cp.cout_pad;
        }
    }
    // ---- smem / tmem budget ----
    const int a_bytes = 128 * block_k * 2;
    a.b_stage_bytes = ((a.pair ? a.block_n / 2 : a.block_n) * block_k * 2 + 1023) / 1024 * 1024;
    const int aux_bytes = 512 + 2 * 512 * 4;
    int budget = 227 * 1024 - 1024 - aux_bytes;
    a.staged = 0;
    if (want_staged && (use_ring || stageable64)) {
        const long long need = could_reside ? slab_ + 3LL * a_bytes : 3LL * (a_bytes + a.b_stage_bytes);
        if (need + sc_bytes <= budget) { a.staged = 1; budget -= sc_bytes; }
    }
    a.ring = (a.staged && use_ring) ? 1 : 0;
    a.c_bufs = ring_slots;
    a.sub_cols = stageable64 ? 64 : 32;
    a.out_f32 = out.dtype == DT_F32 ? 1 : 0;
    if (local && !a.ring) { delete p; return nullptr; }      // only the ring epilogue knows the per-location bias rows
    // weight-stationary variant: when one filter tile covers all filters and its whole [block_n x K] slab fits next
    // to >= 4 activation stages, load it once per CTA and stream only activations (halves the TMA rows per k-block)
    const long long slab = (long long)a.num_kblocks * a.b_stage_bytes;
    a.resident_b = (!a.pair && could_reside && slab + 3LL * a_bytes <= budget) ? 1 : 0;
    int stages;
    if (a.resident_b) {
        stages = (int)((budget - slab) / a_bytes);
        if (stages > 8) stages = 8;
        p->smem_bytes = (size_t)stages * a_bytes + (size_t)slab + (a.staged ? sc_bytes : 0) + aux_bytes + 1024;
    } else {
        stages = budget / (a_bytes + a.b_stage_bytes);
        if (stages > 8) stages = 8;
        if (stages < 2) { delete p; return nullptr; }
        p->smem_bytes = (size_t)stages * (a_bytes + a.b_stage_bytes) + (a.staged ? sc_bytes : 0) + aux_bytes + 1024;
    }
    a.stages = stages;
    if (a.staged) {
        encode_tile_view(&p->maps.c, out, cp.cout_pad, a, a.sub_cols);
        if (residual) encode_tile_view(&p->maps.r, *residual, cp.cout_pad, a, a.sub_cols);
    }
    if (up_out) {
        // the conv's own output is not written: phase (dy, dx) of the upsampled tensor is a strided view with the conv
        // output's geometry, so the same tile coordinates address all four copies
        if (!a.ring || a.mode != 1 || a.sub_cols != 64) { delete p; return nullptr; }
        a.upsample = 1;
        for (int ph = 0; ph < 4; ++ph) {
            const int dy = ph >> 1, dx = ph & 1;
            const unsigned long long esz = 2, ld = (unsigned long long)up_out->ld, W2 = (unsigned long long)up_out->w, H2 = (unsigned long long)up_out->h;
            unsigned long long dims[4] = {(unsigned long long)cp.cout_pad, (unsigned long long)out.w, (unsigned long long)out.h, (unsigned long long)out.n};
            unsigned long long strides[3] = {2 * ld * esz, 2 * W2 * ld * esz, H2 * W2 * ld * esz};
            unsigned box[4] = {64, (unsigned)a.TW, (unsigned)a.TH, (unsigned)a.TN};
            void *base = (unsigned char *)up_out->p + ((size_t)dy * W2 + dx) * ld * esz;
            tc_encode_tiled(ph == 0 ? (void *)&p->maps.c : (void *)&p->maps.cu[ph - 1], 0, 4, base, dims, strides, box, 128);
        }
    }
    a.acc_stages = 2;
    // deeper accumulator rings were measured (YOLOv3-416 b64) to give no gain on the small-filter layers: they are bound
    // by the TMA row rate, not by the MMA->epilogue hand-off.  Kept switchable for experiments.
    if (a.ring && !getenv("B200_NO_DEEP_TMEM")) {        // the ring epilogue takes its constants from global memory: any depth works,
        int fit = 512 / a.block_n;                       // and short K passes (1x1 layers) need the MMA to run tiles ahead
        a.acc_stages = fit >= 8 ? 8 : (fit >= 4 ? 4 : 2);
    } else if (a.n_tiles == 1 && getenv("B200_DEEP_TMEM")) {   // per-tile constants are hoisted, so any number of buffers works
        int fit = 512 / a.block_n;
        a.acc_stages = fit >= 8 ? 8 : (fit >= 4 ? 4 : 2);
    }
    int cols = a.acc_stages * a.block_n;
    a.tmem_cols = cols <= 32 ? 32 : (cols <= 64 ? 64 : (cols <= 128 ? 128 : (cols <= 256 ? 256 : 512)));
    int tiles = a.m_tiles * a.n_tiles;
    p->grid = tiles < 148 ? tiles : 148;
    a.split = 1; a.split_from = 0x7fffffff; a.vtiles = tiles;
    if (a.pair) {
        int pair_tiles = ((a.m_tiles + 1) / 2) * a.n_tiles;
        p->grid = 2 * (pair_tiles < 74 ? pair_tiles : 74);
        a.vtiles = pair_tiles;
        // tail splitting: when the last wave would keep at most half of the 74 pairs busy, its tiles are cut into 2 or 4
        // filter slices (>= 64 filters each) so that all pairs share it: e.g. 184 tiles = 2 waves + 36 -> 2 waves + 72 halves
        const int rem = pair_tiles > 74 ? pair_tiles % 74 : 0;
        if (a.ring && rem > 0 && 2 * rem <= 74 && !getenv("B200_NO_TAIL_SPLIT")) {
            int sp = (4 * rem <= 74 && a.block_n % 256 == 0) ? 4 : 2;
            if (getenv("B200_TAIL_SPLIT")) { int f = atoi(getenv("B200_TAIL_SPLIT")); if (f == 2 || f == 4) sp = f; }
            if ((a.block_n / sp) % 64 == 0 && a.block_n / sp >= 64 && a.sub_cols == 64) {
                a.split = sp; a.split_from = pair_tiles - rem; a.vtiles = pair_tiles - rem + rem * sp;
                cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)cp.cout_pad};
                cuuint64_t strides[1] = {(cuuint64_t)K * esz};
                cuuint32_t box[2] = {(cuuint32_t)block_k, (cuuint32_t)(a.block_n / sp / 2)};
                encode(&p->maps.b2, (void *)cp.w, 2, dims, strides, box, block_k);
            }
        }
    }
    p->flops = 2.0 * (double)a.npix * out.c * K;
    char buf[256];
    snprintf(buf, sizeof buf, "conv_tc mode%d k%d n%d tile %dx%dx%d rows %d m_tiles %d n_tiles %d stages %d smem %zu grid %d%s%s",
             a.mode, block_k, a.block_n, a.TW, a.TH, a.TN, a.a_rows, a.m_tiles, a.n_tiles, stages, p->smem_bytes, p->grid,
             a.pair ? " PAIR(cta_group::2)" : (a.resident_b ? " residentB" : ""), a.res ? " +residual" : "");

    p->desc = buf;
    if (a.ring) p->desc += " ringEpilogue(" + std::to_string(a.c_bufs) + ")";
    else if (a.staged) p->desc += " stagedEpilogue";
    if (a.upsample) p->desc += " +upsample2x";
    if (a.local) p->desc += " unshared(local)";
    if (a.split > 1) p->desc += " tailSplit(" + std::to_string(a.split) + "x" + std::to_string(a.vtiles - a.split_from) + ")";
    return p;
}
