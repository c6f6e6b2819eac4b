// conv_tc_plan.cu — host side of the tcgen05 convolution family (overview in conv_tc.cu): TMA tensor maps, the per-layer planner
// (tile shape, kernel variant, pipeline depth, epilogue, tail splitting) and the launch dispatch.
#include "conv_tc_plan.h"
#include <vector>

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

typedef CUresult (*EncodeIm2colFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                   const int *, const int *, cuuint32_t, cuuint32_t, const cuuint32_t *, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeIm2colFn encode_im2col_fn()
{
    static EncodeIm2colFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        B200_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &q));
        if (!p || q != cudaDriverEntryPointSuccess) { fprintf(stderr, "b200-darknet: cuTensorMapEncodeIm2col unavailable\n"); abort(); }
        fn = (EncodeIm2colFn)p;
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

// im2col-mode view of an NHWC bf16 activation for a size x size convolution (stride, pad as in convolutional_layer.c:325-335):
// one load gathers `block_k` channels of 128 consecutive output pixels for one filter tap.  The bounding box [-pad, W-1+pad-(size-1)]
// holds the tap origins of all output positions; the traversal stride is the convolution's.
static void encode_im2col(CUtensorMap *map, const TView &in, int size, int stride, int pad, int block_k)
{
    const cuuint64_t esz = 2;
    cuuint64_t dims[4] = {(cuuint64_t)in.c, (cuuint64_t)in.w, (cuuint64_t)in.h, (cuuint64_t)in.n};
    cuuint64_t strides[3] = {(cuuint64_t)in.ld * esz, (cuuint64_t)in.w * in.ld * esz, (cuuint64_t)in.h * in.w * in.ld * esz};
    int lower[2] = {-pad, -pad};
    int upper[2] = {pad - (size - 1), pad - (size - 1)};
    cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    CUresult r = encode_im2col_fn()(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, in.p, dims, strides, lower, upper, (cuuint32_t)block_k, 128, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(block_k), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        fprintf(stderr, "b200-darknet: cuTensorMapEncodeIm2col failed (%d) %dx%dx%dx%d size %d stride %d pad %d\n", (int)r, in.n, in.h, in.w, in.c, size, stride, pad);
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

void launch_conv_tc(ConvTcPlan *p, cudaStream_t s)
{
    if (p->args.block) conv_tc_launch_block(p, s);
    else if (p->args.mode == 2) conv_tc_launch_patch(p, s);
    else conv_tc_launch_tap(p, s);
    B200_LAUNCHED();
    if (p->args.ksplit > 1) { conv_tc_launch_splitk_finalize(p, s); B200_LAUNCHED(); }
}

void conv_tc_plan_destroy(ConvTcPlan *p)
{
    if (!p) return;
    if (p->sk.ws) cudaFree(p->sk.ws);
    if (p->sk.ones) cudaFree(p->sk.ones);
    delete p;
}

static int pair_tiles_for_split(const ConvTcArgs &a) { return ((a.m_tiles + 1) / 2) * a.n_tiles; }

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
    a.b_stages = 0;                                                   // L2 prefetch distance of the x patches, in tiles (measured: no effect)
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
    p->in = in; p->out = out; p->res = residual ? *residual : TView{nullptr, 0, 0, 0, 0, 0, 0};
    p->w = cp.w; p->K = K; p->flow_ok = false;
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
        const int n_split = cp.cout_pad / N;
        if (148 % n_split != 0) N = 0;
        const int nb = s2 ? 6 : 9;
        const int b_tile = (N * b_k * 2 + 1023) / 1024 * 1024;
        int groups = 2;
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
        if (N == 0) bTW = 0;
        if (bTW) {
            const int TWv = bTW, P = TWv + halo_x;
            int TH = 128 / P; if (TH > out.h) TH = out.h;
            const int stage_bytes = stage_bytes_for(P), patch_bytes = stage_bytes / np;
            // output ring: a fused residual is prefetched into its slot tiles ahead of the epilogue, so it wants the deeper
            // ring; whatever is left goes to patch stages (3 are enough to cover the load latency, more do not help)
            int c_bufs = residual ? 6 : 3;      // measured on YOLOv3 layer 3: 6 slots + 6 stages beat 8 + 4 and 4 + 8
            while (c_bufs > 2 && 3 * stage_bytes + c_bufs * slot_bytes > room) --c_bufs;
            const int sc_bytes = c_bufs * slot_bytes;
            int st = (room - sc_bytes) / stage_bytes; if (st > 8) st = 8;
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
    // ---- TMA im2col mode: every other filter size > 1 --------------------------------------------------------------------
    // The tile becomes ANY 128 consecutive output pixels (no rectangle has to divide the map: 13 x 13 maps lose no MMA rows,
    // 85 instead of 91 tiles at batch 64), stride 2 needs no parity-phase views, and the output / residual tiles are the dense
    // 2-D boxes of the 1x1 layers.  A fused upsample and the unshared (local) layers keep their spatial tiles.
    if (a.mode == 1 && !local && !up_out && !getenv("B200_NO_IM2COL") && cp.pad < 128 && cp.size - 1 - cp.pad < 128 &&
        (long long)out.h * out.w * in.n < (1LL << 31) - 256 && out.h == (in.h + 2 * cp.pad - cp.size) / cp.stride + 1 &&
        out.w == (in.w + 2 * cp.pad - cp.size) / cp.stride + 1) {
        a.mode = 0; a.im2col = 1;
        a.TW = a.TH = a.TN = 0; a.tiles_x = a.tiles_y = 1;
        a.a_rows = 128;
        a.m_tiles = (int)((a.npix + 127) / 128);
        memset(&p->maps.a, 0, sizeof p->maps.a);
        encode_im2col(&p->maps.a[0], in, cp.size, cp.stride, cp.pad, block_k);
    }
    // ---- epilogue staging / weight residency / CTA pairing -------------------------------------------------
    // staged epilogue (TMEM -> registers -> swizzled smem tile -> TMA store, residual TMA-loaded into the same tile) is used
    // where a shortcut is fused: the per-row residual reads of the direct epilogue are what made fused layers slow.
    const int a_bytes_ = 128 * block_k * 2;
    const int budget_all = 227 * 1024 - 1024 - (512 + 4096);
    // staged epilogue = the ring epilogue (ring_roles) of the tap-per-box kernels
    const bool stageable64 = out.dtype == DT_BF16 && a.block_n % 64 == 0 && cp.cout_pad % 64 == 0 && out.c == cp.cout_pad;
    // 32-filter sub-tiles: bf16 layers with 32 (mod 64) filters and the fp32 head convolutions (255 -> 256 padded filters:
    // the pad column lands in the row's own padding, never in a neighbour's slice of a concat buffer)
    const bool stageable32 = !residual && a.block_n % 32 == 0 && cp.cout_pad % 32 == 0 && (out.c == cp.cout_pad || out.ld == cp.cout_pad) &&
                             a.mode != 2;
    const bool stageable = stageable64 || stageable32;
    // (items shorter than 4 k-blocks of 128 filters are all per-item overhead: those layers keep their own kernels)
    p->flow_ok = block_k == 64 && a.mode == 0 && !local && !up_out && stageable64 && (cp.cout_pad <= 256 || cp.cout_pad % 256 == 0) &&
                 out.ld % 8 == 0 && (!residual || residual->ld % 8 == 0) && a.num_kblocks * (cp.cout_pad < 256 ? cp.cout_pad : 256) >= 4 * 128;
    // MEASURED (YOLOv3-416 b64): the ring epilogue wins on every stageable layer (1x1 layers -10..-20 %, fused shortcuts
    // -8 %) except the stride-2 3x3 layers without a residual, which lose the pipeline stage the ring's slots cost (+3 %).
    // split-K (small batches): a mode-0 layer whose pixel tiles occupy at most half of the 74 CTA pairs and whose K loop is long
    // is cut along K so that the idle pairs shorten its critical path; the partial sums go through an fp32 workspace
    // (ring epilogue with 32-column fp32 sub-tiles) and splitk_finalize_kernel applies batch-norm, activation and shortcut.
    int want_ksplit = 1;
    {
        const long long tiles_pair = (long long)((a.m_tiles + 1) / 2) * a.n_tiles;
        const bool flows_requested = getenv("B200_FLOW") && atoi(getenv("B200_FLOW")) != 0;     // a flow member keeps its whole K loop
        if (a.mode == 0 && !local && !up_out && !getenv("B200_NO_SPLITK") && !flows_requested && stageable64 && out.dtype == DT_BF16 && block_k == 64 &&
            a.block_n % 64 == 0 && a.block_n >= 64 && a.m_tiles >= 2 && tiles_pair <= 37 && a.num_kblocks >= 18) {
            // MEASURED (batch 1, forward pass): at most 8 ranges of at least 4 k-blocks — YOLOv3 1.120 -> 0.979 ms, YOLOv3-tiny 0.206 ->
            // 0.177 ms, YOLOv2 0.542 -> 0.380 ms; up to 32 ranges 1.037 / 0.185 / 0.421 (the finalize pass reads every slab);
            // splitting layers from 8 k-blocks on loses (1.054 ms): below 18 the second launch costs more than the K loop saves
            int s_max = (int)(74 / tiles_pair), by_k = a.num_kblocks / 4;
            int S = s_max < by_k ? s_max : by_k;
            if (S > 8) S = 8;
            if (S >= 2) want_ksplit = S;
        }
    }
    const bool ring_pays = want_ksplit > 1 || residual || !(cp.size == 3 && cp.stride == 2 && a.block_n == 256);
    const bool want_staged = stageable && ring_pays;
    const bool use_ring = want_staged && a.mode != 2 && !getenv("B200_NO_RING");
    int ring_slots = 3;                                  // measured: 3 slots beat 2 and 4 (a 4th costs a pipeline stage)
    const int sc_bytes = use_ring ? ring_slots * 16384 : (a.block_n / 64) * 16384;
    const long long slab_ = (long long)a.num_kblocks * ((a.block_n * block_k * 2 + 1023) / 1024 * 1024);
    const bool could_reside = a.n_tiles == 1 && !local && !getenv("B200_NO_RESIDENT_B") &&
                              slab_ + (want_staged ? 3LL * a_bytes_ + sc_bytes : 4LL * a_bytes_) <= budget_all;
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
#ifdef B200_EXPERIMENTS
        if (getenv("B200_EXP_STAGES") && atoi(getenv("B200_EXP_STAGES")) < stages) stages = atoi(getenv("B200_EXP_STAGES"));   // timing experiment
#endif
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
    if (a.ring) {                                        // the ring epilogue takes its constants from global memory: any depth works,
        int fit = 512 / a.block_n;                       // and short K passes (1x1 layers) need the MMA to run tiles ahead
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
        // filter slices (>= 64 filters each) so that all pairs share it: e.g. 184 tiles = 2 waves + 36 -> 2 waves + 72 halves.
        // MEASURED (round 2): splitting fuller tails does not pay — 338 tiles = 4 waves + 42 cut into 168 quarters (0.75 of a
        // wave by the arithmetic) ran at 0.099 ms against 0.079 unsplit: every slice re-reads its A tile and pays its own fill.
        const int rem = pair_tiles > 74 ? pair_tiles % 74 : 0;
        if (a.ring && rem > 0 && 2 * rem <= 74 && !getenv("B200_NO_TAIL_SPLIT")) {
            int sp = (4 * rem <= 74 && a.block_n % 256 == 0) ? 4 : 2;
            if ((a.block_n / sp) % 64 == 0 && a.block_n / sp >= 64 && a.sub_cols == 64) {
                a.split = sp; a.split_from = pair_tiles - rem; a.vtiles = pair_tiles - rem + rem * sp;
                cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)cp.cout_pad};
                cuuint64_t strides[1] = {(cuuint64_t)K * esz};
                cuuint32_t box[2] = {(cuuint32_t)block_k, (cuuint32_t)(a.block_n / sp / 2)};
                encode(&p->maps.b2, (void *)cp.w, 2, dims, strides, box, block_k);
            }
        }
    }
#ifdef B200_EXPERIMENTS
    if (getenv("B200_EXP") && a.pair && (a.mode == 0 || cp.stride == 1)) a.exp = atoi(getenv("B200_EXP"));
#endif
    memset(&p->sk, 0, sizeof p->sk);
    if (want_ksplit > 1 && a.pair && a.ring && a.split == 1) {
        const int kb_per = (a.num_kblocks + want_ksplit - 1) / want_ksplit;
        const int S = (a.num_kblocks + kb_per - 1) / kb_per;
        const long long slab_rows = (long long)a.m_tiles * 128;
        const size_t ws_bytes = (size_t)S * slab_rows * cp.cout_pad * sizeof(float);
        if (S >= 2 && ws_bytes <= (64u << 20)) {
            ConvTcPlan::SplitK &k = p->sk;
            B200_CHECK(cudaMalloc(&k.ws, ws_bytes));
            B200_CHECK(cudaMalloc(&k.ones, 2 * (size_t)cp.cout_pad * sizeof(float)));
            std::vector<float> host(2 * (size_t)cp.cout_pad, 0.f);
            for (int i = 0; i < cp.cout_pad; ++i) host[i] = 1.f;
            B200_CHECK(cudaMemcpy(k.ones, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice));
            k.scale = cp.scale; k.shift = cp.shift; k.act = cp.act;
            k.res = a.res; k.ldr = a.ldr; k.res_alpha = a.res_alpha; k.res_beta = a.res_beta;
            k.out = (bf16 *)out.p; k.ldo = out.ld; k.npix = a.npix; k.cout_pad = cp.cout_pad;
            // the tile kernel now produces raw fp32 partial sums: identity epilogue, 32-column fp32 sub-tiles, slab-addressed stores
            a.ksplit = S; a.kb_per = kb_per; a.slab_rows = (int)slab_rows;
            p->flow_ok = false;                                  // (a flow would read the identity epilogue constants)
            a.scale = k.ones; a.shift = k.ones + cp.cout_pad; a.act = ACT_LINEAR; a.res = nullptr;
            a.sub_cols = 32; a.out_f32 = 1;
            a.vtiles = pair_tiles_for_split(a) * S;
            p->grid = 2 * (a.vtiles < 74 ? a.vtiles : 74);
            unsigned long long dims[2] = {(unsigned long long)cp.cout_pad, (unsigned long long)(S * slab_rows)};
            unsigned long long strides[1] = {(unsigned long long)cp.cout_pad * sizeof(float)};
            unsigned box[2] = {32, 128};
            tc_encode_tiled(&p->maps.c, 1, 2, k.ws, dims, strides, box, 128);
        }
    }
    p->flops = 2.0 * (double)a.npix * out.c * K;
    char buf[256];
    snprintf(buf, sizeof buf, "conv_tc mode%d k%d n%d tile %dx%dx%d rows %d m_tiles %d n_tiles %d stages %d smem %zu grid %d%s%s",
             a.mode, block_k, a.block_n, a.TW, a.TH, a.TN, a.a_rows, a.m_tiles, a.n_tiles, stages, p->smem_bytes, p->grid,
             a.pair ? " PAIR(cta_group::2)" : (a.resident_b ? " residentB" : ""), a.res ? " +residual" : "");

    p->desc = buf;
    if (a.im2col) p->desc += " im2colTMA";
    if (a.ring) p->desc += " ringEpilogue(" + std::to_string(a.c_bufs) + ")";
    else if (a.staged) p->desc += " stagedEpilogue";
    if (a.upsample) p->desc += " +upsample2x";
    if (a.local) p->desc += " unshared(local)";
    if (a.ksplit > 1) p->desc += " splitK(" + std::to_string(a.ksplit) + "x" + std::to_string(a.kb_per) + ")";
    if (a.split > 1) p->desc += " tailSplit(" + std::to_string(a.split) + "x" + std::to_string(a.vtiles - a.split_from) + ")";
    return p;
}

// ---------------------------------------------------------------------------------------------------
// flows (kernel in conv_tc_flow.cu)
// ---------------------------------------------------------------------------------------------------
// Flows are planned only on request (env B200_FLOW=1 when the network is parsed).  MEASURED (round 2, YOLOv3-416 batch 64, one
// B200): bit-identical results, tensor-issue busy 73 -> 82 % inside the flow, yet the forward pass is 1-3 % SLOWER than one
// launch per layer (4.78 vs 4.68 ms cold, 5.30 vs 5.17 ms sustained): the chip is power-managed at sub-millisecond scale, the SM
// clock inside the flow reads 1660-1750 MHz against ~1920 in the per-layer kernels, i.e. the idle tails the flow removes are
// what lets the per-layer kernels boost (profiles/r2_flow_*.txt).
bool conv_tc_plan_flow_ok(const ConvTcPlan *p)
{
    const char *on = getenv("B200_FLOW");
    return p && p->flow_ok && on && atoi(on) != 0 && !getenv("B200_DISABLE_TC");
}
const char *conv_tc_flow_desc(ConvTcFlow *f) { return f->desc.c_str(); }
// profiling: the 4 stamps (+ pair | position << 16) per item of the last launch (nullptr / 0 unless the flow was planned under B200_FLOW_TRACE=1) and the
// first item number of every member layer (n + 1 entries)
int conv_tc_flow_trace(ConvTcFlow *f, unsigned long long *out, int max_items, int *item0, int max_layers)
{
    if (!f->fp.trace) return 0;
    const int n = f->fp.total_items < max_items ? f->fp.total_items : max_items;
    B200_CHECK(cudaMemcpy(out, f->fp.trace, (size_t)n * 5 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    for (int k = 0; k < (int)f->item0.size() && k < max_layers; ++k) item0[k] = f->item0[k];
    return n;
}
void conv_tc_flow_read_stats(ConvTcFlow *f, unsigned long long *out3)
{
    B200_CHECK(cudaMemcpy(out3, f->fp.stats, 5 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    B200_CHECK(cudaMemset(f->fp.stats, 0, 3 * sizeof(unsigned long long)));
}
void conv_tc_flow_destroy(ConvTcFlow *f)
{
    if (!f) return;
    if (f->blob) cudaFree(f->blob);
    delete f;
}

// ---- the schedule of a flow: list scheduling of all items on the CTA pairs in simulated time ------------------------------------
// A free pair takes the oldest item whose inputs are complete (or, if none is, the one that completes first).  Item cost
// model: tensor time 0.375 us per 256 x 256 x 64 k-block (measured marginal rate of the pair kernel) against the bytes
// the item moves at 1/74 of the HBM rate, plus 0.5 us.  The model only shapes the lists; correctness never depends on it:
// any order produced here is a valid execution in simulated time, so the kernel (all pairs co-resident) cannot deadlock.
static void flow_schedule(const std::vector<FlowLayerArgs> &L, int ctr, int pairs, std::vector<unsigned> &sched, std::vector<int> &sched_off,
                          double &makespan, double &work)
{
    const int n = (int)L.size();
    int total_items = 0;
    for (const FlowLayerArgs &a : L) total_items += a.items;
    sched.clear();
    sched_off.assign(pairs + 1, 0);
    makespan = 0; work = 0;
    std::vector<double> dur(n);
    for (int k = 0; k < n; ++k) {
        const FlowLayerArgs &a = L[k];
        const double t_mma = a.num_kblocks * (a.block_n / 256.0) * 0.375;
        const double bytes = 256.0 * a.cin_blocks * 64 * 2 + 256.0 * a.block_n * 2 * (1 + a.has_res);
        const double t_mem = bytes / 88.0e3;
        dur[k] = (t_mma > t_mem ? t_mma : t_mem) + 0.5;
        work += dur[k] * a.items;
    }
    std::vector<unsigned short> layer_of(total_items);
    for (int k = 0; k < n; ++k) for (int t = 0; t < L[k].items; ++t) layer_of[L[k].item0 + t] = (unsigned short)k;
    std::vector<double> ctr_done(ctr, 0.0);
    std::vector<int> ctr_left(ctr);
    for (int k = 0; k < n; ++k) for (int mp = 0; mp < L[k].items / L[k].n_tiles; ++mp) ctr_left[L[k].done_off + mp] = L[k].n_tiles;
    std::vector<char> taken(total_items, 0);
    std::vector<double> pair_free(pairs, 0.0);
    std::vector<std::vector<unsigned>> lists(pairs);
    auto ready = [&](int id, double &rt) {
        const int k = layer_of[id];
        const FlowLayerArgs &a = L[k];
        const int mp = (id - a.item0) / a.n_tiles;
        rt = 0;
        if (a.dep >= 0) {
            int jlo, jhi;
            flow_dep_range(a, mp, jlo, jhi);
            for (int j = jlo; j <= jhi; ++j) {
                if (ctr_left[a.dep_off + j] > 0) return false;
                if (ctr_done[a.dep_off + j] > rt) rt = ctr_done[a.dep_off + j];
            }
        }
        if (a.res_dep >= 0) {
            if (ctr_left[a.res_off + mp] > 0) return false;
            if (ctr_done[a.res_off + mp] > rt) rt = ctr_done[a.res_off + mp];
        }
        return true;
    };
    int cursor = 0;
    const int window = 2048;
    for (int done = 0; done < total_items; ++done) {
        int p = 0;
        for (int q = 1; q < pairs; ++q) if (pair_free[q] < pair_free[p]) p = q;
        const double T = pair_free[p];
        while (cursor < total_items && taken[cursor]) ++cursor;
        int best = -1; double best_rt = 1e300;
        for (int id = cursor; id < total_items && id < cursor + window; ++id) {
            if (taken[id]) continue;
            double rt;
            if (!ready(id, rt)) continue;
            if (rt <= T) { best = id; best_rt = rt; break; }
            if (rt < best_rt) { best = id; best_rt = rt; }
        }
        if (best < 0) { fprintf(stderr, "b200-darknet: internal error: flow schedule has no runnable item\n"); abort(); }
        const int k = layer_of[best];
        const FlowLayerArgs &a = L[k];
        const double start = T > best_rt ? T : best_rt, fin = start + dur[k];
        taken[best] = 1;
        lists[p].push_back(((unsigned)k << 24) | (unsigned)(best - a.item0));
        pair_free[p] = fin;
        const int c = a.done_off + (best - a.item0) / a.n_tiles;
        if (fin > ctr_done[c]) ctr_done[c] = fin;
        --ctr_left[c];
        if (fin > makespan) makespan = fin;
    }
    for (int q = 0; q < pairs; ++q) {
        sched_off[q] = (int)sched.size();
        sched.insert(sched.end(), lists[q].begin(), lists[q].end());
    }
    sched_off[pairs] = (int)sched.size();
}

// Replays per-pair item lists WITHOUT the cost model: a pair may run its next item as soon as the counters that item waits on
// are complete.  Returns -1 when every item ran exactly once, else the first item that could never run (a deadlock or a hole).
static int flow_schedule_check(const std::vector<FlowLayerArgs> &L, int ctr, int pairs, const std::vector<unsigned> &sched, const std::vector<int> &sched_off)
{
    std::vector<int> left(ctr, 0), pos(pairs);
    int total = 0;
    for (const FlowLayerArgs &a : L) { for (int mp = 0; mp < a.items / a.n_tiles; ++mp) left[a.done_off + mp] = a.n_tiles; total += a.items; }
    if ((int)sched.size() != total) return 0;
    std::vector<char> ran(total, 0);
    for (int q = 0; q < pairs; ++q) pos[q] = sched_off[q];
    int done = 0;
    for (bool progress = true; progress;) {
        progress = false;
        for (int q = 0; q < pairs; ++q) {
            while (pos[q] < sched_off[q + 1]) {
                const unsigned ent = sched[pos[q]];
                const FlowLayerArgs &a = L[ent >> 24];
                const int t = (int)(ent & 0xffffffu), mp = t / a.n_tiles;
                bool ok = t < a.items && !ran[a.item0 + t];
                if (ok && a.dep >= 0) {
                    int jlo, jhi;
                    flow_dep_range(a, mp, jlo, jhi);
                    for (int j = jlo; j <= jhi && ok; ++j) ok = left[a.dep_off + j] == 0;
                }
                if (ok && a.res_dep >= 0) ok = left[a.res_off + mp] == 0;
                if (!ok) break;
                ran[a.item0 + t] = 1; --left[a.done_off + mp]; ++pos[q]; ++done; progress = true;
            }
        }
    }
    if (done == total) return -1;
    for (int q = 0; q < pairs; ++q) if (pos[q] < sched_off[q + 1]) { const unsigned ent = sched[pos[q]]; return L[ent >> 24].item0 + (int)(ent & 0xffffffu); }
    return 0;
}

ConvTcFlow *conv_tc_flow_create(const ConvTcFlowMember *members, int n)
{
    if (n < 2 || n > kFlowMaxLayers) return nullptr;
    std::vector<FlowLayerArgs> L(n);
    std::vector<CUtensorMap> maps(4 * (size_t)n);
    memset(maps.data(), 0, maps.size() * sizeof(CUtensorMap));
    int item = 0, ctr = 0;
    double flops = 0;
    for (int k = 0; k < n; ++k) {
        const ConvTcPlan *p = members[k].plan;
        if (!p || !p->flow_ok) return nullptr;
        const ConvTcArgs &pa = p->args;
        FlowLayerArgs &a = L[k];
        memset(&a, 0, sizeof a);
        a.block_n = pa.cout_pad < 256 ? pa.cout_pad : 256;
        a.n_tiles = pa.cout_pad / a.block_n;
        a.m_tiles = (int)((pa.npix + 127) / 128);
        const int m_pairs = (a.m_tiles + 1) / 2;
        a.item0 = item; a.items = m_pairs * a.n_tiles; item += a.items;
        a.num_kblocks = pa.num_kblocks; a.cin_blocks = pa.cin_blocks;
        a.im2col = pa.im2col; a.size = pa.size; a.stride = pa.stride; a.pad = pa.pad; a.OW = pa.OW; a.OH = pa.OH; a.npix = (int)pa.npix;
        a.in_W = p->in.w; a.in_H = p->in.h;
        a.act = pa.act; a.has_res = p->res.p ? 1 : 0; a.res_alpha = pa.res_alpha; a.res_beta = pa.res_beta;
        a.scale = pa.scale; a.shift = pa.shift;
        a.done_off = ctr; ctr += m_pairs;
        a.dep = members[k].dep; a.res_dep = a.has_res ? members[k].res_dep : -1;
        if (a.dep >= k || a.res_dep >= k) return nullptr;                  // producers come first: the item order is the dependency order
        if (a.dep >= 0) {
            const FlowLayerArgs &d = L[a.dep];
            if ((long long)d.npix != (long long)p->in.n * p->in.h * p->in.w) return nullptr;
            a.dep_off = d.done_off; a.dep_unit = 2 * d.n_tiles;
        }
        if (a.res_dep >= 0) {
            const FlowLayerArgs &d = L[a.res_dep];
            if (d.npix != a.npix) return nullptr;
            a.res_off = d.done_off; a.res_unit = 2 * d.n_tiles;
        }
        // A: the plan's own view (dense 2-D rows or the im2col-mode map); B: half a filter tile per CTA of the pair
        maps[4 * k + 0] = p->maps.a[0];
        {
            cuuint64_t dims[2] = {(cuuint64_t)p->K, (cuuint64_t)pa.cout_pad};
            cuuint64_t strides[1] = {(cuuint64_t)p->K * 2};
            cuuint32_t box[2] = {64, (cuuint32_t)(a.block_n / 2)};
            encode(&maps[4 * k + 1], (void *)p->w, 2, dims, strides, box, 64);
        }
        ConvTcArgs dense; memset(&dense, 0, sizeof dense); dense.mode = 0; dense.npix = pa.npix;
        encode_tile_view(&maps[4 * k + 2], p->out, pa.cout_pad, dense, 64);
        if (a.has_res) encode_tile_view(&maps[4 * k + 3], p->res, pa.cout_pad, dense, 64);
        flops += p->flops;
    }
    const int pairs = 74;
    std::vector<unsigned> sched;
    std::vector<int> sched_off;
    double makespan = 0, work = 0;
    flow_schedule(L, ctr, pairs, sched, sched_off, makespan, work);
    for (int k = 0; k < n; ++k) if (L[k].items >= (1 << 24)) return nullptr;

    ConvTcFlow *f = new ConvTcFlow();
    const size_t maps_bytes = maps.size() * sizeof(CUtensorMap), layers_bytes = (size_t)n * sizeof(FlowLayerArgs);
    const size_t layers_off = (maps_bytes + 255) / 256 * 256, done_off = (layers_off + layers_bytes + 255) / 256 * 256;
    const size_t sched_at = (done_off + (size_t)ctr * sizeof(unsigned) + 255) / 256 * 256;
    const size_t off_at = (sched_at + sched.size() * sizeof(unsigned) + 255) / 256 * 256;
    const size_t stats_at = (off_at + sched_off.size() * sizeof(int) + 255) / 256 * 256;
    const bool tracing = getenv("B200_FLOW_TRACE") != nullptr;
    const size_t trace_at = stats_at + 8 * sizeof(unsigned long long);
    const size_t total = trace_at + (tracing ? (size_t)item * 5 * sizeof(unsigned long long) : 0);
    B200_CHECK(cudaMalloc(&f->blob, total));
    B200_CHECK(cudaMemset(f->blob, 0, total));
    B200_CHECK(cudaMemcpy(f->blob, maps.data(), maps_bytes, cudaMemcpyHostToDevice));
    B200_CHECK(cudaMemcpy((unsigned char *)f->blob + layers_off, L.data(), layers_bytes, cudaMemcpyHostToDevice));
    B200_CHECK(cudaMemcpy((unsigned char *)f->blob + sched_at, sched.data(), sched.size() * sizeof(unsigned), cudaMemcpyHostToDevice));
    B200_CHECK(cudaMemcpy((unsigned char *)f->blob + off_at, sched_off.data(), sched_off.size() * sizeof(int), cudaMemcpyHostToDevice));
    f->fp.maps = (const CUtensorMap *)f->blob;
    f->fp.layers = (const FlowLayerArgs *)((unsigned char *)f->blob + layers_off);
    f->fp.done = (unsigned *)((unsigned char *)f->blob + done_off);
    f->fp.sched = (const unsigned *)((unsigned char *)f->blob + sched_at);
    f->fp.sched_off = (const int *)((unsigned char *)f->blob + off_at);
    f->fp.stats = (unsigned long long *)((unsigned char *)f->blob + stats_at);
    f->fp.trace = tracing ? (unsigned long long *)((unsigned char *)f->blob + trace_at) : nullptr;
    f->item0.resize(n + 1);
    for (int k = 0; k < n; ++k) f->item0[k] = L[k].item0;
    f->item0[n] = item;
    f->fp.nl = n; f->fp.total_items = item; f->fp.epoch = 0;
    f->smem_bytes = (size_t)kFlowStages * 32768 + (size_t)kFlowRing * 16384 + 512 + (size_t)kFlowMaxLayers * sizeof(FlowLayerArgs) + 1024;
    f->flops = flops;
    char buf[200];
    snprintf(buf, sizeof buf, "conv_tc FLOW %d layers %d pair tiles %d counters stages %d smem %zu grid 148 model %.0f us (%.0f us of work per pair)",
             n, item, ctr, kFlowStages, f->smem_bytes, makespan, work / pairs);
    f->desc = buf;
    return f;
}

// Host-only self-test of the flow scheduler (no device needed; tests/test_flow_schedule.py): a chain of `n` layers over a
// batch x hw x hw map — spec[4k .. 4k+3] = {filter size (1 or 3), stride, input channels, filters}, 3x3 layers with stride 1 from
// the third layer on take the layer two back as their shortcut operand — is cut into items, scheduled on `pairs` pairs and the
// lists replayed without the cost model.  Returns -1 when the schedule is complete and deadlock-free, else the stuck item.
extern "C" int b200_flow_schedule_selftest(int n, const int *spec, int batch, int hw, int pairs, double *makespan_us, double *work_us)
{
    std::vector<FlowLayerArgs> L(n);
    int item = 0, ctr = 0, h = hw;
    for (int k = 0; k < n; ++k) {
        const int size = spec[4 * k], stride = spec[4 * k + 1], cin = spec[4 * k + 2], cout = spec[4 * k + 3];
        FlowLayerArgs &a = L[k];
        memset(&a, 0, sizeof a);
        const int pad = size / 2, oh = (h + 2 * pad - size) / stride + 1;
        a.block_n = cout < 256 ? cout : 256; a.n_tiles = cout / a.block_n;
        a.npix = batch * oh * oh; a.m_tiles = (a.npix + 127) / 128;
        const int m_pairs = (a.m_tiles + 1) / 2;
        a.item0 = item; a.items = m_pairs * a.n_tiles; item += a.items;
        a.cin_blocks = cin / 64; a.num_kblocks = size * size * a.cin_blocks;
        a.im2col = size > 1; a.size = size; a.stride = stride; a.pad = pad; a.OW = a.OH = oh; a.in_W = a.in_H = h;
        a.done_off = ctr; ctr += m_pairs;
        a.dep = k - 1; a.res_dep = -1;
        if (k >= 2 && size == 3 && stride == 1 && L[k - 2].npix == a.npix && L[k - 2].n_tiles * L[k - 2].block_n == cout) { a.res_dep = k - 2; a.has_res = 1; }
        if (a.dep >= 0) { a.dep_off = L[a.dep].done_off; a.dep_unit = 2 * L[a.dep].n_tiles; }
        if (a.res_dep >= 0) { a.res_off = L[a.res_dep].done_off; a.res_unit = 2 * L[a.res_dep].n_tiles; }
        h = oh;
    }
    std::vector<unsigned> sched; std::vector<int> off;
    double mk = 0, wk = 0;
    flow_schedule(L, ctr, pairs, sched, off, mk, wk);
    if (makespan_us) *makespan_us = mk;
    if (work_us) *work_us = wk / pairs;
    return flow_schedule_check(L, ctr, pairs, sched, off);
}

// Host-only check of flow_dep_range (the counters a consuming tile waits on) against the brute-force footprint of the tile:
// for every 256-pixel tile of a size x size / stride convolution over a batch x in_h x in_w input, every input pixel that any tap
// of any output pixel of the tile reads must belong to a counter inside [jlo, jhi].  Returns -1 when that holds for all tiles,
// else the first tile that would read data it did not wait for.
extern "C" int b200_flow_dep_selftest(int batch, int in_h, int in_w, int size, int stride)
{
    FlowLayerArgs a;
    memset(&a, 0, sizeof a);
    a.size = size; a.stride = stride; a.pad = size / 2; a.im2col = size > 1;
    a.in_H = in_h; a.in_W = in_w;
    a.OH = (in_h + 2 * a.pad - size) / stride + 1; a.OW = (in_w + 2 * a.pad - size) / stride + 1;
    a.npix = batch * a.OH * a.OW;
    const int m_pairs = (a.npix + 255) / 256;
    for (int mp = 0; mp < m_pairs; ++mp) {
        int jlo, jhi;
        flow_dep_range(a, mp, jlo, jhi);
        for (int p = mp * 256; p < (mp + 1) * 256 && p < a.npix; ++p) {
            const int n = p / (a.OH * a.OW), r = p % (a.OH * a.OW), oy = r / a.OW, ox = r % a.OW;
            for (int ky = 0; ky < size; ++ky)
                for (int kx = 0; kx < size; ++kx) {
                    const int iy = oy * stride - a.pad + ky, ix = ox * stride - a.pad + kx;
                    if (iy < 0 || iy >= in_h || ix < 0 || ix >= in_w) continue;            // zero fill: nothing is read
                    const int q = ((n * in_h + iy) * in_w + ix) >> 8;
                    if (size == 1 ? q != mp || jlo != mp || jhi != mp : (q < jlo || q > jhi)) return mp;
                }
        }
    }
    return -1;
}
