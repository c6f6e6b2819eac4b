// conv_tc_plan.cu — host side of the tcgen05 convolution family (overview in conv_tc.cu): TMA tensor maps, the per-layer planner
// (tile shape, kernel variant, pipeline depth, epilogue, tail splitting) and the launch dispatch.
#include "conv_tc_plan.h"

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
    // MEASURED (YOLOv3-416 b64): the ring epilogue wins on every stageable layer (1x1 layers -10..-20 %, fused shortcuts
    // -8 %) except the stride-2 3x3 layers without a residual, which lose the pipeline stage the ring's slots cost (+3 %).
    const bool ring_pays = residual || !(cp.size == 3 && cp.stride == 2 && a.block_n == 256);
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
    if (a.split > 1) p->desc += " tailSplit(" + std::to_string(a.split) + "x" + std::to_string(a.vtiles - a.split_from) + ")";
    return p;
}
