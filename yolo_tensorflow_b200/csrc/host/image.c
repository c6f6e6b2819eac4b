/*
 * image.c — the few host image helpers a detection driver calls right before network_predict
 * (plain host C re-statements; the device-side preprocessing is a later row).
 * Semantics follow image.c:960-979 (letterbox_image: aspect-preserving resize, 0.5 grey fill, centred embed)
 * and image.c:1347-1390 (resize_image: two-pass separable bilinear with (src-1)/(dst-1) scale).
 * Images are fp32 planar CHW in [0,1].
 */
#include "darknet.h"

image make_image(int w, int h, int c)
{
    image m = { w, h, c, calloc((size_t)w * h * c, sizeof(float)) };
    return m;
}

void free_image(image m) { free(m.data); }

static inline float px(const image *m, int x, int y, int k) { return m->data[((size_t)k * m->h + y) * m->w + x]; }

image resize_image(image im, int w, int h)
{
    image out = make_image(w, h, im.c);
    image rows = make_image(w, im.h, im.c);          /* horizontally resampled intermediate */
    const float xs = (float)(im.w - 1) / (w - 1), ys = (float)(im.h - 1) / (h - 1);
    for (int k = 0; k < im.c; ++k)
        for (int y = 0; y < im.h; ++y)
            for (int x = 0; x < w; ++x) {
                float v;
                if (x == w - 1 || im.w == 1) v = px(&im, im.w - 1, y, k);
                else {
                    float sx = x * xs;
                    int ix = (int)sx;
                    float dx = sx - ix;
                    v = (1 - dx) * px(&im, ix, y, k) + dx * px(&im, ix + 1, y, k);
                }
                rows.data[((size_t)k * rows.h + y) * w + x] = v;
            }
    for (int k = 0; k < im.c; ++k)
        for (int y = 0; y < h; ++y) {
            float sy = y * ys;
            int iy = (int)sy;
            float dy = sy - iy;
            float *dst = out.data + ((size_t)k * h + y) * w;
            for (int x = 0; x < w; ++x) dst[x] = (1 - dy) * px(&rows, x, iy, k);
            if (y == h - 1 || im.h == 1) continue;
            for (int x = 0; x < w; ++x) dst[x] += dy * px(&rows, x, iy + 1, k);
        }
    free_image(rows);
    return out;
}

image letterbox_image(image im, int w, int h)
{
    int nw, nh;
    if (((float)w / im.w) < ((float)h / im.h)) { nw = w; nh = (im.h * w) / im.w; }
    else { nh = h; nw = (im.w * h) / im.h; }
    image scaled = resize_image(im, nw, nh);
    image boxed = make_image(w, h, im.c);
    for (size_t i = 0; i < (size_t)w * h * im.c; ++i) boxed.data[i] = .5f;
    const int ox = (w - nw) / 2, oy = (h - nh) / 2;
    for (int k = 0; k < im.c; ++k)
        for (int y = 0; y < nh; ++y)
            memcpy(boxed.data + ((size_t)k * h + oy + y) * w + ox, scaled.data + ((size_t)k * nh + y) * nw, (size_t)nw * sizeof(float));
    free_image(scaled);
    return boxed;
}
