/* writers.c — result files of the detector validation drivers, written from the batched path's records.
 *
 * The reference keeps these as statics of examples/detector.c: print_cocos (:165-188, COCO result json lines through the
 * 80 -> 91 category map and the image id parsed from the file name, :157-163), print_detector_detections (:190-209, VOC
 * "comp4" files, one per class, boxes shifted by +1 pixel and clamped to [1, w] x [1, h]) and print_imagenet_detections
 * (:211-232).  They walk a detection array box by box, class by class; here the input is the record list of
 * b200_detect_batch (pixel coordinates: relative = 0), sorted by (image, box, class) so the files are deterministic.
 * Corner arithmetic is done in double and narrowed to float exactly as the C expressions of the reference do. */
#include "darknet.h"
#include "b200_engine.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static const int coco_category[80] = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 27, 28, 31, 32, 33, 34,
                                      35, 36, 37, 38, 39, 40, 41, 42, 43, 44, 46, 47, 48, 49, 50, 51, 52, 53, 54, 55, 56, 57, 58, 59, 60, 61, 62, 63,
                                      64, 65, 67, 70, 72, 73, 74, 75, 76, 77, 78, 79, 80, 81, 82, 84, 85, 86, 87, 88, 89, 90};

int b200_coco_image_id(const char *filename)             /* "…/COCO_val2014_000000000042.jpg" -> 42 */
{
    const char *p = strrchr(filename, '/');
    const char *c = strrchr(filename, '_');
    if (c) p = c;
    return atoi(p ? p + 1 : filename);
}

static int by_image_box_class(const void *a, const void *b)
{
    const b200_det *x = a, *y = b;
    if (x->image != y->image) return x->image < y->image ? -1 : 1;
    if (x->box_id != y->box_id) return x->box_id < y->box_id ? -1 : 1;
    return (x->cls > y->cls) - (x->cls < y->cls);
}

void b200_sort_records(b200_det *rec, int n) { qsort(rec, (size_t)n, sizeof *rec, by_image_box_class); }

typedef struct { float xmin, ymin, xmax, ymax; } corners;

static corners clamp_corners(box b, float shift, float lo, int w, int h)
{
    corners c;
    c.xmin = b.x - b.w / 2. + shift; c.xmax = b.x + b.w / 2. + shift;
    c.ymin = b.y - b.h / 2. + shift; c.ymax = b.y + b.h / 2. + shift;
    if (c.xmin < lo) c.xmin = lo;
    if (c.ymin < lo) c.ymin = lo;
    if (c.xmax > w) c.xmax = w;
    if (c.ymax > h) c.ymax = h;
    return c;
}

/* print_cocos for every record: image i is image_paths[i] with original size widths[i] x heights[i] */
int b200_write_coco(FILE *fp, const b200_det *rec, int n, const char *const *image_paths, const int *widths, const int *heights)
{
    for (int r = 0; r < n; ++r) {
        const b200_det *d = &rec[r];
        if (!d->prob || d->cls < 0 || d->cls >= 80) continue;
        corners c = clamp_corners(d->bbox, 0.f, 0.f, widths[d->image], heights[d->image]);
        float bx = c.xmin, by = c.ymin, bw = c.xmax - c.xmin, bh = c.ymax - c.ymin;
        fprintf(fp, "{\"image_id\":%d, \"category_id\":%d, \"bbox\":[%f, %f, %f, %f], \"score\":%f},\n",
                b200_coco_image_id(image_paths[d->image]), coco_category[d->cls], bx, by, bw, bh, d->prob);
    }
    return ferror(fp) ? -1 : 0;
}

/* print_detector_detections: fps[class] are the per-class files, ids[i] the image identifiers */
int b200_write_voc(FILE **fps, const b200_det *rec, int n, const char *const *ids, const int *widths, const int *heights)
{
    for (int r = 0; r < n; ++r) {
        const b200_det *d = &rec[r];
        if (!d->prob) continue;
        corners c = clamp_corners(d->bbox, 1.f, 1.f, widths[d->image], heights[d->image]);
        fprintf(fps[d->cls], "%s %f %f %f %f %f\n", ids[d->image], d->prob, c.xmin, c.ymin, c.xmax, c.ymax);
    }
    return 0;
}

/* print_imagenet_detections: numeric image ids, classes numbered from 1 */
int b200_write_imagenet(FILE *fp, const b200_det *rec, int n, const int *image_ids, const int *widths, const int *heights)
{
    for (int r = 0; r < n; ++r) {
        const b200_det *d = &rec[r];
        if (!d->prob) continue;
        corners c = clamp_corners(d->bbox, 0.f, 0.f, widths[d->image], heights[d->image]);
        fprintf(fp, "%d %d %f %f %f %f %f\n", image_ids[d->image], d->cls + 1, d->prob, c.xmin, c.ymin, c.xmax, c.ymax);
    }
    return ferror(fp) ? -1 : 0;
}

/* path-based conveniences for bindings that cannot pass a FILE* (ctypes): append to the named file(s) */
int b200_append_coco(const char *path, b200_det *rec, int n, const char *const *image_paths, const int *widths, const int *heights)
{
    FILE *fp = fopen(path, "a");
    if (!fp) return -1;
    b200_sort_records(rec, n);
    int rc = b200_write_coco(fp, rec, n, image_paths, widths, heights);
    fclose(fp);
    return rc;
}

int b200_append_voc(const char *prefix, const char *const *names, int classes, b200_det *rec, int n, const char *const *ids,
                    const int *widths, const int *heights)
{
    FILE **fps = calloc((size_t)classes, sizeof(FILE *));
    if (!fps) return -1;
    int rc = 0;
    for (int j = 0; j < classes; ++j) {                  /* "<prefix><class name>.txt", as validate_detector names them (:416-419) */
        char buff[1024];
        snprintf(buff, sizeof buff, "%s%s.txt", prefix, names[j]);
        fps[j] = fopen(buff, "a");
        if (!fps[j]) rc = -1;
    }
    if (rc == 0) {
        b200_sort_records(rec, n);
        rc = b200_write_voc(fps, rec, n, ids, widths, heights);
    }
    for (int j = 0; j < classes; ++j) if (fps[j]) fclose(fps[j]);
    free(fps);
    return rc;
}

int b200_append_imagenet(const char *path, b200_det *rec, int n, const int *image_ids, const int *widths, const int *heights)
{
    FILE *fp = fopen(path, "a");
    if (!fp) return -1;
    b200_sort_records(rec, n);
    int rc = b200_write_imagenet(fp, rec, n, image_ids, widths, heights);
    fclose(fp);
    return rc;
}
