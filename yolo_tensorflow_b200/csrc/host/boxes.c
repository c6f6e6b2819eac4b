/*
 * boxes.c — detection extraction and NMS entry points of the darknet C API (host C).
 *
 * get_network_boxes / make_network_boxes / fill_network_boxes / num_detections / free_detections
 * (reference network.c:510-577) and do_nms_sort / do_nms_obj (box.c:21-89).  The arithmetic — anchor
 * decode, thresholding, compaction, IoU, suppression — runs on the device (dev/decode.cu, dev/nms.cu);
 * this file only owns the `detection` records the API hands to the caller: one calloc'd array plus one
 * calloc'd prob[classes] per record, freed by free_detections(dets, n), exactly the reference ownership.
 */
#include "darknet.h"
#include "b200_engine.h"

/* internal C ABI of the engine (dev/engine.cu) */
int b200_engine_count_boxes(b200_engine *e, network *net, int image, float thresh);
int b200_engine_decode_image(b200_engine *e, network *net, int image, int w, int h, float thresh, int relative,
                             const float **box, const float **obj, const float **prob, const int **id, float hier, const int *map);
void b200_engine_hierarchy(b200_engine *e, network *net, int image);
int b200_engine_has_tree(b200_engine *e);
int b200_engine_classes(b200_engine *e);
void b200_engine_push_heads(b200_engine *e, network *net, int items);
void b200_engine_avg_flipped(b200_engine *e, network *net);

static int last_layer_classes(network *net) { return net->layers[net->n - 1].classes; }

static int count_for_image(network *net, int image, float thresh)
{
    return b200_engine_count_boxes(b200_engine_of(net), net, image, thresh);
}

int num_detections(network *net, float thresh) { return count_for_image(net, 0, thresh); }

static detection *alloc_dets(network *net, int nboxes)
{
    layer l = net->layers[net->n - 1];
    detection *dets = calloc(nboxes > 0 ? nboxes : 1, sizeof(detection));
    for (int i = 0; i < nboxes; ++i) {
        dets[i].prob = calloc(l.classes > 0 ? l.classes : 1, sizeof(float));
        if (l.coords > 4) dets[i].mask = calloc(l.coords - 4, sizeof(float));
    }
    return dets;
}

/* l.batch == 2 on a [yolo] / [region] head: get_*_detections averages item 0 with the mirrored item 1 (`detector valid2`) */
static int heads_flip(const network *net)
{
    for (int i = 0; i < net->n; ++i) {
        const layer *l = &net->layers[i];
        if ((l->type == YOLO || l->type == REGION) && l->batch == 2) return 1;
    }
    return 0;
}

/* the reference API reads the heads' HOST buffers; callers may have rewritten them since the last predict (demo.c:54-83) */
static int push_heads(network *net)
{
    int flip = heads_flip(net);
    b200_engine_push_heads(b200_engine_of(net), net, flip ? 2 : 1);
    return flip;
}

detection *make_network_boxes(network *net, float thresh, int *num)
{
    push_heads(net);
    int nboxes = num_detections(net, thresh);
    if (num) *num = nboxes;
    return alloc_dets(net, nboxes);
}

static int fill_for_image(network *net, int image, int w, int h, float thresh, float hier, int *map, int relative, detection *dets, int room)
{
    const float *box, *obj, *prob;
    const int *id;
    int classes = last_layer_classes(net);
    b200_engine *e = b200_engine_of(net);
    /* YOLO9000: get_region_detections turns the conditional class probabilities into absolute ones in place before it scores
     * the boxes (hierarchy_predictions, region_layer.c:412-414) */
    if (b200_engine_has_tree(e)) b200_engine_hierarchy(e, net, image);
    int n = b200_engine_decode_image(e, net, image, w, h, thresh, relative, &box, &obj, &prob, &id, hier, map);
    if (room >= 0 && n > room) n = room;
    for (int i = 0; i < n; ++i) {
        dets[i].bbox.x = box[4 * i + 0]; dets[i].bbox.y = box[4 * i + 1];
        dets[i].bbox.w = box[4 * i + 2]; dets[i].bbox.h = box[4 * i + 3];
        dets[i].objectness = obj[i];
        dets[i].classes = classes;
        memcpy(dets[i].prob, prob + (size_t)i * classes, (size_t)classes * sizeof(float));
    }
    return n;
}

void fill_network_boxes(network *net, int w, int h, float thresh, float hier, int *map, int relative, detection *dets)
{
    if (push_heads(net)) b200_engine_avg_flipped(b200_engine_of(net), net);
    fill_for_image(net, 0, w, h, thresh, hier, map, relative, dets, -1);
}

detection *get_network_boxes_batch(network *net, int b, int w, int h, float thresh, float hier, int *map, int relative, int *num)
{
    int nboxes = count_for_image(net, b, thresh);
    detection *dets = alloc_dets(net, nboxes);
    int got = fill_for_image(net, b, w, h, thresh, hier, map, relative, dets, nboxes);
    (void)got;
    if (num) *num = nboxes;
    return dets;
}

/* network.c:559-567.  Like the reference this reads the heads' HOST buffers (a caller may have rewritten l.output since the
 * last predict, demo.c:54-83), counts before and fills after the batch == 2 flip-average (make_network_boxes runs first,
 * network.c:563-564; get_yolo_detections / get_region_detections average inside the fill, yolo_layer.c:320,
 * region_layer.c:368-390): *num is the count of the un-averaged item 0 and records the averaged output no longer fills stay
 * zero, exactly what the calloc'd reference array holds.  (Where the averaged output has MORE boxes over the threshold than
 * were counted, the reference writes past its allocation; here the surplus is dropped.) */
detection *get_network_boxes(network *net, int w, int h, float thresh, float hier, int *map, int relative, int *num)
{
    int flip = push_heads(net);
    if (!flip) return get_network_boxes_batch(net, 0, w, h, thresh, hier, map, relative, num);
    int nboxes = count_for_image(net, 0, thresh);
    detection *dets = alloc_dets(net, nboxes);
    b200_engine_avg_flipped(b200_engine_of(net), net);
    fill_for_image(net, 0, w, h, thresh, hier, map, relative, dets, nboxes);
    if (num) *num = nboxes;
    return dets;
}

void free_detections(detection *dets, int n)
{
    for (int i = 0; i < n; ++i) {
        free(dets[i].prob);
        if (dets[i].mask) free(dets[i].mask);
    }
    free(dets);
}

/* same two-pointer partition as box.c:60-70: detections with objectness 0 go to the tail */
static int partition_live(detection *dets, int total)
{
    int k = total - 1;
    for (int i = 0; i <= k; ++i) {
        if (dets[i].objectness == 0) {
            detection swap = dets[i];
            dets[i] = dets[k];
            dets[k] = swap;
            --k;
            --i;
        }
    }
    return k + 1;
}

/* stable descending order as an index permutation (bottom-up merge sort).  `key` has `stride` floats per row; rows are
 * compared on columns ncols-1, ncols-2, ..., 0 in that order (ncols = 1: a plain sort by one value). */
static int row_after(const float *key, int stride, int ncols, int a, int b)      /* does row b sort strictly before row a? */
{
    const float *ka = key + (size_t)a * stride, *kb = key + (size_t)b * stride;
    for (int c = ncols - 1; c >= 0; --c) {
        if (kb[c] > ka[c]) return 1;
        if (kb[c] < ka[c]) return 0;
    }
    return 0;
}

static int *stable_order_desc(const float *key, int stride, int ncols, int n)
{
    int *perm = malloc((size_t)n * sizeof(int)), *tmp = malloc((size_t)n * sizeof(int));
    for (int i = 0; i < n; ++i) perm[i] = i;
    for (int width = 1; width < n; width *= 2) {
        for (int lo = 0; lo < n; lo += 2 * width) {
            int mid = lo + width < n ? lo + width : n, hi = lo + 2 * width < n ? lo + 2 * width : n;
            int a = lo, b = mid, o = lo;
            while (a < mid && b < hi) tmp[o++] = row_after(key, stride, ncols, perm[a], perm[b]) ? perm[b++] : perm[a++];
            while (a < mid) tmp[o++] = perm[a++];
            while (b < hi) tmp[o++] = perm[b++];
        }
        memcpy(perm, tmp, (size_t)n * sizeof(int));
    }
    free(tmp);
    return perm;
}

static void permute_dets(detection *dets, const int *perm, int n)
{
    detection *sorted = malloc((size_t)n * sizeof(detection));
    for (int i = 0; i < n; ++i) sorted[i] = dets[perm[i]];
    memcpy(dets, sorted, (size_t)n * sizeof(detection));
    free(sorted);
}

void do_nms_sort(detection *dets, int total, int classes, float thresh)
{
    total = partition_live(dets, total);
    if (total <= 0 || classes <= 0) return;
    float *boxes = malloc((size_t)total * 4 * sizeof(float));
    float *probs = malloc((size_t)total * classes * sizeof(float));
    float *key = malloc((size_t)total * classes * sizeof(float));
    for (int i = 0; i < total; ++i) {
        boxes[4 * i + 0] = dets[i].bbox.x; boxes[4 * i + 1] = dets[i].bbox.y;
        boxes[4 * i + 2] = dets[i].bbox.w; boxes[4 * i + 3] = dets[i].bbox.h;
        memcpy(probs + (size_t)i * classes, dets[i].prob, (size_t)classes * sizeof(float));
    }
    memcpy(key, probs, (size_t)total * classes * sizeof(float));
    b200_nms_sort_arrays(boxes, probs, total, classes, thresh);      /* device: all classes in parallel */
    for (int i = 0; i < total; ++i) {
        memcpy(dets[i].prob, probs + (size_t)i * classes, (size_t)classes * sizeof(float));
        dets[i].sort_class = classes - 1;
    }
    /* observable side effect of the reference: the array is re-sorted once per class (box.c:72-77), each time by that class's
     * score as it is BEFORE that class's suppression, and glibc's qsort is a merge sort (stable) — so the array ends up ordered
     * by the last class's score, ties by the class before it, and so on down to class 0, then by the partitioned input order
     * (tests/test_oracle.py pins this on the reference build).  draw_detections prints in that order.  prob/mask pointers
     * travel with their structs. */
    int *perm = stable_order_desc(key, classes, classes, total);
    permute_dets(dets, perm, total);
    free(perm); free(key); free(boxes); free(probs);
}

void do_nms_obj(detection *dets, int total, int classes, float thresh)
{
    total = partition_live(dets, total);
    if (total <= 0) return;
    /* reference order after the call: sorted by objectness descending (box.c:39) */
    {
        float *key = malloc((size_t)total * sizeof(float));
        for (int i = 0; i < total; ++i) { dets[i].sort_class = -1; key[i] = dets[i].objectness; }
        int *perm = stable_order_desc(key, 1, 1, total);
        permute_dets(dets, perm, total);
        free(perm); free(key);
    }
    float *boxes = malloc((size_t)total * 4 * sizeof(float));
    float *obj = malloc((size_t)total * sizeof(float));
    unsigned char *supp = calloc(total, 1);
    for (int i = 0; i < total; ++i) {
        boxes[4 * i + 0] = dets[i].bbox.x; boxes[4 * i + 1] = dets[i].bbox.y;
        boxes[4 * i + 2] = dets[i].bbox.w; boxes[4 * i + 3] = dets[i].bbox.h;
        obj[i] = dets[i].objectness;
    }
    b200_nms_obj_arrays(boxes, obj, total, thresh, supp);
    for (int i = 0; i < total; ++i) {
        if (!supp[i]) continue;
        dets[i].objectness = 0;
        for (int k = 0; k < classes; ++k) dets[i].prob[k] = 0;
    }
    free(boxes); free(obj); free(supp);
}

float box_iou(box a, box b)
{
    /* box.c:152-182; host helper for drivers, not used by the device path */
    float l = (a.x - a.w / 2 > b.x - b.w / 2) ? a.x - a.w / 2 : b.x - b.w / 2;
    float r = (a.x + a.w / 2 < b.x + b.w / 2) ? a.x + a.w / 2 : b.x + b.w / 2;
    float t = (a.y - a.h / 2 > b.y - b.h / 2) ? a.y - a.h / 2 : b.y - b.h / 2;
    float d = (a.y + a.h / 2 < b.y + b.h / 2) ? a.y + a.h / 2 : b.y + b.h / 2;
    float w = r - l, h = d - t;
    float inter = (w < 0 || h < 0) ? 0 : w * h;
    return inter / (a.w * a.h + b.w * b.h - inter);
}
