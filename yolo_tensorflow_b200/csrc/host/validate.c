/* validate.c — the detector validation loop as a batched driver.
 *
 * The reference's validate_detector (examples/detector.c:364-487) walks the image list one image at a time: four loader
 * threads letterbox on the CPU, then network_predict (batch 1), get_network_boxes(net, im.w, im.h, .005, .5, map, 0, ..),
 * do_nms_sort(.45) and one of print_cocos / print_imagenet_detections / print_detector_detections.  Here the same list is
 * consumed net->batch images at a time: b200_letterbox_batch_u8 resizes batch k+1 on the device while the records of batch k
 * are read back (b200_detect_submitted with B200_INPUT_RESIDENT), boxes come back corrected with each image's own size in
 * pixel coordinates, and the records go through the writers of writers.c in (image, box, class) order — the order the
 * reference's loops produce.  File names, the json brackets and the imagenet numbering follow :393-412 and :478-482.
 * Image decoding stays with the caller (stb is out of scope): images are decoded RGB, HWC, 8 bit. */
#include "darknet.h"
#include "b200_engine.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* utils.c:179-191: the file name without directories, cut at its first '.' */
static char *image_id_of(const char *path)
{
    const char *c = path, *next;
    while ((next = strchr(c, '/'))) c = next + 1;
    char *id = malloc(strlen(c) + 1);
    if (!id) return NULL;
    strcpy(id, c);
    char *dot = strchr(id, '.');
    if (dot) *dot = 0;
    return id;
}

int b200_validate_images(network *net, const unsigned char *const *rgb_hwc, const int *widths, const int *heights,
                         const char *const *paths, int m, const char *eval, const char *prefix, const char *outfile,
                         const char *const *names, float thresh, float nms)
{
    if (!net || m < 1 || !rgb_hwc || !widths || !heights || !paths || !prefix) return -1;
    const int batch = net->batch;
    const int classes = net->layers[net->n - 1].classes;
    const int coco = eval && 0 == strcmp(eval, "coco"), imagenet = eval && 0 == strcmp(eval, "imagenet");
    const int max_out = 1 << 20;
    char buff[1024];
    FILE *fp = NULL, **fps = NULL;
    int rc = 0, total = 0;

    b200_det *rec = malloc((size_t)max_out * sizeof *rec);
    char **ids = calloc((size_t)m, sizeof *ids);
    int *numbers = malloc((size_t)m * sizeof *numbers);
    if (!rec || !ids || !numbers) { rc = -1; goto done; }
    for (int i = 0; i < m; ++i) {
        numbers[i] = i + 1;                                          /* :463, i+t-nthreads+1 */
        if (!(ids[i] = image_id_of(paths[i]))) { rc = -1; goto done; }
    }

    if (coco) {
        snprintf(buff, sizeof buff, "%s/%s.json", prefix, outfile ? outfile : "coco_results");
        if (!(fp = fopen(buff, "w"))) { rc = -1; goto done; }
        fprintf(fp, "[\n");
    } else if (imagenet) {
        snprintf(buff, sizeof buff, "%s/%s.txt", prefix, outfile ? outfile : "imagenet-detection");
        if (!(fp = fopen(buff, "w"))) { rc = -1; goto done; }
    } else {
        if (!names || !(fps = calloc((size_t)classes, sizeof *fps))) { rc = -1; goto done; }
        for (int j = 0; j < classes; ++j) {
            snprintf(buff, sizeof buff, "%s/%s%s.txt", prefix, outfile ? outfile : "comp4_det_test_", names[j]);
            if (!(fps[j] = fopen(buff, "w"))) { rc = -1; goto done; }
        }
    }

    const int nb = (m + batch - 1) / batch;
    if (b200_letterbox_batch_u8(net, rgb_hwc, widths, heights, m < batch ? m : batch)) { rc = -1; goto done; }
    b200_submit_batch(net, B200_INPUT_RESIDENT);
    for (int k = 0; k < nb; ++k) {
        const int base = k * batch, have = m - base < batch ? m - base : batch;
        const int more = k + 1 < nb;
        if (more) {
            const int nbase = base + batch, nhave = m - nbase < batch ? m - nbase : batch;
            if (b200_letterbox_batch_u8(net, rgb_hwc + nbase, widths + nbase, heights + nbase, nhave)) rc = -1;
        }
        /* always collect the submitted batch, also after an error, so that the engine is left idle */
        int n = b200_detect_submitted(net, (more && rc == 0) ? B200_INPUT_RESIDENT : NULL, 0, 0, thresh, nms, 0, rec, max_out, NULL);
        if (rc) break;
        if (n >= max_out) fprintf(stderr, "b200-darknet: b200_validate_images: more than %d records in one batch, list truncated\n", max_out);
        int kept = 0;                                                /* a short last batch leaves stale images in the other slots */
        for (int r = 0; r < n; ++r) if (rec[r].image < have) rec[kept++] = rec[r];
        b200_sort_records(rec, kept);
        if (coco) rc = b200_write_coco(fp, rec, kept, paths + base, widths + base, heights + base);
        else if (imagenet) rc = b200_write_imagenet(fp, rec, kept, numbers + base, widths + base, heights + base);
        else rc = b200_write_voc(fps, rec, kept, (const char *const *)ids + base, widths + base, heights + base);
        if (rc) {
            if (more) b200_detect_submitted(net, NULL, 0, 0, thresh, nms, 0, rec, max_out, NULL);      /* drain the batch in flight */
            break;
        }
        total += kept;
    }
    if (coco && rc == 0) {                                           /* :478-482: the last ",\n" becomes "\n]\n" */
        fseek(fp, -2, SEEK_CUR);
        fprintf(fp, "\n]\n");
    }

done:
    if (fp) fclose(fp);
    if (fps) {
        for (int j = 0; j < classes; ++j) if (fps[j]) fclose(fps[j]);
        free(fps);
    }
    if (ids) for (int i = 0; i < m; ++i) free(ids[i]);
    free(ids); free(numbers); free(rec);
    return rc ? -1 : total;
}
