/* layers.h — host-side construction of the `layer` records for the YOLO inference path. */
#ifndef B200_LAYERS_H
#define B200_LAYERS_H
#include "darknet.h"
#include "cfg.h"

typedef struct {            /* running shape while walking the cfg (parser.c:119-128 size_params) */
    int batch, inputs, h, w, c, index;
    network *net;
} shape_cursor;

/* returns 1 and fills *out when `type` (e.g. "[convolutional]") is a layer of the inference path */
int  build_layer(const char *type, cfg_section *opt, shape_cursor cur, layer *out);
ACTIVATION activation_from_name(const char *s);
void release_layer_host(layer l);
#endif
