/*
 * weights.c — darknet `.weights` reader (host C).
 *
 * File format (reference parser.c:1241-1345, writer side :992-1068):
 *   int32 major, minor, revision; then `seen` as 8 bytes when major*10+minor >= 2 (and both < 1000),
 *   else as int32; then, for every layer in cfg order that carries parameters:
 *     CONVOLUTIONAL : biases[n], (scales[n], rolling_mean[n], rolling_variance[n] if batch_normalize),
 *                     weights[n * c/groups * size * size]  (OIHW)            (parser.c:1163-1238)
 *     CONNECTED     : biases[outputs], weights[outputs*inputs], (BN triplet) (parser.c:1083-1119)
 *     LOCAL         : biases[outputs], weights[size*size*c*n*locations]      (parser.c:1315-1341)
 * The fp32 arrays land in the public `layer` fields exactly as in the reference (drivers may read
 * them); the engine then folds BN, repacks for NHWC and uploads (b200_engine_upload_weights).
 * Unlike the reference fork, nothing is dumped to stdout.  The reference ignores fread's result, so a file that ends at a
 * layer boundary (a backbone-only file such as darknet53.conv.74 loaded with plain load_weights) simply leaves the remaining
 * layers as initialised: the same here, with a notice.  A file that ends INSIDE an array is corrupt and stays fatal.
 */
#include "darknet.h"
#include "b200_engine.h"

static int g_at_layer_start, g_file_ended;

static void read_floats(float *dst, size_t n, FILE *fp, const char *what, int layer_index)
{
    if (g_file_ended) return;
    size_t got = fread(dst, sizeof(float), n, fp);
    if (got == 0 && g_at_layer_start && feof(fp)) {
        fprintf(stderr, "\nb200-darknet: weights file ends before layer %d: it and the layers after it keep their initial values\n", layer_index);
        g_file_ended = 1;
        return;
    }
    g_at_layer_start = 0;
    if (got != n) {
        fprintf(stderr, "\nb200-darknet: weights file too short: layer %d %s wanted %zu floats, got %zu\n",
                layer_index, what, n, got);
        exit(-1);
    }
}

/* in-place transpose of a rows x cols matrix (used for `flipped` convs and pre-0.2 files: parser.c:1070-1081) */
static void transpose_in_place(float *a, int rows, int cols)
{
    float *t = calloc((size_t)rows * cols, sizeof(float));
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) t[(size_t)c * rows + r] = a[(size_t)r * cols + c];
    memcpy(a, t, (size_t)rows * cols * sizeof(float));
    free(t);
}

void load_weights_upto(network *net, char *filename, int start, int cutoff)
{
    fprintf(stderr, "Loading weights from %s...", filename);
    FILE *fp = fopen(filename, "rb");
    if (!fp) { fprintf(stderr, "Couldn't open file: %s\n", filename); exit(0); }

    int major = 0, minor = 0, revision = 0;
    if (fread(&major, sizeof(int), 1, fp) != 1 || fread(&minor, sizeof(int), 1, fp) != 1 ||
        fread(&revision, sizeof(int), 1, fp) != 1) {
        fprintf(stderr, "\nb200-darknet: %s has no header\n", filename);
        exit(-1);
    }
    if ((major * 10 + minor) >= 2 && major < 1000 && minor < 1000) {
        unsigned long long seen = 0;
        if (fread(&seen, sizeof(seen), 1, fp) != 1) exit(-1);
        *net->seen = (size_t)seen;
    } else {
        int seen = 0;
        if (fread(&seen, sizeof(int), 1, fp) != 1) exit(-1);
        *net->seen = (size_t)seen;
    }
    int transpose = (major > 1000) || (minor > 1000);
    g_file_ended = 0;

    for (int i = start; i < net->n && i < cutoff; ++i) {
        layer l = net->layers[i];
        if (l.dontload) continue;
        g_at_layer_start = 1;
        if (l.type == CONVOLUTIONAL) {
            read_floats(l.biases, l.n, fp, "biases", i);
            if (l.batch_normalize && !l.dontloadscales) {
                read_floats(l.scales, l.n, fp, "scales", i);
                read_floats(l.rolling_mean, l.n, fp, "rolling_mean", i);
                read_floats(l.rolling_variance, l.n, fp, "rolling_variance", i);
            }
            read_floats(l.weights, l.nweights, fp, "weights", i);
            if (l.flipped) transpose_in_place(l.weights, l.c * l.size * l.size, l.n);
        } else if (l.type == CONNECTED) {
            read_floats(l.biases, l.outputs, fp, "biases", i);
            read_floats(l.weights, (size_t)l.outputs * l.inputs, fp, "weights", i);
            if (transpose) transpose_in_place(l.weights, l.inputs, l.outputs);
            if (l.batch_normalize && !l.dontloadscales) {
                read_floats(l.scales, l.outputs, fp, "scales", i);
                read_floats(l.rolling_mean, l.outputs, fp, "rolling_mean", i);
                read_floats(l.rolling_variance, l.outputs, fp, "rolling_variance", i);
            }
        } else if (l.type == LOCAL) {
            int locations = l.out_w * l.out_h;
            read_floats(l.biases, l.outputs, fp, "biases", i);
            read_floats(l.weights, (size_t)l.size * l.size * l.c * l.n * locations, fp, "weights", i);
        }
    }
    fprintf(stderr, "Done!\n");
    fclose(fp);
    b200_engine_upload_weights(b200_engine_of(net), net);
}

void load_weights(network *net, char *filename)
{
    load_weights_upto(net, filename, 0, net->n);
}
