/*
 * compat.c — small host symbols that the reference's drivers and ctypes wrapper resolve at load time
 * (python/darknet.py:48-115 binds free_ptrs, reset_rnn, get_metadata, load_image_color, rgbgr_image).
 * They sit next to the accelerated path, not on it (SURVEY.md §8b "host helpers ... may be plain re-statements").
 */
#include "darknet.h"
#include "cfg.h"
#include <ctype.h>

void free_ptrs(void **ptrs, int n)
{
    for (int i = 0; i < n; ++i) free(ptrs[i]);
    free(ptrs);
}

void reset_rnn(network *net) { (void)net; }      /* no recurrent layers on the YOLO path */

void rgbgr_image(image im)
{
    if (im.c < 3) return;
    size_t plane = (size_t)im.w * im.h;
    for (size_t i = 0; i < plane; ++i) {
        float t = im.data[i];
        im.data[i] = im.data[i + 2 * plane];
        im.data[i + 2 * plane] = t;
    }
}

/* `.data` files: key=value lines without sections (option_list.c:7-33) */
static char *data_value(const char *path, const char *key)
{
    FILE *fp = fopen(path, "r");
    if (!fp) { fprintf(stderr, "Couldn't open file: %s\n", path); exit(0); }
    char line[4096], *found = NULL;
    while (fgets(line, sizeof line, fp)) {
        char *w = line;
        for (char *r = line; *r; ++r) if (!isspace((unsigned char)*r)) *w++ = *r;
        *w = 0;
        if (line[0] == '#' || line[0] == ';' || line[0] == 0) continue;
        char *eq = strchr(line, '=');
        if (!eq) continue;
        *eq = 0;
        if (strcmp(line, key) == 0) { free(found); found = strdup(eq + 1); }
    }
    fclose(fp);
    return found;
}

static char **read_lines(const char *path, int *count)
{
    FILE *fp = fopen(path, "r");
    if (!fp) { fprintf(stderr, "Couldn't open file: %s\n", path); exit(0); }
    int cap = 64, n = 0;
    char **v = malloc(cap * sizeof(char *));
    char line[4096];
    while (fgets(line, sizeof line, fp)) {
        size_t len = strlen(line);
        while (len && (line[len - 1] == '\n' || line[len - 1] == '\r')) line[--len] = 0;
        if (n == cap) v = realloc(v, (cap *= 2) * sizeof(char *));
        v[n++] = strdup(line);
    }
    fclose(fp);
    *count = n;
    return v;
}

metadata get_metadata(char *file)
{
    /* option_list.c:35-50 */
    metadata m = {0};
    char *names = data_value(file, "names");
    if (!names) names = data_value(file, "labels");
    if (!names) fprintf(stderr, "No names or labels found\n");
    else { int n = 0; m.names = read_lines(names, &n); free(names); }
    char *classes = data_value(file, "classes");
    if (classes) { m.classes = atoi(classes); free(classes); }
    else { fprintf(stderr, "classes: Using default '2'\n"); m.classes = 2; }
    return m;
}

/* Binary PPM (P6) / PGM (P5) reader.  The reference decodes JPEG/PNG through the vendored stb_image
 * (image.c:1442-1464), which is out of scope here; feed other formats through your own decoder and
 * make_image(). */
image load_image_color(char *filename, int w, int h)
{
    FILE *fp = fopen(filename, "rb");
    if (!fp) { fprintf(stderr, "Cannot load image \"%s\"\n", filename); exit(0); }
    char magic[3] = {0};
    int iw = 0, ih = 0, maxv = 0;
    if (fscanf(fp, "%2s", magic) != 1 || (strcmp(magic, "P6") && strcmp(magic, "P5"))) {
        fprintf(stderr, "b200-darknet: load_image_color reads binary PPM/PGM only (\"%s\")\n", filename);
        exit(0);
    }
    int ch = fgetc(fp);
    while (ch == '#' || isspace(ch)) { if (ch == '#') while (ch != '\n' && ch != EOF) ch = fgetc(fp); ch = fgetc(fp); }
    ungetc(ch, fp);
    if (fscanf(fp, "%d %d %d", &iw, &ih, &maxv) != 3 || maxv <= 0 || maxv > 255) { fprintf(stderr, "bad PNM header in %s\n", filename); exit(0); }
    fgetc(fp);
    int src_c = magic[1] == '6' ? 3 : 1;
    unsigned char *raw = malloc((size_t)iw * ih * src_c);
    if (fread(raw, 1, (size_t)iw * ih * src_c, fp) != (size_t)iw * ih * src_c) { fprintf(stderr, "short read in %s\n", filename); exit(0); }
    fclose(fp);
    image im = make_image(iw, ih, 3);
    for (int k = 0; k < 3; ++k)
        for (int y = 0; y < ih; ++y)
            for (int x = 0; x < iw; ++x)
                im.data[((size_t)k * ih + y) * iw + x] = raw[((size_t)y * iw + x) * src_c + (src_c == 3 ? k : 0)] / 255.f;
    free(raw);
    if (w && h && (w != iw || h != ih)) { image r = resize_image(im, w, h); free_image(im); im = r; }
    return im;
}
