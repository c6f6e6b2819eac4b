#include "cfg.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static void die_missing(const char *path)
{
    /* utils.c:281-285 file_error(): message then exit(0) */
    fprintf(stderr, "Couldn't open file: %s\n", path);
    exit(0);
}

/* remove every blank, tab and newline in place (utils.c:302-313 strips ALL whitespace, not just the ends) */
static void squeeze(char *s)
{
    char *w = s;
    for (; *s; ++s) if (*s != ' ' && *s != '\t' && *s != '\n' && *s != '\r') *w++ = *s;
    *w = 0;
}

static char *read_line(FILE *fp)
{
    size_t cap = 256, len = 0;
    char *buf = malloc(cap);
    int ch;
    if (feof(fp)) { free(buf); return NULL; }
    while ((ch = fgetc(fp)) != EOF && ch != '\n') {
        if (len + 2 > cap) buf = realloc(buf, cap *= 2);
        buf[len++] = (char)ch;
    }
    if (ch == EOF && len == 0) { free(buf); return NULL; }
    buf[len] = 0;
    return buf;
}

static cfg_section *push_section(cfg_file *f, char *type)
{
    if (f->n == f->cap) f->sec = realloc(f->sec, (f->cap = f->cap ? 2 * f->cap : 64) * sizeof(cfg_section));
    cfg_section *s = &f->sec[f->n++];
    memset(s, 0, sizeof *s);
    s->type = type;
    return s;
}

static int push_option(cfg_section *s, char *line)
{
    char *eq = strchr(line, '=');
    if (!eq || eq[1] == 0) return 0;     /* option_list.c:63 rejects a trailing '=' */
    *eq = 0;
    if (s->n == s->cap) s->kv = realloc(s->kv, (s->cap = s->cap ? 2 * s->cap : 16) * sizeof(cfg_kv));
    s->kv[s->n].key = line;          /* key and val share one heap block owned by .key */
    s->kv[s->n].val = eq + 1;
    s->kv[s->n].used = 0;
    s->n++;
    return 1;
}

cfg_file *cfg_read(const char *path)
{
    FILE *fp = fopen(path, "r");
    if (!fp) die_missing(path);
    cfg_file *f = calloc(1, sizeof *f);
    cfg_section *cur = NULL;
    char *line;
    int lineno = 0;
    while ((line = read_line(fp)) != NULL) {
        ++lineno;
        squeeze(line);
        switch (line[0]) {
        case '[':
            cur = push_section(f, line);
            break;
        case 0: case '#': case ';':
            free(line);
            break;
        default:
            if (!cur || !push_option(cur, line)) {
                fprintf(stderr, "Config file error line %d, could parse: %s\n", lineno, line);
                free(line);
            }
        }
    }
    fclose(fp);
    return f;
}

void cfg_free(cfg_file *f)
{
    if (!f) return;
    for (int i = 0; i < f->n; ++i) {
        for (int j = 0; j < f->sec[i].n; ++j) free(f->sec[i].kv[j].key);
        free(f->sec[i].kv);
        free(f->sec[i].type);
    }
    free(f->sec);
    free(f);
}

const char *cfg_find(cfg_section *s, const char *key)
{
    for (int i = 0; i < s->n; ++i)
        if (strcmp(s->kv[i].key, key) == 0) { s->kv[i].used = 1; return s->kv[i].val; }
    return NULL;
}

const char *cfg_str(cfg_section *s, const char *key, const char *def)
{
    const char *v = cfg_find(s, key);
    if (v) return v;
    if (def) fprintf(stderr, "%s: Using default '%s'\n", key, def);
    return def;
}

int cfg_int(cfg_section *s, const char *key, int def)
{
    const char *v = cfg_find(s, key);
    if (v) return atoi(v);
    fprintf(stderr, "%s: Using default '%d'\n", key, def);
    return def;
}

int cfg_int_quiet(cfg_section *s, const char *key, int def)
{
    const char *v = cfg_find(s, key);
    return v ? atoi(v) : def;
}

float cfg_float(cfg_section *s, const char *key, float def)
{
    const char *v = cfg_find(s, key);
    if (v) return (float)atof(v);
    fprintf(stderr, "%s: Using default '%lf'\n", key, def);
    return def;
}

float cfg_float_quiet(cfg_section *s, const char *key, float def)
{
    const char *v = cfg_find(s, key);
    return v ? (float)atof(v) : def;
}

void cfg_report_unused(cfg_section *s)
{
    for (int i = 0; i < s->n; ++i)
        if (!s->kv[i].used) fprintf(stderr, "Unused field: '%s = %s'\n", s->kv[i].key, s->kv[i].val);
}
