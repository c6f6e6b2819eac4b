/* cfg.h — INI-style darknet .cfg reader (host C).  Behaviour follows the reference's read_cfg /
 * option_find* (parser.c:876-909, option_list.c:52-140): `[section]` headers, key=value lines,
 * '#' ';' comments, whitespace stripped everywhere, typed lookups that mark a key as used and log
 * "Using default" / "Unused field" notices on stderr with the reference's wording. */
#ifndef B200_CFG_H
#define B200_CFG_H

typedef struct { char *key, *val; int used; } cfg_kv;
typedef struct { char *type; cfg_kv *kv; int n, cap; } cfg_section;
typedef struct { cfg_section *sec; int n, cap; } cfg_file;

cfg_file   *cfg_read(const char *path);            /* exits like file_error() when the file is missing */
void        cfg_free(cfg_file *f);
const char *cfg_find(cfg_section *s, const char *key);                       /* option_find          */
const char *cfg_str(cfg_section *s, const char *key, const char *def);      /* option_find_str      */
int         cfg_int(cfg_section *s, const char *key, int def);              /* option_find_int      */
int         cfg_int_quiet(cfg_section *s, const char *key, int def);        /* option_find_int_quiet*/
float       cfg_float(cfg_section *s, const char *key, float def);          /* option_find_float    */
float       cfg_float_quiet(cfg_section *s, const char *key, float def);    /* option_find_float_quiet */
void        cfg_report_unused(cfg_section *s);                              /* option_unused        */
#endif
