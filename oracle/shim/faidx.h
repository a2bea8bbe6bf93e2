/* oracle/shim/faidx.h -- TEST INFRASTRUCTURE ONLY.  In-memory stand-in for htslib's faidx:
 * the reference only calls fai_fetch(fai, "{ctg}:b-e", &len) with 1-based inclusive
 * coordinates (ptMarker.c:739-741). */
#ifndef ORACLE_SHIM_FAIDX_H
#define ORACLE_SHIM_FAIDX_H
#ifdef __cplusplus
extern "C" {
#endif
typedef struct faidx_t {
    int n;
    char **names;
    const char **seqs; /* ASCII bases, not owned */
    long *lens;
} faidx_t;
char *fai_fetch(const faidx_t *fai, const char *reg, int *len);
faidx_t *fai_load(const char *fn);
void fai_destroy(faidx_t *fai);
#ifdef __cplusplus
}
#endif
#endif
