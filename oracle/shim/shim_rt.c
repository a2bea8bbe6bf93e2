/* oracle/shim/shim_rt.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Runtime for the header stand-ins in this directory: the handful of htslib record helpers,
 * the nucleotide tables, an in-memory fai_fetch, and sonLib's stList/stHash.  Together with
 * oracle/probaln_port.c this is everything the reference's marker-path sources need to link
 * (see oracle/Makefile, target _ref).  Written from the public BAM specification and from the
 * semantics the reference's call sites rely on; no reference or htslib source is copied.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "sam.h"
#include "faidx.h"
#include "sonLib.h"

/* ------------------------------------------------------------------ bam1_t helpers */

bam1_t *bam_init1(void) { return (bam1_t *) calloc(1, sizeof(bam1_t)); }

void bam_destroy1(bam1_t *b) {
    if (!b) return;
    free(b->data);
    free(b);
}

bam1_t *bam_copy1(bam1_t *dst, const bam1_t *src) {
    uint8_t *buf = (uint8_t *) realloc(dst->data, src->l_data > 0 ? (size_t) src->l_data : 1);
    if (!buf) return NULL;
    memcpy(buf, src->data, (size_t) src->l_data);
    dst->core = src->core;
    dst->id = src->id;
    dst->data = buf;
    dst->l_data = src->l_data;
    dst->m_data = (uint32_t) src->l_data;
    return dst;
}

/* Aux area: tag[2] type[1] payload.  Returns a pointer to the TYPE byte (htslib contract,
 * relied upon by cigar_it.c:47,56 which skip one char).  Only the types a long-read BAM
 * carries are sized here. */
static int aux_payload_size(const uint8_t *p, const uint8_t *end) {
    switch (*p) {
        case 'A': case 'c': case 'C': return 1;
        case 's': case 'S': return 2;
        case 'i': case 'I': case 'f': return 4;
        case 'd': return 8;
        case 'Z': case 'H': {
            const uint8_t *q = p + 1;
            while (q < end && *q) q++;
            return (int) (q - (p + 1)) + 1;
        }
        case 'B': {
            int esz;
            uint32_t n;
            switch (p[1]) {
                case 'c': case 'C': esz = 1; break;
                case 's': case 'S': esz = 2; break;
                default: esz = 4;
            }
            memcpy(&n, p + 2, 4);
            return 1 + 4 + esz * (int) n;
        }
        default: return -1;
    }
}

uint8_t *bam_aux_get(const bam1_t *b, const char tag[2]) {
    uint8_t *p = bam_get_aux(b);
    uint8_t *end = b->data + b->l_data;
    while (p + 3 <= end) {
        int sz = aux_payload_size(p + 2, end);
        if (sz < 0) return NULL;
        if (p[0] == (uint8_t) tag[0] && p[1] == (uint8_t) tag[1]) return p + 2;
        p += 3 + sz;
    }
    return NULL;
}

const char *sam_hdr_tid2name(const sam_hdr_t *h, int tid) {
    if (!h || tid < 0 || tid >= h->n_targets) return NULL;
    return h->target_name[tid];
}

/* ------------------------------------------------------------------ nucleotide tables */

#define X 15
const unsigned char seq_nt16_table[256] = {
    X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X,
    X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X,
    X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X,
    1, 2, 4, 8, X, X, X, X, X, X, X, X, X, 0, X, X,
    X, 1, 14, 2, 13, X, X, 4, 11, X, X, 12, X, 3, X, X,
    X, X, 5, 6, 8, X, 7, 9, X, 10, X, X, X, X, X, X,
    X, 1, 14, 2, 13, X, X, 4, 11, X, X, 12, X, 3, X, X,
    X, X, 5, 6, 8, X, 7, 9, X, 10, X, X, X, X, X, X,
    X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X,
    X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X,
    X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X,
    X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X,
    X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X,
    X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X,
    X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X,
    X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X};
#undef X
const int seq_nt16_int[] = {4, 0, 1, 4, 2, 4, 4, 4, 3, 4, 4, 4, 4, 4, 4, 4};

/* ------------------------------------------------------------------ in-memory faidx */

faidx_t *fai_load(const char *fn) {
    (void) fn;
    return NULL; /* the oracle driver builds faidx_t itself */
}
void fai_destroy(faidx_t *fai) { (void) fai; }

/* region syntax used by the reference: "{name}:b-e", 1-based inclusive (ptMarker.c:739) */
char *fai_fetch(const faidx_t *fai, const char *reg, int *len) {
    *len = -1;
    if (!fai || reg[0] != '{') return NULL;
    const char *close = strrchr(reg, '}');
    if (!close || close[1] != ':') return NULL;
    size_t nlen = (size_t) (close - (reg + 1));
    long b = 0, e = 0;
    if (sscanf(close + 2, "%ld-%ld", &b, &e) != 2) return NULL;
    for (int i = 0; i < fai->n; i++) {
        if (strlen(fai->names[i]) == nlen && strncmp(fai->names[i], reg + 1, nlen) == 0) {
            long s = b - 1, t = e; /* 0-based half open */
            if (s < 0) s = 0;
            if (t > fai->lens[i]) t = fai->lens[i];
            if (t < s) t = s;
            char *out = (char *) malloc((size_t) (t - s) + 1);
            memcpy(out, fai->seqs[i] + s, (size_t) (t - s));
            out[t - s] = '\0';
            *len = (int) (t - s);
            return out;
        }
    }
    return NULL;
}

/* ------------------------------------------------------------------ stList */

struct stList {
    void **items;
    int64_t n, cap;
    void (*destruct)(void *);
};

stList *stList_construct3(int64_t size, void (*destructElement)(void *)) {
    stList *l = (stList *) malloc(sizeof(stList));
    l->cap = size > 8 ? size : 8;
    l->n = size;
    l->items = (void **) calloc((size_t) l->cap, sizeof(void *));
    l->destruct = destructElement;
    return l;
}

void stList_destruct(stList *l) {
    if (!l) return;
    if (l->destruct)
        for (int64_t i = 0; i < l->n; i++)
            if (l->items[i]) l->destruct(l->items[i]);
    free(l->items);
    free(l);
}

void stList_append(stList *l, void *item) {
    if (l->n == l->cap) {
        l->cap *= 2;
        l->items = (void **) realloc(l->items, (size_t) l->cap * sizeof(void *));
    }
    l->items[l->n++] = item;
}

void *stList_get(stList *l, int64_t i) {
    if (i < 0 || i >= l->n) {
        fprintf(stderr, "[oracle shim] stList_get index %ld out of range (%ld)\n", (long) i, (long) l->n);
        abort();
    }
    return l->items[i];
}

int64_t stList_length(stList *l) { return l ? l->n : 0; }

static __thread int (*g_cmp)(const void *, const void *);
static int cmp_ptr(const void *a, const void *b) { return g_cmp(*(void *const *) a, *(void *const *) b); }

void stList_sort(stList *l, int (*cmpFn)(const void *a, const void *b)) {
    g_cmp = cmpFn;
    qsort(l->items, (size_t) l->n, sizeof(void *), cmp_ptr);
}

stList *stList_copy(stList *l, void (*destructItem)(void *)) {
    stList *c = stList_construct3(0, destructItem);
    for (int64_t i = 0; i < l->n; i++) stList_append(c, l->items[i]);
    return c;
}

/* ------------------------------------------------------------------ stHash (insertion-ordered assoc list) */

struct stHash {
    void **keys, **vals;
    int64_t n, cap;
    uint64_t (*hk)(const void *);
    int (*eq)(const void *, const void *);
    void (*dk)(void *);
    void (*dv)(void *);
};
struct stHashIterator {
    stHash *h;
    int64_t i;
};

uint64_t stHash_stringKey(const void *k) {
    uint64_t h = 1469598103934665603ull;
    for (const unsigned char *p = (const unsigned char *) k; *p; p++) h = (h ^ *p) * 1099511628211ull;
    return h;
}
int stHash_stringEqualKey(const void *a, const void *b) { return strcmp((const char *) a, (const char *) b) == 0; }

stHash *stHash_construct3(uint64_t (*hashKey)(const void *), int (*eq)(const void *, const void *),
                          void (*dk)(void *), void (*dv)(void *)) {
    stHash *h = (stHash *) calloc(1, sizeof(stHash));
    h->cap = 8;
    h->keys = (void **) calloc(8, sizeof(void *));
    h->vals = (void **) calloc(8, sizeof(void *));
    h->hk = hashKey;
    h->eq = eq;
    h->dk = dk;
    h->dv = dv;
    return h;
}

void stHash_destruct(stHash *h) {
    if (!h) return;
    for (int64_t i = 0; i < h->n; i++) {
        if (h->dk) h->dk(h->keys[i]);
        if (h->dv) h->dv(h->vals[i]);
    }
    free(h->keys);
    free(h->vals);
    free(h);
}

static int64_t hash_find(stHash *h, const void *key) {
    for (int64_t i = 0; i < h->n; i++)
        if (h->eq(h->keys[i], key)) return i;
    return -1;
}

void stHash_insert(stHash *h, void *key, void *value) {
    int64_t i = hash_find(h, key);
    if (i >= 0) {
        h->vals[i] = value;
        return;
    }
    if (h->n == h->cap) {
        h->cap *= 2;
        h->keys = (void **) realloc(h->keys, (size_t) h->cap * sizeof(void *));
        h->vals = (void **) realloc(h->vals, (size_t) h->cap * sizeof(void *));
    }
    h->keys[h->n] = key;
    h->vals[h->n] = value;
    h->n++;
}

void *stHash_search(stHash *h, void *key) {
    int64_t i = hash_find(h, key);
    return i >= 0 ? h->vals[i] : NULL;
}

stHashIterator *stHash_getIterator(stHash *h) {
    stHashIterator *it = (stHashIterator *) malloc(sizeof(stHashIterator));
    it->h = h;
    it->i = 0;
    return it;
}
void *stHash_getNext(stHashIterator *it) { return it->i < it->h->n ? it->h->keys[it->i++] : NULL; }
void stHash_destructIterator(stHashIterator *it) { free(it); }

stList *stHash_getKeys(stHash *h) {
    stList *l = stList_construct3(0, NULL);
    for (int64_t i = 0; i < h->n; i++) stList_append(l, h->keys[i]);
    return l;
}
