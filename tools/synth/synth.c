/* tools/synth/synth.c -- synthetic assembly + read-group generator (test / bench data only).
 * See synth.h.  Nothing here is derived from the reference; the record layout follows the
 * public SAM/BAM specification (CIGAR packing len<<4|op, 4-bit SEQ high nibble first) and
 * minimap2's documented short-form cs tag (":n", "*<ref><read>", "+<ins>", "-<del>").
 */
#include "synth.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ rng */
typedef struct { uint64_t s; } rng_t;
static inline uint64_t rng_next(rng_t *r) {
    uint64_t z = (r->s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline double rng_u(rng_t *r) { return (double) (rng_next(r) >> 11) * (1.0 / 9007199254740992.0); }
static inline uint64_t rng_below(rng_t *r, uint64_t n) { return n ? rng_next(r) % n : 0; }
static inline double rng_normal(rng_t *r) {
    double u1 = rng_u(r), u2 = rng_u(r);
    if (u1 < 1e-300) u1 = 1e-300;
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}
static inline int64_t rng_geom(rng_t *r, double p) { /* trials until success, >= 1 */
    if (p <= 0) return INT64_MAX / 4;
    if (p >= 1) return 1;
    double u = rng_u(r);
    if (u < 1e-300) u = 1e-300;
    return 1 + (int64_t) floor(log(u) / log1p(-p));
}
static const char BASES[4] = {'A', 'C', 'G', 'T'};
static inline char comp(char c) {
    switch (c) {
        case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
        case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a';
        default: return c;
    }
}
static inline char other_base(rng_t *r, char c) {
    for (;;) {
        char b = BASES[rng_below(r, 4)];
        if (b != c) return b;
    }
}

/* ------------------------------------------------------------------ vectors */
#define VEC(T) struct { T *p; int64_t n, cap; }
#define VPUSH(v, x)                                                                \
    do {                                                                           \
        if ((v).n == (v).cap) {                                                    \
            (v).cap = (v).cap ? (v).cap * 2 : 256;                                 \
            (v).p = realloc((v).p, (size_t) (v).cap * sizeof(*(v).p));             \
        }                                                                          \
        (v).p[(v).n++] = (x);                                                      \
    } while (0)
#define VRESERVE(v, extra)                                                         \
    do {                                                                           \
        if ((v).n + (extra) > (v).cap) {                                           \
            while ((v).n + (extra) > (v).cap) (v).cap = (v).cap ? (v).cap * 2 : 256; \
            (v).p = realloc((v).p, (size_t) (v).cap * sizeof(*(v).p));             \
        }                                                                          \
    } while (0)

/* ------------------------------------------------------------------ assembly */
enum { EV_SNV = 1, EV_DEL = 2, EV_INS = 3 };
typedef struct {
    int64_t pos;   /* copy-0 coordinate the event applies to */
    int64_t cum;   /* length shift accumulated by the events before this one */
    int64_t ins_off;
    int32_t len;
    uint8_t type;
    char alt;
} event_t;
typedef struct {
    event_t *ev;
    int64_t n;
    char *ins_pool;
} copyvar_t;
typedef struct {
    char *seq; /* as stored in the FASTA (possibly reverse-complemented) */
    int64_t len;
    int rc;
    char name[64];
} contig_t;

struct synth {
    synth_cfg cfg;
    int32_t n_contigs;
    contig_t *ctg;  /* [locus * n_copies + copy] */
    copyvar_t *var; /* same indexing; copy 0 has no events */
};

typedef struct {
    const copyvar_t *v;
    int64_t ei;
    int32_t del_remain;
} cit_t;

static void cit_init(cit_t *it, const copyvar_t *v, int64_t s1, int64_t *tpos_out) {
    it->v = v;
    it->ei = 0;
    it->del_remain = 0;
    if (!v || v->n == 0) {
        if (tpos_out) *tpos_out = s1;
        return;
    }
    int64_t lo = 0, hi = v->n; /* first event with pos >= s1 */
    while (lo < hi) {
        int64_t mid = (lo + hi) / 2;
        if (v->ev[mid].pos >= s1) hi = mid; else lo = mid + 1;
    }
    it->ei = lo;
    int64_t tpos = s1;
    if (lo > 0) {
        const event_t *e = &v->ev[lo - 1];
        int64_t after = e->cum + (e->type == EV_INS ? e->len : e->type == EV_DEL ? -e->len : 0);
        if (e->type == EV_DEL && e->pos + e->len > s1) {
            it->del_remain = (int32_t) (e->pos + e->len - s1);
            tpos = e->pos + e->cum;
        } else {
            tpos = s1 + after;
        }
    }
    if (tpos_out) *tpos_out = tpos;
}

/* element(s) of this copy at copy-0 position p */
static inline void cit_step(cit_t *it, int64_t p, char anc, int *present, char *base, const char **ins, int *ins_len) {
    *ins = NULL;
    *ins_len = 0;
    *present = 1;
    *base = anc;
    if (it->del_remain > 0) {
        it->del_remain--;
        *present = 0;
        return;
    }
    const copyvar_t *v = it->v;
    if (!v || it->ei >= v->n || v->ev[it->ei].pos != p) return;
    const event_t *e = &v->ev[it->ei++];
    if (e->type == EV_SNV) {
        *base = e->alt;
    } else if (e->type == EV_DEL) {
        *present = 0;
        it->del_remain = e->len - 1;
    } else {
        *ins = v->ins_pool + e->ins_off;
        *ins_len = e->len;
    }
}

void synth_default_cfg(synth_cfg *c, int preset) {
    memset(c, 0, sizeof(*c));
    c->seed = 20240600;
    c->n_loci = 1;
    c->n_copies = 2;
    c->locus_len = 5000000;
    c->snv_rate = 1e-3;
    c->indel_rate = 1e-4;
    c->long_indel_rate = 5e-6;
    c->hp_frac = 0.05;
    c->n_rate = 0;
    c->rc_contig_prob = 0.5;
    c->len_mean = 15000;
    c->len_sd = 2000;
    c->len_min = 5000;
    c->len_max = 25000;
    c->err_sub = 2e-4;
    c->err_ins = 5e-4;
    c->err_del = 5e-4;
    c->hp_indel_mult = 5;
    c->long_err_indel_prob = 0;
    c->qual_model = 0;
    c->max_secondaries = 1;
    c->wrong_primary_prob = 0.5;
    c->clip_prob = 0.2;
    c->clip_max = 200;
    c->hard_clip_prob = 0.5;
    c->eqx = 0;
    c->use_md = 0;
    if (preset == 1) { /* ONT */
        c->len_mean = 30000;
        c->len_sd = 8000;
        c->len_min = 5000;
        c->len_max = 80000;
        c->err_sub = 0.01;
        c->err_ins = 0.015;
        c->err_del = 0.02;
        c->hp_indel_mult = 2;
        c->long_err_indel_prob = 0.01;
        c->qual_model = 1;
    } else if (preset == 2) { /* stress: many near-identical copies, homopolymer rich */
        c->n_copies = 9;
        c->max_secondaries = 8;
        c->snv_rate = 1e-3;
        c->indel_rate = 2e-4;
        c->long_indel_rate = 2e-5;
        c->hp_frac = 0.30;
        c->locus_len = 2000000;
    }
}

static void build_locus(synth *s, int locus, rng_t *r) {
    const synth_cfg *c = &s->cfg;
    int64_t L = c->locus_len;
    contig_t *c0 = &s->ctg[locus * c->n_copies];
    c0->seq = (char *) malloc((size_t) L + 1);
    c0->len = L;
    c0->rc = 0;
    double hp_start = c->hp_frac / 12.5;
    for (int64_t i = 0; i < L;) {
        if (rng_u(r) < hp_start) {
            int run = 5 + (int) rng_below(r, 16);
            char b = BASES[rng_below(r, 4)];
            for (int j = 0; j < run && i < L; j++) c0->seq[i++] = b;
        } else {
            c0->seq[i++] = (c->n_rate > 0 && rng_u(r) < c->n_rate) ? 'N' : BASES[rng_below(r, 4)];
        }
    }
    c0->seq[L] = 0;
    snprintf(c0->name, sizeof(c0->name), "syn#1#ctg%d", locus);

    double rate = c->snv_rate + c->indel_rate + c->long_indel_rate;
    for (int cp = 1; cp < c->n_copies; cp++) {
        copyvar_t *v = &s->var[locus * c->n_copies + cp];
        VEC(event_t) evs = {0};
        VEC(char) pool = {0};
        int64_t cum = 0;
        int64_t p = 50 + rng_geom(r, rate);
        while (p < L - 100) {
            event_t e;
            memset(&e, 0, sizeof(e));
            e.pos = p;
            e.cum = cum;
            double u = rng_u(r) * rate;
            int64_t span = 1;
            if (u < c->snv_rate) {
                if (c0->seq[p] == 'N') { p += 1; continue; }
                e.type = EV_SNV;
                e.len = 1;
                e.alt = other_base(r, c0->seq[p]);
            } else {
                int islong = u >= c->snv_rate + c->indel_rate;
                int len = islong ? 11 + (int) rng_below(r, 50) : 1 + (int) rng_below(r, 8);
                if (rng_below(r, 2)) {
                    e.type = EV_DEL;
                    e.len = len;
                    cum -= len;
                    span = len;
                } else {
                    e.type = EV_INS;
                    e.len = len;
                    e.ins_off = pool.n;
                    for (int j = 0; j < len; j++) VPUSH(pool, BASES[rng_below(r, 4)]);
                    cum += len;
                }
            }
            VPUSH(evs, e);
            p += span + 2 + rng_geom(r, rate);
        }
        v->ev = evs.p;
        v->n = evs.n;
        v->ins_pool = pool.p;
        /* materialise */
        contig_t *cc = &s->ctg[locus * c->n_copies + cp];
        int64_t Lc = L + cum;
        cc->seq = (char *) malloc((size_t) Lc + 1);
        cit_t it;
        cit_init(&it, v, 0, NULL);
        int64_t w = 0;
        for (int64_t q = 0; q < L; q++) {
            int pr, il;
            char b;
            const char *ins;
            cit_step(&it, q, c0->seq[q], &pr, &b, &ins, &il);
            if (pr) cc->seq[w++] = b;
            for (int j = 0; j < il; j++) cc->seq[w++] = ins[j];
        }
        cc->len = w;
        cc->seq[w] = 0;
        cc->rc = rng_u(r) < c->rc_contig_prob;
        if (cc->rc) {
            for (int64_t a = 0, b2 = w - 1; a < b2; a++, b2--) {
                char t = comp(cc->seq[a]);
                cc->seq[a] = comp(cc->seq[b2]);
                cc->seq[b2] = t;
            }
            if (w & 1) cc->seq[w / 2] = comp(cc->seq[w / 2]);
        }
        snprintf(cc->name, sizeof(cc->name), "syn#%d#ctg%d", cp + 1, locus);
    }
}

synth *synth_create(const synth_cfg *c) {
    if (c->n_copies < 1 || c->n_loci < 1 || c->locus_len < 2000) return NULL;
    synth *s = (synth *) calloc(1, sizeof(*s));
    s->cfg = *c;
    s->n_contigs = c->n_loci * c->n_copies;
    s->ctg = (contig_t *) calloc((size_t) s->n_contigs, sizeof(contig_t));
    s->var = (copyvar_t *) calloc((size_t) s->n_contigs, sizeof(copyvar_t));
    for (int l = 0; l < c->n_loci; l++) {
        rng_t r = {c->seed * 0x2545F4914F6CDD1Dull + 0x1234567ull * (uint64_t) (l + 1)};
        build_locus(s, l, &r);
    }
    return s;
}

void synth_destroy(synth *s) {
    if (!s) return;
    for (int i = 0; i < s->n_contigs; i++) {
        free(s->ctg[i].seq);
        free(s->var[i].ev);
        free(s->var[i].ins_pool);
    }
    free(s->ctg);
    free(s->var);
    free(s);
}
int32_t synth_n_contigs(const synth *s) { return s->n_contigs; }
const char *synth_contig_name(const synth *s, int32_t tid) { return s->ctg[tid].name; }
const char *synth_contig_seq(const synth *s, int32_t tid) { return s->ctg[tid].seq; }
int64_t synth_contig_len(const synth *s, int32_t tid) { return s->ctg[tid].len; }

/* ------------------------------------------------------------------ batch under construction */
struct synth_batch {
    sp_flat_batch view;
    VEC(int32_t) grp_aln_off, flag, tid, pos, l_qseq, n_cigar, tag_kind;
    VEC(int64_t) qname_off, cigar_off, tag_off, seq_off, qual_off;
    VEC(char) qname_pool, tag_pool;
    VEC(uint32_t) cigar_pool;
    VEC(uint8_t) seq_pool, qual_pool;
};

typedef struct { char type, rd, rf; } col_t; /* '=' 'X' 'I' 'D' ; read base ; ref base */
typedef VEC(col_t) colvec;

typedef struct {
    /* the read in walk orientation */
    VEC(char) rbase;
    VEC(uint8_t) rqual;
    /* per S element */
    VEC(uint8_t) op;      /* 0 copy, 1 sub, 2 del */
    VEC(char) subbase;
    VEC(int32_t) ins_len; /* read insertion after the element */
    VEC(int64_t) ins_off; /* into ins_pool */
    VEC(char) ins_pool;
} readsim_t;

static uint8_t draw_qual(const synth_cfg *c, rng_t *r) {
    if (c->qual_model == 1) {
        double q = 14 + 6 * rng_normal(r);
        if (q < 2) q = 2;
        if (q > 40) q = 40;
        return (uint8_t) (q + 0.5);
    }
    double u = rng_u(r);
    if (u < 0.70) return 93;
    if (u < 0.90) return (uint8_t) (40 + rng_below(r, 21));
    if (u < 0.98) return (uint8_t) (10 + rng_below(r, 21));
    return (uint8_t) rng_below(r, 10);
}

static int err_indel_len(const synth_cfg *c, rng_t *r) {
    if (c->long_err_indel_prob > 0 && rng_u(r) < c->long_err_indel_prob) return 5 + (int) rng_below(r, 21);
    int len = 1;
    while (len < 4 && rng_u(r) < 0.25) len++;
    return len;
}

/* simulate the read over the source copy's elements in [s1,e1) */
static void simulate_read(const synth *s, int locus, int S, int64_t s1, int64_t e1, rng_t *r, readsim_t *rs) {
    const synth_cfg *c = &s->cfg;
    const char *anc = s->ctg[locus * c->n_copies].seq;
    cit_t it;
    cit_init(&it, S ? &s->var[locus * c->n_copies + S] : NULL, s1, NULL);
    char prev = 0;
    int skip_del = 0;
    for (int64_t p = s1; p < e1; p++) {
        int pr, il;
        char b;
        const char *ins;
        cit_step(&it, p, anc[p], &pr, &b, &ins, &il);
        for (int j = -1; j < il; j++) {
            char sb;
            if (j < 0) {
                if (!pr) continue;
                sb = b;
            } else {
                sb = ins[j];
            }
            double mult = (sb == prev) ? c->hp_indel_mult : 1.0;
            prev = sb;
            uint8_t op = 0;
            char sub = 0;
            if (skip_del > 0) {
                op = 2;
                skip_del--;
            } else {
                double u = rng_u(r);
                if (u < c->err_del * mult) {
                    op = 2;
                    skip_del = err_indel_len(c, r) - 1;
                } else if (u < c->err_del * mult + c->err_sub && sb != 'N') {
                    op = 1;
                    sub = other_base(r, sb);
                }
            }
            VPUSH(rs->op, op);
            VPUSH(rs->subbase, sub);
            if (op != 2) {
                VPUSH(rs->rbase, op == 1 ? sub : sb);
                VPUSH(rs->rqual, draw_qual(c, r));
            }
            int32_t ilen = 0;
            int64_t ioff = rs->ins_pool.n;
            if (rng_u(r) < c->err_ins * mult) {
                ilen = err_indel_len(c, r);
                for (int t = 0; t < ilen; t++) {
                    char ib = (mult > 1.0 && rng_u(r) < 0.7) ? sb : BASES[rng_below(r, 4)];
                    if (ib == 'N') ib = 'A';
                    VPUSH(rs->ins_pool, ib);
                    VPUSH(rs->rbase, ib);
                    VPUSH(rs->rqual, draw_qual(c, r));
                }
            }
            VPUSH(rs->ins_len, ilen);
            VPUSH(rs->ins_off, ioff);
        }
    }
}

/* columns of the read against target copy T (walk orientation); returns T index of the first T element */
static int64_t build_columns(const synth *s, int locus, int S, int T, int64_t s1, int64_t e1, const readsim_t *rs,
                             colvec *cols) {
    const synth_cfg *c = &s->cfg;
    const char *anc = s->ctg[locus * c->n_copies].seq;
    cit_t is, itg;
    int64_t tpos0;
    cit_init(&is, S ? &s->var[locus * c->n_copies + S] : NULL, s1, NULL);
    cit_init(&itg, T ? &s->var[locus * c->n_copies + T] : NULL, s1, &tpos0);
    int64_t si = 0;
    cols->n = 0;
    for (int64_t p = s1; p < e1; p++) {
        int spr, sil, tpr, til;
        char sb, tb;
        const char *sins, *tins;
        cit_step(&is, p, anc[p], &spr, &sb, &sins, &sil);
        cit_step(&itg, p, anc[p], &tpr, &tb, &tins, &til);
        int m = sil > til ? sil : til;
        for (int j = -1; j < m; j++) {
            int shas, thas;
            char sbase = 0, tbase = 0;
            if (j < 0) {
                shas = spr; thas = tpr; sbase = sb; tbase = tb;
            } else {
                shas = j < sil; thas = j < til;
                if (shas) sbase = sins[j];
                if (thas) tbase = tins[j];
            }
            if (shas) {
                uint8_t op = rs->op.p[si];
                if (op != 2) {
                    char rb = op == 1 ? rs->subbase.p[si] : sbase;
                    col_t cc;
                    cc.rd = rb;
                    if (thas) {
                        cc.rf = tbase;
                        cc.type = (rb == tbase) ? '=' : 'X';
                    } else {
                        cc.rf = 0;
                        cc.type = 'I';
                    }
                    VPUSH(*cols, cc);
                } else if (thas) {
                    col_t cc = {'D', 0, tbase};
                    VPUSH(*cols, cc);
                }
                for (int t = 0; t < rs->ins_len.p[si]; t++) {
                    col_t cc = {'I', rs->ins_pool.p[rs->ins_off.p[si] + t], 0};
                    VPUSH(*cols, cc);
                }
                si++;
            } else if (thas) {
                col_t cc = {'D', 0, tbase};
                VPUSH(*cols, cc);
            }
        }
    }
    return tpos0;
}

static inline int nt4(char c) {
    switch (c) {
        case 'A': case 'a': return 1;
        case 'C': case 'c': return 2;
        case 'G': case 'g': return 4;
        case 'T': case 't': return 8;
        default: return 15;
    }
}
static inline char lower(char c) { return (c >= 'A' && c <= 'Z') ? (char) (c + 32) : c; }

static void push_num(synth_batch *b, long v) {
    char tmp[24];
    int n = snprintf(tmp, sizeof(tmp), "%ld", v);
    for (int i = 0; i < n; i++) VPUSH(b->tag_pool, tmp[i]);
}

/* serialise one alignment record */
static void emit_alignment(synth_batch *b, const synth *s, int tid, int is_secondary, int read_strand,
                           colvec *cols, int64_t tpos0, const readsim_t *rs, int clipL, int clipR, int hard, rng_t *r) {
    (void) r;
    const synth_cfg *c = &s->cfg;
    const contig_t *ct = &s->ctg[tid];
    int64_t n_read = rs->rbase.n;
    /* locate the aligned column range: drop clipped read bases, then trim to M columns */
    int64_t c0 = 0, c1 = cols->n; /* [c0,c1) */
    int64_t rdL = 0, tL = 0;      /* read bases / T elements consumed before c0 */
    while (c0 < c1 && !((cols->p[c0].type == '=' || cols->p[c0].type == 'X') && rdL >= clipL)) {
        if (cols->p[c0].type != 'D') rdL++;
        if (cols->p[c0].type != 'I') tL++;
        c0++;
    }
    int64_t rdR = 0;
    while (c1 > c0 && !((cols->p[c1 - 1].type == '=' || cols->p[c1 - 1].type == 'X') && rdR >= clipR)) {
        if (cols->p[c1 - 1].type != 'D') rdR++;
        c1--;
    }
    int64_t tn = 0, rn = 0;
    for (int64_t i = c0; i < c1; i++) {
        if (cols->p[i].type != 'I') tn++;
        if (cols->p[i].type != 'D') rn++;
    }
    int64_t tstart = tpos0 + tL;
    int64_t lclip = rdL, rclip = n_read - rdL - rn;
    /* orientation of the stored contig */
    int rc = ct->rc;
    int64_t pos = rc ? ct->len - (tstart + tn) : tstart;
    int flag = ((rc ^ read_strand) ? 0x10 : 0) | (is_secondary ? 0x100 : 0);
    if (rc) {
        for (int64_t a = c0, z = c1 - 1; a < z; a++, z--) {
            col_t t = cols->p[a];
            cols->p[a] = cols->p[z];
            cols->p[z] = t;
        }
        for (int64_t i = c0; i < c1; i++) {
            cols->p[i].rd = comp(cols->p[i].rd);
            cols->p[i].rf = comp(cols->p[i].rf);
        }
        int64_t t = lclip; lclip = rclip; rclip = t;
    }
    /* CIGAR */
    int64_t cig0 = b->cigar_pool.n;
    int clip_op = hard ? 5 : 4;
    if (lclip > 0) VPUSH(b->cigar_pool, (uint32_t) (lclip << 4 | clip_op));
    {
        int cur = -1;
        int64_t run = 0;
        for (int64_t i = c0; i <= c1; i++) {
            int op = -1;
            if (i < c1) {
                char t = cols->p[i].type;
                op = t == 'I' ? 1 : t == 'D' ? 2 : c->eqx ? (t == '=' ? 7 : 8) : 0;
            }
            if (op != cur) {
                if (cur >= 0) VPUSH(b->cigar_pool, (uint32_t) (run << 4 | cur));
                cur = op;
                run = 0;
            }
            run++;
        }
    }
    if (rclip > 0) VPUSH(b->cigar_pool, (uint32_t) (rclip << 4 | clip_op));
    /* SEQ / QUAL in stored orientation */
    int64_t q_lo = hard ? (rc ? rclip : lclip) : 0;                 /* walk-orientation read range kept */
    int64_t q_hi = hard ? n_read - (rc ? lclip : rclip) : n_read;
    int64_t l_qseq = q_hi - q_lo;
    int64_t seq0 = b->seq_pool.n, qual0 = b->qual_pool.n;
    VRESERVE(b->seq_pool, (l_qseq + 1) / 2 + 1);
    VRESERVE(b->qual_pool, l_qseq + 1);
    for (int64_t i = 0; i < l_qseq; i++) {
        int64_t w = rc ? q_hi - 1 - i : q_lo + i;
        char base = rc ? comp(rs->rbase.p[w]) : rs->rbase.p[w];
        int code = nt4(base);
        if ((i & 1) == 0) b->seq_pool.p[b->seq_pool.n++] = (uint8_t) (code << 4);
        else b->seq_pool.p[b->seq_pool.n - 1] |= (uint8_t) code;
        b->qual_pool.p[b->qual_pool.n++] = rs->rqual.p[w];
    }
    /* tag */
    int64_t tag0 = b->tag_pool.n;
    if (!c->use_md) {
        for (int64_t i = c0; i < c1;) {
            char t = cols->p[i].type;
            int64_t j = i;
            while (j < c1 && cols->p[j].type == t) j++;
            if (t == '=') {
                VPUSH(b->tag_pool, ':');
                push_num(b, (long) (j - i));
            } else if (t == 'X') {
                for (int64_t k = i; k < j; k++) {
                    VPUSH(b->tag_pool, '*');
                    VPUSH(b->tag_pool, lower(cols->p[k].rf));
                    VPUSH(b->tag_pool, lower(cols->p[k].rd));
                }
            } else if (t == 'I') {
                VPUSH(b->tag_pool, '+');
                for (int64_t k = i; k < j; k++) VPUSH(b->tag_pool, lower(cols->p[k].rd));
            } else {
                VPUSH(b->tag_pool, '-');
                for (int64_t k = i; k < j; k++) VPUSH(b->tag_pool, lower(cols->p[k].rf));
            }
            i = j;
        }
    } else {
        long run = 0;
        int prev_del = 0;
        for (int64_t i = c0; i < c1; i++) {
            char t = cols->p[i].type;
            if (t == 'I') {  /* an insertion splits a deletion run: samtools calmd writes "^A0^C" */
                prev_del = 0;
                continue;
            }
            if (t == '=') {
                run++;
                prev_del = 0;
            } else if (t == 'X') {
                push_num(b, run);
                run = 0;
                VPUSH(b->tag_pool, cols->p[i].rf);
                prev_del = 0;
            } else {
                if (!prev_del) {
                    push_num(b, run);
                    run = 0;
                    VPUSH(b->tag_pool, '^');
                }
                VPUSH(b->tag_pool, cols->p[i].rf);
                prev_del = 1;
            }
        }
        push_num(b, run);
    }
    VPUSH(b->flag, flag);
    VPUSH(b->tid, tid);
    VPUSH(b->pos, (int32_t) pos);
    VPUSH(b->l_qseq, (int32_t) l_qseq);
    VPUSH(b->n_cigar, (int32_t) (b->cigar_pool.n - cig0));
    VPUSH(b->tag_kind, c->use_md ? 1 : 0);
    VPUSH(b->cigar_off, cig0);
    VPUSH(b->tag_off, tag0);
    VPUSH(b->seq_off, seq0);
    VPUSH(b->qual_off, qual0);
}

synth_batch *synth_generate(const synth *s, int64_t first_group, int32_t n_groups) {
    const synth_cfg *c = &s->cfg;
    synth_batch *b = (synth_batch *) calloc(1, sizeof(*b));
    readsim_t rs;
    memset(&rs, 0, sizeof(rs));
    colvec cols = {0};
    VPUSH(b->grp_aln_off, 0);
    VPUSH(b->qname_off, 0);
    for (int64_t g = first_group; g < first_group + n_groups; g++) {
        rng_t r = {(c->seed ^ 0xA5A5A5A55A5A5A5Aull) + 0x9E3779B97F4A7C15ull * (uint64_t) (g + 1)};
        rng_next(&r);
        int locus = (int) rng_below(&r, (uint64_t) c->n_loci);
        int S = (int) rng_below(&r, (uint64_t) c->n_copies);
        int64_t L = c->locus_len;
        double lenf = c->len_mean + c->len_sd * rng_normal(&r);
        if (lenf < c->len_min) lenf = c->len_min;
        if (lenf > c->len_max) lenf = c->len_max;
        int64_t len = (int64_t) lenf;
        if (len > L - 200) len = L - 200;
        int64_t s1 = 100 + (int64_t) rng_below(&r, (uint64_t) (L - len - 199));
        int64_t e1 = s1 + len;
        int read_strand = (int) rng_below(&r, 2);

        rs.rbase.n = rs.rqual.n = rs.op.n = rs.subbase.n = rs.ins_len.n = rs.ins_off.n = rs.ins_pool.n = 0;
        simulate_read(s, locus, S, s1, e1, &r, &rs);

        /* which copies get an alignment; index 0 = primary */
        int n_aln = 1 + (c->max_secondaries < c->n_copies - 1 ? c->max_secondaries : c->n_copies - 1);
        int targets[16];
        int used[16] = {0};
        int P = S;
        if (c->n_copies > 1 && rng_u(&r) < c->wrong_primary_prob) {
            P = (S + 1 + (int) rng_below(&r, (uint64_t) (c->n_copies - 1))) % c->n_copies;
        }
        targets[0] = P;
        used[P] = 1;
        int nt = 1;
        if (P != S && nt < n_aln) { targets[nt++] = S; used[S] = 1; }
        while (nt < n_aln) {
            int t = (int) rng_below(&r, (uint64_t) c->n_copies);
            if (used[t]) continue;
            used[t] = 1;
            targets[nt++] = t;
        }
        /* BAM order: random permutation */
        int order[16];
        for (int i = 0; i < n_aln; i++) order[i] = i;
        for (int i = n_aln - 1; i > 0; i--) {
            int j = (int) rng_below(&r, (uint64_t) (i + 1));
            int t = order[i]; order[i] = order[j]; order[j] = t;
        }
        for (int oi = 0; oi < n_aln; oi++) {
            int ai = order[oi];
            int T = targets[ai];
            int64_t tpos0 = build_columns(s, locus, S, T, s1, e1, &rs, &cols);
            int clipL = 0, clipR = 0, hard = 0;
            if (rng_u(&r) < c->clip_prob) {
                if (rng_below(&r, 2)) clipL = 1 + (int) rng_below(&r, (uint64_t) c->clip_max);
                if (rng_below(&r, 2)) clipR = 1 + (int) rng_below(&r, (uint64_t) c->clip_max);
                hard = (ai != 0) && (rng_u(&r) < c->hard_clip_prob);
            }
            emit_alignment(b, s, locus * c->n_copies + T, ai != 0, read_strand, &cols, tpos0, &rs, clipL, clipR,
                           hard, &r);
        }
        char nm[48];
        int nn = snprintf(nm, sizeof(nm), "read%09ld", (long) g);
        for (int i = 0; i < nn; i++) VPUSH(b->qname_pool, nm[i]);
        VPUSH(b->qname_off, b->qname_pool.n);
        VPUSH(b->grp_aln_off, (int32_t) b->flag.n);
    }
    /* closing offsets */
    VPUSH(b->cigar_off, b->cigar_pool.n);
    VPUSH(b->tag_off, b->tag_pool.n);
    VPUSH(b->seq_off, b->seq_pool.n);
    VPUSH(b->qual_off, b->qual_pool.n);
    free(rs.rbase.p); free(rs.rqual.p); free(rs.op.p); free(rs.subbase.p);
    free(rs.ins_len.p); free(rs.ins_off.p); free(rs.ins_pool.p); free(cols.p);

    sp_flat_batch *v = &b->view;
    v->n_groups = n_groups;
    v->n_alns = (int32_t) b->flag.n;
    v->grp_aln_off = b->grp_aln_off.p;
    v->qname_off = b->qname_off.p;
    v->qname_pool = b->qname_pool.p;
    v->flag = b->flag.p;
    v->tid = b->tid.p;
    v->pos = b->pos.p;
    v->l_qseq = b->l_qseq.p;
    v->n_cigar = b->n_cigar.p;
    v->tag_kind = b->tag_kind.p;
    v->cigar_off = b->cigar_off.p;
    v->tag_off = b->tag_off.p;
    v->seq_off = b->seq_off.p;
    v->qual_off = b->qual_off.p;
    v->cigar_pool = b->cigar_pool.p;
    v->tag_pool = b->tag_pool.p;
    v->seq_pool = b->seq_pool.p;
    v->qual_pool = b->qual_pool.p;
    return b;
}

const sp_flat_batch *synth_batch_view(const synth_batch *b) { return &b->view; }

void synth_batch_free(synth_batch *b) {
    if (!b) return;
    free(b->grp_aln_off.p); free(b->flag.p); free(b->tid.p); free(b->pos.p); free(b->l_qseq.p);
    free(b->n_cigar.p); free(b->tag_kind.p); free(b->qname_off.p); free(b->cigar_off.p); free(b->tag_off.p);
    free(b->seq_off.p); free(b->qual_off.p); free(b->qname_pool.p); free(b->tag_pool.p);
    free(b->cigar_pool.p); free(b->seq_pool.p); free(b->qual_pool.p);
    free(b);
}
