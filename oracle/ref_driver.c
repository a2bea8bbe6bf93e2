/* oracle/ref_driver.c -- TEST INFRASTRUCTURE ONLY (checker / reported CPU baseline).
 *
 * Drives the reference's OWN marker-path functions (compiled unmodified from /root/reference by
 * oracle/Makefile into oracle/_ref/) over a flat batch of read groups, the way one worker job
 * does in the reference: secphase.c:156-219 (marker branch of runOneThread) after
 * ptAlignment_construct (secphase.c:338).  Records every intermediate the CUDA path has to
 * match: marker lists per stage, consensus blocks, every HMM instance, scores, selected index.
 *
 * Quirk handling ("Quirks"): Q1 conf_blocks_length is read uninitialised
 * by the reference when the x0.8 loop never runs (secphase.c:107,170); it is defined here as 1.
 */
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ptAlignment.h"
#include "ptBlock.h"
#include "ptMarker.h"

#include "probaln_port.h"
#include "ref_driver.h"

/* ---------------------------------------------------------------- growable tables */
typedef struct { void *p; int64_t n, cap; size_t esz; } vec;
static void vec_init(vec *v, size_t esz) { v->p = 0; v->n = 0; v->cap = 0; v->esz = esz; }
static void *vec_push(vec *v, int64_t cnt) {
    if (v->n + cnt > v->cap) {
        int64_t nc = v->cap ? v->cap * 2 : 1024;
        while (nc < v->n + cnt) nc *= 2;
        v->p = realloc(v->p, (size_t) nc * v->esz);
        v->cap = nc;
    }
    void *r = (char *) v->p + (size_t) v->n * v->esz;
    v->n += cnt;
    return r;
}

struct oracle_out {
    int keep_hmm;
    /* optional text-output sink (oracle_out_enable_outputs): what secphase.c:194-216 does when the
     * selected alignment is a secondary, using the reference's own block functions */
    int sink;
    vec log;                 /* bytes of <prefix>.out.log */
    stHash *mod_blocks;      /* modified_blocks_by_marker_per_contig */
    stHash *marker_blocks;   /* marker_blocks_all_haps_per_contig */
    int reads_modified_by_marker;
    vec groups, scores, extents, markers[3], marker_off[3], blocks, block_off, hmm, hmm_state, hmm_q;
    vec qual;                /* every record's quality array as the job leaves it (what sam_write1 emits, secphase.c:182-189) */
    int32_t cur_aln_global; /* for the HMM trace */
};

oracle_out *oracle_out_create(int keep_hmm_arrays) {
    oracle_out *o = (oracle_out *) calloc(1, sizeof(*o));
    o->keep_hmm = keep_hmm_arrays;
    vec_init(&o->groups, sizeof(int32_t));
    vec_init(&o->scores, sizeof(double));
    vec_init(&o->extents, sizeof(int32_t));
    for (int s = 0; s < 3; s++) {
        vec_init(&o->markers[s], sizeof(int32_t));
        vec_init(&o->marker_off[s], sizeof(int64_t));
    }
    vec_init(&o->blocks, sizeof(int32_t));
    vec_init(&o->block_off, sizeof(int64_t));
    vec_init(&o->hmm, sizeof(int32_t));
    vec_init(&o->hmm_state, sizeof(int32_t));
    vec_init(&o->hmm_q, sizeof(uint8_t));
    vec_init(&o->qual, sizeof(uint8_t));
    return o;
}

void oracle_out_enable_outputs(oracle_out *o) {
    if (o->sink) return;
    o->sink = 1;
    vec_init(&o->log, 1);
    o->mod_blocks = stHash_construct3(stHash_stringKey, stHash_stringEqualKey, NULL, (void (*)(void *)) stList_destruct);
    o->marker_blocks = stHash_construct3(stHash_stringKey, stHash_stringEqualKey, NULL, (void (*)(void *)) stList_destruct);
}

static void log_printf(oracle_out *o, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    int n = vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    memcpy(vec_push(&o->log, n), buf, (size_t) n);
}

/* merge_and_save_blocks, secphase.c:59-72, with the reference's own ptBlock functions */
static int save_blocks(stHash *blocks, const char *bed_path, int print_count, int *total_len, int *total_n) {
    ptBlock_sort_stHash_by_rfs(blocks);
    stHash *merged = ptBlock_merge_blocks_per_contig_by_rf_v2(blocks);
    *total_len = ptBlock_get_total_length_by_rf(merged);
    *total_n = ptBlock_get_total_number(merged);
    ptBlock_save_in_bed(merged, (char *) bed_path, print_count != 0);
    stHash_destruct(merged);
    return 0;
}

/* Writes <dir>/<prefix>.out.log, .modified_read_blocks.markers.bed and .marker_blocks.bed the way
 * secphase.c:681-732 does; totals[4] = length/number of the two merged tables. */
int oracle_out_save(oracle_out *o, const char *dir, const char *prefix, int32_t *totals, int32_t *n_modified) {
    if (!o->sink) return -1;
    char path[2048];
    snprintf(path, sizeof(path), "%s/%s.out.log", dir, prefix);
    FILE *fp = fopen(path, "w");
    if (!fp) return -2;
    if (o->log.n) fwrite(o->log.p, 1, (size_t) o->log.n, fp);
    fclose(fp);
    int tl, tn;
    snprintf(path, sizeof(path), "%s/%s.modified_read_blocks.markers.bed", dir, prefix);
    save_blocks(o->mod_blocks, path, 1, &tl, &tn);
    if (totals) { totals[0] = tl; totals[1] = tn; }
    snprintf(path, sizeof(path), "%s/%s.marker_blocks.bed", dir, prefix);
    save_blocks(o->marker_blocks, path, 0, &tl, &tn);
    if (totals) { totals[2] = tl; totals[3] = tn; }
    if (n_modified) *n_modified = o->reads_modified_by_marker;
    return 0;
}

/* The reference's block merges on caller-supplied intervals (rows of start, end, count):
 * mode 0 = ptBlock_merge_blocks (union), 1 = ptBlock_merge_blocks_v2; count < 0 = no count data.
 * Returns the number of merged rows (written to out3 up to max_rows). */
int64_t oracle_merge_blocks(const int32_t *rows3, int64_t n, int mode, int32_t *out3, int64_t max_rows) {
    stHash *h = stHash_construct3(stHash_stringKey, stHash_stringEqualKey, NULL, (void (*)(void *)) stList_destruct);
    stList *l = stList_construct3(0, ptBlock_destruct);
    for (int64_t i = 0; i < n; i++) {
        const int32_t *r = rows3 + 3 * i;
        stList_append(l, r[2] < 0 ? ptBlock_construct(r[0], r[1], -1, -1, -1, -1)
                                  : ptBlock_construct_with_count(r[0], r[1], -1, -1, -1, -1, r[2]));
    }
    stHash_insert(h, copyString("ctg"), l);
    ptBlock_sort_stHash_by_rfs(h);
    stHash *m = mode == 0 ? ptBlock_merge_blocks_per_contig_by_rf(h) : ptBlock_merge_blocks_per_contig_by_rf_v2(h);
    stList *ml = (stList *) stHash_search(m, "ctg");
    int64_t k = stList_length(ml);
    for (int64_t i = 0; i < k && i < max_rows; i++) {
        ptBlock *b = (ptBlock *) stList_get(ml, i);
        out3[3 * i] = b->rfs;
        out3[3 * i + 1] = b->rfe;
        out3[3 * i + 2] = b->data ? ptBlock_get_count(b) : -1;
    }
    stHash_destruct(m);
    stHash_destruct(h);
    return k;
}

void oracle_out_destroy(oracle_out *o) {
    if (!o) return;
    if (o->sink) {
        free(o->log.p);
        stHash_destruct(o->mod_blocks);
        stHash_destruct(o->marker_blocks);
    }
    free(o->groups.p); free(o->scores.p); free(o->extents.p);
    for (int s = 0; s < 3; s++) { free(o->markers[s].p); free(o->marker_off[s].p); }
    free(o->blocks.p); free(o->block_off.p); free(o->hmm.p); free(o->hmm_state.p); free(o->hmm_q.p); free(o->qual.p);
    free(o);
}

#define GETTER(name, field, type, width)                                  \
    const type *name(const oracle_out *o, int64_t *n) {                   \
        if (n) *n = o->field.n / (width);                                 \
        return (const type *) o->field.p;                                 \
    }
GETTER(oracle_out_groups, groups, int32_t, ORACLE_GROUP_W)
GETTER(oracle_out_scores, scores, double, 1)
GETTER(oracle_out_extents, extents, int32_t, 4)
GETTER(oracle_out_blocks, blocks, int32_t, ORACLE_BLOCK_W)
GETTER(oracle_out_block_off, block_off, int64_t, 1)
GETTER(oracle_out_hmm, hmm, int32_t, ORACLE_HMM_W)
GETTER(oracle_out_hmm_state, hmm_state, int32_t, 1)
GETTER(oracle_out_hmm_q, hmm_q, uint8_t, 1)
GETTER(oracle_out_qual, qual, uint8_t, 1)
const int32_t *oracle_out_markers(const oracle_out *o, int stage, int64_t *n) {
    if (n) *n = o->markers[stage].n / ORACLE_MARKER_W;
    return (const int32_t *) o->markers[stage].p;
}
const int64_t *oracle_out_marker_off(const oracle_out *o, int stage, int64_t *n) {
    if (n) *n = o->marker_off[stage].n;
    return (const int64_t *) o->marker_off[stage].p;
}
const char *oracle_kind(void) { return "reference"; }
void oracle_srand(unsigned seed) { srand(seed); }

/* The reference's own get_best_record_index (ptAlignment.c:137-177) replayed over already-scored
 * groups in group order with the process's rand() stream: lets a multi-threaded oracle run (whose
 * rand() draws interleave arbitrarily) be given the selection a single-worker run would make. */
int oracle_select(int32_t n_groups, const int32_t *grp_aln_off, const int32_t *flag, const double *score,
                  const oracle_params *p, int32_t *best_out) {
    ptAlignment store[16];
    bam1_t recs[16];
    ptAlignment *alns[16];
    for (int32_t g = 0; g < n_groups; g++) {
        const int a0 = grp_aln_off[g], n = grp_aln_off[g + 1] - a0;
        if (n < 1 || n > 16) return -1;
        for (int i = 0; i < n; i++) {
            memset(&store[i], 0, sizeof(store[i]));
            memset(&recs[i], 0, sizeof(recs[i]));
            recs[i].core.flag = (uint16_t) flag[a0 + i];
            store[i].record = &recs[i];
            store[i].score = score[a0 + i];
            alns[i] = &store[i];
        }
        best_out[g] = get_best_record_index(alns, n, p->prim_margin_score, (double) p->min_score, p->prim_margin_random);
    }
    return 0;
}

/* ---------------------------------------------------------------- helpers */
static uint32_t fnv32(const void *p, size_t n) {
    const uint8_t *b = (const uint8_t *) p;
    uint32_t h = 2166136261u;
    for (size_t i = 0; i < n; i++) h = (h ^ b[i]) * 16777619u;
    return h;
}

static void hmm_trace(void *ud, const uint8_t *ref, int l_ref, const uint8_t *query, int l_query,
                      const uint8_t *iqual, float d, float e, int bw, const int *state, const uint8_t *q) {
    (void) ref; (void) query; (void) iqual; (void) d; (void) e;
    oracle_out *o = (oracle_out *) ud;
    int32_t *row = (int32_t *) vec_push(&o->hmm, ORACLE_HMM_W);
    long cells = oracle_probaln_cells(l_ref, l_query, bw);
    row[0] = o->cur_aln_global;
    row[1] = l_ref;
    row[2] = l_query;
    row[3] = bw;
    row[4] = (int32_t) (cells & 0x7fffffff);
    row[5] = (int32_t) (cells >> 31);
    row[6] = (int32_t) fnv32(state, sizeof(int) * (size_t) l_query);
    row[7] = (int32_t) fnv32(q, (size_t) l_query);
    if (o->keep_hmm) {
        memcpy(vec_push(&o->hmm_state, l_query), state, sizeof(int) * (size_t) l_query);
        memcpy(vec_push(&o->hmm_q, l_query), q, (size_t) l_query);
    }
}

static void record_markers(oracle_out *o, int stage, stList *markers) {
    int64_t n = markers ? stList_length(markers) : 0;
    for (int64_t i = 0; i < n; i++) {
        ptMarker *m = (ptMarker *) stList_get(markers, i);
        int32_t *row = (int32_t *) vec_push(&o->markers[stage], ORACLE_MARKER_W);
        row[0] = m->alignment_idx;
        row[1] = m->read_pos_f;
        row[2] = m->base_idx;
        row[3] = m->base_q;
        row[4] = m->is_match ? 1 : 0;
        row[5] = m->ref_pos;
    }
    *(int64_t *) vec_push(&o->marker_off[stage], 1) = o->markers[stage].n / ORACLE_MARKER_W;
}

static bam1_t *make_record(const sp_flat_batch *b, int g, int a) {
    int64_t qn_len = b->qname_off[g + 1] - b->qname_off[g];
    int l_qname = (int) qn_len + 1;
    int extranul = 0;
    while ((l_qname + extranul) % 4) extranul++; /* htslib keeps the CIGAR 4-byte aligned */
    l_qname += extranul;
    int n_cigar = b->n_cigar[a], l_qseq = b->l_qseq[a];
    int64_t tag_len = b->tag_off[a + 1] - b->tag_off[a];
    int kind = b->tag_kind ? b->tag_kind[a] : 0;
    size_t l_data = (size_t) l_qname + 4u * (size_t) n_cigar + (size_t) ((l_qseq + 1) / 2) + (size_t) l_qseq +
                    3 + (size_t) tag_len + 1;
    bam1_t *r = bam_init1();
    r->data = (uint8_t *) calloc(l_data, 1);
    r->l_data = (int) l_data;
    r->m_data = (uint32_t) l_data;
    r->core.pos = b->pos[a];
    r->core.tid = b->tid[a];
    r->core.flag = (uint16_t) b->flag[a];
    r->core.l_qname = (uint16_t) l_qname;
    r->core.l_extranul = (uint8_t) extranul;
    r->core.n_cigar = (uint32_t) n_cigar;
    r->core.l_qseq = l_qseq;
    uint8_t *p = r->data;
    memcpy(p, b->qname_pool + b->qname_off[g], (size_t) qn_len);
    p += l_qname;
    memcpy(p, b->cigar_pool + b->cigar_off[a], 4u * (size_t) n_cigar);
    p += 4u * (size_t) n_cigar;
    memcpy(p, b->seq_pool + b->seq_off[a], (size_t) ((l_qseq + 1) / 2));
    p += (l_qseq + 1) / 2;
    memcpy(p, b->qual_pool + b->qual_off[a], (size_t) l_qseq);
    p += l_qseq;
    p[0] = kind == 0 ? 'c' : 'M';
    p[1] = kind == 0 ? 's' : 'D';
    p[2] = 'Z';
    memcpy(p + 3, b->tag_pool + b->tag_off[a], (size_t) tag_len);
    return r;
}

/* ---------------------------------------------------------------- the job */
int oracle_run(const sp_flat_batch *b, const oracle_params *p, const oracle_refseq *ref, oracle_out *out) {
    sam_hdr_t hdr;
    hdr.n_targets = ref->n_contigs;
    hdr.target_name = (char **) ref->names;
    hdr.target_len = NULL;
    faidx_t fai;
    fai.n = ref->n_contigs;
    fai.names = (char **) ref->names;
    fai.seqs = (const char **) ref->seqs;
    fai.lens = (long *) ref->lens;

    oracle_probaln_set_trace(hmm_trace, out);
    if (out->block_off.n == 0) *(int64_t *) vec_push(&out->block_off, 1) = 0;
    for (int s = 0; s < 3; s++)
        if (out->marker_off[s].n == 0) *(int64_t *) vec_push(&out->marker_off[s], 1) = 0;

    for (int g = 0; g < b->n_groups; g++) {
        int a0 = b->grp_aln_off[g], n = b->grp_aln_off[g + 1] - a0;
        ptAlignment **alns = (ptAlignment **) malloc(sizeof(ptAlignment *) * (size_t) (n > 0 ? n : 1));
        for (int i = 0; i < n; i++) {
            bam1_t *rec = make_record(b, g, a0 + i);
            alns[i] = ptAlignment_construct(rec, &hdr); /* copies the record (ptAlignment.c:30-40) */
            bam_destroy1(rec);
        }
        int32_t *grow = (int32_t *) vec_push(&out->groups, ORACLE_GROUP_W);
        memset(grow, 0, sizeof(int32_t) * ORACLE_GROUP_W);

        /* ---- marker branch, secphase.c:157-180 ---- */
        stList *markers = ptMarker_get_initial_markers(alns, n, p->min_q);
        grow[2] = (int32_t) stList_length(markers);
        remove_all_mismatch_markers(&markers, n);
        grow[3] = (int32_t) stList_length(markers);
        sort_and_fill_markers(&markers, alns, n);
        grow[4] = (int32_t) stList_length(markers);
        filter_ins_markers(&markers, alns, n);
        grow[5] = (int32_t) stList_length(markers);
        record_markers(out, 0, markers);
        int conf_blocks_length = 1; /* Q1 */
        int margin_eff = p->flank_margin;
        int scored = 0;
        if (markers && stList_length(markers) > 0) {
            set_confident_blocks(alns, n, p->indel_threshold);
            while (p->consensus && needs_to_find_blocks(alns, n, 1000, &hdr)) {
                margin_eff *= 0.8;
                set_flanking_blocks(alns, n, markers, margin_eff);
                conf_blocks_length = correct_conf_blocks(alns, n, p->indel_threshold);
                if (conf_blocks_length == 0) break;
            }
            if (conf_blocks_length > 0 || !p->consensus) {
                if (p->baq_flag) {
                    /* calc_update_baq_all (ptMarker.c:811-831), unrolled so that each HMM call can
                     * be attributed to its alignment in the trace */
                    for (int i = 0; i < n; i++) {
                        out->cur_aln_global = a0 + i;
                        const char *ctg = sam_hdr_tid2name(&hdr, alns[i]->record->core.tid);
                        calc_local_baq(&fai, ctg, alns[i], i, markers, p->conf_d, p->conf_e, p->conf_b, p->set_q);
                    }
                    for (int64_t i = 0; i < stList_length(markers); i++) {
                        ptMarker *m = (ptMarker *) stList_get(markers, i);
                        m->base_q = bam_get_qual(alns[m->alignment_idx]->record)[m->base_idx];
                    }
                }
                record_markers(out, 1, markers);
                filter_lowq_markers(&markers, p->min_q);
                calc_alignment_score(markers, alns);
                scored = 1;
            }
        }
        if (!scored) record_markers(out, 1, markers);
        record_markers(out, 2, markers);
        grow[6] = margin_eff;
        grow[7] = conf_blocks_length;
        grow[8] = (int32_t) stList_length(markers);
        grow[9] = scored;

        /* blocks as left on the alignments */
        for (int i = 0; i < n; i++) {
            stList *bl = alns[i]->conf_blocks;
            int64_t nb = bl ? stList_length(bl) : 0;
            for (int64_t k = 0; k < nb; k++) {
                ptBlock *blk = (ptBlock *) stList_get(bl, k);
                int32_t *row = (int32_t *) vec_push(&out->blocks, ORACLE_BLOCK_W);
                row[0] = blk->rfs; row[1] = blk->rfe; row[2] = blk->sqs;
                row[3] = blk->sqe; row[4] = blk->rds_f; row[5] = blk->rde_f;
            }
            *(int64_t *) vec_push(&out->block_off, 1) = out->blocks.n / ORACLE_BLOCK_W;
        }

        /* ---- selection, secphase.c:191-193 ---- */
        int best = n > 0 ? get_best_record_index(alns, n, p->prim_margin_score, (double) p->min_score,
                                                 p->prim_margin_random)
                         : -1;
        grow[0] = best;
        grow[1] = n > 0 ? get_primary_index(alns, n) : -1;
        if (out->sink && best >= 0 && (alns[best]->record->core.flag & BAM_FSECONDARY)) {
            /* secphase.c:194-216; print_alignment_scores (secphase.c:32-57), SCORE_TYPE_MARKER */
            log_printf(out, "#MARKER SCORE\n");
            log_printf(out, "$\t%s\n", bam_get_qname(alns[0]->record));
            for (int i = 0; i < n; i++) {
                if ((alns[i]->record->core.flag & BAM_FSECONDARY) == 0) log_printf(out, "*\t");
                else if (i == best) log_printf(out, "@\t");
                else log_printf(out, "!\t");
                log_printf(out, "%.2f\t%s\t%ld\t%d\n", alns[i]->score, alns[i]->contig,
                           (long) alns[i]->record->core.pos, alns[i]->rfe);
            }
            log_printf(out, "\n");
            int primary_idx = get_primary_index(alns, n);
            ptBlock_add_alignment(out->mod_blocks, alns[primary_idx], true);
            ptBlock_add_alignment(out->mod_blocks, alns[best], true);
            ptMarker_add_marker_blocks_by_contig(out->marker_blocks, alns[primary_idx]->contig, primary_idx, markers);
            ptMarker_add_marker_blocks_by_contig(out->marker_blocks, alns[best]->contig, best, markers);
            out->reads_modified_by_marker += 1;
        }
        for (int i = 0; i < n; i++) {
            *(double *) vec_push(&out->scores, 1) = alns[i]->score;
            int32_t *e = (int32_t *) vec_push(&out->extents, 4);
            e[0] = alns[i]->rfs; e[1] = alns[i]->rfe; e[2] = alns[i]->rds_f; e[3] = alns[i]->rde_f;
            /* the qualities the record carries when secphase.c:182-189 writes it with -w */
            int lq = alns[i]->record->core.l_qseq;
            if (lq > 0) memcpy(vec_push(&out->qual, lq), bam_get_qual(alns[i]->record), (size_t) lq);
        }
        stList_destruct(markers);
        for (int i = 0; i < n; i++) ptAlignment_destruct(alns[i]);
        free(alns);
    }
    oracle_probaln_set_trace(0, 0);
    return 0;
}
