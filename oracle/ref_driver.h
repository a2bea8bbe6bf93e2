/* oracle/ref_driver.h -- TEST INFRASTRUCTURE ONLY.
 *
 * C interface of the two CPU checkers:
 *   oracle/_ref/libsecphase_ref.so   the reference's OWN marker-path sources
 *                                    (/root/reference/programs/submodules/{cigar_it,ptAlignment,
 *                                    ptMarker,ptBlock,common}/ *.c, compiled unmodified) + shim +
 *                                    probaln_port.c, driven the way secphase.c:156-219 drives them;
 *   oracle/liboracle_port.so         reserved name for a plain-C restatement of the same path exporting the
 *                                    same entry points (not written; see oracle/pyoracle.py).
 */
#ifndef ORACLE_REF_DRIVER_H
#define ORACLE_REF_DRIVER_H
#include <stdint.h>
#include "../include/sp_flat_batch.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_params { /* work_arg_t scalars, tpool.h:26-55 */
    int32_t baq_flag, consensus, indel_threshold, min_q, min_score, set_q, flank_margin;
    double prim_margin_score, prim_margin_random, conf_d, conf_e, conf_b;
} oracle_params;

typedef struct oracle_refseq { /* assembly as ASCII, one string per BAM tid */
    int32_t n_contigs;
    const char *const *names;
    const char *const *seqs;
    const int64_t *lens;
} oracle_refseq;

#define ORACLE_MARKER_W 6 /* alignment_idx, read_pos_f, base_idx, base_q, is_match, ref_pos */
#define ORACLE_BLOCK_W 6  /* rfs, rfe, sqs, sqe, rds_f, rde_f */
#define ORACLE_HMM_W 8    /* global alignment index, l_ref, l_query, par_bw, cells_lo, cells_hi, fnv(state), fnv(q) */
#define ORACLE_GROUP_W 10 /* best_idx, prim_idx, n_init, n_after_allmm, n_filled, n_after_ins, margin_eff,
                             n_consensus_blocks, n_final, scored */

typedef struct oracle_out oracle_out;

oracle_out *oracle_out_create(int keep_hmm_arrays);
void oracle_out_destroy(oracle_out *o);

/* run every group of `b` in order; rand() state is whatever the process has (see oracle_srand) */
int oracle_run(const sp_flat_batch *b, const oracle_params *p, const oracle_refseq *ref, oracle_out *out);
void oracle_srand(unsigned seed);
/* get_best_record_index (ptAlignment.c:137-177) over scored groups, in order, on the current rand() stream */
int oracle_select(int32_t n_groups, const int32_t *grp_aln_off, const int32_t *flag, const double *score,
                  const oracle_params *p, int32_t *best_out);

/* Text outputs (reference kind only): after oracle_out_enable_outputs, oracle_run also does what
 * secphase.c:194-216 does for every group whose selected alignment is a secondary, with the
 * reference's own ptBlock/ptMarker block functions; oracle_out_save then writes <prefix>.out.log,
 * .modified_read_blocks.markers.bed and .marker_blocks.bed as secphase.c:681-732 does. */
void oracle_out_enable_outputs(oracle_out *o);
int oracle_out_save(oracle_out *o, const char *dir, const char *prefix, int32_t *totals4, int32_t *n_modified);
/* the reference's ptBlock_merge_blocks (mode 0) / ptBlock_merge_blocks_v2 (mode 1) on rows of
 * (start, end, count); count < 0 = block without count data */
int64_t oracle_merge_blocks(const int32_t *rows3, int64_t n, int mode, int32_t *out3, int64_t max_rows);

/* flat result tables; *n receives the number of ROWS */
const int32_t *oracle_out_groups(const oracle_out *o, int64_t *n);                 /* [n][ORACLE_GROUP_W] */
const double *oracle_out_scores(const oracle_out *o, int64_t *n);                  /* [n_alns] */
const int32_t *oracle_out_extents(const oracle_out *o, int64_t *n);                /* [n_alns][4] rfs,rfe,rds_f,rde_f */
const int32_t *oracle_out_markers(const oracle_out *o, int stage, int64_t *n);     /* stage 0 pre-BAQ, 1 post-BAQ, 2 final */
const int64_t *oracle_out_marker_off(const oracle_out *o, int stage, int64_t *n);  /* [n_groups+1] */
const int32_t *oracle_out_blocks(const oracle_out *o, int64_t *n);                 /* [n][ORACLE_BLOCK_W] */
const int64_t *oracle_out_block_off(const oracle_out *o, int64_t *n);              /* [n_alns+1] */
const int32_t *oracle_out_hmm(const oracle_out *o, int64_t *n);                    /* [n][ORACLE_HMM_W] */
const int32_t *oracle_out_hmm_state(const oracle_out *o, int64_t *n);              /* concatenated state[] */
const uint8_t *oracle_out_hmm_q(const oracle_out *o, int64_t *n);                  /* concatenated q[] */
const uint8_t *oracle_out_qual(const oracle_out *o, int64_t *n);                   /* records' final quality arrays, batch qual_pool layout */
const char *oracle_kind(void); /* "reference" or "port" */

#ifdef __cplusplus
}
#endif
#endif
