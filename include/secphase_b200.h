/* include/secphase_b200.h -- C ABI of libsecphase_b200.so
 *
 * B200 (sm_100a) implementation of Secphase's marker-mode scoring of read groups.
 *
 * What it replaces in the reference (paths relative to /root/reference/):
 *   - the per-read-group job  runOneThread(), marker branch   programs/src/secphase.c:156-219
 *     handed to the thread pool by  tpool_add_work(tm, runOneThread, arg_cp)   secphase.c:303
 *     with the parameter bag  work_arg_t   programs/submodules/tpool/tpool.h:26-55;
 *   - everything that job calls: ptCigarIt_* (cigar_it.c), ptAlignment_init_coordinates
 *     (ptAlignment.c:42-95), the ptMarker_* pipeline (ptMarker.c), htslib's probaln_glocal()
 *     (call site ptMarker.c:754-757) and get_best_record_index (ptAlignment.c:137-177).
 * The reference has no FFI/plugin API; this ABI is the batch form of that one job type: many
 * read groups per call instead of one job per group, plain pointers and sizes only.
 *
 * Threading: one sp_ctx per GPU, driven by one host thread.  Results come back in submission
 * order so that the caller can emit out.log deterministically; the tie-break random draws of
 * get_best_record_index (ptAlignment.c:156-171: glibc rand(), never seeded) are replayed on the
 * host in that order (sp_wait), bit-compatible with glibc's TYPE_3 generator seeded with 1.
 *
 * Error behaviour: every int function returns 0 on success or a negative SP_E* code; nothing
 * calls exit().  sp_last_error() returns a thread-local message.  There is NO CPU fallback: all
 * entry points that compute fail with SP_ENODEV when no CUDA device is usable.
 */
#ifndef SECPHASE_B200_H
#define SECPHASE_B200_H
#include <stdint.h>
#include "sp_flat_batch.h"
#ifdef __cplusplus
extern "C" {
#endif

#define SP_OK 0
#define SP_EINVAL (-1)   /* bad argument / malformed batch */
#define SP_ENODEV (-2)   /* no usable CUDA device */
#define SP_ECUDA (-3)    /* CUDA runtime error (message in sp_last_error) */
#define SP_ENOMEM (-4)
#define SP_ESTATE (-5)   /* call sequence error (e.g. wait on an idle slot) */
#define SP_EUNSUPPORTED (-6) /* CIGAR N/P ops (undefined in the reference, cigar_it.c:225-291) */
#define SP_ECAPACITY (-7) /* an internal per-group bound was exceeded even after the retry */

#define SP_MAX_ALN_PER_GROUP 10 /* secphase.c:286 */
#define SP_N_SLOTS 3            /* batches in flight per context */

/* Scalars of work_arg_t (tpool.h:26-55), same names and meaning. */
typedef struct sp_params {
    int32_t baq_flag;        /* -q */
    int32_t consensus;       /* -c */
    int32_t indel_threshold; /* -t */
    int32_t min_q;           /* -m */
    int32_t min_score;       /* -n */
    int32_t set_q;           /* -s */
    int32_t flank_margin;    /* -F (default 500) */
    double prim_margin_score;  /* -p */
    double prim_margin_random; /* -r */
    double conf_d;             /* -d */
    double conf_e;             /* -e */
    double conf_b;             /* -b */
} sp_params;

/* defaults of secphase.c:420-449 */
void sp_params_default(sp_params *p);
/* presets of secphase.c:477-504; name is "hifi" or "ont"; overwrites the preset fields only */
int sp_params_preset(sp_params *p, const char *name);

typedef struct sp_ctx sp_ctx;

sp_ctx *sp_create(const sp_params *p, int cuda_device); /* NULL on error */
void sp_destroy(sp_ctx *ctx);
const char *sp_last_error(void);
const char *sp_version(void);

/* Reference assembly, once per context; copied to HBM as 1 byte/base codes A,C,G,T -> 0..3,
 * everything else -> 4 (seq_nt16_int[seq_nt16_table[c]], ptMarker.c:744).  Replaces the
 * per-job fai_load + per-block fai_fetch (secphase.c:101, ptMarker.c:736-744).
 * Contig order must be the BAM header's tid order. */
int sp_set_reference_ascii(sp_ctx *ctx, int32_t n_contigs, const char *const *seqs, const int64_t *lens);
int sp_set_reference_codes(sp_ctx *ctx, int32_t n_contigs, const uint8_t *codes, const int64_t *contig_off);

#define SP_MARKER_W 6 /* alignment_idx, read_pos_f, base_idx, base_q, is_match, ref_pos (ptMarker.h:34-41) */
#define SP_BLOCK_W 6  /* rfs, rfe, sqs, sqe, rds_f, rde_f (ptBlock.h:23-38) */
#define SP_GROUP_W 10 /* best_idx, prim_idx, n_init, n_after_allmm, n_filled, n_after_ins, margin_eff,
                         n_consensus_blocks, n_final, scored */
#define SP_HMM_W 8    /* global alignment index, l_ref, l_query, par_bw, block index, first row slot,
                         n_rows, reserved */

typedef struct sp_result {
    int32_t n_groups;
    int32_t n_alns;
    const int32_t *group;      /* [n_groups][SP_GROUP_W] */
    const double *score;       /* [n_alns]  alignment->score after calc_alignment_score */
    const int32_t *extent;     /* [n_alns][4] rfs, rfe, rds_f, rde_f (ptAlignment.h:26-36) */
    const int64_t *marker_off; /* [n_groups+1] */
    const int32_t *marker;     /* final markers, [marker_off[n_groups]][SP_MARKER_W] */
    /* work statistics of this batch */
    int64_t hmm_instances;
    int64_t hmm_cells;         /* band cells, SURVEY.md 8(d) definition */
    int64_t h2d_bytes, d2h_bytes;
    float ms_total;            /* CUDA-event time H2D start -> D2H end on the slot's stream */
    float ms_hmm;              /* CUDA-event time of the BAQ-HMM kernels only */
    float ms_stage[8];         /* h2d, walk, group (markers+blocks+count), emit+sort, hmm, score, d2h, 0 */
    int32_t gpu_launches;      /* kernels launched for this batch */
    /* --writeBam mode only (sp_set_write_qual), else NULL / 0: the quality arrays of every record as
     * calc_update_baq_all leaves them (ptMarker.c:786, 709-720, 797-806), laid out exactly like the
     * batch's qual_pool / qual_off -- what sam_write1 emits at secphase.c:182-189. */
    const uint8_t *baq_qual;
    int64_t baq_qual_bytes;
    /* BAQ-HMM arithmetic this batch ran with (sp_set_hmm_mode) and, in fast mode, how many HMM instances its
     * guard band sent to the strict kernel */
    int32_t hmm_mode;
    int32_t pad0;
    int64_t hmm_strict_reruns;
} sp_result;

/* Page-locked host memory for the pools of an sp_flat_batch (cigar_pool, tag_pool, seq_pool,
 * qual_pool).  The reference hands each worker a pointer to records the reader thread already
 * holds in memory (secphase.c:303, 338); the equivalent here is that the BAM reader decodes
 * straight into these buffers, from which sp_submit copies to the device without an intermediate
 * host copy.  Pools in ordinary (pageable) memory are accepted too and are staged first.
 * The caller must keep a submitted batch's pools unchanged until sp_wait returns: a page-locked quality pool of a
 * batch whose markers are sparse (HiFi) is not copied at all -- the kernels read the few dozen quality bytes per
 * alignment they need in place, over PCIe (sp_result.h2d_bytes then counts a 32-byte sector per such read instead
 * of the pool; SECPHASE_B200_NO_ZERO_COPY=1 turns it off). */
void *sp_host_alloc(size_t bytes); /* NULL on error */
void sp_host_free(void *p);

/* Asynchronous: stages the per-alignment metadata (and any pageable pool) into the slot's pinned
 * buffer, enqueues the H2D copies and the first kernels (walk, group) on the slot's stream and
 * returns without waiting for the device.  The rest of the batch (HMM tables sized from the
 * device-side totals, HMM kernels, scoring, D2H) is enqueued by the context's launcher thread as
 * soon as those totals arrive, so the caller's thread is free to submit the next batch at once.
 * slot in [0, SP_N_SLOTS). */
int sp_submit(sp_ctx *ctx, const sp_flat_batch *batch, int slot);
/* Blocks until the slot's batch is complete, replays the tie-break RNG for its groups and fills
 * *out; the pointers stay valid until the next sp_submit on the same slot. */
int sp_wait(sp_ctx *ctx, int slot, sp_result *out);
/* Non-blocking: 1 if the slot's batch has finished on the device (sp_wait will not block on
 * the GPU), 0 if it is still running, negative on error.  Lets one host thread keep feeding a
 * device while it retires finished batches (the reference's tpool_wait has no such need: its
 * workers write their own output, secphase.c:194-216). */
int sp_poll(sp_ctx *ctx, int slot);

/* -w/--writeBam (secphase.c:182-189, 643-657): when on, the HMM evaluates the MAP state / quality
 * of every base of every window's write-back range (ptMarker.c:763-786) instead of the marker
 * rows only, and sp_wait returns the modified quality arrays in sp_result.baq_qual.  Scores,
 * markers and the selected alignment are unchanged by the mode.  Costs ~700 B of HBM per window
 * row while a batch is in flight: submit smaller batches (the CLI uses 512 read groups).
 * Must be called with no batch in flight; batches uploaded with sp_upload before the switch have to
 * be uploaded again, and results of completed batches must have been read. */
int sp_set_write_qual(sp_ctx *ctx, int on);

/* BAQ-HMM arithmetic (the replacement of htslib's probaln_glocal, call site ptMarker.c:754-757).
 *   0 "strict": every double operation in the reference's order, no FMA contraction -- the bits of every
 *      intermediate posterior equal the reference's.
 *   1 "fast" (default): FMA-contracted, scale-free evaluation (csrc/sp_hmmf.cuh) -- about half the FP64
 *      instructions.  Intermediate posteriors drift (measured <= 1e-11 relative on 1 - pmax, <= 11 units of
 *      2^-53 absolute); the integers that are consumed (MAP state, q) are protected by a guard band of
 *      1e-9 relative + 2^-47 absolute around every decision threshold and 1e-9 relative between the two best
 *      posteriors: an instance with a consumed row inside a band is recomputed by the strict kernel, so
 *      BAQ values, markers, scores and the selected alignment are the reference's in both modes.
 * --writeBam batches (sp_set_write_qual) always run strict.  SECPHASE_B200_HMM=strict|fast in the environment
 * of sp_create sets the initial mode.  Must be called with no batch in flight. */
int sp_set_hmm_mode(sp_ctx *ctx, int mode);
int sp_get_hmm_mode(sp_ctx *ctx);

/* Device-side stopwatch over several batches (bench.py): sp_mark records a CUDA event on slot 0's
 * stream (call it when the device is idle); sp_elapsed_since_mark gives the CUDA-event time from
 * that mark to the end of the last batch completed (sp_wait) on `slot`.  The span of a pipelined
 * run is the maximum over the slots used. */
int sp_mark(sp_ctx *ctx);
int sp_elapsed_since_mark(sp_ctx *ctx, int slot, float *ms);

/* How the context split the GPU: the latency-bound integer stages of a batch run on *int_sms SMs of
 * their own (a CUDA green context) while the FP64 HMM kernels of the batches ahead of it keep the
 * other *hmm_sms.  Returns 1 when partitioned, 0 when not (both counts are then the whole GPU);
 * SECPHASE_B200_INT_SMS=<n> in the environment of sp_create changes the request (0 = off).
 * Other switches read by sp_create, for tests and measurements (results are the same in every setting):
 *   SECPHASE_B200_WALK=serial|warp|<n>   CIGAR/cs walk: a thread per alignment for all / a warp per alignment for
 *                                        all / a warp for op tables planned at >= n entries (default 600)
 *   SECPHASE_B200_GROUP=serial|lanes     marker merge + consensus blocks: a thread per read group for all / a lane
 *                                        per alignment for all (default: lanes for groups of <= 4 alignments)
 *   SECPHASE_B200_SCORE=serial|lanes     scoring pass likewise (default: lanes)
 *   SECPHASE_B200_NO_PRIORITY=1          emit/sort/score kernels on the slot's ordinary stream */
int sp_sm_partition(sp_ctx *ctx, int32_t *int_sms, int32_t *hmm_sms);

/* --- device-resident variant used to time the kernels alone (bench `value`): the batch is
 * uploaded once with sp_upload, sp_run_resident enqueues kernels only. */
int sp_upload(sp_ctx *ctx, const sp_flat_batch *batch, int slot);
int sp_run_resident(sp_ctx *ctx, int slot);

/* --- intermediate tables of the last completed batch of a slot (parity tests). `what`:
 *   0 markers before BAQ   [n][SP_MARKER_W], off per group
 *   1 markers after BAQ    [n][SP_MARKER_W], off per group
 *   2 consensus blocks     [n][SP_BLOCK_W],  off per alignment
 *   3 HMM instances        [n][SP_HMM_W],    off = NULL
 *   4 HMM marker rows      [n][4] = instance, row t, state, q ; off = NULL
 * Returns the number of rows, or a negative error.  Pointers are host memory owned by ctx.
 * Test switches (rows/off ignored): what = -1 keeps the post-BAQ table of later batches; what = -2 returns
 * how many batches were re-run with the provable table bounds so far (SP_ECAPACITY is only reported when
 * that second run overflows too); what <= -100 clamps the FIRST plan's per-group block workspace to
 * -(what+100) entries so that tests can force that retry; what = -3 returns how many alignments of the slot's
 * last batch the warp-cooperative CIGAR/cs walker left to the serial walker (MD tags, cs text that is not the
 * run structure of the CIGAR; -1 when SECPHASE_B200_WALK=serial made the serial walker the only one). */
int64_t sp_debug_table(sp_ctx *ctx, int slot, int what, const int32_t **rows, const int64_t **off);

/* --- the HMM alone, batch form of probaln_glocal(ref,l_ref,query,l_query,iqual,&conf,state,q)
 * (call site ptMarker.c:755-757) for uniform iqual == set_q.  All pointers are HOST memory.
 * For instance j: ref codes ref_pool[ref_off[j] .. +l_ref[j]), query codes likewise, band
 * parameter conf.bw = par_bw[j].  Only the query rows listed in rows[row_off[j]..row_off[j+1])
 * (0-based t, ascending) are evaluated by the MAP step; state_out/q_out/pmax_out are indexed
 * like rows.  pmax_out (normalised max posterior) may be NULL. */
int sp_hmm_batch(sp_ctx *ctx, int32_t n, const uint8_t *ref_pool, const int64_t *ref_off, const int32_t *l_ref,
                 const uint8_t *query_pool, const int64_t *query_off, const int32_t *l_query,
                 const int32_t *par_bw, const int64_t *row_off, const int32_t *rows, int32_t *state_out,
                 uint8_t *q_out, double *pmax_out, float *ms_kernel);

/* FP64 pipe micro-benchmark for the roofline denominator: register-resident
 * chains of DFMA (mode 0) or alternating DADD/DMUL (mode 1).  Returns instructions/second
 * (per thread-instruction, i.e. lane-ops/s) in *ops_per_s. */
int sp_fp64_peak(sp_ctx *ctx, int mode, double *ops_per_s, float *ms);

/* Emulation of glibc rand() (TYPE_3, seed 1) used for the tie-breaks; exposed for tests. */
void sp_rng_seed(sp_ctx *ctx, unsigned seed);
int sp_rng_next(sp_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif
