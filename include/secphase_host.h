/* include/secphase_host.h -- C ABI of libsecphase_host.so (host side of the drop-in, no CUDA inside)
 *
 * The pieces of the reference's process that surround the marker-mode hot path and that stay on
 * the host (paths relative to /root/reference/):
 *   - BAM ingest + grouping by query name + eligibility filter
 *        parseAlignmentsAndScatterJobs()                  programs/src/secphase.c:230-351
 *        (htslib sam_open/sam_hdr_read/sam_read1, secphase.c:236-237,268; htslib is absent from
 *        this image, so BGZF/BAM are decoded here directly over zlib, multi-threaded)
 *   - FASTA load (fai_load / fai_fetch, secphase.c:101,615; ptMarker.c:736-744) -> 0..4 codes
 *   - the text outputs of a selected secondary: out.log record  secphase.c:194-200,32-57
 *     block tables  ptBlock_add_alignment (ptBlock.c:551-571),
 *                   ptMarker_add_marker_blocks_by_contig (ptMarker.c:851-865),
 *     merge + BED   merge_and_save_blocks (secphase.c:59-72) -> ptBlock_merge_blocks_v2
 *                   (ptBlock.c:274-428) and ptBlock_save_in_bed (ptBlock.c:573-602)
 * The compute itself is behind include/secphase_b200.h; this library never touches a GPU and
 * never computes a score.  The `secphase` executable (secphase_b200/host/secphase_main.cpp) links
 * both libraries.
 *
 * Error behaviour: int functions return 0 (or a count) on success, a negative SPH_E* code on
 * failure; sph_last_error() gives a thread-local message.  Nothing calls exit().
 */
#ifndef SECPHASE_HOST_H
#define SECPHASE_HOST_H
#include <stddef.h>
#include <stdint.h>
#include "sp_flat_batch.h"
#ifdef __cplusplus
extern "C" {
#endif

#define SPH_OK 0
#define SPH_EINVAL (-1)
#define SPH_EIO (-2)      /* open/read/write failure */
#define SPH_EFORMAT (-3)  /* not BGZF/BAM/FASTA, truncated file, CRC mismatch */
#define SPH_ENOMEM (-4)
#define SPH_ENOTAG (-5)   /* a mapped record of an eligible group carries neither cs:Z nor MD:Z
                             (the reference exits here, cigar_it.c:64-67) */

const char *sph_last_error(void);

/* ------------------------------------------------------------------ BAM ingest */
typedef struct sph_bam sph_bam;
typedef struct sph_batch sph_batch;

/* threads = inflate / pack workers (>=1).  Reads and parses the header. */
sph_bam *sph_bam_open(const char *path, int threads);
void sph_bam_close(sph_bam *r);
int32_t sph_bam_n_targets(const sph_bam *r);
const char *sph_bam_target_name(const sph_bam *r, int32_t tid); /* sam_hdr_tid2name */
int64_t sph_bam_target_len(const sph_bam *r, int32_t tid);
const char *sph_bam_header_text(const sph_bam *r, int64_t *len);

/* Sequence length available per BAM tid (0 = contig absent from the FASTA).  Groups with an
 * alignment that leaves [0, len) -- where the reference's fai_fetch would fail its assert
 * (ptMarker.c:741-742) -- are skipped and counted by sph_bam_skipped_groups. */
int sph_bam_set_contig_limits(sph_bam *r, const int64_t *len_by_tid);

/* A reusable host batch whose four pools (cigar/tag/seq/qual) live in memory obtained from
 * `alloc` -- pass sp_host_alloc / sp_host_free (include/secphase_b200.h) to get page-locked
 * pools that sp_submit copies to the device without staging; NULL,NULL = malloc/free. */
sph_batch *sph_batch_create(void *(*alloc)(size_t), void (*release)(void *));
void sph_batch_destroy(sph_batch *b);
const sp_flat_batch *sph_batch_view(const sph_batch *b);
/* per alignment of the batch: 0-based record ordinal in the BAM (for diagnostics) */
const int64_t *sph_batch_record_index(const sph_batch *b);

/* -w/--writeBam support: when on, sph_bam_next_batch also keeps every alignment's whole BAM record
 * body (the bytes after block_size: refID .. aux) so that the record can be written back out
 * (secphase.c:182-189).  sph_batch_records returns the pool and its [n_alns+1] offsets. */
void sph_batch_keep_records(sph_batch *b, int on);
const uint8_t *sph_batch_records(const sph_batch *b, const int64_t **off);

/* Fills `b` with the next eligible read groups in file order: at most max_groups groups and
 * about max_bytes pool bytes (a single group larger than max_bytes is still taken alone).
 * Eligibility as secphase.c:285-288,336-337: consecutive records of one query name, unmapped
 * records skipped, at most 11 kept, then 2..10 alignments, no supplementary, exactly one
 * non-secondary.  Returns the number of groups (0 = end of file) or a negative error. */
int32_t sph_bam_next_batch(sph_bam *r, sph_batch *b, int32_t max_groups, int64_t max_bytes);
/* counters of secphase.c:246-247 after the records consumed so far */
void sph_bam_counts(const sph_bam *r, int64_t *parsed_alignments, int64_t *parsed_reads);

/* Test / benchmark data: writes `b` as a BAM file (BGZF, `level` = zlib level, 0..9) with an
 * @SQ header built from the names/lengths.  Every alignment carries its cs:Z or MD:Z tag. */
int sph_bam_write(const char *path, int32_t n_targets, const char *const *names, const int64_t *lens,
                  const sp_flat_batch *b, int level, int threads);
/* the same, incrementally (files larger than one batch) */
typedef struct sph_bamw sph_bamw;
sph_bamw *sph_bamw_open(const char *path, int32_t n_targets, const char *const *names, const int64_t *lens,
                        int level, int threads);
int sph_bamw_add(sph_bamw *w, const sp_flat_batch *b);
int sph_bamw_close(sph_bamw *w); /* also frees w */
/* groups skipped because the marker path is undefined for them in the reference (CIGAR N/P ops,
 * SEQ/QUAL absent or inconsistent with the CIGAR) */
int64_t sph_bam_skipped_groups(const sph_bam *r);

/* ------------------------------------------------------------------ FASTA */
typedef struct sph_fasta sph_fasta;
sph_fasta *sph_fasta_load(const char *path, int threads);
void sph_fasta_free(sph_fasta *f);
int32_t sph_fasta_n(const sph_fasta *f);
const char *sph_fasta_name(const sph_fasta *f, int32_t i);
int64_t sph_fasta_len(const sph_fasta *f, int32_t i);
/* codes A,C,G,T -> 0..3 (either case), everything else 4 == seq_nt16_int[seq_nt16_table[c]]
 * (ptMarker.c:744); contigs concatenated, offsets in sph_fasta_offsets ([n+1]) */
const uint8_t *sph_fasta_codes(const sph_fasta *f);
const int64_t *sph_fasta_offsets(const sph_fasta *f);
int sph_fasta_write(const char *path, int32_t n, const char *const *names, const char *const *seqs,
                    const int64_t *lens, int line_width);

/* ------------------------------------------------------------------ block tables + BED */
typedef struct sph_blocks sph_blocks; /* stHash contig -> list of (rfs, rfe[, count]) */
sph_blocks *sph_blocks_create(int with_count);
void sph_blocks_destroy(sph_blocks *t);
int sph_blocks_add(sph_blocks *t, const char *contig, int32_t rfs, int32_t rfe); /* count = 1 */
int sph_blocks_add_count(sph_blocks *t, const char *contig, int32_t rfs, int32_t rfe, int32_t count);
/* sort by rfs then ptBlock_merge_blocks_v2 per contig, in place (secphase.c:62-63) */
int sph_blocks_merge_v2(sph_blocks *t);
/* plain union merge ptBlock_merge_blocks (ptBlock.c:238-272), in place */
int sph_blocks_merge(sph_blocks *t);
int64_t sph_blocks_total_length(const sph_blocks *t); /* ptBlock_get_total_length_by_rf */
int64_t sph_blocks_total_number(const sph_blocks *t); /* ptBlock_get_total_number */
/* contigs in strcmp order, "%s\t%d\t%d[\t%d]\n" with end = rfe+1, rows with rfe<rfs skipped */
int sph_blocks_save_bed(const sph_blocks *t, const char *path);
/* flat copy for tests: contig index (strcmp order), rfs, rfe, count per row; returns rows */
int64_t sph_blocks_export(const sph_blocks *t, int32_t *rows4, int64_t max_rows);

/* ------------------------------------------------------------------ out.log */
/* One record of secphase.c:197-200 + print_alignment_scores (secphase.c:32-57), score type
 * MARKER: "#MARKER SCORE\n$\t<qname>\n" then per alignment "<*|@|!>\t%.2f\t%s\t%ld\t%d\n", then
 * "\n".  Writes into buf (cap bytes); returns the length, or the needed length if > cap. */
int64_t sph_format_marker_record(char *buf, int64_t cap, const char *qname, int32_t qname_len, int32_t n_alns,
                                 const int32_t *flag, const double *score, const char *const *contig,
                                 const int32_t *pos, const int32_t *rfe, int32_t best_idx);

/* ------------------------------------------------------------------ -w/--writeBam output */
/* One alignment line as htslib's sam_format1 prints it (the reference opens the output with
 * sam_open(path, "w"), secphase.c:651 -- mode "w" without 'b' is SAM *text*, whatever the file is
 * called): QNAME FLAG RNAME POS MAPQ CIGAR RNEXT PNEXT TLEN SEQ QUAL [TAG:TYPE:VALUE...] '\n'.
 * rec/rec_len = BAM record body (after block_size).  qual_override (l_seq bytes, raw Phred) replaces
 * the record's own QUAL when not NULL.  Returns the line length, or the needed length if > cap;
 * negative on a malformed record. */
int64_t sph_format_sam_record(char *buf, int64_t cap, const uint8_t *rec, int64_t rec_len, int32_t n_targets,
                              const char *const *target_names, const uint8_t *qual_override);
/* <outDir>/<prefix>.quality_modified.out.bam of secphase.c:643-657: header text of the input
 * (sam_hdr_write), then the records of every scored read group with the qualities
 * calc_update_baq_all left (sp_result.baq_qual).  The batch must have been read with
 * sph_batch_keep_records on. */
typedef struct sph_samw sph_samw;
sph_samw *sph_samw_open(const char *path, const sph_bam *header_from);
/* the same with `threads` workers formatting the records of each batch in parallel (output unchanged) */
sph_samw *sph_samw_open_mt(const char *path, const sph_bam *header_from, int threads);
int sph_samw_write_batch(sph_samw *w, const sph_batch *b, const uint8_t *baq_qual);
int sph_samw_close(sph_samw *w); /* also frees w */

#ifdef __cplusplus
}
#endif
#endif
