/* include/sp_flat_batch.h -- plain-data description of a set of read groups in host memory.
 *
 * A "read group" is what the reference hands to one worker job: all alignment records of one
 * query name, one primary plus its secondaries (secphase.c:279-303, work_arg_t.alignments in
 * tpool.h:26-55).  Fields mirror the BAM record members the marker path reads through htslib
 * macros (core.flag/tid/pos/l_qseq/n_cigar, bam_get_cigar/seq/qual, aux cs:Z or MD:Z).
 * This header only describes data; it is shared by the product library, the synthetic-data
 * generator and the test oracle's driver.
 */
#ifndef SP_FLAT_BATCH_H
#define SP_FLAT_BATCH_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct sp_flat_batch {
    int32_t n_groups;
    int32_t n_alns;
    const int32_t *grp_aln_off; /* [n_groups+1] first alignment of each group (BAM order inside a group) */
    const int64_t *qname_off;   /* [n_groups+1] into qname_pool (names are NOT NUL-terminated) */
    const char *qname_pool;
    /* per alignment, [n_alns] */
    const int32_t *flag;    /* BAM FLAG (0x10 reverse, 0x100 secondary, 0x800 supplementary) */
    const int32_t *tid;     /* reference id, index into the contig table given to sp_set_reference */
    const int32_t *pos;     /* 0-based leftmost reference coordinate */
    const int32_t *l_qseq;  /* stored SEQ length (hard clips excluded) */
    const int32_t *n_cigar;
    const int32_t *tag_kind; /* 0 = cs:Z (short form), 1 = MD:Z ; may be NULL => all cs */
    /* per alignment offsets, [n_alns+1], into the pools below */
    const int64_t *cigar_off; /* in uint32 units */
    const int64_t *tag_off;   /* in bytes; tag text without the leading 'Z', no NUL needed */
    const int64_t *seq_off;   /* in bytes; 4-bit packed, high nibble first, (l_qseq+1)/2 bytes */
    const int64_t *qual_off;  /* in bytes; raw Phred, l_qseq bytes */
    const uint32_t *cigar_pool;
    const char *tag_pool;
    const uint8_t *seq_pool;
    const uint8_t *qual_pool;
} sp_flat_batch;

#ifdef __cplusplus
}
#endif
#endif
