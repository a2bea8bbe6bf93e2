/* oracle/shim/sam.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Stand-in for the slice of htslib 1.17's <htslib/sam.h> that the reference's marker-path
 * sources (programs/submodules/{cigar_it,ptAlignment,ptMarker,ptBlock,common}) touch, so that
 * those files can be compiled UNMODIFIED from /root/reference into oracle/_ref/ (see
 * oracle/Makefile).  Layout and macro semantics follow the public SAM/BAM specification and
 * htslib's documented accessor macros; nothing here is copied from the reference tree.
 */
#ifndef ORACLE_SHIM_SAM_H
#define ORACLE_SHIM_SAM_H
#include <stdint.h>
#include <stddef.h>
#include <limits.h> /* htslib/sam.h pulls this in transitively; ptMarker.c:757 uses INT_MIN */

#ifdef __cplusplus
extern "C" {
#endif

typedef int64_t hts_pos_t;

typedef struct bam1_core_t {
    hts_pos_t pos;
    int32_t tid;
    uint16_t bin;
    uint8_t qual;
    uint8_t l_extranul;
    uint16_t flag;
    uint16_t l_qname;
    uint32_t n_cigar;
    int32_t l_qseq;
    int32_t mtid;
    hts_pos_t mpos;
    hts_pos_t isize;
} bam1_core_t;

typedef struct bam1_t {
    bam1_core_t core;
    uint64_t id;
    uint8_t *data;
    int l_data;
    uint32_t m_data;
    uint32_t mempolicy;
} bam1_t;

#define BAM_CMATCH 0
#define BAM_CINS 1
#define BAM_CDEL 2
#define BAM_CREF_SKIP 3
#define BAM_CSOFT_CLIP 4
#define BAM_CHARD_CLIP 5
#define BAM_CPAD 6
#define BAM_CEQUAL 7
#define BAM_CDIFF 8
#define BAM_CBACK 9

#define BAM_CIGAR_SHIFT 4
#define BAM_CIGAR_MASK 0xf
#define bam_cigar_op(c) ((c) & BAM_CIGAR_MASK)
#define bam_cigar_oplen(c) ((c) >> BAM_CIGAR_SHIFT)

#define BAM_FPAIRED 1
#define BAM_FPROPER_PAIR 2
#define BAM_FUNMAP 4
#define BAM_FMUNMAP 8
#define BAM_FREVERSE 16
#define BAM_FMREVERSE 32
#define BAM_FREAD1 64
#define BAM_FREAD2 128
#define BAM_FSECONDARY 256
#define BAM_FQCFAIL 512
#define BAM_FDUP 1024
#define BAM_FSUPPLEMENTARY 2048

#define bam_is_rev(b) (((b)->core.flag & BAM_FREVERSE) != 0)
#define bam_get_qname(b) ((char *)(b)->data)
#define bam_get_cigar(b) ((uint32_t *)((b)->data + (b)->core.l_qname))
#define bam_get_seq(b) ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname)
#define bam_get_qual(b) ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname + (((b)->core.l_qseq + 1) >> 1))
#define bam_get_aux(b) ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname + (((b)->core.l_qseq + 1) >> 1) + (b)->core.l_qseq)
#define bam_get_l_aux(b) ((b)->l_data - ((b)->core.n_cigar << 2) - (b)->core.l_qname - (b)->core.l_qseq - (((b)->core.l_qseq + 1) >> 1))
#define bam_seqi(s, i) ((s)[(i) >> 1] >> ((~(i) & 1) << 2) & 0xf)

bam1_t *bam_init1(void);
void bam_destroy1(bam1_t *b);
bam1_t *bam_copy1(bam1_t *bdst, const bam1_t *bsrc);
uint8_t *bam_aux_get(const bam1_t *b, const char tag[2]);

/* header: only tid -> name is needed */
typedef struct sam_hdr_t {
    int32_t n_targets;
    char **target_name;
    uint32_t *target_len;
} sam_hdr_t;
const char *sam_hdr_tid2name(const sam_hdr_t *h, int tid);

/* opaque file handle, only referenced by type in tpool.h / secphase.c */
typedef struct samFile samFile;

extern const unsigned char seq_nt16_table[256];
extern const int seq_nt16_int[];

/* htslib declares the BAQ HMM in sam.h */
typedef struct {
    float d, e;
    int bw;
} probaln_par_t;
int probaln_glocal(const uint8_t *ref, int l_ref, const uint8_t *query, int l_query,
                   const uint8_t *iqual, const probaln_par_t *c, int *state, uint8_t *q);

#ifdef __cplusplus
}
#endif
#endif
