/* tools/synth/synth.h -- synthetic diploid (or k-copy) assembly + simulated long-read
 * alignment groups with KNOWN CIGAR and cs tags.
 *
 * Test / benchmark data only.  There is no network and no real BAM in this image, so every
 * input the hot path sees is produced here, deterministically from a seed, directly in the
 * flat read-group layout of include/sp_flat_batch.h (the same fields a BAM record carries).
 * Group `g` depends only on (seed, g), so any rank can generate its own shard.
 */
#ifndef SP_SYNTH_H
#define SP_SYNTH_H
#include <stdint.h>
#include "../../include/sp_flat_batch.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct synth_cfg {
    uint64_t seed;
    /* assembly */
    int32_t n_loci;        /* independent loci (contig "pairs"); each has n_copies contigs */
    int32_t n_copies;      /* 2 = diploid; up to 9 for the stress config */
    int64_t locus_len;     /* length of copy 0 of every locus */
    double snv_rate;       /* per base, copies 1.. vs copy 0 */
    double indel_rate;     /* 1-8 bp */
    double long_indel_rate;/* 11-60 bp */
    double hp_frac;        /* fraction of sequence inside homopolymer runs of 5-20 */
    double n_rate;         /* per base probability of an 'N' in the assembly */
    double rc_contig_prob; /* probability that a copy>0 contig is stored reverse-complemented */
    /* reads */
    double len_mean, len_sd;
    int32_t len_min, len_max;
    double err_sub, err_ins, err_del;
    double hp_indel_mult;  /* indel error multiplier inside homopolymers */
    double long_err_indel_prob; /* probability that an error indel is long (up to 25 bp) */
    int32_t qual_model;    /* 0 = HiFi mixture, 1 = ONT N(14,6) clipped to [2,40] */
    int32_t max_secondaries; /* alignments per group = 1 + min(max_secondaries, n_copies-1) */
    double wrong_primary_prob; /* primary placed on a copy other than the source */
    double clip_prob;      /* probability that an alignment gets end clipping */
    int32_t clip_max;      /* max clipped bases per end */
    double hard_clip_prob; /* given clipping on a secondary: use H instead of S */
    int32_t eqx;           /* 1: CIGAR uses =/X, 0: M */
    int32_t use_md;        /* 1: emit MD:Z instead of cs:Z */
} synth_cfg;

typedef struct synth synth;
typedef struct synth_batch synth_batch;

void synth_default_cfg(synth_cfg *c, int preset /*0 HiFi, 1 ONT, 2 stress*/);
synth *synth_create(const synth_cfg *c);
void synth_destroy(synth *s);

int32_t synth_n_contigs(const synth *s);
const char *synth_contig_name(const synth *s, int32_t tid);
const char *synth_contig_seq(const synth *s, int32_t tid); /* ASCII, NUL-terminated */
int64_t synth_contig_len(const synth *s, int32_t tid);

synth_batch *synth_generate(const synth *s, int64_t first_group, int32_t n_groups);
const sp_flat_batch *synth_batch_view(const synth_batch *b);
void synth_batch_free(synth_batch *b);

#ifdef __cplusplus
}
#endif
#endif
