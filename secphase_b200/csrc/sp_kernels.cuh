// sp_kernels.cuh -- __global__ kernels of the marker-mode scoring pipeline (sm_100a).
//
//   k_walk            K1  thread per alignment     CIGAR/cs walk -> op table, extents, markers, confident blocks
//   k_group           K2+K3 thread per read group  marker merge/filter, consensus-block loop, HMM count pass
//   k_scan_groups         one CTA                  exclusive scans -> per-group item/row/s offsets + totals
//   k_emit            K3b thread per read group    HMM instances + marker rows (deterministic placement)
//   k_sort_*              counting sort of instances by (band class, length) for lock-step warps
//   k_hmm             K4  lane per HMM instance    the FP64 forward/backward kernel (sp_hmm.cuh)
//   k_score           K5  thread per read group    BAQ at markers, filter, scores, selection
// No tensor cores: none of this is a dense contraction.
#pragma once
#include <cuda_runtime.h>

#include "sp_blocks.cuh"
#include "sp_common.h"
#include "sp_hmm.cuh"
#include "sp_hmm2.cuh"
#include "sp_hmmf.cuh"
#include "sp_markers.cuh"
#include "sp_score.cuh"
#include "sp_walk.cuh"
#include "sp_walk_warp.cuh"
#include "sp_group_warp.cuh"

#define SP_SORT_LBINS 2048

struct SpTotals {  // device-side counters of one batch, read back once mid-pipeline
    int32_t n_items, n_rows;
    int64_t cells;
    int64_t s_doubles;
    int32_t max_bw, pad;
    int32_t class_count[SP_N_CLASSES];
    int32_t fin_rows;  // compact final-marker rows reserved so far
    int32_t err;
    int32_t max_lq, pad1;
    int64_t class_rows[SP_N_CLASSES];
};

struct SpBatchPtrs {
    int32_t G, A;
    // uploaded
    const int32_t *grp_aln_off, *flag, *tid, *pos, *l_qseq, *n_cigar, *tag_kind, *aln_grp, *gblk_cap, *glist;
    const int64_t *cigar_off, *tag_off, *seq_off, *qual_off, *ops_off, *imk_off, *gpos_off, *gent_off, *gblk_off,
        *giv_off;
    const uint32_t *cigar_pool;
    const uint8_t *tag_pool, *seq_pool, *qual_pool;
    // device work tables
    SpOp *ops;
    SpInitMarker *imk;
    SpAlnInfo *info;
    SpBlock *blk;
    SpIv *iv;
    int32_t *nb;
    int32_t *gpos;
    SpEntry *ent;
    int32_t *res;
    int32_t *baq;      // optional (debug)
    int32_t *gP;
    SpGroupOut *gout;
    SpEmitCounts *gcnt;
    SpEmitCounts *acnt;  // per alignment: what its HMM windows add to the group's counts (k_count)
    int32_t *walk_fb;    // [0] count, [1..] alignments the warp walker left to the serial one (k_walk_warp -> k_walk_list)
    int32_t *item_off, *row_off;
    int64_t *sdbl_off;
    double *score;
    int32_t *fin_wide;  // [total_ent][6] worst-case placement
    int32_t *fin;       // compact
    // reference
    const uint8_t *ref;
    const int64_t *contig_off;
    int32_t n_contigs;
};

__device__ __forceinline__ SpGroupAlnView sp_make_view(const SpBatchPtrs &B, int g) {
    SpGroupAlnView V;
    V.a0 = B.grp_aln_off[g];
    V.n = B.grp_aln_off[g + 1] - V.a0;
    V.flag = B.flag;
    V.l_qseq = B.l_qseq;
    V.qual_off = B.qual_off;
    V.qual_pool = B.qual_pool;
    V.info = B.info;
    V.ops_off = B.ops_off;
    V.ops = B.ops;
    V.imk_off = B.imk_off;
    V.imk = B.imk;
    return V;
}

__device__ __forceinline__ void sp_walk_one(const SpBatchPtrs &B, const SpConst *__restrict__ Cp, int a) {
    const int g = B.aln_grp[a];
    const int i = a - B.grp_aln_off[g];
    const int cap = B.gblk_cap[g];
    SpBlock *cb = B.blk + B.gblk_off[g] + (int64_t) i * cap;
    SpAlnInfo info;
    sp_walk_alignment(Cp->indel_threshold, Cp->min_q, B.flag[a], B.pos[a], B.l_qseq[a], B.n_cigar[a],
                      B.cigar_pool + B.cigar_off[a], B.tag_pool, B.tag_off[a], B.tag_off[a + 1],
                      B.tag_kind[a], B.qual_pool + B.qual_off[a], B.ops + B.ops_off[a],
                      (int) (B.ops_off[a + 1] - B.ops_off[a] - 1), B.imk + B.imk_off[a],
                      (int) (B.imk_off[a + 1] - B.imk_off[a]), cb, cap, &info);
    B.info[a] = info;
    B.nb[a] = info.n_cb;
}

// thread per alignment: the serial walker (SECPHASE_B200_WALK=serial, and the specification of the warp walker)
__global__ void __launch_bounds__(128) k_walk(SpBatchPtrs B, const SpConst *__restrict__ Cp) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= B.A) return;
    sp_walk_one(B, Cp, a);
}

// warp per alignment (sp_walk_warp.cuh); what it declines (MD tags, text that is not what an aligner writes for
// this CIGAR) is listed for k_walk_list
// min_ops: alignments whose op table is planned smaller than this go to the list at once -- with a few tokens per
// 32 bytes of text and a hundred-odd tokens in all (HiFi) the serial walker executes fewer warp instructions per
// alignment (32 alignments a warp) than a warp that scans every byte; with thousands of tokens (ONT) it is the
// other way round by a factor of three to four.
__global__ void __launch_bounds__(128, 8) k_walk_warp(SpBatchPtrs B, const SpConst *__restrict__ Cp, int min_ops) {
    __shared__ uint8_t s_code[256];  // class of every byte value (sp_cs_code)
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_code[i] = (uint8_t) sp_cs_code((uint32_t) i);
    __syncthreads();
    const int a = (int) ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (a >= B.A) return;
    if ((int) (B.ops_off[a + 1] - B.ops_off[a] - 1) < min_ops) {
        if ((threadIdx.x & 31) == 0) B.walk_fb[1 + atomicAdd(B.walk_fb, 1)] = a;
        return;
    }
    const int g = B.aln_grp[a];
    const int i = a - B.grp_aln_off[g];
    const int cap = B.gblk_cap[g];
    SpBlock *cb = B.blk + B.gblk_off[g] + (int64_t) i * cap;
    const bool ok = sp_walk_alignment_warp(Cp->indel_threshold, Cp->min_q, B.flag[a], B.pos[a], B.l_qseq[a], B.n_cigar[a],
                                           B.cigar_pool + B.cigar_off[a], B.tag_pool, B.tag_off[a], B.tag_off[a + 1],
                                           B.tag_kind[a], B.qual_pool + B.qual_off[a], B.ops + B.ops_off[a],
                                           (int) (B.ops_off[a + 1] - B.ops_off[a] - 1), B.imk + B.imk_off[a],
                                           (int) (B.imk_off[a + 1] - B.imk_off[a]), cb, cap, &B.info[a], s_code);
    if ((threadIdx.x & 31) == 0) {
        if (ok) B.nb[a] = B.info[a].n_cb;
        else B.walk_fb[1 + atomicAdd(B.walk_fb, 1)] = a;
    }
}

__global__ void __launch_bounds__(128) k_walk_list(SpBatchPtrs B, const SpConst *__restrict__ Cp) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B.walk_fb[0]) return;
    sp_walk_one(B, Cp, B.walk_fb[1 + t]);
}

#ifdef SP_PROFILE_GROUP  // tuning aid: clock64 split of the thread-per-group stages (summed over threads)
#define SP_PROF_T0() long long prof_t = clock64()
#define SP_PROF(k) do { long long t_ = clock64(); atomicAdd(&sp_prof[k], (unsigned long long) (t_ - prof_t)); prof_t = t_; } while (0)
#else
#define SP_PROF_T0() ((void) 0)
#define SP_PROF(k) ((void) 0)
#endif

// thread per read group: groups glist[first .. first + count) (the whole batch: first 0, count G)
__global__ void __launch_bounds__(64) k_group(SpBatchPtrs B, const SpConst *__restrict__ Cp, int first, int count) {
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= count) return;
    const int g = B.glist[first + slot];
    SP_PROF_T0();
    const SpConst &C = *Cp;
    SpGroupAlnView V = sp_make_view(B, g);
    SpGroupOut o;
    memset(&o, 0, sizeof(o));
    int err = 0;
    for (int i = 0; i < V.n; i++) err |= B.info[V.a0 + i].err;
    int32_t counts[4];
    int32_t *gpos = B.gpos + B.gpos_off[g];
    SpEntry *ent = B.ent + B.gent_off[g];
    const int P = sp_group_markers(V, gpos, ent, (int) (B.gpos_off[g + 1] - B.gpos_off[g]), counts, &err);
    SP_PROF(0);
    B.gP[g] = P;
    o.n_init = counts[0];
    o.n_after_allmm = counts[1];
    o.n_filled = counts[2];
    o.n_after_ins = counts[3];
    SpBlockWork W;
    W.cap = B.gblk_cap[g];
    W.ab = B.blk + B.gblk_off[g];
    W.nb = B.nb + V.a0;
    W.cons_a = B.iv + B.giv_off[g];
    W.cons_b = W.cons_a + W.cap;
    W.flank = W.cons_b + W.cap;
    int margin = C.flank_margin, conf_len = 1;
    bool scored = false;
    if (P > 0) {
        conf_len = sp_consensus_loop(C, V, P, gpos, W, &margin, &err);
        SP_PROF(1);
        if (conf_len > 0 || !C.consensus) scored = true;  // (the HMM windows are counted per alignment, k_count)
    } else {
        for (int i = 0; i < V.n; i++) W.nb[i] = 0;  // secphase.c:161: no markers, no confident blocks
    }
    SP_PROF(2);
    o.margin_eff = margin;
    o.conf_len = conf_len;
    o.scored = scored ? 1 : 0;
    o.err = err;
    B.gout[g] = o;
}

// K2 + K3 with a lane per alignment (sp_group_warp.cuh): WID lanes per read group, the groups of one lane class
// (glist[first .. first + count)).
template <int WID>
__global__ void __launch_bounds__(128) k_group_lanes(SpBatchPtrs B, const SpConst *__restrict__ Cp, int first, int count) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if ((t & ~31) / WID >= count) return;  // (whole warps only)
    const int slot = t / WID;
    const bool has = slot < count;
    const int g = has ? B.glist[first + slot] : 0;
    SpGroupAlnView V = sp_make_view(B, g);
    SpBlockWork W;
    W.cap = B.gblk_cap[g];
    W.ab = B.blk + B.gblk_off[g];
    W.nb = B.nb + V.a0;
    W.cons_a = B.iv + B.giv_off[g];
    W.cons_b = W.cons_a + W.cap;
    W.flank = W.cons_b + W.cap;
    sp_group_lanes<WID>(*Cp, V, has, B.gpos + B.gpos_off[g], B.ent + B.gent_off[g], (int) (B.gpos_off[g + 1] - B.gpos_off[g]), W,
                        &B.gout[g], &B.gP[g]);
}

// K3b count pass, thread per ALIGNMENT: the HMM windows calc_local_baq would run for this alignment (ptMarker.c:
// 670-809) -- how many instances / rows / band cells they are.  Alignments of a group are independent here, so the
// lay-out walk of a group's windows is spread over as many threads as it has alignments (up to ten in the
// many-secondaries case) instead of running them one after the other in the group's thread.
__global__ void __launch_bounds__(128) k_count(SpBatchPtrs B, const SpConst *__restrict__ Cp) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= B.A) return;
    const SpConst &C = *Cp;
    const int g = B.aln_grp[a];
    SpEmitCounts cnt;
    memset(&cnt, 0, sizeof(cnt));
    if (B.gout[g].scored && C.baq_flag) {
        SpGroupAlnView V = sp_make_view(B, g);
        const int i = a - V.a0, cap = B.gblk_cap[g];
        sp_emit_alignment<false>(C, V, i, B.gP[g], B.ent + B.gent_off[g], B.blk + B.gblk_off[g] + (int64_t) i * cap, B.nb[a],
                                 B.contig_off[B.tid[a]], 0, cnt, nullptr, nullptr, 0, nullptr, 0, 0);
    }
    B.acnt[a] = cnt;
}
// per group: the sum over its alignments
__global__ void __launch_bounds__(128) k_group_counts(SpBatchPtrs B) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= B.G) return;
    SpEmitCounts s;
    memset(&s, 0, sizeof(s));
    for (int a = B.grp_aln_off[g]; a < B.grp_aln_off[g + 1]; a++) {
        const SpEmitCounts c = B.acnt[a];
        s.n_items += c.n_items;
        s.n_rows += c.n_rows;
        s.cells += c.cells;
        s.s_doubles += c.s_doubles;
        s.max_bw = max(s.max_bw, c.max_bw);
        s.max_lq = max(s.max_lq, c.max_lq);
        for (int k = 0; k < SP_N_CLASSES; k++) { s.class_count[k] += c.class_count[k]; s.class_rows[k] += c.class_rows[k]; }
    }
    B.gcnt[g] = s;
}

// exclusive scans over groups (one CTA of 1024 threads); also accumulates the batch totals
__global__ void __launch_bounds__(1024) k_scan_groups(SpBatchPtrs B, SpTotals *tot) {
    __shared__ int32_t s_items[1024], s_rows[1024];
    __shared__ int64_t s_sd[1024], s_cells[1024];
    __shared__ int32_t s_cls[SP_N_CLASSES], s_maxbw, s_maxlq;
    __shared__ unsigned long long s_crows[SP_N_CLASSES];
    const int t = threadIdx.x, G = B.G;
    const int per = (G + 1023) / 1024;
    const int g0 = t * per, g1 = min(G, g0 + per);
    if (t < SP_N_CLASSES) { s_cls[t] = 0; s_crows[t] = 0; }
    if (t == 0) { s_maxbw = 0; s_maxlq = 0; }
    __syncthreads();
    int32_t li = 0, lr = 0, lbw = 0, llq = 0;
    int64_t ls = 0, lc = 0;
    int32_t lcls[SP_N_CLASSES];
    int64_t lcr[SP_N_CLASSES];
    for (int k = 0; k < SP_N_CLASSES; k++) { lcls[k] = 0; lcr[k] = 0; }
    for (int g = g0; g < g1; g++) {
        const SpEmitCounts c = B.gcnt[g];
        li += c.n_items;
        lr += c.n_rows;
        ls += c.s_doubles;
        lc += c.cells;
        lbw = max(lbw, c.max_bw);
        llq = max(llq, c.max_lq);
        for (int k = 0; k < SP_N_CLASSES; k++) { lcls[k] += c.class_count[k]; lcr[k] += c.class_rows[k]; }
    }
    s_items[t] = li;
    s_rows[t] = lr;
    s_sd[t] = ls;
    s_cells[t] = lc;
    atomicMax(&s_maxbw, lbw);
    atomicMax(&s_maxlq, llq);
    for (int k = 0; k < SP_N_CLASSES; k++)
        if (lcls[k]) {
            atomicAdd(&s_cls[k], lcls[k]);
            atomicAdd(&s_crows[k], (unsigned long long) lcr[k]);
        }
    __syncthreads();
    // Hillis-Steele inclusive scan
    for (int d = 1; d < 1024; d <<= 1) {
        int32_t a = 0, b = 0;
        int64_t c = 0, e = 0;
        if (t >= d) { a = s_items[t - d]; b = s_rows[t - d]; c = s_sd[t - d]; e = s_cells[t - d]; }
        __syncthreads();
        s_items[t] += a; s_rows[t] += b; s_sd[t] += c; s_cells[t] += e;
        __syncthreads();
    }
    int32_t bi = s_items[t] - li, br = s_rows[t] - lr;
    int64_t bs = s_sd[t] - ls;
    for (int g = g0; g < g1; g++) {
        const SpEmitCounts c = B.gcnt[g];
        B.item_off[g] = bi;
        B.row_off[g] = br;
        B.sdbl_off[g] = bs;
        bi += c.n_items;
        br += c.n_rows;
        bs += c.s_doubles;
    }
    if (t == 1023) {
        tot->n_items = s_items[1023];
        tot->n_rows = s_rows[1023];
        tot->s_doubles = s_sd[1023];
        tot->cells = s_cells[1023];
        tot->max_bw = s_maxbw;
        tot->max_lq = s_maxlq;
        for (int k = 0; k < SP_N_CLASSES; k++) { tot->class_count[k] = s_cls[k]; tot->class_rows[k] = (int64_t) s_crows[k]; }
        B.item_off[G] = s_items[1023];
        B.row_off[G] = s_rows[1023];
        B.sdbl_off[G] = s_sd[1023];
    }
}

// K3b emit pass, thread per alignment: places the alignment's instances and rows behind those of the alignments
// before it in the group (the order the reference runs them, calc_update_baq_all ptMarker.c:811-831)
__global__ void __launch_bounds__(128) k_emit(SpBatchPtrs B, const SpConst *__restrict__ Cp, SpItem *items, SpRow *rows) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= B.A) return;
    const SpConst &C = *Cp;
    const int g = B.aln_grp[a];
    if (!B.gout[g].scored || !C.baq_flag) return;
    SP_PROF_T0();
    SpGroupAlnView V = sp_make_view(B, g);
    const int i = a - V.a0, cap = B.gblk_cap[g];
    SpEmitCounts cnt;
    memset(&cnt, 0, sizeof(cnt));
    for (int k = V.a0; k < a; k++) {  // what the alignments before this one placed
        const SpEmitCounts c = B.acnt[k];
        cnt.n_items += c.n_items;
        cnt.n_rows += c.n_rows;
        cnt.s_doubles += c.s_doubles;
    }
    sp_emit_alignment<true>(C, V, i, B.gP[g], B.ent + B.gent_off[g], B.blk + B.gblk_off[g] + (int64_t) i * cap, B.nb[a],
                            B.contig_off[B.tid[a]], 0, cnt, B.res + B.gent_off[g], items, B.item_off[g], rows,
                            B.row_off[g], B.sdbl_off[g]);
    SP_PROF(3);
}

// -w mode: the rows of every HMM window (warp per instance, lane per row), see sp_fill_row
__global__ void __launch_bounds__(256) k_fill_rows(SpBatchPtrs B, const SpItem *__restrict__ items, int n_items, SpRow *rows) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n_items) return;
    const SpItem it = items[w];
    const int a = it.aln;
    const SpOp *ops = B.ops + B.ops_off[a];
    const bool rev = (B.flag[a] & SP_FREVERSE) != 0;
    const int blk_rfs = (int) (it.ref_off - B.contig_off[B.tid[a]]);
    for (int k = lane; k < it.n_rows; k += 32) rows[it.row0 + k] = sp_fill_row(it, w, k, ops, rev, blk_rfs);
}

// ---- instance ordering: by band half-width (all widths beyond the shared-memory kernels share one bin), then by
// descending window length: key = min(bw, MAXBW+1) * LBINS + (LBINS-1 - min(l_query, LBINS-1)).  Band classes are
// ranges of widths, so every class is a contiguous stretch of the order; inside it the 32-instance sets a warp
// pulls are uniform in width (what the fast kernel's virtual band wants) except where two widths meet.
#define SP_SORT_BWBINS (SP_H2_MAXBW + 3)
__device__ __forceinline__ int sp_item_key(const SpItem &it) {
    const int bw = sp_hmm_bw(it.l_ref, it.l_query, it.par_bw);
    const int lq = it.l_query < SP_SORT_LBINS - 1 ? it.l_query : SP_SORT_LBINS - 1;
    const int b = it.n_rows > 0 ? (bw <= SP_H2_MAXBW ? bw : SP_H2_MAXBW + 1) : SP_H2_MAXBW + 2;  // row-less instances are not run
    return b * SP_SORT_LBINS + (SP_SORT_LBINS - 1 - lq);
}
__global__ void k_sort_hist(const SpItem *items, int n, int32_t *bins) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(&bins[sp_item_key(items[i])], 1);
}
// exclusive scan over (SP_N_CLASSES+1)*LBINS bins, one CTA; class_start[c] = first slot of class c
__global__ void __launch_bounds__(1024) k_sort_scan(int32_t *bins, int nbins, int32_t *class_start) {
    __shared__ int32_t s[1024];
    const int t = threadIdx.x;
    const int per = (nbins + 1023) / 1024;
    const int b0 = t * per, b1 = min(nbins, b0 + per);
    int32_t l = 0;
    for (int b = b0; b < b1; b++) l += bins[b];
    s[t] = l;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        int32_t a = t >= d ? s[t - d] : 0;
        __syncthreads();
        s[t] += a;
        __syncthreads();
    }
    int32_t run = s[t] - l;
    for (int b = b0; b < b1; b++) {
        const int32_t c = bins[b];
        bins[b] = run;
        run += c;
    }
    (void) class_start;
}
__global__ void k_sort_scatter(const SpItem *items, int n, int32_t *bins, int32_t *order) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) order[atomicAdd(&bins[sp_item_key(items[i])], 1)] = i;
}

// ---- K4 ---------------------------------------------------------------------------------
// One warp per CTA, one HMM instance per lane.  Dynamic shared memory: W*3*32 doubles of band
// state followed by W*32 uint32 reference codes (lane-interleaved).  For the widest class the
// band lives in global memory instead (gband != nullptr).
__global__ void __launch_bounds__(32) k_hmm(const SpConst *__restrict__ Cp, const SpItem *__restrict__ items,
                                            const int32_t *__restrict__ order, int first, int count, int W,
                                            const uint8_t *__restrict__ ref, const uint8_t *__restrict__ qbytes,
                                            const uint8_t *__restrict__ seq_pool, const int64_t *__restrict__ seq_off,
                                            double *__restrict__ s_pool, double *__restrict__ fsave, int64_t fs_stride,
                                            SpRow *rows, double *gband) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x;
    const int slot = blockIdx.x * 32 + lane;
    if (slot >= count) return;
    const SpItem it = items[order ? order[first + slot] : first + slot];
    SpBand<32> B;
    if (gband) {
        double *base = gband + (int64_t) blockIdx.x * ((int64_t) W * 3 * 32 + (int64_t) W * 16);
        B.row = base + lane;
        B.code = reinterpret_cast<uint32_t *>(base + (int64_t) W * 3 * 32) + lane;
    } else {
        B.row = smem + lane;
        B.code = reinterpret_cast<uint32_t *>(smem + W * 3 * 32) + lane;
    }
    B.W = W;
    SpHmmIn in;
    in.ref = ref + it.ref_off;
    if (it.query_off >= 0) {
        in.qbytes = qbytes;
        in.qseq4 = nullptr;
        in.q0 = it.query_off;
    } else {
        in.qbytes = nullptr;
        in.qseq4 = seq_pool + seq_off[it.aln];
        in.q0 = it.q_sqs;
    }
    in.l_ref = it.l_ref;
    in.l_query = it.l_query;
    in.par_bw = it.par_bw;
    sp_hmm_instance<32, 1>(*Cp, in, B, s_pool + it.s_off, fsave + (int64_t) it.row0 * fs_stride, fs_stride,
                           rows + it.row0, it.n_rows);
}


// Main K4 kernel (band half-width <= SP_H2_MAXBW): persistent CTAs (one per SM, as many warps as
// band slabs fit the SM's shared memory), one HMM instance per lane.  Warps pull sets of 32
// instances from a global counter in (band class, length)-sorted order, i.e. longest first, so the
// SMs drain together.  Per-warp slab in dynamic shared memory: ncell x 32 double2 (M,I) followed
// by ncell x 32 double (D), lane-interleaved; see sp_hmm2.cuh.
template <int NW, int NC, bool IL = false>
__global__ void __launch_bounds__(224, 1) k_hmm2(const SpConst *__restrict__ Cp, const SpItem *__restrict__ items,
                                              const int32_t *__restrict__ order, int first, int count, int ncell,
                                              const uint8_t *__restrict__ ref, const uint8_t *__restrict__ qbytes,
                                              const uint8_t *__restrict__ seq_pool, const int64_t *__restrict__ seq_off,
                                              double *__restrict__ s_pool, double *__restrict__ fsave, int64_t fs_stride,
                                              SpRow *rows, int *work_counter, const int64_t *__restrict__ set_base,
                                              int set_first, const int *__restrict__ count_ptr) {
    extern __shared__ double2 smem2[];
    if (count_ptr) count = *count_ptr;  // strict re-run of the instances the fast kernel's guard band flagged
    const int lane = threadIdx.x & 31;
    double2 *slab = smem2 + (size_t) (threadIdx.x >> 5) * ncell * 48;  // ncell*32*(16+8) bytes per warp
    double2 *mi = slab + lane;
    double *d = reinterpret_cast<double *>(slab + ncell * 32) + lane;
    const int nwork = (count + 31) >> 5;
    for (;;) {
        int w = 0;
        if (lane == 0) w = atomicAdd(work_counter, 1);
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w >= nwork) break;
        const int slot = w * 32 + lane;
        if (slot < count) {
            const SpItem it = items[order ? order[first + slot] : first + slot];
            for (int c = 0; c < ncell; c++) {
                mi[c * 32] = make_double2(0., 0.);
                d[c * 32] = 0.;
            }
            SpBand2<32> B;
            B.mi = mi + 32;  // cell -1 is the permanent zero cell
            B.d = d + 32;
            SpHmmIn in;
            in.ref = ref + it.ref_off;
            if (it.query_off >= 0) {
                in.qbytes = qbytes;
                in.qseq4 = nullptr;
                in.q0 = it.query_off;
            } else {
                in.qbytes = nullptr;
                in.qseq4 = seq_pool + seq_off[it.aln];
                in.q0 = it.q_sqs;
            }
            in.l_ref = it.l_ref;
            in.l_query = it.l_query;
            in.par_bw = it.par_bw;
            if constexpr (IL)  // -w mode: forward rows of the warp's 32 instances interleaved lane by lane (fs_stride = cells*64)
                sp_hmm2_instance<32, NW, NC, 64>(*Cp, in, B, s_pool + it.s_off, fsave + set_base[set_first + w] + 2 * lane,
                                                 fs_stride, rows + it.row0, it.n_rows, w * 32 + 32 <= count);
            else
                sp_hmm2_instance<32, NW, NC>(*Cp, in, B, s_pool + it.s_off, fsave + (int64_t) it.row0 * fs_stride,
                                             fs_stride, rows + it.row0, it.n_rows, w * 32 + 32 <= count);
        }
        __syncwarp();
    }
}

// Fast-arithmetic K4 kernel (sp_hmmf.cuh): persistent CTAs, one instance per lane, per-warp slab of (NC+2) x 32
// 16-byte cells and nothing else.  Every lane of a warp runs an instance (the last, partial set repeats its last
// instance with no rows) because the row bodies are chosen by warp votes.  An instance whose guard band fired is
// recomputed at once, by its own lane, with the strict body (sp_hmm2.cuh, the reference's rounding order) in a
// private corner of the warp's slab -- a strict instance needs ~1 KB when it does not have to share banks with 31
// others -- so no second launch trails the class.
// warps per CTA = 16-byte-cell slabs that fit the 227 KB of shared memory of an SM, at most SP_HMMF_MAX_WARPS
// (two per scheduler; with eight warps a thread may use the whole 255-register budget)
#ifndef SP_HMMF_MAX_WARPS
#define SP_HMMF_MAX_WARPS 8
#endif
constexpr int sp_hmmf_warps(int nc) {
    return 232448 / ((nc + 2) * 512) > SP_HMMF_MAX_WARPS ? SP_HMMF_MAX_WARPS : 232448 / ((nc + 2) * 512);
}
template <int NC>
__global__ void __launch_bounds__(32 * sp_hmmf_warps(NC), 1) k_hmmf(const SpConst *__restrict__ Cp, const SpItem *__restrict__ items,
                                                 const int32_t *__restrict__ order, int first, int count,
                                                 const uint8_t *__restrict__ ref, const uint8_t *__restrict__ qbytes,
                                                 const uint8_t *__restrict__ seq_pool, const int64_t *__restrict__ seq_off,
                                                 double *__restrict__ s_pool, double *__restrict__ fsave, int64_t fs_stride,
                                                 SpRow *rows, int *work_counter, int *rerun_count, int guard_all) {
    extern __shared__ double2 smem2[];
    const int lane = threadIdx.x & 31;
    double2 *slab = smem2 + (size_t) (threadIdx.x >> 5) * (NC + 2) * 32;
    double2 *mi = slab + 32 + lane;  // cell -1 sits one row below
    const int nwork = (count + 31) >> 5;
    for (;;) {
        int w = 0;
        if (lane == 0) w = atomicAdd(work_counter, 1);
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w >= nwork) break;
        const int slot = w * 32 + lane;
        const bool dup = slot >= count;
        const int idx = order ? order[first + (dup ? count - 1 : slot)] : first + (dup ? count - 1 : slot);
        const SpItem it = items[idx];
        SpHmmIn in;
        in.ref = ref + it.ref_off;
        if (it.query_off >= 0) {
            in.qbytes = qbytes;
            in.qseq4 = nullptr;
            in.q0 = it.query_off;
        } else {
            in.qbytes = nullptr;
            in.qseq4 = seq_pool + seq_off[it.aln];
            in.q0 = it.q_sqs;
        }
        in.l_ref = it.l_ref;
        in.l_query = it.l_query;
        in.par_bw = it.par_bw;
        // The virtual band is as wide as the widest instance of the set.  Instances are ordered by width, so a set is
        // uniform except where two widths meet; there the narrower lanes run every row through the masked pass.
        const int bww = __reduce_max_sync(0xffffffffu, sp_hmm_bw(it.l_ref, it.l_query, it.par_bw));
        const int flag = sp_hmmf_instance<32, (NC + 63) / 64>(*Cp, in, mi, bww, fsave + (int64_t) it.row0 * fs_stride, fs_stride,
                                                             rows + it.row0, dup ? 0 : it.n_rows, guard_all != 0);
        bool flagged = !dup && flag != 0;
        __syncwarp();
        // guard band fired: the lane recomputes its instance in the reference's order (generic strict body, no votes)
        unsigned fl = __ballot_sync(0xffffffffu, flagged);
        if (fl) {
            constexpr int BWC = (NC - 1) / 2, NCELL = 2 * BWC + 3;     // strict band: cells -1 .. 2bw+1
            constexpr int PER = NCELL + (NCELL + 1) / 2;                 // (M,I) pairs + the D plane, in 16-byte units
            constexpr int CAP = (NC + 2) * 32 / PER;                     // instances the slab holds at a time
            if (lane == 0) atomicAdd(rerun_count, __popc(fl));
            int rank = __popc(fl & ((1u << lane) - 1));
            while (fl) {
                __syncwarp();
                if (flagged && rank < CAP) {
                    double2 *reg = slab + rank * PER;
                    for (int c = 0; c < PER; c++) reg[c] = make_double2(0., 0.);
                    SpBand2<1> B;
                    B.mi = reg + 1;
                    B.d = reinterpret_cast<double *>(reg + NCELL) + 1;
                    // (exact-width classes: the unrolled strict body, D state in registers -- 2-3x the generic one's speed)
                    constexpr int UNC = (NC == 41 || NC == 43 || NC == 45) ? NC : 0;
                    sp_hmm2_instance<1, (NC + 63) / 64, UNC, 2, (UNC > 45 ? SP_H2_NCRF_WIDE : SP_H2_NCRF), SP_H2_NCRB, true>(
                        *Cp, in, B, s_pool + it.s_off, fsave + (int64_t) it.row0 * fs_stride, fs_stride, rows + it.row0, it.n_rows, true);
                    flagged = false;
                }
                rank -= CAP;
                fl = __ballot_sync(0xffffffffu, flagged);
            }
        }
        __syncwarp();
    }
}

// -w mode: offset (in doubles) of every 32-instance set's lane-interleaved forward-row block.  Sets are
// numbered class by class in launch order; within a class instances are sorted by descending window
// length, so a set's longest instance is its first (valid while all windows are below the sort's last
// length bin, which the launcher checks).  One CTA; set_base[n_sets] receives the total.
struct SpSetPlan {
    int32_t first_item[SP_N_CLASSES], count[SP_N_CLASSES], first_set[SP_N_CLASSES + 1], cells[SP_N_CLASSES];
};
__global__ void __launch_bounds__(1024) k_fs_sets(SpSetPlan pl, const SpItem *__restrict__ items,
                                                  const int32_t *__restrict__ order, int64_t *set_base) {
    __shared__ int64_t s[1024];
    const int t = threadIdx.x;
    const int n_sets = pl.first_set[SP_N_CLASSES];
    const int per = (n_sets + 1023) / 1024;
    const int s0 = min(n_sets, t * per), s1 = min(n_sets, s0 + per);
    auto size_of = [&](int k) -> int64_t {  // rows of the set's longest window (a set may span two band widths)
        int c = 0;
        while (c + 1 < SP_N_CLASSES && k >= pl.first_set[c + 1]) c++;
        const int i0 = (k - pl.first_set[c]) * 32, i1 = min(i0 + 32, pl.count[c]);
        int nr = 0;
        for (int i = i0; i < i1; i++) nr = max(nr, items[order[pl.first_item[c] + i]].n_rows);
        return (int64_t) nr * pl.cells[c] * 64;
    };
    int64_t l = 0;
    for (int k = s0; k < s1; k++) l += size_of(k);
    s[t] = l;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        int64_t a = t >= d ? s[t - d] : 0;
        __syncthreads();
        s[t] += a;
        __syncthreads();
    }
    int64_t run = s[t] - l;
    for (int k = s0; k < s1; k++) {
        set_base[k] = run;
        run += size_of(k);
    }
    if (t == 1023) set_base[n_sets] = s[1023];
}

__global__ void __launch_bounds__(64) k_score(SpBatchPtrs B, const SpConst *__restrict__ Cp, const SpRow *rows,
                                              double prim_margin, double min_score, SpTotals *tot, int first, int count) {
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= count) return;
    const int g = B.glist[first + slot];
    const SpConst &C = *Cp;
    SpGroupAlnView V = sp_make_view(B, g);
    SpGroupOut o = B.gout[g];
    int32_t *wide = B.fin_wide + B.gent_off[g] * 6;
    const int nf = sp_score_group(C, V, B.gP[g], B.gpos + B.gpos_off[g], B.ent + B.gent_off[g],
                                  B.res + B.gent_off[g], rows, o.scored != 0, B.score + V.a0, wide,
                                  B.baq ? B.baq + B.gent_off[g] : nullptr);
    o.n_final = nf;
    sp_select(V, B.score + V.a0, prim_margin, min_score, &o);
    const int off = atomicAdd(&tot->fin_rows, nf);
    o.fin_off = off;
    int32_t *dst = B.fin + (int64_t) off * 6;
    for (int k = 0; k < nf * 6; k++) dst[k] = wide[k];
    if (o.err) atomicOr(&tot->err, o.err);
    B.gout[g] = o;
}

// ---- --writeBam: BAQ-modified quality arrays (full_baq mode) --------------------------------
// thread per HMM row: fully parallel, one byte read + one byte written per base of every window
// K5 with a lane per alignment (sp_score_lanes), the groups of one lane class
template <int WID>
__global__ void __launch_bounds__(128) k_score_lanes(SpBatchPtrs B, const SpConst *__restrict__ Cp, const SpRow *rows,
                                                     double prim_margin, double min_score, SpTotals *tot, int first, int count) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if ((t & ~31) / WID >= count) return;  // (whole warps only)
    const int slot = t / WID, sub = threadIdx.x & (WID - 1), gbase = (threadIdx.x & 31) - sub;
    const bool has = slot < count;
    const int g = has ? B.glist[first + slot] : 0;
    SpGroupAlnView V = sp_make_view(B, g);
    SpGroupOut o = B.gout[g];
    int32_t *wide = B.fin_wide + B.gent_off[g] * 6;
    const int nf = sp_score_lanes<WID>(*Cp, V, has, B.gP[g], B.gpos + B.gpos_off[g], B.ent + B.gent_off[g], B.res + B.gent_off[g],
                                       rows, o.scored != 0, B.score + V.a0, wide, B.baq ? B.baq + B.gent_off[g] : nullptr);
    __syncwarp();
    int off = 0;
    if (has && sub == 0) {
        o.n_final = nf;
        sp_select(V, B.score + V.a0, prim_margin, min_score, &o);
        off = atomicAdd(&tot->fin_rows, nf);
        o.fin_off = off;
        if (o.err) atomicOr(&tot->err, o.err);
        B.gout[g] = o;
    }
    off = __shfl_sync(SP_FULL, off, gbase);
    if (has) {
        int32_t *dst = B.fin + (int64_t) off * 6;
        for (int k = sub; k < nf * 6; k += WID) dst[k] = wide[k];
    }
}

__global__ void __launch_bounds__(256) k_baq_rows(const SpConst *__restrict__ Cp, const SpItem *__restrict__ items,
                                                  const SpRow *__restrict__ rows, int n_rows,
                                                  const int64_t *__restrict__ qual_off,
                                                  const uint8_t *__restrict__ qual_pool, uint8_t *__restrict__ qual_out) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const SpRow R = rows[r];
    const SpItem it = items[R.item];
    const int64_t p = qual_off[it.aln] + it.q_sqs + R.t;
    qual_out[p] = sp_baq_row_qual(*Cp, R, qual_pool[p]);
}
__global__ void __launch_bounds__(64) k_baq_zero(SpBatchPtrs B, uint8_t *qual_out) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= B.G) return;
    SpGroupAlnView V = sp_make_view(B, g);
    sp_baq_zero_group(V, B.gP[g], B.ent + B.gent_off[g], B.res + B.gent_off[g], qual_out);
}

// ---- reference encoding: ASCII -> codes 0..4 (seq_nt16_int[seq_nt16_table[c]], ptMarker.c:744)
__global__ void k_encode_ref(const uint8_t *__restrict__ ascii, uint8_t *__restrict__ codes, int64_t n) {
    int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const uint8_t c = ascii[i];
        uint8_t v = 4;
        switch (c) {
            case 'A': case 'a': case '0': v = 0; break;
            case 'C': case 'c': case '1': v = 1; break;
            case 'G': case 'g': case '2': v = 2; break;
            case 'T': case 't': case '3': v = 3; break;
            default: v = 4;
        }
        codes[i] = v;
    }
}

// ---- FP64 pipe micro-benchmark (roofline denominator) -----------------------------------
__global__ void k_fp64_peak(double *out, int iters, int mode) {
    double a0 = 1.0 + threadIdx.x * 1e-9, a1 = a0 + 1e-3, a2 = a0 + 2e-3, a3 = a0 + 3e-3;
    double a4 = a0 + 4e-3, a5 = a0 + 5e-3, a6 = a0 + 6e-3, a7 = a0 + 7e-3;
    const double m = 1.0000001, c = 1e-9;
    if (mode == 0) {
        for (int i = 0; i < iters; i++) {
            a0 = __fma_rn(a0, m, c); a1 = __fma_rn(a1, m, c); a2 = __fma_rn(a2, m, c); a3 = __fma_rn(a3, m, c);
            a4 = __fma_rn(a4, m, c); a5 = __fma_rn(a5, m, c); a6 = __fma_rn(a6, m, c); a7 = __fma_rn(a7, m, c);
        }
    } else {
        for (int i = 0; i < iters; i++) {
            a0 = __dmul_rn(a0, m); a1 = __dadd_rn(a1, c); a2 = __dmul_rn(a2, m); a3 = __dadd_rn(a3, c);
            a4 = __dmul_rn(a4, m); a5 = __dadd_rn(a5, c); a6 = __dmul_rn(a6, m); a7 = __dadd_rn(a7, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
