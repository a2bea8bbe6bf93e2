// sp_blocks.cuh -- stage K3: consensus blocks of one read group and the HMM work list.
//
// Restates, for one read group,
//   needs_to_find_blocks      (ptMarker.c:649-667)
//   find_flanking_blocks      (ptMarker.c:446-484)   via set_flanking_blocks 487-492
//   intersect_by_rd_f         (ptMarker.c:398-435)
//   correct_conf_blocks       (ptMarker.c:495-647)   intersection + projection to seq/ref coords
//   the x0.8 margin loop      (secphase.c:162-169)
// and the control flow of
//   calc_local_baq            (ptMarker.c:670-809)   which blocks run probaln_glocal, which
//                                                    markers are zeroed / keep raw quality
// The projection walks the stored op table (SpOp) instead of re-tokenising the cs tag.
#pragma once
#include "sp_common.h"
#include "sp_markers.cuh"
#include "sp_walk.cuh"

#if defined(SP_PROFILE_GROUP) && defined(__CUDACC__)  // tuning aid (see sp_kernels.cuh): [0..7] stages, [8..15] inside the consensus rounds
__device__ unsigned long long sp_prof[16];
#endif
#if defined(SP_PROFILE_GROUP) && defined(__CUDA_ARCH__)
#define SP_PROFB_T0() long long profb_t = clock64()
#define SP_PROFB(k) do { long long t_ = clock64(); atomicAdd(&sp_prof[8 + (k)], (unsigned long long) (t_ - profb_t)); profb_t = t_; } while (0)
#else
#define SP_PROFB_T0() ((void) 0)
#define SP_PROFB(k) ((void) 0)
#endif

struct SpIv {  // interval in read-forward coordinates
    int32_t s, e;
};

SP_HD void sp_sort_blocks_by_rds(SpBlock *b, int n) {
    // The lists arrive sorted by stored-SEQ start, i.e. ascending in read-forward coordinates for a forward
    // alignment (nothing to do) and strictly descending for a reverse one (reverse in place: same result as
    // the insertion sort below, which would need n^2/2 moves for it -- a quarter of the block traffic of the
    // margin loop on read groups with many secondaries).
    bool asc = true, desc = true;
    for (int i = 1; i < n; i++) {
        const bool gt = b[i - 1].rds_f > b[i].rds_f;
        asc = asc && !gt;
        desc = desc && gt;
    }
    if (asc) return;
    if (desc) {
        for (int i = 0, j = n - 1; i < j; i++, j--) {
            const SpBlock t = b[i];
            b[i] = b[j];
            b[j] = t;
        }
        return;
    }
    for (int i = 1; i < n; i++) {
        SpBlock x = b[i];
        int j = i - 1;
        while (j >= 0 && b[j].rds_f > x.rds_f) {
            b[j + 1] = b[j];
            j--;
        }
        b[j + 1] = x;
    }
}
SP_HD void sp_sort_blocks_by_sqs(SpBlock *b, int n) {
    for (int i = 1; i < n; i++) {
        SpBlock x = b[i];
        int j = i - 1;
        while (j >= 0 && b[j].sqs > x.sqs) {
            b[j + 1] = b[j];
            j--;
        }
        b[j + 1] = x;
    }
}

// intersect_by_rd_f, ptMarker.c:398-435 (strict comparisons kept).  get2(j) reads list 2.
template <class Get2>
SP_HD int sp_intersect(const SpIv *l1, int n1, Get2 get2, int n2, SpIv *out, int cap, int *err) {
    if (n1 == 0 || n2 == 0) return 0;
    int j = 0, m = 0;
    bool have2 = true;
    SpIv b2 = get2(0);
    for (int i = 0; i < n1; i++) {
        const SpIv b1 = l1[i];
        while (have2 && b2.e < b1.s) {
            j++;
            have2 = j < n2;
            if (have2) b2 = get2(j);
        }
        while (have2 && b2.s < b1.e) {
            if (m < cap) {
                out[m].s = sp_max(b1.s, b2.s);
                out[m].e = sp_min(b1.e, b2.e);
            } else {
                *err |= SP_GERR_BLOCK_CAP;
            }
            m++;
            if (b2.e <= b1.e) {
                j++;
                have2 = j < n2;
                if (have2) b2 = get2(j);
            } else {
                break;
            }
        }
    }
    return m <= cap ? m : cap;
}

struct SpBlockWork {
    SpIv *cons_a, *cons_b, *flank;  // [cap] each
    SpBlock *ab;                    // [n][cap] per-alignment block lists
    int32_t *nb;                    // [n]
    int cap;
};

// find_flanking_blocks (ptMarker.c:446-484) for one alignment: the windows of +-margin around the group's marker
// positions, clipped to the alignment and merged.  Returns the list length (<= cap).
SP_HD int sp_flank_list(const SpAlnInfo &ai, int n, int P, const int32_t *gpos, int margin, SpIv *flank, int cap, int *err) {
    int nf = 0;
    if (P > 0) {
        int start = sp_max(ai.rds_f, gpos[0] - margin);
        int end = sp_min(ai.rde_f, gpos[0] + margin);
        // The marker list holds n entries per position (ptMarker.c:455-476 iterates all of them).
        // After the first entry of a position `end` equals that position's own clipped end, so the
        // other n-1 entries are no-ops whenever the clipped interval is non-degenerate (cs < ce);
        // only degenerate ones (cs >= ce) are replayed entry by entry.
        for (int pi = 0; pi < P; pi++) {
            const int p = gpos[pi];
            const int cs = sp_max(ai.rds_f, p - margin);
            const int ce = sp_min(ai.rde_f, p + margin);
            const int reps = (cs < ce) ? 1 : n;
            for (int r = (pi == 0 ? 1 : 0); r < (pi == 0 ? sp_max(reps, 1) : reps); r++) {
                if (cs < end) {
                    end = ce;
                } else {
                    if (nf < cap) { flank[nf].s = start; flank[nf].e = end; } else *err |= SP_GERR_BLOCK_CAP;
                    nf++;
                    start = cs;
                    end = ce;
                }
            }
        }
        if (nf < cap) { flank[nf].s = start; flank[nf].e = end; } else *err |= SP_GERR_BLOCK_CAP;
        nf++;
        if (nf > cap) nf = cap;
    }
    return nf;
}

// The projection of correct_conf_blocks (ptMarker.c:528-643) for alignment i of the group: the consensus blocks
// cur[0..nc) (read-forward coordinates) in the alignment's seq/ref coordinates, written to `out` in SEQ order.
// Returns the list length (<= cap).
SP_HD int sp_project_blocks(const SpGroupAlnView &G, int i, const SpIv *cur, int nc, int indel_threshold, SpBlock *out,
                            int cap, int *err) {
    const int a = G.a0 + i;
    const bool rev = (G.flag[a] & SP_FREVERSE) != 0;
    const SpOp *ops = G.ops + G.ops_off[a];
    const int n_ops = G.info[a].n_ops;
    int m = 0;
    int j = rev ? nc - 1 : 0;
    bool have = true, del_flag = false;
    int b_s, b_e;
    int rfs = 0, rfe = 0, sqs = 0, sqe = 0;  // the reference leaves these uninitialised
    if (rev) { b_s = -cur[j].e; b_e = -cur[j].s; } else { b_s = cur[j].s; b_e = cur[j].e; }
    for (int o = 0; o < n_ops; o++) {
        // Ops that cannot touch the current block are jumped over with a binary search in the op
        // table instead of being visited (the reference walks every op of the record in every
        // round of the x0.8 loop, ptMarker.c:545-641).  In walking coordinates an op spans
        // [w[o], w[o+1]-1]; with del_flag clear, an op matters only if it ends at or after b_s-1
        // (it may contain b_s, or be a deletion sitting at b_s) and, once ops start past b_s,
        // only if it reaches b_e (the emitting branch).  Everything skipped would at most have
        // cleared del_flag, which is already clear.
        if (!del_flag) {
            const int w0 = rev ? -ops[o].rdx : ops[o].rdx;
            const int target = (w0 > b_s) ? b_e + 1 : b_s;
            // (bisecting the whole remaining table beats galloping from o here: the first levels of the search hit
            // the same few entries every time and stay in L1; measured, profiles/r02_prof_group_gallop_v29.txt)
            int lo = o, hi = n_ops;  // first idx in [o, n_ops) with w[idx+1] >= target
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                const int w1 = rev ? -ops[mid + 1].rdx : ops[mid + 1].rdx;
                if (w1 >= target) hi = mid; else lo = mid + 1;
            }
            o = lo;
            if (o >= n_ops) break;
        }
        const SpOpView v = sp_op_view(ops, o, rev);
        int c_s, c_e;
        if (rev) { c_s = -v.rde_f; c_e = -v.rds_f; } else { c_s = v.rds_f; c_e = v.rde_f; }
        if (sp_op_is_match(v.op) || v.op == SP_CINS) {
            const bool ins = v.op == SP_CINS;
            while (have && b_e <= c_e) {
                if (c_s <= b_s && !(del_flag && c_s == b_s)) {
                    rfs = ins ? v.rfs : v.rfs + (b_s - c_s);
                    sqs = v.sqs + (b_s - c_s);
                }
                rfe = ins ? v.rfe : v.rfs + (b_e - c_s);
                sqe = v.sqs + (b_e - c_s);
                if (m < cap) {
                    SpBlock nb;
                    nb.rfs = rfs; nb.rfe = rfe; nb.sqs = sqs; nb.sqe = sqe;
                    nb.rds_f = cur[j].s; nb.rde_f = cur[j].e;
                    out[m] = nb;
                } else {
                    *err |= SP_GERR_BLOCK_CAP;
                }
                m++;
                if (rev && j > 0) {
                    j--;
                    b_s = -cur[j].e; b_e = -cur[j].s;
                } else if (!rev && j < nc - 1) {
                    j++;
                    b_s = cur[j].s; b_e = cur[j].e;
                } else {
                    have = false;
                }
            }
            if (!have) break;
            if (c_s <= b_s && b_s <= c_e && !(del_flag && c_s == b_s)) {
                rfs = ins ? v.rfs : v.rfs + (b_s - c_s);
                sqs = v.sqs + (b_s - c_s);
            }
            del_flag = false;
        } else if (v.op == SP_CDEL) {
            // 615-625 only writes prev_block->rfe of the temporary consensus list (no effect, Q6)
            if (have && b_s == c_s && v.len <= indel_threshold) {
                del_flag = true;
                rfs = v.rfs;
                sqs = v.sqs;
            }
        }
    }
    if (m > cap) m = cap;
    sp_sort_blocks_by_sqs(out, m);
    return m;
}

// correct_conf_blocks (ptMarker.c:495-647) preceded by set_flanking_blocks (487-492).
SP_HD int sp_correct_conf_blocks(const SpGroupAlnView &G, int P, const int32_t *gpos, int margin,
                                 int indel_threshold, SpBlockWork &W, int *err) {
    const int n = G.n, cap = W.cap;
    SP_PROFB_T0();
    // --- intersect the alignments' confident blocks (495-507)
    sp_sort_blocks_by_rds(W.ab, W.nb[0]);
    SpIv *cur = W.cons_a, *nxt = W.cons_b;
    int nc = W.nb[0];
    for (int k = 0; k < nc; k++) {
        cur[k].s = W.ab[k].rds_f;
        cur[k].e = W.ab[k].rde_f;
    }
    for (int i = 1; i < n; i++) {
        SpBlock *bi = W.ab + (int64_t) i * cap;
        sp_sort_blocks_by_rds(bi, W.nb[i]);
        nc = sp_intersect(cur, nc, [&](int j) { SpIv v; v.s = bi[j].rds_f; v.e = bi[j].rde_f; return v; },
                          W.nb[i], nxt, cap, err);
        SpIv *t = cur; cur = nxt; nxt = t;
    }
    SP_PROFB(0);
    // --- intersect with every alignment's flanking blocks (508-514)
    for (int i = 0; i < n; i++) {
        const int nf = sp_flank_list(G.info[G.a0 + i], n, P, gpos, margin, W.flank, cap, err);
        SP_PROFB(1);
        const SpIv *fl = W.flank;
        nc = sp_intersect(cur, nc, [&](int j) { return fl[j]; }, nf, nxt, cap, err);
        SpIv *t = cur; cur = nxt; nxt = t;
        SP_PROFB(2);
    }
    if (nc == 0) {  // 515-523
        for (int i = 0; i < n; i++) W.nb[i] = 0;
        return 0;
    }
    // --- project the consensus blocks into each alignment's seq/ref coordinates (528-643)
    for (int i = 0; i < n; i++) {
        W.nb[i] = sp_project_blocks(G, i, cur, nc, indel_threshold, W.ab + (int64_t) i * cap, cap, err);
        SP_PROFB(3);
    }
    return nc;
}

// needs_to_find_blocks, ptMarker.c:649-667
SP_HD bool sp_needs_to_find_blocks(const SpBlockWork &W, int n, int threshold) {
    bool flag = false;
    for (int j = 0; j < n; j++) {
        if (W.nb[j] == 0) return true;
        const SpBlock *b = W.ab + (int64_t) j * W.cap;
        for (int i = 0; i < W.nb[j]; i++)
            if ((b[i].sqe - b[i].sqs) > threshold || (b[i].rfe - b[i].rfs) > threshold) flag = true;
    }
    return flag;
}

// The loop of secphase.c:161-169.  On entry W.ab/nb hold the confident blocks of the walk.
// Returns conf_blocks_length (defined as 1 when the loop never runs, quirk Q1); *margin_out
// receives the effective flank margin.
SP_HD int sp_consensus_loop(const SpConst &C, const SpGroupAlnView &G, int P, const int32_t *gpos, SpBlockWork &W,
                            int *margin_out, int *err) {
    int margin = C.flank_margin;
    int conf_len = 1;
    while (C.consensus && sp_needs_to_find_blocks(W, G.n, SP_MAX_BLOCK_LEN)) {
        margin = (int) (margin * 0.8);  // "flank_margin_eff *= 0.8" on an int
        conf_len = sp_correct_conf_blocks(G, P, gpos, margin, C.indel_threshold, W, err);
        if (conf_len == 0) break;
    }
    *margin_out = margin;
    return conf_len;
}

// ---------------------------------------------------------------------------------------
// calc_local_baq control flow for ONE alignment (ptMarker.c:670-809).  RES_RAW: the marker
// keeps its raw quality; RES_ZERO: qual[base_idx] = 0 (709-720, 797-806); RES_SETQ: inside an HMM
// window but not under an M/=/X op, bq stays set_q (763-764); >=0: index of the SpRow whose HMM
// state/q decides (772-786).
enum { SP_RES_RAW = -1, SP_RES_ZERO = -2, SP_RES_SETQ = -3 };

SP_HD int sp_hmm_bw(int l_ref, int l_query, int par_bw) {  // band half-width probaln_glocal derives
    int bw = l_ref > l_query ? l_ref : l_query;
    if (bw > par_bw) bw = par_bw;
    int d = l_ref - l_query;
    if (d < 0) d = -d;
    if (bw < d) bw = d;
    return bw;
}
// band cells of one instance: sum over rows i = 1..l_query of the columns max(1,i-bw)..min(l_ref,i+bw).
// Closed form; every row is non-empty because bw >= |l_ref - l_query| (sp_hmm_bw).
SP_HD int64_t sp_hmm_cells(int l_ref, int l_query, int bw) {
    const int64_t Lr = l_ref, Lq = l_query, w = bw;
    int64_t k1 = Lr - w;  // rows i <= k1 end at i+bw, the others at l_ref
    k1 = k1 < 0 ? 0 : (k1 > Lq ? Lq : k1);
    const int64_t s_end = k1 * (k1 + 1) / 2 + k1 * w + (Lq - k1) * Lr;
    const int64_t k2 = Lq < w + 1 ? Lq : w + 1;  // rows i <= k2 begin at column 1, the others at i-bw
    const int64_t s_beg = k2 + (Lq * (Lq + 1) / 2 - k2 * (k2 + 1) / 2) - (Lq - k2) * w;
    return s_end - s_beg + Lq;
}

struct SpEmitCounts {
    int32_t n_items, n_rows;
    int64_t cells;
    int64_t s_doubles;  // sum over items of (l_query + 2)
    int32_t max_bw;
    int32_t class_count[SP_N_CLASSES];  // instances with at least one marker row, per band class
    int32_t class_rows[SP_N_CLASSES];   // their rows (sizes the lane-interleaved forward-row pool of the -w mode)
    int32_t max_lq;                     // longest window of the group
};

template <bool EMIT>
SP_HD void sp_emit_alignment(const SpConst &C, const SpGroupAlnView &G, int i, int P, const SpEntry *entries,
                             const SpBlock *blocks, int n_blocks, int64_t ref_base, int64_t entry_base,
                             SpEmitCounts &cnt, int32_t *res, SpItem *items, int item_base, SpRow *rows,
                             int row_base, int64_t s_base) {
    const int n = G.n, a = G.a0 + i;
    const bool rev = (G.flag[a] & SP_FREVERSE) != 0;
    const SpOp *ops = G.ops + G.ops_off[a];
    const int n_ops = G.info[a].n_ops;
    // marker cursor over this alignment's entries in list order (reverse strand: backwards)
    int j = rev ? P - 1 : 0;
    const int jstep = rev ? -1 : 1;
#define SP_MK_VALID(jj) ((jj) >= 0 && (jj) < P)
#define SP_MK_BASE(jj) (entries[(int64_t) (jj) * n + i].base_idx)
    // iterator state "constructed, not advanced" (cigar_it.c:24-31)
    int o = -1;
    int it_sqs = 0, it_sqe = -1, it_rfs = 0, it_rfe = -1, it_op = -1, it_len = 0;
    bool it_init = false;
    auto it_next = [&]() -> int {
        if (o + 1 >= n_ops) return 0;
        o++;
        SpOpView v = sp_op_view(ops, o, rev);
        it_sqs = v.sqs; it_sqe = v.sqe; it_rfs = v.rfs; it_rfe = v.rfe; it_op = v.op; it_len = v.len;
        return v.len;
    };
    if (!it_init) {
        // before the first next(): sqs=0, sqe=-1, rfs=pos, rfe=pos-1; pos == rfs of the first op
        it_rfs = n_ops > 0 ? ops[0].rfs : 0;
        it_rfe = it_rfs - 1;
        it_init = true;
    }
    for (int b = 0; b < n_blocks; b++) {
        const SpBlock blk = blocks[b];
        // advance to the first op that reaches the block in both coordinates (ptMarker.c:703-705).  The ops'
        // ends only grow along the table, so after a few steps the rest of the way is bisected: ONT
        // alignments have ~130 refined ops between consecutive windows, HiFi ~5.
        int skipped = 0;
        while (it_sqe < blk.sqs || it_rfe < blk.rfs) {
            if (++skipped > 8 && o + 2 < n_ops) {
                int lo = o + 1, hi = n_ops - 1;  // first op in [lo, hi] with sqe >= blk.sqs and rfe >= blk.rfs, hi if none
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    const SpOp nx = ops[mid + 1];
                    if (nx.sqs - 1 >= blk.sqs && nx.rfs - 1 >= blk.rfs) hi = mid; else lo = mid + 1;
                }
                o = lo - 1;
            }
            if (it_next() == 0) break;
        }
        while (SP_MK_VALID(j) && SP_MK_BASE(j) < blk.sqs + SP_BLOCK_MARGIN) {
            if (blk.sqs <= SP_MK_BASE(j)) {
                if (EMIT) res[entry_base + (int64_t) j * n + i] = SP_RES_ZERO;
            }
            j += jstep;
        }
        if (SP_MK_VALID(j) && SP_MK_BASE(j) <= blk.sqe - SP_BLOCK_MARGIN && blk.sqs + SP_BLOCK_MARGIN <= SP_MK_BASE(j)) {
            const int l_query = blk.sqe - blk.sqs + 1;
            const int l_ref = blk.rfe - blk.rfs + 1;
            int dlen = l_ref - l_query;
            if (dlen < 0) dlen = -dlen;
            const int par_bw = (int) (dlen + C.conf_b);  // ptMarker.c:754, int + double -> int
            const int bw = sp_hmm_bw(l_ref, l_query, par_bw);
            const int item_idx = item_base + cnt.n_items;
            const int first_row = row_base + cnt.n_rows;
            int n_rows = 0;
            // rows: this alignment's markers with base_idx in [sqs+10, sqe-11]; sqe-10 is written by the
            // HMM and then zeroed again (800-803), so it needs no row.
            // full_baq (--writeBam): one row per base of the write-back range t in [10, l_query-10)
            // (ptMarker.c:786), slot = first_row + t - 10; rows no M/=/X op visits keep
            // expected == SP_INT_MIN, i.e. bq = set_q (763-764).
            // The rows themselves are then written by sp_fill_row (one thread per row) from the range of
            // refined ops this loop examines, recorded in the SpItem.
            const bool full = C.full_baq != 0;
            if (full) {
                n_rows = l_query - 2 * SP_BLOCK_MARGIN;
                if (n_rows < 0) n_rows = 0;
            }
            int jj = j;
            const int op_first = o < 0 ? 0 : o;
            int op_last = op_first - 1;
            // the bq loop, ptMarker.c:767-785
            while (it_sqs <= blk.sqe || it_rfs <= blk.rfe) {
                int x = it_rfs - blk.rfs;
                x = x < 0 ? 0 : x;
                int y = it_sqs - blk.sqs;
                y = y < 0 ? 0 : y;
                if (o > op_last) op_last = o;
                if (sp_op_is_match(it_op)) {
                    const int len = sp_min(it_len, sp_min(it_sqe, blk.sqe) - sp_max(it_sqs, blk.sqs) + 1);
                    // markers with t in [y, y+len)
                    while (SP_MK_VALID(jj) && SP_MK_BASE(jj) - blk.sqs < y) jj += jstep;
                    while (SP_MK_VALID(jj) && SP_MK_BASE(jj) - blk.sqs < y + len) {
                        const int t = SP_MK_BASE(jj) - blk.sqs;
                        if (t >= SP_BLOCK_MARGIN && t < l_query - SP_BLOCK_MARGIN - 1) {
                            if (full) {
                                if (EMIT) res[entry_base + (int64_t) jj * n + i] = first_row + t - SP_BLOCK_MARGIN;
                            } else {
                                if (EMIT) {
                                    SpRow r;
                                    r.item = item_idx;
                                    r.t = t;
                                    r.entry = (int32_t) (entry_base + (int64_t) jj * n + i);
                                    r.expected = x + (t - y);
                                    r.state = 0;
                                    r.q = 0;
                                    r.pmax = 0.0;
                                    rows[first_row + n_rows] = r;
                                    res[entry_base + (int64_t) jj * n + i] = first_row + n_rows;
                                }
                                n_rows++;
                            }
                        }
                        jj += jstep;
                    }
                }
                if (it_sqe <= blk.sqe || it_rfe <= blk.rfe) {
                    if (it_next() == 0) break;
                } else {
                    break;
                }
            }
            if (EMIT) {
                // markers of the write-back range that no M/=/X op visited keep bq = set_q (763-764):
                // no HMM row is needed for them
                int j2 = j;
                while (SP_MK_VALID(j2) && SP_MK_BASE(j2) <= blk.sqe - SP_BLOCK_MARGIN - 1) {
                    const int64_t e = entry_base + (int64_t) j2 * n + i;
                    if (SP_MK_BASE(j2) - blk.sqs >= SP_BLOCK_MARGIN && res[e] == SP_RES_RAW) res[e] = SP_RES_SETQ;
                    j2 += jstep;
                }
                SpItem it;
                it.ref_off = ref_base + blk.rfs;
                it.aln = a;
                it.blk = b;
                it.l_ref = l_ref;
                it.l_query = l_query;
                it.q_sqs = blk.sqs;
                it.par_bw = par_bw;
                it.row0 = first_row;
                it.n_rows = n_rows;
                it.query_off = -1;
                it.s_off = s_base + cnt.s_doubles;
                it.op_first = op_first;
                it.op_last = op_last;
                items[item_idx] = it;
            }
            cnt.n_items += 1;
            cnt.n_rows += n_rows;
            cnt.cells += sp_hmm_cells(l_ref, l_query, bw);
            cnt.s_doubles += l_query + 2;
            if (bw > cnt.max_bw) cnt.max_bw = bw;
            if (n_rows > 0) {
                cnt.class_count[sp_band_class(bw)] += 1;
                cnt.class_rows[sp_band_class(bw)] += n_rows;
            }
            if (l_query > cnt.max_lq) cnt.max_lq = l_query;
        }
        while (SP_MK_VALID(j) && SP_MK_BASE(j) <= blk.sqe) {
            if (blk.sqe - SP_BLOCK_MARGIN <= SP_MK_BASE(j)) {
                if (EMIT) res[entry_base + (int64_t) j * n + i] = SP_RES_ZERO;
            }
            j += jstep;
        }
    }
#undef SP_MK_VALID
#undef SP_MK_BASE
}

// --writeBam mode: row k (query row t = k + 10) of an HMM window, as the bq loop of calc_local_baq sees
// the base (ptMarker.c:767-785): `expected` = x + (t - y) if one of the refined ops that loop examined
// (SpItem::op_first..op_last) is an M/=/X op covering the base, else SP_INT_MIN (bq stays set_q, 763-764).
// One thread per row; the op is found by binary search on the ops' stored-SEQ starts.
SP_HD SpRow sp_fill_row(const SpItem &it, int item_idx, int k, const SpOp *ops, bool rev, int blk_rfs) {
    SpRow r;
    r.item = item_idx;
    r.t = k + SP_BLOCK_MARGIN;
    r.entry = -1;
    r.expected = SP_INT_MIN;
    r.state = 0;
    r.q = 0;
    r.pmax = 0.0;
    const int blk_sqs = it.q_sqs, blk_sqe = it.q_sqs + it.l_query - 1;
    const int p = blk_sqs + r.t;
    int lo = it.op_first, hi = it.op_last;  // last op in [lo, hi] with sqs <= p (zero-length ops sort before the op that owns the base)
    if (lo > hi || ops[lo].sqs > p) return r;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (ops[mid].sqs <= p) lo = mid; else hi = mid - 1;
    }
    const SpOpView v = sp_op_view(ops, lo, rev);
    if (!sp_op_is_match(v.op) || p > v.sqe) return r;
    int x = v.rfs - blk_rfs;
    x = x < 0 ? 0 : x;
    int y = v.sqs - blk_sqs;
    y = y < 0 ? 0 : y;
    const int len = sp_min(v.len, sp_min(v.sqe, blk_sqe) - sp_max(v.sqs, blk_sqs) + 1);
    if (r.t >= y && r.t < y + len) r.expected = x + (r.t - y);
    return r;
}
