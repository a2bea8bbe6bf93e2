// sp_group_warp.cuh -- stages K2 + K3 with one LANE per alignment: W lanes (2, 4 or 16) work on one read group,
// 32 / W read groups share a warp.
//
// Same outputs as the thread-per-group form (sp_group_markers, sp_consensus_loop -- which stay the specification,
// the form the host-side tests run, and what k_group launches when SECPHASE_B200_GROUP=serial), i.e. the results of
//   remove_all_mismatch_markers / sort_and_fill_markers / filter_ins_markers   (ptMarker.c:209-295, 156-206)
//   the x0.8 margin loop around correct_conf_blocks                            (secphase.c:161-169, ptMarker.c:495-647).
// What a thread did for the group's alignments one after the other is done by the alignments' lanes side by side:
//   - the merge of the marker lists: every lane holds the head of its own list, the smallest head is a segmented
//     minimum, the op search at that position (insertion / clip filter, ref_pos of matches) runs in every lane for
//     its own alignment, a ballot collects the verdict, every lane writes its own entry;
//   - in a round of the margin loop: sorting the alignment's block list, its flanking windows (find_flanking_blocks),
//     the projection of the consensus blocks into its coordinates and the check for over-long blocks -- two thirds
//     to nine tenths of the round; the chain of interval intersections in between is order-dependent and stays with
//     the group's first lane.
// The stages are chains of dependent loads (binary searches in op tables): their time is the length of the chain,
// which shrinks by the number of alignments, and a batch keeps n times more threads in flight.
#pragma once
#include "sp_blocks.cuh"
#include "sp_score.cuh"
#include "sp_walk_warp.cuh"  // SP_FULL, SP_LANE, SP_WD

#if defined(__CUDACC__) || defined(SP_WARP_EMU)

// minimum over the W lanes of a group (W a power of two; every lane of the warp takes part)
template <int W> SP_WD int sp_seg_min(int v) {
#pragma unroll
    for (int o = 1; o < W; o <<= 1) {
        const int t = __shfl_xor_sync(SP_FULL, v, o);
        v = t < v ? t : v;
    }
    return v;
}
template <int W> SP_WD int sp_seg_sum(int v) {
#pragma unroll
    for (int o = 1; o < W; o <<= 1) v += __shfl_xor_sync(SP_FULL, v, o);
    return v;
}
template <int W> SP_WD int sp_seg_or(int v) {
#pragma unroll
    for (int o = 1; o < W; o <<= 1) v |= __shfl_xor_sync(SP_FULL, v, o);
    return v;
}

// Work of one lane: alignment `sub` of read group V (valid == false: a lane without a group, or beyond the group's
// alignments -- it only takes part in the warp's collectives).  gpos / entries / W as for the serial functions;
// W.flank must hold V.n lists of W.cap intervals here.  The group's first lane (sub == 0) writes *o and *gP.
template <int WID>
SP_WDN void sp_group_lanes(const SpConst &C, const SpGroupAlnView &V, bool has_group, int32_t *gpos, SpEntry *entries,
                           int pos_cap, SpBlockWork &W, SpGroupOut *o, int32_t *gP) {
    const int lane = SP_LANE();
    const int sub = lane & (WID - 1);
    const int gbase = lane - sub;
    const uint32_t gmask = (WID == 32 ? 0xffffffffu : ((1u << WID) - 1u)) << gbase;
    const int n = has_group ? V.n : 0;
    const bool active = sub < n;
    const int a = V.a0 + sub;
    const bool rev = active && (V.flag[a] & SP_FREVERSE) != 0;
    const SpOp *ops = active ? V.ops + V.ops_off[a] : nullptr;
    const int n_ops = active ? V.info[a].n_ops : 0;
    int err = active ? V.info[a].err : 0;
    // ------------------------------------------------------------------ K2: the marker list (sp_group_markers)
    const int cnt = active ? V.info[a].n_imk : 0;
    const SpInitMarker *imk = active ? V.imk + V.imk_off[a] : nullptr;
    int cur = rev ? cnt - 1 : 0;  // every list is walked in ascending read_pos_f
    int head = (active && cur >= 0 && cur < cnt) ? imk[cur].read_pos_f : 0x7fffffff;
    const int n_init = sp_seg_sum<WID>(cnt);
    int P = 0, n_after_allmm = 0, n_filled = 0;
    for (;;) {
        const int p = sp_seg_min<WID>(head);  // smallest head position of the group
        const bool live = p != 0x7fffffff;
        if (!__any_sync(SP_FULL, live)) break;
        const bool mine = live && active && head == p;  // this alignment mismatches at p
        const int occ = __popc(__ballot_sync(SP_FULL, mine) & gmask);
        int has = -1;
        if (mine) {
            has = cur;
            cur += rev ? -1 : 1;
            head = (cur >= 0 && cur < cnt) ? imk[cur].read_pos_f : 0x7fffffff;
        }
        const bool cand = live && occ != n;  // occ == n: a mismatch in every alignment, a read error (ptMarker.c:223-225)
        if (cand) {
            n_after_allmm += occ;
            n_filled += n;
        }
        // insertion / clip filter (ptMarker.c:172-190) + ref_pos of match markers inside '=' ops
        bool bad = false;
        int refpos_eq = SP_INT_MIN;
        if (cand && active) {
            const int j = sp_find_op_by_read_pos(ops, n_ops, rev, p);
            if (j >= 0) {
                const SpOpView v = sp_op_view(ops, j, rev);
                if (v.op == SP_CINS || v.op == SP_CSOFT || v.op == SP_CHARD) bad = true;
                if (v.op == SP_CEQUAL) refpos_eq = rev ? v.rfs + v.rde_f - p : v.rfs + p - v.rds_f;
            }
        }
        const bool dropped = (__ballot_sync(SP_FULL, bad) & gmask) != 0;
        if (!cand || dropped) continue;
        if (P >= pos_cap) {
            err |= SP_GERR_MARKER_CAP;
            P++;
            continue;
        }
        if (sub == 0) gpos[P] = p;
        if (active) {
            SpEntry e;
            if (has >= 0) {
                const SpInitMarker m = imk[has];
                e.base_idx = m.base_idx;
                e.ref_pos = m.ref_pos;
                e.q = m.q;
                e.flags = 0;
            } else {  // ptMarker_construct_match, ptMarker.c:77-107
                const int lq = V.l_qseq[a];
                const int bi = rev ? lq + V.info[a].rclip_h - p - 1 : p - V.info[a].lclip_h;
                e.base_idx = bi;
                e.q = (bi >= 0 && bi < lq) ? (int) V.qual_pool[V.qual_off[a] + bi] : 0;  // (Q5, see sp_group_markers)
                e.ref_pos = -1;
                e.flags = 1;
            }
            if (refpos_eq != SP_INT_MIN) e.ref_pos = refpos_eq;  // ptMarker.c:184-187
            entries[(int64_t) P * n + sub] = e;
        }
        P++;
    }
    if (P > pos_cap) P = pos_cap;
    // ------------------------------------------------------------------ K3: the margin loop (sp_consensus_loop)
    const int cap = W.cap;
    SpBlock *my_blocks = W.ab + (int64_t) sub * cap;
    SpIv *my_flank = W.flank + (int64_t) sub * cap;
    int margin = C.flank_margin, conf_len = 1;
    bool stopped = !(has_group && P > 0 && C.consensus);
    if (has_group && P == 0 && active) W.nb[sub] = 0;  // secphase.c:161: no markers, no confident blocks
    __syncwarp();
    for (;;) {
        // needs_to_find_blocks (ptMarker.c:649-667): an alignment without blocks, or a block longer than the limit
        bool need = false;
        if (!stopped && active) {
            const int nb = W.nb[sub];
            need = nb == 0;
            for (int k = 0; k < nb; k++)
                if ((my_blocks[k].sqe - my_blocks[k].sqs) > SP_MAX_BLOCK_LEN || (my_blocks[k].rfe - my_blocks[k].rfs) > SP_MAX_BLOCK_LEN)
                    need = true;
        }
        const uint32_t need_m = __ballot_sync(SP_FULL, need);  // (every lane votes: no short-circuit around a collective)
        const bool go = !stopped && (need_m & gmask) != 0;
        if (!go) stopped = true;
        if (!__any_sync(SP_FULL, go)) break;
        if (go) margin = (int) (margin * 0.8);  // "flank_margin_eff *= 0.8" on an int
        // ---- one round of correct_conf_blocks (ptMarker.c:495-647)
        int nf = 0;
        if (go && active) {
            sp_sort_blocks_by_rds(my_blocks, W.nb[sub]);
            nf = sp_flank_list(V.info[a], n, P, gpos, margin, my_flank, cap, &err);
        }
        __syncwarp();
        // the first lane intersects: the alignments' blocks in turn, then their flanking windows in turn
        int nfs[WID];
#pragma unroll
        for (int i = 0; i < WID; i++) nfs[i] = __shfl_sync(SP_FULL, nf, gbase + i);
        int nc = 0, in_b = 0;
        if (go && sub == 0) {
            SpIv *curl = W.cons_a, *nxt = W.cons_b;
            nc = W.nb[0];
            for (int k = 0; k < nc; k++) {
                curl[k].s = W.ab[k].rds_f;
                curl[k].e = W.ab[k].rde_f;
            }
            for (int i = 1; i < n; i++) {
                const SpBlock *bi = W.ab + (int64_t) i * cap;
                nc = sp_intersect(curl, nc, [&](int j) { SpIv v; v.s = bi[j].rds_f; v.e = bi[j].rde_f; return v; }, W.nb[i], nxt,
                                  cap, &err);
                SpIv *t = curl; curl = nxt; nxt = t;
            }
#pragma unroll
            for (int i = 0; i < WID; i++) {
                if (i < n) {
                    const SpIv *fl = W.flank + (int64_t) i * cap;
                    nc = sp_intersect(curl, nc, [&](int j) { return fl[j]; }, nfs[i], nxt, cap, &err);
                    SpIv *t = curl; curl = nxt; nxt = t;
                }
            }
            in_b = curl == W.cons_b;
        }
        __syncwarp();
        nc = __shfl_sync(SP_FULL, nc, gbase);
        in_b = __shfl_sync(SP_FULL, in_b, gbase);
        if (go) {
            if (nc == 0) {  // 515-523
                if (active) W.nb[sub] = 0;
                stopped = true;
            } else if (active) {
                W.nb[sub] = sp_project_blocks(V, sub, in_b ? W.cons_b : W.cons_a, nc, C.indel_threshold, my_blocks, cap, &err);
            }
            conf_len = nc;
        }
        __syncwarp();
    }
    err = sp_seg_or<WID>(err);
    if (has_group && sub == 0) {
        SpGroupOut out;
        out.best_pre = 0; out.prim_idx = 0;
        out.n_init = n_init;
        out.n_after_allmm = n_after_allmm;
        out.n_filled = n_filled;
        out.n_after_ins = P * n;
        out.margin_eff = margin;
        out.conf_len = conf_len;
        out.n_final = 0;
        out.scored = (P > 0 && (conf_len > 0 || !C.consensus)) ? 1 : 0;
        out.max_idx = 0; out.tie_mask = 0;
        out.err = err;
        out.fin_off = 0; out.pad0 = 0; out.pad1 = 0;
        *o = out;
        *gP = P;
    }
}

// K5 with a lane per alignment: sp_score_group's pass over the group's marker list (ptMarker.c:110-153 filter, the
// score of ptAlignment.c / secphase.c:171-177).  Every lane resolves the BAQ of its own entry, the minimum over the
// alignments is a segmented minimum, every lane adds its own alignment's terms in list order (the order the serial
// form adds them in, so the sums are the same doubles) and writes its own row of the final-marker table.
// Returns the number of rows (group-uniform).
template <int WID>
SP_WDN int sp_score_lanes(const SpConst &C, const SpGroupAlnView &V, bool has_group, int P, const int32_t *gpos,
                          const SpEntry *entries, const int32_t *res, const SpRow *rows, bool scored, double *score,
                          int32_t *fin, int32_t *baq_out) {
    const int sub = SP_LANE() & (WID - 1);
    const int n = has_group ? V.n : 0;
    const bool active = sub < n;
    if (!has_group) P = 0;
    double sc = 0.0;
    int nf = 0;
    for (int p = 0; __any_sync(SP_FULL, p < P); p++) {
        const bool on = p < P && active;
        SpEntry e;
        e.base_idx = 0; e.ref_pos = 0; e.q = 0; e.flags = 0;
        int qv = 100;  // ptMarker.c:119
        if (on) {
            e = entries[(int64_t) p * n + sub];
            qv = (scored && C.baq_flag) ? sp_resolve_q(C, e, res[(int64_t) p * n + sub], rows) : e.q;
            if (baq_out) baq_out[(int64_t) p * n + sub] = qv;
        }
        int mq = sp_seg_min<WID>(qv);
        if (mq > 100) mq = 100;
        if (p >= P || (scored && !(mq > C.min_q))) continue;  // ptMarker.c:124,144 (strict)
        if (active) {
            const int q = scored ? mq : qv;
            int32_t *row = fin + (int64_t) (nf + sub) * 6;
            row[0] = sub;
            row[1] = gpos[p];
            row[2] = e.base_idx;
            row[3] = q;
            row[4] = e.flags & 1;
            row[5] = e.ref_pos;
            if (scored) {
                const int qi = q & 255;  // reverse_quality takes a uint8_t
                sc = SP_DADD(sc, (e.flags & 1) ? C.sc_match[qi] : C.sc_mis[qi]);
            }
        }
        nf += n;
    }
    if (active) score[sub] = sc;
    return nf;
}

#endif  // __CUDACC__ || SP_WARP_EMU
