// sp_score.cuh -- stage K5: BAQ write-back at the markers, low-quality filter, scores, selection.
//
// Restates, for one read group,
//   the bq[] rule and write-back of calc_local_baq   (ptMarker.c:763-786)  at marker bases only
//   calc_update_baq_all's marker update              (ptMarker.c:823-830)
//   filter_lowq_markers                              (ptMarker.c:110-153)
//   calc_alignment_score / reverse_quality           (ptMarker.c:307-325, 298-304)
//   the deterministic part of get_best_record_index  (ptAlignment.c:137-177)
// log()/pow() never run on the device: the two score terms come from 256-entry tables the host
// fills with its own libm (SpConst::sc_match / sc_mis), the device only performs the IEEE adds
// in list order, which is what makes alignment->score bit-identical.
#pragma once
#include "sp_blocks.cuh"
#include "sp_common.h"
#include "sp_hmm.cuh"

struct SpGroupOut {  // per group, device -> host
    int32_t best_pre;   // result of ptAlignment.c:175 given max_idx = first maximum (before tie-break RNG)
    int32_t prim_idx;
    int32_t n_init, n_after_allmm, n_filled, n_after_ins;
    int32_t margin_eff, conf_len, n_final, scored;
    int32_t max_idx;    // first secondary with the maximal score (ptAlignment.c:149-153)
    int32_t tie_mask;   // secondaries with score >= max (ptAlignment.c:157-163)
    int32_t err;
    int32_t fin_off;    // start of this group's rows in the compact final-marker table
    int32_t pad0, pad1;
};

// BAQ of one marker entry after calc_update_baq_all
SP_HD int sp_resolve_q(const SpConst &C, const SpEntry &e, int res, const SpRow *rows) {
    if (res == SP_RES_RAW) return e.q;
    if (res == SP_RES_ZERO) return 0;
    int bq;
    if (res == SP_RES_SETQ) {
        bq = C.set_q & 255;  // uint8_t block_qual[] = set_q (ptMarker.c:747-748)
    } else {
        const SpRow r = rows[res];
        if (((r.state & 3) != 0) || ((r.state >> 2) != r.expected)) bq = 0;
        else bq = e.q < r.q ? e.q : r.q;
    }
    return bq < 94 ? bq : 93;
}

// --writeBam (secphase.c:182-189 writes the records after calc_update_baq_all has modified their
// qualities in place).  One row of the write-back range of an HMM window, ptMarker.c:763-786:
// qual[sqs+t] = min(bq[t], 93) with bq[t] = set_q where no M/=/X op visited the base, 0 where the
// MAP state disagrees with the alignment, else min(raw quality, q[t]).
SP_HD uint8_t sp_baq_row_qual(const SpConst &C, const SpRow &r, int raw_q) {
    int bq;
    if (r.expected == SP_INT_MIN) bq = C.set_q & 255;
    else if (((r.state & 3) != 0) || ((r.state >> 2) != r.expected)) bq = 0;
    else bq = raw_q < r.q ? raw_q : r.q;
    return (uint8_t) (bq < 94 ? bq : 93);
}
// The markers calc_local_baq zeroes next to block edges (ptMarker.c:709-720, 797-806); runs after the
// rows above (the trailing margin is zeroed after the block's write-back, 786 vs 797).
SP_HD void sp_baq_zero_group(const SpGroupAlnView &G, int P, const SpEntry *entries, const int32_t *res,
                             uint8_t *qual_out) {
    const int n = G.n;
    for (int p = 0; p < P; p++)
        for (int i = 0; i < n; i++)
            if (res[(int64_t) p * n + i] == SP_RES_ZERO)
                qual_out[G.qual_off[G.a0 + i] + entries[(int64_t) p * n + i].base_idx] = 0;
}

// Writes the final list to fin[] as SP_MARKER_W-wide rows and returns the number of rows.
// baq_out[P*n] (optional) receives every entry's quality after calc_update_baq_all.
SP_HD int sp_score_group(const SpConst &C, const SpGroupAlnView &G, int P, const int32_t *gpos, const SpEntry *entries,
                         const int32_t *res, const SpRow *rows, bool scored, double *score, int32_t *fin,
                         int32_t *baq_out) {
    const int n = G.n;
    for (int i = 0; i < n; i++) score[i] = 0.0;
    int nf = 0;
    for (int p = 0; p < P; p++) {
        int qv[SP_MAX_ALN_PER_GROUP_C];
        int mq = 100;  // ptMarker.c:119
        for (int i = 0; i < n; i++) {
            const SpEntry e = entries[(int64_t) p * n + i];
            qv[i] = (scored && C.baq_flag) ? sp_resolve_q(C, e, res[(int64_t) p * n + i], rows) : e.q;
            if (baq_out) baq_out[(int64_t) p * n + i] = qv[i];
            if (mq > qv[i]) mq = qv[i];
        }
        if (scored && !(mq > C.min_q)) continue;  // ptMarker.c:124,144 (strict)
        for (int i = 0; i < n; i++) {
            const SpEntry e = entries[(int64_t) p * n + i];
            const int q = scored ? mq : qv[i];
            int32_t *row = fin + (int64_t) nf * 6;
            row[0] = i;
            row[1] = gpos[p];
            row[2] = e.base_idx;
            row[3] = q;
            row[4] = e.flags & 1;
            row[5] = e.ref_pos;
            nf++;
            if (scored) {
                const int qi = q & 255;  // reverse_quality takes a uint8_t
                score[i] = SP_DADD(score[i], (e.flags & 1) ? C.sc_match[qi] : C.sc_mis[qi]);
            }
        }
    }
    return nf;
}

// deterministic part of get_best_record_index, ptAlignment.c:137-177
SP_HD void sp_select(const SpGroupAlnView &G, const double *score, double prim_margin, double min_score,
                     SpGroupOut *out) {
    const int n = G.n;
    double max_score = -1.7976931348623157e308, prim_score = -1.7976931348623157e308;
    int max_idx = -1, prim_idx = -1;
    for (int i = 0; i < n; i++) {
        if ((G.flag[G.a0 + i] & SP_FSECONDARY) == 0) {
            prim_idx = i;
            prim_score = score[i];
        } else if (max_score < score[i]) {
            max_idx = i;
            max_score = score[i];
        }
    }
    int mask = 0;
    for (int i = 0; i < n; i++)
        if (((G.flag[G.a0 + i] & SP_FSECONDARY) != 0) && (max_score <= score[i])) mask |= 1 << i;
    out->prim_idx = prim_idx;
    out->max_idx = max_idx;
    out->tie_mask = mask;
    if (n == 1) {
        out->best_pre = 0;
        return;
    }
    out->best_pre = (prim_idx == -1 || max_score <= (prim_score + prim_margin) || max_score < min_score) ? prim_idx
                                                                                                        : max_idx;
}
