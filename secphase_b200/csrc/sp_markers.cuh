// sp_markers.cuh -- stage K2: the marker list of one read group.
//
// One pass replaces, for one read group of n alignments,
//   remove_all_mismatch_markers   (ptMarker.c:209-248, comparator ptMarker_cmp 31-39)
//   sort_and_fill_markers         (ptMarker.c:251-295, ptMarker_construct_match 77-107)
//   filter_ins_markers            (ptMarker.c:156-206)
// The reference sorts one list of (read_pos_f, alignment_idx) keys three times; each
// alignment's initial markers already come out of the walk monotone in read_pos_f (ascending
// on the forward strand, descending on the reverse strand), so an n-way merge (n <= 10) yields
// the same order without a sort.  The insertion filter and the ref_pos fill-in are a binary
// search in each alignment's op table instead of a lock-step re-walk.
#pragma once
#include "sp_common.h"
#include "sp_walk.cuh"

struct SpGroupAlnView {  // what K2..K5 need to know about the alignments of one group
    int n;               // alignments in the group
    int a0;              // global index of the first
    const int32_t *flag;
    const int32_t *l_qseq;
    const int64_t *qual_off;
    const uint8_t *qual_pool;
    const SpAlnInfo *info;     // indexed by global alignment index
    const int64_t *ops_off;    // per alignment, into ops (each table has cap+1 slots)
    const SpOp *ops;
    const int64_t *imk_off;
    const SpInitMarker *imk;
};

// Returns the number of surviving positions P; writes gpos[P] and entries[P*n] (position-major,
// alignment-minor == the order of the reference's sorted list).  counts[0..3] = list lengths
// after get_initial / remove_all_mismatch / fill / filter_ins (for the parity table).
SP_HD int sp_group_markers(const SpGroupAlnView &G, int32_t *gpos, SpEntry *entries, int pos_cap, int32_t *counts,
                           int *err) {
    const int n = G.n;
    // per alignment: cursor into its initial markers (walked in ascending read_pos_f) and the position under the
    // cursor (the heads are compared n times per position).  The op that holds a position is found by bisecting the
    // whole op table every time: galloping from the previous hit was measured slower (the first levels of the
    // bisection stay in L1; profiles/r02_prof_group_gallop_v29.txt).
    int cur[SP_MAX_ALN_PER_GROUP_C], head[SP_MAX_ALN_PER_GROUP_C];
    int n_init = 0;
    for (int i = 0; i < n; i++) {
        const int a = G.a0 + i;
        const bool rev = (G.flag[a] & SP_FREVERSE) != 0;
        const int cnt = G.info[a].n_imk;
        n_init += cnt;
        cur[i] = rev ? cnt - 1 : 0;  // walk every list in ascending read_pos_f
        head[i] = (cur[i] >= 0 && cur[i] < cnt) ? G.imk[G.imk_off[a] + cur[i]].read_pos_f : 0x7fffffff;
    }
    int P = 0, n_after_allmm = 0, n_filled = 0;
    for (;;) {
        // smallest head position
        int p = 0x7fffffff;
        for (int i = 0; i < n; i++) p = head[i] < p ? head[i] : p;
        if (p == 0x7fffffff) break;
        // which alignments mismatch at p
        int occ = 0;
        int has[SP_MAX_ALN_PER_GROUP_C];
        for (int i = 0; i < n; i++) {
            has[i] = -1;
            if (head[i] == p) {
                const int a = G.a0 + i;
                const int cnt = G.info[a].n_imk;
                has[i] = cur[i];
                occ++;
                cur[i] += ((G.flag[a] & SP_FREVERSE) != 0) ? -1 : 1;
                head[i] = (cur[i] >= 0 && cur[i] < cnt) ? G.imk[G.imk_off[a] + cur[i]].read_pos_f : 0x7fffffff;
            }
        }
        if (occ == n) continue;  // mismatch in every alignment: a read error (ptMarker.c:223-225)
        n_after_allmm += occ;
        n_filled += n;
        // insertion / clip filter (ptMarker.c:172-190) + ref_pos of match markers inside '=' ops
        bool keep = true;
        int refpos_eq[SP_MAX_ALN_PER_GROUP_C];
        for (int i = 0; i < n; i++) {
            const int a = G.a0 + i;
            const bool rev = (G.flag[a] & SP_FREVERSE) != 0;
            const SpOp *ops = G.ops + G.ops_off[a];
            refpos_eq[i] = SP_INT_MIN;
            int j = sp_find_op_by_read_pos(ops, G.info[a].n_ops, rev, p);
            if (j < 0) continue;
            SpOpView v = sp_op_view(ops, j, rev);
            if (v.op == SP_CINS || v.op == SP_CSOFT || v.op == SP_CHARD) keep = false;
            if (v.op == SP_CEQUAL) refpos_eq[i] = rev ? v.rfs + v.rde_f - p : v.rfs + p - v.rds_f;
        }
        if (!keep) continue;
        if (P >= pos_cap) {
            *err |= SP_GERR_MARKER_CAP;
            P++;
            continue;
        }
        gpos[P] = p;
        for (int i = 0; i < n; i++) {
            const int a = G.a0 + i;
            SpEntry e;
            if (has[i] >= 0) {
                const SpInitMarker m = G.imk[G.imk_off[a] + has[i]];
                e.base_idx = m.base_idx;
                e.ref_pos = m.ref_pos;
                e.q = m.q;
                e.flags = 0;
            } else {  // ptMarker_construct_match, ptMarker.c:77-107
                const bool rev = (G.flag[a] & SP_FREVERSE) != 0;
                const int lq = G.l_qseq[a];
                int bi = rev ? lq + G.info[a].rclip_h - p - 1 : p - G.info[a].lclip_h;
                e.base_idx = bi;
                // Q5: the reference reads qual[] out of bounds for positions inside this
                // alignment's hard clip; such positions never survive the filter above.
                e.q = (bi >= 0 && bi < lq) ? (int) G.qual_pool[G.qual_off[a] + bi] : 0;
                e.ref_pos = -1;
                e.flags = 1;
            }
            if (refpos_eq[i] != SP_INT_MIN) e.ref_pos = refpos_eq[i];  // ptMarker.c:184-187
            entries[(int64_t) P * n + i] = e;
        }
        P++;
    }
    counts[0] = n_init;
    counts[1] = n_after_allmm;
    counts[2] = n_filled;
    counts[3] = (P <= pos_cap ? P : pos_cap) * n;
    return P <= pos_cap ? P : pos_cap;
}
