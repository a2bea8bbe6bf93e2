// sp_plan.h -- host-side batch planning shared by the launcher (sp_api.cu) and tests/hostsim.
//
// Everything the kernels write has a size that depends on the data (refined ops, markers,
// blocks).  The host knows cheap upper bounds from the record text alone and lays the
// per-alignment / per-group tables out with prefix sums, so kernels never allocate:
//   refined ops      <= n_cigar + (token start characters of the cs/MD text) + 1
//   initial markers  <= number of '*' in cs  (mismatched bases)   / upper-case letters in MD
//   confident blocks <= 1 + (#I/D longer than the indel threshold) + (#clip ops)
//   positions P      <= sum of the group's initial markers
//   consensus blocks <= P + sum(confident blocks) + 2n + 8   ("tight"; intersecting k interval
//                       lists can in theory exceed it -> kernels flag SP_GERR_BLOCK_CAP and the
//                       launcher re-plans that batch with the provable bound  n*P + sum(cb))
// Also holds the host-evaluated constants (float sub-expressions, libm tables) and the glibc
// rand() emulation used for the tie-breaks of get_best_record_index (ptAlignment.c:156-171).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "../../include/secphase_b200.h"
#include "sp_common.h"

struct SpPlan {
    int32_t G = 0, A = 0;
    std::vector<int32_t> aln_grp;   // [A] group of each alignment
    std::vector<int64_t> ops_off;   // [A+1]; each table has cap+1 slots (sentinel)
    std::vector<int64_t> imk_off;   // [A+1]
    std::vector<int32_t> cb_cap;    // [A]
    std::vector<int64_t> gpos_off;  // [G+1] positions
    std::vector<int64_t> gent_off;  // [G+1] entries (n per position)
    std::vector<int64_t> gblk_off;  // [G+1] SpBlock workspace: n lists of gblk_cap[g]
    std::vector<int64_t> giv_off;   // [G+1] SpIv workspace: 2 + n lists of gblk_cap[g] (two consensus lists, a flank list per alignment)
    std::vector<int32_t> glist;     // [G] the groups ordered by lane class (sp_group_lane_class): k_group_lanes<2|4|16>
    int32_t glist_n[3] = {0, 0, 0}; // groups per lane class
    std::vector<int32_t> gblk_cap;  // [G]
    std::vector<int64_t> g_msum, g_cbsum;  // [G] the bounds gblk_cap was derived from (sp_plan_block_caps)
    int64_t total_ops = 0, total_imk = 0, total_pos = 0, total_ent = 0, total_blk = 0, total_iv = 0;
};

// Per-alignment counts from the record text alone (the only O(bytes) part of planning; independent
// per alignment, so large batches are counted by a few threads).
struct SpAlnCounts {
    int64_t ntok, nmis;
    int32_t cb, rc;
};
inline void sp_count_alignment(const sp_flat_batch *b, int a, int indel_threshold, SpAlnCounts &o) {
    // byte classes: bit0 = starts a cs token (: * + -), bit1 = cs mismatch (*), bit2 = MD token
    // (upper-case letter, ^, digit), bit3 = MD mismatch (upper-case letter)
    static const struct Tab {
        uint8_t t[256];
        Tab() {
            memset(t, 0, sizeof(t));
            t[(uint8_t) ':'] |= 1; t[(uint8_t) '*'] |= 1 | 2; t[(uint8_t) '+'] |= 1; t[(uint8_t) '-'] |= 1;
            for (int c = 'A'; c <= 'Z'; c++) t[c] |= 4 | 8;
            for (int c = '0'; c <= '9'; c++) t[c] |= 4;
            t[(uint8_t) '^'] |= 4;
        }
    } tab;
    o.ntok = o.nmis = 0;
    o.cb = 1;
    o.rc = SP_OK;
    const int nc = b->n_cigar[a];
    if (nc < 1) { o.rc = SP_EINVAL; return; }
    const uint32_t *cig = b->cigar_pool + b->cigar_off[a];
    for (int k = 0; k < nc; k++) {
        const int op = (int) (cig[k] & 15), len = (int) (cig[k] >> 4);
        if (op == SP_CSOFT || op == SP_CHARD) o.cb++;
        else if ((op == SP_CINS || op == SP_CDEL) && len > indel_threshold) o.cb++;
        else if (op == SP_CREF_SKIP || op == SP_CPAD || op > SP_CDIFF) { o.rc = SP_EUNSUPPORTED; return; }
    }
    const uint8_t *t = (const uint8_t *) b->tag_pool + b->tag_off[a];
    const int64_t tl = b->tag_off[a + 1] - b->tag_off[a];
    const int kind = b->tag_kind ? b->tag_kind[a] : 0;
    int64_t ntok = 0, nmis = 0;
    if (kind == 0) {
        for (int64_t k = 0; k < tl; k++) {
            const uint8_t f = tab.t[t[k]];
            ntok += f & 1;
            nmis += (f >> 1) & 1;
        }
    } else {
        for (int64_t k = 0; k < tl; k++) {
            const uint8_t f = tab.t[t[k]];
            ntok += (f >> 2) & 1;
            nmis += (f >> 3) & 1;
        }
    }
    o.ntok = ntok;
    o.nmis = nmis;
}

// Block / interval workspaces of every group from the per-group bounds: the tight cap, or (safe) the
// provable one.  Needs nothing but the counts, so a batch whose kernels flagged SP_GERR_BLOCK_CAP can
// be laid out again without another pass over its text.
inline int sp_plan_block_caps(SpPlan &pl, const int32_t *grp_aln_off, bool safe_caps, int test_cap = 0) {
    int64_t blk = 0, iv = 0;
    for (int g = 0; g < pl.G; g++) {
        const int64_t n = grp_aln_off[g + 1] - grp_aln_off[g], P = pl.g_msum[(size_t) g], cbsum = pl.g_cbsum[(size_t) g];
        int64_t cap = safe_caps ? n * P + cbsum + 2 * n + 8 : P + cbsum + 2 * n + 8;
        if (test_cap > 0 && !safe_caps && cap > test_cap) cap = test_cap;  // tests: force the retry path
        if (cap > 0x3fffffff) return SP_ENOMEM;
        pl.gblk_cap[(size_t) g] = (int32_t) cap;
        pl.gblk_off[(size_t) g] = blk;
        blk += cap * n;
        pl.giv_off[(size_t) g] = iv;
        iv += cap * (2 + n);
    }
    pl.gblk_off[(size_t) pl.G] = blk;
    pl.giv_off[(size_t) pl.G] = iv;
    pl.total_blk = blk;
    pl.total_iv = iv;
    return SP_OK;
}

// lanes a read group of n alignments gets in k_group_lanes: 2, 4 or 16 (classes 0, 1, 2)
inline int sp_group_lane_class(int n) { return n <= 2 ? 0 : n <= 4 ? 1 : 2; }

// returns SP_OK or a negative error.  n_threads > 1: the per-alignment text scan runs on that many threads.
inline int sp_make_plan(const sp_flat_batch *b, int indel_threshold, bool safe_caps, SpPlan &pl, int n_threads = 1,
                        int test_cap = 0) {
    const int G = b->n_groups, A = b->n_alns;
    if (G < 0 || A < 0) return SP_EINVAL;
    pl.G = G;
    pl.A = A;
    pl.aln_grp.assign((size_t) A, 0);
    pl.ops_off.assign((size_t) A + 1, 0);
    pl.imk_off.assign((size_t) A + 1, 0);
    pl.cb_cap.assign((size_t) A, 0);
    pl.gpos_off.assign((size_t) G + 1, 0);
    pl.gent_off.assign((size_t) G + 1, 0);
    pl.gblk_off.assign((size_t) G + 1, 0);
    pl.giv_off.assign((size_t) G + 1, 0);
    pl.gblk_cap.assign((size_t) G, 0);
    pl.g_msum.assign((size_t) G, 0);
    pl.g_cbsum.assign((size_t) G, 0);
    for (int g = 0; g < G; g++) {
        const int a0 = b->grp_aln_off[g], a1 = b->grp_aln_off[g + 1];
        const int n = a1 - a0;
        if (n < 1 || n > SP_MAX_ALN_PER_GROUP || a0 < 0 || a1 > A) return SP_EINVAL;
    }
    {
        pl.glist.assign((size_t) G, 0);
        int32_t start[3] = {0, 0, 0};
        pl.glist_n[0] = pl.glist_n[1] = pl.glist_n[2] = 0;
        for (int g = 0; g < G; g++) pl.glist_n[sp_group_lane_class(b->grp_aln_off[g + 1] - b->grp_aln_off[g])]++;
        start[1] = pl.glist_n[0];
        start[2] = pl.glist_n[0] + pl.glist_n[1];
        for (int g = 0; g < G; g++) pl.glist[(size_t) start[sp_group_lane_class(b->grp_aln_off[g + 1] - b->grp_aln_off[g])]++] = g;
    }
    std::vector<SpAlnCounts> cnt((size_t) A);
    if (n_threads > 1 && A >= 4 * n_threads) {
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; t++)
            th.emplace_back([&, t] {
                // contiguous ranges balanced by tag bytes
                const int64_t total = b->tag_off[A], lo = total * t / n_threads, hi = total * (t + 1) / n_threads;
                int a = (int) (std::lower_bound(b->tag_off, b->tag_off + A, lo) - b->tag_off);
                const int a_end = t + 1 == n_threads ? A : (int) (std::lower_bound(b->tag_off, b->tag_off + A, hi) - b->tag_off);
                for (; a < a_end; a++) sp_count_alignment(b, a, indel_threshold, cnt[(size_t) a]);
            });
        for (auto &x : th) x.join();
    } else {
        for (int a = 0; a < A; a++) sp_count_alignment(b, a, indel_threshold, cnt[(size_t) a]);
    }
    int64_t ops = 0, imk = 0, pos = 0, ent = 0;
    for (int g = 0; g < G; g++) {
        const int a0 = b->grp_aln_off[g], a1 = b->grp_aln_off[g + 1];
        const int n = a1 - a0;
        int64_t msum = 0, cbsum = 0;
        for (int a = a0; a < a1; a++) {
            const SpAlnCounts &c = cnt[(size_t) a];
            if (c.rc != SP_OK) return c.rc;
            pl.aln_grp[(size_t) a] = g;
            pl.ops_off[(size_t) a] = ops;
            ops += b->n_cigar[a] + c.ntok + 2;  // +1 spare, +1 sentinel
            pl.imk_off[(size_t) a] = imk;
            imk += c.nmis;
            pl.cb_cap[(size_t) a] = c.cb;
            msum += c.nmis;
            cbsum += c.cb;
        }
        const int64_t P = msum;
        pl.g_msum[(size_t) g] = msum;
        pl.g_cbsum[(size_t) g] = cbsum;
        pl.gpos_off[(size_t) g] = pos;
        pos += P;
        pl.gent_off[(size_t) g] = ent;
        ent += P * n;
    }
    pl.ops_off[(size_t) A] = ops;
    pl.imk_off[(size_t) A] = imk;
    pl.gpos_off[(size_t) G] = pos;
    pl.gent_off[(size_t) G] = ent;
    pl.total_ops = ops;
    pl.total_imk = imk;
    pl.total_pos = pos;
    pl.total_ent = ent;
    return sp_plan_block_caps(pl, b->grp_aln_off, safe_caps, test_cap);
}

// ---- constants the device must not derive itself (host libm / C float semantics) ------------
inline void sp_fill_const(const sp_params &p, SpConst &C) {
    memset(&C, 0, sizeof(C));
    C.baq_flag = p.baq_flag;
    C.consensus = p.consensus;
    C.indel_threshold = p.indel_threshold;
    C.min_q = p.min_q;
    C.set_q = p.set_q;
    C.flank_margin = p.flank_margin;
    C.conf_b = p.conf_b;
    // probaln_par_t conf = {conf_d, conf_e, conf_bw} narrows to float (ptMarker.c:680)
    const float d = (float) p.conf_d, e = (float) p.conf_e;
    volatile float one_m_2d = 1 - d - d;  // htslib evaluates (1 - c->d - c->d) in float
    volatile float one_m_d = 1 - d;
    volatile float one_m_e = 1 - e;
    C.m0f = (double) one_m_2d;
    C.omd_f = (double) one_m_d;
    C.omd_ff = one_m_d;
    C.d_f = d;
    C.d_d = (double) d;
    C.ome_f = (double) one_m_e;
    C.e_d = (double) e;
    // qual[i] = g_qual2prob[iqual[i]] with g_qual2prob[q] = (float) pow(10, -q/10.)
    const float qf = (float) pow(10, -(p.set_q & 255) / 10.);
    C.em_match = 1. - qf;
    C.em_mis = qf * .33333333333;
    // q thresholds: f(t) = (int)(-4.343*log(t) + .499) is non-increasing in t; qthr[n] is the
    // largest double t with f(t) >= n, found by bisection over the bit patterns of (0, 1].
    for (int n = 1; n <= 101; n++) {
        uint64_t lo = 1, hi = 0x3ff0000000000000ull;  // smallest subnormal .. 1.0
        // invariant: f(lo) >= n (true for the smallest positive double: -4.343*log(4.9e-324) ~ 3235)
        while (lo < hi) {
            uint64_t mid = lo + (hi - lo + 1) / 2;
            double t;
            memcpy(&t, &mid, 8);
            int k = (int) (-4.343 * log(t) + .499);
            if (k >= n) lo = mid; else hi = mid - 1;
        }
        memcpy(&C.qthr[n], &lo, 8);
    }
    C.qthr[0] = 1.0;
    for (int q = 0; q < 256; q++) {
        // reverse_quality(uint8_t q), ptMarker.c:298-304
        double rq;
        if (q >= 93) rq = 0;
        else if (q == 0) rq = 93;
        else {
            double pr = 1 - pow(10, (double) q / -10);
            rq = -10 * log(pr);
        }
        C.sc_match[q] = -1 * rq;
        C.sc_mis[q] = -1 * q - 10 * log(3);
    }
}

// ---- glibc rand() (TYPE_3 additive feedback, r[i] = r[i-3] + r[i-31]), bit-compatible -------
struct SpRng {
    int32_t r[34];
    uint32_t ring[31];  // the 31 most recent 32-bit sums
    int idx = 0;
    void seed(unsigned s) {
        if (s == 0) s = 1;
        int32_t t[344 + 31];
        t[0] = (int32_t) s;
        for (int i = 1; i < 31; i++) {
            int64_t v = (16807LL * t[i - 1]) % 2147483647;
            if (v < 0) v += 2147483647;
            t[i] = (int32_t) v;
        }
        for (int i = 31; i < 34; i++) t[i] = t[i - 31];
        for (int i = 34; i < 344; i++) t[i] = (int32_t) ((uint32_t) t[i - 31] + (uint32_t) t[i - 3]);
        for (int i = 0; i < 31; i++) ring[i] = (uint32_t) t[344 - 31 + i];
        idx = 0;
    }
    int next() {
        // ring[(idx + j) % 31] holds o[k-31+j]; new = o[k-31] + o[k-3]
        uint32_t v = ring[idx % 31] + ring[(idx + 28) % 31];
        ring[idx % 31] = v;
        idx = (idx + 1) % 31;
        return (int) (v >> 1);
    }
};

// get_best_record_index, ptAlignment.c:137-177, given the device's deterministic part.
inline int sp_finalize_best(SpRng &rng, int n, const double *score, int prim_idx, int max_idx, int tie_mask,
                            double prim_margin, double min_score, double prim_margin_random) {
    if (n == 1) return 0;
    double max_score = max_idx >= 0 ? score[max_idx] : -1.7976931348623157e308;
    double prim_score = prim_idx >= 0 ? score[prim_idx] : -1.7976931348623157e308;
    int ties[SP_MAX_ALN_PER_GROUP], nt = 0;
    for (int i = 0; i < n; i++)
        if (tie_mask & (1 << i)) ties[nt++] = i;
    if (nt > 1) max_idx = ties[rng.next() % nt];
    const int rnd = rng.next() % 2;
    // "abs(max_score - prim_score)" is the INTEGER abs (ptAlignment.c:172): the double is
    // truncated to int first (cvttsd2si semantics on the reference's x86-64 build).
    double diff = max_score - prim_score;
    int di = (diff >= 2147483648.0 || diff < -2147483648.0 || diff != diff) ? (int) 0x80000000 : (int) diff;
    int ad = di < 0 ? (int) (0u - (unsigned) di) : di;
    if ((double) ad < prim_margin_random) return rnd == 0 ? prim_idx : max_idx;
    if (prim_idx == -1 || max_score <= (prim_score + prim_margin) || max_score < min_score) return prim_idx;
    return max_idx;
}
