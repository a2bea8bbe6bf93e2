// tests/hostsim/hostsim.cpp -- TEST HARNESS ONLY, not part of the product.
//
// The GPU-less build container cannot execute kernels, so the device stage logic
// (secphase_b200/csrc/sp_*.cuh, written as SP_HD functions) is compiled here with g++ and run
// thread-by-thread in plain loops, with the same table layout the CUDA launcher uses.  This
// lets `pytest -m "not gpu"` diff the kernels' logic against the CPU oracle on thousands of
// fuzzed read groups before any GPU time is spent.  libsecphase_b200.so never contains or calls
// this file; the product path fails loudly without a CUDA device.
//
// Build: g++ -O2 -ffp-contract=off -fPIC -shared (see tests/conftest.py).
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../secphase_b200/csrc/sp_blocks.cuh"
#include "../../secphase_b200/csrc/sp_common.h"
#include "../../secphase_b200/csrc/sp_hmm.cuh"
#include "../../secphase_b200/csrc/sp_hmm2.cuh"
#include "../../secphase_b200/csrc/sp_hmmf.cuh"
#include "../../secphase_b200/csrc/sp_markers.cuh"
#include "../../secphase_b200/csrc/sp_plan.h"
#include "../../secphase_b200/csrc/sp_score.cuh"
#include "../../secphase_b200/csrc/sp_walk.cuh"
#define SP_WARP_EMU 1
#include "warp_emu.h"
#include "../../secphase_b200/csrc/sp_walk_warp.cuh"
#include "../../secphase_b200/csrc/sp_group_warp.cuh"

struct HsOut {
    std::vector<int32_t> group;   // [G][SP_GROUP_W]
    std::vector<double> score;    // [A]
    std::vector<int32_t> extent;  // [A][4]
    std::vector<int32_t> mk[3];   // pre, baq, final [.][6]
    std::vector<int64_t> mk_off[3];
    std::vector<int32_t> blocks;  // [.][6]
    std::vector<int64_t> block_off;
    std::vector<int32_t> items;  // [.][SP_HMM_W]
    std::vector<int32_t> rows;   // [.][4] item, t, state, q
    std::vector<double> rows_pmax;  // [.] normalised max posterior of each row
    std::vector<uint8_t> qual;   // full_baq mode: quality pool after the write-back (k_baq_rows, k_baq_zero)
    int64_t cells = 0;
    int32_t err = 0;
    // fast HMM arithmetic (hs_run3, hmm_mode 1): instances the guard band sent to the strict kernel
    int64_t fast_instances = 0, rerun_instances = 0, rerun_threshold = 0, rerun_tie = 0, rerun_numeric = 0;
    // hmm_mode 2, over the consumed rows of un-flagged instances, t = 1 - pmax: largest |t_fast - t_strict| in units
    // of 2^-53 (the grid the reference's own 1 - max/sum lives on) and largest relative difference among rows
    // with t >= 1e-4 (where that grid is finer than 1e-12 relative)
    double max_abs_drift_ulp = 0, max_rel_drift = 0;
};

// the instantiation launch_hmm() (sp_api.cu) picks for a band half-width; `unrolled` = what a full
// warp of that class runs, otherwise what a partial last warp runs
// `il_lane` >= 0: the -w mode's lane-interleaved forward-row block (k_hmm2<.,.,true>), this instance
// playing lane il_lane of its 32-instance set; fsave then points at the set's block
static void hs_hmm2_dispatch(const SpConst &C, const SpHmmIn &in, const SpBand2<1> &B, int bw, double *rinv,
                             double *fsave, SpRow *rows, int n_rows, bool unrolled, int il_lane = -1) {
    const int64_t fss = 2 * (2 * bw + 1);
    const int cls = sp_band_class(bw);
    if (il_lane >= 0) {
        const int64_t rs = (int64_t) (2 * sp_class_bw(cls) + 1) * 64;
        double *fl = fsave + 2 * il_lane;
        switch (sp_class_unrolled_cells(cls)) {
            case 41: sp_hmm2_instance<1, 1, 41, 64>(C, in, B, rinv, fl, rs, rows, n_rows, unrolled); return;
            case 43: sp_hmm2_instance<1, 1, 43, 64>(C, in, B, rinv, fl, rs, rows, n_rows, unrolled); return;
            case 45: sp_hmm2_instance<1, 1, 45, 64>(C, in, B, rinv, fl, rs, rows, n_rows, unrolled); return;
            case 47: sp_hmm2_instance<1, 1, 47, 64>(C, in, B, rinv, fl, rs, rows, n_rows, unrolled); return;
            case 49: sp_hmm2_instance<1, 1, 49, 64>(C, in, B, rinv, fl, rs, rows, n_rows, unrolled); return;
            case 51: sp_hmm2_instance<1, 1, 51, 64>(C, in, B, rinv, fl, rs, rows, n_rows, unrolled); return;
            default: break;
        }
        switch (sp_h2_words(sp_class_bw(cls))) {
            case 1: sp_hmm2_instance<1, 1, 0, 64>(C, in, B, rinv, fl, rs, rows, n_rows, false); break;
            case 2: sp_hmm2_instance<1, 2, 0, 64>(C, in, B, rinv, fl, rs, rows, n_rows, false); break;
            default: sp_hmm2_instance<1, 3, 0, 64>(C, in, B, rinv, fl, rs, rows, n_rows, false); break;
        }
        return;
    }
    switch (sp_class_unrolled_cells(cls)) {
        case 41: sp_hmm2_instance<1, 1, 41>(C, in, B, rinv, fsave, fss, rows, n_rows, unrolled); return;
        case 43: sp_hmm2_instance<1, 1, 43>(C, in, B, rinv, fsave, fss, rows, n_rows, unrolled); return;
        case 45: sp_hmm2_instance<1, 1, 45>(C, in, B, rinv, fsave, fss, rows, n_rows, unrolled); return;
        case 47: sp_hmm2_instance<1, 1, 47>(C, in, B, rinv, fsave, fss, rows, n_rows, unrolled); return;
        case 49: sp_hmm2_instance<1, 1, 49>(C, in, B, rinv, fsave, fss, rows, n_rows, unrolled); return;
        case 51: sp_hmm2_instance<1, 1, 51>(C, in, B, rinv, fsave, fss, rows, n_rows, unrolled); return;
        default: break;
    }
    switch (sp_h2_words(sp_class_bw(cls))) {
        case 1: sp_hmm2_instance<1, 1, 0>(C, in, B, rinv, fsave, fss, rows, n_rows, false); break;
        case 2: sp_hmm2_instance<1, 2, 0>(C, in, B, rinv, fsave, fss, rows, n_rows, false); break;
        default: sp_hmm2_instance<1, 3, 0>(C, in, B, rinv, fsave, fss, rows, n_rows, false); break;
    }
}

// the fast kernel body (sp_hmmf.cuh); bwv = half-width of the virtual band (the product passes the warp's widest
// band; >= bw).  Returns the guard flags, or -1 when the band is beyond the shared-memory kernels.
static int hs_hmmf_dispatch(const SpConst &C, const SpHmmIn &in, int bwv, double *fsave, int64_t fss, SpRow *rows, int n_rows,
                            bool guard_all) {
    if (bwv > SP_H2_MAXBW) return -1;
    const int nc = 2 * bwv + 1;
    std::vector<SpD2> mi((size_t) nc + 2);
    switch (sp_h2_words(bwv)) {
        case 1: return sp_hmmf_instance<1, 1>(C, in, mi.data() + 1, bwv, fsave, fss, rows, n_rows, guard_all);
        case 2: return sp_hmmf_instance<1, 2>(C, in, mi.data() + 1, bwv, fsave, fss, rows, n_rows, guard_all);
        default: return sp_hmmf_instance<1, 3>(C, in, mi.data() + 1, bwv, fsave, fss, rows, n_rows, guard_all);
    }
}

static int g_group_lanes = 0;  // hs_set_group_lanes: K2 + K3 through sp_group_warp.cuh on the emulated warp

extern "C" {

void hs_set_group_lanes(int on) { g_group_lanes = on; }


HsOut *hs_out_create() { return new HsOut(); }
void hs_out_destroy(HsOut *o) { delete o; }

#define HS_GET(name, field, T)                         \
    const T *name(const HsOut *o, int64_t *n) {        \
        *n = (int64_t) o->field.size();                \
        return o->field.data();                        \
    }
HS_GET(hs_group, group, int32_t)
HS_GET(hs_score, score, double)
HS_GET(hs_extent, extent, int32_t)
HS_GET(hs_blocks, blocks, int32_t)
HS_GET(hs_block_off, block_off, int64_t)
HS_GET(hs_items, items, int32_t)
HS_GET(hs_rows, rows, int32_t)
HS_GET(hs_qual, qual, uint8_t)
HS_GET(hs_rows_pmax, rows_pmax, double)
const int32_t *hs_markers(const HsOut *o, int st, int64_t *n) { *n = (int64_t) o->mk[st].size(); return o->mk[st].data(); }
const int64_t *hs_marker_off(const HsOut *o, int st, int64_t *n) { *n = (int64_t) o->mk_off[st].size(); return o->mk_off[st].data(); }
int64_t hs_cells(const HsOut *o) { return o->cells; }
void hs_fast_stats(const HsOut *o, int64_t *st5, double *drift) {
    drift[1] = o->max_abs_drift_ulp;
    st5[0] = o->fast_instances; st5[1] = o->rerun_instances; st5[2] = o->rerun_threshold; st5[3] = o->rerun_tie;
    st5[4] = o->rerun_numeric;
    drift[0] = o->max_rel_drift;
}
int32_t hs_err(const HsOut *o) { return o->err; }

void hs_fill_const(const sp_params *p, SpConst *C) { sp_fill_const(*p, *C); }
int hs_sizeof_const() { return (int) sizeof(SpConst); }

// glibc rand() emulation checks
void *hs_rng_create(unsigned seed) { SpRng *r = new SpRng(); r->seed(seed); return r; }
int hs_rng_next(void *r) { return ((SpRng *) r)->next(); }
void hs_rng_destroy(void *r) { delete (SpRng *) r; }

// The HMM alone (same contract as sp_hmm_batch, one instance).
int hs_hmm(const sp_params *p, const uint8_t *ref, int l_ref, const uint8_t *query, int l_query, int par_bw,
           const int32_t *rows_t, int n_rows, int32_t *state, uint8_t *q, double *pmax, double *s_out) {
    SpConst C;
    sp_fill_const(*p, C);
    int bw = sp_hmm_bw(l_ref, l_query, par_bw);
    int W = 2 * bw + 2;
    std::vector<double> band((size_t) W * 3, 0.0);
    std::vector<uint32_t> code((size_t) W, 0);
    std::vector<double> s((size_t) l_query + 2, 0.0);
    std::vector<double> fsave((size_t) n_rows * 2 * (2 * bw + 1) + 2, 0.0);
    std::vector<SpRow> rows((size_t) (n_rows > 0 ? n_rows : 1));
    for (int i = 0; i < n_rows; i++) {
        rows[i].item = 0; rows[i].t = rows_t[i]; rows[i].entry = -1; rows[i].expected = 0;
        rows[i].state = 0; rows[i].q = 0; rows[i].pmax = 0;
    }
    SpHmmIn in;
    in.ref = ref; in.qbytes = query; in.qseq4 = nullptr; in.q0 = 0;
    in.l_ref = l_ref; in.l_query = l_query; in.par_bw = par_bw;
    SpBand<1> B;
    B.row = band.data(); B.code = code.data(); B.W = W;
    sp_hmm_instance<1, 1>(C, in, B, s.data(), fsave.data(), 2 * (2 * bw + 1), rows.data(), n_rows);
    for (int i = 0; i < n_rows; i++) {
        state[i] = rows[i].state;
        q[i] = (uint8_t) rows[i].q;
        if (pmax) pmax[i] = rows[i].pmax;
    }
    if (s_out) memcpy(s_out, s.data(), sizeof(double) * ((size_t) l_query + 2));
    return 0;
}

// Same, through the shared-memory-band kernel body (sp_hmm2.cuh); bw must be <= SP_H2_MAXBW.
int hs_hmm2(const sp_params *p, const uint8_t *ref, int l_ref, const uint8_t *query, int l_query, int par_bw,
            const int32_t *rows_t, int n_rows, int32_t *state, uint8_t *q, double *pmax, int unrolled) {
    SpConst C;
    sp_fill_const(*p, C);
    const int bw = sp_hmm_bw(l_ref, l_query, par_bw);
    if (bw > SP_H2_MAXBW) return -1;
    const int ncell = sp_h2_cells(bw);
    std::vector<SpD2> mi((size_t) ncell);
    std::vector<double> d((size_t) ncell, 0.0);
    for (auto &v : mi) v.x = v.y = 0.0;
    std::vector<double> rinv((size_t) l_query + 2, 0.0);
    std::vector<double> fsave((size_t) n_rows * 2 * (2 * bw + 1) + 2, 0.0);
    std::vector<SpRow> rows((size_t) (n_rows > 0 ? n_rows : 1));
    for (int i = 0; i < n_rows; i++) {
        rows[i].item = 0; rows[i].t = rows_t[i]; rows[i].entry = -1; rows[i].expected = 0;
        rows[i].state = 0; rows[i].q = 0; rows[i].pmax = 0;
    }
    SpHmmIn in;
    in.ref = ref; in.qbytes = query; in.qseq4 = nullptr; in.q0 = 0;
    in.l_ref = l_ref; in.l_query = l_query; in.par_bw = par_bw;
    SpBand2<1> B;
    B.mi = mi.data() + 1; B.d = d.data() + 1;
    hs_hmm2_dispatch(C, in, B, bw, rinv.data(), fsave.data(), rows.data(), n_rows, unrolled != 0);
    for (int i = 0; i < n_rows; i++) {
        state[i] = rows[i].state;
        q[i] = (uint8_t) rows[i].q;
        if (pmax) pmax[i] = rows[i].pmax;
    }
    return 0;
}

// The fast-arithmetic body (sp_hmmf.cuh) alone; returns the guard flags (>= 0) or -1 when the band class has none.
int hs_hmmf(const sp_params *p, const uint8_t *ref, int l_ref, const uint8_t *query, int l_query, int par_bw,
            const int32_t *rows_t, int n_rows, int32_t *state, uint8_t *q, double *pmax, int extra_bw) {
    SpConst C;
    sp_fill_const(*p, C);
    int bw = sp_hmm_bw(l_ref, l_query, par_bw);
    if (bw > SP_H2_MAXBW) return -1;
    bw = bw + extra_bw > SP_H2_MAXBW ? SP_H2_MAXBW : bw + extra_bw;  // a wider virtual band: the lane of a mixed warp
    const int nc = 2 * bw + 1;
    std::vector<double> fsave((size_t) n_rows * 2 * nc + 2, 0.0);
    std::vector<SpRow> rows((size_t) (n_rows > 0 ? n_rows : 1));
    for (int i = 0; i < n_rows; i++) {
        rows[i].item = 0; rows[i].t = rows_t[i]; rows[i].entry = -1; rows[i].expected = 0;
        rows[i].state = 0; rows[i].q = 0; rows[i].pmax = 0;
    }
    SpHmmIn in;
    in.ref = ref; in.qbytes = query; in.qseq4 = nullptr; in.q0 = 0;
    in.l_ref = l_ref; in.l_query = l_query; in.par_bw = par_bw;
    const int flag = hs_hmmf_dispatch(C, in, bw, fsave.data(), 2 * nc, rows.data(), n_rows, true);
    for (int i = 0; i < n_rows; i++) {
        state[i] = rows[i].state;
        q[i] = (uint8_t) rows[i].q;
        if (pmax) pmax[i] = rows[i].pmax;
    }
    return flag;
}

// Signature of the batch plan (every table offset) computed with n_threads text-scan threads: the plan must
// not depend on the thread count.
uint64_t hs_plan_sig(const sp_flat_batch *b, int indel_threshold, int safe_caps, int n_threads, int *rc_out) {
    SpPlan pl;
    const int rc = sp_make_plan(b, indel_threshold, safe_caps != 0, pl, n_threads);
    if (rc_out) *rc_out = rc;
    if (rc != SP_OK) return 0;
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](const void *p, size_t n) {
        const uint8_t *q = (const uint8_t *) p;
        for (size_t i = 0; i < n; i++) h = (h ^ q[i]) * 1099511628211ull;
    };
    mix(pl.aln_grp.data(), 4 * pl.aln_grp.size()); mix(pl.ops_off.data(), 8 * pl.ops_off.size());
    mix(pl.imk_off.data(), 8 * pl.imk_off.size()); mix(pl.cb_cap.data(), 4 * pl.cb_cap.size());
    mix(pl.gpos_off.data(), 8 * pl.gpos_off.size()); mix(pl.gent_off.data(), 8 * pl.gent_off.size());
    mix(pl.gblk_off.data(), 8 * pl.gblk_off.size()); mix(pl.giv_off.data(), 8 * pl.giv_off.size());
    mix(pl.gblk_cap.data(), 4 * pl.gblk_cap.size());
    return h;
}

// Diagnostic: refined-op / initial-marker counts of every alignment against the planned capacities.
int hs_walk_stats(const sp_flat_batch *b, const sp_params *p, int32_t *out /* [A][5]: n_ops, cap, n_imk, cap, err */) {
    SpConst C;
    sp_fill_const(*p, C);
    SpPlan pl;
    int rc = sp_make_plan(b, p->indel_threshold, false, pl);
    if (rc != SP_OK) return rc;
    const int64_t tag_bytes = b->tag_off[pl.A];
    std::vector<uint8_t> tagbuf((size_t) tag_bytes + 32, 0);
    uint8_t *tag_pool = tagbuf.data();
    while (((uintptr_t) tag_pool) & 15) tag_pool++;
    memcpy(tag_pool, b->tag_pool, (size_t) tag_bytes);
    for (int a = 0; a < pl.A; a++) {
        const int ocap = (int) (pl.ops_off[a + 1] - pl.ops_off[a] - 1), mcap = (int) (pl.imk_off[a + 1] - pl.imk_off[a]);
        std::vector<SpOp> ops((size_t) ocap + 1);
        std::vector<SpInitMarker> imk((size_t) mcap + 1);
        std::vector<SpBlock> cb((size_t) pl.cb_cap[a] + 1);
        SpAlnInfo info;
        sp_walk_alignment(C.indel_threshold, C.min_q, b->flag[a], b->pos[a], b->l_qseq[a], b->n_cigar[a],
                          b->cigar_pool + b->cigar_off[a], tag_pool, b->tag_off[a], b->tag_off[a + 1],
                          b->tag_kind ? b->tag_kind[a] : 0, b->qual_pool + b->qual_off[a], ops.data(), ocap, imk.data(),
                          mcap, cb.data(), pl.cb_cap[a], &info);
        out[a * 5 + 0] = info.n_ops; out[a * 5 + 1] = ocap; out[a * 5 + 2] = info.n_imk; out[a * 5 + 3] = mcap;
        out[a * 5 + 4] = info.err;
    }
    return 0;
}

// Self-test of the warp emulator against what the intrinsics are defined to return: 0 = all good, else the number of
// the first check that failed.
int hs_warp_emu_selftest(void) {
    int bad = 0;
    auto fail = [&](int k) { if (!bad) bad = k; };
    warp_emu::run_warp([&]() {
        const int lane = warp_emu::lane_id();
        if (__shfl_sync(SP_FULL, lane * 3, 7) != 21) fail(1);
        if (__shfl_up_sync(SP_FULL, lane, 2) != (lane >= 2 ? lane - 2 : lane)) fail(2);
        if (__shfl_down_sync(SP_FULL, lane, 5) != (lane + 5 < 32 ? lane + 5 : lane)) fail(3);
        if (__shfl_xor_sync(SP_FULL, lane, 9) != (lane ^ 9)) fail(4);
        if (__ballot_sync(SP_FULL, (lane % 3) == 0) != 0x49249249u) fail(5);
        if (!__any_sync(SP_FULL, lane == 31) || __any_sync(SP_FULL, lane == 32)) fail(6);
        if (__all_sync(SP_FULL, lane < 31) || !__all_sync(SP_FULL, lane < 32)) fail(7);
        if (__reduce_min_sync(SP_FULL, 100 - lane) != 69 || __reduce_max_sync(SP_FULL, lane * lane) != 961) fail(8);
        if (__reduce_add_sync(SP_FULL, lane) != 496 || __reduce_or_sync(SP_FULL, 1u << (lane & 7)) != 0xffu) fail(9);
        if (sp_warp_incl_scan(lane + 1, lane) != (lane + 1) * (lane + 2) / 2) fail(10);
        if (sp_seg_min<4>(31 - lane) != 31 - (lane | 3) || sp_seg_sum<8>(1) != 8 || sp_seg_or<2>(1 << (lane & 1)) != 3) fail(11);
        // lanes may run different amounts of private work between two rendezvous
        int acc = 0;
        for (int k = 0; k < lane * 10; k++) acc += k;
        if (__shfl_sync(SP_FULL, acc, 3) != 435) fail(12);
        __syncwarp();
        if (__fns(0xf0f0u, 0, 3) != 6 || __popc(0xf0f0u) != 8 || __clz(1) != 31 || __ffs(8) != 4) fail(13);
    });
    return bad;
}

// The warp-cooperative walker (sp_walk_warp.cuh, run on warp_emu.h's 32 coroutine lanes) against the serial walker
// for every alignment of a batch.  out[0] alignments, out[1] handled by the warp walker (the others fell back),
// out[2] alignments whose tables differ, out[3] the first of them (-1 none), out[4] what differed there
// (1 ops, 2 markers, 4 blocks, 8 scalars).
int hs_walk_warp_check(const sp_flat_batch *b, const sp_params *p, int64_t *out) {
    SpConst C;
    sp_fill_const(*p, C);
    SpPlan pl;
    int rc = sp_make_plan(b, p->indel_threshold, false, pl);
    if (rc != SP_OK) return rc;
    const int64_t tag_bytes = b->tag_off[pl.A];
    std::vector<uint8_t> tagbuf((size_t) tag_bytes + 64, 0);
    uint8_t *tag_pool = tagbuf.data();
    while (((uintptr_t) tag_pool) & 15) tag_pool++;
    memcpy(tag_pool, b->tag_pool, (size_t) tag_bytes);
    out[0] = pl.A; out[1] = 0; out[2] = 0; out[3] = -1; out[4] = 0;
    for (int a = 0; a < pl.A; a++) {
        const int ocap = (int) (pl.ops_off[a + 1] - pl.ops_off[a] - 1), mcap = (int) (pl.imk_off[a + 1] - pl.imk_off[a]);
        const int ccap = pl.cb_cap[a];
        std::vector<SpOp> ops((size_t) ocap + 1), ops2((size_t) ocap + 1);
        std::vector<SpInitMarker> imk((size_t) mcap + 1), imk2((size_t) mcap + 1);
        std::vector<SpBlock> cb((size_t) ccap + 1), cb2((size_t) ccap + 1);
        SpAlnInfo info, info2;
        memset(&info, 0, sizeof info); memset(&info2, 0, sizeof info2);
        const int tk = b->tag_kind ? b->tag_kind[a] : 0;
        sp_walk_alignment(C.indel_threshold, C.min_q, b->flag[a], b->pos[a], b->l_qseq[a], b->n_cigar[a],
                          b->cigar_pool + b->cigar_off[a], tag_pool, b->tag_off[a], b->tag_off[a + 1], tk,
                          b->qual_pool + b->qual_off[a], ops.data(), ocap, imk.data(), mcap, cb.data(), ccap, &info);
        bool handled = false;
        static uint8_t code_table[256];  // (as k_walk_warp's shared-memory table)
        for (int c = 0; c < 256; c++) code_table[c] = (uint8_t) sp_cs_code((uint32_t) c);
        warp_emu::run_warp([&]() {
            const bool ok = sp_walk_alignment_warp(C.indel_threshold, C.min_q, b->flag[a], b->pos[a], b->l_qseq[a], b->n_cigar[a],
                                                   b->cigar_pool + b->cigar_off[a], tag_pool, b->tag_off[a], b->tag_off[a + 1], tk,
                                                   b->qual_pool + b->qual_off[a], ops2.data(), ocap, imk2.data(), mcap,
                                                   cb2.data(), ccap, &info2, (a & 1) ? code_table : nullptr);
            if (warp_emu::lane_id() == 0) handled = ok;
        });
        if (!handled) continue;
        out[1]++;
        int diff = 0;
        if (memcmp(&info, &info2, sizeof info)) diff |= 8;
        else {
            if (memcmp(ops.data(), ops2.data(), sizeof(SpOp) * (size_t) (info.n_ops + 1))) diff |= 1;
            if (memcmp(imk.data(), imk2.data(), sizeof(SpInitMarker) * (size_t) info.n_imk)) diff |= 2;
            if (memcmp(cb.data(), cb2.data(), sizeof(SpBlock) * (size_t) info.n_cb)) diff |= 4;
        }
        if (diff) {
            if (out[3] < 0) { out[3] = a; out[4] = diff; }
            out[2]++;
        }
    }
    return 0;
}

// Whole marker path for a batch; ref_codes: concatenated contigs (codes 0..4), contig_off[n+1].
int hs_run2(const sp_flat_batch *b, const sp_params *p, const uint8_t *ref_codes, const int64_t *contig_off,
            int n_contigs, int safe_caps, unsigned rng_seed, int full_baq, HsOut *out);
int hs_run3(const sp_flat_batch *b, const sp_params *p, const uint8_t *ref_codes, const int64_t *contig_off,
            int n_contigs, int safe_caps, unsigned rng_seed, int full_baq, int hmm_mode, HsOut *out);
int hs_run(const sp_flat_batch *b, const sp_params *p, const uint8_t *ref_codes, const int64_t *contig_off,
           int n_contigs, int safe_caps, unsigned rng_seed, HsOut *out) {
    return hs_run2(b, p, ref_codes, contig_off, n_contigs, safe_caps, rng_seed, 0, out);
}
// full_baq != 0: the --writeBam mode of sp_set_write_qual (rows for every base of the write-back range)
int hs_run2(const sp_flat_batch *b, const sp_params *p, const uint8_t *ref_codes, const int64_t *contig_off,
            int n_contigs, int safe_caps, unsigned rng_seed, int full_baq, HsOut *out) {
    return hs_run3(b, p, ref_codes, contig_off, n_contigs, safe_caps, rng_seed, full_baq, 0, out);
}
// hmm_mode: 0 strict kernels only; 1 the product's default: fast kernel where the band class has one (never in
// -w mode), instances flagged by its guard band re-run by the strict kernel; 2 = 1 plus statistics: every fast
// instance is ALSO run strictly and the integers / drift of the un-flagged ones compared (out->err bit 0x200 on
// a difference the guard band missed)
int hs_run3(const sp_flat_batch *b, const sp_params *p, const uint8_t *ref_codes, const int64_t *contig_off,
            int n_contigs, int safe_caps, unsigned rng_seed, int full_baq, int hmm_mode, HsOut *out) {
    (void) n_contigs;
    SpConst C;
    sp_fill_const(*p, C);
    C.full_baq = full_baq != 0;
    SpPlan pl;
    int rc = sp_make_plan(b, p->indel_threshold, safe_caps != 0, pl);
    if (rc != SP_OK) return rc;
    const int G = pl.G, A = pl.A;
    // 16-byte padded copy of the tag pool (SpByteReader reads aligned 16-byte words)
    const int64_t tag_bytes = b->tag_off[A];
    std::vector<uint8_t> tagbuf((size_t) tag_bytes + 32, 0);
    uint8_t *tag_pool = tagbuf.data();
    while (((uintptr_t) tag_pool) & 15) tag_pool++;
    memcpy(tag_pool, b->tag_pool, (size_t) tag_bytes);

    std::vector<SpOp> ops((size_t) pl.total_ops + 1);
    std::vector<SpInitMarker> imk((size_t) pl.total_imk + 1);
    std::vector<SpAlnInfo> info((size_t) A);
    std::vector<SpBlock> blk((size_t) pl.total_blk + 1);
    std::vector<SpIv> iv((size_t) pl.total_iv + 1);
    std::vector<int32_t> nb((size_t) A, 0);
    std::vector<int32_t> gpos((size_t) pl.total_pos + 1);
    std::vector<SpEntry> ent((size_t) pl.total_ent + 1);
    std::vector<int32_t> res((size_t) pl.total_ent + 1, SP_RES_RAW);
    std::vector<int32_t> baq((size_t) pl.total_ent + 1, 0);
    std::vector<int32_t> fin((size_t) pl.total_ent * 6 + 6);
    std::vector<SpGroupOut> gout((size_t) G);
    std::vector<int32_t> gP((size_t) G, 0);
    std::vector<double> score((size_t) A, 0.0);
    std::vector<SpEmitCounts> gcnt((size_t) G);

    // K1: one "thread" per alignment; confident blocks land in the group's block workspace
    for (int a = 0; a < A; a++) {
        const int g = pl.aln_grp[a];
        const int i = a - b->grp_aln_off[g];
        SpBlock *cb = blk.data() + pl.gblk_off[g] + (int64_t) i * pl.gblk_cap[g];
        sp_walk_alignment(C.indel_threshold, C.min_q, b->flag[a], b->pos[a], b->l_qseq[a], b->n_cigar[a],
                          b->cigar_pool + b->cigar_off[a], tag_pool, b->tag_off[a], b->tag_off[a + 1],
                          b->tag_kind ? b->tag_kind[a] : 0, b->qual_pool + b->qual_off[a],
                          ops.data() + pl.ops_off[a], (int) (pl.ops_off[a + 1] - pl.ops_off[a] - 1),
                          imk.data() + pl.imk_off[a], (int) (pl.imk_off[a + 1] - pl.imk_off[a]), cb,
                          pl.gblk_cap[g], &info[a]);
        nb[a] = info[a].n_cb;
    }
    auto view = [&](int g) {
        SpGroupAlnView V;
        V.a0 = b->grp_aln_off[g];
        V.n = b->grp_aln_off[g + 1] - V.a0;
        V.flag = b->flag;
        V.l_qseq = b->l_qseq;
        V.qual_off = b->qual_off;
        V.qual_pool = b->qual_pool;
        V.info = info.data();
        V.ops_off = pl.ops_off.data();
        V.ops = ops.data();
        V.imk_off = pl.imk_off.data();
        V.imk = imk.data();
        return V;
    };
    out->mk_off[0].push_back(0);
    auto work = [&](int g, const SpGroupAlnView &V) {
        SpBlockWork W;
        W.cap = pl.gblk_cap[g];
        W.ab = blk.data() + pl.gblk_off[g];
        W.nb = nb.data() + V.a0;
        W.cons_a = iv.data() + pl.giv_off[g];
        W.cons_b = W.cons_a + W.cap;
        W.flank = W.cons_b + W.cap;
        return W;
    };
    // K2 + K3 with a lane per alignment (sp_group_warp.cuh on the emulated warp), as k_group_lanes<2|4|16> runs them
    std::vector<SpGroupOut> lane_out((size_t) G);
    if (g_group_lanes) {
        int first = 0;
        for (int cls = 0; cls < 3; cls++) {
            const int wid = cls == 0 ? 2 : cls == 1 ? 4 : 16, count = pl.glist_n[cls], per_warp = 32 / wid;
            for (int w0 = 0; w0 < count; w0 += per_warp) {
                warp_emu::run_warp([&]() {
                    const int slot = w0 + warp_emu::lane_id() / wid;
                    const bool has = slot < count;
                    const int g = has ? pl.glist[(size_t) (first + slot)] : 0;
                    SpGroupAlnView V = view(g);
                    SpBlockWork W = work(g, V);
                    int32_t *gp = gpos.data() + pl.gpos_off[g];
                    SpEntry *en = ent.data() + pl.gent_off[g];
                    const int pc = (int) (pl.gpos_off[g + 1] - pl.gpos_off[g]);
                    if (wid == 2) sp_group_lanes<2>(C, V, has, gp, en, pc, W, &lane_out[g], &gP[g]);
                    else if (wid == 4) sp_group_lanes<4>(C, V, has, gp, en, pc, W, &lane_out[g], &gP[g]);
                    else sp_group_lanes<16>(C, V, has, gp, en, pc, W, &lane_out[g], &gP[g]);
                });
            }
            first += count;
        }
    }
    // K2 + K3 (consensus + count pass): one "thread" per group
    for (int g = 0; g < G; g++) {
        SpGroupAlnView V = view(g);
        SpGroupOut &o = gout[g];
        memset(&o, 0, sizeof(o));
        int err = 0;
        int P;
        if (g_group_lanes) {
            o = lane_out[g];
            P = gP[g];
            err = o.err;
        } else {
            for (int i = 0; i < V.n; i++) err |= info[V.a0 + i].err;
            int32_t counts[4];
            P = sp_group_markers(V, gpos.data() + pl.gpos_off[g], ent.data() + pl.gent_off[g],
                                 (int) (pl.gpos_off[g + 1] - pl.gpos_off[g]), counts, &err);
            gP[g] = P;
            o.n_init = counts[0]; o.n_after_allmm = counts[1]; o.n_filled = counts[2]; o.n_after_ins = counts[3];
        }
        for (int pi = 0; pi < P; pi++)
            for (int i = 0; i < V.n; i++) {
                const SpEntry &e = ent[pl.gent_off[g] + (int64_t) pi * V.n + i];
                int32_t row[6] = {i, gpos[pl.gpos_off[g] + pi], e.base_idx, e.q, e.flags & 1, e.ref_pos};
                out->mk[0].insert(out->mk[0].end(), row, row + 6);
            }
        out->mk_off[0].push_back((int64_t) out->mk[0].size() / 6);
        SpBlockWork W = work(g, V);
        int margin = C.flank_margin, conf_len = 1;
        bool scored = false;
        SpEmitCounts cnt;
        memset(&cnt, 0, sizeof(cnt));
        if (g_group_lanes) {
            margin = o.margin_eff;
            conf_len = o.conf_len;
        } else if (P > 0) {
            conf_len = sp_consensus_loop(C, V, P, gpos.data() + pl.gpos_off[g], W, &margin, &err);
        }
        if (P > 0) {
            if (conf_len > 0 || !C.consensus) {
                scored = true;
                if (C.baq_flag) {
                    for (int i = 0; i < V.n; i++) {
                        const int a = V.a0 + i;
                        sp_emit_alignment<false>(C, V, i, P, ent.data() + pl.gent_off[g],
                                                 W.ab + (int64_t) i * W.cap, W.nb[i], contig_off[b->tid[a]],
                                                 0, cnt, nullptr, nullptr, 0, nullptr, 0, 0);
                    }
                }
            }
        }
        if (P == 0)  // the reference never builds confident blocks for a group without markers (secphase.c:161)
            for (int i = 0; i < V.n; i++) W.nb[i] = 0;
        gcnt[g] = cnt;
        o.margin_eff = margin;
        o.conf_len = conf_len;
        if (g_group_lanes && (o.scored != (scored ? 1 : 0))) err |= 0x4000;  // (the lanes' own verdict must agree)
        o.scored = scored ? 1 : 0;
        o.err = err;
    }
    // scan
    std::vector<int32_t> item_off((size_t) G + 1, 0), row_off((size_t) G + 1, 0);
    for (int g = 0; g < G; g++) {
        item_off[g + 1] = item_off[g] + gcnt[g].n_items;
        row_off[g + 1] = row_off[g] + gcnt[g].n_rows;
        out->cells += gcnt[g].cells;
    }
    std::vector<SpItem> items((size_t) item_off[G] + 1);
    std::vector<SpRow> rows((size_t) row_off[G] + 1);
    // emit
    for (int g = 0; g < G; g++) {
        SpGroupAlnView V = view(g);
        if (!gout[g].scored || !C.baq_flag) continue;
        SpEmitCounts cnt;
        memset(&cnt, 0, sizeof(cnt));
        for (int i = 0; i < V.n; i++) {
            const int a = V.a0 + i;
            sp_emit_alignment<true>(C, V, i, gP[g], ent.data() + pl.gent_off[g],
                                    blk.data() + pl.gblk_off[g] + (int64_t) i * pl.gblk_cap[g], nb[a],
                                    contig_off[b->tid[a]], 0, cnt, res.data() + pl.gent_off[g], items.data(),
                                    item_off[g], rows.data(), row_off[g], 0);
        }
        if (cnt.n_items != gcnt[g].n_items || cnt.n_rows != gcnt[g].n_rows) out->err |= 0x100;
    }
    if (full_baq) {  // k_fill_rows
        for (int w = 0; w < item_off[G]; w++) {
            const SpItem &it = items[w];
            const int a = it.aln;
            const int blk_rfs = (int) (it.ref_off - contig_off[b->tid[a]]);
            for (int k = 0; k < it.n_rows; k++)
                rows[it.row0 + k] = sp_fill_row(it, w, k, ops.data() + pl.ops_off[a], (b->flag[a] & SP_FREVERSE) != 0, blk_rfs);
        }
    }
    // K4: one "lane" per item
    bool all_h2 = true;
    for (int it = 0; it < item_off[G]; it++) {
        const SpItem &I = items[it];
        if (sp_hmm_bw(I.l_ref, I.l_query, I.par_bw) > SP_H2_MAXBW || I.l_query > 2046) all_h2 = false;
    }
    for (int it = 0; it < item_off[G]; it++) {
        const SpItem &I = items[it];
        const int bw = sp_hmm_bw(I.l_ref, I.l_query, I.par_bw);
        const int W = 2 * bw + 2;
        std::vector<double> band((size_t) W * 3, 0.0);
        std::vector<uint32_t> code((size_t) W, 0);
        std::vector<double> s((size_t) I.l_query + 2, 0.0);
        // the launcher's choice (run_phase_b): lane-interleaved rows in -w mode when every instance runs
        // the shared-memory-band kernel
        const int il_lane = (full_baq && all_h2) ? it % 32 : -1;
        std::vector<double> fsave(il_lane >= 0 ? (size_t) I.n_rows * (2 * sp_class_bw(sp_band_class(bw)) + 1) * 64 + 2
                                               : (size_t) I.n_rows * 2 * (2 * bw + 1) + 2, 0.0);
        SpHmmIn in;
        in.ref = ref_codes + I.ref_off;
        in.qbytes = nullptr;
        in.qseq4 = b->seq_pool + b->seq_off[I.aln];
        in.q0 = I.q_sqs;
        in.l_ref = I.l_ref; in.l_query = I.l_query; in.par_bw = I.par_bw;
        bool strict = true;
        std::vector<SpRow> fast_rows;
        if (hmm_mode != 0 && !full_baq && sp_hmmf_class_cells(sp_band_class(bw)) != 0 && I.n_rows > 0) {  // launch_hmm's choice
            // every third instance plays a lane of a mixed warp (virtual band = its class's widest)
            const int bwv = it % 3 == 2 ? sp_class_bw(sp_band_class(bw)) : bw;
            const int nc = 2 * bwv + 1;
            std::vector<double> ff((size_t) I.n_rows * 2 * nc + 2, 0.0);
            const int fl = hs_hmmf_dispatch(C, in, bwv, ff.data(), 2 * nc, rows.data() + I.row0, I.n_rows, false);  // as k_hmmf does
            out->fast_instances++;
            strict = fl != 0;
            if (fl) {
                out->rerun_instances++;
                if (fl & SP_HMMF_NEAR_THRESHOLD) out->rerun_threshold++;
                if (fl & SP_HMMF_NEAR_TIE) out->rerun_tie++;
                if (fl & SP_HMMF_NUMERIC) out->rerun_numeric++;
            } else if (hmm_mode == 2) {
                fast_rows.assign(rows.begin() + I.row0, rows.begin() + I.row0 + I.n_rows);
                strict = true;
            }
        }
        if (!strict) {
        } else if (bw <= SP_H2_MAXBW) {  // same dispatch as launch_hmm() in sp_api.cu
            const int ncell = sp_h2_cells(bw);
            std::vector<SpD2> mi((size_t) ncell);
            std::vector<double> dd((size_t) ncell, 0.0);
            for (auto &v : mi) v.x = v.y = 0.0;
            SpBand2<1> B2;
            B2.mi = mi.data() + 1; B2.d = dd.data() + 1;
            hs_hmm2_dispatch(C, in, B2, bw, s.data(), fsave.data(), rows.data() + I.row0, I.n_rows, true, il_lane);
        } else {
            SpBand<1> B;
            B.row = band.data(); B.code = code.data(); B.W = W;
            sp_hmm_instance<1, 1>(C, in, B, s.data(), fsave.data(), 2 * (2 * bw + 1), rows.data() + I.row0, I.n_rows);
        }
        for (size_t k = 0; k < fast_rows.size(); k++) {  // hmm_mode 2: what the guard band let through must be exact
            const SpRow &S = rows[I.row0 + k], &F = fast_rows[k];
            // what the pipeline consumes of a row (sp_resolve_q): 0 unless the MAP state is M at the expected column
            auto used = [](const SpRow &R) { return ((R.state & 3) != 0 || (R.state >> 2) != R.expected) ? 0 : (R.q < 93 ? R.q : 93); };
            if (used(S) != used(F)) out->err |= 0x200;
            const double ts = 1. - S.pmax, tf = 1. - F.pmax;
            const double ad = tf > ts ? tf - ts : ts - tf;
            if (ad * 9007199254740992. > out->max_abs_drift_ulp) out->max_abs_drift_ulp = ad * 9007199254740992.;
            if (ts >= 1e-4 && ad / ts > out->max_rel_drift) out->max_rel_drift = ad / ts;
        }
        int32_t irow[SP_HMM_W] = {I.aln, I.l_ref, I.l_query, I.par_bw, I.blk, I.row0, I.n_rows, 0};
        out->items.insert(out->items.end(), irow, irow + SP_HMM_W);
    }
    for (int r = 0; r < row_off[G]; r++) {
        int32_t rr[4] = {rows[r].item, rows[r].t, rows[r].state, rows[r].q};
        out->rows.insert(out->rows.end(), rr, rr + 4);
        out->rows_pmax.push_back(rows[r].pmax);
    }
    if (full_baq) {  // run_phase_b: D2D copy of the raw pool, k_baq_rows, k_baq_zero
        out->qual.assign(b->qual_pool, b->qual_pool + b->qual_off[A]);
        for (int r = 0; r < row_off[G]; r++) {
            const SpItem &I = items[rows[r].item];
            const int64_t q = b->qual_off[I.aln] + I.q_sqs + rows[r].t;
            out->qual[(size_t) q] = sp_baq_row_qual(C, rows[r], b->qual_pool[q]);
        }
        for (int g = 0; g < G; g++) {
            SpGroupAlnView V = view(g);
            sp_baq_zero_group(V, gP[g], ent.data() + pl.gent_off[g], res.data() + pl.gent_off[g], out->qual.data());
        }
    }
    // K5 + host finalisation
    SpRng rng;
    rng.seed(rng_seed);
    out->mk_off[1].push_back(0);
    out->mk_off[2].push_back(0);
    out->block_off.push_back(0);
    std::vector<int> lane_nf((size_t) G, 0);
    if (g_group_lanes) {  // K5 as k_score_lanes<2|4|16> runs it
        int first = 0;
        for (int cls = 0; cls < 3; cls++) {
            const int wid = cls == 0 ? 2 : cls == 1 ? 4 : 16, count = pl.glist_n[cls], per_warp = 32 / wid;
            for (int w0 = 0; w0 < count; w0 += per_warp) {
                warp_emu::run_warp([&]() {
                    const int slot = w0 + warp_emu::lane_id() / wid;
                    const bool has = slot < count;
                    const int g = has ? pl.glist[(size_t) (first + slot)] : 0;
                    SpGroupAlnView V = view(g);
                    const int32_t *gp = gpos.data() + pl.gpos_off[g];
                    const SpEntry *en = ent.data() + pl.gent_off[g];
                    const int32_t *rs = res.data() + pl.gent_off[g];
                    double *sc = score.data() + V.a0;
                    int32_t *fn = fin.data() + pl.gent_off[g] * 6, *bq = baq.data() + pl.gent_off[g];
                    const bool scd = gout[g].scored != 0;
                    int nf;
                    if (wid == 2) nf = sp_score_lanes<2>(C, V, has, gP[g], gp, en, rs, rows.data(), scd, sc, fn, bq);
                    else if (wid == 4) nf = sp_score_lanes<4>(C, V, has, gP[g], gp, en, rs, rows.data(), scd, sc, fn, bq);
                    else nf = sp_score_lanes<16>(C, V, has, gP[g], gp, en, rs, rows.data(), scd, sc, fn, bq);
                    if (has && warp_emu::lane_id() % wid == 0) lane_nf[(size_t) g] = nf;
                });
            }
            first += count;
        }
    }
    for (int g = 0; g < G; g++) {
        SpGroupAlnView V = view(g);
        SpGroupOut &o = gout[g];
        const int P = gP[g];
        int nf = g_group_lanes ? lane_nf[(size_t) g]
                               : sp_score_group(C, V, P, gpos.data() + pl.gpos_off[g], ent.data() + pl.gent_off[g],
                                                res.data() + pl.gent_off[g], rows.data(), o.scored != 0, score.data() + V.a0,
                                                fin.data() + pl.gent_off[g] * 6, baq.data() + pl.gent_off[g]);
        o.n_final = nf;
        sp_select(V, score.data() + V.a0, p->prim_margin_score, (double) p->min_score, &o);
        int best = sp_finalize_best(rng, V.n, score.data() + V.a0, o.prim_idx, o.max_idx, o.tie_mask,
                                    p->prim_margin_score, (double) p->min_score, p->prim_margin_random);
        int32_t grow[SP_GROUP_W] = {best, o.prim_idx, o.n_init, o.n_after_allmm, o.n_filled, o.n_after_ins,
                                    o.margin_eff, o.conf_len, o.n_final, o.scored};
        out->group.insert(out->group.end(), grow, grow + SP_GROUP_W);
        out->err |= o.err;
        for (int pi = 0; pi < P; pi++)
            for (int i = 0; i < V.n; i++) {
                const SpEntry &e = ent[pl.gent_off[g] + (int64_t) pi * V.n + i];
                int32_t row[6] = {i, gpos[pl.gpos_off[g] + pi], e.base_idx, baq[pl.gent_off[g] + (int64_t) pi * V.n + i],
                                  e.flags & 1, e.ref_pos};
                out->mk[1].insert(out->mk[1].end(), row, row + 6);
            }
        out->mk_off[1].push_back((int64_t) out->mk[1].size() / 6);
        out->mk[2].insert(out->mk[2].end(), fin.data() + pl.gent_off[g] * 6, fin.data() + pl.gent_off[g] * 6 + (int64_t) nf * 6);
        out->mk_off[2].push_back((int64_t) out->mk[2].size() / 6);
        for (int i = 0; i < V.n; i++) {
            const int a = V.a0 + i;
            out->score.push_back(score[a]);
            int32_t ex[4] = {info[a].rfs, info[a].rfe, info[a].rds_f, info[a].rde_f};
            out->extent.insert(out->extent.end(), ex, ex + 4);
            const SpBlock *bl = blk.data() + pl.gblk_off[g] + (int64_t) i * pl.gblk_cap[g];
            for (int k = 0; k < nb[a]; k++) {
                int32_t br[6] = {bl[k].rfs, bl[k].rfe, bl[k].sqs, bl[k].sqe, bl[k].rds_f, bl[k].rde_f};
                out->blocks.insert(out->blocks.end(), br, br + 6);
            }
            out->block_off.push_back((int64_t) out->blocks.size() / 6);
        }
    }
    return 0;
}
}  // extern "C"
