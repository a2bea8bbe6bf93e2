// sp_walk_warp.cuh -- stage K1, warp-cooperative: one warp walks one alignment's CIGAR + cs tag.
//
// Same outputs as sp_walk_alignment (sp_walk.cuh), which stays the specification and the fallback: the refined op
// table, alignment extents, initial markers and confident blocks of
//   ptCigarIt_next / _next_cs          (cigar_it.c:213-308, 145-211)
//   ptAlignment_init_coordinates       (ptAlignment.c:42-95)
//   ptMarker_get_initial_markers       (ptMarker.c:42-75)
//   find_confident_blocks              (ptMarker.c:328-395).
// The serial walker spends ~10^3 instructions per cs token with a quarter of its lanes active (32 different
// alignments per warp, every lane in a different token); here the 32 lanes of a warp look at 32 consecutive
// bytes of ONE alignment's cs text:
//   1. tokens are found by ballots (a token starts at ':', '+', '-' or at a '*' that does not continue a run of
//      "*xy" groups), ':' numbers are evaluated by a short segmented scan, each token end writes its op;
//   2. the ops, 32 at a time: coordinates are prefix sums; the tokens are checked against the CIGAR -- between its
//      end clips the CIGAR must be exactly the run structure of the tokens (an M op = a maximal run of ':'/'*'
//      tokens of the same total length, an I/D op = a '+'/'-' token of the same length; with =/X CIGARs every
//      token is one op).  For such input -- everything an aligner writes -- the refined ops ARE the end clips plus
//      the tokens;
//   3. in the same pass markers, extents and confident blocks follow from the coordinates with
//      ballots and scans.
// Anything else (MD tags, text outside the cs grammar, tokens that disagree with the CIGAR, N/P ops, clips in odd
// places, ':0') makes the function return false -- what it wrote by then lies in the alignment's own tables, which
// the serial walker, whose behaviour on such input is the reference's, then rewrites (k_walk_list).
#pragma once
#include "sp_walk.cuh"

// (tests/hostsim compiles this file for the host with SP_WARP_EMU: tests/hostsim/warp_emu.h runs the 32 lanes as
// coroutines and supplies the *_sync intrinsics)
#if defined(__CUDACC__) || defined(SP_WARP_EMU)

#define SP_FULL 0xffffffffu
#if defined(__CUDACC__)
#define SP_WD __device__ __forceinline__
#define SP_WDN __device__
#define SP_LANE() ((int) (threadIdx.x & 31))
#else
#define SP_WD inline
#define SP_WDN inline
#define SP_LANE() (warp_emu::lane_id())
#endif

SP_WD int sp_warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(SP_FULL, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// One M/=/X/I/D/S/H op's steps: bit0 advances the stored SEQ, bit1 the reference, bit2 the read (cigar_it.c:224-291)
SP_WD int sp_op_step_mask(int op) {
    return (op == SP_CMATCH || op == SP_CEQUAL || op == SP_CDIFF) ? 7 : op == SP_CINS ? 5 : op == SP_CDEL ? 2
           : op == SP_CSOFT ? 5 : op == SP_CHARD ? 4 : 0;
}

// true: tables written; false: run the serial walker instead.
SP_WDN bool sp_walk_alignment_warp(int indel_threshold, int min_q, int flag, int pos, int l_qseq, int n_cigar,
                                       const uint32_t *__restrict__ cigar, const uint8_t *__restrict__ tag, int64_t tag_beg,
                                       int64_t tag_end, int tag_kind, const uint8_t *__restrict__ qual, SpOp *ops, int ops_cap,
                                       SpInitMarker *imk, int imk_cap, SpBlock *cb, int cb_cap, SpAlnInfo *info) {
    const int lane = SP_LANE();
    const bool rev = (flag & SP_FREVERSE) != 0;
    if (tag_kind != 0 || n_cigar < 1 || tag_end <= tag_beg) return false;
    // ---------------------------------------------------------------- 0. end clips, kind of CIGAR
    int n_lead = 0, n_trail = 0, lead_h = 0, lead_s = 0, trail_s = 0, trail_h = 0;
    {
        const uint32_t c0 = cigar[0], cl = cigar[n_cigar - 1];
        if ((c0 & 15) == SP_CHARD) {
            lead_h = (int) (c0 >> 4);
            n_lead = 1;
            if (n_cigar > 1 && (cigar[1] & 15) == SP_CSOFT) { lead_s = (int) (cigar[1] >> 4); n_lead = 2; }
        } else if ((c0 & 15) == SP_CSOFT) {
            lead_s = (int) (c0 >> 4);
            n_lead = 1;
        }
        if (n_cigar > n_lead) {
            if ((cl & 15) == SP_CHARD) {
                trail_h = (int) (cl >> 4);
                n_trail = 1;
                if (n_cigar - 1 > n_lead && (cigar[n_cigar - 2] & 15) == SP_CSOFT) { trail_s = (int) (cigar[n_cigar - 2] >> 4); n_trail = 2; }
            } else if ((cl & 15) == SP_CSOFT) {
                trail_s = (int) (cl >> 4);
                n_trail = 1;
            }
        }
    }
    const int n_core = n_cigar - n_lead - n_trail;
    if (n_core < 1) return false;
    const uint32_t *core = cigar + n_lead;
    bool eqx = false;
    {
        bool bad = false, has_m = false, has_eqx = false;
        for (int k = lane; k < n_core; k += 32) {
            const int op = (int) (core[k] & 15);
            const bool mt = op == SP_CMATCH || op == SP_CEQUAL || op == SP_CDIFF;
            if (!(mt || op == SP_CINS || op == SP_CDEL)) bad = true;  // clip inside, N, P, B
            if (op == SP_CMATCH) {
                has_m = true;
                if (k > 0) {
                    const int pv = (int) (core[k - 1] & 15);
                    if (pv == SP_CMATCH || pv == SP_CEQUAL || pv == SP_CDIFF) bad = true;  // adjacent M-type ops: token run = several ops
                }
            }
            if (op == SP_CEQUAL || op == SP_CDIFF) {
                has_eqx = true;
                if (k > 0 && (core[k - 1] & 15) == SP_CMATCH) bad = true;
            }
            if ((core[k] >> 4) == 0) bad = true;  // zero-length op: the serial walker stops there
        }
        if (__any_sync(SP_FULL, bad)) return false;
        has_m = __any_sync(SP_FULL, has_m);
        has_eqx = __any_sync(SP_FULL, has_eqx);
        if (has_m && has_eqx) return false;
        eqx = has_eqx;
    }
    // ---------------------------------------------------------------- 1. tokens
    // Bytes are classified once (codes below); a lane sees its own code, the three before it and the one after.
    // Carried across chunks (warp-uniform): codes of the last three bytes, the token the chunk begins in, the value
    // of the number it begins in.
    enum { K_OTHER = 0, K_DIGIT = 1, K_LOWER = 2, K_COLON = 3, K_STAR = 4, K_PLUS = 5, K_MINUS = 6 };
    auto code_of = [](uint32_t c) -> uint32_t {  // (selects, not branches)
        uint32_t k = (c - '0') < 10u ? K_DIGIT : K_OTHER;
        k = (c - 'a') < 26u ? K_LOWER : k;
        k = c == ':' ? K_COLON : k;
        k = c == '*' ? K_STAR : k;
        k = c == '+' ? K_PLUS : k;
        k = c == '-' ? K_MINUS : k;
        return k;
    };
    uint32_t pw = 0;                         // codes of the three bytes before the chunk, 4 bits each, nearest lowest
    int cur_kind = 0;                        // code of the start character of the token the chunk begins in
    int cur_start = 0;                       // (text offsets are relative to tag_beg)
    uint32_t num_in = 0;                     // value of the number the chunk begins in
    int n_tok = 0;
    bool invalid = false;
    const uint8_t *text = tag + tag_beg;
    const int n_text = (int) (tag_end - tag_beg);
    const int n_chunks = (n_text + 1 + 31) >> 5;       // one virtual terminator byte
    uint32_t c_nx = lane < n_text ? text[lane] : 0;    // the chunk ahead is loaded and classified one iteration early
    uint32_t k_nx = code_of(c_nx);
    for (int ch = 0; ch < n_chunks; ch++) {
        const int cbase = ch << 5;
        const int p = cbase + lane;
        const bool in_range = p < n_text;
        const uint32_t c0 = c_nx, k0 = k_nx;
        c_nx = p + 32 < n_text ? text[p + 32] : 0;
        k_nx = code_of(c_nx);
        // codes at p, p-1, p-2, p-3 in 4-bit fields
        uint32_t kw = k0 | (__shfl_up_sync(SP_FULL, k0, 1) << 4);
        if (lane == 0) kw = k0 | ((pw & 15u) << 4);
        uint32_t kw4 = kw | (__shfl_up_sync(SP_FULL, kw, 2) << 8);
        if (lane == 0) kw4 = kw | ((pw >> 4) << 8);
        if (lane == 1) kw4 = kw | ((pw & 255u) << 8);
        const uint32_t k1 = (kw4 >> 4) & 15u, k2 = (kw4 >> 8) & 15u, k3 = (kw4 >> 12) & 15u;
        uint32_t kx = __shfl_down_sync(SP_FULL, k0, 1);
        const uint32_t kx31 = __shfl_sync(SP_FULL, k_nx, 0);
        if (lane == 31) kx = kx31;
        const bool isd = k0 == K_DIGIT, isl = k0 == K_LOWER;
        const bool start = k0 == K_COLON || k0 >= K_PLUS || (k0 == K_STAR && !(k3 == K_STAR && k2 == K_LOWER && k1 == K_LOWER));
        const bool start_nx = kx == K_COLON || kx >= K_PLUS || (kx == K_STAR && !(k2 == K_STAR && k1 == K_LOWER && k0 == K_LOWER));
        const bool end = in_range && (p + 1 == n_text || start_nx);
        // the cs grammar  (:[0-9]+ | \*[a-z][a-z] | [+-][a-z]+)*  checked looking backwards (the byte at tag_end is a
        // virtual terminator)
        if (p <= n_text) {
            bool bad = false;
            if (in_range) {
                if (k0 == K_OTHER) bad = true;
                if (p == 0 && k0 < K_COLON) bad = true;
                if (isd && !(k1 == K_COLON || k1 == K_DIGIT)) bad = true;
                if (isl && !(k1 >= K_STAR || k1 == K_LOWER)) bad = true;
                if (isl && k1 == K_LOWER && k2 == K_LOWER && k3 == K_STAR) bad = true;  // third letter of a '*' group
            }
            if (!isd && k1 == K_COLON) bad = true;                   // ':' without a number
            if (!isl && k1 >= K_STAR) bad = true;                    // sign or '*' without a letter
            if (!isl && k1 == K_LOWER && k2 == K_STAR) bad = true;   // '*' group with a single letter
            if (bad) invalid = true;
        }
        // which token does my byte belong to
        const uint32_t le = lane == 31 ? 0xffffffffu : ((2u << lane) - 1);
        const uint32_t sm = __ballot_sync(SP_FULL, start);
        const uint32_t below = sm & le;
        const int src = below ? 31 - __clz(below) : 0;
        const uint32_t ck = __shfl_sync(SP_FULL, k0, src);
        const int tk_kind = below ? (int) ck : cur_kind;
        const int tk_start = below ? cbase + src : cur_start;
        // numbers: value = num_in * m + a after a scan of (m, a) over 8 bytes (longer numbers: not a read's)
        uint32_t m = isd ? 10u : 0u, a = isd ? c0 - '0' : 0u;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            const uint32_t pm = __shfl_up_sync(SP_FULL, m, o), pa = __shfl_up_sync(SP_FULL, a, o);
            if (lane >= o) { a = pa * m + a; m = pm * m; }
        }
        const uint32_t num = num_in * m + a;
        const uint32_t em = __ballot_sync(SP_FULL, end);
        if (end) {
            int tlen, op;
            if (tk_kind == K_COLON) { tlen = (int) num; op = SP_CEQUAL; if (p - tk_start > 8) invalid = true; }
            else if (tk_kind == K_STAR) { tlen = (p - tk_start + 1) / 3; op = SP_CDIFF; }
            else { tlen = p - tk_start; op = tk_kind == K_PLUS ? SP_CINS : SP_CDEL; }
            if (tlen <= 0) invalid = true;  // ":0" ends the serial walk
            const int idx = n_lead + n_tok + __popc(em & (le >> 1));
            if (idx < ops_cap) ops[idx].oplen = (uint32_t) op | ((uint32_t) tlen << 4);  // coordinates follow in step 2
        }
        // carries
        n_tok += __popc(em);
        if (sm) {
            const int ls = 31 - __clz(sm);
            cur_kind = (int) __shfl_sync(SP_FULL, k0, ls);
            cur_start = cbase + ls;
        }
        num_in = __shfl_sync(SP_FULL, num, 31);
        pw = __shfl_sync(SP_FULL, kw4, 31) & 0xfffu;
    }
    if (__any_sync(SP_FULL, invalid)) return false;
    if (eqx && n_tok != n_core) return false;
    const int n_ops = n_lead + n_tok + n_trail;
    if (n_ops > ops_cap) return false;  // (cannot happen with the plan's bound; the serial walker flags it)
    // a zero-length clip would end the serial walk (ptCigarIt_next returns its len)
    if ((n_lead > 0 && ((cigar[0] >> 4) == 0 || (n_lead == 2 && (cigar[1] >> 4) == 0))) ||
        (n_trail > 0 && ((cigar[n_cigar - 1] >> 4) == 0 || (n_trail == 2 && (cigar[n_cigar - 2] >> 4) == 0))))
        return false;
    if (lane == 0) {
        int k = 0;
        if (n_lead > 0 && (cigar[0] & 15) == SP_CHARD) ops[k++].oplen = (uint32_t) SP_CHARD | ((uint32_t) lead_h << 4);
        if (k < n_lead) ops[k++].oplen = (uint32_t) SP_CSOFT | ((uint32_t) lead_s << 4);
        k = n_lead + n_tok;
        if (n_trail == 2 || (n_trail == 1 && (cigar[n_cigar - 1] & 15) == SP_CSOFT)) ops[k++].oplen = (uint32_t) SP_CSOFT | ((uint32_t) trail_s << 4);
        if (k < n_ops) ops[k++].oplen = (uint32_t) SP_CHARD | ((uint32_t) trail_h << 4);
    }
    __syncwarp();
    // ---------------------------------------------------------------- 2. one pass over the ops, 32 at a time:
    //   coordinates (prefix sums), the check against the CIGAR, initial markers (ptMarker.c:50-70), confident
    //   blocks (ptMarker.c:328-395)
    const int T = lead_h + trail_h + l_qseq;  // cigar_it.c:41-42
    int c_sq = 0, c_rf = 0;                   // running totals (read bases = SEQ bases + hard clips passed)
    const int lh = (cigar[0] & 15) == SP_CHARD ? lead_h : 0;
    int first_match = 0x7fffffff;
    int run_c = 0, gap_from = lead_s;         // M/I/D CIGARs: runs matched so far, SEQ offset after the last indel
    int n_imk = 0, n_cb = 0, err = 0;
    // A delimiter (clip, or an indel longer than the threshold) closes the block that began at the op after the
    // previous delimiter; (d_sq, d_rf, d_rd) are that op's coordinates (ptMarker.c:332-334 before the first).
    int d_sq = 0, d_rf = pos, d_rd = rev ? T - 1 : 0;
    for (int base = 0; base < n_ops; base += 32) {
        const int k = base + lane;
        const bool on = k < n_ops;
        const uint32_t ol = on ? ops[k].oplen : 0;
        const int op = (int) (ol & 15), len = (int) (ol >> 4);
        const int smk = on ? sp_op_step_mask(op) : 0;
        const int sq = (smk & 1) ? len : 0, rf = (smk & 2) ? len : 0, rd = (smk & 4) ? len : 0;
        const int i_sq = c_sq + sp_warp_incl_scan(sq, lane), i_rf = c_rf + sp_warp_incl_scan(rf, lane);
        const int i_rd = i_sq + lh + ((on && op == SP_CHARD && k >= n_lead) ? len : 0);
        SpOp o;
        o.oplen = ol;
        o.sqs = i_sq - sq;
        o.rfs = pos + i_rf - rf;
        o.rdx = rev ? T - (i_rd - rd) - 1 : i_rd - rd;
        const int nx_rdx = rev ? T - i_rd - 1 : i_rd;  // the next op's
        if (on) {
            ops[k] = o;
            if ((op == SP_CEQUAL || op == SP_CDIFF) && k < first_match) first_match = k;
        }
        const uint32_t lt = (1u << lane) - 1;
        // ---- check against the CIGAR
        const bool is_tok = on && k >= n_lead && k < n_lead + n_tok;
        if (eqx) {
            if (is_tok && ol != core[k - n_lead]) invalid = true;
        } else {
            const bool indel = is_tok && (op == SP_CINS || op == SP_CDEL);
            const uint32_t im = __ballot_sync(SP_FULL, indel);
            const uint32_t before = im & lt;
            const int pi = before ? 31 - __clz(before) : 0;
            const int after_prev = __shfl_sync(SP_FULL, i_sq, pi);
            const int gap = o.sqs - (before ? after_prev : gap_from);  // only M-type tokens in between
            const uint32_t gm = __ballot_sync(SP_FULL, indel && gap > 0);
            if (indel) {
                const int r = run_c + __popc(before) + __popc(gm & (lt | (1u << lane)));
                if (r >= n_core || core[r] != ol) invalid = true;
                else if (gap > 0 && core[r - 1] != ((uint32_t) SP_CMATCH | ((uint32_t) gap << 4))) invalid = true;
            }
            if (im) {
                run_c += __popc(im) + __popc(gm);
                gap_from = __shfl_sync(SP_FULL, i_sq, 31 - __clz(im));
            }
        }
        // ---- initial markers
        {
            int cnt = 0;
            if (on && op == SP_CDIFF) {
                if (o.sqs + len > l_qseq) invalid = true;  // (tokens that overrun the read: not checked yet, stay inside QUAL)
                else
                    for (int j = 0; j < len; j++) cnt += (int) qual[o.sqs + j] >= min_q;
            }
            int incl, total;
            if (!__any_sync(SP_FULL, cnt > 1)) {  // (one base per '*' token is the rule)
                const uint32_t bm = __ballot_sync(SP_FULL, cnt == 1);
                incl = __popc(bm & (lt | (1u << lane)));
                total = __popc(bm);
            } else {
                incl = sp_warp_incl_scan(cnt, lane);
                total = __shfl_sync(SP_FULL, incl, 31);
            }
            if (cnt) {
                int w = n_imk + incl - cnt;
                for (int j = 0; j < len; j++) {
                    const int q = qual[o.sqs + j];
                    if (q < min_q) continue;
                    if (w < imk_cap) {
                        SpInitMarker mk;
                        mk.read_pos_f = rev ? o.rdx - j : o.rdx + j;
                        mk.base_idx = o.sqs + j;
                        mk.ref_pos = o.rfs + j;
                        mk.q = q;
                        imk[w] = mk;
                    } else {
                        err |= SP_GERR_MARKER_CAP;
                    }
                    w++;
                }
            }
            n_imk += total;
        }
        // ---- confident blocks
        {
            const bool delim = on && (op == SP_CSOFT || op == SP_CHARD || ((op == SP_CINS || op == SP_CDEL) && len > indel_threshold));
            const uint32_t dm = __ballot_sync(SP_FULL, delim);
            if (dm) {  // (warp-uniform)
                const uint32_t before = dm & lt;
                const int pd = before ? 31 - __clz(before) : 0;
                const int p_sq = __shfl_sync(SP_FULL, i_sq, pd), p_rf = __shfl_sync(SP_FULL, pos + i_rf, pd),
                          p_rd = __shfl_sync(SP_FULL, nx_rdx, pd);
                const int s_sq = before ? p_sq : d_sq, s_rf = before ? p_rf : d_rf, s_rd = before ? p_rd : d_rd;
                const bool emit = delim && s_sq < o.sqs && s_rf < o.rfs;
                const uint32_t emk = __ballot_sync(SP_FULL, emit);
                if (emit) {
                    SpBlock b;
                    b.rfs = s_rf; b.rfe = o.rfs - 1; b.sqs = s_sq; b.sqe = o.sqs - 1;
                    if (rev) { b.rds_f = o.rdx + 1; b.rde_f = s_rd; }   // it.rde_f + 1 .. conf_rd
                    else { b.rds_f = s_rd; b.rde_f = o.rdx - 1; }        // conf_rd .. it.rds_f - 1
                    const int w = n_cb + __popc(emk & lt);
                    if (w < cb_cap) cb[w] = b; else err |= SP_GERR_BLOCK_CAP;
                }
                n_cb += __popc(emk);
                const int ld = 31 - __clz(dm);
                d_sq = __shfl_sync(SP_FULL, i_sq, ld);
                d_rf = __shfl_sync(SP_FULL, pos + i_rf, ld);
                d_rd = __shfl_sync(SP_FULL, nx_rdx, ld);
            }
        }
        c_sq = __shfl_sync(SP_FULL, i_sq, 31);
        c_rf = __shfl_sync(SP_FULL, i_rf, 31);
    }
    const int c_rd = c_sq + lh + trail_h;
    if (!eqx) {  // the M run after the last indel, and the run count
        const int gap = (c_sq - trail_s) - gap_from;
        if (gap > 0) {
            if (run_c + 1 != n_core || core[run_c] != ((uint32_t) SP_CMATCH | ((uint32_t) gap << 4))) invalid = true;
        } else if (run_c != n_core) {
            invalid = true;
        }
    }
    // (a declined alignment has written only into its own tables, which the serial walker now rewrites)
    if (__any_sync(SP_FULL, invalid)) return false;
    first_match = __reduce_min_sync(SP_FULL, first_match);
    if (lane == 0) {  // sentinel closes the table
        SpOp o;
        o.oplen = SP_CSENTINEL;
        o.sqs = c_sq;
        o.rfs = pos + c_rf;
        o.rdx = rev ? T - c_rd - 1 : c_rd;
        ops[n_ops] = o;
    }
    __syncwarp();
    if (lane == 0) {  // the last block (ptMarker.c:380-392) and the scalars
        const SpOp lastop = ops[n_ops - 1], sen = ops[n_ops];
        const int s_sq = d_sq, s_rf = d_rf, s_rd = d_rd;
        const int it_sqe = sen.sqs - 1, it_rfe = sen.rfs - 1;
        const int last_rds = rev ? sen.rdx + 1 : lastop.rdx, last_rde = rev ? lastop.rdx : sen.rdx - 1;
        if (s_sq <= it_sqe) {
            SpBlock b;
            b.rfs = s_rf; b.rfe = it_rfe; b.sqs = s_sq; b.sqe = it_sqe;
            if (rev) { b.rds_f = last_rds; b.rde_f = s_rd; }
            else { b.rds_f = s_rd; b.rde_f = last_rde; }
            if (n_cb < cb_cap) cb[n_cb] = b; else err |= SP_GERR_BLOCK_CAP;
            n_cb++;
        }
        // extents (ptAlignment.c:52-93)
        int a_rfs = -1, a_rfe = -1, a_rds = -1, a_rde = -1;
        if (first_match < n_ops) {
            const SpOp f = ops[first_match];
            a_rfs = f.rfs;
            if (rev) a_rde = f.rdx; else a_rds = f.rdx;
            if (n_trail > 0) {  // the first clip after the aligned part
                const SpOp c = ops[n_lead + n_tok];
                a_rfe = c.rfs - 1;
                if (rev) a_rds = c.rdx + 1; else a_rde = c.rdx - 1;
            }
        }
        const int lop = (int) (lastop.oplen & 15);
        if (a_rfe == -1 && sp_op_is_match(lop)) {
            a_rfe = it_rfe;
            if (rev) a_rds = last_rds; else a_rde = last_rde;
        }
        info->n_ops = n_ops;
        info->rfs = a_rfs; info->rfe = a_rfe; info->rds_f = a_rds; info->rde_f = a_rde;
        info->lclip_h = (cigar[0] & 15) == SP_CHARD ? (int) (cigar[0] >> 4) : 0;
        info->rclip_h = (cigar[n_cigar - 1] & 15) == SP_CHARD ? (int) (cigar[n_cigar - 1] >> 4) : 0;
        info->pad0 = info->pad1 = 0;
    }
    err = __reduce_or_sync(SP_FULL, err);
    n_cb = __shfl_sync(SP_FULL, n_cb, 0);
    if (lane == 0) {
        info->n_imk = n_imk <= imk_cap ? n_imk : imk_cap;
        info->n_cb = n_cb <= cb_cap ? n_cb : cb_cap;
        info->err = err;
    }
    return true;
}

#endif  // __CUDACC__ || SP_WARP_EMU
