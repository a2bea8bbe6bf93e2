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
//      "*xy" groups), ':' numbers are evaluated by a segmented scan, token ends are compacted to dense lanes;
//   2. while they stream by, the tokens are checked against the CIGAR: between its end clips the CIGAR must be
//      exactly the run structure of the tokens (an M op = a maximal run of ':'/'*' tokens of the same total
//      length, an I/D op = a '+'/'-' token of the same length; with =/X CIGARs every token is one op).  For such
//      input -- everything an aligner writes -- the refined ops ARE the end clips plus the tokens;
//   3. coordinates are prefix sums over the op table; markers, extents and confident blocks follow from it with
//      ballots and scans.
// Anything else (MD tags, text outside the cs grammar, tokens that disagree with the CIGAR, N/P ops, clips in odd
// places, ':0') makes the function return false before it has written anything that matters, and lane 0 runs the
// serial walker, whose behaviour on such input is the reference's.
#pragma once
#include "sp_walk.cuh"

#if defined(__CUDACC__)

#define SP_FULL 0xffffffffu

__device__ __forceinline__ int sp_warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(SP_FULL, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// One M/=/X/I/D/S/H op's steps: bit0 advances the stored SEQ, bit1 the reference, bit2 the read (cigar_it.c:224-291)
__device__ __forceinline__ int sp_op_step_mask(int op) {
    return (op == SP_CMATCH || op == SP_CEQUAL || op == SP_CDIFF) ? 7 : op == SP_CINS ? 5 : op == SP_CDEL ? 2
           : op == SP_CSOFT ? 5 : op == SP_CHARD ? 4 : 0;
}

// true: tables written; false: run the serial walker instead.
__device__ bool sp_walk_alignment_warp(int indel_threshold, int min_q, int flag, int pos, int l_qseq, int n_cigar,
                                       const uint32_t *__restrict__ cigar, const uint8_t *__restrict__ tag, int64_t tag_beg,
                                       int64_t tag_end, int tag_kind, const uint8_t *__restrict__ qual, SpOp *ops, int ops_cap,
                                       SpInitMarker *imk, int imk_cap, SpBlock *cb, int cb_cap, SpAlnInfo *info) {
    const int lane = threadIdx.x & 31;
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
    // ---------------------------------------------------------------- 1. tokens (+ check against the CIGAR)
    // carried across chunks (warp-uniform)
    uint32_t pb1 = 0, pb2 = 0, pb3 = 0;      // the three bytes before the chunk
    int cur_kind = 0;                        // start character of the token the chunk begins in
    int64_t cur_start = tag_beg;
    uint32_t num_in = 0;                     // value of the number the chunk begins in
    int n_tok = 0;
    int run_idx = -1, run_len = 0, prev_tok_class = 3;  // class: 0 M-type, 1 I, 2 D, 3 none yet
    int prev_tok_kind = 0, prev_tok_len = 0;            // (=/X CIGARs: the last token itself is the run)
    bool invalid = false;
    const int n_chunks = (int) ((tag_end - tag_beg + 1 + 31) >> 5);  // one virtual terminator byte
    for (int ch = 0; ch < n_chunks; ch++) {
        const int64_t p = tag_beg + ((int64_t) ch << 5) + lane;
        const bool in_range = p < tag_end;
        const uint32_t c0 = in_range ? tag[p] : 0;
        uint32_t b1 = __shfl_up_sync(SP_FULL, c0, 1), b2 = __shfl_up_sync(SP_FULL, c0, 2), b3 = __shfl_up_sync(SP_FULL, c0, 3);
        if (lane < 1) b1 = pb1;
        if (lane < 2) b2 = lane == 0 ? pb2 : pb1;
        if (lane < 3) b3 = lane == 0 ? pb3 : lane == 1 ? pb2 : pb1;
        uint32_t nx = __shfl_down_sync(SP_FULL, c0, 1);
        if (lane == 31) nx = p + 1 < tag_end ? tag[p + 1] : 0;
        auto dig = [](uint32_t c) { return c >= '0' && c <= '9'; };
        auto low = [](uint32_t c) { return c >= 'a' && c <= 'z'; };
        auto sgn = [](uint32_t c) { return c == '+' || c == '-'; };
        const bool isd = dig(c0), isl = low(c0), iss = c0 == ':' || c0 == '*' || sgn(c0);
        const bool start = in_range && (c0 == ':' || sgn(c0) || (c0 == '*' && !(b3 == '*' && low(b2) && low(b1))));
        const bool start_nx = p + 1 < tag_end && (nx == ':' || sgn(nx) || (nx == '*' && !(b2 == '*' && low(b1) && low(c0))));
        const bool end = in_range && (p + 1 == tag_end || start_nx);
        // the cs grammar  (:[0-9]+ | \*[a-z][a-z] | [+-][a-z]+)*  checked looking backwards (the byte at tag_end is a
        // virtual terminator)
        if (p <= tag_end) {
            bool bad = false;
            if (in_range) {
                if (!(isd || isl || iss)) bad = true;
                if (p == tag_beg && !iss) bad = true;
                if (isd && !(b1 == ':' || dig(b1))) bad = true;
                if (isl && !(b1 == '*' || sgn(b1) || low(b1))) bad = true;
                if (isl && low(b1) && low(b2) && b3 == '*') bad = true;  // third letter of a '*' group
            }
            if (!isd && b1 == ':') bad = true;                      // ':' without a number
            if (!isl && (b1 == '*' || sgn(b1))) bad = true;         // sign or '*' without a letter
            if (!isl && low(b1) && b2 == '*') bad = true;           // '*' group with a single letter
            if (bad) invalid = true;
        }
        // which token does my byte belong to
        const uint32_t sm = __ballot_sync(SP_FULL, start);
        const uint32_t below = sm & (lane == 31 ? 0xffffffffu : ((2u << lane) - 1));
        int tk_kind = cur_kind;
        int64_t tk_start = cur_start;
        {
            const int src = below ? 31 - __clz(below) : 0;
            const uint32_t ck = __shfl_sync(SP_FULL, c0, src);
            if (below) { tk_kind = (int) ck; tk_start = tag_beg + ((int64_t) ch << 5) + src; }
        }
        // numbers: affine scan (m, a): value = num_in * m + a
        uint32_t m = isd ? 10u : 0u, a = isd ? c0 - '0' : 0u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t pm = __shfl_up_sync(SP_FULL, m, o), pa = __shfl_up_sync(SP_FULL, a, o);
            if (lane >= o) { a = pa * m + a; m = pm * m; }
        }
        const uint32_t num = num_in * m + a;
        // token at its end byte
        int tlen = 0, tclass = 3;
        if (end) {
            if (tk_kind == ':') { tlen = (int) num; tclass = 0; if (p - tk_start > 9) invalid = true; }
            else if (tk_kind == '*') { tlen = (int) ((p - tk_start + 1) / 3); tclass = 0; }
            else { tlen = (int) (p - tk_start); tclass = tk_kind == '+' ? 1 : 2; }
            if (tlen <= 0) invalid = true;  // ":0" ends the serial walk
        }
        // dense lanes: token r of this chunk
        const uint32_t em = __ballot_sync(SP_FULL, end);
        const int nt = __popc(em);
        const int srcl = lane < nt ? (int) __fns(em, 0, lane + 1) : 0;
        const int d_len = __shfl_sync(SP_FULL, tlen, srcl), d_class = __shfl_sync(SP_FULL, tclass, srcl);
        const int d_kind = __shfl_sync(SP_FULL, tk_kind, srcl);
        const bool d_valid = lane < nt;
        // runs: a token starts a new run if it is an indel or follows one (=/X CIGARs: every token is a run)
        int pcl = __shfl_up_sync(SP_FULL, d_class, 1), pkd = __shfl_up_sync(SP_FULL, d_kind, 1), pln = __shfl_up_sync(SP_FULL, d_len, 1);
        if (lane == 0) { pcl = prev_tok_class; pkd = prev_tok_kind; pln = prev_tok_len; }
        const bool run_start = d_valid && (eqx || d_class != 0 || pcl != 0);
        const uint32_t rs = __ballot_sync(SP_FULL, run_start);
        const int my_run = run_idx + __popc(rs & (lane == 31 ? 0xffffffffu : ((2u << lane) - 1)));
        // length of the M run a token closes: segmented sum of the M tokens' lengths
        int seg = (d_valid && d_class == 0) ? d_len : 0;
        {
            bool head = run_start;  // a segment head stops the propagation
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int ps = __shfl_up_sync(SP_FULL, seg, o);
                const bool ph = __shfl_up_sync(SP_FULL, head, o);
                if (lane >= o && !head) { seg += ps; head = ph; }
            }
            if (!head && d_valid) seg += run_len;  // the run began in an earlier chunk
        }
        int pseg = __shfl_up_sync(SP_FULL, seg, 1);
        if (lane == 0) pseg = run_len;
        // at a run start the run before it is complete: compare it with its CIGAR op
        if (run_start && my_run >= 1) {
            const int k = my_run - 1;
            if (k >= n_core) invalid = true;
            else {
                const uint32_t c = core[k];
                const int op = (int) (c & 15), len = (int) (c >> 4);
                if (eqx) {
                    const int want = pkd == ':' ? SP_CEQUAL : pkd == '*' ? SP_CDIFF : pkd == '+' ? SP_CINS : SP_CDEL;
                    if (op != want || len != pln) invalid = true;
                } else if (pcl == 0) {
                    if (op != SP_CMATCH || len != pseg) invalid = true;
                } else {
                    if (op != (pcl == 1 ? SP_CINS : SP_CDEL) || len != pln) invalid = true;
                }
            }
        }
        // temporary op records: kind and length (coordinates follow in step 2)
        if (d_valid) {
            const int idx = n_lead + n_tok + lane;
            if (idx < ops_cap) {
                const int op = d_kind == ':' ? SP_CEQUAL : d_kind == '*' ? SP_CDIFF : d_kind == '+' ? SP_CINS : SP_CDEL;
                ops[idx].oplen = (uint32_t) op | ((uint32_t) d_len << 4);
            }
        }
        // carries
        if (nt > 0) {
            const int last = nt - 1;
            run_idx = __shfl_sync(SP_FULL, my_run, last);
            run_len = __shfl_sync(SP_FULL, seg, last);
            prev_tok_class = __shfl_sync(SP_FULL, d_class, last);
            prev_tok_kind = __shfl_sync(SP_FULL, d_kind, last);
            prev_tok_len = __shfl_sync(SP_FULL, d_len, last);
            n_tok += nt;
        }
        if (sm) {
            const int src = 31 - __clz(sm);
            cur_kind = (int) __shfl_sync(SP_FULL, c0, src);
            cur_start = tag_beg + ((int64_t) ch << 5) + src;
        }
        num_in = __shfl_sync(SP_FULL, num, 31);
        pb1 = __shfl_sync(SP_FULL, c0, 31);
        pb2 = __shfl_sync(SP_FULL, c0, 30);
        pb3 = __shfl_sync(SP_FULL, c0, 29);
        if (__any_sync(SP_FULL, invalid)) return false;
    }
    // the last run, and the run count
    {
        bool bad = run_idx != n_core - 1;
        if (!bad) {
            const uint32_t c = core[run_idx];
            const int op = (int) (c & 15), len = (int) (c >> 4);
            if (eqx) {
                const int want = prev_tok_kind == ':' ? SP_CEQUAL : prev_tok_kind == '*' ? SP_CDIFF : prev_tok_kind == '+' ? SP_CINS : SP_CDEL;
                bad = op != want || len != prev_tok_len;
            } else if (prev_tok_class == 0) {
                bad = op != SP_CMATCH || len != run_len;
            } else {
                bad = op != (prev_tok_class == 1 ? SP_CINS : SP_CDEL) || len != prev_tok_len;
            }
        }
        if (bad) return false;
    }
    const int n_ops = n_lead + n_tok + n_trail;
    if (n_ops > ops_cap) return false;  // (cannot happen with the plan's bound; the serial walker flags it)
    if (lane == 0) {
        int k = 0;
        if (lead_h || (n_lead > 0 && (cigar[0] & 15) == SP_CHARD)) ops[k++].oplen = (uint32_t) SP_CHARD | ((uint32_t) lead_h << 4);
        if (k < n_lead) ops[k++].oplen = (uint32_t) SP_CSOFT | ((uint32_t) lead_s << 4);
        k = n_lead + n_tok;
        if (n_trail == 2 || (n_trail == 1 && (cigar[n_cigar - 1] & 15) == SP_CSOFT)) ops[k++].oplen = (uint32_t) SP_CSOFT | ((uint32_t) trail_s << 4);
        if (k < n_ops) ops[k++].oplen = (uint32_t) SP_CHARD | ((uint32_t) trail_h << 4);
    }
    __syncwarp();
    // a zero-length clip would end the serial walk (ptCigarIt_next returns its len)
    if ((n_lead > 0 && ((cigar[0] >> 4) == 0 || (n_lead == 2 && (cigar[1] >> 4) == 0))) ||
        (n_trail > 0 && ((cigar[n_cigar - 1] >> 4) == 0 || (n_trail == 2 && (cigar[n_cigar - 2] >> 4) == 0))))
        return false;
    // ---------------------------------------------------------------- 2. coordinates: prefix sums over the ops
    const int T = lead_h + trail_h + l_qseq;  // cigar_it.c:41-42
    int c_sq = 0, c_rf = 0, c_rd = 0;         // running totals
    int first_match = 0x7fffffff;
    for (int base = 0; base < n_ops; base += 32) {
        const int k = base + lane;
        const bool on = k < n_ops;
        const uint32_t ol = on ? ops[k].oplen : 0;
        const int op = (int) (ol & 15), len = (int) (ol >> 4);
        const int sm = on ? sp_op_step_mask(op) : 0;
        const int sq = (sm & 1) ? len : 0, rf = (sm & 2) ? len : 0, rd = (sm & 4) ? len : 0;
        const int i_sq = sp_warp_incl_scan(sq, lane), i_rf = sp_warp_incl_scan(rf, lane), i_rd = sp_warp_incl_scan(rd, lane);
        if (on) {
            SpOp o;
            o.oplen = ol;
            o.sqs = c_sq + i_sq - sq;
            o.rfs = pos + c_rf + i_rf - rf;
            const int erd = c_rd + i_rd - rd;  // read bases before this op
            o.rdx = rev ? T - erd - 1 : erd;
            ops[k] = o;
            if ((op == SP_CEQUAL || op == SP_CDIFF) && k < first_match) first_match = k;
        }
        c_sq += __shfl_sync(SP_FULL, i_sq, 31);
        c_rf += __shfl_sync(SP_FULL, i_rf, 31);
        c_rd += __shfl_sync(SP_FULL, i_rd, 31);
    }
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
    // ---------------------------------------------------------------- 3. initial markers (ptMarker.c:50-70)
    int n_imk = 0, err = 0;
    for (int base = 0; base < n_ops; base += 32) {
        const int k = base + lane;
        int cnt = 0;
        SpOp o;
        o.oplen = 0; o.sqs = 0; o.rfs = 0; o.rdx = 0;
        if (k < n_ops) {
            o = ops[k];
            if ((int) (o.oplen & 15) == SP_CDIFF) {
                const int len = (int) (o.oplen >> 4);
                for (int j = 0; j < len; j++) cnt += (int) qual[o.sqs + j] >= min_q;
            }
        }
        const int incl = sp_warp_incl_scan(cnt, lane);
        if (cnt) {
            int w = n_imk + incl - cnt;
            const int len = (int) (o.oplen >> 4);
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
        n_imk += __shfl_sync(SP_FULL, incl, 31);
    }
    // ---------------------------------------------------------------- 4. confident blocks (ptMarker.c:328-395)
    // A delimiter (clip, or an indel longer than the threshold) closes the block that began at the op after the
    // previous delimiter; the op after a delimiter is where the next block starts.
    int n_cb = 0;
    int last_delim = -1;  // (warp-uniform) index of the last delimiter seen so far
    for (int base = 0; base < n_ops; base += 32) {
        const int k = base + lane;
        bool delim = false;
        SpOp o;
        o.oplen = 0; o.sqs = 0; o.rfs = 0; o.rdx = 0;
        if (k < n_ops) {
            o = ops[k];
            const int op = (int) (o.oplen & 15), len = (int) (o.oplen >> 4);
            delim = op == SP_CSOFT || op == SP_CHARD || ((op == SP_CINS || op == SP_CDEL) && len > indel_threshold);
        }
        const uint32_t dm = __ballot_sync(SP_FULL, delim);
        bool emit = false;
        SpBlock b;
        if (delim) {
            const uint32_t before = dm & ((1u << lane) - 1);
            const int pd = before ? base + 31 - __clz(before) : last_delim;
            int s_sq = 0, s_rf = pos, s_rd = rev ? T - 1 : 0;  // ptMarker.c:332-334
            if (pd >= 0) {
                const SpOp nx = ops[pd + 1];
                s_sq = nx.sqs; s_rf = nx.rfs; s_rd = nx.rdx;
            }
            if (s_sq < o.sqs && s_rf < o.rfs) {
                emit = true;
                b.rfs = s_rf; b.rfe = o.rfs - 1; b.sqs = s_sq; b.sqe = o.sqs - 1;
                if (rev) { b.rds_f = o.rdx + 1; b.rde_f = s_rd; }   // it.rde_f + 1 .. conf_rd
                else { b.rds_f = s_rd; b.rde_f = o.rdx - 1; }        // conf_rd .. it.rds_f - 1
            }
        }
        const uint32_t emk = __ballot_sync(SP_FULL, emit);
        if (emit) {
            const int w = n_cb + __popc(emk & ((1u << lane) - 1));
            if (w < cb_cap) cb[w] = b; else err |= SP_GERR_BLOCK_CAP;
        }
        n_cb += __popc(emk);
        if (dm) last_delim = base + 31 - __clz(dm);
    }
    if (lane == 0) {  // the last block (ptMarker.c:380-392) and the scalars
        const SpOp lastop = ops[n_ops - 1], sen = ops[n_ops];
        int s_sq = 0, s_rf = pos, s_rd = rev ? T - 1 : 0;
        if (last_delim >= 0) {
            const SpOp nx = ops[last_delim + 1];
            s_sq = nx.sqs; s_rf = nx.rfs; s_rd = nx.rdx;
        }
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

#endif  // __CUDACC__
