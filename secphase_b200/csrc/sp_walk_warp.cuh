// sp_walk_warp.cuh -- stage K1, warp-cooperative: one warp walks one alignment's CIGAR + cs tag.
//
// Same outputs as sp_walk_alignment (sp_walk.cuh), which stays the specification and the fallback: the refined op
// table, alignment extents, initial markers and confident blocks of
//   ptCigarIt_next / _next_cs          (cigar_it.c:213-308, 145-211)
//   ptAlignment_init_coordinates       (ptAlignment.c:42-95)
//   ptMarker_get_initial_markers       (ptMarker.c:42-75)
//   find_confident_blocks              (ptMarker.c:328-395).
// The serial walker spends ~10^3 instructions per cs token with a quarter of its lanes active (32 different
// alignments per warp, every lane in a different token); here the 32 lanes of a warp look at 128 consecutive
// bytes of ONE alignment's cs text:
//   1. every lane classifies four bytes; tokens are found from the codes (a token starts at ':', '+', '-' or at a
//      '*' that does not continue a run of "*xy" groups), ':' numbers are evaluated by a short scan of (x10, +digit)
//      pairs across lanes and serially within a lane, each token end writes its op;
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

// classes of cs text bytes
enum { K_OTHER = 0, K_DIGIT = 1, K_LOWER = 2, K_COLON = 3, K_STAR = 4, K_PLUS = 5, K_MINUS = 6 };
SP_WD uint32_t sp_cs_code(uint32_t c) {  // (selects, not branches)
    uint32_t k = (c - '0') < 10u ? K_DIGIT : K_OTHER;
    k = (c - 'a') < 26u ? K_LOWER : k;
    k = c == ':' ? K_COLON : k;
    k = c == '*' ? K_STAR : k;
    k = c == '+' ? K_PLUS : k;
    k = c == '-' ? K_MINUS : k;
    return k;
}

// true: tables written; false: run the serial walker instead.  lut: sp_cs_code of every byte value (shared memory
// in the kernel), or null.
SP_WDN bool sp_walk_alignment_warp(int indel_threshold, int min_q, int flag, int pos, int l_qseq, int n_cigar,
                                       const uint32_t *__restrict__ cigar, const uint8_t *__restrict__ tag, int64_t tag_beg,
                                       int64_t tag_end, int tag_kind, const uint8_t *__restrict__ qual, SpOp *ops, int ops_cap,
                                       SpInitMarker *imk, int imk_cap, SpBlock *cb, int cb_cap, SpAlnInfo *info,
                                       const uint8_t *lut = nullptr) {
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
    // A lane owns FOUR consecutive text bytes (one aligned word), an iteration of the warp covers 128: the bytes are
    // classified once (a 256-entry table in shared memory), a lane sees the codes of its own four bytes, of the three
    // before and of the one after in one 32-bit window, walks its four bytes serially (token tracking, number
    // value) and meets the other lanes only for what crosses lane borders -- the token and the number a lane begins
    // in, the rank of its token ends -- so the shuffles, ballots and carries of an iteration are shared by 128 bytes.
    // Carried across iterations (warp-uniform): the codes of the last lane, the token and the value of the number
    // the next iteration begins in.
    const int n_text = (int) (tag_end - tag_beg);
    const int64_t word0 = tag_beg & ~(int64_t) 3;           // first aligned word that holds text
    const int p_first = (int) (word0 - tag_beg);           // text offset of its first byte (-3 .. 0)
    const int n_iter = (n_text + 1 - p_first + 127) >> 7;  // (one virtual terminator byte)
    auto classify = [&](uint32_t w, int p0) -> uint32_t {  // 4-bit codes of the four bytes of word w at text offset p0
        uint32_t pk = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t c = (w >> (8 * j)) & 255u;
            const uint32_t k = lut ? lut[c] : sp_cs_code(c);
            pk |= ((p0 + j >= 0 && p0 + j < n_text) ? k : (uint32_t) K_OTHER) << (4 * j);
        }
        return pk;
    };
    auto load_word = [&](int p0) -> uint32_t {  // (words that hold no text are not read)
        return (p0 + 3 >= 0 && p0 < n_text) ? *reinterpret_cast<const uint32_t *>(tag + tag_beg + p0) : 0u;
    };
    uint32_t carry_pk = 0;  // codes of the last lane of the iteration before
    int cur_kind = 0;       // code of the start character of the token the iteration begins in
    int cur_start = 0;      // its text offset
    uint32_t num_in = 0;    // value of the number the iteration begins in
    int n_tok = 0;
    bool invalid = false;
    uint32_t w_nx = load_word(p_first + 4 * lane);  // the words ahead are loaded and classified one iteration early
    uint32_t pk_nx = classify(w_nx, p_first + 4 * lane);
    const uint32_t lt = (1u << lane) - 1;
    for (int it = 0; it < n_iter; it++) {
        const int p0 = p_first + (it << 7) + 4 * lane;  // text offset of my first byte
        const uint32_t w = w_nx, pk = pk_nx;
        w_nx = load_word(p0 + 128);
        pk_nx = classify(w_nx, p0 + 128);
        uint32_t prev = __shfl_up_sync(SP_FULL, pk, 1), next = __shfl_down_sync(SP_FULL, pk, 1);
        const uint32_t next31 = __shfl_sync(SP_FULL, pk_nx, 0);
        if (lane == 0) prev = carry_pk;
        if (lane == 31) next = next31;
        // window of codes: nibbles 0..2 = the three bytes before mine, 3..6 = mine, 7 = the byte after
        const uint32_t win = ((prev >> 4) & 0xfffu) | (pk << 12) | ((next & 15u) << 28);
        auto nib = [&](int i) -> uint32_t { return (win >> (4 * i)) & 15u; };
        auto starts = [&](int i) -> bool {  // does the byte at window index i (3..7) start a token
            const uint32_t k0 = nib(i);
            return k0 == K_COLON || k0 >= K_PLUS ||
                   (k0 == K_STAR && !(nib(i - 3) == K_STAR && nib(i - 2) == K_LOWER && nib(i - 1) == K_LOWER));
        };
        bool st[5];
#pragma unroll
        for (int j = 0; j < 5; j++) st[j] = starts(3 + j);
        bool en[4];
        int n_end = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int p = p0 + j;
            en[j] = p >= 0 && p < n_text && (p + 1 == n_text || st[j + 1]);
            n_end += en[j];
            // the cs grammar  (:[0-9]+ | \*[a-z][a-z] | [+-][a-z]+)*  checked looking backwards (the byte at n_text is
            // a virtual terminator)
            if (p >= 0 && p <= n_text) {
                const uint32_t k0 = nib(3 + j), k1 = nib(2 + j), k2 = nib(1 + j), k3 = nib(j);
                const bool isd = k0 == K_DIGIT, isl = k0 == K_LOWER;
                bool bad = false;
                if (p < n_text) {
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
        }
        // the token my first byte lies in: the last start in the lanes below, else the iteration's carry
        int my_last = -1;  // my own last start
#pragma unroll
        for (int j = 0; j < 4; j++) my_last = st[j] ? j : my_last;
        const uint32_t hs = __ballot_sync(SP_FULL, my_last >= 0);
        const uint32_t my_tok = my_last >= 0 ? (nib(3 + my_last) | ((uint32_t) my_last << 4)) : 0u;
        const uint32_t below = hs & lt;
        const int src = below ? 31 - __clz(below) : 0;
        const uint32_t from = __shfl_sync(SP_FULL, my_tok, src);
        int tk_kind = below ? (int) (from & 15u) : cur_kind;
        int tk_start = below ? p_first + (it << 7) + 4 * src + (int) (from >> 4) : cur_start;
        // the number my first byte lies in: my four bytes as (m, a) with value' = value * m + a, composed over the
        // four lanes below (16 bytes; numbers longer than 8 digits are not a read's and are declined)
        uint32_t m = 1, a = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const bool isd = nib(3 + j) == K_DIGIT;
            const uint32_t d = ((w >> (8 * j)) & 255u) - '0';
            a = isd ? a * 10u + d : 0u;
            m = isd ? m * 10u : 0u;
        }
        uint32_t im = m, ia = a;  // inclusive over lanes l-3 .. l
#pragma unroll
        for (int o = 1; o < 4; o <<= 1) {
            const uint32_t pm = __shfl_up_sync(SP_FULL, im, o), pa = __shfl_up_sync(SP_FULL, ia, o);
            if (lane >= o) { ia = pa * im + ia; im = pm * im; }
        }
        uint32_t em_ = __shfl_up_sync(SP_FULL, im, 1), ea = __shfl_up_sync(SP_FULL, ia, 1);
        if (lane == 0) { em_ = 1; ea = 0; }
        uint32_t v = num_in * em_ + ea;  // value of the number at my first byte
        // where my token ends go
        const uint32_t e1 = __ballot_sync(SP_FULL, n_end >= 1), e2 = __ballot_sync(SP_FULL, n_end >= 2);
        const int idx = n_lead + n_tok + __popc(e1 & lt) + __popc(e2 & lt);
        // my four bytes in turn: which token, which number value; a lane sees at most two token ends (a token has at
        // least two bytes -- text where it has fewer fails the grammar above and is declined), kept for the writes below
        int e_p[2] = {0, 0}, e_kind[2] = {0, 0}, e_start[2] = {0, 0}, n_e = 0;
        uint32_t e_v[2] = {0, 0};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t k0 = nib(3 + j);
            if (st[j]) { tk_kind = (int) k0; tk_start = p0 + j; }
            v = k0 == K_DIGIT ? v * 10u + (((w >> (8 * j)) & 255u) - '0') : 0u;
            if (en[j]) {
                const int c = n_e < 1 ? 0 : 1;
                e_p[c] = p0 + j; e_kind[c] = tk_kind; e_start[c] = tk_start; e_v[c] = v;
                n_e++;
            }
        }
#pragma unroll
        for (int c = 0; c < 2; c++) {
            if (c < n_e) {
                int tlen, op;
                if (e_kind[c] == K_COLON) { tlen = (int) e_v[c]; op = SP_CEQUAL; if (e_p[c] - e_start[c] > 8) invalid = true; }
                else if (e_kind[c] == K_STAR) { tlen = (e_p[c] - e_start[c] + 1) / 3; op = SP_CDIFF; }
                else { tlen = e_p[c] - e_start[c]; op = e_kind[c] == K_PLUS ? SP_CINS : SP_CDEL; }
                if (tlen <= 0) invalid = true;  // ":0" ends the serial walk
                if (idx + c < ops_cap) ops[idx + c].oplen = (uint32_t) op | ((uint32_t) tlen << 4);  // coordinates follow in step 2
            }
        }
        if (n_e > 2) invalid = true;
        // carries
        n_tok += __popc(e1) + __popc(e2);
        if (hs) {
            const int ls = 31 - __clz(hs);
            const uint32_t lastt = __shfl_sync(SP_FULL, my_tok, ls);
            cur_kind = (int) (lastt & 15u);
            cur_start = p_first + (it << 7) + 4 * ls + (int) (lastt >> 4);
        }
        num_in = __shfl_sync(SP_FULL, v, 31);
        carry_pk = __shfl_sync(SP_FULL, pk, 31);
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
