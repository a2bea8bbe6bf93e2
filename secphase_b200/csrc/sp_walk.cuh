// sp_walk.cuh -- stage K1: one pass over an alignment's CIGAR + cs/MD tag.
//
// Replaces, for one alignment, the three separate iterator walks the reference performs with
//   ptCigarIt_construct / ptCigarIt_next / _next_cs / _next_md   (cigar_it.c:14-69,213-308,145-211,72-141)
// inside
//   ptAlignment_init_coordinates      (ptAlignment.c:42-95)     -> alignment extents
//   ptMarker_get_initial_markers      (ptMarker.c:42-75)        -> mismatch markers with q >= min_q
//   find_confident_blocks             (ptMarker.c:328-395)      -> indel/clip-free runs
// and additionally materialises the refined op table (SpOp) so that no later stage re-walks or
// re-tokenises anything (the reference re-walks ~10x per alignment, regcomp each time).
#pragma once
#include "sp_common.h"

// Sequential byte reader with a 16-byte register window (global loads are 128-bit, aligned).
struct SpByteReader {
    const uint8_t *base;  // 16-byte aligned pool start
    int64_t wbase;
    uint32_t w0, w1, w2, w3;
    SP_HD void init(const uint8_t *pool) {
        base = pool;
        wbase = -16;
        w0 = w1 = w2 = w3 = 0;
    }
    SP_HD uint8_t at(int64_t i) {
        int64_t a = i & ~(int64_t) 15;
        if (a != wbase) {
#if defined(__CUDA_ARCH__)
            uint4 v = *reinterpret_cast<const uint4 *>(base + a);
            w0 = v.x; w1 = v.y; w2 = v.z; w3 = v.w;
#else
            const uint32_t *p = reinterpret_cast<const uint32_t *>(base + a);
            w0 = p[0]; w1 = p[1]; w2 = p[2]; w3 = p[3];
#endif
            wbase = a;
        }
        int k = (int) (i & 15);
        uint32_t x = (k < 8) ? ((k < 4) ? w0 : w1) : ((k < 12) ? w2 : w3);
        return (uint8_t) (x >> ((k & 3) * 8));
    }
};

SP_HD bool sp_is_lower(uint8_t c) { return c >= 'a' && c <= 'z'; }
SP_HD bool sp_is_upper(uint8_t c) { return c >= 'A' && c <= 'Z'; }
SP_HD bool sp_is_digit(uint8_t c) { return c >= '0' && c <= '9'; }

// The iterator state of cigar_it.h:12-41.
struct SpCigarIt {
    const uint32_t *cigar;
    int n, idx, is_rev;
    int op, len;
    int match_remain;
    int rds_f, rde_f, sqs, sqe, rfs, rfe;
    int kind;  // 0 cs, 1 MD
    SpByteReader tag;
    int64_t tag_pos, tag_end;
    int bad;  // an op the reference has no case for (N, P, B): undefined there, rejected here
};

SP_HD void sp_it_init(SpCigarIt &it, const uint32_t *cigar, int n_cigar, int flag, int pos, int l_qseq,
                      const uint8_t *tag_pool, int64_t tag_beg, int64_t tag_end, int kind) {
    // cigar_it.c:14-69
    it.cigar = cigar;
    it.n = n_cigar;
    it.idx = -1;
    it.is_rev = (flag & SP_FREVERSE) != 0;
    it.op = -1;
    it.len = 0;
    it.match_remain = 0;
    it.sqs = 0;
    it.sqe = -1;
    it.rfs = pos;
    it.rfe = pos - 1;
    int lclip = 0, rclip = 0;
    if (n_cigar > 0) {
        if ((cigar[0] & 15) == SP_CHARD) lclip = (int) (cigar[0] >> 4);
        if ((cigar[n_cigar - 1] & 15) == SP_CHARD) rclip = (int) (cigar[n_cigar - 1] >> 4);
    }
    it.rds_f = it.is_rev ? rclip + lclip + l_qseq : 0;
    it.rde_f = it.is_rev ? rclip + lclip + l_qseq - 1 : -1;
    it.kind = kind;
    it.tag.init(tag_pool);
    it.tag_pos = tag_beg;
    it.tag_end = tag_end;
    it.bad = 0;
}

// ptCigarIt_next_cs, cigar_it.c:145-211.  The reference runs the POSIX regex
//   (:([0-9]+))|(([+-])([a-z]+)|([\*]([a-z]+))+)
// un-anchored from the current offset (leftmost-longest); tokens are recognised the same way.
// No match leaves op/len untouched and returns 0.
SP_HD int sp_it_next_cs(SpCigarIt &it) {
    int64_t i = it.tag_pos, n = it.tag_end;
    while (i < n) {
        uint8_t c = it.tag.at(i);
        if (c == ':') {
            int64_t j = i + 1;
            int v = 0;
            while (j < n) {
                uint8_t d = it.tag.at(j);
                if (!sp_is_digit(d)) break;
                v = v * 10 + (d - '0');
                j++;
            }
            if (j > i + 1) {
                it.op = SP_CEQUAL;
                it.len = v;
                it.tag_pos = j;
                return it.len;
            }
        } else if (c == '*') {
            int64_t j = i;
            while (j < n && it.tag.at(j) == '*') {
                int64_t k = j + 1;
                while (k < n && sp_is_lower(it.tag.at(k))) k++;
                if (k == j + 1) break;
                j = k;
            }
            if (j > i) {
                it.op = SP_CDIFF;
                it.len = (int) ((j - i + 1) / 3);
                it.tag_pos = j;
                return it.len;
            }
        } else if (c == '+' || c == '-') {
            int64_t k = i + 1;
            while (k < n && sp_is_lower(it.tag.at(k))) k++;
            if (k > i + 1) {
                it.op = (c == '+') ? SP_CINS : SP_CDEL;
                it.len = (int) (k - i - 1);
                it.tag_pos = k;
                return it.len;
            }
        }
        i++;
    }
    return 0;
}

// ptCigarIt_next_md, cigar_it.c:72-141, regex (([A-Z])([0][A-Z])*)|([0-9]+)|([\^]([A-Z]+)).
SP_HD int sp_it_next_md(SpCigarIt &it) {
    for (;;) {
        int64_t i = it.tag_pos, n = it.tag_end;
        bool found = false;
        while (i < n) {
            uint8_t c = it.tag.at(i);
            if (sp_is_digit(c)) {
                int64_t j = i;
                int v = 0;
                while (j < n) {
                    uint8_t d = it.tag.at(j);
                    if (!sp_is_digit(d)) break;
                    v = v * 10 + (d - '0');
                    j++;
                }
                if (c == '0') {  // "two consecutive mismatches" separator, cigar_it.c:86-93
                    it.op = SP_CDIFF;
                    it.len = 0;
                } else {
                    it.op = SP_CEQUAL;
                    it.len = v;
                }
                it.tag_pos = j;
                found = true;
                break;
            } else if (sp_is_upper(c)) {
                int64_t j = i + 1;
                while (j + 1 < n && it.tag.at(j) == '0' && sp_is_upper(it.tag.at(j + 1))) j += 2;
                if (c < 90) {  // cigar_it.c:104 ('Z' itself falls through every branch)
                    it.op = SP_CDIFF;
                    it.len = 1 + (int) ((j - i - 1) / 2);
                }
                it.tag_pos = j;
                found = true;
                break;
            } else if (c == '^') {
                int64_t j = i + 1;
                while (j < n && sp_is_upper(it.tag.at(j))) j++;
                if (j > i + 1) {
                    it.op = SP_CDEL;
                    it.len = (int) (j - i - 1);
                    it.tag_pos = j;
                    found = true;
                    break;
                }
            }
            i++;
        }
        if (!found) return 0;   // also the failed recursive call of cigar_it.c:138: len is 0 there
        if (it.len != 0) return it.len;
        // len == 0: "between consecutive mismatches in MD tag there placed an additional zero";
        // the reference recurses once more (cigar_it.c:138) -- same as taking the next token.
    }
}

// ptCigarIt_next, cigar_it.c:213-308.  Returns the iterator's len; 0 ends iteration.
SP_HD int sp_it_next(SpCigarIt &it) {
    if (it.idx == it.n - 1) return 0;
    it.idx += 1;
    int rd_step = 0, sq_step = 0, rf_step = 0;
    int op = (int) (it.cigar[it.idx] & 15);
    int len = (int) (it.cigar[it.idx] >> 4);
    switch (op) {
        case SP_CMATCH:
        case SP_CEQUAL:
        case SP_CDIFF:
            if (it.match_remain == 0) it.match_remain = len;
            if (it.kind == 0) {
                sp_it_next_cs(it);
                it.match_remain -= it.len;
                if (0 < it.match_remain) it.idx -= 1;
            } else {
                if (0 <= it.match_remain) sp_it_next_md(it);
                if (it.match_remain < 0) {
                    it.op = SP_CEQUAL;
                    it.len = sp_min(len, -1 * it.match_remain);
                    it.match_remain += len;
                } else {
                    int md_len = it.len;
                    it.len = sp_min(it.len, it.match_remain);
                    it.match_remain -= md_len;
                }
                if (0 < it.match_remain) it.idx -= 1;
            }
            rd_step = sq_step = rf_step = it.len;
            break;
        case SP_CINS:
            if (it.kind == 0) sp_it_next_cs(it);
            it.len = len;
            it.op = op;
            rd_step = len;
            sq_step = len;
            rf_step = 0;
            break;
        case SP_CDEL:
            if (it.kind == 0) sp_it_next_cs(it);
            else sp_it_next_md(it);
            rd_step = 0;
            sq_step = 0;
            rf_step = len;
            break;
        case SP_CSOFT:
            it.len = len;
            it.op = op;
            rd_step = len;
            sq_step = len;
            rf_step = 0;
            break;
        case SP_CHARD:
            it.len = len;
            it.op = op;
            rd_step = len;
            sq_step = 0;
            rf_step = 0;
            break;
        default:
            it.bad = 1;  // N / P / B: the reference reads uninitialised steps here (Q15)
            return 0;
    }
    if (it.is_rev) {
        it.rde_f = it.rds_f - 1;
        it.rds_f -= rd_step;
    } else {
        it.rds_f = it.rde_f + 1;
        it.rde_f += rd_step;
    }
    it.sqs = it.sqe + 1;
    it.sqe += sq_step;
    it.rfs = it.rfe + 1;
    it.rfe += rf_step;
    return it.len;
}

SP_HD bool sp_op_is_match(int op) { return op == SP_CMATCH || op == SP_CEQUAL || op == SP_CDIFF; }

// per-alignment scalar results of the walk
struct SpAlnInfo {
    int32_t n_ops;
    int32_t rfs, rfe, rds_f, rde_f;  // ptAlignment extents
    int32_t lclip_h, rclip_h;        // hard clips in CIGAR order (ptMarker.c:82-89)
    int32_t n_imk;                   // initial markers
    int32_t n_cb;                    // confident blocks
    int32_t err;
    int32_t pad0, pad1;
};

SP_HD void sp_walk_alignment(int indel_threshold, int min_q, int flag, int pos, int l_qseq, int n_cigar,
                             const uint32_t *cigar, const uint8_t *tag_pool, int64_t tag_beg, int64_t tag_end,
                             int tag_kind, const uint8_t *qual, SpOp *ops, int ops_cap, SpInitMarker *imk,
                             int imk_cap, SpBlock *cb, int cb_cap, SpAlnInfo *info) {
    SpCigarIt it;
    sp_it_init(it, cigar, n_cigar, flag, pos, l_qseq, tag_pool, tag_beg, tag_end, tag_kind);
    const bool rev = it.is_rev != 0;
    int err = 0;
    int n_ops = 0, n_imk = 0, n_cb = 0;
    // ptAlignment_init_coordinates state
    int a_rfs = -1, a_rfe = -1, a_rds = -1, a_rde = -1;
    // find_confident_blocks state (ptMarker.c:332-334)
    int conf_sqs = 0, conf_rfs = pos, conf_rd = rev ? it.rde_f : it.rds_f;

    while (sp_it_next(it)) {
        // ---- op table
        if (n_ops < ops_cap) {
            SpOp o;
            o.oplen = (uint32_t) it.op | ((uint32_t) it.len << 4);
            o.sqs = it.sqs;
            o.rfs = it.rfs;
            o.rdx = rev ? it.rde_f : it.rds_f;
            ops[n_ops] = o;
        } else {
            err |= SP_GERR_OP_CAP;
        }
        n_ops++;
        const bool is_m = sp_op_is_match(it.op);
        // ---- extents, ptAlignment.c:52-81
        if (a_rfs == -1 && is_m) {
            a_rfs = it.rfs;
            if (rev) a_rde = it.rde_f; else a_rds = it.rds_f;
        }
        if (a_rfe == -1 && a_rfs != -1 && (it.op == SP_CHARD || it.op == SP_CSOFT)) {
            a_rfe = it.rfe;
            if (rev) a_rds = it.rde_f + 1; else a_rde = it.rds_f - 1;
        }
        // ---- initial markers, ptMarker.c:50-70
        if (it.op == SP_CDIFF) {
            for (int j = 0; j < it.len; j++) {
                int q = qual[it.sqs + j];
                if (q < min_q) continue;
                if (n_imk < imk_cap) {
                    SpInitMarker m;
                    m.read_pos_f = rev ? it.rde_f - j : it.rds_f + j;
                    m.base_idx = it.sqs + j;
                    m.ref_pos = it.rfs + j;
                    m.q = q;
                    imk[n_imk] = m;
                } else {
                    err |= SP_GERR_MARKER_CAP;
                }
                n_imk++;
            }
        }
        // ---- confident blocks, ptMarker.c:337-377
        const bool is_indel = (it.op == SP_CINS || it.op == SP_CDEL);
        const bool is_clip = (it.op == SP_CSOFT || it.op == SP_CHARD);
        if ((is_indel && it.len > indel_threshold) || is_clip) {
            if (conf_sqs < it.sqs && conf_rfs < it.rfs) {
                SpBlock b;
                b.rfs = conf_rfs;
                b.rfe = it.rfs - 1;
                b.sqs = conf_sqs;
                b.sqe = it.sqs - 1;
                if (rev) {
                    b.rds_f = it.rde_f + 1;
                    b.rde_f = conf_rd;
                } else {
                    b.rds_f = conf_rd;
                    b.rde_f = it.rds_f - 1;
                }
                if (n_cb < cb_cap) cb[n_cb] = b; else err |= SP_GERR_BLOCK_CAP;
                n_cb++;
            }
            conf_sqs = it.sqe + 1;
            conf_rfs = it.rfe + 1;
            conf_rd = rev ? it.rds_f - 1 : it.rde_f + 1;
        }
    }
    if (it.bad) err |= SP_GERR_BADOP;
    // extents when the alignment ends with mis/matches, ptAlignment.c:83-93
    if (a_rfe == -1 && sp_op_is_match(it.op)) {
        a_rfe = it.rfe;
        if (rev) a_rds = it.rds_f; else a_rde = it.rde_f;
    }
    // last confident block, ptMarker.c:380-392
    if (conf_sqs <= it.sqe) {
        SpBlock b;
        b.rfs = conf_rfs;
        b.rfe = it.rfe;
        b.sqs = conf_sqs;
        b.sqe = it.sqe;
        if (rev) {
            b.rds_f = it.rds_f;
            b.rde_f = conf_rd;
        } else {
            b.rds_f = conf_rd;
            b.rde_f = it.rde_f;
        }
        if (n_cb < cb_cap) cb[n_cb] = b; else err |= SP_GERR_BLOCK_CAP;
        n_cb++;
    }
    // sentinel closes the op table (ends of the last op)
    {
        SpOp o;
        o.oplen = SP_CSENTINEL;
        o.sqs = it.sqe + 1;
        o.rfs = it.rfe + 1;
        o.rdx = rev ? it.rds_f - 1 : it.rde_f + 1;
        ops[n_ops <= ops_cap ? n_ops : ops_cap] = o;  // the table has ops_cap+1 slots
    }
    info->n_ops = n_ops <= ops_cap ? n_ops : ops_cap;
    info->rfs = a_rfs;
    info->rfe = a_rfe;
    info->rds_f = a_rds;
    info->rde_f = a_rde;
    info->lclip_h = (n_cigar > 0 && (cigar[0] & 15) == SP_CHARD) ? (int) (cigar[0] >> 4) : 0;
    info->rclip_h = (n_cigar > 0 && (cigar[n_cigar - 1] & 15) == SP_CHARD) ? (int) (cigar[n_cigar - 1] >> 4) : 0;
    info->n_imk = n_imk <= imk_cap ? n_imk : imk_cap;
    info->n_cb = n_cb <= cb_cap ? n_cb : cb_cap;
    info->err = err;
    info->pad0 = info->pad1 = 0;
}

// ---- op table accessors (see SpOp) -------------------------------------------------
struct SpOpView {
    int op, len, sqs, sqe, rfs, rfe, rds_f, rde_f;
};
SP_HD SpOpView sp_op_view(const SpOp *ops, int j, bool rev) {
    SpOpView v;
    const SpOp a = ops[j], b = ops[j + 1];
    v.op = (int) (a.oplen & 15);
    v.len = (int) (a.oplen >> 4);
    v.sqs = a.sqs;
    v.sqe = b.sqs - 1;
    v.rfs = a.rfs;
    v.rfe = b.rfs - 1;
    if (rev) {
        v.rde_f = a.rdx;
        v.rds_f = b.rdx + 1;
    } else {
        v.rds_f = a.rdx;
        v.rde_f = b.rdx - 1;
    }
    return v;
}

// index of the op whose read-forward interval contains p, or -1
SP_HD int sp_find_op_by_read_pos(const SpOp *ops, int n_ops, bool rev, int p) {
    int lo = 0, hi = n_ops;  // largest j with rdx<=p (fwd) / rdx>=p (rev)
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        int x = ops[mid].rdx;
        bool ok = rev ? (x >= p) : (x <= p);
        if (ok) lo = mid + 1; else hi = mid;
    }
    int j = lo - 1;
    if (j < 0) return -1;
    SpOpView v = sp_op_view(ops, j, rev);
    if (v.rds_f <= p && p <= v.rde_f) return j;
    return -1;
}

