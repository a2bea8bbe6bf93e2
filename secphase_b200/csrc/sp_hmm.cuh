// sp_hmm.cuh -- stage K4: banded glocal forward-backward HMM (BAQ), one HMM instance per lane.
//
// Replaces htslib 1.17 probaln_glocal() as called by calc_local_baq (ptMarker.c:754-757) with a
// uniform iqual == set_q, plus the state/q consumption of ptMarker.c:772-780 restricted to the
// query rows that hold a marker (the only rows the default CLI ever reads, ptMarker.c:826-830).
// Arithmetic specification: SURVEY.md section 8(a) A10; CPU statement: oracle/probaln_port.c.
//
// Bit-exactness: every double operation below is an explicit round-to-nearest IEEE add / mul /
// div in the order the reference evaluates it (no FMA contraction: __dmul_rn/__dadd_rn are never
// fused), the row sum and the D-state recurrence run as serial chains along the band, and the
// row scaling is "store unscaled, multiply by 1/s on the next read", which performs the same
// single rounding as the reference's in-place rescale.  The only algebraic folds are the exact
// ones (EI*m1, EI*m4 hoisted; EI = 2^-2).
//
// Parallel layout: a warp runs 32 independent instances in lock step (instances are sorted by
// band class and length, so trip counts agree); each lane owns a circular band buffer of W cells
// x {M,I,D} doubles in shared memory, interleaved by lane (cell stride = 32 doubles) so that any
// per-lane cell index is bank-conflict free.  Serial chains cost latency, not throughput: the
// M/I updates of neighbouring cells and the other resident warps fill the FP64 pipe.
#pragma once
#include "sp_common.h"

#if defined(__CUDA_ARCH__)
#define SP_DMUL(a, b) __dmul_rn((a), (b))
#define SP_DADD(a, b) __dadd_rn((a), (b))
#define SP_DDIV(a, b) __ddiv_rn((a), (b))
#define SP_FDIV(a, b) __fdiv_rn((a), (b))
#else  // host simulation build: compiled with -ffp-contract=off on x86-64 (SSE2 doubles)
#define SP_DMUL(a, b) ((a) * (b))
#define SP_DADD(a, b) ((a) + (b))
#define SP_DDIV(a, b) ((a) / (b))
#define SP_FDIV(a, b) ((a) / (b))
#endif

#define SP_HMM_EI .25
#define SP_HMM_EM .33333333333

// per-lane view of the band buffer: cell c, state s  ->  row[(c*3+s)*STRIDE]
template <int STRIDE>
struct SpBand {
    double *row;     // lane offset already applied
    uint32_t *code;  // ref code of the column held in each cell, same interleave
    int W;
    SP_HD double ld(int c, int s) const { return row[(c * 3 + s) * STRIDE]; }
    SP_HD void st(int c, int s, double v) const { row[(c * 3 + s) * STRIDE] = v; }
    SP_HD uint32_t ldc(int c) const { return code[c * STRIDE]; }
    SP_HD void stc(int c, uint32_t v) const { code[c * STRIDE] = v; }
    SP_HD int inc(int c) const { return c + 1 == W ? 0 : c + 1; }
    SP_HD int dec(int c) const { return c == 0 ? W - 1 : c - 1; }
};

struct SpHmmIn {
    const uint8_t *ref;    // l_ref codes 0..4
    const uint8_t *qbytes; // byte-per-base query codes, or NULL
    const uint8_t *qseq4;  // 4-bit packed stored SEQ (BAM), used when qbytes == NULL
    int64_t q0;            // first query base (index into qbytes, or base index into qseq4)
    int l_ref, l_query, par_bw;
};

SP_HD int sp_query_code(const SpHmmIn &in, int i0) {
    if (in.qbytes) return in.qbytes[in.q0 + i0];
    const int64_t b = in.q0 + i0;
    const int nib = (in.qseq4[b >> 1] >> ((~b & 1) << 2)) & 0xf;  // bam_seqi
    // seq_nt16_int: 1->0 (A) 2->1 (C) 4->2 (G) 8->3 (T) everything else -> 4
    return nib == 1 ? 0 : nib == 2 ? 1 : nib == 4 ? 2 : nib == 8 ? 3 : 4;
}

SP_HD double sp_emis(const SpConst &C, int rc, int qc) {
    return (rc > 3 || qc > 3) ? 1. : (rc == qc ? C.em_match : C.em_mis);
}

// q = (int)(-4.343*log(t)+.499) with t = 1-max, k>100 -> 99, via the host-built threshold
// table (glibc log is only evaluated on the host; see SpConst::qthr).
SP_HD int sp_q_from_t(const SpConst &C, double t) {
    if (!(t > 0.)) return 0;  // log(0) = -inf -> (int)(+inf) is INT_MIN on x86-64 -> uint8 0; NaN likewise
    int lo = 0, hi = 101;     // count n in 1..101 with t <= qthr[n]; qthr is decreasing in n
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (t <= C.qthr[mid]) lo = mid; else hi = mid - 1;
    }
    return lo > 100 ? 99 : lo;
}

// One instance.  s_arr[(i)*SSTRIDE], i = 0..l_query+1 : scaling factors.
// fsave + r*fs_stride : scaled forward M,I of marker row r, [c*2+{0,1}], c = k - beg(i).
template <int STRIDE, int SSTRIDE>
SP_HD void sp_hmm_instance(const SpConst &C, const SpHmmIn &in, SpBand<STRIDE> B, double *s_arr, double *fsave,
                           int64_t fs_stride, SpRow *rows, int n_rows) {
    const int Lr = in.l_ref, Lq = in.l_query;
    int bw = Lr > Lq ? Lr : Lq;
    if (bw > in.par_bw) bw = in.par_bw;
    {
        int d = Lr - Lq;
        if (d < 0) d = -d;
        if (bw < d) bw = d;
    }
    // transition matrix; float-typed sub-expressions were folded on the host
    const double sM = SP_DDIV(1., (double) (2 * Lq + 2));
    const double sI = sM;
    const double oms = SP_DADD(1., -sM);
    const double m0 = SP_DMUL(C.m0f, oms);
    const double m1 = SP_DMUL(C.d_d, oms);
    const double m2 = m1;
    const double m3 = SP_DMUL(C.ome_f, oms);
    const double m4 = SP_DMUL(C.e_d, oms);
    const double m6 = C.ome_f;
    const double m8 = C.e_d;
    const double bM = (double) SP_FDIV(C.omd_ff, (float) Lr);
    const double bI = (double) SP_FDIV(C.d_f, (float) Lr);
    const double eim1 = SP_DMUL(SP_HMM_EI, m1), eim4 = SP_DMUL(SP_HMM_EI, m4);

    // ------------------------------------------------------------------ forward
    s_arr[0] = 1.;
    int nr = 0;  // next marker row (rows are ascending in t)
    {            // row 1: k = 1..min(Lr, bw+1)
        const int end = Lr < bw + 1 ? Lr : bw + 1;
        const int qc = sp_query_code(in, 0);
        double sum = 0.;
        int c = 1 % B.W;
        for (int k = 1; k <= end; k++) {
            const int rc = in.ref[k - 1];
            B.stc(c, (uint32_t) rc);
            const double M = SP_DMUL(sp_emis(C, rc, qc), bM);
            const double I = SP_DMUL(SP_HMM_EI, bI);
            B.st(c, 0, M);
            B.st(c, 1, I);
            B.st(c, 2, 0.);
            sum = SP_DADD(sum, SP_DADD(M, I));
            c = B.inc(c);
        }
        s_arr[1 * SSTRIDE] = sum;
        if (nr < n_rows && rows[nr].t == 0) {  // only the stand-alone API asks for row 1 (true division there)
            double *fs = fsave + (int64_t) nr * fs_stride;
            int cc = 1 % B.W;
            for (int k = 1; k <= end; k++) {
                fs[(k - 1) * 2 + 0] = SP_DDIV(B.ld(cc, 0), sum);
                fs[(k - 1) * 2 + 1] = SP_DDIV(B.ld(cc, 1), sum);
                cc = B.inc(cc);
            }
            nr++;
        }
    }
    double s_prev = s_arr[1 * SSTRIDE];
    int cbeg = 1 % B.W;  // cell of column beg of the previous row
    int beg_prev = 1;
    for (int i = 2; i <= Lq; i++) {
        const int beg = i - bw > 1 ? i - bw : 1;
        const int end = i + bw < Lr ? i + bw : Lr;
        const int end_prev = (i - 1) + bw < Lr ? (i - 1) + bw : Lr;
        const int qc = sp_query_code(in, i - 1);
        const bool first = (i == 2);  // row 1 was rescaled by a true division
        const double r = first ? 0. : SP_DDIV(1., s_prev);
        // cell of column beg (beg advances by at most one per row)
        int c = (beg == beg_prev) ? cbeg : B.inc(cbeg);
        double pM, pI, pD;  // scaled row i-1 at column k-1
        if (beg > 1) {
            const int cm = (beg == beg_prev) ? B.dec(cbeg) : cbeg;  // column beg-1 is in the band of row i-1
            const double a = B.ld(cm, 0), b = B.ld(cm, 1), d = B.ld(cm, 2);
            if (first) { pM = SP_DDIV(a, s_prev); pI = SP_DDIV(b, s_prev); pD = SP_DDIV(d, s_prev); }
            else { pM = SP_DMUL(a, r); pI = SP_DMUL(b, r); pD = SP_DMUL(d, r); }
        } else {
            pM = pI = pD = 0.;  // f[i-1][0] == 0 for i-1 >= 1
        }
        cbeg = c;
        beg_prev = beg;
        double cM = 0., cD = 0., sum = 0.;
        for (int k = beg; k <= end; k++) {
            double qM, qI, qD;
            int rc;
            if (k <= end_prev) {
                const double a = B.ld(c, 0), b = B.ld(c, 1), d = B.ld(c, 2);
                if (first) { qM = SP_DDIV(a, s_prev); qI = SP_DDIV(b, s_prev); qD = SP_DDIV(d, s_prev); }
                else { qM = SP_DMUL(a, r); qI = SP_DMUL(b, r); qD = SP_DMUL(d, r); }
                rc = (int) B.ldc(c);
            } else {  // column entering the band: row i-1 holds zeros there
                qM = qI = qD = 0.;
                rc = in.ref[k - 1];
                B.stc(c, (uint32_t) rc);
            }
            const double e = sp_emis(C, rc, qc);
            const double M = SP_DMUL(e, SP_DADD(SP_DADD(SP_DMUL(m0, pM), SP_DMUL(m3, pI)), SP_DMUL(m6, pD)));
            const double I = SP_DMUL(SP_HMM_EI, SP_DADD(SP_DMUL(m1, qM), SP_DMUL(m4, qI)));
            const double D = SP_DADD(SP_DMUL(m2, cM), SP_DMUL(m8, cD));
            sum = SP_DADD(sum, SP_DADD(SP_DADD(M, I), D));
            B.st(c, 0, M);
            B.st(c, 1, I);
            B.st(c, 2, D);
            pM = qM; pI = qI; pD = qD;
            cM = M; cD = D;
            c = B.inc(c);
        }
        s_arr[(int64_t) i * SSTRIDE] = sum;
        s_prev = sum;
        if (nr < n_rows && rows[nr].t + 1 == i) {  // marker row: keep the scaled forward M,I
            const double ri = SP_DDIV(1., sum);
            double *fs = fsave + (int64_t) nr * fs_stride;
            int cc = cbeg;
            for (int k = beg; k <= end; k++) {
                fs[(k - beg) * 2 + 0] = SP_DMUL(B.ld(cc, 0), ri);
                fs[(k - beg) * 2 + 1] = SP_DMUL(B.ld(cc, 1), ri);
                cc = B.inc(cc);
            }
            nr++;
        }
    }
    // s[Lq+1] = sum_k ( M'[Lq,k]*sM + I'[Lq,k]*sI ) over the band of row Lq, k ascending
    {
        const int i = Lq;
        const int beg = i - bw > 1 ? i - bw : 1;
        const int end = i + bw < Lr ? i + bw : Lr;
        double sum = 0.;
        int c = cbeg;
        if (Lq == 1) {
            for (int k = beg; k <= end; k++) {
                const double a = SP_DDIV(B.ld(c, 0), s_prev), b = SP_DDIV(B.ld(c, 1), s_prev);
                sum = SP_DADD(sum, SP_DADD(SP_DMUL(a, sM), SP_DMUL(b, sI)));
                c = B.inc(c);
            }
        } else {
            const double r = SP_DDIV(1., s_prev);
            for (int k = beg; k <= end; k++) {
                const double a = SP_DMUL(B.ld(c, 0), r), b = SP_DMUL(B.ld(c, 1), r);
                sum = SP_DADD(sum, SP_DADD(SP_DMUL(a, sM), SP_DMUL(b, sI)));
                c = B.inc(c);
            }
        }
        s_arr[(int64_t) (Lq + 1) * SSTRIDE] = sum;
    }
    if (n_rows == 0) return;

    // ------------------------------------------------------------------ backward (+ MAP at marker rows)
    const int i_stop = rows[0].t + 1;  // nothing below the lowest marker row is consumed
    const double sLq = s_prev, sLq1 = s_arr[(int64_t) (Lq + 1) * SSTRIDE];
    // row Lq: constant inside the band, already in its final scale
    {
        const int beg = Lq - bw > 1 ? Lq - bw : 1;
        const int end = Lq + bw < Lr ? Lq + bw : Lr;
        const double vM = SP_DDIV(SP_DDIV(sM, sLq), sLq1);
        const double vI = SP_DDIV(SP_DDIV(sI, sLq), sLq1);
        int c = cbeg;  // cell of column beg of row Lq; codes of the band are still in place
        for (int k = beg; k <= end; k++) {
            B.st(c, 0, vM);
            B.st(c, 1, vI);
            c = B.inc(c);
        }
    }
    // MAP of one row (the state/q the reference reads at ptMarker.c:778-779)
    auto map_row = [&](int ri, int beg, int end, int cb, double y, bool scale) {
        const double *fs = fsave + (int64_t) ri * fs_stride;
        double sum = 0., mx = 0.;
        int max_k = -1;
        int cc = cb;
        for (int k = beg; k <= end; k++) {
            double bm = B.ld(cc, 0), bi = B.ld(cc, 1);
            if (scale) { bm = SP_DMUL(bm, y); bi = SP_DMUL(bi, y); }
            double z = SP_DMUL(fs[(k - beg) * 2 + 0], bm);
            if (z > mx) { mx = z; max_k = (k - 1) << 2 | 0; }
            sum = SP_DADD(sum, z);
            z = SP_DMUL(fs[(k - beg) * 2 + 1], bi);
            if (z > mx) { mx = z; max_k = (k - 1) << 2 | 1; }
            sum = SP_DADD(sum, z);
            cc = B.inc(cc);
        }
        mx = SP_DDIV(mx, sum);
        rows[ri].state = max_k;
        rows[ri].pmax = mx;
        rows[ri].q = sp_q_from_t(C, SP_DADD(1., -mx));
    };
    nr = n_rows - 1;
    if (rows[nr].t + 1 == Lq) {  // stand-alone API only (pipeline rows satisfy t <= Lq-12)
        const int beg = Lq - bw > 1 ? Lq - bw : 1;
        const int end = Lq + bw < Lr ? Lq + bw : Lr;
        map_row(nr, beg, end, cbeg, 1., false);
        nr--;
        if (nr < 0) return;
    }
    int beg_next = Lq - bw > 1 ? Lq - bw : 1;  // beg of row i+1
    int cbeg_next = cbeg;
    for (int i = Lq - 1; i >= i_stop; i--) {
        const int beg = i - bw > 1 ? i - bw : 1;
        const int end = i + bw < Lr ? i + bw : Lr;
        const int end_next = (i + 1) + bw < Lr ? (i + 1) + bw : Lr;
        const int qc = sp_query_code(in, i);  // query[i] (0-based) == base of row i+1
        const double r1 = (i + 1 == Lq) ? 1. : SP_DDIV(1., s_arr[(int64_t) (i + 1) * SSTRIDE]);
        // cell of column beg of this row
        const int cb = (beg == beg_next) ? cbeg_next : B.dec(cbeg_next);
        if (beg != beg_next) {  // column beg enters the band from the left
            B.stc(cb, (uint32_t) in.ref[beg - 1]);
        }
        // start at column end, walk down
        int c = cb;
        for (int k = beg; k < end; k++) c = B.inc(c);  // cell of column end
        double nM;   // scaled b'M[i+1][k+1]
        int rc_up;   // ref code of column k+1  (ref[k] 0-based)
        {
            const int cn = B.inc(c);
            if (end + 1 <= end_next) {
                nM = SP_DMUL(B.ld(cn, 0), r1);
                rc_up = (int) B.ldc(cn);
            } else {
                nM = 0.;
                rc_up = 4;  // k >= l_ref: emission forced to 0 below
            }
        }
        double cD = 0.;
        const bool ygt1 = i > 1;
        for (int k = end; k >= beg; k--) {
            double qM, qI;
            if (k >= beg_next) {
                qM = SP_DMUL(B.ld(c, 0), r1);
                qI = SP_DMUL(B.ld(c, 1), r1);
            } else {
                qM = qI = 0.;
            }
            const double em = (k >= Lr) ? 0. : sp_emis(C, rc_up, qc);
            const double e = SP_DMUL(em, nM);
            const double bMv = SP_DADD(SP_DADD(SP_DMUL(e, m0), SP_DMUL(eim1, qI)), SP_DMUL(m2, cD));
            const double bIv = SP_DADD(SP_DMUL(e, m3), SP_DMUL(eim4, qI));
            double bDv = SP_DADD(SP_DMUL(e, m6), SP_DMUL(m8, cD));
            bDv = ygt1 ? bDv : SP_DMUL(bDv, 0.);
            rc_up = (int) B.ldc(c);
            B.st(c, 0, bMv);
            B.st(c, 1, bIv);
            nM = qM;
            cD = bDv;
            c = B.dec(c);
        }
        beg_next = beg;
        cbeg_next = cb;
        if (nr >= 0 && rows[nr].t + 1 == i) {
            map_row(nr, beg, end, cb, SP_DDIV(1., s_arr[(int64_t) i * SSTRIDE]), true);
            nr--;
        }
    }
}
