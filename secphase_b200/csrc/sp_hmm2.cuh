// sp_hmm2.cuh -- stage K4, main kernel body: banded glocal forward-backward HMM (BAQ) for band
// half-widths up to SP_H2_MAXBW, one HMM instance per lane, band in shared memory.
//
// Same contract and the same arithmetic as sp_hmm.cuh (htslib 1.17 probaln_glocal as called at
// ptMarker.c:754-757): every double operation is an explicit round-to-nearest
// IEEE add / mul / div, in the reference's order, never contracted to an FMA; the row sum and the
// D-state recurrence stay serial chains along the band; rows are stored unscaled and multiplied by
// 1/s on the next read (one rounding, as the reference's in-place rescale).  What changes is only
// how the work is laid out for the SM:
//
//  * band-offset addressing: cell o of the buffer holds column beg(i)+o of the row written last,
//    so neighbours are reached through compile-time immediates from one per-row base (the band
//    origin shift sh = beg(i)-beg(i-1) is folded into that base); a permanent zero cell at o = -1
//    and never-written cells above the band stand for the reference's zero padding;
//  * (M,I) of a cell are one 16-byte word (LDS.128/STS.128), D a separate 8-byte plane, both
//    lane-interleaved, hence bank-conflict free for any per-lane offset;
//  * cells are processed U at a time: the "parallel part" of chunk c+1 (rescale, M, I, m2*M, M+I:
//    independent across cells) is issued next to the two serial chains of chunk c (D recurrence,
//    row sum), so the FP64 pipe always has independent work while a chain waits on its latency;
//  * the emission term is picked from two 64-bit masks per row (match / N), built from three
//    bit-planes of the reference codes that slide with the band: no per-cell code loads;
//  * forward stores 1/s[i] (it needs it anyway); backward never divides.
// Exact algebraic folds used: EI*m1 and EI*m4 hoisted (EI = 2^-2; identical unless an intermediate
// is below 2^-1020, far under anything that reaches a posterior), y=(i>1) folded into m6,m8 for
// row 1 ((x)*0 == e*0 + 0*bD == +0 for finite x).
#pragma once
#include "sp_blocks.cuh"
#include "sp_common.h"
#include "sp_hmm.cuh"

#define SP_H2_U 4
// Unrolled row bodies: number of low band cells whose forward D / backward bI live in registers
// (the remaining cells keep them in shared memory).  Bounded by the 255-register file: a single
// spill is ruinous here because local memory competes for the ~28 KB of L1 left beside 227 KB of
// shared memory.
#ifndef SP_H2_NCRF
#define SP_H2_NCRF 64  // forward: D of cells < NCRF in registers (clamped to NC)
#endif
#ifndef SP_H2_NCRF_WIDE
#define SP_H2_NCRF_WIDE 24  // the same for the unrolled bodies of bands wider than 45 cells (register budget)
#endif
#ifndef SP_H2_NCRB
#define SP_H2_NCRB -1  // backward: bI of cells < NCRB in registers; -1 = no unrolled backward body at all
#endif
#define SP_H2_MAXBW 94  // 2*bw+1 <= 189 cells: the per-row emission masks are NW <= 3 64-bit words
SP_HD int sp_h2_words(int bw) { return (2 * bw + 1 + 63) >> 6; }

// NW x 64 bits, bit o <-> band cell o.  All indexing is by unrolled compare-and-select so that the
// words stay in registers.
template <int NW>
struct SpBits {
    uint64_t w[NW];
    SP_HD void clear() {
#pragma unroll
        for (int k = 0; k < NW; k++) w[k] = 0;
    }
    SP_HD void shr1() {
#pragma unroll
        for (int k = 0; k < NW; k++) w[k] = (w[k] >> 1) | (k + 1 < NW ? w[k + 1 < NW ? k + 1 : k] << 63 : 0);
    }
    SP_HD void shl1_in(uint64_t bit) {
#pragma unroll
        for (int k = NW - 1; k >= 0; k--) w[k] = (w[k] << 1) | (k ? w[k ? k - 1 : 0] >> 63 : bit);
    }
    SP_HD void or_bit(int pos, uint64_t bit) {
#pragma unroll
        for (int k = 0; k < NW; k++)
            if ((pos >> 6) == k) w[k] |= bit << (pos & 63);
    }
    SP_HD uint32_t from(int o) const {  // bits o, o+1, ... of the word holding o (o multiple of 4 => 4 bits valid)
        uint64_t x = w[0];
#pragma unroll
        for (int k = 1; k < NW; k++)
            if ((o >> 6) == k) x = w[k];
        return (uint32_t) (x >> (o & 63));
    }
    SP_HD bool any_below(int n) const {
        uint64_t acc = 0;
#pragma unroll
        for (int k = 0; k < NW; k++) {
            const int m = n - 64 * k;
            acc |= m <= 0 ? 0 : (m >= 64 ? w[k] : (w[k] & (((uint64_t) 1 << m) - 1)));
        }
        return acc != 0;
    }
};

// match / N masks of one row from the three code bit-planes and the row's query base
template <int NW>
SP_HD void sp_h2_row_masks(const SpBits<NW> &p0, const SpBits<NW> &p1, const SpBits<NW> &p2, int qc, SpBits<NW> &mm,
                           SpBits<NW> &nn) {
#pragma unroll
    for (int k = 0; k < NW; k++) {
        if (qc > 3) { mm.w[k] = 0; nn.w[k] = ~(uint64_t) 0; }
        else {
            mm.w[k] = ((qc & 1) ? p0.w[k] : ~p0.w[k]) & ((qc & 2) ? p1.w[k] : ~p1.w[k]) & ~p2.w[k];
            nn.w[k] = p2.w[k];
        }
    }
}

// Warp-uniform decisions of the unrolled fast path (every lane of a full warp is inside
// sp_hmm2_instance together); the host simulation runs one lane at a time.
#if defined(__CUDA_ARCH__)
#define SP_WARP_ANY(p) __any_sync(0xffffffffu, (p))
#define SP_WARP_ALL(p) __all_sync(0xffffffffu, (p))
#define SP_WARP_MIN(x) __reduce_min_sync(0xffffffffu, (x))
#define SP_WARP_MAX(x) __reduce_max_sync(0xffffffffu, (x))
// scheduling fence inside the fully unrolled row bodies: stops ptxas hoisting shared-memory loads
// of far-away cells (it otherwise runs out of the 255 registers and spills)
#define SP_SCHED_FENCE() asm volatile("" ::: "memory")
#else
#define SP_WARP_ANY(p) (p)
#define SP_WARP_ALL(p) (p)
#define SP_WARP_MIN(x) (x)
#define SP_WARP_MAX(x) (x)
#define SP_SCHED_FENCE() ((void) 0)
#endif

#if defined(__CUDACC__)
typedef double2 SpD2;
#else
struct alignas(16) SpD2 {
    double x, y;
};
#endif

// per-lane view: cell o (-1 <= o <= 2*bw+1) lives at mi[o*STRIDE], d[o*STRIDE]; all cells zero on entry
template <int STRIDE>
struct SpBand2 {
    SpD2 *mi;
    double *d;
};

SP_HD int sp_h2_cells(int bw) { return 2 * bw + 3; }  // cells -1 .. 2bw+1

// One byte from global memory into a 32-bit register with nothing hanging off the load (the
// compiler otherwise appends a zero-extension right behind it, which makes a load issued a whole
// row ahead of its use stall at once).
SP_HD uint32_t sp_ldg_u8(const uint8_t *p) {
#if defined(__CUDA_ARCH__)
    uint32_t v;
    asm("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
#else
    return *p;
#endif
}
// Query base i0 in two steps, so that the load can be issued a whole row before the decode needs it.
SP_HD uint32_t sp_query_raw(const SpHmmIn &in, int i0) {
    return sp_ldg_u8(in.qbytes ? in.qbytes + in.q0 + i0 : in.qseq4 + ((in.q0 + i0) >> 1));
}
SP_HD int sp_query_decode(const SpHmmIn &in, int i0, uint32_t raw) {
    if (in.qbytes) return (int) raw;
    const uint32_t nib = (raw >> ((~(in.q0 + i0) & 1) << 2)) & 0xf;  // bam_seqi
    // seq_nt16_int = {4,0,1,4,2,4,4,4,3,4,4,4,4,4,4,4}, one hex digit per nibble value
    return (int) ((0x4444444344424104ull >> (nib << 2)) & 0xf);
}

// rinv[i], i = 1..l_query : 1/s[i].   fsave + r*fs_stride + o*fs_cs : scaled forward (M,I) of consumed row r,
// cell o, as one 16-byte pair.  Marker-row mode keeps each instance's rows to itself (fs_cs = 2, rows
// back to back); the --writeBam mode, where every lane saves and re-reads every row in lock step,
// interleaves the 32 lanes of a warp (fs_cs = 64, lane offset folded into fsave) so that each
// (row, cell) is one coalesced 512-byte access.
//
// NC > 0 selects the fully unrolled row bodies for bands of exactly NC = 2*bw+1 cells: for the rows
// where every lane of the warp has a full, sliding band (bw+2 <= i <= Lr-bw) the D plane (forward)
// lives in registers and every cell index is a compile-time constant, which removes a third of
// the shared-memory traffic and all per-chunk address / mask arithmetic.  full_warp says that all
// 32 lanes are inside this function (the fast path takes warp-uniform decisions by vote).
// SOLO: the lane runs on its own (the other lanes of the warp are elsewhere): the unrolled bodies' warp-uniform
// decisions are then simply this lane's (k_hmmf's in-place recomputation of a guard-banded instance).
template <int STRIDE, int NW, int NC, int FS_CS = 2, int NCRF_ = (NC > 45 ? SP_H2_NCRF_WIDE : SP_H2_NCRF),
          int NCRB_ = SP_H2_NCRB, bool SOLO = false>
SP_HD void sp_hmm2_instance(const SpConst &C, const SpHmmIn &in, const SpBand2<STRIDE> B, double *rinv, double *fsave,
                            int64_t fs_stride, SpRow *rows, int n_rows, bool full_warp) {
    constexpr int fs_cs = FS_CS;  // doubles between consecutive cells of a saved forward row
    constexpr int U = SP_H2_U;
    const int Lr = in.l_ref, Lq = in.l_query;
    const int bw = sp_hmm_bw(Lr, Lq, in.par_bw);
    // rows [fi0, fi1] run the unrolled forward body (empty range when NC == 0 or the warp is mixed)
    int fi0 = 1, fi1 = 0;
    bool unrolled_ok = false;
    if constexpr (NC > 0) {
        if constexpr (SOLO) unrolled_ok = 2 * bw + 1 == NC;
        else unrolled_ok = full_warp && SP_WARP_ALL(2 * bw + 1 == NC);
        if (unrolled_ok) {
            fi0 = bw + 2;
            if constexpr (SOLO) fi1 = Lq < Lr - bw ? Lq : Lr - bw;
            else fi1 = SP_WARP_MIN(Lq < Lr - bw ? Lq : Lr - bw);
        }
    }
    constexpr int NCR = NCRF_ < NC ? NCRF_ : NC;               // forward cells with D in registers
    constexpr int NCB = NCRB_ < NC ? (NCRB_ < 0 ? 0 : NCRB_) : NC;  // backward cells with bI in registers
    constexpr bool BWD_UNROLLED = NC > 0 && NCRB_ >= 0;
    double Dr[(NCR > NCB ? NCR : NCB) > 0 ? (NCR > NCB ? NCR : NCB) : 1];  // see the unrolled bodies
    bool d_regs = false;
    (void) Dr; (void) d_regs; (void) fi0; (void) fi1; (void) unrolled_ok;
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
    const double emA = C.em_match, emB = C.em_mis;

    // bit-planes of the reference codes under the band: bit o <-> ref[beg-1+o] (column beg+o)
    SpBits<NW> p0, p1, p2;
    p0.clear(); p1.clear(); p2.clear();
    int nr = 0;  // next marker row (rows are ascending in t)

    // ------------------------------------------------------------------ forward, row 1
    int n_prev = Lr < bw + 1 ? Lr : bw + 1;
    double s_last;
    {
        const int qc = sp_query_code(in, 0);
        double sum = 0.;
        for (int o = 0; o < n_prev; o++) {
            const int rc = in.ref[o];
            p0.or_bit(o, (uint64_t) (rc & 1));
            p1.or_bit(o, (uint64_t) ((rc >> 1) & 1));
            p2.or_bit(o, (uint64_t) ((rc >> 2) & 1));
            SpD2 v;
            v.x = SP_DMUL(sp_emis(C, rc, qc), bM);
            v.y = SP_DMUL(SP_HMM_EI, bI);
            B.mi[o * STRIDE] = v;
            sum = SP_DADD(sum, SP_DADD(v.x, v.y));
        }
        // the reference rescales row 1 by a true division; store it scaled, later rows read it with r = 1
        double *fs = (nr < n_rows && rows[nr].t == 0) ? fsave + (int64_t) nr * fs_stride : nullptr;
        for (int o = 0; o < n_prev; o++) {
            SpD2 v = B.mi[o * STRIDE];
            v.x = SP_DDIV(v.x, sum);
            v.y = SP_DDIV(v.y, sum);
            B.mi[o * STRIDE] = v;
            if (fs) *reinterpret_cast<SpD2 *>(fs + (int64_t) o * fs_cs) = v;
        }
        if (fs) nr++;  // only the stand-alone API asks for row 1
        rinv[1] = SP_DDIV(1., sum);
        s_last = sum;
    }
    // ------------------------------------------------------------------ forward, rows 2..Lq
    // Global reads (query base, entering reference base, next marker row) are issued one row ahead
    // of their use so that their latency hides behind a whole row of arithmetic.
    double r = 1.;
    int beg_prev = 1, end_prev = n_prev;
    int t_next = nr < n_rows ? rows[nr].t : 0x7fffffff;
    uint32_t qraw_next = Lq >= 2 ? sp_query_raw(in, 1) : 0;
    uint32_t rc_next = 2 + bw <= Lr ? sp_ldg_u8(in.ref + 1 + bw) : 0;  // the column entering at row 2, if any
    for (int i = 2; i <= Lq; i++) {
        const int beg = i - bw > 1 ? i - bw : 1;
        const int end = i + bw < Lr ? i + bw : Lr;
        const int n = end - beg + 1;
        const int sh = beg - beg_prev;
        const int qc = sp_query_decode(in, i - 1, qraw_next);
        const uint32_t rc_in = rc_next;
        if (i < Lq) qraw_next = sp_query_raw(in, i);
        if (i + 1 + bw <= Lr) rc_next = sp_ldg_u8(in.ref + i + bw);
        if (sh) { p0.shr1(); p1.shr1(); p2.shr1(); }
        if (end > end_prev) {  // one column enters the band on the right
            p0.or_bit(n - 1, (uint64_t) (rc_in & 1));
            p1.or_bit(n - 1, (uint64_t) ((rc_in >> 1) & 1));
            p2.or_bit(n - 1, (uint64_t) ((rc_in >> 2) & 1));
        }
        SpBits<NW> mm, nn;  // bit o: reference base of column beg+o equals the query base / emission is 1 (an N)
        sp_h2_row_masks(p0, p1, p2, qc, mm, nn);
        // rows whose band holds an N (rare) take the one-cell-at-a-time code, which knows about nn
        const bool has_n = nn.any_below(n);

        double sum = 0.;
        bool fast = false;
        if constexpr (NC > 0) {
            if constexpr (SOLO) fast = i >= fi0 && i <= fi1 && !has_n;
            else fast = i >= fi0 && i <= fi1 && !SP_WARP_ANY(has_n);  // warp-uniform: fi0, fi1 are
            if (fast != d_regs) {                                 // move D between its plane and registers
#pragma unroll
                for (int o = 0; o < NCR; o++) {
                    if (fast) Dr[o] = B.d[o * STRIDE];
                    else B.d[o * STRIDE] = Dr[o];
                }
                d_regs = fast;
            }
        }
        if (NC > 0 && fast) {
            if constexpr (NC > 0) {
                // full sliding band: n == NC, sh == 1; M[i,o] reads old cell o, I[i,o] old cell o+1
                SpD2 a = B.mi[0];
                double pM = SP_DMUL(a.x, r), pI = SP_DMUL(a.y, r), pD = SP_DMUL(NCR > 0 ? Dr[0] : B.d[0], r);
                double Mlast = 0., cD = 0.;
#pragma unroll
                for (int o = 0; o < NC; o++) {
                    double qM = 0., qI = 0., qD = 0.;
                    SpD2 v;
                    if (o + 1 < NC) {
                        a = B.mi[(o + 1 < NC ? o + 1 : 0) * STRIDE];
                        qM = SP_DMUL(a.x, r); qI = SP_DMUL(a.y, r);
                        qD = SP_DMUL(o + 1 < NCR ? Dr[o + 1 < NCR ? o + 1 : 0] : B.d[(o + 1) * STRIDE], r);
                        v.y = SP_DADD(SP_DMUL(eim1, qM), SP_DMUL(eim4, qI));
                    } else {
                        v.y = 0.;  // the column that just entered: row i-1 holds zeros there
                    }
                    const double e = ((mm.w[o >> 6] >> (o & 63)) & 1) ? emA : emB;
                    v.x = SP_DMUL(e, SP_DADD(SP_DADD(SP_DMUL(m0, pM), SP_DMUL(m3, pI)), SP_DMUL(m6, pD)));
                    cD = SP_DADD(SP_DMUL(m2, Mlast), SP_DMUL(m8, cD));
                    sum = SP_DADD(sum, SP_DADD(SP_DADD(v.x, v.y), cD));
                    B.mi[o * STRIDE] = v;
                    if (o < NCR) Dr[o < NCR ? o : 0] = cD;
                    else B.d[o * STRIDE] = cD;
                    Mlast = v.x;
                    pM = qM; pI = qI; pD = qD;
                }
            }
        } else {
        // old row (i-1): M[i,o] reads old cell o-1+sh, I[i,o] reads old cell o+sh
        const SpD2 *omi = B.mi + sh * STRIDE;
        const double *od = B.d + sh * STRIDE;
        double pM, pI, pD;  // scaled old cell o-1+sh (carried between cells)
        {
            const SpD2 a = omi[-STRIDE];
            pM = SP_DMUL(a.x, r); pI = SP_DMUL(a.y, r); pD = SP_DMUL(od[-STRIDE], r);
        }
        double Mlast = 0., cD = 0.;

        // parallel part of U cells starting at o0: everything that does not depend on this row's D chain
        auto fwdP = [&](int o0, double (&t)[U], double (&u)[U]) {
            const uint32_t mb = mm.from(o0);
            double qM[U], qI[U], qD[U];
#pragma unroll
            for (int j = 0; j < U; j++) {
                const SpD2 a = omi[(o0 + j) * STRIDE];
                const double dd = od[(o0 + j) * STRIDE];
                qM[j] = SP_DMUL(a.x, r); qI[j] = SP_DMUL(a.y, r); qD[j] = SP_DMUL(dd, r);
            }
#pragma unroll
            for (int j = 0; j < U; j++) {
                const double aM = j ? qM[j ? j - 1 : 0] : pM, aI = j ? qI[j ? j - 1 : 0] : pI,
                             aD = j ? qD[j ? j - 1 : 0] : pD;
                const double e = ((mb >> j) & 1) ? emA : emB;
                SpD2 v;
                v.x = SP_DMUL(e, SP_DADD(SP_DADD(SP_DMUL(m0, aM), SP_DMUL(m3, aI)), SP_DMUL(m6, aD)));
                v.y = SP_DADD(SP_DMUL(eim1, qM[j]), SP_DMUL(eim4, qI[j]));
                t[j] = SP_DMUL(m2, Mlast);
                u[j] = SP_DADD(v.x, v.y);
                B.mi[(o0 + j) * STRIDE] = v;
                Mlast = v.x;
            }
            pM = qM[U - 1]; pI = qI[U - 1]; pD = qD[U - 1];
        };
        // the two serial chains over those U cells: D[o] = m2*M[o-1] + m8*D[o-1];  sum += (M+I)+D
        auto fwdC = [&](int o0, const double (&t)[U], const double (&u)[U]) {
#pragma unroll
            for (int j = 0; j < U; j++) {
                cD = SP_DADD(t[j], SP_DMUL(m8, cD));
                sum = SP_DADD(sum, SP_DADD(u[j], cD));
                B.d[(o0 + j) * STRIDE] = cD;
            }
        };
        const int nfull = has_n ? 0 : n / U;
        if (nfull > 0) {
            double tA[U], uA[U], tB[U], uB[U];
            fwdP(0, tA, uA);
            int c = 1;
            for (; c + 1 < nfull; c += 2) {
                fwdP(c * U, tB, uB);
                fwdC((c - 1) * U, tA, uA);
                fwdP((c + 1) * U, tA, uA);
                fwdC(c * U, tB, uB);
            }
            if (c < nfull) {
                fwdP(c * U, tB, uB);
                fwdC((c - 1) * U, tA, uA);
                fwdC(c * U, tB, uB);
            } else {
                fwdC((c - 1) * U, tA, uA);
            }
        }
        for (int o = nfull * U; o < n; o++) {  // remainder cells, one at a time
            const SpD2 a = omi[o * STRIDE];
            const double qM = SP_DMUL(a.x, r), qI = SP_DMUL(a.y, r), qD = SP_DMUL(od[o * STRIDE], r);
            const double e = (nn.from(o) & 1) ? 1. : ((mm.from(o) & 1) ? emA : emB);
            SpD2 v;
            v.x = SP_DMUL(e, SP_DADD(SP_DADD(SP_DMUL(m0, pM), SP_DMUL(m3, pI)), SP_DMUL(m6, pD)));
            v.y = SP_DADD(SP_DMUL(eim1, qM), SP_DMUL(eim4, qI));
            cD = SP_DADD(SP_DMUL(m2, Mlast), SP_DMUL(m8, cD));
            sum = SP_DADD(sum, SP_DADD(SP_DADD(v.x, v.y), cD));
            B.mi[o * STRIDE] = v;
            B.d[o * STRIDE] = cD;
            Mlast = v.x;
            pM = qM; pI = qI; pD = qD;
        }
        }  // generic row body
        const double ri = SP_DDIV(1., sum);
        rinv[i] = ri;
        s_last = sum;
        if (t_next + 1 == i) {  // marker row: keep the scaled forward M,I
            double *fs = fsave + (int64_t) nr * fs_stride;
            for (int o = 0; o < n; o++) {
                const SpD2 a = B.mi[o * STRIDE];
                SpD2 v;
                v.x = SP_DMUL(a.x, ri);
                v.y = SP_DMUL(a.y, ri);
                *reinterpret_cast<SpD2 *>(fs + (int64_t) o * fs_cs) = v;
            }
            nr++;
            t_next = nr < n_rows ? rows[nr].t : 0x7fffffff;
        }
        r = ri;
        beg_prev = beg;
        end_prev = end;
        n_prev = n;
    }
    // s[Lq+1] = sum_k ( M'[Lq,k]*sM + I'[Lq,k]*sI ) over the band of row Lq, k ascending
    double sLq1 = 0.;
    for (int o = 0; o < n_prev; o++) {
        const SpD2 a = B.mi[o * STRIDE];
        sLq1 = SP_DADD(sLq1, SP_DADD(SP_DMUL(SP_DMUL(a.x, r), sM), SP_DMUL(SP_DMUL(a.y, r), sI)));
    }
    // ------------------------------------------------------------------ backward (+ MAP at marker rows)
    // (no early return below: the lanes of a full warp vote together further down)
    int i_stop = n_rows > 0 ? rows[0].t + 1 : Lq;  // nothing below the lowest marker row is consumed
    if (n_rows > 0) {  // row Lq: constant inside the band, already in its final scale
        SpD2 v;
        v.x = SP_DDIV(SP_DDIV(sM, s_last), sLq1);
        v.y = SP_DDIV(SP_DDIV(sI, s_last), sLq1);
        for (int o = 0; o < n_prev; o++) B.mi[o * STRIDE] = v;
    }
    // MAP of one row (the state/q the reference reads at ptMarker.c:778-779)
    auto map_row = [&](int ri, int beg, int n, double y, bool scale) {
        const double *fs = fsave + (int64_t) ri * fs_stride;
        double sum = 0., mx = 0.;
        int max_k = -1;
        // the saved forward row comes from global memory: fetch MAPC cells ahead of the (serial, in
        // the reference's order) max/sum chain so that the loads overlap instead of each stalling it
        constexpr int MAPC = FS_CS == 64 ? 4 : 1;  // (marker-row mode: a handful of rows, keep the registers)
        for (int o0 = 0; o0 < n; o0 += MAPC) {
            SpD2 fv[MAPC];
#pragma unroll
            for (int j = 0; j < MAPC; j++)
                if (o0 + j < n) fv[j] = *reinterpret_cast<const SpD2 *>(fs + (int64_t) (o0 + j) * fs_cs);
#pragma unroll
            for (int j = 0; j < MAPC; j++) {
                const int o = o0 + j;
                if (o < n) {
                    const SpD2 a = B.mi[o * STRIDE];
                    double bm = a.x, bi = a.y;
                    if (scale) { bm = SP_DMUL(bm, y); bi = SP_DMUL(bi, y); }
                    double z = SP_DMUL(fv[j].x, bm);
                    if (z > mx) { mx = z; max_k = (beg + o - 1) << 2 | 0; }
                    sum = SP_DADD(sum, z);
                    z = SP_DMUL(fv[j].y, bi);
                    if (z > mx) { mx = z; max_k = (beg + o - 1) << 2 | 1; }
                    sum = SP_DADD(sum, z);
                }
            }
        }
        mx = SP_DDIV(mx, sum);
        rows[ri].state = max_k;
        rows[ri].pmax = mx;
        rows[ri].q = sp_q_from_t(C, SP_DADD(1., -mx));
    };
    nr = n_rows - 1;
    if (nr >= 0 && rows[nr].t + 1 == Lq) {  // stand-alone API only (pipeline rows satisfy t <= Lq-12)
        map_row(nr, beg_prev, n_prev, 1., false);
        nr--;
    }
    if (nr < 0) i_stop = Lq;  // nothing (left) to do for this lane: no live step below
    // planes re-aligned for the backward sweep: bit o <-> ref[beg+o] (the base of column beg+o+1)
    p0.shr1(); p1.shr1(); p2.shr1();
    int beg_next = beg_prev, end_next = end_prev;  // band of row i+1
    t_next = nr >= 0 ? rows[nr].t : -2;
    qraw_next = Lq >= 2 ? sp_query_raw(in, Lq - 1) : 0;
    {
        const int b = Lq - 1 - bw > 1 ? Lq - 1 - bw : 1;
        rc_next = (b != beg_prev && b < Lr) ? sp_ldg_u8(in.ref + b) : 0;
    }
    double rinv_i = Lq >= 2 ? rinv[Lq - 1] : 1.;  // 1/s[i], fetched one row before it scales row i
    double r1 = 1.;                               // row Lq is stored in its final scale
    // The sweep is counted in steps j (row i = Lq-1-j) so that a full warp stays in lock step even
    // though its lanes start and stop at different rows: a lane that has passed its lowest marker
    // row ("dead") keeps executing the unrolled body on its own slab -- which costs nothing, the
    // warp is busy anyway -- so that the warp-uniform choice of the unrolled body does not end when
    // the first lane finishes.
    int jmax = Lq - 1 - i_stop;
    if constexpr (BWD_UNROLLED) {
        if (unrolled_ok && !SOLO) jmax = SP_WARP_MAX(jmax);
    }
#define Ir Dr  // backward: the same registers hold bI of the last written row while in the unrolled body
    bool i_regs = false;
    (void) i_regs;
    for (int j = 0; j <= jmax; j++) {
        const int i = Lq - 1 - j;
        const bool live = i >= i_stop;
        const int beg = i - bw > 1 ? i - bw : 1;
        const int end = i + bw < Lr ? i + bw : Lr;
        const int n = end - beg + 1;
        const int sh = beg_next - beg;
        int qc = 0;
        uint32_t rc_in = 0;
        double y = 0.;
        if (live) {
            qc = sp_query_decode(in, i, qraw_next);  // query[i] (0-based) == base of row i+1
            rc_in = rc_next;
            y = rinv_i;
            if (i > i_stop) {
                qraw_next = sp_query_raw(in, i - 1);
                const int b = i - 1 - bw > 1 ? i - 1 - bw : 1;
                rc_next = (b != beg && b < Lr) ? sp_ldg_u8(in.ref + b) : 0;  // (the bit of column Lr+1 is never consumed)
                rinv_i = rinv[i - 1];
            }
        }
        const double m6e = i > 1 ? m6 : 0., m8e = i > 1 ? m8 : 0.;
        if (live && sh) {  // one column enters the band on the left
            p0.shl1_in((uint64_t) (rc_in & 1));
            p1.shl1_in((uint64_t) ((rc_in >> 1) & 1));
            p2.shl1_in((uint64_t) ((rc_in >> 2) & 1));
        }
        SpBits<NW> mm, nn;
        sp_h2_row_masks(p0, p1, p2, qc, mm, nn);
        const bool has_n = nn.any_below(n);
        bool fast = false;
        if constexpr (BWD_UNROLLED) {
            if (unrolled_ok) {
                // full sliding band with an in-range top cell: i >= bw+1 and i+bw < Lr
                if constexpr (SOLO) fast = !live || (i >= bw + 1 && i + bw < Lr && !has_n);
                else fast = SP_WARP_ALL(!live || (i >= bw + 1 && i + bw < Lr && !has_n));
                if (fast != i_regs) {  // unrolled body, cells < NCB: bM in the dense 8-byte plane, bI in registers
#pragma unroll
                    for (int o = 0; o < NCB; o++) {
                        if (fast) {
                            const SpD2 a = B.mi[o * STRIDE];
                            B.d[o * STRIDE] = a.x;
                            Ir[o] = a.y;
                        } else {
                            SpD2 a;
                            a.x = B.d[o * STRIDE];
                            a.y = Ir[o];
                            B.mi[o * STRIDE] = a;
                        }
                    }
                    i_regs = fast;
                }
            }
        }
        if (BWD_UNROLLED && fast) {
            if constexpr (BWD_UNROLLED) {
                // cell o needs bM of old cell o and bI of old cell o-1.  Cells >= NCB keep (bM,bI) as one
                // double2: the word of cell o-1 is loaded at step o (its bI is needed now, its bM at the
                // next step); cells < NCB read bM from the dense plane and bI from registers.
                double cD = 0.;
                double nMraw = NC - 1 >= NCB ? B.mi[(NC - 1) * STRIDE].x : 0.;  // bM of old cell o, cells >= NCB
#pragma unroll
                for (int o = NC - 1; o >= 0; o--) {
                    double bm_old, bi_old = 0.;
                    if (o >= NCB) {
                        bm_old = nMraw;
                        if (o - 1 >= NCB) {
                            const SpD2 a = B.mi[(o > 0 ? o - 1 : 0) * STRIDE];
                            bi_old = a.y;
                            nMraw = a.x;
                        } else if (o > 0) {
                            bi_old = Ir[o - 1 >= 0 && o - 1 < NCB ? o - 1 : 0];
                        }
                    } else {
                        bm_old = B.d[o * STRIDE];
                        if (o > 0) bi_old = Ir[o > 0 && o - 1 < NCB ? o - 1 : 0];
                    }
                    const double nM = SP_DMUL(bm_old, r1);
                    const double em = ((mm.w[o >> 6] >> (o & 63)) & 1) ? emA : emB;
                    const double e = SP_DMUL(em, nM);
                    double X, bIv;
                    if (o > 0) {
                        const double qI = SP_DMUL(bi_old, r1);
                        X = SP_DADD(SP_DMUL(e, m0), SP_DMUL(eim1, qI));
                        bIv = SP_DADD(SP_DMUL(e, m3), SP_DMUL(eim4, qI));
                    } else {  // column beg is outside the band of row i+1: bI there is zero
                        X = SP_DADD(SP_DMUL(e, m0), 0.);
                        bIv = SP_DADD(SP_DMUL(e, m3), 0.);
                    }
                    const double bMv = SP_DADD(X, SP_DMUL(m2, cD));
                    if (o >= NCB) {
                        SpD2 v;
                        v.x = bMv;
                        v.y = bIv;
                        B.mi[o * STRIDE] = v;
                    } else {
                        B.d[o * STRIDE] = bMv;
                        Ir[o < NCB ? o : 0] = bIv;
                    }
                    cD = SP_DADD(SP_DMUL(e, m6e), SP_DMUL(m8e, cD));
                }
            }
        } else if (live) {
        // old row (i+1): cell o needs bI of old cell o-sh and bM of old cell o+1-sh
        const SpD2 *omi = B.mi - sh * STRIDE;
        double nM = (end + 1 <= end_next) ? SP_DMUL(omi[n * STRIDE].x, r1) : 0.;  // scaled bM[i+1][end+1]
        double cD = 0.;

        auto bwdP = [&](int o0, double (&X)[U], double (&bIv)[U], double (&w)[U]) {
            const uint32_t mb = mm.from(o0);
            double qM[U], qI[U];
#pragma unroll
            for (int j = 0; j < U; j++) {
                const SpD2 a = omi[(o0 + j) * STRIDE];
                qM[j] = SP_DMUL(a.x, r1); qI[j] = SP_DMUL(a.y, r1);
            }
#pragma unroll
            for (int j = U - 1; j >= 0; j--) {
                const double up = (j == U - 1) ? nM : qM[j < U - 1 ? j + 1 : 0];
                const double em = ((mb >> j) & 1) ? emA : emB;
                const double e = SP_DMUL(em, up);
                X[j] = SP_DADD(SP_DMUL(e, m0), SP_DMUL(eim1, qI[j]));
                bIv[j] = SP_DADD(SP_DMUL(e, m3), SP_DMUL(eim4, qI[j]));
                w[j] = SP_DMUL(e, m6e);
            }
            nM = qM[0];
        };
        auto bwdC = [&](int o0, const double (&X)[U], const double (&bIv)[U], const double (&w)[U]) {
#pragma unroll
            for (int j = U - 1; j >= 0; j--) {
                SpD2 v;
                v.x = SP_DADD(X[j], SP_DMUL(m2, cD));
                v.y = bIv[j];
                cD = SP_DADD(w[j], SP_DMUL(m8e, cD));
                B.mi[(o0 + j) * STRIDE] = v;
            }
        };
        const int nfull = has_n ? 0 : (n - 1) / U;  // the top cell always goes through the generic single-cell code
        for (int o = n - 1; o >= nfull * U; o--) {
            const SpD2 a = omi[o * STRIDE];
            const double qM = SP_DMUL(a.x, r1), qI = SP_DMUL(a.y, r1);
            const double em = (beg + o >= Lr) ? 0. : ((nn.from(o) & 1) ? 1. : ((mm.from(o) & 1) ? emA : emB));
            const double e = SP_DMUL(em, nM);
            SpD2 v;
            v.x = SP_DADD(SP_DADD(SP_DMUL(e, m0), SP_DMUL(eim1, qI)), SP_DMUL(m2, cD));
            v.y = SP_DADD(SP_DMUL(e, m3), SP_DMUL(eim4, qI));
            cD = SP_DADD(SP_DMUL(e, m6e), SP_DMUL(m8e, cD));
            B.mi[o * STRIDE] = v;
            nM = qM;
        }
        if (nfull > 0) {
            double XA[U], IA[U], wA[U], XB[U], IB[U], wB[U];
            int c = nfull - 1;
            bwdP(c * U, XA, IA, wA);
            c--;
            for (; c - 1 >= 0; c -= 2) {
                bwdP(c * U, XB, IB, wB);
                bwdC((c + 1) * U, XA, IA, wA);
                bwdP((c - 1) * U, XA, IA, wA);
                bwdC(c * U, XB, IB, wB);
            }
            if (c >= 0) {
                bwdP(c * U, XB, IB, wB);
                bwdC((c + 1) * U, XA, IA, wA);
                bwdC(c * U, XB, IB, wB);
            } else {
                bwdC((c + 1) * U, XA, IA, wA);
            }
        }
        }  // generic row body
        if (live) {
            beg_next = beg;
            end_next = end;
            r1 = y;
            if (t_next + 1 == i) {
                if constexpr (BWD_UNROLLED) {
                    if (i_regs) {  // map_row reads (bM,bI) from the double2 plane
#pragma unroll
                        for (int o = 0; o < NCB; o++) {
                            SpD2 a;
                            a.x = B.d[o * STRIDE];
                            a.y = Ir[o];
                            B.mi[o * STRIDE] = a;
                        }
                    }
                }
                map_row(nr, beg, n, y, true);
                nr--;
                t_next = nr >= 0 ? rows[nr].t : -2;
            }
        }
    }
}
#undef Ir
