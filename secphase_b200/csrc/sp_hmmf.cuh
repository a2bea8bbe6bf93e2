// sp_hmmf.cuh -- stage K4, fast arithmetic: the same banded glocal forward-backward HMM as
// sp_hmm2.cuh (htslib 1.17 probaln_glocal as called at ptMarker.c:754-757), one
// instance per lane, band in shared memory -- but evaluated as the cheapest mathematically
// equivalent program instead of in the reference's rounding order:
//
//  * products and sums are contracted to FMAs;
//  * no per-row normalising sum: the posterior the MAP step needs, f*b / sum_k(f*b), is invariant
//    under any per-row scale, so rows are only kept in range by an EXACT power-of-two factor that is
//    derived every SP_HMMF_RS rows from the largest exponent of the row and folded into the next row's
//    transition coefficients (power-of-two scaling commutes with rounding: the results are those of
//    an unscaled evaluation in a wider exponent range); hence no 1/s[i] array in HBM and no division
//    per row;
//  * a virtual band of NC = 2*BW+1 cells that slides by exactly one column per row for every row
//    (cell o of row i is column i-BW+o); columns outside [1, l_ref] or outside the instance's own
//    band |k-i| <= bw are held at exact zeros -- which is what the reference's zero padding means --
//    by an edge variant of the row body that runs only for rows that have such columns (or an N, or a
//    lane that has to keep the row for the MAP step, or a pending rescale);
//  * the forward sweep does not store the states (M,I,D) of a row but the two sums the next row
//    needs from them, G = m0*M + m3*I + m6*D (what flows into M one column to the right) and
//    H = EI*(m1*M + m4*I) (what flows into I of the same column): M[i,k] = e*G[i-1,k-1],
//    I[i,k] = H[i-1,k], and D lives only as the running chain value inside the row.  One 16-byte
//    word per cell, no D plane, no state registers between rows -- so the bodies are short loops over
//    chunks of cells with compile-time offsets and ten warps fit an SM;
//    8 FP64 instructions per cell forward, 8 backward (the strict kernel issues 18 + 14).
//
// What this costs: the bits of the intermediate posteriors differ from the reference's (relative
// drift of 1 - pmax measured <= 2e-11 on the benchmark workloads).  What is consumed
// downstream are integers -- the MAP state and q = (int)(-4.343*log(1-pmax)+.499) -- so every consumed
// row is GUARD-BANDED: if 1 - pmax lies within SP_HMMF_GUARD_ABS + SP_HMMF_GUARD_REL*(1-pmax) of one of
// the 101 decision thresholds (or of the pmax == 1 cliff), or the runner-up posterior is within
// SP_HMMF_TIE_REL of the maximum, or anything non-finite shows up, the instance is queued for the
// strict kernel (sp_hmm2.cuh), which recomputes it in the reference's order.  The integers that
// leave K4 are therefore the reference's.
#pragma once
#include <string.h>

#include "sp_hmm2.cuh"

#if defined(__CUDA_ARCH__)
#define SP_FMA(a, b, c) __fma_rn((a), (b), (c))
#else
#define SP_FMA(a, b, c) __builtin_fma((a), (b), (c))
#endif

#define SP_HMMF_RS 16                            // rows between two range checks
#define SP_HMMF_GUARD_ABS 7.105427357601002e-15  // 2^-47: 64 ulps of a posterior next to 1
#define SP_HMMF_GUARD_REL 1e-9
#define SP_HMMF_TIE_REL 1e-9
enum { SP_HMMF_NEAR_THRESHOLD = 1, SP_HMMF_NEAR_TIE = 2, SP_HMMF_NUMERIC = 4 };

// cells of the fast kernel's virtual band for a band class (sp_common.h): every class the shared-memory
// kernels cover (bw <= SP_H2_MAXBW); 0 = no fast body (the generic class)
SP_HD int sp_hmmf_class_cells(int cls) {
    const int bw = sp_class_bw(cls);
    return bw > 0 ? 2 * bw + 1 : 0;
}

SP_HD int sp_dbl_hi(double x) {
#if defined(__CUDA_ARCH__)
    return __double2hiint(x);
#else
    uint64_t u;
    memcpy(&u, &x, 8);
    return (int) (u >> 32);
#endif
}
SP_HD double sp_dbl_from_hi(int hi) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double(hi, 0);
#else
    uint64_t u = (uint64_t) (uint32_t) hi << 32;
    double x;
    memcpy(&x, &u, 8);
    return x;
#endif
}

template <int NW>
SP_HD void sp_bits_range(SpBits<NW> &b, int lo, int hi) {  // bits lo..hi (inclusive), empty when lo > hi
#pragma unroll
    for (int k = 0; k < NW; k++) {
        int l = lo - 64 * k, h = hi - 64 * k;
        l = l < 0 ? 0 : l;
        h = h > 63 ? 63 : h;
        b.w[k] = (l > h) ? 0 : ((~(uint64_t) 0 >> (63 - h)) & (~(uint64_t) 0 << l));
    }
}
template <int NW, int NC>
SP_HD bool sp_bits_all(const SpBits<NW> &b) {  // bits 0..NC-1 all set
    bool ok = true;
#pragma unroll
    for (int k = 0; k < NW; k++) {
        const int m = NC - 64 * k;
        const uint64_t want = m >= 64 ? ~(uint64_t) 0 : (((uint64_t) 1 << (m > 0 ? m : 0)) - 1);
        ok = ok && ((b.w[k] & want) == want);
    }
    return ok;
}
#define SP_BIT(b, o) (((b).w[(o) >> 6] >> ((o) & 63)) & 1)
template <int N>
struct SpInt {
    static constexpr int value = N;
};

// Returns the guard flags of the instance (0: every consumed row's state / q is safe to use).
// mi: this lane's cells, cell c (-1 <= c <= NC) at mi[c*STRIDE]; overwritten.
// fsave + r*fs_stride + 2*o: raw forward (M,I) of consumed row r, cell o (fs_stride >= 2*NC).
// Every lane of the warp must call this together (warp-uniform votes pick the row body); a lane that has
// nothing to do passes n_rows = 0.
// guard_all: band every one of the 101 thresholds (the stand-alone HMM API hands q itself to the caller); the
// pipeline only consumes min(q, 93) (ptMarker.c:786 clamps the written quality), so there the thresholds above
// 93 -- where 1 - pmax ~ 1e-10 is only a few hundred units of 2^-53 -- decide nothing and are not banded.
template <int STRIDE, int NC>
SP_HD int sp_hmmf_instance(const SpConst &C, const SpHmmIn &in, SpD2 *mi, double *fsave, int64_t fs_stride, SpRow *rows,
                           int n_rows, bool guard_all = true) {
    constexpr int NW = (NC + 63) / 64;
    constexpr int BW = (NC - 1) / 2;
    constexpr int CH = 8;  // cells per chunk: mask bits of a chunk never straddle a 64-bit word
    // the plain bodies of the narrow classes are unrolled completely (the chunks' independent parts overlap the
    // serial D chain of their neighbours); wide bands keep a loop of two chunks so that the code stays small
    constexpr int UF = NC <= 45 ? NC / CH : 2;
    const int Lr = in.l_ref, Lq = in.l_query;
    const int bw = sp_hmm_bw(Lr, Lq, in.par_bw);
    int flag = 0;
    if (bw > BW || Lq < 1 || Lr < 1) {  // not this class's instance (the launcher never sends one)
        flag = SP_HMMF_NUMERIC;
        n_rows = 0;
    }
    // transition matrix; float-typed sub-expressions were folded on the host
    const double sM = 1. / (double) (2 * Lq + 2);
    const double oms = 1. - sM;
    const double m0 = C.m0f * oms, m1 = C.d_d * oms, m2 = m1, m3 = C.ome_f * oms, m4 = C.e_d * oms;
    const double m6 = C.ome_f, m8 = C.e_d;
    const double eim1 = SP_HMM_EI * m1, eim4 = SP_HMM_EI * m4;
    const double emA = C.em_match, emB = C.em_mis;
    const SpD2 zero2 = {0., 0.};
    const bool narrow = bw < BW;  // the instance's own band is narrower than the class's: every row is an edge row

    SpBits<NW> p0, p1, p2;  // bit-planes of the reference codes under the band
    p0.clear(); p1.clear(); p2.clear();
    for (int c = -1; c <= NC; c++) mi[c * STRIDE] = zero2;

    // valid cells of row i: columns 1..Lr inside the instance's own band
    auto valid_range = [&](int i, int &lo, int &hi) {
        lo = BW - bw > BW + 1 - i ? BW - bw : BW + 1 - i;
        hi = BW + bw < Lr - i + BW ? BW + bw : Lr - i + BW;
    };

    // ------------------------------------------------------------------ forward, row 1
    int nr = 0;
    {
        const double bM = (double) SP_FDIV(C.omd_ff, (float) Lr), bI = (double) SP_FDIV(C.d_f, (float) Lr);
        const int qc = sp_query_code(in, 0);
        const int kend = Lr < bw + 1 ? Lr : bw + 1;
        const int kpl = Lr < BW + 1 ? Lr : BW + 1;  // columns under the virtual band of row 1
        double *fs = (nr < n_rows && rows[nr].t == 0) ? fsave : nullptr;  // only the stand-alone API asks for row 1
        if (fs)
            for (int o = 0; o < NC; o++) *reinterpret_cast<SpD2 *>(fs + 2 * o) = zero2;
        for (int k = 1; k <= kpl; k++) {
            const int rc = in.ref[k - 1], o = k + BW - 1;
            p0.or_bit(o, (uint64_t) (rc & 1));
            p1.or_bit(o, (uint64_t) ((rc >> 1) & 1));
            p2.or_bit(o, (uint64_t) ((rc >> 2) & 1));
            if (k <= kend) {
                const double M = sp_emis(C, rc, qc) * bM, I = SP_HMM_EI * bI;  // D[1,k] = 0
                SpD2 v;
                v.x = SP_FMA(m3, I, m0 * M);
                v.y = SP_FMA(eim4, I, eim1 * M);
                mi[o * STRIDE] = v;
                if (fs) {
                    SpD2 f = {M, I};
                    *reinterpret_cast<SpD2 *>(fs + 2 * o) = f;
                }
            }
        }
        if (fs) nr++;
    }
    // ------------------------------------------------------------------ forward, rows 2..Lq
    // All lanes of the warp walk the rows together up to the longest instance; a lane past its own last
    // row ("dead") keeps executing the body on its own cells, which nobody reads any more.
    const int LqW = SP_WARP_MAX(Lq);
    int t_next = nr < n_rows ? rows[nr].t : 0x7fffffff;
    uint32_t qraw_next = Lq >= 2 ? sp_query_raw(in, 1) : 0;
    uint32_t rc_next = 2 + BW <= Lr ? sp_ldg_u8(in.ref + 1 + BW) : 0;  // the column entering at row 2, if any
    double r = 1.;  // pending power-of-two scale, applied by the (edge) body of the row after a range check
    for (int i = 2; i <= LqW; i++) {
        const bool live = i <= Lq;
        int qc = 0;
        if (live) {
            qc = sp_query_decode(in, i - 1, qraw_next);
            const uint32_t rc_in = rc_next;
            if (i < Lq) qraw_next = sp_query_raw(in, i);
            if (i + 1 + BW <= Lr) rc_next = sp_ldg_u8(in.ref + i + BW);
            p0.shr1(); p1.shr1(); p2.shr1();
            if (i + BW <= Lr) {  // one column enters the band on the right
                p0.or_bit(NC - 1, (uint64_t) (rc_in & 1));
                p1.or_bit(NC - 1, (uint64_t) ((rc_in >> 1) & 1));
                p2.or_bit(NC - 1, (uint64_t) ((rc_in >> 2) & 1));
            }
        }
        SpBits<NW> mm, nn;
        sp_h2_row_masks(p0, p1, p2, qc, mm, nn);
        const bool save = live && t_next + 1 == i;
        const bool rescale = i % SP_HMMF_RS == 1;  // (warp-uniform) r was derived from the row before
        // rows whose every cell is a valid column and that need nothing special run the plain body
        const bool plain = !rescale && SP_WARP_ALL(!live || (!narrow && !save && i > BW && i + BW <= Lr && !nn.any_below(NC)));
        // cell o: M = e * G_old[o], I = H_old[o+1], D = m8*D[o-1] + m2*M[o-1]; stores G, H of the new row
        double Gcur = mi[0].x, Mlast = 0., cD = 0.;
        if (plain) {
            auto chunk = [&](int o0, auto nc_tag) {
                constexpr int N = decltype(nc_tag)::value;
                const uint32_t mb = mm.from(o0);
#pragma unroll
                for (int j = 0; j < N; j++) {
                    const SpD2 a = mi[(o0 + j + 1) * STRIDE];  // (cell NC holds zeros)
                    const double M = ((mb >> j) & 1 ? emA : emB) * Gcur;
                    const double I = a.y;
                    cD = SP_FMA(m8, cD, m2 * Mlast);
                    SpD2 v;
                    v.x = SP_FMA(m6, cD, SP_FMA(m3, I, m0 * M));
                    v.y = SP_FMA(eim4, I, eim1 * M);
                    mi[(o0 + j) * STRIDE] = v;
                    Mlast = M;
                    Gcur = a.x;
                }
            };
#pragma unroll UF
            for (int o0 = 0; o0 + CH <= NC; o0 += CH) chunk(o0, SpInt<CH>());
            if constexpr (NC % CH != 0) chunk(NC - NC % CH, SpInt<NC % CH>());
        } else {
            SpBits<NW> vm;
            int lo, hi;
            valid_range(i, lo, hi);
            sp_bits_range(vm, lo, hi);
            const double eA = emA * r, eB = emB * r, eN = r;
            double *fs = save ? fsave + (int64_t) nr * fs_stride : nullptr;
            auto chunk = [&](int o0, auto nc_tag) {
                constexpr int N = decltype(nc_tag)::value;
                const uint32_t mb = mm.from(o0), nb = nn.from(o0), vb = vm.from(o0);
#pragma unroll
                for (int j = 0; j < N; j++) {
                    const SpD2 a = mi[(o0 + j + 1) * STRIDE];
                    const bool ok = (vb >> j) & 1;
                    const double e = (nb >> j) & 1 ? eN : ((mb >> j) & 1 ? eA : eB);
                    const double M = ok ? e * Gcur : 0.;
                    const double I = ok ? r * a.y : 0.;
                    cD = ok ? SP_FMA(m8, cD, m2 * Mlast) : 0.;
                    SpD2 v;
                    v.x = SP_FMA(m6, cD, SP_FMA(m3, I, m0 * M));
                    v.y = SP_FMA(eim4, I, eim1 * M);
                    mi[(o0 + j) * STRIDE] = v;
                    if (fs) {
                        SpD2 f = {M, I};
                        *reinterpret_cast<SpD2 *>(fs + 2 * (o0 + j)) = f;
                    }
                    Mlast = M;
                    Gcur = a.x;
                }
            };
#pragma unroll 2
            for (int o0 = 0; o0 + CH <= NC; o0 += CH) chunk(o0, SpInt<CH>());
            if constexpr (NC % CH != 0) chunk(NC - NC % CH, SpInt<NC % CH>());
            r = 1.;
            if (save) {
                nr++;
                t_next = nr < n_rows ? rows[nr].t : 0x7fffffff;
            }
        }
        if (i % SP_HMMF_RS == 0) {  // range check: largest exponent of the row -> exact power-of-two scale
            int mh = 0;
#pragma unroll 8
            for (int o = 0; o < NC; o++) {
                const SpD2 a = mi[o * STRIDE];
                const int hx = sp_dbl_hi(a.x), hy = sp_dbl_hi(a.y);
                mh = hx > mh ? hx : mh;
                mh = hy > mh ? hy : mh;
            }
            const int ex = (mh >> 20) & 0x7ff;
            if (ex == 0 || ex == 0x7ff || mh < 0) {
                if (live) flag |= SP_HMMF_NUMERIC;
                r = 1.;
            } else {
                r = sp_dbl_from_hi((2046 - ex) << 20);  // 2^(1023-ex)
            }
        }
    }
    // ------------------------------------------------------------------ backward (+ MAP at consumed rows)
    int i_stop = n_rows > 0 ? rows[0].t + 1 : Lq;  // nothing below the lowest consumed row is needed
    // MAP of one row: max / runner-up / sum of f*b over the band, the decision and its guard band
    auto map_row = [&](int ri) {
        const double *fs = fsave + (int64_t) ri * fs_stride;
        double sum = 0., mx = 0., mx2 = 0.;
        int max_o = -1;
#pragma unroll 1
        for (int o = 0; o < NC; o++) {  // (a handful of rows per instance: keep it small)
            const SpD2 f = *reinterpret_cast<const SpD2 *>(fs + 2 * o);
            const SpD2 b = mi[o * STRIDE];
            double z = f.x * b.x;
            if (z > mx) { mx2 = mx; mx = z; max_o = o << 2; }
            else if (z > mx2) mx2 = z;
            sum = sum + z;
            z = f.y * b.y;
            if (z > mx) { mx2 = mx; mx = z; max_o = o << 2 | 1; }
            else if (z > mx2) mx2 = z;
            sum = sum + z;
        }
        const int i = rows[ri].t + 1;
        const double pm = SP_DDIV(mx, sum);
        const double t = 1. - pm;
        // q, and how close t is to a decision threshold: qthr[lo] >= t > qthr[lo+1]
        int lo = 0, hi = 101;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (t <= C.qthr[mid]) lo = mid; else hi = mid - 1;
        }
        const double g = SP_HMMF_GUARD_ABS + SP_HMMF_GUARD_REL * t;
        if (!(t > g)) flag |= SP_HMMF_NEAR_THRESHOLD;  // next to the pmax == 1 cliff (q = 0), or NaN
        const int top = guard_all ? 101 : 93;
        if (lo >= 1 && lo <= top && C.qthr[lo] - t <= g) flag |= SP_HMMF_NEAR_THRESHOLD;
        if (lo < top && t - C.qthr[lo + 1] <= g) flag |= SP_HMMF_NEAR_THRESHOLD;
        if (!(mx2 < mx * (1. - SP_HMMF_TIE_REL))) flag |= SP_HMMF_NEAR_TIE;  // also mx == 0 and NaN
        if (!(sum > 0.) || !(sum < 1.7976931348623157e308)) flag |= SP_HMMF_NUMERIC;
        rows[ri].state = max_o < 0 ? -1 : (((i - BW + (max_o >> 2)) - 1) << 2 | (max_o & 3));
        rows[ri].pmax = pm;
        rows[ri].q = !(t > 0.) ? 0 : (lo > 100 ? 99 : lo);
    };
    for (int c = -1; c <= NC; c++) mi[c * STRIDE] = zero2;
    if (n_rows > 0) {  // row Lq: constant inside the band (any constant: the posterior is scale-free)
        int lo, hi;
        valid_range(Lq, lo, hi);
        const SpD2 one2 = {1., 1.};
        for (int o = lo < 0 ? 0 : lo; o <= hi && o < NC; o++) mi[o * STRIDE] = one2;
    }
    nr = n_rows - 1;
    if (nr >= 0 && rows[nr].t + 1 == Lq) {  // stand-alone API only (pipeline rows satisfy t <= Lq-12)
        map_row(nr);
        nr--;
    }
    if (nr < 0) i_stop = Lq;  // nothing (left) to do for this lane: no live step below
    t_next = nr >= 0 ? rows[nr].t : -2;
    // the planes are where the lane's last forward row left them: bit o <-> ref[Lq-1-BW+o], which is the
    // base of column k+1 for cell o of row Lq-1
    qraw_next = Lq >= 2 ? sp_query_raw(in, Lq - 1) : 0;
    rc_next = (Lq - 2 - BW >= 0 && Lq - 2 - BW < Lr) ? sp_ldg_u8(in.ref + (Lq - 2 - BW)) : 0;
    const int jmax = SP_WARP_MAX(Lq - 1 - i_stop);
    double r1 = 1.;
    for (int j = 0; j <= jmax; j++) {
        const int i = Lq - 1 - j;
        const bool live = i >= i_stop && i >= 1;
        int qc = 0;
        if (live) {
            qc = sp_query_decode(in, i, qraw_next);  // query[i] (0-based) == base of row i+1
            const uint32_t rc_in = rc_next;  // (step 0 does not shift: rc_next then already holds step 1's code)
            if (i > i_stop) {
                qraw_next = sp_query_raw(in, i - 1);
                if (j > 0) {
                    const int x = i - 1 - BW;  // ref index entering at the next step
                    rc_next = (x >= 0 && x < Lr) ? sp_ldg_u8(in.ref + x) : 0;
                }
            }
            if (j > 0) {  // one column enters the band on the left: bit 0 <-> ref[i-BW]
                const int x = i - BW;
                const uint32_t rc = (x >= 0 && x < Lr) ? rc_in : 0;
                p0.shl1_in((uint64_t) (rc & 1));
                p1.shl1_in((uint64_t) ((rc >> 1) & 1));
                p2.shl1_in((uint64_t) ((rc >> 2) & 1));
            }
        }
        SpBits<NW> mm, nn;
        sp_h2_row_masks(p0, p1, p2, qc, mm, nn);
        const bool rescale = j % SP_HMMF_RS == 0 && j > 0;  // (warp-uniform) r1 was derived at the step before
        // Cells left of column 1 or right of l_ref need no mask here: with row i+1 zero outside its valid cells
        // the right side stays zero by itself and what appears left of column 1 never flows back into a valid
        // cell (dependencies only run towards smaller columns) nor into the MAP (f is zero there).  Only an
        // instance narrower than its class, or an N, takes the edge body.
        const bool plain = !rescale && SP_WARP_ALL(!live || (!narrow && !nn.any_below(NC)));
        const double m6e = i > 1 ? m6 : 0., m8e = i > 1 ? m8 : 0.;
        // cell o needs bM of old cell o (column k+1 of row i+1) and bI of old cell o-1 (column k)
        double cD = 0., bMo = mi[(NC - 1) * STRIDE].x;
        if (plain) {
            auto chunk = [&](int o0, auto nc_tag) {
                constexpr int N = decltype(nc_tag)::value;
                const uint32_t mb = mm.from(o0);
#pragma unroll
                for (int jj = N - 1; jj >= 0; jj--) {
                    const SpD2 a = mi[(o0 + jj - 1) * STRIDE];  // (cell -1 holds zeros)
                    const double e = ((mb >> jj) & 1 ? emA : emB) * bMo;
                    SpD2 v;
                    v.x = SP_FMA(m2, cD, SP_FMA(e, m0, eim1 * a.y));
                    v.y = SP_FMA(e, m3, eim4 * a.y);
                    cD = SP_FMA(m8e, cD, e * m6e);
                    mi[(o0 + jj) * STRIDE] = v;
                    bMo = a.x;
                }
            };
            if constexpr (NC % CH != 0) chunk(NC - NC % CH, SpInt<NC % CH>());
#pragma unroll UF
            for (int o0 = NC - NC % CH - CH; o0 >= 0; o0 -= CH) chunk(o0, SpInt<CH>());
        } else {
            SpBits<NW> vm;
            int lo, hi;
            valid_range(i, lo, hi);
            sp_bits_range(vm, lo, hi);
            const double eA = emA * r1, eB = emB * r1, eN = r1, c1 = eim1 * r1, c4 = eim4 * r1;
            auto chunk = [&](int o0, auto nc_tag) {
                constexpr int N = decltype(nc_tag)::value;
                const uint32_t mb = mm.from(o0), nb = nn.from(o0), vb = vm.from(o0);
#pragma unroll
                for (int jj = N - 1; jj >= 0; jj--) {
                    const SpD2 a = mi[(o0 + jj - 1) * STRIDE];
                    const double e = ((nb >> jj) & 1 ? eN : ((mb >> jj) & 1 ? eA : eB)) * bMo;
                    SpD2 v;
                    v.x = SP_FMA(m2, cD, SP_FMA(e, m0, c1 * a.y));
                    v.y = SP_FMA(e, m3, c4 * a.y);
                    cD = SP_FMA(m8e, cD, e * m6e);
                    if (!((vb >> jj) & 1)) { v.x = 0.; v.y = 0.; cD = 0.; }
                    mi[(o0 + jj) * STRIDE] = v;
                    bMo = a.x;
                }
            };
            if constexpr (NC % CH != 0) chunk(NC - NC % CH, SpInt<NC % CH>());
#pragma unroll 2
            for (int o0 = NC - NC % CH - CH; o0 >= 0; o0 -= CH) chunk(o0, SpInt<CH>());
            r1 = 1.;
        }
        if (live && t_next + 1 == i) {
            map_row(nr);
            nr--;
            t_next = nr >= 0 ? rows[nr].t : -2;
        }
        if (j % SP_HMMF_RS == SP_HMMF_RS - 1) {
            int mh = 0;
#pragma unroll 8
            for (int o = 0; o < NC; o++) {
                const SpD2 a = mi[o * STRIDE];
                const int hx = sp_dbl_hi(a.x), hy = sp_dbl_hi(a.y);
                mh = hx > mh ? hx : mh;
                mh = hy > mh ? hy : mh;
            }
            const int ex = (mh >> 20) & 0x7ff;
            if (ex == 0 || ex == 0x7ff || mh < 0) {
                if (live) flag |= SP_HMMF_NUMERIC;
                r1 = 1.;
            } else {
                r1 = sp_dbl_from_hi((2046 - ex) << 20);
            }
        }
    }
    return flag;
}
