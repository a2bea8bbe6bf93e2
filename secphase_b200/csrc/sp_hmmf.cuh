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
//  * a virtual band of NC = 2*BW+1 cells (BW = the widest band among the warp's 32 instances; instances
//    are sorted by band width, so almost every warp is uniform) that slides by exactly one column per row
//    (cell o of row i is column i-BW+o); columns outside [1, l_ref] or outside the instance's own
//    band |k-i| <= bw are held at exact zeros -- which is what the reference's zero padding means --
//    by an edge variant of the row body that runs only for rows that have such columns (or an N, or a
//    lane that has to keep the row for the MAP step, or a pending rescale);
//  * the forward sweep does not store the states (M,I,D) of a row but the two sums the next row
//    needs from them, G = m0*M + m3*I + m6*D (what flows into M one column to the right) and
//    H = EI*(m1*M + m4*I) (what flows into I of the same column): M[i,k] = e*G[i-1,k-1],
//    I[i,k] = H[i-1,k], and D lives only as the running chain value inside the row.  One 16-byte
//    word per cell, no D plane, no state registers between rows, ten warps per SM;
//    8 FP64 instructions per cell forward, 8 backward (the strict kernel issues 18 + 14);
//  * with that little arithmetic per cell the kernel would be bound by shared-memory bandwidth (one
//    16-byte load and store per cell and row: 8 SM cycles per warp against 4 of FP64 issue), so up to
//    SP_HMMF_RB consecutive rows are computed in ONE pass over the band: row r of the block runs r-1
//    cells behind row r-1 and takes its inputs (G of the same cell, H of the next) straight from
//    registers; only the block's last row is written back.  Blocks end where a lane needs a whole
//    row in memory (a consumed row) and never contain an edge row.  The passes are loops over chunks of
//    eight cells (the band width is a run-time value), small enough to live in the instruction cache.
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

#ifndef SP_HMMF_RS
#define SP_HMMF_RS 16
#endif
#ifndef SP_HMMF_RB
#define SP_HMMF_RB 4  // rows per pass of the plain bodies
#endif
#ifndef SP_HMMF_CHUNK
#define SP_HMMF_CHUNK 4  // cells per iteration of a pass's steady loop (one mask extraction per row and chunk)
#endif
// a global byte load that stays where it is written (the rows' codes are fetched BEFORE the pass that hides their
// latency; a plain load would be sunk to its first use behind the pass)
SP_HD uint32_t sp_ldg_u8_here(const uint8_t *p) {
#if defined(__CUDA_ARCH__)
    uint32_t v;
    asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
#else
    return *p;
#endif
}                            // rows between two range checks
#define SP_HMMF_GUARD_ABS 7.105427357601002e-15  // 2^-47: 64 ulps of a posterior next to 1
#define SP_HMMF_GUARD_REL 1e-9
#define SP_HMMF_TIE_REL 1e-9
enum { SP_HMMF_NEAR_THRESHOLD = 1, SP_HMMF_NEAR_TIE = 2, SP_HMMF_NUMERIC = 4 };

// cells of the fast kernel's slab for a band class (sp_common.h), 0 = the class always runs the strict kernel.
// Row masks of one (bw <= 27: 99 % of the HiFi band cells) or two 64-bit words (bw <= 62); the launcher sends a
// two-word class to the fast kernel only when it is populated enough for its 32-instance sets to be uniform in
// width (ONT), see launch_hmm.
SP_HD int sp_hmmf_class_cells(int cls) {
    const int bw = sp_class_bw(cls);
    return (bw > 0 && 2 * bw + 1 <= 128) ? 2 * bw + 1 : 0;
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
template <int NW>
SP_HD uint32_t sp_bit_rt(const SpBits<NW> &b, int o) {  // bit o, run-time position (selects, no indexing)
    uint64_t x = b.w[0];
#pragma unroll
    for (int k = 1; k < NW; k++)
        if ((o >> 6) == k) x = b.w[k];
    return (uint32_t) (x >> (o & 63)) & 1;
}
template <int NW>
SP_HD uint32_t sp_bits8_rt(const SpBits<NW> &b, int o) {  // bits o..o+7 (o >= 0), may straddle two words
    uint64_t lo = b.w[0], hi = NW > 1 ? b.w[NW > 1 ? 1 : 0] : 0;
#pragma unroll
    for (int k = 1; k < NW; k++)
        if ((o >> 6) == k) { lo = b.w[k]; hi = k + 1 < NW ? b.w[k + 1 < NW ? k + 1 : k] : 0; }
    const int sh = o & 63;
    uint64_t x = lo >> sh;
    if (sh > 56) x |= hi << (64 - sh);
    return (uint32_t) x & 0xff;
}
template <int N>
struct SpInt {
    static constexpr int value = N;
};
// f(SpInt<A>()), f(SpInt<A+1>()), ... f(SpInt<B-1>()): a loop whose counter is a compile-time constant in the body
template <int A, int B, class F>
SP_HD void sp_static_for(F &&f) {
    if constexpr (A < B) {
        f(SpInt<A>());
        sp_static_for<A + 1, B>(f);
    }
}

// Returns the guard flags of the instance (0: every consumed row's state / q is safe to use).
// mi: this lane's cells, cell c (-1 <= c <= 2*BW+1) at mi[c*STRIDE]; overwritten.
// BW: half-width of the virtual band, >= the instance's own band half-width and the same for every lane of the
// warp (the kernel passes the warp's maximum); NW: 64-bit words of a row mask, 64*NW >= 2*BW+1.
// fsave + r*fs_stride + 2*o: raw forward (M,I) of consumed row r, cell o (fs_stride >= 2*(2*BW+1)).
// Every lane of the warp must call this together (warp-uniform votes pick the row bodies); a lane that has
// nothing to do passes n_rows = 0.
// guard_all: band every one of the 101 thresholds and every near-tie (the stand-alone HMM API hands state and q
// themselves to the caller).  The pipeline (guard_all == false) consumes of a row only
//     (state is M at the alignment's own column `expected`) ? min(raw quality, min(q, 93)) : 0
// (ptMarker.c:772-786, sp_resolve_q), so there only two things can change the outcome: a runner-up within the
// tie band when the expected state is the winner or the runner-up, and q's thresholds <= 93 when it wins.
template <int STRIDE, int NW>
SP_HD int sp_hmmf_instance(const SpConst &C, const SpHmmIn &in, SpD2 *mi, int BW, double *fsave, int64_t fs_stride,
                           SpRow *rows, int n_rows, bool guard_all = true) {
    constexpr int RB = NW == 1 ? SP_HMMF_RB : 2;  // wide bands: fewer mask registers
    const int NC = 2 * BW + 1;
    const int Lr = in.l_ref, Lq = in.l_query;
    const int bw = sp_hmm_bw(Lr, Lq, in.par_bw);
    int flag = 0;
    if (bw > BW || NC > 64 * NW || Lq < 1 || Lr < 1) {  // (the launcher never sends such an instance)
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
    const bool narrow = bw < BW;  // the instance's own band is narrower than the warp's: every row is an edge row

    SpBits<NW> p0, p1, p2;  // bit-planes of the reference codes under the band
    p0.clear(); p1.clear(); p2.clear();
    for (int c = -1; c <= NC; c++) mi[c * STRIDE] = zero2;

    // valid cells of row i: columns 1..Lr inside the instance's own band
    auto valid_range = [&](int i, int &lo, int &hi) {
        lo = BW - bw > BW + 1 - i ? BW - bw : BW + 1 - i;
        hi = BW + bw < Lr - i + BW ? BW + bw : Lr - i + BW;
    };
    // range check of the row in memory: largest exponent -> rescale the row by an exact power of two when it
    // has drifted far from 1 (it rarely has: a row loses ~1e-4 per mismatch)
    auto range_check = [&](bool live) {
        int mh = 0;
#pragma unroll 4
        for (int o = 0; o < NC; o++) {
            const SpD2 a = mi[o * STRIDE];
            const int hx = sp_dbl_hi(a.x), hy = sp_dbl_hi(a.y);
            mh = hx > mh ? hx : mh;
            mh = hy > mh ? hy : mh;
        }
        const int ex = (mh >> 20) & 0x7ff;
        if (ex == 0 || ex == 0x7ff || mh < 0) {
            if (live) flag |= SP_HMMF_NUMERIC;
        } else if (ex < 1023 - 64 || ex > 1023 + 64) {
            const double sc = sp_dbl_from_hi((2046 - ex) << 20);  // 2^(1023-ex)
#pragma unroll 4
            for (int o = 0; o < NC; o++) {
                SpD2 a = mi[o * STRIDE];
                a.x *= sc;
                a.y *= sc;
                mi[o * STRIDE] = a;
            }
        }
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
    // row ("dead") keeps executing the bodies on its own cells, which nobody reads any more.
    const int LqW = SP_WARP_MAX(Lq);
    int t_next = nr < n_rows ? rows[nr].t : 0x7fffffff;
    // raw query bytes and entering reference codes of the next RB rows (kept apart: nothing may hang off a load
    // before the pass that hides its latency has run)
    uint32_t qn[RB], rn[RB], q4 = 0, r4 = 0;
    auto fetch_fwd = [&](int i0) {
#pragma unroll
        for (int rr = 0; rr < RB; rr++) {
            const int row = i0 + rr;
            qn[rr] = row <= Lq ? sp_ldg_u8_here(in.qbytes ? in.qbytes + in.q0 + (row - 1) : in.qseq4 + ((in.q0 + row - 1) >> 1)) : 0;
            rn[rr] = row + BW <= Lr ? sp_ldg_u8_here(in.ref + row + BW - 1) : 0;
        }
    };
    auto pack = [&]() {  // byte r <-> row i+r.  Called after the pass: the barrier keeps the compiler from pulling the
        q4 = 0;          // shifts up to the loads (where they would wait for them) -- volatile asms keep their order
        r4 = 0;
#pragma unroll
        for (int rr = 0; rr < RB; rr++) {
#if defined(__CUDA_ARCH__)
            asm volatile("" : "+r"(qn[rr]), "+r"(rn[rr])::"memory");
#endif
            q4 |= qn[rr] << (8 * rr);
            r4 |= rn[rr] << (8 * rr);
        }
    };
    fetch_fwd(2);
    pack();
    int since_check = 0;
    for (int i = 2; i <= LqW;) {
        // ---- rows of this pass: a consumed row ends the block (it is kept in memory), else RB rows
        int lane_r = RB;
        if (t_next + 1 >= i && t_next + 1 < i + RB && t_next + 1 <= Lq) lane_r = t_next + 2 - i;
        const int R = SP_WARP_MIN(lane_r);
        // ---- advance the bit-planes over those rows; emission / N masks of each; does any row need the masked pass
        // (an invalid column -- outside 1..l_ref or the instance's own band -- or an N)
        SpBits<NW> mmr[RB], nnr[RB];
        bool mask_lane = false;
#pragma unroll
        for (int rr = 0; rr < RB; rr++) {
            if (rr < R) {
                const int row = i + rr;
                const bool live = row <= Lq;
                int qc = 0;
                if (live) {
                    qc = sp_query_decode(in, row - 1, (q4 >> (8 * rr)) & 0xff);
                    const uint32_t rc = (r4 >> (8 * rr)) & 0xff;
                    p0.shr1(); p1.shr1(); p2.shr1();
                    if (row + BW <= Lr) {  // one column enters the band on the right
                        p0.or_bit(NC - 1, (uint64_t) (rc & 1));
                        p1.or_bit(NC - 1, (uint64_t) ((rc >> 1) & 1));
                        p2.or_bit(NC - 1, (uint64_t) ((rc >> 2) & 1));
                    }
                }
                sp_h2_row_masks(p0, p1, p2, qc, mmr[rr], nnr[rr]);
                if (live && (narrow || row <= BW || row + BW > Lr || nnr[rr].any_below(NC))) mask_lane = true;
            } else {
                mmr[rr].clear();
                nnr[rr].clear();
            }
        }
        fetch_fwd(i + R);  // the codes of the rows after this pass: their latency hides behind it
        const bool masked = SP_WARP_ANY(mask_lane);
        SpBits<NW> vmr[RB];
        if (masked) {
#pragma unroll
            for (int rr = 0; rr < RB; rr++) {
                int lo, hi;
                valid_range(i + rr, lo, hi);
                sp_bits_range(vmr[rr], lo, hi);
            }
        } else {
#pragma unroll
            for (int rr = 0; rr < RB; rr++) vmr[rr].clear();
        }
        const int last = i + R - 1;
        const bool save = last <= Lq && t_next + 1 == last;
        double *fs = save ? fsave + (int64_t) nr * fs_stride : nullptr;
        // ---- one pass over R rows.  Row r (0-based) computes cell s - r at step s:
        //   M = e*G_above[o], I = H_above[o+1], D = m8*D[o-1] + m2*M[o-1], G = m0*M + m3*I + m6*D, H = EI*(m1*M + m4*I)
        // with "above" = the stored row for r = 0, else row r-1 of this pass (G from its previous step, H fresh).
        // MK: invalid cells are forced to zero and an N emits 1 (rows at the ends of the window, N bases).
        auto pass = [&](auto rtag, auto mtag) {
            constexpr int RR = decltype(rtag)::value;
            constexpr bool MK = decltype(mtag)::value != 0;
            double Gc[RR], Ml[RR], cD[RR];
#pragma unroll
            for (int rr = 0; rr < RR; rr++) { Gc[rr] = 0.; Ml[rr] = 0.; cD[rr] = 0.; }
            double g0 = mi[0].x;  // G of the stored row at the cell row 0 computes next
            // one step; rows [RA, RBB) are inside the band, rows < RA have finished; bits(rr) = mask bits of row rr's cell:
            // bit 0 match, bit 1 N, bit 2 valid
            auto step = [&](SpD2 *cell, int sidx, auto ra_tag, auto rb_tag, auto bits) {  // cell = &mi[sidx]
                constexpr int RA = decltype(ra_tag)::value, RBB = decltype(rb_tag)::value;
                double Gup = 0., Hup = 0.;
                if constexpr (RA == 0) {
                    const SpD2 a = cell[STRIDE];  // cell s+1 of the stored row (cell NC holds zeros)
                    Gup = g0;
                    Hup = a.y;
                    g0 = a.x;
                } else {
                    Gup = Gc[RA - 1];  // the row above has finished: its last cell's G, H = 0 beyond the band
                }
#pragma unroll
                for (int rr = RA; rr < RBB; rr++) {
                    const uint32_t b = bits(rr);
                    double M, I;
                    if constexpr (MK) {
                        const bool ok = (b & 4) != 0;
                        const double e = (b & 2) ? 1. : ((b & 1) ? emA : emB);
                        M = ok ? e * Gup : 0.;
                        I = ok ? Hup : 0.;
                        cD[rr] = ok ? SP_FMA(m8, cD[rr], m2 * Ml[rr]) : 0.;
                    } else {
                        M = ((b & 1) ? emA : emB) * Gup;
                        I = Hup;
                        cD[rr] = SP_FMA(m8, cD[rr], m2 * Ml[rr]);
                    }
                    const double G = SP_FMA(m6, cD[rr], SP_FMA(m3, I, m0 * M));
                    const double H = SP_FMA(eim4, I, eim1 * M);
                    Ml[rr] = M;
                    Gup = Gc[rr];  // what row rr+1 reads: G of ITS cell (this row's previous cell), H of this cell
                    Hup = H;
                    Gc[rr] = G;
                    if (rr == RR - 1) {
                        SpD2 v = {G, H};
                        cell[-(RR - 1) * STRIDE] = v;  // cell s-(RR-1)
                        if (fs) {
                            SpD2 f = {M, I};
                            *reinterpret_cast<SpD2 *>(fs + 2 * (sidx - (RR - 1))) = f;
                        }
                    }
                }
            };
            auto bit_rt = [&](int rr, int o) -> uint32_t {
                uint32_t v = sp_bit_rt(mmr[rr], o);
                if constexpr (MK) v |= sp_bit_rt(nnr[rr], o) << 1 | sp_bit_rt(vmr[rr], o) << 2;
                return v;
            };
            // fill: steps 0..RR-2, rows 0..s
            sp_static_for<0, RR - 1>([&](auto st) {
                constexpr int S = decltype(st)::value;
                step(mi + S * STRIDE, S, SpInt<0>(), SpInt<S + 1>(), [&](int rr) { return bit_rt(rr, S - rr); });
            });
            // steady: steps RR-1..NC-1, all rows; chunks of 8 steps share one mask extraction per row
            int s = RR - 1;
            constexpr int CK = SP_HMMF_CHUNK;
            for (; s + CK <= NC; s += CK) {
                uint32_t mb[RR], nb[MK ? RR : 1], vb[MK ? RR : 1];
#pragma unroll
                for (int rr = 0; rr < RR; rr++) {
                    mb[rr] = sp_bits8_rt(mmr[rr], s - rr);
                    if constexpr (MK) {
                        nb[rr] = sp_bits8_rt(nnr[rr], s - rr);
                        vb[rr] = sp_bits8_rt(vmr[rr], s - rr);
                    }
                }
                SpD2 *cell = mi + s * STRIDE;
                sp_static_for<0, CK>([&](auto jt) {
                    constexpr int J = decltype(jt)::value;
                    step(cell + J * STRIDE, s + J, SpInt<0>(), SpInt<RR>(), [&](int rr) -> uint32_t {
                        uint32_t v = (mb[rr] >> J) & 1;
                        if constexpr (MK) v |= ((nb[rr] >> J) & 1) << 1 | ((vb[rr] >> J) & 1) << 2;
                        return v;
                    });
                });
            }
            for (; s < NC; s++) step(mi + s * STRIDE, s, SpInt<0>(), SpInt<RR>(), [&](int rr) { return bit_rt(rr, s - rr); });
            // drain: steps NC..NC+RR-2, rows d+1..RR-1
            sp_static_for<0, RR - 1>([&](auto dt) {
                constexpr int D = decltype(dt)::value;
                step(mi + (NC + D) * STRIDE, NC + D, SpInt<D + 1>(), SpInt<RR>(), [&](int rr) { return bit_rt(rr, NC + D - rr); });
            });
        };
        if (!masked) {
            if (R == 1) pass(SpInt<1>(), SpInt<0>());
            else if (RB >= 2 && R == 2) pass(SpInt<(RB >= 2 ? 2 : 1)>(), SpInt<0>());
            else if (RB >= 3 && R == 3) pass(SpInt<(RB >= 3 ? 3 : 1)>(), SpInt<0>());
            else pass(SpInt<RB>(), SpInt<0>());
        } else {
            if (R == 1) pass(SpInt<1>(), SpInt<1>());
            else if (RB >= 2 && R == 2) pass(SpInt<(RB >= 2 ? 2 : 1)>(), SpInt<1>());
            else if (RB >= 3 && R == 3) pass(SpInt<(RB >= 3 ? 3 : 1)>(), SpInt<1>());
            else pass(SpInt<RB>(), SpInt<1>());
        }
        if (save) {
            nr++;
            t_next = nr < n_rows ? rows[nr].t : 0x7fffffff;
        }
        i += R;
        pack();
        since_check += R;
        if (since_check >= SP_HMMF_RS) {
            range_check(i - 1 <= Lq);
            since_check = 0;
        }
    }
    // ------------------------------------------------------------------ backward (+ MAP at consumed rows)
    int i_stop = n_rows > 0 ? rows[0].t + 1 : Lq;  // nothing below the lowest consumed row is needed
    auto map_row = [&](int ri) {
        const double *fs = fsave + (int64_t) ri * fs_stride;
        const int i = rows[ri].t + 1;
        const int o_exp = guard_all ? -1 : rows[ri].expected + 1 - (i - BW);  // cell of column expected+1
        double sum = 0., mx = 0., mx2 = 0., zE = 0.;
        int max_o = -1;
#pragma unroll 1
        for (int o = 0; o < NC; o++) {  // (a handful of rows per instance: keep it small)
            const SpD2 f = *reinterpret_cast<const SpD2 *>(fs + 2 * o);
            const SpD2 b = mi[o * STRIDE];
            double z = f.x * b.x;
            if (o == o_exp) zE = z;
            if (z > mx) { mx2 = mx; mx = z; max_o = o << 2; }
            else if (z > mx2) mx2 = z;
            sum = sum + z;
            z = f.y * b.y;
            if (z > mx) { mx2 = mx; mx = z; max_o = o << 2 | 1; }
            else if (z > mx2) mx2 = z;
            sum = sum + z;
        }
        const double pm = SP_DDIV(mx, sum);
        const double t = 1. - pm;
        // q, and how close t is to a decision threshold: qthr[lo] >= t > qthr[lo+1]
        int lo = 0, hi = 101;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (t <= C.qthr[mid]) lo = mid; else hi = mid - 1;
        }
        const double g = SP_HMMF_GUARD_ABS + SP_HMMF_GUARD_REL * t;
        const double tie = mx * (1. - SP_HMMF_TIE_REL);
        const bool win_exp = !guard_all && max_o == (o_exp << 2) && o_exp >= 0;
        if (guard_all || win_exp) {
            if (!(t > g)) flag |= SP_HMMF_NEAR_THRESHOLD;  // next to the pmax == 1 cliff (q = 0), or NaN
            const int top = guard_all ? 101 : 93;
            if (lo >= 1 && lo <= top && C.qthr[lo] - t <= g) flag |= SP_HMMF_NEAR_THRESHOLD;
            if (lo < top && t - C.qthr[lo + 1] <= g) flag |= SP_HMMF_NEAR_THRESHOLD;
            if (!(mx2 < tie)) flag |= SP_HMMF_NEAR_TIE;  // also mx == 0 and NaN
        } else if (!(zE < tie)) {
            flag |= SP_HMMF_NEAR_TIE;  // the expected state is (nearly) as likely as the winner
        }
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
    // The planes are where the lane's last forward row left them: bit o <-> ref[Lq-1-BW+o], which is the base
    // of column k+1 for cell o of row Lq-1.  Step j of the sweep is row Lq-1-j of every lane.
    auto fetch_bwd = [&](int x0) {  // entry r <-> row x0-r: query[x0-r], ref[x0-r-BW]
#pragma unroll
        for (int rr = 0; rr < RB; rr++) {
            const int x = x0 - rr;
            const bool on = x >= 1 && x >= i_stop;
            const int y = x - BW;
            qn[rr] = on ? sp_ldg_u8_here(in.qbytes ? in.qbytes + in.q0 + x : in.qseq4 + ((in.q0 + x) >> 1)) : 0;
            rn[rr] = (on && y >= 0 && y < Lr) ? sp_ldg_u8_here(in.ref + y) : 0;
        }
    };
    const int jmax = SP_WARP_MAX(Lq - 1 - i_stop);
    fetch_bwd(Lq - 1);
    pack();
    since_check = 0;
    for (int j = 0; j <= jmax;) {
        const int i = Lq - 1 - j;  // first (highest) row of this pass
        int lane_r = RB;  // a consumed row ends the block: the MAP reads it from memory
        if (t_next + 1 <= i && t_next + 1 > i - RB && t_next + 1 >= i_stop && t_next + 1 >= 1) lane_r = i - t_next;
        const int R = SP_WARP_MIN(lane_r);
        // Cells left of column 1 or right of l_ref need no mask here: with the row above zero outside its valid
        // cells the right side stays zero by itself, and what appears left of column 1 never flows back into a valid
        // cell (dependencies only run towards smaller columns) nor into the MAP (f is zero there).  Only an instance
        // narrower than its warp's band, an N, or row 1 (no D state there) takes the masked pass.
        SpBits<NW> mmr[RB], nnr[RB];
        bool mask_lane = false;
        double m6r[RB], m8r[RB];
#pragma unroll
        for (int rr = 0; rr < RB; rr++) {
            const int x = i - rr;
            m6r[rr] = x > 1 ? m6 : 0.;
            m8r[rr] = x > 1 ? m8 : 0.;
            if (rr < R) {
                const bool live = x >= i_stop && x >= 1;
                int qc = 0;
                if (live) {
                    qc = sp_query_decode(in, x, (q4 >> (8 * rr)) & 0xff);  // query[x] (0-based) == base of row x+1
                    if (j + rr > 0) {  // one column enters the band on the left: bit 0 <-> ref[x-BW]
                        const uint32_t rc = (r4 >> (8 * rr)) & 0xff;
                        p0.shl1_in((uint64_t) (rc & 1));
                        p1.shl1_in((uint64_t) ((rc >> 1) & 1));
                        p2.shl1_in((uint64_t) ((rc >> 2) & 1));
                    }
                }
                sp_h2_row_masks(p0, p1, p2, qc, mmr[rr], nnr[rr]);
                if (live && (narrow || x <= 1 || nnr[rr].any_below(NC))) mask_lane = true;
            } else {
                mmr[rr].clear();
                nnr[rr].clear();
            }
        }
        fetch_bwd(i - R);
        const bool masked = SP_WARP_ANY(mask_lane);
        SpBits<NW> vmr[RB];
        if (masked) {
#pragma unroll
            for (int rr = 0; rr < RB; rr++) {
                int lo, hi;
                valid_range(i - rr, lo, hi);
                sp_bits_range(vmr[rr], lo, hi);
            }
        } else {
#pragma unroll
            for (int rr = 0; rr < RB; rr++) vmr[rr].clear();
        }
        // row a (0-based) computes cell NC-1-s+a at step s; it needs bM of the row above at its own cell
        // (that row's previous step) and bI of the row above one cell to the left (fresh)
        auto pass = [&](auto rtag, auto mtag) {
            constexpr int RR = decltype(rtag)::value;
            constexpr bool MK = decltype(mtag)::value != 0;
            double Mc[RR], cD[RR];
#pragma unroll
            for (int rr = 0; rr < RR; rr++) { Mc[rr] = 0.; cD[rr] = 0.; }
            double b0 = mi[(NC - 1) * STRIDE].x;  // bM of the stored row at the cell row 0 computes next
            auto step = [&](SpD2 *cell, auto ra_tag, auto rb_tag, auto bits) {  // cell = &mi[NC-1-s]: row 0's cell
                constexpr int RA = decltype(ra_tag)::value, RBB = decltype(rb_tag)::value;
                double Mup = 0., Iup = 0.;
                if constexpr (RA == 0) {
                    const SpD2 a = cell[-STRIDE];  // (cell -1 holds zeros)
                    Mup = b0;
                    Iup = a.y;
                    b0 = a.x;
                } else {
                    Mup = Mc[RA - 1];  // the row above has finished: its cell 0, bI = 0 left of the band
                }
#pragma unroll
                for (int rr = RA; rr < RBB; rr++) {
                    const uint32_t b = bits(rr);
                    double bMv, bIv;
                    if constexpr (MK) {
                        const double e = ((b & 2) ? 1. : ((b & 1) ? emA : emB)) * Mup;
                        bMv = SP_FMA(m2, cD[rr], SP_FMA(e, m0, eim1 * Iup));
                        bIv = SP_FMA(e, m3, eim4 * Iup);
                        cD[rr] = SP_FMA(m8r[rr], cD[rr], e * m6r[rr]);
                        if (!(b & 4)) { bMv = 0.; bIv = 0.; cD[rr] = 0.; }
                    } else {
                        const double e = ((b & 1) ? emA : emB) * Mup;
                        bMv = SP_FMA(m2, cD[rr], SP_FMA(e, m0, eim1 * Iup));
                        bIv = SP_FMA(e, m3, eim4 * Iup);
                        cD[rr] = SP_FMA(m8, cD[rr], e * m6);
                    }
                    Mup = Mc[rr];
                    Iup = bIv;
                    Mc[rr] = bMv;
                    if (rr == RR - 1) {
                        SpD2 v = {bMv, bIv};
                        cell[(RR - 1) * STRIDE] = v;  // cell NC-1-s+(RR-1)
                    }
                }
            };
            auto bit_rt = [&](int rr, int o) -> uint32_t {
                uint32_t v = sp_bit_rt(mmr[rr], o);
                if constexpr (MK) v |= sp_bit_rt(nnr[rr], o) << 1 | sp_bit_rt(vmr[rr], o) << 2;
                return v;
            };
            sp_static_for<0, RR - 1>([&](auto st) {
                constexpr int S = decltype(st)::value;
                step(mi + (NC - 1 - S) * STRIDE, SpInt<0>(), SpInt<S + 1>(), [&](int rr) { return bit_rt(rr, NC - 1 - S + rr); });
            });
            int s = RR - 1;
            constexpr int CK = SP_HMMF_CHUNK;
            for (; s + CK <= NC; s += CK) {
                // rows' cells at step s+J: NC-1-(s+J)+rr; the CK cells of row rr are bits (NC-CK-s+rr)..(NC-1-s+rr)
                uint32_t mb[RR], nb[MK ? RR : 1], vb[MK ? RR : 1];
#pragma unroll
                for (int rr = 0; rr < RR; rr++) {
                    mb[rr] = sp_bits8_rt(mmr[rr], NC - CK - s + rr);
                    if constexpr (MK) {
                        nb[rr] = sp_bits8_rt(nnr[rr], NC - CK - s + rr);
                        vb[rr] = sp_bits8_rt(vmr[rr], NC - CK - s + rr);
                    }
                }
                SpD2 *cell = mi + (NC - 1 - s) * STRIDE;
                sp_static_for<0, CK>([&](auto jt) {
                    constexpr int J = decltype(jt)::value;
                    step(cell - J * STRIDE, SpInt<0>(), SpInt<RR>(), [&](int rr) -> uint32_t {
                        uint32_t v = (mb[rr] >> (CK - 1 - J)) & 1;
                        if constexpr (MK) v |= ((nb[rr] >> (CK - 1 - J)) & 1) << 1 | ((vb[rr] >> (CK - 1 - J)) & 1) << 2;
                        return v;
                    });
                });
            }
            for (; s < NC; s++)
                step(mi + (NC - 1 - s) * STRIDE, SpInt<0>(), SpInt<RR>(), [&](int rr) { return bit_rt(rr, NC - 1 - s + rr); });
            sp_static_for<0, RR - 1>([&](auto dt) {
                constexpr int D = decltype(dt)::value;
                step(mi + (-1 - D) * STRIDE, SpInt<D + 1>(), SpInt<RR>(), [&](int rr) { return bit_rt(rr, -1 - D + rr); });
            });
        };
        if (!masked) {
            if (R == 1) pass(SpInt<1>(), SpInt<0>());
            else if (RB >= 2 && R == 2) pass(SpInt<(RB >= 2 ? 2 : 1)>(), SpInt<0>());
            else if (RB >= 3 && R == 3) pass(SpInt<(RB >= 3 ? 3 : 1)>(), SpInt<0>());
            else pass(SpInt<RB>(), SpInt<0>());
        } else {
            if (R == 1) pass(SpInt<1>(), SpInt<1>());
            else if (RB >= 2 && R == 2) pass(SpInt<(RB >= 2 ? 2 : 1)>(), SpInt<1>());
            else if (RB >= 3 && R == 3) pass(SpInt<(RB >= 3 ? 3 : 1)>(), SpInt<1>());
            else pass(SpInt<RB>(), SpInt<1>());
        }
        const int last = i - R + 1;
        if (last >= i_stop && last >= 1 && t_next + 1 == last) {
            map_row(nr);
            nr--;
            t_next = nr >= 0 ? rows[nr].t : -2;
        }
        j += R;
        pack();
        since_check += R;
        if (since_check >= SP_HMMF_RS) {
            range_check(last >= i_stop && last >= 1);
            since_check = 0;
        }
    }
    return flag;
}
