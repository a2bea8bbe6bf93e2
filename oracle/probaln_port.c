/* oracle/probaln_port.c -- TEST INFRASTRUCTURE ONLY (checker; never measured as product,
 * never linked into libsecphase_b200.so).
 *
 * CPU restatement of the banded glocal profile-HMM forward/backward + MAP that Secphase calls
 * as htslib's probaln_glocal() (reference call site: programs/submodules/ptMarker/ptMarker.c:
 * 754-757; parameter struct built at ptMarker.c:680).  The function lives in the third-party
 * dependency htslib, pinned at 1.17 by the reference's Dockerfile:16-24, which is NOT vendored
 * under /root/reference and is not installed in this image.  It is therefore restated from its
 * published algorithm (Li 2011, "Improving SNP discovery by base alignment quality", and the
 * long-stable kprobaln/probaln implementation) as specified in SURVEY.md section 8(a) row A10.
 *
 * PARITY UNPINNED against htslib itself: no htslib binary or source is available here, and the
 * reference holds no golden vectors for this function.  What IS checked (tests/test_oracle_hmm.py):
 *   - the forward/backward identities  b[0][0] == 1  and  sum_k f*b * s[i] == 1  (to 1e-12);
 *   - the structural constants .25, .33333333333, -4.343, .499, k>100 -> 99, sM=sI=1/(2Lq+2),
 *     float-typed d, e and per-base error probability;
 *   - agreement with an independent O(Lq*Lr) un-banded log-space forward-backward on small cases.
 *
 * Arithmetic contract (what the CUDA kernel must reproduce bit-for-bit): IEEE-754 binary64,
 * round-to-nearest-even, NO fused multiply-add (build with -ffp-contract=off), every
 * expression evaluated in the order written below.
 */
#include <limits.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "probaln_port.h"

#define HMM_EI .25
#define HMM_EM .33333333333

/* band column of (i,k): rows keep bw2 cells x 3 states plus one zero cell of padding each side */
static inline int band_u(int bw, int i, int k) {
    int x = i - bw;
    if (x < 0) x = 0;
    return (k - x + 1) * 3;
}

static float g_q2p[256];
static int g_q2p_ready = 0;

static inline double emission(const float *qual, const uint8_t *ref, const uint8_t *query, int i0, int k0) {
    /* i0,k0 are 0-based indices into query/ref */
    if (ref[k0] > 3 || query[i0] > 3) return 1.;
    return ref[k0] == query[i0] ? 1. - qual[i0] : qual[i0] * HMM_EM;
}

long oracle_probaln_cells(int l_ref, int l_query, int c_bw) {
    int bw = l_ref > l_query ? l_ref : l_query;
    if (bw > c_bw) bw = c_bw;
    if (bw < abs(l_ref - l_query)) bw = abs(l_ref - l_query);
    long cells = 0;
    for (int i = 1; i <= l_query; i++) {
        int beg = i - bw > 1 ? i - bw : 1;
        int end = i + bw < l_ref ? i + bw : l_ref;
        if (end >= beg) cells += end - beg + 1;
    }
    return cells;
}

int oracle_probaln_glocal_ex(const uint8_t *ref, int l_ref, const uint8_t *query, int l_query,
                             const uint8_t *iqual, float par_d, float par_e, int par_bw,
                             int *state, uint8_t *q, double *s_out, double *pmax_out, double *pb_out) {
    if (l_ref <= 0 || l_query <= 0) return 0;
    if (!g_q2p_ready) {
        for (int i = 0; i < 256; i++) g_q2p[i] = (float) pow(10, -i / 10.);
        g_q2p_ready = 1;
    }
    int bw = l_ref > l_query ? l_ref : l_query;
    if (bw > par_bw) bw = par_bw;
    if (bw < abs(l_ref - l_query)) bw = abs(l_ref - l_query);
    const int bw2 = bw * 2 + 1;
    const size_t stride = (size_t) (bw2 < l_ref ? bw2 : l_ref) * 3 + 6;

    /* +8: row l_query's right neighbour of the last cell is read (and multiplied by 0) when l_ref < bw2 */
    double *f = (double *) calloc((size_t) (l_query + 1) * stride + 8, sizeof(double));
    double *b = (double *) calloc((size_t) (l_query + 1) * stride + 8, sizeof(double));
    double *s = (double *) calloc((size_t) l_query + 2, sizeof(double));
    float *qual = (float *) calloc((size_t) l_query, sizeof(float));
    if (!f || !b || !s || !qual) {
        free(f); free(b); free(s); free(qual);
        return INT_MIN;
    }
    for (int i = 0; i < l_query; i++) qual[i] = g_q2p[iqual ? iqual[i] : 30];

    /* transition matrix; d and e are floats widened to double */
    const double sM = 1. / (2 * l_query + 2), sI = sM;
    double m[9];
    m[0] = (1 - par_d - par_d) * (1 - sM);
    m[1] = m[2] = par_d * (1 - sM);
    m[3] = (1 - par_e) * (1 - sI);
    m[4] = par_e * (1 - sI);
    m[5] = 0.;
    m[6] = 1 - par_e;
    m[7] = 0.;
    m[8] = par_e;
    const double bM = (1 - par_d) / l_ref, bI = par_d / l_ref;

    /* ---- forward ---- */
    f[band_u(bw, 0, 0)] = 1.;
    s[0] = 1.;
    {
        double *fi = f + stride;
        int end = l_ref < bw + 1 ? l_ref : bw + 1;
        double sum = 0.;
        for (int k = 1; k <= end; k++) {
            int u = band_u(bw, 1, k);
            fi[u] = emission(qual, ref, query, 0, k - 1) * bM;
            fi[u + 1] = HMM_EI * bI;
            sum += fi[u] + fi[u + 1];
        }
        s[1] = sum;
        int lo = band_u(bw, 1, 1), hi = band_u(bw, 1, end) + 2;
        for (int u = lo; u <= hi; u++) fi[u] /= sum;
    }
    for (int i = 2; i <= l_query; i++) {
        double *fi = f + (size_t) i * stride;
        const double *fp = f + (size_t) (i - 1) * stride;
        int beg = i - bw > 1 ? i - bw : 1;
        int end = i + bw < l_ref ? i + bw : l_ref;
        double sum = 0.;
        for (int k = beg; k <= end; k++) {
            double e = emission(qual, ref, query, i - 1, k - 1);
            int u = band_u(bw, i, k);
            int v11 = band_u(bw, i - 1, k - 1), v10 = band_u(bw, i - 1, k), v01 = band_u(bw, i, k - 1);
            fi[u] = e * (m[0] * fp[v11] + m[3] * fp[v11 + 1] + m[6] * fp[v11 + 2]);
            fi[u + 1] = HMM_EI * (m[1] * fp[v10] + m[4] * fp[v10 + 1]);
            fi[u + 2] = m[2] * fi[v01] + m[8] * fi[v01 + 2];
            sum += fi[u] + fi[u + 1] + fi[u + 2];
        }
        s[i] = sum;
        double r = 1. / sum;
        int lo = band_u(bw, i, beg), hi = band_u(bw, i, end) + 2;
        for (int u = lo; u <= hi; u++) fi[u] *= r;
    }
    {
        const double *fl = f + (size_t) l_query * stride;
        double sum = 0.;
        for (int k = 1; k <= l_ref; k++) {
            int u = band_u(bw, l_query, k);
            if (u < 3 || u >= bw2 * 3 + 3) continue;
            sum += fl[u] * sM + fl[u + 1] * sI;
        }
        s[l_query + 1] = sum;
    }
    int Pr;
    {
        double p = 1., acc = 0.;
        for (int i = 0; i <= l_query + 1; i++) {
            p *= s[i];
            if (p < 1e-100) acc += -4.343 * log(p), p = 1.;
        }
        acc += -4.343 * log(p * l_ref * l_query);
        Pr = (int) (acc + .499);
    }

    /* ---- backward ---- */
    {
        double *bl = b + (size_t) l_query * stride;
        for (int k = 1; k <= l_ref; k++) {
            int u = band_u(bw, l_query, k);
            if (u < 3 || u >= bw2 * 3 + 3) continue;
            bl[u] = sM / s[l_query] / s[l_query + 1];
            bl[u + 1] = sI / s[l_query] / s[l_query + 1];
        }
    }
    for (int i = l_query - 1; i >= 1; i--) {
        double *bi = b + (size_t) i * stride;
        const double *bn = b + (size_t) (i + 1) * stride;
        int beg = i - bw > 1 ? i - bw : 1;
        int end = i + bw < l_ref ? i + bw : l_ref;
        double y = (i > 1);
        for (int k = end; k >= beg; k--) {
            int u = band_u(bw, i, k);
            int v11 = band_u(bw, i + 1, k + 1), v10 = band_u(bw, i + 1, k), v01 = band_u(bw, i, k + 1);
            double e = (k >= l_ref ? 0 : emission(qual, ref, query, i, k)) * bn[v11];
            bi[u] = e * m[0] + HMM_EI * m[1] * bn[v10 + 1] + m[2] * bi[v01 + 2];
            bi[u + 1] = e * m[3] + HMM_EI * m[4] * bn[v10 + 1];
            bi[u + 2] = (e * m[6] + m[8] * bi[v01 + 2]) * y;
        }
        double r = 1. / s[i];
        int lo = band_u(bw, i, beg), hi = band_u(bw, i, end) + 2;
        for (int u = lo; u <= hi; u++) bi[u] *= r;
    }
    {
        int end = l_ref < bw + 1 ? l_ref : bw + 1;
        double sum = 0.;
        const double *b1 = b + stride;
        for (int k = end; k >= 1; k--) {
            int u = band_u(bw, 1, k);
            double e = emission(qual, ref, query, 0, k - 1);
            if (u < 3 || u >= bw2 * 3 + 3) continue;
            sum += e * b1[u] * bM + HMM_EI * b1[u + 1] * bI;
        }
        b[band_u(bw, 0, 0)] = sum / s[0];
        if (pb_out) *pb_out = sum / s[0];
    }

    /* ---- MAP ---- */
    for (int i = 1; i <= l_query; i++) {
        const double *fi = f + (size_t) i * stride, *bi = b + (size_t) i * stride;
        int beg = i - bw > 1 ? i - bw : 1;
        int end = i + bw < l_ref ? i + bw : l_ref;
        double sum = 0., max = 0.;
        int max_k = -1;
        for (int k = beg; k <= end; k++) {
            int u = band_u(bw, i, k);
            double z = fi[u] * bi[u];
            if (z > max) max = z, max_k = (k - 1) << 2 | 0;
            sum += z;
            z = fi[u + 1] * bi[u + 1];
            if (z > max) max = z, max_k = (k - 1) << 2 | 1;
            sum += z;
        }
        max /= sum;
        sum *= s[i]; /* == 1 up to rounding; kept for the identity test */
        if (pmax_out) pmax_out[i - 1] = max;
        if (state) state[i - 1] = max_k;
        if (q) {
            /* (int) of +inf / NaN is undefined in ISO C; on the reference's x86-64 build
             * cvttsd2si returns INT_MIN, which then stores 0 into the uint8_t.  Make that
             * explicit so the oracle does not depend on UB. */
            double v = -4.343 * log(1. - max) + .499;
            int kq = (v >= 2147483648. || v != v) ? INT_MIN : (int) v;
            q[i - 1] = (uint8_t) (kq > 100 ? 99 : kq);
        }
    }
    if (s_out) memcpy(s_out, s, ((size_t) l_query + 2) * sizeof(double));
    free(f); free(b); free(s); free(qual);
    return Pr;
}

/* ---- call trace (lets the _ref driver see every HMM instance calc_local_baq launches) ---- */
static __thread oracle_hmm_trace_fn g_trace_fn = 0;
static __thread void *g_trace_ud = 0;
void oracle_probaln_set_trace(oracle_hmm_trace_fn fn, void *ud) {
    g_trace_fn = fn;
    g_trace_ud = ud;
}

#ifndef ORACLE_NO_HTSLIB_SYMBOL
/* The exact symbol/signature the reference links against (htslib/sam.h). */
typedef struct {
    float d, e;
    int bw;
} probaln_par_t;
int probaln_glocal(const uint8_t *ref, int l_ref, const uint8_t *query, int l_query,
                   const uint8_t *iqual, const probaln_par_t *c, int *state, uint8_t *q) {
    int r = oracle_probaln_glocal_ex(ref, l_ref, query, l_query, iqual, c->d, c->e, c->bw, state, q, 0, 0, 0);
    if (g_trace_fn) g_trace_fn(g_trace_ud, ref, l_ref, query, l_query, iqual, c->d, c->e, c->bw, state, q);
    return r;
}
#endif
