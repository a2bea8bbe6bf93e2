"""not gpu: pins the restated probaln_glocal (oracle/probaln_port.c) as far as it can be pinned
without htslib (absent from /root/reference and from this image -- see the "PARITY UNPINNED" note
in that file and in DESIGN.md):

  * forward/backward identities of the scaled recursion (b[0][0] == 1);
  * agreement of every row's MAP state, max posterior and integer q with an independent un-banded,
    un-scaled, extended-precision forward-backward of the same model;
  * the structural constants (k>100 -> 99, +.499 rounding, N bases emit 1).
"""
import numpy as np
import pytest

EI, EM = .25, .33333333333
LD = np.longdouble


def full_forward_backward(ref, query, set_q, d, e):
    """Un-banded, un-scaled restatement in extended precision; returns (state, pmax) per query row."""
    Lr, Lq = len(ref), len(query)
    # d and e are floats in probaln_par_t: sub-expressions made only of them are float arithmetic in C
    df, ef, one = np.float32(d), np.float32(e), np.float32(1)
    qual = LD(np.float32(10 ** (-set_q / 10.)))
    sM = sI = LD(1) / (2 * Lq + 2)
    d, e = LD(df), LD(ef)
    m0, m1, m2 = LD(one - df - df) * (1 - sM), d * (1 - sM), d * (1 - sM)
    m3, m4, m6, m8 = LD(one - ef) * (1 - sI), e * (1 - sI), LD(one - ef), e
    bM, bI = LD((one - df) / np.float32(Lr)), LD(df / np.float32(Lr))

    def E(i, k):  # 1-based row / column
        r, q = ref[k - 1], query[i - 1]
        if r > 3 or q > 3:
            return LD(1)
        return 1 - qual if r == q else qual * LD(EM)

    fM = np.zeros((Lq + 2, Lr + 2), LD); fI = fM.copy(); fD = fM.copy()
    for k in range(1, Lr + 1):
        fM[1, k] = E(1, k) * bM
        fI[1, k] = LD(EI) * bI
    for i in range(2, Lq + 1):
        for k in range(1, Lr + 1):
            fM[i, k] = E(i, k) * (m0 * fM[i - 1, k - 1] + m3 * fI[i - 1, k - 1] + m6 * fD[i - 1, k - 1])
            fI[i, k] = LD(EI) * (m1 * fM[i - 1, k] + m4 * fI[i - 1, k])
            fD[i, k] = m2 * fM[i, k - 1] + m8 * fD[i, k - 1]
    P = sum(fM[Lq, k] * sM + fI[Lq, k] * sI for k in range(1, Lr + 1))
    gM = np.zeros((Lq + 2, Lr + 2), LD); gI = gM.copy(); gD = gM.copy()
    for k in range(1, Lr + 1):
        gM[Lq, k], gI[Lq, k] = sM, sI
    for i in range(Lq - 1, 0, -1):
        y = LD(1 if i > 1 else 0)
        for k in range(Lr, 0, -1):
            em = (E(i + 1, k + 1) * gM[i + 1, k + 1]) if k < Lr else LD(0)
            gM[i, k] = em * m0 + LD(EI) * m1 * gI[i + 1, k] + m2 * gD[i, k + 1]
            gI[i, k] = em * m3 + LD(EI) * m4 * gI[i + 1, k]
            gD[i, k] = (em * m6 + m8 * gD[i, k + 1]) * y
    state = np.zeros(Lq, np.int64); pmax = np.zeros(Lq, np.float64); gap = np.zeros(Lq, np.float64)
    for i in range(1, Lq + 1):
        zs = []
        for k in range(1, Lr + 1):
            zs.append((fM[i, k] * gM[i, k], (k - 1) << 2 | 0))
            zs.append((fI[i, k] * gI[i, k], (k - 1) << 2 | 1))
        tot = sum(z for z, _ in zs)
        best = max(zs, key=lambda t: t[0])
        second = sorted((z for z, _ in zs), reverse=True)[1] if len(zs) > 1 else LD(0)
        state[i - 1] = best[1]
        pmax[i - 1] = float(best[0] / tot)
        gap[i - 1] = float((best[0] - second) / tot)
    return state, pmax, gap, float(P)


def random_pair(rng, lr):
    ref = rng.integers(0, 4, lr).astype(np.uint8)
    q = []
    for c in ref:
        u = rng.random()
        if u < 0.05:
            continue
        if u < 0.10:
            q += [int(rng.integers(0, 4)), int(c)]
        elif u < 0.18:
            q.append((int(c) + 1) % 4)
        else:
            q.append(int(c))
    if not q:
        q = [0]
    return ref, np.array(q, np.uint8)


@pytest.mark.parametrize("preset", ["hifi", "ont"])
def test_port_matches_independent_unbanded_forward_backward(oracle, preset):
    op = oracle.preset_params(preset)
    rng = np.random.default_rng(3)
    checked = 0
    for trial in range(14):
        ref, query = random_pair(rng, int(rng.integers(3, 40)))
        if trial % 5 == 0:
            ref[len(ref) // 2] = 4  # an N in the reference
        if trial % 6 == 0 and len(query) > 2:
            query[1] = 4
        iq = np.full(len(query), op.set_q, np.uint8)
        bw = max(len(ref), len(query)) + 5  # band covers the whole matrix
        o = oracle.probaln(ref, query, iq, np.float32(op.conf_d), np.float32(op.conf_e), bw, want_s=True)
        st, pm, gap, P = full_forward_backward(ref, query, op.set_q, op.conf_d, op.conf_e)
        assert abs(o["pb"] - 1.0) < 1e-10  # b[0][0] of the scaled recursion
        assert abs(np.prod(o["s"][: len(query) + 2]) / P - 1.0) < 1e-9  # likelihood = product of the scaling factors
        for i in range(len(query)):
            assert abs(o["pmax"][i] - pm[i]) < 1e-10
            if gap[i] > 1e-9:  # skip exact posterior ties
                assert o["state"][i] == st[i]
            v = -4.343 * np.log(1. - pm[i]) + .499 if pm[i] < 1 else np.inf
            if np.isfinite(v) and abs(v - round(v)) > 1e-6:
                k = int(v)
                assert o["q"][i] == (99 if k > 100 else k)
            checked += 1
    assert checked > 150


def test_band_restricts_columns_and_counts_cells(oracle):
    op = oracle.preset_params("hifi")
    rng = np.random.default_rng(9)
    ref, query = random_pair(rng, 300)
    iq = np.full(len(query), op.set_q, np.uint8)
    bw = abs(len(ref) - len(query)) + 20  # what calc_local_baq passes (ptMarker.c:754)
    o = oracle.probaln(ref, query, iq, np.float32(op.conf_d), np.float32(op.conf_e), bw)
    k = o["state"] >> 2
    i = np.arange(len(query))
    ok = o["state"] >= 0
    eff = max(min(max(len(ref), len(query)), bw), abs(len(ref) - len(query)))
    assert np.all(np.abs(k[ok] + 1 - (i[ok] + 1)) <= eff)
    cells = sum(min(len(ref), r + eff) - max(1, r - eff) + 1 for r in range(1, len(query) + 1))
    assert oracle.probaln_cells(len(ref), len(query), bw) == cells


def test_perfect_match_saturates_q(oracle):
    op = oracle.preset_params("hifi")
    rng = np.random.default_rng(1)
    ref = rng.integers(0, 4, 400).astype(np.uint8)
    query = ref[20:380].copy()
    iq = np.full(len(query), op.set_q, np.uint8)
    o = oracle.probaln(ref, query, iq, np.float32(op.conf_d), np.float32(op.conf_e), abs(len(ref) - len(query)) + 20)
    mid = slice(40, len(query) - 40)
    assert np.all((o["state"][mid] & 3) == 0)
    assert np.all((o["state"][mid] >> 2) == np.arange(len(query))[mid] + 20)
    assert o["q"].max() <= 100 and o["q"][mid].min() >= 30
