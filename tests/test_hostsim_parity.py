"""not gpu: the kernels' stage logic (secphase_b200/csrc/sp_*.cuh compiled for the host by
tests/hostsim) against the CPU oracle, bit for bit, on seeded synthetic read groups."""
import numpy as np
import pytest

from tests.conftest import CASES, make_case, oracle_refseq
from tools.parity import compare_results


@pytest.fixture(scope="module")
def hostsim():
    from tests.hostsim import pyhostsim
    pyhostsim.lib()
    return pyhostsim


@pytest.mark.parametrize("name,spreset,ppreset,ng,over", CASES, ids=[c[0] for c in CASES])
def test_stage_logic_bit_exact_vs_oracle(hostsim, oracle, name, spreset, ppreset, ng, over):
    s, b, codes, off = make_case(spreset, ng, **over)
    op = oracle.preset_params(ppreset)
    exp = oracle.run(b, op, oracle_refseq(oracle, s))
    got = hostsim.run(b, hostsim.params_from_oracle(op), codes, off)
    assert got["err"] == 0
    bad = compare_results(exp, got, label="hostsim")
    assert not bad, "\n".join(bad)
    assert len(got["items"]) == len(exp["hmm"])


@pytest.mark.parametrize("name,spreset,ppreset,ng,over", CASES, ids=[c[0] for c in CASES])
def test_write_qual_mode_bit_exact_vs_oracle(hostsim, oracle, name, spreset, ppreset, ng, over):
    """-w/--writeBam mode (SpConst::full_baq): every record's quality array as calc_update_baq_all leaves
    it (ptMarker.c:786, 709-720, 797-806) equals the reference's records byte for byte, and nothing else
    the job returns changes with the mode."""
    s, b, codes, off = make_case(spreset, ng, **over)
    op = oracle.preset_params(ppreset)
    exp = oracle.run(b, op, oracle_refseq(oracle, s))
    got = hostsim.run(b, hostsim.params_from_oracle(op), codes, off, full_baq=True)
    assert got["err"] == 0
    bad = compare_results(exp, got, label="hostsim-full")
    assert not bad, "\n".join(bad)
    assert exp["qual"].shape == got["qual"].shape == b.qual_pool.shape
    diff = np.flatnonzero(exp["qual"] != got["qual"])
    assert diff.size == 0, f"{diff.size} quality bytes differ, first at {diff[:5]}: oracle={exp['qual'][diff[:5]]} got={got['qual'][diff[:5]]}"
    # the mode really rewrites qualities where an HMM ran
    if len(exp["hmm"]):
        assert (exp["qual"] != b.qual_pool).any()
    # every marker's post-BAQ quality is the byte of its record (calc_update_baq_all, ptMarker.c:823-830)
    mk, mo = got["markers_baq"], got["markers_baq_off"]
    for g in range(b.n_groups):
        a0 = int(b.grp_aln_off[g])
        rows = mk[int(mo[g]):int(mo[g + 1])]
        if got["groups"][g, 9] and len(rows):
            pos = b.qual_off[a0 + rows[:, 0]] + rows[:, 2]
            assert np.array_equal(got["qual"][pos].astype(np.int32), rows[:, 3])


@pytest.mark.parametrize("preset,over", [("hifi", {}), ("ont", {}), ("ont", dict(use_md=1)), ("stress", {})])
def test_plan_independent_of_scan_threads(hostsim, preset, over):
    """The launcher scans text-heavy batches (ONT) with several threads: same plan as with one."""
    s, b, _, _ = make_case(preset, 37, locus_len=300000, **over)
    thr = 20 if preset == "ont" else 10
    rc1, h1 = hostsim.plan_signature(b, thr, 1)
    assert rc1 == 0
    for nt in (2, 3, 8, 64):
        assert hostsim.plan_signature(b, thr, nt) == (0, h1)
    assert hostsim.plan_signature(b, thr, 4, safe_caps=True)[0] == 0
    assert hostsim.plan_signature(b.group_slice(0, 0), thr, 8)[0] == 0
    # an unsupported CIGAR op is reported the same way from any thread
    bad = b.group_slice(0, 12)
    bad.cigar_pool = bad.cigar_pool.copy()
    bad.cigar_pool[int(bad.cigar_off[5])] = (7 << 4) | 3   # 7N
    assert hostsim.plan_signature(bad, thr, 1)[0] == hostsim.plan_signature(bad, thr, 8)[0] == -6   # SP_EUNSUPPORTED


def test_glibc_rand_emulation(hostsim):
    import ctypes
    libc = ctypes.CDLL("libc.so.6")
    L = hostsim.lib()
    for seed in (1, 7, 123456):
        libc.srand(seed)
        r = L.hs_rng_create(seed)
        for _ in range(2000):
            assert L.hs_rng_next(r) == libc.rand()
        L.hs_rng_destroy(r)


def _random_instances(rng, n_trials):
    for trial in range(n_trials):
        lr = int(rng.integers(1, 120)) if trial % 3 == 0 else int(rng.integers(100, 1001))
        ref = rng.integers(0, 4, lr).astype(np.uint8)
        q = []
        for c in ref:
            u = rng.random()
            if u < 0.02:
                continue
            if u < 0.04:
                q += [int(rng.integers(0, 4)), int(c)]
            elif u < 0.06:
                q.append((int(c) + 1) % 4)
            else:
                q.append(int(c))
        if trial % 7 == 0 and len(q) > 5:
            q[3] = 4
        if trial % 11 == 0:
            ref[min(5, lr - 1)] = 4
        if not q:
            q = [0]
        query = np.array(q, np.uint8)
        bw = abs(lr - len(query)) + 20 if trial % 5 else int(rng.integers(1, 70))
        yield ref, query, bw


@pytest.mark.parametrize("preset", ["hifi", "ont"])
def test_hmm_kernel_bodies_all_rows_bit_exact_vs_port(hostsim, oracle, preset):
    """Both K4 bodies (sp_hmm.cuh generic, sp_hmm2.cuh shared-memory band) against the restated
    probaln_glocal: state, q and the bits of the normalised max posterior at EVERY row."""
    rng = np.random.default_rng(5)
    op = oracle.preset_params(preset)
    hp = hostsim.params_from_oracle(op)
    n2 = 0
    for ref, query, bw in _random_instances(rng, 120):
        rows = np.arange(len(query), dtype=np.int32)
        iq = np.full(len(query), op.set_q, np.uint8)
        o = oracle.probaln(ref, query, iq, np.float32(op.conf_d), np.float32(op.conf_e), bw)
        for got in (hostsim.hmm(hp, ref, query, bw, rows), hostsim.hmm2(hp, ref, query, bw, rows),
                    hostsim.hmm2(hp, ref, query, bw, rows, unrolled=False)):
            if got is None:
                continue
            n2 += 1
            assert np.array_equal(o["state"], got["state"])
            assert np.array_equal(o["q"], got["q"])
            assert np.array_equal(o["pmax"].view(np.int64), got["pmax"].view(np.int64))
        # sparse marker rows (the pipeline's use): same answers as the all-rows run
        if len(query) > 30:
            sel = np.sort(rng.choice(len(query), size=3, replace=False)).astype(np.int32)
            g2 = hostsim.hmm2(hp, ref, query, bw, sel)
            if g2 is not None:
                assert np.array_equal(o["state"][sel], g2["state"])
                assert np.array_equal(o["q"][sel], g2["q"])
    assert n2 > 250


def _fast_rows_check(o, got, rows):
    """Per consumed row: is it outside every guard band (recomputed here from the oracle's pmax)?  Returns the
    rows on which the fast kernel's integers may be used without a strict re-run."""
    t = 1.0 - o["pmax"][rows]
    return t


@pytest.mark.parametrize("preset", ["hifi", "ont"])
def test_hmm_fast_body_guard_band(hostsim, oracle, preset):
    """sp_hmmf.cuh (FMA, no per-row sum, power-of-two rescaling, virtual sliding band) against the restated
    probaln_glocal.  Contract of the guard band: an instance that comes back UNFLAGGED has the reference's
    state and q at every consumed row; 1 - pmax agrees within 16 units of 2^-53 + 1e-10 relative (the band
    itself is 64 units + 1e-9).  Instances with random bands, N bases, short references, sparse and dense
    rows; the flagged fraction stays small on sparse (marker-like) rows."""
    rng = np.random.default_rng(17)
    op = oracle.preset_params(preset)
    hp = hostsim.params_from_oracle(op)
    n_fast = n_flag_dense = n_sparse = n_flag_sparse = 0
    worst_ulp = worst_rel = 0.0
    for ref, query, bw in _random_instances(rng, 160):
        iq = np.full(len(query), op.set_q, np.uint8)
        o = oracle.probaln(ref, query, iq, np.float32(op.conf_d), np.float32(op.conf_e), bw)
        dense = np.arange(len(query), dtype=np.int32)
        picks = [dense]
        if len(query) > 40:
            picks += [np.sort(rng.choice(np.arange(10, len(query) - 11), size=3, replace=False)).astype(np.int32)
                      for _ in range(3)]
        for k, rows in enumerate(picks):
            got = hostsim.hmmf(hp, ref, query, bw, rows, extra_bw=(0, 3, 0, 7)[k % 4])
            if got is None:   # band class without a fast body: the launcher uses the strict kernel
                continue
            n_fast += 1
            sparse = len(rows) == 3
            n_sparse += sparse
            if got["flags"]:
                n_flag_sparse += sparse
                n_flag_dense += not sparse
                continue
            assert np.array_equal(o["state"][rows], got["state"]), (len(ref), len(query), bw)
            assert np.array_equal(o["q"][rows], got["q"]), (len(ref), len(query), bw)
            ts, tf = 1.0 - o["pmax"][rows], 1.0 - got["pmax"]
            ad = np.abs(ts - tf)
            assert np.all(ad <= 16 * 2.0 ** -53 + 1e-10 * ts)
            worst_ulp = max(worst_ulp, float(ad.max() * 2.0 ** 53))
            big = ts >= 1e-4
            if big.any():
                worst_rel = max(worst_rel, float((ad[big] / ts[big]).max()))
    assert n_fast > 150 and n_sparse > 100
    assert n_flag_sparse <= 0.05 * n_sparse, (n_flag_sparse, n_sparse)
    print(f"fast instances {n_fast}, flagged dense {n_flag_dense}, flagged sparse {n_flag_sparse}/{n_sparse}, "
          f"worst |dt| {worst_ulp:.1f} x 2^-53, worst relative {worst_rel:.2e}")


@pytest.mark.parametrize("name,spreset,ppreset,ng,over", CASES, ids=[c[0] for c in CASES])
def test_fast_hmm_pipeline_bit_exact_vs_oracle(hostsim, oracle, name, spreset, ppreset, ng, over):
    """The product's default K4 arrangement -- fast kernel, guard band, strict re-run of the flagged instances --
    gives the oracle's tables; in cross-check mode (every fast instance also run strictly) no integer
    difference slipped through the band and the drift stays 10x inside it."""
    s, b, codes, off = make_case(spreset, ng, **over)
    ref = oracle_refseq(oracle, s)
    op = oracle.preset_params(ppreset)
    exp = oracle.run(b, op, ref, keep_hmm=False)
    hp = hostsim.params_from_oracle(op)
    got = hostsim.run(b, hp, codes, off, hmm_mode=1)
    assert got["err"] == 0
    bad = compare_results(exp, got, label="hostsim-fast")
    assert not bad, "\n".join(bad)
    x = hostsim.run(b, hp, codes, off, hmm_mode=2)
    assert x["err"] == 0, "an unflagged fast instance differs from the strict kernel"
    f = x["fast"]
    assert f["instances"] > 0.5 * len(x["items"])
    assert f["rerun"] <= 0.01 * f["instances"] + 2
    assert f["max_abs_drift_ulp"] <= 16 and f["max_rel_drift"] <= 1e-10


@pytest.mark.parametrize("name,spreset,ppreset,ng,over", CASES, ids=[c[0] for c in CASES])
def test_warp_walker_equals_serial_walker(hostsim, oracle, name, spreset, ppreset, ng, over):
    """sp_walk_warp.cuh (the 32 lanes emulated by tests/hostsim/warp_emu.h) writes the tables of sp_walk.cuh --
    ops, extents, initial markers, confident blocks -- for every alignment it accepts, and accepts everything the
    generator writes with cs tags."""
    s, b, codes, off = make_case(spreset, min(ng, 12), **over)
    r = hostsim.walk_warp_check(b, hostsim.params_from_oracle(oracle.preset_params(ppreset)))
    assert r["different"] == 0, r
    assert r["handled"] == (0 if over.get("use_md") else r["alignments"]), r


def test_warp_walker_declines_or_agrees_on_damaged_input(hostsim, oracle):
    """Damaged cs text and CIGARs (substituted and swapped bytes, changed lengths and ops, zero-length ops): the
    warp walker either hands the alignment to the serial walker or writes exactly the serial walker's tables."""
    rng = np.random.default_rng(11)
    alpha = np.frombuffer(b":*+-acgtn0123456789Z^ ", dtype=np.uint8)
    handled = total = 0
    for spreset, ppreset, over in (("hifi", "hifi", dict(locus_len=300000)),
                                   ("hifi", "hifi", dict(locus_len=300000, eqx=1, clip_prob=0.8)),
                                   ("ont", "ont", dict(locus_len=300000, len_mean=3000, len_sd=1000, len_min=800, clip_prob=0.9,
                                                       hard_clip_prob=0.9))):
        s, b0, codes, off = make_case(spreset, 12, **over)
        P = hostsim.params_from_oracle(oracle.preset_params(ppreset))
        for it in range(12):
            b = b0.group_slice(0, b0.n_groups)
            b.tag_pool = b.tag_pool.copy()
            b.cigar_pool = b.cigar_pool.copy()
            for a in range(len(b.n_cigar)):
                mode = int(rng.integers(0, 5))
                t0, t1 = int(b.tag_off[a]), int(b.tag_off[a + 1])
                c0, nc = int(b.cigar_off[a]), int(b.n_cigar[a])
                if mode == 0 and t1 > t0:
                    for _ in range(int(rng.integers(1, 4))):
                        b.tag_pool[rng.integers(t0, t1)] = alpha[rng.integers(0, len(alpha))]
                elif mode == 1:
                    k = c0 + int(rng.integers(0, nc))
                    v = int(b.cigar_pool[k])
                    b.cigar_pool[k] = (max(0, (v >> 4) + int(rng.integers(-2, 3))) << 4) | (v & 15)
                elif mode == 2:
                    k = c0 + int(rng.integers(0, nc))
                    b.cigar_pool[k] = (int(b.cigar_pool[k]) & ~15) | int(rng.choice([0, 1, 2, 4, 5, 7, 8]))
                elif mode == 3 and t1 - t0 > 4:
                    i, j = rng.integers(t0, t1, 2)
                    b.tag_pool[i], b.tag_pool[j] = b.tag_pool[j], b.tag_pool[i]
            r = hostsim.walk_warp_check(b, P)
            assert r["different"] == 0, (spreset, it, r)
            handled += r["handled"]
            total += r["alignments"]
    assert 0 < handled < total


@pytest.mark.parametrize("name,spreset,ppreset,ng,over", CASES, ids=[c[0] for c in CASES])
def test_lane_per_alignment_group_stages_bit_exact_vs_oracle(hostsim, oracle, name, spreset, ppreset, ng, over):
    """K2 + K3 in the form k_group_lanes<2|4|16> runs them (sp_group_warp.cuh: a lane per alignment, the group's
    lanes voting and handing lists to each other), on the emulated warp: every table of the job equals the
    reference's, as with the thread-per-group form."""
    s, b, codes, off = make_case(spreset, ng, **over)
    op = oracle.preset_params(ppreset)
    exp = oracle.run(b, op, oracle_refseq(oracle, s))
    got = hostsim.run(b, hostsim.params_from_oracle(op), codes, off, group_lanes=True)
    assert got["err"] == 0
    bad = compare_results(exp, got, label="hostsim-lanes")
    assert not bad, "\n".join(bad)
    # and with the provable (larger) block workspaces of the capacity retry
    got2 = hostsim.run(b, hostsim.params_from_oracle(op), codes, off, group_lanes=True, safe_caps=True)
    assert got2["err"] == 0 and not compare_results(exp, got2, label="hostsim-lanes-safe")


def test_warp_emulator_selftest(hostsim):
    """tests/hostsim/warp_emu.h returns what the CUDA warp intrinsics are defined to return (shuffles, votes,
    reductions, the repo's scans on top of them), with lanes doing unequal private work between rendezvous."""
    assert hostsim.lib().hs_warp_emu_selftest() == 0


def _handmade_batch(alns):
    """One read group per alignment from (flag, pos, CIGAR string, cs string, l_qseq)."""
    import re
    from secphase_b200.flatbatch import FlatBatch
    ops = "MIDNSHP=X"
    cig, cig_off, tag, tag_off, seq_off, qual_off = [], [0], bytearray(), [0], [0], [0]
    for flag, pos, cigar, cs, lq in alns:
        for n, o in re.findall(r"(\d+)([MIDNSHP=X])", cigar):
            cig.append((int(n) << 4) | ops.index(o))
        cig_off.append(len(cig))
        tag += cs.encode()
        tag_off.append(len(tag))
        seq_off.append(seq_off[-1] + (lq + 1) // 2)
        qual_off.append(qual_off[-1] + lq)
    n = len(alns)
    rng = np.random.default_rng(3)
    return FlatBatch(
        grp_aln_off=np.arange(n + 1), qname_off=np.arange(n + 1) * 2, qname_pool=np.frombuffer(b"r\0" * n, np.uint8),
        flag=[a[0] for a in alns], tid=np.zeros(n), pos=[a[1] for a in alns], l_qseq=[a[4] for a in alns],
        n_cigar=np.diff(cig_off), tag_kind=np.zeros(n), cigar_off=cig_off, tag_off=tag_off, seq_off=seq_off,
        qual_off=qual_off, cigar_pool=cig, tag_pool=np.frombuffer(bytes(tag), np.uint8),
        seq_pool=np.zeros(seq_off[-1], np.uint8), qual_pool=rng.integers(0, 60, qual_off[-1]).astype(np.uint8))


def test_warp_walker_handmade_alignments(hostsim, oracle):
    """Hand-written CIGAR/cs pairs: every shape of the cs grammar at every byte alignment of the text (the warp walker
    reads aligned words), tokens that span lanes and 128-byte iterations, both strands -- handled and identical to the
    serial walker; text outside the grammar or disagreeing with the CIGAR -- declined (or, where the serial walker
    gives the same answer anyway, identical)."""
    ins300 = "acgt" * 75
    good = [
        ("5M", ":5", 5), ("3S5M2S", ":5", 10), ("2H3S4M1I3M2D2M4S1H", ":4+a:3-cg:2", 17),
        ("10M", ":3*ag:2*ct*ga:2", 10), ("3=1X2=2X2=", ":3*ag:2*ct*ga:2", 10), ("1M", "*ag", 1), ("12M", ":12", 12),
        ("5M300I5M", ":5+" + ins300 + ":5", 310), ("5=300I5=", ":5+" + ins300 + ":5", 310),
        ("60M", ":5" + "*ag" * 50 + ":5", 60), ("5=50X5=", ":5" + "*ag" * 50 + ":5", 60),
        ("1234567M", ":1234567", 1234567), ("4M1D4M", ":4-a:4", 8), ("1X1I1X", "*ag+c*tc", 3),
        ("2M1I1M1D3M", "*ag:1+t*ca-g:2*gt", 7), ("100M", ":9" * 10 + ":10", 100),
        ("7H40M7H", ":13*ag:13*ct:12", 40),
    ]
    # text padding so that consecutive tags start at all four byte alignments
    alns = []
    for k, (cigar, cs, lq) in enumerate(good * 4):
        alns.append((16 if (k % 2) else 0, 1000 + k, cigar, cs, lq))
        if k % len(good) == len(good) - 1:
            alns.append((0, 5, "1M", ":1", 1))  # (2 bytes: shifts the alignment of everything after it)
            alns.append((0, 5, "1M", "*ag", 1) if (k // len(good)) % 2 else (0, 5, "1M", ":1", 1))
    b = _handmade_batch(alns)
    assert len(set(int(x) % 4 for x in b.tag_off[:-1])) == 4
    P = hostsim.params_from_oracle(oracle.preset_params("hifi"))
    r = hostsim.walk_warp_check(b, P)
    assert r["different"] == 0 and r["handled"] == r["alignments"] == len(alns), r
    bad = [
        ("5M", ":5:", 5), ("5M", ":5*a", 5), ("5M", ":4*agc", 5), ("5M", ":5+", 5), ("5M", "5", 5), ("5M", ":2 :3", 5),
        ("5M", ":0:5", 5), ("5M", "::5", 5), ("5M", ":5-", 5), ("5M", ":2*a*g:2", 5), ("5M", ":2+:3", 5),
        ("123456789M", ":123456789", 10), ("5M", ":4", 5), ("5M", ":6", 5), ("4M1I", ":5", 5), ("2M1I2M", ":2-a:2", 5),
        ("2M1I2M", ":2+ac:2", 6), ("3=2X", ":5", 5), ("3M2=", ":5", 5), ("2M3S2M", ":2:2", 7), ("5M", "Z5", 5),
        ("3M0I2M", ":5", 5), ("5M", ":5\t", 5),
    ]
    b2 = _handmade_batch([(16 if (k % 2) else 0, 77, cigar, cs, lq) for k, (cigar, cs, lq) in enumerate(bad * 4)])
    r2 = hostsim.walk_warp_check(b2, P)
    assert r2["different"] == 0, r2
    assert r2["handled"] == 0, r2   # every one of them goes to the serial walker


@pytest.mark.parametrize("seed", [7919, 15838, 23757])
def test_warp_cooperative_stages_other_seeds(hostsim, oracle, seed):
    """The warp walker and the lane-per-alignment group/score stages on read groups from other seeds than the fixed
    cases (HiFi, ONT, many secondaries): tables equal to the reference's, every cs alignment handled by the warp walker."""
    for spreset, ppreset, ng, over in (("hifi", "hifi", 24, dict(locus_len=300000)), ("ont", "ont", 8, dict(locus_len=400000)),
                                       ("stress", "hifi", 8, dict(locus_len=300000))):
        s, b, codes, off = make_case(spreset, ng, seed=20240603 + seed, **over)
        op = oracle.preset_params(ppreset)
        P = hostsim.params_from_oracle(op)
        exp = oracle.run(b, op, oracle_refseq(oracle, s))
        got = hostsim.run(b, P, codes, off, group_lanes=True, hmm_mode=1)
        assert got["err"] == 0 and not compare_results(exp, got, label=f"lanes-{spreset}-{seed}")
        r = hostsim.walk_warp_check(b, P)
        assert r["different"] == 0 and r["handled"] == r["alignments"], r
