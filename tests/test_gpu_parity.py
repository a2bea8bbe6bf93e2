"""-m gpu: the CUDA path, called through the C ABI, against the CPU oracle and committed goldens."""
import numpy as np
import pytest

from tests.conftest import CASES, count_top_score_ties, make_case, oracle_refseq
from tools.parity import compare_results

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sp():
    import secphase_b200
    return secphase_b200


@pytest.mark.parametrize("name,spreset,ppreset,ng,over", CASES, ids=[c[0] for c in CASES])
def test_pipeline_bit_exact_vs_oracle(sp, oracle, name, spreset, ppreset, ng, over):
    s, b, codes, off = make_case(spreset, ng, **over)
    ref = oracle_refseq(oracle, s)
    exp = oracle.run(b, oracle.preset_params(ppreset), ref, keep_hmm=False)
    with sp.Secphase(ppreset) as eng:
        eng.set_reference_codes(codes, off)
        got = eng.run_debug(b)          # default: fast HMM arithmetic + guard band + strict re-run
        assert got["hmm_mode"] == "fast"
        eng.set_hmm_mode("strict")
        eng.rng_seed(1)
        got_s = eng.run_debug(b)
        assert got_s["hmm_mode"] == "strict" and got_s["hmm_rerun"] == 0
    for g, label in ((got, "cuda-fast"), (got_s, "cuda-strict")):
        bad = compare_results(exp, g, label=label)
        assert not bad, "\n".join(bad)
        assert g["hmm_instances"] == len(exp["hmm"])
        cells = int(exp["hmm"][:, 4].astype(np.int64).sum() + (exp["hmm"][:, 5].astype(np.int64) << 31).sum())
        assert g["hmm_cells"] == cells
        assert g["gpu_launches"] >= 5
    # (what is consumed of a marker row -- 0 unless the MAP state is M at the alignment's own column, else
    # min(q, 93) -- is the BAQ compared above in markers_baq; the guard band protects exactly that)
    assert np.array_equal(got["rows"][:, :2], got_s["rows"][:, :2])
    assert got["hmm_rerun"] <= 0.01 * got["hmm_instances"] + 2


def test_top_score_ties_follow_the_reference_rand_stream(sp, oracle):
    """Several secondaries with the same top score: the reference picks rand() % count of them
    (ptAlignment.c:156-170) on a never-seeded glibc stream.  The case must really contain ties, the
    selection must equal the reference's own get_best_record_index in group order, and a group's
    selection must not depend on how the batch was cut (the stream position is carried over)."""
    name, spreset, ppreset, ng, over = [c for c in CASES if c[0] == "stress_ties"][0]
    s, b, codes, off = make_case(spreset, ng, **over)
    ref = oracle_refseq(oracle, s)
    op = oracle.preset_params(ppreset)
    exp = oracle.run(b, op, ref)
    n_ties = count_top_score_ties(b, exp["scores"])
    assert n_ties >= 1, "the tie case holds no tie"
    with sp.Secphase(ppreset) as eng:
        eng.set_reference_codes(codes, off)
        got = eng.run(b)
    assert np.array_equal(got["scores"].view(np.int64), exp["scores"].view(np.int64))
    assert np.array_equal(got["groups"][:, 0], exp["groups"][:, 0])
    assert np.array_equal(got["groups"][:, 0], oracle.select(b.grp_aln_off, b.flag, exp["scores"], op))
    tied_swaps = int((got["groups"][:, 0] != got["groups"][:, 1]).sum())
    assert tied_swaps >= 1
    with sp.Secphase(ppreset) as eng:  # two batches on two slots: one rand() stream in submission order
        eng.set_reference_codes(codes, off)
        h = ng // 2
        eng.submit(b.group_slice(0, h), slot=0)
        eng.submit(b.group_slice(h, ng), slot=1)
        parts = [eng.wait(0), eng.wait(1)]
    assert np.array_equal(np.concatenate([p["groups"][:, 0] for p in parts]), exp["groups"][:, 0])


@pytest.mark.parametrize("name,spreset,ppreset,ng,over", CASES, ids=[c[0] for c in CASES])
def test_write_qual_mode_bit_exact_vs_oracle(sp, oracle, name, spreset, ppreset, ng, over):
    """-w/--writeBam mode (sp_set_write_qual): the quality array of every record as the reference's
    calc_update_baq_all leaves it (what sam_write1 emits, secphase.c:182-189), byte for byte; scores,
    markers and selection are unchanged by the mode; switching the mode off again restores the
    marker-rows-only path on the same context."""
    s, b, codes, off = make_case(spreset, ng, **over)
    ref = oracle_refseq(oracle, s)
    exp = oracle.run(b, oracle.preset_params(ppreset), ref, keep_hmm=False)
    with sp.Secphase(ppreset) as eng:
        eng.set_reference_codes(codes, off)
        eng.set_write_qual(True)
        got = eng.run_debug(b)
        bad = compare_results(exp, got, label="cuda-full")
        assert not bad, "\n".join(bad)
        assert got["baq_qual"].shape == exp["qual"].shape == b.qual_pool.shape
        diff = np.flatnonzero(exp["qual"] != got["baq_qual"])
        assert diff.size == 0, f"{diff.size} quality bytes differ, first at {diff[:5]}"
        if len(exp["hmm"]):
            assert (got["baq_qual"] != b.qual_pool).any()
            # one HMM row per base of each window's write-back range [10, l_query-10)
            assert len(got["rows"]) == int((exp["hmm"][:, 2] - 20).sum())
        assert got["hmm_instances"] == len(exp["hmm"])
        # through a second slot and with the pools in page-locked memory
        got2 = eng.run(sp.pin_batch(b), slot=1)
        assert np.array_equal(got2["baq_qual"], exp["qual"])
        eng.set_write_qual(False)
        eng.rng_seed(1)  # (the tie-break stream runs on across the batches of a context)
        got3 = eng.run_debug(b)
        assert "baq_qual" not in got3
        assert not compare_results(exp, got3, label="cuda-after-full")
        assert len(got3["rows"]) < len(got["rows"]) or len(exp["hmm"]) == 0


def test_write_qual_lane_private_layout(sp, oracle, monkeypatch):
    """The -w mode's fallback layout (instances keep their forward rows to themselves: used when a batch
    has bands too wide for the shared-memory kernel or windows longer than the length sort resolves)."""
    monkeypatch.setenv("SECPHASE_B200_NO_INTERLEAVE", "1")
    for spreset, ppreset, ng, over in (("hifi", "hifi", 60, dict(locus_len=300000)), ("stress", "hifi", 16, dict(locus_len=300000))):
        s, b, codes, off = make_case(spreset, ng, **over)
        exp = oracle.run(b, oracle.preset_params(ppreset), oracle_refseq(oracle, s))
        with sp.Secphase(ppreset) as eng:
            eng.set_reference_codes(codes, off)
            eng.set_write_qual(True)
            got = eng.run(b)
        assert np.array_equal(got["baq_qual"], exp["qual"])
        assert np.array_equal(got["scores"].view(np.int64), exp["scores"].view(np.int64))


def test_write_qual_without_consensus_long_windows(sp, oracle):
    """-q without -c: the HMM windows are whole confident blocks (thousands of bases), beyond the
    length-sorted range of the interleaved layout -> the launcher must fall back, results unchanged."""
    s, b, codes, off = make_case("hifi", 6, locus_len=200000, len_mean=4000, len_sd=500, len_min=3000)
    op = oracle.preset_params("hifi", consensus=0)
    exp = oracle.run(b, op, oracle_refseq(oracle, s))
    with sp.Secphase("hifi", consensus=0) as eng:
        eng.set_reference_codes(codes, off)
        eng.set_write_qual(True)
        got = eng.run(b)
    assert np.array_equal(got["baq_qual"], exp["qual"])
    assert np.array_equal(got["scores"].view(np.int64), exp["scores"].view(np.int64))
    assert int(exp["hmm"][:, 2].max()) > 2046 or len(exp["hmm"]) == 0


@pytest.mark.parametrize("name", ["hifi", "stress"])
def test_block_capacity_retry(sp, oracle, name):
    """A batch whose kernels overflow the tight per-group bound on consensus blocks is re-run inside sp_wait
    with the provable bound (sp_plan.h); the caller sees the same results, not SP_ECAPACITY.  The overflow is
    forced by clamping the first plan's workspaces; submit and resident paths, and a second batch after it."""
    _, spreset, ppreset, ng, over = [c for c in CASES if c[0] == name][0]
    s, b, codes, off = make_case(spreset, ng, **over)
    exp = oracle.run(b, oracle.preset_params(ppreset), oracle_refseq(oracle, s), keep_hmm=False)
    with sp.Secphase(ppreset) as eng:
        eng.set_reference_codes(codes, off)
        eng.debug_force_block_cap(2)
        got = eng.run_debug(b)
        assert eng.cap_retries() == 1
        bad = compare_results(exp, got, label="cuda-after-retry")
        assert not bad, "\n".join(bad)
        eng.upload(b, slot=1)
        eng.run_resident(1)
        r2 = eng.wait(1)
        assert eng.cap_retries() == 2
        assert np.array_equal(r2["scores"].view(np.int64), exp["scores"].view(np.int64))
        eng.run_resident(1)  # the slot now holds the safe plan: no further retry
        r3 = eng.wait(1)
        assert eng.cap_retries() == 2
        assert np.array_equal(r3["scores"].view(np.int64), exp["scores"].view(np.int64))
        eng.debug_force_block_cap(0)
        r4 = eng.run(b, slot=2)
        assert eng.cap_retries() == 2
        assert np.array_equal(r4["scores"].view(np.int64), exp["scores"].view(np.int64))


def test_reference_windows_beyond_2_32(sp, oracle):
    """BASELINE configs[3] addressing (a 2 x 3.1 Gb assembly is a 6.2 GB replica per GPU): with 18 filler
    contigs of N (4.32e9 codes > 2^32) ahead of the real ones, every window the kernels fetch -- contig_off,
    SpItem::ref_off, the HMM's reference pointer -- lies beyond 32 bits.  Results must equal the
    reference's on the un-shifted assembly (fai_fetch of {ctg}:s-e addresses by contig, ptMarker.c:736-744)."""
    import copy
    s, b, codes, off = make_case("hifi", 40, locus_len=300000)
    exp = oracle.run(b, oracle.preset_params("hifi"), oracle_refseq(oracle, s), keep_hmm=False)
    n_fill, fill_len = 18, 240_000_000
    big = np.empty(n_fill * fill_len + len(codes), np.uint8)
    big[:n_fill * fill_len] = 4
    big[n_fill * fill_len:] = codes
    off2 = np.concatenate([np.arange(n_fill, dtype=np.int64) * fill_len, off + n_fill * fill_len])
    assert off2[n_fill] > 2 ** 32
    b2 = copy.copy(b)
    b2.tid = (b.tid + n_fill).astype(np.int32)
    with sp.Secphase("hifi") as eng:
        eng.set_reference_codes(big, off2)
        del big
        got = eng.run_debug(b2)
        bad = compare_results(exp, got, label="cuda-6GB-replica")
        assert not bad, "\n".join(bad)
        assert got["hmm_instances"] == len(exp["hmm"]) > 0
        eng.set_write_qual(True)
        full = eng.run(b2)
        assert np.array_equal(full["baq_qual"], exp["qual"])


def test_reference_ascii_upload_matches_codes(sp, oracle):
    s, b, codes, off = make_case("hifi", 30, locus_len=200000, n_rate=1e-3)
    with sp.Secphase("hifi") as e1, sp.Secphase("hifi") as e2:
        e1.set_reference_codes(codes, off)
        e2.set_reference_ascii([s.contig_ptr(i) for i in range(s.n_contigs)], s.lens)
        r1, r2 = e1.run(b), e2.run(b)
    assert np.array_equal(r1["scores"].view(np.int64), r2["scores"].view(np.int64))
    assert np.array_equal(r1["groups"], r2["groups"])


def _random_hmm_instances(rng, n):
    refs, queries, bws, rows = [], [], [], []
    for trial in range(n):
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
        refs.append(ref)
        queries.append(query)
        bws.append(abs(lr - len(query)) + 20 if trial % 5 else int(rng.integers(1, 70)))
        rows.append(np.arange(len(query), dtype=np.int32))
    return refs, queries, bws, rows


def test_hmm_fast_mode_all_rows_integers_exact(sp, oracle):
    """Default K4 arrangement through sp_hmm_batch: fast kernel (FMA, scale-free), guard band on all 101
    thresholds and on the two best posteriors, strict re-run of the flagged instances.  state and q of EVERY
    row equal the restated probaln_glocal; 1 - pmax is within the band's own tolerance where it was not
    recomputed."""
    rng = np.random.default_rng(23)
    for preset in ("hifi", "ont"):
        op = oracle.preset_params(preset)
        refs, queries, bws, rows = _random_hmm_instances(rng, 200)
        with sp.Secphase(preset) as eng:
            assert eng.hmm_mode() == "fast"
            st, qq, pm, ms = eng.hmm_batch(refs, queries, bws, rows)
        worst = 0.0
        for j in range(len(refs)):
            iq = np.full(len(queries[j]), op.set_q, np.uint8)
            o = oracle.probaln(refs[j], queries[j], iq, np.float32(op.conf_d), np.float32(op.conf_e), bws[j])
            assert np.array_equal(o["state"], st[j]), (preset, j)
            assert np.array_equal(o["q"], qq[j]), (preset, j)
            ts, tf = 1.0 - o["pmax"], 1.0 - pm[j]
            assert np.all(np.abs(ts - tf) <= 64 * 2.0 ** -53 + 1e-9 * ts), (preset, j)
            worst = max(worst, float(np.abs(ts - tf).max() * 2.0 ** 53))
        assert worst <= 64


def test_hmm_all_rows_bit_exact_vs_port(sp, oracle):
    """Strict mode: every row's state, q and the normalised max posterior (bitwise) for random instances."""
    rng = np.random.default_rng(11)
    for preset in ("hifi", "ont"):
        op = oracle.preset_params(preset)
        refs, queries, bws, rows = [], [], [], []
        for trial in range(160):
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
            refs.append(ref)
            queries.append(query)
            bws.append(abs(lr - len(query)) + 20 if trial % 5 else int(rng.integers(1, 70)))
            rows.append(np.arange(len(query), dtype=np.int32))
        with sp.Secphase(preset) as eng:
            eng.set_hmm_mode("strict")
            st, qq, pm, ms = eng.hmm_batch(refs, queries, bws, rows)
        for j in range(len(refs)):
            iq = np.full(len(queries[j]), op.set_q, np.uint8)
            o = oracle.probaln(refs[j], queries[j], iq, np.float32(op.conf_d), np.float32(op.conf_e), bws[j])
            assert np.array_equal(o["state"], st[j]), (preset, j)
            assert np.array_equal(o["q"], qq[j]), (preset, j)
            assert np.array_equal(o["pmax"].view(np.int64), pm[j].view(np.int64)), (preset, j)


def test_empty_and_degenerate_batches(sp, oracle):
    s, b, codes, off = make_case("hifi", 8, locus_len=200000)
    with sp.Secphase("hifi") as eng:
        eng.set_reference_codes(codes, off)
        empty = b.group_slice(0, 0)
        r = eng.run(empty)
        assert r["n_groups"] == 0 and len(r["scores"]) == 0
        one = b.group_slice(3, 4)
        r1 = eng.run(one)
        ref = oracle_refseq(oracle, s)
        e1 = oracle.run(one, oracle.preset_params("hifi"), ref)
        assert np.array_equal(e1["scores"].view(np.int64), r1["scores"].view(np.int64))


def test_slots_pipeline_in_order_matches_single_run(sp, oracle):
    s, b, codes, off = make_case("hifi", 90, locus_len=300000)
    with sp.Secphase("hifi") as eng:
        eng.set_reference_codes(codes, off)
        whole = eng.run(b)
    with sp.Secphase("hifi") as eng:
        eng.set_reference_codes(codes, off)
        parts = [b.group_slice(0, 30), b.group_slice(30, 60), b.group_slice(60, 90)]
        for k, p in enumerate(parts):
            eng.submit(p, slot=k)
        outs = [eng.wait(slot=k) for k in range(3)]
    scores = np.concatenate([o["scores"] for o in outs])
    groups = np.concatenate([o["groups"] for o in outs])
    assert np.array_equal(scores.view(np.int64), whole["scores"].view(np.int64))
    assert np.array_equal(groups, whole["groups"])


def test_rng_matches_glibc(sp):
    import ctypes
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(1)
    with sp.Secphase("hifi") as eng:
        eng.rng_seed(1)
        for _ in range(1000):
            assert eng.rng_next() == libc.rand()


def test_warp_walker_and_serial_walker(sp, oracle, monkeypatch):
    """K1 has two forms: a warp per alignment (sp_walk_warp.cuh; what an aligner writes) and a thread per alignment
    (sp_walk.cuh; everything, and whatever the first declines).  Both against the reference's tables, and the
    hand-over between them: MD tags and CIGARs mixing M with =/X go to the serial walker, nothing else does."""
    cases = [("ont", "ont", 24, dict(locus_len=400000)),
             ("hifi", "hifi", 48, dict(locus_len=300000, eqx=1, n_rate=1e-4, clip_prob=0.8)),
             ("ont", "ont", 40, dict(locus_len=300000, len_mean=8000, len_sd=3000, len_min=1500, clip_prob=0.9, hard_clip_prob=0.9)),
             ("hifi", "hifi", 40, dict(locus_len=300000, use_md=1)),
             ("stress", "hifi", 12, dict(locus_len=300000))]
    for mode in ("warp", "serial"):
        monkeypatch.setenv("SECPHASE_B200_WALK", mode)   # ("warp": also for alignments the default leaves to the serial walker)
        # K2/K3 likewise: a lane per alignment (default, k_group_lanes) and a thread per read group (k_group)
        monkeypatch.setenv("SECPHASE_B200_GROUP", "lanes" if mode == "warp" else "serial")
        for k, (spreset, ppreset, ng, over) in enumerate(cases):
            s, b, codes, off = make_case(spreset, ng, **over)
            exp = oracle.run(b, oracle.preset_params(ppreset), oracle_refseq(oracle, s))
            A = len(b.n_cigar)
            with sp.Secphase(ppreset) as eng:
                eng.set_reference_codes(codes, off)
                got = eng.run_debug(b)
                fb = eng.walk_fallbacks()
                assert not compare_results(exp, got, label=f"cuda-{mode}-walk")
                assert fb == (-1 if mode == "serial" else A if over.get("use_md") else 0)
                if k == 1:
                    # '=' ops rewritten as 'M' next to '=' / 'X' ops: same alignment, but no longer the token
                    # structure the warp walker expects -> handed to the serial walker, same tables
                    b2 = b.group_slice(0, b.n_groups)
                    b2.cigar_pool = b2.cigar_pool.copy()
                    changed = set()
                    for a in range(0, A, 3):
                        c0, nc = int(b2.cigar_off[a]), int(b2.n_cigar[a])
                        eq = [j for j in range(c0, c0 + nc) if (int(b2.cigar_pool[j]) & 15) == 7]
                        if len(eq) > 1:
                            j = eq[len(eq) // 2]
                            b2.cigar_pool[j] = int(b2.cigar_pool[j]) & ~15
                            changed.add(a)
                    eng.rng_seed(1)
                    got2 = eng.run_debug(b2)
                    assert not compare_results(exp, got2, label=f"cuda-{mode}-walk-mixed-cigar")
                    assert eng.walk_fallbacks() == (-1 if mode == "serial" else len(changed)) and len(changed) > 3
