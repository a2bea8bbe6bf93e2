"""bench.py's reference arm runs without a GPU: it must print exactly one JSON line on stdout with the
keys of the benchmark contract (the GPU arm is exercised on the B200 box by tools/gpu_round.sh)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-sample", "48", "--locus-len", "2000000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "read_groups_per_sec" and d["unit"] == "read-groups/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_parity_machinery_of_the_bench_matches_a_single_oracle_run():
    """bench.py checks the GPU at size against a thread-pooled oracle run whose per-chunk tables are concatenated and
    whose selection is the reference's get_best_record_index replayed in group order: that composite must equal one
    sequential oracle run (same tables, same selection, ties included), also with filler contigs ahead of the real
    ones (the configs[3] workload)."""
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench
    from oracle import pyoracle
    from tools.parity import compare_results
    if "reference" not in pyoracle.available_kinds():
        pyoracle.build()
    sp = dict(bench.SPECS["stress"], locus_len=200_000, synth_over=dict(snv_rate=1e-4, indel_rate=2e-5, long_indel_rate=2e-6))
    wl = bench.Workload(sp)
    b = wl.generate(0, 40)
    keep = {}
    g, cells, dt, kind = bench.cpu_reference_run(wl, [b], "hifi", threads=3, keep=keep)
    assert g == 40 and cells > 0 and kind == "reference"
    one = pyoracle.run(b, pyoracle.preset_params("hifi"), wl.oracle_refseq(pyoracle))
    assert not compare_results(one, keep, label="pooled")
    sc, gao = one["scores"], b.grp_aln_off
    ties = sum(1 for gi in range(b.n_groups)
               if (lambda s: len(s) > 1 and (s == s.max()).sum() > 1)(sc[gao[gi]:gao[gi + 1]][(b.flag[gao[gi]:gao[gi + 1]] & 256) != 0]))
    assert ties >= 1
    # filler contigs: same groups, every tid shifted, same tables
    sp2 = dict(bench.SPECS["hifi"], locus_len=200_000, filler_bp=3_000_000)
    wl2 = bench.Workload(sp2)
    assert wl2.tid_shift >= 1 and wl2.off[wl2.tid_shift] == sum(wl2.lens[:wl2.tid_shift]) and len(wl2.codes) == wl2.off[-1]
    assert (wl2.codes[:wl2.off[wl2.tid_shift]] == 4).all()
    b2 = wl2.generate(0, 12)
    plain = bench.Workload(dict(sp2, filler_bp=0))
    b0 = plain.generate(0, 12)
    assert np.array_equal(b2.tid, b0.tid + wl2.tid_shift)
    r2 = pyoracle.run(b2, pyoracle.preset_params("hifi"), wl2.oracle_refseq(pyoracle))
    r0 = pyoracle.run(b0, pyoracle.preset_params("hifi"), plain.oracle_refseq(pyoracle))
    assert not compare_results(r0, r2, label="filler")
