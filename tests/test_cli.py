"""The `secphase` executable (bin/secphase): option handling without a GPU, and -- on the B200 box --
whole runs on synthetic BAM + FASTA whose out.log / BED files must equal, byte for byte, what the
reference's own marker path writes for the same read groups (oracle/_ref driven as secphase.c:
156-219 + 681-732 drive it, one worker, input order)."""
import json
import os
import subprocess

import pytest

from secphase_b200 import hostlib
from tests.conftest import make_case, oracle_refseq

CLI = hostlib.CLI_PATH


def run_cli(args, **kw):
    return subprocess.run([CLI] + args, capture_output=True, text=True, timeout=600, **kw)


@pytest.fixture(scope="module")
def small_inputs(tmp_path_factory):
    d = tmp_path_factory.mktemp("cli")
    s, b, _, _ = make_case("hifi", 24, locus_len=150000)
    bam, fa = str(d / "in.bam"), str(d / "asm.fa")
    hostlib.write_bam(bam, s.names, s.lens, b)
    hostlib.write_fasta(fa, s.names, [s.contig_ptr(i) for i in range(s.n_contigs)], s.lens)
    return d, s, b, bam, fa


def test_help_and_bad_option():
    r = run_cli(["-h"])
    assert r.returncode == 1 and "Usage: secphase  -i <INPUT_BAM> -f <FASTA>" in r.stderr   # secphase.c:563,599
    r = run_cli(["--no-such-option"])
    assert r.returncode == 1 and "Usage:" in r.stderr
    for opt in ("--hifi, -x", "--ont, -y", "--primMarginScore, -p", "--flankMargin, -F", "--threads, -@"):
        assert opt in r.stderr


def test_both_presets_rejected(small_inputs):
    d, _, _, bam, fa = small_inputs
    r = run_cli(["-i", bam, "-f", fa, "--hifi", "--ont", "-o", str(d / "o1")])
    assert r.returncode == 1   # EXIT_FAILURE, secphase.c:633-637
    assert "Presets --hifi and --ont cannot be enabled at the same time" in r.stderr


def test_out_of_scope_modes_say_so(small_inputs):
    d, _, _, bam, fa = small_inputs
    r = run_cli(["-i", bam, "-f", fa, "-v", "x.vcf", "-o", str(d / "o2")])
    assert r.returncode == 1 and "variant mode" in r.stderr


def test_missing_inputs(small_inputs):
    d, _, _, bam, fa = small_inputs
    r = run_cli(["-i", str(d / "nope.bam"), "-f", fa, "--hifi", "-o", str(d / "o4")])
    assert r.returncode == 1 and "nope.bam" in r.stderr
    r = run_cli(["-i", bam, "-f", str(d / "nope.fa"), "--hifi", "-o", str(d / "o5")])
    assert r.returncode == 1 and "nope.fa" in r.stderr


def test_no_gpu_is_a_loud_failure(small_inputs):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    d, _, _, bam, fa = small_inputs
    r = run_cli(["-i", bam, "-f", fa, "--hifi", "-o", str(d / "o6")])
    assert r.returncode == 1 and "cannot initialise GPU 0" in r.stderr   # no CPU fallback


# ------------------------------------------------------------------------------------------ GPU
OUT_FILES = ["out.log", "modified_read_blocks.markers.bed", "marker_blocks.bed"]
EMPTY_FILES = ["initial_variant_blocks.bed", "modified_read_blocks.variants.bed", "variant_blocks.bed"]

GPU_CASES = [
    ("hifi", "hifi", ["--hifi"], 160, dict(locus_len=300000)),
    ("ont", "ont", ["--ont"], 40, dict(locus_len=300000)),
    ("stress", "hifi", ["--hifi"], 32, dict(locus_len=200000)),
    ("hifi_p20_n50", "hifi", ["--hifi", "-p", "20", "-n", "-50"], 100, dict(locus_len=300000)),   # README values
    ("ont_md_clips", "ont", ["--ont"], 30, dict(locus_len=300000, use_md=1, clip_prob=0.9, hard_clip_prob=0.9)),
]


@pytest.mark.gpu
@pytest.mark.parametrize("name,preset,flags,n_groups,over", GPU_CASES)
def test_cli_outputs_equal_the_reference(tmp_path, oracle, name, preset, flags, n_groups, over):
    if "reference" not in oracle.available_kinds():
        pytest.skip("oracle/_ref not built")
    s, b, _, _ = make_case(preset, n_groups, **over)
    bam, fa = str(tmp_path / "in.bam"), str(tmp_path / "asm.fa")
    hostlib.write_bam(bam, s.names, s.lens, b)
    hostlib.write_fasta(fa, s.names, [s.contig_ptr(i) for i in range(s.n_contigs)], s.lens)
    pover = {}
    if "-p" in flags:
        pover = dict(prim_margin_score=20.0, min_score=-50)
    exp_dir = tmp_path / "exp"
    exp_dir.mkdir()
    exp = oracle.run(b, oracle.preset_params("ont" if "--ont" in flags else "hifi", **pover), oracle_refseq(oracle, s),
                     kind="reference", outputs=(str(exp_dir), "ref"))
    for variant, extra in (("one_batch", []), ("many_batches", ["--batchGroups", "7", "-@", "2"])):
        out_dir = tmp_path / f"got_{variant}"
        r = run_cli(["-i", bam, "-f", fa, "-o", str(out_dir), "-P", "t"] + flags + extra)
        assert r.returncode == 0, r.stderr
        for f in OUT_FILES:
            got = (out_dir / f"t.{f}").read_bytes()
            want = (exp_dir / f"ref.{f}").read_bytes()
            assert got == want, (variant, f)
        for f in EMPTY_FILES:
            assert (out_dir / f"t.{f}").read_bytes() == b""
        assert f"Number of reads modified by marker score = {exp['reads_modified_by_marker']}" in r.stderr
        assert f"Total length of read blocks modified by markers: {exp['totals'][0]}." in r.stderr
        assert f"Total number of projected marker blocks on all haplotypes: {exp['totals'][3]}." in r.stderr
        summary = json.loads(r.stderr.split("[secphase_b200] ")[1].splitlines()[0])
        assert summary["read_groups"] == b.n_groups and summary["gpu_launches"] > 0
    assert exp["reads_modified_by_marker"] > 0, "the case must exercise the output path"


@pytest.mark.gpu
def test_cli_contig_order_and_extra_contigs(tmp_path, oracle):
    """Names, not order, link the BAM header and the FASTA (ptMarker.c:739,819-820)."""
    if "reference" not in oracle.available_kinds():
        pytest.skip("oracle/_ref not built")
    s, b, _, _ = make_case("hifi", 60, locus_len=200000)
    bam, fa = str(tmp_path / "in.bam"), str(tmp_path / "asm.fa")
    hostlib.write_bam(bam, s.names, s.lens, b)
    order = list(range(s.n_contigs))[::-1]
    names = [s.names[i] for i in order] + ["unrelated"]
    seqs = [s.contig_ptr(i) for i in order] + [b"ACGTACGTAC"]
    hostlib.write_fasta(fa, names, seqs, [s.lens[i] for i in order] + [10], line_width=80)
    exp_dir = tmp_path / "exp"
    exp_dir.mkdir()
    oracle.run(b, oracle.preset_params("hifi"), oracle_refseq(oracle, s), kind="reference", outputs=(str(exp_dir), "ref"))
    r = run_cli(["-i", bam, "-f", fa, "-o", str(tmp_path / "got"), "--hifi"])
    assert r.returncode == 0, r.stderr
    for f in OUT_FILES:
        assert (tmp_path / "got" / f"secphase.{f}").read_bytes() == (exp_dir / f"ref.{f}").read_bytes(), f


@pytest.mark.gpu
def _two_contexts_env():
    """--gpus 2 on this box: two physical devices when present, else two contexts on device 0
    (SECPHASE_B200_DEVICES; above the C ABI the two cases are the same code path)."""
    import torch
    env = dict(os.environ)
    if torch.cuda.device_count() < 2:
        env["SECPHASE_B200_DEVICES"] = "0,0"
    return env


@pytest.mark.gpu
def test_cli_two_gpus_failure_exits_instead_of_hanging(tmp_path):
    """A GPU thread that fails must not leave the reader, the other GPU thread and the ordered output
    waiting on each other: the run ends with the error on stderr and exit code 1."""
    s, b, _, _ = make_case("hifi", 200, locus_len=300000)
    bam, fa = str(tmp_path / "in.bam"), str(tmp_path / "asm.fa")
    hostlib.write_bam(bam, s.names, s.lens, b)
    hostlib.write_fasta(fa, s.names, [s.contig_ptr(i) for i in range(s.n_contigs)], s.lens)
    for where in ("0,1", "1,0", "1,3"):
        env = dict(_two_contexts_env(), SECPHASE_B200_FAIL_AT=where)
        r = subprocess.run([CLI, "-i", bam, "-f", fa, "-o", str(tmp_path / ("f" + where.replace(",", "_"))), "--hifi",
                            "--gpus", "2", "--batchGroups", "16"], capture_output=True, text=True, timeout=120, env=env)
        assert r.returncode == 1, (where, r.stderr[-500:])
        assert "injected failure" in r.stderr


@pytest.mark.gpu
def test_cli_two_gpus_same_output(tmp_path, oracle):
    env2 = _two_contexts_env()
    s, b, _, _ = make_case("hifi", 200, locus_len=300000)
    bam, fa = str(tmp_path / "in.bam"), str(tmp_path / "asm.fa")
    hostlib.write_bam(bam, s.names, s.lens, b)
    hostlib.write_fasta(fa, s.names, [s.contig_ptr(i) for i in range(s.n_contigs)], s.lens)
    outs = []
    for n in (1, 2):
        d = tmp_path / f"g{n}"
        r = run_cli(["-i", bam, "-f", fa, "-o", str(d), "--hifi", "--gpus", str(n), "--batchGroups", "16"], env=env2)
        assert r.returncode == 0, r.stderr
        outs.append([(d / f"secphase.{f}").read_bytes() for f in OUT_FILES])
        assert f'"gpus": {n}' in r.stderr
    assert outs[0] == outs[1]
    assert len(outs[0][0]) > 0  # out.log is not empty: secondaries were selected


@pytest.mark.gpu
@pytest.mark.parametrize("preset,flags,n_groups,over", [
    ("hifi", ["--hifi"], 120, dict(locus_len=300000)),
    ("ont", ["--ont"], 30, dict(locus_len=300000, clip_prob=0.8, hard_clip_prob=0.5)),
])
def test_cli_write_bam(tmp_path, oracle, preset, flags, n_groups, over):
    """-w/--writeBam (secphase.c:182-189, 643-657): <prefix>.quality_modified.out.bam holds the input
    header and every record of every scored read group with the qualities the reference's own
    calc_update_baq_all leaves in the records; the other outputs do not change with -w."""
    if "reference" not in oracle.available_kinds():
        pytest.skip("oracle/_ref not built")
    from tests.test_host_sam import py_sam_line
    s, b, _, _ = make_case(preset, n_groups, **over)
    bam, fa = str(tmp_path / "in.bam"), str(tmp_path / "asm.fa")
    hostlib.write_bam(bam, s.names, s.lens, b)
    hostlib.write_fasta(fa, s.names, [s.contig_ptr(i) for i in range(s.n_contigs)], s.lens)
    exp_dir = tmp_path / "exp"
    exp_dir.mkdir()
    exp = oracle.run(b, oracle.preset_params(preset), oracle_refseq(oracle, s), kind="reference",
                     outputs=(str(exp_dir), "ref"))
    want = []
    for g in range(b.n_groups):
        for a in range(int(b.grp_aln_off[g]), int(b.grp_aln_off[g + 1])):
            q = exp["qual"][int(b.qual_off[a]):int(b.qual_off[a + 1])]
            want.append(py_sam_line(b, g, a, s.names, qual=q))
    assert (exp["qual"] != b.qual_pool).any()
    for variant, extra in (("one_batch", []), ("many_batches", ["--batchGroups", "7", "-@", "2"])):
        out_dir = tmp_path / f"got_{variant}"
        r = run_cli(["-i", bam, "-f", fa, "-o", str(out_dir), "-P", "t", "-w"] + flags + extra)
        assert r.returncode == 0, r.stderr
        lines = (out_dir / "t.quality_modified.out.bam").read_text().splitlines(keepends=True)
        head = [ln for ln in lines if ln.startswith("@")]
        assert [ln.split("\t")[1] for ln in head if ln.startswith("@SQ")] == [f"SN:{n}" for n in s.names]
        assert lines[len(head):] == want, variant
        for f in OUT_FILES:
            assert (out_dir / f"t.{f}").read_bytes() == (exp_dir / f"ref.{f}").read_bytes(), (variant, f)
