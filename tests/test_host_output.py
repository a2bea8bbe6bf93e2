"""Output side of the drop-in (libsecphase_host): block tables, the two merges, BED files and
out.log records, pinned on the reference's own known-answer tests (programs/src/secphase_test.c:
30-231) and, where oracle/_ref is built, on the reference's own ptBlock functions."""
import numpy as np
import pytest

from secphase_b200 import hostlib

CTG1 = [(30, 50), (5, 6), (10, 20), (0, 10), (50, 60)]   # secphase_test.c:46-50,101-105,168-172
CTG2 = [(50, 60), (5, 15), (10, 10), (8, 8), (0, 10)]    # secphase_test.c:52-56,107-111,174-178


def table(with_count):
    t = hostlib.Blocks(with_count)
    for s, e in CTG1:
        t.add("ctg1", s, e)
    for s, e in CTG2:
        t.add("ctg2", s, e)
    return t


def test_kat_merge_without_count():
    # test_mergingBlocksWithoutCount, truth at secphase_test.c:124-128
    t = table(False)
    t.merge()
    rows = t.rows()
    assert rows[rows[:, 0] == 0][:, 1:3].tolist() == [[0, 20], [30, 60]]
    assert rows[rows[:, 0] == 1][:, 1:3].tolist() == [[0, 15], [50, 60]]


def test_kat_merge_with_count_v2():
    # test_mergingBlocksWithCount_v2, truth at secphase_test.c:197-203
    t = table(True)
    t.merge_v2()
    rows = t.rows()
    c1, c2 = rows[rows[:, 0] == 0], rows[rows[:, 0] == 1]
    assert c1[:, 1].tolist() == [0, 5, 7, 10, 11, 30, 50, 51]
    assert c1[:, 2].tolist() == [4, 6, 9, 10, 20, 49, 50, 60]
    assert c1[:, 3].tolist() == [1, 2, 1, 2, 1, 1, 2, 1]
    assert c2[:, 1].tolist() == [0, 5, 8, 9, 10, 11, 50]
    assert c2[:, 2].tolist() == [4, 7, 8, 9, 10, 15, 60]
    assert c2[:, 3].tolist() == [1, 2, 3, 2, 3, 1, 1]
    assert t.total_number() == 15
    assert t.total_length() == 52 + 27  # ptBlock_get_total_length_by_rf: sum of (rfe - rfs + 1)


def test_bed_writer(tmp_path):
    t = hostlib.Blocks(True)
    t.add("b#2#ctg", 10, 19)
    t.add("b#1#ctg", 5, 5)
    t.add("B", 7, 6)          # rfe < rfs: skipped on output (ptBlock.c:589-591)
    t.add("b#2#ctg", 15, 30)
    t.merge_v2()
    p = tmp_path / "x.bed"
    t.save_bed(str(p))
    # contigs in strcmp order; end is rfe + 1
    assert p.read_text() == "b#1#ctg\t5\t6\t1\nb#2#ctg\t10\t15\t1\nb#2#ctg\t15\t20\t2\nb#2#ctg\t20\t31\t1\n"
    u = hostlib.Blocks(False)
    u.add("c", 3, 3)
    u.add("c", 3, 3)
    u.add("c", 4, 4)
    u.add("c", -1, -1)        # a match marker whose ref_pos was never filled
    u.merge_v2()
    u.save_bed(str(p))
    assert p.read_text() == "c\t-1\t0\nc\t3\t4\nc\t4\t5\n"


def test_out_log_record_format():
    # README.md:136-140 shows the shape of one record; print_alignment_scores is secphase.c:32-57
    rec = hostlib.format_marker_record("m64043_200710_174426/2353/ccs", [0, 0x100, 0x110],
                                       [-35.0, -0.0, -12.3456], ["HG00438#1#JAHBCB010000044.1",
                                                                "HG00438#2#JAHBCA010000036.1", "c3"],
                                       [4904283, 3898995, 7], [4924283, 3918995, 9], 1)
    assert rec == ("#MARKER SCORE\n$\tm64043_200710_174426/2353/ccs\n"
                   "*\t-35.00\tHG00438#1#JAHBCB010000044.1\t4904283\t4924283\n"
                   "@\t-0.00\tHG00438#2#JAHBCA010000036.1\t3898995\t3918995\n"
                   "!\t-12.35\tc3\t7\t9\n\n")


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("with_count", [False, True])
def test_merges_match_the_reference_functions(oracle, mode, with_count):
    if "reference" not in oracle.available_kinds():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(1234 + 2 * mode + with_count)
    for trial in range(300):
        n = int(rng.integers(0, 14))
        span = int(rng.choice([6, 30, 200]))
        s = rng.integers(0, span, n)
        e = s + rng.integers(0, max(2, span // 3), n) * (rng.random(n) < 0.8)
        cnt = rng.integers(1, 4, n) if with_count else np.full(n, -1)
        rows = np.stack([s, e, cnt], 1).astype(np.int32)
        exp = oracle.merge_blocks(rows, mode)
        t = hostlib.Blocks(with_count)
        for a, b, c in rows.tolist():
            t.add("ctg", a, b, max(c, 0))
        (t.merge_v2 if mode else t.merge)()
        got = t.rows()[:, 1:]
        if not with_count:
            got = got.copy()
            got[:, 2] = -1
        assert got.tolist() == exp.tolist(), (trial, rows.tolist())
