"""Host ingest (libsecphase_host): BGZF/BAM -> read groups -> flat batches, FASTA -> codes.

Checked against the writer's input (round trips on every synthetic preset, with the chunk size
shrunk so that records, groups and the header straddle chunk boundaries) and against a Python
restatement of the reference's grouping/eligibility loop (secphase.c:266-340)."""
import os

import numpy as np
import pytest

from secphase_b200 import hostlib
from secphase_b200.flatbatch import _FIELDS, FlatBatch
from tests.conftest import make_case

ROOT_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def assert_same_batch(got, exp):
    for name, _ in _FIELDS:
        a, e = getattr(got, name), getattr(exp, name)
        assert a.shape == e.shape, name
        assert (a == e).all(), name


def read_all(path, max_groups=4096, max_bytes=256 << 20, threads=3):
    out = []
    with hostlib.BamReader(path, threads=threads) as r:
        while True:
            b = r.next_batch(max_groups, max_bytes)
            if b is None:
                break
            assert 0 < b.n_groups <= max_groups
            out.append(b)
        counts, skipped = r.counts(), r.skipped_groups()
    return out, counts, skipped


@pytest.mark.parametrize("preset,over", [("hifi", {}), ("ont", {}), ("stress", {}),
                                         ("hifi", dict(use_md=1, clip_prob=0.9, hard_clip_prob=0.9))])
@pytest.mark.parametrize("chunk", [None, 70_000])
def test_bam_round_trip(tmp_path, monkeypatch, preset, over, chunk):
    s, b, _, _ = make_case(preset, 40 if preset != "hifi" else 90, locus_len=200000, **over)
    if chunk:
        monkeypatch.setenv("SPH_CHUNK_BYTES", str(chunk))
        monkeypatch.setenv("SPH_HEAD_ROOM", "512")
    p = str(tmp_path / "t.bam")
    hostlib.write_bam(p, s.names, s.lens, b, level=1, threads=2)
    got, counts, skipped = read_all(p, max_groups=17)
    assert_same_batch(FlatBatch.concat(got), b)
    assert counts == (b.n_alns + 1, b.n_groups)  # the reference counts the failing read too (secphase.c:268-269)
    assert skipped == 0
    # byte-bounded batches give the same groups
    got2, _, _ = read_all(p, max_groups=1 << 20, max_bytes=200_000)
    assert len(got2) > 1
    assert_same_batch(FlatBatch.concat(got2), b)


def reference_grouping(records):
    """secphase.c:266-340 on (qname, flag) records -> list of (qname, [record indices]) handed to the pool."""
    out, cur, name = [], [], None
    for i, (qn, fl) in enumerate(list(records) + [(None, None)]):
        if name is None and qn is not None:
            name = qn
        if qn != name or qn is None:
            n = len(cur)
            supp = sum(1 for j in cur if records[j][1] & 0x800)
            prim = sum(1 for j in cur if not records[j][1] & 0x100)
            if 1 < n <= 10 and supp == 0 and prim == 1:
                out.append((name, cur))
            cur, name = [], qn
        if qn is None:
            break
        if fl & 0x4:
            continue
        if len(cur) > 10:
            continue
        cur.append(i)
    return out


def test_grouping_and_eligibility(tmp_path):
    s, b, _, _ = make_case("stress", 12, locus_len=150000)
    a0 = [int(x) for x in b.grp_aln_off]
    first = lambda g: a0[g]  # noqa: E731
    sec = lambda g, k: a0[g] + 1 + (k % (a0[g + 1] - a0[g] - 1))  # noqa: E731
    P, S = 0x0, 0x100
    groups = [
        ("r00_ok", [(first(0), P), (sec(0, 0), S)]),
        ("r01_single", [(first(1), P)]),
        ("r02_unmapped_inside", [(first(2), P), (sec(2, 0), S | 0x4), (sec(2, 1), S)]),
        ("r03_supp", [(first(3), P), (sec(3, 0), S), (sec(3, 1), 0x800)]),
        ("r04_two_primaries", [(first(4), P), (sec(4, 0), P)]),
        ("r05_no_primary", [(sec(5, 0), S), (sec(5, 1), S)]),
        ("r06_ten", [(first(6), P)] + [(sec(6, k), S) for k in range(9)]),
        ("r07_eleven", [(first(7), P)] + [(sec(7, k), S) for k in range(10)]),
        ("r08_thirteen", [(first(8), P)] + [(sec(8, k), S) for k in range(12)]),
        ("r09_all_unmapped", [(first(9), 0x4), (sec(9, 0), S | 0x4)]),
        ("r10_reverse_sec", [(first(10), P), (sec(10, 0), S | (int(b.flag[sec(10, 0)]) & 0x10))]),
        ("r11_ok_last", [(first(11), P), (sec(11, 0), S), (sec(11, 1), S)]),
    ]
    full = FlatBatch.assemble(b, groups)
    p = str(tmp_path / "e.bam")
    hostlib.write_bam(p, s.names, s.lens, full)
    records = [(qn, fl if fl is not None else 0) for qn, alns in groups for _, fl in alns]
    exp = reference_grouping(records)
    got, counts, skipped = read_all(p, max_groups=3)
    cat = FlatBatch.concat(got)
    assert [cat.qname(g) for g in range(cat.n_groups)] == [qn for qn, _ in exp]
    assert [qn for qn, _ in exp] == ["r00_ok", "r02_unmapped_inside", "r06_ten", "r10_reverse_sec", "r11_ok_last"]
    # the kept alignments are exactly the reference's, in file order
    rec_idx = np.concatenate([g.record_index for g in got])
    assert rec_idx.tolist() == [j for _, idx in exp for j in idx]
    assert counts == (len(records) + 1, len(groups))
    assert skipped == 0
    # and they carry the right data
    want = FlatBatch.assemble(full, [(qn, [(j, None) for j in idx]) for qn, idx in exp])
    assert_same_batch(cat, want)


def test_groups_the_reference_cannot_score_are_skipped(tmp_path):
    s, b, _, _ = make_case("hifi", 6, locus_len=150000)
    bad_n = FlatBatch.assemble(b, [(b.qname(g), [(a, None) for a in range(b.grp_aln_off[g], b.grp_aln_off[g + 1])])
                                   for g in range(b.n_groups)])
    # group 1: an N (ref-skip) CIGAR op; group 3: alignment running past the contig end
    a1 = int(bad_n.grp_aln_off[1])
    bad_n.cigar_pool[int(bad_n.cigar_off[a1])] = (5 << 4) | 3
    a3 = int(bad_n.grp_aln_off[3])
    bad_n.pos[a3] = s.lens[int(bad_n.tid[a3])] - 10
    p = str(tmp_path / "n.bam")
    hostlib.write_bam(p, s.names, s.lens, bad_n)
    got, _, skipped = read_all(p)
    cat = FlatBatch.concat(got)
    assert [cat.qname(g) for g in range(cat.n_groups)] == [b.qname(g) for g in (0, 2, 4, 5)]
    assert skipped == 2


def test_missing_tag_is_an_error(tmp_path):
    s, b, _, _ = make_case("hifi", 3, locus_len=150000)
    b.tag_kind[:] = 2  # written as XX:Z
    p = str(tmp_path / "x.bam")
    hostlib.write_bam(p, s.names, s.lens, b)
    with hostlib.BamReader(p) as r:
        with pytest.raises(hostlib.HostError, match="MD or CS"):
            r.next_batch()


def test_corrupt_inputs(tmp_path):
    s, b, _, _ = make_case("hifi", 30, locus_len=150000)
    p = str(tmp_path / "ok.bam")
    hostlib.write_bam(p, s.names, s.lens, b, level=6)
    raw = open(p, "rb").read()
    # truncated in the middle of a block
    t = str(tmp_path / "trunc.bam")
    open(t, "wb").write(raw[:len(raw) // 2])
    with pytest.raises(hostlib.HostError):
        read_all(t)
    # a flipped payload byte: inflate error or CRC mismatch
    c = bytearray(raw)
    c[len(c) // 2] ^= 0x5A
    cpath = str(tmp_path / "crc.bam")
    open(cpath, "wb").write(bytes(c))
    with pytest.raises(hostlib.HostError):
        read_all(cpath)
    # not BGZF at all / gzip but not BAM
    npath = str(tmp_path / "plain.bam")
    open(npath, "wb").write(b"@HD\tVN:1.6\n" * 100)
    with pytest.raises(hostlib.HostError, match="BGZF"):
        hostlib.BamReader(npath)
    with pytest.raises(hostlib.HostError):
        hostlib.BamReader(str(tmp_path / "does_not_exist.bam"))


def test_empty_bam(tmp_path):
    s, b, _, _ = make_case("hifi", 1, locus_len=150000)
    p = str(tmp_path / "empty.bam")
    hostlib.write_bam(p, s.names, s.lens, [])
    with hostlib.BamReader(p) as r:
        assert r.names == s.names and r.lens == s.lens
        assert r.next_batch() is None
        assert r.counts() == (1, 1)  # secphase.c:268-281 on an empty file


def test_fasta_codes(tmp_path):
    rng = np.random.default_rng(5)
    seqs = []
    for n in (1, 59, 60, 61, 1000, 70001):
        seqs.append(bytes(rng.choice(np.frombuffer(b"ACGTacgtNnRYKM-", np.uint8), size=n)))
    names = [f"ctg{i}#x" for i in range(len(seqs))]
    table = np.full(256, 4, np.uint8)
    for ch, v in zip(b"ACGTacgt", [0, 1, 2, 3, 0, 1, 2, 3]):
        table[ch] = v
    exp = np.concatenate([table[np.frombuffer(x, np.uint8)] for x in seqs])
    for lw in (60, 7, 100000):
        p = str(tmp_path / f"a{lw}.fa")
        hostlib.write_fasta(p, names, seqs, [len(x) for x in seqs], line_width=lw)
        got_names, codes, off = hostlib.load_fasta(p, threads=3)
        assert got_names == names
        assert off.tolist() == np.concatenate([[0], np.cumsum([len(x) for x in seqs])]).tolist()
        assert (codes == exp).all()
    # CRLF line ends, description after the name, no final newline, empty contig
    p = str(tmp_path / "crlf.fa")
    open(p, "wb").write(b">a desc here\r\nACGT\r\nNN\r\n>empty\r\n>b\tx\r\nggtt")
    got_names, codes, off = hostlib.load_fasta(p)
    assert got_names == ["a", "empty", "b"]
    assert off.tolist() == [0, 6, 6, 10]
    assert codes.tolist() == [0, 1, 2, 3, 4, 4, 2, 2, 3, 3]
    with pytest.raises(hostlib.HostError):
        hostlib.load_fasta(str(tmp_path / "missing.fa"))
    bad = str(tmp_path / "bad.fa")
    open(bad, "wb").write(b"ACGT\n")
    with pytest.raises(hostlib.HostError, match="FASTA"):
        hostlib.load_fasta(bad)


def test_long_cigars_in_the_cg_tag_are_restored(tmp_path):
    """BAM keeps a CIGAR of more than 65535 operations in CG:B,I behind the placeholder <l_seq>S<ref_len>N; htslib's
    bam_read1 (through which the reference reads, secphase.c:268) swaps it back in.  The writer is told to use that
    form for every CIGAR longer than 8 operations (SPH_BAM_WRITE_LONG_CIGAR is read once per process: a child
    process writes), the reader must return the very same batches -- none skipped because of the N operation."""
    import subprocess, sys, textwrap
    s, b, _, _ = make_case("ont", 24, locus_len=200000, len_mean=6000, len_sd=1500, len_min=2000)
    assert int(b.n_cigar.max()) > 100
    p = str(tmp_path / "cg.bam")
    child = textwrap.dedent("""
        import sys
        sys.path.insert(0, %r)
        from secphase_b200 import hostlib
        from tests.conftest import make_case
        s, b, _, _ = make_case("ont", 24, locus_len=200000, len_mean=6000, len_sd=1500, len_min=2000)
        hostlib.write_bam(sys.argv[1], s.names, s.lens, b, level=1, threads=1)
    """ % ROOT_DIR)
    r = subprocess.run([sys.executable, "-c", child, p], env=dict(os.environ, SPH_BAM_WRITE_LONG_CIGAR="8"),
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-1000:]
    plain = str(tmp_path / "plain.bam")
    hostlib.write_bam(plain, s.names, s.lens, b, level=1, threads=1)
    assert os.path.getsize(p) != os.path.getsize(plain)   # the CG form really was written
    got, counts, skipped = read_all(p, max_groups=1 << 20)
    assert skipped == 0
    assert_same_batch(FlatBatch.concat(got), b)
