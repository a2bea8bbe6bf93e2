"""Malformed BAM records (structurally valid BGZF, valid CRCs, garbage in the record headers, truncated
streams) through the reader, correct_bam and secphase_index: an error or a consistent batch, never a crash.
The mutations run in child processes so that a segfault shows up as a failed test, not a dead runner."""
import os
import struct
import subprocess
import sys
import textwrap
import zlib

import numpy as np

from secphase_b200 import hostlib
from secphase_b200.build import CORRECT_BAM, INDEX_TOOL
from tests.conftest import make_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def bgzf_write(path, body):
    with open(path, "wb") as fh:
        for i in range(0, len(body), 60000):
            chunk = body[i:i + 60000]
            co = zlib.compressobj(1, zlib.DEFLATED, -15)
            comp = co.compress(chunk) + co.flush()
            fh.write(b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", len(comp) + 25) + comp +
                     struct.pack("<II", zlib.crc32(chunk), len(chunk)))
        fh.write(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))


def mutants(tmp_path, n, seed0):
    import gzip
    s, b, _, _ = make_case("hifi", 10, locus_len=100000, len_mean=3000, len_sd=500, len_min=1500)
    src = str(tmp_path / "src.bam")
    hostlib.write_bam(src, s.names, s.lens, b)
    raw = gzip.open(src, "rb").read()
    o = 8 + struct.unpack_from("<i", raw, 4)[0]
    n_ref = struct.unpack_from("<i", raw, o)[0]
    o += 4
    for _ in range(n_ref):
        o += 4 + struct.unpack_from("<i", raw, o)[0] + 4
    hdr_end, recs = o, []
    while o < len(raw):
        recs.append(o)
        o += 4 + struct.unpack_from("<i", raw, o)[0]
    paths = []
    for k in range(n):
        rng = np.random.default_rng(seed0 + k)
        m = bytearray(raw)
        for _ in range(int(rng.integers(1, 4))):
            r = recs[int(rng.integers(0, len(recs)))]
            f = int(rng.integers(0, 8))
            if f == 0: struct.pack_into("<i", m, r, int(rng.integers(-5, 1 << 20)))            # block_size
            elif f == 1: m[r + 12] = int(rng.integers(0, 256))                                  # l_read_name
            elif f == 2: struct.pack_into("<H", m, r + 16, int(rng.integers(0, 65536)))         # n_cigar_op
            elif f == 3: struct.pack_into("<i", m, r + 20, int(rng.integers(-10, 1 << 22)))     # l_seq
            elif f == 4: struct.pack_into("<i", m, r + 4, int(rng.integers(-3, 10)))            # refID
            elif f == 5: struct.pack_into("<i", m, r + 8, int(rng.integers(-10, 1 << 30)))      # pos
            elif f == 6: struct.pack_into("<H", m, r + 18, int(rng.integers(0, 65536)))         # flag
            else: m[r + int(rng.integers(36, 400))] = int(rng.integers(0, 256))                 # name / cigar / seq byte
        if k % 7 == 0:
            m = m[:int(rng.integers(hdr_end // 2, len(m)))]
        p = str(tmp_path / f"m{k}.bam")
        bgzf_write(p, bytes(m))
        paths.append(p)
    return paths


def test_reader_survives_malformed_records(tmp_path):
    paths = mutants(tmp_path, 40, 4242)
    child = textwrap.dedent("""
        import sys
        sys.path.insert(0, %r)
        from secphase_b200 import hostlib
        ok = err = 0
        for p in sys.argv[1:]:
            try:
                with hostlib.BamReader(p, threads=2, keep_records=True) as rd:
                    for fb in rd:
                        assert fb.n_groups > 0
                        assert int(fb.cigar_off[-1]) == len(fb.cigar_pool) and int(fb.qual_off[-1]) == len(fb.qual_pool)
                        assert int(fb.seq_off[-1]) == len(fb.seq_pool) and int(fb.tag_off[-1]) == len(fb.tag_pool)
                        for a in range(fb.n_alns):     # kept records must format or be rejected, never crash
                            try:
                                hostlib.format_sam_record(fb.rec_pool[int(fb.rec_off[a]):int(fb.rec_off[a + 1])], rd.names)
                            except hostlib.HostError:
                                pass
                ok += 1
            except hostlib.HostError:
                err += 1
        print("ok", ok, "err", err)
    """ % ROOT)
    r = subprocess.run([sys.executable, "-c", child] + paths, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.returncode, r.stderr[-1500:])
    ok, err = int(r.stdout.split()[1]), int(r.stdout.split()[3])
    assert ok + err == len(paths) and err > 0


def test_tools_survive_malformed_records(tmp_path):
    for tool in (CORRECT_BAM, INDEX_TOOL):
        if not os.path.exists(tool):
            from secphase_b200.build import build_host
            build_host()
    for p in mutants(tmp_path, 24, 777):
        r = subprocess.run([CORRECT_BAM, "-i", p, "-o", str(tmp_path / "o.bam"), "-m", "0", "-a", "0"],
                           capture_output=True, text=True, timeout=60)
        assert r.returncode in (0, 1), (p, r.returncode, r.stderr[-300:])      # a signal would be negative
        r = subprocess.run([INDEX_TOOL, "-i", p, "--stepSize", "2"], capture_output=True, text=True, timeout=60)
        assert r.returncode in (0, 1), (p, r.returncode, r.stderr[-300:])


def test_header_name_without_terminator(tmp_path):
    """A reference name with no NUL inside its l_name bytes (malformed header): correct_bam and the reader take
    exactly l_name bytes instead of reading on through l_ref and the records (run under ASan in a debug build;
    here: no crash, and the records still come out)."""
    import gzip
    s, b, _, _ = make_case("hifi", 4, locus_len=100000, len_mean=3000, len_sd=500, len_min=1500)
    src = str(tmp_path / "src.bam")
    hostlib.write_bam(src, s.names, s.lens, b)
    raw = bytearray(gzip.open(src, "rb").read())
    o = 8 + struct.unpack_from("<i", raw, 4)[0]
    n_ref = struct.unpack_from("<i", raw, o)[0]
    o += 4
    for _ in range(n_ref):
        l_name = struct.unpack_from("<i", raw, o)[0]
        assert raw[o + 4 + l_name - 1] == 0
        raw[o + 4 + l_name - 1] = ord("X")   # the terminator becomes part of the name
        o += 4 + l_name + 4
    p = str(tmp_path / "noterm.bam")
    bgzf_write(p, bytes(raw))
    if not os.path.exists(CORRECT_BAM):
        from secphase_b200.build import build_host
        build_host()
    r = subprocess.run([CORRECT_BAM, "-i", p, "-o", str(tmp_path / "o.bam"), "-m", "0", "-a", "0"],
                       capture_output=True, text=True, timeout=60)
    assert r.returncode in (0, 1), (r.returncode, r.stderr[-300:])
    with hostlib.BamReader(p, threads=1) as rd:
        assert all(n.endswith("X") and "\0" not in n for n in rd.names)
