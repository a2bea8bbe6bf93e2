"""`bin/secphase_index` (host only): <BAM>.secphase.index as programs/src/secphase_index.c writes it --
int64 count + BGZF virtual offsets -- against a Python restatement that walks the BGZF blocks itself."""
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

from secphase_b200 import hostlib
from secphase_b200.build import INDEX_TOOL
from tests.conftest import make_case


def bgzf_blocks(path):
    """[(file offset, inflated bytes)] of every BGZF block, and the file size."""
    raw = open(path, "rb").read()
    out, o = [], 0
    while o < len(raw):
        assert raw[o:o + 4] == b"\x1f\x8b\x08\x04"
        xlen = struct.unpack_from("<H", raw, o + 10)[0]
        extra = raw[o + 12:o + 12 + xlen]
        bsize, e = None, 0
        while e + 4 <= xlen:
            slen = struct.unpack_from("<H", extra, e + 2)[0]
            if extra[e:e + 2] == b"BC":
                bsize = struct.unpack_from("<H", extra, e + 4)[0] + 1
            e += 4 + slen
        data = zlib.decompress(raw[o + 12 + xlen:o + bsize - 8], -15)
        out.append((o, data))
        o += bsize
    return out, len(raw)


def expected_index(path, step):
    """secphase_index.c:60-118 restated, with htslib's bgzf_tell rule (a used-up block hands over to the next)."""
    blocks, fsize = bgzf_blocks(path)
    stream = b"".join(d for _, d in blocks)
    starts = np.cumsum([0] + [len(d) for _, d in blocks])   # uncompressed offset of every block

    def tell(u):   # virtual offset of uncompressed position u
        k = int(np.searchsorted(starts, u, side="right")) - 1
        while k < len(blocks) and u == starts[k + 1]:   # at a block end: the next block, offset 0 (skipping none: empty blocks too)
            k += 1
            if k == len(blocks):
                return fsize << 16
            if len(blocks[k][1]) > 0 or k == len(blocks) - 1:
                break
        if k >= len(blocks):
            return fsize << 16
        return (blocks[k][0] << 16) | (u - int(starts[k]))

    l_text = struct.unpack_from("<i", stream, 4)[0]
    o = 8 + l_text
    n_ref = struct.unpack_from("<i", stream, o)[0]
    o += 4
    for _ in range(n_ref):
        o += 4 + struct.unpack_from("<i", stream, o)[0] + 4
    addrs = [tell(o)]
    first, count, idx = None, 0, 1
    while o < len(stream):
        bs = struct.unpack_from("<i", stream, o)[0]
        rec = stream[o + 4:o + 4 + bs]
        name = rec[32:32 + rec[8]].split(b"\0")[0]
        o += 4 + bs
        if first is None:
            first = name
        if name != first:
            count += 1
        if count == idx * step:
            addrs.append(tell(o))
            idx += 1
    addrs.append(fsize << 16)
    return addrs


@pytest.mark.parametrize("step,args", [(10, []), (3, ["--stepSize", "3"]), (25, ["-s", "25"])])
def test_index_file_matches_the_reference_layout(tmp_path, step, args):
    if not os.path.exists(INDEX_TOOL):
        from secphase_b200.build import build_host
        build_host()
    s, b, _, _ = make_case("hifi", 60, locus_len=200000, len_mean=6000, len_sd=1500, len_min=2000)
    bam = str(tmp_path / "in.bam")
    hostlib.write_bam(bam, s.names, s.lens, b, level=1, threads=2)
    r = subprocess.run([INDEX_TOOL, "-i", bam] + args, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert "Writing index file" in r.stderr and "Done!" in r.stderr
    raw = open(bam + ".secphase.index", "rb").read()
    n = struct.unpack_from("<q", raw, 0)[0]
    got = list(struct.unpack_from(f"<{n}q", raw, 8))
    assert len(raw) == 8 * (n + 1)
    want = expected_index(bam, step)
    assert got == want
    assert got == sorted(got) and len(set(got)) == len(got)
    # every address but the last is the start of a record: the reader of the hot path, pointed at the
    # inflated stream there, must find a plausible block_size/refID pair
    blocks, fsize = bgzf_blocks(bam)
    by_addr = {a: d for a, d in blocks}
    for v in got[:-1]:
        data, off = by_addr[v >> 16], v & 0xffff
        if off + 8 <= len(data):
            bs, tid = struct.unpack_from("<ii", data, off)
            assert 32 < bs < 1 << 24 and 0 <= tid < len(s.names)
    assert got[-1] == fsize << 16


def test_index_usage_and_errors(tmp_path):
    if not os.path.exists(INDEX_TOOL):
        from secphase_b200.build import build_host
        build_host()
    r = subprocess.run([INDEX_TOOL, "-h"], capture_output=True, text=True)
    assert r.returncode == 1 and "Usage: secphase_index  -i <INPUT_BAM>" in r.stderr   # secphase_index.c:52-56
    r = subprocess.run([INDEX_TOOL, "-i", str(tmp_path / "nope.bam")], capture_output=True, text=True)
    assert r.returncode == 1 and "nope.bam" in r.stderr
    bad = tmp_path / "bad.bam"
    bad.write_bytes(b"\x1f\x8b\x08\x04" + b"\0" * 40)
    r = subprocess.run([INDEX_TOOL, "-i", str(bad)], capture_output=True, text=True)
    assert r.returncode == 1
