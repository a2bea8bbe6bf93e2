"""The reader's own DEFLATE decoder (secphase_b200/host/sph_inflate.cpp): a compiled sweep against zlib under
ASan/UBSan, and BAM round trips at every compression level with the fast path on and off."""
import os
import subprocess

import pytest

from secphase_b200 import hostlib
from secphase_b200.flatbatch import FlatBatch
from tests.conftest import make_case
from tests.test_host_ingest import assert_same_batch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_decoder_sweep_under_sanitizers(tmp_path):
    exe = str(tmp_path / "inflate_sweep")
    src = [os.path.join(ROOT, "tests", "inflate", "inflate_sweep.cpp"),
           os.path.join(ROOT, "secphase_b200", "host", "sph_inflate.cpp")]
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-omit-frame-pointer",
                           "-o", exe] + src + ["-lz", "-pthread"])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900, env=dict(os.environ, ASAN_OPTIONS="detect_leaks=0"))
    assert r.returncode == 0, (r.stdout[-800:], r.stderr[-2000:])
    assert "wrong 0" in r.stdout and "refused 0" in r.stdout, r.stdout
    assert "ERROR" not in r.stderr and "runtime error" not in r.stderr, r.stderr[-2000:]


@pytest.mark.parametrize("level", [0, 1, 4, 9])
def test_bam_round_trip_fast_and_zlib_paths(tmp_path, level):
    s, b, _, _ = make_case("hifi", 40, locus_len=200000, len_mean=6000, len_sd=1500, len_min=2000)
    p = str(tmp_path / "t.bam")
    hostlib.write_bam(p, s.names, s.lens, b, level=level, threads=2)
    child = ("import sys; sys.path.insert(0, %r)\n"
             "from secphase_b200 import hostlib\n"
             "from secphase_b200.flatbatch import FlatBatch\n"
             "import numpy as np\n"
             "with hostlib.BamReader(%r, threads=3) as r:\n"
             "    fb = FlatBatch.concat(list(r))\n"
             "np.savez(%r, **{k: getattr(fb, k) for k in ('flag', 'pos', 'cigar_pool', 'tag_pool', 'seq_pool', 'qual_pool', 'qual_off')})\n")
    outs = []
    for env in ({}, {"SPH_ZLIB_INFLATE": "1"}):
        out = str(tmp_path / ("o%d.npz" % len(outs)))
        subprocess.check_call([os.sys.executable, "-c", child % (ROOT, p, out)], env=dict(os.environ, **env), timeout=300)
        outs.append(out)
    import numpy as np
    a, z = np.load(outs[0]), np.load(outs[1])
    for k in a.files:
        assert np.array_equal(a[k], z[k]), k
    with hostlib.BamReader(p, threads=2) as r:
        assert_same_batch(FlatBatch.concat(list(r)), b)
