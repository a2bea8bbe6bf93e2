"""Committed golden fixtures (tests/golden/*.npz, written by tools/make_golden.py from the
reference's own marker-path code): the oracle and the host simulation of the kernels (not gpu) and
the CUDA path through the C ABI (gpu) must all reproduce them bit for bit.  Nothing here reads
/root/reference."""
import glob
import os

import numpy as np
import pytest

from secphase_b200.flatbatch import _FIELDS, FlatBatch
from tools.parity import ASCII2CODE, compare_results

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
OUT_TABLES = ["groups", "scores", "extents", "blocks", "block_off", "markers_pre", "markers_pre_off",
              "markers_baq", "markers_baq_off", "markers_final", "markers_final_off"]


def load(path):
    z = np.load(path)
    batch = FlatBatch(**{f: z["in_" + f] for f, _ in _FIELDS})
    exp = {t: (z["out_" + t].view(np.float64) if t == "scores" else z["out_" + t]) for t in OUT_TABLES}
    exp["hmm"] = z["out_hmm"]
    exp["qual"] = z["out_qual"]
    lens = [int(x) for x in z["ref_lens"]]
    off = np.zeros(len(lens) + 1, np.int64)
    off[1:] = np.cumsum(lens)
    ascii_ = np.ascontiguousarray(z["ref_ascii"])
    return batch, exp, ascii_, off, [str(x) for x in z["ref_names"]], lens, str(z["preset"])


def test_fixtures_exist():
    assert len(GOLDEN) >= 5


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_reproduces_golden(oracle, path):
    batch, exp, ascii_, off, names, lens, preset = load(path)
    ptrs = [ascii_.ctypes.data + int(off[i]) for i in range(len(lens))]
    got = oracle.run(batch, oracle.preset_params(preset), oracle.make_refseq(names, ptrs, lens))
    bad = compare_results(exp, got, label="oracle")
    assert not bad, "\n".join(bad)
    assert np.array_equal(exp["hmm"], got["hmm"])  # incl. the hashes of every state[] / q[] array
    assert np.array_equal(exp["qual"], got["qual"])


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_kernel_logic_on_host_reproduces_golden(oracle, path):
    from tests.hostsim import pyhostsim
    batch, exp, ascii_, off, names, lens, preset = load(path)
    got = pyhostsim.run(batch, pyhostsim.params_from_oracle(oracle.preset_params(preset)), ASCII2CODE[ascii_], off)
    assert got["err"] == 0
    bad = compare_results(exp, got, label="hostsim")
    assert not bad, "\n".join(bad)
    # -w/--writeBam mode: the records' quality arrays
    full = pyhostsim.run(batch, pyhostsim.params_from_oracle(oracle.preset_params(preset)), ASCII2CODE[ascii_], off,
                         full_baq=True)
    assert np.array_equal(full["qual"], exp["qual"])
    assert not compare_results(exp, full, label="hostsim-full")


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_cuda_reproduces_golden(path):
    import secphase_b200
    batch, exp, ascii_, off, names, lens, preset = load(path)
    with secphase_b200.Secphase(preset) as eng:
        eng.set_reference_ascii([ascii_.ctypes.data + int(off[i]) for i in range(len(lens))], lens)
        got = eng.run_debug(batch)
        assert got["hmm_instances"] == len(exp["hmm"])
        cells = int(exp["hmm"][:, 4].astype(np.int64).sum() + (exp["hmm"][:, 5].astype(np.int64) << 31).sum())
        assert got["hmm_cells"] == cells
        bad = compare_results(exp, got, label="cuda")
        assert not bad, "\n".join(bad)
        # same answer with the pools in page-locked memory (zero-copy H2D path)
        eng.rng_seed(1)  # the tie-break stream continues across batches of a context: rewind it for the repeat
        again = eng.run(secphase_b200.pin_batch(batch))
        assert np.array_equal(again["scores"].view(np.int64), exp["scores"].view(np.int64))
        assert np.array_equal(again["groups"], exp["groups"])
        # -w/--writeBam mode: the records' quality arrays as the reference leaves them
        eng.set_write_qual(True)
        eng.rng_seed(1)
        full = eng.run(batch)
        assert np.array_equal(full["groups"], exp["groups"])
        assert np.array_equal(full["baq_qual"], exp["qual"])
        assert np.array_equal(full["scores"].view(np.int64), exp["scores"].view(np.int64))
