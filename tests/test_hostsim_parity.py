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
