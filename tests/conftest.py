import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def oracle():
    """CPU checker: the reference's own marker-path code when oracle/_ref was built, else the port."""
    from oracle import pyoracle
    if not pyoracle.available_kinds():
        try:
            pyoracle.build()
        except Exception as e:  # pragma: no cover
            pytest.skip(f"oracle could not be built: {e}")
    if not pyoracle.available_kinds():
        pytest.skip("no oracle library available")
    return pyoracle


def make_case(preset, n_groups, first=0, **cfg_over):
    """(synth, batch, ref codes, contig offsets) for a seeded synthetic case."""
    from tools.parity import encode_reference
    from tools.synth.pysynth import Synth, default_cfg
    cfg = default_cfg(preset, **cfg_over)
    s = Synth(cfg)
    b = s.generate(first, n_groups)
    codes, off = encode_reference(s)
    return s, b, codes, off


def oracle_refseq(pyoracle, s):
    return pyoracle.make_refseq(s.names, [s.contig_ptr(i) for i in range(s.n_contigs)], s.lens)


CASES = [
    # name, synth preset, params preset, groups, synth overrides
    ("hifi", "hifi", "hifi", 120, dict(locus_len=300000)),
    ("ont", "ont", "ont", 40, dict(locus_len=400000)),
    ("stress", "stress", "hifi", 24, dict(locus_len=300000)),
    ("hifi_eqx_clip_N", "hifi", "hifi", 80, dict(locus_len=300000, eqx=1, n_rate=1e-4, clip_prob=0.8)),
    ("ont_short_hardclip", "ont", "ont", 60,
     dict(locus_len=300000, len_mean=8000, len_sd=3000, len_min=1500, clip_prob=0.9, hard_clip_prob=0.9)),
    ("hifi_md", "hifi", "hifi", 60, dict(locus_len=300000, use_md=1)),
    # several secondaries tie at the top score: the rand() tie-break (ptAlignment.c:156-170) is exercised
    ("stress_ties", "stress", "hifi", 48, dict(locus_len=300000, snv_rate=1e-4, indel_rate=2e-5, long_indel_rate=2e-6)),
]


def count_top_score_ties(batch, scores):
    """Read groups in which more than one secondary holds the maximum secondary score."""
    t = 0
    for g in range(batch.n_groups):
        a0, a1 = int(batch.grp_aln_off[g]), int(batch.grp_aln_off[g + 1])
        sec = [scores[a] for a in range(a0, a1) if batch.flag[a] & 256]
        if len(sec) > 1 and sec.count(max(sec)) > 1:
            t += 1
    return t
