#!/usr/bin/env python
"""Writes tests/golden/*.npz: inputs AND outputs of the CPU oracle built from the reference's own
marker-path sources (oracle/_ref, kind "reference") on small seeded synthetic cases.  Needs
/root/reference (this container); the fixtures are what travels.  Re-run only when the oracle or the
case list changes:   python tools/make_golden.py
Each file is self-contained: the read groups (sp_flat_batch arrays), the assembly (ASCII), the
parameter preset, and every table the oracle records (markers per stage, consensus blocks, HMM
instances with hashes of state[]/q[], score bit patterns, selected index)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from secphase_b200.flatbatch import _FIELDS  # noqa: E402

CASES = [
    # name, synth preset, params preset, groups, synth overrides
    ("hifi", "hifi", "hifi", 24, dict(locus_len=60000, len_mean=9000, len_sd=2000, len_min=3000)),
    ("ont", "ont", "ont", 10, dict(locus_len=60000, len_mean=7000, len_sd=2500, len_min=2000)),
    ("stress", "stress", "hifi", 6, dict(locus_len=60000, len_mean=7000, len_sd=2000, len_min=3000)),
    ("hifi_eqx_clip_N", "hifi", "hifi", 16,
     dict(locus_len=60000, len_mean=8000, len_sd=2000, len_min=3000, eqx=1, n_rate=2e-4, clip_prob=0.8)),
    ("ont_hardclip_md", "ont", "ont", 10,
     dict(locus_len=60000, len_mean=5000, len_sd=2000, len_min=1500, clip_prob=0.9, hard_clip_prob=0.9, use_md=1)),
    # near-identical repeat copies: several secondaries share the top score, so get_best_record_index draws
    # rand() % count (ptAlignment.c:156-170) -- pins the tie-break replay
    ("stress_ties", "stress", "hifi", 12,
     dict(locus_len=60000, len_mean=7000, len_sd=2000, len_min=3000, snv_rate=1e-4, indel_rate=2e-5, long_indel_rate=2e-6)),
]
OUT_TABLES = ["groups", "scores", "extents", "blocks", "block_off", "hmm", "markers_pre", "markers_pre_off",
              "markers_baq", "markers_baq_off", "markers_final", "markers_final_off",
              "qual"]  # qual: every record's quality array as the job leaves it (the -w/--writeBam output)
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def main():
    from oracle import pyoracle
    from tools.synth.pysynth import Synth, default_cfg
    if "reference" not in pyoracle.available_kinds():
        pyoracle.build()
    assert "reference" in pyoracle.available_kinds(), "oracle/_ref needs /root/reference"
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    only = set(sys.argv[1:])  # `python tools/make_golden.py <name>...` (re)writes just those fixtures
    for name, spreset, ppreset, ng, over in CASES:
        if only and name not in only:
            continue
        s = Synth(default_cfg(spreset, **over))
        b = s.generate(0, ng)
        ref = pyoracle.make_refseq(s.names, [s.contig_ptr(i) for i in range(s.n_contigs)], s.lens)
        r = pyoracle.run(b, pyoracle.preset_params(ppreset), ref, kind="reference")
        data = {"in_" + f: getattr(b, f) for f, _ in _FIELDS}
        data["ref_ascii"] = np.concatenate([s.contig_ascii(i) for i in range(s.n_contigs)])
        data["ref_lens"] = np.array(s.lens, np.int64)
        data["ref_names"] = np.array(s.names)
        data["preset"] = np.array(ppreset)
        for t in OUT_TABLES:
            data["out_" + t] = r[t].view(np.int64) if t == "scores" else r[t]
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **data)
        swaps = int((r["groups"][:, 0] != r["groups"][:, 1]).sum())
        print(f"{name}: {ng} groups, {len(r['hmm'])} HMM instances, {len(r['markers_final'])} final markers, "
              f"{swaps} swaps -> {os.path.getsize(path) / 1024:.0f} KB")


if __name__ == "__main__":
    main()
