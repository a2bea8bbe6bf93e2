#!/usr/bin/env python
"""Clean (one batch in flight, nothing overlapped) per-stage device times of the pipeline on one
resident synthetic batch: the number to read when tuning a kernel.  GPU only.

  python tools/stage_bench.py [--preset hifi|ont|stress] [--groups N] [--iters K] [--locus-len L]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--preset", default="hifi")
    ap.add_argument("--groups", type=int, default=4096)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--locus-len", type=int, default=20_000_000)
    ap.add_argument("--write-qual", action="store_true", help="-w/--writeBam mode: BAQ of every window base")
    ap.add_argument("--hmm", default=None, choices=["fast", "strict"])
    args = ap.parse_args()
    import secphase_b200
    from tools.parity import encode_reference
    from tools.synth.pysynth import Synth, default_cfg
    s = Synth(default_cfg(args.preset, locus_len=args.locus_len, seed=20240603))
    codes, off = encode_reference(s)
    batch = s.generate(0, args.groups)
    ppreset = "ont" if args.preset == "ont" else "hifi"
    with secphase_b200.Secphase(ppreset) as eng:
        eng.set_reference_codes(codes, off)
        if args.hmm:
            eng.set_hmm_mode(args.hmm)
        if args.write_qual:
            eng.set_write_qual(True)
        eng.upload(batch, 0)
        st = []
        for i in range(args.iters + 2):
            eng.run_resident(0)
            r = eng.wait(0, copy=False)
            if i >= 2:
                st.append(r)
        ms = np.mean([x["ms_stage"] for x in st], axis=0)
        cells = st[0]["hmm_cells"]
        out = {"preset": args.preset, "write_qual": bool(args.write_qual), "groups": args.groups, "hmm_mode": st[0]["hmm_mode"], "hmm_rerun": st[0]["hmm_rerun"], "hmm_instances": st[0]["hmm_instances"], "cells": cells,
               "stage_ms": dict(zip(["h2d", "walk", "group", "emit_sort", "hmm", "score", "d2h"], [round(float(x), 3) for x in ms[:7]])),
               "hmm_gcups": cells / (ms[4] * 1e-3) / 1e9, "total_ms": float(np.mean([x["ms_total"] for x in st])),
               "groups_per_s_serial": args.groups / (float(np.mean([x["ms_total"] for x in st])) * 1e-3)}
        print(json.dumps(out))


if __name__ == "__main__":
    main()
