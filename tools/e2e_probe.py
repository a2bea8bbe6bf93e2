#!/usr/bin/env python
"""End-to-end (sp_submit with host buffers) step time vs batch size, with and without the in-place quality reads.
  python tools/e2e_probe.py [--preset hifi] [--groups 2048 8192] """
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--preset", default="hifi")
    ap.add_argument("--groups", type=int, nargs="+", default=[2048, 8192])
    ap.add_argument("--locus-len", type=int, default=20_000_000)
    ap.add_argument("--steps", type=int, default=12)
    args = ap.parse_args()
    import secphase_b200
    from tools.parity import encode_reference
    from tools.synth.pysynth import Synth, default_cfg
    s = Synth(default_cfg(args.preset, locus_len=args.locus_len, seed=20240603))
    codes, off = encode_reference(s)
    for g in args.groups:
        batches = [secphase_b200.pin_batch(s.generate(i * g, g)) for i in range(3)]
        for zc in (1, 0):
            if zc: os.environ.pop("SECPHASE_B200_NO_ZERO_COPY", None)
            else: os.environ["SECPHASE_B200_NO_ZERO_COPY"] = "1"
            with secphase_b200.Secphase("ont" if args.preset == "ont" else "hifi") as eng:
                eng.set_reference_codes(codes, off)
                def run(n):
                    out, infl = [], []
                    for i in range(n):
                        sl = i % 3
                        if len(infl) == 3: out.append(eng.wait(infl.pop(0)))
                        eng.submit(batches[i % 3], slot=sl); infl.append(sl)
                    while infl: out.append(eng.wait(infl.pop(0)))
                    return out
                run(4)
                reps = []
                for _ in range(3):
                    t0 = time.perf_counter(); st = run(args.steps); reps.append((time.perf_counter() - t0) / args.steps * 1e3)
                ms = np.mean([x["ms_stage"] for x in st], axis=0)
                print(json.dumps({"preset": args.preset, "groups": g, "zero_copy": bool(zc), "ms_per_step": [round(x, 2) for x in reps],
                                  "groups_per_s": round(g / (min(reps) * 1e-3)), "h2d_bytes": st[0]["h2d_bytes"],
                                  "stage_ms": dict(zip(["h2d", "walk", "group", "emit_sort", "hmm", "score", "d2h"], [round(float(x), 3) for x in ms[:7]]))}))


if __name__ == "__main__":
    main()
