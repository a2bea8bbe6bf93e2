#!/bin/bash
tag=${1:-v15}
out=gpurun_out; mkdir -p $out
export PYTHONUNBUFFERED=1
( timeout 600 python bench.py --no-cpu-baseline --steps 32 ) > $out/${tag}_bench.json 2> $out/${tag}_bench.err
python - <<PY
import json
try:
    d=json.load(open("$out/${tag}_bench.json"))
    print("value %.0f e2e %.0f gcups_kernel %.1f hmm_ms %.3f frac %.3f issue %.3f" % (d["value"], d["e2e"]["value"], d["gcups_kernel"], d["roofline"]["kernel_ms_per_step"], d["roofline"]["frac"], d["roofline"]["issue_slot_frac"]), d["stage_ms_isolated"])
except Exception as e:
    print("failed", e); print(open("$out/${tag}_bench.err").read()[-1500:])
PY
