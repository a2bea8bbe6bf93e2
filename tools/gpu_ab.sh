#!/bin/bash
# A/B of kernel variants on one box: parity tests first, then bench.py per variant.
#   gpurun --timeout 900 -- 'bash tools/gpu_ab.sh v8 "SECPHASE_B200_DPLANE=smem" "SECPHASE_B200_DPLANE=global"'
tag=$1; shift
out=gpurun_out; mkdir -p $out
export PYTHONUNBUFFERED=1
( timeout 600 python -m pytest tests -m gpu -x -q ) > $out/${tag}_pytest_gpu.log 2>&1
tail -3 $out/${tag}_pytest_gpu.log
i=0
for v in "$@"; do
  echo "== $v"
  ( env $v timeout 400 python bench.py --no-cpu-baseline --steps 24 ) > $out/${tag}_bench_$i.json 2> $out/${tag}_bench_$i.err
  python - <<PY
import json
try:
    d=json.load(open("$out/${tag}_bench_$i.json"))
    print("value %.0f e2e %.0f gcups_kernel %.1f hmm_ms %.3f frac %.3f" % (d["value"], d["e2e"]["value"], d["gcups_kernel"], d["roofline"]["kernel_ms_per_step"], d["roofline"]["frac"]), d["stage_ms_isolated"])
except Exception as e:
    print("failed", e); print(open("$out/${tag}_bench_$i.err").read()[-1500:])
PY
  i=$((i+1))
done
