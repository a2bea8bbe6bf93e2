#!/bin/bash
tag=${1:-v17}
out=gpurun_out; mkdir -p $out
export PYTHONUNBUFFERED=1
i=0
for lib in "" "secphase_b200/lib/variants/lib_exact3.so"; do
  for rep in 1 2; do
  ( SECPHASE_B200_LIB=$lib timeout 600 python bench.py --no-cpu-baseline --steps 32 ) > $out/${tag}_bench_$i.json 2> $out/${tag}_bench_$i.err
  python - <<PY
import json
try:
    d=json.load(open("$out/${tag}_bench_$i.json"))
    print("lib='$lib' value %.0f e2e %.0f gcups_kernel %.1f hmm_ms %.3f frac %.3f issue %.3f" % (d["value"], d["e2e"]["value"], d["gcups_kernel"], d["roofline"]["kernel_ms_per_step"], d["roofline"]["frac"], d["roofline"]["issue_slot_frac"]))
except Exception as e:
    print("failed", e); print(open("$out/${tag}_bench_$i.err").read()[-1500:])
PY
  i=$((i+1))
  done
done
