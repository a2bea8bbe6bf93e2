#!/bin/bash
# Pipelined throughput of the ONT and stress configs (3 batches in flight), with batch-size sweep for ONT
tag=${1:-v12}
out=gpurun_out; mkdir -p $out
export PYTHONUNBUFFERED=1
for g in 2048 8192; do
  ( timeout 400 python bench.py --preset ont --groups $g --locus-len 20000000 --steps 9 --warmup 3 --no-cpu-baseline ) > $out/${tag}_bench_ont_$g.json 2> $out/${tag}_bench_ont_$g.err
  python - <<PY
import json
try:
    d=json.load(open("$out/${tag}_bench_ont_$g.json"))
    print("ont $g: value %.0f e2e %.0f gcups %.1f gcups_kernel %.1f" % (d["value"], d["e2e"]["value"], d["gcups"], d["gcups_kernel"]), d["stage_ms_isolated"])
except Exception as e:
    print("failed", e); print(open("$out/${tag}_bench_ont_$g.err").read()[-1500:])
PY
done
( timeout 400 python bench.py --preset stress --groups 2048 --locus-len 20000000 --steps 9 --warmup 3 --no-cpu-baseline ) > $out/${tag}_bench_stress.json 2> $out/${tag}_bench_stress.err
python - <<PY
import json
try:
    d=json.load(open("$out/${tag}_bench_stress.json"))
    print("stress: value %.0f e2e %.0f gcups %.1f gcups_kernel %.1f" % (d["value"], d["e2e"]["value"], d["gcups"], d["gcups_kernel"]), d["stage_ms_isolated"])
except Exception as e:
    print("failed", e); print(open("$out/${tag}_bench_stress.err").read()[-1500:])
PY
( timeout 300 python bench.py --preset ont --groups 1024 --locus-len 20000000 --impl reference --steps 2 --warmup 1 ) > $out/${tag}_bench_ont_ref.json 2>/dev/null
cat $out/${tag}_bench_ont_ref.json | cut -c1-300
