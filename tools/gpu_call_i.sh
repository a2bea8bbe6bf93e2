#!/bin/bash
tag=${1:-v20}
out=gpurun_out; mkdir -p $out
export PYTHONUNBUFFERED=1
( timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 16 --warmup 3 ) > $out/${tag}_bench_n2.json 2> $out/${tag}_bench_n2.err
echo "stdout lines: $(wc -l < $out/${tag}_bench_n2.json)"; head -c 200 $out/${tag}_bench_n2.json; echo
( timeout 500 python bench.py ) > $out/${tag}_bench.json 2> $out/${tag}_bench.err
echo "stdout lines: $(wc -l < $out/${tag}_bench.json)"
python - <<PY
import json
d=json.load(open("$out/${tag}_bench.json"))
print("value %.0f e2e %.0f hmm_ms %.3f frac %.3f" % (d["value"], d["e2e"]["value"], d["roofline"]["kernel_ms_per_step"], d["roofline"]["frac"]), d["clocks"], d["cpu_baseline"]["value"])
PY
