#!/bin/bash
# exact unrolled classes for bw 23..25: parity + bench A/B
tag=${1:-v14}
out=gpurun_out; mkdir -p $out
export PYTHONUNBUFFERED=1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $out/${tag}_pytest_gpu.log 2>&1
tail -4 $out/${tag}_pytest_gpu.log
( timeout 600 python bench.py --no-cpu-baseline --steps 32 ) > $out/${tag}_bench.json 2> $out/${tag}_bench.err
python - <<PY
import json
try:
    d=json.load(open("$out/${tag}_bench.json"))
    print("value %.0f e2e %.0f gcups_kernel %.1f hmm_ms %.3f frac %.3f issue %.3f" % (d["value"], d["e2e"]["value"], d["gcups_kernel"], d["roofline"]["kernel_ms_per_step"], d["roofline"]["frac"], d["roofline"]["issue_slot_frac"]), d["stage_ms_isolated"])
except Exception as e:
    print("failed", e); print(open("$out/${tag}_bench.err").read()[-1500:])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_hmm -s 12 -c 9 --csv --log-file $out/${tag}_hmm_launches.csv python tools/stage_bench.py --preset hifi --groups 8192 --locus-len 150000000 --iters 1 > $out/${tag}_l.log 2>&1
grep -o "k_hmm2<[^>]*>.*" $out/${tag}_hmm_launches.csv | sed 's/(con.*"(/ (/' | cut -c1-120
