#!/bin/bash
# GPU visit B: parity (incl. the lane-interleaved -w layout and its fallbacks), -w stage times, main-path regression check
tag=${1:-v10}
out=gpurun_out; mkdir -p $out
export PYTHONUNBUFFERED=1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $out/${tag}_pytest_gpu.log 2>&1
tail -5 $out/${tag}_pytest_gpu.log
for a in "--preset hifi --groups 1024" "--preset hifi --groups 4096" "--preset ont --groups 1024" "--preset stress --groups 512"; do
  timeout 300 python tools/stage_bench.py $a --write-qual >> $out/${tag}_stage_wq.json 2>> $out/${tag}_stage_wq.err
done
SECPHASE_B200_NO_INTERLEAVE=1 timeout 300 python tools/stage_bench.py --preset hifi --groups 1024 --write-qual >> $out/${tag}_stage_wq.json 2>> $out/${tag}_stage_wq.err
cat $out/${tag}_stage_wq.json; tail -3 $out/${tag}_stage_wq.err
( timeout 300 python tools/cli_bench.py --groups 8192 --extra=-w --repeat 1 ) > $out/${tag}_cli_w.json 2> $out/${tag}_cli_w.err
cat $out/${tag}_cli_w.json; tail -3 $out/${tag}_cli_w.err
( timeout 600 python bench.py --no-cpu-baseline --steps 24 ) > $out/${tag}_bench.json 2> $out/${tag}_bench.err
python - <<PY
import json
try:
    d=json.load(open("$out/${tag}_bench.json"))
    print("value %.0f e2e %.0f gcups_kernel %.1f hmm_ms %.3f frac %.3f traffic %s" % (d["value"], d["e2e"]["value"], d["gcups_kernel"], d["roofline"]["kernel_ms_per_step"], d["roofline"]["frac"], d["roofline"]["traffic"]))
except Exception as e:
    print("failed", e); print(open("$out/${tag}_bench.err").read()[-1500:])
PY
