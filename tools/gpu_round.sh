#!/bin/bash
# One GPU-box visit: parity tests, smoke, both bench arms, ncu launch list + one full capture of the
# HMM kernel, stage times of the ONT / stress configs and the file-to-file CLI run.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh v7'
# Everything lands in gpurun_out/<tag>_*; numbers printed under ncu are never bench values.
tag=${1:-vX}
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_smi.txt 2>&1
nproc >> $out/${tag}_smi.txt
( time timeout 600 python -m pytest tests -m gpu -x -q ) > $out/${tag}_pytest_gpu.log 2>&1
tail -3 $out/${tag}_pytest_gpu.log
( timeout 300 python __graft_entry__.py smoke ) > $out/${tag}_smoke.log 2>&1
tail -2 $out/${tag}_smoke.log
( timeout 600 python bench.py ) > $out/${tag}_bench.json 2> $out/${tag}_bench.err
tail -c 600 $out/${tag}_bench.json
( timeout 400 python bench.py --impl reference --steps 3 --warmup 1 ) > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err
cat $out/${tag}_bench_ref.json
for p in ont stress; do
  timeout 300 python tools/stage_bench.py --preset $p --groups 2048 >> $out/${tag}_stage.json 2>> $out/${tag}_stage.err
done
timeout 300 python tools/stage_bench.py --preset hifi --groups 4096 --write-qual >> $out/${tag}_stage.json 2>> $out/${tag}_stage.err
cat $out/${tag}_stage.json
# pipelined throughput of the other BASELINE configs (ONT, stress): bench.py --preset, 3 batches in flight
( timeout 300 python bench.py --preset ont --groups 8192 --locus-len 20000000 --steps 9 --warmup 3 --no-cpu-baseline ) > $out/${tag}_bench_ont.json 2> $out/${tag}_bench_ont.err
( timeout 300 python bench.py --preset stress --groups 8192 --locus-len 20000000 --steps 6 --warmup 3 --no-cpu-baseline ) > $out/${tag}_bench_stress.json 2> $out/${tag}_bench_stress.err
python - <<PY
import json
for n in ("ont", "stress"):
    try:
        d = json.load(open("$out/${tag}_bench_%s.json" % n))
        print(n, "value %.0f e2e %.0f gcups %.1f" % (d["value"], d["e2e"]["value"], d["gcups"]), d["config"]["sm_partition"])
    except Exception as e:
        print(n, "failed", e)
PY
( timeout 400 python tools/cli_bench.py --groups 32768 ) > $out/${tag}_cli.json 2> $out/${tag}_cli.err
tail -c 1500 $out/${tag}_cli.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_launch.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_hmm2 -s 6 -c 12 -f -o $out/${tag}_k_hmm2 \
  python tools/stage_bench.py --preset hifi --groups 8192 --locus-len 150000000 --iters 1 > $out/${tag}_ncu_full.log 2>&1
ls -la $out | tail -20
