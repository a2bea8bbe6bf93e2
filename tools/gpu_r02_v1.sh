#!/bin/bash
# round 2, visit 1: baseline state of HEAD + the captures VERDICT r01 asked for before any kernel work
tag=r2v1
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_smi.txt 2>&1
nproc >> $out/${tag}_smi.txt
( time timeout 600 python -m pytest tests -m gpu -x -q ) > $out/${tag}_pytest_gpu.log 2>&1
tail -3 $out/${tag}_pytest_gpu.log
# integer stages on ONT / stress: full ncu of k_walk / k_group / k_emit (divergence, occupancy)
for p in ont stress; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_walk|k_group|k_emit' -s 3 -c 3 -f -o $out/${tag}_int_$p \
    python tools/stage_bench.py --preset $p --groups 2048 --iters 1 > $out/${tag}_ncu_int_$p.log 2>&1
done
# CLI with the own inflate decoder
( timeout 500 python tools/cli_bench.py --groups 32768 ) > $out/${tag}_cli.json 2> $out/${tag}_cli.err
tail -c 1200 $out/${tag}_cli.json
# sustained run (>= 30 s of FP64 load) with the clock record
( timeout 400 python bench.py --steps 2500 --warmup 3 --no-cpu-baseline ) > $out/${tag}_bench_sustained.json 2> $out/${tag}_bench_sustained.err
tail -c 800 $out/${tag}_bench_sustained.json
ls -la $out | tail
