#!/bin/bash
# Two-GPU check of the sharded path: bench.py under torchrun (both arms), then the CLI with --gpus 2.
#   gpurun --gpus 2 --timeout 700 -- 'bash tools/gpu_scale2.sh v8'
tag=${1:-vX}
n=${2:-2}
out=gpurun_out; mkdir -p $out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > $out/${tag}_n${n}_smi.txt 2>&1
( timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $n --steps 24 --warmup 3 --no-cpu-baseline ) > $out/${tag}_bench_n${n}.json 2> $out/${tag}_bench_n${n}.err
tail -c 1200 $out/${tag}_bench_n${n}.json; tail -5 $out/${tag}_bench_n${n}.err
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --impl reference --gpus $n --steps 2 --warmup 1 ) > $out/${tag}_bench_ref_n${n}.json 2> $out/${tag}_bench_ref_n${n}.err
cat $out/${tag}_bench_ref_n${n}.json
( timeout 300 python tools/cli_bench.py --groups 32768 --gpus $n ) > $out/${tag}_cli_n${n}.json 2> $out/${tag}_cli_n${n}.err
cat $out/${tag}_cli_n${n}.json; tail -3 $out/${tag}_cli_n${n}.err
