#!/bin/bash
# GPU visit A of this session: parity tests (incl. the -w/--writeBam mode), stage times of the
# write-qual mode, one source-level ncu capture of the unrolled HMM kernel.
tag=${1:-v9}
out=gpurun_out; mkdir -p $out
export PYTHONUNBUFFERED=1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $out/${tag}_pytest_gpu.log 2>&1
tail -5 $out/${tag}_pytest_gpu.log
for p in hifi ont; do
  timeout 300 python tools/stage_bench.py --preset $p --groups 1024 --write-qual >> $out/${tag}_stage_wq.json 2>> $out/${tag}_stage_wq.err
  timeout 300 python tools/stage_bench.py --preset $p --groups 1024 >> $out/${tag}_stage_wq.json 2>> $out/${tag}_stage_wq.err
done
cat $out/${tag}_stage_wq.json; tail -3 $out/${tag}_stage_wq.err
( timeout 300 python tools/cli_bench.py --groups 8192 --extra=-w --repeat 1 ) > $out/${tag}_cli_w.json 2> $out/${tag}_cli_w.err
cat $out/${tag}_cli_w.json; tail -3 $out/${tag}_cli_w.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_hmm2 -s 9 -c 4 -f -o $out/${tag}_k_hmm2_src \
  python tools/stage_bench.py --preset hifi --groups 4096 --locus-len 20000000 --iters 1 > $out/${tag}_ncu_src.log 2>&1
tail -3 $out/${tag}_ncu_src.log
ls -la $out | tail
