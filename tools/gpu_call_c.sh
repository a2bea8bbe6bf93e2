#!/bin/bash
tag=${1:-v11}
out=gpurun_out; mkdir -p $out
export PYTHONUNBUFFERED=1
( time timeout 600 python -m pytest tests -m gpu -x -q -k "write_qual or write_bam or hmm_all_rows" ) > $out/${tag}_pytest_gpu.log 2>&1
tail -4 $out/${tag}_pytest_gpu.log
for a in "--preset hifi --groups 1024" "--preset hifi --groups 4096" "--preset ont --groups 1024" "--preset stress --groups 512"; do
  timeout 300 python tools/stage_bench.py $a --write-qual >> $out/${tag}_stage_wq.json 2>> $out/${tag}_stage_wq.err
done
cat $out/${tag}_stage_wq.json; tail -3 $out/${tag}_stage_wq.err
