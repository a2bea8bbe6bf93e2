#!/bin/bash
# SM-partition sweep for the integer-heavy configs (ONT, stress): SECPHASE_B200_INT_SMS
tag=${1:-v13}
out=gpurun_out; mkdir -p $out
export PYTHONUNBUFFERED=1
run() { # name preset groups int_sms
  ( SECPHASE_B200_INT_SMS=$4 timeout 400 python bench.py --preset $2 --groups $3 --locus-len 20000000 --steps 9 --warmup 3 --no-cpu-baseline ) > $out/${tag}_$1.json 2> $out/${tag}_$1.err
  python - <<PY
import json
try:
    d=json.load(open("$out/${tag}_$1.json"))
    print("$1: value %.0f e2e %.0f gcups %.1f gcups_kernel %.1f" % (d["value"], d["e2e"]["value"], d["gcups"], d["gcups_kernel"]), d["config"]["sm_partition"], {k: round(v,1) for k,v in d["stage_ms_isolated"].items()})
except Exception as e:
    print("$1 failed", e); print(open("$out/${tag}_$1.err").read()[-800:])
PY
}
run ont_8 ont 8192 8
run ont_24 ont 8192 24
run ont_48 ont 8192 48
run ont_0 ont 8192 0
run stress_24 stress 2048 24
run stress_48 stress 2048 48
run hifi_16 hifi 8192 16
