#!/bin/bash
# One GPU-box visit: `gpurun --timeout T -- 'bash tools/gpu_visit.sh <tag> <step>...'`
# steps: test smoke bench ref ncu_launch ncu_hmm cli cli2 stage
tag=$1; shift
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_smi.txt 2>&1
nproc >> $out/${tag}_smi.txt
for step in "$@"; do
case $step in
test)
  ( time timeout 900 python -m pytest tests -m gpu -x -q ) > $out/${tag}_pytest_gpu.log 2>&1
  tail -5 $out/${tag}_pytest_gpu.log
  grep -E "^E  |Error|FAILED" $out/${tag}_pytest_gpu.log | head -20 ;;
smoke)
  ( timeout 300 python __graft_entry__.py smoke ) > $out/${tag}_smoke.log 2>&1
  tail -2 $out/${tag}_smoke.log ;;
bench)
  ( time timeout 900 python bench.py ) > $out/${tag}_bench.json 2> $out/${tag}_bench.err
  tail -3 $out/${tag}_bench.err
  python - <<PY
import json
try:
    d = json.load(open("$out/${tag}_bench.json"))
    r = d["roofline"]
    print("value %.0f e2e %.0f kernel_ms %.2f frac %.3f mode %s rerun %s" % (d["value"], d["e2e"]["value"], r["kernel_ms_per_step"], r["frac"], r.get("hmm_mode"), r.get("strict_rerun_instances_per_step")))
    print("parity", d.get("parity_vs_cpu_reference"))
    for k, v in d.get("per_config", {}).items():
        if "error" in v: print(k, v); continue
        p = v.get("parity_vs_cpu_reference") or {}
        print(k, "value %.0f e2e %.0f frac %.3f kernel_ms %.2f wall %.0fs" % (v["value"], v["e2e"]["value"], v["roofline"]["frac"], v["roofline"]["kernel_ms_per_step"], v["wall_s"]), v["stage_ms_isolated"], "all_equal", p.get("all_equal"), "ties", p.get("groups_with_tied_top_secondaries"), p.get("mismatches"))
except Exception as e:
    print("bench parse failed", e)
PY
  ;;
benchq)  # the bench line only (no CPU legs)
  ( timeout 600 python bench.py --no-cpu-baseline ) > $out/${tag}_benchq.json 2> $out/${tag}_benchq.err
  tail -c 900 $out/${tag}_benchq.json ;;
benchm)  # the bench line only, one fast launch for all classes
  ( SECPHASE_B200_HMM_MERGE=1 timeout 600 python bench.py --no-cpu-baseline ) > $out/${tag}_benchm.json 2> $out/${tag}_benchm.err
  tail -c 900 $out/${tag}_benchm.json ;;
ref)
  ( timeout 400 python bench.py --impl reference --steps 3 --warmup 1 ) > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err
  cat $out/${tag}_bench_ref.json ;;
stage)
  for m in fast strict; do
    for p in ont stress; do
      timeout 300 python tools/stage_bench.py --preset $p --groups 2048 --hmm $m >> $out/${tag}_stage.json 2>> $out/${tag}_stage.err
    done
    timeout 300 python tools/stage_bench.py --preset hifi --groups 8192 --locus-len 150000000 --hmm $m >> $out/${tag}_stage.json 2>> $out/${tag}_stage.err
  done
  cat $out/${tag}_stage.json ;;
prof)  # clock64 split of k_group / k_emit (lib/variants/lib_prof.so = -DSP_PROFILE_GROUP)
  for p in stress ont; do for v in ${PROF_LIBS:-prof}; do
    echo "variant $v preset $p" >> $out/${tag}_prof.err
    SECPHASE_B200_LIB=secphase_b200/lib/variants/lib_$v.so timeout 300 python tools/stage_bench.py --preset $p --groups 2048 --iters 3 >> $out/${tag}_prof.json 2>> $out/${tag}_prof.err
  done; done
  cat $out/${tag}_prof.err | tail -24; cut -c1-330 $out/${tag}_prof.json ;;
ontsms)  # bench line of one config (PRESET, default ont) vs the SM split (integer stages | HMM)
  for n in ${SMS_LIST:-16 24 32 40}; do
    ( SECPHASE_B200_INT_SMS=$n timeout 300 python bench.py --preset ${PRESET:-ont} --no-cpu-baseline --no-per-config --steps ${STEPS:-6} ) >> $out/${tag}_ontsms.json 2>> $out/${tag}_ontsms.err
  done
  python - <<EOF
import json
for l in open('$out/${tag}_ontsms.json'):
    try: d=json.loads(l)
    except Exception: continue
    print(d['config'].get('sm_partition'), round(d['value']), round(d['e2e']['value']), d['roofline']['kernel_ms_per_step'], d.get('stage_ms_isolated'))
EOF
  ;;
stagew)  # fast-arithmetic stage times with a warp per alignment in K1 for every alignment (default: only long op tables)
  for p in ont stress; do
    SECPHASE_B200_WALK=warp timeout 300 python tools/stage_bench.py --preset $p --groups 2048 >> $out/${tag}_stagew.json 2>> $out/${tag}_stagew.err
    timeout 300 python tools/stage_bench.py --preset $p --groups 2048 >> $out/${tag}_stagew.json 2>> $out/${tag}_stagew.err
  done
  SECPHASE_B200_WALK=warp timeout 300 python tools/stage_bench.py --preset hifi --groups 8192 --locus-len 150000000 >> $out/${tag}_stagew.json 2>> $out/${tag}_stagew.err
  timeout 300 python tools/stage_bench.py --preset hifi --groups 8192 --locus-len 150000000 >> $out/${tag}_stagew.json 2>> $out/${tag}_stagew.err
  timeout 300 python tools/stage_bench.py --preset ont --groups 10240 --locus-len 5000000 >> $out/${tag}_stagew.json 2>> $out/${tag}_stagew.err
  cut -c1-400 $out/${tag}_stagew.json ;;
stagehifi)  # HiFi bench-size stage times: fast (default lib and every lib under lib/variants), then strict
  for lib in "" secphase_b200/lib/variants/*.so; do
    [ -n "$lib" ] && [ ! -f "$lib" ] && continue
    echo "lib=${lib:-default}" >> $out/${tag}_stagehifi.json
    SECPHASE_B200_LIB=$lib timeout 300 python tools/stage_bench.py --preset hifi --groups 8192 --locus-len 150000000 --hmm fast >> $out/${tag}_stagehifi.json 2>> $out/${tag}_stagehifi.err
  done
  timeout 300 python tools/stage_bench.py --preset hifi --groups 8192 --locus-len 150000000 --hmm strict >> $out/${tag}_stagehifi.json 2>> $out/${tag}_stagehifi.err
  cat $out/${tag}_stagehifi.json ;;
traffic)  # DRAM bytes of the HMM launch set of one step (tools/traffic_from_ncu.py turns it into profiles/r02_traffic_k_hmm.json)
  timeout 500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_hmm --csv \
    --log-file $out/${tag}_traffic.csv python tools/stage_bench.py --preset hifi --groups 8192 --locus-len 150000000 --iters 1 \
    > $out/${tag}_traffic_stage.json 2> $out/${tag}_traffic.err
  tail -c 300 $out/${tag}_traffic_stage.json ;;
ontsize)  # ONT stage times vs batch size (thread-per-alignment integer stages want many alignments in flight)
  for g in 4096 10240 20480; do
    timeout 300 python tools/stage_bench.py --preset ont --groups $g --locus-len 5000000 >> $out/${tag}_ontsize.json 2>> $out/${tag}_ontsize.err
  done
  cut -c1-330 $out/${tag}_ontsize.json ;;
ontlaunch)  # per-class HMM launch times of one ONT batch
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_hmm --csv --log-file $out/${tag}_ont_hmm_launches.csv \
    python tools/stage_bench.py --preset ont --groups 10240 --locus-len 5000000 --iters 1 > $out/${tag}_ont_l.log 2>&1
  tail -c 400 $out/${tag}_ont_l.log ;;
cli)
  ( timeout 500 python tools/cli_bench.py --groups 32768 ) > $out/${tag}_cli.json 2> $out/${tag}_cli.err
  tail -c 1200 $out/${tag}_cli.json ;;
ncu_launch)
  timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_launch.log 2>&1 ;;
ncu_hmm)  # launch list of one stage-bench run, then a full capture of the bulk class kernels of the 3rd batch (<= 64 MiB come back)
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_hmm --csv --log-file $out/${tag}_hmm_launches.csv \
    python tools/stage_bench.py --preset hifi --groups 8192 --locus-len 150000000 --iters 1 > $out/${tag}_ncu_l.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_hmmf|k_hmm2' -s 16 -c 8 -f -o $out/${tag}_k_hmm \
    python tools/stage_bench.py --preset hifi --groups 8192 --locus-len 150000000 --iters 1 > $out/${tag}_ncu_full.log 2>&1
  ls -la $out/${tag}_k_hmm.ncu-rep ;;
ncu_int)
  # (4 matching launches per batch: k_walk_warp, k_walk_list, k_group_lanes<W> | k_group, k_emit; the third batch is captured)
  for p in ${NCU_PRESETS:-ont stress}; do
    timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_walk|k_group_lanes|k_group$|k_emit' -s 8 -c 4 -f -o $out/${tag}_int_$p \
      python tools/stage_bench.py --preset $p --groups ${NCU_GROUPS:-2048} --iters 1 > $out/${tag}_ncu_int_$p.log 2>&1
  done ;;
*) echo "unknown step $step" ;;
esac
done
ls -la $out | grep ${tag}_ | tail -20
