#!/bin/bash
# A/B of one environment switch on ONE box (boxes differ by 1-2 %): tools/ab.sh NAME "v1 v2 ..." [reps] [extra bench args]
# every run under its own timeout; prints ms/step per run
name=$1; vals=$2; reps=${3:-2}; shift 3
mkdir -p gpurun_out
for r in $(seq 1 $reps); do
  for v in $vals; do
    env $name=$v timeout 150 python bench.py --no-cpu-baseline --no-roofline "$@" > gpurun_out/ab_tmp.json 2>gpurun_out/ab_tmp.err
    rc=$?
    ms=$(python -c "import json;print('%.3f'%json.load(open('gpurun_out/ab_tmp.json'))['ms_per_step'])" 2>/dev/null)
    echo "$name=$v rep$r rc=$rc ms/step=$ms" | tee -a gpurun_out/ab_log.txt
  done
done
