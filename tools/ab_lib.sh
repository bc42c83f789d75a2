#!/bin/bash
# A/B of two builds of the library on ONE box: tools/ab_lib.sh base.so new.so [reps] [bench args]; restores new.so at the end
base=$1; new=$2; reps=${3:-2}; shift 3
L=speech-to-image-translation-without-text_b200/libsg2b200.so
for r in $(seq 1 $reps); do
  for w in new base; do
    if [ $w = new ]; then cp $new $L; else cp $base $L; fi
    timeout 150 python bench.py --no-cpu-baseline --no-roofline "$@" > gpurun_out/ab_tmp.json 2>gpurun_out/ab_tmp.err
    rc=$?
    ms=$(python -c "import json;print('%.3f'%json.load(open('gpurun_out/ab_tmp.json'))['ms_per_step'])" 2>/dev/null)
    echo "lib=$w rep$r rc=$rc ms/step=$ms" | tee -a gpurun_out/ab_log.txt
  done
done
cp $new $L
