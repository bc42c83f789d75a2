#!/bin/bash
# round-2 final evidence run (one B200): every step bounded, no pipes
mkdir -p gpurun_out/final; O=gpurun_out/final
SG2_BENCH_WATCHDOG=150 timeout 200 python bench.py > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "bench rc=$?"
timeout 200 python -m pytest tests -m gpu -x -q > $O/gputests.log 2>&1; echo "pytest rc=$?"; tail -1 $O/gputests.log
timeout 100 python tools/ncu_conv_labels.py run $O/conv_labels.json > $O/conv_labels.log 2>&1; echo "labels rc=$?"
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread \
  --clock-control none -k regex:"tile_conv|tile_wgrad|igemm" --csv --log-file $O/conv_ncu.csv \
  python tools/ncu_conv_labels.py run $O/conv_labels.json > $O/conv_ncu.log 2>&1; echo "conv ncu rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 6500 --csv --log-file $O/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-roofline > $O/launches_bench.log 2>&1; echo "launch list rc=$?"
timeout 100 python tools/step_trace.py $O/step_trace.json > $O/step_trace.txt 2>&1; echo "trace rc=$?"
SG2_BENCH_WATCHDOG=80 timeout 100 python bench.py --mode synth --batch 256 > $O/bench_cfg4_synth.json 2>/dev/null; echo "synth rc=$?"
