#!/bin/bash
mkdir -p gpurun_out/final2; O=gpurun_out/final2
timeout 150 python -m pytest tests -m gpu -x -q > $O/gputests.log 2>&1; echo "pytest rc=$?"; tail -1 $O/gputests.log
SG2_BENCH_WATCHDOG=60 timeout 100 python bench.py --no-cpu-baseline > $O/bench_cfg2.json 2> $O/bench.err; echo "bench rc=$?"
LOAD=1 WATCHDOG=12 timeout 60 python tools/stress_replay.py 2500 > $O/stress.txt 2>&1; echo "stress rc=$?"; tail -1 $O/stress.txt
