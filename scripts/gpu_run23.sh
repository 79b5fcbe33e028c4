#!/bin/bash
# new full-width / full-size Flux tests, launch list of the C2 retrieve job
mkdir -p gpurun_out
run() { local t=$1 name=$2; shift 2; local t0=$SECONDS; timeout $t "$@" > gpurun_out/r23_$name.log 2>&1; local rc=$?; echo "rc=$rc secs=$((SECONDS-t0))" >> gpurun_out/r23_$name.log; tail -n 8 gpurun_out/r23_$name.log | cut -c1-1500; return $rc; }
run 400 tests python -m pytest tests/test_flux_gpu.py tests/test_pipelines_gpu.py -m gpu -x -q -s -k "full or generator_list"
DRAG_BENCH_LAUNCH_LIST_ONLY=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 700 --csv \
  --log-file gpurun_out/r23_retrieve_launches.csv python bench.py --workload retrieve --steps 1 --warmup 1 > gpurun_out/r23_retrieve_launches_bench.log 2>&1
wc -l gpurun_out/r23_retrieve_launches.csv
exit 0
