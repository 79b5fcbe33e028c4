#!/bin/bash
# batch experiment (C4 per-GPU slice as one batch of 4 vs 2 vs 1), batched-fill parity test, C2 retrieve bench
mkdir -p gpurun_out
run() { local t=$1 name=$2; shift 2; local t0=$SECONDS; timeout $t "$@" > gpurun_out/r16_$name.log 2>&1; local rc=$?; echo "rc=$rc secs=$((SECONDS-t0))" >> gpurun_out/r16_$name.log; tail -n 6 gpurun_out/r16_$name.log | cut -c1-2500; return $rc; }
run 300 tests python -m pytest tests/test_pipelines_gpu.py -m gpu -x -q
run 400 bench_b4 python bench.py --steps 2 --warmup 3 --batch 4
run 300 bench_b2 python bench.py --steps 2 --warmup 3 --batch 2
run 300 retrieve python bench.py --workload retrieve --steps 2 --warmup 2
exit 0
