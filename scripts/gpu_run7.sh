#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/r7_bench_compose.log 2>&1; echo "rc=$?" >> gpurun_out/r7_bench_compose.log
tail -n 20 gpurun_out/r7_bench_compose.log
exit 0
