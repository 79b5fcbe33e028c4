#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vit_gpu.py -m gpu -q > gpurun_out/r9_vit.log 2>&1; echo "rc=$?" >> gpurun_out/r9_vit.log
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/r9_bench_compose.log 2>&1; echo "rc=$?" >> gpurun_out/r9_bench_compose.log
tail -n 30 gpurun_out/r9_vit.log; tail -n 3 gpurun_out/r9_bench_compose.log
exit 0
