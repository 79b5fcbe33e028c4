#!/bin/bash
# attention tail rows on CUDA cores (S = k*128 + 1|2), launch counter, batched generation test
mkdir -p gpurun_out
run() { local t=$1 name=$2; shift 2; local t0=$SECONDS; timeout $t "$@" > gpurun_out/r24_$name.log 2>&1; local rc=$?; echo "rc=$rc secs=$((SECONDS-t0))" >> gpurun_out/r24_$name.log; tail -n 8 gpurun_out/r24_$name.log | cut -c1-2500; return $rc; }
run 400 tests python -m pytest tests/test_flux_gpu.py tests/test_vit_gpu.py tests/test_pipelines_gpu.py tests/test_siglip_gpu.py -m gpu -x -q -k "not full_size" || exit 0
run 120 attn_bench python scripts/bench_attn.py
run 300 retrieve python bench.py --workload retrieve
exit 0
