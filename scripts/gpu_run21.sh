#!/bin/bash
# column-group raster + L2 hints microbench; bench with the same-box cuBLAS yardstick
mkdir -p gpurun_out
run() { local t=$1 name=$2; shift 2; local t0=$SECONDS; timeout $t "$@" > gpurun_out/r21_$name.log 2>&1; local rc=$?; echo "rc=$rc secs=$((SECONDS-t0))" >> gpurun_out/r21_$name.log; tail -n 12 gpurun_out/r21_$name.log | cut -c1-2500; return $rc; }
run 200 tests python -m pytest tests/test_gemm_gpu.py -m gpu -x -q || exit 0
run 300 raster2 python scripts/bench_raster2.py
exit 0
