#!/bin/bash
# QKV epilogue with bias / norm weights staged in shared memory: parity, then the QKV microbench (run 20: 1179 / 1216 TFLOP/s)
mkdir -p gpurun_out
run() { local t=$1 name=$2; shift 2; local t0=$SECONDS; timeout $t "$@" > gpurun_out/r28_$name.log 2>&1; local rc=$?; echo "rc=$rc secs=$((SECONDS-t0))" >> gpurun_out/r28_$name.log; tail -n 5 gpurun_out/r28_$name.log | cut -c1-1500; return $rc; }
run 200 tests python -m pytest tests/test_gemm_gpu.py tests/test_flux_gpu.py tests/test_vit_gpu.py -m gpu -x -q || exit 0
run 60 qkv python scripts/bench_qkv.py
exit 0
