#!/bin/bash
mkdir -p gpurun_out
run() { local t=$1 name=$2; shift 2; timeout $t "$@" > gpurun_out/r14_$name.log 2>&1; local rc=$?; echo "rc=$rc" >> gpurun_out/r14_$name.log; tail -n 14 gpurun_out/r14_$name.log | cut -c1-1700; return $rc; }
run 240 kern_test python -m pytest tests/test_flux_gpu.py tests/test_gemm_gpu.py -m gpu -x -q || exit 0
run 120 attn_bench python scripts/bench_attn.py
run 240 gemm_bench python scripts/bench_gemm.py --sustain
run 400 newtests python -m pytest tests/test_vae_gpu.py tests/test_siglip_gpu.py tests/test_pipelines_gpu.py tests/test_retrieval_cli_gpu.py -m gpu -q
run 300 bench_loop python bench.py --steps 3 --warmup 3 --scope loop
run 500 bench_full python bench.py --steps 3 --warmup 3 --scope full
exit 0
