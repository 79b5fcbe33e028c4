#!/bin/bash
# 12-warp GEMM (setmaxnreg 72/216) + software-pipelined RoPE epilogue; attention poly mask back to 2/8
mkdir -p gpurun_out
run() { local t=$1 name=$2; shift 2; local t0=$SECONDS; timeout $t "$@" > gpurun_out/r20_$name.log 2>&1; local rc=$?; echo "rc=$rc secs=$((SECONDS-t0))" >> gpurun_out/r20_$name.log; tail -n 9 gpurun_out/r20_$name.log | cut -c1-2500; return $rc; }
run 400 tests python -m pytest tests/test_gemm_gpu.py tests/test_flux_gpu.py tests/test_vit_gpu.py tests/test_vae_gpu.py tests/test_siglip_gpu.py tests/test_pipelines_gpu.py -m gpu -x -q || exit 0
run 120 qkv python scripts/bench_qkv.py
run 120 attn_bench python scripts/bench_attn.py
run 600 bench_default python bench.py --steps 3
exit 0
