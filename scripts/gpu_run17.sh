#!/bin/bash
# attention v3 (packed fp32x2 softmax, split P arrival) + grouped GEMM raster: parity first, then microbenches, then bench at B=1/2/4
mkdir -p gpurun_out
run() { local t=$1 name=$2; shift 2; local t0=$SECONDS; timeout $t "$@" > gpurun_out/r17_$name.log 2>&1; local rc=$?; echo "rc=$rc secs=$((SECONDS-t0))" >> gpurun_out/r17_$name.log; tail -n 9 gpurun_out/r17_$name.log | cut -c1-2500; return $rc; }
run 300 tests python -m pytest tests/test_flux_gpu.py tests/test_gemm_gpu.py tests/test_vit_gpu.py tests/test_siglip_gpu.py tests/test_vae_gpu.py -m gpu -x -q || exit 0
run 120 attn_bench python scripts/bench_attn.py
run 240 raster python scripts/bench_raster.py
run 300 bench_b2 python bench.py --steps 2 --warmup 3 --batch 2
run 300 bench_b1 python bench.py --steps 3 --warmup 3 --batch 1
run 400 bench_b4 python bench.py --steps 2 --warmup 3 --batch 4
exit 0
