#!/bin/bash
# fail-fast, tight timeouts: attention correctness first (a hang costs 3 minutes, not 40)
mkdir -p gpurun_out
run() { local t=$1 name=$2; shift 2; timeout $t "$@" > gpurun_out/r13_$name.log 2>&1; local rc=$?; echo "rc=$rc" >> gpurun_out/r13_$name.log; tail -n 12 gpurun_out/r13_$name.log | cut -c1-1600; return $rc; }
run 180 attn_test python -m pytest tests/test_flux_gpu.py -m gpu -x -q -k "attention" || exit 0
run 120 attn_bench python scripts/bench_attn.py
run 400 newtests python -m pytest tests/test_vae_gpu.py tests/test_siglip_gpu.py tests/test_pipelines_gpu.py tests/test_retrieval_cli_gpu.py -m gpu -q
run 400 alltests python -m pytest tests -m gpu -x -q --deselect tests/test_vae_gpu.py --deselect tests/test_siglip_gpu.py --deselect tests/test_pipelines_gpu.py --deselect tests/test_retrieval_cli_gpu.py
run 400 bench python bench.py --steps 3 --warmup 3
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16|attention_tcgen05" -s 4 -c 4 \
  -o gpurun_out/r13_kernels python scripts/prof_kernels.py > gpurun_out/r13_ncu.log 2>&1
tail -n 3 gpurun_out/r13_ncu.log
exit 0
