#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_flux_gpu.py tests/test_vit_gpu.py tests/test_retrieval_cli_gpu.py -m gpu -x -q > gpurun_out/r12_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r12_pytest.log
tail -n 15 gpurun_out/r12_pytest.log
timeout 300 python scripts/bench_attn.py > gpurun_out/r12_attn.log 2>&1; tail -n 7 gpurun_out/r12_attn.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r12_bench.log 2>&1; echo "rc=$?" >> gpurun_out/r12_bench.log
tail -n 2 gpurun_out/r12_bench.log | cut -c1-1800
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16|attention_tcgen05" -s 4 -c 4 \
  -o gpurun_out/r12_kernels python scripts/prof_kernels.py > gpurun_out/r12_ncu.log 2>&1
tail -n 3 gpurun_out/r12_ncu.log
exit 0
