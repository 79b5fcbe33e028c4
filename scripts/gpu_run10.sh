#!/bin/bash
# full GPU test-suite, default bench, ncu launch list of the bench's timed region, ncu --set full of GEMM + attention
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r10_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r10_pytest.log
tail -n 5 gpurun_out/r10_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r10_bench.log 2>&1; echo "rc=$?" >> gpurun_out/r10_bench.log
tail -n 2 gpurun_out/r10_bench.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 1000 --csv \
  --log-file gpurun_out/r10_launches.csv python bench.py --steps 1 --warmup 3 > gpurun_out/r10_ncu_launch.log 2>&1
tail -n 2 gpurun_out/r10_ncu_launch.log
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:gemm_bf16_tcgen05 -s 40 -c 4 -o gpurun_out/r10_gemm python bench.py --steps 1 --warmup 3 > gpurun_out/r10_ncu_gemm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:attention_tcgen05 -s 20 -c 2 -o gpurun_out/r10_attn python bench.py --steps 1 --warmup 3 > gpurun_out/r10_ncu_attn.log 2>&1
ls -la gpurun_out | tail -n 8
exit 0
