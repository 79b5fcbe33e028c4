#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -x -q > gpurun_out/r11_pytest_gemm.log 2>&1; echo "rc=$?" >> gpurun_out/r11_pytest_gemm.log
tail -n 15 gpurun_out/r11_pytest_gemm.log
if grep -q "rc=0" gpurun_out/r11_pytest_gemm.log; then
  timeout 600 python scripts/bench_gemm.py > gpurun_out/r11_gemm.log 2>&1; tail -n 16 gpurun_out/r11_gemm.log
  timeout 600 python scripts/bench_gemm.py --sustain > gpurun_out/r11_gemm_sustain.log 2>&1; tail -n 16 gpurun_out/r11_gemm_sustain.log
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r11_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r11_pytest.log
  tail -n 8 gpurun_out/r11_pytest.log
  timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r11_bench.log 2>&1; echo "rc=$?" >> gpurun_out/r11_bench.log
  tail -n 2 gpurun_out/r11_bench.log
fi
exit 0
