#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_flux_gpu.py -m gpu -q -x > gpurun_out/r8_flux.log 2>&1; echo "rc=$?" >> gpurun_out/r8_flux.log
timeout 600 python scripts/bench_attn.py > gpurun_out/r8_attn_bench.log 2>&1
tail -n 25 gpurun_out/r8_flux.log; cat gpurun_out/r8_attn_bench.log
exit 0
