#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/attn_variants.py > gpurun_out/r6_attn_variants.log 2>&1
timeout 900 python -m pytest tests/test_flux_gpu.py -m gpu -q > gpurun_out/r6_flux.log 2>&1; echo "rc=$?" >> gpurun_out/r6_flux.log
cat gpurun_out/r6_attn_variants.log; tail -n 40 gpurun_out/r6_flux.log
exit 0
