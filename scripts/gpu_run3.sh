#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x > gpurun_out/r3_gemm.log 2>&1; echo "gemm rc=$?" >> gpurun_out/r3_gemm.log
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gemm_gpu.py > gpurun_out/r3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3_pytest.log
timeout 600 python bench.py --workload scan --steps 20 --warmup 3 > gpurun_out/r3_bench_scan.log 2>&1; echo "bench rc=$?" >> gpurun_out/r3_bench_scan.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r3_launches.csv python bench.py --workload scan --steps 3 --warmup 3 > gpurun_out/r3_ncu.log 2>&1
tail -n 12 gpurun_out/r3_gemm.log; tail -n 6 gpurun_out/r3_pytest.log; tail -n 2 gpurun_out/r3_bench_scan.log
exit 0
