#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_topk_gpu.py -m gpu -q > gpurun_out/r5_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r5_pytest.log
timeout 600 python bench.py --workload scan --steps 20 --warmup 3 > gpurun_out/r5_bench_scan.log 2>&1; echo "bench rc=$?" >> gpurun_out/r5_bench_scan.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ip_scan_topk -s 4 -c 1 -o gpurun_out/r5_scan_prof python bench.py --workload scan --steps 3 --warmup 3 > gpurun_out/r5_ncu_scan.log 2>&1
tail -n 6 gpurun_out/r5_pytest.log; tail -n 2 gpurun_out/r5_bench_scan.log
exit 0
