#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/bench_gemm.py > gpurun_out/r4_gemm_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ip_scan_topk -s 4 -c 1 -o gpurun_out/r4_scan_prof python bench.py --workload scan --steps 3 --warmup 3 > gpurun_out/r4_ncu_scan.log 2>&1
tail -n 12 gpurun_out/r4_gemm_bench.log; tail -n 3 gpurun_out/r4_ncu_scan.log
exit 0
