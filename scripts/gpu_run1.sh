#!/bin/bash
# first GPU trip: parity tests, smoke, scan bench, launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r1_gpuinfo.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r1_smoke.log
timeout 600 python bench.py --workload scan --steps 20 --warmup 3 > gpurun_out/r1_bench_scan.log 2>&1; echo "bench rc=$?" >> gpurun_out/r1_bench_scan.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r1_launches.csv python bench.py --workload scan --steps 3 --warmup 3 > gpurun_out/r1_ncu.log 2>&1
tail -5 gpurun_out/r1_pytest.log gpurun_out/r1_smoke.log gpurun_out/r1_bench_scan.log
