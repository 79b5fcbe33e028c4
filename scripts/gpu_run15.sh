#!/bin/bash
# re-entry verification: full GPU suite, fresh bench lines, launch list + ncu --set full of the top kernels
mkdir -p gpurun_out
run() { local t=$1 name=$2; shift 2; local t0=$SECONDS; timeout $t "$@" > gpurun_out/r15_$name.log 2>&1; local rc=$?; echo "rc=$rc secs=$((SECONDS-t0))" >> gpurun_out/r15_$name.log; tail -n 8 gpurun_out/r15_$name.log | cut -c1-1800; return $rc; }
run 900 tests python -m pytest tests -m gpu -x -q --durations=15
run 120 attn_bench python scripts/bench_attn.py
run 500 bench_full python bench.py --steps 3 --warmup 3
run 200 bench_scan python bench.py --workload scan
run 300 bench_ref python bench.py --impl reference --steps 1 --warmup 0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 40000 --csv \
  --log-file gpurun_out/r15_launches.csv python bench.py --steps 1 --warmup 3 > gpurun_out/r15_launches_bench.log 2>&1
wc -l gpurun_out/r15_launches.csv
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16|attention_tcgen05" -s 4 -c 4 \
  -o gpurun_out/r15_kernels python scripts/prof_kernels.py > gpurun_out/r15_ncu.log 2>&1
tail -n 3 gpurun_out/r15_ncu.log
exit 0
