#!/bin/bash
# attention v3b (3/8 polynomial exp2), default bench line at batch 4, launch list + ncu --set full at the batch-4 shapes
mkdir -p gpurun_out
run() { local t=$1 name=$2; shift 2; local t0=$SECONDS; timeout $t "$@" > gpurun_out/r19_$name.log 2>&1; local rc=$?; echo "rc=$rc secs=$((SECONDS-t0))" >> gpurun_out/r19_$name.log; tail -n 9 gpurun_out/r19_$name.log | cut -c1-2500; return $rc; }
run 300 tests python -m pytest tests/test_flux_gpu.py tests/test_pipelines_gpu.py -m gpu -x -q || exit 0
run 120 attn_bench python scripts/bench_attn.py
run 600 bench_default python bench.py
run 200 retrieve python bench.py --workload retrieve
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16|attention_tcgen05" -s 4 -c 4 \
  -o gpurun_out/r19_kernels_b4 python scripts/prof_kernels.py 4 > gpurun_out/r19_ncu.log 2>&1
tail -n 2 gpurun_out/r19_ncu.log
DRAG_BENCH_LAUNCH_LIST_ONLY=1 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 1600 --csv \
  --log-file gpurun_out/r19_launches.csv python bench.py --steps 1 --warmup 3 > gpurun_out/r19_launches_bench.log 2>&1
wc -l gpurun_out/r19_launches.csv; tail -n 2 gpurun_out/r19_launches_bench.log
exit 0
