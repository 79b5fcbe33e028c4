#!/bin/bash
# round-end style verification: whole GPU suite, smoke, every bench line, ncu full at the batch-4 shapes with the final raster
mkdir -p gpurun_out
run() { local t=$1 name=$2; shift 2; local t0=$SECONDS; timeout $t "$@" > gpurun_out/r22_$name.log 2>&1; local rc=$?; echo "rc=$rc secs=$((SECONDS-t0))" >> gpurun_out/r22_$name.log; tail -n 6 gpurun_out/r22_$name.log | cut -c1-2600; return $rc; }
run 900 tests python -m pytest tests -m gpu -x -q
run 300 smoke python -c "import __graft_entry__ as g; g.smoke()"
run 200 raster2 python scripts/bench_raster2.py
run 600 bench_default python bench.py
run 200 bench_scan python bench.py --workload scan
run 200 bench_retrieve python bench.py --workload retrieve
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16|attention_tcgen05" -s 4 -c 4 \
  -o gpurun_out/r22_kernels_b4 python scripts/prof_kernels.py 4 > gpurun_out/r22_ncu.log 2>&1
tail -n 2 gpurun_out/r22_ncu.log
exit 0
