#!/bin/bash
# final verification of the round: whole GPU suite, smoke, every bench line (default compose, loop scope, scan, retrieve, reference arm)
mkdir -p gpurun_out
run() { local t=$1 name=$2; shift 2; local t0=$SECONDS; timeout $t "$@" > gpurun_out/r25_$name.log 2>&1; local rc=$?; echo "rc=$rc secs=$((SECONDS-t0))" >> gpurun_out/r25_$name.log; tail -n 5 gpurun_out/r25_$name.log | cut -c1-2600; return $rc; }
run 900 tests python -m pytest tests -m gpu -x -q --durations=8
run 300 smoke python -c "import __graft_entry__ as g; g.smoke()"
run 600 bench_default python bench.py --steps 3
run 300 bench_loop python bench.py --scope loop --steps 1
run 200 bench_scan python bench.py --workload scan
run 200 bench_retrieve python bench.py --workload retrieve
run 200 bench_ref python bench.py --impl reference --steps 1 --warmup 0
exit 0
