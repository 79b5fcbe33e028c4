#!/bin/bash
# packed single all-gather in the sharded search: N=2 scan bench (was 9876 GB/s, 0.4148 ms/step with two all-gathers)
mkdir -p gpurun_out
run() { local t=$1 name=$2; shift 2; local t0=$SECONDS; timeout $t "$@" > gpurun_out/r27_$name.log 2>&1; local rc=$?; echo "rc=$rc secs=$((SECONDS-t0))" >> gpurun_out/r27_$name.log; tail -n 3 gpurun_out/r27_$name.log | cut -c1-1500; return $rc; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544"
run 120 n2_scan $TR bench.py --gpus 2 --workload scan
exit 0
