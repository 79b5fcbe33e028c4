#!/bin/bash
# N=8 launch contract: NCCL init, sharded scan (all-gather of per-shard top-k across 8 ranks), reference arm under torchrun
mkdir -p gpurun_out
run() { local t=$1 name=$2; shift 2; local t0=$SECONDS; timeout $t "$@" > gpurun_out/r26_$name.log 2>&1; local rc=$?; echo "rc=$rc secs=$((SECONDS-t0))" >> gpurun_out/r26_$name.log; tail -n 4 gpurun_out/r26_$name.log | cut -c1-1800; return $rc; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
run 150 n8_scan $TR bench.py --gpus 8 --workload scan
run 60 n8_ref $TR bench.py --gpus 8 --impl reference --workload scan --steps 1 --warmup 0
exit 0
