#!/bin/bash
# hd-64 single-tile attention variant (parity + microbench), raster heuristic v2, then the N=2 launch contract (torchrun)
mkdir -p gpurun_out
run() { local t=$1 name=$2; shift 2; local t0=$SECONDS; timeout $t "$@" > gpurun_out/r18_$name.log 2>&1; local rc=$?; echo "rc=$rc secs=$((SECONDS-t0))" >> gpurun_out/r18_$name.log; tail -n 9 gpurun_out/r18_$name.log | cut -c1-2500; return $rc; }
run 300 tests python -m pytest tests/test_flux_gpu.py tests/test_vit_gpu.py tests/test_gemm_gpu.py -m gpu -x -q
run 120 attn_bench python scripts/bench_attn.py
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
run 200 n2_scan $TR bench.py --gpus 2 --workload scan
run 300 n2_retrieve $TR bench.py --gpus 2 --workload retrieve --steps 2 --warmup 2
run 200 n2_ref $TR bench.py --gpus 2 --impl reference --steps 1 --warmup 0
run 500 n2_compose $TR bench.py --gpus 2 --steps 1 --warmup 3
exit 0
