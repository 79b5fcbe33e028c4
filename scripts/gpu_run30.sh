#!/bin/bash
# C2 embed batch size: 500 vs the default 250 (same box, back to back)
mkdir -p gpurun_out
for eb in 500 250; do
  timeout 100 python bench.py --workload retrieve --embed-batch $eb --steps 2 --warmup 2 > gpurun_out/r30_retrieve_eb$eb.log 2>&1
  grep -o '"value": [0-9.]*, "unit": "images/s", "n_gpus"' gpurun_out/r30_retrieve_eb$eb.log | head -1
done
exit 0
