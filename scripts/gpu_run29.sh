#!/bin/bash
# C2 with the reference script's default tower (ViT-B/32) next to the BASELINE tower (ViT-L/14)
mkdir -p gpurun_out
timeout 150 python bench.py --workload retrieve --clip-model ViT-B/32 --steps 3 > gpurun_out/r29_retrieve_b32.log 2>&1; echo "rc=$?" >> gpurun_out/r29_retrieve_b32.log
tail -n 2 gpurun_out/r29_retrieve_b32.log | cut -c1-2500
exit 0
