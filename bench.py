#!/usr/bin/env python
"""bench.py - the driver's measurement contract for the Domain-RAG retrieve-then-compose hot path.

    python bench.py --gpus N --steps K --warmup W [--workload compose|c3|scan|retrieve] [--impl reference]

Prints ONE JSON line on rank 0. Workloads (each module's docstring has the details):
  compose  (default) BASELINE metric: composed 1024^2 images/sec, 50 Flux-Redux steps per image (C4 per-GPU slice, batch 4).
           The line also carries `gpu_baseline` (the oracle modules in bf16 torch-eager on the same B200 = the reference's
           torch path), `cpu_baseline` (N = 1) and `secondary` = the second half of the metric and the other configs measured
           in the same run: C5 scan GB/s vs HBM peak, C2 retrieve images/s, C3 seconds per batch.  [bench_compose.py]
  c3       BASELINE config C3: Flux-Redux outpainting 512^2, 20 steps, batch 8.                      [bench_compose.py]
  scan     C5: corpus cosine-top-k, 1M x 512 fp32 rows per GPU, top-100, achieved HBM GB/s; `--sweep` = the whole
           N x D x nq grid. At N > 1 every search is verified against a replicated single index.     [bench_scan.py]
  retrieve C2: 10k-image corpus -> CLIP ViT-L/14 embed -> resident index -> top-100 -> style re-rank. [bench_retrieve.py]
`--impl reference` times the CPU oracle (the reference's algorithm; its own third-party packages are not installable
offline) on the host cores for the same metric/config.
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

DEFAULT_STEPS = {"scan": 20, "retrieve": 3, "c3": 3, "compose": 4}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="default: 4 (compose: 4 x --batch images), 3 (c3, retrieve), 20 (scan)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="compose", choices=["compose", "c3", "scan", "retrieve"])
    ap.add_argument("--batch", type=int, default=None,
                    help="compose / c3: compositions per GPU and step, run as one batch (default 4 = C4: 32 compositions / "
                         "8 GPUs; c3: 8)")
    ap.add_argument("--sweep", action="store_true", help="scan: the whole C5 grid N x D x nq in one line (`sweep` list)")
    ap.add_argument("--no-secondary", dest="no_secondary", action="store_true",
                    help="compose: skip the secondary block (scan / C2 / C3 records)")
    ap.add_argument("--no-gpu-baseline", dest="no_gpu_baseline", action="store_true",
                    help="compose / c3: skip the torch-eager GPU baseline leg")
    ap.add_argument("--clip-model", dest="clip_model", default=None, choices=["ViT-L/14", "ViT-B/32", "ViT-B/16"],
                    help="retrieve: image tower (default ViT-L/14 = BASELINE config C2; ViT-B/32 = the reference script's default)")
    ap.add_argument("--embed-batch", dest="embed_batch", type=int, default=None, help="retrieve: images per encode_image call (default 500)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.steps is None:
        args.steps = DEFAULT_STEPS[args.workload]
    if args.workload == "retrieve":
        import bench_retrieve as W
        out = W.run_reference(args) if args.impl == "reference" else W.run(args)
    elif args.workload == "scan":
        import bench_scan as W
        out = W.run_reference(args) if args.impl == "reference" else W.run(args)
    else:
        import bench_compose as W
        out = W.run_reference(args, args.workload) if args.impl == "reference" else W.run(args, args.workload)
    if out is not None:
        print(json.dumps(out), flush=True)
    try:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


if __name__ == "__main__":
    main()
