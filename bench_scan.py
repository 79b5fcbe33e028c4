"""Scan workloads of bench.py: the corpus cosine-similarity x top-k of the first retrieval stage
(retrieval/clip100_resnet_style_all_shots.py:425-434: faiss.IndexFlatIP add / search) - BASELINE config C5.

  --workload scan           one point: 1 M x 512 fp32 rows per GPU, nq = 1, top-100 (the metric's "achieved HBM GB/s vs peak")
  --workload scan --sweep   N in {1e4, 3e4, 1e5, 3e5, 1e6} x D in {512, 768} x nq in {1, 64}, one JSON line
                            whose `sweep` list carries every point

Multi-GPU (weak scaling): every rank owns n rows of an (n * world)-row corpus; a search = local scan x top-k -> ONE
all-gather of the per-shard (id, score) pairs (NCCL) -> merge kernel. At world > 1 every timed search result is compared,
bit for bit, with a replicated single-index search over the all-gathered corpus on planted queries (`verified_*` keys): the
number is only reported if sharded == single.
"""
from __future__ import annotations

import json
import os
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent

SCAN_N, SCAN_D, SCAN_K, SCAN_NQ = 1_000_000, 512, 100, 1
SWEEP_N = (10_000, 30_000, 100_000, 300_000, 1_000_000)
SWEEP_D = (512, 768)
SWEEP_NQ = (1, 64)
L2_BYTES = 126 * 1024 * 1024
METRIC = "corpus cosine-top-k scan throughput (algorithmic bytes / device time)"


def scan_algorithmic_bytes(n, d, nq, k):
    return n * d * 4 + nq * d * 4 + nq * k * 12


def scan_traffic_from_profile():
    """dram read+write bytes per launch of the scan kernel from the committed ncu --set full capture."""
    for name in ("r02_scan_traffic.json", "r01_scan_traffic.json"):
        try:
            return json.loads((REPO / "profiles" / name).read_text())["traffic_bytes_per_launch"]
        except Exception:
            continue
    return None


def make_corpus_device(n, d, seed, device):
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    x = torch.randn(n, d, generator=g, device=device)
    x /= x.norm(dim=1, keepdim=True)
    return x


def planted_queries(x_all, nq, seed, device):
    """nq unit queries, each a corpus row (evenly spaced global ids, so every shard owns some) plus 5 % noise: the planted
    row must come back as hit 0."""
    import torch
    n = x_all.shape[0]
    ids = torch.linspace(0, n - 1, nq + 2, device=device)[1:-1].round().long() if nq > 1 else \
        torch.tensor([int(n * 0.7)], device=device)
    g = torch.Generator(device=device).manual_seed(seed)
    q = x_all[ids] + 0.05 * torch.randn(nq, x_all.shape[1], generator=g, device=device) / x_all.shape[1] ** 0.5
    return (q / q.norm(dim=1, keepdim=True)).contiguous(), ids


def measure(rank, world, local, n, d, nq, k, steps, warmup, flush_l2=False, e2e=True):
    """One scan point on an initialised process group. Returns a dict on every rank (rank 0 uses it)."""
    import torch
    import torch.distributed as dist

    from domain_rag_b200 import _lib
    from domain_rag_b200.benchutil import barrier, max_over_ranks
    from domain_rag_b200.index import IndexFlatIP, ShardedIndexFlatIP

    dev = torch.device("cuda", local)
    x = make_corpus_device(n, d, 4006 + rank, dev)
    six = ShardedIndexFlatIP(d, rank, world, device=local)
    six.add_local(x, lo=rank * n, ntotal_global=n * world)
    ix = six._index
    ix.set_timing(True)

    single = None
    if world > 1:           # replicated single index for the on-hardware sharded == single check
        x_all = torch.empty((world * n, d), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(x_all, x)
        single = IndexFlatIP(d, local)
        single.add_device(x_all, base_id=0)
        q_dev, planted = planted_queries(x_all, nq, 4999, dev)
        D1, I1 = single.search_device(q_dev, k)
        torch.cuda.synchronize()
    else:
        q_dev, planted = planted_queries(x, nq, 4999, dev)
    q_host = q_dev.cpu().pin_memory()
    flush = torch.empty(2 * L2_BYTES // 4, dtype=torch.float32, device=dev) if flush_l2 else None

    # exchange step at world > 1: the packed NCCL all-gather + merge kernel is timed first (nccl_ms_per_search), then the
    # fused NVLink peer-memory kernel, which the line reports when symmetric memory is available on this box
    nccl_ms = None
    results = []
    if world > 1:
        for _ in range(max(warmup, 3)):
            six.search(q_dev, k)
        barrier(world)
        n0, n1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0.record()
        for _ in range(steps):
            results.append(six.search(q_dev, k))
        n1.record()
        barrier(world)
        nccl_ms = max_over_ranks(n0.elapsed_time(n1), world) / steps
        six.enable_p2p()
    # Warm-up keeps as many result tensors alive as the timed loop will, then frees them: the timed searches then take their
    # output blocks from torch's caching allocator instead of a fresh cudaMalloc (seen as a 3 ms hiccup in a 20-search bracket
    # right after torch.cuda.empty_cache()).
    keep = [six.search(q_dev, k) for _ in range(max(warmup, 3, steps))]
    torch.cuda.synchronize(dev)
    del keep
    barrier(world)
    _lib.launch_count(reset=True)
    if flush is None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier(world)
        e0.record()
        for _ in range(steps):
            results.append(six.search(q_dev, k))
        e1.record()
        n_launches = _lib.launch_count()
        barrier(world)
        total_ms = max_over_ranks(e0.elapsed_time(e1), world)
    else:                   # corpus smaller than ~2x L2: evict it between timed searches, time each search on its own
        evs = []
        barrier(world)
        for _ in range(steps):
            flush.fill_(1.0)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            results.append(six.search(q_dev, k))
            b.record()
            evs.append((a, b))
        n_launches = _lib.launch_count()
        barrier(world)
        total_ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in evs), world)
    kern_ms = []
    for _ in range(steps):
        if flush is not None:
            flush.fill_(1.0)
        six.search(q_dev, k)
        kern_ms.append(ix.last_scan_ms())
    barrier(world)
    kern_avg = sum(kern_ms) / len(kern_ms)

    verified = None
    if single is not None:
        ok = all(bool(torch.equal(I, I1)) and bool(torch.equal(D, D1)) for D, I in results)
        hit0 = bool(torch.equal(results[-1][1][:, 0], planted))
        flag = torch.tensor([int(ok and hit0)], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        verified = bool(flag.item())
        if not verified:
            raise AssertionError(f"rank {rank}: NCCL-sharded search != replicated single-index search "
                                 f"(n={n}, d={d}, nq={nq}, world={world}; equal={ok}, planted hit0={hit0})")
        del single, x_all
    else:
        assert bool(torch.equal(results[-1][1][:, 0], planted + rank * n)), "planted rows did not come back as hit 0"

    e2e_s = None
    if e2e:                 # public host-buffer API: H2D of the queries, D2H of (D, I) every step
        def step_e2e():
            D, I = six.search(q_host.to(dev, non_blocking=True), k)
            return D.cpu(), I.cpu()
        for _ in range(3):
            step_e2e()
        barrier(world)
        t0 = time.perf_counter()
        for _ in range(steps):
            step_e2e()
        barrier(world)
        e2e_s = max_over_ranks(time.perf_counter() - t0, world) / steps

    alg = scan_algorithmic_bytes(n, d, nq, k)
    ms = total_ms / steps
    geo = ix.last_launch()
    out = {"n_per_gpu": n, "d": d, "nq": nq, "k": k, "ms_per_search": round(ms, 4), "kernel_ms": round(kern_avg, 4),
           "gbs": round(alg * world / (ms * 1e-3) / 1e9, 1), "kernel_gbs": round(alg / (kern_avg * 1e-3) / 1e9, 1),
           "e2e_gbs": None if e2e_s is None else round(alg * world / e2e_s / 1e9, 1),
           "l2": "flushed between searches" if flush is not None else "corpus larger than L2",
           "exchange": six.exchange if world > 1 else None,
           "nccl_ms_per_search": None if nccl_ms is None else round(nccl_ms, 4),
           "launches": int(n_launches), "grid": geo["grid"], "ring_stages": geo["stages"],
           "rows_per_stage": geo["rows_per_stage"], "nq_batch": geo["nq_batch"],
           "verified_sharded_equals_single": verified}
    del six, ix, x
    torch.cuda.empty_cache()
    return out


def _line(pt, world, steps, warmup, clocks, peaks, extra_cfg=None):
    n, d, nq, k = pt["n_per_gpu"], pt["d"], pt["nq"], pt["k"]
    alg = scan_algorithmic_bytes(n, d, nq, k)
    cfg = {"workload": f"C5 scan: {n} x {d} fp32 embeddings per GPU, nq={nq}, top-{k}; index row-sharded, per-shard top-k "
                       f"exchanged and merged on every rank", "l2_policy": pt["l2"] + f" ({alg / 1e6:.0f} MB per GPU vs 126 MB)",
           "grid": pt["grid"], "ring_stages": pt["ring_stages"], "rows_per_stage": pt["rows_per_stage"],
           "exchange": pt["exchange"], "nccl_ms_per_search": pt["nccl_ms_per_search"],
           "verified_sharded_equals_single": pt["verified_sharded_equals_single"]}
    cfg.update(extra_cfg or {})
    return {"metric": METRIC, "value": pt["gbs"], "unit": "GB/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": pt["ms_per_search"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": cfg,
            "e2e": {"value": pt["e2e_gbs"], "unit": "GB/s", "h2d_bytes_per_step": nq * d * 4, "d2h_bytes_per_step": nq * k * 12},
            "gpu_launches": pt["launches"], "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": pt["kernel_gbs"], "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": round(pt["kernel_gbs"] / peaks["hbm_gbs"], 4), "traffic": scan_traffic_from_profile(),
                         "kernel": "ip_scan_topk_kernel", "kernel_ms": pt["kernel_ms"], "peak_source": peaks["source"]}}


def run(args):
    import torch  # noqa: F401

    from domain_rag_b200.benchutil import ClockSampler, dist_setup, measured_peaks
    rank, world, local = dist_setup(args.gpus)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if getattr(args, "sweep", False):
        pts = []
        for d in SWEEP_D:
            for nq in SWEEP_NQ:
                for n in SWEEP_N:
                    pts.append(measure(rank, world, local, n, d, nq, SCAN_K, args.steps, args.warmup,
                                       flush_l2=n * d * 4 < 2 * L2_BYTES, e2e=False))
        head = next(p for p in pts if (p["n_per_gpu"], p["d"], p["nq"]) == (SCAN_N, SCAN_D, SCAN_NQ))
    else:
        pts = None
        head = measure(rank, world, local, SCAN_N, SCAN_D, SCAN_NQ, SCAN_K, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else {}
    if rank != 0:
        return None
    out = _line(head, world, args.steps, args.warmup, clocks, measured_peaks())
    if pts is not None:
        out["e2e"]["value"] = head["gbs"] if out["e2e"]["value"] is None else out["e2e"]["value"]
        out["e2e"]["note"] = "sweep mode times the device-resident search only (e2e leg skipped per point)"
        out["sweep"] = pts
    if world == 1:
        out["cpu_baseline"] = cpu_baseline(SCAN_D, SCAN_NQ, SCAN_K)
    return out


def cpu_baseline(d, nq, k, n_sample=200_000, min_seconds=2.0):
    """The oracle (numpy fp64-accumulate scan + top-k) on a bounded sample of the same workload."""
    import numpy as np

    from oracle import ip_topk as O
    g = np.random.default_rng(1)
    x = g.standard_normal((n_sample, d), dtype=np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    q = g.standard_normal((nq, d), dtype=np.float32)
    O.ip_topk(x[:1000], q, k)
    t0 = time.perf_counter()
    reps = 0
    while True:
        O.ip_topk(x, q, k)
        reps += 1
        if time.perf_counter() - t0 > min_seconds:
            break
    dt = (time.perf_counter() - t0) / reps
    return {"value": round(scan_algorithmic_bytes(n_sample, d, nq, k) / dt / 1e9, 3), "unit": "GB/s",
            "cores": os.cpu_count(), "kind": "port",
            "sample": f"oracle.ip_topk over {n_sample} x {d} rows ({reps} reps), BLAS threads = all cores"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    n, d, k, nq = SCAN_N, SCAN_D, SCAN_K, SCAN_NQ
    import numpy as np

    from oracle import ip_topk as O
    n_sample = 200_000
    g = np.random.default_rng(1)
    x = g.standard_normal((n_sample, d), dtype=np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    q = g.standard_normal((nq, d), dtype=np.float32)
    for _ in range(args.warmup):
        O.ip_topk(x, q, k)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.ip_topk(x, q, k)
    dt = (time.perf_counter() - t0) / args.steps
    val = round(scan_algorithmic_bytes(n_sample, d, nq, k) / dt / 1e9, 3)
    sample = f"each step = oracle.ip_topk over a {n_sample}-row sample of the {n}-row corpus"
    return {"impl": "reference", "metric": METRIC,
            "value": val, "unit": "GB/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"C5 scan: {n} x {d} fp32 embeddings per GPU, nq={nq}, top-{k}; index row-sharded, "
                                   f"per-shard top-k exchanged and merged on every rank", "sample": sample},
            "cpu_baseline": {"value": val, "unit": "GB/s", "cores": os.cpu_count(), "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
