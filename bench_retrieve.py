"""Retrieve workload of bench.py (BASELINE config C2): embed a 10 000-image synthetic corpus with the CLIP
ViT-L/14 image tower (+ L2 normalise), keep the embeddings resident as the inner-product index of the GPU that made them,
embed the 7 one-shot ArTaxOr queries, exact top-100, then the ResNet-50-stem style re-rank of the 7 x (1 + 100) images.

Reference path this replaces (retrieval/clip100_resnet_style_all_shots.py): compute_coco_clip_features :270-296
(batch-1 encode_image + normalise per image), clip_first_stage_retrieval :396-451 (faiss IndexFlatIP rebuilt per query),
resnet_second_stage_rerank :454-497 (101 batch-1 stem passes per query).

One bench "step" = the whole C2 job on every GPU (weak scaling: every rank embeds its own 10 000-image shard; the index
is row-sharded, a search = local scan x top-k -> all-gather of the per-shard top-k -> merge). `value`: preprocessed
images resident in HBM; `e2e`: pinned host tensors, H2D per batch, D2H of the ranked results. Roofline: the encoder is
tensor-bound, 162.0 GFLOP per ViT-L/14 image.
"""
from __future__ import annotations

import os
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent

N_CORPUS, N_QUERY, TOP_K, EMBED_BATCH = 10_000, 7, 100, 500      # 500 vs 250 per encode call: 5308 vs 5188 img/s (run 30)
# images per encode_image call by tower: ViT-B/32 has 50 tokens per image, so 2500 images make the GEMM M of 500 ViT-L/14 images
# (measured, same box: 74 352 img/s at 500 per call, 78 096 at 2500)
EMBED_BATCH_BY_MODEL = {"ViT-B/32": 2500}


def embed_batch_for(model: str, requested=None) -> int:
    return int(requested or EMBED_BATCH_BY_MODEL.get(model, EMBED_BATCH))
MODEL = "ViT-L/14"


def vit_flops_per_image(cfg) -> float:
    """per layer 2*L*d*3d (qkv) + 4*L^2*d (attention) + 2*L*d^2 (out) + 16*L*d^2 (MLP); + patch embed + proj."""
    L, d, g = cfg.tokens, cfg.width, cfg.grid
    layer = 2.0 * L * d * 3 * d + 4.0 * L * L * d + 2.0 * L * d * d + 16.0 * L * d * d
    return cfg.layers * layer + 2.0 * g * g * (3 * cfg.patch ** 2) * d + 2.0 * d * cfg.out_dim


def c2_workload(model: str, embed_batch: int) -> str:
    """config.workload of the retrieve line - shared by the b200 arm and the reference arm."""
    return (f"C2 per GPU: {N_CORPUS} synthetic 224^2 images (uint8 pixels, normalised on the GPU) -> CLIP {model} embed (batches of {embed_batch}) + "
            f"L2 normalise -> resident fp32 index shard; {N_QUERY} queries -> exact top-{TOP_K} "
            f"(sharded: per-shard top-k exchanged and merged) -> ResNet-50-stem style statistics of "
            f"{N_QUERY}x(1+{TOP_K}) 256^2 images; random-init weights")


def synth_images(n, res, seed, device, chunk=500, u8=False):
    """Image tensors [n,3,res,res] generated on the device (low-frequency structure + noise so embeddings are spread out):
    fp32 = what `preprocess(PIL)` stacks to; u8 = the resized / cropped uint8 pixels of the ingest path (`preprocess_u8`)."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    out = torch.empty((n, 3, res, res), dtype=torch.uint8 if u8 else torch.float32, device=device)
    for i in range(0, n, chunk):
        m = min(chunk, n - i)
        low = torch.randn((m, 3, 8, 8), generator=g, device=device)
        x = torch.nn.functional.interpolate(low, size=(res, res), mode="bilinear") \
            + 0.3 * torch.randn((m, 3, res, res), generator=g, device=device)
        out[i:i + m] = (x * 50 + 128).clamp_(0, 255).to(torch.uint8) if u8 else x
    return out


def measure(rank, world, local, MODEL, EMBED_BATCH, steps, warmup, with_e2e=True, launch_list_only=False):
    """The C2 job on an initialised process group; returns raw measurements (every rank)."""
    import ctypes as C

    import torch

    from domain_rag_b200 import _lib, clip
    from domain_rag_b200 import benchutil as B
    from domain_rag_b200.index import ShardedIndexFlatIP
    from domain_rag_b200.ingest import device_batches
    from domain_rag_b200.resnet import ResNetEncoder
    from domain_rag_b200.retrieval import rerank_by_style

    dev = torch.device("cuda", local)
    lib = _lib.load()
    model, _ = clip.load(MODEL, device=dev, seed=2000, max_batch=max(512, EMBED_BATCH))   # workspace for one whole encode call
    cfg = clip.CONFIGS[MODEL]
    stem = ResNetEncoder(seed=2000).to(dev).eval()
    corpus = synth_images(N_CORPUS, cfg.image, 1001 + rank, dev, u8=True)     # uint8 ingest
    queries = synth_images(N_QUERY, cfg.image, 1002, dev, u8=True)
    style_imgs = synth_images(N_QUERY * (1 + TOP_K), 256, 1003, dev, u8=True)      # [707,3,256,256] uint8 (/ 255 in the kernel)
    torch.cuda.synchronize()

    def job(corpus_src, query_src, style_src, to_host):
        """The C2 job. *_src are device tensors (value leg) or pinned host tensors (e2e leg)."""
        emb = torch.empty((N_CORPUS, cfg.out_dim), dtype=torch.float32, device=dev)
        i = 0
        for x in device_batches(corpus_src, EMBED_BATCH, dev):      # host source: batch i+1 crosses PCIe while batch i encodes
            emb[i:i + x.shape[0]] = model.encode_image(x, normalize=True)
            i += x.shape[0]
        ix = ShardedIndexFlatIP(cfg.out_dim, rank, world, device=local)
        ix.add_local(emb, lo=rank * N_CORPUS, ntotal_global=N_CORPUS * world)     # zero-copy: embeddings ARE the shard
        q = model.encode_image(query_src.to(dev, non_blocking=True), normalize=True)
        D, I = ix.search(q, TOP_K)
        feats = stem.style_features(style_src.to(dev, non_blocking=True))          # one launch for all 707 images
        if not to_host:
            return D, I, feats
        Dh, Ih, fh = D.cpu().numpy(), I.cpu().numpy(), feats.cpu().numpy()
        ranked = []
        for qi in range(N_QUERY):                                                    # host sort exactly as the reference
            first = [{"similarity": float(Dh[qi, j]), "image_path": str(int(Ih[qi, j])), "source_dataset": "coco"}
                     for j in range(TOP_K)]
            base = qi * (1 + TOP_K)
            ranked.append(rerank_by_style(fh[base], list(fh[base + 1: base + 1 + TOP_K]), first))
        return Dh, Ih, ranked

    for _ in range(max(1, min(warmup, 3))):
        job(corpus, queries, style_imgs, False)
    B.barrier(world)
    sampler = B.ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    B.barrier(world)
    torch.cuda.cudart().cudaProfilerStart()
    _lib.launch_count(reset=True)
    e0.record()
    for _ in range(steps):
        job(corpus, queries, style_imgs, False)
    e1.record()
    n_launches = _lib.launch_count()
    torch.cuda.cudart().cudaProfilerStop()
    B.barrier(world)
    total_ms = B.max_over_ranks(e0.elapsed_time(e1), world)
    m = {"ms_per_step": total_ms / steps, "clocks": sampler.stop() if rank == 0 else {}, "gpu_launches": int(n_launches),
         "cfg": cfg}
    if launch_list_only:
        return m

    # dominant kernels: event bracket around every GEMM / attention launch of one more job (same stream)
    lib.drag_prof_enable(1)
    job(corpus, queries, style_imgs, False)
    torch.cuda.synchronize()
    ms, work, cnt = (C.c_double * 2)(), (C.c_double * 2)(), (C.c_int * 2)()
    _lib.check(lib.drag_prof_collect(ms, work, cnt, 2), "drag_prof_collect")
    lib.drag_prof_enable(0)
    (m["g_ms"], m["g_fl"], m["g_n"]), (m["a_ms"], m["a_fl"], m["a_n"]) = [(ms[i], work[i], cnt[i]) for i in range(2)]

    # the stem-statistics kernel on its own (one launch over the 707 style images), CUDA events on the launching stream
    for _ in range(2):
        stem.style_features(style_imgs)
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(10):
        stem.style_features(style_imgs)
    s1.record()
    torch.cuda.synchronize()
    m["stem_ms"] = s0.elapsed_time(s1) / 10
    # SURVEY 8d counts 786 432 B in (fp32 pixels) + 512 B out per image; with uint8 ingest the kernel reads a quarter of that.
    # The roofline keeps the SURVEY figure as the algorithmic bytes (what the reference's fp32 tensor holds).
    m["stem_bytes"] = style_imgs.shape[0] * (3 * 256 * 256 * 4 + 128 * 4)

    if with_e2e:            # end to end: pinned host tensors in, ranked lists out
        corpus_h = torch.empty(corpus.shape, dtype=corpus.dtype, pin_memory=True)
        corpus_h.copy_(corpus)
        queries_h, style_h = queries.cpu().pin_memory(), style_imgs.cpu().pin_memory()
        job(corpus_h, queries_h, style_h, True)
        B.barrier(world)
        n_e2e = max(1, min(steps, 3))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            Dh, Ih, ranked = job(corpus_h, queries_h, style_h, True)
        B.barrier(world)
        m["e2e_s"] = B.max_over_ranks(time.perf_counter() - t0, world) / n_e2e
        assert len(ranked) == N_QUERY and len(ranked[0]) == TOP_K and ranked[0][0]["rank"] == 1
        m["h2d"] = (corpus_h.numel() * corpus_h.element_size() + queries_h.numel() * queries_h.element_size()
                    + style_h.numel() * style_h.element_size())
    return m


def stem_roofline(m, peaks):
    gbs = m["stem_bytes"] / (m["stem_ms"] * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": round(gbs, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": round(gbs / peaks["hbm_gbs"], 4), "kernel": "stem_stats_tc_kernel (707 images of 256^2, one launch; split-bf16 implicit GEMM on tcgen05)",
            "tensor_tflops": round(2 * 2.0 * 128 * 128 * 64 * 224 * 707 / (m["stem_ms"] * 1e-3) / 1e12, 1),
            "tensor_note": "uint8 pixels are exact bf16 operands: 2 split terms (x.wh + x.wl) x 14 k-steps of 16 per conv row",
            "kernel_ms": round(m["stem_ms"], 4), "traffic": None}


def run(args):
    from domain_rag_b200 import benchutil as B
    rank, world, local = B.dist_setup(args.gpus)
    MODEL = getattr(args, "clip_model", None) or globals()["MODEL"]     # ViT-L/14 (BASELINE) or ViT-B/32 (reference default)
    EMBED_BATCH = embed_batch_for(MODEL, getattr(args, "embed_batch", None))
    launch_only = os.environ.get("DRAG_BENCH_LAUNCH_LIST_ONLY") == "1"
    m = measure(rank, world, local, MODEL, EMBED_BATCH, args.steps, args.warmup, launch_list_only=launch_only)
    if launch_only:     # profiler runs: nothing after the timed region matters
        if rank == 0:
            print('{"launch_list_only": true}', flush=True)
        return None
    if rank != 0:
        return None
    cfg, ms_per_step = m["cfg"], m["ms_per_step"]
    g_ms, g_fl, g_n, a_ms, a_fl, a_n = m["g_ms"], m["g_fl"], m["g_n"], m["a_ms"], m["a_fl"], m["a_n"]
    n_img = N_CORPUS + N_QUERY
    flops = vit_flops_per_image(cfg) * n_img
    peaks = B.measured_peaks()
    peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
    return {
        "metric": f"C2 retrieval: corpus images embedded + indexed + queried per second ({MODEL}, top-100, style re-rank)",
        "value": round(n_img * world / (ms_per_step * 1e-3), 1), "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_per_step, 2), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": c2_workload(MODEL, EMBED_BATCH),
                   "l2_policy": "1.5 GB of uint8 images + ~90 GB of activations stream through per step (>> 126 MB L2)",
                   "flops_per_image": vit_flops_per_image(cfg),
                   "achieved_tflops": round(flops / (ms_per_step * 1e-3) / 1e12, 1),
                   "whole_job_frac_of_tensor_peak": round(flops / (ms_per_step * 1e-3) / 1e12 / peak, 4)},
        "e2e": {"value": round(n_img * world / m["e2e_s"], 1), "unit": "images/s", "h2d_bytes_per_step": int(m["h2d"]),
                "d2h_bytes_per_step": int(N_QUERY * TOP_K * 12 + N_QUERY * (1 + TOP_K) * 128 * 4)},
        "gpu_launches": m["gpu_launches"],
        "gpu_launches_note": f"every kernel of libdomainrag_b200.so launched inside the timed region; {int(g_n + a_n)} per step "
                             "are tcgen05 GEMM + attention",
        "clocks": m["clocks"],
        "roofline": {"bound": "tensor", "achieved": round(g_fl / (g_ms * 1e-3) / 1e12, 1), "peak": peak, "unit": "TFLOP/s",
                     "frac": round(g_fl / (g_ms * 1e-3) / 1e12 / peak, 4), "traffic": None,
                     "kernel": "gemm_bf16_tcgen05_2cta_kernel (all GEMM launches of one C2 job)",
                     "kernel_ms": round(g_ms / max(g_n, 1), 4), "launches": g_n, "share_of_step": round(g_ms / ms_per_step, 4),
                     "peak_source": peaks["source"] + " (sustained)",
                     "attention": {"achieved": round(a_fl / (a_ms * 1e-3) / 1e12, 1), "kernel_ms": round(a_ms / max(a_n, 1), 4),
                                   "launches": a_n, "share_of_step": round(a_ms / ms_per_step, 4)},
                     "stem_stats": stem_roofline(m, peaks)},
        **({"cpu_baseline": cpu_baseline(model=MODEL)} if world == 1 else {}),   # rank 0 at N = 1 only
    }


def measure_compact(rank, world, local, model: str = MODEL):
    """Compact C2 record for the default line's `secondary` block (2 timed jobs, no e2e leg)."""
    from domain_rag_b200 import benchutil as B
    m = measure(rank, world, local, model, embed_batch_for(model), 2, 1, with_e2e=False)
    peaks = B.measured_peaks()
    peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
    n_img = N_CORPUS + N_QUERY
    flops = vit_flops_per_image(m["cfg"]) * n_img
    ms = m["ms_per_step"]
    return {"workload": f"C2 per GPU: {N_CORPUS} images -> CLIP {model} embed -> resident index -> {N_QUERY} queries top-{TOP_K} "
                        "-> style re-rank statistics",
            "images_per_s": round(n_img * world / (ms * 1e-3), 1), "ms_per_job": round(ms, 2),
            "frac_of_tensor_peak": round(flops / (ms * 1e-3) / 1e12 / peak, 4),
            "gemm_tflops": round(m["g_fl"] / (m["g_ms"] * 1e-3) / 1e12, 1),
            "attention_tflops": round(m["a_fl"] / (m["a_ms"] * 1e-3) / 1e12, 1),
            "stem_stats": stem_roofline(m, peaks)}


def cpu_baseline(n_sample: int = 16, model: str = MODEL):
    """Oracle on the host cores: ViT-L/14 fp32 embed of a bounded sample, scaled linearly to the corpus (the scan and
    the stem statistics are < 1 % of the CPU time at C2 sizes and are timed on their full sizes)."""
    import numpy as np
    import torch

    from oracle import ip_topk as OI
    from oracle import stem as OS
    from oracle import vit as OV
    torch.set_num_threads(os.cpu_count())
    MODEL = model
    cfg = OV.CONFIGS[MODEL]
    state = OV.init_state(cfg, 2000)
    x = torch.randn(n_sample, 3, cfg.image, cfg.image, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        OV.embed(state, cfg, x[:2])
        t0 = time.perf_counter()
        OV.embed(state, cfg, x)
        t_img = (time.perf_counter() - t0) / n_sample
    g = np.random.default_rng(1)
    X = g.standard_normal((N_CORPUS, cfg.out_dim)).astype(np.float32)
    q = g.standard_normal((N_QUERY, cfg.out_dim)).astype(np.float32)
    t0 = time.perf_counter()
    OI.ip_topk(X, q, TOP_K)
    t_scan = time.perf_counter() - t0
    from domain_rag_b200.resnet import random_stem_state
    imgs = torch.rand(32, 3, 256, 256, generator=torch.Generator().manual_seed(2))
    t0 = time.perf_counter()
    OS.style_features(imgs, random_stem_state(2000))
    t_stem = (time.perf_counter() - t0) / 32 * N_QUERY * (1 + TOP_K)
    total = t_img * (N_CORPUS + N_QUERY) + t_scan + t_stem
    return {"value": round((N_CORPUS + N_QUERY) / total, 2), "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"oracle.vit.embed fp32 on {n_sample} images ({t_img * 1e3:.0f} ms/img) scaled to {N_CORPUS + N_QUERY}; "
                      f"oracle.ip_topk full size ({t_scan * 1e3:.0f} ms); oracle.stem on 32 images scaled to "
                      f"{N_QUERY * (1 + TOP_K)} ({t_stem:.1f} s)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    vals, cb = [], None
    m = getattr(args, "clip_model", None) or MODEL
    for _ in range(max(1, min(args.steps, 2))):
        cb = cpu_baseline(model=m)
        vals.append(cb["value"])
    val = sum(vals) / len(vals)
    cb["value"] = val
    return {"impl": "reference",
            "metric": f"C2 retrieval: corpus images embedded + indexed + queried per second ({m}, top-100, style re-rank)",
            "value": val, "unit": "images/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round((N_CORPUS + N_QUERY) / val * 1e3, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": c2_workload(m, embed_batch_for(m, getattr(args, "embed_batch", None))),
                       "sample": cb["sample"]},
            "cpu_baseline": cb, "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
