"""Tile-raster A/B on the B200: plain row-fastest order (drag_debug_set(4, 1<<20)) vs the grouped order (default)
at the batched Flux shapes, sustained (~1 s per point), cuBLAS beside it."""
import json, sys, torch
sys.path.insert(0, '.')
from domain_rag_b200 import _lib, ops
lib = _lib.load()

def t_ms(fn, iters, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

res = []
for M, N, K in [(5337, 12288, 3072), (5337, 3072, 15360), (10674, 12288, 3072), (10674, 3072, 15360), (21348, 12288, 3072),
                (21348, 9216, 3072), (21348, 3072, 15360), (21348, 3072, 12288)]:
    a = torch.randn(M, K, device='cuda').bfloat16(); w = (torch.randn(N, K, device='cuda') * K ** -0.5).bfloat16()
    bias = torch.randn(N, device='cuda').bfloat16()
    out = torch.empty(M, N, device='cuda', dtype=torch.bfloat16)
    fl = 2.0 * M * N * K
    iters = max(10, int(1.0 / (fl / 1.2e15)))
    row = {"M": M, "N": N, "K": K}
    for name, g in (("rowfast", 1 << 20), ("grouped", 0), ("g8", 8), ("g4", 4)):
        lib.drag_debug_set(4, g)
        row[name + "_tflops"] = round(fl / t_ms(lambda: ops.linear(a, w, bias, out=out), iters) / 1e9, 1)
    lib.drag_debug_set(4, 0)
    row["cublas_tflops"] = round(fl / t_ms(lambda: torch.matmul(a, w.t(), out=out), iters) / 1e9, 1)
    res.append(row); print(row, flush=True)
    del a, w, out
json.dump(res, open('gpurun_out/raster_bench.json', 'w'), indent=1)
