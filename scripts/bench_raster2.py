"""Column-group raster A/B (drag_debug_set key 6) on the N = 3072 projections of a batch-4 Flux step, sustained."""
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
for M, N, K in [(21348, 3072, 15360), (21348, 3072, 12288), (21348, 3072, 3072), (16384, 3072, 12288), (4964, 3072, 12288),
                (5337, 3072, 15360), (21348, 12288, 3072), (21348, 9216, 3072)]:
    a = torch.randn(M, K, device='cuda').bfloat16(); w = (torch.randn(N, K, device='cuda') * K ** -0.5).bfloat16()
    bias = torch.randn(N, device='cuda').bfloat16()
    out = torch.empty(M, N, device='cuda', dtype=torch.bfloat16)
    fl = 2.0 * M * N * K
    iters = max(10, int(0.8 / (fl / 1.2e15)))
    row = {"M": M, "N": N, "K": K}
    for name, gn in (("auto", 0), ("n4", 4), ("n6", 6), ("n12", 12), ("n24", 24)):
        if gn > N // 256 and gn != 12: continue
        lib.drag_debug_set(6, gn)
        row[name + "_tflops"] = round(fl / t_ms(lambda: ops.linear(a, w, bias, out=out), iters) / 1e9, 1)
    for name, gn in (("n6h", 6), ("n12h", 12)):
        lib.drag_debug_set(6, gn); lib.drag_debug_set(7, 1)
        row[name + "_tflops"] = round(fl / t_ms(lambda: ops.linear(a, w, bias, out=out), iters) / 1e9, 1)
    lib.drag_debug_set(6, 0); lib.drag_debug_set(7, 0)
    row["cublas_tflops"] = round(fl / t_ms(lambda: torch.matmul(a, w.t(), out=out), iters) / 1e9, 1)
    res.append(row); print(row, flush=True)
    del a, w, out
json.dump(res, open('gpurun_out/raster2_bench.json', 'w'), indent=1)
