"""GEMM microbenchmark on the B200: Flux / ViT shapes, CUDA-event timed, torch.matmul beside it."""
import sys, json, torch
sys.path.insert(0, '.')
from domain_rag_b200 import ops

def t_ms(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

shapes = [(5337, 9216, 3072), (4096, 12288, 3072), (4096, 3072, 12288), (5337, 3072, 15360), (5337, 21504, 3072),
          (2265, 9216, 3072), (8192, 8192, 8192), (21348, 3072, 3072), (12800, 2304, 768), (64, 1056768, 3072)]
res = []
for M, N, K in shapes:
    a = torch.randn(M, K, device='cuda').bfloat16(); w = (torch.randn(N, K, device='cuda') * K ** -0.5).bfloat16()
    out = torch.empty(M, N, device='cuda', dtype=torch.bfloat16)
    ms = t_ms(lambda: ops.linear(a, w, out=out))
    ms_t = t_ms(lambda: torch.matmul(a, w.t(), out=out))
    fl = 2.0 * M * N * K
    res.append({"M": M, "N": N, "K": K, "ours_ms": round(ms, 4), "ours_tflops": round(fl / ms / 1e9, 1),
                "torch_ms": round(ms_t, 4), "torch_tflops": round(fl / ms_t / 1e9, 1)})
    print(res[-1], flush=True)
    del a, w, out
json.dump(res, open('gpurun_out/gemm_bench.json', 'w'), indent=1)
