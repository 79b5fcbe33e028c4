"""QKV GEMM with the fused RMSNorm + RoPE epilogue vs the same GEMM with a plain bias epilogue (sustained, ~1 s per point)."""
import sys, torch
sys.path.insert(0, '.')
from domain_rag_b200 import ops

def t_ms(fn, iters, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

d, H, S = 3072, 24, 5337
for B in (1, 4):
    M = B * S
    rnd = lambda *s, sc=1.0: (torch.randn(*s, device='cuda') * sc).bfloat16()
    a, wq, bq = rnd(M, d), rnd(3 * d, d, sc=d ** -0.5), rnd(3 * d)
    q = torch.empty(B, H, S, 128, device='cuda', dtype=torch.bfloat16); k = torch.empty_like(q); v = torch.empty_like(q)
    qn, kn = rnd(128).float().mul(0.1).add(1).bfloat16(), rnd(128).float().mul(0.1).add(1).bfloat16()
    ang = torch.rand(S, 64, device='cuda') * 6.28
    cos, sin = ang.cos().contiguous(), ang.sin().contiguous()
    out = torch.empty(M, 3 * d, device='cuda', dtype=torch.bfloat16)
    fl = 2.0 * M * 3 * d * d
    iters = max(10, int(1.0 / (fl / 1.2e15)))
    ms_r = t_ms(lambda: ops.qkv_rope(a, wq, bq, q, k, v, qn, kn, cos, sin, 0, S), iters)
    ms_p = t_ms(lambda: ops.linear(a, wq, bq, out=out), iters)
    ms_c = t_ms(lambda: torch.matmul(a, wq.t(), out=out), iters)
    print(f"M={M}: qkv+rmsnorm+rope {fl/ms_r/1e9:.0f} TFLOP/s | plain bias {fl/ms_p/1e9:.0f} | cuBLAS {fl/ms_c/1e9:.0f}", flush=True)
