"""A/B of the GEMM epilogue store width (drag_debug_set key 9: 1 = 16-byte stores, 0 = 32-byte STG.256 / LDG.256) on the
store-bound K = 1024 shapes of the CLIP ViT-L/14 blocks (B = 500 images) and on two Flux shapes (batch 4)."""
import sys, torch
sys.path.insert(0, '.')
from domain_rag_b200 import ops

def t_ms(fn, iters=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

rnd = lambda *s, sc=1.0: (torch.randn(*s, device='cuda') * sc).bfloat16()
shapes = [("vit fc     quick_gelu", 128500, 4096, 1024, ops.EPI_QUICK_GELU, False),
          ("vit qkv    bias      ", 128500, 3072, 1024, ops.EPI_BIAS, False),
          ("vit out    +resid    ", 128500, 1024, 1024, ops.EPI_GATE_RESID, True),
          ("vit proj   +resid    ", 128500, 1024, 4096, ops.EPI_GATE_RESID, True),
          ("flux mlp-up gelu_tanh", 21348, 12288, 3072, ops.EPI_GELU_TANH, False),
          ("flux out   gate+resid", 21348, 3072, 15360, ops.EPI_GATE_RESID, True)]
for name, M, N, K, mode, resid in shapes:
    a, w, b = rnd(M, K), rnd(N, K, sc=K ** -0.5), rnd(N)
    out = torch.empty(M, N, device='cuda', dtype=torch.bfloat16)
    x = rnd(M, N) if resid else None
    res = {}
    for key in (1, 0):
        ops.debug_set(9, key)
        res[key] = t_ms(lambda: ops.linear(a, w, b, mode=mode, out=(x if resid else out), resid=x))
    ops.debug_set(9, 0)
    fl = 2.0 * M * N * K
    print(f"{name} M={M} N={N} K={K}: 16-byte stores {res[1]:.3f} ms {fl/res[1]/1e9:.0f} TFLOP/s | 32-byte {res[0]:.3f} ms {fl/res[0]/1e9:.0f} TFLOP/s", flush=True)
    del a, w, b, out, x
