"""Attention micro-benchmark: the Flux shapes (head dim 128) against torch SDPA, and every head-dim-64 variant of the CLIP ViT
shapes selected through drag_debug_set (see include/domainrag_b200.h): same box, same tensors, CUDA-event timed."""
import sys, torch
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


def with_keys(keys, fn):
    """Time fn with drag_debug_set(key, value) applied, then restore the defaults."""
    defaults = {10: 0, 11: 1, 12: 1, 13: 1, 14: 0, 16: 0, 17: 1}
    for k, v in keys.items(): ops.debug_set(k, v)
    try:
        return t_ms(fn)
    finally:
        for k in keys: ops.debug_set(k, defaults[k])


for (B, H, S, hd) in [(1, 24, 5337, 128), (4, 24, 5337, 128), (1, 24, 2265, 128), (8, 24, 2265, 128),
                      (500, 16, 257, 64), (500, 12, 197, 64), (1024, 12, 50, 64)]:
    q = torch.randn(B, H, S, hd, device='cuda').bfloat16(); k = torch.randn_like(q); v = torch.randn_like(q)
    out = torch.empty(B * S, H * hd, device='cuda', dtype=torch.bfloat16)
    run = lambda: ops.attention(q, k, v, 0, out1=out)
    fl = 4.0 * B * H * S * S * hd
    ms = t_ms(run)
    ms_t = t_ms(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v))
    print(f"B={B} H={H} S={S} hd={hd}: ours {ms:.3f} ms {fl/ms/1e9:.0f} TFLOP/s | torch sdpa {ms_t:.3f} ms {fl/ms_t/1e9:.0f} TFLOP/s", flush=True)
    if hd == 64:
        rows = [("tiled online-softmax kernels (key 10 = 1)", {10: 1}),
                ("whole-row, one tile per CTA, L2 prefetch (12 = 0, 17 = 0)", {12: 0, 17: 0}),
                ("whole-row, one tile per CTA, no L2 prefetch (+ 11 = 0)", {12: 0, 17: 0, 11: 0}),
                ("persistent, query tiles start together (13 = 0)", {13: 0}),
                ("persistent, 2 of 8 exponentials on the FMA pipe (14 = 1)", {14: 1}),
                ("persistent, two softmax threads per row (16 = 1)", {16: 1})]
        for name, keys in rows:
            print(f"     {name}: {with_keys(keys, run):.3f} ms", flush=True)
