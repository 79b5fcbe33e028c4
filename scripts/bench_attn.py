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
for (B, H, S, hd) in [(1, 24, 5337, 128), (4, 24, 5337, 128), (1, 24, 2265, 128), (8, 24, 2265, 128), (500, 16, 257, 64), (1024, 12, 50, 64), (500, 12, 197, 64)]:
    q = torch.randn(B, H, S, hd, device='cuda').bfloat16(); k = torch.randn_like(q); v = torch.randn_like(q)
    out = torch.empty(B * S, H * hd, device='cuda', dtype=torch.bfloat16)
    ms = t_ms(lambda: ops.attention(q, k, v, 0, out1=out))
    ms_one = ms
    if hd == 64:      # A/B: the tiled online-softmax kernels instead of the whole-row kernel
        ops.debug_set(10, 1)
        ms_one = t_ms(lambda: ops.attention(q, k, v, 0, out1=out))
        ops.debug_set(10, 0)
        ops.debug_set(11, 0)
        ms_np = t_ms(lambda: ops.attention(q, k, v, 0, out1=out))
        ops.debug_set(11, 1)
        ops.debug_set(13, 0)
        ms_ns = t_ms(lambda: ops.attention(q, k, v, 0, out1=out))
        ops.debug_set(13, 1); ops.debug_set(14, 1)
        ms_poly = t_ms(lambda: ops.attention(q, k, v, 0, out1=out))
        ops.debug_set(14, 0); ops.debug_set(16, 1)
        ms_two = t_ms(lambda: ops.attention(q, k, v, 0, out1=out))
        ops.debug_set(16, 0)
        print(f"   persistent, two softmax threads per row: {ms_two:.3f} ms", flush=True)
        ops.debug_set(12, 0); ops.debug_set(17, 0)
        ms_row1 = t_ms(lambda: ops.attention(q, k, v, 0, out1=out))
        ops.debug_set(12, 1); ops.debug_set(17, 1)
        print(f"   variants: no L2 prefetch (one-tile kernel) {ms_np:.3f} | persistent without stagger {ms_ns:.3f} | persistent with 2/8 poly exp2 {ms_poly:.3f} | one-tile-per-CTA whole-row kernel {ms_row1:.3f} ms", flush=True)
    ms_t = t_ms(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v))
    fl = 4.0 * B * H * S * S * hd
    print(f"B={B} H={H} S={S} hd={hd}: ours {ms:.3f} ms {fl/ms/1e9:.0f} TFLOP/s (hd 64 tiled kernels: {ms_one:.3f} ms {fl/ms_one/1e9:.0f}) | torch sdpa {ms_t:.3f} ms {fl/ms_t/1e9:.0f} TFLOP/s", flush=True)
