"""Bring-up helper: try MN-major V descriptor (LBO, SBO) variants of the attention kernel vs SDPA."""
import sys, torch
sys.path.insert(0, '.')
from domain_rag_b200 import ops
torch.manual_seed(0)
B, H, S = 1, 2, 384
q = torch.randn(B, H, S, 128, device='cuda').bfloat16(); k = torch.randn_like(q); v = torch.randn_like(q)
ref = torch.nn.functional.scaled_dot_product_attention(q.float(), k.float(), v.float()).permute(0, 2, 1, 3).reshape(B, S, -1)
for lbo, sbo in [(0, 1024), (1024, 1024), (1024, 0), (16384, 1024), (1024, 16384), (2048, 1024), (128, 1024), (1024, 128)]:
    ops.debug_set(1, lbo); ops.debug_set(2, sbo)
    try:
        _, o = ops.attention(q, k, v, 0)
        torch.cuda.synchronize()
        rel = ((o.view(B, S, -1).float() - ref).norm() / ref.norm()).item()
        print(f"lbo={lbo} sbo={sbo} rel_l2={rel:.5f}", flush=True)
    except Exception as e:
        print(f"lbo={lbo} sbo={sbo} ERROR {e}", flush=True); break
