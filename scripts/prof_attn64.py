"""Launch sequence for ncu: head-dim-64 attention at the ViT-L/14 shape (500 images x 16 heads x 257 tokens)."""
import sys, torch
sys.path.insert(0, '.')
from domain_rag_b200 import ops
S = int(sys.argv[1]) if len(sys.argv) > 1 else 257
B, H = 500, 16
q = torch.randn(B, H, S, 64, device='cuda').bfloat16(); k = torch.randn_like(q); v = torch.randn_like(q)
out = torch.empty(B * S, H * 64, device='cuda', dtype=torch.bfloat16)
for _ in range(3):
    ops.attention(q, k, v, 0, out1=out)
torch.cuda.synchronize()
print("ok")
