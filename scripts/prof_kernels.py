"""Short launch sequence for `ncu --set full`: the MLP-up GEMM (GELU epilogue), the single-block out projection
(K = 15360, gate+residual epilogue), the QKV GEMM with RMSNorm+RoPE epilogue and the joint attention, at the
shapes of one Flux step at 1024^2 (S = 5337)."""
import sys, torch
sys.path.insert(0, '.')
from domain_rag_b200 import ops
torch.manual_seed(0)
BATCH = int(sys.argv[1]) if len(sys.argv) > 1 else 1      # compositions per GPU (bench.py --batch)
S_TOK, d, H = 5337, 3072, 24
M = BATCH * S_TOK
dev = 'cuda'
rnd = lambda *s, sc=1.0: (torch.randn(*s, device=dev) * sc).bfloat16()
a, w1, b1 = rnd(M, d), rnd(4 * d, d, sc=d ** -0.5), rnd(4 * d)
wide = torch.empty(M, 5 * d, device=dev, dtype=torch.bfloat16)
w2, b2, x, gate = rnd(d, 5 * d, sc=(5 * d) ** -0.5), rnd(d), rnd(M, d), rnd(BATCH, d)
wq, bq = rnd(3 * d, d, sc=d ** -0.5), rnd(3 * d)
q = torch.empty(BATCH, H, S_TOK, 128, device=dev, dtype=torch.bfloat16); k = torch.empty_like(q); v = torch.empty_like(q)
qn, kn = rnd(128).float().mul(0.1).add(1).bfloat16(), rnd(128).float().mul(0.1).add(1).bfloat16()
ang = torch.rand(S_TOK, 64, device=dev) * 6.28
cos, sin = ang.cos().contiguous(), ang.sin().contiguous()
for it in range(3):
    ops.linear(a, w1, b1, mode=ops.EPI_GELU_TANH, out=wide[:, d:])
    ops.qkv_rope(a, wq, bq, q, k, v, qn, kn, cos, sin, 0, S_TOK)
    ops.attention(q, k, v, 0, out1=wide[:, :d])
    ops.linear(wide, w2, b2, mode=ops.EPI_GATE_RESID, resid=x, gate=gate, rows_per_batch=S_TOK, out=x)
torch.cuda.synchronize()
print("ok")
