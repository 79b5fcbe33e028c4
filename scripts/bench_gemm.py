"""GEMM microbenchmark on the B200: Flux / ViT shapes, CUDA-event timed, single-CTA vs CTA-pair kernel,
torch.matmul (cuBLAS) beside it. `--sustain` runs each shape for ~1.5 s (power-capped clocks)."""
import ctypes as C, json, sys, torch
sys.path.insert(0, '.')
from domain_rag_b200 import _lib, ops

sustain = "--sustain" in sys.argv
lib = _lib.load()

def t_ms(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

shapes = [(5337, 9216, 3072), (5337, 12288, 3072), (5337, 3072, 15360), (4096, 9216, 3072), (4096, 12288, 3072),
          (4096, 3072, 12288), (4096, 3072, 3072), (1241, 9216, 3072), (1241, 12288, 3072), (1241, 3072, 12288),
          (1241, 3072, 3072), (8192, 8192, 8192), (12800, 2304, 768), (16448, 4096, 1024)]
res = []
for M, N, K in shapes:
    a = torch.randn(M, K, device='cuda').bfloat16(); w = (torch.randn(N, K, device='cuda') * K ** -0.5).bfloat16()
    bias = torch.randn(N, device='cuda').bfloat16()
    out = torch.empty(M, N, device='cuda', dtype=torch.bfloat16)
    fl = 2.0 * M * N * K
    iters = max(10, int(1.5 / (fl / 1.2e15))) if sustain else 20
    row = {"M": M, "N": N, "K": K}
    for name, force in (("1cta", 1), ("2cta", 0)):
        lib.drag_debug_set(3, force)
        ms = t_ms(lambda: ops.linear(a, w, bias, mode=ops.EPI_GELU_TANH, out=out), iters)
        row[name + "_gelu_tflops"] = round(fl / ms / 1e9, 1)
        ms = t_ms(lambda: ops.linear(a, w, bias, out=out), iters)
        row[name + "_tflops"] = round(fl / ms / 1e9, 1)
    lib.drag_debug_set(3, 0)
    ms_t = t_ms(lambda: torch.matmul(a, w.t(), out=out), iters)
    row["cublas_tflops"] = round(fl / ms_t / 1e9, 1)
    res.append(row)
    print(row, flush=True)
    del a, w, out
json.dump(res, open('gpurun_out/gemm_bench%s.json' % ("_sustain" if sustain else ""), 'w'), indent=1)
