#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/scan_big.py <<'PY'
import sys, torch, numpy as np
sys.path.insert(0, '.')
from domain_rag_b200.index import IndexFlatIP
n = int(sys.argv[1]); d = 512
x = torch.randn(n, d, device='cuda'); x /= x.norm(dim=1, keepdim=True)
ix = IndexFlatIP(d); ix.add_device(x)
q = torch.randn(1, d, device='cuda'); q /= q.norm(dim=1, keepdim=True)
D, I = ix.search_device(q, 100); torch.cuda.synchronize()
ref = torch.topk(x @ q[0], 100)
print('n', n, 'match', torch.equal(ref.indices, I[0]), float((ref.values - D[0]).abs().max()))
PY
timeout 300 python /tmp/scan_big.py 1000000 > gpurun_out/r2_scan_big.log 2>&1; echo "rc=$?" >> gpurun_out/r2_scan_big.log
if ! grep -q "match True" gpurun_out/r2_scan_big.log; then
  timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python /tmp/scan_big.py 200000 > gpurun_out/r2_sanitizer.log 2>&1
fi
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
timeout 600 python bench.py --workload scan --steps 20 --warmup 3 > gpurun_out/r2_bench_scan.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2_bench_scan.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches.csv python bench.py --workload scan --steps 3 --warmup 3 > gpurun_out/r2_ncu.log 2>&1
tail -n 5 gpurun_out/r2_scan_big.log gpurun_out/r2_pytest.log gpurun_out/r2_bench_scan.log
exit 0
