"""CPU checks of the measurement contract: the reference arm prints one JSON line with the agreed keys (it times the
CPU oracle, never the CUDA library), and the algorithmic-work figures the rooflines are computed from match SURVEY 8d."""
import json
import subprocess
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline", "impl"}


def _run(*flags):
    res = subprocess.run([sys.executable, str(REPO / "bench.py"), "--impl", "reference", *flags], capture_output=True,
                         text=True, timeout=600, cwd=REPO)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, res.stdout
    return json.loads(lines[0])


def test_reference_arm_scan_line():
    d = _run("--workload", "scan", "--steps", "1", "--warmup", "0")
    assert BASE_KEYS <= set(d) and d["impl"] == "reference" and d["unit"] == "GB/s" and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["vs_baseline"] is None and "workload" in d["config"]


def test_reference_arm_other_ranks_stay_silent(monkeypatch):
    env = dict(**__import__("os").environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, str(REPO / "bench.py"), "--impl", "reference", "--workload", "scan", "--gpus", "2",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300, cwd=REPO, env=env)
    assert res.returncode == 0 and not [ln for ln in res.stdout.splitlines() if ln.startswith("{")]


def test_algorithmic_work_matches_survey():
    sys.path.insert(0, str(REPO))
    import bench_compose
    import bench_scan
    import bench_retrieve
    from domain_rag_b200 import clip
    # scan bytes = N*D*4 + nq*D*4 + nq*k*12 (2.048 GB at N=1e6, D=512)
    assert bench_scan.scan_algorithmic_bytes(1_000_000, 512, 1, 100) == 2_048_000_000 + 2048 + 1200
    # 88.85 TFLOP per MMDiT forward at 1024^2 (GEMM 68.90 + attention 19.95); 32.83 TFLOP at 512^2
    g, a = bench_compose.flops_per_forward(4096)
    assert abs(g / 1e12 - 68.90) < 0.01 and abs(a / 1e12 - 19.95) < 0.01
    g, a = bench_compose.flops_per_forward(1024)
    assert abs((g + a) / 1e12 - 32.83) < 0.01
    # 8 x 512^2 x 20 steps = 5.253 PFLOP; headline: 4.443 PFLOP per 1024^2 image at 50 steps
    assert abs(8 * 20 * (g + a) / 1e15 - 5.253) < 0.002
    g, a = bench_compose.flops_per_forward(4096)
    assert abs(50 * (g + a) / 1e15 - 4.443) < 0.002
    assert bench_compose.WORKLOADS["c3"] == dict(side=512, T=20, batch=8)
    assert bench_scan.SWEEP_N == (10_000, 30_000, 100_000, 300_000, 1_000_000) and bench_scan.SWEEP_D == (512, 768)
    # 162.0 GFLOP per ViT-L/14 image, 8.82 GFLOP per ViT-B/32 image
    assert abs(bench_retrieve.vit_flops_per_image(clip.CONFIGS["ViT-L/14"]) / 1e9 - 162.0) < 0.1
    assert abs(bench_retrieve.vit_flops_per_image(clip.CONFIGS["ViT-B/32"]) / 1e9 - 8.82) < 0.02


def test_reference_arm_c3_line_shares_the_workload_string():
    """The reference arm of every compose-type workload names the same config.workload as the b200 arm."""
    sys.path.insert(0, str(REPO))
    import bench_compose
    assert "512^2" in bench_compose.full_workload(8, 512, 20) and "20 MMDiT steps" in bench_compose.full_workload(8, 512, 20)
    assert "C4 per-GPU slice" in bench_compose.full_workload(4)
