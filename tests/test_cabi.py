"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol the header declares
(no compute calls - there is no GPU here)."""
import re
import subprocess
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]


def header_symbols():
    text = (REPO / "include" / "domainrag_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(drag_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_symbols():
    syms = header_symbols()
    assert "drag_index_search" in syms and "drag_stem_stats" in syms and len(syms) >= 10


def test_library_exports_every_header_symbol(lib):
    from domain_rag_b200 import _lib
    syms = header_symbols()
    out = subprocess.run(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (drag_[a-z0-9_]+)", out))
    missing = [s for s in syms if s not in exported]
    assert not missing, f"not exported: {missing}"
    for s in syms:
        assert hasattr(lib, s)
    # the ctypes table covers the whole header
    assert sorted(_lib.SIGNATURES) == syms


def test_library_targets_sm100a_and_uses_bulk_copy(lib):
    from domain_rag_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    scan = sass[sass.index("ip_scan_topk_kernel"):]
    assert "UBLKCP" in scan  # cp.async.bulk staging of the corpus rows


def test_hot_kernels_are_blackwell_native_sass(lib):
    """the GEMM and attention kernels carry tcgen05 MMA (UTCHMMA), TMEM loads / stores (LDTM / STTM), TMA
    tensor loads (UTMALDG) and tcgen05.commit barriers (UTCBAR); the CTA-pair GEMM uses the .2CTA forms. Per-kernel
    counts are committed under profiles/r02_sass_summary.txt (scripts/sass_summary.py)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("sass_summary", str(REPO / "scripts" / "sass_summary.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    counts = mod.kernel_counts()
    gemm = {k: c for k, c in counts.items() if "gemm_bf16_tcgen05" in k}
    attn = {k: c for k, c in counts.items() if "attention_tcgen05_kernel" in k}
    assert len(gemm) >= 4 and len(attn) >= 2
    for name, c in {**gemm, **attn}.items():
        assert c["UTCHMMA"] >= 4 and c["LDTM"] >= 4 and c["UTMALDG"] >= 2 and c["UTCBAR"] >= 3, (name, dict(c))
        assert c["HMMA"] == 0, (name, "legacy mma.sync in a tcgen05 kernel")
    for name, c in attn.items():
        assert c["STTM"] >= 1 and c["FFMA2"] >= 32, (name, dict(c))          # P written back to TMEM; packed fp32x2 softmax
    row = {k: c for k, c in counts.items() if "attention_row" in k}          # whole-row kernels of the CLIP ViT path
    assert len(row) >= 2
    for name, c in row.items():
        assert c["UTCHMMA"] >= 2 and c["LDTM"] >= 3 and c["STTM"] >= 1 and c["UTMALDG"] >= 3 and c["HMMA"] == 0, (name, dict(c))
    pair = [c for k, c in gemm.items() if "2cta" in k]
    assert pair and all(c["UTCHMMA.2CTA"] >= 4 and c["UTCBAR.2CTA.MULTICAST"] >= 1 for c in pair)
    summary = (REPO / "profiles" / "r02_sass_summary.txt").read_text()
    assert "UTCHMMA" in summary and "attention_tcgen05_kernel" in summary


def test_no_device_is_reported_not_faked(lib):
    import ctypes as C
    import torch
    if torch.cuda.is_available():
        return
    h = C.c_void_p()
    rc = lib.drag_index_create(8, 0, C.byref(h))
    assert rc != 0 and b"no CUDA device" in lib.drag_last_error()


def test_product_package_never_imports_oracle():
    for p in (REPO / "domain_rag_b200").rglob("*.py"):
        src = p.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), p


def test_debug_set_environment_hook():
    """DRAG_DEBUG_SET="key=value,..." applies drag_debug_set knobs when the binding loads the library (same-box A/B runs of
    bench.py); an unknown key fails loudly instead of being ignored."""
    import subprocess
    import sys
    code = "from domain_rag_b200 import _lib; _lib.load(); print('loaded')"
    ok = subprocess.run([sys.executable, "-c", code], cwd=str(REPO), capture_output=True, text=True,
                        env={**__import__("os").environ, "DRAG_DEBUG_SET": "13=1,15=3"})
    assert ok.returncode == 0 and "loaded" in ok.stdout, ok.stderr[-400:]
    bad = subprocess.run([sys.executable, "-c", code], cwd=str(REPO), capture_output=True, text=True,
                         env={**__import__("os").environ, "DRAG_DEBUG_SET": "999=1"})
    assert bad.returncode != 0 and "unknown key" in bad.stderr
