"""Generate tests/golden/* by running the REFERENCE's own functions (TEST INFRASTRUCTURE ONLY).

Run in the authoring container, where /root/reference exists:  python oracle/make_golden.py
The reference scripts cannot be imported as they stand (clip, faiss, matplotlib, diffusers are not
installed), so empty stand-in modules are placed in sys.modules for exactly those names; none of
the functions exercised below touches them. torchvision's `resnet50(pretrained=True)` (needs a
download) is replaced by a random-init resnet50 carrying domain_rag_b200.resnet.random_stem_state.
Outputs are small fixtures committed under tests/golden/ (the reference does not travel to the
GPU box).
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys
import tempfile
import types
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parents[1]
REF = Path(os.environ.get("DRAG_REFERENCE", "/root/reference"))
GOLD = REPO / "tests" / "golden"
sys.path.insert(0, str(REPO))


def _stub(name: str, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def _load(path: Path, modname: str):
    spec = importlib.util.spec_from_file_location(modname, str(path))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_reference_retrieval(workdir: str):
    _stub("clip")
    _stub("faiss")
    mpl = _stub("matplotlib")
    mpl.pyplot = _stub("matplotlib.pyplot")
    cwd = os.getcwd()
    os.chdir(workdir)  # the script creates ./retrieval_results at import
    saved_env = os.environ.get("CUDA_VISIBLE_DEVICES")
    try:
        mod = _load(REF / "retrieval" / "clip100_resnet_style_all_shots.py", "ref_retrieval")
    finally:
        os.chdir(cwd)
        if saved_env is None:
            os.environ.pop("CUDA_VISIBLE_DEVICES", None)
        else:
            os.environ["CUDA_VISIBLE_DEVICES"] = saved_env
    return mod


def load_reference_outpaint():
    class _Dummy:  # names imported by the script but unused by the helpers below
        pass
    _stub("diffusers", FluxPriorReduxPipeline=_Dummy, FluxFillPipeline=_Dummy, ControlNetModel=_Dummy,
          StableDiffusionControlNetPipeline=_Dummy, FluxPipeline=_Dummy)
    _stub("diffusers.utils", load_image=lambda p: None)
    _stub("diffusers.pipelines")
    _stub("diffusers.pipelines.stable_diffusion", StableDiffusionSafetyChecker=_Dummy)
    return _load(REF / "outpainting_updown_sampling_redux.py", "ref_outpaint")


def synth_image_u8(seed: int, h: int, w: int) -> np.ndarray:
    """Smooth-ish synthetic RGB image (blobs + noise) so that resize / JPEG-free IO is non-trivial."""
    g = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.zeros((h, w, 3), np.float32)
    for c in range(3):
        for _ in range(4):
            cx, cy, s = g.uniform(0, w), g.uniform(0, h), g.uniform(10, 60)
            img[..., c] += g.uniform(0.2, 1.0) * np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * s * s))
    img += g.normal(0, 0.05, img.shape).astype(np.float32)
    img = (img - img.min()) / (img.max() - img.min() + 1e-8)
    return (img * 255).round().astype(np.uint8)


def main():
    import cv2
    import torchvision.models as tvm
    from PIL import Image

    from domain_rag_b200.resnet import random_stem_state

    GOLD.mkdir(parents=True, exist_ok=True)
    img_dir = GOLD / "images"
    img_dir.mkdir(exist_ok=True)

    # ---------------------------------------------------------------- retrieval-side functions
    with tempfile.TemporaryDirectory() as tmp:
        ref = load_reference_retrieval(tmp)
    state = random_stem_state(2000)
    real_resnet50 = tvm.resnet50

    def fake_resnet50(pretrained=False, **kw):
        torch.manual_seed(0)
        net = real_resnet50(weights=None)
        net.conv1.weight.data.copy_(state["conv1.weight"])
        net.bn1.weight.data.copy_(state["bn1.weight"])
        net.bn1.bias.data.copy_(state["bn1.bias"])
        net.bn1.running_mean.data.copy_(state["bn1.running_mean"])
        net.bn1.running_var.data.copy_(state["bn1.running_var"])
        return net

    ref.models.resnet50 = fake_resnet50
    model = ref.ResNetEncoder().eval()  # reference :51-64, :228
    ref.models.resnet50 = real_resnet50

    # (1) tensor path: ResNetEncoder.forward + calc_mean_std on seeded [4,3,256,256] inputs
    g = torch.Generator().manual_seed(1000)
    x = torch.rand(4, 3, 256, 256, generator=g)
    with torch.no_grad():
        feats = model(x)
        mean, std = ref.calc_mean_std(feats)
    stats = torch.cat([mean.flatten(1), std.flatten(1)], 1).numpy()

    # (2) file path: compute_resnet_features (cv2.imread -> RGB -> resize 256 -> /255 -> stem -> stats)
    paths, file_feats = [], []
    for i, (h, w) in enumerate([(256, 256), (300, 420), (180, 240), (512, 384), (256, 320), (200, 200)]):
        p = img_dir / f"synth_{i}.png"
        cv2.imwrite(str(p), cv2.cvtColor(synth_image_u8(10 + i, h, w), cv2.COLOR_RGB2BGR))
        paths.append(p)
        file_feats.append(ref.compute_resnet_features(str(p), model, "cpu"))
    np.savez(GOLD / "stem_stats.npz", input_seed=np.int64(1000), stats=stats.astype(np.float32),
             file_feats=np.stack(file_feats).astype(np.float32),
             file_names=np.array([p.name for p in paths]))

    # (3) resnet_second_stage_rerank on those files (query = image 0, candidates = 1..5, one missing)
    first_stage = [{"similarity": 0.9 - 0.1 * i, "image_path": str(paths[i]), "source_dataset": "coco",
                    "index": i} for i in range(1, 6)]
    first_stage.insert(2, {"similarity": 0.5, "image_path": str(img_dir / "missing.png"),
                           "source_dataset": "coco", "index": 99})
    ref.tqdm = lambda it, **kw: it
    rer = ref.resnet_second_stage_rerank(str(paths[0]), first_stage, model, "cpu")
    for r in rer:
        r["image_path"] = Path(r["image_path"]).name
    for r in first_stage:
        r["image_path"] = Path(r["image_path"]).name
    json.dump({"query": paths[0].name, "first_stage": first_stage, "reranked": rer},
              open(GOLD / "rerank.json", "w"), indent=1)

    # ---------------------------------------------------------------- composition-side helpers
    out = load_reference_outpaint()
    helpers = {"split": [], "resolution": [], "mask": []}
    for n, gpus in [(0, 4), (1, 4), (7, 1), (7, 2), (8, 8), (10, 3), (32, 8), (33, 8), (5, 8)]:
        helpers["split"].append({"n": n, "gpus": gpus,
                                 "out": out.split_samples_for_gpus([f"s{i}" for i in range(n)], gpus)})
    for (w, h) in [(640, 480), (1024, 768), (500, 375), (3000, 2000), (1024, 1024), (2800, 1200),
                   (333, 517), (800, 3000)]:
        im = Image.fromarray(synth_image_u8(77, min(h, 64), min(w, 64))).resize((w, h))
        try:
            p_im, up, down, need_up, need_down = out.process_image_resolution(im)
            helpers["resolution"].append({"size": [w, h], "out_size": list(p_im.size), "up": up, "down": down,
                                          "need_up": need_up, "need_down": need_down})
        except ValueError:
            helpers["resolution"].append({"size": [w, h], "error": True})
    mask_arrays = {}
    for i, (size, boxes) in enumerate([((64, 48), [(10, 12, 20, 16)]),
                                       ((64, 48), [(-5, -5, 20, 20), (50, 40, 30, 30)]),
                                       ((100, 80), [(0, 0, 100, 80)]),
                                       ((33, 21), [(3, 4, 5, 6), (20, 10, 8, 8), (32, 20, 4, 4)])]):
        m, _ = out.generate_outpaint_mask(Image.new("RGB", size), boxes)
        mask_arrays[f"mask_{i}"] = np.array(m)
        helpers["mask"].append({"size": list(size), "boxes": [list(b) for b in boxes], "key": f"mask_{i}"})
    # bicubic resize parity for up/down-scale helpers
    im = Image.fromarray(synth_image_u8(5, 48, 64))
    mask_arrays["upscale_1p7"] = np.array(out.upscale_image(im, 1.7))
    mask_arrays["downscale_1p7"] = np.array(out.downscale_image(out.upscale_image(im, 1.7), 1.7))
    mask_arrays["resize_src"] = np.array(im)
    json.dump(helpers, open(GOLD / "host_helpers.json", "w"), indent=1)
    np.savez_compressed(GOLD / "host_helpers_arrays.npz", **mask_arrays)
    print("golden fixtures written to", GOLD)


if __name__ == "__main__":
    main()
