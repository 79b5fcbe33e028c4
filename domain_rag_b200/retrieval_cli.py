"""Driver of the two-stage retrieval - the logic behind the drop-in entry point
`retrieval/clip100_resnet_style_all_shots.py` (same flags, same cache / result files as the reference).

Reference behaviour mirrored (retrieval/clip100_resnet_style_all_shots.py):
  main :966-1107                          flags, per-(dataset, shot) loop, all_shots_retrieval_results.json
  load_or_compute_coco_features :499-657  cache order: global .pt -> --pretrained-* -> local .npy cache -> compute
  retrieve_by_category_multi_source :773-898   per-sample query, per-query JSON + visual, per-(dataset, shot) JSON
  get_inpainted_images :89-158            <lamainpaint>/<dataset>/<k>_shot/*.jpg (+ optional category_mapping.json)

What differs underneath: images are decoded by a thread pool and embedded in batches with one
H2D copy per batch (the reference runs batch 1 with a .cpu() sync per image); the corpus index is built once and
stays in HBM across queries (the reference re-adds N*D floats per query); the 101 style vectors of a query are one
launch. The query is embedded once (Appendix B). `--clip-top-k` is honoured (the reference ignores it; the
default 100 gives identical output). Visualisations are composed with PIL (matplotlib is not needed).
"""
from __future__ import annotations

import argparse
import glob
import json
import os
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import retrieval as R

RESULTS_DIR = "./retrieval_results"
LAMAINPAINT_DIR = "../lamainpaint"

# Tag of the encoder weights the feature caches of this run are computed with ("<model>:<sha1 of the weights file>" or
# "<model>:random-init"). Written next to every cache this driver creates (<cache>.meta.json); a cache whose tag differs is
# recomputed instead of reused, so features from a random-init dry run can never leak into a run with real weights.
# Caches without a sidecar (written by the reference itself) are trusted as they are.
_WEIGHTS_TAG = {"clip": None}


def _file_sha1(path: str) -> str:
    import hashlib
    h = hashlib.sha1()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 22), b""):
            h.update(chunk)
    return h.hexdigest()[:16]


def _write_cache_tag(cache_file: str) -> None:
    with open(cache_file + ".meta.json", "w") as f:
        json.dump({"clip_weights": _WEIGHTS_TAG["clip"]}, f)


def _cache_tag_ok(cache_file: str) -> bool:
    meta = cache_file + ".meta.json"
    if not os.path.exists(meta):
        return True
    try:
        with open(meta, "r") as f:
            tag = json.load(f).get("clip_weights")
    except Exception:
        return False
    if tag != _WEIGHTS_TAG["clip"]:
        print(f"缓存 {cache_file} 由不同的CLIP权重生成 ({tag} != {_WEIGHTS_TAG['clip']})，将重新计算")
        return False
    return True


# ------------------------------------------------------------------------------------ query side
def get_inpainted_images(dataset_name: str, shot_count: int, lamainpaint_dir: Optional[str] = None):
    """-> (sample_id -> jpg path, sample_id -> category); category = sample id unless category_mapping.json
    maps it (reference :89-158)."""
    shot_dir = os.path.join(lamainpaint_dir or LAMAINPAINT_DIR, dataset_name, f"{shot_count}_shot")
    if not os.path.exists(shot_dir):
        print(f"错误：找不到数据集 {dataset_name} 的 {shot_count}_shot 目录: {shot_dir}")
        return {}, {}
    files = glob.glob(os.path.join(shot_dir, "*.jpg"))
    if not files:
        print(f"错误：在 {shot_dir} 中找不到任何jpg图像")
        return {}, {}
    mapping = {}
    mapping_file = os.path.join(shot_dir, "category_mapping.json")
    if os.path.exists(mapping_file):
        try:
            with open(mapping_file, "r") as f:
                mapping = json.load(f)
            print(f"已加载类别映射文件: {mapping_file}")
        except Exception as e:  # the reference prints and carries on
            print(f"加载类别映射文件时出错: {e}")
    sample_to_image, sample_to_category = {}, {}
    for p in files:
        sid = os.path.splitext(os.path.basename(p))[0]
        sample_to_image[sid] = p
        sample_to_category[sid] = mapping.get(sid, sid)
    print(f"找到 {len(sample_to_image)} 个inpainted图像")
    return sample_to_image, sample_to_category


# ------------------------------------------------------------------------------------ embedding
def _load_preprocessed(path: str, preprocess):
    from PIL import Image
    try:
        return preprocess(Image.open(R.clean_image_path(path)).convert("RGB"))
    except Exception as e:
        print(f"处理图像时出错 {path}: {e}")
        return None


def embed_images(image_paths: Sequence[str], model, preprocess, device, batch_size: int = 256,
                 workers: int = 8) -> Tuple[np.ndarray, List[str]]:
    """CLIP-embed and L2-normalise `image_paths` in batches: decode/resize on a thread pool, ONE pinned H2D copy
    and one encode (one C call) per batch. Images cross PCIe as uint8 crops (a quarter of the float tensor's bytes);
    ToTensor + Normalize run inside the patch kernel with the formula `preprocess` uses, so the embeddings equal those of
    the reference's preprocess(PIL) -> float path. Unreadable images are skipped like the reference (:291-293).
    -> (float32 [n_valid, D], valid paths)."""
    import torch

    from . import clip as _clip
    if hasattr(model, "visual") and isinstance(model.visual, _clip.CLIPVisual):
        preprocess = _clip.preprocess_u8(model)          # same resize / crop; normalisation moves to the GPU
    feats, valid = [], []
    with ThreadPoolExecutor(max_workers=workers) as ex:
        for b0 in range(0, len(image_paths), batch_size):
            chunk = list(image_paths[b0:b0 + batch_size])
            tensors = list(ex.map(lambda p: _load_preprocessed(p, preprocess), chunk))
            keep = [(p, t) for p, t in zip(chunk, tensors) if t is not None]
            if not keep:
                continue
            x = torch.stack([t for _, t in keep]).pin_memory().to(device, non_blocking=True)
            f = model.encode_image(x, normalize=True)
            feats.append(f.float().cpu().numpy())
            valid.extend(R.clean_image_path(p) for p, _ in keep)
    if not feats:
        return np.zeros((0, 0), np.float32), []
    return np.concatenate(feats, 0), valid


def list_corpus_images(kind: str, root: str) -> List[str]:
    """File discovery of the reference: COCO = first existing of images/train2017/val2017, recursive
    jpg/jpeg/png (:263-281); Mini-ImageNet = train/<class>/*.{jpg,jpeg,png} (:909-924)."""
    paths: List[str] = []
    if kind == "coco":
        base = next((os.path.join(root, s) for s in ("images", "train2017", "val2017")
                     if os.path.isdir(os.path.join(root, s))), None)
        if base is None:
            print(f"错误：找不到COCO图像目录，已检查 {root}")
            return []
        for ext in ("*.jpg", "*.jpeg", "*.png"):
            paths.extend(str(p) for p in Path(base).glob(f"**/{ext}"))
    else:
        base = os.path.join(root, "train")
        if not os.path.isdir(base):
            print(f"错误：找不到Mini ImageNet图像目录，已检查 {base}")
            return []
        for cls in os.listdir(base):
            cdir = os.path.join(base, cls)
            if os.path.isdir(cdir):
                for ext in ("*.jpg", "*.jpeg", "*.png"):
                    paths.extend(str(p) for p in Path(cdir).glob(ext))
    return paths


def _load_pretrained(path: str, paths_json: Optional[str]):
    """--pretrained-*-features: .npy (+ json paths) or a .pt dict with embeddings/features (+ image_paths/paths).
    .pt files are read with weights_only=True (Appendix B: no arbitrary pickles)."""
    import torch
    feats, paths = None, None
    if path.endswith(".npy"):
        feats = np.load(path)
    elif path.endswith(".pt"):
        data = torch.load(path, map_location="cpu", weights_only=True)
        if isinstance(data, dict):
            feats = data.get("embeddings", data.get("features"))
            paths = data.get("image_paths", data.get("paths"))
        else:
            feats = data
        if feats is not None and hasattr(feats, "numpy"):
            feats = feats.float().numpy()
    if paths_json and os.path.exists(paths_json):
        with open(paths_json, "r") as f:
            paths = json.load(f)
    if paths is not None:
        paths = [R.clean_image_path(p) for p in paths]
    return feats, paths


def load_or_compute_corpus_features(kind: str, args, device, model, preprocess, results_dir: str):
    """-> (float32 [N,D], paths) or (None, None). kind in {"coco", "mini-imagenet"}."""
    stem = "coco" if kind == "coco" else "mini_imagenet"
    feat_file = os.path.join(results_dir, f"{stem}_clip_features.npy")
    path_file = os.path.join(results_dir, f"{stem}_image_paths.json")
    pre = args.pretrained_coco_features if kind == "coco" else args.pretrained_mini_imagenet_features
    pre_paths = args.pretrained_coco_paths if kind == "coco" else args.pretrained_mini_imagenet_paths
    feats, paths = None, None
    if args.force_recompute:
        print(f"强制重新计算{kind}特征...")
    else:
        if kind == "coco" and args.global_features:
            here = Path(sys.argv[0]).resolve().parent.parent
            for cand in (here / "coco_embeddings_global.pt", here / "result_clip_vision" / "coco_embeddings_global.pt"):
                if cand.exists():
                    try:
                        feats, paths = _load_pretrained(str(cand), None)
                        break
                    except Exception as e:
                        print(f"加载全局特征文件 {cand} 时出错: {e}")
        if feats is None and pre and os.path.exists(pre):
            try:
                feats, paths = _load_pretrained(pre, pre_paths)
                print(f"成功加载 {len(feats)} 个预提取特征")
            except Exception as e:
                print(f"加载预提取特征时出错: {e}")
                feats, paths = None, None
        if feats is None and os.path.exists(feat_file) and os.path.exists(path_file) and _cache_tag_ok(feat_file):
            try:
                feats = np.load(feat_file)
                with open(path_file, "r") as f:
                    paths = [R.clean_image_path(p) for p in json.load(f)]
                print(f"成功从本地缓存加载并清理 {len(feats)} 个{kind}特征")
            except Exception as e:
                print(f"加载本地缓存特征时出错: {e}")
                feats, paths = None, None
    if feats is None or paths is None or len(feats) == 0 or len(paths) == 0:
        root = args.coco_dir if kind == "coco" else args.mini_imagenet_dir
        images = list_corpus_images(kind, root)
        if not images:
            print(f"警告：没有{kind}特征")
            return None, None
        print(f"找到 {len(images)} 张{kind}图像，重新计算特征...")
        feats, paths = embed_images(images, model, preprocess, device)
        if len(feats) > 0:
            os.makedirs(results_dir, exist_ok=True)
            np.save(feat_file, feats)           # fp32 (the reference saves fp16 when computed on CUDA, then casts)
            with open(path_file, "w") as f:
                json.dump(paths, f)
            _write_cache_tag(feat_file)
    if feats is None or len(feats) == 0:
        print(f"警告：没有{kind}特征")
        return None, None
    return np.asarray(feats, dtype=np.float32), list(paths)


# ------------------------------------------------------------------------------------ outputs
def visualize_results(query_image_path: str, result_image_paths: Sequence[str], output_path: str, cell: int = 256):
    """Query + top results on a 3 x 4 contact sheet (the reference draws the same grid with matplotlib :352-393)."""
    from PIL import Image, ImageDraw
    sheet = Image.new("RGB", (4 * cell, 3 * cell + 3 * 18), "white")
    draw = ImageDraw.Draw(sheet)
    items = [("Query Image", query_image_path)] + [(f"Top {i + 1}", p) for i, p in enumerate(result_image_paths[:11])]
    for slot, (title, p) in enumerate(items):
        r, c = divmod(slot, 4)
        try:
            im = Image.open(R.clean_image_path(p)).convert("RGB")
            im.thumbnail((cell, cell))
        except Exception:
            print(f"警告：无法读取图像 {p}")
            continue
        y0 = r * (cell + 18)
        draw.text((c * cell + 4, y0 + 2), title, fill="black")
        sheet.paste(im, (c * cell + (cell - im.width) // 2, y0 + 18 + (cell - im.height) // 2))
    sheet.save(output_path)
    print(f"已保存可视化结果到 {output_path}")


def retrieve_by_category_multi_source(dataset_name, shot_count, clip_model, clip_preprocess, resnet_model,
                                      dataset_features, dataset_paths, device, force_recompute_inpainted=False,
                                      results_dir: Optional[str] = None, lamainpaint_dir: Optional[str] = None,
                                      top_k: int = 100, visualize: bool = True):
    """Reference :773-898. Returns {category: [{sample_id, image_path, category, similar_images}]} or None."""
    results_dir = results_dir or RESULTS_DIR
    sample_to_image, sample_to_category = get_inpainted_images(dataset_name, shot_count, lamainpaint_dir)
    if not sample_to_image:
        print(f"错误：找不到数据集 {dataset_name} 的 {shot_count}_shot inpainted图像")
        return None
    category_to_samples: Dict[str, List[str]] = {}
    for sid, cat in sample_to_category.items():
        category_to_samples.setdefault(cat, []).append(sid)
    print(f"为数据集 {dataset_name} ({shot_count}_shot) 找到 {len(category_to_samples)} 个类别")
    os.makedirs(results_dir, exist_ok=True)

    # query embeddings: cached exactly where the reference caches them (and, unlike it, reused as the queries)
    feat_file = os.path.join(results_dir, f"{dataset_name}_{shot_count}_shot_inpainted_clip_features.npy")
    path_file = os.path.join(results_dir, f"{dataset_name}_{shot_count}_shot_inpainted_image_paths.json")
    q_feats, q_paths = None, None
    if not force_recompute_inpainted and os.path.exists(feat_file) and os.path.exists(path_file) and _cache_tag_ok(feat_file):
        try:
            q_feats = np.load(feat_file)
            with open(path_file, "r") as f:
                q_paths = json.load(f)
        except Exception as e:
            print(f"加载inpainted特征时出错: {e}")
            q_feats, q_paths = None, None
    wanted = list(sample_to_image.values())
    if q_feats is None or q_paths is None or set(q_paths) != set(R.clean_image_path(p) for p in wanted):
        q_feats, q_paths = embed_images(wanted, clip_model, clip_preprocess, device)
        if len(q_feats) > 0:
            np.save(feat_file, q_feats)
            with open(path_file, "w") as f:
                json.dump(q_paths, f)
            _write_cache_tag(feat_file)
    q_by_path = {p: q_feats[i] for i, p in enumerate(q_paths)}

    all_results: Dict[str, list] = {}
    for category_name, samples in category_to_samples.items():
        print(f"\n处理类别: {category_name}")
        category_results = []
        for sid in samples:
            image_path = sample_to_image[sid]
            q = q_by_path.get(R.clean_image_path(image_path))
            if q is None:
                print(f"警告：无法从 {image_path} 提取CLIP特征")
                continue
            clip_results = R.clip_first_stage_retrieval(q, dataset_features, dataset_paths, top_k=top_k,
                                                        device=_device_index(device))
            if not clip_results:
                print("警告：CLIP检索未返回结果")
                continue
            final_results = R.resnet_second_stage_rerank(image_path, clip_results, resnet_model, device)
            if not final_results:
                print("警告：ResNet重排序未返回结果")
                continue
            result_file = os.path.join(results_dir, f"{dataset_name}_{shot_count}_shot_{category_name}_{sid}_retrieval_results.json")
            with open(result_file, "w", encoding="utf-8") as f:
                json.dump(final_results, f, indent=2, ensure_ascii=False)
            print(f"已保存样本 {sid} 的检索结果到 {result_file}")
            if visualize:
                visualize_results(image_path, [r["image_path"] for r in final_results[:10]],
                                  os.path.join(results_dir, f"{dataset_name}_{shot_count}_shot_{category_name}_{sid}_visual.jpg"))
            category_results.append({"sample_id": sid, "image_path": image_path, "category": category_name,
                                     "similar_images": final_results})
        if category_results:
            all_results.setdefault(category_name, []).extend(category_results)
    out_file = os.path.join(results_dir, f"{dataset_name}_{shot_count}_shot_retrieval_results.json")
    with open(out_file, "w", encoding="utf-8") as f:
        json.dump(all_results, f, indent=2, ensure_ascii=False)
    print(f"{dataset_name} {shot_count}_shot 所有类别的检索结果已合并保存到 {out_file}")
    return all_results


def _device_index(device) -> int:
    idx = getattr(device, "index", None)
    return int(idx) if idx is not None else 0


# ------------------------------------------------------------------------------------ CLI
def build_parser() -> argparse.ArgumentParser:
    """Flags of the reference (:967-996), same names, defaults and meaning."""
    p = argparse.ArgumentParser(description="CLIP+ResNet图像检索 - 多shot版本 (B200-native)")
    p.add_argument("--datasets", type=str, nargs="+", default=["ArTaxOr", "DIOR", "FISH", "NEU-DET", "UODD", "clipart1k"])
    p.add_argument("--shots", type=int, nargs="+", default=[1, 5, 10])
    p.add_argument("--coco-dir", type=str, default="./coco")
    p.add_argument("--mini-imagenet-dir", type=str, default="./miniimagenet")
    p.add_argument("--dataset-source", type=str, choices=["coco", "mini-imagenet", "both"], default="coco")
    p.add_argument("--clip-top-k", type=int, default=100)
    p.add_argument("--output-dir", type=str, default=None)
    p.add_argument("--gpu-id", type=int, default=0)
    p.add_argument("--pretrained-coco-features", type=str, default="./coco_embeddings_global.pt")
    p.add_argument("--pretrained-coco-paths", type=str, default=None)
    p.add_argument("--pretrained-mini-imagenet-features", type=str, default=None)
    p.add_argument("--pretrained-mini-imagenet-paths", type=str, default=None)
    p.add_argument("--global-features", action="store_true")
    p.add_argument("--force-recompute", action="store_true")
    p.add_argument("--lamainpaint-dir", type=str, default=None)
    p.add_argument("--force-recompute-inpainted", action="store_true")
    # additions (absent from the reference; defaults keep its behaviour)
    p.add_argument("--clip-model", type=str, default="ViT-B/32", help="ViT-B/32 (reference) or ViT-L/14")
    p.add_argument("--clip-weights", type=str, default=None,
                   help="OpenAI-format CLIP state dict (.pt, read with weights_only=True; `visual.*` keys) - what clip.load "
                        "downloads in the reference (:209). Required unless --allow-random-init")
    p.add_argument("--resnet-weights", type=str, default=None,
                   help="torchvision ResNet-50 state dict (.pt, weights_only=True; only conv1.weight and bn1.* are read) - "
                        "what resnet50(pretrained=True) downloads in the reference (:54). Required unless --allow-random-init")
    p.add_argument("--allow-random-init", action="store_true",
                   help="dry runs / CI without checkpoints: seeded random-init encoders. The feature caches written by such "
                        "a run are tagged and never reused by a run with real weights")
    p.add_argument("--no-visual", action="store_true", help="skip the *_visual.jpg contact sheets")
    return p


def main(argv: Optional[Sequence[str]] = None) -> int:
    import torch

    from . import clip
    from .resnet import ResNetEncoder
    args = build_parser().parse_args(argv)
    results_dir = args.output_dir or RESULTS_DIR
    os.makedirs(results_dir, exist_ok=True)
    lama_dir = args.lamainpaint_dir or LAMAINPAINT_DIR
    if not torch.cuda.is_available():
        print("错误：需要CUDA设备 (B200)；此实现没有CPU路径")
        return 1
    device = torch.device(f"cuda:{args.gpu_id}")
    torch.cuda.set_device(device)
    print(f"使用设备: {device}")
    missing = [flag for flag, v in (("--clip-weights", args.clip_weights), ("--resnet-weights", args.resnet_weights)) if not v]
    if missing and not args.allow_random_init:
        print(f"错误：缺少权重文件参数 {' '.join(missing)}；参考实现在此处下载预训练的CLIP / ResNet-50权重。"
              "如只需用随机初始化的编码器做流程测试，请显式传入 --allow-random-init")
        return 2
    state = None
    if args.clip_weights:
        state = torch.load(args.clip_weights, map_location="cpu", weights_only=True)
        _WEIGHTS_TAG["clip"] = f"{args.clip_model}:{_file_sha1(args.clip_weights)}"
        print(f"成功加载CLIP模型 {args.clip_model}")
    else:
        _WEIGHTS_TAG["clip"] = f"{args.clip_model}:random-init"
        print(f"警告：CLIP模型 {args.clip_model} 使用随机初始化权重 (--allow-random-init)，检索结果没有语义意义")
    clip_model, clip_preprocess = clip.load(args.clip_model, device=device, state_dict=state)
    stem_state = None
    if args.resnet_weights:
        sd = torch.load(args.resnet_weights, map_location="cpu", weights_only=True)
        need = ("conv1.weight", "bn1.weight", "bn1.bias", "bn1.running_mean", "bn1.running_var")
        lacking = [k for k in need if k not in sd]
        if lacking:
            print(f"错误：{args.resnet_weights} 缺少ResNet-50 stem参数: {lacking}")
            return 2
        stem_state = {k: sd[k].float() for k in need}
        print("成功加载ResNet特征提取器")
    else:
        print("警告：ResNet特征提取器使用随机初始化权重 (--allow-random-init)，风格重排序没有语义意义")
    resnet_model = ResNetEncoder(stem_state).to(device).eval()

    dataset_features, dataset_paths = {}, {}
    for kind in ("coco", "mini-imagenet"):
        if args.dataset_source in (kind, "both"):
            f, p = load_or_compute_corpus_features(kind, args, device, clip_model, clip_preprocess, results_dir)
            if f is not None and len(f) > 0:
                dataset_features[kind], dataset_paths[kind] = f, p
    if not dataset_features:
        print("错误：没有可用的数据集特征，无法进行检索")
        return 0

    all_shots_results: Dict[str, dict] = {}
    for dataset_name in args.datasets:
        all_shots_results[dataset_name] = {}
        for shot_count in args.shots:
            print(f"\n====== 处理数据集: {dataset_name}, {shot_count}_shot ======")
            shot_dir = os.path.join(lama_dir, dataset_name, f"{shot_count}_shot")
            if not os.path.isdir(shot_dir) or not glob.glob(os.path.join(shot_dir, "*.jpg")):
                print(f"警告：找不到数据集 {dataset_name} 的 {shot_count}_shot 目录或其中没有jpg图像: {shot_dir}")
                print(f"跳过数据集 {dataset_name} 的 {shot_count}_shot")
                continue
            res = retrieve_by_category_multi_source(dataset_name, shot_count, clip_model, clip_preprocess, resnet_model,
                                                    dataset_features, dataset_paths, device,
                                                    args.force_recompute_inpainted, results_dir=results_dir,
                                                    lamainpaint_dir=lama_dir, top_k=args.clip_top_k,
                                                    visualize=not args.no_visual)
            if res:
                all_shots_results[dataset_name][f"{shot_count}_shot"] = res
                print(f"完成数据集 {dataset_name} 的 {shot_count}_shot 检索")
            else:
                print(f"数据集 {dataset_name} 的 {shot_count}_shot 检索失败")
    if any(all_shots_results.values()):
        out = os.path.join(results_dir, "all_shots_retrieval_results.json")
        with open(out, "w", encoding="utf-8") as f:
            json.dump(all_shots_results, f, indent=2, ensure_ascii=False)
        print(f"所有数据集和所有shot的检索结果已合并保存到 {out}")
    else:
        print("没有成功检索任何数据集")
    return 0
