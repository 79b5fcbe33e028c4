"""Composition driver - the logic behind the drop-in entry point `outpainting_updown_sampling_redux.py`: keep the
foreground boxes of a few-shot image, regenerate everything else with Flux-Fill conditioned on a Redux prompt of each
retrieved-and-generated background, restore the original resolution, write the reference's file tree.

Reference behaviour mirrored (outpainting_updown_sampling_redux.py):
  flags                                   :1894-1911 (--multi_bbox / --resume family accepted)
  get_dataset_results / sample discovery  :799-826, :1716-1748  ./result/<ds>_<k>shot_retrieval/results_*/<sample>/
  annotations                             :545-642   ./datasets/<ds>/annotations/<k>_shot.json (COCO format), images in train/
  per-sample flow                         :1113-1322 original/bbox/upscaled/mask/bg/hires/final/params files, int() bbox scaling
  per-dataset tables                      :31-95     (hostlogic.dataset_params)
  result JSON + merge + collection        :1381-1603, :1750-1767, :1813-1886
  multi-GPU                               :157-177, :1605-1715 contiguous split, one process per GPU, merged JSON
Not reproduced: reloading all models for every sample (:1185) - pipelines are built once per process.
"""
from __future__ import annotations

import argparse
import glob
import json
import os
import random
import shutil
import socket
import time
import traceback
import uuid
from datetime import datetime
from typing import Callable, Dict, List, Optional, Sequence

from . import hostlogic as H

RESULT_DIR = "./result"
DATASETS_DIR = "./datasets"
DATASETS_1 = ["clipart1k", "NEU-DET", "ArTaxOr", "coco"]
DATASETS_2 = ["FISH", "UODD", "DIOR"]


def generate_process_id() -> str:
    return f"{socket.gethostname()}_{datetime.now().strftime('%Y%m%d_%H%M%S')}_{str(uuid.uuid4())[:8]}"


def dataset_path_name(dataset_name: str) -> str:
    return "NWPU_VHR_10" if dataset_name == "NWPU_VHR-10" else dataset_name      # reference :801-804


def get_dataset_results(dataset_name: str, shot_number: int, result_dir: str = RESULT_DIR) -> List[str]:
    base = os.path.join(result_dir, f"{dataset_path_name(dataset_name)}_{shot_number}shot_retrieval")
    if not os.path.isdir(base):
        print(f"警告：找不到数据集 {dataset_name} 的 {shot_number}shot 结果目录: {base}")
        return []
    return sorted(os.path.join(base, f) for f in os.listdir(base)
                  if f.startswith("results_") and os.path.isdir(os.path.join(base, f)))


def load_annotation_file(dataset_name: str, shot_number: int, datasets_dir: str = DATASETS_DIR):
    path = os.path.join(datasets_dir, dataset_name, "annotations", f"{shot_number}_shot.json")
    if not os.path.exists(path):
        print(f"警告：找不到注释文件: {path}")
        return None
    try:
        with open(path, "r") as f:
            return json.load(f)
    except Exception as e:
        print(f"读取注释文件出错: {e}")
        return None


def get_bbox_and_original_image(dataset_name, sample_id, shot_number, datasets_dir: str = DATASETS_DIR):
    """-> (original PIL, [bbox crops], [xywh], image_id, [category names]) or five Nones (reference :561-642)."""
    from PIL import Image
    ann = load_annotation_file(dataset_name, shot_number, datasets_dir)
    none5 = (None, None, None, None, None)
    if not ann:
        return none5
    by_name = {os.path.splitext(im["file_name"])[0]: im for im in ann["images"]}
    info = by_name.get(sample_id)
    if info is None:
        info = next((v for k, v in by_name.items() if sample_id in k or k in sample_id), None)
    if info is None:
        print(f"警告：找不到与样本ID {sample_id} 匹配的图像")
        return none5
    cats = {c["id"]: c["name"] for c in ann["categories"]}
    matched = [a for a in ann["annotations"] if str(a["image_id"]) == str(info["id"])]
    if not matched:
        print(f"警告：找不到图像ID {info['id']} 的注释")
        return none5
    path = os.path.join(datasets_dir, dataset_name, "train", info["file_name"])
    if not os.path.exists(path):
        print(f"警告：找不到图像文件: {path}")
        return none5
    original = Image.open(path).convert("RGB")
    crops, boxes, names = [], [], []
    for a in matched:
        boxes.append(a["bbox"])
        names.append(cats.get(a["category_id"], "unknown"))
        x, y, w, h = (int(float(c)) for c in a["bbox"])
        x, y = max(0, min(x, original.width - 1)), max(0, min(y, original.height - 1))
        w, h = max(1, min(w, original.width - x)), max(1, min(h, original.height - y))
        crops.append(original.crop((x, y, x + w, y + h)))
    return original, crops, boxes, info["id"], names


def find_backgrounds(sample_dir: str) -> List[str]:
    """Generated backgrounds of a sample: generated_image_rank{r}.png (or generated_image.png), rank order."""
    ranked = glob.glob(os.path.join(sample_dir, "generated_image_rank*.png"))
    if ranked:
        return sorted(ranked, key=lambda p: int("".join(c for c in os.path.basename(p).split("rank")[1] if c.isdigit()) or 0))
    single = os.path.join(sample_dir, "generated_image.png")
    return [single] if os.path.exists(single) else []


def process_sample_hires(dataset_name, sample_id, pipes, process_id, category_name=None, sample_dir=None, shot_number=1,
                         datasets_dir=DATASETS_DIR, result_dir=RESULT_DIR, outpaint_base="./outpaint_hires",
                         upscale_override: Optional[Dict[str, int]] = None, seed_fn: Optional[Callable[[], int]] = None,
                         num_inference_steps: int = 50, compose_batch: int = 1, seed: Optional[int] = None) -> dict:
    """One sample end to end; returns the reference's log record (status completed / error)."""
    from PIL import Image
    import torch
    print(f"处理样本 {sample_id} 从数据集 {dataset_name}，shot数: {shot_number}")
    t0 = time.time()
    prefix = f"{dataset_name}_{sample_id}_{shot_number}shot"
    log = {"dataset": dataset_name, "sample_id": sample_id, "sample_prefix": prefix, "category": category_name,
           "shot_number": shot_number, "status": "started", "error": None, "original_resolution": None,
           "upscaled_resolution": None, "downscaled_resolution": None, "up_scale_factor": None, "down_scale_factor": None,
           "was_upscaled": False, "was_downscaled": False, "image_id": None, "original_image_size": None,
           "bbox_image_sizes": None, "bbox_coords_list": None, "outpainted_images": [], "process_time_seconds": 0}
    try:
        original, crops, boxes, image_id, categories = get_bbox_and_original_image(dataset_name, sample_id, shot_number,
                                                                                   datasets_dir)
        if sample_dir is None:
            for rf in get_dataset_results(dataset_name, shot_number, result_dir):
                if os.path.isdir(os.path.join(rf, sample_id)):
                    sample_dir = os.path.join(rf, sample_id)
                    break
        if original is None:
            if sample_dir is None or not os.path.exists(os.path.join(sample_dir, "target_input.png")):
                raise ValueError(f"找不到样本 {sample_id} 的原始图像")
            original = Image.open(os.path.join(sample_dir, "target_input.png")).convert("RGB")
            W0, H0 = original.size                       # default box: centred 30 % (reference :943-947)
            bw, bh = int(W0 * 0.3), int(H0 * 0.3)
            boxes = [[(W0 - bw) // 2, (H0 - bh) // 2, bw, bh]]
            crops = [original.crop((boxes[0][0], boxes[0][1], boxes[0][0] + bw, boxes[0][1] + bh))]
            categories = [category_name] if category_name else ["unknown"]
        if sample_dir is None:
            raise ValueError(f"找不到样本 {sample_id} 的结果目录")
        bg_images = find_backgrounds(sample_dir)
        if not bg_images:
            raise ValueError(f"样本 {sample_id} 没有可用的背景图像")
        log.update(image_id=image_id, original_resolution=original.size, original_image_size=original.size,
                   bbox_image_sizes=[c.size if c else None for c in crops], bbox_coords_list=boxes)
        out_dir = os.path.join(outpaint_base, f"process_{process_id}", dataset_name, f"{shot_number}_shot", sample_id)
        os.makedirs(out_dir, exist_ok=True)
        orig_path = os.path.join(out_dir, f"{prefix}_original.png")
        original.save(orig_path)
        saved = []
        for i, c in enumerate(crops):
            p = os.path.join(out_dir, f"{prefix}_bbox{i + 1}_original.jpg")
            c.save(p)
            saved.append(p)
        log["bbox_saved_paths"] = saved

        prm = H.dataset_params(dataset_name)
        min_dim = (upscale_override or {}).get(dataset_name, prm.upscale_dimension)
        try:
            processed, up, down, was_up, was_down = H.process_image_resolution(original, min_dimension=min_dim,
                                                                               max_dimension=H.MAX_DIMENSION)
        except ValueError as e:
            raise ValueError(f"样本 {sample_id} 处理失败: {e}")
        log.update(upscaled_resolution=processed.size, up_scale_factor=up, down_scale_factor=down, was_upscaled=was_up,
                   was_downscaled=was_down, min_dimension_used=min_dim)
        if was_down:
            processed.save(os.path.join(out_dir, f"{prefix}_downscaled_bg.png"))
            log["downscaled_resolution"] = processed.size
        if was_up:
            processed.save(os.path.join(out_dir, f"{prefix}_upscaled_bg.png"))
        factor = up if was_up else (down if was_down else None)
        proc_boxes = [[int(c * factor) for c in b] for b in boxes] if factor is not None else boxes
        mask_image, _ = H.generate_outpaint_mask(processed, proc_boxes)

        # Every background of a sample is composed onto the SAME processed image and mask, so they run as batches of up
        # to `compose_batch` through one FluxFillPipeline call, each composition with its own seed / CPU generator
        # exactly as in the reference's one-at-a-time loop (outpainting...:1185-1300).
        jobs = []
        for bg_idx, bg_path in enumerate(bg_images):
            bg_name = os.path.basename(bg_path)
            suffix = f"_{bg_name.split('rank')[1].split('.')[0]}" if "rank" in bg_name else f"_{bg_idx + 1}"
            mask_path = os.path.join(out_dir, f"{prefix}_mask{suffix}.png")
            mask_image.save(mask_path)
            try:
                bg_image = Image.open(bg_path).convert("RGB")
            except Exception as e:
                print(f"加载背景图像 {bg_path} 失败: {e}")
                continue
            bg_saved = os.path.join(out_dir, f"{prefix}_bg{suffix}_original.png")
            shutil.copy(bg_path, bg_saved)
            # `seed` (--seed, an int: crosses the spawn boundary of --multi_gpu) > seed_fn > a fresh random seed per
            # composition like the reference (:1230-1231)
            job_seed = seed if seed is not None else (seed_fn() if seed_fn else random.randint(0, 2 ** 32 - 1))
            jobs.append(dict(bg_idx=bg_idx, bg_path=bg_path, bg_name=bg_name, suffix=suffix, mask_path=mask_path,
                             bg_image=bg_image, bg_saved=bg_saved, seed=job_seed))
        step = max(1, int(compose_batch))
        for c0 in range(0, len(jobs), step):
            chunk = jobs[c0:c0 + step]
            priors = [pipes.prior_redux([j["bg_image"]], prompt=prm.redux_prompt, prompt_2="",
                                        prompt_embeds_scale=[prm.image_prompt_scale], pooled_prompt_embeds_scale=[1.0])
                      for j in chunk]
            gens = [torch.Generator("cpu").manual_seed(j["seed"]) for j in chunk]
            if len(chunk) == 1:
                results = pipes.pipe_fill(image=processed, mask_image=mask_image, height=processed.height,
                                          width=processed.width, guidance_scale=prm.guidance_scale,
                                          num_inference_steps=num_inference_steps, prompt_embeds=priors[0].prompt_embeds,
                                          pooled_prompt_embeds=priors[0].pooled_prompt_embeds, generator=gens[0],
                                          strength=prm.strength).images
            else:
                results = pipes.pipe_fill(image=[processed] * len(chunk), mask_image=[mask_image] * len(chunk),
                                          height=processed.height, width=processed.width,
                                          guidance_scale=prm.guidance_scale, num_inference_steps=num_inference_steps,
                                          prompt_embeds=torch.cat([p_.prompt_embeds for p_ in priors]),
                                          pooled_prompt_embeds=torch.cat([p_.pooled_prompt_embeds for p_ in priors]),
                                          generator=gens, strength=prm.strength).images
            for j, result in zip(chunk, results):
                bg_idx, bg_path, bg_name, suffix = j["bg_idx"], j["bg_path"], j["bg_name"], j["suffix"]
                mask_path, bg_saved, job_seed = j["mask_path"], j["bg_saved"], j["seed"]
                hires_path = os.path.join(out_dir, f"{prefix}_hires_result{suffix}.png")
                result.save(hires_path)
                final = result
                if was_up:
                    final = H.downscale_image(result, up)
                elif was_down:
                    final = H.upscale_image(result, 1.0 / down)
                final_path = os.path.join(out_dir, f"{prefix}_final_result{suffix}.png")
                final.save(final_path)
                params = {"categories": categories, "image_scale": 1.0, "prompt_scale": 1.0,
                          "image_prompt_scale": prm.image_prompt_scale, "guidance_scale": prm.guidance_scale,
                          "num_inference_steps": num_inference_steps, "strength": prm.strength,
                          "redux_prompt": prm.redux_prompt, "seed": job_seed, "process_id": process_id,
                          "shot_number": shot_number, "bg_index": bg_idx, "bg_filename": bg_name,
                          "original_bg_path": bg_path, "copied_bg_path": bg_saved,
                          "original_resolution": {"width": original.width, "height": original.height},
                          "processed_resolution": {"width": processed.width, "height": processed.height},
                          "min_dimension_used": min_dim, "up_scale_factor": up, "down_scale_factor": down,
                          "was_upscaled": was_up, "was_downscaled": was_down, "bbox_coords_list": boxes,
                          "processed_bbox_coords_list": proc_boxes, "image_id": image_id if image_id else "unknown",
                          "num_bbox": len(boxes)}
                params_path = os.path.join(out_dir, f"{prefix}_params{suffix}.json")
                with open(params_path, "w") as f:
                    json.dump(params, f, indent=2)
                log["outpainted_images"].append({"original_bg_path": bg_path, "copied_bg_path": bg_saved,
                                                 "hires_result_path": hires_path, "final_result_path": final_path,
                                                 "mask_path": mask_path, "params_path": params_path,
                                                 "bbox_coords_list": boxes, "processed_bbox_coords_list": proc_boxes,
                                                 "params": params})
        log["original_saved_path"] = orig_path
        log["status"] = "completed"
    except Exception as e:
        log["status"], log["error"] = "error", str(e)
        print(f"处理样本 {sample_id} 时出错: {e}")
        traceback.print_exc()
    finally:
        log["process_time_seconds"] = time.time() - t0
        done = "处理完成" if log["status"] == "completed" else "处理失败"
        print(f"样本 {sample_id} {done}，耗时 {log['process_time_seconds']:.2f} 秒")
    return log


def get_all_sample_ids(dataset_name, shot_number, result_dir=RESULT_DIR, datasets_dir=DATASETS_DIR) -> List[str]:
    ids = set()
    for rf in get_dataset_results(dataset_name, shot_number, result_dir):
        ids.update(d for d in os.listdir(rf) if os.path.isdir(os.path.join(rf, d)))
    ann = load_annotation_file(dataset_name, shot_number, datasets_dir) if not ids else None
    if ann:
        ids.update(os.path.splitext(im.get("file_name", ""))[0] for im in ann.get("images", []))
    return sorted(ids)


def formatted_result_json(dataset_name, logs, shot_number, process_id) -> dict:
    ok = [l for l in logs if l["status"] == "completed"]
    return {"dataset": dataset_name, "timestamp": datetime.now().strftime("%Y-%m-%d %H:%M:%S"), "process_id": process_id,
            "shot_number": shot_number, "total_samples": len(logs), "successful_samples": len(ok),
            "failed_samples": len(logs) - len(ok), "samples": logs}


def save_formatted_result_json(dataset_name, result_json, shot_number, process_id, outpaint_base="./outpaint_hires") -> str:
    d = os.path.join(outpaint_base, f"process_{process_id}", dataset_name, f"{shot_number}_shot")
    os.makedirs(d, exist_ok=True)
    path = os.path.join(d, f"outpaint_results_{shot_number}shot.json")
    with open(path, "w") as f:
        json.dump(result_json, f, indent=2)
    print(f"已保存格式化的结果JSON文件: {path}")
    return path


def merge_gpu_results(dataset_name, gpu_jsons: Sequence[dict], shot_number, process_id) -> dict:
    if not gpu_jsons:
        return formatted_result_json(dataset_name, [], shot_number, process_id)
    merged = dict(gpu_jsons[0])
    merged.update(samples=[], multi_gpu=True, num_gpus=len(gpu_jsons), gpu_process_ids=[])
    for g in gpu_jsons:
        merged["samples"].extend(g.get("samples", []))
        merged["gpu_process_ids"].append(g.get("gpu_process_id", ""))
    merged["total_samples"] = len(merged["samples"])
    merged["successful_samples"] = sum(1 for s in merged["samples"] if s["status"] == "completed")
    merged["failed_samples"] = merged["total_samples"] - merged["successful_samples"]
    return merged


def copy_final_results_to_collection(process_id, shot_number=None, outpaint_base="./outpaint_hires",
                                     final_base="./final_results") -> str:
    coll = os.path.join(final_base, f"process_{process_id}") + (f"/{shot_number}_shot" if shot_number is not None else "")
    os.makedirs(coll, exist_ok=True)
    root = os.path.join(outpaint_base, f"process_{process_id}")
    n = 0
    for ds in (os.listdir(root) if os.path.isdir(root) else []):
        for shot_dir in os.listdir(os.path.join(root, ds)):
            sp = os.path.join(root, ds, shot_dir)
            if not os.path.isdir(sp) or (shot_number is not None and not shot_dir.startswith(f"{shot_number}_shot")):
                continue
            dst = os.path.join(coll, ds, shot_dir)
            os.makedirs(dst, exist_ok=True)
            for f in glob.glob(os.path.join(sp, "*", "*_final_result*.png")):
                shutil.copy(f, os.path.join(dst, os.path.basename(f)))
                n += 1
    print(f"已将 {n} 个最终结果文件复制到集合目录: {coll}")
    return coll


def parse_resume_log(log_file: str):
    """Sample ids a previous run finished / failed. Accepts the script's own lines ("样本 X 处理完成，耗时…" /
    "样本 X 处理失败…" / "处理样本 X 时出错"), which the reference's parser never matched."""
    done, failed = set(), set()
    with open(log_file, "r", errors="replace") as f:
        for line in f:
            if line.startswith("样本 ") and " 处理完成" in line:
                done.add(line.split()[1])
            elif line.startswith("样本 ") and " 处理失败" in line:
                failed.add(line.split()[1])
            elif line.startswith("处理样本 ") and "时出错" in line:
                failed.add(line.split()[1])
    return done, failed - done


def _worker(rank, gpu_lists, dataset_name, shot_number, process_id, kwargs, load_kwargs, out_queue):
    import torch
    from .models import load_model
    torch.cuda.set_device(rank)
    pipes = load_model(device=f"cuda:{rank}", want=("fill",), **load_kwargs)
    logs = [process_sample_hires(dataset_name, sid, pipes, process_id, shot_number=shot_number, **kwargs)
            for sid in gpu_lists[rank]]
    res = formatted_result_json(dataset_name, logs, shot_number, process_id)
    res["gpu_process_id"] = f"{process_id}_gpu{rank}"
    out_queue.put((rank, res))


def _dead_rank_result(dataset_name, sample_ids, shot_number, process_id, rank, exitcode) -> dict:
    """Result JSON of a worker that died before reporting (OOM, load_model error ...): every one of its samples is an
    error record, so the merged JSON and --failed_only still account for them."""
    msg = f"GPU {rank} 工作进程异常退出 (exit code {exitcode})"
    logs = [{"dataset": dataset_name, "sample_id": sid, "sample_prefix": f"{dataset_name}_{sid}_{shot_number}shot",
             "category": None, "shot_number": shot_number, "status": "error", "error": msg, "outpainted_images": [],
             "process_time_seconds": 0} for sid in sample_ids]
    res = formatted_result_json(dataset_name, logs, shot_number, process_id)
    res["gpu_process_id"] = f"{process_id}_gpu{rank}"
    return res


def collect_worker_results(procs: Dict[int, object], q, poll_seconds: float = 1.0) -> Dict[int, Optional[dict]]:
    """Drain one (rank, result) per worker from `q` while watching the workers: a rank whose process has exited without
    reporting maps to None instead of blocking the parent forever. `procs`: rank -> object with is_alive() / exitcode."""
    import queue as _queue
    results: Dict[int, Optional[dict]] = {}
    pending = set(procs)
    while pending:
        try:
            rank, res = q.get(timeout=poll_seconds)
            results[rank] = res
            pending.discard(rank)
            continue
        except _queue.Empty:
            pass
        for r in list(pending):
            if not procs[r].is_alive():
                try:                                   # its result may have landed between the timeout and this check
                    while True:
                        rank, res = q.get(timeout=0.2)
                        results[rank] = res
                        pending.discard(rank)
                except _queue.Empty:
                    pass
                if r in pending:
                    print(f"错误：GPU {r} 的工作进程已退出 (exit code {procs[r].exitcode})，未返回结果")
                    results[r] = None
                    pending.discard(r)
    return results


def process_dataset_samples_multi_gpu(dataset_name, sample_ids, shot_number, num_gpus, process_id, kwargs, load_kwargs):
    import torch.multiprocessing as mp
    lists = H.split_samples_for_gpus(sample_ids, num_gpus)
    for i, s in enumerate(lists):
        print(f"GPU {i}: {len(s)} 个样本")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = {r: ctx.Process(target=_worker, args=(r, lists, dataset_name, shot_number, process_id, kwargs, load_kwargs, q))
             for r in range(num_gpus) if lists[r]}
    for p in procs.values():
        p.start()
    results = collect_worker_results(procs, q)
    for p in procs.values():
        p.join()
    jsons = [results[r] if results[r] is not None else
             _dead_rank_result(dataset_name, lists[r], shot_number, process_id, r, procs[r].exitcode) for r in sorted(results)]
    return merge_gpu_results(dataset_name, jsons, shot_number, process_id)


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="高分辨率Outpainting处理脚本 (B200-native)")
    p.add_argument("--dataset", type=str)
    p.add_argument("--dataset_group", type=str, choices=["1", "2", "all"], default="all")
    p.add_argument("--sample_id", type=str)
    p.add_argument("--shot", type=int, default=1, choices=[1, 2, 3, 5, 10, 20])
    p.add_argument("--min_dimension", type=int, default=1024)
    p.add_argument("--custom_upscale", type=str)
    p.add_argument("--process_id", type=str)
    p.add_argument("--collect_only", action="store_true")
    p.add_argument("--multi_bbox", action="store_true")
    p.add_argument("--resume", action="store_true")
    p.add_argument("--log_file", type=str)
    p.add_argument("--failed_only", action="store_true")
    p.add_argument("--multi_gpu", action="store_true")
    p.add_argument("--num_gpus", type=int)
    # additions (defaults keep the reference's behaviour)
    p.add_argument("--weights_dir", type=str, default="./model")
    p.add_argument("--model_size", type=str, default="full", choices=["full", "tiny"])
    p.add_argument("--num_inference_steps", type=int, default=50)
    p.add_argument("--seed", type=int, default=None, help="fixed seed instead of a random one per composition")
    p.add_argument("--allow_random_init", action="store_true",
                   help="dry run without checkpoints: seeded random-init models and synthetic text tokens (noise images); "
                        "without it a missing weight file is an error")
    p.add_argument("--compose_batch", type=int, default=4,
                   help="backgrounds of one sample composed per FluxFillPipeline call (same image and mask; 1 = the "
                        "reference's one-at-a-time loop; seeds stay per composition either way)")
    return p


def main(argv=None) -> int:
    import torch
    args = build_parser().parse_args(argv)
    process_id = args.process_id or generate_process_id()
    if args.collect_only:
        copy_final_results_to_collection(process_id, args.shot)
        return 0
    upscale_override = {}
    if args.custom_upscale:
        try:
            ds, dim = args.custom_upscale.split(":")
            upscale_override[ds] = int(dim)
        except Exception as e:
            print(f"解析自定义上采样维度时出错: {e}")
    if args.multi_bbox:
        print("已启用多bbox支持，将处理同一图像上的多个物体")
    datasets = [args.dataset] if args.dataset else {"1": DATASETS_1, "2": DATASETS_2, "all": DATASETS_1 + DATASETS_2}[args.dataset_group]
    done, failed = set(), set()
    if (args.resume or args.failed_only) and args.log_file and os.path.exists(args.log_file):
        done, failed = parse_resume_log(args.log_file)
    n_gpus = args.num_gpus or (torch.cuda.device_count() if args.multi_gpu else 1)
    load_kwargs = dict(weights_dir=args.weights_dir, size=args.model_size, max_side=H.MAX_DIMENSION,
                       max_batch=max(1, args.compose_batch), allow_random_init=args.allow_random_init)
    from .models import missing_files
    lacking = missing_files(args.weights_dir, ("fill",))
    if lacking and not args.allow_random_init:
        print(f"错误：权重目录 {args.weights_dir} 缺少 {lacking}；如只做流程测试请显式传入 --allow_random_init")
        return 2
    kwargs = dict(upscale_override=upscale_override, num_inference_steps=args.num_inference_steps, seed=args.seed,
                  compose_batch=max(1, args.compose_batch))
    pipes = None
    for ds in datasets:
        ids = [args.sample_id] if args.sample_id else get_all_sample_ids(ds, args.shot)
        if args.failed_only:
            ids = [i for i in ids if i in failed]
        elif args.resume:
            ids = [i for i in ids if i not in done]
        if not ids:
            print(f"警告：数据集 {ds} 没有找到任何样本")
            continue
        if args.multi_gpu and n_gpus > 1:
            res = process_dataset_samples_multi_gpu(ds, ids, args.shot, n_gpus, process_id, kwargs, load_kwargs)
        else:
            if pipes is None:
                from .models import load_model
                pipes = load_model(want=("fill",), **load_kwargs)
            logs = [process_sample_hires(ds, sid, pipes, process_id, shot_number=args.shot, **kwargs) for sid in ids]
            res = formatted_result_json(ds, logs, args.shot, process_id)
        save_formatted_result_json(ds, res, args.shot, process_id)
    copy_final_results_to_collection(process_id, args.shot)
    return 0
