"""Background generation driver - the logic behind the drop-in entry point `batch_generate_flux_kshot.py`: for every
few-shot sample and each of its top-5 retrieved images, Redux-blend (retrieved 0.8, target 1.0) and sample a 1024^2
background with 50 Flux steps at guidance 2.5, seed 0.

Reference behaviour mirrored (batch_generate_flux_kshot.py):
  flags                          :33-45
  generate_image                 :439-524   prior([ref, target], ["", ""], scales [0.8, 1.0] / [1.0, 1.0]) -> pipe(2.5, 50,
                                            1024, 1024, Generator("cpu").manual_seed(0)); generated_image_rank{r}.png,
                                            params.txt (once), ref_info{rank}{_sim}.txt, ref_input{rank}.jpg, target_input.png
  k-shot driver                  :766-1046  <out>/<ds>_<k>shot_retrieval/results_coco_0.8_target_1.0_cocotext_1.0_targettext_1.0_<ts>/
                                            <sample>/…, batch_params.txt (the size-statistics section stays empty because the
                                            reader looks for ref_info{rank}.txt, a name never written - SURVEY 8f N4), error.txt /
                                            generation_failed.txt
Retrieval input: <retrieval_results_dir>/<ds>_<k>_shot_retrieval_results.json or all_shots_retrieval_results.json as written
by retrieval/clip100_resnet_style_all_shots.py ({category: [{sample_id, image_path, similar_images: [{rank, similarity,
image_path}]}]}).
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
from datetime import datetime
from typing import List, Optional, Tuple

COCO_IMAGE_SCALE, TARGET_IMAGE_SCALE, COCO_TEXT_SCALE, TARGET_TEXT_SCALE = 0.8, 1.0, 1.0, 1.0
PROMPT_RETRIEVAL = ""
LAMAINPAINT_DIR = "./lamainpaint"
DATASET_GROUPS = {"dataset1": ["UODD", "ArTaxOr", "FISH", "coco"], "dataset2": ["DIOR", "NEU-DET", "clipart1k"]}
ALL_DATASETS = ["FISH", "DIOR", "ArTaxOr", "UODD", "NEU-DET", "clipart1k", "coco"]


def _redux_prior(pipe_prior_redux, coco_image, target_image):
    """The two-image Redux call of reference :459-465 (scaled rows summed; pooled = 2 x CLIP(""))."""
    return pipe_prior_redux([coco_image, target_image], prompt=["", PROMPT_RETRIEVAL], prompt_2=["", PROMPT_RETRIEVAL],
                            prompt_embeds_scale=[COCO_IMAGE_SCALE, TARGET_IMAGE_SCALE],
                            pooled_prompt_embeds_scale=[COCO_TEXT_SCALE, TARGET_TEXT_SCALE])


def _write_outputs(image, coco_image_path, target_image_path, target_image, output_path, rank, similarity, database_type,
                   num_inference_steps) -> None:
    """Files of reference :476-519 for one generated image."""
    width, height = target_image.size
    height, width = max((height // 16) * 16, 64), max((width // 16) * 16, 64)     # derived but unused (Appendix B)
    out_dir = os.path.dirname(output_path)
    os.makedirs(out_dir, exist_ok=True)
    image.save(output_path)
    params_file = os.path.join(out_dir, "params.txt")
    if not os.path.exists(params_file):
        with open(params_file, "w") as f:
            f.write(f"数据库类型: {database_type}\n参考图像权重: {COCO_IMAGE_SCALE}\n目标图像权重: {TARGET_IMAGE_SCALE}\n"
                    f"参考文本权重: {COCO_TEXT_SCALE}\n目标文本权重: {TARGET_TEXT_SCALE}\n提示词: {PROMPT_RETRIEVAL}\n"
                    f"指导比例: 2.5\n推理步数: {num_inference_steps}\n生成图像尺寸: {width}x{height}\n"
                    f"原始图像尺寸: {target_image.size[0]}x{target_image.size[1]}\n")
    rank_str = f"rank{rank}" if rank is not None else ""
    sim_str = f"_sim{similarity:.4f}" if similarity is not None else ""
    with open(os.path.join(out_dir, f"ref_info{rank_str}{sim_str}.txt"), "w") as f:
        f.write(f"数据库类型: {database_type}\n参考图像: {coco_image_path}\n目标图像: {target_image_path}\n"
                f"生成图像尺寸: {width}x{height}\n原始图像尺寸: {target_image.size[0]}x{target_image.size[1]}\n")
        if rank is not None:
            f.write(f"排名: {rank}\n")
        if similarity is not None:
            f.write(f"相似度: {similarity}\n")
    target_out = os.path.join(out_dir, "target_input.png")
    if not os.path.exists(target_out):
        shutil.copy(target_image_path, target_out)
    shutil.copy(coco_image_path, os.path.join(out_dir, f"ref_input{rank_str}.jpg"))


def generate_image(pipe_prior_redux, pipe, coco_image_path, target_image_path, output_path, rank=None, similarity=None,
                   database_type="coco", num_inference_steps=50, size=1024) -> bool:
    """Reference :439-524. Returns False (after printing) on any failure, like the reference."""
    from PIL import Image
    import torch
    try:
        coco_image = Image.open(coco_image_path).convert("RGB")
        target_image = Image.open(target_image_path).convert("RGB")
        prior = _redux_prior(pipe_prior_redux, coco_image, target_image)
        images = pipe(guidance_scale=2.5, num_inference_steps=num_inference_steps, height=size, width=size,
                      generator=torch.Generator("cpu").manual_seed(0), **prior).images
        _write_outputs(images[0], coco_image_path, target_image_path, target_image, output_path, rank, similarity,
                       database_type, num_inference_steps)
        return True
    except Exception as e:
        print(f"生成图像时出错: {e}")
        return False


def generate_images_batch(pipe_prior_redux, pipe, jobs, database_type="coco", num_inference_steps=50, size=1024) -> List[bool]:
    """Several (reference image, target image) generations of the same 1024^2 shape as ONE FluxPipeline call. `jobs` =
    [(coco_image_path, target_image_path, output_path, rank, similarity)]. Every generation keeps the reference's fixed
    seed 0 through its own CPU generator (reference :472), so the images equal the one-at-a-time loop's. A failing batch
    falls back to generate_image per job, which keeps the reference's print-and-continue behaviour."""
    from PIL import Image
    import torch
    if len(jobs) == 1:
        c, t, o, r, sm = jobs[0]
        return [generate_image(pipe_prior_redux, pipe, c, t, o, r, sm, database_type, num_inference_steps, size)]
    try:
        loaded = [(Image.open(c).convert("RGB"), Image.open(t).convert("RGB")) for c, t, _, _, _ in jobs]
        priors = [_redux_prior(pipe_prior_redux, ci, ti) for ci, ti in loaded]
        images = pipe(guidance_scale=2.5, num_inference_steps=num_inference_steps, height=size, width=size,
                      generator=[torch.Generator("cpu").manual_seed(0) for _ in jobs],
                      prompt_embeds=torch.cat([p_["prompt_embeds"] for p_ in priors]),
                      pooled_prompt_embeds=torch.cat([p_["pooled_prompt_embeds"] for p_ in priors])).images
        for (c, t, o, r, sm), (_, ti), im in zip(jobs, loaded, images):
            _write_outputs(im, c, t, ti, o, r, sm, database_type, num_inference_steps)
        return [True] * len(jobs)
    except Exception as e:
        print(f"批量生成图像时出错 ({e})，改为逐张生成")
        return [generate_image(pipe_prior_redux, pipe, c, t, o, r, sm, database_type, num_inference_steps, size)
                for c, t, o, r, sm in jobs]


def load_retrieval_results(retrieval_results_dir: str, dataset_name: str, shot_number: int) -> Optional[dict]:
    """{category: [records]} for one (dataset, shot) from the per-shot file or the all-shots file."""
    per = os.path.join(retrieval_results_dir, f"{dataset_name}_{shot_number}_shot_retrieval_results.json")
    if os.path.exists(per):
        with open(per, "r") as f:
            return json.load(f)
    allf = os.path.join(retrieval_results_dir, "all_shots_retrieval_results.json")
    if os.path.exists(allf):
        with open(allf, "r") as f:
            return json.load(f).get(dataset_name, {}).get(f"{shot_number}_shot")
    return None


def top_similar_images(results: dict, sample_name: str, limit: int = 5) -> List[Tuple[float, str, int]]:
    """[(similarity, image_path, rank)] of the sample's record, ranks 1..limit, existing files only."""
    for records in results.values():
        for rec in records:
            if rec.get("sample_id") == sample_name:
                out = []
                for s in rec.get("similar_images", []):
                    p, r = s.get("image_path", ""), int(s.get("rank", 0))
                    if p and os.path.exists(p) and 1 <= r <= limit:
                        out.append((float(s.get("similarity", 0)), p, r))
                return sorted(out, key=lambda t: t[2])
    return []


def process_kshot_dataset_with_retrieval(dataset_name, pipe_prior_redux, pipe, results, shot_number, output_dir,
                                         lamainpaint_dir=LAMAINPAINT_DIR, database_type="coco", num_inference_steps=50,
                                         size=1024, sample_filter=None, gen_batch: int = 1) -> Optional[str]:
    shot_dir = os.path.join(lamainpaint_dir, dataset_name, f"{shot_number}_shot")
    if not os.path.isdir(shot_dir):
        print(f"错误：找不到k-shot目录 {shot_dir}")
        return None
    names = sorted(os.path.splitext(f)[0] for f in os.listdir(shot_dir) if f.endswith(".jpg"))
    if sample_filter is not None:
        names = [n for n in names if n in sample_filter]
    if not names:
        print(f"跳过数据集 {dataset_name} {shot_number}-shot，因为找不到样本")
        return None
    result_dir = f"{output_dir}/{dataset_name}_{shot_number}shot_retrieval"
    ts = datetime.now().strftime("%Y%m%d_%H%M%S")
    base = os.path.join(result_dir, f"results_coco_{COCO_IMAGE_SCALE}_target_{TARGET_IMAGE_SCALE}_cocotext_{COCO_TEXT_SCALE}"
                                    f"_targettext_{TARGET_TEXT_SCALE}_{ts}")
    os.makedirs(base, exist_ok=True)
    with open(os.path.join(base, "batch_params.txt"), "w") as f:
        f.write(f"数据集: {dataset_name} ({shot_number}-shot，使用检索结果)\nCOCO图像权重: {COCO_IMAGE_SCALE}\n"
                f"目标图像权重: {TARGET_IMAGE_SCALE}\nCOCO文本权重: {COCO_TEXT_SCALE}\n目标文本权重: {TARGET_TEXT_SCALE}\n"
                f"提示词: {PROMPT_RETRIEVAL}\n指导比例: 2.5\n推理步数: {num_inference_steps}\n处理样本数: {len(names)}\n"
                f"为每个样本生成: 最多10张图像 (基于相似度最高的COCO图像)\n图像尺寸: 动态调整至与目标图像匹配 (保证是16的倍数)\n")
    ok = bad = total = 0
    for name in names:
        target = os.path.join(shot_dir, f"{name}.jpg")
        sdir = os.path.join(base, name)
        os.makedirs(sdir, exist_ok=True)
        tops = top_similar_images(results, name)
        if not tops:
            print(f"跳过样本 {name}，因为找不到相似图像")
            with open(os.path.join(sdir, "error.txt"), "w") as f:
                f.write(f"处理样本 {name} 时出错: 找不到相似图像\n时间: {datetime.now().strftime('%Y-%m-%d %H:%M:%S')}\n")
            bad += 1
            continue
        any_ok = False
        jobs = [(ref, target, os.path.join(sdir, f"generated_image_rank{rank}.png"), rank, sim) for sim, ref, rank in tops]
        step = max(1, int(gen_batch))
        for c0 in range(0, len(jobs), step):
            chunk = jobs[c0:c0 + step]
            done = generate_images_batch(pipe_prior_redux, pipe, chunk, database_type=database_type,
                                         num_inference_steps=num_inference_steps, size=size)
            for (_, _, _, rank, _), good in zip(chunk, done):
                if good:
                    print(f"成功生成样本 {name} 的图像 (rank {rank})")
                    total += 1
                    any_ok = True
                else:
                    print(f"生成样本 {name} 的图像失败 (rank {rank})")
        if any_ok:
            ok += 1
        else:
            bad += 1
            with open(os.path.join(sdir, "generation_failed.txt"), "w") as f:
                f.write(f"生成样本 {name} 的图像失败\n时间: {datetime.now().strftime('%Y-%m-%d %H:%M:%S')}\n")
    with open(os.path.join(base, "batch_params.txt"), "a") as f:
        f.write(f"成功处理样本数: {ok}\n失败处理样本数: {bad}\n总共生成图像数: {total}\n\n生成图像尺寸统计:\n"
                f"\n完成时间: {datetime.now().strftime('%Y-%m-%d %H:%M:%S')}\n")
    print(f"数据集 {dataset_name} {shot_number}-shot处理完成：成功 {ok} 个样本，失败 {bad} 个样本，总共生成 {total} 张图像")
    return base


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="批量生成Flux k-shot图像 (B200-native)")
    p.add_argument("--dataset", type=str, default=None)
    p.add_argument("--shots", nargs="+", type=int, default=None)
    p.add_argument("--output_dir", type=str, default="result")
    p.add_argument("--database", type=str, default="coco", choices=["coco", "miniimagenet"])
    p.add_argument("--retrieval_results_dir", type=str, default="./retrieval/retrieval_results")
    p.add_argument("--dataset_group", type=str, default=None, choices=["dataset1", "dataset2", "dataset3", "dataset4"])
    # additions
    p.add_argument("--lamainpaint_dir", type=str, default=LAMAINPAINT_DIR)
    p.add_argument("--weights_dir", type=str, default="./model")
    p.add_argument("--model_size", type=str, default="full", choices=["full", "tiny"])
    p.add_argument("--num_inference_steps", type=int, default=50)
    p.add_argument("--image_size", type=int, default=1024)
    p.add_argument("--gen_batch", type=int, default=4,
                   help="generations of one sample (its top-ranked reference images) per FluxPipeline call; 1 = the "
                        "reference's one-at-a-time loop (every generation keeps seed 0 either way)")
    p.add_argument("--allow_random_init", action="store_true",
                   help="dry run without checkpoints: seeded random-init models and synthetic text tokens (noise images); "
                        "without it a missing weight file is an error")
    p.add_argument("--rank", type=int, default=None, help="data-parallel rank (default: RANK env or 0)")
    p.add_argument("--world_size", type=int, default=None, help="data-parallel world size (default: WORLD_SIZE env or 1)")
    return p


def main(argv=None) -> int:
    from . import hostlogic as H
    from .models import load_model
    args = build_parser().parse_args(argv)
    if args.dataset:
        datasets = [args.dataset]
    elif args.dataset_group:
        datasets = DATASET_GROUPS.get(args.dataset_group, [])
    else:
        datasets = ALL_DATASETS
    shots = args.shots or [1, 5, 10]
    rank = args.rank if args.rank is not None else int(os.environ.get("RANK", "0"))
    world = args.world_size if args.world_size is not None else int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    try:
        pipes = load_model(device=f"cuda:{local}", want=("dev",), weights_dir=args.weights_dir, size=args.model_size,
                           max_side=args.image_size, max_batch=max(1, args.gen_batch),
                           allow_random_init=args.allow_random_init)
    except FileNotFoundError as e:
        print(f"错误：{e}")
        return 2
    for ds in datasets:
        for k in shots:
            results = load_retrieval_results(args.retrieval_results_dir, ds, k)
            if not results:
                print(f"跳过数据集 {ds} {k}-shot，因为找不到检索结果")
                continue
            shot_dir = os.path.join(args.lamainpaint_dir, ds, f"{k}_shot")
            names = sorted(os.path.splitext(f)[0] for f in os.listdir(shot_dir) if f.endswith(".jpg")) if os.path.isdir(shot_dir) else []
            mine = set(H.split_samples_for_gpus(names, world)[rank]) if world > 1 else None
            process_kshot_dataset_with_retrieval(ds, pipes.prior_redux, pipes.pipe, results, k, args.output_dir,
                                                 args.lamainpaint_dir, args.database, args.num_inference_steps,
                                                 args.image_size, sample_filter=mine, gen_batch=max(1, args.gen_batch))
    return 0
