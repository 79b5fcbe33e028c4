"""GPU: the three drop-in entry points chained on a synthetic dataset tree with tiny models:

    retrieval/clip100_resnet_style_all_shots.py  ->  batch_generate_flux_kshot.py  ->  outpainting_updown_sampling_redux.py

Checks the full file trees against the committed golden listing (tests/golden/cli_tree.json; timestamps and similarity
digits normalised), the JSON schemas of the reference, that image pixels equal a direct pipeline call with the same inputs,
`--resume`, and `--multi_gpu --num_gpus 2` when two GPUs are visible."""
import json
import os
import re

import numpy as np
import pytest
import torch
from PIL import Image

pytestmark = pytest.mark.gpu

STEPS, SIZE = 3, 256


def rand_img(path, w, h, seed=0, smooth=False):
    g = np.random.default_rng(seed)
    a = g.random((h, w, 3))
    if smooth:
        yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
        a = np.stack([np.sin(xx / (9 + c)) * np.cos(yy / (11 - c)) for c in range(3)], -1) * 0.4 + 0.5 + 0.05 * a
    Image.fromarray((a.clip(0, 1) * 255).astype(np.uint8)).save(path)


def make_tree(root):
    (root / "coco" / "train2017").mkdir(parents=True)
    for i in range(12):
        rand_img(root / "coco" / "train2017" / f"im{i}.jpg", 60 + i, 50, i)
    shot = root / "lamainpaint" / "DIOR" / "5_shot"
    shot.mkdir(parents=True)
    ds = root / "datasets" / "DIOR"
    (ds / "annotations").mkdir(parents=True)
    (ds / "train").mkdir()
    images, anns = [], []
    for k, (name, w, h) in enumerate([("a1", 320, 240), ("b2", 300, 300)]):
        rand_img(shot / f"{name}.jpg", w, h, 100 + k, smooth=True)
        rand_img(ds / "train" / f"{name}.jpg", w, h, 200 + k, smooth=True)
        images.append({"id": k + 1, "file_name": f"{name}.jpg"})
        anns.append({"image_id": k + 1, "category_id": 1, "bbox": [40 + 10 * k, 30, 90, 80]})
    json.dump({"images": images, "categories": [{"id": 1, "name": "airport"}], "annotations": anns},
              open(ds / "annotations" / "5_shot.json", "w"))


def listing(root):
    """Relative file paths under root, volatile parts normalised."""
    out = []
    for d, _, files in os.walk(root):
        for f in files:
            p = os.path.relpath(os.path.join(d, f), root)
            p = re.sub(r"_targettext_1\.0_\d{8}_\d{6}", "_targettext_1.0_<TS>", p)
            p = re.sub(r"_sim\d+\.\d{4}", "_sim<S>", p)
            out.append(p)
    return sorted(out)


@pytest.fixture(scope="module")
def chain(lib, tmp_path_factory):
    """Runs retrieval -> generate once for the module; compose variants run in the tests."""
    from domain_rag_b200 import generate_cli as GC
    from domain_rag_b200 import retrieval_cli as RC
    root = tmp_path_factory.mktemp("n4")
    make_tree(root)
    cwd = os.getcwd()
    os.chdir(root)
    try:
        assert RC.main(["--datasets", "DIOR", "--shots", "5", "--coco-dir", "coco", "--lamainpaint-dir", "lamainpaint",
                        "--output-dir", "retrieval/retrieval_results", "--pretrained-coco-features", "none.pt",
                        "--clip-top-k", "3", "--allow-random-init", "--no-visual"]) == 0
        assert GC.main(["--dataset", "DIOR", "--shots", "5", "--output_dir", "result", "--lamainpaint_dir", "lamainpaint",
                        "--model_size", "tiny", "--allow_random_init", "--num_inference_steps", str(STEPS),
                        "--image_size", str(SIZE), "--gen_batch", "2"]) == 0
    finally:
        os.chdir(cwd)
    return root


def test_generate_tree_and_pixels(chain, golden_dir):
    from domain_rag_b200 import generate_cli as GC
    from domain_rag_b200.models import load_model
    gold = json.load(open(golden_dir / "cli_tree.json"))
    assert listing(chain / "result") == gold["generate"]
    res_dir = next((chain / "result" / "DIOR_5shot_retrieval").iterdir())
    bp = open(res_dir / "batch_params.txt").read()
    assert "成功处理样本数: 2" in bp and "总共生成图像数: 6" in bp and "生成图像尺寸统计:\n\n完成时间" in bp
    assert f"推理步数: {STEPS}" in open(res_dir / "a1" / "params.txt").read()
    # pixels of a CLI image == the direct pipeline calls the reference makes (:459-474) on the same files, batch 1
    rec = json.load(open(chain / "retrieval" / "retrieval_results" / "DIOR_5_shot_retrieval_results.json"))["a1"][0]
    ref = [s for s in rec["similar_images"] if s["rank"] == 2][0]["image_path"]
    pipes = load_model(device="cuda", want=("dev",), weights_dir="./model", size="tiny", max_side=SIZE, max_batch=2,
                       allow_random_init=True)
    prior = GC._redux_prior(pipes.prior_redux, Image.open(os.path.join(chain, ref) if not os.path.isabs(ref) else ref).convert("RGB"),
                            Image.open(chain / "lamainpaint" / "DIOR" / "5_shot" / "a1.jpg").convert("RGB"))
    img = pipes.pipe(guidance_scale=2.5, num_inference_steps=STEPS, height=SIZE, width=SIZE,
                     generator=torch.Generator("cpu").manual_seed(0), **prior).images[0]
    got = np.asarray(Image.open(res_dir / "a1" / "generated_image_rank2.png")).astype(np.int32)
    d = np.abs(got - np.asarray(img).astype(np.int32))
    assert got.shape == (SIZE, SIZE, 3) and d.max() <= 2, d.max()       # batch-of-2 vs batch-1 run: row independent kernels


def test_compose_tree_schema_pixels_resume(chain, golden_dir, monkeypatch, capsys):
    from domain_rag_b200 import compose_cli as CC
    from domain_rag_b200 import hostlogic as H
    from domain_rag_b200.models import load_model
    monkeypatch.chdir(chain)
    common = ["--dataset", "DIOR", "--shot", "5", "--model_size", "tiny", "--allow_random_init", "--num_inference_steps",
              str(STEPS), "--custom_upscale", "DIOR:256", "--compose_batch", "2", "--seed", "5"]
    assert CC.main([*common, "--process_id", "PID", "--sample_id", "a1"]) == 0
    log_text = capsys.readouterr().out
    assert "样本 a1 处理完成" in log_text
    gold = json.load(open(golden_dir / "cli_tree.json"))
    assert listing(chain / "outpaint_hires" / "process_PID") == gold["compose_a1"]
    assert listing(chain / "final_results" / "process_PID") == gold["final_a1"]
    res = json.load(open(chain / "outpaint_hires" / "process_PID" / "DIOR" / "5_shot" / "outpaint_results_5shot.json"))
    assert set(res) >= {"dataset", "timestamp", "process_id", "shot_number", "total_samples", "successful_samples",
                        "failed_samples", "samples"}
    assert (res["total_samples"], res["successful_samples"], res["shot_number"], res["process_id"]) == (1, 1, 5, "PID")
    s = res["samples"][0]
    assert s["status"] == "completed" and len(s["outpainted_images"]) == 3 and s["was_upscaled"] is True
    out = chain / "outpaint_hires" / "process_PID" / "DIOR" / "5_shot" / "a1"
    prm = json.load(open(out / "DIOR_a1_5shot_params_2.json"))
    assert set(prm) >= {"categories", "image_prompt_scale", "guidance_scale", "num_inference_steps", "strength", "redux_prompt",
                        "seed", "process_id", "shot_number", "bg_index", "bg_filename", "original_resolution",
                        "processed_resolution", "up_scale_factor", "bbox_coords_list", "processed_bbox_coords_list", "num_bbox"}
    assert (prm["strength"], prm["guidance_scale"], prm["seed"], prm["categories"]) == (0.8, 30.0, 5, ["airport"])
    f = 256 / 240
    assert prm["processed_resolution"] == {"width": int(320 * f), "height": 256}
    assert prm["processed_bbox_coords_list"] == [[int(c * f) for c in [40, 30, 90, 80]]]
    # pixels == direct FluxFillPipeline call (the reference's :1237-1257) on the same processed image / mask / background
    pipes = load_model(device="cuda", want=("fill",), weights_dir="./model", size="tiny", max_side=H.MAX_DIMENSION,
                       max_batch=2, allow_random_init=True)
    processed = Image.open(out / "DIOR_a1_5shot_upscaled_bg.png").convert("RGB")
    mask = Image.open(out / "DIOR_a1_5shot_mask_2.png")
    bg = Image.open(out / "DIOR_a1_5shot_bg_2_original.png").convert("RGB")
    prior = pipes.prior_redux([bg], prompt="", prompt_2="", prompt_embeds_scale=[1.0], pooled_prompt_embeds_scale=[1.0])
    img = pipes.pipe_fill(image=processed, mask_image=mask, height=processed.height, width=processed.width,
                          guidance_scale=30.0, num_inference_steps=STEPS, generator=torch.Generator("cpu").manual_seed(5),
                          strength=0.8, **prior).images[0]
    hires = Image.open(out / "DIOR_a1_5shot_hires_result_2.png")
    d = np.abs(np.asarray(hires).astype(np.int32) - np.asarray(img).astype(np.int32))
    assert hires.size == (16 * (processed.width // 16), 16 * (processed.height // 16)) and d.max() <= 2, d.max()
    final = Image.open(out / "DIOR_a1_5shot_final_result_2.png")
    assert final.size == (int(hires.width / f), int(hires.height / f))
    np.testing.assert_array_equal(np.asarray(final), np.asarray(H.downscale_image(hires, f)))
    # --resume: the finished sample is skipped, the other one runs
    (chain / "run.log").write_text(log_text)
    assert CC.main([*common, "--process_id", "PID2", "--resume", "--log_file", "run.log"]) == 0
    res2 = json.load(open(chain / "outpaint_hires" / "process_PID2" / "DIOR" / "5_shot" / "outpaint_results_5shot.json"))
    assert [x["sample_id"] for x in res2["samples"]] == ["b2"] and res2["successful_samples"] == 1
    # a weights directory without files and no --allow_random_init: refuse loudly
    assert CC.main(["--dataset", "DIOR", "--shot", "5", "--model_size", "tiny", "--process_id", "PID3"]) == 2


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_compose_multi_gpu_two_workers(chain, monkeypatch):
    from domain_rag_b200 import compose_cli as CC
    monkeypatch.chdir(chain)
    assert CC.main(["--dataset", "DIOR", "--shot", "5", "--model_size", "tiny", "--allow_random_init", "--num_inference_steps",
                    str(STEPS), "--custom_upscale", "DIOR:256", "--seed", "5", "--process_id", "PIDM", "--multi_gpu",
                    "--num_gpus", "2"]) == 0
    res = json.load(open(chain / "outpaint_hires" / "process_PIDM" / "DIOR" / "5_shot" / "outpaint_results_5shot.json"))
    assert res["multi_gpu"] is True and res["num_gpus"] == 2 and res["gpu_process_ids"] == ["PIDM_gpu0", "PIDM_gpu1"]
    assert sorted(x["sample_id"] for x in res["samples"]) == ["a1", "b2"] and res["successful_samples"] == 2
    # --seed crossed the spawn boundary: same seed => the multi-GPU a1 composition equals the single-GPU one
    a = np.asarray(Image.open(chain / "outpaint_hires" / "process_PIDM" / "DIOR" / "5_shot" / "a1" / "DIOR_a1_5shot_final_result_1.png"))
    b = np.asarray(Image.open(chain / "outpaint_hires" / "process_PID" / "DIOR" / "5_shot" / "a1" / "DIOR_a1_5shot_final_result_1.png"))
    assert np.abs(a.astype(np.int32) - b.astype(np.int32)).max() <= 2
