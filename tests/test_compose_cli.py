"""CPU: file surface and host logic of the generation / composition entry points with stub
pipelines in place of the GPU ones - flags equal the reference's, output trees and JSON records have the reference's
names, bbox scaling / mask / resolution restore follow the reference's integer rules."""
import json
import os

import numpy as np
import pytest
from PIL import Image

from domain_rag_b200 import compose_cli as CC
from domain_rag_b200 import generate_cli as GC
from domain_rag_b200 import hostlogic as H


class _Out(dict):
    __getattr__ = dict.__getitem__


class StubPrior:
    def __init__(self):
        self.calls = []

    def __call__(self, image, prompt=None, prompt_2=None, prompt_embeds_scale=1.0, pooled_prompt_embeds_scale=1.0):
        self.calls.append(dict(n=len(image), prompt=prompt, prompt_2=prompt_2, se=prompt_embeds_scale, sp=pooled_prompt_embeds_scale))
        import torch
        return _Out(prompt_embeds=torch.full((1, 3, 4), float(len(self.calls))), pooled_prompt_embeds=torch.zeros(1, 4))


class StubPipe:
    def __init__(self):
        self.calls = []

    def __call__(self, **kw):
        self.calls.append(kw)
        w, h = kw.get("width", 64), kw.get("height", 64)
        n = len(kw["image"]) if isinstance(kw.get("image"), list) else 1
        if "image" not in kw and isinstance(kw.get("generator"), list):
            n = len(kw["generator"])
        return _Out(images=[Image.new("RGB", (16 * (w // 16), 16 * (h // 16)), (10, 200, 30)) for _ in range(n)])


class Pipes:
    def __init__(self):
        self.prior_redux, self.pipe, self.pipe_fill = StubPrior(), StubPipe(), StubPipe()


def rand_img(path, w, h, seed=0):
    Image.fromarray((np.random.default_rng(seed).random((h, w, 3)) * 255).astype(np.uint8)).save(path)


def test_generate_cli_file_surface(tmp_path):
    ns = vars(GC.build_parser().parse_args([]))
    assert (ns["dataset"], ns["shots"], ns["output_dir"], ns["database"], ns["dataset_group"]) == (None, None, "result", "coco", None)
    shot = tmp_path / "lamainpaint" / "DIOR" / "5_shot"
    shot.mkdir(parents=True)
    refs = tmp_path / "coco"
    refs.mkdir()
    rand_img(shot / "airport_1.jpg", 100, 70)
    rand_img(shot / "bridge_2.jpg", 90, 90, 1)
    sims = []
    for r in range(1, 8):
        rand_img(refs / f"r{r}.jpg", 40, 40, r)
        sims.append({"rank": r, "similarity": 1.0 / (1 + r), "image_path": str(refs / f"r{r}.jpg"), "source_dataset": "coco"})
    results = {"airport_1": [{"sample_id": "airport_1", "image_path": str(shot / "airport_1.jpg"), "category": "airport_1",
                              "similar_images": sims}]}
    pipes = Pipes()
    base = GC.process_kshot_dataset_with_retrieval("DIOR", pipes.prior_redux, pipes.pipe, results, 5, str(tmp_path / "result"),
                                                   str(tmp_path / "lamainpaint"))
    assert os.path.basename(os.path.dirname(base)) == "DIOR_5shot_retrieval"
    assert os.path.basename(base).startswith("results_coco_0.8_target_1.0_cocotext_1.0_targettext_1.0_")
    files = set(os.listdir(os.path.join(base, "airport_1")))
    assert {f"generated_image_rank{r}.png" for r in range(1, 6)} <= files and "generated_image_rank6.png" not in files
    assert {"params.txt", "target_input.png", "ref_inputrank1.jpg", "ref_inforank1_sim0.5000.txt"} <= files   # no separator (sic)
    assert os.listdir(os.path.join(base, "bridge_2")) == ["error.txt"]
    bp = open(os.path.join(base, "batch_params.txt")).read()
    assert "成功处理样本数: 1" in bp and "失败处理样本数: 1" in bp and "总共生成图像数: 5" in bp and "生成图像尺寸统计:\n\n完成时间" in bp
    c = pipes.prior_redux.calls[0]
    assert c == dict(n=2, prompt=["", ""], prompt_2=["", ""], se=[0.8, 1.0], sp=[1.0, 1.0])
    k = pipes.pipe.calls[0]
    assert (k["guidance_scale"], k["num_inference_steps"], k["height"], k["width"]) == (2.5, 50, 1024, 1024)
    assert k["generator"].initial_seed() == 0 and tuple(k["prompt_embeds"].shape) == (1, 3, 4)
    assert "生成图像尺寸: 96x64" in open(os.path.join(base, "airport_1", "params.txt")).read()
    assert len(pipes.pipe.calls) == 5
    # the same job with the ranks of a sample generated in batches of 4: same files, 2 pipeline calls, seed 0 per image
    pipes_b = Pipes()
    base_b = GC.process_kshot_dataset_with_retrieval("DIOR", pipes_b.prior_redux, pipes_b.pipe, results, 5,
                                                     str(tmp_path / "result_b"), str(tmp_path / "lamainpaint"), gen_batch=4)
    files_b = set(os.listdir(os.path.join(base_b, "airport_1")))
    assert files_b == files
    assert [len(c["generator"]) if isinstance(c["generator"], list) else 1 for c in pipes_b.pipe.calls] == [4, 1]
    kb = pipes_b.pipe.calls[0]
    assert all(g.initial_seed() == 0 for g in kb["generator"]) and tuple(kb["prompt_embeds"].shape) == (4, 3, 4)
    assert len(pipes_b.prior_redux.calls) == 5 and GC.build_parser().parse_args([]).gen_batch == 4


def test_compose_cli_file_surface(tmp_path):
    ns = vars(CC.build_parser().parse_args([]))
    for k, v in dict(dataset=None, dataset_group="all", sample_id=None, shot=1, min_dimension=1024, custom_upscale=None,
                     process_id=None, collect_only=False, multi_bbox=False, resume=False, log_file=None, failed_only=False,
                     multi_gpu=False, num_gpus=None).items():
        assert ns[k] == v, k
    ds_dir = tmp_path / "datasets" / "DIOR"
    (ds_dir / "annotations").mkdir(parents=True)
    (ds_dir / "train").mkdir()
    rand_img(ds_dir / "train" / "airport_1.jpg", 640, 480)
    json.dump({"images": [{"id": 7, "file_name": "airport_1.jpg"}], "categories": [{"id": 3, "name": "airport"}],
               "annotations": [{"image_id": 7, "category_id": 3, "bbox": [100.5, 50, 200, 120]},
                               {"image_id": "7", "category_id": 3, "bbox": [10, 400, 50, 60]}]},
              open(ds_dir / "annotations" / "5_shot.json", "w"))
    sdir = tmp_path / "result" / "DIOR_5shot_retrieval" / "results_x" / "airport_1"
    sdir.mkdir(parents=True)
    for r in (1, 2):
        rand_img(sdir / f"generated_image_rank{r}.png", 64, 64, r)
    pipes = Pipes()
    log = CC.process_sample_hires("DIOR", "airport_1", pipes, "PID", shot_number=5, datasets_dir=str(tmp_path / "datasets"),
                                  result_dir=str(tmp_path / "result"), outpaint_base=str(tmp_path / "outpaint_hires"),
                                  seed_fn=lambda: 123)
    assert log["status"] == "completed", log["error"]
    out = tmp_path / "outpaint_hires" / "process_PID" / "DIOR" / "5_shot" / "airport_1"
    p = "DIOR_airport_1_5shot"
    want = {f"{p}_original.png", f"{p}_bbox1_original.jpg", f"{p}_bbox2_original.jpg", f"{p}_upscaled_bg.png"}
    for r in (1, 2):
        want |= {f"{p}_mask_{r}.png", f"{p}_bg_{r}_original.png", f"{p}_hires_result_{r}.png", f"{p}_final_result_{r}.png",
                 f"{p}_params_{r}.json"}
    assert want <= set(os.listdir(out))
    # 640x480 -> upsampled by 1024/480 to 1365x1024 (int truncation), bbox scaled with int(); final restored to int(size / f)
    f = 1024 / 480
    assert log["up_scale_factor"] == pytest.approx(f) and tuple(log["upscaled_resolution"]) == (int(640 * f), 1024)
    prm = json.load(open(out / f"{p}_params_1.json"))
    assert prm["processed_bbox_coords_list"] == [[int(c * f) for c in [100.5, 50, 200, 120]], [int(c * f) for c in [10, 400, 50, 60]]]
    assert (prm["strength"], prm["guidance_scale"], prm["seed"], prm["num_bbox"], prm["categories"]) == (0.8, 30.0, 123, 2, ["airport", "airport"])
    k = pipes.pipe_fill.calls[0]
    assert (k["width"], k["height"], k["strength"], k["guidance_scale"], k["num_inference_steps"]) == (1365, 1024, 0.8, 30.0, 50)
    assert k["generator"].initial_seed() == 123
    mask = np.array(Image.open(out / f"{p}_mask_1.png"))
    want_mask, _ = H.generate_outpaint_mask(Image.new("RGB", (1365, 1024)), prm["processed_bbox_coords_list"])
    np.testing.assert_array_equal(mask, np.array(want_mask))
    hires = Image.open(out / f"{p}_hires_result_1.png")
    final = Image.open(out / f"{p}_final_result_1.png")
    assert hires.size == (1360, 1024) and final.size == (int(1360 / f), int(1024 / f))
    assert pipes.prior_redux.calls[0] == dict(n=1, prompt="", prompt_2="", se=[1.0], sp=[1.0])
    # result JSON, merge, collection, resume log
    res = CC.formatted_result_json("DIOR", [log], 5, "PID")
    path = CC.save_formatted_result_json("DIOR", res, 5, "PID", str(tmp_path / "outpaint_hires"))
    assert path.endswith("DIOR/5_shot/outpaint_results_5shot.json") and json.load(open(path))["successful_samples"] == 1
    merged = CC.merge_gpu_results("DIOR", [dict(res, gpu_process_id="a"), dict(res, gpu_process_id="b")], 5, "PID")
    assert merged["multi_gpu"] and merged["num_gpus"] == 2 and len(merged["samples"]) == 2 and merged["gpu_process_ids"] == ["a", "b"]
    coll = CC.copy_final_results_to_collection("PID", 5, str(tmp_path / "outpaint_hires"), str(tmp_path / "final_results"))
    assert sorted(os.listdir(os.path.join(coll, "DIOR", "5_shot"))) == [f"{p}_final_result_1.png", f"{p}_final_result_2.png"]
    logf = tmp_path / "run.log"
    logf.write_text("样本 a 处理完成，耗时 1.00 秒\n样本 b 处理失败，耗时 2.00 秒\n处理样本 c 时出错: boom\n样本 c 处理完成，耗时 3 秒\n")
    assert CC.parse_resume_log(str(logf)) == ({"a", "c"}, {"b"})
    # the same sample with its backgrounds composed as ONE batch: same files, one pipeline call, per-composition seeds
    pipes_b = Pipes()
    seeds = iter([11, 22])
    log_b = CC.process_sample_hires("DIOR", "airport_1", pipes_b, "PIDB", shot_number=5, datasets_dir=str(tmp_path / "datasets"),
                                    result_dir=str(tmp_path / "result"), outpaint_base=str(tmp_path / "outpaint_hires"),
                                    seed_fn=lambda: next(seeds), compose_batch=4)
    assert log_b["status"] == "completed", log_b["error"]
    out_b = tmp_path / "outpaint_hires" / "process_PIDB" / "DIOR" / "5_shot" / "airport_1"
    assert want <= set(os.listdir(out_b))
    assert len(pipes_b.pipe_fill.calls) == 1 and len(pipes_b.prior_redux.calls) == 2
    kb = pipes_b.pipe_fill.calls[0]
    assert len(kb["image"]) == 2 and len(kb["mask_image"]) == 2 and kb["prompt_embeds"].shape == (2, 3, 4)
    assert [g.initial_seed() for g in kb["generator"]] == [11, 22]
    assert [json.load(open(out_b / f"{p}_params_{r}.json"))["seed"] for r in (1, 2)] == [11, 22]
    assert CC.build_parser().parse_args([]).compose_batch == 4
    # a sample with neither annotation nor backgrounds is an error record, not an exception
    bad = CC.process_sample_hires("DIOR", "nope", pipes, "PID", shot_number=5, datasets_dir=str(tmp_path / "datasets"),
                                  result_dir=str(tmp_path / "result"), outpaint_base=str(tmp_path / "outpaint_hires"))
    assert bad["status"] == "error" and bad["error"]


# ------------------------------------------------------------------ round-2 host logic: loud failures, seeds, dead workers
def test_load_model_refuses_missing_weights(tmp_path):
    """A missing weight file is an error naming the files (no silent noise models); the CLIs return 2 before touching a GPU."""
    from domain_rag_b200 import models as M
    assert M.missing_files(str(tmp_path), ("fill",)) == ["siglip.pt", "redux.pt", "text_embeds.pt", "vae.pt", "flux_fill.pt"]
    (tmp_path / "vae.safetensors").write_bytes(b"")
    (tmp_path / "siglip.pt").write_bytes(b"")
    assert M.missing_files(str(tmp_path), ("dev", "fill")) == ["redux.pt", "text_embeds.pt", "flux_dev.pt", "flux_fill.pt"]
    with pytest.raises(FileNotFoundError, match="flux_fill.pt"):
        M.load_model(device="cpu", want=("fill",), weights_dir=str(tmp_path))
    assert CC.main(["--dataset", "DIOR", "--shot", "5", "--weights_dir", str(tmp_path)]) == 2
    assert GC.main(["--dataset", "DIOR", "--shots", "5", "--weights_dir", str(tmp_path)]) == 2


def test_text_table_is_strict_unless_synthetic_is_allowed():
    import torch
    from domain_rag_b200 import redux as R
    t = R.TextEmbeddingTable("cpu", txt_dim=8, pooled_dim=4, tokens=3)
    with pytest.raises(KeyError, match="make_text_embeds"):
        t.lookup("", "")
    t2 = R.TextEmbeddingTable("cpu", txt_dim=8, pooled_dim=4, tokens=3, allow_synthetic=True)
    a, b = t2.lookup("", ""), t2.lookup("", None)
    assert a[0].shape == (3, 8) and a[1].shape == (4,) and torch.equal(a[0], b[0])
    t3 = R.TextEmbeddingTable("cpu", txt_dim=8, pooled_dim=4, tokens=3, t5_encode=lambda s: torch.full((3, 8), float(len(s))),
                              clip_encode=lambda s: torch.full((4,), float(len(s))))
    e = t3.lookup("ab", "abcd")      # diffusers: prompt -> CLIP (pooled), prompt_2 -> T5
    assert float(e[0][0, 0]) == 4.0 and float(e[1][0]) == 2.0


def test_multi_gpu_seed_crosses_the_spawn_boundary(tmp_path, monkeypatch):
    """--seed reaches the per-GPU workers as a plain int (it used to be dropped with the lambda)."""
    seen = {}

    def fake_multi(ds, ids, shot, n_gpus, process_id, kwargs, load_kwargs):
        import pickle
        pickle.dumps(kwargs)                      # must survive mp spawn
        seen.update(kwargs=kwargs, load_kwargs=load_kwargs, n=n_gpus)
        return CC.formatted_result_json(ds, [], shot, process_id)

    monkeypatch.setattr(CC, "process_dataset_samples_multi_gpu", fake_multi)
    monkeypatch.chdir(tmp_path)
    assert CC.main(["--dataset", "DIOR", "--shot", "5", "--sample_id", "s1", "--multi_gpu", "--num_gpus", "2", "--seed", "77",
                    "--allow_random_init", "--model_size", "tiny"]) == 0
    assert seen["kwargs"]["seed"] == 77 and seen["n"] == 2 and seen["load_kwargs"]["allow_random_init"] is True


def test_dead_worker_does_not_block_the_parent():
    import multiprocessing as mp
    import time

    class P:
        def __init__(self, alive, code=None):
            self.alive, self.exitcode = alive, code

        def is_alive(self):
            return self.alive

    q = mp.Queue()
    q.put((0, {"samples": [{"status": "completed"}], "gpu_process_id": "p_gpu0"}))
    procs = {0: P(True), 1: P(False, -9)}
    t0 = time.time()
    res = CC.collect_worker_results(procs, q, poll_seconds=0.1)
    assert time.time() - t0 < 5 and res[0]["gpu_process_id"] == "p_gpu0" and res[1] is None
    dead = CC._dead_rank_result("DIOR", ["a", "b"], 5, "p", 1, -9)
    merged = CC.merge_gpu_results("DIOR", [res[0], dead], 5, "p")
    assert merged["total_samples"] == 3 and merged["failed_samples"] == 2 and merged["num_gpus"] == 2
    assert all("exit code -9" in s["error"] for s in merged["samples"] if s["status"] == "error")


def test_text_embeds_script_prompts_and_table_roundtrip(tmp_path):
    """scripts/make_text_embeds.py encodes exactly the prompt pairs the two scripts ever pass; its file layout is what
    TextEmbeddingTable.load_file reads; without encoder checkpoints it refuses (no synthetic fallback)."""
    import importlib.util
    import torch
    from domain_rag_b200 import redux as R
    spec = importlib.util.spec_from_file_location("mte", str(CC.__file__).replace("domain_rag_b200/compose_cli.py",
                                                                                 "scripts/make_text_embeds.py"))
    mte = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mte)
    pairs = mte.prompt_pairs(["x"])
    assert pairs == [("", ""), (H.FISH_PROMPT, ""), ("x", "x")]
    assert mte.main(["--flux_dir", str(tmp_path / "nope"), "--out", str(tmp_path / "t.pt")]) == 2
    torch.save({"prompts": [list(p) for p in pairs], "t5": torch.randn(3, 5, 8).bfloat16(),
                "pooled": torch.randn(3, 4).bfloat16()}, tmp_path / "t.pt")
    t = R.TextEmbeddingTable("cpu", txt_dim=8, pooled_dim=4, tokens=5)
    t.load_file(str(tmp_path / "t.pt"))
    assert t.lookup(H.FISH_PROMPT, "")[0].shape == (5, 8) and t.lookup("", None)[1].shape == (4,)
    with pytest.raises(KeyError):
        t.lookup("something else", "")
