"""CPU: product host logic (domain_rag_b200/hostlogic.py) against the golden vectors produced by the
reference's own functions (oracle/make_golden.py) - reference outpainting_updown_sampling_redux.py:31-95,157-177,403-498,836-870."""
import json

import numpy as np
import pytest
from PIL import Image

from domain_rag_b200 import hostlogic as H


def test_against_reference_golden(golden_dir):
    g = json.load(open(golden_dir / "host_helpers.json"))
    arrs = np.load(golden_dir / "host_helpers_arrays.npz")
    for c in g["split"]:
        assert H.split_samples_for_gpus([f"s{i}" for i in range(c["n"])], c["gpus"]) == c["out"]
    for c in g["resolution"]:
        im = Image.new("RGB", tuple(c["size"]))
        if c.get("error"):
            with pytest.raises(ValueError):
                H.process_image_resolution(im)
            continue
        out, up, down, nu, nd = H.process_image_resolution(im)
        assert list(out.size) == c["out_size"] and (nu, nd) == (c["need_up"], c["need_down"])
        assert up == pytest.approx(c["up"]) and down == pytest.approx(c["down"])
    for c in g["mask"]:
        m, boxes = H.generate_outpaint_mask(Image.new("RGB", tuple(c["size"])), [tuple(b) for b in c["boxes"]])
        np.testing.assert_array_equal(np.array(m), arrs[c["key"]])
        assert boxes == [tuple(b) for b in c["boxes"]]
    src = Image.fromarray(arrs["resize_src"])
    up = H.upscale_image(src, 1.7)
    np.testing.assert_array_equal(np.array(up), arrs["upscale_1p7"])
    np.testing.assert_array_equal(np.array(H.downscale_image(up, 1.7)), arrs["downscale_1p7"])
    assert H.upscale_image(src, 1.0) is src and H.downscale_image(src, 0.5) is src


def test_tables_and_steps():
    assert H.dataset_params("DIOR") == H.DatasetParams(0.8, 30.0, 1.0, 1024, "")
    assert H.dataset_params("UODD").upscale_dimension == 2048 and H.dataset_params("NEU-DET").strength == 0.3
    assert H.dataset_params("FISH").redux_prompt.startswith("wihout fish") and H.dataset_params("FISH").image_prompt_scale == 1.2
    assert H.dataset_params("unknown") == H.DatasetParams(0.75, 30.0, 1.0, 1024, "")
    assert [H.executed_steps(50, s) for s in (0.3, 0.4, 0.8, 0.9, 1.0, 1.5)] == [15, 20, 40, 45, 50, 50]
    # fractional T * s: diffusers truncates the START index (int(max(T - min(T*s, T), 0))), so the step count rounds UP;
    # the reference's default strength 0.75 (every dataset outside the table) runs 38 of 50 steps, not 37
    assert [H.executed_steps(50, s) for s in (0.75, 0.35, 0.6, 0.01)] == [38, 18, 30, 1]
    assert [H.executed_steps(28, s) for s in (0.75, 0.5, 0.33)] == [21, 14, 10]
    from domain_rag_b200.flux import executed_range
    from oracle.pipelines import executed_start
    for T in (4, 20, 28, 50):
        for s in (0.0, 0.01, 0.3, 0.33, 0.35, 0.5, 0.6, 0.75, 0.8, 0.9, 0.99, 1.0, 1.5):
            assert executed_range(T, s) == executed_start(T, s) == T - H.executed_steps(T, s)
            assert executed_range(T, s) == int(max(T - min(T * s, T), 0))
    assert H.scale_bbox((10.6, 3.2, 7.9, 5.5), 1.7) == (18, 5, 13, 9)
