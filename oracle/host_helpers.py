"""Oracle: host-side helpers of the composition script. TEST INFRASTRUCTURE ONLY.

Follows outpainting_updown_sampling_redux.py: split_samples_for_gpus :157-177,
process_image_resolution :403-458, downscale_image/upscale_image :460-498,
generate_outpaint_mask :836-870. PINNED by tests/golden/host_helpers.json (outputs of the
reference's own functions, oracle/make_golden.py).
"""
from __future__ import annotations

from PIL import Image, ImageDraw

MIN_DIMENSION = 1024
MAX_DIMENSION = 2800


def split_samples_for_gpus(sample_list, num_gpus):
    if num_gpus <= 1:
        return [sample_list]
    per, rem = divmod(len(sample_list), num_gpus)
    out, s = [], 0
    for g in range(num_gpus):
        e = s + per + (1 if g < rem else 0)
        out.append(sample_list[s:e])
        s = e
    return out


def process_image_resolution(image, min_dimension=MIN_DIMENSION, max_dimension=MAX_DIMENSION):
    w, h = image.size
    mx, mn = max(w, h), min(w, h)
    if mn < min_dimension and mx > max_dimension:
        raise ValueError(f"image needs both up- and down-sampling: {w}x{h}")
    if mn < min_dimension:
        sw = min_dimension / w if w < min_dimension else 1.0
        sh = min_dimension / h if h < min_dimension else 1.0
        f = max(sw, sh)
        return image.resize((int(w * f), int(h * f)), Image.BICUBIC), f, 1.0, True, False
    if mx > max_dimension:
        f = max_dimension / mx
        return image.resize((int(w * f), int(h * f)), Image.BICUBIC), 1.0, f, False, True
    return image, 1.0, 1.0, False, False


def generate_outpaint_mask(size, bbox_coords_list):
    w, h = size
    mask = Image.new("L", (w, h), 255)
    draw = ImageDraw.Draw(mask)
    for (x, y, bw, bh) in bbox_coords_list:
        x2, y2 = x + bw, y + bh
        x = max(0, min(x, w - 1))
        y = max(0, min(y, h - 1))
        x2 = max(0, min(x2, w))
        y2 = max(0, min(y2, h))
        draw.rectangle([x, y, x2, y2], fill=0)
    return mask
