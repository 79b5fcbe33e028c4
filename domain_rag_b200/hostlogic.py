"""Host-side logic of the composition path (integer / PIL work, no tensors): per-dataset sampler
tables, sample partitioning over GPUs, resolution rules and the outpaint mask.

Mirrors outpainting_updown_sampling_redux.py of the reference:
  strength / guidance / image-prompt-scale / upscale / redux-prompt tables   :31-95
  split_samples_for_gpus                                                       :157-177
  process_image_resolution, downscale_image, upscale_image                     :403-498
  generate_outpaint_mask (returns (mask, boxes) like the reference)            :836-870
Pinned by tests/golden/host_helpers.json (outputs of the reference's own functions).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence, Tuple

from PIL import Image, ImageDraw

MIN_DIMENSION = 1024
MAX_DIMENSION = 2800
RESAMPLE = Image.BICUBIC          # UPSCALE_METHOD == DOWNSCALE_METHOD == Image.BICUBIC in the reference

FISH_PROMPT = ("wihout fish, A crystal-clear underwater environment, crisp and in sharp focus, foreground clarity "
               "is high; natural lighting and color continuity.")   # sic (reference :86)


@dataclass(frozen=True)
class DatasetParams:
    strength: float = 0.75            # default_strength
    guidance_scale: float = 30.0      # default_guidance_scale
    image_prompt_scale: float = 1.0
    upscale_dimension: int = 1024
    redux_prompt: str = ""


_TABLE = {
    "FISH": DatasetParams(0.8, 35.0, 1.2, 1024, FISH_PROMPT),
    "DIOR": DatasetParams(0.8, 30.0, 1.0, 1024, ""),
    "ArTaxOr": DatasetParams(0.9, 30.0, 1.0, 1024, ""),
    "UODD": DatasetParams(0.4, 30.0, 1.0, 2048, ""),
    "NEU-DET": DatasetParams(0.3, 30.0, 1.0, 1024, ""),
    "clipart1k": DatasetParams(0.9, 40.0, 1.0, 1024, ""),
    "NWPU_VHR-10": DatasetParams(0.8, 30.0, 1.0, 1024, ""),
    "Camouflage": DatasetParams(0.6, 30.0, 1.0, 1024, ""),
    "coco": DatasetParams(0.8, 30.0, 1.0, 1024, ""),
}


def dataset_params(dataset_name: str) -> DatasetParams:
    """Sampler parameters of a dataset (the reference looks each table up with a default)."""
    return _TABLE.get(dataset_name, DatasetParams())


def executed_steps(num_inference_steps: int, strength: float) -> int:
    """Steps an img2img-style Flux pipeline actually runs: T - t_start with
    t_start = int(max(T - min(T * s, T), 0)) (diffusers get_timesteps; T = 50, s = 0.75 -> 38 steps)."""
    return num_inference_steps - int(max(num_inference_steps - min(num_inference_steps * strength, num_inference_steps), 0))


def split_samples_for_gpus(sample_list: Sequence, num_gpus: int) -> List[list]:
    """Contiguous balanced partition: the first len % num_gpus parts get one extra element."""
    items = list(sample_list)
    if num_gpus <= 1:
        return [items]
    base, extra = divmod(len(items), num_gpus)
    parts, lo = [], 0
    for g in range(num_gpus):
        hi = lo + base + (1 if g < extra else 0)
        parts.append(items[lo:hi])
        lo = hi
    return parts


def process_image_resolution(image: Image.Image, min_dimension: int = MIN_DIMENSION,
                             max_dimension: int = MAX_DIMENSION):
    """-> (image, up_factor, down_factor, upsampled?, downsampled?). Sizes are truncated with int() like the
    reference; an image needing both directions raises ValueError."""
    w, h = image.size
    if min(w, h) < min_dimension and max(w, h) > max_dimension:
        raise ValueError(f"图像既需要上采样又需要下采样：尺寸 {w}x{h}，最小维度 {min(w, h)}，最大维度 {max(w, h)}")
    if min(w, h) < min_dimension:
        f = max(min_dimension / w if w < min_dimension else 1.0, min_dimension / h if h < min_dimension else 1.0)
        return image.resize((int(w * f), int(h * f)), RESAMPLE), f, 1.0, True, False
    if max(w, h) > max_dimension:
        f = max_dimension / max(w, h)
        return image.resize((int(w * f), int(h * f)), RESAMPLE), 1.0, f, False, True
    return image, 1.0, 1.0, False, False


def downscale_image(image: Image.Image, scale_factor: float) -> Image.Image:
    if scale_factor <= 1.0:
        return image
    w, h = image.size
    return image.resize((int(w / scale_factor), int(h / scale_factor)), RESAMPLE)


def upscale_image(image: Image.Image, scale_factor: float) -> Image.Image:
    if scale_factor <= 1.0:
        return image
    w, h = image.size
    return image.resize((int(w * scale_factor), int(h * scale_factor)), RESAMPLE)


def scale_bbox(bbox: Sequence[float], factor: float) -> Tuple[int, int, int, int]:
    """xywh box to the resampled image: every coordinate truncated with int() (reference :1172-1175)."""
    x, y, w, h = bbox
    return int(x * factor), int(y * factor), int(w * factor), int(h * factor)


def generate_outpaint_mask(original_image: Image.Image, bbox_coords_list):
    """L-mode mask: 255 = repaint, 0 = keep (every xywh box, clamped; PIL rectangles include both corners)."""
    W, H = original_image.size
    mask = Image.new("L", (W, H), 255)
    pen = ImageDraw.Draw(mask)
    for (x, y, w, h) in bbox_coords_list:
        x0, y0 = max(0, min(x, W - 1)), max(0, min(y, H - 1))
        x1, y1 = max(0, min(x + w, W)), max(0, min(y + h, H))
        pen.rectangle([x0, y0, x1, y1], fill=0)
    return mask, bbox_coords_list
