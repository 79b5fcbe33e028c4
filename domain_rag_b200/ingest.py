"""GPU-side ingest helpers: pinned host batches cross PCIe on a copy stream while the previous batch is
being encoded, so the H2D time of the reference's per-image `.to(device)` (retrieval/clip100_resnet_style_all_shots.py:
168-171, 284) disappears behind the tensor-core work. Data movement only - no arithmetic."""
from __future__ import annotations

from typing import Iterator

import torch

_COPY_STREAMS: dict = {}


def copy_stream(device) -> "torch.cuda.Stream":
    dev = torch.device(device)
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    if key not in _COPY_STREAMS:
        _COPY_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _COPY_STREAMS[key]


def device_batches(src: torch.Tensor, batch: int, device) -> Iterator[torch.Tensor]:
    """Yield `src[i:i+batch]` on `device`. Device-resident sources are sliced in place; host sources (pin them) are copied
    one batch AHEAD on a side stream: batch i+1 is in flight while the caller's kernels for batch i run on the current
    stream. Each yielded tensor is safe to use on the current stream (event wait + record_stream)."""
    n = src.shape[0]
    if src.is_cuda:
        for i in range(0, n, batch):
            yield src[i:i + batch]
        return
    dev = torch.device(device)
    side, main = copy_stream(dev), torch.cuda.current_stream(dev)
    pending = None
    for i in range(0, n, batch):
        with torch.cuda.stream(side):
            t = src[i:i + batch].to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(side)
        if pending is not None:
            pt, pev = pending
            main.wait_event(pev)
            pt.record_stream(main)
            yield pt
        pending = (t, ev)
    if pending is not None:
        pt, pev = pending
        main.wait_event(pev)
        pt.record_stream(main)
        yield pt
