"""Exact inner-product index - drop-in for the faiss calls of the reference's first-stage retrieval.

Reference call site (retrieval/clip100_resnet_style_all_shots.py:425-434):

    index = faiss.IndexFlatIP(d); index.add(features_np); D, I = index.search(query_np, k)

`IndexFlatIP` here has the same constructor / add / search / ntotal / reset surface and returns the
same (D float32 [nq,k], I int64 [nq,k]) tuple, scores descending. Differences, all supersets:
the corpus lives in HBM and persists across queries; ties are broken by lower id (faiss leaves it
unspecified); `add_device` / `search_device` take torch CUDA tensors without host round trips;
`ShardedIndexFlatIP` partitions rows over ranks and merges per-shard top-k after one all-gather.
Everything executes in libdomainrag_b200.so; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional

import numpy as np

from . import _lib


class IndexFlatIP:
    """faiss.IndexFlatIP mirror backed by the sm_100a scan x top-k kernel."""

    def __init__(self, d: int, device: int = 0):
        lib = _lib.load()
        self.d = int(d)
        self.device = int(device)
        self._h = C.c_void_p()
        _lib.check(lib.drag_index_create(self.d, self.device, C.byref(self._h)), "drag_index_create")
        self._keep = []  # adopted tensors must outlive the index

    # -- faiss surface -------------------------------------------------------------------------
    @property
    def ntotal(self) -> int:
        n = C.c_int64(0)
        _lib.check(_lib.load().drag_index_ntotal(self._h, C.byref(n)), "drag_index_ntotal")
        return n.value

    def add(self, x: np.ndarray) -> None:
        """index.add(features_np): copy host rows [n, d] float32 into HBM."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 2 or x.shape[1] != self.d:
            raise ValueError(f"add: expected [n, {self.d}] array, got {x.shape}")
        base = self.ntotal
        _lib.check(_lib.load().drag_index_add(self._h, _lib.ptr(x), x.shape[0], base, 0, None),
                   "drag_index_add")

    def search(self, q: np.ndarray, k: int):
        """D, I = index.search(query_np, k) with host buffers (synchronous)."""
        q = np.ascontiguousarray(q, dtype=np.float32)
        if q.ndim != 2 or q.shape[1] != self.d:
            raise ValueError(f"search: expected [nq, {self.d}] array, got {q.shape}")
        nq = q.shape[0]
        D = np.empty((nq, k), dtype=np.float32)
        I = np.empty((nq, k), dtype=np.int64)
        _lib.check(_lib.load().drag_index_search(self._h, _lib.ptr(q), nq, int(k), _lib.ptr(D),
                                                 _lib.ptr(I)), "drag_index_search")
        return D, I

    def reset(self) -> None:
        _lib.check(_lib.load().drag_index_reset(self._h), "drag_index_reset")
        self._keep.clear()

    # -- device-resident surface ---------------------------------------------------------------
    def add_device(self, x, base_id: Optional[int] = None, copy: bool = False) -> None:
        """Add rows that already live on this GPU (torch float32 CUDA tensor [n, d]).

        copy=False adopts the tensor zero-copy (kept alive by the index): this is how embeddings
        stay resident on the GPU that produced them."""
        import torch
        if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32):
            raise TypeError("add_device: need a float32 CUDA tensor")
        if x.dim() != 2 or x.shape[1] != self.d or x.device.index != self.device:
            raise ValueError(f"add_device: expected [n, {self.d}] on cuda:{self.device}")
        x = x.contiguous()
        base = self.ntotal if base_id is None else int(base_id)
        lib = _lib.load()
        if copy:
            _lib.check(lib.drag_index_add(self._h, _lib.ptr(x), x.shape[0], base, 1,
                                          _lib.current_stream_ptr(x.device)), "drag_index_add")
        else:
            _lib.check(lib.drag_index_adopt(self._h, _lib.ptr(x), x.shape[0], base), "drag_index_adopt")
            self._keep.append(x)

    def search_device(self, q, k: int, out=None):
        """Asynchronous search on the current stream; q float32 CUDA [nq, d] -> (D, I) CUDA tensors."""
        import torch
        if not (isinstance(q, torch.Tensor) and q.is_cuda and q.dtype == torch.float32):
            raise TypeError("search_device: need a float32 CUDA tensor")
        q = q.contiguous()
        nq = q.shape[0]
        if out is None:
            D = torch.empty((nq, k), dtype=torch.float32, device=q.device)
            I = torch.empty((nq, k), dtype=torch.int64, device=q.device)
        else:
            D, I = out
        _lib.check(_lib.load().drag_index_search_device(self._h, _lib.ptr(q), nq, int(k), _lib.ptr(D),
                                                        _lib.ptr(I), _lib.current_stream_ptr(q.device)),
                   "drag_index_search_device")
        return D, I

    def last_launch(self) -> dict:
        g, s, r, b = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _lib.check(_lib.load().drag_index_last_launch(self._h, C.byref(g), C.byref(s), C.byref(r),
                                                      C.byref(b)), "drag_index_last_launch")
        return {"grid": g.value, "stages": s.value, "rows_per_stage": r.value, "nq_batch": b.value}

    def set_timing(self, enable: bool = True) -> None:
        _lib.check(_lib.load().drag_index_set_timing(self._h, int(enable)), "drag_index_set_timing")

    def last_scan_ms(self) -> float:
        ms = C.c_float(0)
        _lib.check(_lib.load().drag_index_last_scan_ms(self._h, C.byref(ms)), "drag_index_last_scan_ms")
        return ms.value

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None and self._h.value:
                _lib.load().drag_index_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass


def merge_topk_device(scores, ids, k_out: int):
    """Merge [nq, lists, k_in] per-shard results into the global top-k_out (CUDA tensors)."""
    import torch
    nq, lists, k_in = scores.shape
    scores = scores.contiguous()
    ids = ids.contiguous()
    D = torch.empty((nq, k_out), dtype=torch.float32, device=scores.device)
    I = torch.empty((nq, k_out), dtype=torch.int64, device=scores.device)
    _lib.check(_lib.load().drag_topk_merge_device(_lib.ptr(scores), _lib.ptr(ids), nq, lists, k_in, k_out,
                                                  _lib.ptr(D), _lib.ptr(I),
                                                  _lib.current_stream_ptr(scores.device)),
               "drag_topk_merge_device")
    return D, I


def shard_bounds(n: int, world: int) -> list:
    """Contiguous balanced row ranges, first `n % world` ranks get one extra row - the same rule as
    the reference's split_samples_for_gpus (outpainting_updown_sampling_redux.py:157-177)."""
    per, rem = divmod(n, world)
    out, s = [], 0
    for r in range(world):
        e = s + per + (1 if r < rem else 0)
        out.append((s, e))
        s = e
    return out


class ShardedIndexFlatIP:
    """Row-sharded exact IP index: rank r owns rows [lo_r, hi_r) of the global corpus.

    search() = local scan x top-k on every rank -> exchange of the per-shard (score, id) [nq, k] lists -> identical
    merge on every rank. The merged result is bit-identical to a single index over the whole corpus because per-row
    scores do not depend on the partition and the merge uses the same (score desc, id asc) order. Two exchanges:
      "p2p"  (after enable_p2p(): ranks of ONE node with NVLink peer access) one fused kernel - every rank stores its
             lists straight into every peer's symmetric buffer, raises a flag, waits for the peers' flags and merges
             (drag_topk_exchange_merge): no collective launch, no packing kernels;
      "nccl" one packed all_gather + drag_topk_merge_device (any topology; also the fallback when nq or k exceed the
             symmetric buffer's capacity).

    `local_search` / `merge` default to the CUDA kernels; tests inject oracle callables to exercise
    the partition / id-offset / gather logic under gloo on CPU.
    """

    def __init__(self, d: int, rank: int, world: int, device: Optional[int] = None,
                 local_search: Optional[Callable] = None, merge: Optional[Callable] = None,
                 all_gather: Optional[Callable] = None):
        self.d, self.rank, self.world = int(d), int(rank), int(world)
        self._local_search = local_search
        self._merge = merge
        self._all_gather = all_gather
        self._index = None
        if local_search is None:
            self._index = IndexFlatIP(d, device if device is not None else 0)
        self._rows = None
        self.lo = self.hi = 0
        self.ntotal_global = 0
        self._p2p = None            # (symmetric buffer, handle, ctypes pointer array, nq_cap, k_cap)
        self._epoch = 0
        self.exchange = "nccl"

    def enable_p2p(self, group=None, nq_cap: int = 64, k_cap: int = 1024) -> bool:
        """Collective. Allocate one symmetric buffer per rank (torch.distributed._symmetric_memory: the same allocation
        mapped into every peer of the node) for the fused NVLink exchange. Returns False (and keeps the NCCL exchange) when
        symmetric memory is unavailable for this group - e.g. ranks on different nodes."""
        import ctypes as C

        import torch
        import torch.distributed as dist
        if self.world == 1 or self._index is None:
            return False
        try:
            import torch.distributed._symmetric_memory as symm
            lib = _lib.load()
            nbytes = C.c_int64(0)
            _lib.check(lib.drag_topk_exchange_buffer_bytes(self.world, nq_cap, k_cap, C.byref(nbytes)),
                       "drag_topk_exchange_buffer_bytes")
            dev = torch.device("cuda", self._index.device)
            buf = symm.empty(int(nbytes.value), dtype=torch.uint8, device=dev)
            buf.zero_()
            hdl = symm.rendezvous(buf, group if group is not None else dist.group.WORLD)
            ptrs = (C.c_void_p * self.world)(*[int(p) for p in hdl.buffer_ptrs])
            torch.cuda.synchronize(dev)
            dist.barrier(group)                      # every rank's flags are zero before anyone pushes
            self._p2p = (buf, hdl, ptrs, int(nq_cap), int(k_cap))
            ok = True
        except Exception as e:                       # noqa: BLE001 - report and keep the NCCL exchange
            print(f"[ShardedIndexFlatIP] symmetric-memory exchange unavailable ({type(e).__name__}: {e}); using NCCL all_gather")
            self._p2p = None
            ok = False
        # every rank must take the same exchange: agree on the minimum
        flag = torch.tensor([int(ok)], device=torch.device("cuda", self._index.device))
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        ok = bool(flag.item())
        if not ok:
            self._p2p = None
        self._epoch = 0
        self.exchange = "p2p" if ok else "nccl"
        return ok

    def add_local(self, x_local, lo: int, ntotal_global: int) -> None:
        """Register this rank's shard: rows [lo, lo + len(x_local)) of a corpus of ntotal_global."""
        self.lo, self.hi = int(lo), int(lo) + int(x_local.shape[0])
        self.ntotal_global = int(ntotal_global)
        if self._index is not None:
            self._index.add_device(x_local, base_id=self.lo)
        else:
            self._rows = x_local

    def search(self, q, k: int):
        import torch
        import torch.distributed as dist
        p2p = (self.world > 1 and self._index is not None and self._p2p is not None and self.exchange == "p2p"
               and q.shape[0] <= self._p2p[3] and k <= self._p2p[4] and self.world * k <= 8192)
        if p2p:                                   # local scan + fused NVLink exchange + merge behind ONE C call
            _, _, ptrs, nq_cap, k_cap = self._p2p
            if not (isinstance(q, torch.Tensor) and q.is_cuda and q.dtype == torch.float32):
                raise TypeError("search: need a float32 CUDA tensor")
            q = q.contiguous()
            nq = q.shape[0]
            self._epoch += 1
            D = torch.empty((nq, k), dtype=torch.float32, device=q.device)
            I = torch.empty((nq, k), dtype=torch.int64, device=q.device)
            Do, Io = torch.empty_like(D), torch.empty_like(I)
            _lib.check(_lib.load().drag_index_search_sharded(self._index._h, _lib.ptr(q), nq, int(k), ptrs, self.world,
                                                             self.rank, nq_cap, k_cap, self._epoch, _lib.ptr(D), _lib.ptr(I),
                                                             _lib.ptr(Do), _lib.ptr(Io), _lib.current_stream_ptr(q.device)),
                       "drag_index_search_sharded")
            return Do, Io
        if self._index is not None:
            D, I = self._index.search_device(q, k)
        else:
            D, I = self._local_search(self._rows, q, k, self.lo)
        if self.world == 1:
            return D, I
        if self._all_gather is not None:
            Dg, Ig = self._all_gather(D), self._all_gather(I)
        else:
            # ONE collective per search: scores bit-cast to int32, widened to int64 and sent as the second column of
            # the [nq, k, 2] int64 (id, score) pairs (two separate all-gathers cost two NCCL latencies; at nq = 1 the
            # exchange is pure latency: nq * k * 16 B per rank)
            pairs = torch.stack([I, D.contiguous().view(torch.int32).to(torch.int64)], dim=-1).contiguous()
            out = torch.empty((self.world,) + tuple(pairs.shape), dtype=torch.int64, device=pairs.device)
            dist.all_gather(list(out.unbind(0)), pairs)          # contiguous views of one buffer: one collective
            out = out.permute(1, 0, 2, 3)                                   # [nq, world, k, 2]
            Ig = out[..., 0].contiguous()
            Dg = out[..., 1].to(torch.int32).contiguous().view(torch.float32)
        if self._merge is not None:
            return self._merge(Dg, Ig, k)
        return merge_topk_device(Dg, Ig, k)
