"""CPU, world_size 2, gloo: host-side logic of the sharded index (row partition, id offsets, all-gather of
per-shard top-k, merge) with the oracle injected as the per-shard scan/merge - the CUDA kernels that
replace them on the GPU are covered by tests/test_topk_gpu.py::test_sharded_merge_equals_single."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from domain_rag_b200.index import ShardedIndexFlatIP, shard_bounds
from oracle import ip_topk as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, d, nq, k, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = np.random.default_rng(0)
        x = g.standard_normal((n, d)).astype(np.float32)
        x[n // 2 + 3] = x[5]                      # a duplicate across shards: tie -> lower id first
        q = g.standard_normal((nq, d)).astype(np.float32)
        lo, hi = shard_bounds(n, world)[rank]

        def local_search(rows, qq, kk, base):
            D, I = O.ip_topk(rows.numpy(), qq.numpy(), kk, base_id=base)
            return torch.from_numpy(D), torch.from_numpy(I)

        def merge(Dg, Ig, kk):
            D, I = O.merge_topk(Dg.numpy(), Ig.numpy(), kk)
            return torch.from_numpy(D), torch.from_numpy(I)

        six = ShardedIndexFlatIP(d, rank, world, local_search=local_search, merge=merge)
        six.add_local(torch.from_numpy(x[lo:hi]), lo=lo, ntotal_global=n)
        D, I = six.search(torch.from_numpy(q), k)
        Do, Io = O.ip_topk(x, q, k)
        ok = np.array_equal(I.numpy(), Io) and np.array_equal(D.numpy(), Do)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,k,world", [(1001, 10, 2), (37, 100, 2), (501, 16, 3)])
def test_sharded_search_equals_single_index_under_gloo(n, k, world):
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, n, 32, 3, k, ret), nprocs=world, join=True)
    assert dict(ret) == {r: True for r in range(world)}


def test_shard_bounds_matches_reference_split_rule():
    from oracle import host_helpers
    for n, w in [(10, 3), (7, 8), (32, 8), (33, 8), (0, 4)]:
        parts = host_helpers.split_samples_for_gpus(list(range(n)), w) if w > 1 else [list(range(n))]
        got = shard_bounds(n, w)
        assert [len(p) for p in parts] == [hi - lo for lo, hi in got]
        assert got[0][0] == 0 and got[-1][1] == n
