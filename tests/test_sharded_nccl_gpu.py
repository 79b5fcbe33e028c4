"""GPU, world_size 2 (skipped on a one-GPU box): the sharded index ON HARDWARE - local scan x top-k on each GPU, then both
exchanges (the fused NVLink peer-memory kernel and the packed NCCL all-gather + merge kernel) - must return, bit for bit,
what ONE index over the whole corpus returns, and what the CPU oracle returns."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    import torch.distributed as dist

    from domain_rag_b200.index import IndexFlatIP, ShardedIndexFlatIP, shard_bounds
    from oracle import ip_topk as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        out = {}
        for (n, d, nq, k) in [(20001, 512, 7, 100), (3000, 768, 1, 10), (513, 128, 64, 100), (50, 64, 3, 100)]:
            g = np.random.default_rng(n)
            x = g.standard_normal((n, d)).astype(np.float32)
            x /= np.linalg.norm(x, axis=1, keepdims=True)
            x[n // 2 + 3] = x[5]                              # duplicate across shards: tie -> lower id first
            q = x[[5] + list(g.integers(0, n, nq - 1))] + 0.01 * g.standard_normal((nq, d)).astype(np.float32)
            lo, hi = shard_bounds(n, world)[rank]
            six = ShardedIndexFlatIP(d, rank, world, device=rank)
            six.add_local(torch.from_numpy(x[lo:hi]).cuda(), lo=lo, ntotal_global=n)
            qd = torch.from_numpy(q).cuda()
            Dn, In = six.search(qd, k)                        # NCCL all-gather + merge kernel
            p2p = six.enable_p2p()
            res = [six.search(qd, k) for _ in range(5)]       # fused peer-memory exchange, several epochs
            single = IndexFlatIP(d, rank)
            single.add(x)
            D1, I1 = single.search(q, k)
            Do, Io = O.ip_topk(x, q, min(k, n))
            kk = min(k, n)
            ok = (np.array_equal(In.cpu().numpy(), I1) and np.array_equal(Dn.cpu().numpy(), D1)
                  and all(np.array_equal(I.cpu().numpy(), I1) and np.array_equal(D.cpu().numpy(), D1) for D, I in res)
                  and np.array_equal(I1[:, :kk], Io) and np.allclose(D1[:, :kk], Do, rtol=0, atol=1e-6))
            out[(n, d, nq, k)] = (bool(ok), six.exchange if p2p else "nccl-only")
        ret[rank] = out
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_search_on_two_gpus_equals_single_index(lib):
    import torch.multiprocessing as mp
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    got = dict(ret)
    print("sharded search on 2 GPUs:", got)
    assert set(got) == {0, 1}
    for r in (0, 1):
        assert all(ok for ok, _ in got[r].values()), got[r]
    assert all(path == "p2p" for _, path in got[0].values()), "the fused NVLink exchange was not exercised"
