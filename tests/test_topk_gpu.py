"""GPU parity: scan x top-k kernel (through the C ABI) vs the CPU oracle. Indices bit-exact."""
import numpy as np
import pytest
import torch

from oracle import ip_topk as O
from tests.synth import well_separated

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def IndexFlatIP(lib):
    from domain_rag_b200.index import IndexFlatIP
    return IndexFlatIP


def run_host(IndexFlatIP, x, q, k):
    ix = IndexFlatIP(x.shape[1])
    ix.add(x)
    assert ix.ntotal == x.shape[0]
    return ix.search(q, k)


def assert_parity(D, I, x, q, k):
    Do, Io = O.ip_topk(x, q, k)
    np.testing.assert_array_equal(I, Io)
    np.testing.assert_allclose(D, Do, rtol=0, atol=1e-6)


@pytest.mark.parametrize("n,d,nq,k", [
    (128, 768, 8, 10),        # BASELINE config C1 (ViT-L/14 width)
    (10000, 512, 7, 100),     # C2: ArTaxOr 1-shot, 7 queries, ViT-B/32 width
    (10000, 768, 7, 100),     # C2 at ViT-L/14 width
    (5000, 1024, 3, 100),
    (3001, 128, 1, 1),
    (2500, 640, 2, 50),       # NV=5 register path
    (4097, 100, 5, 33),       # D%128 != 0 -> generic bulk path
    (777, 36, 9, 64),         # nq > batch width 8
    (1500, 130, 2, 17),       # D%4 != 0 -> direct path
    (64, 512, 1, 100),        # k > N: padded with (-FLT_MAX, -1)
    (1, 512, 2, 5),
    (20000, 512, 1, 1000),    # large k
])
def test_search_matches_oracle(IndexFlatIP, n, d, nq, k):
    x, q = well_separated(n, d, nq, k, seed=4000 + n + d)
    D, I = run_host(IndexFlatIP, x, q, k)
    assert D.shape == (nq, k) and I.shape == (nq, k) and I.dtype == np.int64
    assert_parity(D, I, x, q, k)


def test_ties_lower_index_first(IndexFlatIP):
    # integer-valued vectors: every dot product is exact in any summation order -> real ties
    g = np.random.default_rng(7)
    base = g.integers(-3, 4, size=(300, 256)).astype(np.float32)
    x = np.concatenate([base, base[:100], base[50:150]], 0)   # duplicates far apart
    q = g.integers(-3, 4, size=(4, 256)).astype(np.float32)
    D, I = run_host(IndexFlatIP, x, q, 100)
    Do, Io = O.ip_topk(x, q, 100)
    np.testing.assert_array_equal(I, Io)
    np.testing.assert_array_equal(D, Do)
    # all-equal corpus: top-k must be ids 0..k-1
    x0 = np.ones((5000, 128), np.float32)
    D, I = run_host(IndexFlatIP, x0, np.ones((1, 128), np.float32), 100)
    np.testing.assert_array_equal(I[0], np.arange(100))


def test_empty_and_segments(IndexFlatIP):
    ix = IndexFlatIP(64)
    D, I = ix.search(np.zeros((2, 64), np.float32), 5)
    assert np.all(I == -1) and np.all(D == O.FAISS_MISSING_SCORE)
    # several add() calls (the reference vstacks per-source arrays, :406-419) == one big add
    x, q = well_separated(6000, 64, 3, 20, seed=11)
    for lo, hi in [(0, 1000), (1000, 1001), (1001, 4500), (4500, 6000)]:
        ix.add(x[lo:hi])
    D, I = ix.search(q, 20)
    assert_parity(D, I, x, q, 20)
    ix.reset()
    assert ix.ntotal == 0


def test_device_api_and_idempotence(IndexFlatIP):
    x, q = well_separated(30000, 512, 4, 100, seed=21)
    ix = IndexFlatIP(512)
    xt = torch.from_numpy(x).cuda()
    ix.add_device(xt)                      # zero-copy adoption of resident embeddings
    qt = torch.from_numpy(q).cuda()
    D1, I1 = ix.search_device(qt, 100)
    D2, I2 = ix.search_device(qt, 100)     # same launch twice -> identical
    torch.cuda.synchronize()
    assert torch.equal(I1, I2) and torch.equal(D1, D2)
    assert_parity(D1.cpu().numpy(), I1.cpu().numpy(), x, q, 100)
    # self-retrieval: a corpus row used as the query must come back first
    D3, I3 = ix.search_device(xt[123:125].contiguous(), 1)
    assert I3.flatten().tolist() == [123, 124]


def test_sharded_merge_equals_single(IndexFlatIP):
    """Shards scanned independently (as separate ranks would) + merge kernel == single index."""
    from domain_rag_b200.index import merge_topk_device, shard_bounds
    x, q = well_separated(40000, 512, 5, 100, seed=31)
    qt = torch.from_numpy(q).cuda()
    Ds, Is = [], []
    for lo, hi in shard_bounds(len(x), 8):
        ix = IndexFlatIP(512)
        ix.add_device(torch.from_numpy(x[lo:hi]).cuda(), base_id=lo)
        D, I = ix.search_device(qt, 100)
        Ds.append(D)
        Is.append(I)
    Dm, Im = merge_topk_device(torch.stack(Ds, 1), torch.stack(Is, 1), 100)
    assert_parity(Dm.cpu().numpy(), Im.cpu().numpy(), x, q, 100)
    single = IndexFlatIP(512)
    single.add(x)
    D1, I1 = single.search(q, 100)
    np.testing.assert_array_equal(Im.cpu().numpy(), I1)
    np.testing.assert_array_equal(Dm.cpu().numpy(), D1)   # scores are partition independent


def test_full_size_properties(IndexFlatIP):
    """BASELINE C5 full size (1M x 512): size-independent properties instead of a CPU re-scan of
    every query: sortedness, planted rows recovered in order, scores equal fp64 recomputation."""
    n, d, k = 1_000_000, 512, 100
    g = torch.Generator(device="cuda").manual_seed(4006)
    x = torch.randn(n, d, generator=g, device="cuda")
    x /= x.norm(dim=1, keepdim=True)
    q = torch.randn(2, d, generator=g, device="cuda")
    q /= q.norm(dim=1, keepdim=True)
    planted = [17, 499_999, 999_999, 250_000]
    for j, r in enumerate(planted):          # strictly decreasing, far above the random scores
        x[r] = q[0] * (0.99 - 0.05 * j) + x[r] * 0.01
    ix = IndexFlatIP(d)
    ix.add_device(x)
    D, I = ix.search_device(q, k)
    torch.cuda.synchronize()
    assert I[0, :4].tolist() == planted
    assert bool((D[:, 1:] <= D[:, :-1]).all())
    assert len(set(I[0].tolist())) == k and len(set(I[1].tolist())) == k
    ref = (x[I[1]].double() @ q[1].double()).float()
    assert torch.allclose(D[1], ref, rtol=0, atol=1e-6)
    # exact agreement with the oracle on the returned candidates + the true k-th threshold
    s = (x.double() @ q[1].double())
    kth = torch.topk(s, k).values[-1]
    assert abs(float(D[1, -1]) - float(kth)) <= 1e-6


def test_first_stage_retrieval_matches_reference_call_site_golden(lib, golden_dir):
    """domain_rag_b200.retrieval.clip_first_stage_retrieval (the drop-in for :396-451) on the GPU reproduces, record for
    record, what the reference's own function emitted for a seeded multi-source corpus (tests/golden/ref_extra.json,
    oracle/make_golden_extra.py): source order, ids, duplicate across sources, k clamp, empty and None sources."""
    import json

    from domain_rag_b200 import retrieval as R
    g = json.load(open(golden_dir / "ref_extra.json"))
    arrs = np.load(golden_dir / "ref_extra_arrays.npz")
    feats = {"coco": arrs["coco"], "empty": np.zeros((0, 64), np.float32), "none": None,
             "mini_imagenet": arrs["mini_imagenet"]}
    paths = {"coco": [f"../../datasets/coco/train2017/{i:012d}.jpg" for i in range(300)], "empty": [], "none": [],
             "mini_imagenet": [f"./mini/n{i:05d}.JPEG" for i in range(150)]}
    for case in g["first_stage"]:
        q = arrs["queries"][case["query"]]
        if case.get("only_empty"):
            assert R.clip_first_stage_retrieval(q, {"empty": feats["empty"], "none": None}, {"empty": [], "none": []},
                                                top_k=case["top_k"]) == []
            continue
        got = R.clip_first_stage_retrieval(q, feats, paths, top_k=case["top_k"])
        want = case["records"]
        assert [(r["index"], r["image_path"], r["source_dataset"]) for r in got] == \
               [(r["index"], r["image_path"], r["source_dataset"]) for r in want]
        np.testing.assert_allclose([r["similarity"] for r in got], [r["similarity"] for r in want], rtol=0, atol=1e-6)
