"""GPU parity tests for IndexFlatIP / IndexIDMap: CUDA path (through the C-ABI) vs the CPU oracle.
Tolerances (BASELINE.json north_star): scores within 1e-5 absolute of the fp64-accumulated oracle;
id lists identical outside near-tie bands of 2e-6 (oracle.compare_topk), ties -> lowest position."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def faiss():
    from wise_b200 import faiss_compat
    return faiss_compat


def _flat(faiss, xb, ids=None):
    if ids is None:
        idx = faiss.IndexFlatIP(xb.shape[1])
        idx.add(xb)
    else:
        idx = faiss.IndexIDMap(faiss.IndexFlatIP(xb.shape[1]))
        idx.add_with_ids(xb, ids)
    return idx


def test_golden_fixture(faiss):
    g = np.load(os.path.join(GOLD, "flat_small.npz"))
    idx = _flat(faiss, g["xb"], g["ids"])
    D, I = idx.search(g["xq"], int(g["k"]))
    r = O.compare_topk(D, I, g["D"], g["I"])
    assert r["rows"] == 4
    # exact duplicate rows 5 / 205 have bit-identical scores on the GPU too: lowest position first
    assert I[3, 0] == 5 * 7 + 3 and I[3, 1] == 205 * 7 + 3 and D[3, 0] == D[3, 1]


def test_config1_100k_x_512_16q_top10(faiss):
    """BASELINE config[0] in full."""
    xb = O.unit_gaussian(100000, 512, 1234)
    xq = O.unit_gaussian(16, 512, 4321)
    idx = _flat(faiss, xb)
    assert idx.ntotal == 100000 and idx.d == 512 and idx.is_trained
    D, I = idx.search(xq, 10)
    assert D.dtype == np.float32 and I.dtype == np.int64 and D.shape == (16, 10)
    Dr, Ir = O.flat_search(xb, xq, 10)
    r = O.compare_topk(D, I, Dr, Ir)
    assert r["exact_rows"] >= 15, r


def test_config1_with_one_percent_duplicates(faiss):
    xb = O.unit_gaussian(100000, 512, 1234)
    xb[50000:51000] = xb[:1000]
    xq = np.concatenate([O.unit_gaussian(12, 512, 4321), xb[10:14]])
    ids = np.arange(100000, dtype=np.int64) * 2 + 1
    idx = _flat(faiss, xb, ids)
    D, I = idx.search(xq, 10)
    Dr, Ir = O.flat_search(xb, xq, 10, ids)
    O.compare_topk(D, I, Dr, Ir)
    for j in range(4):  # a query equal to a duplicated row: the two copies tie exactly, lowest position wins
        assert I[12 + j, 0] == ids[10 + j] and I[12 + j, 1] == ids[50010 + j]


@pytest.mark.parametrize("n,d,nq,k", [
    (1, 4, 1, 1), (5, 64, 2, 10), (31, 8, 3, 4), (33, 12, 1, 40), (1000, 64, 1, 5), (20000, 100, 5, 7),
    (20000, 30, 9, 2048), (33333, 1024, 3, 1000), (50000, 768, 8, 100), (40000, 1536, 2, 10),
    (3000, 4096, 4, 16), (100000, 512, 40, 1), (70000, 768, 1, 100), (25000, 2048, 7, 33)])
def test_shapes(faiss, n, d, nq, k):
    xb = O.unit_gaussian(n, d, n + d)
    xq = O.unit_gaussian(nq, d, nq + k)
    idx = _flat(faiss, xb)
    D, I = idx.search(xq, k)
    Dr, Ir = O.flat_search(xb, xq, k)
    O.compare_topk(D, I, Dr, Ir)
    if k > n:
        assert np.all(I[:, n:] == -1) and np.all(D[:, n:] == O.NEG_FLT_MAX)


def test_clustered_clip_like_data_top100(faiss):
    xb = O.clustered_unit(200000, 768, 512, 2024)
    xq = O.clustered_unit(8, 768, 512, 2025)
    idx = _flat(faiss, xb)
    D, I = idx.search(xq, 100)
    Dr, Ir = O.flat_search(xb, xq, 100)
    O.compare_topk(D, I, Dr, Ir)
    D1, I1 = idx.search(xq[:1], 100)  # n=1 (the API's case) takes a different kernel instance
    O.compare_topk(D1, I1, Dr[:1], Ir[:1])


def test_empty_index_and_incremental_adds(faiss):
    idx = faiss.IndexIDMap(faiss.IndexFlatIP(32))
    xq = O.unit_gaussian(2, 32, 1)
    D, I = idx.search(xq, 5)
    assert np.all(I == -1) and np.all(D == O.NEG_FLT_MAX)
    xb = O.unit_gaussian(3000, 32, 2)
    ids = np.arange(3000, dtype=np.int64)[::-1].copy() + 10
    for s in range(0, 3000, 512):  # the reference adds in 512-row batches (feature_search_index.py:79-82)
        idx.add_with_ids(xb[s:s + 512], ids[s:s + 512])
    assert idx.ntotal == 3000
    D, I = idx.search(xq, 5)
    O.compare_topk(D, I, *O.flat_search(xb, xq, 5, ids))


def test_api_surface_and_errors(faiss):
    flat = faiss.IndexFlatIP(16)
    assert not hasattr(flat, "nprobe") and not hasattr(flat, "direct_map")  # api/routes.py:899,1317 use hasattr
    with pytest.raises(RuntimeError):
        flat.add_with_ids(O.unit_gaussian(2, 16, 0), np.array([1, 2]))
    idm = faiss.IndexIDMap(flat)
    assert not hasattr(idm, "nprobe") and not hasattr(idm, "direct_map")
    with pytest.raises(TypeError):
        idm.search(np.zeros((1, 16), np.float64), 3)
    with pytest.raises(AssertionError):
        idm.search(np.zeros((1, 8), np.float32), 3)
    with pytest.raises(RuntimeError):
        idm.search(np.zeros((1, 16), np.float32), 5000)  # k > WB_MAX_K
    flat2 = faiss.IndexFlatIP(16)
    flat2.add(O.unit_gaussian(4, 16, 0))
    with pytest.raises(RuntimeError):
        faiss.IndexIDMap(flat2)  # "index must be empty on input"


def test_device_pointer_api_and_merge(faiss):
    """wb_search_dev / wb_add_with_ids_dev / wb_merge_topk_dev with torch-owned device memory."""
    import torch
    from wise_b200 import _capi
    L = _capi.lib()
    n, d, nq, k = 60000, 256, 5, 50
    xb = O.unit_gaussian(n, d, 3)
    xq = O.unit_gaussian(nq, d, 4)
    Dr, Ir = O.flat_search(xb, xq, k)
    st = torch.cuda.current_stream().cuda_stream
    parts_D, parts_I = [], []
    bounds = [0, 20000, 20001, 45000, n]
    xd = torch.from_numpy(xb).cuda()
    qd = torch.from_numpy(xq).cuda()
    for a, b in zip(bounds[:-1], bounds[1:]):  # 4 "ranks" holding contiguous row ranges
        idx = faiss.IndexFlatIP(d)
        ids = torch.arange(a, b, dtype=torch.int64, device="cuda")
        _capi.check(L.wb_add_with_ids_dev(idx._h, b - a, xd[a:b].contiguous().data_ptr(), ids.data_ptr(), st))
        D = torch.empty(nq, k, device="cuda")
        I = torch.empty(nq, k, dtype=torch.int64, device="cuda")
        _capi.check(L.wb_search_dev(idx._h, nq, qd.data_ptr(), k, 1, D.data_ptr(), I.data_ptr(), st))
        torch.cuda.synchronize()
        parts_D.append(D); parts_I.append(I)
    Dp = torch.stack(parts_D).contiguous(); Ip = torch.stack(parts_I).contiguous()
    Dm = torch.empty(nq, k, device="cuda"); Im = torch.empty(nq, k, dtype=torch.int64, device="cuda")
    _capi.check(L.wb_merge_topk_dev(0, nq, k, 4, Dp.data_ptr(), Ip.data_ptr(), Dm.data_ptr(), Im.data_ptr(), st))
    torch.cuda.synchronize()
    O.compare_topk(Dm.cpu().numpy(), Im.cpu().numpy(), Dr, Ir)
    # the single-row shard returns k-1 padded slots; merged result must not contain them
    assert (Ip[1, :, 1:] == -1).all() and (Im >= 0).all()


def test_sharded_result_is_bit_identical_to_single(faiss):
    """Row scores do not depend on where a row lives => merge of shards == single index, bit for bit."""
    import torch
    from wise_b200 import _capi
    L = _capi.lib()
    n, d, nq, k = 80000, 768, 4, 100
    xb = O.clustered_unit(n, d, 64, 7)
    xq = O.clustered_unit(nq, d, 64, 8)
    one = _flat(faiss, xb)
    D1, I1 = one.search(xq, k)
    st = torch.cuda.current_stream().cuda_stream
    G = 8
    Dp = torch.empty(G, nq, k, device="cuda"); Ip = torch.empty(G, nq, k, dtype=torch.int64, device="cuda")
    keep = []
    for r in range(G):
        a, b = n * r // G, n * (r + 1) // G
        idx = faiss.IndexIDMap(faiss.IndexFlatIP(d))
        idx.add_with_ids(xb[a:b], np.arange(a, b, dtype=np.int64))
        keep.append(idx)
        qd = torch.from_numpy(xq).cuda()
        _capi.check(L.wb_search_dev(idx._h, nq, qd.data_ptr(), k, 1, Dp[r].data_ptr(), Ip[r].data_ptr(), st))
    Dm = torch.empty(nq, k, device="cuda"); Im = torch.empty(nq, k, dtype=torch.int64, device="cuda")
    _capi.check(L.wb_merge_topk_dev(0, nq, k, G, Dp.data_ptr(), Ip.data_ptr(), Dm.data_ptr(), Im.data_ptr(), st))
    torch.cuda.synchronize()
    assert np.array_equal(Im.cpu().numpy(), I1) and np.array_equal(Dm.cpu().numpy(), D1)


def test_reconstruct_and_export(faiss):
    xb = O.unit_gaussian(5000, 70, 5)  # d % 4 != 0: padded row stride in HBM
    ids = np.arange(5000, dtype=np.int64) * 3 + 7
    idx = _flat(faiss, xb, ids)
    x, i, _ = idx._export(100, 50)
    assert np.array_equal(x, xb[100:150]) and np.array_equal(i, ids[100:150])
    rec = faiss._reconstruct(idx, [7, 3 * 4999 + 7, 3 * 17 + 7])
    assert np.array_equal(rec, xb[[0, 4999, 17]])
    with pytest.raises(RuntimeError):
        faiss._reconstruct(idx, [8])


def test_large_scan_properties_2m_x_768(faiss):
    """Size-independent checks at a size the oracle cannot brute-force quickly: returned scores equal
    fp64 recomputation within 1e-5, rows sorted, and no row in a 200k-row sample beats the k-th score."""
    import torch
    from wise_b200 import _capi
    L = _capi.lib()
    n, d, k, nq = 2_000_000, 768, 100, 3
    gen = torch.Generator(device="cuda"); gen.manual_seed(2024)
    x = torch.randn(n, d, device="cuda", generator=gen)
    x /= x.norm(dim=1, keepdim=True)
    q = x[[5, 1_000_000, n - 1]].clone() * 0.7 + 0.3 * torch.nn.functional.normalize(torch.randn(nq, d, device="cuda", generator=gen), dim=1)
    q = torch.nn.functional.normalize(q, dim=1).contiguous()
    idx = faiss.IndexFlatIP(d)
    idx.reserve(n)
    _capi.check(L.wb_add_with_ids_dev(idx._h, n, x.data_ptr(), None, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    D, I = idx.search(q.cpu().numpy(), k)
    assert np.all(np.diff(D, axis=1) <= 0) and np.all(I >= 0)
    assert I[0, 0] == 5 and I[1, 0] == 1_000_000 and I[2, 0] == n - 1
    qh = q.cpu().numpy().astype(np.float64)
    for j in range(nq):
        rows = x[torch.from_numpy(I[j]).cuda()].cpu().numpy().astype(np.float64)
        assert np.abs(rows @ qh[j] - D[j]).max() <= 1e-5
        assert len(set(I[j].tolist())) == k
    samp = torch.randint(0, n, (200_000,), device="cuda", generator=gen)
    s = (x[samp].double() @ q.double().T).cpu().numpy()  # [200k, nq]
    samp_h = samp.cpu().numpy()
    for j in range(nq):
        better = samp_h[s[:, j] > D[j, -1] + 2e-6]
        assert set(better.tolist()) <= set(I[j].tolist())


# ---- K2: tensor-core path (batches of 5+ queries; WB_GEMM_FORCE=1 lifts the store-size rule) ------------
def _gemm_stats(idx):
    from wise_b200 import _capi
    a, b = C.c_int64(), C.c_int64()
    _capi.lib().wb_gemm_stats(idx._h, C.byref(a), C.byref(b))
    return a.value, b.value


@pytest.mark.parametrize("n,d,nq,k,clustered", [
    (40000, 768, 16, 10, False), (40000, 768, 128, 10, False), (100000, 512, 200, 100, False),
    (50000, 100, 33, 5, False), (131072, 768, 300, 100, True), (70001, 1024, 64, 50, False),
    (300000, 64, 1000, 10, False), (60000, 70, 129, 100, True), (200000, 768, 1024, 100, True),
    (300000, 32, 10, 2048, False), (270000, 48, 40, 1000, True), (99999, 260, 65, 7, False)])
def test_gemm_path_matches_oracle(faiss, monkeypatch, n, d, nq, k, clustered):
    """WB_GEMM_FORCE=1 bypasses the size heuristic so that small stores exercise K2 too.  3xTF32 on tcgen05: scores within 1e-5 of the fp64 oracle (observed ~1.5e-6), ids identical outside
    near-tie bands; the band is widened to 4e-6 here because the tensor-core accumulation noise is ~1.5e-6."""
    monkeypatch.setenv("WB_GEMM_FORCE", "1")
    xb = O.clustered_unit(n, d, 64, 1) if clustered else O.unit_gaussian(n, d, 100)
    xq = O.clustered_unit(nq, d, 64, 2) if clustered else O.unit_gaussian(nq, d, 200)
    ids = np.arange(n, dtype=np.int64) * 3 + 1
    idx = _flat(faiss, xb, ids)
    D, I = idx.search(xq, k)
    epochs, fallbacks = _gemm_stats(idx)
    assert epochs >= 3 and fallbacks == 0, "the tensor-core path must have served this batch"
    Dr, Ir = O.flat_search(xb, xq, k, ids)
    O.compare_topk(D, I, Dr, Ir, band=4e-6)
    # the CUDA-core path on the same index gives the same members (different kernel, same contract)
    D8, I8 = idx.search(xq[:4], k)
    O.compare_topk(D8, I8, Dr[:4], Ir[:4])


@pytest.mark.parametrize("filt", ["0", "1"])
def test_gemm_filter_and_refine_unnormalised(faiss, monkeypatch, filt):
    """K2's later epochs issue only the hi x hi term and keep a row when its one-term score exceeds
    thr - margin, margin = c * |q| * max|x|; candidates are re-scored exactly.  Rows and queries with very different
    norms (0.25 .. 4) make a too-small margin lose true neighbours; WB_GEMM_FILTER=0 is the all-3xTF32 path."""
    monkeypatch.setenv("WB_GEMM_FORCE", "1")
    monkeypatch.setenv("WB_GEMM_FILTER", filt)
    n, d, nq, k = 150000, 384, 48, 50
    rng = np.random.default_rng(11)
    xb = O.unit_gaussian(n, d, 21) * rng.uniform(0.25, 4.0, size=(n, 1)).astype(np.float32)
    xq = O.unit_gaussian(nq, d, 22) * rng.uniform(0.5, 2.0, size=(nq, 1)).astype(np.float32)
    idx = _flat(faiss, xb)
    D, I = idx.search(xq, k)
    epochs, fallbacks = _gemm_stats(idx)
    assert epochs >= 3 and fallbacks == 0
    Dr, Ir = O.flat_search(xb, xq, k)
    scale = 4.0 * 2.0  # rounding errors scale with |q||x| (1 for the unit-norm features the tolerances are quoted for)
    O.compare_topk(D, I, Dr, Ir, score_tol=1e-5 * scale, band=4e-6 * scale)


def test_kernel_choice_by_store_size(faiss):
    """Dispatch heuristic: 16 queries over a 200 MB store stay on the CUDA-core scan (K2's fixed cost would
    dominate); 256 queries over the same store go to the tensor cores.  Same results either way."""
    n, d, k = 100000, 512, 10
    xb = O.unit_gaussian(n, d, 3)
    xq = O.unit_gaussian(256, d, 4)
    idx = _flat(faiss, xb)
    Dr, Ir = O.flat_search(xb, xq, k)
    D, I = idx.search(xq[:16], k)
    assert _gemm_stats(idx)[0] == 0
    O.compare_topk(D, I, Dr[:16], Ir[:16])
    D, I = idx.search(xq, k)
    assert _gemm_stats(idx)[0] > 0
    O.compare_topk(D, I, Dr, Ir, band=4e-6)


def test_gemm_path_duplicates_tie_rule(faiss, monkeypatch):
    monkeypatch.setenv("WB_GEMM_FORCE", "1")
    n, d, k = 60000, 256, 20
    xb = O.unit_gaussian(n, d, 5)
    xb[30000:30500] = xb[:500]
    xq = np.concatenate([O.unit_gaussian(28, d, 6), xb[100:104]])
    idx = _flat(faiss, xb)
    D, I = idx.search(xq, k)
    assert _gemm_stats(idx)[0] > 0
    O.compare_topk(D, I, *O.flat_search(xb, xq, k), band=4e-6)
    for j in range(4):  # exact duplicates: bit-identical scores, lowest position first
        assert I[28 + j, 0] == 100 + j and I[28 + j, 1] == 30100 + j and D[28 + j, 0] == D[28 + j, 1]


def test_gemm_overflow_falls_back_exactly(faiss, monkeypatch):
    """Adversarial order: rows sorted by ascending score for every query, so every row beats the running
    threshold and the candidate lists overflow; the batch is then repaired by the CUDA-core scan."""
    monkeypatch.setenv("WB_GEMM_FORCE", "1")
    n, d, k = 120000, 64, 10
    rng = np.random.default_rng(0)
    base = O.unit_gaussian(1, d, 1)[0]
    t = np.linspace(-0.9, 0.9, n).astype(np.float32)  # score of row i against `base` grows with i
    noise = O.unit_gaussian(n, d, 2)
    noise -= (noise @ base)[:, None] * base[None, :]
    noise /= np.linalg.norm(noise, axis=1, keepdims=True)
    xb = (t[:, None] * base[None, :] + np.sqrt(1 - t[:, None] ** 2) * noise).astype(np.float32)
    xq = np.repeat(base[None, :], 16, axis=0) + 1e-3 * rng.standard_normal((16, d)).astype(np.float32)
    xq /= np.linalg.norm(xq, axis=1, keepdims=True)
    idx = _flat(faiss, xb)
    D, I = idx.search(xq.astype(np.float32), k)
    epochs, fallbacks = _gemm_stats(idx)
    assert epochs > 0 and fallbacks == 1
    O.compare_topk(D, I, *O.flat_search(xb, xq.astype(np.float32), k))


def test_gemm_score_audit_repeated(faiss, monkeypatch):
    """The audit that caught the raw-stage release race in round 1 (profiles/r01/gemm_experiments.md), kept in the
    suite: six fresh indices, every returned score of the tensor-core path recomputed in fp64, no true member missing.
    A corrupted tile shows up as a handful of wrong entries in one 128-row tile."""
    monkeypatch.setenv("WB_GEMM_FORCE", "1")
    n, d, nq, k = 40000, 768, 300, 100  # 300 queries: 128-query TS blocks in epoch 0, 256-query SS blocks afterwards
    xb = O.unit_gaussian(n, d, 100)
    xq = O.unit_gaussian(nq, d, 200)
    Dr, Ir = O.flat_search(xb, xq, k)
    for trial in range(6):
        idx = faiss.IndexFlatIP(d)
        idx.add(xb)
        D, I = idx.search(xq, k)
        true = np.einsum('qkd,qd->qk', xb[I].astype(np.float64), xq.astype(np.float64))
        assert np.abs(true - D).max() <= 1e-5, (trial, float(np.abs(true - D).max()))
        O.compare_topk(D, I, Dr, Ir, band=4e-6)


def test_winners_clustered_in_a_few_partial_lists(faiss):
    """The merge of sorted partial lists (scan tail, K3, exchange) looks at the best few entries of every list first and
    then follows the lists that reach deeper (binary search for the prefix that beats the threshold).  Here ALL winners
    sit in 100 consecutive rows (3-4 row groups, i.e. 3-4 of the 148 per-CTA lists), and in one part of an 8-part merge."""
    import torch
    from wise_b200 import _capi
    L = _capi.lib()
    n, d, k = 200000, 64, 100
    xb = O.unit_gaussian(n, d, 5)
    q = O.unit_gaussian(2, d, 6)
    rng = np.random.default_rng(7)
    for j, start in enumerate((5000, 150000)):  # rows close to query j: all of its top-100
        near = q[j] + 0.05 * rng.standard_normal((100, d)).astype(np.float32)
        xb[start:start + 100] = near / np.linalg.norm(near, axis=1, keepdims=True)
    ids = np.arange(n, dtype=np.int64) + 11
    idx = _flat(faiss, xb, ids)
    D, I = idx.search(q, k)
    Dr, Ir = O.flat_search(xb, q, k, ids)
    O.compare_topk(D, I, Dr, Ir)
    assert set(I[0]) == set(range(5011, 5111)) and set(I[1]) == set(range(150011, 150111))
    D1, I1 = idx.search(q[:1], k)  # batch 1: the fused tail of the one-query scan
    O.compare_topk(D1, I1, Dr[:1], Ir[:1])
    # 8 sorted parts, the whole answer in part 3 (the shape of the multi-GPU exchange when one shard holds the winners)
    G, nq = 8, 3
    Dp = torch.full((G, nq, k), 0.0, device="cuda")
    Ip = torch.zeros((G, nq, k), dtype=torch.int64, device="cuda")
    for g in range(G):
        base = 0.9 if g == 3 else 0.5
        Dp[g] = (base - 1e-3 * torch.arange(k, device="cuda").float() - 1e-5 * g).expand(nq, k)
        Ip[g] = (g * 1000 + torch.arange(k, device="cuda")).expand(nq, k)
    Dm = torch.empty(nq, k, device="cuda"); Im = torch.empty(nq, k, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    _capi.check(L.wb_merge_topk_dev(0, nq, k, G, Dp.data_ptr(), Ip.data_ptr(), Dm.data_ptr(), Im.data_ptr(), st))
    torch.cuda.synchronize()
    assert torch.equal(Im, Ip[3]) and torch.equal(Dm, Dp[3])
    # two parts that interleave entry by entry
    Dp2 = torch.stack([(0.9 - 2e-3 * torch.arange(k, device="cuda").float()).expand(nq, k),
                       (0.899 - 2e-3 * torch.arange(k, device="cuda").float()).expand(nq, k)]).contiguous()
    Ip2 = torch.stack([(torch.arange(k, device="cuda") * 2).expand(nq, k), (torch.arange(k, device="cuda") * 2 + 1).expand(nq, k)]).contiguous()
    _capi.check(L.wb_merge_topk_dev(0, nq, k, 2, Dp2.data_ptr(), Ip2.data_ptr(), Dm.data_ptr(), Im.data_ptr(), st))
    torch.cuda.synchronize()
    assert Im[0].tolist() == list(range(k))


@pytest.mark.parametrize("nq,force_gemm", [(1, False), (3, False), (16, True), (300, True)])
def test_nan_rows_and_nan_queries_are_never_returned(faiss, monkeypatch, nq, force_gemm):
    """faiss keeps its result heaps with `if (C::cmp(simi[0], ip))` (a float '<'): a NaN score is never inserted, so
    rows with a NaN component never surface and a NaN query returns the empty (-FLT_MAX, -1) row.  The kernels compare
    floats against their thresholds before they build a key, which gives the same rule; a store with a few corrupt
    embeddings must answer every other query exactly as if those rows were absent."""
    if force_gemm:
        monkeypatch.setenv("WB_GEMM_FORCE", "1")
    n, d, k = 60000, 256, 50
    xb = O.unit_gaussian(n, d, 77)
    xq = O.unit_gaussian(nq, d, 78)
    bad = np.array([0, 17, 4096, 31337, n - 1])
    xb[bad, 5] = np.nan
    xb[bad[1], :] = np.nan
    ids = np.arange(n, dtype=np.int64) * 3 + 1
    idx = _flat(faiss, xb, ids)
    clean = np.setdiff1d(np.arange(n), bad)
    Dr, Ir = O.flat_search(xb[clean], xq, k, ids[clean])
    D, I = idx.search(xq, k)
    if force_gemm:
        epochs, fallbacks = _gemm_stats(idx)
        assert epochs >= 1 and fallbacks == 0
    O.compare_topk(D, I, Dr, Ir, band=4e-6 if force_gemm else O.NEAR_TIE_BAND)
    assert not np.isin(I, ids[bad]).any() and np.isfinite(D).all()
    xq2 = xq.copy()
    xq2[0, 3] = np.nan  # a NaN query: every score is NaN, nothing is ever inserted
    D2, I2 = idx.search(xq2, k)
    assert np.all(I2[0] == -1) and np.all(D2[0] == O.NEG_FLT_MAX)
    if nq > 1:
        O.compare_topk(D2[1:], I2[1:], Dr[1:], Ir[1:], band=4e-6 if force_gemm else O.NEAR_TIE_BAND)
