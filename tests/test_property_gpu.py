"""Property tests (hypothesis) of the search contract on random small shapes: ragged dimensions (d % 4 != 0),
k > ntotal (-1 padding), exact duplicates (lowest position first), arbitrary int64 ids - flat, tensor-core and IVF paths."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def faiss():
    from wise_b200 import faiss_compat
    return faiss_compat


def _data(seed, n, d, nq, dup):
    rng = np.random.default_rng(seed)
    xb = rng.standard_normal((n, d)).astype(np.float32)
    xb /= np.maximum(np.linalg.norm(xb, axis=1, keepdims=True), 1e-12)
    if dup and n >= 4:
        xb[n // 2:n // 2 + n // 4] = xb[:n // 4]
    xq = rng.standard_normal((nq, d)).astype(np.float32)
    xq /= np.maximum(np.linalg.norm(xq, axis=1, keepdims=True), 1e-12)
    if dup and n >= 1:
        xq[0] = xb[0]
    ids = rng.permutation(10 * n + 7)[:n].astype(np.int64) - 3  # arbitrary, unsorted, may be negative
    ids[ids == -1] = 10 * n + 100  # -1 is the padding value
    return xb, xq, ids


@settings(max_examples=30, deadline=None, suppress_health_check=list(HealthCheck))
@given(seed=st.integers(0, 2 ** 31), n=st.integers(1, 3000), d=st.integers(1, 200), nq=st.integers(1, 24),
       k=st.integers(1, 300), dup=st.booleans())
def test_flat_search_contract(faiss, seed, n, d, nq, k, dup):
    xb, xq, ids = _data(seed, n, d, nq, dup)
    idx = faiss.IndexIDMap(faiss.IndexFlatIP(d))
    idx.add_with_ids(xb, ids)
    D, I = idx.search(xq, k)
    Dr, Ir = O.flat_search(xb, xq, k, ids)
    O.compare_topk(D, I, Dr, Ir, band=4e-6)
    if k > n:
        assert np.all(I[:, n:] == -1) and np.all(D[:, n:] == O.NEG_FLT_MAX)
    if dup and n >= 4 and d >= 8:  # (tiny d makes unrelated rows collide too)
        assert I[0, 0] == ids[0]  # the query equals row 0 and its duplicate at n//2: lowest position first
        if k >= 2:
            assert I[0, 1] == ids[n // 2] and D[0, 0] == D[0, 1]


@settings(max_examples=12, deadline=None, suppress_health_check=list(HealthCheck))
@given(seed=st.integers(0, 2 ** 31), n=st.integers(200, 4000), d=st.integers(2, 96), nlist=st.integers(1, 40),
       nq=st.integers(1, 12), k=st.integers(1, 64), nprobe=st.integers(1, 64))
def test_ivf_search_contract(faiss, seed, n, d, nlist, nq, k, nprobe):
    xb, xq, ids = _data(seed, n, d, nq, False)
    cent = O.kmeans_init(xb, nlist, seed=seed % 1000)
    idx = faiss.IndexIVFFlat(faiss.IndexFlatIP(d), d, nlist, faiss.METRIC_INNER_PRODUCT)
    idx.set_centroids(cent)
    idx.add_with_ids(xb, ids)
    _, _, a = idx._export(0, n, want_assign=True)  # GPU assignment (argmax ties / fp32 noise can differ from fp64)
    idx.nprobe = nprobe
    D, I = idx.search(xq, k)
    Dr, Ir = O.ivf_search(xb, ids, a.astype(np.int64), cent, xq, k, nprobe)
    O.compare_topk(D, I, Dr, Ir, band=4e-6)
