"""A numpy model of K2's filter-and-refine search checked against the oracle.

It restates, in fp64/fp32 numpy, what the CUDA path does per query: geometrically growing epochs; a first epoch
selected on full-precision scores; later epochs that only see the ONE-TERM score (both operands truncated to
tf32) and keep a row when it exceeds `thr - margin`; exact re-scoring of the candidates; compaction into the
top-k with the (score desc, position asc) key.  The model has no GPU in it: it guards the ALGORITHM (margin,
threshold updates, tie rule) so that kernel work cannot silently change what the search returns."""
import numpy as np
import pytest

from oracle import oracle as O


def _tf32(a):
    return (np.ascontiguousarray(a, np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def k2_model_search(xb, xq, k, cap=4096, growth=None, margin_scale=1.0):
    n, d = xb.shape
    growth = growth or min(12.0, max(0.25, cap / (4.0 * k)))
    c_margin = margin_scale * (2.0 ** -9 + d * 2.0 ** -23 * 1.1 + 1e-5)
    max_norm = float(np.sqrt((xb.astype(np.float64) ** 2).sum(1).max()))
    xh, qh = _tf32(xb).astype(np.float64), _tf32(xq).astype(np.float64)
    D = np.full((len(xq), k), -np.inf, np.float32)
    I = np.full((len(xq), k), -1, np.int64)
    stats = {"candidates": 0, "overflow": False}
    for qi, q in enumerate(xq):
        margin = c_margin * float(np.linalg.norm(q.astype(np.float64))) * max_norm
        best = []  # (score fp32, position) sorted by (-score, position)
        thr = -np.inf
        r0 = 0
        while r0 < n:
            ln = cap if r0 == 0 else int(r0 * growth)
            ln = max(128, (ln + 127) // 128 * 128)
            r1 = min(n, r0 + ln)
            rows = np.arange(r0, r1)
            if r0 == 0:
                cand = rows  # threshold is -inf: everything is a candidate
            else:
                one_term = xh[r0:r1] @ qh[qi]
                cand = rows[one_term > thr - margin]
            stats["candidates"] += len(cand)
            if len(cand) > cap:
                stats["overflow"] = True  # the CUDA path repairs such a batch with the exact scan
            exact = (xb[cand].astype(np.float64) @ q.astype(np.float64)).astype(np.float32)  # the re-scoring pass
            best = sorted(best + list(zip(exact.tolist(), cand.tolist())), key=lambda t: (-t[0], t[1]))[:k]
            if len(best) == k:
                thr = best[-1][0]
            r0 = r1
        for j, (s, p) in enumerate(best):
            D[qi, j], I[qi, j] = s, p
    return D, I, stats


@pytest.mark.parametrize("n,d,nq,k,clustered", [(30000, 64, 6, 10, False), (50000, 128, 4, 100, True), (20000, 32, 5, 50, False)])
def test_model_matches_oracle(n, d, nq, k, clustered):
    xb = O.clustered_unit(n, d, 32, 1) if clustered else O.unit_gaussian(n, d, 1)
    xq = O.clustered_unit(nq, d, 32, 2) if clustered else O.unit_gaussian(nq, d, 2)
    D, I, stats = k2_model_search(xb, xq, k)
    Dr, Ir = O.flat_search(xb, xq, k)
    O.compare_topk(D, I, Dr, Ir)
    assert not stats["overflow"]
    # the filter keeps the candidate volume near k * growth per epoch: far below one candidate per row
    assert stats["candidates"] < 0.35 * n * nq


def test_model_duplicates_and_unnormalised_rows():
    """Rows of very different norms (the margin scales with the largest one) and exact duplicates that live in
    different epochs: the duplicates come back with bit-identical scores, lowest position first."""
    n, d, k = 40000, 48, 20
    rng = np.random.default_rng(9)
    xb = O.unit_gaussian(n, d, 5) * rng.uniform(0.5, 3.0, size=(n, 1)).astype(np.float32)
    xb[100:140] *= (5.0 / np.linalg.norm(xb[100:140], axis=1, keepdims=True)).astype(np.float32)  # the largest rows
    xb[30000:30040] = xb[100:140]  # exact duplicates far apart: first epoch vs a filter epoch
    xq = np.concatenate([O.unit_gaussian(6, d, 6), xb[100:104] / np.linalg.norm(xb[100:104], axis=1, keepdims=True)])
    Dr, Ir = O.flat_search(xb, xq, k)
    D, I, stats = k2_model_search(xb, xq, k)
    O.compare_topk(D, I, Dr, Ir, score_tol=5e-5, band=2e-5)  # scores reach 5: tolerances scale with |q||x|
    assert not stats["overflow"]
    for j in range(4):
        assert I[6 + j, 0] == 100 + j and I[6 + j, 1] == 30000 + j and D[6 + j, 0] == D[6 + j, 1]


def test_model_margin_bound_is_what_the_kernel_uses():
    """The constant in capi.cu (`c_margin`) and the model agree; a search with HALF the margin constant is no longer
    guaranteed, a search with the full constant never drops a row whose exact score beats the threshold."""
    n, d = 20000, 256
    xb = np.abs(O.unit_gaussian(n, d, 3))  # all-positive rows and query: truncation errors add up coherently
    xq = np.abs(O.unit_gaussian(3, d, 4))
    exact = xb.astype(np.float64) @ xq.astype(np.float64).T
    one = _tf32(xb).astype(np.float64) @ _tf32(xq).astype(np.float64).T
    c = 2.0 ** -9 + d * 2.0 ** -23 * 1.1 + 1e-5
    bound = c * np.linalg.norm(xb.astype(np.float64), axis=1)[:, None] * np.linalg.norm(xq.astype(np.float64), axis=1)[None, :]
    assert np.all(exact - one <= bound) and np.all(exact - one >= 0)
    assert (exact - one).max() > 0.2 * bound.min()  # coherent errors come within a small factor of the bound
