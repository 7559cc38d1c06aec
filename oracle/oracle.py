"""TEST INFRASTRUCTURE ONLY - numpy restatement of the faiss semantics WISE uses.

PARITY UNPINNED (see oracle/__init__.py): faiss itself is absent; every
function below cites the reference call site it serves and the upstream faiss
routine it restates ([faiss-upstream], from the published algorithm).

Conventions:
  * scores are exact inner products: accumulated in float64 from the float32
    inputs, then rounded once to float32;
  * result rows are ordered by (score desc, position asc) where `position` is
    the insertion order of the row in the index (for IndexIDMap the external
    id is looked up *after* ranking, as faiss does: IndexIDMap::search maps
    labels of the wrapped IndexFlat);
  * unfilled slots hold (score=-FLT_MAX, id=-1) - the callers test id == -1
    (/root/reference/search.py:140-143, /root/reference/api/routes.py:1411).
"""
from __future__ import annotations

import numpy as np

NEG_FLT_MAX = np.float32(-3.4028234663852886e38)
# scores closer than this are a "near tie": fp32 accumulation-order noise for
# unit vectors with d <= 1024.
NEAR_TIE_BAND = 2e-6
SCORE_TOL = 1e-5  # BASELINE.json north_star: "Scores must agree within 1e-5 absolute"


# --------------------------------------------------------------------------- #
# synthetic data
# --------------------------------------------------------------------------- #
def unit_gaussian(n: int, d: int, seed: int) -> np.ndarray:
    """i.i.d. N(0,1) rows, L2-normalised (config C1)."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d), dtype=np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return np.ascontiguousarray(x, dtype=np.float32)


def clustered_unit(n: int, d: int, ncentres: int, seed: int, noise: float = 0.6) -> np.ndarray:
    """'CLIP-like' clustered rows: normalise(c_j + noise*g), g ~ N(0, I/d) (configs C2-C5)."""
    rng = np.random.default_rng(seed)
    centres = rng.standard_normal((ncentres, d), dtype=np.float32)
    centres /= np.linalg.norm(centres, axis=1, keepdims=True)
    j = rng.integers(0, ncentres, size=n)
    g = rng.standard_normal((n, d), dtype=np.float32) / np.float32(np.sqrt(d))
    x = centres[j] + np.float32(noise) * g
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return np.ascontiguousarray(x, dtype=np.float32)


def combined_query(pos_text, pos_image=None, neg=None, text_w=2.0, neg_w=0.2) -> np.ndarray:
    """Mixed query vector of /root/reference/api/routes.py:759-850: weighted mean of
    (+/-) unit vectors (text x2.0, negatives x0.2, /root/reference/config.py:13-14), renormalised."""
    vecs, w = [np.asarray(pos_text, np.float32)], [text_w]
    if pos_image is not None:
        vecs.append(np.asarray(pos_image, np.float32)); w.append(1.0)
    if neg is not None:
        vecs.append(-np.asarray(neg, np.float32)); w.append(neg_w)
    avg = np.average(np.stack(vecs), axis=0, weights=np.asarray(w, np.float32))
    avg = avg / np.linalg.norm(avg, axis=-1, keepdims=True)
    return avg.astype(np.float32)


# --------------------------------------------------------------------------- #
# IndexFlatIP / IndexIDMap search
# --------------------------------------------------------------------------- #
def _scores_f64(xq: np.ndarray, xb: np.ndarray) -> np.ndarray:
    return (xq.astype(np.float64) @ xb.astype(np.float64).T).astype(np.float32)


def _topk_rows(scores: np.ndarray, pos: np.ndarray, k: int):
    """top-k of each row by (score desc, pos asc); returns (D[n,k'], P[n,k']) k'=min(k,m)."""
    n, m = scores.shape
    kk = min(k, m)
    D = np.empty((n, kk), np.float32)
    P = np.empty((n, kk), np.int64)
    for i in range(n):
        s = scores[i]
        if kk < m:
            # keep everything tied with the kk-th score so the tie rule decides
            kth = np.partition(s, m - kk)[m - kk]
            cand = np.nonzero(s >= kth)[0]
        else:
            cand = np.arange(m)
        order = np.lexsort((pos[cand], -s[cand].astype(np.float64)))[:kk]
        D[i] = s[cand][order]
        P[i] = pos[cand][order]
    return D, P


def flat_search(xb: np.ndarray, xq: np.ndarray, k: int, ids: np.ndarray | None = None,
                chunk: int = 65536):
    """IndexFlatIP(+IndexIDMap).search  [faiss-upstream: knn_inner_product ->
    exhaustive_inner_product_seq/_blas + heap_reorder; IndexIDMap::search].
    Call sites: /root/reference/src/index/feature_search_index.py:113,
    /root/reference/api/routes.py:1407.
    Returns (D float32[n,k] desc, I int64[n,k]); -FLT_MAX/-1 padded when k > ntotal."""
    xb = np.ascontiguousarray(xb, np.float32)
    xq = np.ascontiguousarray(xq, np.float32)
    n, N = xq.shape[0], xb.shape[0]
    D = np.full((n, k), NEG_FLT_MAX, np.float32)
    I = np.full((n, k), -1, np.int64)
    if N == 0 or k == 0:
        return D, I
    bestD = np.empty((n, 0), np.float32)
    bestP = np.empty((n, 0), np.int64)
    for s in range(0, N, chunk):
        e = min(N, s + chunk)
        sc = _scores_f64(xq, xb[s:e])
        cD, cP = _topk_rows(sc, np.arange(s, e, dtype=np.int64), k)
        allD = np.concatenate([bestD, cD], axis=1)
        allP = np.concatenate([bestP, cP], axis=1)
        kk = min(k, allD.shape[1])
        nD = np.empty((n, kk), np.float32)
        nP = np.empty((n, kk), np.int64)
        for i in range(n):
            order = np.lexsort((allP[i], -allD[i].astype(np.float64)))[:kk]
            nD[i], nP[i] = allD[i][order], allP[i][order]
        bestD, bestP = nD, nP
    kk = bestD.shape[1]
    D[:, :kk] = bestD
    I[:, :kk] = bestP if ids is None else np.asarray(ids, np.int64)[bestP]
    return D, I


# --------------------------------------------------------------------------- #
# IndexIVFFlat (METRIC_INNER_PRODUCT) : add / search
# --------------------------------------------------------------------------- #
def ivf_assign(x: np.ndarray, centroids: np.ndarray, chunk: int = 16384) -> np.ndarray:
    """Coarse assignment at add time [faiss-upstream: IndexIVF::add_core ->
    quantizer->assign = IndexFlatIP.search(k=1)]: argmax inner product, ties ->
    lowest centroid index.  Call site: /root/reference/src/index/feature_search_index.py:81."""
    out = np.empty(x.shape[0], np.int64)
    for s in range(0, x.shape[0], chunk):
        sc = _scores_f64(x[s:s + chunk], centroids)
        out[s:s + chunk] = np.argmax(sc, axis=1)  # first max = lowest index
    return out


def ivf_coarse(xq: np.ndarray, centroids: np.ndarray, nprobe: int):
    """quantizer->search(n, x, nprobe): the nprobe best centroids by IP (desc, ties lowest index)."""
    return flat_search(centroids, xq, min(nprobe, centroids.shape[0]))


def ivf_search(xb: np.ndarray, ids: np.ndarray | None, assign: np.ndarray, centroids: np.ndarray,
               xq: np.ndarray, k: int, nprobe: int):
    """IndexIVFFlat.search [faiss-upstream: IndexIVF::search -> search_preassigned ->
    IVFFlatScanner::scan_codes + heap]: candidates = every row whose list is among the
    min(nprobe,nlist) best centroids; same ranking rule as flat_search (ties -> lowest
    insertion position).  Call sites as flat_search; nprobe from
    /root/reference/api/routes.py:899-902 (default 1 when never set, search.py)."""
    n = xq.shape[0]
    D = np.full((n, k), NEG_FLT_MAX, np.float32)
    I = np.full((n, k), -1, np.int64)
    _, probes = ivf_coarse(xq, centroids, nprobe)
    for i in range(n):
        pl = probes[i][probes[i] >= 0]
        rows = np.nonzero(np.isin(assign, pl))[0]
        if rows.size == 0:
            continue
        sc = _scores_f64(xq[i:i + 1], xb[rows])
        d1, p1 = _topk_rows(sc, rows.astype(np.int64), k)
        kk = d1.shape[1]
        D[i, :kk] = d1[0]
        I[i, :kk] = p1[0] if ids is None else np.asarray(ids, np.int64)[p1[0]]
    return D, I


# --------------------------------------------------------------------------- #
# k-means (faiss Clustering; spherical=False is what a plain-constructor IndexIVFFlat uses) - IndexIVFFlat.train
# --------------------------------------------------------------------------- #
class FaissRandom:
    """[faiss-upstream: utils/random.cpp RandomGenerator] std::mt19937(seed);
    rand_int(max) = mt() % max ; rand_float() = mt() / float(mt.max())."""

    def __init__(self, seed: int):
        self.bg = np.random.MT19937()
        self.bg._legacy_seeding(int(seed) & 0xFFFFFFFF)

    def raw(self) -> int:
        return int(self.bg.random_raw())

    def rand_int(self, mx: int) -> int:
        return self.raw() % mx

    def rand_float(self) -> float:
        return float(np.float32(self.raw()) / np.float32(4294967295.0))


def rand_perm(n: int, seed: int) -> np.ndarray:
    """[faiss-upstream: rand_perm] Fisher-Yates with RandomGenerator(seed)."""
    rng = FaissRandom(seed)
    perm = np.arange(n, dtype=np.int64)
    for i in range(n - 1):
        i2 = i + rng.rand_int(n - i)
        perm[i], perm[i2] = perm[i2], perm[i]
    return perm


def kmeans_init(x: np.ndarray, k: int, seed: int = 1234, spherical: bool = False) -> np.ndarray:
    """[faiss-upstream: Clustering::train] initial centroids = first k rows of
    rand_perm(nx, seed + 1) (redo 0); spherical => L2-renormalised (no-op for unit rows)."""
    perm = rand_perm(x.shape[0], seed + 1)
    c = x[perm[:k]].astype(np.float32).copy()
    return _renorm(c) if spherical else c


def _renorm(c: np.ndarray) -> np.ndarray:
    nrm = np.sqrt((c.astype(np.float64) ** 2).sum(axis=1))
    nrm[nrm == 0] = 1.0
    return (c / nrm[:, None]).astype(np.float32)


def kmeans_iteration(x: np.ndarray, centroids: np.ndarray, rng: FaissRandom | None = None, spherical: bool = False):
    """One Clustering iteration [faiss-upstream]: assign by max-IP (IndexFlatIP.search k=1),
    compute_centroids (mean of members; empty clusters keep their old centroid),
    split_clusters (eps=1/1024), post_process_centroids (L2 renorm only when cp.spherical; the IndexIVFFlat the
    reference builds with the plain constructor has spherical=False, index_factory is what turns it on for IP).
    Reached from index.train(), /root/reference/src/index/feature_search_index.py:75.
    Returns (new_centroids, assign, objective=sum of max-IP, nsplit)."""
    k, d = centroids.shape
    n = x.shape[0]
    assign = ivf_assign(x, centroids)
    best = np.einsum('ij,ij->i', x.astype(np.float64), centroids[assign].astype(np.float64))
    obj = float(best.sum())
    sums = np.zeros((k, d), np.float64)
    np.add.at(sums, assign, x.astype(np.float64))
    cnt = np.bincount(assign, minlength=k).astype(np.float64)
    newc = centroids.astype(np.float64).copy()
    nz = cnt > 0
    newc[nz] = sums[nz] / cnt[nz, None]
    newc = newc.astype(np.float32)
    # split_clusters
    EPS = np.float32(1.0 / 1024.0)
    nsplit = 0
    if rng is None:
        rng = FaissRandom(1234)
    hass = cnt.copy()
    for ci in range(k):
        if hass[ci] == 0:
            cj = 0
            while True:
                p = (hass[cj] - 1.0) / float(n - k)
                r = rng.rand_float()
                if r < p:
                    break
                cj = (cj + 1) % k
            newc[ci] = newc[cj]
            sign = np.where(np.arange(d) % 2 == 0, 1.0, -1.0).astype(np.float32)
            newc[ci] = newc[ci] * (1 + sign * EPS)
            newc[cj] = newc[cj] * (1 - sign * EPS)
            hass[ci] = hass[cj] // 2
            hass[cj] -= hass[ci]
            nsplit += 1
    return (_renorm(newc) if spherical else newc.astype(np.float32)), assign, obj, nsplit


def kmeans_train(x: np.ndarray, k: int, niter: int = 10, seed: int = 1234, spherical: bool = False):
    """IndexIVFFlat.train -> Level1Quantizer::train_q1 -> Clustering::train
    (niter=10, nredo=1, cp.spherical False unless the caller sets it) [faiss-upstream]."""
    c = kmeans_init(x, k, seed, spherical)
    rng = FaissRandom(seed)
    objs = []
    for _ in range(niter):
        c, _, obj, _ = kmeans_iteration(x, c, rng, spherical)
        objs.append(obj)
    return c, objs


def ivf_train_params(feature_count: int):
    """nlist / train_count rule of /root/reference/src/index/feature_search_index.py:55-59."""
    import math
    if feature_count < 200000:
        cells = 3 * round(math.sqrt(feature_count))
    else:
        cells = 10 * round(math.sqrt(feature_count))
    return cells, min(feature_count, 100 * cells)


# --------------------------------------------------------------------------- #
# comparison harness
# --------------------------------------------------------------------------- #
def compare_topk(D, I, D_ref, I_ref, score_tol: float = SCORE_TOL, band: float = NEAR_TIE_BAND):
    """Parity rule of BASELINE.json: scores within `score_tol` absolute, id lists
    identical except inside near-tie bands.  Formally, for each query row:
      (1) |D[j] - D_ref[j]| <= score_tol for every rank j (padding matches exactly);
      (2) at every rank j where I[j] != I_ref[j], the candidate returned has a
          reference-rank score within `band` of D_ref[j]'s neighbourhood: i.e. both ids
          belong to a run of reference scores that differ pairwise by < band, or the
          id sits across the k-th boundary with |score - D_ref[k-1]| < band.
    Returns a dict with counts; raises AssertionError on a violation."""
    D, I, D_ref, I_ref = map(np.asarray, (D, I, D_ref, I_ref))
    assert D.shape == D_ref.shape and I.shape == I_ref.shape, (D.shape, D_ref.shape)
    n, k = I.shape
    exact_rows = 0
    band_swaps = 0
    for q in range(n):
        pad_ref = I_ref[q] == -1
        assert np.array_equal(I[q] == -1, pad_ref), f"row {q}: -1 padding differs"
        v = ~pad_ref
        assert np.all(D[q][pad_ref] == NEG_FLT_MAX), f"row {q}: padded scores must be -FLT_MAX"
        err = np.abs(D[q][v].astype(np.float64) - D_ref[q][v].astype(np.float64))
        assert err.size == 0 or err.max() <= score_tol, f"row {q}: score err {err.max():.3e} > {score_tol}"
        assert np.all(np.diff(D[q][v].astype(np.float64)) <= 0), f"row {q}: scores not descending"
        if np.array_equal(I[q], I_ref[q]):
            exact_rows += 1
            continue
        ref_pos = {int(i): j for j, i in enumerate(I_ref[q][v])}
        kv = int(v.sum())
        kth = float(D_ref[q, kv - 1])
        ours = set(int(i) for i in I[q][v])
        for j in np.nonzero(I[q] != I_ref[q])[0]:
            gid = int(I[q, j])
            if gid in ref_pos:
                # same member, different rank: the two reference scores must be a near tie
                jr = ref_pos[gid]
                assert abs(float(D_ref[q, jr]) - float(D_ref[q, j])) < band, \
                    f"row {q} rank {j}: id {gid} is at ref rank {jr}, scores differ by >= band"
            else:
                # intruder across the k-th boundary: its score must be a near tie with the ref k-th score
                assert abs(float(D[q, j]) - kth) < 2 * band, \
                    f"row {q} rank {j}: id {gid} not in reference top-k and not a boundary near-tie"
            band_swaps += 1
        for gid, jr in ref_pos.items():
            if gid not in ours:  # displaced reference member: must itself sit in the boundary band
                assert abs(float(D_ref[q, jr]) - kth) < band, \
                    f"row {q}: reference id {gid} (rank {jr}) missing and not a boundary near-tie"
    return {"rows": n, "exact_rows": exact_rows, "band_swaps": band_swaps}
