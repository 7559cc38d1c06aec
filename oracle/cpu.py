"""TEST INFRASTRUCTURE ONLY - ctypes loader for oracle/liboracle_cpu.so (cpu_flat.c).
PARITY UNPINNED (see oracle/__init__.py)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build() -> str:
    so = os.path.join(_HERE, "liboracle_cpu.so")
    src = os.path.join(_HERE, "cpu_flat.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle_cpu.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int64)
        for name in ("orc_flat_search_seq", "orc_flat_search_mt"):
            f = getattr(L, name)
            f.restype = C.c_int
            f.argtypes = [fp, C.c_int64, C.c_int64, ip, fp, C.c_int64, C.c_int64, fp, ip]
        L.orc_ivf_search.restype = C.c_int
        L.orc_ivf_search.argtypes = [fp, C.c_int64, ip, ip, ip, fp, C.c_int64, fp, C.c_int64, C.c_int64,
                                     C.c_int64, fp, ip]
        L.orc_simd_name.restype = C.c_char_p
        L.orc_max_threads.restype = C.c_int
        L.orc_set_threads.restype = None
        L.orc_set_threads.argtypes = [C.c_int]
        _LIB = L
    return _LIB


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _i(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int64))


def flat_search(xb, xq, k, ids=None, mt=False):
    """faiss-equivalent IndexFlatIP(+IDMap) search on host cores; mt=False is faiss's own
    threading (parallel over queries only), mt=True splits the database over all threads."""
    xb = np.ascontiguousarray(xb, np.float32)
    xq = np.ascontiguousarray(xq, np.float32)
    ids = None if ids is None else np.ascontiguousarray(ids, np.int64)
    n, d = xq.shape
    D = np.empty((n, k), np.float32)
    I = np.empty((n, k), np.int64)
    fn = lib().orc_flat_search_mt if mt else lib().orc_flat_search_seq
    rc = fn(_f(xb), xb.shape[0], d, _i(ids), _f(xq), n, k, _f(D), _i(I))
    assert rc == 0
    return D, I


def ivf_search(xb, ids, list_off, perm, centroids, xq, k, nprobe):
    xb = np.ascontiguousarray(xb, np.float32)
    xq = np.ascontiguousarray(xq, np.float32)
    centroids = np.ascontiguousarray(centroids, np.float32)
    list_off = np.ascontiguousarray(list_off, np.int64)
    perm = np.ascontiguousarray(perm, np.int64)
    ids = None if ids is None else np.ascontiguousarray(ids, np.int64)
    n, d = xq.shape
    D = np.empty((n, k), np.float32)
    I = np.empty((n, k), np.int64)
    rc = lib().orc_ivf_search(_f(xb), d, _i(ids), _i(list_off), _i(perm), _f(centroids), centroids.shape[0],
                              _f(xq), n, k, nprobe, _f(D), _i(I))
    assert rc == 0
    return D, I


def simd_name() -> str:
    return lib().orc_simd_name().decode()


def max_threads() -> int:
    return int(lib().orc_max_threads())


def set_threads(n: int) -> None:
    """OpenMP thread count of the C oracle (overrides an inherited OMP_NUM_THREADS)."""
    lib().orc_set_threads(int(n))


def blas_flat_search(xb, xq, k, ids=None, tile=1024, threads=None):
    """faiss exhaustive_inner_product_blas restated [faiss-upstream]: sgemm over (all queries) x (1024 database rows)
    tiles (numpy -> OpenBLAS), each tile's scores folded into the running top-k (argpartition = the reservoir
    handler faiss uses for k >= 100).  The n >= 20 branch of knn_inner_product, reached from
    /root/reference/src/index/feature_search_index.py:113.  Order inside exact ties is NOT the oracle's rule (timing leg)."""
    from threadpoolctl import threadpool_limits
    xb = np.ascontiguousarray(xb, np.float32)
    xq = np.ascontiguousarray(xq, np.float32)
    n = xq.shape[0]
    N = xb.shape[0]
    kk = min(k, N)
    with threadpool_limits(limits=threads):
        best_s = np.full((n, kk), -np.inf, np.float32)
        best_p = np.full((n, kk), -1, np.int64)
        # faiss merges once per 1024-row tile; numpy's per-call overhead makes that unfair to the CPU, so the
        # sgemm runs over 16 tiles at a time and the partial top-k is taken once per 16K rows
        step = tile * 16
        for j0 in range(0, N, step):
            j1 = min(N, j0 + step)
            s = xq @ xb[j0:j1].T
            m = j1 - j0
            if m > kk:
                part = np.argpartition(-s, kk - 1, axis=1)[:, :kk]
                ps = np.take_along_axis(s, part, axis=1)
            else:
                part = np.broadcast_to(np.arange(m), (n, m))
                ps = s
            cs = np.concatenate([best_s, ps], axis=1)
            cp = np.concatenate([best_p, part + j0], axis=1)
            sel = np.argpartition(-cs, kk - 1, axis=1)[:, :kk]
            best_s = np.take_along_axis(cs, sel, axis=1)
            best_p = np.take_along_axis(cp, sel, axis=1)
        order = np.argsort(-best_s, axis=1, kind="stable")
        D = np.take_along_axis(best_s, order, axis=1)
        P = np.take_along_axis(best_p, order, axis=1)
    I = P if ids is None else np.where(P >= 0, np.asarray(ids)[np.maximum(P, 0)], -1)
    if kk < k:
        D = np.concatenate([D, np.full((n, k - kk), np.float32(-3.4028234663852886e38))], axis=1)
        I = np.concatenate([I, np.full((n, k - kk), -1, np.int64)], axis=1)
    return D, I
