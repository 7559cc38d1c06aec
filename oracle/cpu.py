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
