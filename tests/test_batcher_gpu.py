"""SearchBatcher on a real index: 64 concurrent single-query requests - what the patched REST handler
issues (integration/wise_b200.patch, /root/reference/api/routes.py:1407) - are coalesced into index.search(n > 1)
calls that the tensor-core kernel (K2) serves, and every request gets exactly what its own n = 1 search returns."""
import asyncio
import ctypes as C
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _gemm_stats(idx):
    from wise_b200 import _capi
    a, b = C.c_int64(), C.c_int64()
    _capi.lib().wb_gemm_stats(idx._h, C.byref(a), C.byref(b))
    return a.value, b.value


def test_concurrent_requests_reach_the_tensor_core_path():
    from wise_b200 import faiss_compat as faiss
    from wise_b200.batcher import SearchBatcher
    n, d, k, nreq = 120000, 512, 20, 64
    xb = O.clustered_unit(n, d, 200, 41)
    ids = np.arange(n, dtype=np.int64) + 1
    index = faiss.IndexIDMap(faiss.IndexFlatIP(d))
    index.add_with_ids(xb, ids)
    xq = O.clustered_unit(nreq, d, 200, 42)
    single = [index.search(xq[i:i + 1], k) for i in range(nreq)]  # n = 1: the bandwidth-bound scan (K1)
    e0, f0 = _gemm_stats(index)
    assert e0 == 0
    b = SearchBatcher(index, max_batch=64, max_wait_ms=50.0)
    try:
        with ThreadPoolExecutor(max_workers=nreq) as pool:  # 64 clients at once, like 64 in-flight HTTP requests
            futs = [pool.submit(b.search_blocking, xq[i:i + 1], k) for i in range(nreq)]
            got = [f.result(timeout=60) for f in futs]

        async def client(i):
            return await b.search(xq[i:i + 1], k)

        async def many():
            return await asyncio.gather(*[client(i) for i in range(nreq)])

        got_async = asyncio.run(many())
    finally:
        b.close()
    e1, f1 = _gemm_stats(index)
    assert b.requests == 2 * nreq and b.batches < nreq // 4, (b.requests, b.batches)
    assert e1 > e0 and f1 == f0, "coalesced batches must have run on the tensor-core kernel without fallback"
    for res in (got, got_async):
        for i in range(nreq):
            D, I = res[i]
            assert D.shape == (1, k) and I.shape == (1, k)
            O.compare_topk(D, I, single[i][0], single[i][1], band=4e-6)
    Dr, Ir = O.flat_search(xb, xq, k, ids)
    O.compare_topk(np.concatenate([r[0] for r in got]), np.concatenate([r[1] for r in got]), Dr, Ir, band=4e-6)
