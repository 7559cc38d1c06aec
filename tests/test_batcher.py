"""CPU tests of the request micro-batcher (host logic; the index is a test double with faiss-style search)."""
import asyncio
import threading

import numpy as np
import pytest

from oracle import oracle as O
from wise_b200.batcher import SearchBatcher


class _Index:
    def __init__(self, xb):
        self.xb, self.calls = xb, []

    def search(self, x, k):
        self.calls.append(x.shape[0])
        return O.flat_search(self.xb, x, k)


def test_concurrent_requests_share_one_call_and_get_their_own_rows():
    xb = O.unit_gaussian(2000, 16, 1)
    idx = _Index(xb)
    b = SearchBatcher(idx, max_batch=32, max_wait_ms=200)
    qs = O.unit_gaussian(12, 16, 2)
    ks = [5 if i % 2 else 9 for i in range(12)]
    futs = [b.submit(qs[i:i + 1], ks[i]) for i in range(12)]
    res = [f.result(timeout=30) for f in futs]
    b.close()
    assert sum(idx.calls) == 12 and len(idx.calls) <= 2  # coalesced
    for i, (D, I) in enumerate(res):
        Dr, Ir = O.flat_search(xb, qs[i:i + 1], ks[i])
        assert D.shape == (1, ks[i]) and np.array_equal(I, Ir) and np.array_equal(D, Dr)


def test_async_api_and_max_batch():
    xb = O.unit_gaussian(500, 8, 3)
    idx = _Index(xb)
    b = SearchBatcher(idx, max_batch=4, max_wait_ms=50)
    qs = O.unit_gaussian(10, 8, 4)

    async def main():
        return await asyncio.gather(*[b.search(qs[i:i + 1], 3) for i in range(10)])

    out = asyncio.run(main())
    b.close()
    assert max(idx.calls) <= 4 and sum(idx.calls) == 10
    for i, (D, I) in enumerate(out):
        assert np.array_equal(I, O.flat_search(xb, qs[i:i + 1], 3)[1])


def test_errors_reach_every_waiter():
    class Bad:
        def search(self, x, k):
            raise RuntimeError("boom")
    b = SearchBatcher(Bad(), max_batch=8, max_wait_ms=20)
    futs = [b.submit(np.zeros((1, 4), np.float32), 2) for _ in range(3)]
    for f in futs:
        with pytest.raises(RuntimeError):
            f.result(timeout=10)
    b.close()
    with pytest.raises(RuntimeError):
        b.submit(np.zeros((1, 4), np.float32), 2)
