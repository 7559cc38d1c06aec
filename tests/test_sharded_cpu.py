"""World-size-2 gloo test (CPU) of the multi-GPU host logic in wise_b200/sharded.py: contiguous
shard ranges, replicated queries, the all-gather exchange and the (rank, local order) tie rule.
The CUDA kernels are replaced by oracle-backed test doubles here (the product defaults are CUDA-only);
the GPU version of the same property is tests/test_flat_gpu.py::test_sharded_result_is_bit_identical_to_single."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as O
from wise_b200.sharded import ShardedIndex, shard_range


class _CpuShard:
    """Test double with the attributes ShardedIndex reads from a faiss_compat index."""

    def __init__(self, xb, ids):
        self.xb, self.ids, self.d = xb, ids, xb.shape[1]

    @property
    def ntotal(self):
        return self.xb.shape[0]


def _oracle_local_search(index, q, k, nprobe):
    D, I = O.flat_search(index.xb, q.numpy(), k, index.ids)
    return torch.from_numpy(D), torch.from_numpy(I)


def _oracle_merge(Dp, Ip):
    """(score desc, part asc, slot asc) - the rule wb_merge_topk_dev implements."""
    parts, nq, k = Dp.shape
    D = np.full((nq, k), O.NEG_FLT_MAX, np.float32)
    I = np.full((nq, k), -1, np.int64)
    for q in range(nq):
        d = Dp[:, q, :].reshape(-1).numpy()
        i = Ip[:, q, :].reshape(-1).numpy()
        valid = np.nonzero(i >= 0)[0]
        order = valid[np.lexsort((valid, -d[valid].astype(np.float64)))][:k]
        D[q, :order.size] = d[order]
        I[q, :order.size] = i[order]
    return torch.from_numpy(D), torch.from_numpy(I)


def _worker(rank, world, port, n, d, nq, k, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        xb = O.unit_gaussian(n, d, 77)
        xb[n - 50:] = xb[:50]  # duplicates living on DIFFERENT ranks: exact ties across shards
        xq = np.concatenate([O.unit_gaussian(nq - 1, d, 78), xb[3:4]])
        ids = np.arange(n, dtype=np.int64) * 2 + 5
        lo, hi = shard_range(n, rank, world)
        sh = ShardedIndex(_CpuShard(xb[lo:hi], ids[lo:hi]), local_search=_oracle_local_search, merge=_oracle_merge)
        assert sh.ntotal == n and sh.world == world and sh.rank == rank
        D, I = sh.search(xq, k)
        np.savez(os.path.join(out, f"r{rank}.npz"), D=D, I=I)
    finally:
        dist.destroy_process_group()


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 10_000_000):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, i, w) for i in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


@pytest.mark.timeout(300)
def test_two_rank_gloo_equals_single_index(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    n, d, nq, k = 4000, 32, 5, 20
    mp.spawn(_worker, args=(2, port, n, d, nq, k, str(tmp_path)), nprocs=2, join=True)
    xb = O.unit_gaussian(n, d, 77)
    xb[n - 50:] = xb[:50]
    xq = np.concatenate([O.unit_gaussian(nq - 1, d, 78), xb[3:4]])
    ids = np.arange(n, dtype=np.int64) * 2 + 5
    Dr, Ir = O.flat_search(xb, xq, k, ids)
    r0, r1 = (np.load(tmp_path / f"r{r}.npz") for r in (0, 1))
    assert np.array_equal(r0["I"], r1["I"]) and np.array_equal(r0["D"], r1["D"])  # every rank gets the answer
    assert np.array_equal(r0["I"], Ir) and np.array_equal(r0["D"], Dr)  # == single logical index, ties included
    assert Ir[-1, 0] == ids[3] and Ir[-1, 1] == ids[n - 50 + 3]  # cross-rank exact tie: lowest position first
