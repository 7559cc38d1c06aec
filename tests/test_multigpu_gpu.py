"""2-GPU test (skipped with fewer GPUs): sharded search through the peer-memory exchange kernel and through
the NCCL all-gather path must both equal the single-index oracle result, ties included."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out, mode):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WISE_B200_EXCHANGE=mode,
                      WISE_B200_DEVICE=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from wise_b200 import faiss_compat as faiss
        from wise_b200.sharded import ShardedIndex, shard_range
        n, d = 60000, 256
        xb = O.unit_gaussian(n, d, 77)
        xb[n - 50:] = xb[:50]
        ids = np.arange(n, dtype=np.int64) * 2 + 5
        lo, hi = shard_range(n, rank, world)
        idx = faiss.IndexIDMap(faiss.IndexFlatIP(d, device=rank))
        idx.add_with_ids(xb[lo:hi], ids[lo:hi])
        sh = ShardedIndex(idx)
        assert (sh.exchange is not None) == (mode == "peer")
        res = {}
        for nq, k in ((1, 100), (5, 20), (40, 10), (3, 1000)):
            xq = np.concatenate([O.unit_gaussian(nq - 1, d, 78 + nq), xb[3:4]]) if nq > 1 else xb[3:4].copy()
            for rep in range(3):  # repeated calls exercise the double-buffered mailboxes
                D, I = sh.search(xq, k)
            res[f"D{nq}_{k}"], res[f"I{nq}_{k}"] = D, I
            # device-buffer entry point (wb_exch_search_dev: one launch when the scan kernel serves the batch)
            Dd, Id = sh.search_dev(torch.from_numpy(xq).cuda(rank), k)
            torch.cuda.synchronize()
            assert np.array_equal(Dd.cpu().numpy(), D) and np.array_equal(Id.cpu().numpy(), I)
            if sh.exchange is not None:
                assert not sh.exchange.timed_out()
        np.savez(os.path.join(out, f"{mode}_r{rank}.npz"), **res)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("mode", ["peer", "nccl"])
def test_two_gpu_sharded_equals_single(tmp_path, mode):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path), mode), nprocs=2, join=True)
    n, d = 60000, 256
    xb = O.unit_gaussian(n, d, 77)
    xb[n - 50:] = xb[:50]
    ids = np.arange(n, dtype=np.int64) * 2 + 5
    r0, r1 = (np.load(tmp_path / f"{mode}_r{r}.npz") for r in (0, 1))
    for nq, k in ((1, 100), (5, 20), (40, 10), (3, 1000)):
        xq = np.concatenate([O.unit_gaussian(nq - 1, d, 78 + nq), xb[3:4]]) if nq > 1 else xb[3:4].copy()
        Dr, Ir = O.flat_search(xb, xq, k, ids)
        D0, I0, D1, I1 = r0[f"D{nq}_{k}"], r0[f"I{nq}_{k}"], r1[f"D{nq}_{k}"], r1[f"I{nq}_{k}"]
        assert np.array_equal(I0, I1) and np.array_equal(D0, D1)
        O.compare_topk(D0, I0, Dr, Ir, band=4e-6)
        assert I0[-1, 0] == ids[3] and I0[-1, 1] == ids[n - 50 + 3]  # exact tie across ranks: lowest position first


def _train_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WISE_B200_DEVICE=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from wise_b200 import faiss_compat as faiss
        from wise_b200.sharded import ShardedIndex, shard_range, train_ivf_sharded
        n, d, k = 40000, 64, 100
        x = O.clustered_unit(n, d, 100, 31)
        lo, hi = shard_range(n, rank, world)
        ivf = faiss.IndexIVFFlat(faiss.IndexFlatIP(d, device=rank), d, k, faiss.METRIC_INNER_PRODUCT)
        objs = train_ivf_sharded(ivf, torch.from_numpy(x[lo:hi]).cuda(rank))
        assert ivf.is_trained and ivf.quantizer.ntotal == k
        cent = ivf.centroids()
        ivf.add_with_ids(x[lo:hi], np.arange(lo, hi, dtype=np.int64))
        ivf.nprobe = 8
        sh = ShardedIndex(ivf)
        D, I = sh.search(x[:6], 10)
        np.savez(os.path.join(out, f"train_r{rank}.npz"), cent=cent, objs=np.asarray(objs), D=D, I=I)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_sharded_kmeans_and_ivf_search(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_train_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (np.load(tmp_path / f"train_r{r}.npz") for r in (0, 1))
    assert np.array_equal(r0["cent"], r1["cent"])  # replicated centroids stay bit-identical
    objs = r0["objs"]
    assert len(objs) == 10 and np.all(np.linalg.norm(r0["cent"], axis=1) <= 1.0 + 1e-4)  # cp.spherical = False: plain means
    n, d, k = 40000, 64, 100
    x = O.clustered_unit(n, d, 100, 31)
    c_ref, objs_ref = O.kmeans_train(x, k)
    obj_gpu = float(np.max(x.astype(np.float64) @ r0["cent"].astype(np.float64).T, axis=1).sum())
    obj_ref = float(np.max(x.astype(np.float64) @ c_ref.astype(np.float64).T, axis=1).sum())
    assert obj_gpu >= 0.99 * obj_ref, (obj_gpu, obj_ref)
    # sharded IVF search == single logical IVF index with the same centroids (oracle), on both ranks
    a = O.ivf_assign(x, r0["cent"])
    Dr, Ir = O.ivf_search(x, None, a, r0["cent"], x[:6], 10, 8)
    assert np.array_equal(r0["I"], r1["I"])
    O.compare_topk(r0["D"], r0["I"], Dr, Ir, band=4e-6)
