"""BASELINE.json configs 3 and 4 at scale, one process per GPU (torchrun) or a single GPU.

  c4: IVF k-means training, 10M x 1024 -> nlist 16384, 10 iterations (rows sharded, all-reduce of sums/counts)
  c3: IndexIVFFlat nlist=4096 over 50M x 512, rows sharded, nprobe 8..128, peer-memory top-k exchange

    python scripts/bench_configs.py --config c4 [--rows N]
    python -m torch.distributed.run --nproc-per-node 8 scripts/bench_configs.py --config c3
Prints JSON lines (rank 0)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--config", required=True, choices=["c3", "c4"])
ap.add_argument("--rows", type=int, default=0)
ap.add_argument("--niter", type=int, default=10)
a = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lrank = int(os.environ.get("LOCAL_RANK", "0"))
os.environ["WISE_B200_DEVICE"] = str(lrank)
torch.cuda.set_device(lrank); dev = torch.device("cuda", lrank)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
from wise_b200 import faiss_compat as faiss, _capi
from wise_b200.sharded import ShardedIndex, shard_range, train_ivf_sharded
from bench import fill_index_clustered, make_queries
L = _capi.lib()
def say(**kw):
    if rank == 0: print(json.dumps(kw), flush=True)
def gen_rows(lo, hi, d, seed, ncentres):
    """clustered unit rows [lo,hi) as one CUDA tensor (chunk-addressed seeds: independent of sharding)"""
    g = torch.Generator(device=dev); g.manual_seed(seed)
    centres = torch.nn.functional.normalize(torch.randn(ncentres, d, device=dev, generator=g), dim=1)
    out = torch.empty((hi - lo, d), device=dev)
    chunk = 500_000
    for s in range(lo, hi, chunk):
        e = min(hi, s + chunk)
        g2 = torch.Generator(device=dev); g2.manual_seed(seed * 1_000_003 + s)
        j = torch.randint(0, ncentres, (e - s,), device=dev, generator=g2)
        x = centres[j] + 0.6 * torch.randn(e - s, d, device=dev, generator=g2) / (d ** 0.5)
        out[s - lo:e - lo] = torch.nn.functional.normalize(x, dim=1)
    return centres, out

if a.config == "c4":
    n, d, k = a.rows or 10_000_000, 1024, 16384
    lo, hi = shard_range(n, rank, world)
    _, x = gen_rows(lo, hi, d, 7, k)
    ivf = faiss.IndexIVFFlat(faiss.IndexFlatIP(d, device=lrank), d, k, faiss.METRIC_INNER_PRODUCT)
    if world > 1:  # communicator set-up (hundreds of ms on the first collective) is not part of an iteration
        w = torch.zeros(k * d + k, device=dev); dist.all_reduce(w); dist.all_reduce(w.long()); del w
    torch.cuda.synchronize(); t0 = time.time()
    objs = train_ivf_sharded(ivf, x, niter=a.niter, verbose=False)
    torch.cuda.synchronize(); dt = time.time() - t0
    flop = 2.0 * n * k * d * a.niter
    say(config="c4", n_gpus=world, rows=n, d=d, nlist=k, niter=a.niter, seconds=dt, s_per_iter=dt / a.niter,
        assign_tflops_algorithmic=flop / dt / 1e12, objective_first=objs[0], objective_last=objs[-1],
        mean_best_ip=objs[-1] / n)
else:
    n, d, nlist, k = a.rows or 50_000_000, 512, 4096, 100
    lo, hi = shard_range(n, rank, world)
    centres, x = gen_rows(lo, hi, d, 50, 4096)
    ntrain = 100 * nlist  # the reference's rule (feature_search_index.py:55-59)
    tl, th = shard_range(ntrain, rank, world)
    sel = torch.randperm(hi - lo, device=dev)[: th - tl]
    ivf = faiss.IndexIVFFlat(faiss.IndexFlatIP(d, device=lrank), d, nlist, faiss.METRIC_INNER_PRODUCT)
    torch.cuda.synchronize(); t0 = time.time()
    train_ivf_sharded(ivf, x[sel].contiguous())
    torch.cuda.synchronize(); t_train = time.time() - t0
    ivf.reserve(hi - lo)
    st = torch.cuda.current_stream().cuda_stream
    t0 = time.time()
    for s in range(0, hi - lo, 1 << 20):
        e = min(hi - lo, s + (1 << 20))
        ids = torch.arange(lo + s, lo + e, dtype=torch.int64, device=dev)
        _capi.check(L.wb_add_with_ids_dev(ivf._h, e - s, x[s:e].data_ptr(), ids.data_ptr(), st))
    torch.cuda.synchronize(); t_add = time.time() - t0
    say(config="c3", phase="build", n_gpus=world, rows=n, rows_per_gpu=hi - lo, train_s=t_train, add_s=t_add,
        add_rows_per_s_per_gpu=(hi - lo) / t_add)
    flat = faiss.IndexIDMap(faiss.IndexFlatIP(d, device=lrank))
    flat.reserve(hi - lo)
    ids = torch.arange(lo, hi, dtype=torch.int64, device=dev)
    _capi.check(L.wb_add_with_ids_dev(flat._h, hi - lo, x.data_ptr(), ids.data_ptr(), st)); torch.cuda.synchronize()
    del x
    sh_ivf, sh_flat = ShardedIndex(ivf), ShardedIndex(flat)
    for nq in (1, 256):
        q = make_queries(centres, nq, d, 51, dev)
        Df, If = sh_flat.search_dev(q, k)
        for nprobe in (8, 16, 32, 64, 128):
            for _ in range(3): sh_ivf.search_dev(q, k, nprobe)
            if world > 1: dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            steps = 20 if nq == 1 else 5
            e0.record()
            for _ in range(steps): D, I = sh_ivf.search_dev(q, k, nprobe)
            e1.record(); torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
            if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            rec = float(np.mean([len(set(I[i].tolist()) & set(If[i].tolist())) / k for i in range(nq)]))
            say(config="c3", phase="search", n_gpus=world, nq=nq, nprobe=nprobe, k=k, ms_per_batch=float(ms.item()),
                qps=nq / float(ms.item()) * 1e3, recall_vs_flat=rec)
if world > 1:
    dist.destroy_process_group()
