"""BASELINE.json configs 3 and 4 at full size, one process per GPU (torchrun) or a single GPU - with the same
evidence as bench.py: roofline, clocks, an in-run parity check at size, and the CPU restatement timed beside it.

  c3: IndexIVFFlat nlist=4096 over 50M x 512, rows sharded, nprobe 8..128, nq 1 and 256, NVLink top-k exchange
      roofline : HBM - bytes of the DISTINCT probed lists of the batch (each list counted once) / list-scan time
      parity   : fp64 pass over every rank's rows restricted to the lists the coarse quantizer must probe
                 (bench.parity_check with a candidate mask): returned scores within 1e-5, no candidate outside the
                 result beats the k-th score beyond 4e-6, identical bytes on all ranks; plus recall vs exhaustive
      cpu      : bench.cpu_ivf - the restated IndexIVFFlat.search in C on a row sample with the same centroids
  c4: IVF k-means training, 10M x 1024 -> nlist 16384 (rows sharded, one all-reduce of sums + counts per iteration)
      roofline : tensor - 2*n*nlist*d flop per iteration / iteration time vs the TF32 peak measured in the run
      parity   : a 100k-point sample per rank: the assigned centroid's exact score is within the TF32 band of the
                 best one; centroids of 64 lists recomputed in fp64 from their members
      cpu      : bench.cpu_kmeans_iteration - one restated iteration (numpy/OpenBLAS) on a bounded sample, scaled

    python scripts/bench_configs.py --config c4 [--rows N] [--niter K]
    python -m torch.distributed.run --nproc-per-node 8 scripts/bench_configs.py --config c3
Prints JSON lines (rank 0)."""
import argparse, ctypes, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--config", required=True, choices=["c3", "c4"])
ap.add_argument("--rows", type=int, default=0)
ap.add_argument("--niter", type=int, default=10)
ap.add_argument("--no-cpu-baseline", action="store_true")
a = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lrank = int(os.environ.get("LOCAL_RANK", "0"))
os.environ["WISE_B200_DEVICE"] = str(lrank)
torch.cuda.set_device(lrank); dev = torch.device("cuda", lrank)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
import bench
from wise_b200 import faiss_compat as faiss, _capi
from wise_b200.sharded import ShardedIndex, shard_range, train_ivf_sharded
L = _capi.lib()
PEAKS = json.load(open(os.path.join(bench.ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(bench.ROOT, "MEASURED_PEAKS.json")) else {}
HBM_PEAK = float(PEAKS.get("hbm_gbs", bench.HBM_FALLBACK_GBS))


def say(**kw):
    if rank == 0:
        print(json.dumps(kw), flush=True)


def allmax(v):
    t = torch.tensor([v], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def allsum(v):
    t = torch.tensor([v], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t)
    return float(t.item())


class TensorRows:
    """bench.RowSource-compatible view of rows that already sit in a tensor (global rows [lo, lo+len))."""
    def __init__(self, x, lo):
        self.x, self.lo = x, lo
    def chunks(self, lo, hi, step=250_000):
        for s in range(lo, hi, step):
            e = min(hi, s + step)
            yield s, e, self.x[s - self.lo:e - self.lo]


if a.config == "c4":
    n, d, k = a.rows or 10_000_000, 1024, 16384
    lo, hi = shard_range(n, rank, world)
    src = bench.RowSource(n, d, 7, dev, ncentres=k)
    x = torch.cat([c for _, _, c in src.chunks(lo, hi)])
    ivf = faiss.IndexIVFFlat(faiss.IndexFlatIP(d, device=lrank), d, k, faiss.METRIC_INNER_PRODUCT)
    if world > 1:  # communicator set-up (hundreds of ms on the first collective) is not part of an iteration
        w = torch.zeros(k * d + k, device=dev); dist.all_reduce(w); dist.all_reduce(w.long()); del w
    smp = bench.ClockSampler(lrank) if rank == 0 else None
    if smp: smp.start()
    torch.cuda.synchronize(); t0 = time.time()
    objs = train_ivf_sharded(ivf, x, niter=a.niter, verbose=False)
    torch.cuda.synchronize(); dt = allmax(time.time() - t0)
    clocks = smp.finish() if smp else None
    # the all-reduce alone (same sizes as an iteration's): its share of the iteration
    t_ar = 0.0
    if world > 1:
        sums = torch.zeros(k, d, device=dev); cnts = torch.zeros(k, dtype=torch.int64, device=dev)
        dist.all_reduce(sums); torch.cuda.synchronize(); t1 = time.time()
        for _ in range(3): dist.all_reduce(sums); dist.all_reduce(cnts)
        torch.cuda.synchronize(); t_ar = allmax((time.time() - t1) / 3)
    # parity at size: one more assignment from the FINAL centroids on a sample, checked in fp64
    cent = torch.from_numpy(ivf.centroids()).to(dev)
    ns = min(100_000, hi - lo)
    xs = x[:ns].contiguous()
    asg = torch.empty(ns, dtype=torch.int32, device=dev)
    obj = ctypes.c_double(0)
    st = torch.cuda.current_stream().cuda_stream
    _capi.check(L.wb_kmeans_assign_fast_dev(ivf._h, ns, xs.data_ptr(), asg.data_ptr(), ctypes.byref(obj), st))
    torch.cuda.synchronize()
    worst, agree = 0.0, 0
    for s0 in range(0, ns, 10_000):
        s64 = xs[s0:s0 + 10_000].double() @ cent.double().T
        best, arg = s64.max(dim=1)
        got = s64.gather(1, asg[s0:s0 + 10_000].long().unsqueeze(1)).squeeze(1)
        worst = max(worst, float((best - got).max())); agree += int((arg == asg[s0:s0 + 10_000].long()).sum())
    # centroid update: one full iteration's sums for 64 lists, recomputed in fp64 from the assignment
    asg_all = torch.empty(hi - lo, dtype=torch.int32, device=dev)
    _capi.check(L.wb_kmeans_assign_fast_dev(ivf._h, hi - lo, x.data_ptr(), asg_all.data_ptr(), None, st))
    sums = torch.empty(k, d, device=dev); cnts = torch.empty(k, dtype=torch.int64, device=dev)
    _capi.check(L.wb_kmeans_accumulate_dev(ivf._h, hi - lo, x.data_ptr(), asg_all.data_ptr(), sums.data_ptr(), cnts.data_ptr(), st))
    torch.cuda.synchronize()
    upd_err = 0.0
    for l in range(0, k, k // 64):
        mem = (asg_all == l).nonzero().squeeze(1)
        ref = x[mem].double().sum(dim=0)
        upd_err = max(upd_err, float((ref - sums[l].double()).abs().max()))
        assert int(cnts[l]) == mem.numel()
    parity = {"sample_points_per_rank": ns, "max_score_gap_to_best": allmax(worst), "tf32_band": 2e-3,
              "assignment_agreement": allsum(agree) / (ns * world), "max_abs_err_of_sums_64_lists": allmax(upd_err),
              "ok": bool(allmax(worst) <= 2e-3 and allmax(upd_err) <= 1e-2)}
    tb, ts = ctypes.c_double(), ctypes.c_double()
    _capi.check(L.wb_tf32_peak(lrank, 20000, 400, ctypes.byref(tb), ctypes.byref(ts)))
    flop_iter = 2.0 * n * k * d
    ach = flop_iter * a.niter / dt / 1e12 / world
    line = dict(config="c4", metric="s per k-means iteration, 10M x 1024 -> 16384", value=dt / a.niter, unit="s", n_gpus=world, rows=n, d=d,
                nlist=k, niter=a.niter, seconds=dt, dtype="tf32 assignment (training only), fp32 update", scaling="strong",
                roofline={"bound": "tensor", "achieved": ach, "peak": ts.value, "unit": "TFLOP/s per GPU", "frac": ach / ts.value,
                          "peak_burst": tb.value, "flop_per_iteration": flop_iter,
                          "note": "whole iteration (assignment + grouping + sums + all-reduce + update) timed by the wall clock, max over ranks",
                          "allreduce_s_per_iter": t_ar, "allreduce_share": t_ar / (dt / a.niter) if world > 1 else 0.0},
                clocks=clocks, parity_check=parity, objective_first=objs[0], objective_last=objs[-1], mean_best_ip=objs[-1] / n)
    if rank == 0 and not a.no_cpu_baseline:
        ncpu = 20_000
        sec, tc, th = bench.cpu_kmeans_iteration(x[:ncpu].cpu().numpy(), cent.cpu().numpy(), n)
        line["cpu_baseline"] = {"value": sec, "unit": "s per iteration", "cores": th, "kind": "port",
                                "sample": f"{ncpu} of {n} points against all {k} centroids, one restated Clustering iteration "
                                          f"(numpy/OpenBLAS fp64 assignment + update), {tc:.2f} s scaled x{n / ncpu:g}",
                                "host": bench.cpu_info()}
    say(**line)
else:
    n, d, nlist, k = a.rows or 50_000_000, 512, 4096, 100
    lo, hi = shard_range(n, rank, world)
    src = bench.RowSource(n, d, 50, dev)
    centres = src.centres
    x = torch.cat([c for _, _, c in src.chunks(lo, hi)])
    ntrain = 100 * nlist  # the reference's rule (feature_search_index.py:55-59)
    tl, th = shard_range(ntrain, rank, world)
    sel = torch.randperm(hi - lo, device=dev)[: th - tl]
    ivf = faiss.IndexIVFFlat(faiss.IndexFlatIP(d, device=lrank), d, nlist, faiss.METRIC_INNER_PRODUCT)
    if world > 1:
        w = torch.zeros(nlist * d + nlist, device=dev); dist.all_reduce(w); dist.all_reduce(w.long()); del w
    torch.cuda.synchronize(); t0 = time.time()
    train_ivf_sharded(ivf, x[sel].contiguous())
    torch.cuda.synchronize(); t_train = allmax(time.time() - t0)
    ivf.reserve(hi - lo)
    st = torch.cuda.current_stream().cuda_stream
    t0 = time.time()
    for s in range(0, hi - lo, 1 << 20):
        e = min(hi - lo, s + (1 << 20))
        ids = torch.arange(lo + s, lo + e, dtype=torch.int64, device=dev)
        _capi.check(L.wb_add_with_ids_dev(ivf._h, e - s, x[s:e].data_ptr(), ids.data_ptr(), st))
    torch.cuda.synchronize(); t_add = allmax(time.time() - t0)
    # list of every row (insertion order: exported before the first search regroups the store)
    assign = torch.empty(hi - lo, dtype=torch.int32)
    for s in range(0, hi - lo, 1 << 22):
        m = min(1 << 22, hi - lo - s)
        _capi.check(L.wb_export_rows(ivf._h, s, m, None, None, _capi.ptr(assign[s:s + m].numpy())))
    assign = assign.to(dev).long()
    sizes = torch.bincount(assign, minlength=nlist)
    say(config="c3", phase="build", n_gpus=world, rows=n, rows_per_gpu=hi - lo, train_s=t_train, add_s=t_add,
        add_rows_per_s_per_gpu=(hi - lo) / t_add, list_rows_min=int(sizes.min()), list_rows_max=int(sizes.max()))
    flat = faiss.IndexIDMap(faiss.IndexFlatIP(d, device=lrank))
    flat.reserve(hi - lo)
    ids = torch.arange(lo, hi, dtype=torch.int64, device=dev)
    _capi.check(L.wb_add_with_ids_dev(flat._h, hi - lo, x.data_ptr(), ids.data_ptr(), st)); torch.cuda.synchronize()
    cent = torch.from_numpy(ivf.centroids()).to(dev)
    rows_src = TensorRows(x, lo)
    sh_ivf, sh_flat = ShardedIndex(ivf), ShardedIndex(flat)
    L.wb_set_timing(ivf._h, 1)
    cpu_sample = None
    for nq in (1, 256):
        q = bench.make_queries(centres, nq, d, 51, dev)
        Df, If = sh_flat.search_dev(q, k)
        sc = q.double() @ cent.double().T  # exact coarse scores [nq, nlist]
        for nprobe in (8, 16, 32, 64, 128):
            for _ in range(3): sh_ivf.search_dev(q, k, nprobe)
            if world > 1: dist.barrier()
            torch.cuda.synchronize()
            smp = bench.ClockSampler(lrank) if rank == 0 else None
            if smp: smp.start()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            steps = 50 if nq == 1 else 10
            e0.record()
            for _ in range(steps): D, I = sh_ivf.search_dev(q, k, nprobe)
            e1.record(); torch.cuda.synchronize()
            clocks = smp.finish() if smp else None
            ms = allmax(e0.elapsed_time(e1) / steps)
            buf = (ctypes.c_float * 128)()
            ns_ = L.wb_scan_ms_history(ivf._h, buf, min(steps, 128))
            scan_ms = allmax(float(np.mean([buf[i] for i in range(ns_)])))
            # algorithmic bytes: every DISTINCT probed list of the batch once (this rank's slice of it)
            top = torch.topk(sc, nprobe, dim=1)
            probed = torch.zeros(nlist, dtype=torch.bool, device=dev); probed[top.indices.reshape(-1)] = True
            distinct_bytes = float(sizes[probed].sum()) * d * 4
            pair_bytes = float(sizes[top.indices.reshape(-1)].sum()) * d * 4
            gbs = distinct_bytes / (scan_ms * 1e-3) / 1e9
            # parity at size: candidates = rows of the lists whose exact coarse score clears the nprobe-th by 1e-6
            # (must be probed); returned rows may also come from lists within 1e-6 below it
            thr = top.values[:, -1]
            strict_l = sc >= (thr + 1e-6).unsqueeze(1)
            loose_l = sc >= (thr - 1e-6).unsqueeze(1)
            def allowed(s, e):
                a_ = assign[s - lo:e - lo]
                return strict_l[:, a_].T, loose_l[:, a_].T
            par = bench.parity_check(rows_src, lo, hi, q, D, I, k, world, dev, allowed=allowed)
            rec = float(np.mean([len(set(I[i].tolist()) & set(If[i].tolist())) / k for i in range(nq)]))
            line = dict(config="c3", phase="search", metric="QPS, IndexIVFFlat nlist=4096 top-100 (50Mx512 fp32)", value=nq / ms * 1e3,
                        unit="queries/s", n_gpus=world, nq=nq, nprobe=nprobe, k=k, ms_per_batch=ms, scaling="strong",
                        roofline={"bound": "hbm", "achieved": gbs, "peak": HBM_PEAK, "unit": "GB/s per GPU", "frac": gbs / HBM_PEAK,
                                  "kernel": "ivf_listmajor_kernel" if nq >= 8 and nq * nprobe >= nlist else "scan_topk_kernel<1,RW,true>",
                                  "launch_ms": scan_ms, "distinct_list_bytes_per_gpu": distinct_bytes,
                                  "query_major_bytes_per_gpu": pair_bytes,
                                  "note": "list-scan kernel only (CUDA events inside the library); ms_per_batch also holds the "
                                          "coarse quantizer, the merge and the NVLink exchange"},
                        clocks=clocks, parity_check=par, recall_vs_flat=rec)
            if rank == 0 and not a.no_cpu_baseline and nprobe in (8, 128):
                if cpu_sample is None:  # 500k rows of rank 0 with their lists
                    m = min(500_000, hi - lo)
                    cpu_sample = (x[:m].cpu().numpy(), assign[:m].cpu().numpy(), cent.cpu().numpy(), m)
                xs_, as_, cc, m = cpu_sample
                qps_c, tc, th = bench.cpu_ivf(xs_, as_, cc, q.cpu().numpy(), k, nprobe, n)
                line["cpu_baseline"] = {"value": qps_c, "unit": "queries/s", "cores": th, "kind": "port",
                                        "sample": f"{m} of {n} rows (same centroids and lists), median {tc * 1e3:.2f} ms scaled x{n / m:g}; "
                                                  "the restated IndexIVFFlat.search (C, AVX-512), parallel over queries like faiss parallel_mode 0",
                                        "host": bench.cpu_info()}
            say(**line)
if world > 1:
    dist.destroy_process_group()
