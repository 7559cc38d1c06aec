"""Round-2 profiling targets: one mode per family of shipped kernels, small enough for `ncu --set full` replays.
    python scripts/prof_r02.py flat1|flat16|flat1024|ivf_search|ivf_train|ivf_add|ivf_fused [rows]
Each mode prints its own CUDA-event timings, so the plain run doubles as a sanity check."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from wise_b200 import faiss_compat as faiss, _capi

L = _capi.lib()
mode = sys.argv[1]
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream


def search_dev(idx, q, k, nprobe=1, reps=3):
    D = torch.empty(q.shape[0], k, device=dev); I = torch.empty(q.shape[0], k, dtype=torch.int64, device=dev)
    L.wb_set_timing(idx._h, 1)
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        _capi.check(L.wb_search_dev(idx._h, q.shape[0], q.data_ptr(), k, nprobe, D.data_ptr(), I.data_ptr(), st))
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    print(f"{mode}: nq={q.shape[0]} nprobe={nprobe}: host {min(ts):.3f} ms, kernel {L.wb_last_scan_ms(idx._h):.3f} ms", flush=True)
    return D, I


if mode.startswith("flat"):
    d = 768
    src = bench.RowSource(rows, d, 2024, dev)
    idx = faiss.IndexIDMap(faiss.IndexFlatIP(d)); bench.fill_index(idx, src, 0, rows)
    nq = int(mode[4:])
    search_dev(idx, bench.make_queries(src.centres, nq, d, 2025, dev), 100, reps=4 if nq == 1 else 2)
elif mode == "ivf_fused":
    # one-query IVF search as ONE cooperative launch (coarse quantizer inside the list scan, coarse.cuh): nlist 4096
    # like BASELINE config 3; centroids = rows (no training), so every scan_topk_kernel launch here is a fused search
    d, nlist = 512, 4096
    src = bench.RowSource(rows, d, 50, dev)
    ivf = faiss.IndexIVFFlat(faiss.IndexFlatIP(d), d, nlist, faiss.METRIC_INNER_PRODUCT)
    xs = torch.cat([x for _, _, x in src.chunks(0, nlist)])[:nlist]
    ivf.set_centroids(np.ascontiguousarray(xs.cpu().numpy()))
    ivf.reserve(rows)
    for s, e, x in src.chunks(0, rows):
        ids = torch.arange(s, e, dtype=torch.int64, device=dev)
        _capi.check(L.wb_add_with_ids_dev(ivf._h, e - s, x.data_ptr(), ids.data_ptr(), st))
    torch.cuda.synchronize()
    q = bench.make_queries(src.centres, 1, d, 51, dev)
    for nprobe in (8, 32):
        f0 = L.wb_ivf_fused_searches(ivf._h)
        search_dev(ivf, q, 100, nprobe=nprobe, reps=4)
        print(f"ivf_fused: nprobe={nprobe}: {L.wb_ivf_fused_searches(ivf._h) - f0} of 4 searches were one launch", flush=True)
else:
    d, nlist = 512, 1024
    src = bench.RowSource(rows, d, 50, dev)
    ivf = faiss.IndexIVFFlat(faiss.IndexFlatIP(d), d, nlist, faiss.METRIC_INNER_PRODUCT)
    xs = torch.cat([x for _, _, x in src.chunks(0, min(rows, 100 * nlist))])
    if mode == "ivf_train":
        assign = torch.empty(xs.shape[0], dtype=torch.int32, device=dev)
        sums = torch.empty(nlist, d, device=dev); cnts = torch.empty(nlist, dtype=torch.int64, device=dev)
        init = np.ascontiguousarray(xs[:nlist].cpu().numpy())
        _capi.check(L.wb_ivf_set_centroids(ivf._h, _capi.ptr(init)))
        for it in range(3):
            obj = C.c_double(0); ns = C.c_int64(0)
            t0 = time.perf_counter()
            _capi.check(L.wb_kmeans_assign_fast_dev(ivf._h, xs.shape[0], xs.data_ptr(), assign.data_ptr(), C.byref(obj), st))
            _capi.check(L.wb_kmeans_accumulate_dev(ivf._h, xs.shape[0], xs.data_ptr(), assign.data_ptr(), sums.data_ptr(), cnts.data_ptr(), st))
            _capi.check(L.wb_kmeans_update_dev(ivf._h, sums.data_ptr(), cnts.data_ptr(), xs.shape[0], 1234, C.byref(ns), st))
            torch.cuda.synchronize()
            print(f"ivf_train: iteration {it}: {(time.perf_counter() - t0) * 1e3:.2f} ms, objective {obj.value:.1f}", flush=True)
    else:
        ivf.train(xs.cpu().numpy())
        ivf.reserve(rows)
        t0 = time.perf_counter()
        for s, e, x in src.chunks(0, rows):
            ids = torch.arange(s, e, dtype=torch.int64, device=dev)
            _capi.check(L.wb_add_with_ids_dev(ivf._h, e - s, x.data_ptr(), ids.data_ptr(), st))
        torch.cuda.synchronize()
        print(f"{mode}: add {rows} rows in {time.perf_counter() - t0:.3f} s", flush=True)
        if mode == "ivf_search":
            search_dev(ivf, bench.make_queries(src.centres, 1, d, 51, dev), 100, nprobe=32, reps=4)
            search_dev(ivf, bench.make_queries(src.centres, 256, d, 52, dev), 100, nprobe=64, reps=3)
