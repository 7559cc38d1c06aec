"""IVFFlat measurement (SURVEY.md 8d, config C3 scaled to one GPU): k-means train, add, list-scan GB/s.
    python scripts/bench_ivf.py [--rows 5000000 --dim 512 --nlist 4096]
Prints one JSON line per (nq, nprobe) with the scan-kernel time (CUDA events inside the library),
the bytes of the probed lists (algorithmic) and recall@k against the exhaustive flat result."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from wise_b200 import faiss_compat as faiss, _capi
from bench import fill_index_clustered, make_queries

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=5_000_000); ap.add_argument("--dim", type=int, default=512)
ap.add_argument("--nlist", type=int, default=4096); ap.add_argument("--k", type=int, default=100)
a = ap.parse_args()
L = _capi.lib(); dev = torch.device("cuda", 0)
flat = faiss.IndexIDMap(faiss.IndexFlatIP(a.dim))
t = time.time(); centres, _ = fill_index_clustered(flat, 0, a.rows, a.dim, 50, dev); torch.cuda.synchronize()
print(f"# generated {a.rows}x{a.dim} in {time.time()-t:.1f}s", flush=True)
# training sample: the reference's rule, 100 points per centroid (feature_search_index.py:55-59)
ntrain = min(a.rows, 100 * a.nlist)
sel = np.sort(np.random.default_rng(1).choice(a.rows, ntrain, replace=False))
xs = np.empty((ntrain, a.dim), np.float32)
rows_ptr = _capi.C.c_void_p(); ld = _capi.C.c_int64()
L.wb_storage(flat._h, _capi.C.byref(rows_ptr), _capi.C.byref(ld))
store = torch.empty(0)  # view the row store through torch for gathers
import ctypes
class _Ptr:  # minimal __cuda_array_interface__ wrapper
    def __init__(s, p, shape): s.__cuda_array_interface__ = {"data": (p, False), "shape": shape, "typestr": "<f4", "version": 2}
xb = torch.as_tensor(_Ptr(rows_ptr.value, (a.rows, a.dim)), device=dev)
xs = xb[torch.from_numpy(sel).to(dev)].cpu().numpy()
ivf = faiss.IndexIVFFlat(faiss.IndexFlatIP(a.dim), a.dim, a.nlist, faiss.METRIC_INNER_PRODUCT)
t = time.time(); ivf.train(xs); t_train = time.time() - t
print(json.dumps({"phase": "train", "n": ntrain, "nlist": a.nlist, "d": a.dim, "niter": 10, "seconds": t_train,
                  "s_per_iter": t_train / 10}), flush=True)
ivf.reserve(a.rows)
t = time.time()
st = torch.cuda.current_stream().cuda_stream
for s in range(0, a.rows, 1 << 18):
    e = min(a.rows, s + (1 << 18))
    ids = torch.arange(s, e, dtype=torch.int64, device=dev)
    _capi.check(L.wb_add_with_ids_dev(ivf._h, e - s, xb[s:e].data_ptr(), ids.data_ptr(), st))
torch.cuda.synchronize(); t_add = time.time() - t
print(json.dumps({"phase": "add", "rows": a.rows, "seconds": t_add, "rows_per_s": a.rows / t_add}), flush=True)
_, _, assign = ivf._export(0, a.rows, want_assign=True)
sizes = np.bincount(assign, minlength=a.nlist)
print(json.dumps({"phase": "lists", "min": int(sizes.min()), "max": int(sizes.max()), "mean": float(sizes.mean()),
                  "imbalance": float((sizes.astype(np.float64) ** 2).sum() * a.nlist / a.rows ** 2)}), flush=True)
cent = ivf.centroids()
L.wb_set_timing(ivf._h, 1); L.wb_set_timing(flat._h, 1)
for nq in (1, 16, 256):
    q = make_queries(centres, nq, a.dim, 51, dev)
    qh = q.cpu().numpy()
    Df, If = flat.search(qh, a.k)
    coarse = np.argsort(-(qh @ cent.T), axis=1, kind="stable")
    for nprobe in (8, 32, 128):
        ivf.nprobe = nprobe
        D = torch.empty(nq, a.k, device=dev); I = torch.empty(nq, a.k, dtype=torch.int64, device=dev)
        ts, wall = [], []
        for it in range(6):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            _capi.check(L.wb_search_dev(ivf._h, nq, q.data_ptr(), a.k, nprobe, D.data_ptr(), I.data_ptr(), st))
            torch.cuda.synchronize(); wall.append(time.perf_counter() - t0)
            ts.append(L.wb_last_scan_ms(ivf._h))
        ms = float(np.median(ts[2:])); wms = float(np.median(wall[2:])) * 1e3
        rows_scanned = int(sizes[coarse[:, :nprobe]].sum())
        Ih = I.cpu().numpy()
        recall = float(np.mean([len(set(Ih[i]) & set(If[i])) / a.k for i in range(nq)]))
        print(json.dumps({"phase": "search", "nq": nq, "nprobe": nprobe, "k": a.k, "scan_ms": ms, "call_ms": wms,
                          "rows_scanned": rows_scanned, "scan_GBs": rows_scanned * a.dim * 4 / ms / 1e6,
                          "qps": nq / (wms / 1e3), "recall_vs_flat": recall}), flush=True)
