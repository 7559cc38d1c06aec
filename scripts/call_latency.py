"""Per-call latency of the host API at one query (the REST API's case, /root/reference/api/routes.py:1407): wall clock
around `index.search(numpy, k)` and the kernel time inside it (CUDA events in the library), flat and IVF, with the
direct-to-pinned result write on / off (WB_DIRECT_RESULTS) and, for IVF, the coarse quantizer fused into the list-scan
launch on / off (WB_IVF_FUSE_COARSE; `device_ms` = CUDA events around 200 back-to-back wb_search_dev calls with a
device-resident query, i.e. the whole device path of a search including the coarse quantizer).  One JSON line per case.
    python scripts/call_latency.py [--ivf-only]"""
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from wise_b200 import _capi, faiss_compat as faiss  # noqa: E402
from bench import fill_index_clustered, make_queries  # noqa: E402

L = _capi.lib()
dev = torch.device("cuda", 0)


def measure(idx, q, k, reps=200):
    L.wb_set_timing(idx._h, 1)
    for _ in range(20):
        idx.search(q, k)
    wall = []
    for _ in range(reps):
        t0 = time.perf_counter()
        idx.search(q, k)
        wall.append(time.perf_counter() - t0)
    buf = (ctypes.c_float * 128)()
    n = L.wb_scan_ms_history(idx._h, buf, 128)
    wall = np.array(wall) * 1e3
    return {"call_ms_median": float(np.median(wall)), "call_ms_p10": float(np.percentile(wall, 10)),
            "call_ms_p90": float(np.percentile(wall, 90)), "kernel_ms": float(np.median([buf[i] for i in range(n)]))}


def device_ms(idx, q, k, nprobe, reps=200):
    """ms per search of the device path: back-to-back wb_search_dev calls on the current stream, CUDA events around."""
    qd = torch.from_numpy(q).to(dev)
    D = torch.empty(q.shape[0], k, device=dev)
    I = torch.empty(q.shape[0], k, dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    call = lambda: _capi.check(L.wb_search_dev(idx._h, q.shape[0], qd.data_ptr(), k, nprobe, D.data_ptr(), I.data_ptr(), st))
    for _ in range(20):
        call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        call()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    k = 100
    for n, d in (() if "--ivf-only" in sys.argv else ((100_000, 512), (1_000_000, 768))):
        flat = faiss.IndexIDMap(faiss.IndexFlatIP(d))
        centres, _ = fill_index_clustered(flat, 0, n, d, 50, dev)
        for nq in (1, 16):
            q = make_queries(centres, nq, d, 7, dev).cpu().numpy()
            for direct in ("0", "1"):
                os.environ["WB_DIRECT_RESULTS"] = direct
                print(json.dumps({"index": "IndexFlatIP", "rows": n, "d": d, "nq": nq, "k": k, "direct_results": int(direct),
                                  **measure(flat, q, k)}), flush=True)
        del flat
    n, d, nlist = 2_000_000, 512, 1024
    flat = faiss.IndexIDMap(faiss.IndexFlatIP(d))
    centres, _ = fill_index_clustered(flat, 0, n, d, 50, dev)
    rows_ptr, ld = ctypes.c_void_p(), ctypes.c_int64()
    L.wb_storage(flat._h, ctypes.byref(rows_ptr), ctypes.byref(ld))

    class _Ptr:
        def __init__(s, p, shape):
            s.__cuda_array_interface__ = {"data": (p, False), "shape": shape, "typestr": "<f4", "version": 2}
    xb = torch.as_tensor(_Ptr(rows_ptr.value, (n, d)), device=dev)
    ivf = faiss.IndexIVFFlat(faiss.IndexFlatIP(d), d, nlist, faiss.METRIC_INNER_PRODUCT)
    sel = torch.randperm(n, device=dev)[:100 * nlist]
    ivf.train(xb[sel].cpu().numpy())
    st = torch.cuda.current_stream().cuda_stream
    ivf.reserve(n)
    for s in range(0, n, 1 << 18):
        e = min(n, s + (1 << 18))
        idt = torch.arange(s, e, dtype=torch.int64, device=dev)
        _capi.check(L.wb_add_with_ids_dev(ivf._h, e - s, xb[s:e].data_ptr(), idt.data_ptr(), st))
    torch.cuda.synchronize()
    for nq in (1, 4):
        q = make_queries(centres, nq, d, 8, dev).cpu().numpy()
        for nprobe in (8, 32, 128):
            ivf.nprobe = nprobe
            ref = None
            for fused in ("0", "1"):
                os.environ["WB_IVF_FUSE_COARSE"] = fused
                f0 = L.wb_ivf_fused_searches(ivf._h)
                rec = {"index": "IndexIVFFlat", "rows": n, "d": d, "nlist": nlist, "nprobe": nprobe, "nq": nq, "k": k,
                       "fused_coarse": int(fused), **measure(ivf, q, k), "device_ms": device_ms(ivf, q, k, nprobe)}
                rec["fused_searches"] = int(L.wb_ivf_fused_searches(ivf._h) - f0)
                D, I = ivf.search(q, k)
                if ref is None:
                    ref = (D, I)
                else:
                    rec["same_bytes_as_two_launches"] = bool(np.array_equal(ref[1], I) and
                                                             np.array_equal(ref[0].view(np.uint32), D.view(np.uint32)))
                print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
