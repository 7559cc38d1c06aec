"""Per-call latency of the host API at one query (the REST API's case, /root/reference/api/routes.py:1407): wall clock
around `index.search(numpy, k)` and the kernel time inside it (CUDA events in the library), flat and IVF, with the
direct-to-pinned result write on / off (WB_DIRECT_RESULTS).  One JSON line per case.
    python scripts/call_latency.py"""
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


def main():
    k = 100
    for n, d in ((100_000, 512), (1_000_000, 768)):
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
    q = make_queries(centres, 1, d, 8, dev).cpu().numpy()
    for nprobe in (8, 32):
        ivf.nprobe = nprobe
        for direct in ("0", "1"):
            os.environ["WB_DIRECT_RESULTS"] = direct
            print(json.dumps({"index": "IndexIVFFlat", "rows": n, "d": d, "nlist": nlist, "nprobe": nprobe, "nq": 1, "k": k,
                              "direct_results": int(direct), **measure(ivf, q, k)}), flush=True)


if __name__ == "__main__":
    main()
