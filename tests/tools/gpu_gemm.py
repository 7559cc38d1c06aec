"""Dev script: first contact of the tcgen05 path (K2) - parity vs the oracle, then timing."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from oracle import oracle as O
from wise_b200 import faiss_compat as faiss, _capi
L = _capi.lib()
def stats(idx):
    a, b = C.c_int64(), C.c_int64(); L.wb_gemm_stats(idx._h, C.byref(a), C.byref(b)); return a.value, b.value
def check(n, d, nq, k, clustered=False):
    xb = O.clustered_unit(n, d, 64, 1) if clustered else O.unit_gaussian(n, d, 100)
    xq = O.clustered_unit(nq, d, 64, 2) if clustered else O.unit_gaussian(nq, d, 200)
    idx = faiss.IndexFlatIP(d); idx.add(xb)
    t = time.time(); D, I = idx.search(xq, k); dt = time.time() - t
    Dr, Ir = O.flat_search(xb, xq, k)
    err = np.abs(D - Dr)[Ir >= 0].max()
    try:
        r = O.compare_topk(D, I, Dr, Ir)
    except AssertionError as e:
        r = "FAIL " + str(e)[:200]
    print(f"n={n} d={d} nq={nq} k={k}: {r} maxerr={err:.2e} {dt*1e3:.2f} ms gemm(epochs,fallbacks)={stats(idx)}", flush=True)
for cfg in [(40000, 768, 16, 10), (40000, 768, 128, 10), (100000, 512, 200, 100), (50000, 100, 33, 5), (131072, 768, 300, 100, True),
            (70001, 1024, 64, 50), (300000, 64, 1000, 10)]:
    check(*cfg)
if len(sys.argv) > 1 and sys.argv[1] == "time":
    import torch
    n, d = 4_000_000, 768
    gen = torch.Generator(device="cuda"); gen.manual_seed(1)
    x = torch.randn(n, d, device="cuda", generator=gen); x /= x.norm(dim=1, keepdim=True)
    idx = faiss.IndexFlatIP(d); idx.reserve(n)
    _capi.check(L.wb_add_with_ids_dev(idx._h, n, x.data_ptr(), None, None)); torch.cuda.synchronize(); del x
    L.wb_set_timing(idx._h, 1)
    st = torch.cuda.current_stream().cuda_stream
    for nq in (16, 64, 128, 256, 1024):
        q = torch.randn(nq, d, device="cuda"); q /= q.norm(dim=1, keepdim=True)
        D = torch.empty(nq, 100, device="cuda"); I = torch.empty(nq, 100, dtype=torch.int64, device="cuda")
        ts = []
        for _ in range(4):
            _capi.check(L.wb_search_dev(idx._h, nq, q.data_ptr(), 100, 1, D.data_ptr(), I.data_ptr(), st)); torch.cuda.synchronize()
            ts.append(L.wb_last_scan_ms(idx._h))
        ms = min(ts[1:])
        print(f"nq={nq}: {ms:.3f} ms  {nq/ms*1e3:.0f} QPS  {2*nq*n*d/ms/1e9:.1f} TFLOP/s algorithmic ({6*nq*n*d/ms/1e9:.1f} issued)  stats={stats(idx)}", flush=True)
