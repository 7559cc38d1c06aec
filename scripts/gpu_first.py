"""Dev script: first contact with the GPU - correctness over a grid of shapes + scan timing."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as O
from wise_b200 import faiss_compat as faiss, _capi

def check(n, d, nq, k, seed=0, dup=False):
    xb = O.unit_gaussian(n, d, 100 + seed)
    if dup and n > 100:
        xb[n // 2: n // 2 + n // 100] = xb[: n // 100]
    xq = O.unit_gaussian(nq, d, 200 + seed)
    idx = faiss.IndexFlatIP(d)
    idx.add(xb)
    t = time.time(); D, I = idx.search(xq, k); dt = time.time() - t
    Dr, Ir = O.flat_search(xb, xq, k)
    r = O.compare_topk(D, I, Dr, Ir)
    print(f"n={n} d={d} nq={nq} k={k} dup={dup}: {r} maxerr={np.abs(D-Dr)[Ir>=0].max() if (Ir>=0).any() else 0:.2e} {dt*1e3:.2f} ms", flush=True)

for (n, d, nq, k) in [(1000, 64, 1, 5), (5, 64, 2, 10), (100000, 512, 16, 10), (100000, 768, 1, 100), (50000, 768, 8, 100),
                      (33333, 1024, 3, 1000), (20000, 100, 5, 7), (20000, 30, 9, 2048), (100000, 512, 40, 1)]:
    check(n, d, nq, k)
check(100000, 512, 16, 10, dup=True)
import __graft_entry__ as g
g.smoke()

# timing: 2M x 768 (6.1 GB) - larger than L2
import torch
n, d = 2_000_000, 768
gen = torch.Generator(device="cuda"); gen.manual_seed(1)
x = torch.randn(n, d, device="cuda", generator=gen); x /= x.norm(dim=1, keepdim=True)
idx = faiss.IndexFlatIP(d)
L = _capi.lib()
idx.reserve(n)
_capi.check(L.wb_add_with_ids_dev(idx._h, n, x.data_ptr(), None, None))
torch.cuda.synchronize()
L.wb_set_timing(idx._h, 1)
for nq in (1, 2, 4, 8, 16):
    q = O.unit_gaussian(nq, d, 5)
    for k in (10, 100, 1000):
        ts = []
        for it in range(6):
            t = time.time(); D, I = idx.search(q, k); dt = time.time() - t
            ts.append((L.wb_last_scan_ms(idx._h), dt * 1e3))
        ms = min(t[0] for t in ts[2:]); e2e = min(t[1] for t in ts[2:])
        print(f"nq={nq} k={k}: scan {ms:.3f} ms = {n*d*4/ms/1e6:.0f} GB/s (per pass x{(nq+7)//8}), e2e {e2e:.3f} ms", flush=True)
