"""Dev script: sweep scan-kernel launch knobs (env-driven) on a >L2 database."""
import os, sys, time, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from wise_b200 import faiss_compat as faiss, _capi
L = _capi.lib()
n = int(os.environ.get("SWEEP_N", 4_000_000)); d = int(os.environ.get("SWEEP_D", 768))
gen = torch.Generator(device="cuda"); gen.manual_seed(1)
x = torch.randn(n, d, device="cuda", generator=gen); x /= x.norm(dim=1, keepdim=True)
idx = faiss.IndexFlatIP(d); idx.reserve(n)
_capi.check(L.wb_add_with_ids_dev(idx._h, n, x.data_ptr(), None, None)); torch.cuda.synchronize()
del x
L.wb_set_timing(idx._h, 1)
def run(nq, k, iters=8):
    q = torch.randn(nq, d, device="cuda"); q /= q.norm(dim=1, keepdim=True)
    D = torch.empty(nq, k, device="cuda"); I = torch.empty(nq, k, dtype=torch.int64, device="cuda")
    ts = []
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(iters):
        _capi.check(L.wb_search_dev(idx._h, nq, q.data_ptr(), k, 1, D.data_ptr(), I.data_ptr(), st))
        torch.cuda.synchronize()
        ts.append(L.wb_last_scan_ms(idx._h))
    return float(np.median(ts[2:])), min(ts[2:])
cfgs = eval(os.environ.get("SWEEP_CFGS", "None")) or [
    dict(CK=c, STAGES=s, WAVES=w) for c in (128, 256, 384, 768) for s in (2, 3, 4, 8) for w in (1, 2)]
for cfg in cfgs:
    for key, v in cfg.items(): os.environ["WB_SCAN_" + key] = str(v)
    out = []
    for nq in eval(os.environ.get("SWEEP_NQ", "(1, 8)")):
        try:
            med, mn = run(nq, 100)
            out.append(f"nq={nq}: {med:.3f} ms {n*d*4/med/1e6:.0f} GB/s (best {n*d*4/mn/1e6:.0f})")
        except RuntimeError as e:
            out.append(f"nq={nq}: ERR {e}")
    print(cfg, " | ".join(out), flush=True)
