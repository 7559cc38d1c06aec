"""Dev script for ncu: one index (2M x 768), a few searches at batch size argv[1]."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wise_b200 import faiss_compat as faiss, _capi
L = _capi.lib()
nq = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n, d = (int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000), 768
gen = torch.Generator(device="cuda"); gen.manual_seed(1)
x = torch.randn(n, d, device="cuda", generator=gen); x /= x.norm(dim=1, keepdim=True)
idx = faiss.IndexFlatIP(d); idx.reserve(n)
_capi.check(L.wb_add_with_ids_dev(idx._h, n, x.data_ptr(), None, None)); torch.cuda.synchronize(); del x
L.wb_set_timing(idx._h, 1)
st = torch.cuda.current_stream().cuda_stream
q = torch.randn(nq, d, device="cuda"); q /= q.norm(dim=1, keepdim=True)
D = torch.empty(nq, 100, device="cuda"); I = torch.empty(nq, 100, dtype=torch.int64, device="cuda")
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ts = []
for _ in range(iters):
    _capi.check(L.wb_search_dev(idx._h, nq, q.data_ptr(), 100, 1, D.data_ptr(), I.data_ptr(), st)); torch.cuda.synchronize()
    ts.append(L.wb_last_scan_ms(idx._h))
    if iters <= 3: print(nq, ts[-1], flush=True)
if iters > 3:
    tail = sorted(ts[iters // 2:])
    print(nq, "min %.3f  median(last half) %.3f ms over %d searches" % (min(ts), tail[len(tail) // 2], iters), flush=True)
