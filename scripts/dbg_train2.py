import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from oracle import oracle as O
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
os.environ["WISE_B200_DEVICE"] = str(rank)
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
from wise_b200 import faiss_compat as faiss, _capi
from wise_b200.sharded import shard_range
import ctypes as C
L = _capi.lib()
n, d, k = 40000, 64, 100
x = O.clustered_unit(n, d, 100, 31)
lo, hi = shard_range(n, rank, world)
ivf = faiss.IndexIVFFlat(faiss.IndexFlatIP(d, device=rank), d, k, faiss.METRIC_INNER_PRODUCT)
ivf.set_centroids(O.kmeans_init(x, k))
xl = torch.from_numpy(x[lo:hi]).cuda(rank)
st = torch.cuda.current_stream().cuda_stream
assign = torch.empty(hi - lo, dtype=torch.int32, device=xl.device)
for it in range(3):
    obj = C.c_double(0)
    _capi.check(L.wb_kmeans_assign_dev(ivf._h, hi - lo, xl.data_ptr(), assign.data_ptr(), C.byref(obj), st))
    torch.cuda.synchronize()
    a = assign.cpu().numpy()
    ref = O.ivf_assign(x[lo:hi], ivf.centroids())
    print(f"rank {rank} it {it}: obj {obj.value:.4e} assign match {(a == ref).mean():.4f} min {a.min()} max {a.max()}", flush=True)
dist.destroy_process_group()
