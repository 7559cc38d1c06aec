"""Small end-to-end pass over every kernel for compute-sanitizer (memcheck / racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from oracle import oracle as O
from wise_b200 import faiss_compat as faiss
d = 96
xb = O.clustered_unit(9000, d, 40, 1); ids = np.arange(9000, dtype=np.int64) * 2 + 1
flat = faiss.IndexIDMap(faiss.IndexFlatIP(d)); flat.add_with_ids(xb, ids)
for nq, k in ((1, 10), (3, 100), (16, 10), (130, 5)):
    xq = O.clustered_unit(nq, d, 40, 2 + nq)
    D, I = flat.search(xq, k)
    O.compare_topk(D, I, *O.flat_search(xb, xq, k, ids), band=4e-6)
print("flat ok", flush=True)
ivf = faiss.IndexIVFFlat(faiss.IndexFlatIP(d), d, 64, faiss.METRIC_INNER_PRODUCT)
ivf.train(xb[:4000]); ivf.add_with_ids(xb, ids); ivf.nprobe = 8
D, I = ivf.search(xb[:3], 10)
assert (I[:, 0] == ids[:3]).all()
print("ivf ok", flush=True)
print(faiss._reconstruct(flat, [ids[5]]).shape)
