import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from oracle import oracle as O
from wise_b200 import faiss_compat as faiss
n, d, nq, k = 40000, 768, 128, 100
xb = O.unit_gaussian(n, d, 100); xq = O.unit_gaussian(nq, d, 200)
for trial in range(6):
    idx = faiss.IndexFlatIP(d); idx.add(xb)
    D, I = idx.search(xq, k)
    true = np.einsum('qkd,qd->qk', xb[I].astype(np.float64), xq.astype(np.float64))
    err = np.abs(true - D)
    bad = np.argwhere(err > 1e-5)
    print(f"trial {trial}: bad entries {len(bad)} of {nq*k}; max err {err.max():.2e}")
    if len(bad):
        rows = I[bad[:, 0], bad[:, 1]]
        tiles = rows // 128
        print("  queries:", sorted(set(bad[:, 0].tolist()))[:40])
        print("  tiles  :", sorted(set(tiles.tolist()))[:40], "n distinct", len(set(tiles.tolist())))
        print("  row%128:", sorted(set((rows % 128).tolist()))[:64])
        print("  tile%148:", sorted(set((tiles % 148).tolist()))[:40], " epochs(row ranges):", sorted(set(np.digitize(rows, [4096, 16384, 65536]).tolist())))
        # also: are there MISSING true top-k members?
    Dr, Ir = O.flat_search(xb, xq, k)
    miss = sum(len(set(Ir[q]) - set(I[q])) for q in range(nq))
    print("  missing true members:", miss)
