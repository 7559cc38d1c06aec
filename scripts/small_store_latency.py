"""Per-call search latency on small stores (C1-like: 100k x 512) for K1 vs forced K2.
Usage: python scripts/small_store_latency.py  (prints a small table; needs a GPU)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from wise_b200 import faiss_compat as faiss  # noqa: E402


def main():
    rng = np.random.default_rng(0)
    for n, d in ((100_000, 512), (1_000_000, 512), (2_000_000, 768)):
        xb = rng.standard_normal((n, d), dtype=np.float32)
        xb /= np.linalg.norm(xb, axis=1, keepdims=True)
        idx = faiss.IndexFlatIP(d)
        idx.add(xb)
        for nq in (1, 8, 16, 32, 64, 256):
            xq = xb[:nq].copy()
            row = []
            for force in ("0", "1"):
                os.environ["WB_GEMM_FORCE"] = force
                for _ in range(5):
                    idx.search(xq, 10)
                t0 = time.perf_counter()
                for _ in range(50):
                    idx.search(xq, 10)
                row.append((time.perf_counter() - t0) / 50 * 1e3)
            print(f"n={n} d={d} nq={nq:4d}  heuristic {row[0]:.3f} ms   forced-K2 {row[1]:.3f} ms", flush=True)


if __name__ == "__main__":
    main()
