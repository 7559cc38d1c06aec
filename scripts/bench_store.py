"""Host-side loader throughput: WebdatasetStore shards -> (ids, vectors) batches, C++ shard reader vs python loop.
Usage: python scripts/bench_store.py [rows] [dim]   (CPU only; writes a temporary store under ./gpurun_out/)."""
import os
import shutil
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from wise_b200.store import WebdatasetStore  # noqa: E402


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
    dim = int(sys.argv[2]) if len(sys.argv) > 2 else 768
    root = os.path.join(os.path.dirname(__file__), "..", "gpurun_out", "store_bench")
    shutil.rmtree(root, ignore_errors=True)
    os.makedirs(root)
    rng = np.random.default_rng(0)
    w = WebdatasetStore("image", root)
    w.enable_write(shard_maxcount=50_000, shard_maxsize=0)  # the reference's extract-features writes one sample per add()
    t0 = time.perf_counter()
    for b in range(0, rows, 10_000):
        n = min(10_000, rows - b)
        x = rng.standard_normal((n, dim), dtype=np.float32)
        for i in range(n):
            w.add(b + i, x[i:i + 1])
    w.close()
    t_w = time.perf_counter() - t0
    size = sum(os.path.getsize(os.path.join(root, f)) for f in os.listdir(root))
    print(f"wrote {rows} x {dim} in {t_w:.2f} s ({size / 1e6:.0f} MB on disk)")
    for fast in ("1", "1") + (("0",) if os.environ.get("SLOW", "1") == "1" else ()):
        os.environ["WISE_B200_FAST_STORE"] = fast
        r = WebdatasetStore("image", root)
        t0 = time.perf_counter()
        r.enable_read(shard_shuffle=False)
        t_open = time.perf_counter() - t0
        t0 = time.perf_counter()
        got = 0
        for ids, x in r.iter_batch(batch_size=int(os.environ.get("BATCH", "65536"))):
            got += len(ids)
        dt = time.perf_counter() - t0
        print(f"fast={fast}: open {t_open:.2f} s, iter_batch {dt:.2f} s = {got / dt / 1e6:.2f} M rows/s = "
              f"{got * dim * 4 / dt / 1e9:.2f} GB/s of features")
    shutil.rmtree(root, ignore_errors=True)


if __name__ == "__main__":
    main()
