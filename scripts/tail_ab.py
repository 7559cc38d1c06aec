"""Same-box A/B of the fused scan tail: merge of the per-CTA top-k lists through their heads (WB_MERGE_HEADS=1, default)
against the sort-based merge (0), and the rank-counting sort of a CTA's [list | queue] (WB_RANK_SORT=1, default) against
the bitonic sort (0).  Device path of one-query searches (CUDA events around back-to-back wb_search_dev
calls) on flat stores of several sizes - 1.25M x 768 is one GPU's share of BASELINE config 2 at 8 GPUs - and on an IVF
index; results must be byte-identical.  One JSON line per case.
    python scripts/tail_ab.py"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from wise_b200 import _capi, faiss_compat as faiss  # noqa: E402
from bench import fill_index_clustered, make_queries  # noqa: E402

L = _capi.lib()
dev = torch.device("cuda", 0)


def device_ms(idx, qd, k, nprobe, reps):
    D = torch.empty(qd.shape[0], k, device=dev)
    I = torch.empty(qd.shape[0], k, dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    call = lambda: _capi.check(L.wb_search_dev(idx._h, qd.shape[0], qd.data_ptr(), k, nprobe, D.data_ptr(), I.data_ptr(), st))
    for _ in range(10):
        call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        call()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, D.cpu().numpy(), I.cpu().numpy()


def ab(tag, idx, qd, k, nprobe, reps, extra):
    out = {}
    arms = {"old": ("0", "0"), "heads": ("1", "0"), "rank": ("0", "1"), "both": ("1", "1")}
    for rnd in range(2):  # A B A B: drift between the arms shows up as a difference between the rounds
        for name, (heads, rank) in arms.items():
            os.environ["WB_MERGE_HEADS"], os.environ["WB_RANK_SORT"] = heads, rank
            ms, D, I = device_ms(idx, qd, k, nprobe, reps)
            out.setdefault(name, []).append((ms, D, I))
    ref = out["old"][0]
    same = all(np.array_equal(ref[2], r[2]) and np.array_equal(ref[1].view(np.uint32), r[1].view(np.uint32))
               for v in out.values() for r in v)
    print(json.dumps({"case": tag, **extra, "nq": int(qd.shape[0]), "k": k,
                      **{name + "_ms": [round(r[0], 5) for r in v] for name, v in out.items()},
                      "same_bytes": bool(same)}), flush=True)


def main():
    for n, d, reps in ((100_000, 512, 400), (1_250_000, 768, 200), (10_000_000, 768, 40)):
        flat = faiss.IndexIDMap(faiss.IndexFlatIP(d))
        centres, _ = fill_index_clustered(flat, 0, n, d, 50, dev)
        for nq, k in ((1, 100), (4, 100), (1, 10), (1, 250)):
            ab("flat", flat, make_queries(centres, nq, d, 7, dev), k, 1, reps, {"rows": n, "d": d})
        del flat
        torch.cuda.empty_cache()
    n, d, nlist = 2_000_000, 512, 1024
    ivf = faiss.IndexIVFFlat(faiss.IndexFlatIP(d), d, nlist, faiss.METRIC_INNER_PRODUCT)
    flat = faiss.IndexIDMap(faiss.IndexFlatIP(d))
    centres, _ = fill_index_clustered(flat, 0, nlist, d, 50, dev)
    x, _, _ = flat._export(0, nlist)  # centroids = the first nlist rows (no training: this script times searches)
    ivf.set_centroids(x)
    fill_index_clustered(ivf, 0, n, d, 50, dev)
    for nq in (1, 4):
        for nprobe in (8, 32, 128):
            ab("ivf", ivf, make_queries(centres, nq, d, 8, dev), 100, nprobe, 300, {"rows": n, "d": d, "nlist": nlist, "nprobe": nprobe})


if __name__ == "__main__":
    main()
