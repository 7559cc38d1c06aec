"""Generates the committed fixtures under tests/golden/.  Run in the build container only
(it imports the reference from /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

1. numpy_store/*.npz + numpy_store_expected.npz : written by the REFERENCE's own NumpySaveStore
   (/root/reference/src/feature/store/numpy_save_store.py) - pins our reader to the reference's
   on-disk layout.  (WebdatasetStore cannot be imported: the `webdataset` package is absent.)
2. flat_small.npz : a small IndexFlatIP case (with duplicate rows) and its oracle answer - a
   regression pin for the two oracle implementations.  NOT a faiss output: parity unpinned.
"""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

REF_STORE = "/root/reference/src/feature/store"


def load_reference_numpy_store():
    pkg = types.ModuleType("refstore")
    pkg.__path__ = [REF_STORE]
    sys.modules["refstore"] = pkg
    spec = importlib.util.spec_from_file_location("refstore.numpy_save_store", os.path.join(REF_STORE, "numpy_save_store.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    return mod.NumpySaveStore


def main():
    NumpySaveStore = load_reference_numpy_store()
    out = os.path.join(HERE, "numpy_store")
    os.makedirs(out, exist_ok=True)
    for f in os.listdir(out):
        os.remove(os.path.join(out, f))
    rng = np.random.default_rng(20240518)
    feats = rng.standard_normal((7, 8)).astype(np.float32)
    feats /= np.linalg.norm(feats, axis=1, keepdims=True)
    ids = np.array([1, 2, 5, 8, 13, 21, 34], np.int64)
    w = NumpySaveStore("video", out)
    w.enable_write(shard_maxcount=3, shard_maxsize=-1, verbose=0)
    for i, v in zip(ids, feats):
        w.add(int(i), v[None, :])
    w.close()
    del w
    r = NumpySaveStore("video", out)
    r.enable_read()
    got_ids, got = [], []
    for fid, vec in r:
        got_ids.append(int(fid))
        got.append(vec)
    np.savez(os.path.join(HERE, "numpy_store_expected.npz"), ids=np.asarray(got_ids, np.int64),
             features=np.concatenate(got, 0), feature_count=r.feature_count, feature_dim=r.feature_dim)

    from oracle import oracle as O
    xb = O.unit_gaussian(300, 32, 11)
    xb[200:210] = xb[0:10]  # exact duplicates: exercises the lowest-position tie rule
    xq = np.concatenate([O.unit_gaussian(3, 32, 12), xb[5:6]])
    ext = np.arange(300, dtype=np.int64) * 7 + 3
    D, I = O.flat_search(xb, xq, 12, ext)
    np.savez(os.path.join(HERE, "flat_small.npz"), xb=xb, xq=xq, ids=ext, k=12, D=D, I=I)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
