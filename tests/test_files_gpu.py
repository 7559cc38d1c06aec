"""GPU tests of the on-disk contract: .faiss files (byte layout + round trip) and the
FeatureSearchIndex build/load/search recipe over a WISE feature store."""
import struct

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def faiss():
    from wise_b200 import faiss_compat
    return faiss_compat


def test_idmap_file_bytes_and_round_trip(faiss, tmp_path):
    xb = O.unit_gaussian(10, 4, 1)
    ids = np.array([3, 1, 4, 15, 9, 2, 6, 5, 35, 8], np.int64)
    idx = faiss.IndexIDMap(faiss.IndexFlatIP(4))
    idx.add_with_ids(xb, ids)
    fn = str(tmp_path / "image-IndexFlatIP.faiss")
    faiss.write_index(idx, fn)
    hdr = struct.pack("<iqqqBi", 4, 10, 1 << 20, 1 << 20, 1, 0)
    expect = b"IxMp" + hdr + b"IxFI" + hdr + struct.pack("<Q", 40) + xb.tobytes() + struct.pack("<Q", 10) + ids.tobytes()
    assert open(fn, "rb").read() == expect
    back = faiss.read_index(fn, faiss.IO_FLAG_READ_ONLY)
    assert isinstance(back, faiss.IndexIDMap) and back.ntotal == 10 and back.d == 4
    q = O.unit_gaussian(2, 4, 2)
    assert all(np.array_equal(a, b) for a, b in zip(idx.search(q, 5), back.search(q, 5)))


def test_ivf_file_round_trip(faiss, tmp_path):
    n, d, nlist = 3000, 24, 16
    xb = O.clustered_unit(n, d, 20, 3)
    ids = np.arange(n, dtype=np.int64) + 1
    idx = faiss.IndexIVFFlat(faiss.IndexFlatIP(d), d, nlist, faiss.METRIC_INNER_PRODUCT)
    idx.set_centroids(O.kmeans_init(xb, nlist))
    idx.add_with_ids(xb, ids)
    idx.nprobe = 4
    fn = str(tmp_path / "video-IndexIVFFlat.faiss")
    faiss.write_index(idx, fn)
    raw = open(fn, "rb").read()
    assert raw[:4] == b"IwFl" and b"ilar" in raw and b"full" in raw
    back = faiss.read_index(fn, faiss.IO_FLAG_READ_ONLY)
    assert isinstance(back, faiss.IndexIVFFlat) and back.ntotal == n and back.nlist == nlist and back.nprobe == 4
    assert np.array_equal(back.centroids(), idx.centroids())
    q = O.clustered_unit(5, d, 20, 4)
    for nprobe in (1, 4, 16):
        idx.nprobe = back.nprobe = nprobe
        D0, I0 = idx.search(q, 10)
        D1, I1 = back.search(q, 10)
        O.compare_topk(D1, I1, D0, I0)  # same rows, list-major order on disk: only tie order may move
    with pytest.raises(RuntimeError):  # faiss: "could not open ... for reading"
        faiss.read_index(str(tmp_path / "missing.faiss"), 0)


class RandomFeatures:
    """The debug extractor of /root/reference/docs/FeatureExtractor.md: deterministic unit vectors per text."""

    def __init__(self, extractor_id, d=64):
        self.d = d

    def extract_text_features(self, texts):
        out = []
        for t in texts:
            rng = np.random.default_rng(abs(hash(t)) % (2 ** 32))
            v = rng.standard_normal(self.d).astype(np.float32)
            out.append(v / np.linalg.norm(v))
        return np.stack(out)


@pytest.mark.parametrize("index_type", ["IndexFlatIP", "IndexIVFFlat"])
def test_feature_search_index_recipe(faiss, tmp_path, index_type):
    """create-index.py -> search.py flow of the reference, on a synthetic WebdatasetStore."""
    from wise_b200.feature_search_index import FeatureSearchIndex, SearchIndexFactory
    from wise_b200.store import WebdatasetStore
    feats = tmp_path / "store" / "a" / "b" / "c" / "d" / "features"
    feats.mkdir(parents=True)
    n, d = 2600, 64
    x = O.clustered_unit(n, d, 30, 5)
    w = WebdatasetStore("video", feats)
    w.enable_write(1000, 1 << 30)
    for i in range(n):
        w.add(i + 1, x[i:i + 1])
    w.close()
    asset = {"features_dir": feats, "index_dir": tmp_path / "store" / "a" / "b" / "c" / "d" / "index"}
    si = SearchIndexFactory("video", "a/b/c/d", asset, feature_extractor_factory=RandomFeatures, verbose=False)
    assert isinstance(si, FeatureSearchIndex) and not si.is_index_loaded()
    assert si.load_index(index_type) is False
    si.create_index(index_type)
    fn = si.get_index_filename(index_type)
    assert fn.name == f"video-{index_type}.faiss" and fn.exists()
    mtime = fn.stat().st_mtime_ns
    si.create_index(index_type)  # exists + overwrite=False: untouched
    assert fn.stat().st_mtime_ns == mtime
    assert si.load_index(index_type) and si.is_index_loaded()
    assert si.index.ntotal == n
    if index_type == "IndexIVFFlat":
        assert si.index.nlist == 3 * round(np.sqrt(n))
        si.index.nprobe = si.index.nlist
    dist, ids = si.search("video", "cooking", topk=20)
    assert dist.shape == (20,) and ids.shape == (20,)
    qv = RandomFeatures("x").extract_text_features(["This is a photo of a cooking"])
    Dr, Ir = O.flat_search(x, qv, 20, np.arange(1, n + 1))
    O.compare_topk(dist[None], ids[None], Dr, Ir)


def test_pipelined_ingest_equals_batch_loop(faiss, tmp_path, monkeypatch):
    """Shards -> pinned ring -> HBM (wise_b200.ingest, the default of create_index) builds byte-for-byte the same
    index file as the reference-shaped batch loop (feature_search_index.py:78-82), including a shard the C++ reader
    must hand to the python reader (a multi-row sample among single-row ones)."""
    from wise_b200.feature_search_index import FeatureSearchIndex
    from wise_b200.store import WebdatasetStore
    feats = tmp_path / "features"
    feats.mkdir()
    n, d = 5200, 48
    x = O.clustered_unit(n, d, 30, 6)
    w = WebdatasetStore("image", feats)
    w.enable_write(1000, 1 << 30)
    i = 0
    while i < n:
        if i == 2500:  # one irregular member: 3 rows under one key
            w.add(i + 1, x[i:i + 3]); i += 3
        else:
            w.add(i + 1, x[i:i + 1]); i += 1
    w.close()
    files = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("WISE_B200_PIPELINED_INGEST", mode)
        asset = {"features_dir": feats, "index_dir": tmp_path / ("index" + mode)}
        si = FeatureSearchIndex("image", "a/b/c/d", asset, feature_extractor_factory=RandomFeatures, verbose=False)
        si.create_index("IndexFlatIP")
        files[mode] = si.get_index_filename("IndexFlatIP").read_bytes()
        assert si.load_index("IndexFlatIP") and si.index.ntotal == n
    assert files["1"] == files["0"]
    D, I = si.index.search(x[[0, 2501, n - 1]], 3)
    assert I[0, 0] == 1 and I[1, 0] == 2501 and I[2, 0] == n  # rows 2500..2502 share key 2501 (one sample of 3 rows)
