"""FeatureStore layout tests, mirroring /root/reference/src/feature/store/test_feature_store.py,
plus the fixture written by the reference's own NumpySaveStore (tests/golden/make_golden.py)."""
import itertools
import os
import pickle
import tarfile

import numpy as np
import pytest

from wise_b200.store import FeatureStoreFactory, NumpySaveStore, WebdatasetStore, decode_features

GOLD = os.path.join(os.path.dirname(__file__), "golden")
A = np.array([[1, 2, 3, 4]])
B = np.array([[5, 6, 7, 8]])
Cc = np.array([[9, 10, 11, 12]])


def test_reads_reference_written_numpy_store():
    exp = np.load(os.path.join(GOLD, "numpy_store_expected.npz"))
    s = FeatureStoreFactory.load_store("video", os.path.join(GOLD, "numpy_store"))
    assert isinstance(s, NumpySaveStore)
    s.enable_read()
    assert s.feature_count == int(exp["feature_count"]) == 7 and s.feature_dim == int(exp["feature_dim"]) == 8
    ids, feats = zip(*[(int(i), v) for i, v in s])
    assert list(ids) == exp["ids"].tolist()
    assert np.array_equal(np.concatenate(feats, 0), exp["features"])
    bi, bf = zip(*s.iter_batch(4))
    assert np.array_equal(np.concatenate(bi), exp["ids"]) and np.array_equal(np.concatenate(bf), exp["features"])
    assert bi[0].dtype == np.int64 and bf[0].dtype == np.float32


def test_numpy_save_store_reads_reference_layout(tmp_path):  # reader side of the reference's test_numpy_save_store
    """Shards in the layout numpy_save_store.py:80-87 writes (feature_id int32[n], features float32[n, d]; the fixture
    under tests/golden/ is written by the reference's own class): read back sample by sample and in batches."""
    seq = [A, B, Cc, Cc, B, A, B]
    for shard, lo in enumerate(range(0, len(seq), 3)):
        part = seq[lo:lo + 3]
        np.savez(tmp_path / ("test-store-%06d" % shard), feature_id=np.arange(lo, lo + len(part), dtype=np.int32),
                 features=np.concatenate(part).astype(np.float32))
    r = NumpySaveStore("test-store", tmp_path)
    r.enable_read()
    loaded = {int(i): v for i, v in r}
    assert r.feature_count == 7 and r.feature_dim == 4
    for i, f in enumerate(seq):
        assert np.all(np.equal(loaded[i], f))
    bi, bf = zip(*r.iter_batch(2))
    assert np.array_equal(np.concatenate(bi), np.arange(7)) and np.array_equal(np.concatenate(bf), np.concatenate(seq))
    with pytest.raises(NotImplementedError):
        r.enable_write(3, -1)


def test_webdataset_store_batch_write(tmp_path):  # reference test_webdataset_store_batch_write
    f0 = np.concatenate((A, B, Cc), axis=0)
    f3 = np.concatenate((Cc, B, A), axis=0)
    w = WebdatasetStore("test-store", tmp_path)
    w.enable_write(3, 256)
    w.add(0, f0)
    w.add(3, f3)
    w.close()
    r = WebdatasetStore("test-store", tmp_path)
    r.enable_read(shard_shuffle=False, shuffle_values=False)
    got = {int(i): v for i, v in r}
    assert np.all(np.equal(got[0], f0)) and np.all(np.equal(got[3], f3))


def test_webdataset_store_read_order(tmp_path):  # reference test_webdataset_store_read_order
    f0 = np.concatenate((A, B, Cc), axis=0)
    f3 = np.concatenate((Cc, B, A), axis=0)
    w = WebdatasetStore("test-store", tmp_path)
    w.enable_write(3, 256, verbose=0)
    for i, f in ((0, f0), (3, f3), (6, A), (7, B), (8, Cc)):
        w.add(i, f)
    w.close()
    r = WebdatasetStore("test-store", tmp_path)
    r.enable_read(shard_shuffle=False, shuffle_values=False)
    assert [int(i) for i, _ in r] == [0, 3, 6, 7, 8]
    assert r.feature_count == 5 and r.feature_dim == 4


def test_webdataset_member_layout_and_batches(tmp_path):
    """`%010d.features.pyd` members holding pickle.dumps(ndarray (1,d) f32) - webdataset_store.py:93-99."""
    rng = np.random.default_rng(0)
    x = rng.standard_normal((1100, 16)).astype(np.float32)
    w = WebdatasetStore("video", tmp_path)
    w.enable_write(500, 1 << 30)
    for i in range(x.shape[0]):
        w.add(i + 1, x[i:i + 1])
    w.close()
    files = sorted(os.listdir(tmp_path))
    assert files == ["video-000000.tar", "video-000001.tar", "video-000002.tar"]
    with tarfile.open(tmp_path / files[0]) as tf:
        m = tf.getmembers()[0]
        assert m.name == "0000000001.features.pyd"
        v = pickle.loads(tf.extractfile(m).read())
        assert v.shape == (1, 16) and v.dtype == np.float32
    s = FeatureStoreFactory.load_store("video", tmp_path)
    assert isinstance(s, WebdatasetStore)
    s.enable_read()
    assert (s.feature_count, s.feature_dim) == (1100, 16)
    sizes, ids_all, xs = [], [], []
    for ids, feats in s.iter_batch():
        sizes.append(len(ids)); ids_all.append(ids); xs.append(feats)
    assert sizes == [512, 512, 76]
    assert np.array_equal(np.concatenate(ids_all), np.arange(1, 1101)) and np.array_equal(np.concatenate(xs), x)


def test_unpickler_rejects_non_numpy():
    class Evil:
        def __reduce__(self):
            return (os.system, ("true",))
    with pytest.raises(pickle.UnpicklingError):
        decode_features(pickle.dumps(Evil()))


def test_factory_rejects_mixed_or_missing(tmp_path):
    with pytest.raises(ValueError):
        FeatureStoreFactory.load_store("video", tmp_path)


# ---- C++ shard reader (wb_tar_scan / wb_tar_read): host code, runs without a GPU ---------------------
def _write_store(tmp_path, n, d, maxcount, protocol=None, dtype=np.float32):
    rng = np.random.default_rng(1)
    x = rng.standard_normal((n, d)).astype(dtype)
    w = WebdatasetStore("video", tmp_path)
    w.enable_write(maxcount, 1 << 30)
    for i in range(n):
        w.add(3 * i + 1, x[i:i + 1])
    w.close()
    return x


def test_fast_reader_matches_python_reader(tmp_path, monkeypatch):
    x = _write_store(tmp_path, 2500, 48, 1000)
    fast = WebdatasetStore("video", tmp_path)
    fast.enable_read()
    assert fast._shard_rows and sum(fast._shard_rows.values()) == 2500
    fi, fx = zip(*fast.iter_batch())
    monkeypatch.setenv("WISE_B200_FAST_STORE", "0")
    slow = WebdatasetStore("video", tmp_path)
    slow.enable_read()
    assert not slow._shard_rows and (slow.feature_count, slow.feature_dim) == (fast.feature_count, fast.feature_dim) == (2500, 48)
    si, sx = zip(*slow.iter_batch())
    assert [len(b) for b in fi] == [len(b) for b in si] == [512, 512, 512, 512, 452]
    assert np.array_equal(np.concatenate(fi), np.concatenate(si)) and np.array_equal(np.concatenate(fx), np.concatenate(sx))
    assert np.array_equal(np.concatenate(fx), x) and np.array_equal(np.concatenate(fi), 3 * np.arange(2500) + 1)


def test_fast_reader_c_abi_directly(tmp_path):
    import ctypes as C
    from wise_b200 import _capi
    x = _write_store(tmp_path, 300, 16, 1000)
    L = _capi.lib()
    fn = str(tmp_path / "video-000000.tar").encode()
    rows, members, d = C.c_int64(), C.c_int64(), C.c_int64()
    assert L.wb_tar_scan(fn, C.byref(rows), C.byref(members), C.byref(d)) == 0
    assert (rows.value, members.value, d.value) == (300, 300, 16)
    ids = np.empty(300, np.int64); out = np.empty((300, 16), np.float32); n = C.c_int64()
    assert L.wb_tar_read(fn, 16, 300, _capi.ptr(ids), _capi.ptr(out), C.byref(n)) == 0
    assert n.value == 300 and np.array_equal(out, x) and ids[7] == 22
    assert L.wb_tar_read(fn, 16, 10, _capi.ptr(ids), _capi.ptr(out), C.byref(n)) != 0  # buffer too small
    assert L.wb_tar_scan(str(tmp_path / "nope.tar").encode(), C.byref(rows), C.byref(members), C.byref(d)) != 0


def test_fast_reader_falls_back_on_foreign_samples(tmp_path):
    """float64 arrays (the reference's own unit test writes int arrays) are not the fp32 fast path: python reads them."""
    w = WebdatasetStore("test-store", tmp_path)
    w.enable_write(3, 256)
    f0 = np.concatenate((A, B, Cc), axis=0)
    for i, f in ((0, f0), (3, f0[::-1].copy()), (6, A)):
        w.add(i, f)
    w.close()
    r = WebdatasetStore("test-store", tmp_path)
    r.enable_read()
    assert [int(i) for i, _ in r] == [0, 3, 6] and r.feature_dim == 4 and r.feature_count == 3


@pytest.mark.parametrize("protocol", [2, 3, 4, 5])
def test_fast_reader_pickle_protocols(tmp_path, protocol):
    """The scanner understands the ndarray pickles of every protocol a WISE install may have written."""
    import ctypes as C
    import io
    from wise_b200 import _capi
    rng = np.random.default_rng(protocol)
    x = rng.standard_normal((5, 1, 33)).astype(np.float32)
    fn = tmp_path / "video-000000.tar"
    with tarfile.open(fn, "w") as tf:
        for i in range(5):
            data = pickle.dumps(x[i], protocol=protocol)
            ti = tarfile.TarInfo("%010d.features.pyd" % (i + 10))
            ti.size = len(data)
            tf.addfile(ti, io.BytesIO(data))
    L = _capi.lib()
    rows, members, d = C.c_int64(), C.c_int64(), C.c_int64()
    rc = L.wb_tar_scan(str(fn).encode(), C.byref(rows), C.byref(members), C.byref(d))
    if protocol == 2:
        # protocol 2 stores the payload as a latin-1 *text* string (UTF-8 in the file): not raw bytes, so the
        # C++ reader must decline and the python reader takes the shard
        assert rc != 0
        r = WebdatasetStore("video", tmp_path)
        r.enable_read()
        got = np.concatenate([v for _, v in r.iter_batch()])
        assert not r._shard_rows and np.array_equal(got, x[:, 0, :])
        return
    assert rc == 0, L.wb_last_error()
    assert (rows.value, d.value) == (5, 33)
    ids = np.empty(5, np.int64); out = np.empty((5, 33), np.float32); n = C.c_int64()
    assert L.wb_tar_read(str(fn).encode(), 33, 5, _capi.ptr(ids), _capi.ptr(out), C.byref(n)) == 0
    assert np.array_equal(out, x[:, 0, :]) and ids.tolist() == [10, 11, 12, 13, 14]


def _read_c_abi(fn, d, cap):
    import ctypes as C
    from wise_b200 import _capi
    L = _capi.lib()
    rows, members, dd = C.c_int64(), C.c_int64(), C.c_int64()
    assert L.wb_tar_scan(str(fn).encode(), C.byref(rows), C.byref(members), C.byref(dd)) == 0, L.wb_last_error()
    ids = np.full(cap, -7, np.int64); out = np.zeros((cap, d), np.float32); n = C.c_int64()
    assert L.wb_tar_read(str(fn).encode(), d, cap, _capi.ptr(ids), _capi.ptr(out), C.byref(n)) == 0, L.wb_last_error()
    return rows.value, members.value, dd.value, ids[: n.value], out[: n.value]


def test_fast_reader_threads_agree(tmp_path, monkeypatch):
    """The fixed-stride plan decodes a regular shard on several threads (one per 2048 samples, at most 16):
    the result does not depend on the thread count."""
    x = _write_store(tmp_path, 9000, 24, 100000)
    fn = tmp_path / "video-000000.tar"
    got = {}
    for t in ("1", "3", "16"):
        monkeypatch.setenv("WISE_B200_LOADER_THREADS", t)
        got[t] = _read_c_abi(fn, 24, 9000)
        assert got[t][:3] == (9000, 9000, 24)
        assert np.array_equal(got[t][4], x) and np.array_equal(got[t][3], 3 * np.arange(9000) + 1)


def test_fast_reader_irregular_shards_take_the_sequential_walk(tmp_path):
    """Anything that breaks the fixed stride - samples with different row counts, a foreign member between two
    samples - fails the per-member verification of the plan and is decoded by the sequential walker instead."""
    import io
    rng = np.random.default_rng(3)
    parts = [rng.standard_normal((m, 8)).astype(np.float32) for m in (1, 2, 1, 3, 1)]
    fn = tmp_path / "video-000000.tar"
    with tarfile.open(fn, "w") as tf:
        for i, a in enumerate(parts):
            data = pickle.dumps(a)
            ti = tarfile.TarInfo("%010d.features.pyd" % (i * 5))
            ti.size = len(data)
            ti.mtime = 1700000000.25 + i  # float mtime: python writes a pax extended header per member
            tf.addfile(ti, io.BytesIO(data))
            if i == 2:
                junk = b'{"note": "not a feature"}'
                tj = tarfile.TarInfo("%010d.json" % (i * 5))
                tj.size = len(junk)
                tf.addfile(tj, io.BytesIO(junk))
    rows, members, d, ids, out = _read_c_abi(fn, 8, 8)
    assert (rows, members, d) == (8, 5, 8)
    assert np.array_equal(out, np.concatenate(parts)) and ids.tolist() == [0, 5, 5, 10, 15, 15, 15, 20]
    # same key width and payload, but one member is not a feature file: the plan's name check must reject it
    fn2 = tmp_path / "video-000001.tar"
    a = rng.standard_normal((6, 1, 8)).astype(np.float32)
    with tarfile.open(fn2, "w") as tf:
        for i in range(6):
            data = pickle.dumps(a[i])
            ti = tarfile.TarInfo(("%010d.features.pyd" if i != 3 else "%010d.featurez.pyd") % i)
            ti.size = len(data)
            tf.addfile(ti, io.BytesIO(data))
    rows, members, d, ids, out = _read_c_abi(fn2, 8, 6)
    assert (rows, members) == (5, 5) and ids.tolist() == [0, 1, 2, 4, 5]
    assert np.array_equal(out, a[[0, 1, 2, 4, 5], 0, :])


def test_iter_batch_shard_aligned(tmp_path):
    """exact=False never copies across shards: batches end at shard boundaries and still cover every row in order."""
    x = _write_store(tmp_path, 2500, 16, 1000)
    r = WebdatasetStore("video", tmp_path)
    r.enable_read()
    bi, bx = zip(*r.iter_batch(batch_size=600, exact=False))
    assert [len(b) for b in bi] == [600, 400, 600, 400, 500]
    assert np.array_equal(np.concatenate(bx), x) and np.array_equal(np.concatenate(bi), 3 * np.arange(2500) + 1)


def test_shard_shuffled_batches_follow_the_sample_stream(tmp_path):
    """shard_shuffle=True (the IVF training sample of create_index): iter_batch visits the shards in the same
    shuffled order as the per-sample iterator, through the C++ reader."""
    import random
    x = _write_store(tmp_path, 2300, 12, 500)
    a = WebdatasetStore("video", tmp_path)
    a.enable_read(shard_shuffle=True)
    random.seed(77)
    ids_stream = [i for i, _ in itertools.islice(a, 1200)]
    b = WebdatasetStore("video", tmp_path)
    b.enable_read(shard_shuffle=True)
    assert b._shard_rows
    random.seed(77)
    got_i, got_x = [], []
    for bi, bx in b.iter_batch(batch_size=400, exact=False):
        got_i.append(bi); got_x.append(bx)
        if sum(len(g) for g in got_i) >= 1200:
            break
    got_i, got_x = np.concatenate(got_i)[:1200], np.concatenate(got_x)[:1200]
    assert got_i.tolist() == ids_stream
    assert np.array_equal(got_x, x[(got_i - 1) // 3])
