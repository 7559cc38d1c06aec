"""FeatureStore readers/writers that keep WISE's on-disk layout, without the `webdataset` package.

Mirrors /root/reference/src/feature/store/{feature_store.py, webdataset_store.py,
numpy_save_store.py, feature_store_factory.py}: same class names, methods, file names and sample
encoding, so stores written by either implementation are read by the other.

  WebdatasetStore : shards `<name>-%06d.tar`; one tar member per sample,
                    `%010d.features.pyd` = pickle.dumps(np.ndarray (m, d))   (webdataset_store.py:33-35,93-99)
  NumpySaveStore  : shards `<name>-%06d.npz` with `feature_id`, `features`    (numpy_save_store.py:80-87)

Differences, all on the safe side: `feature_count` is exact (the reference estimates it from
tar sizes, webdataset_store.py:83-91); NumpySaveStore gains `iter_batch` (create_index needs it,
feature_search_index.py:80); unpickling is restricted to numpy arrays.
"""
from __future__ import annotations

import enum
import glob
import io
import os
import pickle
import random
import tarfile
import time
from pathlib import Path

import numpy as np


class FeatureStore:
    def __init__(self, store_name, store_data_dir):
        raise NotImplementedError

    def add(self, index, features):
        raise NotImplementedError

    def load(self, start_index=0, count=-1):
        raise NotImplementedError


class _NumpyOnlyUnpickler(pickle.Unpickler):
    """np.load(..., allow_pickle=True) equivalent that refuses anything but ndarray pickles."""

    _ALLOWED = {
        ("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"),
        ("numpy", "ndarray"), ("numpy", "dtype"),
        ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar"),
        ("numpy.core.numeric", "_frombuffer"), ("numpy._core.numeric", "_frombuffer"),
        ("_codecs", "encode"),  # protocol <= 2 pickles carry the payload as latin-1 text
    }

    def find_class(self, module, name):
        if (module, name) in self._ALLOWED:
            return super().find_class(module, name)
        raise pickle.UnpicklingError(f"feature store sample references {module}.{name}; only numpy arrays are allowed")


def decode_features(payload: bytes) -> np.ndarray:
    """webdataset_store.py:113,129: np.load(BytesIO(payload), allow_pickle=True) of a `.pyd` member."""
    if payload[:6] == b"\x93NUMPY":
        return np.load(io.BytesIO(payload), allow_pickle=False)
    return _NumpyOnlyUnpickler(io.BytesIO(payload)).load()


def _fast_tar():
    """libwiseb200's host-side shard reader (wb_tar_scan / wb_tar_read) or None if the library is not built.
    Reading shards is host logic, so the python tarfile reader below stays as the portable path."""
    try:
        from . import _capi
        return _capi.lib(), _capi
    except Exception:
        return None


class WebdatasetStore(FeatureStore):
    EXTENSION = "tar"

    def __init__(self, store_name, store_data_dir):
        self.store_name = store_name
        self.store_data_dir = str(store_data_dir)
        self.store_data_filename = os.path.join(self.store_data_dir, self.store_name + "-%06d." + self.EXTENSION)
        self.feature_count = -1
        self.feature_dim = -1
        self._tar = None

    # ---- writing: webdataset.ShardWriter semantics ------------------------------------------------
    def enable_write(self, shard_maxcount, shard_maxsize, verbose=0):
        self.shard_maxcount = shard_maxcount
        self.shard_maxsize = shard_maxsize if shard_maxsize and shard_maxsize > 0 else float("inf")
        self.verbose = verbose
        self._shard = 0
        self._count = 0
        self._size = 0
        self._tar = None

    def _next_stream(self):
        if self._tar is not None:
            self._tar.close()
        fname = self.store_data_filename % self._shard
        if self.verbose:
            print("# writing", fname, self._count)
        self._shard += 1
        self._tar = tarfile.open(fname, "w")  # plain tar, like ShardWriter
        self._count = 0
        self._size = 0

    def add(self, id, features):
        if getattr(self, "shard_maxcount", None) is None:
            raise ValueError("enable_write() must be activated before invoking add() method")
        if self._tar is None or self._count >= self.shard_maxcount or self._size >= self.shard_maxsize:
            self._next_stream()
        data = pickle.dumps(features)  # webdataset's default encoder for the "pyd" extension
        ti = tarfile.TarInfo("%010d" % id + ".features.pyd")
        ti.size = len(data)
        ti.mtime = time.time()
        ti.mode = 0o444
        ti.uname = "bigdata"
        ti.gname = "bigdata"
        self._tar.addfile(ti, io.BytesIO(data))
        self._count += 1
        self._size += len(data) + 512 + (-len(data)) % 512

    def close(self):
        if self._tar is not None:
            self._tar.close()
            self._tar = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- reading ----------------------------------------------------------------------------------
    def shard_files(self):
        """Shards {first..last} in ascending index order (webdataset brace expansion, :53-67)."""
        prefix = os.path.join(self.store_data_dir, self.store_name + "-")
        found = {}
        for fn in glob.iglob(prefix + "*.tar"):
            tok = fn[len(prefix):].split(".tar")[0]
            found[int(tok)] = (tok, fn)
        if not found:
            return []
        lo, hi = min(found), max(found)
        width = len(found[lo][0])
        return [prefix + str(i).zfill(width) + ".tar" for i in range(lo, hi + 1)]

    def enable_read(self, shard_shuffle=False, shuffle_values=False, shuffle_bufsize=10000):
        self.shard_shuffle = shard_shuffle
        self.shuffle_values = shuffle_values
        self.shuffle_bufsize = shuffle_bufsize
        files = self.shard_files()
        self.feature_count = 0
        self.feature_dim = -1
        self._shard_rows = {}
        fast = _fast_tar() if os.environ.get("WISE_B200_FAST_STORE", "1") != "0" else None
        for fn in files:
            if fast is not None:
                import ctypes as C
                rows, members, d = C.c_int64(), C.c_int64(), C.c_int64()
                if fast[0].wb_tar_scan(fn.encode(), C.byref(rows), C.byref(members), C.byref(d)) == 0:
                    # feature_count counts samples (tar members), like the reference; rows feed the index
                    self.feature_count += members.value
                    self._shard_rows[fn] = rows.value
                    if self.feature_dim < 0 and d.value > 0:
                        self.feature_dim = d.value
                    continue
            with tarfile.open(fn) as tf:
                for m in tf:
                    if not m.isreg():
                        continue
                    if self.feature_dim < 0:
                        self.feature_dim = decode_features(tf.extractfile(m).read()).shape[1]
                    self.feature_count += 1

    def _samples(self):
        files = self.shard_files()
        if self.shard_shuffle:
            files = list(files)
            random.shuffle(files)  # webdataset shuffles shard order with a time-seeded RNG
        for fn in files:
            with tarfile.open(fn) as tf:
                for m in tf:
                    if not m.isreg():
                        continue
                    key, _, ext = os.path.basename(m.name).partition(".")
                    if ext != "features.pyd":
                        continue
                    yield int(key), decode_features(tf.extractfile(m).read())

    def _maybe_shuffled(self):
        it = self._samples()
        if not self.shuffle_values:
            yield from it
            return
        buf = []
        for s in it:  # webdataset .shuffle(bufsize): reservoir-style buffer shuffle
            buf.append(s)
            if len(buf) >= self.shuffle_bufsize:
                yield buf.pop(random.randrange(len(buf)))
        random.shuffle(buf)
        yield from buf

    def __iter__(self):
        yield from self._maybe_shuffled()

    def _decode_shard(self, fn):
        """One whole shard through the C++ reader: (ids int64[n], x float32[n, d]), or None when the shard needs
        the python reader.  The ctypes call releases the GIL, so a background thread can run it."""
        import ctypes as C
        fast = _fast_tar()
        rows = getattr(self, "_shard_rows", {}).get(fn)
        if fast is None or rows is None or self.feature_dim <= 0:
            return None
        ids = np.empty(rows, np.int64)
        x = np.empty((rows, self.feature_dim), np.float32)
        n = C.c_int64()
        rc = fast[0].wb_tar_read(fn.encode(), self.feature_dim, rows, fast[1].ptr(ids), fast[1].ptr(x), C.byref(n))
        return (ids[: n.value], x[: n.value]) if rc == 0 else None

    def decode_shard_into(self, fn, ids_out, x_out):
        """Decode one whole shard with the C++ reader straight into caller-owned buffers (pinned host memory in
        wise_b200.ingest): ids_out int64[cap], x_out float32[cap, d].  Returns the number of rows, or None when
        the shard needs the python reader.  The ctypes call releases the GIL."""
        import ctypes as C
        fast = _fast_tar()
        rows = getattr(self, "_shard_rows", {}).get(fn)
        if fast is None or rows is None or self.feature_dim <= 0 or rows > ids_out.shape[0]:
            return None
        n = C.c_int64()
        rc = fast[0].wb_tar_read(fn.encode(), self.feature_dim, ids_out.shape[0], fast[1].ptr(ids_out), fast[1].ptr(x_out),
                                 C.byref(n))
        return int(n.value) if rc == 0 else None

    def _fast_shards(self):
        """(fn, decoded) per shard, in shard order (shuffled when shard_shuffle is set); shard i+1 is decoded on a
        helper thread while the caller consumes shard i (only for the fp32 layout WISE writes; others yield None)."""
        from concurrent.futures import ThreadPoolExecutor
        files = list(self.shard_files())
        if self.shard_shuffle:
            random.shuffle(files)  # same shard-order shuffle as _samples(); rows stay in order inside a shard
        if not files:
            return
        with ThreadPoolExecutor(max_workers=1) as pool:
            nxt = pool.submit(self._decode_shard, files[0])
            for i, fn in enumerate(files):
                dec = nxt.result()
                if i + 1 < len(files):
                    nxt = pool.submit(self._decode_shard, files[i + 1])
                yield fn, dec

    def _python_shard(self, fn):
        li, lx = [], []
        with tarfile.open(fn) as tf:
            for m in tf:
                key, _, ext = os.path.basename(m.name).partition(".")
                if m.isreg() and ext == "features.pyd":
                    v = decode_features(tf.extractfile(m).read())
                    li.append(int(key)); lx.append(np.squeeze(v, axis=0))
        if not li:
            return np.empty(0, np.int64), np.empty((0, self.feature_dim), np.float32)
        return np.asarray(li, np.int64), np.stack(lx).astype(np.float32)

    def iter_batch(self, batch_size=512, exact=True):
        """(ids int64[b], features float32[b, d]) batches; (1, d) samples squeezed like :131-139.
        Un-shuffled reads go through the C++ shard reader (one pass per shard, no per-vector python objects);
        batches are views into the decoded shard, only a batch that straddles two shards is copied.
        exact=False (index builds, which accept any batch length) ends a batch at every shard boundary instead."""
        if (not self.shuffle_values and getattr(self, "_shard_rows", None)
                and os.environ.get("WISE_B200_FAST_STORE", "1") != "0"):
            carry_i, carry_x = None, None
            for fn, dec in self._fast_shards():
                ids, x = dec if dec is not None else self._python_shard(fn)
                pos = 0
                if carry_i is not None and carry_i.shape[0]:
                    take = min(batch_size - carry_i.shape[0], ids.shape[0])
                    carry_i = np.concatenate([carry_i, ids[:take]])
                    carry_x = np.concatenate([carry_x, x[:take]])
                    pos = take
                    if carry_i.shape[0] < batch_size:
                        continue  # the shard was too small to complete the batch
                    yield carry_i, carry_x
                    carry_i, carry_x = None, None
                full = pos + ((ids.shape[0] - pos) // batch_size) * batch_size
                for s0 in range(pos, full, batch_size):
                    yield ids[s0:s0 + batch_size], x[s0:s0 + batch_size]
                if not exact:
                    if full < ids.shape[0]:
                        yield ids[full:], x[full:]
                    continue
                carry_i, carry_x = ids[full:], x[full:]
            if carry_i is not None and carry_i.shape[0]:
                yield carry_i, carry_x
            return
        ids, vecs = [], []
        for fid, vec in self._maybe_shuffled():
            ids.append(fid)
            vecs.append(np.squeeze(vec, axis=0))
            if len(ids) == batch_size:
                yield np.asarray(ids, np.int64), np.stack(vecs)
                ids, vecs = [], []
        if ids:
            yield np.asarray(ids, np.int64), np.stack(vecs)


class NumpySaveStore(FeatureStore):
    """READ side of /root/reference/src/feature/store/numpy_save_store.py (SURVEY.md section 2 row 5 scopes this store
    as a loader input only): shards `<name>-%06d.npz` holding `feature_id int32[n]` and `features float32[n, d]`
    (:80-87).  Writing stays with WISE's extract-features.py and the reference's own class."""

    def __init__(self, store_name, store_data_dir):
        self.store_name = store_name
        self.store_data_dir = Path(store_data_dir)

    def enable_read(self, shard_shuffle=False, shuffle_values=False, shuffle_bufsize=10000):
        self.shard_shuffle = shard_shuffle
        self.shuffle_values = shuffle_values
        self.shuffle_bufsize = shuffle_bufsize
        pattern = self.store_data_dir / (self.store_name + "-*.npz")
        self.npz_filename_list = list(glob.iglob(pattern.as_posix()))
        if self.shard_shuffle:
            random.shuffle(self.npz_filename_list)
        else:
            self.npz_filename_list.sort()
        self.feature_count = 0
        self.feature_dim = -1
        for fn in self.npz_filename_list:
            with np.load(fn) as payload:
                self.feature_count += payload["feature_id"].shape[0]
                if self.feature_dim < 0:
                    f0 = payload["features"][0]
                    if f0.ndim not in (1, 2):
                        raise ValueError(f"unrecognized feature shape {f0.shape}")
                    self.feature_dim = f0.shape[-1]

    def _shards(self):
        for fn in self.npz_filename_list:
            with np.load(fn) as payload:
                yield payload["feature_id"], payload["features"]

    def __iter__(self):
        for ids, feats in self._shards():
            n = ids.shape[0]
            order = random.sample(range(n), n) if self.shuffle_values else range(n)
            for i in order:
                yield ids[i], np.take(feats, [i], 0)  # (1, d), like the reference

    def iter_batch(self, batch_size=512):
        """Not in the reference (its create_index could only consume WebdatasetStore); same contract."""
        for ids, feats in self._shards():
            feats = feats.reshape(ids.shape[0], -1)
            for s in range(0, ids.shape[0], batch_size):
                yield ids[s:s + batch_size].astype(np.int64), np.ascontiguousarray(feats[s:s + batch_size], np.float32)

    def enable_write(self, shard_maxcount, shard_maxsize, verbose=0):
        raise NotImplementedError("wise_b200 only reads NumpySaveStore shards; WISE's own class writes them")

    def add(self, id, features):
        raise NotImplementedError("wise_b200 only reads NumpySaveStore shards; WISE's own class writes them")


class FeatureStoreType(str, enum.Enum):
    WEBDATASET = "webdataset"
    NUMPY = "numpy"


class FeatureStoreFactory:
    @classmethod
    def create_store(cls, feature_store_type, media_type, features_dir):
        if feature_store_type == FeatureStoreType.WEBDATASET:
            return WebdatasetStore(media_type, features_dir)
        if feature_store_type == FeatureStoreType.NUMPY:
            return NumpySaveStore(media_type, features_dir)
        raise ValueError(f"unknown feature_store_type {feature_store_type}")

    @classmethod
    def load_store(cls, media_type, features_dir):
        """Infer the store type from the shard extension (feature_store_factory.py:23-38)."""
        features_dir = Path(features_dir)
        exts = []
        for fn in glob.iglob((features_dir / (media_type + "-*.*")).as_posix()):
            suffix = Path(fn).suffix
            if suffix not in exts:
                exts.append(suffix)
        if len(exts) != 1:
            raise ValueError(f"failed to infer type of {media_type} feature store in {features_dir}")
        if exts[0] == ".tar":
            return WebdatasetStore(media_type, features_dir)
        if exts[0] == ".npz":
            return NumpySaveStore(media_type, features_dir)
        raise ValueError(f"unknown store containing shard filenames with extension {exts[0]}")
