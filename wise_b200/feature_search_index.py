"""FeatureSearchIndex on the B200 backend: the index wrapper WISE's CLI and REST API hold.

Mirrors /root/reference/src/index/feature_search_index.py:13-114 (same constructor, methods,
file naming and build/query recipe) and /root/reference/src/index/search_index.py:8-24, with
`faiss` replaced by wise_b200.faiss_compat.  create-index.py, search.py and api/routes.py call
exactly these methods and poke `.index` / `.feature_extractor` directly.
"""
from __future__ import annotations

import itertools
import math
import os
from pathlib import Path

import numpy as np

from . import faiss_compat as faiss
from .store import FeatureStoreFactory, WebdatasetStore


class SearchIndex:
    """The 6-method contract of /root/reference/src/index/search_index.py."""

    def __init__(self, media_type, asset_id, assets):
        raise NotImplementedError

    def get_index_filename(self, index_type):
        raise NotImplementedError

    def create_index(self, index_type, overwrite=False):
        raise NotImplementedError

    def is_index_loaded(self):
        raise NotImplementedError

    def load_index(self, index_type):
        raise NotImplementedError

    def search(self, media_type, query, topk=5, query_type="text"):
        raise NotImplementedError


def _default_extractor_factory(extractor_id):
    """The reference resolves extractors with src.feature.FeatureExtractorFactory (open_clip / CLAP).
    Model inference is out of scope here: use WISE's own factory if it is importable."""
    try:
        from src.feature.feature_extractor_factory import FeatureExtractorFactory  # type: ignore
    except Exception as e:  # pragma: no cover - depends on the host application
        raise RuntimeError(
            "no feature extractor factory: pass feature_extractor_factory=... or run inside a WISE checkout") from e
    return FeatureExtractorFactory(extractor_id)


def ivf_cell_and_train_count(feature_count: int):
    """nlist / training-sample rule of feature_search_index.py:55-59."""
    if feature_count < 200000:
        cell_count = 3 * round(math.sqrt(feature_count))
    else:
        cell_count = 10 * round(math.sqrt(feature_count))
    return cell_count, min(feature_count, 100 * cell_count)


class FeatureSearchIndex(SearchIndex):
    def __init__(self, media_type, asset_id, asset, feature_extractor_factory=None, verbose=True):
        self.media_type = media_type
        self.feature_extractor_id = asset_id
        assert "features_dir" in asset, "features_dir missing in assets"
        self.features_dir = Path(asset["features_dir"])
        assert "index_dir" in asset, "index_dir missing in assets"
        self.index_dir = Path(asset["index_dir"])
        self.prompt = {
            "image": "This is a photo of a ",
            "video": "This is a photo of a ",
            "audio": "this is the sound of ",
        }
        self._extractor_factory = feature_extractor_factory or _default_extractor_factory
        self._verbose = verbose

    def _say(self, msg):
        if self._verbose:
            print(msg)

    def get_index_filename(self, index_type):
        return self.index_dir / (self.media_type + "-" + index_type + ".faiss")

    def create_index(self, index_type, overwrite=False):
        self.index_dir.mkdir(parents=True, exist_ok=True)
        index_fn = self.get_index_filename(index_type)
        if index_fn.exists() and overwrite is False:
            self._say(f"{index_type} for {self.media_type} already exists")
            return
        if index_type not in ("IndexFlatIP", "IndexIVFFlat"):
            raise ValueError(f"unsupported index type {index_type}")
        self.index_type = index_type

        feature_store = FeatureStoreFactory.load_store(self.media_type, self.features_dir)
        feature_store.enable_read(shard_shuffle=False)
        feature_count = feature_store.feature_count
        feature_dim = feature_store.feature_dim

        index = faiss.IndexFlatIP(feature_dim)
        if index_type == "IndexFlatIP":
            index = faiss.IndexIDMap(index)  # IndexFlatIP has no add_with_ids (reference :48-52)
        if index_type == "IndexIVFFlat":
            quantizer = index
            cell_count, train_count = ivf_cell_and_train_count(feature_count)
            index = faiss.IndexIVFFlat(quantizer, feature_dim, cell_count, faiss.METRIC_INNER_PRODUCT)
            self._say(f"  loading a random sample of {train_count} features from {feature_count} features ...")
            shuffled = type(feature_store)(self.media_type, self.features_dir)
            if isinstance(feature_store, WebdatasetStore) and getattr(feature_store, "_shard_rows", None):
                # same files: reuse the shard scan instead of walking the store a second time
                shuffled.shard_shuffle, shuffled.shuffle_values, shuffled.shuffle_bufsize = True, False, 10000
                shuffled.feature_count, shuffled.feature_dim = feature_count, feature_dim
                shuffled._shard_rows = dict(feature_store._shard_rows)
            else:
                shuffled.enable_read(shard_shuffle=True)
            train_features = np.ndarray((train_count, feature_dim), dtype=np.float32)
            if isinstance(shuffled, WebdatasetStore):
                # same sample as the reference's islice over the shard-shuffled stream (:66-71), read shard-wise
                got = 0
                for _, xb in shuffled.iter_batch(batch_size=65536, exact=False):
                    take = min(xb.shape[0], train_count - got)
                    train_features[got:got + take] = xb[:take]
                    got += take
                    if got == train_count:
                        break
                assert got == train_count
            else:
                for i, (_, vec) in enumerate(itertools.islice(shuffled, train_count)):
                    train_features[i, :] = vec
            assert not index.is_trained
            self._say(f"  training {index_type} index with {train_count} features with {cell_count} clusters ...")
            index.train(train_features)
            assert index.is_trained

        self._say("Adding feature vectors to index")
        index.reserve(feature_count)
        if (isinstance(feature_store, WebdatasetStore) and getattr(feature_store, "_shard_rows", None)
                and os.environ.get("WISE_B200_PIPELINED_INGEST", "1") != "0"):
            # shards -> pinned ring -> HBM: decode of shard i+1 overlaps the copy and the IVF assignment of shard i
            from .ingest import add_store_pipelined
            add_store_pipelined(index, feature_store)
        else:
            # the reference adds 512 rows per call (iter_batch default); larger batches amortise the call overhead
            batches = (feature_store.iter_batch(batch_size=65536, exact=False) if isinstance(feature_store, WebdatasetStore)
                       else feature_store.iter_batch(batch_size=65536))
            for ids_batch, vectors_batch in batches:
                index.add_with_ids(np.ascontiguousarray(vectors_batch, np.float32), np.ascontiguousarray(ids_batch, np.int64))
        faiss.write_index(index, index_fn.as_posix())
        self._say(f"  saved index to {index_fn}")

    def is_index_loaded(self):
        return hasattr(self, "index")

    def load_index(self, index_type):
        index_fn = self.get_index_filename(index_type)
        if not index_fn.exists():
            self._say(f"  index {index_fn} does not exist")
            self._say("  use create-index.py script to create an index")
            return False  # the reference has a bare `False` here (:95) and then raises from read_index
        self.index = faiss.read_index(index_fn.as_posix(), faiss.IO_FLAG_READ_ONLY)
        self.feature_extractor = self._extractor_factory(self.feature_extractor_id)
        return True

    def search(self, media_type, query, topk=5, query_type="text"):
        if query_type != "text":
            raise ValueError(f"query_type={query_type} not implemented")
        if media_type == "audio":
            if isinstance(query, str):
                media_query_text = [query]
            else:
                media_query_text = [(self.prompt[media_type] + x) for x in query]
        else:
            media_query_text = [(self.prompt[media_type] + query)]
        query_features = self.feature_extractor.extract_text_features(media_query_text)
        dist, ids = self.index.search(np.ascontiguousarray(query_features, np.float32), topk)
        return dist[0], ids[0]


def SearchIndexFactory(media_type, asset_id, asset, **kw):
    """/root/reference/src/index/search_index_factory.py:4-21 for the feature media types."""
    if media_type in ("audio", "video", "image"):
        return FeatureSearchIndex(media_type, asset_id, asset, **kw)
    raise ValueError(f"Unknown media_type {media_type} (the metadata index is SQLite, not this backend)")
