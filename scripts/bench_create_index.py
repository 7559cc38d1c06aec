"""create-index.py's work on a REAL on-disk WebdatasetStore: wall time of FeatureSearchIndex.create_index
for IndexFlatIP and IndexIVFFlat, phase by phase, through the pinned, overlapped shard -> HBM pipeline
(wise_b200/ingest.py).  Replaces /root/reference/src/index/feature_search_index.py:33-85.

    python scripts/bench_create_index.py [--rows 10000000 --dim 768 --dir /dev/shm/wise_b200_store --types IndexFlatIP,IndexIVFFlat]

The store is written in the exact byte layout WebdatasetStore.add produces (one pax header + one ustar header + one
pickled (1, d) float32 array per member, shards of 100k members) by a vectorised writer: a template member comes
from the real writer, then only the key digits, the header checksum and the payload change.  Rows are the
benchmark's clustered unit vectors (bench.RowSource), generated on the GPU.  Prints JSON lines."""
import argparse, json, os, shutil, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=10_000_000)
ap.add_argument("--dim", type=int, default=768)
ap.add_argument("--dir", default="/dev/shm/wise_b200_store")
ap.add_argument("--types", default="IndexFlatIP,IndexIVFFlat")
ap.add_argument("--shard-rows", type=int, default=100_000)
ap.add_argument("--keep", action="store_true")
a = ap.parse_args()

import bench
from wise_b200.store import WebdatasetStore
from wise_b200.feature_search_index import FeatureSearchIndex
from wise_b200 import faiss_compat as faiss


def say(**kw):
    print(json.dumps(kw), flush=True)


def member_template(d):
    """One member exactly as WebdatasetStore.add writes it, and where its key digits / checksum / payload sit."""
    tmp = a.dir + ".tmpl"
    shutil.rmtree(tmp, ignore_errors=True); os.makedirs(tmp)
    w = WebdatasetStore("image", tmp); w.enable_write(10, 0)
    probe = (np.arange(d, dtype=np.float32) + 0.5).reshape(1, d)
    w.add(0, probe); w.add(1, probe); w.close()
    b = open(os.path.join(tmp, "image-000000.tar"), "rb").read()
    shutil.rmtree(tmp)
    pay = b.find(probe.tobytes())
    stride = b.find(probe.tobytes(), pay + 1) - pay
    hdr = b.find(b"0000000000.features.pyd")
    assert pay > 0 and stride > 0 and stride % 512 == 0 and hdr % 512 == 0
    t = np.frombuffer(b[:stride], np.uint8).copy()
    h = t[hdr:hdr + 512].astype(np.int64)
    base = int(h.sum() - h[148:156].sum() + 8 * 32 - h[0:10].sum())  # checksum without the key digits
    return t, stride, hdr, pay, base


def write_store(rows, d):
    shutil.rmtree(a.dir, ignore_errors=True)
    feat = os.path.join(a.dir, "store", "mlfoundation", "openclip", "model", "pretrained", "features")
    os.makedirs(feat)
    tmpl, stride, hdr, pay, base = member_template(d)
    dev = torch.device("cuda", 0)
    src = bench.RowSource(rows, d, 2024, dev)
    t0 = time.time()
    shard, fh, in_shard = 0, None, 0
    pow10 = 10 ** np.arange(9, -1, -1, dtype=np.int64)
    pow8 = 8 ** np.arange(5, -1, -1, dtype=np.int64)
    for s, e, x in src.chunks(0, rows):
        xh = x.cpu().numpy()
        pos = 0
        while pos < e - s:
            if fh is None:
                fh = open(os.path.join(feat, "image-%06d.tar" % shard), "wb"); in_shard = 0
            n = min(a.shard_rows - in_shard, e - s - pos)
            ids = np.arange(s + pos + 1, s + pos + n + 1, dtype=np.int64)  # WISE vector ids start at 1
            M = np.tile(tmpl, (n, 1))
            digs = (ids[:, None] // pow10) % 10
            M[:, hdr:hdr + 10] = (digs + 48).astype(np.uint8)
            chk = base + (digs + 48).sum(axis=1)
            M[:, hdr + 148:hdr + 154] = (((chk[:, None] // pow8) % 8) + 48).astype(np.uint8)
            M[:, pay:pay + d * 4] = xh[pos:pos + n].view(np.uint8)
            fh.write(M.tobytes())
            pos += n; in_shard += n
            if in_shard == a.shard_rows:
                end = fh.tell() + 1024
                fh.write(b"\0" * (1024 + (-end) % 10240)); fh.close(); fh = None; shard += 1
    if fh is not None:
        end = fh.tell() + 1024
        fh.write(b"\0" * (1024 + (-end) % 10240)); fh.close(); shard += 1
    dt = time.time() - t0
    size = sum(os.path.getsize(os.path.join(feat, f)) for f in os.listdir(feat))
    say(phase="write_store", rows=rows, dim=d, shards=shard, bytes=size, seconds=dt, dir=a.dir)
    del src
    torch.cuda.empty_cache()
    return feat


feat = write_store(a.rows, a.dim)
# the python tar reader must accept the vectorised shards too (first shard, first members)
import tarfile
with tarfile.open(os.path.join(feat, "image-000000.tar")) as tf:
    m = tf.next()
    assert m.name == "0000000001.features.pyd", m.name
index_dir = os.path.join(a.dir, "index")
for itype in a.types.split(","):
    fsi = FeatureSearchIndex("image", "mlfoundation/openclip/model/pretrained", {"features_dir": feat, "index_dir": index_dir},
                             feature_extractor_factory=lambda _id: None, verbose=False)
    # phases, replicated from create_index so that each can be timed (the full call is timed below)
    t0 = time.time(); st = WebdatasetStore("image", feat); st.enable_read(shard_shuffle=False); t_scan = time.time() - t0
    say(phase="enable_read", index_type=itype, seconds=t_scan, feature_count=st.feature_count, dim=st.feature_dim)
    torch.cuda.synchronize(); t0 = time.time()
    fsi.create_index(itype, overwrite=True)
    torch.cuda.synchronize(); t_total = time.time() - t0
    fn = fsi.get_index_filename(itype)
    say(phase="create_index", index_type=itype, rows=a.rows, dim=a.dim, seconds=t_total, rows_per_s=a.rows / t_total,
        index_bytes=os.path.getsize(fn), pipelined=os.environ.get("WISE_B200_PIPELINED_INGEST", "1") != "0")
    # add phase alone (shards -> pinned ring -> HBM), on a fresh index
    from wise_b200.ingest import add_store_pipelined
    if itype == "IndexFlatIP":
        idx = faiss.IndexIDMap(faiss.IndexFlatIP(a.dim)); idx.reserve(st.feature_count)
        t0 = time.time(); n = add_store_pipelined(idx, st); t_add = time.time() - t0
        say(phase="add_pipelined", index_type=itype, rows=n, seconds=t_add, rows_per_s=n / t_add,
            tar_GBps=sum(os.path.getsize(f) for f in st.shard_files()) / t_add / 1e9)
        del idx
    # the file loads and finds its own rows
    t0 = time.time(); assert fsi.load_index(itype); t_load = time.time() - t0
    src = bench.RowSource(a.rows, a.dim, 2024, torch.device("cuda", 0))
    probe = [0, a.rows // 2, a.rows - 2]
    chunk0 = next(src.chunks(0, 1))[2]
    q = np.stack([next(src.chunks(i, i + 1))[2][0].cpu().numpy() for i in probe])
    if hasattr(fsi.index, "nprobe"):
        fsi.index.nprobe = 32
    D, I = fsi.index.search(q, 5)
    say(phase="load_and_search", index_type=itype, load_seconds=t_load, top1_ids=I[:, 0].tolist(), expect=[p + 1 for p in probe],
        ok=bool(all(int(I[j, 0]) == probe[j] + 1 or int(I[j, 1]) == probe[j] + 1 for j in range(3))))
    del fsi, src
    torch.cuda.empty_cache()
    os.remove(fn)
if not a.keep:
    shutil.rmtree(a.dir, ignore_errors=True)
