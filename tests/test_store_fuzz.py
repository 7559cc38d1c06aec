"""Robustness of the C++ shard reader (host code, no GPU): truncated and corrupted shards must yield an error code or
a consistent decode - never a crash or a write past the caller's buffers.  Runs in a subprocess so that a segfault is
a test failure, not the end of the pytest session."""
import subprocess
import sys
import textwrap

import numpy as np

from wise_b200.store import WebdatasetStore


def _write(tmp_path, n=400, d=20):
    x = np.random.default_rng(0).standard_normal((n, d)).astype(np.float32)
    w = WebdatasetStore("video", tmp_path)
    w.enable_write(100000, 1 << 30)
    for i in range(n):
        w.add(i, x[i:i + 1])
    w.close()
    return tmp_path / "video-000000.tar"


FUZZ = textwrap.dedent("""
    import ctypes as C, os, sys
    import numpy as np
    sys.path.insert(0, {root!r})
    from wise_b200 import _capi
    L = _capi.lib()
    src = open({shard!r}, "rb").read()
    rng = np.random.default_rng(1)
    n, d, guard = 400, 20, 64
    i0 = src.find(b"PaxHeader")
    stride = src.find(b"PaxHeader", i0 + 1) - i0   # pax header block + its records + member header + padded payload
    assert stride > 0 and stride % 512 == 0
    bad = ok = 0
    for trial in range(300):
        b = bytearray(src)
        kind = trial % 6
        if kind == 0:
            b = b[: int(rng.integers(0, len(b)))]                      # truncation anywhere
        elif kind == 1:
            for _ in range(int(rng.integers(1, 8))):                    # random byte flips
                b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        elif kind == 2:
            j = int(rng.integers(0, n)) * stride + 1024 + 124             # a member's size field
            b[j:j + 11] = (b"%011o" % int(rng.integers(0, 1 << 33)))
        elif kind == 3:
            j = int(rng.integers(0, n)) * stride + 1024 + 512             # inside a pickle prologue
            b[j + int(rng.integers(0, 140))] = int(rng.integers(0, 256))
        elif kind == 4:
            j = int(rng.integers(0, n)) * stride + 1024 + 124           # GNU base-256 size, up to 2^64 and beyond
            b[j] = 0x80 | int(rng.integers(0, 2))
            b[j + 1:j + 12] = bytes(int(v) for v in rng.integers(0, 256, 11))
        else:
            cut = int(rng.integers(0, n)) * stride                        # drop a whole block somewhere
            b = b[:cut] + b[cut + 512:]
        fn = {shard!r} + ".fuzz"
        open(fn, "wb").write(bytes(b))
        rows, members, dd = C.c_int64(), C.c_int64(), C.c_int64()
        rc = L.wb_tar_scan(fn.encode(), C.byref(rows), C.byref(members), C.byref(dd))
        if rc != 0:
            bad += 1
            continue
        cap = rows.value
        assert 0 <= cap <= 4 * n and dd.value > 0, (cap, dd.value)
        ids = np.full(cap + guard, -12345, np.int64)
        out = np.full((cap + guard) * dd.value, 7.25, np.float32)
        got = C.c_int64()
        rc = L.wb_tar_read(fn.encode(), dd.value, cap, _capi.ptr(ids), _capi.ptr(out), C.byref(got))
        assert np.all(ids[cap:] == -12345) and np.all(out[cap * dd.value:] == 7.25), "write past the caller's buffer"
        if rc == 0:
            assert 0 <= got.value <= cap  # the scan samples 64 members: its count is an upper bound, the read is exact
            ok += 1
        else:
            bad += 1
    os.remove(fn)
    print("fuzz done", ok, bad)
    assert ok > 0 and bad > 0
""")


def test_shard_reader_survives_corrupt_shards(tmp_path):
    import os
    shard = _write(tmp_path)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", FUZZ.format(root=root, shard=str(shard))], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout[-2000:], r.stderr[-2000:])
    assert "fuzz done" in r.stdout
