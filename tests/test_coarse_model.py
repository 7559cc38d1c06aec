"""CPU model of the selection the fused coarse quantizer runs in every CTA (wise_b200/csrc/coarse.cuh
block_select_top_keys): range-adaptive radix selection of the top-`need` (score, ~list id) keys.  The model restates
the kernel's arithmetic pass by pass (bin width, boundary bin, early exit, the recursion into a bin) so that the
invariants the device code relies on - a range always holds >= need keys, the loop ends within 8 passes, exactly `need`
keys come out, ties go to the LOWER list id - are checked here, without a GPU.

Replaces nothing in the reference: the selection is the top-nprobe of faiss's quantizer->search at the top of
IndexIVF::search [faiss-upstream], reached from /root/reference/src/index/feature_search_index.py:113."""
import zlib

import numpy as np
import pytest

BINS = 1024
RESOLVE_MAX = 256


def order_f32(x: np.ndarray) -> np.ndarray:
    """common.cuh order_f32: monotone map float32 -> uint32 (NaN -> 0 like coarse_score_slice)."""
    x = np.asarray(x, np.float32) + np.float32(0.0)
    u = x.view(np.uint32)
    o = np.where(u >> 31 != 0, ~u, u ^ np.uint32(0x80000000)).astype(np.uint32)
    return np.where(np.isnan(x), np.uint32(0), o).astype(np.uint32)


def key64(okeys: np.ndarray) -> list[int]:
    return [(int(k) << 32) | (0xFFFFFFFF - i) for i, k in enumerate(okeys)]


def select_top_keys(okeys: np.ndarray, need0: int) -> tuple[list[int], int]:
    """The kernel's loop on python ints; returns (selected keys in arbitrary order, passes)."""
    keys = key64(okeys)
    lo = int(okeys.min()) << 32
    hi = (int(okeys.max()) << 32) | 0xFFFFFFFF
    need = need0
    sel: list[int] = []
    for p in range(8):
        rng = hi - lo
        sh = 0 if rng < BINS else rng.bit_length() - 10
        assert (rng >> sh) < BINS
        inside = [k for k in keys if lo <= k <= hi]
        assert len(inside) >= need >= 1, "a range holds at least `need` keys"
        hist = [0] * BINS
        for k in inside:
            hist[(k - lo) >> sh] += 1
        # largest bin b with (keys in bins >= b) >= need
        suffix = 0
        bstar = None
        for b in range(BINS - 1, -1, -1):
            if suffix + hist[b] >= need:
                bstar = b
                break
            suffix += hist[b]
        above, cnt = suffix, hist[bstar]
        assert above < need
        last = cnt == need - above
        for k in inside:
            b = (k - lo) >> sh
            if b > bstar or (last and b == bstar):
                sel.append(k)
        if last:
            return sel, p + 1
        if cnt <= RESOLVE_MAX:  # small boundary bin: the best `need - above` of its keys by rank counting, no further pass
            binkeys = [k for k in inside if (k - lo) >> sh == bstar]
            assert len(binkeys) == cnt
            want = need - above
            for k in binkeys:
                if sum(1 for u in binkeys if u > k) < want:
                    sel.append(k)
            return sel, p + 1
        assert sh > 0, "a width-1 bin holds one key (keys are distinct)"
        nlo = lo + (bstar << sh)
        nhi = min(hi, nlo + (1 << sh) - 1)
        lo, hi, need = nlo, nhi, need - above
    raise AssertionError("selection did not end within 8 passes")


def expected(okeys: np.ndarray, need: int) -> list[int]:
    return sorted(key64(okeys), reverse=True)[:need]


CASES = {
    "gaussian_4096": lambda r: r.standard_normal(4096).astype(np.float32) * 0.05,
    "gaussian_31620": lambda r: r.standard_normal(31620).astype(np.float32) * 0.04,
    "small_37": lambda r: r.standard_normal(37).astype(np.float32),
    "all_equal": lambda r: np.zeros(4096, np.float32),
    "three_values": lambda r: r.choice(np.array([-0.25, 0.0, 0.5], np.float32), 5000),
    "with_nan": lambda r: np.where(r.random(2048) < 0.1, np.nan, r.standard_normal(2048)).astype(np.float32),
    "huge_range": lambda r: (r.standard_normal(3000) * 1e30).astype(np.float32),
    "denormal_ties": lambda r: np.concatenate([np.full(100, 1e-45, np.float32), np.full(100, -0.0, np.float32),
                                               np.full(100, 0.0, np.float32)]),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_selection_model_matches_sort(name):
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    okeys = order_f32(CASES[name](rng))
    n = okeys.size
    for need in sorted({1, 2, 8, 32, 100, 128, min(2048, n), n} & set(range(1, n + 1))):
        sel, passes = select_top_keys(okeys, need)
        assert len(sel) == need
        assert sorted(sel, reverse=True) == expected(okeys, need), (name, need)
        assert passes <= 7


def test_ties_go_to_the_lower_list_id():
    okeys = order_f32(np.zeros(300, np.float32))
    sel, _ = select_top_keys(okeys, 10)
    ids = sorted(0xFFFFFFFF - (k & 0xFFFFFFFF) for k in sel)
    assert ids == list(range(10))


def test_gaussian_scores_take_one_pass():
    rng = np.random.default_rng(5)
    passes = []
    for _ in range(20):
        okeys = order_f32(rng.standard_normal(4096).astype(np.float32) * 0.05)
        passes.append(select_top_keys(okeys, 32)[1])
    assert max(passes) == 1, passes


def test_large_boundary_bins_still_recurse():
    okeys = order_f32(np.zeros(4096, np.float32))  # every key in one bin of 4096 > RESOLVE_MAX
    sel, passes = select_top_keys(okeys, 10)
    assert passes >= 2 and sorted(0xFFFFFFFF - (k & 0xFFFFFFFF) for k in sel) == list(range(10))
