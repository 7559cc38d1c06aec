"""CPU model of the merge the last CTA of a scan runs (wise_b200/csrc/merge.cuh block_merge_heads): the top-k of
`nlists` sorted per-CTA lists through their heads.  The model restates heads_plan and the four steps - stage J entries
per list, T0 = k-th largest of the first j entries of every list, candidates = the staged prefix >= T0 of every list,
rank by counting - and checks the facts the device code relies on: T0 is a lower bound of the final k-th key, the
candidates always contain the true top-k unless the overflow flag is raised (then the kernel falls back to the
sort-based merge), ranks of distinct keys are a permutation, and the plan's sizes stay inside the buffers.

Replaces nothing in the reference: this is the merge of faiss's per-thread result heaps [faiss-upstream], reached from
/root/reference/src/index/feature_search_index.py:113."""
import numpy as np
import pytest

CAND_CAP = 1024


def heads_plan(k: int, nlists: int):
    """merge.cuh heads_plan: (j, J, stride) or None when the sort-based merge is used."""
    if k < 1 or k > 128 or nlists < 1 or nlists > 256:
        return None
    j = max(1, min(k, (k + k // 4 + nlists - 1) // nlists))
    if nlists * j > 256:
        return None
    lg = 3
    while (1 << lg) < 4 * j:
        lg += 1
    J = 1 << lg
    return j, J, J | 1


def merge_heads(lists: np.ndarray, k: int):
    """lists: (nlists, k) uint64, each row sorted descending, 0 = empty.  Returns (top-k sorted, ok)."""
    nlists = lists.shape[0]
    plan = heads_plan(k, nlists)
    assert plan is not None
    j, J, _ = plan
    staged = np.zeros((nlists, J), np.uint64)
    staged[:, :min(J, k)] = lists[:, :min(J, k)]        # entries beyond k are staged as empty
    heads = staged[:, :j].reshape(-1)
    nz = heads[heads != 0]
    T0 = np.uint64(0)
    if nz.size >= k:
        T0 = np.sort(nz)[::-1][k - 1]                      # the key with exactly k - 1 larger heads
    cand = []
    overflow = False
    for row in staged:
        c = 0
        while c < J and row[c] != 0 and row[c] >= T0:
            c += 1
        if c == J and J < k:
            overflow = True                                # the list may reach deeper than its staged prefix
        cand.extend(row[:c])
    if len(cand) > CAND_CAP:
        overflow = True
    if overflow:
        return None, False
    cand = np.array(cand, np.uint64)
    out = np.zeros(k, np.uint64)
    ranks = (cand[None, :] > cand[:, None]).sum(axis=1)   # rank = number of larger candidates
    assert np.unique(ranks).size == ranks.size, "distinct keys have distinct ranks"
    for key, r in zip(cand, ranks):
        if r < k:
            out[r] = key
    return out, True


def make_lists(rng, nlists, k, rows_per_list, clustered=0):
    """Distinct random keys dealt to nlists lists; `clustered` > 0 puts the globally best keys into that many lists."""
    total = nlists * rows_per_list
    keys = (rng.permutation(4 * total)[:total].astype(np.uint64) + np.uint64(1)) << np.uint64(20)  # distinct, non-zero
    if clustered:
        keys = np.sort(keys)[::-1]
        per = [list() for _ in range(nlists)]
        for i, key in enumerate(keys):
            per[(i % clustered) if i < 4 * k else int(rng.integers(0, nlists))].append(key)
    else:
        per = np.array_split(rng.permutation(keys), nlists)
    lists = np.zeros((nlists, k), np.uint64)
    for l, p in enumerate(per):
        top = np.sort(np.array(p, np.uint64))[::-1][:k]
        lists[l, :top.size] = top
    return lists


@pytest.mark.parametrize("nlists,k,rows", [(148, 100, 675), (148, 10, 50), (148, 100, 3), (37, 100, 400), (74, 64, 200),
                                           (8, 100, 1000), (2, 100, 500), (1, 5, 3), (256, 128, 130)])
def test_heads_merge_model_equals_a_full_sort(nlists, k, rows):
    rng = np.random.default_rng(nlists * 1000 + k)
    fell_back = 0
    for rep in range(5):
        lists = make_lists(rng, nlists, k, rows)
        want = np.zeros(k, np.uint64)
        allk = np.sort(lists[lists != 0])[::-1][:k]
        want[:allk.size] = allk
        got, ok = merge_heads(lists, k)
        if not ok:
            fell_back += 1
            continue
        assert np.array_equal(got, want)
    assert fell_back == 0, "random placement must not need the fallback"


def test_clustered_winners_raise_the_overflow_flag_or_stay_exact():
    rng = np.random.default_rng(9)
    flagged = 0
    for clustered in (1, 2, 3, 5):
        lists = make_lists(rng, 148, 100, 300, clustered=clustered)
        want = np.sort(lists[lists != 0])[::-1][:100]
        got, ok = merge_heads(lists, 100)
        if ok:
            assert np.array_equal(got, want)
        else:
            flagged += 1
    assert flagged >= 1, "winners held by one or two lists reach deeper than 8 staged entries"


def test_plan_sizes():
    for k in (1, 5, 10, 64, 100, 128):
        for nlists in (1, 2, 8, 37, 74, 148, 256):
            plan = heads_plan(k, nlists)
            if plan is None:
                continue
            j, J, stride = plan
            assert nlists * j >= min(k, nlists * j) and nlists * j <= 256 and J >= max(4 * j, 8) and stride % 2 == 1
            assert nlists * j >= k or j == k, "at least k heads define T0 unless every list is staged whole"
    assert heads_plan(250, 148) is None and heads_plan(100, 300) is None
