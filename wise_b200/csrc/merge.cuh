// merge.cuh - K3 topk_merge: merge per-CTA (or per-GPU) partial top-k lists into the final
// (D, I) rows.  Replaces faiss heap_reorder / HeapResultHandler::end and the parallel_mode-1
// merge of IndexIVF::search [faiss-upstream]; runs inside every index.search()
// (/root/reference/src/index/feature_search_index.py:113, /root/reference/api/routes.py:1407).
// Latency-bound: one CTA per query, bitonic sort in shared memory over rounds of candidates.
#pragma once
#include <float.h>
#include "common.cuh"
#include "radix_select.cuh"

namespace wb {

constexpr int kMergeThreads = 1024;

struct MergeParams {
    int64_t nq;
    int k;
    int64_t nparts;
    int S;  // sort buffer entries: power of two >= k + 2 * kMergeThreads (or >= all candidates)
    // source A: keys [nq][nparts][k] from scan_topk_kernel; positions -> ids via ids/id_base
    const uint64_t* keys;
    const int64_t* ids;  // may be null: id = position
    // source B: (D, I) parts [nparts][nq][k] (multi-GPU merge)
    const float* Dp;
    const int64_t* Ip;
    float* D;    // [nq][k]
    int64_t* I;  // [nq][k]
};

// Candidate c of query q as a sortable key (0 = empty slot).
template <bool FROM_KEYS>
__device__ __forceinline__ uint64_t merge_load(const MergeParams& p, int64_t q, int64_t M, int64_t c) {
    if constexpr (FROM_KEYS) {
        return p.keys[q * M + c];
    } else {
        const int64_t part = c / p.k, slot = c - part * p.k;
        const int64_t src = (part * p.nq + q) * p.k + slot;
        return p.Ip[src] >= 0 ? make_key(p.Dp[src], (uint32_t)c) : 0ull;
    }
}

// Barrier among the NT threads that run a selection: the whole CTA (bar_id < 0) or a named barrier (the consumer
// warps of scan_topk_kernel, whose producer warp has already left).
template <int NT>
__device__ __forceinline__ void sel_sync(int bar_id) {
    if (bar_id < 0) __syncthreads();
    else named_bar_sync(bar_id, NT);
}

// Block-level top-k of M candidates produced by load(i), i in [0, M): leaves the k best keys sorted
// (descending) in buf[0, k).  buf = [ best k | queue ]: round 0 sorts the first S candidates; after
// that a candidate is queued only if it beats the current k-th best, so almost everything dies on
// one compare and the buffer is re-sorted only when the queue could overflow.
// NT threads (tid in [0, NT)) sharing barrier bar_id; S >= k + 2 * NT or S >= M.
template <int NT = kMergeThreads, class Load>
__device__ __forceinline__ void block_select_topk(uint64_t* buf, int S, int k, int64_t M, Load load, int* cnt,
                                                  int tid = threadIdx.x, int bar_id = -1) {
    const int first = (int)min((int64_t)S, M);
    for (int i = tid; i < S; i += NT) buf[i] = i < first ? load((int64_t)i) : 0ull;
    if (tid == 0) *cnt = 0;
    sel_sync<NT>(bar_id);
    // sort only as much as is filled (the rest is zeros = minimal keys): a K2 compaction typically folds a few
    // hundred candidates, and a 4096-key bitonic sort costs ~40 us against ~8 us for 512 keys
    bitonic_sort_desc<NT>(buf, min(S, max(2, pow2_ceil(first))), 1, tid, bar_id);
    const int qcap = S - k;
    for (int64_t base = first; base < M; base += NT) {
        if (base == first)
            for (int i = k + tid; i < S; i += NT) buf[i] = 0ull;
        sel_sync<NT>(bar_id);
        const uint64_t thr = buf[k - 1];
        const int64_t c = base + tid;
        const uint64_t key = c < M ? load(c) : 0ull;
        const bool pass = key > thr;  // also rejects empty slots (0)
        const unsigned m = __ballot_sync(0xffffffffu, pass);
        if (m) {
            int slot0 = 0;
            if ((tid & 31) == 0) slot0 = atomicAdd(cnt, __popc(m));
            slot0 = __shfl_sync(0xffffffffu, slot0, 0);
            if (pass) buf[k + slot0 + __popc(m & ((1u << (tid & 31)) - 1))] = key;
        }
        sel_sync<NT>(bar_id);
        if (*cnt > qcap - NT) {  // uniform: cnt is read after the barrier
            bitonic_sort_desc<NT>(buf, S, 1, tid, bar_id);
            for (int i = k + tid; i < S; i += NT) buf[i] = 0ull;
            if (tid == 0) *cnt = 0;
            sel_sync<NT>(bar_id);
        }
    }
    sel_sync<NT>(bar_id);
    // buf = [k sorted best | *cnt queued | zeros]
    if (*cnt > 0) bitonic_sort_desc<NT>(buf, min(S, max(2, pow2_ceil(k + *cnt))), 1, tid, bar_id);
    sel_sync<NT>(bar_id);
}

// Top-k of `nlists` SORTED (descending) lists of k keys each, load(list, rank) -> key.  Round 0 sorts only
// the best j = S / nlists entries of every list; the k-th best of those is already (almost always) the final
// threshold, and because the lists are sorted one compare per list decides whether anything deeper can
// matter.  Lists that do reach deeper push their surviving tail; if that ever overflows the queue the
// caller falls back to the generic block_select_topk.  Returns false on overflow (uniform).
template <int NT = kMergeThreads, class Load2>
__device__ __forceinline__ bool block_select_topk_lists(uint64_t* buf, int S, int k, int nlists, Load2 load, int* cnt,
                                                        int tid = threadIdx.x, int bar_id = -1) {
    // round 0 depth: ~4x the expected share of a list in the global top-k (+ slack), at most what fits
    const int j_fit = max(1, S / max(nlists, 1));
    const int j0 = min(k, min(j_fit, max(4, (4 * k + nlists - 1) / max(nlists, 1) + 3)));
    const int n0 = nlists * j0;
    if (n0 > S) return false;  // more lists than buffer entries (uniform)
    for (int i = tid; i < S; i += NT) buf[i] = i < n0 ? load(i / j0, i % j0) : 0ull;
    if (tid == 0) *cnt = 0;
    sel_sync<NT>(bar_id);
    int Ps = 2;
    while (Ps < n0) Ps <<= 1;  // sort only as much as is filled
    Ps = max(Ps, 2);
    bitonic_sort_desc<NT>(buf, min(Ps, S), 1, tid, bar_id);
    if (j0 >= k) return true;
    const int qcap = S - k;
    for (int i = k + tid; i < S; i += NT) buf[i] = 0ull;
    sel_sync<NT>(bar_id);
    const uint64_t thr = buf[k - 1];
    for (int l = tid; l < nlists; l += NT) {
        // The entries of a sorted list that beat thr are a prefix [j0, e).  Almost always it is empty (one load).  When
        // the winners cluster in a few lists (an IVF query whose best lists were scanned by a few CTAs, a shard that
        // holds most of the global top-k) e is found by binary search - 7 dependent loads instead of up to k - and the
        // prefix is copied with independent loads.
        if (!(load(l, j0) > thr)) continue;
        int lo = j0 + 1, hi = k;  // every r < lo beats thr; no r >= hi does
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (load(l, mid) > thr) lo = mid + 1;
            else hi = mid;
        }
        const int c = lo - j0;
        const int slot0 = atomicAdd(cnt, c);
        if (slot0 + c <= qcap)
            for (int r = j0; r < lo; ++r) buf[k + slot0 + (r - j0)] = load(l, r);
    }
    sel_sync<NT>(bar_id);
    const int c = *cnt;
    if (c > qcap) return false;
    if (c > 0) {
        int P2 = 2;
        while (P2 < k + c) P2 <<= 1;
        bitonic_sort_desc<NT>(buf, min(P2, S), 1, tid, bar_id);
    }
    sel_sync<NT>(bar_id);
    return true;
}

// ---- merge of sorted lists through their heads ---------------------------------------------------------------
// Top-k of `nlists` SORTED (descending, 0-padded) lists, load(list, rank) -> key, without sorting anything large.
// The k-th largest of the first j entries of every list (nlists * j >= 1.25 k of them) is a lower bound T0 of the final
// k-th key - at least k keys are >= T0 - and a sharp one: with the winners spread over 148 lists, ~1.7 k keys pass it.
// The candidates >= T0 are a short prefix of every list; they are found in the J staged entries of each list, compacted,
// and ranked by counting (rank = number of larger candidates = output position: the keys are distinct).  Three
// broadcast-read loops over a few hundred keys and five barriers replace a 1024-key bitonic sort, per-list boundary
// loads from L2, a binary search and a second sort (12.4 us per search in the last CTA of a scan; NOTES.md).
// Returns false (uniform, nothing written) when a list may reach deeper than its staged prefix or the candidates
// overflow - winners clustered in a few lists - and the caller runs block_select_topk_lists instead.
constexpr int kHeadsCandCap = 1024;
struct HeadsPlan {
    int j, J, lgJ, stride, ok;
    size_t bytes;
};
__host__ __device__ inline HeadsPlan heads_plan(int k, int nlists) {
    HeadsPlan P{};
    if (k < 1 || k > 128 || nlists < 1 || nlists > 256) return P;  // (k = 250: 8 us slower than the sort-based merge)
    int j = (k + k / 4 + nlists - 1) / nlists;  // heads per list that define T0
    if (j < 1) j = 1;
    if (j > k) j = k;
    if ((int64_t)nlists * j > 256) return P;
    int lgJ = 3;                                // staged entries per list: a power of two >= max(4 j, 8) - a list holds
    while ((1 << lgJ) < 4 * j) ++lgJ;           // ~1.7 j candidates on average; with 2 j, 37 lists fell back 1 time in 5
    P.j = j;
    P.lgJ = lgJ;
    P.J = 1 << lgJ;                             // (entries beyond k are staged as empty)
    P.stride = P.J | 1;                         // odd stride: the per-list scans of step 3 spread over the banks
    P.ok = 1;
    // staged[nlists][stride] | heads[nlists * j] | hsorted[k] | cand[cap] | out[k] | overflow flag
    P.bytes = (((size_t)nlists * P.stride + (size_t)nlists * j + kHeadsCandCap + 2 * (size_t)k) * 8 + 64 + 15) & ~(size_t)15;
    return P;
}

template <int NT, bool WARP_RANK, class Load2>
__device__ __forceinline__ bool block_merge_heads(unsigned char* region, int k, int nlists, Load2 load, int* cnt, int tid,
                                                  int bar_id, const uint64_t** best_out,
                                                  unsigned long long* ts = nullptr /* diagnostics: 4 phase stamps */) {
    auto stamp = [&](int i) {
        if (ts && tid == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            ts[i] = t;
        }
    };
    const HeadsPlan hp = heads_plan(k, nlists);
    if (!hp.ok) return false;  // (uniform; the host only sets the mode when the plan is valid)
    const int j = hp.j, J = hp.J, lgJ = hp.lgJ, stride = hp.stride, nh = nlists * j;
    const int warp = tid >> 5, lane = tid & 31;
    uint64_t* staged = reinterpret_cast<uint64_t*>(region);
    uint64_t* heads = staged + (size_t)nlists * stride;
    uint64_t* hsorted = heads + nh;
    uint64_t* cand = hsorted + k;
    uint64_t* out = cand + kHeadsCandCap;
    int* ovf_s = reinterpret_cast<int*>(out + k);                // [0] overflow flag
    // 1. stage the first J entries of every list (the first j of them also densely: the heads).  All loads of a thread
    //    are issued before its first store: the keys come from L2 at ~0.5 us per round trip.
    for (int base = 0; base < (nlists << lgJ); base += NT * 8) {
        uint64_t v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = base + u * NT + tid;
            const int l = i >> lgJ, r = i & (J - 1);
            v[u] = (l < nlists && r < k) ? load(l, r) : 0ull;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = base + u * NT + tid;
            const int l = i >> lgJ, r = i & (J - 1);
            if (l < nlists) {
                staged[l * stride + r] = v[u];
                if (r < j) heads[l * j + r] = v[u];
            }
        }
    }
    for (int i = tid; i < 2 * k; i += NT) {  // hsorted, out
        if (i < k) hsorted[i] = 0ull;
        else out[i - k] = 0ull;
    }
    if (tid == 0) {
        *cnt = 0;
        *ovf_s = 0;
    }
    sel_sync<NT>(bar_id);
    stamp(0);
    // 2. T0 = the k-th largest of the heads (0 if fewer than k heads are non-empty: everything is a candidate)
    block_rank_scatter(heads, nh, hsorted, k, warp, NT / 32);
    sel_sync<NT>(bar_id);
    stamp(1);
    // 3. candidates: the prefix >= T0 of every list (one atomicAdd per warp: the lanes' counts are scanned first)
    const uint64_t T0 = hsorted[k - 1];
    for (int l0 = 0; l0 < nlists; l0 += NT) {  // (uniform trip count: whole warps take part in the shuffles)
        const int l = l0 + tid;
        const uint64_t* row = staged + (l < nlists ? l : 0) * stride;
        int c = 0;
        if (l < nlists) {
            while (c < J && row[c] != 0ull && row[c] >= T0) ++c;
            if (c == J && J < k) *ovf_s = 1;  // the list may hold more candidates than were staged
        }
        int inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        const int wtot = __shfl_sync(0xffffffffu, inc, 31);
        int wbase = 0;
        if (lane == 31 && wtot) wbase = atomicAdd(cnt, wtot);
        wbase = __shfl_sync(0xffffffffu, wbase, 31);
        const int off = wbase + inc - c;
        if (c) {
            if (off + c <= kHeadsCandCap) {
                for (int r = 0; r < c; ++r) cand[off + r] = row[r];
            } else {
                *ovf_s = 1;
            }
        }
    }
    sel_sync<NT>(bar_id);
    stamp(2);
    if (*ovf_s) {
        sel_sync<NT>(bar_id);  // everyone has read the flag before the caller's fallback reuses the region
        return false;
    }
    // 4. rank by counting: position = number of larger candidates.  Two codings, chosen per kernel by measurement
    //    (~170 candidates; NOTES.md section 6): one thread per candidate with broadcast loads - 2.0-2.9 us in the flat
    //    scan kernel but 5.2 us in the gather instantiation - or block_rank_scatter (lanes split the array, a warp
    //    ranks 4 keys per pass) - 4.2 us in the flat kernel, 3.4 us in the gather one.
    if constexpr (WARP_RANK) {
        block_rank_scatter(cand, *cnt, out, k, warp, NT / 32);
    } else {
        const int C = *cnt;
        for (int t = tid; t < C; t += NT) {
            const uint64_t key = cand[t];
            int rank = 0;
#pragma unroll 8
            for (int u = 0; u < C; ++u) rank += cand[u] > key;
            if (rank < k) out[rank] = key;
        }
    }
    sel_sync<NT>(bar_id);
    stamp(3);
    *best_out = out;
    return true;
}

template <bool FROM_KEYS>
__global__ void __launch_bounds__(kMergeThreads) merge_topk_kernel(const MergeParams p) {
    extern __shared__ __align__(16) unsigned char smem_merge[];
    uint64_t* buf = reinterpret_cast<uint64_t*>(smem_merge);
    __shared__ int cnt;
    const int tid = threadIdx.x;
    const int64_t q = blockIdx.x;
    const int k = p.k;
    const int64_t M = p.nparts * (int64_t)k;
    // both sources are nparts sorted lists of k keys
    const bool ok = block_select_topk_lists(
        buf, p.S, k, (int)p.nparts, [&](int l, int r) { return merge_load<FROM_KEYS>(p, q, M, (int64_t)l * k + r); }, &cnt);
    if (!ok) {
        __syncthreads();
        block_select_topk(buf, p.S, k, M, [&](int64_t c) { return merge_load<FROM_KEYS>(p, q, M, c); }, &cnt);
    }
    for (int j = tid; j < k; j += kMergeThreads) {
        const uint64_t key = buf[j];
        float d = -FLT_MAX;
        int64_t id = -1;
        if (key) {
            d = key_score(key);
            const uint32_t pos = key_pos(key);
            if constexpr (FROM_KEYS) {
                id = p.ids ? p.ids[pos] : (int64_t)pos;
            } else {
                const int64_t part = pos / k, slot = pos - part * k;
                id = p.Ip[(part * p.nq + q) * k + slot];
            }
        }
        p.D[q * k + j] = d;
        p.I[q * k + j] = id;
    }
}

// K2 epoch boundary: fold the candidates a query collected during the epoch into its top-k,
// tighten its threshold, reset its candidate counter; on the last epoch also emit (D, I).
struct CompactParams {
    int k, cap, kstride, S;
    uint64_t* keys;   // [nq][kstride]: [0,k) top-k so far (sorted), [k, k+cap) candidates
    int* cnt;         // [nq]
    float* thr;       // [nq_pad]
    const int64_t* ids;
    float* D;         // null unless this is the last epoch
    int64_t* I;
    // filter-and-refine searches (one-term K2 epochs): exact fp32 re-scoring of keys against the row store
    int rescore;        // 0: keys carry final scores; 1: re-score the epoch's candidates BEFORE the selection
                        // (they carry one-term filter scores); 2: re-score the k winners AFTER the selection (first
                        // epoch: selected by 3xTF32 scores) so that every key in the list carries the same arithmetic
    const float* rows;  // [*, ld] row store the key positions index
    const float* q;     // [nq, ld] zero-padded queries
    int ld;
};

// keys[i] <- make_key(<rows[pos(keys[i])], q>, pos) for i in [0, n): one warp per key, fp32 FMA over float4 lanes and
// a fixed-order xor-shuffle reduction - the same row and query always give the same bits (exact duplicates tie).
__device__ __forceinline__ void rescore_keys(uint64_t* keys, int n, const float* rows, const float* qrow, int ld,
                                             int warp, int nwarps) {
    const int lane = threadIdx.x & 31;
    const float4* qv = reinterpret_cast<const float4*>(qrow);
    const int ld4 = ld >> 2;
    for (int i = warp; i < n; i += nwarps) {
        const uint64_t key = keys[i];
        if (key == 0ull) continue;  // empty slot (uniform per warp)
        const uint32_t pos = key_pos(key);
        const float4* rv = reinterpret_cast<const float4*>(rows + (size_t)pos * ld);
        float acc = 0.f;
        // all loads of a 256-float4 block are issued before the first FMA (a candidate is one random 2-4 KB row: the
        // kernel is bound by memory latency x loads in flight); the FMA order is the plain ascending one
        for (int t0 = 0; t0 < ld4; t0 += 256) {
            float4 a[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int t = t0 + lane + 32 * u;
                a[u] = t < ld4 ? __ldg(rv + t) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int t = t0 + lane + 32 * u;
                if (t < ld4) {
                    const float4 b = qv[t];
                    acc = fmaf(a[u].x, b.x, acc);
                    acc = fmaf(a[u].y, b.y, acc);
                    acc = fmaf(a[u].z, b.z, acc);
                    acc = fmaf(a[u].w, b.w, acc);
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) keys[i] = make_key(acc, pos);
    }
}

// rescore == 1: exact scores for the candidates of a one-term epoch.  grid = (nq, gy): the candidates of query
// blockIdx.x are spread over gy blocks of 8 warps, so the ~k*growth random 3 KB row reads per query run at HBM
// bandwidth instead of at one SM's latency (inside the compaction CTA they cost ~70 us per epoch).
__global__ void __launch_bounds__(256) rescore_candidates_kernel(const CompactParams p) {
    const int q = blockIdx.x;
    const int ncand = min(p.cnt[q], p.cap);
    rescore_keys(p.keys + (size_t)q * p.kstride + p.k, ncand, p.rows, p.q + (size_t)q * p.ld, p.ld,
                 blockIdx.y * 8 + (threadIdx.x >> 5), gridDim.y * 8);
}

__global__ void __launch_bounds__(kMergeThreads) compact_topk_kernel(const CompactParams p) {
    extern __shared__ __align__(16) unsigned char smem_merge[];
    uint64_t* buf = reinterpret_cast<uint64_t*>(smem_merge);
    __shared__ int cnt;
    const int tid = threadIdx.x;
    const int64_t q = blockIdx.x;
    const int k = p.k;
    uint64_t* base = p.keys + q * p.kstride;
    const int ncand = min(p.cnt[q], p.cap);
    const int64_t M = k + ncand;
    // (rescore == 1: rescore_candidates_kernel has already rewritten the candidates' scores, grid-wide)
    block_select_topk(buf, p.S, k, M, [&](int64_t c) { return base[c]; }, &cnt);
    if (p.rescore == 2) {
        const int P = max(2, pow2_ceil(k));  // <= S
        for (int j = k + tid; j < P; j += kMergeThreads) buf[j] = 0ull;
        rescore_keys(buf, k, p.rows, p.q + (size_t)q * p.ld, p.ld, tid >> 5, kMergeThreads / 32);
        __syncthreads();
        bitonic_sort_desc<kMergeThreads>(buf, P, 1, tid, -1);
        __syncthreads();
    }
    for (int j = tid; j < k; j += kMergeThreads) {
        const uint64_t key = buf[j];
        base[j] = key;
        if (p.D) {
            p.D[q * k + j] = key ? key_score(key) : -FLT_MAX;
            const uint32_t pos = key_pos(key);
            p.I[q * k + j] = key ? (p.ids ? p.ids[pos] : (int64_t)pos) : -1;
        }
    }
    if (tid == 0) {
        const uint64_t kth = buf[k - 1];
        p.thr[q] = kth ? key_score(kth) : -INFINITY;
        p.cnt[q] = 0;
    }
}

// ---- filter margins ----------------------------------------------------------------------------
// max over the rows of |x|^2, folded into *out with an integer atomicMax (non-negative floats order like their bits)
__global__ void row_norm2_max_kernel(const float* rows, int64_t n, int ld, float* out) {
    const int lane = threadIdx.x & 31;
    const int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float best = 0.f;
    for (int64_t r = w; r < n; r += nw) {
        const float4* rv = reinterpret_cast<const float4*>(rows + (size_t)r * ld);
        float acc = 0.f;
        for (int t = lane; t < (ld >> 2); t += 32) {
            const float4 a = rv[t];
            acc += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        best = fmaxf(best, acc);
    }
    if (lane == 0 && best > 0.f) atomicMax(reinterpret_cast<unsigned int*>(out), __float_as_uint(best));
}

// margin[i] = c * |q_i| * sqrt(max |x|^2) for real queries, 0 for padding (whose threshold is +inf anyway)
__global__ void query_margin_kernel(const float* q, int nq, int nq_pad, int ld, const float* norm2_max, float c,
                                    float* margin) {
    const int lane = threadIdx.x & 31;
    const int i = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
    if (i >= nq_pad) return;
    float acc = 0.f;
    if (i < nq)
        for (int t = lane; t < ld; t += 32) acc += q[(size_t)i * ld + t] * q[(size_t)i * ld + t];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    // 1.001: the two square roots and the sums above are themselves rounded
    if (lane == 0) margin[i] = i < nq ? c * 1.001f * sqrtf(acc) * sqrtf(*norm2_max) : 0.f;
}

// ---- small utility kernels ---------------------------------------------------------------
__global__ void iota_ids_kernel(int64_t* ids, int64_t start, int64_t n) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) ids[i] = start + i;
}

// *flag <- 1 when x[0..count) holds a NaN or an Inf (faiss Clustering::train refuses such a training set)
__global__ void nonfinite_flag_kernel(const float* x, int64_t count, int* flag) {
    bool bad = false;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
        bad |= !isfinite(x[i]);
    if (__syncthreads_or(bad) && threadIdx.x == 0) *flag = 1;
}

// Coarse labels of added rows -> list numbers.  A row whose scores are all NaN has no best centroid (label -1; faiss
// add_core leaves such a row out of every list but still counts it): it is filed under list 0, where no search can
// return it (a NaN score never passes a threshold compare), so the store stays searchable.
__global__ void labels_to_lists_kernel(const int64_t* labels, int32_t* out, int64_t n) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = labels[i] < 0 ? 0 : (int32_t)labels[i];
}

// (query, probe slot) pairs of a batch -> list numbers for the list-major scan.  A slot without a list (-1: a NaN
// query has no best centroids) gets an EMPTY result list - no work item will ever write it.
__global__ void probes_to_pairs_kernel(const int64_t* probes, int32_t* out, int64_t npairs, int k, uint64_t* parts) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= npairs) return;
    const int64_t l = probes[i];
    out[i] = (int32_t)l;
    if (l < 0)
        for (int j = 0; j < k; ++j) parts[(size_t)i * k + j] = 0ull;
}

// dst[n, ld] <- src[n, d], zero padding columns d..ld
__global__ void pad_rows_kernel(const float* src, float* dst, int64_t n, int d, int ld) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n * ld) {
        const int64_t r = i / ld;
        const int c = (int)(i - r * ld);
        dst[i] = c < d ? src[r * d + c] : 0.f;
    }
}

// dst[r, :] <- src[perm[r], :] in float4 units (IVF: group the row store by list)
__global__ void permute_rows_kernel(const float4* src, float4* dst, const uint32_t* perm, int64_t n, int ld4) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n * ld4) {
        const int64_t r = i / ld4;
        const int c = (int)(i - r * ld4);
        dst[i] = src[(size_t)perm[r] * ld4 + c];
    }
}
// out[i] = in[src[i]] (IVF re-grouping: list assignment of each moved row)
__global__ void permute_i32_kernel(const int32_t* in, int32_t* out, const uint32_t* src, int64_t n) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[src[i]];
}

// out[i] = ids[pos ? pos[start + i] : start + i]: external ids of the storage rows [start, start + n)
__global__ void gather_ids_kernel(const int64_t* ids, const uint32_t* pos, int64_t start, int64_t n, int64_t* out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = ids[pos ? (int64_t)pos[start + i] : start + i];
}

__global__ void iota_u32_kernel(uint32_t* v, int64_t start, int64_t n) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) v[start + i] = (uint32_t)(start + i);
}

// out[m, d] <- rows[pos[m], 0:d]
__global__ void gather_rows_kernel(const float* rows, int ld, int d, const int64_t* pos, int64_t m, float* out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < m * d) {
        const int64_t r = i / d;
        const int c = (int)(i - r * d);
        out[i] = rows[(size_t)pos[r] * ld + c];
    }
}

// pos[j] = lowest position whose value equals targets[j] (-1 if none); m <= 64 per launch
template <class T>
__global__ void find_ids_kernel(const T* ids, int64_t n, const int64_t* targets, int m, unsigned long long* pos) {
    __shared__ int64_t t[64];
    if (threadIdx.x < m) t[threadIdx.x] = targets[threadIdx.x];
    __syncthreads();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v = (int64_t)ids[i];
        for (int j = 0; j < m; ++j)
            if (v == t[j]) atomicMin(&pos[j], (unsigned long long)i);
    }
}

}  // namespace wb
