// radix_select.cuh - exact block-level top-k of n DISTINCT 64-bit keys (0 = empty slot) without sorting them.
//
// Used where a thread block must pick k winners out of many more UNORDERED candidates and only the winners need an
// order: the top-nprobe of the centroid scores in the fused coarse quantizer (coarse.cuh) - the selection inside faiss
// quantizer->search [faiss-upstream], reached from /root/reference/src/index/feature_search_index.py:113.
// (Tried for the merge of the per-CTA top-k lists at the tail of a scan as well: 148 x 100 staged keys took 16.6 us
// against 12.4 us for the sort-based merge, which touches only the heads of the SORTED lists - profiles/r02/NOTES.md;
// that merge is block_merge_heads in merge.cuh.)
// Every pass histograms the keys of the current range [lo, hi] into <= 1024 equal-width bins, takes every bin above the
// one that holds the need-th largest key, and either finishes (that bin is taken whole) or recurses into it.  The first
// range is [min key, max key], so the bins resolve the score distribution at once; a range shrinks by >= 2^10 per pass
// and a width-1 bin holds one key, so the loop ends after at most 7 passes; in practice after ONE, because a boundary bin
// of up to 256 keys is resolved on the spot by rank counting.  The winners are then ordered by the caller.
// tests/test_coarse_model.py checks the arithmetic on the CPU.
#pragma once
#include "common.cuh"

namespace wb {

constexpr int kSelBins = 1024;  // histogram bins of one selection pass (32 chunks of 32)
constexpr int kRadixResolveMax = 256;  // a boundary bin of up to this many keys is resolved by rank counting (fits the histogram's memory)

// Shared-memory scratch of one selection (the caller places it; 16-byte aligned).
constexpr size_t kRadixScratchBytes = (size_t)kSelBins * 4 + 32 * 4 + 1024;
struct RadixScratch {
    uint32_t* hist;      // [kSelBins]
    uint32_t* chunk;     // [32] per-chunk totals
    unsigned char* ctl;  // [1024]: lo, hi (u64) | sel_n, need, bstar, above, cnt (int) | per-warp min / max / count
    __device__ __forceinline__ explicit RadixScratch(unsigned char* base)
        : hist(reinterpret_cast<uint32_t*>(base)), chunk(reinterpret_cast<uint32_t*>(base) + kSelBins),
          ctl(base + (size_t)kSelBins * 4 + 32 * 4) {}
};

template <int NT>
__device__ __forceinline__ void rs_sync(int bar_id) {
    if (bar_id < 0) __syncthreads();
    else named_bar_sync(bar_id, NT);
}

// Top-min(need0, #non-empty) of the keys key_at(i), i in [0, n), written to sel[] in arbitrary order; returns how many.
// Non-empty keys must be distinct.  Runs on NT threads (whole warps, tid in [0, NT)) that share barrier bar_id
// (< 0: the whole CTA); key_at must be cheap and side-effect free (it is evaluated in every pass).
template <int NT, class KeyAt>
__device__ __forceinline__ int block_radix_select(int n, int need0, KeyAt key_at, RadixScratch sc, uint64_t* sel, int tid,
                                                  int bar_id) {
    static_assert(NT % 32 == 0 && NT <= 1024, "whole warps");
    const int warp = tid >> 5, lane = tid & 31;
    uint32_t* hist = sc.hist;
    uint32_t* chunk = sc.chunk;
    uint64_t* ctl64 = reinterpret_cast<uint64_t*>(sc.ctl);            // [0] lo, [1] hi
    int* ctl = reinterpret_cast<int*>(sc.ctl + 16);                   // [0] sel_n [1] need [2] bstar [3] above [4] cnt
    uint64_t* wmin = reinterpret_cast<uint64_t*>(sc.ctl + 64);        // [32]
    uint64_t* wmax = wmin + 32;                                       // [32]
    uint32_t* wcnt = reinterpret_cast<uint32_t*>(sc.ctl + 64 + 512);  // [32]
    // ---- range and number of the non-empty keys -------------------------------------------------------------
    uint64_t mn = ~0ull, mx = 0ull;
    uint32_t nz = 0;
    for (int i = tid; i < n; i += NT) {
        const uint64_t v = key_at(i);
        if (v) {
            mn = v < mn ? v : mn;
            mx = v > mx ? v : mx;
            ++nz;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const uint64_t a = __shfl_xor_sync(0xffffffffu, mn, o), b = __shfl_xor_sync(0xffffffffu, mx, o);
        mn = a < mn ? a : mn;
        mx = b > mx ? b : mx;
    }
    nz = __reduce_add_sync(0xffffffffu, nz);
    if (lane == 0) {
        wmin[warp] = mn;
        wmax[warp] = mx;
        wcnt[warp] = nz;
    }
    rs_sync<NT>(bar_id);
    if (tid == 0) {
        for (int w = 1; w < NT / 32; ++w) {
            mn = wmin[w] < mn ? wmin[w] : mn;
            mx = wmax[w] > mx ? wmax[w] : mx;
            nz += wcnt[w];
        }
        ctl64[0] = mn;
        ctl64[1] = mx;
        ctl[0] = 0;
        ctl[1] = min(need0, (int)nz);
    }
    rs_sync<NT>(bar_id);
    const int need_total = ctl[1];
    if (need_total <= 0) return 0;  // (uniform)
    for (int pass = 0; pass < 8; ++pass) {
        const uint64_t lo = ctl64[0], hi = ctl64[1];
        const int need = ctl[1];
        const uint64_t range = hi - lo;
        const int sh = range < (uint64_t)kSelBins ? 0 : (64 - __clzll((long long)range)) - 10;  // (range >> sh) < 1024
        for (int i = tid; i < kSelBins; i += NT) hist[i] = 0u;
        rs_sync<NT>(bar_id);
        for (int i = tid; i < n; i += NT) {
            const uint64_t kk = key_at(i);
            if (kk >= lo && kk <= hi) atomicAdd(&hist[(uint32_t)((kk - lo) >> sh)], 1u);  // lo > 0: empty keys never count
        }
        rs_sync<NT>(bar_id);
        for (int c = warp; c < 32; c += NT / 32) {
            const uint32_t s = __reduce_add_sync(0xffffffffu, hist[c * 32 + lane]);
            if (lane == 0) chunk[c] = s;
        }
        rs_sync<NT>(bar_id);
        if (warp == 0) {
            // suffix sums (inclusive) over the 32 chunk totals, then over the 32 bins of the boundary chunk
            const uint32_t cv = chunk[lane];
            uint32_t cs = cv;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_down_sync(0xffffffffu, cs, o);
                if (lane + o < 32) cs += t;
            }
            const unsigned mc = __ballot_sync(0xffffffffu, cs >= (uint32_t)need);  // lane 0 always: the range holds >= need keys
            const int cstar = 31 - __clz(mc);
            const uint32_t above_chunks = __shfl_sync(0xffffffffu, cs, cstar) - __shfl_sync(0xffffffffu, cv, cstar);
            const uint32_t hv = hist[cstar * 32 + lane];
            uint32_t hs = hv;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_down_sync(0xffffffffu, hs, o);
                if (lane + o < 32) hs += t;
            }
            hs += above_chunks;  // keys in bins >= cstar * 32 + lane
            const unsigned mb = __ballot_sync(0xffffffffu, hs >= (uint32_t)need);
            const int bl = 31 - __clz(mb);
            const uint32_t ge = __shfl_sync(0xffffffffu, hs, bl), cnt = __shfl_sync(0xffffffffu, hv, bl);
            if (lane == 0) {
                ctl[2] = cstar * 32 + bl;
                ctl[3] = (int)(ge - cnt);
                ctl[4] = (int)cnt;
                ctl[5] = 0;  // keys collected from the boundary bin
            }
        }
        rs_sync<NT>(bar_id);
        const uint32_t bstar = (uint32_t)ctl[2];
        const int above = ctl[3], cnt = ctl[4];
        const bool last = cnt == need - above;  // the boundary bin is taken whole: done
        // A small boundary bin is resolved on the spot instead of by another histogram pass (a pass costs ~2 us of
        // barriers whatever it counts): its keys are collected - the histogram is dead by now, its memory holds them -
        // and the best `need - above` of them are found by rank counting.
        const bool resolve = !last && cnt <= kRadixResolveMax;
        uint64_t* binkeys = reinterpret_cast<uint64_t*>(hist);
        for (int i = tid; i < n; i += NT) {
            const uint64_t kk = key_at(i);
            if (kk >= lo && kk <= hi) {
                const uint32_t b = (uint32_t)((kk - lo) >> sh);
                if (b > bstar || (last && b == bstar)) sel[atomicAdd(&ctl[0], 1)] = kk;
                else if (resolve && b == bstar) binkeys[atomicAdd(&ctl[5], 1)] = kk;
            }
        }
        rs_sync<NT>(bar_id);
        if (last) break;
        if (resolve) {
            const int want = need - above;
            for (int t = tid; t < cnt; t += NT) {
                const uint64_t kk = binkeys[t];
                int rank = 0;
                for (int u = 0; u < cnt; ++u) rank += binkeys[u] > kk;
                if (rank < want) sel[atomicAdd(&ctl[0], 1)] = kk;
            }
            rs_sync<NT>(bar_id);
            break;
        }
        if (tid == 0) {
            const uint64_t nlo = lo + ((uint64_t)bstar << sh);
            uint64_t nhi = nlo + (((uint64_t)1 << sh) - 1);
            if (nhi > hi) nhi = hi;
            ctl64[0] = nlo;
            ctl64[1] = nhi;
            ctl[1] = need - above;
        }
        rs_sync<NT>(bar_id);
    }
    return need_total;
}

}  // namespace wb
