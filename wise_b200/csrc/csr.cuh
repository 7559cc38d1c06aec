// csr.cuh - K8 ivf_add_scatter: group n items by their list on the device - histogram, exclusive scan, STABLE scatter.
//
// Replaces faiss IndexIVFFlat::add_core -> invlists->add_entry [faiss-upstream] (rows appended to their inverted list
// in insertion order), reached from index.add_with_ids, /root/reference/src/index/feature_search_index.py:81, and the
// per-iteration grouping of faiss Clustering::compute_centroids (feature_search_index.py:75).  Round 1 did this with a
// D2H copy of all assignments and a single-threaded counting sort on the host.
//
// The items are cut into B contiguous blocks.  cnt[b][l] = number of items of list l in block b (global atomics on a
// row only block b touches); a column-wise exclusive scan turns it into the first slot of (b, l); every block then
// places its items IN ORDER, tile by tile: inside a 256-item tile the rank of an item among the earlier items of the
// same list is counted with one pass over the tile's keys in shared memory, and the last item of each list in the tile
// advances the block's cursor.  No float math, no sorting network; output order = (list, original index), i.e. the
// insertion order inside every list is kept bit for bit.
// HBM-bound on paper (n * 12 bytes), in practice bound by the 256-compare rank loop: ~2 ms for 50M items.
#pragma once
#include "common.cuh"

namespace wb {

constexpr int kCsrTile = 256;

// cnt[b][l] += 1 for the items of block b; *bad is raised for an assignment outside [0, nlist)
__global__ void __launch_bounds__(kCsrTile) csr_hist_kernel(const int32_t* assign, int64_t n, int64_t nlist,
                                                            int64_t items_per_block, uint32_t* cnt, int* bad) {
    const int64_t b = blockIdx.x;
    const int64_t i0 = b * items_per_block, i1 = min(n, i0 + items_per_block);
    uint32_t* row = cnt + b * nlist;
    for (int64_t i = i0 + threadIdx.x; i < i1; i += kCsrTile) {
        const int32_t a = assign[i];
        if (a < 0 || a >= nlist) *bad = 1;
        else atomicAdd(&row[a], 1u);
    }
}

// column-wise exclusive scan over the blocks: cnt[b][l] <- sum of cnt[b'][l], b' < b;  total[l] = list size
__global__ void csr_colscan_kernel(uint32_t* cnt, int B, int64_t nlist, int64_t* total) {
    const int64_t l = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (l >= nlist) return;
    uint32_t run = 0;
    for (int b = 0; b < B; ++b) {
        const uint32_t t = cnt[(size_t)b * nlist + l];
        cnt[(size_t)b * nlist + l] = run;
        run += t;
    }
    total[l] = run;
}

// list_off[0..nlist] = exclusive scan of total[0..nlist) (one CTA; nlist is at most a few hundred thousand)
__global__ void __launch_bounds__(1024) csr_offsets_kernel(const int64_t* total, int64_t nlist, int64_t* list_off) {
    __shared__ int64_t warp_sum[32];
    __shared__ int64_t carry_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int64_t base = 0; base < nlist; base += 1024) {
        const int64_t l = base + tid;
        const int64_t v = l < nlist ? total[l] : 0;
        int64_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int64_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) warp_sum[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int64_t w = warp_sum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int64_t t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            warp_sum[lane] = w;  // inclusive over warps
        }
        __syncthreads();
        const int64_t before = carry_s + (warp ? warp_sum[warp - 1] : 0) + inc - v;
        if (l < nlist) list_off[l] = before;
        __syncthreads();
        if (tid == 1023) carry_s = before + v;
        __syncthreads();
    }
    if (tid == 0) list_off[nlist] = carry_s;
}

// Stable scatter.  src[slot] = index of the item that lands in slot; pos_out[slot] = pos_in[item] (or the item index
// when pos_in is null) - the IVF store uses it to carry the ORIGINAL insertion position through a re-grouping.
__global__ void __launch_bounds__(kCsrTile) csr_scatter_kernel(const int32_t* assign, int64_t n, int64_t nlist,
                                                               int64_t items_per_block, uint32_t* cnt,
                                                               const int64_t* list_off, const uint32_t* pos_in,
                                                               uint32_t* src, uint32_t* pos_out) {
    __shared__ int32_t keys[kCsrTile];
    const int64_t b = blockIdx.x;
    const int64_t i0 = b * items_per_block, i1 = min(n, i0 + items_per_block);
    uint32_t* cur = cnt + b * nlist;  // first free slot (relative to list_off) of every list for this block
    const int tid = threadIdx.x;
    for (int64_t t0 = i0; t0 < i1; t0 += kCsrTile) {
        const int64_t i = t0 + tid;
        const int nv = (int)min((int64_t)kCsrTile, i1 - t0);
        const int32_t key = i < i1 ? assign[i] : -1;
        keys[tid] = key;
        __syncthreads();
        const bool ok = i < i1 && key >= 0 && key < nlist;
        int rank = 0, later = 0;
        uint32_t base = 0;
        if (ok) {
            for (int j = 0; j < nv; ++j) {  // broadcast reads: one wavefront per j
                const bool same = keys[j] == key;
                rank += (same && j < tid);
                later += (same && j > tid);
            }
            base = __ldcg(&cur[key]);  // written by an earlier tile of THIS block, through L2
            const int64_t slot = list_off[key] + base + rank;
            src[slot] = (uint32_t)i;
            if (pos_out) pos_out[slot] = pos_in ? pos_in[i] : (uint32_t)i;
        }
        __syncthreads();  // every item of the tile has read its list's cursor
        if (ok && later == 0) __stcg(&cur[key], base + (uint32_t)rank + 1u);
        __syncthreads();  // the cursors are in L2 before the next tile reads them; keys[] is reused
    }
}

}  // namespace wb
