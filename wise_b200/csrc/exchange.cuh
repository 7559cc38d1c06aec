// exchange.cuh - multi-GPU top-k exchange + merge over NVLink peer memory.
//
// The only exchange step of the sharded search: every rank holds the local top-k of the
// replicated queries; the global top-k is the merge of `world` lists.  Instead of an NCCL all-gather
// followed by a merge kernel, ONE CTA per query on every rank
//   1. stores its rank's k (score, id) pairs of query q straight into the mailbox of EVERY peer
//      (st.global on cudaIpc-mapped peer pointers -> NVLink / NVSwitch), fences, raises a per-query flag,
//   2. spins until the `world` flags of query q have arrived in its own mailbox,
//   3. merges the world*k candidates (same block_select_topk as K3; ties: score desc, rank asc,
//      local order asc == lowest global insertion position) and writes the final (D, I) row.
// exch_push_wait_merge is that sequence as a device function: exchange_merge_kernel runs it on a finished local
// (D, I), and the scan kernel's tail (scan.cuh, fused mode) runs it on the top-k it has just selected, so a sharded
// batch-1 search is a single launch per GPU.
// A CTA depends only on the CTA that handles the same query on the other ranks, never on another local CTA, and
// grids are capped at the resident CTA count, so the wait cannot deadlock on residency as long as every rank
// launches; the wait is bounded by a wall-clock timeout that raises *status instead of trapping (a trap would
// poison the CUDA context of every peer).  Mailboxes are double-buffered by the parity of the call sequence number.
#pragma once
#include "merge.cuh"

namespace wb {

constexpr int kExchMaxWorld = 16;
constexpr unsigned long long kExchTimeoutNs = 20ull * 1000 * 1000 * 1000;  // 20 s: a peer that never launched

struct ExchParams {
    int rank, world;
    int64_t nq;
    int k;
    int S;                          // sort buffer entries
    uint32_t seq;                   // call sequence number (> 0)
    const float* D_local;           // [nq][k] this rank's results (exchange_merge_kernel only)
    const int64_t* I_local;
    unsigned char* mailbox[kExchMaxWorld];  // mailbox base on every rank (peer-mapped); [rank] is local
    size_t region_bytes;            // one (buffer, sender) region
    size_t flags_bytes;             // bytes reserved for the per-query flags at the start of a region
    size_t cap_entries;             // D/I capacity of a region
    float* D;                       // [nq][k] merged
    int64_t* I;
    int* status;                    // pinned host word: set to 1 when a wait timed out (results are then invalid)
};

__device__ __forceinline__ unsigned char* exch_region(const ExchParams& p, int owner, int buf, int sender) {
    return p.mailbox[owner] + ((size_t)buf * p.world + sender) * p.region_bytes;
}

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Push this rank's k pairs of query q (local(j) -> score, id; id < 0 = empty slot) to every rank, wait for the
// other ranks' rows, merge, write p.D / p.I row q.  Runs on NT threads (tid in [0, NT)) that share barrier bar_id
// (< 0: the whole CTA); buf holds p.S keys, cnt is a shared counter.
template <int NT, class Local>
__device__ __forceinline__ void exch_push_wait_merge(const ExchParams& p, int64_t q, uint64_t* buf, int* cnt, int tid,
                                                     int bar_id, Local local) {
    const int k = p.k;
    const int b = (int)(p.seq & 1u);
    // ---- 1. push this rank's row of query q to every rank (including itself) -------------------
    for (int j = tid; j < k; j += NT) {
        float d;
        int64_t id;
        local(j, d, id);
        for (int r = 0; r < p.world; ++r) {
            unsigned char* reg = exch_region(p, r, b, p.rank);
            float* dD = reinterpret_cast<float*>(reg + p.flags_bytes);
            int64_t* dI = reinterpret_cast<int64_t*>(reg + p.flags_bytes + p.cap_entries * sizeof(float));
            dD[q * k + j] = d;
            dI[q * k + j] = id;
        }
    }
    __threadfence_system();
    sel_sync<NT>(bar_id);
    if (tid < p.world) {
        volatile uint32_t* flag = reinterpret_cast<volatile uint32_t*>(exch_region(p, tid, b, p.rank)) + q;
        *flag = p.seq;
    }
    // ---- 2. wait for every rank's row of query q ---------------------------------------------------
    if (tid < p.world) {
        const volatile uint32_t* flag = reinterpret_cast<const volatile uint32_t*>(exch_region(p, p.rank, b, tid)) + q;
        uint32_t spins = 0;
        unsigned long long t0 = 0;
        while (*flag != p.seq) {
            if ((++spins & 0xFFFu) == 0) {
                const unsigned long long now = global_ns();
                if (t0 == 0) t0 = now;
                else if (now - t0 > kExchTimeoutNs) {
                    if (p.status) *reinterpret_cast<volatile int*>(p.status) = 1;  // pinned host word
                    break;
                }
            }
        }
    }
    __threadfence_system();
    sel_sync<NT>(bar_id);
    // ---- 3. merge world * k candidates ---------------------------------------------------------------
    const int64_t M = (int64_t)p.world * k;
    auto load = [&](int64_t c) -> uint64_t {
        const int r = (int)(c / k);
        const int slot = (int)(c - (int64_t)r * k);
        const unsigned char* reg = exch_region(p, p.rank, b, r);
        const volatile float* vD = reinterpret_cast<const volatile float*>(reg + p.flags_bytes);
        const volatile int64_t* vI =
            reinterpret_cast<const volatile int64_t*>(reg + p.flags_bytes + p.cap_entries * sizeof(float));
        return vI[q * k + slot] >= 0 ? make_key(vD[q * k + slot], (uint32_t)c) : 0ull;
    };
    // (block_merge_heads was tried here too: with 2 ranks the sort-based merge already sorts all 2k keys in one
    // 256-key pass and the heads variant cost 4 us more per search - profiles/r02/NOTES.md)
    if (!block_select_topk_lists<NT>(buf, p.S, k, p.world, [&](int l, int r) { return load((int64_t)l * k + r); }, cnt,
                                     tid, bar_id)) {
        sel_sync<NT>(bar_id);
        block_select_topk<NT>(buf, p.S, k, M, load, cnt, tid, bar_id);
    }
    const unsigned char* myreg0 = exch_region(p, p.rank, b, 0);
    for (int j = tid; j < k; j += NT) {
        const uint64_t key = buf[j];
        float d = -FLT_MAX;
        int64_t id = -1;
        if (key) {
            d = key_score(key);
            const uint32_t c = key_pos(key);
            const int r = (int)(c / k);
            const int slot = (int)(c - (uint32_t)r * k);
            const unsigned char* reg = myreg0 + (size_t)r * p.region_bytes;
            const volatile int64_t* sI =
                reinterpret_cast<const volatile int64_t*>(reg + p.flags_bytes + p.cap_entries * sizeof(float));
            id = sI[q * k + slot];
        }
        p.D[q * k + j] = d;
        p.I[q * k + j] = id;
    }
    sel_sync<NT>(bar_id);  // buf is reused by the caller's next query
}

__global__ void __launch_bounds__(kMergeThreads) exchange_merge_kernel(const ExchParams p) {
    extern __shared__ __align__(16) unsigned char smem_merge[];
    uint64_t* buf = reinterpret_cast<uint64_t*>(smem_merge);
    __shared__ int cnt;
    const int k = p.k;
    // The grid is capped at the number of CTAs that are resident at once and every CTA walks its queries in the same
    // order on every rank: CTA c only ever waits for CTA c of the peers, which is resident and on the same query or
    // ahead - the wait never depends on the order in which the hardware dispatches CTAs.
    for (int64_t q = blockIdx.x; q < p.nq; q += gridDim.x) {
        exch_push_wait_merge<kMergeThreads>(p, q, buf, &cnt, threadIdx.x, -1, [&](int j, float& d, int64_t& id) {
            d = p.D_local[q * k + j];
            id = p.I_local[q * k + j];
        });
    }
}

}  // namespace wb
