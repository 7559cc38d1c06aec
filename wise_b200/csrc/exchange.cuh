// exchange.cuh - multi-GPU top-k exchange + merge in ONE kernel over NVLink peer memory.
//
// The only exchange step of the sharded search: every rank holds the local top-k of the
// replicated queries; the global top-k is the merge of `world` lists.  Instead of an NCCL all-gather
// followed by a merge kernel, CTA q of every rank
//   1. stores its rank's k (score, id) pairs of query q straight into the mailbox of EVERY peer
//      (st.global on cudaIpc-mapped peer pointers -> NVLink / NVSwitch), fences, raises a per-query flag,
//   2. spins until the `world` flags of query q have arrived in its own mailbox,
//   3. merges the world*k candidates (same block_select_topk as K3; ties: score desc, rank asc,
//      local order asc == lowest global insertion position) and writes the final (D, I) row.
// A CTA depends only on the same-numbered CTA of the other ranks, never on another local CTA, so the
// kernel cannot deadlock on residency; spins are bounded and trap.  Mailboxes are double-buffered by
// the parity of the call sequence number.
#pragma once
#include "merge.cuh"

namespace wb {

constexpr int kExchMaxWorld = 16;

struct ExchParams {
    int rank, world;
    int64_t nq;
    int k;
    int S;                          // sort buffer entries
    uint32_t seq;                   // call sequence number (> 0)
    const float* D_local;           // [nq][k] this rank's results
    const int64_t* I_local;
    unsigned char* mailbox[kExchMaxWorld];  // mailbox base on every rank (peer-mapped); [rank] is local
    size_t region_bytes;            // one (buffer, sender) region
    size_t flags_bytes;             // bytes reserved for the per-query flags at the start of a region
    size_t cap_entries;             // D/I capacity of a region
    float* D;                       // [nq][k] merged
    int64_t* I;
};

__device__ __forceinline__ unsigned char* exch_region(const ExchParams& p, int owner, int buf, int sender) {
    return p.mailbox[owner] + ((size_t)buf * p.world + sender) * p.region_bytes;
}

__global__ void __launch_bounds__(kMergeThreads) exchange_merge_kernel(const ExchParams p) {
    extern __shared__ __align__(16) unsigned char smem_merge[];
    uint64_t* buf = reinterpret_cast<uint64_t*>(smem_merge);
    __shared__ int cnt;
    const int tid = threadIdx.x;
    const int64_t q = blockIdx.x;
    const int k = p.k;
    const int b = (int)(p.seq & 1u);
    // ---- 1. push this rank's row of query q to every rank (including itself) -------------------
    for (int r = 0; r < p.world; ++r) {
        unsigned char* reg = exch_region(p, r, b, p.rank);
        float* dD = reinterpret_cast<float*>(reg + p.flags_bytes);
        int64_t* dI = reinterpret_cast<int64_t*>(reg + p.flags_bytes + p.cap_entries * sizeof(float));
        for (int j = tid; j < k; j += kMergeThreads) {
            dD[q * k + j] = p.D_local[q * k + j];
            dI[q * k + j] = p.I_local[q * k + j];
        }
    }
    __threadfence_system();
    __syncthreads();
    if (tid < p.world) {
        volatile uint32_t* flag = reinterpret_cast<volatile uint32_t*>(exch_region(p, tid, b, p.rank)) + q;
        *flag = p.seq;
    }
    // ---- 2. wait for every rank's row of query q ---------------------------------------------------
    if (tid < p.world) {
        const volatile uint32_t* flag = reinterpret_cast<const volatile uint32_t*>(exch_region(p, p.rank, b, tid)) + q;
        uint32_t spins = 0;
        while (*flag != p.seq) {
            if (++spins > (1u << 28)) __trap();
        }
    }
    __threadfence_system();
    __syncthreads();
    // ---- 3. merge world * k candidates ---------------------------------------------------------------
    const int64_t M = (int64_t)p.world * k;
    auto load = [&](int64_t c) -> uint64_t {
        const int r = (int)(c / k);
        const int slot = (int)(c - (int64_t)r * k);
        const unsigned char* reg = exch_region(p, p.rank, b, r);
        const float* sD = reinterpret_cast<const float*>(reg + p.flags_bytes);
        const int64_t* sI = reinterpret_cast<const int64_t*>(reg + p.flags_bytes + p.cap_entries * sizeof(float));
        const volatile int64_t* vI = sI;
        const volatile float* vD = sD;
        return vI[q * k + slot] >= 0 ? make_key(vD[q * k + slot], (uint32_t)c) : 0ull;
    };
    if (!block_select_topk_lists(buf, p.S, k, p.world, [&](int l, int r) { return load((int64_t)l * k + r); }, &cnt)) {
        __syncthreads();
        block_select_topk(buf, p.S, k, M, load, &cnt);
    }
    const unsigned char* myreg0 = exch_region(p, p.rank, b, 0);
    for (int j = tid; j < k; j += kMergeThreads) {
        const uint64_t key = buf[j];
        float d = -FLT_MAX;
        int64_t id = -1;
        if (key) {
            d = key_score(key);
            const uint32_t c = key_pos(key);
            const int r = (int)(c / k);
            const int slot = (int)(c - (uint32_t)r * k);
            const unsigned char* reg = myreg0 + (size_t)r * p.region_bytes;
            const volatile int64_t* sI = reinterpret_cast<const volatile int64_t*>(reg + p.flags_bytes + p.cap_entries * sizeof(float));
            id = sI[q * k + slot];
        }
        p.D[q * k + j] = d;
        p.I[q * k + j] = id;
    }
}

}  // namespace wb
