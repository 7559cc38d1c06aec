// coarse.cuh - K4 inside K5: the IVF coarse quantizer as the PROLOGUE of the list-scan kernel.
//
// Replaces the quantizer->search(n, x, nprobe, ...) call at the top of faiss IndexIVF::search
// [faiss-upstream], reached from /root/reference/src/index/feature_search_index.py:113 with nprobe set at :106.
//
// A one-query IVF search used to be two launches: a scan of the centroid table with a fused 148-list merge (top-nprobe),
// then the scan of the probed lists with a second fused merge.  The first launch moves a few MB and costs 12-20 us of
// fixed work (prologue, per-CTA sort, cross-CTA merge, launch gap) - as much as the list scan itself at nprobe 8.
// Here the CTAs of a query group
//   1. score a slice of the centroid table each (warp per 4 centroids, coalesced 128-bit loads, no staging - the table
//      is L2 resident) and publish the ORDERED scores (order_f32) to a global array,
//   2. meet at a group barrier (arrival counter; the launch is cooperative, so every CTA is resident),
//   3. each select the SAME top-nprobe of the nlist scores - exact radix selection on (score, ~list id) keys
//      (radix_select.cuh: range-adaptive 1024-bin histograms, 2-3 passes over keys held in shared memory), then a
//      bitonic sort of the nprobe winners - and go on to scan the lists.
// The selection is redundant per CTA (148 x nlist keys from L2) but needs no second barrier and no broadcast.
// Order of the winners: score descending, ties by LOWER list id - the order merge_topk_kernel gave the probe table.
// A NaN score gets key 0 (never wins against a real score; a winner with key 0 becomes probe -1 = no list).
#pragma once
#include "radix_select.cuh"

namespace wb {

constexpr int kCoarseRowsPerWarp = 4;   // centroids scored together by one warp (independent loads in flight)

struct CoarseSmem {
    size_t keys, scratch, sel, total;
};

// Scratch of the coarse phase inside the (still idle) ring of the scan kernel.
__host__ __device__ inline CoarseSmem coarse_smem_layout(int64_t nlist, int np) {
    CoarseSmem L;
    size_t o = 0;
    L.keys = o;    o += ((size_t)nlist * 4 + 15) & ~(size_t)15;   // ordered scores of every centroid
    L.scratch = o; o += kRadixScratchBytes;                       // block_radix_select
    L.sel = o;     o += (size_t)pow2_ceil(np) * 8;
    L.total = o;
    return L;
}

__device__ __forceinline__ uint64_t coarse_key64(uint32_t okey, uint32_t idx) {
    return ((uint64_t)okey << 32) | (uint64_t)(0xFFFFFFFFu - idx);
}

// Scores of the centroids [c0, c1) against the query in shared memory `q4` (d4 float4s), ordered keys -> out[c].
// Called by all NT threads of the CTA (whole warps).
template <int NT>
__device__ __forceinline__ void coarse_score_slice(const float* centroids, int ld, const float4* q4, int64_t c0, int64_t c1,
                                                   uint32_t* out, int tid) {
    const int warp = tid >> 5, lane = tid & 31;
    const int d4 = ld >> 2;
    constexpr int R = kCoarseRowsPerWarp;
    for (int64_t cb = c0 + (int64_t)warp * R; cb < c1; cb += (int64_t)(NT / 32) * R) {
        const int nv = (int)min((int64_t)R, c1 - cb);
        float acc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = 0.f;
        const float4* base = reinterpret_cast<const float4*>(centroids + (size_t)cb * ld);
#pragma unroll 2
        for (int i = lane; i < d4; i += 32) {
            float4 x[R];
#pragma unroll
            for (int r = 0; r < R; ++r) x[r] = r < nv ? __ldg(base + (size_t)r * d4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 q = q4[i];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float a = acc[r];
                a = fmaf(x[r].x, q.x, a);
                a = fmaf(x[r].y, q.y, a);
                a = fmaf(x[r].z, q.z, a);
                a = fmaf(x[r].w, q.w, a);
                acc[r] = a;
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (lane == r && r < nv) {
                const float s = acc[r];
                out[cb + r] = (s == s) ? order_f32(s) : 0u;
            }
        }
    }
}

// Arrival barrier of the `expected` CTAs that share `counter` (zero before the first arrival; the caller resets it once
// every CTA is known to have left - the scan kernel's tail does).  All threads of the CTA call it.  The launch is
// cooperative, so a missing CTA is a bug, not a scheduling accident: the wait traps after ~4 s instead of hanging.
__device__ __forceinline__ void group_arrive_wait(unsigned int* counter, unsigned int expected) {
    __threadfence();  // this thread's global stores are visible device-wide before the CTA's arrival is
    __syncthreads();
    if (threadIdx.x == 0) {
        atomicAdd(counter, 1u);
        unsigned int seen = 0, spins = 0;
        unsigned long long t0 = 0;
        while (true) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
            if (seen >= expected) break;
            if ((++spins & 0x3FFu) == 0) {
                unsigned long long now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                if (t0 == 0) t0 = now;
                else if (now - t0 > 4000000000ull) __trap();
            }
        }
        __threadfence();
    }
    __syncthreads();
}

}  // namespace wb
