// common.cuh - PTX helpers (mbarrier, bulk async copy, named barriers) and the sortable
// (score, position) key shared by every selection kernel.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace wb {

constexpr int kWarp = 32;

// ---- sortable 64-bit key ---------------------------------------------------------------
// key = orderable(score) << 32 | ~position.  Larger key == better candidate: higher score
// first, then LOWER position (the tie rule).  key 0 is "empty".
__device__ __forceinline__ uint32_t order_f32(float f) {
    f += 0.0f;  // -0.0 -> +0.0
    uint32_t u = __float_as_uint(f);
    return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float unorder_f32(uint32_t o) {
    uint32_t u = (o & 0x80000000u) ? (o ^ 0x80000000u) : ~o;
    return __uint_as_float(u);
}
__device__ __forceinline__ uint64_t make_key(float score, uint32_t pos) {
    return ((uint64_t)order_f32(score) << 32) | (uint64_t)(0xFFFFFFFFu - pos);
}
__device__ __forceinline__ float key_score(uint64_t k) { return unorder_f32((uint32_t)(k >> 32)); }
__device__ __forceinline__ uint32_t key_pos(uint64_t k) { return 0xFFFFFFFFu - (uint32_t)k; }

// ---- shared-memory addresses -------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// ---- mbarrier ----------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
// Spin on a phase parity.  A bounded spin turns a protocol bug into a trap (launch failure)
// instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    uint32_t spins = 0;
    const uint32_t a = smem_u32(bar);
    while (true) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(a), "r"(parity)
            : "memory");
        if (done) break;
        if (++spins > (1u << 26)) __trap();
    }
}

// ---- 1-D bulk async copy global -> shared (TMA engine, SASS UBLKCP) ------------------------
// dst/src 16-byte aligned, bytes a multiple of 16; completion is signalled on `bar` as
// complete_tx(bytes).
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// One lane of a fully converged warp.  Code guarded by this predicate is known to the compiler to run
// in exactly one thread, so warp-uniform instructions (UTCHMMA, UTMALDG, UTCBAR) get their operands in
// uniform registers directly; a plain `if (lane == 0)` makes nvcc wrap every such instruction in an
// ELECT / BRA.U.ANY serialisation loop (~100 cycles per tcgen05.mma - measured, profiles/r01).
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred = 0;
    asm volatile(
        "{\n"
        ".reg .b32 rx;\n"
        ".reg .pred px;\n"
        "elect.sync rx|px, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, px;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// ---- named barriers (sub-CTA sync among the consumer warps) -----------------------------------
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// barrier + OR-reduction of a per-thread predicate; every participant gets the result.
__device__ __forceinline__ bool named_bar_or(int id, int nthreads, bool pred) {
    uint32_t out;
    asm volatile(
        "{\n"
        ".reg .pred p, q;\n"
        "setp.ne.u32 q, %1, 0;\n"
        "bar.red.or.pred p, %2, %3, q;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(out)
        : "r"((uint32_t)pred), "r"(id), "r"(nthreads)
        : "memory");
    return out != 0;
}

__host__ __device__ __forceinline__ int pow2_ceil(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

// ---- bitonic sort of `nlists` arrays of P (power of two) keys, descending -------------------
// Executed by NT threads that share barrier `bar_id` (bar_id < 0 => __syncthreads()).
// Thread i owns the compare-exchange j = i mod (P/2) of list i / (P/2).  For strides <= 32 the 32 consecutive j of a
// warp touch only the warp's own 64 keys, so those steps need a __syncwarp, not a block barrier: a 1024-key sort takes
// 15 block barriers instead of 55, a 256-key flush 6 instead of 36 (each barrier costs a 256-thread group ~100 cycles;
// the sorts sit on the tail of every scan and on every threshold flush).
template <int NT>
__device__ __forceinline__ void bitonic_block_sync(int bar_id) {
    if (bar_id < 0) __syncthreads();
    else named_bar_sync(bar_id, NT);
}

template <int NT>
__device__ __forceinline__ void bitonic_sort_desc(uint64_t* keys, int P, int nlists, int tid, int bar_id) {
    static_assert(NT % 32 == 0, "whole warps");
    const int half = P >> 1;
    const int lg_half = 31 - __clz(half);  // P is a power of two: no integer division in the loops
    const int total = nlists << lg_half;
    auto step = [&](int size, int stride) {
        for (int i = tid; i < total; i += NT) {
            const int l = i >> lg_half;
            const int j = i & (half - 1);
            const int pos = 2 * j - (j & (stride - 1));
            uint64_t* base = keys + ((size_t)l << (lg_half + 1));
            const uint64_t a = base[pos], b = base[pos + stride];
            const bool desc = ((pos & size) == 0);
            if ((a < b) == desc) {
                base[pos] = b;
                base[pos + stride] = a;
            }
        }
    };
    if (half < 32) {  // tiny lists: a warp's pairs are not confined to 64 keys of one list - plain scheme
        for (int size = 2; size <= P; size <<= 1)
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                step(size, stride);
                bitonic_block_sync<NT>(bar_id);
            }
        return;
    }
    // (the loop trip count is warp-uniform: total is a multiple of 32 and NT a multiple of 32)
    for (int size = 2; size <= P; size <<= 1) {
        int stride = size >> 1;
        for (; stride > 32; stride >>= 1) {  // pairs cross warps
            step(size, stride);
            bitonic_block_sync<NT>(bar_id);
        }
        for (; stride > 0; stride >>= 1) {   // pairs stay inside the warp's 64 keys
            step(size, stride);
            __syncwarp();
        }
        if (size >= 64) bitonic_block_sync<NT>(bar_id);  // the next size starts with a stride >= 64 (or the sort is done)
    }
    if (P < 64) bitonic_block_sync<NT>(bar_id);
}

// ---- rank counting ---------------------------------------------------------------------------------------------
// Non-empty keys (0 = empty) are distinct, so "number of larger keys" IS the position of a key in descending order.
// The tail of a scan runs on the 8 warps of ONE SM - two warps per scheduler, every dependent instruction exposed - so
// the loops below rank 4 keys per pass over the array with 4 independent loads in flight (a warp that ranked one key
// at a time with a rolled loop spent ~270 ns per key; an out-of-line copy of this routine was slower than the inlined
// one: per-phase %globaltimer stamps in profiles/r02/NOTES.md).
struct Rank4 {
    int r[4];
};
// Ranks of k0..k3 among the n keys of the shared-memory array `arr`; the lanes of the calling warp split the array,
// every lane gets the ranks.
__device__ __forceinline__ Rank4 warp_rank4_desc(const uint64_t* arr, int n, uint64_t k0, uint64_t k1, uint64_t k2,
                                                 uint64_t k3) {
    const int lane = threadIdx.x & 31;
    int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    for (int u0 = 0; u0 < n; u0 += 128) {
        uint64_t a[4];
#pragma unroll
        for (int x = 0; x < 4; ++x) {
            const int u = u0 + x * 32 + lane;
            a[x] = u < n ? arr[u] : 0ull;
        }
#pragma unroll
        for (int x = 0; x < 4; ++x) {
            c0 += a[x] > k0;
            c1 += a[x] > k1;
            c2 += a[x] > k2;
            c3 += a[x] > k3;
        }
    }
    Rank4 out;
    out.r[0] = (int)__reduce_add_sync(0xffffffffu, (unsigned)c0);
    out.r[1] = (int)__reduce_add_sync(0xffffffffu, (unsigned)c1);
    out.r[2] = (int)__reduce_add_sync(0xffffffffu, (unsigned)c2);
    out.r[3] = (int)__reduce_add_sync(0xffffffffu, (unsigned)c3);
    return out;
}

// dst[rank] = key for every non-empty key of src[0, n) whose rank is below `limit`; `nwarps` whole warps (warp index
// `warp`).  src / dst are shared memory and must not overlap; the caller zero-fills dst and synchronises around the call.
__device__ __forceinline__ void block_rank_scatter(const uint64_t* src, int n, uint64_t* dst, int limit, int warp,
                                                   int nwarps) {
    const int lane = threadIdx.x & 31;
    const uint64_t* a = src;
    for (int t0 = warp * 4; t0 < n; t0 += nwarps * 4) {
        uint64_t key = 0ull;  // lane b < 4 holds key t0 + b
        if (lane < 4 && t0 + lane < n) key = src[t0 + lane];
        const uint64_t k0 = __shfl_sync(0xffffffffu, key, 0), k1 = __shfl_sync(0xffffffffu, key, 1),
                       k2 = __shfl_sync(0xffffffffu, key, 2), k3 = __shfl_sync(0xffffffffu, key, 3);
        const Rank4 rk = warp_rank4_desc(a, n, k0, k1, k2, k3);
        const int myr = lane == 0 ? rk.r[0] : lane == 1 ? rk.r[1] : lane == 2 ? rk.r[2] : rk.r[3];
        if (lane < 4 && key && myr < limit) dst[myr] = key;
    }
}

// In-place sort of ONE array of P <= NT keys, descending (empty keys last): warp w ranks the keys [32 w, 32 w + 32),
// lane i keeps (key, rank) of key 32 w + i in registers across the barrier that separates the reads from the scatter.
// Three barriers instead of the 36 compare-exchange steps of a 256-key bitonic sort.
template <int NT>
__device__ __forceinline__ void rank_sort_desc(uint64_t* keys, int P, int tid, int bar_id) {
    const int warp = tid >> 5, lane = tid & 31;
    const uint64_t* a = keys;
    uint64_t my_key = 0ull;
    int my_rank = 0;
    if (warp * 32 < P) {  // (warp-uniform)
        const uint64_t own = warp * 32 + lane < P ? keys[warp * 32 + lane] : 0ull;
#pragma unroll 1
        for (int b0 = 0; b0 < 32; b0 += 4) {
            const uint64_t k0 = __shfl_sync(0xffffffffu, own, b0), k1 = __shfl_sync(0xffffffffu, own, b0 + 1),
                           k2 = __shfl_sync(0xffffffffu, own, b0 + 2), k3 = __shfl_sync(0xffffffffu, own, b0 + 3);
            const Rank4 rk = warp_rank4_desc(a, P, k0, k1, k2, k3);
            const int x = lane - b0;
            if (x >= 0 && x < 4) my_rank = x == 0 ? rk.r[0] : x == 1 ? rk.r[1] : x == 2 ? rk.r[2] : rk.r[3];
        }
        my_key = own;
    }
    bitonic_block_sync<NT>(bar_id);
    if (tid < P) keys[tid] = 0ull;
    bitonic_block_sync<NT>(bar_id);
    if (my_key) keys[my_rank] = my_key;
    bitonic_block_sync<NT>(bar_id);
}

}  // namespace wb
