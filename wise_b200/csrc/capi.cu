// capi.cu - the C-ABI of libwiseb200.so (include/wise_b200.h): index handles, HBM row store,
// kernel launch configuration.  Host-side plumbing only; every flop happens in scan.cuh /
// merge.cuh / kmeans.cuh.  No CPU fallback: without a CUDA device every compute call fails.
#include <cuda_runtime.h>
#include <float.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <random>
#include <string>
#include <vector>

#include "../../include/wise_b200.h"
#include "csr.cuh"
#include "exchange.cuh"
#include "gemm.cuh"
#include "gemm_ss.cuh"
#include "ivf_lm.cuh"
#include "kmeans.cuh"
#include "merge.cuh"
#include "scan.cuh"
#include "tarstore.h"

using namespace wb;

// ---- errors ------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return 1;
}
#define CK(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            return fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)
#define TRY(call)              \
    do {                       \
        int r_ = (call);       \
        if (r_ != 0) return r_; \
    } while (0)

extern "C" const char* wb_last_error(void) { return g_err.c_str(); }
extern "C" const char* wb_version(void) { return "wise_b200 0.1 (sm_100a)"; }
extern "C" int wb_device_count(int* count) {
    CK(cudaGetDeviceCount(count));
    return 0;
}

// ---- grow-only device scratch --------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) CK(cudaFree(p));
        p = nullptr;
        cap = 0;
        size_t want = std::max(bytes, (size_t)1 << 16);
        CK(cudaMalloc(&p, want));
        cap = want;
        return 0;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T>
    T* as() { return reinterpret_cast<T*>(p); }
};

// Grow-only pinned host buffer (cudaHostAlloc): a copy between pageable memory and the device is staged by the
// driver and blocks the call; small transfers go through this buffer instead and stay asynchronous.
struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) CK(cudaFreeHost(p));
        p = nullptr;
        cap = 0;
        const size_t want = std::max(bytes, (size_t)1 << 16);
        CK(cudaHostAlloc(&p, want, cudaHostAllocMapped | cudaHostAllocPortable));  // kernels may write into it (UVA)
        cap = want;
        return 0;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};
struct wb_index {
    int device = 0;
    int d = 0, ld = 0;
    bool ivf = false;
    int64_t nlist = 0;
    bool trained = true;
    bool spherical = false;  // k-means: L2-renormalise the centroids after every update (faiss cp.spherical)
    cudaStream_t stream = nullptr;
    // row store (insertion order)
    float* rows = nullptr;
    int64_t* ids = nullptr;
    int32_t* assign = nullptr;  // IVF: list of each row
    int64_t n = 0, cap = 0;
    float* norm2_max = nullptr;  // device scalar: max |row|^2 over everything ever added (K2 filter margin)
    // IVF quantizer + inverted lists.  `ids` always stay in INSERTION order.  After a (re)grouping the rows are
    // stored list by list (insertion order kept inside a list): row_pos[row] is then the insertion position of a
    // storage row - the tie rule and the id lookup - and slot_row is null.  When the transient second copy of the rows
    // does not fit, the rows stay where they are and slot_row[CSR slot] is the storage row of each list slot.
    float* centroids = nullptr;   // [nlist, ld]
    uint32_t* slot_row = nullptr; // [n] or null (slot == storage row)
    uint32_t* row_pos = nullptr;  // [map_cap] or null (storage row == insertion position)
    int64_t map_cap = 0;
    int64_t* list_off = nullptr;  // [nlist + 1]
    bool csr_dirty = true;
    // scratch
    DevBuf parts, qbuf, dbuf, ibuf, pD, pI, xbuf, idbuf, misc, kperm, koff, gimg, gimg2, gkeys, gstate, gmargin, eD, eI, tailcnt, ccnt, ctot, lmA, lmB, lmC, lmD, lmE, lmF, lmG;
    PinBuf pin_q, pin_o;  // pinned staging of small query / result transfers of the host API
    static constexpr int kAddSlots = 8;   // wb_add_with_ids_pinned: one event per in-flight pinned buffer
    cudaEvent_t add_ev[kAddSlots] = {};
    bool add_pending[kAddSlots] = {};
    int64_t gemm_launches = 0, gemm_fallbacks = 0;
    DevBuf phase;               // WB_PHASE_TS=1: per-CTA phase stamps of the last fused-tail scan launch (diagnostics)
    DevBuf ckeys;               // fused coarse quantizer: ordered centroid scores of the queries of one launch
    bool coop_ok = false;       // the device takes cooperative launches (fused coarse quantizer, coarse.cuh)
    int64_t coarse_fused = 0;   // IVF searches served by ONE launch (coarse quantizer inside the list scan)
    // device properties
    int sm_count = 148;
    int smem_max = 0;
    // accounting
    int64_t launches = 0;
    bool timing = false;
    static constexpr int kEvRing = 128;  // event pairs of the most recent timed scan launches
    cudaEvent_t ev0[kEvRing] = {}, ev1[kEvRing] = {};
    int64_t ev_count = 0;
};

static int set_dev(const wb_index* h) {
    CK(cudaSetDevice(h->device));
    return 0;
}

// ---- construction --------------------------------------------------------------------------
static int create_common(int d, int device, bool ivf, int64_t nlist, wb_index** out) {
    if (!out) return fail("out is NULL");
    *out = nullptr;
    if (d <= 0 || d > 65536) return fail("invalid dimension %d", d);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail("no CUDA device available (%s): wise_b200 has no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail("device %d out of range (%d devices)", device, ndev);
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    wb_index* h = new wb_index();
    h->device = device;
    h->d = d;
    h->ld = (d + 3) & ~3;
    h->ivf = ivf;
    h->nlist = nlist;
    h->trained = !ivf;
    h->sm_count = prop.multiProcessorCount;
    h->smem_max = (int)prop.sharedMemPerBlockOptin;
    h->coop_ok = prop.cooperativeLaunch != 0;
    CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    for (int i = 0; i < wb_index::kEvRing; ++i) {
        CK(cudaEventCreate(&h->ev0[i]));
        CK(cudaEventCreate(&h->ev1[i]));
    }
    CK(cudaMalloc(&h->norm2_max, sizeof(float)));
    CK(cudaMemsetAsync(h->norm2_max, 0, sizeof(float), h->stream));
    if (ivf) {
        CK(cudaMalloc(&h->centroids, (size_t)nlist * h->ld * sizeof(float)));
        CK(cudaMemsetAsync(h->centroids, 0, (size_t)nlist * h->ld * sizeof(float), h->stream));
        CK(cudaMalloc(&h->list_off, (size_t)(nlist + 1) * sizeof(int64_t)));
        CK(cudaMemsetAsync(h->list_off, 0, (size_t)(nlist + 1) * sizeof(int64_t), h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));  // the zero-fills above must not race a first add on another stream
    *out = h;
    return 0;
}

extern "C" int wb_flat_create(int d, int device, wb_index** out) { return create_common(d, device, false, 0, out); }
extern "C" int wb_ivf_create(int d, int64_t nlist, int device, wb_index** out) {
    if (nlist <= 0 || nlist > (int64_t)1 << 30) return fail("invalid nlist %lld", (long long)nlist);
    return create_common(d, device, true, nlist, out);
}

extern "C" int wb_free(wb_index* h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    cudaFree(h->rows);
    cudaFree(h->ids);
    cudaFree(h->assign);
    cudaFree(h->centroids);
    cudaFree(h->norm2_max);
    cudaFree(h->slot_row);
    cudaFree(h->row_pos);
    cudaFree(h->list_off);
    for (DevBuf* b : {&h->parts, &h->qbuf, &h->dbuf, &h->ibuf, &h->pD, &h->pI, &h->xbuf, &h->idbuf, &h->misc,
                      &h->kperm, &h->koff, &h->gimg, &h->gimg2, &h->gkeys, &h->gstate, &h->gmargin, &h->eD, &h->eI, &h->tailcnt, &h->ccnt, &h->ctot, &h->lmA, &h->lmB, &h->lmC,
                      &h->lmD, &h->lmE, &h->lmF, &h->lmG, &h->ckeys, &h->phase})
        b->release();
    h->pin_q.release();
    h->pin_o.release();
    for (int i = 0; i < wb_index::kAddSlots; ++i)
        if (h->add_ev[i]) cudaEventDestroy(h->add_ev[i]);
    for (int i = 0; i < wb_index::kEvRing; ++i) {
        cudaEventDestroy(h->ev0[i]);
        cudaEventDestroy(h->ev1[i]);
    }
    cudaStreamDestroy(h->stream);
    delete h;
    return 0;
}

extern "C" int64_t wb_dim(const wb_index* h) { return h ? h->d : -1; }
extern "C" int64_t wb_ntotal(const wb_index* h) { return h ? h->n : -1; }
extern "C" int wb_is_trained(const wb_index* h) { return h && h->trained; }
extern "C" int64_t wb_nlist(const wb_index* h) { return h ? h->nlist : -1; }
extern "C" int wb_is_ivf(const wb_index* h) { return h && h->ivf; }
extern "C" int64_t wb_launch_count(const wb_index* h) { return h ? h->launches : -1; }
extern "C" int64_t wb_ivf_fused_searches(const wb_index* h) { return h ? h->coarse_fused : -1; }
constexpr size_t kPhaseCtas = 1024;  // phase stamps: CTAs of query group 0 (the kernel ignores blockIdx.x >= 1024)
extern "C" int wb_phase_stamps(wb_index* h, uint64_t* out_host, int ctas) {
    if (!h || !out_host) return fail("NULL argument");
    if (!h->phase.p) return fail("no phase stamps: set WB_PHASE_TS=1 before the search");
    if (ctas < 1 || (size_t)ctas > kPhaseCtas) return fail("ctas out of range");
    TRY(set_dev(h));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(out_host, h->phase.p, (size_t)ctas * 16 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return 0;
}
extern "C" int wb_gemm_stats(const wb_index* h, int64_t* epochs, int64_t* fallbacks) {
    if (!h) return fail("NULL index");
    if (epochs) *epochs = h->gemm_launches;
    if (fallbacks) *fallbacks = h->gemm_fallbacks;
    return 0;
}
extern "C" int wb_set_timing(wb_index* h, int on) {
    if (!h) return fail("NULL index");
    h->timing = on != 0;
    h->ev_count = 0;
    return 0;
}
static float scan_ms_at(wb_index* h, int64_t i) {
    const int s = (int)(i % wb_index::kEvRing);
    if (cudaEventSynchronize(h->ev1[s]) != cudaSuccess) return -1.f;
    float ms = -1.f;
    if (cudaEventElapsedTime(&ms, h->ev0[s], h->ev1[s]) != cudaSuccess) return -1.f;
    return ms;
}
extern "C" float wb_last_scan_ms(wb_index* h) {
    if (!h || h->ev_count == 0) return -1.f;
    if (cudaSetDevice(h->device) != cudaSuccess) return -1.f;
    return scan_ms_at(h, h->ev_count - 1);
}
extern "C" int wb_scan_ms_history(wb_index* h, float* out_ms, int cap) {
    if (!h || !out_ms || cap <= 0) return 0;
    if (cudaSetDevice(h->device) != cudaSuccess) return 0;
    const int64_t n = std::min<int64_t>(std::min<int64_t>(h->ev_count, wb_index::kEvRing), cap);
    for (int64_t i = 0; i < n; ++i) out_ms[i] = scan_ms_at(h, h->ev_count - n + i);
    return (int)n;
}
extern "C" int wb_storage(wb_index* h, void** rows_dev, int64_t* ld) {
    if (!h) return fail("NULL index");
    if (rows_dev) *rows_dev = h->rows;
    if (ld) *ld = h->ld;
    return 0;
}

// ---- capacity ------------------------------------------------------------------------------
static int ensure_capacity(wb_index* h, int64_t need) {
    if (need <= h->cap) return 0;
    if (need >= (int64_t)0xFFFFFFFEll) return fail("an index shard is limited to 2^32-2 rows per GPU");
    int64_t newcap = std::max<int64_t>(need, std::max<int64_t>(1024, h->cap + h->cap / 2));
    float* nrows = nullptr;
    int64_t* nids = nullptr;
    int32_t* nas = nullptr;
    CK(cudaMalloc(&nrows, (size_t)newcap * h->ld * sizeof(float)));
    CK(cudaMalloc(&nids, (size_t)newcap * sizeof(int64_t)));
    if (h->ivf) CK(cudaMalloc(&nas, (size_t)newcap * sizeof(int32_t)));
    if (h->row_pos && newcap > h->map_cap) {
        uint32_t* np_ = nullptr;
        CK(cudaMalloc(&np_, (size_t)newcap * sizeof(uint32_t)));
        if (h->n > 0) CK(cudaMemcpyAsync(np_, h->row_pos, (size_t)h->n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        cudaFree(h->row_pos);
        h->row_pos = np_;
        h->map_cap = newcap;
    }
    if (h->n > 0) {
        CK(cudaMemcpyAsync(nrows, h->rows, (size_t)h->n * h->ld * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
        CK(cudaMemcpyAsync(nids, h->ids, (size_t)h->n * sizeof(int64_t), cudaMemcpyDeviceToDevice, h->stream));
        if (h->ivf)
            CK(cudaMemcpyAsync(nas, h->assign, (size_t)h->n * sizeof(int32_t), cudaMemcpyDeviceToDevice, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    cudaFree(h->rows);
    cudaFree(h->ids);
    cudaFree(h->assign);
    h->rows = nrows;
    h->ids = nids;
    h->assign = nas;
    h->cap = newcap;
    return 0;
}

extern "C" int wb_reserve(wb_index* h, int64_t n) {
    if (!h) return fail("NULL index");
    TRY(set_dev(h));
    return ensure_capacity(h, n);
}

// ---- scan launch ---------------------------------------------------------------------------
struct ScanCfg {
    int NQ, RW, P, ck, nchunks, stages, single_copy;
    size_t smem;
};

static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

// Pick queries-per-CTA, rows-per-warp and ring depth so that everything fits the 227 KB of
// shared memory: [ring | queries | top-k lists + queues | barriers].
// Measured on B200 (profiles/r01_scan_sweep.md): one bulk copy engine request costs the producer
// ~27 cycles, so stages hold WHOLE rows - a flat row group is then one contiguous copy (7.2 TB/s);
// chunked stages (one 1 KB copy per row and chunk) top out at 4.0 TB/s and are only the fallback
// for rows too wide for two whole-row stages.
static int plan_scan(const wb_index* h, int64_t nq, int k, bool gather, int nprobe, ScanCfg* c) {
    const int ld = h->ld;
    int NQ = gather ? 1 : (nq >= 8 ? 8 : nq > 2 ? 4 : nq > 1 ? 2 : 1);
    NQ = std::min(NQ, std::max(1, env_int("WB_SCAN_NQ", 8)));
    const int P = pow2_ceil(k + kMinQueue);
    const int max_stages = std::max(2, env_int("WB_SCAN_STAGES", 8));
    const int rw_max = std::min(4, std::max(1, env_int("WB_SCAN_RW", 4)));
    const int ck_env = env_int("WB_SCAN_CK", 0) & ~3;
    const int np = gather ? nprobe : 0;
    for (; NQ >= 1; NQ >>= 1) {
        if (ck_env == 0 || ck_env >= ld) {
            for (int RW = rw_max; RW >= 1; RW >>= 1) {  // whole-row stages
                const int gr = kConsumerWarps * RW;
                const size_t fixed = scan_smem_layout(NQ, ld, P, ld, 0, np, gr).total;
                const size_t stage_bytes = (size_t)gr * ld * 4;
                if (fixed + 2 * stage_bytes + 64 > (size_t)h->smem_max) continue;
                if (stage_bytes >= (1u << 20)) continue;  // mbarrier tx-count limit
                const int stages = (int)std::min<size_t>(max_stages, ((size_t)h->smem_max - fixed - 64) / (stage_bytes + 24));
                *c = ScanCfg{NQ, RW, P, ld, 1, stages, gather ? 0 : 1, 0};
                c->smem = scan_smem_layout(NQ, ld, P, ld, stages, np, gr).total;
                return 0;
            }
        }
        const int RW = rw_max;  // chunked fallback: very wide rows
        const int gr = kConsumerWarps * RW;
        for (int ck = std::min(ck_env ? ck_env : 1024, ld) & ~3; ck >= 32; ck = (ck / 2) & ~3) {
            const size_t fixed = scan_smem_layout(NQ, ld, P, ck, 0, np, gr).total;
            const size_t stage_bytes = (size_t)gr * ck * 4;
            if (fixed + 2 * stage_bytes + 64 > (size_t)h->smem_max) continue;
            const int stages = (int)std::min<size_t>(max_stages, ((size_t)h->smem_max - fixed - 64) / (stage_bytes + 24));
            *c = ScanCfg{NQ, RW, P, ck, (ld + ck - 1) / ck, stages, 0, 0};
            c->smem = scan_smem_layout(NQ, ld, P, ck, stages, np, gr).total;
            return 0;
        }
    }
    return fail("scan does not fit shared memory (d=%d, k=%d)", h->d, k);
}

template <int NQ, int RW, bool GATHER>
static int launch_scan_t(const ScanParams& p, dim3 grid, size_t smem, cudaStream_t st) {
    static thread_local bool attr_done[64] = {};
    int dev = 0;
    CK(cudaGetDevice(&dev));
    if (dev >= 64 || !attr_done[dev]) {
        CK(cudaFuncSetAttribute(scan_topk_kernel<NQ, RW, GATHER>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        if (dev < 64) attr_done[dev] = true;
    }
    scan_topk_kernel<NQ, RW, GATHER><<<grid, kScanThreads, smem, st>>>(p);
    CK(cudaGetLastError());
    return 0;
}

template <int NQ, bool GATHER>
static int launch_scan_rw(int RW, const ScanParams& p, dim3 grid, size_t smem, cudaStream_t st) {
    switch (RW) {
        case 4: return launch_scan_t<NQ, 4, GATHER>(p, grid, smem, st);
        case 2: return launch_scan_t<NQ, 2, GATHER>(p, grid, smem, st);
        case 1: return launch_scan_t<NQ, 1, GATHER>(p, grid, smem, st);
    }
    return fail("bad RW %d", RW);
}

// Cooperative launch of the gather kernel (fused coarse quantizer: the CTAs of a query group meet at a barrier, so all
// of them must be resident - the cooperative launch guarantees it or fails).
template <int RW>
static cudaError_t launch_scan_coop_t(const ScanParams& p, dim3 grid, size_t smem, cudaStream_t st) {
    static thread_local bool attr_done[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 64 || !attr_done[dev]) {
        e = cudaFuncSetAttribute(scan_topk_kernel<1, RW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
        if (e != cudaSuccess) return e;
        if (dev < 64) attr_done[dev] = true;
    }
    void* args[] = {const_cast<ScanParams*>(&p)};
    return cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(scan_topk_kernel<1, RW, true>), grid,
                                       dim3(kScanThreads), args, smem, st);
}

static void apply_scan_cfg(const ScanCfg& c, ScanParams& p) {
    p.P = c.P;
    p.ck = c.ck;
    p.nchunks = c.nchunks;
    p.stages = c.stages;
    p.single_copy = c.single_copy;
    p.rw = c.RW;
    p.prefetch_idx = env_int("WB_GATHER_PREFETCH", 1);
    p.rank_sort = c.P <= kConsumerThreads && env_int("WB_RANK_SORT", 1);  // measured: slower than bitonic at P = 512
}

static cudaError_t launch_scan_coop(const ScanCfg& c, ScanParams& p, dim3 grid, cudaStream_t st) {
    apply_scan_cfg(c, p);
    switch (c.RW) {
        case 4: return launch_scan_coop_t<4>(p, grid, c.smem, st);
        case 2: return launch_scan_coop_t<2>(p, grid, c.smem, st);
        default: return launch_scan_coop_t<1>(p, grid, c.smem, st);
    }
}

static int launch_scan(const ScanCfg& c, bool gather, ScanParams& p, dim3 grid, cudaStream_t st) {
    apply_scan_cfg(c, p);
    if (gather) return launch_scan_rw<1, true>(c.RW, p, grid, c.smem, st);
    switch (c.NQ) {
        case 1: return launch_scan_rw<1, false>(c.RW, p, grid, c.smem, st);
        case 2: return launch_scan_rw<2, false>(c.RW, p, grid, c.smem, st);
        case 4: return launch_scan_rw<4, false>(c.RW, p, grid, c.smem, st);
        case 8: return launch_scan_rw<8, false>(c.RW, p, grid, c.smem, st);
    }
    return fail("bad NQ %d", c.NQ);
}

// Sort-buffer size of the merge: everything at once when it is small, else k + a 2048-entry queue.
static int merge_buffer_entries(int k, int64_t nparts, int nthreads = kMergeThreads) {
    const int64_t M = nparts * (int64_t)k;
    const int full = pow2_ceil(k + 2 * nthreads);
    if (M <= full) return std::max(2, pow2_ceil((int)std::max<int64_t>(M, k)));
    return full;
}

template <bool FROM_KEYS>
static int merge_smem_optin() {
    static thread_local bool done[64] = {};
    int dev = 0;
    CK(cudaGetDevice(&dev));
    if (dev >= 64 || !done[dev]) {
        CK(cudaFuncSetAttribute(merge_topk_kernel<FROM_KEYS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 8));
        if (dev < 64) done[dev] = true;
    }
    return 0;
}

static int launch_merge_keys(wb_index* h, int64_t nq, int k, int64_t nparts, const uint64_t* keys, const int64_t* ids,
                             float* D, int64_t* I, cudaStream_t st) {
    MergeParams m{};
    m.nq = nq;
    m.k = k;
    m.nparts = nparts;
    m.S = merge_buffer_entries(k, nparts);
    m.keys = keys;
    m.ids = ids;
    m.D = D;
    m.I = I;
    TRY(merge_smem_optin<true>());
    merge_topk_kernel<true><<<(unsigned)nq, kMergeThreads, (size_t)m.S * 8, st>>>(m);
    CK(cudaGetLastError());
    h->launches++;
    return 0;
}

// Arrival counters of the fused tail merge: zero between launches (the last CTA of a group resets its own).
constexpr int64_t kMaxGridY = 32768;
static int ensure_tail_counters(wb_index* h, cudaStream_t st) {
    if (h->tailcnt.p) return 0;
    // [arrivals | row-group dispensers | coarse-barrier arrivals]
    TRY(h->tailcnt.ensure((size_t)3 * kMaxGridY * sizeof(unsigned int)));
    CK(cudaMemsetAsync(h->tailcnt.p, 0, (size_t)3 * kMaxGridY * sizeof(unsigned int), st));
    return 0;
}

// Can the scan kernel's last CTA run the merge (and the multi-GPU exchange) itself?  The sort buffer and the staged
// local winners must fit the (then idle) ring.
static bool tail_fusable(const ScanCfg& c, int k, int64_t nparts, int world, int* S_merge, int* heads_bytes) {
    *S_merge = 0;
    *heads_bytes = 0;
    if (env_int("WB_FUSE_TAIL", 1) == 0) return false;
    const size_t ring_bytes = (size_t)c.stages * kConsumerWarps * c.RW * c.ck * 4;
    const int S = std::max(merge_buffer_entries(k, nparts, kConsumerThreads),
                           world > 1 ? merge_buffer_entries(k, world, kConsumerThreads) : 2);
    if ((size_t)S * 8 + (size_t)k * 12 + 64 > ring_bytes) return false;
    *S_merge = S;
    // merge through the heads of the sorted lists first (block_merge_heads, merge.cuh); the sort-based merge stays
    // behind it as the fallback, so both regions must fit
    if (env_int("WB_MERGE_HEADS", 1) && nparts <= 256) {
        const HeadsPlan a = heads_plan(k, (int)nparts);
        if (a.ok && std::max(a.bytes, (size_t)S * 8) + (size_t)k * 12 + 64 <= ring_bytes) *heads_bytes = (int)a.bytes;
    }
    return true;
}

static void set_tail(ScanParams& p, wb_index* h, const ScanCfg& c, int S_merge, int heads_bytes, const int64_t* ids, float* D,
                     int64_t* I, int k, const ExchParams* ex) {
    p.fuse_tail = 1;
    p.S_merge = S_merge;
    p.heads_bytes = heads_bytes;
    p.phase_ts = nullptr;
    if (env_int("WB_PHASE_TS", 0) && h->phase.ensure(kPhaseCtas * 16 * sizeof(unsigned long long)) == 0)
        p.phase_ts = h->phase.as<unsigned long long>();
    p.tail_count = h->tailcnt.as<unsigned int>();
    p.group_count = (c.nchunks == 1 && env_int("WB_SCAN_DYNAMIC", 1)) ? h->tailcnt.as<unsigned int>() + kMaxGridY : nullptr;
    p.ids = ids;
    p.D = D;
    p.I = I;
    p.exchange = 0;
    if (ex) {
        p.exchange = 1;
        p.exch = *ex;
        p.exch.S = merge_buffer_entries(k, ex->world, kConsumerThreads);
    }
}

// Exhaustive top-k of `nq` device queries (row stride ld) against `nrows` rows: K1 with the K3 merge (and, when `ex`
// is given, the multi-GPU exchange) fused into its tail; K1 + K3 as two launches when the tail cannot be fused.
// *exchanged tells the caller whether (D, I) already hold the GLOBAL result (written to ex->D / ex->I).
static int run_flat_scan(wb_index* h, const float* rows, int64_t nrows, const float* q_dev, int64_t nq, int k,
                         const int64_t* ids, float* D, int64_t* I, cudaStream_t st, bool timed,
                         const ExchParams* ex = nullptr, bool* exchanged = nullptr) {
    if (exchanged) *exchanged = false;
    ScanCfg c;
    TRY(plan_scan(h, nq, k, false, 0, &c));
    const int gr = kConsumerWarps * c.RW;
    const int64_t ngroups = (nrows + gr - 1) / gr;
    const int64_t qgroups_total = (nq + c.NQ - 1) / c.NQ;
    const int waves = std::max(1, env_int("WB_SCAN_WAVES", 1));
    int64_t S = std::min<int64_t>(std::max<int64_t>(ngroups, 1),
                                  std::max<int64_t>(1, ((int64_t)h->sm_count * waves * (qgroups_total > 1 ? 2 : 1) +
                                                        qgroups_total - 1) / qgroups_total));
    S = std::min<int64_t>(S, (int64_t)h->sm_count * waves);
    TRY(h->parts.ensure((size_t)nq * S * k * sizeof(uint64_t)));
    ScanParams p{};
    p.rows = rows;
    p.nrows = nrows;
    p.ld = h->ld;
    p.k = k;
    p.nparts = (int)S;
    int S_merge = 0, heads_bytes = 0;
    const bool fuse = tail_fusable(c, k, S, ex ? ex->world : 1, &S_merge, &heads_bytes);
    if (fuse) {
        TRY(ensure_tail_counters(h, st));
        set_tail(p, h, c, S_merge, heads_bytes, ids, ex ? ex->D : D, ex ? ex->I : I, k, ex);
    }
    const int evs = (int)(h->ev_count % wb_index::kEvRing);
    if (timed && h->timing) CK(cudaEventRecord(h->ev0[evs], st));
    for (int64_t g0 = 0; g0 < qgroups_total; g0 += kMaxGridY) {
        const int64_t gy = std::min(kMaxGridY, qgroups_total - g0);
        const int64_t qoff = g0 * c.NQ;
        p.queries = q_dev + (size_t)qoff * h->ld;
        p.nq = (int)std::min<int64_t>(nq - qoff, gy * c.NQ);
        p.parts = h->parts.as<uint64_t>() + (size_t)qoff * S * k;
        if (fuse) {  // the tail indexes (D, I) and the mailboxes by the query number inside this launch
            p.D = (ex ? ex->D : D) + (size_t)qoff * k;
            p.I = (ex ? ex->I : I) + (size_t)qoff * k;
            if (ex && qoff != 0) return fail("exchange-fused scans take at most %lld query groups", (long long)kMaxGridY);
        }
        TRY(launch_scan(c, false, p, dim3((unsigned)S, (unsigned)gy), st));
        h->launches++;
    }
    if (timed && h->timing) {
        CK(cudaEventRecord(h->ev1[evs], st));
        h->ev_count++;
    }
    if (fuse) {
        if (ex && exchanged) *exchanged = true;
        return 0;
    }
    return launch_merge_keys(h, nq, k, S, h->parts.as<uint64_t>(), ids, D, I, st);
}

// ---- K2: tensor-core batched scan ---------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int get_tensormap_encoder(PFN_encodeTiled* out) {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
        if (qres != cudaDriverEntryPointSuccess || !p) return fail("cuTensorMapEncodeTiled is not available in this driver");
        fn = (PFN_encodeTiled)p;
    }
    *out = fn;
    return 0;
}

// Cost model behind the K1 / K2 choice, measured on B200 (profiles/r01/small_store_latency.txt, batch sweeps of
// profiles/r01 and r02): an 8-query K1 pass streams the rows at ~4.2 TB/s-equivalent (FMA-bound at 8 queries per CTA);
// K2 streams them once at ~6 TB/s but pays ~0.3 ms per search for its epoch launches, compactions and the one
// stream synchronisation that reads the overflow flag.
constexpr double kK1BatchBytesPerS = 4.2e12;
constexpr double kK2StreamBytesPerS = 6.0e12;
constexpr double kK2FixedSeconds = 0.3e-3;

// Can the tensor-core path take this problem?  (candidate capacity `cap` rows form epoch 0)
static bool gemm_eligible(const wb_index* h, int64_t nrows, int64_t nq, int k, int* cap_out) {
    if (env_int("WB_GEMM", 1) == 0) return false;
    if (nq < env_int("WB_GEMM_MIN_NQ", 5)) return false;  // measured: K2 5.8 ms vs K1 8.1 ms at 5..8 queries (10M x 768)
    if (nrows >= ((int64_t)1 << 31) - 256) return false;  // TMA coordinates are int32
    int64_t cap = std::min<int64_t>(4096, std::max<int64_t>(256, ((nrows / 64) + 127) / 128 * 128));
    if (cap < 2 * (int64_t)k || nrows < 4 * cap) return false;
    // K2 pays ~0.3 ms of fixed cost per search (epoch launches, compactions, one stream sync) and then streams the
    // rows at ~6 TB/s; K1 needs ceil(nq/8) passes at ~4.2 TB/s-equivalent but has no fixed cost.  Small stores with
    // few queries are faster on K1 (e.g. 100k x 512 with 16 queries: 0.1 ms vs 0.35 ms).
    if (env_int("WB_GEMM_FORCE", 0) == 0) {
        const double bytes = (double)nrows * h->ld * 4.0;
        const double passes = (double)((nq + 7) / 8);
        if (passes * bytes / kK1BatchBytesPerS < kK2FixedSeconds + bytes / kK2StreamBytesPerS) return false;
    }
    *cap_out = (int)cap;
    return true;
}

template <int BN>
static int run_flat_gemm_t(wb_index* h, const float* rows, int64_t nrows, const float* q_ld, int64_t nq, int k, int cap,
                         const int64_t* ids, float* D, int64_t* I, cudaStream_t st, bool timed, bool* overflowed) {
    // Filter-and-refine (main row store only, whose norms are tracked): after the first epoch the GEMM issues only
    // the hi x hi term and the compaction re-scores the few hundred candidates per query exactly in fp32.
    const bool filter = rows == h->rows && env_int("WB_GEMM_FILTER", 1) != 0;
    constexpr bool kFold = kGemmFold<BN, false>;
    using Cfg = GemmCfg<BN, kFold>;
    constexpr int kGemmBN = BN;
    constexpr int kGemmBBytes = Cfg::kBBytes;
    constexpr size_t kGemmSmemBytes = Cfg::kSmemBytes;
    const int ld = h->ld;
    const int nchunks = (ld + kGemmBK - 1) / kGemmBK;
    const int nqb = (int)((nq + kGemmBN - 1) / kGemmBN);
    const int kstride = k + cap;
    // measured (10M x 768): 40 / 64 queries 5.9 / 6.3 ms on one CTA vs 6.2 / 6.4 ms on the pair; 128 queries 8.8 vs 8.1 ms
    const bool use2 = BN >= 128 && env_int("WB_GEMM_2CTA", 1) != 0 && (h->sm_count % 2) == 0;
    // Batches above 128 queries run their filter epochs in 256-query blocks on the SS kernel (gemm_ss.cuh): half as many
    // passes over the rows and 1.5x fewer L2 bytes per MAC than 128-query blocks.
    const bool use_f2 = BN >= 128 && filter && use2 && nq > 128 && env_int("WB_GEMM_F2", 1) != 0;
    // a single block of 65..128 queries: the same SS kernel with 64 queries per CTA (one pass over the rows, HBM-bound)
    const bool use_f2h = BN >= 128 && filter && use2 && nq <= 128 && env_int("WB_GEMM_F2", 1) != 0 &&
                         env_int("WB_GEMM_F2H", 1) != 0;
    const int nqb2 = (int)((nq + kF2BN - 1) / kF2BN);
    const int nq_pad = use_f2 ? nqb2 * kF2BN : nqb * kGemmBN;  // thresholds / margins of padding queries: +inf / 0
    TRY(h->gimg.ensure((size_t)nqb * nchunks * kGemmBBytes));
    TRY(h->gkeys.ensure((size_t)nq * kstride * sizeof(uint64_t)));
    TRY(h->gstate.ensure((size_t)nq_pad * 4 + (size_t)nq * 4 + 64 + 64 * 4));  // thr | cnt | overflow | one work dispenser per epoch
    float* thr = h->gstate.as<float>();
    float* margin = nullptr;
    if (filter) {
        TRY(h->gmargin.ensure((size_t)nq_pad * 4));
        margin = h->gmargin.as<float>();
        // |exact - hi.hi| <= 2^-9 |x||q| (two truncations to tf32, Cauchy-Schwarz) + fp32 accumulation slack
        // (n-term fp32 accumulation: <= n * 2^-23 * sum |x_i q_i| even if the tensor core truncates)
        const float c_margin = 1.953125e-3f + (float)ld * 1.1920929e-7f * 1.1f + 1e-5f;
        query_margin_kernel<<<(unsigned)((nq_pad * 32 + 255) / 256), 256, 0, st>>>(q_ld, (int)nq, nq_pad, ld, h->norm2_max,
                                                                                 c_margin, margin);
        CK(cudaGetLastError());
        h->launches++;
    }
    if (use_f2) {
        TRY(h->gimg2.ensure((size_t)nqb2 * 2 * nchunks * kF2BBytes));
        const int64_t n4 = (int64_t)nqb2 * 2 * nchunks * 8 * kF2Half;
        image_queries_f2_kernel<128><<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(q_ld, (int)nq, ld, nchunks, nqb2,
                                                                                  h->gimg2.as<float>());
        CK(cudaGetLastError());
        h->launches++;
    } else if (use_f2h) {
        TRY(h->gimg2.ensure((size_t)nqb * 2 * nchunks * F2Cfg<64>::kBBytes));
        const int64_t n4 = (int64_t)nqb * 2 * nchunks * 8 * 64;
        image_queries_f2_kernel<64><<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(q_ld, (int)nq, ld, nchunks, nqb,
                                                                                 h->gimg2.as<float>());
        CK(cudaGetLastError());
        h->launches++;
    }
    int* cnt = reinterpret_cast<int*>(thr + (size_t)nq_pad);
    int* overflow = cnt + nq;
    unsigned int* dispensers = reinterpret_cast<unsigned int*>(overflow + 16);  // [64], zeroed once per search
    const bool dyn_work = env_int("WB_GEMM_DYNAMIC", 1) != 0;
    if (dyn_work) CK(cudaMemsetAsync(dispensers, 0, 64 * sizeof(unsigned int), st));
    int epoch_no = 0;
    uint64_t* keys = h->gkeys.as<uint64_t>();
    {
        const int64_t n1 = std::max<int64_t>((int64_t)nq_pad, nq * (int64_t)k);
        init_gemm_state_kernel<<<(unsigned)((n1 + 255) / 256), 256, 0, st>>>(thr, (int)nq, nq_pad, cnt, keys, k,
                                                                            kstride, overflow);
        CK(cudaGetLastError());
        const int64_t n2 = (int64_t)nqb * nchunks * 8 * kGemmBN;
        if constexpr (BN >= 128) {
            if (use2) {  // the 2-CTA kernel wants one image per 64 queries (each CTA of a pair holds half a block)
                const int64_t n3 = (int64_t)(2 * nqb) * nchunks * 8 * (BN / 2);
                split_queries_kernel<BN / 2><<<(unsigned)((n3 + 255) / 256), 256, 0, st>>>(q_ld, (int)nq, ld, nchunks,
                                                                                      2 * nqb, h->gimg.as<float>());
            }
        }
        if (!use2) {
            split_queries_kernel<BN, kFold><<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(q_ld, (int)nq, ld, nchunks, nqb,
                                                                                     h->gimg.as<float>());
            if (kFold && filter) {  // the one-term epochs read a plain image (hi half only), the first epoch the folded one
                TRY(h->gimg2.ensure((size_t)nqb * nchunks * kGemmBBytes));
                split_queries_kernel<BN, false><<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(q_ld, (int)nq, ld, nchunks, nqb,
                                                                                         h->gimg2.as<float>());
                h->launches++;
            }
        }
        CK(cudaGetLastError());
        h->launches += 2;
    }
    PFN_encodeTiled encode = nullptr;
    TRY(get_tensormap_encoder(&encode));
    CUtensorMap tmap;
    {
        cuuint64_t gdim[2] = {(cuuint64_t)ld, (cuuint64_t)nrows};
        cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
        cuuint32_t box[2] = {(cuuint32_t)kGemmBK, (cuuint32_t)kGemmBM};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)rows, gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    }
    static thread_local bool attr_done[64] = {};
    int dev = 0;
    CK(cudaGetDevice(&dev));
    if (dev >= 64 || !attr_done[dev]) {
        CK(cudaFuncSetAttribute(gemm_topk_kernel<BN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmemBytes));
        CK(cudaFuncSetAttribute(gemm_topk_kernel<BN, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)GemmCfg<BN, false>::kSmemBytes));
        if (dev < 64) attr_done[dev] = true;
    }
    TRY(merge_smem_optin<true>());
    static thread_local bool cattr_done[64] = {};
    if (dev >= 64 || !cattr_done[dev]) {
        CK(cudaFuncSetAttribute(compact_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 8));
        if (dev < 64) cattr_done[dev] = true;
    }
    GemmParams g{};
    g.nchunks = nchunks;
    g.nq = (int)nq;
    g.nqb = nqb;
    g.bimg = h->gimg.as<float>();
    g.thr = thr;
    g.margin = margin;
    g.keys = keys;
    g.cnt = cnt;
    g.overflow = overflow;
    g.k = k;
    g.cap = cap;
    g.kstride = kstride;
    CompactParams c{};
    c.k = k;
    c.cap = cap;
    c.kstride = kstride;
    {
        const int64_t mmax = (int64_t)k + cap;
        const int full = pow2_ceil(k + 2 * kMergeThreads);
        c.S = mmax <= full ? pow2_ceil((int)mmax) : full;
    }
    c.keys = keys;
    c.cnt = cnt;
    c.thr = thr;
    c.ids = ids;
    c.rows = rows;
    c.q = q_ld;
    c.ld = ld;
    // next epoch = growth x (rows seen so far): a query then collects ~k*ln(1+growth) survivors per epoch, far
    // below `cap`; fewer, larger epochs mean fewer launch gaps and compactions (each costs ~50 us)
    const double growth = std::min(12.0, std::max(0.25, (double)cap / (4.0 * k)));
    const int evs = (int)(h->ev_count % wb_index::kEvRing);
    if (timed && h->timing) CK(cudaEventRecord(h->ev0[evs], st));
    int64_t r0 = 0;
    while (r0 < nrows) {
        int64_t len = r0 == 0 ? cap : (int64_t)((double)r0 * growth);
        len = std::max<int64_t>(kGemmBM, (len + kGemmBM - 1) / kGemmBM * kGemmBM);
        const int64_t r1 = std::min(nrows, r0 + len);
        g.row_begin = r0;
        g.row_end = r1;
        const int64_t nwork = ((r1 - r0 + kGemmBM - 1) / kGemmBM) * nqb;
        const unsigned grid = (unsigned)std::min<int64_t>(nwork, h->sm_count);
        // first epoch: 3xTF32 scores select, the winners are re-scored; later epochs: one-term filter, candidates re-scored
        const bool one_term = filter && r0 > 0;
        c.rescore = filter ? (r0 == 0 ? 2 : 1) : 0;
        bool launched = false;
        if constexpr (BN >= 128) {
            if (use_f2 && one_term) {  // 256-query blocks, both operands from shared memory
                static thread_local bool a3[64] = {};
                if (dev >= 64 || !a3[dev]) {
                    CK(cudaFuncSetAttribute(filter2_topk_kernel<128, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)kF2SmemBytes));
                    if (dev < 64) a3[dev] = true;
                }
                GemmParams g2 = g;
                g2.nqb = nqb2;
                g2.bimg = h->gimg2.as<float>();
                const int64_t ntp = ((r1 - r0 + kGemmBM - 1) / kGemmBM + 1) / 2;
                const unsigned grid3 = 2u * (unsigned)std::min<int64_t>(ntp * nqb2, h->sm_count / 2);
                filter2_topk_kernel<128, false, false><<<grid3, kF2Threads, kF2SmemBytes, st>>>(tmap, g2);
                launched = true;
            } else if (use_f2h && one_term) {
                static thread_local bool a5[64] = {};
                if (dev >= 64 || !a5[dev]) {
                    CK(cudaFuncSetAttribute(filter2_topk_kernel<64, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)F2Cfg<64>::kSmemBytes));
                    if (dev < 64) a5[dev] = true;
                }
                GemmParams g2 = g;
                g2.bimg = h->gimg2.as<float>();
                const int64_t ntp = ((r1 - r0 + kGemmBM - 1) / kGemmBM + 1) / 2;
                const unsigned grid3 = 2u * (unsigned)std::min<int64_t>(ntp * nqb, h->sm_count / 2);
                filter2_topk_kernel<64, false, false><<<grid3, kF2Threads, F2Cfg<64>::kSmemBytes, st>>>(tmap, g2);
                launched = true;
            }
        }
        if constexpr (BN >= 128) {  // the CTA pair exists for 128-query blocks only
            if (use2 && !launched) {
                static thread_local bool a2[64] = {};
                if (dev >= 64 || !a2[dev]) {
                    CK(cudaFuncSetAttribute(gemm2_topk_kernel<BN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)Gemm2Cfg<BN>::kSmemBytes));
                    CK(cudaFuncSetAttribute(gemm2_topk_kernel<BN, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)Gemm2Cfg<BN>::kSmemBytes));
                    if (dev < 64) a2[dev] = true;
                }
                const int64_t ntp = ((r1 - r0 + kGemmBM - 1) / kGemmBM + 1) / 2;
                const unsigned grid2 = 2u * (unsigned)std::min<int64_t>(ntp * nqb, h->sm_count / 2);
                if (one_term) gemm2_topk_kernel<BN, false, 1><<<grid2, kGemmThreads, Gemm2Cfg<BN>::kSmemBytes, st>>>(tmap, g);
                else gemm2_topk_kernel<BN, false><<<grid2, kGemmThreads, Gemm2Cfg<BN>::kSmemBytes, st>>>(tmap, g);
            }
        }
        if (!use2 && !launched) {
            // SMs stream at different rates: (tile, block) items are drawn from a dispenser, not dealt out statically
            g.work_counter = dyn_work && epoch_no < 64 ? dispensers + epoch_no : nullptr;
            if (one_term) {
                GemmParams g1 = g;
                if (kFold) g1.bimg = h->gimg2.as<float>();
                gemm_topk_kernel<BN, false, 1><<<grid, kGemmThreads, GemmCfg<BN, false>::kSmemBytes, st>>>(tmap, g1);
            } else {
                gemm_topk_kernel<BN, false><<<grid, kGemmThreads, kGemmSmemBytes, st>>>(tmap, g);
            }
        }
        CK(cudaGetLastError());
        const bool last = r1 >= nrows;
        c.D = last ? D : nullptr;
        c.I = last ? I : nullptr;
        if (c.rescore == 1) {
            const unsigned gy = (unsigned)std::max<int64_t>(1, std::min<int64_t>(64, (int64_t)h->sm_count * 4 / nq));
            rescore_candidates_kernel<<<dim3((unsigned)nq, gy), 256, 0, st>>>(c);
            CK(cudaGetLastError());
            h->launches++;
        }
        compact_topk_kernel<<<(unsigned)nq, kMergeThreads, (size_t)c.S * 8, st>>>(c);
        CK(cudaGetLastError());
        h->launches += 2;
        h->gemm_launches++;
        epoch_no++;
        r0 = r1;
    }
    if (timed && h->timing) {
        CK(cudaEventRecord(h->ev1[evs], st));
        h->ev_count++;
    }
    int ovf = 0;
    CK(cudaMemcpyAsync(&ovf, overflow, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *overflowed = ovf != 0;
    return 0;
}

static int run_flat_gemm(wb_index* h, const float* rows, int64_t nrows, const float* q_ld, int64_t nq, int k, int cap,
                         const int64_t* ids, float* D, int64_t* I, cudaStream_t st, bool timed, bool* overflowed) {
    const int bn_max = env_int("WB_GEMM_BN", 128);
    // 16-query blocks halve the image traffic and the MMA work of a half-empty 32-query block (K2 is power-capped at
    // small batches: 992 W and 1.37 GHz at batch 16)
    if (nq <= 16 && bn_max >= 16 && env_int("WB_GEMM_BN16", 1))
        return run_flat_gemm_t<16>(h, rows, nrows, q_ld, nq, k, cap, ids, D, I, st, timed, overflowed);
    if (nq <= 32 && bn_max >= 32) return run_flat_gemm_t<32>(h, rows, nrows, q_ld, nq, k, cap, ids, D, I, st, timed, overflowed);
    if ((nq <= 64 && bn_max >= 64) || bn_max < 128)
        return run_flat_gemm_t<64>(h, rows, nrows, q_ld, nq, k, cap, ids, D, I, st, timed, overflowed);
    return run_flat_gemm_t<128>(h, rows, nrows, q_ld, nq, k, cap, ids, D, I, st, timed, overflowed);
}

// K4/K6 on the tensor cores: argmax over the centroid table for n device-resident points (row stride ld).
// One launch, no epochs, no host synchronisation: the points are the A operand (streamed once from HBM), the
// centroids are the query images (L2 resident), each row keeps its running maximum in the epilogue.
static bool assign_gemm_eligible(const wb_index* h, int64_t n) {
    return env_int("WB_GEMM", 1) != 0 && env_int("WB_GEMM_ASSIGN", 1) != 0 && n >= 1024 && h->nlist >= 64 &&
           n < ((int64_t)1 << 31) - 256;
}

static int run_assign_gemm(wb_index* h, const float* x_ld, int64_t n, int32_t* assign_out, float* best_out,
                           cudaStream_t st, bool tf32 = false) {
    constexpr int BN = 128;
    using Cfg = GemmCfg<BN>;
    const int ld = h->ld;
    const int nchunks = (ld + kGemmBK - 1) / kGemmBK;
    if (tf32 && (h->sm_count % 2) == 0 && env_int("WB_GEMM_2CTA", 1) != 0) {
        // k-means TRAINING assignment: plain TF32 on the SS pair kernel, 256 centroids per block (gemm_ss.cuh)
        const int nqb2 = (int)((h->nlist + kF2BN - 1) / kF2BN);
        TRY(h->gimg2.ensure((size_t)nqb2 * 2 * nchunks * kF2BBytes));
        const int64_t n4 = (int64_t)nqb2 * 2 * nchunks * 8 * kF2Half;
        image_queries_f2_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(h->centroids, (int)h->nlist, ld, nchunks, nqb2,
                                                                             h->gimg2.as<float>());
        CK(cudaGetLastError());
        PFN_encodeTiled enc = nullptr;
        TRY(get_tensormap_encoder(&enc));
        CUtensorMap tm;
        cuuint64_t gdim[2] = {(cuuint64_t)ld, (cuuint64_t)n};
        cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
        cuuint32_t box[2] = {(cuuint32_t)kGemmBK, (cuuint32_t)kGemmBM};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)x_ld, gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        static thread_local bool a4[64] = {};
        int dev = 0;
        CK(cudaGetDevice(&dev));
        if (dev >= 64 || !a4[dev]) {
            CK(cudaFuncSetAttribute(filter2_topk_kernel<128, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)kF2SmemBytes));
            if (dev < 64) a4[dev] = true;
        }
        GemmParams g{};
        g.row_begin = 0;
        g.row_end = n;
        g.nchunks = nchunks;
        g.nq = (int)h->nlist;
        g.nqb = nqb2;
        g.bimg = h->gimg2.as<float>();
        g.assign_out = assign_out;
        g.best_out = best_out;
        const int64_t ntp = ((n + kGemmBM - 1) / kGemmBM + 1) / 2;
        filter2_topk_kernel<128, true, false><<<2u * (unsigned)std::min<int64_t>(ntp, h->sm_count / 2), kF2Threads, kF2SmemBytes, st>>>(tm, g);
        CK(cudaGetLastError());
        h->launches += 2;
        h->gemm_launches++;
        return 0;
    }
    const int nqb = (int)((h->nlist + BN - 1) / BN);
    TRY(h->gimg.ensure((size_t)nqb * nchunks * Cfg::kBBytes));
    const int64_t n2 = (int64_t)nqb * nchunks * 8 * BN;
    const bool use2 = env_int("WB_GEMM_2CTA", 1) != 0 && (h->sm_count % 2) == 0;
    if (use2) {
        const int64_t n3 = (int64_t)(2 * nqb) * nchunks * 8 * (BN / 2);
        split_queries_kernel<BN / 2><<<(unsigned)((n3 + 255) / 256), 256, 0, st>>>(h->centroids, (int)h->nlist, ld, nchunks,
                                                                              2 * nqb, h->gimg.as<float>());
    } else {
        split_queries_kernel<BN><<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(h->centroids, (int)h->nlist, ld, nchunks, nqb,
                                                                              h->gimg.as<float>());
    }
    CK(cudaGetLastError());
    PFN_encodeTiled encode = nullptr;
    TRY(get_tensormap_encoder(&encode));
    CUtensorMap tmap;
    cuuint64_t gdim[2] = {(cuuint64_t)ld, (cuuint64_t)n};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)kGemmBK, (cuuint32_t)kGemmBM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)x_ld, gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    static thread_local bool attr_done[64] = {};
    int dev = 0;
    CK(cudaGetDevice(&dev));
    if (dev >= 64 || !attr_done[dev]) {
        CK(cudaFuncSetAttribute(gemm_topk_kernel<BN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmemBytes));
        if (dev < 64) attr_done[dev] = true;
    }
    GemmParams g{};
    g.row_begin = 0;
    g.row_end = n;
    g.nchunks = nchunks;
    g.nq = (int)h->nlist;
    g.nqb = nqb;
    g.bimg = h->gimg.as<float>();
    g.assign_out = assign_out;
    g.best_out = best_out;
    const int64_t ntiles = (n + kGemmBM - 1) / kGemmBM;
    if (use2) {
        static thread_local bool a2[64] = {};
        if (dev >= 64 || !a2[dev]) {
            CK(cudaFuncSetAttribute(gemm2_topk_kernel<BN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)Gemm2Cfg<BN>::kSmemBytes));
            if (dev < 64) a2[dev] = true;
        }
        const int64_t ntp = (ntiles + 1) / 2;
        gemm2_topk_kernel<BN, true><<<2u * (unsigned)std::min<int64_t>(ntp, h->sm_count / 2), kGemmThreads,
                                     Gemm2Cfg<BN>::kSmemBytes, st>>>(tmap, g);
    } else {
        gemm_topk_kernel<BN, true><<<(unsigned)std::min<int64_t>(ntiles, h->sm_count), kGemmThreads, Cfg::kSmemBytes, st>>>(tmap, g);
    }
    CK(cudaGetLastError());
    h->launches += 2;
    h->gemm_launches++;
    return 0;
}

// Exhaustive top-k, any batch size: tensor cores (K2) when the batch is big enough, else the
// bandwidth-bound CUDA-core scan (K1).  A candidate-list overflow in K2 (adversarially ordered data)
// is repaired by re-running the batch through K1.
static int run_flat_any(wb_index* h, const float* rows, int64_t nrows, const float* q_dev, int64_t nq, int k,
                        const int64_t* ids, float* D, int64_t* I, cudaStream_t st, bool timed,
                        const ExchParams* ex = nullptr, bool* exchanged = nullptr) {
    if (exchanged) *exchanged = false;
    int cap = 0;
    if (gemm_eligible(h, nrows, nq, k, &cap)) {
        const int64_t max_q = 65536;  // bounds the candidate buffers
        bool any_overflow = false;
        for (int64_t q0 = 0; q0 < nq; q0 += max_q) {
            const int64_t nqc = std::min(max_q, nq - q0);
            bool ovf = false;
            TRY(run_flat_gemm(h, rows, nrows, q_dev + (size_t)q0 * h->ld, nqc, k, cap, ids, D + (size_t)q0 * k,
                              I + (size_t)q0 * k, st, timed, &ovf));
            any_overflow |= ovf;
        }
        if (!any_overflow) return 0;
        h->gemm_fallbacks++;
    }
    return run_flat_scan(h, rows, nrows, q_dev, nq, k, ids, D, I, st, timed,
                         ex && nq <= kMaxGridY && env_int("WB_FUSE_EXCH", 1) ? ex : nullptr, exchanged);
}

// ---- K8: CSR inverted lists, built on the device (csr.cuh) -------------------------------------------------------
// Stable grouping of n items by list: list_off[nlist + 1], src[slot] = item index, pos_out[slot] = pos_in[item].
// Everything is enqueued on `st`; the only host involvement is reading one error word when `check` is set.
static int device_csr(wb_index* h, const int32_t* assign, int64_t n, const uint32_t* pos_in, int64_t* list_off,
                      uint32_t* src, uint32_t* pos_out, cudaStream_t st, bool check) {
    const int64_t nlist = h->nlist;
    if (n >= (int64_t)0xFFFFFFFFll) return fail("too many items to group");
    // B blocks of items: enough to fill the GPU, few enough for the [B][nlist] counter matrix to stay small
    int64_t B = std::min<int64_t>(1024, std::max<int64_t>(1, (n + 4095) / 4096));
    B = std::max<int64_t>(1, std::min<int64_t>(B, ((int64_t)64 << 20) / std::max<int64_t>(nlist, 1)));
    int64_t ipb = (n + B - 1) / B;
    ipb = std::max<int64_t>(kCsrTile, (ipb + kCsrTile - 1) / kCsrTile * kCsrTile);
    B = std::max<int64_t>(1, (n + ipb - 1) / ipb);
    TRY(h->ccnt.ensure((size_t)B * nlist * sizeof(uint32_t) + 64));
    TRY(h->ctot.ensure((size_t)nlist * sizeof(int64_t) + 64));
    uint32_t* cnt = h->ccnt.as<uint32_t>();
    int* bad = reinterpret_cast<int*>(h->ctot.as<unsigned char>() + (size_t)nlist * sizeof(int64_t));
    CK(cudaMemsetAsync(cnt, 0, (size_t)B * nlist * sizeof(uint32_t), st));
    CK(cudaMemsetAsync(bad, 0, sizeof(int), st));
    if (n > 0) {
        csr_hist_kernel<<<(unsigned)B, kCsrTile, 0, st>>>(assign, n, nlist, ipb, cnt, bad);
        CK(cudaGetLastError());
    }
    csr_colscan_kernel<<<(unsigned)((nlist + 255) / 256), 256, 0, st>>>(cnt, (int)B, nlist, h->ctot.as<int64_t>());
    CK(cudaGetLastError());
    csr_offsets_kernel<<<1, 1024, 0, st>>>(h->ctot.as<int64_t>(), nlist, list_off);
    CK(cudaGetLastError());
    if (n > 0) {
        csr_scatter_kernel<<<(unsigned)B, kCsrTile, 0, st>>>(assign, n, nlist, ipb, cnt, list_off, pos_in, src, pos_out);
        CK(cudaGetLastError());
    }
    h->launches += 4;
    if (check) {
        int hb = 0;
        CK(cudaMemcpyAsync(&hb, bad, sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (hb) return fail("a row has a list assignment outside [0, %lld)", (long long)nlist);
    }
    return 0;
}

// Build the inverted lists.  Default: PHYSICALLY group the row store by list (rows of a list become one
// contiguous span, insertion order kept inside a list), so the list scan is a sequential HBM stream instead of
// a gather of 2-3 KB rows (measured 3.3-4.3 TB/s gathered).  Needs a second copy of the rows while it runs; if
// that does not fit, the lists stay a CSR of row indices (slot_row) over the store as it is.
static int ensure_csr(wb_index* h) {
    if (!h->csr_dirty) return 0;
    cudaStream_t st = h->stream;
    const int64_t n = h->n;
    cudaFree(h->slot_row);
    h->slot_row = nullptr;
    if (n == 0) {
        CK(cudaMemsetAsync(h->list_off, 0, (size_t)(h->nlist + 1) * sizeof(int64_t), st));
        CK(cudaStreamSynchronize(st));
        h->csr_dirty = false;
        return 0;
    }
    uint32_t* src = nullptr;
    uint32_t* pos_out = nullptr;
    const int64_t mcap = std::max<int64_t>(h->cap, n);
    CK(cudaMalloc(&src, (size_t)n * sizeof(uint32_t)));
    CK(cudaMalloc(&pos_out, (size_t)mcap * sizeof(uint32_t)));
    int rc = device_csr(h, h->assign, n, h->row_pos, h->list_off, src, pos_out, st, true);
    if (rc) {
        cudaFree(src);
        cudaFree(pos_out);
        return rc;
    }
    bool grouped = false;
    if (env_int("WB_IVF_CONTIGUOUS", 1)) {
        const int64_t ncap = std::max<int64_t>(h->cap, std::max<int64_t>(n, 1024));
        const size_t need = (size_t)ncap * h->ld * 4 + (size_t)ncap * 4;
        size_t free_b = 0, total_b = 0;
        CK(cudaMemGetInfo(&free_b, &total_b));
        if (free_b > need + ((size_t)1 << 30)) {
            float* nrows = nullptr;
            int32_t* nas = nullptr;
            CK(cudaMalloc(&nrows, (size_t)ncap * h->ld * 4));
            CK(cudaMalloc(&nas, (size_t)ncap * 4));
            const int ld4 = h->ld / 4;
            const int64_t tot = n * ld4;
            permute_rows_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(
                reinterpret_cast<const float4*>(h->rows), reinterpret_cast<float4*>(nrows), src, n, ld4);
            CK(cudaGetLastError());
            permute_i32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(h->assign, nas, src, n);
            CK(cudaGetLastError());
            h->launches += 2;
            CK(cudaStreamSynchronize(st));
            cudaFree(h->rows);
            cudaFree(h->assign);
            h->rows = nrows;
            h->assign = nas;
            if (ncap > h->cap) {  // ids keep their insertion order; they only follow the capacity
                int64_t* nids = nullptr;
                CK(cudaMalloc(&nids, (size_t)ncap * 8));
                CK(cudaMemcpy(nids, h->ids, (size_t)n * 8, cudaMemcpyDeviceToDevice));
                cudaFree(h->ids);
                h->ids = nids;
            }
            h->cap = ncap;
            cudaFree(h->row_pos);
            h->row_pos = pos_out;  // storage row -> insertion position
            h->map_cap = mcap;
            pos_out = nullptr;
            cudaFree(src);
            src = nullptr;
            grouped = true;
        }
    }
    if (!grouped) {  // rows stay put: lists are row indices, row_pos (if any) still describes the storage
        CK(cudaStreamSynchronize(st));
        h->slot_row = src;
        cudaFree(pos_out);
    }
    h->csr_dirty = false;
    return 0;
}

// ---- add -----------------------------------------------------------------------------------
// assign rows [n0, n0+n) of the store to their max-inner-product centroid (K4 with k = 1)
static int assign_rows(wb_index* h, const float* x_dev, int64_t n, int32_t* assign_out, float* best_out,
                       cudaStream_t st, bool tf32 = false) {
    if (assign_gemm_eligible(h, n)) return run_assign_gemm(h, x_dev, n, assign_out, best_out, st, tf32);
    TRY(h->pD.ensure((size_t)n * sizeof(float)));
    TRY(h->pI.ensure((size_t)n * sizeof(int64_t)));
    float* D = best_out ? best_out : h->pD.as<float>();
    TRY(run_flat_any(h, h->centroids, h->nlist, x_dev, n, 1, nullptr, D, h->pI.as<int64_t>(), st, false));
    labels_to_lists_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(h->pI.as<int64_t>(), assign_out, n);
    CK(cudaGetLastError());
    h->launches++;
    return 0;
}

static int add_common(wb_index* h, int64_t n, const float* x, const int64_t* ids, bool host, const int32_t* preassign,
                      cudaStream_t st, bool sync_host = true) {
    if (!h) return fail("NULL index");
    if (n < 0) return fail("negative n");
    if (n == 0) return 0;
    if (!x) return fail("x is NULL");
    TRY(set_dev(h));
    if (h->ivf && !h->trained) return fail("IndexIVFFlat must be trained before adding vectors");
    if (h->n + n > h->cap && st != h->stream) CK(cudaStreamSynchronize(st));  // growth copies run on the handle's stream
    TRY(ensure_capacity(h, h->n + n));
    const int64_t n0 = h->n;
    float* dst = h->rows + (size_t)n0 * h->ld;
    const cudaMemcpyKind kind = host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    if (h->ld == h->d) {
        CK(cudaMemcpyAsync(dst, x, (size_t)n * h->d * sizeof(float), kind, st));
    } else {
        const float* src = x;
        if (host) {
            TRY(h->xbuf.ensure((size_t)n * h->d * sizeof(float)));
            CK(cudaMemcpyAsync(h->xbuf.p, x, (size_t)n * h->d * sizeof(float), kind, st));
            src = h->xbuf.as<float>();
        }
        const int64_t tot = n * h->ld;
        pad_rows_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(src, dst, n, h->d, h->ld);
        CK(cudaGetLastError());
        h->launches++;
    }
    {   // running max of |row|^2: the K2 filter epochs derive their safety margin from it
        const unsigned blocks = (unsigned)std::min<int64_t>((n + 7) / 8, (int64_t)h->sm_count * 8);
        row_norm2_max_kernel<<<blocks, 256, 0, st>>>(dst, n, h->ld, h->norm2_max);
        CK(cudaGetLastError());
        h->launches++;
    }
    if (ids) {
        CK(cudaMemcpyAsync(h->ids + n0, ids, (size_t)n * sizeof(int64_t), kind, st));
    } else {
        iota_ids_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(h->ids + n0, n0, n);
        CK(cudaGetLastError());
        h->launches++;
    }
    if (h->row_pos) {  // the store has been grouped before: new rows sit at their insertion position
        iota_u32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(h->row_pos, n0, n);
        CK(cudaGetLastError());
        h->launches++;
    }
    if (h->ivf) {
        if (preassign) {
            CK(cudaMemcpyAsync(h->assign + n0, preassign, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        } else {
            TRY(assign_rows(h, dst, n, h->assign + n0, nullptr, st));
        }
        h->csr_dirty = true;
    }
    h->n += n;
    if (host && sync_host) CK(cudaStreamSynchronize(st));  // the caller may reuse its buffers on return
    return 0;
}

extern "C" int wb_add_with_ids(wb_index* h, int64_t n, const float* x_host, const int64_t* ids_host) {
    return add_common(h, n, x_host, ids_host, true, nullptr, h ? h->stream : nullptr);
}
// ---- pinned, overlapped ingest (FeatureStore shards -> HBM) ------------------------------------------------------
// The loader decodes shard i+1 into one pinned buffer while the copy engine moves shard i out of another and the
// SMs assign its rows to their lists: wb_add_with_ids_pinned enqueues (H2D, norms, ids, IVF assignment) on the
// handle's stream and returns at once; `slot` names the pinned buffer, wb_add_slot_wait(slot) blocks until the GPU has
// finished reading it.  Replaces the synchronous per-batch add of feature_search_index.py:79-82.
extern "C" int wb_pinned_alloc(int64_t bytes, void** out) {
    if (!out || bytes <= 0) return fail("bad arguments");
    *out = nullptr;
    CK(cudaHostAlloc(out, (size_t)bytes, cudaHostAllocDefault));
    return 0;
}
extern "C" int wb_pinned_free(void* p) {
    if (p) CK(cudaFreeHost(p));
    return 0;
}
extern "C" int wb_add_with_ids_pinned(wb_index* h, int64_t n, const float* x_pinned, const int64_t* ids_pinned, int slot) {
    if (!h) return fail("NULL index");
    if (slot < 0 || slot >= wb_index::kAddSlots) return fail("slot %d out of range [0, %d)", slot, wb_index::kAddSlots);
    TRY(set_dev(h));
    if (!h->add_ev[slot]) CK(cudaEventCreateWithFlags(&h->add_ev[slot], cudaEventDisableTiming));
    TRY(add_common(h, n, x_pinned, ids_pinned, true, nullptr, h->stream, false));
    CK(cudaEventRecord(h->add_ev[slot], h->stream));
    h->add_pending[slot] = true;
    return 0;
}
// Same for rows whose list is already known (faiss.read_index of an IndexIVFFlat file): no coarse quantisation.
extern "C" int wb_ivf_add_preassigned_pinned(wb_index* h, int64_t n, const float* x_pinned, const int64_t* ids_pinned,
                                             const int32_t* assign_pinned, int slot) {
    if (!h || !h->ivf) return fail("not an IVF index");
    if (!assign_pinned) return fail("assign is NULL");
    if (slot < 0 || slot >= wb_index::kAddSlots) return fail("slot %d out of range [0, %d)", slot, wb_index::kAddSlots);
    TRY(set_dev(h));
    if (!h->add_ev[slot]) CK(cudaEventCreateWithFlags(&h->add_ev[slot], cudaEventDisableTiming));
    TRY(add_common(h, n, x_pinned, ids_pinned, true, assign_pinned, h->stream, false));
    CK(cudaEventRecord(h->add_ev[slot], h->stream));
    h->add_pending[slot] = true;
    return 0;
}
extern "C" int wb_add_slot_wait(wb_index* h, int slot) {
    if (!h) return fail("NULL index");
    if (slot < 0 || slot >= wb_index::kAddSlots) return fail("slot %d out of range [0, %d)", slot, wb_index::kAddSlots);
    if (!h->add_pending[slot]) return 0;
    TRY(set_dev(h));
    CK(cudaEventSynchronize(h->add_ev[slot]));
    h->add_pending[slot] = false;
    return 0;
}
extern "C" int wb_sync(wb_index* h) {
    if (!h) return fail("NULL index");
    TRY(set_dev(h));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int wb_add_with_ids_dev(wb_index* h, int64_t n, const float* x_dev, const int64_t* ids_dev, void* stream) {
    return add_common(h, n, x_dev, ids_dev, false, nullptr, (cudaStream_t)stream);
}
extern "C" int wb_ivf_add_preassigned(wb_index* h, int64_t n, const float* x_host, const int64_t* ids_host,
                                      const int32_t* assign_host) {
    if (!h || !h->ivf) return fail("not an IVF index");
    if (!assign_host) return fail("assign is NULL");
    return add_common(h, n, x_host, ids_host, true, assign_host, h->stream);
}

// ---- K5 for batches: list-major scan (ivf_lm.cuh) ------------------------------------------------------------------
template <int RW>
static int launch_lm_t(const LmParams& p, int grid, size_t smem, cudaStream_t st) {
    static thread_local bool attr_done[64] = {};
    int dev = 0;
    CK(cudaGetDevice(&dev));
    if (dev >= 64 || !attr_done[dev]) {
        CK(cudaFuncSetAttribute(ivf_listmajor_kernel<8, RW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        if (dev < 64) attr_done[dev] = true;
    }
    ivf_listmajor_kernel<8, RW><<<grid, kScanThreads, smem, st>>>(p);
    CK(cudaGetLastError());
    return 0;
}

// Taken when the probes of the batch overlap enough for shared reads to pay (on average every list is probed at
// least once) and the rows are physically grouped by list.  WB_IVF_LISTMAJOR=0 / 1 forces it off / on.
static int run_ivf_listmajor(wb_index* h, int64_t nq, const float* q_ld, int k, int np, float* D, int64_t* I,
                             cudaStream_t st, bool* done) {
    *done = false;
    const int mode = env_int("WB_IVF_LISTMAJOR", -1);
    if (mode == 0 || h->slot_row != nullptr || h->n == 0) return 0;
    const int64_t npairs = nq * (int64_t)np;
    if (mode < 0 && (nq < 8 || npairs < h->nlist)) return 0;
    if (npairs >= ((int64_t)1 << 31) || (double)npairs * k * 8.0 > 2e9) return 0;
    ScanCfg c;
    if (plan_scan(h, 8, k, false, 0, &c) != 0 || c.NQ != 8 || !c.single_copy) return 0;
    const int64_t nlist = h->nlist;
    const int64_t max_items = npairs / 8 + nlist + 1;
    TRY(h->lmA.ensure((size_t)npairs * 4));              // list of every (query, probe) pair
    TRY(h->lmB.ensure((size_t)npairs * 4));              // pairs grouped by list
    TRY(h->lmC.ensure((size_t)(nlist + 1) * 8));         // pairs per list (CSR offsets)
    TRY(h->lmD.ensure((size_t)(nlist + 1) * 8));         // work items per list
    TRY(h->lmE.ensure((size_t)(nlist + 1) * 8));         // first work item of every list
    TRY(h->lmF.ensure((size_t)max_items * 4));           // list of every work item
    TRY(h->lmG.ensure((size_t)nq * 4 + 64));             // shared thresholds + the work counter
    TRY(h->parts.ensure((size_t)npairs * k * sizeof(uint64_t)));
    probes_to_pairs_kernel<<<(unsigned)((npairs + 255) / 256), 256, 0, st>>>(h->pI.as<int64_t>(), h->lmA.as<int32_t>(), npairs, k,
                                                                            h->parts.as<uint64_t>());
    CK(cudaGetLastError());
    TRY(device_csr(h, h->lmA.as<int32_t>(), npairs, nullptr, h->lmC.as<int64_t>(), h->lmB.as<uint32_t>(), nullptr, st, false));
    lm_item_count_kernel<<<(unsigned)((nlist + 255) / 256), 256, 0, st>>>(h->lmC.as<int64_t>(), nlist, 8, h->lmD.as<int64_t>());
    CK(cudaGetLastError());
    csr_offsets_kernel<<<1, 1024, 0, st>>>(h->lmD.as<int64_t>(), nlist, h->lmE.as<int64_t>());
    CK(cudaGetLastError());
    lm_item_fill_kernel<<<(unsigned)((nlist + 255) / 256), 256, 0, st>>>(h->lmE.as<int64_t>(), nlist, h->lmF.as<int32_t>());
    CK(cudaGetLastError());
    CK(cudaMemsetAsync(h->lmG.p, 0, (size_t)nq * 4 + 64, st));
    LmParams p{};
    p.rows = h->rows;
    p.ld = h->ld;
    p.queries = q_ld;
    p.nq = (int)nq;
    p.k = k;
    p.P = c.P;
    p.stages = c.stages;
    p.nprobe = np;
    p.nlist = nlist;
    p.list_off = h->list_off;
    p.row_pos = h->row_pos;
    p.pl_off = h->lmC.as<int64_t>();
    p.pair_src = h->lmB.as<uint32_t>();
    p.item_off = h->lmE.as<int64_t>();
    p.item_list = h->lmF.as<int32_t>();
    p.gthr = h->lmG.as<uint32_t>();
    p.counter = reinterpret_cast<unsigned int*>(h->lmG.as<unsigned char>() + (size_t)nq * 4 + 16);
    p.parts = h->parts.as<uint64_t>();
    const int evs = (int)(h->ev_count % wb_index::kEvRing);
    if (h->timing) CK(cudaEventRecord(h->ev0[evs], st));
    const int grid = (int)std::min<int64_t>(h->sm_count, max_items);
    switch (c.RW) {
        case 4: TRY(launch_lm_t<4>(p, grid, c.smem, st)); break;
        case 2: TRY(launch_lm_t<2>(p, grid, c.smem, st)); break;
        default: TRY(launch_lm_t<1>(p, grid, c.smem, st)); break;
    }
    if (h->timing) {
        CK(cudaEventRecord(h->ev1[evs], st));
        h->ev_count++;
    }
    h->launches += 4;
    TRY(launch_merge_keys(h, nq, k, np, h->parts.as<uint64_t>(), h->ids, D, I, st));
    *done = true;
    return 0;
}

// The query-major list scan: one launch per <= 32768 queries (the probe table is in h->pI), tail merge fused or not.
static int run_gather_scan(wb_index* h, const ScanCfg& c, ScanParams& p, int64_t S, bool fuse, int64_t nq, const float* q_ld,
                           int k, int np, float* D, int64_t* I, cudaStream_t st, const ExchParams* ex, bool* exchanged) {
    const int evs = (int)(h->ev_count % wb_index::kEvRing);
    if (h->timing) CK(cudaEventRecord(h->ev0[evs], st));
    for (int64_t q0 = 0; q0 < nq; q0 += kMaxGridY) {
        const int64_t gy = std::min(kMaxGridY, nq - q0);
        p.queries = q_ld + (size_t)q0 * h->ld;
        p.nq = (int)gy;
        p.probes = h->pI.as<int64_t>() + (size_t)q0 * np;
        p.parts = h->parts.as<uint64_t>() + (size_t)q0 * S * k;
        if (fuse) {
            p.D = (ex ? ex->D : D) + (size_t)q0 * k;
            p.I = (ex ? ex->I : I) + (size_t)q0 * k;
        }
        TRY(launch_scan(c, true, p, dim3((unsigned)S, (unsigned)gy), st));
        h->launches++;
    }
    if (h->timing) {
        CK(cudaEventRecord(h->ev1[evs], st));
        h->ev_count++;
    }
    if (fuse) {
        if (ex && exchanged) *exchanged = true;
        return 0;
    }
    return launch_merge_keys(h, nq, (int)k, S, h->parts.as<uint64_t>(), h->ids, D, I, st);
}

// ---- search --------------------------------------------------------------------------------
static int search_dev_impl(wb_index* h, int64_t nq, const float* q_ld /* [nq, ld] */, int64_t k, int64_t nprobe, float* D,
                           int64_t* I, cudaStream_t st, const ExchParams* ex = nullptr, bool* exchanged = nullptr) {
    if (exchanged) *exchanged = false;
    if (!h->ivf) return run_flat_any(h, h->rows, h->n, q_ld, nq, (int)k, h->ids, D, I, st, true, ex, exchanged);
    if (!h->trained) return fail("IndexIVFFlat is not trained");
    int np = (int)std::min<int64_t>(std::max<int64_t>(nprobe, 1), std::min<int64_t>(h->nlist, WB_MAX_K));
    if (h->csr_dirty) {
        if (st != h->stream) CK(cudaStreamSynchronize(st));
        TRY(ensure_csr(h));
    }
    // K5 plan: gather-scan of the probed lists, S CTAs per query
    ScanCfg c;
    TRY(plan_scan(h, 1, (int)k, true, np, &c));
    int64_t S = nq >= h->sm_count ? 1 : std::max<int64_t>(1, (int64_t)h->sm_count / nq);
    if (ex && (nq > kMaxGridY || !env_int("WB_FUSE_EXCH", 1))) ex = nullptr;
    int S_merge = 0, heads_bytes = 0;
    const bool fuse = tail_fusable(c, (int)k, S, ex ? ex->world : 1, &S_merge, &heads_bytes);
    // K4 + K5 in one launch (coarse.cuh): a few queries, every CTA resident, the selection scratch fits the idle ring
    const size_t ring_bytes = (size_t)c.stages * kConsumerWarps * c.RW * c.ck * 4;
    bool fuse_coarse = fuse && h->coop_ok && nq <= env_int("WB_IVF_FUSE_COARSE_MAXQ", 4) && S * nq <= h->sm_count &&
                       h->nlist < ((int64_t)1 << 30) && coarse_smem_layout(h->nlist, np).total <= ring_bytes &&
                       env_int("WB_IVF_FUSE_COARSE", 1) && env_int("WB_IVF_LISTMAJOR", -1) != 1;
    for (int attempt = 0; attempt < 2; ++attempt) {
        if (!fuse_coarse) {
            // K4: coarse quantizer = exhaustive scan of the centroids, top-nprobe
            TRY(h->pD.ensure((size_t)nq * np * sizeof(float)));
            TRY(h->pI.ensure((size_t)nq * np * sizeof(int64_t)));
            TRY(run_flat_any(h, h->centroids, h->nlist, q_ld, nq, np, nullptr, h->pD.as<float>(), h->pI.as<int64_t>(), st, false));
            // K5, batches: list-major - the probe table is inverted on the device and every probed list is streamed once
            // per group of 8 of the queries that probe it (ivf_lm.cuh)
            bool done = false;
            TRY(run_ivf_listmajor(h, nq, q_ld, (int)k, np, D, I, st, &done));
            if (done) return 0;
        }
        TRY(h->parts.ensure((size_t)nq * S * k * sizeof(uint64_t)));
        ScanParams p{};
        p.rows = h->rows;
        p.nrows = 0;
        p.ld = h->ld;
        p.k = (int)k;
        p.nparts = (int)S;
        p.perm = h->slot_row;
        p.row_pos = h->row_pos;
        p.list_off = h->list_off;
        p.nprobe = np;
        if (fuse) {
            TRY(ensure_tail_counters(h, st));
            set_tail(p, h, c, S_merge, heads_bytes, h->ids, ex ? ex->D : D, ex ? ex->I : I, (int)k, ex);
        }
        if (fuse_coarse) {
            TRY(h->ckeys.ensure((size_t)nq * h->nlist * sizeof(uint32_t)));
            p.centroids = h->centroids;
            p.nlist = h->nlist;
            p.coarse_keys = h->ckeys.as<uint32_t>();
            p.coarse_count = h->tailcnt.as<unsigned int>() + 2 * kMaxGridY;
            p.queries = q_ld;
            p.nq = (int)nq;
            p.parts = h->parts.as<uint64_t>();
            const int evs = (int)(h->ev_count % wb_index::kEvRing);
            if (h->timing) CK(cudaEventRecord(h->ev0[evs], st));
            const cudaError_t e = launch_scan_coop(c, p, dim3((unsigned)S, (unsigned)nq), st);
            if (e != cudaSuccess) {
                // (a partitioned device can refuse a full-width cooperative grid: two launches from now on)
                cudaGetLastError();
                h->coop_ok = false;
                fuse_coarse = false;
                continue;
            }
            if (h->timing) {
                CK(cudaEventRecord(h->ev1[evs], st));
                h->ev_count++;
            }
            h->launches++;
            h->coarse_fused++;
            if (ex && exchanged) *exchanged = true;
            return 0;
        }
        return run_gather_scan(h, c, p, S, fuse, nq, q_ld, (int)k, np, D, I, st, ex, exchanged);
    }
    return fail("unreachable");
}

static int check_search_args(const wb_index* h, int64_t nq, const void* q, int64_t k, const void* D, const void* I) {
    if (!h) return fail("NULL index");
    if (nq < 0) return fail("negative nq");
    if (k < 1 || k > WB_MAX_K) return fail("k=%lld out of range [1, %d]", (long long)k, WB_MAX_K);
    if (nq > 0 && (!q || !D || !I)) return fail("NULL buffer");
    return 0;
}

constexpr size_t kPinStageMax = (size_t)4 << 20;  // larger transfers are bandwidth-bound anyway: copy them directly

// queries arrive with row stride d; the kernels want stride ld
static int stage_queries(wb_index* h, int64_t nq, const float* q, bool host, cudaStream_t st, const float** out) {
    if (!host && h->ld == h->d) {
        *out = q;
        return 0;
    }
    const size_t qbytes = (size_t)nq * h->d * sizeof(float);
    if (host && qbytes <= kPinStageMax) {  // pageable -> pinned (host memcpy) -> device (async)
        PinBuf& pb = h->pin_q;
        TRY(pb.ensure(qbytes));
        memcpy(pb.p, q, qbytes);
        q = reinterpret_cast<const float*>(pb.p);
    }
    TRY(h->qbuf.ensure((size_t)nq * h->ld * sizeof(float)));
    if (h->ld == h->d) {
        CK(cudaMemcpyAsync(h->qbuf.p, q, qbytes, cudaMemcpyHostToDevice, st));
    } else {
        const float* src = q;
        if (host) {
            TRY(h->xbuf.ensure(qbytes));
            CK(cudaMemcpyAsync(h->xbuf.p, q, qbytes, cudaMemcpyHostToDevice, st));
            src = h->xbuf.as<float>();
        }
        const int64_t tot = nq * h->ld;
        pad_rows_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(src, h->qbuf.as<float>(), nq, h->d, h->ld);
        CK(cudaGetLastError());
        h->launches++;
    }
    *out = h->qbuf.as<float>();
    return 0;
}

// (D, I) device -> caller's host buffers, then ONE synchronisation of the stream.
// `word_dev` (optional): one more device int fetched with the results (the exchange's time-out flag), so that reading
// it costs no synchronisation of its own.
static int fetch_results(wb_index* h, int64_t nq, int64_t k, const float* D_dev, const int64_t* I_dev, float* D_host,
                         int64_t* I_host, cudaStream_t st) {
    const size_t dbytes = (size_t)nq * k * sizeof(float), ibytes = (size_t)nq * k * sizeof(int64_t);
    PinBuf& pb = h->pin_o;
    if (dbytes + ibytes <= kPinStageMax) {
        TRY(pb.ensure(dbytes + ibytes));
        unsigned char* stage = reinterpret_cast<unsigned char*>(pb.p);
        CK(cudaMemcpyAsync(stage, I_dev, ibytes, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(stage + ibytes, D_dev, dbytes, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        memcpy(I_host, stage, ibytes);
        memcpy(D_host, stage + ibytes, dbytes);
    } else {
        CK(cudaMemcpyAsync(D_host, D_dev, dbytes, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(I_host, I_dev, ibytes, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    return 0;
}

// Small result sets skip the device buffer and its two D2H copies: the kernel that emits the final (D, I) rows writes
// them straight into the pinned staging buffer (a few KB of posted PCIe writes), which the host reads after the
// stream synchronisation.  At one query this takes ~10 us off a 60 us IVF call.
constexpr size_t kDirectResultMax = (size_t)64 << 10;
static bool direct_results(wb_index* h, int64_t nq, int64_t k) {
    return (size_t)nq * k * 12 <= kDirectResultMax && env_int("WB_DIRECT_RESULTS", 1);
}
static int direct_result_buffers(wb_index* h, int64_t nq, int64_t k, float** D_pin, int64_t** I_pin) {
    const size_t dbytes = (size_t)nq * k * sizeof(float), ibytes = (size_t)nq * k * sizeof(int64_t);
    TRY(h->pin_o.ensure(dbytes + ibytes));
    *I_pin = reinterpret_cast<int64_t*>(h->pin_o.p);
    *D_pin = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(h->pin_o.p) + ibytes);
    return 0;
}
static int finish_direct_results(int64_t nq, int64_t k, const float* D_pin, const int64_t* I_pin, float* D_host,
                                 int64_t* I_host, cudaStream_t st) {
    CK(cudaStreamSynchronize(st));
    memcpy(I_host, I_pin, (size_t)nq * k * sizeof(int64_t));
    memcpy(D_host, D_pin, (size_t)nq * k * sizeof(float));
    return 0;
}

extern "C" int wb_search_dev(wb_index* h, int64_t nq, const float* q_dev, int64_t k, int64_t nprobe, float* D_dev,
                             int64_t* I_dev, void* stream) {
    TRY(check_search_args(h, nq, q_dev, k, D_dev, I_dev));
    if (nq == 0) return 0;
    TRY(set_dev(h));
    cudaStream_t st = (cudaStream_t)stream;
    const float* q = nullptr;
    TRY(stage_queries(h, nq, q_dev, false, st, &q));
    return search_dev_impl(h, nq, q, k, nprobe, D_dev, I_dev, st);
}

extern "C" int wb_search(wb_index* h, int64_t nq, const float* q_host, int64_t k, int64_t nprobe, float* D_host,
                         int64_t* I_host) {
    TRY(check_search_args(h, nq, q_host, k, D_host, I_host));
    if (nq == 0) return 0;
    TRY(set_dev(h));
    cudaStream_t st = h->stream;
    const float* q = nullptr;
    TRY(stage_queries(h, nq, q_host, true, st, &q));
    if (direct_results(h, nq, k)) {
        float* Dp = nullptr;
        int64_t* Ip = nullptr;
        TRY(direct_result_buffers(h, nq, k, &Dp, &Ip));
        TRY(search_dev_impl(h, nq, q, k, nprobe, Dp, Ip, st));
        return finish_direct_results(nq, k, Dp, Ip, D_host, I_host, st);
    }
    TRY(h->dbuf.ensure((size_t)nq * k * sizeof(float)));
    TRY(h->ibuf.ensure((size_t)nq * k * sizeof(int64_t)));
    TRY(search_dev_impl(h, nq, q, k, nprobe, h->dbuf.as<float>(), h->ibuf.as<int64_t>(), st));
    return fetch_results(h, nq, k, h->dbuf.as<float>(), h->ibuf.as<int64_t>(), D_host, I_host, st);
}

extern "C" int wb_merge_topk_dev(int device, int64_t nq, int64_t k, int64_t nparts, const float* D_parts_dev,
                                 const int64_t* I_parts_dev, float* D_dev, int64_t* I_dev, void* stream) {
    if (k < 1 || k > WB_MAX_K) return fail("k=%lld out of range [1, %d]", (long long)k, WB_MAX_K);
    if (nq <= 0 || nparts <= 0) return fail("empty merge");
    if (nparts * k >= (int64_t)0xFFFFFFFFll) return fail("too many candidates to merge");
    CK(cudaSetDevice(device));
    MergeParams m{};
    m.nq = nq;
    m.k = (int)k;
    m.nparts = nparts;
    m.S = merge_buffer_entries((int)k, nparts);
    m.Dp = D_parts_dev;
    m.Ip = I_parts_dev;
    m.D = D_dev;
    m.I = I_dev;
    TRY(merge_smem_optin<false>());
    merge_topk_kernel<false><<<(unsigned)nq, kMergeThreads, (size_t)m.S * 8, (cudaStream_t)stream>>>(m);
    CK(cudaGetLastError());
    return 0;
}

// ---- IVF: make the inverted lists current (K8) and describe them ---------------------------------------------------
// faiss keeps its inverted lists current on every add; this backend groups the row store lazily, at the first search
// after an add.  write_index calls this first so that every list is one contiguous run of storage rows.
extern "C" int wb_ivf_finalize(wb_index* h) {
    if (!h || !h->ivf) return fail("not an IVF index");
    TRY(set_dev(h));
    return ensure_csr(h);
}
// list_off_host[nlist + 1]: list l holds the storage rows [list_off[l], list_off[l+1]) once wb_ivf_finalize has run
// and *grouped_out is 1 (0: the rows could not be regrouped for lack of memory; use wb_export_rows + assignments).
extern "C" int wb_ivf_list_offsets(wb_index* h, int64_t* list_off_host, int* grouped_out) {
    if (!h || !h->ivf) return fail("not an IVF index");
    if (!list_off_host) return fail("NULL buffer");
    TRY(set_dev(h));
    TRY(ensure_csr(h));
    CK(cudaMemcpyAsync(list_off_host, h->list_off, (size_t)(h->nlist + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (grouped_out) *grouped_out = h->slot_row == nullptr;
    return 0;
}

// ---- row access ----------------------------------------------------------------------------
extern "C" int wb_reconstruct_batch(wb_index* h, int64_t m, const int64_t* ids_host, float* out_host) {
    if (!h) return fail("NULL index");
    if (m < 0) return fail("negative m");
    if (m == 0) return 0;
    if (!ids_host || !out_host) return fail("NULL buffer");
    TRY(set_dev(h));
    cudaStream_t st = h->stream;
    TRY(h->idbuf.ensure((size_t)m * 2 * sizeof(int64_t)));
    TRY(h->xbuf.ensure((size_t)m * h->d * sizeof(float)));
    int64_t* tgt = h->idbuf.as<int64_t>();
    unsigned long long* pos = reinterpret_cast<unsigned long long*>(tgt + m);
    CK(cudaMemcpyAsync(tgt, ids_host, (size_t)m * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(pos, 0xFF, (size_t)m * sizeof(int64_t), st));
    for (int64_t j0 = 0; j0 < m; j0 += 64) {
        const int mm = (int)std::min<int64_t>(64, m - j0);
        const unsigned grid = (unsigned)std::min<int64_t>(std::max<int64_t>((h->n + 255) / 256, 1), h->sm_count * 8);
        find_ids_kernel<int64_t><<<grid, 256, 0, st>>>(h->ids, h->n, tgt + j0, mm, pos + j0);
        CK(cudaGetLastError());
        h->launches++;
    }
    std::vector<int64_t> hpos((size_t)m);
    CK(cudaMemcpyAsync(hpos.data(), pos, (size_t)m * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (int64_t j = 0; j < m; ++j)
        if (hpos[j] < 0) return fail("reconstruct: id %lld not found in the index", (long long)ids_host[j]);
    if (h->row_pos) {  // grouped IVF store: insertion position -> storage row
        CK(cudaMemcpyAsync(tgt, pos, (size_t)m * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
        CK(cudaMemsetAsync(pos, 0xFF, (size_t)m * sizeof(int64_t), st));
        for (int64_t j0 = 0; j0 < m; j0 += 64) {
            const int mm = (int)std::min<int64_t>(64, m - j0);
            const unsigned grid = (unsigned)std::min<int64_t>(std::max<int64_t>((h->n + 255) / 256, 1), h->sm_count * 8);
            find_ids_kernel<uint32_t><<<grid, 256, 0, st>>>(h->row_pos, h->n, tgt + j0, mm, pos + j0);
            CK(cudaGetLastError());
            h->launches++;
        }
    }
    const int64_t tot = m * h->d;
    gather_rows_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(h->rows, h->ld, h->d, (const int64_t*)pos, m,
                                                                       h->xbuf.as<float>());
    CK(cudaGetLastError());
    h->launches++;
    CK(cudaMemcpyAsync(out_host, h->xbuf.p, (size_t)tot * sizeof(float), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int wb_export_rows(wb_index* h, int64_t start, int64_t n, float* x_host, int64_t* ids_host,
                              int32_t* assign_host) {
    if (!h) return fail("NULL index");
    if (start < 0 || n < 0 || start + n > h->n) return fail("export range [%lld,%lld) outside [0,%lld)",
                                                            (long long)start, (long long)(start + n), (long long)h->n);
    if (n == 0) return 0;
    TRY(set_dev(h));
    cudaStream_t st = h->stream;
    if (x_host)
        CK(cudaMemcpy2DAsync(x_host, (size_t)h->d * 4, h->rows + (size_t)start * h->ld, (size_t)h->ld * 4,
                             (size_t)h->d * 4, (size_t)n, cudaMemcpyDeviceToHost, st));
    if (ids_host) {
        if (h->row_pos) {  // grouped IVF store: the id of storage row r is ids[row_pos[r]]
            TRY(h->idbuf.ensure((size_t)n * 8));
            gather_ids_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(h->ids, h->row_pos, start, n, h->idbuf.as<int64_t>());
            CK(cudaGetLastError());
            h->launches++;
            CK(cudaMemcpyAsync(ids_host, h->idbuf.p, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
        } else {
            CK(cudaMemcpyAsync(ids_host, h->ids + start, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
        }
    }
    if (assign_host) {
        if (!h->ivf) return fail("assignments exist only for IVF indices");
        CK(cudaMemcpyAsync(assign_host, h->assign + start, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    return 0;
}

// ---- IVF quantizer / k-means -----------------------------------------------------------------
extern "C" int wb_ivf_set_centroids(wb_index* h, const float* c) {
    if (!h || !h->ivf) return fail("not an IVF index");
    if (!c) return fail("NULL centroids");
    if (h->n > 0) return fail("cannot replace the centroids of a non-empty index");
    TRY(set_dev(h));
    CK(cudaMemsetAsync(h->centroids, 0, (size_t)h->nlist * h->ld * 4, h->stream));
    CK(cudaMemcpy2DAsync(h->centroids, (size_t)h->ld * 4, c, (size_t)h->d * 4, (size_t)h->d * 4, (size_t)h->nlist,
                         cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->trained = true;
    return 0;
}

// faiss ClusteringParameters::spherical.  Default 0: an IndexIVFFlat built with the plain constructor, as the
// reference does (feature_search_index.py:60), trains with spherical = false [faiss-upstream]; index_factory is what
// switches it on for inner-product indices.
extern "C" int wb_ivf_set_spherical(wb_index* h, int on) {
    if (!h || !h->ivf) return fail("not an IVF index");
    h->spherical = on != 0;
    return 0;
}

extern "C" int wb_ivf_mark_trained(wb_index* h) {
    if (!h || !h->ivf) return fail("not an IVF index");
    h->trained = true;
    return 0;
}

extern "C" int wb_ivf_get_centroids(const wb_index* h, float* c) {
    if (!h || !h->ivf) return fail("not an IVF index");
    if (!c) return fail("NULL centroids");
    TRY(set_dev(h));
    CK(cudaMemcpy2DAsync(c, (size_t)h->d * 4, h->centroids, (size_t)h->ld * 4, (size_t)h->d * 4, (size_t)h->nlist,
                         cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

// x_dev has row stride d; kernels need stride ld
static int stage_points(wb_index* h, int64_t n, const float* x_dev, cudaStream_t st, const float** out) {
    if (h->ld == h->d) {
        *out = x_dev;
        return 0;
    }
    TRY(h->xbuf.ensure((size_t)n * h->ld * sizeof(float)));
    const int64_t tot = n * h->ld;
    pad_rows_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(x_dev, h->xbuf.as<float>(), n, h->d, h->ld);
    CK(cudaGetLastError());
    h->launches++;
    *out = h->xbuf.as<float>();
    return 0;
}

static int kmeans_assign_impl(wb_index* h, int64_t n, const float* x_dev, int32_t* assign_dev, double* objective_host,
                              void* stream, bool tf32);
extern "C" int wb_kmeans_assign_dev(wb_index* h, int64_t n, const float* x_dev, int32_t* assign_dev,
                                    double* objective_host, void* stream) {
    return kmeans_assign_impl(h, n, x_dev, assign_dev, objective_host, stream, false);
}
// TRAINING-only variant: plain-TF32 scores (one tensor-core term instead of three), ~3x less tensor work.
extern "C" int wb_kmeans_assign_fast_dev(wb_index* h, int64_t n, const float* x_dev, int32_t* assign_dev,
                                         double* objective_host, void* stream) {
    return kmeans_assign_impl(h, n, x_dev, assign_dev, objective_host, stream, env_int("WB_KMEANS_EXACT", 0) == 0);
}
static int kmeans_assign_impl(wb_index* h, int64_t n, const float* x_dev, int32_t* assign_dev, double* objective_host,
                              void* stream, bool tf32) {
    if (!h || !h->ivf) return fail("not an IVF index");
    if (n <= 0 || !x_dev || !assign_dev) return fail("bad arguments");
    TRY(set_dev(h));
    cudaStream_t st = (cudaStream_t)stream;
    const float* x = nullptr;
    TRY(stage_points(h, n, x_dev, st, &x));
    TRY(h->dbuf.ensure((size_t)n * sizeof(float)));
    TRY(assign_rows(h, x, n, assign_dev, h->dbuf.as<float>(), st, tf32));
    if (objective_host) {
        TRY(h->misc.ensure(sizeof(double)));
        CK(cudaMemsetAsync(h->misc.p, 0, sizeof(double), st));
        sum_f32_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, 1024), 256, 0, st>>>(h->dbuf.as<float>(), n,
                                                                                          h->misc.as<double>());
        CK(cudaGetLastError());
        h->launches++;
        CK(cudaMemcpyAsync(objective_host, h->misc.p, sizeof(double), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    return 0;
}

extern "C" int wb_kmeans_accumulate_dev(wb_index* h, int64_t n, const float* x_dev, const int32_t* assign_dev,
                                        float* sums_dev, int64_t* counts_dev, void* stream) {
    if (!h || !h->ivf) return fail("not an IVF index");
    if (n < 0 || !sums_dev || !counts_dev) return fail("bad arguments");
    TRY(set_dev(h));
    cudaStream_t st = (cudaStream_t)stream;
    TRY(h->kperm.ensure(std::max<size_t>((size_t)n * 4, 16)));
    TRY(h->koff.ensure((size_t)(h->nlist + 1) * 8));
    // group the points by their centroid on the device (stable: every list is summed in insertion order, so the
    // result does not depend on the schedule); no host copy, no synchronisation
    TRY(device_csr(h, assign_dev, n, nullptr, h->koff.as<int64_t>(), h->kperm.as<uint32_t>(), nullptr, st, false));
    segment_sum_kernel<<<(unsigned)h->nlist, 256, 0, st>>>(x_dev, h->d, h->d, h->kperm.as<uint32_t>(),
                                                          h->koff.as<int64_t>(), sums_dev, counts_dev);
    CK(cudaGetLastError());
    h->launches++;
    return 0;
}

extern "C" int wb_kmeans_update_dev(wb_index* h, const float* sums_dev, const int64_t* counts_dev, int64_t n_total,
                                    int64_t seed, int64_t* nsplit_host, void* stream) {
    if (!h || !h->ivf) return fail("not an IVF index");
    TRY(set_dev(h));
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t k = h->nlist;
    const int d = h->d, ld = h->ld;
    const int64_t tot = k * d;
    centroid_mean_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(sums_dev, counts_dev, h->centroids, ld, d, k);
    CK(cudaGetLastError());
    h->launches++;
    // split_clusters [faiss-upstream]: an empty list takes a perturbed copy of a list picked with probability ~ its
    // size; eps = 1/1024.  The pick is a sequential draw over the list sizes (host, k counters); the copies are applied
    // on the device (split_apply_kernel) - the centroid table never leaves HBM.
    std::vector<int64_t> cnt((size_t)k);
    CK(cudaMemcpyAsync(cnt.data(), counts_dev, (size_t)k * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    int64_t nsplit = 0;
    bool any_empty = false;
    for (int64_t c = 0; c < k; ++c) any_empty |= cnt[c] == 0;
    if (any_empty && n_total > k) {
        std::mt19937 mt((uint32_t)seed);
        std::vector<double> hass(cnt.begin(), cnt.end());
        const double denom = (double)(n_total - k);
        std::vector<float> prob((size_t)k);  // acceptance probability of every list, kept current
        for (int64_t c = 0; c < k; ++c) prob[c] = (float)((hass[c] - 1.0) / denom);
        std::vector<int32_t> pairs;
        for (int64_t ci = 0; ci < k; ++ci) {
            if (hass[ci] != 0) continue;
            int64_t cj = 0;
            for (int64_t guard = 0;; cj = (cj + 1 == k ? 0 : cj + 1)) {
                // (float)mt() / (float)mt.max(): the divisor is 2^32 as a float, so the product below is the same bits
                const float r = (float)mt() * 2.3283064365386963e-10f;
                if (r < prob[cj]) break;
                if (++guard > k * 1000) return fail("split_clusters did not converge");
            }
            pairs.push_back((int32_t)ci);
            pairs.push_back((int32_t)cj);
            hass[ci] = (double)((int64_t)hass[cj] / 2);
            hass[cj] -= hass[ci];
            prob[ci] = (float)((hass[ci] - 1.0) / denom);
            prob[cj] = (float)((hass[cj] - 1.0) / denom);
            nsplit++;
        }
        TRY(h->misc.ensure(pairs.size() * sizeof(int32_t) + 16));
        CK(cudaMemcpyAsync(h->misc.p, pairs.data(), pairs.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        split_apply_kernel<<<(unsigned)((d + 255) / 256), 256, 0, st>>>(h->centroids, ld, d, h->misc.as<int32_t>(), (int)nsplit);
        CK(cudaGetLastError());
        h->launches++;
        CK(cudaStreamSynchronize(st));  // `pairs` (host) is read by the copy above
    }
    if (h->spherical) {
        renorm_rows_kernel<<<(unsigned)k, 256, 0, st>>>(h->centroids, ld, d);
        CK(cudaGetLastError());
        h->launches++;
    }
    if (nsplit_host) *nsplit_host = nsplit;
    return 0;
}

extern "C" int wb_ivf_train(wb_index* h, int64_t n, const float* x_host, int niter, int64_t seed) {
    if (!h || !h->ivf) return fail("not an IVF index");
    if (h->n > 0) return fail("cannot train a non-empty index");
    if (n < h->nlist) return fail("Number of training points (%lld) should be at least as large as number of clusters (%lld)",
                                   (long long)n, (long long)h->nlist);
    if (!x_host) return fail("NULL training set");
    TRY(set_dev(h));
    cudaStream_t st = h->stream;
    const int64_t k = h->nlist;
    const int d = h->d;
    float* x = nullptr;
    CK(cudaMalloc(&x, (size_t)n * d * 4));
    int rc = 0;
    int32_t* assign = nullptr;
    float* sums = nullptr;
    int64_t* counts = nullptr;
    int* nonfinite = nullptr;
    g_err.clear();  // the error path below reports a CUDA error only when no step has already described its failure
    do {
        if ((rc = (cudaMemcpyAsync(x, x_host, (size_t)n * d * 4, cudaMemcpyHostToDevice, st) != cudaSuccess))) break;
        if ((rc = (cudaMalloc(&assign, (size_t)n * 4) != cudaSuccess))) break;
        if ((rc = (cudaMalloc(&sums, (size_t)k * d * 4) != cudaSuccess))) break;
        if ((rc = (cudaMalloc(&counts, (size_t)k * 8) != cudaSuccess))) break;
        // faiss Clustering::train: "input contains NaN's or Inf's" - checked on the device, read at the first sync
        if ((rc = (cudaMalloc(&nonfinite, sizeof(int)) != cudaSuccess))) break;
        if ((rc = (cudaMemsetAsync(nonfinite, 0, sizeof(int), st) != cudaSuccess))) break;
        nonfinite_flag_kernel<<<(unsigned)(h->sm_count * 8), 256, 0, st>>>(x, n * (int64_t)d, nonfinite);
        // initial centroids: first k rows of the faiss-style random permutation, seed + 1
        std::mt19937 mt((uint32_t)(seed + 1));
        std::vector<int64_t> perm((size_t)n);
        for (int64_t i = 0; i < n; ++i) perm[i] = i;
        for (int64_t i = 0; i + 1 < n; ++i) {
            const int64_t i2 = i + (int64_t)(mt() % (uint64_t)(n - i));
            std::swap(perm[i], perm[i2]);
        }
        if ((rc = (cudaMemsetAsync(h->centroids, 0, (size_t)k * h->ld * 4, st) != cudaSuccess))) break;
        {   // the k picked rows, gathered by one kernel
            if ((rc = h->idbuf.ensure((size_t)k * 8))) break;
            if ((rc = (cudaMemcpyAsync(h->idbuf.p, perm.data(), (size_t)k * 8, cudaMemcpyHostToDevice, st) != cudaSuccess))) break;
            if (h->ld == d) {
                gather_rows_kernel<<<(unsigned)((k * d + 255) / 256), 256, 0, st>>>(x, d, d, h->idbuf.as<int64_t>(), k, h->centroids);
            } else {
                if ((rc = h->xbuf.ensure((size_t)k * d * 4))) break;
                gather_rows_kernel<<<(unsigned)((k * d + 255) / 256), 256, 0, st>>>(x, d, d, h->idbuf.as<int64_t>(), k, h->xbuf.as<float>());
                pad_rows_kernel<<<(unsigned)((k * h->ld + 255) / 256), 256, 0, st>>>(h->xbuf.as<float>(), h->centroids, k, d, h->ld);
            }
            h->launches++;
            if ((rc = (cudaStreamSynchronize(st) != cudaSuccess))) break;  // perm (host) is read by the copy above
            int bad = 0;
            if ((rc = (cudaMemcpy(&bad, nonfinite, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess))) break;
            if (bad) {
                rc = fail("input contains NaN's or Inf's");
                break;
            }
        }
        if (h->spherical) {
            renorm_rows_kernel<<<(unsigned)k, 256, 0, st>>>(h->centroids, h->ld, d);
            h->launches++;
        }
        for (int it = 0; it < niter && !rc; ++it) {
            double obj = 0;
            if ((rc = wb_kmeans_assign_fast_dev(h, n, x, assign, &obj, st))) break;
            if ((rc = wb_kmeans_accumulate_dev(h, n, x, assign, sums, counts, st))) break;
            int64_t nsplit = 0;
            if ((rc = wb_kmeans_update_dev(h, sums, counts, n, 1234, &nsplit, st))) break;
        }
    } while (0);
    cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(x);
    cudaFree(assign);
    cudaFree(sums);
    cudaFree(counts);
    cudaFree(nonfinite);
    if (rc) {
        if (g_err.empty()) fail("k-means training failed: %s", cudaGetErrorString(cudaGetLastError()));
        return 1;
    }
    if (e != cudaSuccess) return fail("k-means training failed: %s", cudaGetErrorString(e));
    h->trained = true;
    return 0;
}

// ---- TF32 tensor-pipe peak of this GPU, measured with the library's own MMA shape (gemm_ss.cuh) -----------------
// Denominator for the batched-search roofline in bench.py: `iters` x 4 back-to-back
// tcgen05.mma.cta_group::2.kind::tf32 (M = 256, N = 256, K = 8) per CTA pair, no loads, best of `reps` launches.
extern "C" int wb_tf32_peak(int device, int iters, int reps, double* tflops_burst_out, double* tflops_sustained_out) {
    if (iters < 1 || reps < 1 || !tflops_burst_out) return fail("bad arguments");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return fail("no CUDA device available (%s)", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail("device %d out of range", device);
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail("device %d is not sm_100", device);
    const int npairs = prop.multiProcessorCount / 2;
    const int constant_ops = env_int("WB_PEAK_CONSTANT", 0);  // 1: all-ones operands (no bit toggling: less power, higher clock)
    const size_t smem = (size_t)kF2StageBytes + 1024;
    CK(cudaFuncSetAttribute(tf32_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const double flop = (double)npairs * iters * 4.0 * 2.0 * 256.0 * 256.0 * 8.0;
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {  // burst: best single launch (the first one is the warm-up)
        CK(cudaEventRecord(e0, 0));
        tf32_peak_kernel<<<2 * npairs, 128, smem, 0>>>(iters, constant_ops);
        CK(cudaGetLastError());
        CK(cudaEventRecord(e1, 0));
        CK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (r > 0) best = std::min(best, ms);
    }
    *tflops_burst_out = flop / (best * 1e-3) / 1e12;
    if (tflops_sustained_out) {  // sustained: `reps` launches back to back, the second half timed (power cap settled)
        const int half = std::max(1, reps / 2);
        for (int r = 0; r < reps - half; ++r) tf32_peak_kernel<<<2 * npairs, 128, smem, 0>>>(iters, constant_ops);
        CK(cudaEventRecord(e0, 0));
        for (int r = 0; r < half; ++r) tf32_peak_kernel<<<2 * npairs, 128, smem, 0>>>(iters, constant_ops);
        CK(cudaGetLastError());
        CK(cudaEventRecord(e1, 0));
        CK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        *tflops_sustained_out = flop * half / (ms * 1e-3) / 1e12;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return 0;
}

// ---- multi-GPU exchange over peer memory ----------------------------------------------------------
struct wb_exchange {
    int device = 0, rank = 0, world = 1;
    size_t region_bytes = 0, flags_bytes = 0, cap_entries = 0, nq_cap = 0, total_bytes = 0;
    unsigned char* local = nullptr;
    unsigned char* peers[kExchMaxWorld] = {};
    bool opened[kExchMaxWorld] = {};
    uint32_t seq = 0;       // sequence number of the last exchange that was LAUNCHED (all ranks advance in lockstep)
    int* status = nullptr;  // pinned (device-visible) host word raised by a kernel whose wait for the peers timed out
    int64_t launches = 0;   // exchange_merge_kernel launches (the fused path needs none)
};

extern "C" int wb_exch_create(int device, int rank, int world, int64_t max_queries, int64_t max_entries,
                              wb_exchange** out) {
    if (!out) return fail("out is NULL");
    *out = nullptr;
    if (world < 1 || world > kExchMaxWorld || rank < 0 || rank >= world) return fail("bad rank/world %d/%d", rank, world);
    if (max_queries < 1 || max_entries < max_queries) return fail("bad exchange capacity");
    CK(cudaSetDevice(device));
    wb_exchange* ex = new wb_exchange();
    ex->device = device;
    ex->rank = rank;
    ex->world = world;
    ex->nq_cap = (size_t)max_queries;
    ex->cap_entries = ((size_t)max_entries + 3) & ~(size_t)3;
    ex->flags_bytes = ((size_t)max_queries * 4 + 255) & ~(size_t)255;
    ex->region_bytes = (ex->flags_bytes + ex->cap_entries * 12 + 255) & ~(size_t)255;
    ex->total_bytes = ex->region_bytes * 2 * world;
    CK(cudaMalloc(&ex->local, ex->total_bytes));
    CK(cudaMemset(ex->local, 0, ex->total_bytes));
    CK(cudaHostAlloc(&ex->status, sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable));
    *ex->status = 0;
    CK(cudaDeviceSynchronize());
    ex->peers[rank] = ex->local;
    *out = ex;
    return 0;
}

extern "C" int wb_exch_local_handle(wb_exchange* ex, void* handle64) {
    if (!ex || !handle64) return fail("NULL argument");
    CK(cudaSetDevice(ex->device));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t hnd;
    CK(cudaIpcGetMemHandle(&hnd, ex->local));
    memcpy(handle64, &hnd, 64);
    return 0;
}

extern "C" int wb_exch_open_peers(wb_exchange* ex, const void* handles /* [world][64] */) {
    if (!ex || !handles) return fail("NULL argument");
    CK(cudaSetDevice(ex->device));
    for (int r = 0; r < ex->world; ++r) {
        if (r == ex->rank) continue;
        cudaIpcMemHandle_t hnd;
        memcpy(&hnd, (const char*)handles + (size_t)r * 64, 64);
        void* ptr = nullptr;
        CK(cudaIpcOpenMemHandle(&ptr, hnd, cudaIpcMemLazyEnablePeerAccess));
        ex->peers[r] = (unsigned char*)ptr;
        ex->opened[r] = true;
    }
    return 0;
}

extern "C" int wb_exch_free(wb_exchange* ex) {
    if (!ex) return 0;
    cudaSetDevice(ex->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < ex->world; ++r)
        if (ex->opened[r]) cudaIpcCloseMemHandle(ex->peers[r]);
    cudaFree(ex->local);
    cudaFreeHost(ex->status);
    delete ex;
    return 0;
}

// Parameters of the NEXT exchange (sequence number ex->seq + 1).  The caller commits the sequence number with
// exch_commit() only after the kernel that carries these parameters has been launched successfully: a rank whose
// launch failed must not drift out of lockstep with its peers.
static int exch_prepare(wb_exchange* ex, int64_t nq, int64_t k, float* D_dev, int64_t* I_dev, ExchParams* out) {
    if (!ex) return fail("NULL exchange");
    if (k < 1 || k > WB_MAX_K) return fail("k=%lld out of range [1, %d]", (long long)k, WB_MAX_K);
    if (nq < 1 || (size_t)nq > ex->nq_cap || (size_t)(nq * k) > ex->cap_entries)
        return fail("exchange capacity exceeded (nq=%lld, k=%lld)", (long long)nq, (long long)k);
    for (int r = 0; r < ex->world; ++r)
        if (!ex->peers[r]) return fail("peer %d is not mapped: call wb_exch_open_peers first", r);
    ExchParams p{};
    p.rank = ex->rank;
    p.world = ex->world;
    p.nq = nq;
    p.k = (int)k;
    p.S = merge_buffer_entries((int)k, ex->world);
    p.seq = ex->seq + 1;
    if (p.seq == 0) p.seq = 1;
    for (int r = 0; r < ex->world; ++r) p.mailbox[r] = ex->peers[r];
    p.region_bytes = ex->region_bytes;
    p.flags_bytes = ex->flags_bytes;
    p.cap_entries = ex->cap_entries;
    p.D = D_dev;
    p.I = I_dev;
    p.status = ex->status;
    *out = p;
    return 0;
}
static void exch_commit(wb_exchange* ex, const ExchParams& p) { ex->seq = p.seq; }

static int launch_exchange_kernel(wb_exchange* ex, ExchParams& p, const float* D_local, const int64_t* I_local,
                                  cudaStream_t st) {
    p.D_local = D_local;
    p.I_local = I_local;
    p.S = merge_buffer_entries(p.k, p.world);
    static thread_local bool attr_done[64] = {};
    if (ex->device >= 64 || !attr_done[ex->device]) {
        CK(cudaFuncSetAttribute(exchange_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 8));
        if (ex->device < 64) attr_done[ex->device] = true;
    }
    int sms = 148;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ex->device));
    const unsigned grid = (unsigned)std::min<int64_t>(p.nq, sms);  // one resident CTA per SM at most (1024 threads)
    exchange_merge_kernel<<<grid, kMergeThreads, (size_t)p.S * 8, st>>>(p);
    CK(cudaGetLastError());
    exch_commit(ex, p);
    ex->launches++;
    return 0;
}

// Global top-k from this rank's local (D, I): one kernel does the NVLink exchange and the merge.
// Every rank must call this the same number of times with the same nq and k.
extern "C" int wb_exch_merge_dev(wb_exchange* ex, int64_t nq, int64_t k, const float* D_local_dev,
                                 const int64_t* I_local_dev, float* D_dev, int64_t* I_dev, void* stream) {
    ExchParams p;
    TRY(exch_prepare(ex, nq, k, D_dev, I_dev, &p));
    CK(cudaSetDevice(ex->device));
    return launch_exchange_kernel(ex, p, D_local_dev, I_local_dev, (cudaStream_t)stream);
}

// 1 when a kernel of this exchange gave up waiting for a peer (a rank that never launched): the results of that
// search are invalid.  The word lives in pinned host memory that the kernels write through; it is current for every
// kernel the caller has synchronised with.
extern "C" int64_t wb_exch_launch_count(const wb_exchange* ex) { return ex ? ex->launches : -1; }

extern "C" int wb_exch_status(wb_exchange* ex, int* timed_out) {
    if (!ex || !timed_out) return fail("NULL argument");
    *timed_out = *reinterpret_cast<volatile int*>(ex->status);
    return 0;
}

// ---- FeatureStore fast ingest (host only) ------------------------------------------------------------
// Count the feature rows of one WebdatasetStore shard and report their dimension.
extern "C" int wb_tar_scan(const char* path, int64_t* rows_out, int64_t* members_out, int64_t* d_out) {
    if (!path) return fail("NULL path");
    int64_t rows = 0, members = 0, d = -1;
    std::string why;
    {   // fixed-stride shards (what WebdatasetStore writes): the layout is derived from the first member and checked
        // on up to 64 evenly spaced members - a handful of pages per shard, like the reference, which counts members
        // without reading payloads (webdataset_store.py:83-91; its count is an estimate too).  wb_tar_read verifies
        // EVERY member when it decodes and reports the exact number of rows.
        wbtar::Mapped mp;
        wbtar::Plan pl;
        bool ok = mp.open(path, false) && wbtar::make_plan(mp, &pl);
        if (ok) {  // up to 64 evenly spaced members (all of them in a small shard)
            const int64_t S = std::min<int64_t>(pl.count, 64);
            for (int64_t i = 0; i < S && ok; ++i) {
                int64_t id = 0;
                ok = wbtar::check_member(mp, pl, S > 1 ? i * (pl.count - 1) / (S - 1) : 0, &id);
            }
        }
        if (ok) {
            if (rows_out) *rows_out = pl.count * pl.m;
            if (members_out) *members_out = pl.count;
            if (d_out) *d_out = pl.d;
            return 0;
        }
    }
    const int rc = wbtar::walk(path, [&](const wbtar::Sample& s) {
        if (d < 0) d = s.d;
        if (s.d != d) return 2;
        rows += s.m;
        members++;
        return 0;
    }, &why);
    if (rc) return fail("wb_tar_scan(%s): %s (code %d)", path, why.empty() ? "mixed dimensions" : why.c_str(), rc), rc + 1;
    if (rows_out) *rows_out = rows;
    if (members_out) *members_out = members;
    if (d_out) *d_out = d;
    return 0;
}

// Decode one shard into caller-owned host buffers: ids[cap] (the sample key, repeated for multi-row samples),
// x[cap*d] float32.  *rows_out = rows written.  Fails (non-zero) if the shard needs the python reader.
extern "C" int wb_tar_read(const char* path, int64_t d, int64_t cap, int64_t* ids, float* x, int64_t* rows_out) {
    if (!path || !ids || !x || !rows_out) return fail("NULL argument");
    int64_t n = 0;
    std::string why;
    {
        wbtar::Mapped mp;
        wbtar::Plan pl;
        if (mp.open(path) && wbtar::make_plan(mp, &pl) && pl.d == d && pl.count * pl.m <= cap) {
            const size_t row_bytes = (size_t)pl.m * (size_t)d * 4;
            const bool ok = wbtar::for_each_member(mp, pl, [&](int64_t j, int64_t id) {
                memcpy(reinterpret_cast<unsigned char*>(x) + (size_t)j * row_bytes,
                       mp.p + (size_t)j * pl.stride + pl.pre + 512 + pl.data_off, row_bytes);
                for (int64_t r = 0; r < pl.m; ++r) ids[j * pl.m + r] = id;
            });
            if (ok) {
                *rows_out = pl.count * pl.m;
                return 0;
            }
        }
    }
    const int rc = wbtar::walk(path, [&](const wbtar::Sample& s) {
        if (s.d != d) { why = "dimension changes inside the shard"; return 2; }
        if (n + s.m > cap) { why = "buffer too small"; return 1; }
        memcpy(x + (size_t)n * d, s.data, (size_t)s.m * d * 4);
        for (int64_t r = 0; r < s.m; ++r) ids[n + r] = s.id;
        n += s.m;
        return 0;
    }, &why);
    if (rc) return fail("wb_tar_read(%s): %s (code %d)", path, why.c_str(), rc), rc + 1;
    *rows_out = n;
    return 0;
}

// Search over a row-sharded index, device buffers: local search + NVLink exchange + merge.  Batches that run the
// scan kernel (flat batch <= 4, every IVF list scan) do all of it in ONE launch (scan.cuh, fused tail); the others
// run the local search and then exchange_merge_kernel.  Every rank calls it with the same queries.
static int exch_search_dev_impl(wb_index* h, wb_exchange* ex, int64_t nq, const float* q_ld, int64_t k, int64_t nprobe,
                                float* D_dev, int64_t* I_dev, cudaStream_t st) {
    if (ex->device != h->device) return fail("index and exchange live on different devices");
    ExchParams p;
    TRY(exch_prepare(ex, nq, k, D_dev, I_dev, &p));
    TRY(h->dbuf.ensure((size_t)nq * k * sizeof(float)));
    TRY(h->ibuf.ensure((size_t)nq * k * sizeof(int64_t)));
    bool exchanged = false;
    TRY(search_dev_impl(h, nq, q_ld, k, nprobe, h->dbuf.as<float>(), h->ibuf.as<int64_t>(), st, &p, &exchanged));
    if (exchanged) {
        exch_commit(ex, p);
        return 0;
    }
    return launch_exchange_kernel(ex, p, h->dbuf.as<float>(), h->ibuf.as<int64_t>(), st);
}

extern "C" int wb_exch_search_dev(wb_index* h, wb_exchange* ex, int64_t nq, const float* q_dev, int64_t k, int64_t nprobe,
                                  float* D_dev, int64_t* I_dev, void* stream) {
    TRY(check_search_args(h, nq, q_dev, k, D_dev, I_dev));
    if (!ex) return fail("NULL exchange");
    if (nq == 0) return 0;
    TRY(set_dev(h));
    cudaStream_t st = (cudaStream_t)stream;
    const float* q = nullptr;
    TRY(stage_queries(h, nq, q_dev, false, st, &q));
    return exch_search_dev_impl(h, ex, nq, q, k, nprobe, D_dev, I_dev, st);
}

// Host-buffer search over a sharded index: H2D of the queries, local search, NVLink exchange + merge, D2H of the
// global (D, I) - one call, one stream, one synchronisation.  Every rank calls it with the same queries.
extern "C" int wb_exch_search(wb_index* h, wb_exchange* ex, int64_t nq, const float* q_host, int64_t k, int64_t nprobe,
                              float* D_host, int64_t* I_host) {
    TRY(check_search_args(h, nq, q_host, k, D_host, I_host));
    if (!ex) return fail("NULL exchange");
    if (nq == 0) return 0;
    TRY(set_dev(h));
    cudaStream_t st = h->stream;
    const float* q = nullptr;
    TRY(stage_queries(h, nq, q_host, true, st, &q));
    if (direct_results(h, nq, k)) {
        float* Dp = nullptr;
        int64_t* Ip = nullptr;
        TRY(direct_result_buffers(h, nq, k, &Dp, &Ip));
        TRY(exch_search_dev_impl(h, ex, nq, q, k, nprobe, Dp, Ip, st));
        TRY(finish_direct_results(nq, k, Dp, Ip, D_host, I_host, st));
    } else {
        TRY(h->eD.ensure((size_t)nq * k * sizeof(float)));
        TRY(h->eI.ensure((size_t)nq * k * sizeof(int64_t)));
        TRY(exch_search_dev_impl(h, ex, nq, q, k, nprobe, h->eD.as<float>(), h->eI.as<int64_t>(), st));
        TRY(fetch_results(h, nq, k, h->eD.as<float>(), h->eI.as<int64_t>(), D_host, I_host, st));
    }
    if (*reinterpret_cast<volatile int*>(ex->status)) return fail("sharded search: a peer GPU did not join the exchange within 20 s (results invalid)");
    return 0;
}
