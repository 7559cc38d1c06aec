/* TEST INFRASTRUCTURE ONLY - C restatement of faiss's CPU IndexFlatIP / IVFFlat search.
 *
 * PARITY UNPINNED: faiss (faiss-gpu==1.7.2, /root/reference/torch-faiss-requirements.txt:6)
 * is a third-party dependency that is neither vendored under /root/reference nor installable
 * here; this file restates its published algorithm so that (1) the numpy oracle has an
 * independent second implementation to agree with and (2) bench.py has a "faiss-equivalent"
 * CPU baseline to time on the GPU box's host cores.  Never linked into the product.
 *
 * Restated routines [faiss-upstream]:
 *   fvec_inner_product            (utils/distances_simd.cpp)    -> ip_f32()
 *   exhaustive_inner_product_seq  (utils/distances.cpp)         -> orc_flat_search_seq()
 *       "#pragma omp parallel for" over QUERIES only; per query a sequential scan of all
 *       rows with heap_replace_top on a k-min-heap, then heap_reorder (descending).
 *       => one query == one thread, which is what the reference pays at n=1
 *       (call sites /root/reference/src/index/feature_search_index.py:113,
 *        /root/reference/api/routes.py:1407).
 *   IVFFlatScanner::scan_codes    (IndexIVFFlat.cpp)            -> orc_ivf_search()
 * plus orc_flat_search_mt(): the database range split over all threads (not what faiss does
 * for n < 20, but the fair "all host cores" CPU number BASELINE.md section 3 asks for).
 *
 * Tie rule (ours): score desc, then insertion position asc.
 */
#include <float.h>
#include <immintrin.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- inner product: AVX-512 / AVX2 / scalar, picked at run time ------------------------ */
__attribute__((target("avx512f"))) static float ip_avx512(const float* x, const float* y, size_t d) {
    __m512 a0 = _mm512_setzero_ps(), a1 = _mm512_setzero_ps(), a2 = _mm512_setzero_ps(), a3 = _mm512_setzero_ps();
    size_t i = 0;
    for (; i + 64 <= d; i += 64) {
        a0 = _mm512_fmadd_ps(_mm512_loadu_ps(x + i), _mm512_loadu_ps(y + i), a0);
        a1 = _mm512_fmadd_ps(_mm512_loadu_ps(x + i + 16), _mm512_loadu_ps(y + i + 16), a1);
        a2 = _mm512_fmadd_ps(_mm512_loadu_ps(x + i + 32), _mm512_loadu_ps(y + i + 32), a2);
        a3 = _mm512_fmadd_ps(_mm512_loadu_ps(x + i + 48), _mm512_loadu_ps(y + i + 48), a3);
    }
    for (; i + 16 <= d; i += 16) a0 = _mm512_fmadd_ps(_mm512_loadu_ps(x + i), _mm512_loadu_ps(y + i), a0);
    float s = _mm512_reduce_add_ps(_mm512_add_ps(_mm512_add_ps(a0, a1), _mm512_add_ps(a2, a3)));
    for (; i < d; i++) s += x[i] * y[i];
    return s;
}

__attribute__((target("avx2,fma"))) static float ip_avx2(const float* x, const float* y, size_t d) {
    __m256 a0 = _mm256_setzero_ps(), a1 = _mm256_setzero_ps(), a2 = _mm256_setzero_ps(), a3 = _mm256_setzero_ps();
    size_t i = 0;
    for (; i + 32 <= d; i += 32) {
        a0 = _mm256_fmadd_ps(_mm256_loadu_ps(x + i), _mm256_loadu_ps(y + i), a0);
        a1 = _mm256_fmadd_ps(_mm256_loadu_ps(x + i + 8), _mm256_loadu_ps(y + i + 8), a1);
        a2 = _mm256_fmadd_ps(_mm256_loadu_ps(x + i + 16), _mm256_loadu_ps(y + i + 16), a2);
        a3 = _mm256_fmadd_ps(_mm256_loadu_ps(x + i + 24), _mm256_loadu_ps(y + i + 24), a3);
    }
    for (; i + 8 <= d; i += 8) a0 = _mm256_fmadd_ps(_mm256_loadu_ps(x + i), _mm256_loadu_ps(y + i), a0);
    __m256 t = _mm256_add_ps(_mm256_add_ps(a0, a1), _mm256_add_ps(a2, a3));
    __m128 lo = _mm_add_ps(_mm256_castps256_ps128(t), _mm256_extractf128_ps(t, 1));
    lo = _mm_hadd_ps(lo, lo);
    lo = _mm_hadd_ps(lo, lo);
    float s = _mm_cvtss_f32(lo);
    for (; i < d; i++) s += x[i] * y[i];
    return s;
}

static float ip_scalar(const float* x, const float* y, size_t d) {
    float s = 0.f;
    for (size_t i = 0; i < d; i++) s += x[i] * y[i];
    return s;
}

typedef float (*ip_fn)(const float*, const float*, size_t);
static ip_fn pick_ip(void) {
    __builtin_cpu_init();
    if (__builtin_cpu_supports("avx512f")) return ip_avx512;
    if (__builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma")) return ip_avx2;
    return ip_scalar;
}

const char* orc_simd_name(void) {
    ip_fn f = pick_ip();
    return f == ip_avx512 ? "avx512f" : f == ip_avx2 ? "avx2+fma" : "scalar";
}

/* Thread count for the *_mt / omp-parallel entry points.  torch.distributed.run exports OMP_NUM_THREADS=1 to its
 * workers; the benchmark's reference arm calls this with the size of its CPU affinity set instead. */
void orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---- k-min-heap keyed by (score asc, position desc): the root is the WORST kept entry -- */
typedef struct { float s; int64_t p; } ent_t;
/* a is worse than b  <=>  lower score, or equal score and later position */
static inline int worse(ent_t a, ent_t b) { return a.s < b.s || (a.s == b.s && a.p > b.p); }

static inline void heap_sift_down(ent_t* h, size_t k, size_t i) {
    ent_t v = h[i];
    for (;;) {
        size_t l = 2 * i + 1, r = l + 1, m;
        if (l >= k) break;
        m = (r < k && worse(h[r], h[l])) ? r : l;
        if (!worse(h[m], v)) break;
        h[i] = h[m];
        i = m;
    }
    h[i] = v;
}

static inline void heap_init(ent_t* h, size_t k) {
    for (size_t i = 0; i < k; i++) { h[i].s = -FLT_MAX; h[i].p = INT64_MAX; }
}

/* faiss heap_replace_top: called when the new entry beats the root */
static inline void heap_offer(ent_t* h, size_t k, float s, int64_t p) {
    ent_t e = { s, p };
    if (worse(h[0], e)) { h[0] = e; heap_sift_down(h, k, 0); }
}

static int cmp_best_first(const void* a, const void* b) {
    const ent_t* x = (const ent_t*)a; const ent_t* y = (const ent_t*)b;
    if (x->s != y->s) return x->s > y->s ? -1 : 1;
    return x->p < y->p ? -1 : (x->p > y->p ? 1 : 0);
}

/* faiss heap_reorder + IndexIDMap id translation */
static void heap_emit(ent_t* h, size_t k, const int64_t* ids, float* D, int64_t* I) {
    qsort(h, k, sizeof(ent_t), cmp_best_first);
    for (size_t j = 0; j < k; j++) {
        if (h[j].p == INT64_MAX) { D[j] = -FLT_MAX; I[j] = -1; }
        else { D[j] = h[j].s; I[j] = ids ? ids[h[j].p] : h[j].p; }
    }
}

/* ---- IndexFlatIP.search, faiss threading (parallel over queries only) ------------------- */
int orc_flat_search_seq(const float* xb, int64_t N, int64_t d, const int64_t* ids,
                        const float* xq, int64_t n, int64_t k, float* D, int64_t* I) {
    ip_fn ip = pick_ip();
    if (k <= 0) return 0;
#pragma omp parallel for schedule(static)
    for (int64_t q = 0; q < n; q++) {
        ent_t* h = (ent_t*)malloc(sizeof(ent_t) * (size_t)k);
        heap_init(h, (size_t)k);
        const float* x = xq + q * d;
        for (int64_t j = 0; j < N; j++) heap_offer(h, (size_t)k, ip(x, xb + j * d, (size_t)d), j);
        heap_emit(h, (size_t)k, ids, D + q * k, I + q * k);
        free(h);
    }
    return 0;
}

/* ---- same result, database range split over all threads (fair all-core CPU number) ------ */
int orc_flat_search_mt(const float* xb, int64_t N, int64_t d, const int64_t* ids,
                       const float* xq, int64_t n, int64_t k, float* D, int64_t* I) {
    ip_fn ip = pick_ip();
    if (k <= 0) return 0;
    int T = orc_max_threads();
    ent_t* parts = (ent_t*)malloc(sizeof(ent_t) * (size_t)k * (size_t)T * (size_t)n);
#pragma omp parallel num_threads(T)
    {
#ifdef _OPENMP
        int t = omp_get_thread_num();
#else
        int t = 0;
#endif
        int64_t lo = N * t / T, hi = N * (t + 1) / T;
        for (int64_t q = 0; q < n; q++) heap_init(parts + ((size_t)q * T + t) * k, (size_t)k);
        /* row-major over the slice so each database row is read once for all n queries */
        for (int64_t j = lo; j < hi; j++) {
            const float* y = xb + j * d;
            for (int64_t q = 0; q < n; q++)
                heap_offer(parts + ((size_t)q * T + t) * k, (size_t)k, ip(xq + q * d, y, (size_t)d), j);
        }
    }
    for (int64_t q = 0; q < n; q++) {
        ent_t* all = parts + (size_t)q * T * k;
        qsort(all, (size_t)k * T, sizeof(ent_t), cmp_best_first);
        heap_emit(all, (size_t)k, ids, D + q * k, I + q * k); /* first k of the sorted union */
    }
    free(parts);
    return 0;
}

/* ---- IndexIVFFlat.search over CSR lists (perm = rows of each list in insertion order) --- */
int orc_ivf_search(const float* xb, int64_t d, const int64_t* ids, const int64_t* list_off,
                   const int64_t* perm, const float* centroids, int64_t nlist,
                   const float* xq, int64_t n, int64_t k, int64_t nprobe, float* D, int64_t* I) {
    ip_fn ip = pick_ip();
    if (nprobe > nlist) nprobe = nlist;
    if (k <= 0) return 0;
#pragma omp parallel for schedule(dynamic)
    for (int64_t q = 0; q < n; q++) {
        const float* x = xq + q * d;
        ent_t* ch = (ent_t*)malloc(sizeof(ent_t) * (size_t)nprobe);
        heap_init(ch, (size_t)nprobe);
        for (int64_t c = 0; c < nlist; c++) heap_offer(ch, (size_t)nprobe, ip(x, centroids + c * d, (size_t)d), c);
        qsort(ch, (size_t)nprobe, sizeof(ent_t), cmp_best_first);
        ent_t* h = (ent_t*)malloc(sizeof(ent_t) * (size_t)k);
        heap_init(h, (size_t)k);
        for (int64_t pi = 0; pi < nprobe; pi++) {
            if (ch[pi].p == INT64_MAX) continue;
            int64_t l = ch[pi].p;
            for (int64_t s = list_off[l]; s < list_off[l + 1]; s++) {
                int64_t r = perm[s];
                heap_offer(h, (size_t)k, ip(x, xb + r * d, (size_t)d), r);
            }
        }
        heap_emit(h, (size_t)k, ids, D + q * k, I + q * k);
        free(h); free(ch);
    }
    return 0;
}
