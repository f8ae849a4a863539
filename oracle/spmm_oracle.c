/*
 * spmm_oracle.c — CPU restatement of the GraphBLAS aggregation of the reference's CPU trainer.
 * TEST INFRASTRUCTURE / CPU BASELINE ONLY: never linked into or called by the product path.
 *
 * Follows Parallel-GCN/main.c:
 *   :269-272  AH = A * H            GrB_mxm(AH, NULL, NULL, GxB_PLUS_TIMES_FP32, A, H, GrB_DESC_R)
 *   :292-296  AH += A * Hcap_p      GrB_mxm(AH, NULL, GrB_PLUS_FP32, GxB_PLUS_TIMES_FP32, A, Hcap, NULL)
 *             (one accumulate per peer whose halo rows arrived; same pattern for the gradient at
 *              :374-404 with A applied to G)
 * A is the rank's block of rows in CSR (the reference reads "i j val" triples, main.c:609-647),
 * H is dense row-major fp32 (the reference stores it as a GraphBLAS matrix whose owned rows are
 * fully populated, main.c:650-684, so PLUS_TIMES over it is exactly CSR x dense).
 * The semiring is fp32 PLUS_TIMES: products and sums in float, one row of A at a time, in CSR order.
 *
 * SuiteSparse:GraphBLAS itself is not vendored and cannot be built here: parity at
 * this boundary is unpinned by the reference; this file is cross-checked against the PGCN.py oracle
 * (tests/test_oracle_golden.py) and, on 1 rank, the two paths compute the same product.
 *
 * Build: gcc -O3 -march=native -fopenmp -shared -fPIC spmm_oracle.c -o liboracle_spmm.so
 */
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* AH (accumulate ? += : =) A[:, col_lo:col_hi) * H     — col range selects "own" vs one peer's halo
 * columns so the per-peer accumulate structure of main.c:275-299 can be replayed literally.
 * rowptr/colidx/vals: CSR with m rows; H: ncols x f; AH: m x f. */
void grb_mxm_plus_times_fp32(int64_t m, const int32_t* rowptr, const int32_t* colidx, const float* vals,
                             const float* H, int64_t f, float* AH,
                             int32_t col_lo, int32_t col_hi, int accumulate)
{
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < m; ++i) {
        float* out = AH + i * f;
        if (!accumulate) memset(out, 0, (size_t)f * sizeof(float));
        for (int32_t e = rowptr[i]; e < rowptr[i + 1]; ++e) {
            const int32_t j = colidx[e];
            if (j < col_lo || j >= col_hi) continue;
            const float a = vals[e];
            const float* h = H + (int64_t)j * f;
            for (int64_t c = 0; c < f; ++c) out[c] += a * h[c];
        }
    }
}

/* The whole aggregation of one layer on one rank: local part, then one accumulate per peer block.
 * peer_off[k+1] are column offsets of the halo groups (columns >= m). */
void grb_aggregate(int64_t m, const int32_t* rowptr, const int32_t* colidx, const float* vals,
                   const float* Hcat, int64_t f, float* AH, int32_t k, const int64_t* peer_off)
{
    grb_mxm_plus_times_fp32(m, rowptr, colidx, vals, Hcat, f, AH, 0, (int32_t)m, 0);      /* :271 */
    for (int32_t p = 0; p < k; ++p) {                                                       /* :275 */
        const int32_t lo = (int32_t)(m + peer_off[p]), hi = (int32_t)(m + peer_off[p + 1]);
        if (hi > lo) grb_mxm_plus_times_fp32(m, rowptr, colidx, vals, Hcat, f, AH, lo, hi, 1); /* :295 */
    }
}

/* Plain one-pass CSR x dense (what the two steps above add up to); used as the timed CPU baseline.
 * Written the way a competent CPU port would be: the H rows of the next few edges are software-prefetched
 * (the gather is latency-bound otherwise: the next address depends on colidx[e+1]), the row sum is kept in
 * a local accumulator and the inner loop is left to the vectoriser. Same arithmetic and order as above. */
#define ORACLE_PREFETCH_DIST 12
void spmm_csr_fp32(int64_t m, const int32_t* rowptr, const int32_t* colidx, const float* vals,
                   const float* H, int64_t f, float* Z)
{
    const int64_t nnz = rowptr[m];
#pragma omp parallel
    {
        float* acc = (float*)__builtin_alloca((size_t)f * sizeof(float));
#pragma omp for schedule(dynamic, 256)
        for (int64_t i = 0; i < m; ++i) {
            for (int64_t c = 0; c < f; ++c) acc[c] = 0.0f;
            for (int32_t e = rowptr[i]; e < rowptr[i + 1]; ++e) {
                if ((int64_t)e + ORACLE_PREFETCH_DIST < nnz) {
                    const char* p = (const char*)(H + (int64_t)colidx[e + ORACLE_PREFETCH_DIST] * f);
                    for (int64_t b = 0; b < f * 4; b += 64) __builtin_prefetch(p + b, 0, 0);
                }
                const float a = vals[e];
                const float* h = H + (int64_t)colidx[e] * f;
#pragma omp simd
                for (int64_t c = 0; c < f; ++c) acc[c] += a * h[c];
            }
            float* out = Z + i * f;
            for (int64_t c = 0; c < f; ++c) out[c] = acc[c];
        }
    }
}

void oracle_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
