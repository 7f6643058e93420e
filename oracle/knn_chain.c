/*
 * CPU ORACLE (test infrastructure, not product code): strict brute-force KNN over edge
 * midpoints in the arithmetic of the reference's torch.cdist call
 * (graphem_rapids/backends/embedder_pytorch.py:580 `torch.cdist(query_chunk, reference_points, p=2)`
 *  followed by :583 `torch.topk(distances, k, dim=1, largest=False)`).
 *
 * torch (third party, unpinned `torch>=2.0.0`; 2.11.0 in this image) computes, when either
 * side has more than 25 rows, D = sqrt(max(0, [-2x, |x|^2, 1] . [y, 1, |y|^2]^T)) through an
 * SGEMM with K = d+2, which on CPU is bit-equal to the sequential FMA chain below [probed in
 * this container on 256 x 200 000 pairs, d = 2 and 3].  With <= 25 rows on both sides it
 * uses the direct form sqrt(fma(d2,d2,fma(d1,d1,d0*d0))) [probed].
 *
 * The order is the north star's: ascending (distance, index), distance = correctly rounded
 * sqrtf of the clamped chain value.  torch.topk's own tie order is unspecified.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC knn_chain.c -o _build/liboracle_knn.so -lm
 * (-ffp-contract=off: the squared norms must NOT be fused; the chain uses explicit fmaf.)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static inline float sq_norm(const float *x, int d) {
    /* x.pow(2).sum(-1): non-fused, left to right (bit-equal to torch for d <= 3) */
    float s = x[0] * x[0];
    for (int j = 1; j < d; ++j) s = s + x[j] * x[j];
    return s;
}

static inline float chain_mm(const float *q, float qn, const float *y, float yn, int d) {
    float acc = 0.0f;
    for (int j = 0; j < d; ++j) acc = fmaf(-2.0f * q[j], y[j], acc);
    acc = fmaf(qn, 1.0f, acc);
    acc = fmaf(1.0f, yn, acc);
    return acc < 0.0f ? 0.0f : acc;
}

static inline float chain_direct(const float *q, const float *y, int d) {
    float t = q[0] - y[0];
    float acc = t * t;
    for (int j = 1; j < d; ++j) { t = q[j] - y[j]; acc = fmaf(t, t, acc); }
    return acc;
}

/* mid: (E,d) fp32 row-major; samp: (S) int64 edge ids; out rows sorted ascending by (dist, idx) */
int oracle_knn_strict(const float *mid, int64_t E, int d, const int64_t *samp, int64_t S,
                      int kp1, int mm_mode, int64_t *out_idx, float *out_dist) {
    if (kp1 > E || kp1 <= 0) return -1;
    float *yn = (float *)malloc(sizeof(float) * (size_t)E);
    if (!yn) return -2;
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < E; ++e) yn[e] = sq_norm(mid + e * d, d);

#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t s = 0; s < S; ++s) {
        const float *q = mid + samp[s] * d;
        const float qn = yn[samp[s]];
        int64_t *bi = out_idx + s * kp1;
        float *bd = out_dist + s * kp1;
        int n = 0;
        for (int64_t e = 0; e < E; ++e) {
            float v = mm_mode ? chain_mm(q, qn, mid + e * d, yn[e], d) : chain_direct(q, mid + e * d, d);
            float dist = sqrtf(v);
            /* candidates arrive in ascending index order, so a tie never displaces */
            if (n == kp1 && !(dist < bd[n - 1])) continue;
            int p = (n < kp1) ? n : kp1 - 1;
            while (p > 0 && dist < bd[p - 1]) { bd[p] = bd[p - 1]; bi[p] = bi[p - 1]; --p; }
            bd[p] = dist; bi[p] = e;
            if (n < kp1) ++n;
        }
    }
    free(yn);
    return 0;
}
