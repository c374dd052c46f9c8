/* TEST INFRASTRUCTURE ONLY -- exact fp64 restatement of the reference's retrieval idiom.
 *
 * Reference: torch.cdist(Q, G) -> torch.topk(k, largest=False)
 *   (/root/reference/src/trains/trainers/complementary_item_retrieval_trainer.py:240-242,
 *    src/demo/app.py:189-190), restated as argmax of  q.g - 0.5*|g|^2  (metric 1, same order
 *   as ascending L2 distance) or q.g (metric 0), ranked by the lexicographic key
 *   (-score, index) so ties resolve to the lowest index (torch.topk leaves ties unspecified).
 *
 * Every score is one sequential fp64 accumulation over the embedding dimension, so the value
 * for a gallery row depends only on that row's bytes, never on its position or on chunking
 * (BLAS fp64 GEMMs do not have this property); duplicated rows therefore tie exactly.
 *
 * Parity status: the reference has no golden vectors for this path; this file is pinned to
 * the reference idiom's own output in tests/golden/search_pool3000.npz (tests/test_oracle.py).
 *
 * Build: gcc -O2 -pthread -shared -fPIC  (oracle/Makefile).
 */
#include <stdint.h>
#include <stdlib.h>
#include <pthread.h>

static int better(double s, int64_t i, double s2, int64_t i2) {
    return s > s2 || (s == s2 && i < i2);
}

/* scores (nq, n) fp64 */
void ofx_oracle_scores(const float* q, int64_t nq, const float* g, int64_t n, int64_t dim,
                       int metric, double* out) {
    for (int64_t j = 0; j < n; ++j) {
        const float* gr = g + j * dim;
        double nn = 0.0;
        if (metric == 1)
            for (int64_t d = 0; d < dim; ++d) nn += (double)gr[d] * (double)gr[d];
        for (int64_t i = 0; i < nq; ++i) {
            const float* qr = q + i * dim;
            double acc = 0.0;
            for (int64_t d = 0; d < dim; ++d) acc += (double)qr[d] * (double)gr[d];
            out[i * n + j] = acc - 0.5 * nn;
        }
    }
}

/* exact top-k of each query over rows [0,n) carrying global ids id_offset + j */
typedef struct {
    const float* q; const float* g; const double* half;
    int64_t nq, n, dim, k, id_offset, q_lo, q_hi; int64_t* out_idx; double* out_score;
} job_t;

static void* search_rows(void* arg) {
    job_t* J = (job_t*)arg;
    int64_t k = J->k, dim = J->dim;
    for (int64_t i = J->q_lo; i < J->q_hi; ++i) {
        const float* qr = J->q + i * dim;
        int64_t* bi = J->out_idx + i * k;
        double* bs = J->out_score + i * k;
        int64_t have = 0;
        for (int64_t j = 0; j < J->n; ++j) {
            const float* gr = J->g + j * dim;
            double acc = 0.0;
            for (int64_t d = 0; d < dim; ++d) acc += (double)qr[d] * (double)gr[d];
            acc -= J->half[j];
            int64_t id = J->id_offset + j;
            if (have == k && !better(acc, id, bs[k - 1], bi[k - 1])) continue;
            int64_t p = have < k ? have++ : k - 1;
            while (p > 0 && better(acc, id, bs[p - 1], bi[p - 1])) {
                bs[p] = bs[p - 1]; bi[p] = bi[p - 1]; --p;
            }
            bs[p] = acc; bi[p] = id;
        }
        for (int64_t p = have; p < k; ++p) { bs[p] = -1.0 / 0.0; bi[p] = -1; }
    }
    return 0;
}

void ofx_oracle_search(const float* q, int64_t nq, const float* g, int64_t n, int64_t dim,
                       int metric, int64_t k, int64_t id_offset, int threads,
                       int64_t* out_idx, double* out_score) {
    double* half = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    for (int64_t j = 0; j < n; ++j) {
        double nn = 0.0;
        if (metric == 1)
            for (int64_t d = 0; d < dim; ++d) nn += (double)g[j * dim + d] * (double)g[j * dim + d];
        half[j] = 0.5 * nn;
    }
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    if (threads > nq) threads = nq > 0 ? (int)nq : 1;
    pthread_t tid[256];
    job_t jobs[256];
    for (int t = 0; t < threads; ++t) {
        job_t J = {q, g, half, nq, n, dim, k, id_offset, nq * t / threads, nq * (t + 1) / threads,
                   out_idx, out_score};
        jobs[t] = J;
        pthread_create(&tid[t], 0, search_rows, &jobs[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(tid[t], 0);
    free(half);
}
