// Per-category candidate-pool retrieval: the reference never searches one flat
// gallery -- `compute_recall_metrics` groups the queries by the target item's category and ranks
// each against that category's pool of <= 3000 items with torch.cdist -> torch.topk(50,
// largest=False) (/root/reference/src/trains/trainers/complementary_item_retrieval_trainer.py:
// 192-249; pools: src/trains/datasets/polyvore/polyvore_complementary_item_retrieval_dataset.py:
// 111-153).  Pools are tiny next to a tensor-core tile sweep, so this path is an exact brute
// force: one CTA per query scores every row of the query's pool in fp64 (q.g - 0.5|g|^2, the
// same ranking as ascending L2 distance), sorts (score desc, index asc) in shared memory and
// emits the first k pool-local indices.  HBM/L2-bound: a pool (<= 12 MB fp32) stays in L2 and
// is re-read once per query of that category.
#include "common.h"

namespace ofx {

constexpr int kPoolMax = 4096;      // rows per pool (reference: 3000)
constexpr int kPoolThreads = 256;

__global__ void __launch_bounds__(kPoolThreads)
pool_search_kernel(const float* __restrict__ pools, const long long* __restrict__ pool_off,
                   const float* __restrict__ queries, const int* __restrict__ query_pool, int dim,
                   int k, int metric, long long* __restrict__ out_idx, double* __restrict__ out_score) {
    extern __shared__ __align__(16) unsigned char sm[];
    double* s_score = reinterpret_cast<double*>(sm);                       // [n_pad]
    int* s_idx = reinterpret_cast<int*>(sm + sizeof(double) * kPoolMax);   // [n_pad]
    float* s_q = reinterpret_cast<float*>(sm + (sizeof(double) + sizeof(int)) * kPoolMax);   // [dim]
    const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = query_pool[q];
    const long long lo = pool_off[c];
    const int n = static_cast<int>(pool_off[c + 1] - lo);
    int n_pad = 2;
    while (n_pad < n) n_pad <<= 1;
    for (int d = tid; d < dim; d += kPoolThreads) s_q[d] = queries[static_cast<long long>(q) * dim + d];
    __syncthreads();
    for (int r = warp; r < n_pad; r += kPoolThreads / 32) {
        double acc = 0.0, nn = 0.0;
        if (r < n) {
            const float* g = pools + (lo + r) * dim;
            for (int d = lane * 4; d < dim; d += 128) {
                const float4 gv = *reinterpret_cast<const float4*>(g + d);
                const float4 qv = *reinterpret_cast<const float4*>(s_q + d);
                acc = fma(static_cast<double>(qv.x), static_cast<double>(gv.x), acc);
                acc = fma(static_cast<double>(qv.y), static_cast<double>(gv.y), acc);
                acc = fma(static_cast<double>(qv.z), static_cast<double>(gv.z), acc);
                acc = fma(static_cast<double>(qv.w), static_cast<double>(gv.w), acc);
                nn = fma(static_cast<double>(gv.x), static_cast<double>(gv.x), nn);
                nn = fma(static_cast<double>(gv.y), static_cast<double>(gv.y), nn);
                nn = fma(static_cast<double>(gv.z), static_cast<double>(gv.z), nn);
                nn = fma(static_cast<double>(gv.w), static_cast<double>(gv.w), nn);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                acc += __shfl_xor_sync(0xffffffffu, acc, o);
                nn += __shfl_xor_sync(0xffffffffu, nn, o);
            }
        }
        if (lane == 0) {
            s_score[r] = r < n ? (metric == OFX_METRIC_L2 ? acc - 0.5 * nn : acc) : -INFINITY;
            s_idx[r] = r < n ? r : 0x7fffffff;
        }
    }
    __syncthreads();
    // bitonic sort: (score desc, index asc); padding (-inf, INT_MAX) sinks to the end
    for (int size = 2; size <= n_pad; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int e = tid; e < n_pad / 2; e += kPoolThreads) {
                const int a = 2 * e - (e & (stride - 1));
                const int b = a + stride;
                const bool asc_block = (a & size) != 0;   // blocks alternate direction
                const double sa = s_score[a], sb = s_score[b];
                const int ia = s_idx[a], ib = s_idx[b];
                const bool a_first = sa > sb || (sa == sb && ia < ib);   // a ranks before b
                if (a_first == asc_block) {
                    s_score[a] = sb; s_score[b] = sa;
                    s_idx[a] = ib; s_idx[b] = ia;
                }
            }
            __syncthreads();
        }
    }
    for (int j = tid; j < k; j += kPoolThreads) {
        const bool real = j < n;
        out_idx[static_cast<long long>(q) * k + j] = real ? s_idx[j] : -1;
        out_score[static_cast<long long>(q) * k + j] = real ? s_score[j] : -INFINITY;
    }
}

}  // namespace ofx

using namespace ofx;

extern "C" int ofx_pool_search(const float* pools, const int64_t* pool_offsets, int32_t n_pools,
                               int32_t max_pool_rows, const float* queries, const int32_t* query_pool,
                               int32_t n_query, int32_t dim, int32_t k, int32_t metric, double* out_score,
                               int64_t* out_idx, void* stream) {
    if (n_query < 0 || n_pools < 1) return fail(OFX_E_SHAPE, "ofx_pool_search: n_query %d, n_pools %d", n_query, n_pools);
    if (k < 1 || k > 64) return fail(OFX_E_SHAPE, "ofx_pool_search: k %d not in [1,64]", k);
    if (dim <= 0 || dim % 128 || dim > 2048) return fail(OFX_E_SHAPE, "ofx_pool_search: dim %d must be a multiple of 128, <= 2048", dim);
    if (max_pool_rows < 1 || max_pool_rows > kPoolMax)
        return fail(OFX_E_SHAPE, "ofx_pool_search: pools hold 1..%d rows (got max %d)", kPoolMax, max_pool_rows);
    if (metric != OFX_METRIC_DOT && metric != OFX_METRIC_L2) return fail(OFX_E_ARG, "ofx_pool_search: metric %d", metric);
    if (n_query == 0) return OFX_OK;
    if (!pools || !pool_offsets || !queries || !query_pool || !out_score || !out_idx)
        return fail(OFX_E_ARG, "ofx_pool_search: null argument");
    if (reinterpret_cast<uintptr_t>(pools) % 16 || reinterpret_cast<uintptr_t>(queries) % 16)
        return fail(OFX_E_ARG, "ofx_pool_search: misaligned pointer");
    OFX_TRY(require_sm100());
    const size_t smem = (sizeof(double) + sizeof(int)) * kPoolMax + sizeof(float) * dim;
    static DeviceOnce configured;
    if (configured.need()) {
        OFX_CUDA(cudaFuncSetAttribute(pool_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>((sizeof(double) + sizeof(int)) * kPoolMax + sizeof(float) * 2048)));
    }
    pool_search_kernel<<<n_query, kPoolThreads, smem, static_cast<cudaStream_t>(stream)>>>(
        pools, reinterpret_cast<const long long*>(pool_offsets), queries, query_pool, dim, k, metric,
        reinterpret_cast<long long*>(out_idx), out_score);
    OFX_LAUNCH_CHECK();
    return OFX_OK;
}
