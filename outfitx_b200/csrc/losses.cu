// Evaluation-side scoring of the reference's trainers, forward only:
//
//   FocalLoss            /root/reference/src/losses/focal_loss.py:23-41
//   SetWiseRankingLoss   /root/reference/src/losses/set_wise_ranking_loss.py:14-37
//   compute_cp_metrics   /root/reference/src/trains/trainers/compatibility_prediction_trainer.py:406-436
//                        (sigmoid -> threshold 0.5 -> TP/FP/FN/accuracy, and sklearn's roc_auc_score)
//
// All three are HBM-bound reductions over data the scoring pass has just produced (logits, query
// embeddings) plus labels / negatives; they run on the caller's stream right behind it, so the
// validation loop never copies scores to the host before it has a metric.  Reductions are
// two-stage (per-CTA partials in the caller's workspace, then one CTA in a fixed order) or integer
// atomics, so results are deterministic run to run.
#include <math.h>

#include "common.h"

namespace ofx {

constexpr int kLossThreads = 256;
constexpr int kLossMaxBlocks = 1024;

__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double t = 0.0;
    if (warp == 0) {
        t = lane < (blockDim.x >> 5) ? sh[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    return t;   // valid in warp 0
}

__device__ __forceinline__ float sigmoid_f32(float x) { return 1.f / (1.f + expf(-x)); }

// ---------------------------------------------------------------- focal loss (focal_loss.py:23-41)
// ce = BCE-with-logits(x, y) ; p = sigmoid(x) ; p_t = p y + (1 - p)(1 - y)
// loss = alpha_t * ce * (1 - p_t)^gamma,  alpha_t = alpha y + (1 - alpha)(1 - y)  (when alpha >= 0)
__global__ void __launch_bounds__(kLossThreads)
focal_loss_kernel(const float* __restrict__ logits, const float* __restrict__ labels, long long n, float gamma,
                  float alpha, float* __restrict__ per_elem, double* __restrict__ partial) {
    __shared__ double sh[kLossThreads / 32];
    double acc = 0.0;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float x = logits[i], y = labels[i];
        const float ce = fmaxf(x, 0.f) - x * y + log1pf(expf(-fabsf(x)));
        const float p = sigmoid_f32(x);
        const float pt = p * y + (1.f - p) * (1.f - y);
        const float om = 1.f - pt;
        const float mod = gamma == 2.f ? om * om : (gamma == 0.f ? 1.f : (gamma == 1.f ? om : powf(om, gamma)));
        float l = ce * mod;
        if (alpha >= 0.f) l = (alpha * y + (1.f - alpha) * (1.f - y)) * l;
        if (per_elem) per_elem[i] = l;
        acc += static_cast<double>(l);
    }
    const double t = block_sum(acc, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

// out[0] = sum, out[1] = mean (sum / max(n, 1))
__global__ void __launch_bounds__(kLossThreads)
focal_finish_kernel(const double* __restrict__ partial, int n_partial, long long n, double* __restrict__ out) {
    __shared__ double sh[kLossThreads / 32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n_partial; i += blockDim.x) acc += partial[i];
    const double t = block_sum(acc, sh);
    if (threadIdx.x == 0) {
        out[0] = t;
        out[1] = n > 0 ? t / static_cast<double>(n) : nan("");   // torch: mean of an empty tensor is nan
    }
}

// ---------------------------------------------------------------- set-wise ranking loss
// (set_wise_ranking_loss.py:14-37)  one CTA per outfit b:
//   pos  = || y_hat - y + 1e-6 ||            (F.pairwise_distance, eps added to the difference)
//   neg_k = || y_hat - negatives[b, k] ||
//   hinge_sum[b] = sum_k valid relu(pos - neg_k + margin),  valid[b] = #valid,
//   hard[b] = relu(pos - min_valid_k neg_k + margin)        (min over nothing = +inf -> 0)
struct RankPartial {
    double hinge_sum;
    float hard;
    int valid;
};

__global__ void __launch_bounds__(128)
ranking_loss_kernel(const float* __restrict__ y, const float* __restrict__ y_hat, const float* __restrict__ neg,
                    const uint8_t* __restrict__ neg_mask, int n_neg, int dim, float margin,
                    RankPartial* __restrict__ partial) {
    extern __shared__ __align__(16) float s_q[];            // y_hat row
    __shared__ float s_red[4];
    __shared__ float s_hinge[4], s_min[4];
    __shared__ int s_valid[4];
    const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float* yh = y_hat + static_cast<long long>(b) * dim;
    const float* yp = y + static_cast<long long>(b) * dim;
    float pp = 0.f;
    for (int i = tid * 4; i < dim; i += 512) {
        const float4 q = *reinterpret_cast<const float4*>(yh + i);
        const float4 t = *reinterpret_cast<const float4*>(yp + i);
        *reinterpret_cast<float4*>(s_q + i) = q;
        const float d0 = q.x - t.x + 1e-6f, d1 = q.y - t.y + 1e-6f, d2 = q.z - t.z + 1e-6f, d3 = q.w - t.w + 1e-6f;
        pp += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pp += __shfl_xor_sync(0xffffffffu, pp, o);
    if (lane == 0) s_red[warp] = pp;
    __syncthreads();
    const float pos = sqrtf(s_red[0] + s_red[1] + s_red[2] + s_red[3]);

    float hinge = 0.f, mn = INFINITY;
    int valid = 0;
    for (int k = warp; k < n_neg; k += 4) {
        if (neg_mask[static_cast<long long>(b) * n_neg + k]) continue;     // True = padding
        const float* g = neg + (static_cast<long long>(b) * n_neg + k) * dim;
        float dd = 0.f;
        for (int i = lane * 4; i < dim; i += 128) {
            const float4 q = *reinterpret_cast<const float4*>(s_q + i);
            const float4 t = __ldg(reinterpret_cast<const float4*>(g + i));
            const float d0 = q.x - t.x, d1 = q.y - t.y, d2 = q.z - t.z, d3 = q.w - t.w;
            dd += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dd += __shfl_xor_sync(0xffffffffu, dd, o);
        const float nd = sqrtf(dd);
        hinge += fmaxf(pos - nd + margin, 0.f);
        mn = fminf(mn, nd);
        ++valid;
    }
    if (lane == 0) { s_hinge[warp] = hinge; s_min[warp] = mn; s_valid[warp] = valid; }
    __syncthreads();
    if (tid == 0) {
        RankPartial r;
        r.hinge_sum = static_cast<double>(s_hinge[0]) + s_hinge[1] + s_hinge[2] + s_hinge[3];
        r.valid = s_valid[0] + s_valid[1] + s_valid[2] + s_valid[3];
        const float m = fminf(fminf(s_min[0], s_min[1]), fminf(s_min[2], s_min[3]));
        r.hard = fmaxf(pos - m + margin, 0.f);
        partial[b] = r;
    }
}

// out[0] = L_all + L_hard, out[1] = L_all, out[2] = L_hard
__global__ void __launch_bounds__(kLossThreads)
ranking_finish_kernel(const RankPartial* __restrict__ partial, int batch, double* __restrict__ out) {
    __shared__ double sh[kLossThreads / 32];
    double hs = 0.0, hd = 0.0, vc = 0.0;
    for (int i = threadIdx.x; i < batch; i += blockDim.x) {
        const RankPartial r = partial[i];
        hs += r.hinge_sum; hd += static_cast<double>(r.hard); vc += static_cast<double>(r.valid);
    }
    const double a = block_sum(hs, sh);
    const double b = block_sum(hd, sh);
    const double c = block_sum(vc, sh);
    if (threadIdx.x == 0) {
        const double l_all = a / fmax(c, 1.0);        // valid_count.clamp(min=1)
        const double l_hard = b / static_cast<double>(batch);
        out[0] = l_all + l_hard; out[1] = l_all; out[2] = l_hard;
    }
}

// ---------------------------------------------------------------- CP metrics
// counts[0..5] = TP, FP, FN, correct, n_pos, n_neg  with prediction = sigmoid(logit) > 0.5 and
// label = int(label); counts[6] = 2 * #{(i,j): y_i = 1, y_j = 0, p_j < p_i} + #{...: p_j == p_i},
// i.e. AUC = counts[6] / (2 n_pos n_neg)  -- the Mann-Whitney statistic roc_auc_score computes.
__global__ void __launch_bounds__(kLossThreads)
cp_counts_kernel(const float* __restrict__ logits, const float* __restrict__ labels, long long n,
                 float* __restrict__ probs, unsigned long long* __restrict__ counts) {
    unsigned tp = 0, fp = 0, fn = 0, ok = 0, np_ = 0, nn = 0;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float p = sigmoid_f32(logits[i]);
        probs[i] = p;
        const int y = static_cast<int>(labels[i]);          // labels.int()
        const int pred = p > 0.5f ? 1 : 0;
        tp += (pred == 1) & (y == 1); fp += (pred == 1) & (y == 0); fn += (pred == 0) & (y == 1);
        ok += pred == y; np_ += y == 1; nn += y == 0;
    }
    unsigned v[6] = {tp, fp, fn, ok, np_, nn};
#pragma unroll
    for (int c = 0; c < 6; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[c] += __shfl_xor_sync(0xffffffffu, v[c], o);
        if ((threadIdx.x & 31) == 0 && v[c]) atomicAdd(&counts[c], static_cast<unsigned long long>(v[c]));
    }
}

constexpr int kAucTile = 2048;
__global__ void __launch_bounds__(kLossThreads)
cp_auc_kernel(const float* __restrict__ probs, const float* __restrict__ labels, long long n,
              unsigned long long* __restrict__ counts) {
    __shared__ float s_p[kAucTile];            // probabilities of the NEGATIVES of the j tile; +inf otherwise
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const bool pos = i < n && static_cast<int>(labels[i]) == 1;
    const float pi = i < n ? probs[i] : 0.f;
    const long long j0 = static_cast<long long>(blockIdx.y) * kAucTile;
    for (int t = threadIdx.x; t < kAucTile; t += blockDim.x) {
        const long long j = j0 + t;
        s_p[t] = (j < n && static_cast<int>(labels[j]) == 0) ? probs[j] : INFINITY;
    }
    __syncthreads();
    unsigned c = 0;
    if (pos) {
#pragma unroll 8
        for (int t = 0; t < kAucTile; ++t) {
            const float pj = s_p[t];
            c += (pj < pi ? 2u : 0u) + (pj == pi ? 1u : 0u);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&counts[6], static_cast<unsigned long long>(c));
}

static int loss_blocks(long long n) {
    long long b = (n + kLossThreads - 1) / kLossThreads;
    if (b < 1) b = 1;
    return static_cast<int>(b < kLossMaxBlocks ? b : kLossMaxBlocks);
}

}  // namespace ofx

using namespace ofx;

extern "C" size_t ofx_loss_workspace_bytes(int64_t batch) {
    const size_t a = sizeof(double) * kLossMaxBlocks;
    const size_t b = sizeof(RankPartial) * static_cast<size_t>(batch > 0 ? batch : 0);
    return align_up(a > b ? a : b, 256);
}

extern "C" int ofx_focal_loss(const float* logits, const float* labels, int64_t n, float gamma, float alpha,
                              float* per_elem, double* out, void* workspace, size_t workspace_bytes, void* stream) {
    if (n < 0) return fail(OFX_E_SHAPE, "ofx_focal_loss: n %lld", static_cast<long long>(n));
    if (!(gamma >= 0.f)) return fail(OFX_E_ARG, "ofx_focal_loss: gamma %g should be non-negative", gamma);
    if (!(alpha <= 1.f)) return fail(OFX_E_ARG, "ofx_focal_loss: alpha %g should be in [0, 1]", alpha);
    if (!out || !workspace || (n > 0 && (!logits || !labels))) return fail(OFX_E_ARG, "ofx_focal_loss: null argument");
    if (workspace_bytes < ofx_loss_workspace_bytes(0)) return fail(OFX_E_WORKSPACE, "ofx_focal_loss: workspace too small");
    OFX_TRY(require_sm100());
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int blocks = loss_blocks(n);
    focal_loss_kernel<<<blocks, kLossThreads, 0, s>>>(logits, labels, n, gamma, alpha, per_elem,
                                                      static_cast<double*>(workspace));
    OFX_LAUNCH_CHECK();
    focal_finish_kernel<<<1, kLossThreads, 0, s>>>(static_cast<const double*>(workspace), blocks, n, out);
    OFX_LAUNCH_CHECK();
    return OFX_OK;
}

extern "C" int ofx_set_wise_ranking_loss(const float* y, const float* y_hat, const float* negatives,
                                         const uint8_t* negative_mask, int32_t batch, int32_t n_neg, int32_t dim,
                                         float margin, double* out, void* workspace, size_t workspace_bytes,
                                         void* stream) {
    if (batch < 1 || n_neg < 0) return fail(OFX_E_SHAPE, "ofx_set_wise_ranking_loss: batch %d, n_neg %d", batch, n_neg);
    if (dim < 4 || dim % 4 || dim > 8192) return fail(OFX_E_SHAPE, "ofx_set_wise_ranking_loss: dim %d must be a multiple of 4, <= 8192", dim);
    if (!y || !y_hat || !out || !workspace || (n_neg > 0 && (!negatives || !negative_mask)))
        return fail(OFX_E_ARG, "ofx_set_wise_ranking_loss: null argument");
    if ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(y_hat) | reinterpret_cast<uintptr_t>(negatives)) % 16)
        return fail(OFX_E_ARG, "ofx_set_wise_ranking_loss: misaligned pointer");
    if (workspace_bytes < ofx_loss_workspace_bytes(batch)) return fail(OFX_E_WORKSPACE, "ofx_set_wise_ranking_loss: workspace too small");
    OFX_TRY(require_sm100());
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    ranking_loss_kernel<<<batch, 128, sizeof(float) * dim, s>>>(y, y_hat, negatives, negative_mask, n_neg, dim, margin,
                                                                static_cast<RankPartial*>(workspace));
    OFX_LAUNCH_CHECK();
    ranking_finish_kernel<<<1, kLossThreads, 0, s>>>(static_cast<const RankPartial*>(workspace), batch, out);
    OFX_LAUNCH_CHECK();
    return OFX_OK;
}

extern "C" int ofx_cp_metrics(const float* logits, const float* labels, int64_t n, float* probs, int64_t* counts,
                              void* stream) {
    if (n < 0 || n > (1ll << 22)) return fail(OFX_E_SHAPE, "ofx_cp_metrics: n %lld not in [0, 2^22]", static_cast<long long>(n));
    if (!counts || (n > 0 && (!logits || !labels || !probs))) return fail(OFX_E_ARG, "ofx_cp_metrics: null argument");
    OFX_TRY(require_sm100());
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    OFX_CUDA(cudaMemsetAsync(counts, 0, 7 * sizeof(int64_t), s));
    if (n == 0) return OFX_OK;
    unsigned long long* c = reinterpret_cast<unsigned long long*>(counts);
    cp_counts_kernel<<<loss_blocks(n), kLossThreads, 0, s>>>(logits, labels, n, probs, c);
    OFX_LAUNCH_CHECK();
    dim3 grid(static_cast<unsigned>((n + kLossThreads - 1) / kLossThreads), static_cast<unsigned>((n + kAucTile - 1) / kAucTile));
    cp_auc_kernel<<<grid, kLossThreads, 0, s>>>(probs, labels, n, c);
    OFX_LAUNCH_CHECK();
    return OFX_OK;
}
