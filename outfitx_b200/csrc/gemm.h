#pragma once
#include "common.h"

namespace ofx {

// C[m,n] = A[m,k] . W[n,k]^T  (+bias[n]) (mish) (+residual[m,n] fp32)
struct GemmArgs {
    const void* a;         // bf16 (tc) or fp32 (simt), row pitch lda elements
    long long lda;
    const void* w;         // same element type as a, row pitch ldw
    long long ldw;
    int m;                 // rows (host upper bound)
    const int* m_dev;      // optional device-side row count, <= m
    int n, k;
    const float* bias;
    int act_mish;
    const float* residual; // fp32, may alias out when out_f32
    long long ldr;
    void* out;             // bf16 or fp32 (tc: out_f32 selects; simt: always fp32)
    long long ldo;
    int out_f32;
};

int gemm_bf16(const GemmArgs& g, cudaStream_t stream);
int gemm_f32(const GemmArgs& g, cudaStream_t stream);

}  // namespace ofx
