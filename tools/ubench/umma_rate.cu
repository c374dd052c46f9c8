// Micro-benchmark: sustained tcgen05.mma issue/execute rate of one CTA pair (cta_group::2) or one
// CTA (cta_group::1) for a given instruction shape, operands resident in shared memory (garbage
// data, K-major SW128 descriptors), accumulating into TMEM.  Prints cycles per MMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../outfitx_b200/csrc umma_rate.cu -o umma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace ofx;

template <int PAIR>
__global__ void __launch_bounds__(128, 1)
rate_kernel(int m, int n, int iters, int b_mn_major, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0;
    for (int i = threadIdx.x; i < 128 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) { if (PAIR) tmem_alloc_pair(&slot, 512); else tmem_alloc(&slot, 512); }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (warp == 1 && rank == 0) {
        const uint32_t idesc = umma_idesc_bf16(m, n) | (b_mn_major ? (1u << 16) : 0u);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 64 * 1024);
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint64_t ad = umma_desc_k_sw128(a0 + (k & 3) * 32 + (k >> 2) * 16384);
                    const uint64_t bd = b_mn_major ? umma_desc_k_sw128(b0 + k * 2048) : umma_desc_k_sw128(b0 + (k & 3) * 32 + (k >> 2) * 16384);
                    if (PAIR) umma_bf16_pair(tmem + (it & 1) * 256, ad, bd, idesc, 1u);
                    else umma_bf16(tmem + (it & 1) * 256, ad, bd, idesc, 1u);
                }
            }
            __syncwarp();
        }
        if (elect_one()) { if (PAIR) umma_commit_pair(&bar, 0b01); else umma_commit(&bar); }
        __syncwarp();
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        if (lane == 0) out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();
    if (warp == 0) { tc_fence_after(); if (PAIR) tmem_dealloc_pair(tmem, 512); else tmem_dealloc(tmem, 512); }
}

template <int PAIR>
static void run(const char* name, int m, int n, int b_mn, int grid) {
    long long* d; cudaMalloc(&d, 8 * 256); cudaMemset(d, 0, 8 * 256);
    const int smem = 129 * 1024, iters = 2000;
    cudaFuncSetAttribute(rate_kernel<PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = PAIR ? 2 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, rate_kernel<PAIR>, m, n, iters, b_mn, d);
    cudaError_t e2 = cudaDeviceSynchronize();
    long long h[256]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const double cyc = double(h[0]) / (iters * 8.0);
    const double macs = double(m) * n * 16 / cyc / (PAIR ? 2 : 1);
    printf("%-44s grid %3d: %7.1f cycles/MMA  %6.0f MAC/cycle/SM (%.0f%% of 4096)  [%s %s]\n", name, grid, cyc, macs, macs / 40.96,
           cudaGetErrorString(e), cudaGetErrorString(e2));
    cudaFree(d);
}

int main() {
    for (int grid : {2, 148}) {
        run<0>("cta_group::1 M=128 N=256", 128, 256, 0, grid);
        run<0>("cta_group::1 M=128 N=128", 128, 128, 0, grid);
        run<0>("cta_group::1 M=64  N=256", 64, 256, 0, grid);
        run<1>("cta_group::2 M=256 N=256", 256, 256, 0, grid);
        run<1>("cta_group::2 M=256 N=128", 256, 128, 0, grid);
        run<1>("cta_group::2 M=256 N=128 (B MN-major)", 256, 128, 1, grid);
        run<1>("cta_group::2 M=128 N=256 (ffn_block today)", 128, 256, 0, grid);
        run<1>("cta_group::2 M=128 N=128", 128, 128, 0, grid);
    }
    return 0;
}
