// Micro-benchmark: (1) how many SMs a grid of 4-CTA / 2-CTA clusters with one CTA per SM really occupies,
// (2) the rate at which the warps of a CTA can push register data into a PEER CTA's shared memory
// (st.shared::cluster.v4, SW128-style conflict-free pattern, both directions at once), with the arrival
// signalled by a remote mbarrier arrive (release.cluster) or by st.async complete_tx.
// Decides how the round-2 fused FFN block exchanges hidden chunks between cooperating CTAs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../outfitx_b200/csrc dsmem_rate.cu -o dsmem_rate
#include <cstdio>
#include <cuda_runtime.h>
#include <set>
#include "ptx.cuh"
using namespace ofx;

__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint4 v) {
    asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_async_v4(uint32_t addr, uint4 v, uint32_t bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                 ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(bar) : "memory");
}
__device__ __forceinline__ void arrive_release_cluster(uint32_t bar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory");
}

// MODE 0: plain remote stores + one release.cluster arrive per warp and chunk; MODE 1: st.async complete_tx
template <int MODE>
__global__ void __launch_bounds__(256, 1) dsmem_kernel(int peer_xor, int iters, long long* cycles, int* smid, int* errs) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 2 * 32768);      // [2]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    if (threadIdx.x == 0) {
        uint32_t s; asm volatile("mov.u32 %0, %%smid;" : "=r"(s)); smid[blockIdx.x] = s;
        for (int b = 0; b < 2; ++b) mbar_init(&bar[b], MODE == 0 ? 8 : 1);
        fence_barrier_init();
    }
    __syncthreads();
    cluster_sync_all();
    const uint32_t peer = rank ^ peer_xor;
    // warp w owns rows 32*(w&3)..+31 of a 128-row x 128-byte k-block pair; chunk half (w>>2)
    const int row = (warp & 3) * 32 + lane, half = warp >> 2;
    uint32_t dst[2], pbar[2];
    for (int b = 0; b < 2; ++b) {
        dst[b] = mapa_shared(smem_u32(smem + b * 32768 + half * 16384 + (row >> 3) * 1024 + (row & 7) * 128), peer);
        pbar[b] = mapa_shared(smem_u32(&bar[b]), peer);
    }
    long long t0 = clock64();
    uint32_t ph[2] = {0, 0};
    int bad = 0;
    for (int it = 0; it < iters; ++it) {
        const int b = it & 1;
        if (MODE == 1 && threadIdx.x == 0) mbar_arrive_expect_tx(&bar[b], 32768);
        const uint32_t tag = (it << 16) | (rank << 8);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const uint4 v = make_uint4(tag | c, row, tag, c);
            const uint32_t a = dst[b] + ((c ^ (row & 7)) << 4);
            if (MODE == 0) st_cluster_v4(a, v); else st_async_v4(a, v, pbar[b]);
        }
        if (MODE == 0) {
            __syncwarp();
            if (lane == 0) arrive_release_cluster(pbar[b]);
        }
        // consume what the peer sent (acquire at cluster scope), check one word, then tell nobody: the next write
        // into this buffer is two iterations away and gated by the peer's own wait on OUR data of the same iteration
        mbar_wait_cluster(&bar[b], ph[b]);
        ph[b] ^= 1;
        const uint4 got = *reinterpret_cast<const uint4*>(smem + b * 32768 + half * 16384 + (row >> 3) * 1024 + (row & 7) * 128 + ((3 ^ (row & 7)) << 4));
        if (got.x != (((uint32_t)it << 16) | (peer << 8) | 3u) || got.y != (uint32_t)row) ++bad;
        __syncthreads();      // every warp has read before anyone can be overwritten two iterations later
    }
    long long t1 = clock64();
    if (bad) atomicAdd(errs, bad);
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    __syncthreads();
    cluster_sync_all();
}

template <int MODE>
static void run(const char* name, int cl, int peer_xor, int grid) {
    long long* d; int *smid, *errs;
    cudaMalloc(&d, 8 * 1024); cudaMemset(d, 0, 8 * 1024);
    cudaMalloc(&smid, 4 * 1024); cudaMemset(smid, 0xff, 4 * 1024);
    cudaMalloc(&errs, 4); cudaMemset(errs, 0, 4);
    const int smem = 200 * 1024, iters = 400;     // 200 KB: one CTA per SM, as the real kernel
    cudaFuncSetAttribute(dsmem_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(dsmem_kernel<MODE>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int max_clusters = -1;
    cudaOccupancyMaxActiveClusters(&max_clusters, dsmem_kernel<MODE>, &cfg);
    cudaError_t e = cudaLaunchKernelEx(&cfg, dsmem_kernel<MODE>, peer_xor, iters, d, smid, errs);
    cudaError_t e2 = cudaDeviceSynchronize();
    long long h[1024]; int hs[1024]; int herr = -1;
    cudaMemcpy(h, d, 8 * grid, cudaMemcpyDeviceToHost);
    cudaMemcpy(hs, smid, 4 * grid, cudaMemcpyDeviceToHost);
    cudaMemcpy(&herr, errs, 4, cudaMemcpyDeviceToHost);
    std::set<int> sms(hs, hs + grid);
    long long mx = 0, mn = 1ll << 62;
    for (int i = 0; i < grid; ++i) { mx = h[i] > mx ? h[i] : mx; mn = h[i] < mn ? h[i] : mn; }
    printf("%-34s cluster %d grid %3d: max active clusters %3d, distinct SMs %3zu | %6.0f..%6.0f cycles per 32 KB chunk = %5.1f B/clk out (and in) per SM, %d bad words [%s %s]\n",
           name, cl, grid, max_clusters, sms.size(), double(mn) / iters, double(mx) / iters, 32768.0 * iters / double(mx), herr,
           cudaGetErrorString(e), cudaGetErrorString(e2));
    cudaFree(d); cudaFree(smid); cudaFree(errs);
}

int main() {
    run<0>("st.shared::cluster + release arrive", 2, 1, 148);
    run<1>("st.async complete_tx", 2, 1, 148);
    run<0>("st.shared::cluster + release arrive", 4, 2, 148);
    run<1>("st.async complete_tx", 4, 2, 148);
    run<0>("st.shared::cluster + release arrive", 4, 2, 132);
    run<1>("st.async complete_tx", 4, 2, 132);
    run<1>("st.async complete_tx", 4, 2, 4);
    return 0;
}
