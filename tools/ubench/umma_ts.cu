// Micro-benchmark / known-answer test for tcgen05.mma with the A operand in TENSOR MEMORY
// (cta_group::2, M = 256: 128 rows per CTA) -- the building block of the round-2 fused FFN block,
// where the bf16 hidden chunk written back into TMEM by the mish epilogue feeds GEMM 2 directly.
//
//   part 1  known answer: A (256 x 64, small integers) is written to TMEM with tcgen05.st.32x32b
//           (thread <-> row, register i <-> K elements 2i, 2i+1), B (N x 64) sits in shared memory in
//           the K-major SW128 layout; D = A . B^T is read back with tcgen05.ld and compared exactly.
//   part 2  sustained issue rate, operands resident: TS against SS for N = 256 and N = 128.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../outfitx_b200/csrc umma_ts.cu -o umma_ts
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace ofx;

__device__ __forceinline__ void umma_bf16_pair_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                                  uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
        "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
        "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__host__ __device__ inline int a_val(int row, int k) { return (row * 5 + k * 3) % 7 - 3; }
__host__ __device__ inline int b_val(int n, int k) { return (n * 3 + k * 5) % 5 - 2; }
__device__ __forceinline__ uint32_t bf16_bits(int v) {
    __nv_bfloat16 h = __float2bfloat16_rn(static_cast<float>(v));
    return *reinterpret_cast<unsigned short*>(&h);
}

// ---------------------------------------------------------------- part 1: known answer
template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) ts_check_kernel(int* bad, float* sample) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    // B: this CTA's N/2 rows of the tile (rows n = rank * N/2 + r), 64 K elements each, SW128 K-major
    for (int i = threadIdx.x; i < (N / 2) * 8; i += 128) {
        const int r = i >> 3, c16 = i & 7;
        uint32_t w[4];
        for (int j = 0; j < 4; ++j) {
            const int k = c16 * 8 + 2 * j;
            const int n = static_cast<int>(rank) * (N / 2) + r;
            w[j] = bf16_bits(b_val(n, k)) | (bf16_bits(b_val(n, k + 1)) << 16);
        }
        *reinterpret_cast<uint4*>(smem + (r >> 3) * 1024 + (r & 7) * 128 + ((c16 ^ (r & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc_pair(&slot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = slot;
    const uint32_t t_lane = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    // A: thread <-> row (this CTA's rows rank*128 + threadIdx.x), 64 K elements = 32 packed columns at column 256
    {
        const int row = static_cast<int>(rank) * 128 + threadIdx.x;
        uint32_t v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = bf16_bits(a_val(row, 2 * i)) | (bf16_bits(a_val(row, 2 * i + 1)) << 16);
        tmem_st_32x32(t_lane + 256, v);
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    if (warp == 1 && rank == 0) {
        const uint32_t idesc = umma_idesc_bf16(256, N);
        if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                umma_bf16_pair_ts(tmem, tmem + 256 + 8 * k, umma_desc_k_sw128(smem_u32(smem) + 32 * k), idesc, k ? 1u : 0u);
            umma_commit_pair(&bar, 0b11);
        }
        __syncwarp();
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    const int row = static_cast<int>(rank) * 128 + threadIdx.x;
    int nbad = 0;
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t raw[32];
        tmem_ld_32x32(t_lane + c0, raw);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) {
            int want = 0;
            for (int k = 0; k < 64; ++k) want += a_val(row, k) * b_val(c0 + j, k);
            const float got = __uint_as_float(raw[j]);
            if (got != static_cast<float>(want)) ++nbad;
            if (row == 133 && c0 + j < 8) { sample[c0 + j] = got; sample[8 + c0 + j] = static_cast<float>(want); }
        }
    }
    if (nbad) atomicAdd(bad, nbad);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) { tc_fence_after(); tmem_dealloc_pair(tmem, 512); }
}

// ---------------------------------------------------------------- part 2: rate
template <int TS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) ts_rate_kernel(int n, int iters, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    for (int i = threadIdx.x; i < 128 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc_pair(&slot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = slot;
    {   // something finite in the A columns (384..511)
        uint32_t v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0x3c003c00u;
        for (int c = 384; c < 512; c += 32) tmem_st_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    if (warp == 1 && rank == 0) {
        const uint32_t idesc = umma_idesc_bf16(256, n);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 64 * 1024);
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint64_t bd = umma_desc_k_sw128(b0 + (k & 3) * 32 + (k >> 2) * 16384);
                    if (TS) umma_bf16_pair_ts(tmem + (it & 1) * 128, tmem + 384 + 8 * k, bd, idesc, 1u);
                    else umma_bf16_pair(tmem + (it & 1) * 128, umma_desc_k_sw128(a0 + (k & 3) * 32 + (k >> 2) * 16384), bd, idesc, 1u);
                }
            }
            __syncwarp();
        }
        if (elect_one()) umma_commit_pair(&bar, 0b01);
        __syncwarp();
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        if (lane == 0) out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) { tc_fence_after(); tmem_dealloc_pair(tmem, 512); }
}

template <int TS>
static void rate(const char* name, int n, int grid) {
    long long* d; cudaMalloc(&d, 8 * 256); cudaMemset(d, 0, 8 * 256);
    const int smem = 129 * 1024, iters = 2000;
    cudaFuncSetAttribute(ts_rate_kernel<TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    ts_rate_kernel<TS><<<grid, 128, smem>>>(n, iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[256]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const double cyc = double(h[0]) / (iters * 8.0);
    printf("%-40s grid %3d: %7.1f cycles/MMA (floor %d)  [%s]\n", name, grid, cyc, n / 2, cudaGetErrorString(e));
    cudaFree(d);
}

template <int N>
static void check() {
    int* bad; float* sample;
    cudaMalloc(&bad, 4); cudaMemset(bad, 0, 4);
    cudaMalloc(&sample, 64); cudaMemset(sample, 0, 64);
    const int smem = 64 * 1024;
    cudaFuncSetAttribute(ts_check_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    ts_check_kernel<N><<<2, 128, smem>>>(bad, sample);
    cudaError_t e = cudaDeviceSynchronize();
    int hb = -1; float hs[16];
    cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(hs, sample, 64, cudaMemcpyDeviceToHost);
    printf("TS known answer M=256 N=%d K=64: %d of %d outputs wrong [%s]\n   row 133 got :", N, hb, 256 * N, cudaGetErrorString(e));
    for (int i = 0; i < 8; ++i) printf(" %g", hs[i]);
    printf("\n   row 133 want:");
    for (int i = 0; i < 8; ++i) printf(" %g", hs[8 + i]);
    printf("\n");
    cudaFree(bad); cudaFree(sample);
}

int main() {
    check<256>();
    check<128>();
    for (int grid : {2, 148}) {
        rate<0>("SS cta_group::2 M=256 N=256", 256, grid);
        rate<1>("TS cta_group::2 M=256 N=256", 256, grid);
        rate<0>("SS cta_group::2 M=256 N=128", 128, grid);
        rate<1>("TS cta_group::2 M=256 N=128", 128, grid);
    }
    return 0;
}
