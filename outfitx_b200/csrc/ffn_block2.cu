// Fused feed-forward block of one encoder layer (d_model 512), round-2 form: FOUR SMs per 256 token rows.
//
//     x  <-  x + W2 . mish(W1 . LN2(x) + b1) + b2            (rows of the token matrix, in place)
//     h_next <- bf16(LN1_next(x))                             (optional: norm1 of the following layer)
//
// i.e. torch.nn.TransformerEncoderLayer's `x = x + _ff_block(norm2(x))` (pre-LN slow path,
// torch/nn/modules/transformer.py:950,980-982) as the reference builds it at
// /root/reference/src/models/outfit_x.py:32-45 (activation = F.mish, d_ffn 2024 zero-padded to 2048).
//
// Why a new shape.  The round-1 kernel (ffn_block.cu) gives every CTA 64 token rows because a 128-row x 512-column
// fp32 output tile alone fills the 512 TMEM columns.  With 64 rows per CTA each weight byte parked in shared memory
// feeds one 64-row MMA: operand reads (96 B/clk) plus TMA refill (64 B/clk) exceed the 128 B/clk an SM moves, and the
// kernel sat at 46 % tensor-pipe activity (ncu, profiles/r1_ncu_ffn_block_ln.txt).  Here a CTA keeps 128 rows and HALF
// of the output columns:
//
//   * a CTA PAIR (cluster of 2, tcgen05 cta_group::2, M = 256) owns 256 token rows x 256 output columns and computes
//     HALF of the hidden units for those rows; its PARTNER pair (the neighbouring cluster) owns the other 256 output
//     columns and the other half of the hidden units of the SAME rows.  TMEM per CTA: output 128 lanes x 256 columns +
//     two 128-column hidden-chunk accumulators = 512.
//   * GEMM 1 (hidden chunk of 128 units): A = LN2(x) rows resident in shared memory (128 KB, written by the LayerNorm
//     warps straight into the UMMA layout), B = W1 rows streamed by TMA.
//   * the mish epilogue writes the bf16 hidden chunk BACK INTO TMEM over its own fp32 accumulator (tcgen05.st) and
//     GEMM 2 of that chunk takes its A operand from TMEM (tools/ubench/umma_ts.cu: same issue rate as from shared
//     memory) -- no shared-memory round trip for the pair's own half;
//   * the same bf16 chunk is handed to the partner through L2: staged once in shared memory, written by a TMA bulk
//     store into a small ring of scratch slots (19 MB in total, L2 resident), announced with a release counter; the
//     partner's TMA producer acquires the counter and loads the chunk into its weight ring as the A operand of its
//     "peer" GEMM 2.  (Direct stores into the partner's shared memory were measured first -- tools/ubench/dsmem_rate.cu:
//     8-9 B/clk per SM and 4-CTA clusters leave 16 of the 148 SMs idle -- which is why the exchange goes through L2
//     between two 2-CTA clusters instead.)
//
// Shared-memory traffic per 256-row tile and CTA: GEMM 1 operands 1.5 MB, GEMM 2 operands 0.75 MB, TMA fills 1.25 MB,
// staging + H 0.9 MB = 4.4 MB per 32.8k tensor-pipe cycles = 134 B/clk at the MMA floor (the old shape needed 190).
//
// Warp roles (16 warps): 0 TMA producer (weights, H), 1 MMA issuer (leader CTA), 2 TMEM allocator + peer loader,
// 3 exporter, 4-11 epilogue (two warpgroups
// alternating hidden chunks, then the residual epilogue), 12-15 LayerNorm (prologue LN2 -> H; norm1 of the next layer
// for the previous tile in their idle time).  Registers re-partitioned with setmaxnreg (80 / 152 / 128: they sum to the 512 x 128 the CTA is launched with -- the pool is the CTA's own allocation).
//
// Cross-cluster protocol (global int counters, zeroed per launch):
//   uflag[group][side][rank][slot] += 1 per epilogue warp once its part of a chunk is in that scratch slot (a slot's
//                                     tenants follow one another, so 8 x generation identifies the chunk)
//   uack [group][side]             += 1 per k-block of side's chunks the partner has pulled into its shared memory
//   xflag[group][rank][side]       += 1 per epilogue warp and tile once its new x columns are in global memory
// Every wait is on a counter of STRICTLY earlier chunks, so the two pairs cannot deadlock as long as
// both are resident -- the grid is at most one CTA per SM and all its CTAs are co-resident.
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>

#include "encoder_ops.h"
#include "ptx.cuh"

namespace ofx {
namespace ffn2 {

constexpr int DM = 512;               // d_model (K of GEMM 1, N of GEMM 2 over both partners)
constexpr int KB1 = DM / 64;          // k-blocks of GEMM 1
constexpr int CH = 128;               // hidden units per chunk (N of GEMM 1)
constexpr int ROWS = 128;             // token rows per CTA
constexpr int TILE = 2 * ROWS;        // token rows per CTA pair (and per group of two pairs)
constexpr int NOUT = DM / 2;          // output columns per pair
constexpr int KBLK_BYTES = ROWS * 128;        // one 128-row x 64-element bf16 operand block
constexpr int H_BYTES = KB1 * KBLK_BYTES;     // 128 KB
constexpr int STG_WARP_BYTES = 32 * 128;      // per epilogue warp: 32 rows x 128 B
constexpr int N_EPI_WARPS = 8, N_LN_WARPS = 4;
constexpr int STG_BYTES = N_EPI_WARPS * STG_WARP_BYTES;   // 32 KB
constexpr int SLOT_BYTES = 16384;             // ring slot: 2 k-blocks of W1 (64 rows each) | 1 k-block of W2 | 1 k-block of the partner's U
constexpr int NS = 4;
constexpr int BAR_BYTES = 512;
constexpr int SMEM_BYTES = 1024 + H_BYTES + STG_BYTES + NS * SLOT_BYTES + BAR_BYTES;
constexpr int NTHREADS = 512;
constexpr int EPI_WARP0 = 4, LN_WARP0 = 12;
constexpr uint32_t TM_OUT = 0, TM_ACC1 = 256;   // TMEM columns
constexpr int NSLOT = 16;             // scratch slots per (group, side, rank): two tiles of chunks (nch <= 8), see slot_of()
constexpr int LAG = 2;                // the partner's chunk j is consumed in step j + LAG (its announcement is deferred by one chunk of the producing warpgroup)

struct Params {
    float* x;               // (rows, 512) fp32 residual stream, updated in place
    int rows;               // host-side row count (upper bound when rows_dev != nullptr)
    const int* rows_dev;    // optional device-side row count
    const float* ln_w;      // LayerNorm 2
    const float* ln_b;
    const float* b1;        // (n_chunks * 256) fp32, zero beyond d_ffn
    const float* b2;        // (512)
    int nch;                // hidden chunks per side = padded d_ffn / 256
    __nv_bfloat16* h_next;  // optional (rows, 512) bf16: LayerNorm 1 of the NEXT layer applied to the new x
    const float* lnn_w;
    const float* lnn_b;
    int* uflag;             // [n_groups][side 2][rank 2][NSLOT]
    int* uack;              // [n_groups][side 2]
    int* xflag;             // [n_groups][rank 2][side 2]
    __nv_bfloat16* scratch; // [n_groups][side 2][rank 2][NSLOT][128][128] bf16: the hidden-chunk exchange ring
    __nv_bfloat16* hbuf;    // [n_ctas][2][128][512] bf16: LN2 rows staged one tile ahead by the LayerNorm warps
    long long* prof;        // instrumented builds (OFX_FFN_PROF=1): 32 cycle counters per CTA, else nullptr
    int mode;               // instrumented builds (OFX_FFN2_MODE): bit 0 = the weight producer pulls the partner's chunks itself
};

__device__ __forceinline__ float mish_fast(float x) {
    // x * n / (n + 2), n = w (w + 2), w = e^x   (see gemm.cu)
    float w, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w) : "f"(fminf(x, 40.f) * 1.4426950408889634f));
    const float n = fmaf(w, w, w + w);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(n + 2.f));
    return x * n * r;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Waits.  Product build: the bounded spins of ptx.cuh (trap instead of hanging the GPU).  Instrumented build
// (-DOFX_DEBUG, python -m outfitx_b200.build --debug): a wait that exceeds ~0.5 s reports who is stuck on what and
// gives up, so a wedged protocol prints its wait-for graph instead of dying silently.
#ifdef OFX_DEBUG
__device__ __forceinline__ unsigned long long gtime_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define FFN2_STUCK(tag, a, b) printf("%llu ffn2 stuck: %-10s cta %3d (group %d side %d rank %d) warp %2d  %d %d\n", gtime_ns() / 1000ull, tag, \
                                     static_cast<int>(blockIdx.x), static_cast<int>(blockIdx.x) >> 2, (static_cast<int>(blockIdx.x) >> 1) & 1, \
                                     static_cast<int>(blockIdx.x) & 1, static_cast<int>(threadIdx.x >> 5), a, b)
__device__ __forceinline__ void bar_wait(uint64_t* bar, uint32_t parity, const char* tag, int a = 0, int b = 0) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins == (1u << 22)) { if ((threadIdx.x & 31) == 0) FFN2_STUCK(tag, a, b); break; }
    }
}
__device__ __forceinline__ void flag_wait(const int* ptr, int target, const char* tag, int a = 0, int b = 0) {
    uint32_t spins = 0;
    while (ld_acquire_gpu(ptr) < target) {
        __nanosleep(40);
        if (++spins == (1u << 21)) { FFN2_STUCK(tag, a, ld_acquire_gpu(ptr) * 1000 + target); break; }
    }
}
#else
__device__ __forceinline__ void bar_wait(uint64_t* bar, uint32_t parity, const char*, int = 0, int = 0) { mbar_wait(bar, parity); }
__device__ __forceinline__ void flag_wait(const int* ptr, int target, const char*, int = 0, int = 0) { wait_flag_ge(ptr, target); }
#endif

#ifdef OFX_DEBUG
#define PROF(slot, stmt) do { if (p.prof) { const long long tq__ = clock64(); stmt; prof_acc[slot] += clock64() - tq__; } else { stmt; } } while (0)
#define PROF_DECL() long long prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; const long long prof_t0 = clock64()
#define PROF_DUMP(base) do { if (p.prof && lane == 0) { long long* o__ = p.prof + blockIdx.x * 32 + (base); o__[0] = clock64() - prof_t0; \
        for (int i__ = 0; i__ < 7; ++i__) o__[1 + i__] = prof_acc[i__]; } } while (0)
#else
#define PROF(slot, stmt) do { stmt; } while (0)
#define PROF_DECL() do { } while (0)
#define PROF_DUMP(base) do { } while (0)
#endif

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
ffn_block2_kernel(const __grid_constant__ CUtensorMap tm_w1, const __grid_constant__ CUtensorMap tm_w2,
                  const __grid_constant__ CUtensorMap tm_ust, const __grid_constant__ CUtensorMap tm_uld,
                  const __grid_constant__ CUtensorMap tm_h, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* s_h = smem;
    uint8_t* s_stg = smem + H_BYTES;
    uint8_t* s_w = s_stg + STG_BYTES;                    // ring
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_w + NS * SLOT_BYTES);
    uint64_t* w_full = bars;                    // [NS]  leader's are used (tx from both CTAs)
    uint64_t* w_empty = bars + NS;              // [NS]  per CTA (multicast commit)
    uint64_t* acc1_full = bars + 2 * NS;        // [2]   per CTA (multicast commit)
    uint64_t* u_full = acc1_full + 2;           // [2]   leader's: 8 epilogue-warp arrivals
    uint64_t* out_full = u_full + 2;            //       per CTA (multicast commit)
    uint64_t* out_empty = out_full + 1;         //       leader's: 16 arrivals
    uint64_t* h_full = out_empty + 1;           //       leader's: 16 LN-warp arrivals
    uint64_t* h_empty = h_full + 1;             //       per CTA (multicast commit)
    uint64_t* h_staged = h_empty + 1;           //       per CTA: 4 LN-warp arrivals, the next tile's LN2 rows are in hbuf
    uint64_t* hbuf_free = h_staged + 1;         // [2]   per CTA: producer arrival, that half of hbuf has been pulled into H and consumed
    uint64_t* stg_full = hbuf_free + 2;         // [8]   per CTA: epilogue warp w has staged its part of a hidden chunk
    uint64_t* stg_free = stg_full + N_EPI_WARPS; // [8]  per CTA: the exporter's bulk store has read warp w's staging block
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(stg_free + N_EPI_WARPS);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();          // 0 = leader of the pair
    const int pair = blockIdx.x >> 1;
    const int group = pair >> 1, side = pair & 1, n_groups = gridDim.x >> 2;
    const int n_rows = p.rows_dev ? min(*p.rows_dev, p.rows) : p.rows;
    const int n_tiles = (n_rows + TILE - 1) / TILE;
    const int nch = p.nch;
    int* my_uflag = p.uflag + ((group * 2 + side) * 2 + rank) * NSLOT;              // what THIS CTA's epilogue publishes
    const int* peer_uflag = p.uflag + ((group * 2 + (side ^ 1)) * 2 + rank) * NSLOT; // what the partner's same-rank CTA publishes
    int* xflag = p.xflag + (group * 2 + rank) * 2;                         // [side]: the rows' two column halves
    // scratch rows: ((((group * 2 + side) * 2 + rank) * NSLOT + slot) * 128 + row), 128 hidden columns each
    const int my_srow = ((group * 2 + side) * 2 + static_cast<int>(rank)) * NSLOT * ROWS;
    const int peer_srow = ((group * 2 + (side ^ 1)) * 2 + static_cast<int>(rank)) * NSLOT * ROWS;
    // Chunk seq (running number over this group's tiles) lives in scratch slot (tile parity, j) and is that slot's
    // gen-th tenant.  Two tiles of slots make acknowledgements unnecessary: a pair cannot finish tile t - 1 without ALL of
    // the partner's chunks of tile t - 1, which the partner produces only after its MMA warp has issued every peer GEMM 2
    // of tile t - 2 -- i.e. after the bulk loads that emptied the slots of tile t - 2 have completed.  (An
    // acknowledgement counter was tried first: the release in front of each atomic cost the MMA warp ~20 % of its time.)
    auto slot_of = [&](int seq) { return ((seq / nch) & 1) * nch + seq % nch; };
    auto gen_of = [&](int seq) { return (seq / nch) >> 1; };

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_w1);
        tma_prefetch_desc(&tm_w2);
        tma_prefetch_desc(&tm_ust);
        tma_prefetch_desc(&tm_uld);
        tma_prefetch_desc(&tm_h);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < NS; ++i) {
            mbar_init(&w_full[i], 1);
            mbar_init(&w_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc1_full[i], 1);
            mbar_init(&u_full[i], 2 * N_EPI_WARPS);        // all 8 epilogue warps x 2 CTAs
        }
        mbar_init(out_full, 1);
        mbar_init(out_empty, 2 * N_EPI_WARPS);
        mbar_init(h_full, 1);                   // leader's arrive.expect_tx; the bytes of both CTAs' bulk loads
        mbar_init(h_empty, 1);
        mbar_init(h_staged, N_LN_WARPS);
        mbar_init(&hbuf_free[0], 1);
        mbar_init(&hbuf_free[1], 1);
        for (int i = 0; i < N_EPI_WARPS; ++i) {
            mbar_init(&stg_full[i], 1);
            mbar_init(&stg_free[i], 1);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc_pair(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 80;" ::: "memory");
    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer (weights + the partner's hidden chunks)
        if (lane == 0) {
            PROF_DECL();
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t full_addr0 = mapa_shared(smem_u32(&w_full[0]), 0);
            auto acquire = [&]() -> uint32_t {      // waits for the slot, arms the leader's barrier, returns its cluster address
                PROF(0, bar_wait(&w_empty[stage], phase ^ 1, "w_empty", stage));
                if (rank == 0) mbar_arrive_expect_tx(&w_full[stage], 2 * SLOT_BYTES);   // both CTAs' bytes
                return full_addr0 + stage * 8;
            };
            auto advance = [&]() { if (++stage == NS) { stage = 0; phase ^= 1; } };
            auto g1 = [&](int j) {   // W1 rows of own chunk j: this CTA's 64 of the 128 hidden units, k-blocks (2i, 2i+1) per slot
                const int row = (side * nch + j) * CH + static_cast<int>(rank) * 64;
                for (int i = 0; i < KB1 / 2; ++i) {
                    const uint32_t bar = acquire();
                    tma_load_2d_pair(s_w + stage * SLOT_BYTES, &tm_w1, bar, (2 * i) * 64, row, kEvictLast);
                    tma_load_2d_pair(s_w + stage * SLOT_BYTES + SLOT_BYTES / 2, &tm_w1, bar, (2 * i + 1) * 64, row, kEvictLast);
                    advance();
                }
            };
            auto w2slot = [&](int hid0) {   // W2: this CTA's 128 of the pair's 256 output columns x 64 hidden units
                const uint32_t bar = acquire();
                tma_load_2d_pair(s_w + stage * SLOT_BYTES, &tm_w2, bar, hid0, side * NOUT + static_cast<int>(rank) * 128, kEvictLast);
                advance();
            };
            auto g2own = [&](int j) {
                for (int i = 0; i < 2; ++i) w2slot((side * nch + j) * CH + i * 64);
            };
            auto g2peer = [&](int j, int seq) {      // the A slots (the partner's chunk) belong to the peer loader (warp 2)
                if (p.mode & 1) {
                    flag_wait(peer_uflag + slot_of(seq), N_EPI_WARPS * (gen_of(seq) + 1), "peer_uflag", seq);
                    fence_proxy_async_all();
                }
                for (int i = 0; i < 2; ++i) {
                    if (p.mode & 1) {
                        const uint32_t bar = acquire();
                        tma_load_2d_pair(s_w + stage * SLOT_BYTES, &tm_uld, bar, i * 64, peer_srow + slot_of(seq) * ROWS, kEvictFirst);
                    }
                    advance();
                    w2slot(((side ^ 1) * nch + j) * CH + i * 64);
                }
            };
            int seq0 = 0;       // chunk number of this tile's chunk 0
            uint32_t hph = 0;
            const uint32_t hfull_addr = mapa_shared(smem_u32(h_full), 0);
            const int hrow0 = static_cast<int>(blockIdx.x) * 2 * ROWS;
            for (int t = group; t < n_tiles; t += n_groups, seq0 += nch) {
                // this tile's LN2 rows: staged in hbuf by the LayerNorm warps during the previous tile; H itself is free
                // once GEMM 1 of the previous tile has retired
                PROF(2, bar_wait(h_staged, hph, "h_staged", t));
                PROF(2, bar_wait(h_empty, hph ^ 1, "h_empty", t));
                if (t != group) mbar_arrive(&hbuf_free[hph ^ 1]);     // the previous tile's half of hbuf may be refilled
                if (rank == 0) mbar_arrive_expect_tx(h_full, 2 * H_BYTES);
#pragma unroll 1
                for (int kb = 0; kb < KB1; ++kb)
                    tma_load_2d_pair(s_h + kb * KBLK_BYTES, &tm_h, hfull_addr, kb * 64, hrow0 + static_cast<int>(hph) * ROWS, kEvictFirst);
                hph ^= 1;
                g1(0);
                if (nch > 1) g1(1);
                for (int j = 0; j < nch; ++j) {
                    g2own(j);
                    if (j + 2 < nch) g1(j + 2);
                    if (j >= LAG) g2peer(j - LAG, seq0 + j - LAG);
                }
                for (int j = nch > LAG ? nch - LAG : 0; j < nch; ++j) g2peer(j, seq0 + j);
            }
            PROF_DUMP(0);
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (leader only)
        // The whole warp runs the loop (warp-uniform control flow and addresses, so the descriptors live in
        // uniform registers); one elected lane issues the tcgen05 ops.
        if (rank == 0) {
            constexpr uint32_t idesc_g1 = umma_idesc_bf16(256, CH);
            constexpr uint32_t idesc_g2 = umma_idesc_bf16(256, NOUT);
            constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO, version, SW128
            const uint32_t h_lo = ((smem_u32(s_h) & 0x3FFFF) >> 4) | (1u << 16);
            const uint32_t w_lo = ((smem_u32(s_w) & 0x3FFFF) >> 4) | (1u << 16);
            auto desc = [](uint32_t lo) { return (static_cast<uint64_t>(kDescHi) << 32) | lo; };
            PROF_DECL();
            int stage = 0;
            uint32_t phase = 0, tphase = 0, uph = 0;
            uint32_t ready = 0;     // result of the early poll of w_full[stage]
            // waits for ring slot `stage`, then polls the NEXT slot's barrier: a try_wait costs ~100 cycles even on a
            // completed phase, and the MMAs of a slot run for only 512, so the poll is issued before them and its
            // latency overlaps their issue instead of opening a bubble in the tensor pipe (the round-1 kernel's trick)
            auto wait_slot = [&]() -> uint32_t {     // this ring slot's descriptor base
                if (!ready) PROF(0, bar_wait(&w_full[stage], phase, "w_full", stage));
                tc_fence_after();
                int ns = stage + 1;
                uint32_t nph = phase;
                if (ns == NS) { ns = 0; nph ^= 1; }
                ready = mbar_test_wait(&w_full[ns], nph) ? 1u : 0u;      // a probe, not a wait: it must not hold back this slot's MMAs
                return w_lo + stage * (SLOT_BYTES >> 4);
            };
            auto release_slot = [&]() {              // inside the elected lane: frees the slot in both CTAs when the MMAs retire
                umma_commit_pair(&w_empty[stage], 0b11);
            };
            auto advance = [&]() { if (++stage == NS) { stage = 0; phase ^= 1; } };
            auto g1 = [&](int j, bool last) {
                const uint32_t d = tmem_base + TM_ACC1 + (j & 1) * CH;
                for (int i = 0; i < KB1 / 2; ++i) {
                    const uint32_t b = wait_slot();
                    const uint32_t a = h_lo + (2 * i) * (KBLK_BYTES >> 4);
                    if (elect_one()) {
#pragma unroll
                        for (int sub = 0; sub < 2; ++sub)
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_bf16_pair(d, desc(a + sub * (KBLK_BYTES >> 4) + 2 * k),
                                               desc(b + sub * (SLOT_BYTES >> 5) + 2 * k), idesc_g1,
                                               (i | sub | k) != 0 ? 1u : 0u);
                        release_slot();
                    }
                    __syncwarp();
                    advance();
                }
                if (elect_one()) {
                    umma_commit_pair(&acc1_full[j & 1], 0b11);
                    if (last) umma_commit_pair(h_empty, 0b11);   // H may be refilled
                }
                __syncwarp();
            };
            auto g2own = [&](int j, bool first) {      // A = bf16 hidden chunk in TMEM (written over acc1[j&1] by the epilogue)
                PROF(1, bar_wait(&u_full[j & 1], (uph >> (j & 1)) & 1, "u_full", j));
                uph ^= 1u << (j & 1);
                tc_fence_after();
                const uint32_t a_t = tmem_base + TM_ACC1 + (j & 1) * CH;
                for (int i = 0; i < 2; ++i) {
                    const uint32_t b = wait_slot();
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16_pair_ts(tmem_base + TM_OUT, a_t + i * 32 + k * 8, desc(b + 2 * k), idesc_g2,
                                              (first && i == 0 && k == 0) ? 0u : 1u);
                        release_slot();
                    }
                    __syncwarp();
                    advance();
                }
            };
            auto g2peer = [&]() {                      // A = the partner's chunk, pulled into the ring by the producer
                for (int i = 0; i < 2; ++i) {
                    const uint32_t a = wait_slot();
                    const int a_stage = stage;
                    advance();
                    const uint32_t b = wait_slot();
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16_pair(tmem_base + TM_OUT, desc(a + 2 * k), desc(b + 2 * k), idesc_g2, 1u);
                        umma_commit_pair(&w_empty[a_stage], 0b11);
                        release_slot();
                    }
                    __syncwarp();
                    advance();
                }
            };
            for (int t = group; t < n_tiles; t += n_groups) {
                PROF(2, bar_wait(h_full, tphase, "h_full", t));
                tc_fence_after();
                g1(0, nch == 1);
                if (nch > 1) g1(1, nch == 2);
                for (int j = 0; j < nch; ++j) {
                    if (j == 0) {   // the previous tile's output has left TMEM
                        PROF(3, bar_wait(out_empty, tphase ^ 1, "out_empty", t));
                        tc_fence_after();
                    }
                    g2own(j, j == 0);
                    if (j + 2 < nch) g1(j + 2, j + 3 == nch);
                    if (j >= LAG) g2peer();
                }
                for (int j = nch > LAG ? nch - LAG : 0; j < nch; ++j) g2peer();
                if (elect_one()) umma_commit_pair(out_full, 0b11);
                __syncwarp();
                tphase ^= 1;
            }
            PROF_DUMP(8);
        }
    } else if (warp == 3) {
        // ------------------------------------------------------------ peer loader (one thread): pulls the partner's hidden
        // chunks into the ring.  It walks the same slot schedule as the weight producer and owns the A slots of the peer
        // GEMM 2; waiting for the partner's counter and the proxy fence behind it (acquired generic view -> bulk load)
        // cost ~2.5k cycles per chunk, during which the weight producer -- which used to do this -- fed nothing.
        if (lane == 0 && !(p.mode & 1)) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t full_addr0 = mapa_shared(smem_u32(&w_full[0]), 0);
            // Slots owned by the weight producer are WAITED for as well (not just counted): an mbarrier wait only tells
            // "the phase of this parity has completed", so a waiter must never be more than one revolution away from the
            // ring's real position -- counting ahead made the first wait of this thread pass on the fresh barriers.
            auto skip = [&](int n) {
                for (int i = 0; i < n; ++i) {
                    bar_wait(&w_empty[stage], phase ^ 1, "w_empty(s)", stage);
                    if (++stage == NS) { stage = 0; phase ^= 1; }
                }
            };
            auto g2peer = [&](int seq) {
                flag_wait(peer_uflag + slot_of(seq), N_EPI_WARPS * (gen_of(seq) + 1), "peer_uflag", seq);   // the partner CTA has stored chunk seq
                fence_proxy_async_all();
                for (int i = 0; i < 2; ++i) {
                    bar_wait(&w_empty[stage], phase ^ 1, "w_empty(p)", stage);
                    if (rank == 0) mbar_arrive_expect_tx(&w_full[stage], 2 * SLOT_BYTES);
                    tma_load_2d_pair(s_w + stage * SLOT_BYTES, &tm_uld, full_addr0 + stage * 8, i * 64, peer_srow + slot_of(seq) * ROWS, kEvictFirst);
                    if (++stage == NS) { stage = 0; phase ^= 1; }
                    skip(1);      // the W2 slot behind it
                }
            };
            int seq0 = 0;
            for (int t = group; t < n_tiles; t += n_groups, seq0 += nch) {
                skip(nch > 1 ? 8 : 4);                       // G1(0), G1(1)
                for (int j = 0; j < nch; ++j) {
                    skip(2);                                 // G2own(j)
                    if (j + 2 < nch) skip(4);                // G1(j + 2)
                    if (j >= LAG) g2peer(seq0 + j - LAG);
                }
                for (int j = nch > LAG ? nch - LAG : 0; j < nch; ++j) g2peer(seq0 + j);
            }
        }
    } else if (warp == 2) {
        // ------------------------------------------------------------ exporter (one warp): hands the staged hidden chunks to
        // the partner.  Plain coalesced copies, staging block -> L2 scratch (4 rows of 128 B per instruction), then ONE
        // release on the chunk's counter.  A TMA bulk store was tried first: its completion wait and the proxy fence that
        // must precede the generic-proxy release cost ~3k cycles per chunk (ncu: FENCE.VIEW.ASYNC + ERRBAR + MEMBAR = 9 %
        // of all stall samples), first on the epilogue warps -- i.e. on the G1 -> E -> G2 chain -- then on a dedicated
        // thread that could not keep up with one chunk per 4k cycles.  Generic stores need neither.
        {
            uint32_t ph = 0;
            int seq = 0;
            const int r4 = lane >> 3, c16 = lane & 7;
            for (int t = group; t < n_tiles; t += n_groups) {
                for (int j = 0; j < nch; ++j, ++seq) {
                    uint8_t* slot = reinterpret_cast<uint8_t*>(p.scratch) + (static_cast<size_t>(my_srow) + slot_of(seq) * ROWS) * (CH * 2);
#pragma unroll 1
                    for (int w = 0; w < N_EPI_WARPS; ++w) {
                        bar_wait(&stg_full[w], ph, "stg_full", seq, w);
                        const uint32_t blk = smem_u32(s_stg + w * STG_WARP_BYTES);
                        uint4 v[8];
#pragma unroll
                        for (int it = 0; it < 8; ++it) {
                            const int r = it * 4 + r4;
                            v[it] = lds128(blk + r * 128 + ((c16 ^ (r & 7)) << 4));
                        }
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&stg_free[w]);          // the block may be rewritten
                        uint8_t* dst = slot + static_cast<size_t>((w & 3) * 32 + r4) * (CH * 2) + (w >> 2) * 128 + c16 * 16;
#pragma unroll
                        for (int it = 0; it < 8; ++it)
                            asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(dst + static_cast<size_t>(it) * 4 * (CH * 2)),
                                         "r"(v[it].x), "r"(v[it].y), "r"(v[it].z), "r"(v[it].w) : "memory");
                    }
                    ph ^= 1;
                    __syncwarp();        // every lane's stores happen-before lane 0's release (cumulative at gpu scope)
                    if (lane == 0) red_release_gpu_add(my_uflag + slot_of(seq), N_EPI_WARPS);
                }
            }
        }
    }
    } else if (warp < LN_WARP0) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 152;" ::: "memory");
        // ------------------------------------------------------------ epilogue warps
        const int q = warp & 3;                       // TMEM lane quarter = rows q*32 .. +31 of this CTA's 128
        const int wg = (warp - EPI_WARP0) >> 2;       // warpgroup: which 64 of a chunk's 128 hidden units; output column half in the drain
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        uint8_t* stg = s_stg + (warp - EPI_WARP0) * STG_WARP_BYTES;      // this warp's 32-row x 128-byte staging block
        const uint32_t stg_u = smem_u32(stg);
        const uint32_t stg_row = stg_u + lane * 128;
        const int sw = lane & 7;
        const uint32_t ufull0 = mapa_shared(smem_u32(&u_full[0]), 0);
        const uint32_t ufull1 = mapa_shared(smem_u32(&u_full[1]), 0);
        const uint32_t oempty = mapa_shared(smem_u32(out_empty), 0);
        uint32_t a1ph = 0, tphase = 0;
        int seq0 = 0;
        const int ew = warp - EPI_WARP0;
        uint32_t fph = 0;
        bool pending_free = false;     // a staged block has been handed to the exporter and not yet waited for
        PROF_DECL();
        auto staging_free = [&]() {    // before any write to this warp's staging block
            if (pending_free) {
                PROF(1, bar_wait(&stg_free[ew], fph, "stg_free", ew));
                fph ^= 1;
                pending_free = false;
            }
        };
        // E(j): bf16(mish(acc1 + b1)) of this warp's 32 rows x 64 hidden units -> TMEM (own GEMM 2) and scratch (partner).
        // ALL eight warps work on every chunk (two per lane quarter, 64 columns each): what limits the block is the
        // LATENCY G1(j) -> E(j) -> G2(j) -- the MMA warp can run only one GEMM 1 ahead (two accumulators) -- so a chunk
        // must not sit in one warpgroup for 8k cycles.  The bf16 chunk overwrites its own fp32 accumulator: the warp of
        // the upper 64 columns stores into columns the lower warp reads, hence both load their columns completely and
        // meet at a named barrier before the first store.  The 64 bias values sit two per lane, broadcast by shuffles.
        auto mish_chunk = [&](int j, int seq) {
            const int b = j & 1;
            const float2 bl = __ldg(reinterpret_cast<const float2*>(p.b1 + (side * nch + j) * CH + wg * 64) + lane);
            PROF(0, bar_wait(&acc1_full[b], (a1ph >> b) & 1, "acc1_full", j, seq));
            a1ph ^= 1u << b;
            tc_fence_after();
            const uint32_t t_acc = t_lane + TM_ACC1 + b * CH;
            uint32_t raw[64];
            {
                uint32_t (&lo)[32] = *reinterpret_cast<uint32_t (*)[32]>(&raw[0]);
                uint32_t (&hi)[32] = *reinterpret_cast<uint32_t (*)[32]>(&raw[32]);
                tmem_ld_32x32(t_acc + wg * 64, lo);
                tmem_ld_32x32(t_acc + wg * 64 + 32, hi);
            }
            tmem_ld_wait();
            tc_fence_before();
            asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");      // the quarter's two warps have read their columns
            tc_fence_after();
            staging_free();      // the previous chunk's bulk store has read the staging block
#pragma unroll
            for (int sl = 0; sl < 4; ++sl) {
                uint32_t w[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int src = sl * 8 + i * 2;        // lanes holding bias[16 sl + 4 i .. +3] (two values each)
                    const float bx = __shfl_sync(0xffffffffu, bl.x, src), by = __shfl_sync(0xffffffffu, bl.y, src);
                    const float bz = __shfl_sync(0xffffffffu, bl.x, src + 1), bw = __shfl_sync(0xffffffffu, bl.y, src + 1);
                    const float v0 = mish_fast(__uint_as_float(raw[16 * sl + 4 * i + 0]) + bx);
                    const float v1 = mish_fast(__uint_as_float(raw[16 * sl + 4 * i + 1]) + by);
                    const float v2 = mish_fast(__uint_as_float(raw[16 * sl + 4 * i + 2]) + bz);
                    const float v3 = mish_fast(__uint_as_float(raw[16 * sl + 4 * i + 3]) + bw);
                    w[2 * i] = pack_bf16(v0, v1);
                    w[2 * i + 1] = pack_bf16(v2, v3);
                }
                tmem_st_32x8(t_acc + wg * 32 + sl * 8, w);
                sts128(stg_row + (((sl * 2 + 0) ^ sw) << 4), make_uint4(w[0], w[1], w[2], w[3]));
                sts128(stg_row + (((sl * 2 + 1) ^ sw) << 4), make_uint4(w[4], w[5], w[6], w[7]));
            }
            // this warp's 32 rows x 64 units (one k-block of the chunk) are staged: the exporter hands them to the partner
            __syncwarp();
            if (lane == 0) mbar_arrive(&stg_full[ew]);
            pending_free = true;
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(b ? ufull1 : ufull0);           // own GEMM 2 of this chunk may start
        };
        for (int t = group; t < n_tiles; t += n_groups, seq0 += nch) {
#pragma unroll 1
            for (int j = 0; j < nch; ++j) PROF(4, mish_chunk(j, seq0 + j));
            staging_free();
            // ---- residual epilogue: x[:, side*256 + ...] <- x + out + b2 for this warp's 32 rows x 128 columns, 4 slabs
            // of 32 columns: TMEM -> registers (thread = row) -> swizzled staging block -> lane = (row % 4, 16-byte chunk):
            // every global access of the warp is 4 full 128-byte row segments.  The residual rows of slab s+1 are requested
            // before slab s is touched (two register buffers), so their L2 latency hides behind a whole slab.
            const long long row0 = static_cast<long long>(t) * TILE + rank * ROWS + q * 32;
            const int col0 = side * NOUT + wg * 128;
            const int sub_row = lane >> 3, chunk = lane & 7;
            const int rows_valid = n_rows - static_cast<int>(row0) - sub_row;   // row 4i+sub_row live iff 4i < rows_valid
            float4 res[2][8], b4[2];
            auto load_res = [&](int sl) {
                const float* xp = p.x + (row0 + sub_row) * DM + col0 + sl * 32 + chunk * 4;
                b4[sl & 1] = __ldg(reinterpret_cast<const float4*>(p.b2 + col0 + sl * 32 + chunk * 4));
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    res[sl & 1][i] = 4 * i < rows_valid ? ldg128(xp + static_cast<long long>(4 * i) * DM) : make_float4(0.f, 0.f, 0.f, 0.f);
            };
            load_res(0);
            PROF(5, bar_wait(out_full, tphase, "out_full", t));
            tphase ^= 1;
            tc_fence_after();
#ifdef OFX_DEBUG
            const long long drain_t0 = clock64();
#endif
#pragma unroll
            for (int sl = 0; sl < 4; ++sl) {
                if (sl < 3) load_res(sl + 1);
                float* xp = p.x + (row0 + sub_row) * DM + col0 + sl * 32 + chunk * 4;
                uint32_t raw[32];
                tmem_ld_32x32(t_lane + TM_OUT + wg * 128 + sl * 32, raw);
                tmem_ld_wait();
                __syncwarp();        // the previous slab's phase B has finished reading the staging block
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    sts128(stg_row + ((i ^ sw) << 4), make_uint4(raw[4 * i], raw[4 * i + 1], raw[4 * i + 2], raw[4 * i + 3]));
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = 4 * i + sub_row;
                    const uint4 u = lds128(stg_u + r * 128 + ((chunk ^ (r & 7)) << 4));
                    float4 v = res[sl & 1][i];
                    const float4 bb = b4[sl & 1];
                    v.x += __uint_as_float(u.x) + bb.x; v.y += __uint_as_float(u.y) + bb.y;
                    v.z += __uint_as_float(u.z) + bb.z; v.w += __uint_as_float(u.w) + bb.w;
                    if (4 * i < rows_valid) stg128(xp + static_cast<long long>(4 * i) * DM, v);
                }
            }
            tc_fence_before();
            __syncwarp();        // the lanes' x stores happen-before lane 0's release below (cumulative at gpu scope)
            if (lane == 0) {
                mbar_arrive_cluster(oempty);
                if (p.h_next) red_release_gpu_add(xflag + side, 1);
            }
#ifdef OFX_DEBUG
            prof_acc[6] += clock64() - drain_t0;
#endif
        }
        if (q == 0) PROF_DUMP(16 + 8 * wg);
    } else {
        // ------------------------------------------------------------ LayerNorm warps (128 registers, no setmaxnreg)
        // (1) LN2 of the NEXT tile's x rows -> hbuf (bf16, row-major, global / L2): one tile ahead of its use, so the
        //     128 KB operand tile arrives in H by one bulk load the moment GEMM 1 of the running tile has retired
        //     (computing it in place after h_empty left the tensor pipe idle for ~20k cycles per tile);
        // (2) LN1 of the NEXT layer on the PREVIOUS tile's new rows -> h_next: the two pairs wrote 256 columns each,
        //     so the rows are complete once both sides' 8 epilogue warps have counted in; side s normalises rows
        //     s*64 .. +63 of each CTA's 128.
        const int lw = warp - LN_WARP0;
        int tiles_done = 0;
        PROF_DECL();
        auto load_affine = [&](const float* gw, const float* gb, float4 (&g)[4], float4 (&be)[4]) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                g[i] = __ldg(reinterpret_cast<const float4*>(gw + i * 128 + lane * 4));
                be[i] = __ldg(reinterpret_cast<const float4*>(gb + i * 128 + lane * 4));
            }
        };
        // two-pass LayerNorm of 4 rows held in registers (lane owns columns i*128 + lane*4 .. +3)
        auto norm4 = [&](float4 (&v)[4][4], const float4 (&gam)[4], const float4 (&bet)[4], auto&& emit) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float s = 0.f;
#pragma unroll
                for (int i = 0; i < 4; ++i) s += v[u][i].x + v[u][i].y + v[u][i].z + v[u][i].w;
                const float mu = warp_sum(s) * (1.f / DM);
                float qq = 0.f;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float a = v[u][i].x - mu, b = v[u][i].y - mu, c = v[u][i].z - mu, d = v[u][i].w - mu;
                    qq += a * a + b * b + c * c + d * d;
                }
                const float rstd = rsqrtf(warp_sum(qq) * (1.f / DM) + 1e-5f);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 g = gam[i], be = bet[i];
                    emit(u, i, (v[u][i].x - mu) * rstd * g.x + be.x, (v[u][i].y - mu) * rstd * g.y + be.y,
                         (v[u][i].z - mu) * rstd * g.z + be.z, (v[u][i].w - mu) * rstd * g.w + be.w);
                }
            }
        };
        auto load4 = [&](float4 (&v)[4][4], long long row) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    v[u][i] = row + u < n_rows ? ldg128(p.x + (row + u) * DM + i * 128 + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        // LN2 of tile t's rows (this warp's 32 of the CTA's 128) -> hbuf[parity]
        auto stage = [&](int t, int parity) {
            float4 gam[4], bet[4];
            load_affine(p.ln_w, p.ln_b, gam, bet);
            const long long row0 = static_cast<long long>(t) * TILE + rank * ROWS + lw * 32;
            __nv_bfloat16* dst = p.hbuf + (static_cast<long long>(blockIdx.x) * 2 + parity) * ROWS * DM + static_cast<long long>(lw) * 32 * DM;
#pragma unroll 1
            for (int rb = 0; rb < 32; rb += 4) {
                float4 v[4][4];
                load4(v, row0 + rb);
                norm4(v, gam, bet, [&](int u, int i, float y0, float y1, float y2, float y3) {
                    if (row0 + rb + u >= n_rows) y0 = y1 = y2 = y3 = 0.f;
                    *reinterpret_cast<uint2*>(dst + (rb + u) * DM + i * 128 + lane * 4) = make_uint2(pack_bf16(y0, y1), pack_bf16(y2, y3));
                });
            }
            fence_proxy_async_all();      // generic-proxy global writes -> the producer's bulk loads
            __syncwarp();
            if (lane == 0) mbar_arrive(h_staged);
        };
        auto ln_next = [&](int t) {      // LN1(next layer) of tile t's new rows -> h_next
            float4 gam[4], bet[4];
            load_affine(p.lnn_w, p.lnn_b, gam, bet);
            if (lane == 0) {     // both pairs' epilogue warps have stored their column halves of tile t
                flag_wait(xflag, N_EPI_WARPS * tiles_done, "xflag0", t);
                flag_wait(xflag + 1, N_EPI_WARPS * tiles_done, "xflag1", t);
            }
            __syncwarp();
            const long long row0 = static_cast<long long>(t) * TILE + rank * ROWS + side * 64 + lw * 16;
#pragma unroll 1
            for (int rb = 0; rb < 16; rb += 4) {
                if (row0 + rb >= n_rows) break;
                float4 v[4][4];
                load4(v, row0 + rb);
                norm4(v, gam, bet, [&](int u, int i, float y0, float y1, float y2, float y3) {
                    const long long row = row0 + rb + u;
                    if (row < n_rows)
                        *reinterpret_cast<uint2*>(p.h_next + row * DM + i * 128 + lane * 4) =
                            make_uint2(pack_bf16(y0, y1), pack_bf16(y2, y3));
                });
            }
        };
        int ti = 0;
        if (group < n_tiles) PROF(1, stage(group, 0));
        for (int t = group; t < n_tiles; t += n_groups, ++ti) {
            // the next tile goes into the other half of hbuf; its previous tenant (tile ti - 1) has been pulled into H
            // and consumed once the producer has started on tile ti (hbuf_free: one arrival per use of a half, so the
            // wait can never fall two phases behind)
            if (t + n_groups < n_tiles) {
                if (ti >= 1) PROF(0, bar_wait(&hbuf_free[(ti + 1) & 1], ((ti - 1) >> 1) & 1, "hbuf_free", t));
                PROF(1, stage(t + n_groups, (ti + 1) & 1));
            }
            if (p.h_next && ti >= 1) { ++tiles_done; PROF(2, ln_next(t - n_groups)); }
        }
        if (p.h_next && ti >= 1) { ++tiles_done; PROF(2, ln_next(group + (ti - 1) * n_groups)); }
        if (lw == 0 && rank == 1) PROF_DUMP(8);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 512);
    }
}

}  // namespace ffn2

// scratch: hidden-chunk ring (bf16) + the three counter arrays
static size_t ffn2_scratch_bytes(int n_groups) {
    return static_cast<size_t>(n_groups) * 4 * ffn2::NSLOT * ffn2::ROWS * ffn2::CH * 2;
}
constexpr int kFlagsPerGroup = 4 * ffn2::NSLOT + 2 + 4;
static size_t ffn2_flag_bytes(int n_groups) { return align_up(static_cast<size_t>(n_groups) * kFlagsPerGroup * sizeof(int), 256); }

static size_t ffn2_hbuf_bytes(int n_groups) { return static_cast<size_t>(n_groups) * 4 * 2 * ffn2::ROWS * ffn2::DM * 2; }

size_t ffn_block2_workspace_bytes(int sm) {
    const int n_groups = sm / 4;
    return align_up(ffn2_scratch_bytes(n_groups), 256) + ffn2_flag_bytes(n_groups) + align_up(ffn2_hbuf_bytes(n_groups), 256);
}

bool ffn_block2_supported(int dm, int fp) {
    return dm == ffn2::DM && fp > 0 && fp % 256 == 0 && fp / 256 * 2 <= ffn2::NSLOT && sm_count() >= 4;
}

int ffn_block2_bf16(const FfnBlockArgs& a, cudaStream_t stream) {
    using namespace ffn2;
    if (a.rows <= 0) return OFX_OK;
    if (!ffn_block2_supported(a.dm, a.fp))
        return fail(OFX_E_SHAPE, "ffn_block: needs d_model 512 and padded d_ffn %% 256 == 0 (got %d, %d)", a.dm, a.fp);
    if (a.h_next && (!a.lnn_w || !a.lnn_b)) return fail(OFX_E_ARG, "ffn_block: h_next needs the next layer's LayerNorm terms");
    const int n_groups_max = sm_count() / 4;
    const size_t need = ffn_block2_workspace_bytes(sm_count());
    if (!a.workspace || a.workspace_bytes < need)
        return fail(OFX_E_WORKSPACE, "ffn_block workspace %zu B < required %zu B", a.workspace_bytes, need);
    if (reinterpret_cast<uintptr_t>(a.workspace) % 256) return fail(OFX_E_ARG, "ffn_block: workspace must be 256-byte aligned");
    uint8_t* ws = static_cast<uint8_t*>(a.workspace);
    int* flags = reinterpret_cast<int*>(ws + align_up(ffn2_scratch_bytes(n_groups_max), 256));
    CUtensorMap tm_w1, tm_w2, tm_ust, tm_uld;
    OFX_TRY(make_tmap_bf16(&tm_w1, a.w1, static_cast<uint64_t>(a.fp), DM, DM, 64));
    OFX_TRY(make_tmap_bf16(&tm_w2, a.w2, DM, static_cast<uint64_t>(a.fp), a.fp, 128));
    const uint64_t srows = static_cast<uint64_t>(n_groups_max) * 4 * NSLOT * ROWS;
    OFX_TRY(make_tmap_bf16(&tm_ust, ws, srows, CH, CH, 32));
    OFX_TRY(make_tmap_bf16(&tm_uld, ws, srows, CH, CH, 128));
    uint8_t* hbuf = ws + align_up(ffn2_scratch_bytes(n_groups_max), 256) + ffn2_flag_bytes(n_groups_max);
    CUtensorMap tm_h;
    OFX_TRY(make_tmap_bf16(&tm_h, hbuf, static_cast<uint64_t>(n_groups_max) * 4 * 2 * ROWS, DM, DM, 128));
    static DeviceOnce configured;
    if (configured.need())
        OFX_CUDA(cudaFuncSetAttribute(ffn_block2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    const int n_tiles = (a.rows + TILE - 1) / TILE;
    const int groups = n_tiles < n_groups_max ? n_tiles : n_groups_max;
    OFX_CUDA(cudaMemsetAsync(flags, 0, ffn2_flag_bytes(n_groups_max), stream));
    Params p{a.x, a.rows, a.rows_dev, a.ln_w, a.ln_b, a.b1, a.b2, a.fp / 256,
             static_cast<__nv_bfloat16*>(a.h_next), a.lnn_w, a.lnn_b,
             flags, flags + n_groups_max * 4 * NSLOT, flags + n_groups_max * (4 * NSLOT + 2),
             reinterpret_cast<__nv_bfloat16*>(ws), reinterpret_cast<__nv_bfloat16*>(hbuf), nullptr, 0};
#ifdef OFX_DEBUG
    { static int mode = -1; if (mode < 0) { const char* e = getenv("OFX_FFN2_MODE"); mode = e ? atoi(e) : 0; } p.mode = mode; }
#endif
#ifdef OFX_DEBUG   // instrumented builds only: per-role wait-cycle counters, dumped synchronously (OFX_FFN_PROF=1)
    static int prof_on = -1;
    static long long* prof_dev = nullptr;
    if (prof_on < 0) { const char* e = getenv("OFX_FFN_PROF"); prof_on = (e && e[0] == '1') ? 1 : 0; }
    if (prof_on) {
        if (!prof_dev) OFX_CUDA(cudaMalloc(&prof_dev, 8 * 32 * 160));
        OFX_CUDA(cudaMemsetAsync(prof_dev, 0, 8 * 32 * 160, stream));
        p.prof = prof_dev;
    }
#endif
    ffn_block2_kernel<<<groups * 4, NTHREADS, SMEM_BYTES, stream>>>(tm_w1, tm_w2, tm_ust, tm_uld, tm_h, p);
    OFX_LAUNCH_CHECK();
#ifdef OFX_DEBUG
    if (prof_on) {
        static int dumps = 0;
        static long long h[32 * 160];
        OFX_CUDA(cudaStreamSynchronize(stream));
        OFX_CUDA(cudaMemcpy(h, prof_dev, sizeof(h), cudaMemcpyDeviceToHost));
        if (dumps++ < 3)
            for (int c = 0; c < groups * 4 && c < 148; c += 36) {
                const long long* o = h + c * 32;
                fprintf(stderr, "ffn2 prof cta %3d: producer total %lld  w_empty %lld  peer_flag %lld | %s total %lld  a %lld  b %lld  c %lld  d %lld\n",
                        c, o[0], o[1], o[2], (c & 1) ? "LN warp: (h_empty, prologue, ln_next)" : "MMA: (w_full, u_full, h_full, out_empty)",
                        o[8], o[9], o[10], o[11], o[12]);
                for (int w = 0; w < 2; ++w)
                    fprintf(stderr, "   epilogue wg %d: total %lld  acc1_full %lld  stg_read %lld  uack %lld  publish %lld  mish_chunks(all) %lld  out_full %lld  drain %lld\n",
                            w, o[16 + 8 * w], o[17 + 8 * w], o[18 + 8 * w], o[19 + 8 * w], o[20 + 8 * w], o[21 + 8 * w], o[22 + 8 * w], o[23 + 8 * w]);
            }
    }
#endif
    return OFX_OK;
}

}  // namespace ofx
