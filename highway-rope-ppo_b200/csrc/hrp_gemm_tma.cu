// hrp_gemm_tma.cu -- TMA-fed tcgen05 GEMM (3xTF32) for the policy / value MLP of ppo/agent.py:12-84, 218-252.
//
//   C[M, N] = sum_k A(m, k) * B(n, k)   (+ bias[n], ReLU, ReLU mask), optionally also C_lo = C - trunc_tf32(C)
//
// What is different from hrp_mlp_tc.cu (which stays as the fallback for operands a tensor map cannot describe):
//   * operands arrive by TMA (cp.async.bulk.tensor, SWIZZLE_128B, one elected thread), not through registers: no loader
//     warps, no generic-proxy stores into the operand stages;
//   * the 3xTF32 split is NOT done in the kernel.  Every operand exists in memory twice: the fp32 tensor itself -- the
//     tensor core truncates fp32 to TF32, so it IS the "hi" operand -- and a "lo" tensor x - trunc_tf32(x), written by
//     whoever produced x (this kernel's epilogue for activations, prepare_weights / split_lo for the rest).  A stage
//     is four TMA boxes (A, A_lo, B, B_lo); shared memory sees each operand byte written once and read by the MMAs,
//     which is what bounds the main loop (128 B/cycle/SM);
//   * both operand majors: K-major (row-major [rows, K], the forward and input-gradient GEMMs) and MN-major (row-major
//     [K, rows]: the weight-gradient GEMMs contract over the batch, and the input-gradient GEMMs read W[out, in]
//     as the MN-major B operand instead of a transposed copy);
//   * warp roles: warp 0 = TMA producer, warp 1 = TMEM allocation + MMA issue, warps 2-5 = epilogue (TMEM -> registers
//     -> global, 64 contiguous bytes per thread and column chunk).
// Accumulation: D[:, 0:BN] += A_hi [B_hi; B_lo] (one UMMA with N = 2 BN over the stacked B tile) and D[:, 0:BN] +=
// A_lo B_hi; the two halves are added in the epilogue.
#include <cuda.h>
#include <stdlib.h>

#include "hrp_internal.cuh"

// phase clocks of CTA (0,0,0) when hrp_debug_tma_gemm_clocks(1) switched them on (tools/gemm_bench.py)
__device__ long long g_gt_phase[16];
__device__ int g_gt_phase_on;
// (compiled in only by a profiling build: HRP_PHASE_CLOCKS=1 python highway-rope-ppo_b200/build.py --force)
#ifdef HRP_PHASE_CLOCKS
#define GT_PHASE(i) do { if (g_gt_phase_on && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) g_gt_phase[(i)] = clock64(); } while (0)
#else
#define GT_PHASE(i) do { } while (0)
#endif

namespace {

// K-block of 32 fp32 = 128 bytes per operand row (SWIZZLE_128B for K-major tiles); a stage of a 128 x 64 tile is
// 48 KB, four stages 192 KB: ONE CTA per SM.  Measured alternatives (profiles/r02_gemm_bench.txt): 16-wide K-blocks
// (SWIZZLE_64B, 24 KB stages, two CTAs per SM) run the same tile in 13.7 us instead of 9.2 us (twice the boxes and
// barrier round trips per byte); two 48 KB stages cannot cover the ~1300-cycle TMA latency with a ~700-cycle K-block.
constexpr int BM = 128, BK = 32;
constexpr int GT_THREADS = 192;             // 6 warps
constexpr int A_TILE = BM * BK * 4;         // 16 KB

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "GT_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra GT_DONE;\n\t"
        "bra GT_WAIT;\n\t"
        "GT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap *map, void *dst, uint64_t *bar, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *map, const void *src, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(map)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor), version 1.
//   K-major  tile [rows][32 fp32], SWIZZLE_128B (layout type 2): 8-row x 128-byte atoms 1024 B apart (SBO); LBO unused.
//   MN-major tile: 32-bit operands have ONE legal MN-major layout, SWIZZLE_128B_BASE32B (layout type 1: the 128-byte
//   swizzle with 32-byte atoms, Swizzle<2,5,2>; TMA's CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B), whose atom is 4 k-rows x
//   128 bytes (32 mn values).  The tile is stored as boxes of [32 k][32 mn fp32] = 4 KB each: the next 32 mn values
//   are one box further (LBO = 4096 B), the next 4 k are one atom further (SBO = 512 B); one UMMA (8 k) spans two atoms.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, bool mn_major)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)(mn_major ? ((BK * 128) >> 4) : 1) << 16;
    d |= (uint64_t)((mn_major ? 512 : 1024) >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(mn_major ? 1 : 2) << 61;
    return d;
}
// cute::UMMA::InstrDescriptor: D = F32 (bits 4-5 = 1), A = B = TF32 (2 at bits 7-9, 10-12), a_major bit 15, b_major bit 16
// (0 = K-major, 1 = MN-major), N >> 3 at bits 17-22, M >> 4 at bits 24-28
__host__ __device__ constexpr uint32_t make_idesc(int m, int n, bool a_mn, bool b_mn)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

struct GemmMaps {
    CUtensorMap a, a_lo, b, b_lo;
    CUtensorMap c, c_lo;   // output [splits][M][N] as boxes of [128 rows][32 columns], SWIZZLE_128B (TMA store)
};

template <int BN> __host__ __device__ constexpr int gt_stages() { return BN == 64 ? 4 : 3; }   // 192 KB either way
template <int BN> __host__ __device__ constexpr int gt_stage_bytes() { return 2 * (A_TILE + BN * BK * 4); }

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(GT_THREADS, 1)
tma_gemm_kernel(const __grid_constant__ GemmMaps maps, int M, int N, int K, int k_chunk, float *__restrict__ C,
                float *__restrict__ C_lo, int ldc, const float *__restrict__ bias, int relu,
                const float *__restrict__ mask, int ldm, int tma_store)
{
    constexpr int B_TILE = BN * BK * 4;
    constexpr int STAGE = gt_stage_bytes<BN>();
    constexpr int STAGES = gt_stages<BN>();
    constexpr int TM_COLS = 2 * BN;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t bar_full[4], bar_empty[4], bar_done;
    __shared__ uint32_t tmem_base_s;
    __shared__ float bias_s[BN];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    hrp_pdl_release();
    if (tid == 0) GT_PHASE(0);
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kbeg = blockIdx.z * k_chunk, kend = min(K, kbeg + k_chunk);
    const int nkb = (kend - kbeg + BK - 1) / BK;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&bar_full[s], 1);    // the producer's arrive.expect_tx; the four boxes complete the transaction count
            mbar_init(&bar_empty[s], 1);   // tcgen05.commit arrives when the MMAs have read the stage
        }
        mbar_init(&bar_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.a_lo)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.b)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.b_lo)) : "memory");
        if (tma_store) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.c)) : "memory");
            if (C_lo) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.c_lo)) : "memory");
        }
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "n"(TM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_s;
    if (tid == 0) GT_PHASE(1);   // set up (barriers, TMEM)
    hrp_pdl_wait();   // everything above overlapped the previous kernel's tail; its results are visible from here
    if (tid == 0) GT_PHASE(2);   // predecessor complete

    if (warp == 0) {
        // ===== TMA producer: one thread; a stage = A, A_lo, B, B_lo boxes of one K-block
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % STAGES, k0 = kbeg + kb * BK;
                if (kb >= STAGES) mbar_wait(&bar_empty[s], (uint32_t)(((kb / STAGES) - 1) & 1));
                uint8_t *st = smem + s * STAGE;
                mbar_expect_tx(&bar_full[s], (uint32_t)STAGE);
                if (A_MN) {   // boxes of [BK k][32 m]: coordinates (m, k)
#pragma unroll
                    for (int j = 0; j < BM / 32; ++j) {
                        tma_load_2d(&maps.a, st + j * (BK * 128), &bar_full[s], m0 + 32 * j, k0);
                        tma_load_2d(&maps.a_lo, st + A_TILE + j * (BK * 128), &bar_full[s], m0 + 32 * j, k0);
                    }
                } else {      // one box [128 m][BK k]: coordinates (k, m)
                    tma_load_2d(&maps.a, st, &bar_full[s], k0, m0);
                    tma_load_2d(&maps.a_lo, st + A_TILE, &bar_full[s], k0, m0);
                }
                uint8_t *sb = st + 2 * A_TILE;
                if (B_MN) {
#pragma unroll
                    for (int j = 0; j < BN / 32; ++j) {
                        tma_load_2d(&maps.b, sb + j * (BK * 128), &bar_full[s], n0 + 32 * j, k0);
                        tma_load_2d(&maps.b_lo, sb + B_TILE + j * (BK * 128), &bar_full[s], n0 + 32 * j, k0);
                    }
                } else {
                    tma_load_2d(&maps.b, sb, &bar_full[s], k0, n0);
                    tma_load_2d(&maps.b_lo, sb + B_TILE, &bar_full[s], k0, n0);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: waits for a full stage, one elected lane issues the UMMAs, tcgen05.commit frees the stage
        constexpr uint32_t idesc = make_idesc(BM, BN, A_MN, B_MN);
        constexpr uint32_t idesc2 = make_idesc(BM, 2 * BN, A_MN, B_MN);   // A_hi against the stacked [B_hi ; B_lo] tile
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % STAGES;
            mbar_wait(&bar_full[s], (uint32_t)((kb / STAGES) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0 && kb == 0) GT_PHASE(3);   // first stage landed
            if (lane == 0 && kb == 1) GT_PHASE(4);   // second stage landed (or MMAs of the first issued)
            if (lane == 0) {
                const uint32_t a_s = smem_u32(smem + s * STAGE), b_s = a_s + 2 * A_TILE;
                const uint64_t da_hi = make_desc(a_s, A_MN), db_hi = make_desc(b_s, B_MN);
                const uint64_t da_lo = da_hi + (A_TILE >> 4);
                // one UMMA consumes 8 k: K-major operands advance 32 bytes inside the swizzle row, MN-major operands one
                // 8-row atom (1024 bytes)
                constexpr uint64_t a_step = A_MN ? (1024 >> 4) : (32 >> 4), b_step = B_MN ? (1024 >> 4) : (32 >> 4);
#pragma unroll
                for (int kk = 0; kk < BK / 8; ++kk) {
                    const uint32_t acc = (kb > 0 || kk > 0) ? 1u : 0u;
                    umma_tf32(tmem_d, da_hi + kk * a_step, db_hi + kk * b_step, idesc2, acc);   // [A_hi B_hi | A_hi B_lo]
                    umma_tf32(tmem_d, da_lo + kk * a_step, db_hi + kk * b_step, idesc, 1u);     // + A_lo B_hi
                }
                umma_commit(&bar_empty[s]);
                if (kb + 1 == nkb) { umma_commit(&bar_done); GT_PHASE(5); }   // last MMA issued
            }
            __syncwarp();
        }
    } else {
        // ===== epilogue warps 2..5: TMEM lanes 32 (warp % 4) .., one accumulator row per thread.
        // While the main loop runs they fetch what the epilogue needs from global memory: the bias slice into shared
        // memory and (64-wide tiles) this thread's row of the ReLU mask into registers.
        {
            const int i = tid - 64;
            if (i < BN) bias_s[i] = (bias && n0 + i < N) ? __ldg(bias + n0 + i) : 0.f;
        }
        constexpr int MASK_REGS = BN == 64 ? 16 : 1;
        float4 mk[MASK_REGS];
        const bool mask_pre = BN == 64 && mask != nullptr && ldm % 4 == 0 && ((uintptr_t)mask & 15) == 0 && n0 + BN <= N;
        if (mask_pre) {
            const int r = m0 + 32 * (warp & 3) + lane;
            const float4 *mrow = reinterpret_cast<const float4 *>(mask + (size_t)min(r, M - 1) * ldm + n0);
#pragma unroll
            for (int j = 0; j < MASK_REGS; ++j) mk[j] = __ldg(mrow + j);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");   // bias_s is complete
        if (nkb > 0) mbar_wait(&bar_done, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (tid == 64) GT_PHASE(6);   // accumulator complete
        const int q = warp & 3, row = 32 * q + lane, gm = m0 + row;
        float *Cz = C + (size_t)blockIdx.z * M * ldc;
        float *Clz = C_lo ? C_lo + (size_t)blockIdx.z * M * ldc : nullptr;
        const bool vec_ok = ldc % 4 == 0 && ((uintptr_t)Cz & 15) == 0 && (!Clz || ((uintptr_t)Clz & 15) == 0) &&
                            (!mask || (ldm % 4 == 0 && ((uintptr_t)mask & 15) == 0));
#pragma unroll
        for (int c0 = 0; c0 < BN; c0 += 16) {
            uint32_t v[16], u[16];
            if (nkb > 0) {
                const uint32_t taddr = tmem_d + ((uint32_t)(32 * q) << 16) + (uint32_t)c0;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
                    "%15}, [%16];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                    : "r"(taddr));
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
                    "%15}, [%16];"
                    : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                      "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
                    : "r"(taddr + (uint32_t)BN));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = u[j] = 0u;
            }
            const int gn0 = n0 + c0;
            if (tma_store) {
                // stage the tile (and its lo part) in shared memory -- the operand stages are free once bar_done has
                // fired -- as [128 rows][32 columns] panels in the SWIZZLE_128B layout (16-byte chunk ^ row % 8: the
                // 32 rows of a warp spread over the banks), and let ONE TMA store per panel write it out: full
                // 128-byte rows, edges clipped by the tensor map
                float x[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    x[j] = __uint_as_float(v[j]) + __uint_as_float(u[j]) + bias_s[c0 + j];
                    if (relu) x[j] = fmaxf(x[j], 0.f);
                }
                if (mask_pre) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        const float4 m4 = mk[BN == 64 ? (c0 + j) / 4 : 0];
                        x[j] = m4.x > 0.f ? x[j] : 0.f; x[j + 1] = m4.y > 0.f ? x[j + 1] : 0.f;
                        x[j + 2] = m4.z > 0.f ? x[j + 2] : 0.f; x[j + 3] = m4.w > 0.f ? x[j + 3] : 0.f;
                    }
                } else if (mask && gm < M) {
                    const float *mrow = mask + (size_t)gm * ldm + gn0;
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (gn0 + j < N) x[j] = __ldg(mrow + j) > 0.f ? x[j] : 0.f;
                }
                uint8_t *panel = smem + (c0 >> 5) * (BM * 128) + row * 128;
                uint8_t *panel_lo = panel + BN * BM * 4;
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    const int ch = (((c0 & 31) + j) >> 2) ^ (row & 7);
                    *reinterpret_cast<float4 *>(panel + ch * 16) = make_float4(x[j], x[j + 1], x[j + 2], x[j + 3]);
                    if (C_lo) {
                        float4 l;
                        l.x = x[j] - __uint_as_float(__float_as_uint(x[j]) & 0xFFFFE000u);
                        l.y = x[j + 1] - __uint_as_float(__float_as_uint(x[j + 1]) & 0xFFFFE000u);
                        l.z = x[j + 2] - __uint_as_float(__float_as_uint(x[j + 2]) & 0xFFFFE000u);
                        l.w = x[j + 3] - __uint_as_float(__float_as_uint(x[j + 3]) & 0xFFFFE000u);
                        *reinterpret_cast<float4 *>(panel_lo + ch * 16) = l;
                    }
                }
                continue;
            }
            if (gm < M && gn0 < N) {
                float x[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    x[j] = __uint_as_float(v[j]) + __uint_as_float(u[j]) + bias_s[c0 + j];
                    if (relu) x[j] = fmaxf(x[j], 0.f);
                }
                float *crow = Cz + (size_t)gm * ldc + gn0;
                const float *mrow = mask ? mask + (size_t)gm * ldm + gn0 : nullptr;
                if (vec_ok && gn0 + 16 <= N) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        if (mrow) {
                            const float4 mk = __ldg(reinterpret_cast<const float4 *>(mrow + j));
                            x[j] = mk.x > 0.f ? x[j] : 0.f; x[j + 1] = mk.y > 0.f ? x[j + 1] : 0.f;
                            x[j + 2] = mk.z > 0.f ? x[j + 2] : 0.f; x[j + 3] = mk.w > 0.f ? x[j + 3] : 0.f;
                        }
                        *reinterpret_cast<float4 *>(crow + j) = make_float4(x[j], x[j + 1], x[j + 2], x[j + 3]);
                        if (Clz) {
                            float4 l;
                            l.x = x[j] - __uint_as_float(__float_as_uint(x[j]) & 0xFFFFE000u);
                            l.y = x[j + 1] - __uint_as_float(__float_as_uint(x[j + 1]) & 0xFFFFE000u);
                            l.z = x[j + 2] - __uint_as_float(__float_as_uint(x[j + 2]) & 0xFFFFE000u);
                            l.w = x[j + 3] - __uint_as_float(__float_as_uint(x[j + 3]) & 0xFFFFE000u);
                            *reinterpret_cast<float4 *>(Clz + (size_t)gm * ldc + gn0 + j) = l;
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        if (gn0 + j < N) {
                            float y = x[j];
                            if (mrow) y = __ldg(mrow + j) > 0.f ? y : 0.f;
                            crow[j] = y;
                            if (Clz) Clz[(size_t)gm * ldc + gn0 + j] = y - __uint_as_float(__float_as_uint(y) & 0xFFFFE000u);
                        }
                    }
                }
            }
        }
        if (tid == 64) GT_PHASE(7);   // tile staged / written
        if (tma_store) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> bulk-copy reads
            asm volatile("bar.sync 1, 128;" ::: "memory");                     // the four epilogue warps
            if (warp == 2 && lane == 0) {
#pragma unroll
                for (int p = 0; p < BN / 32; ++p) {
                    if (n0 + 32 * p < N) {
                        tma_store_3d(&maps.c, smem + p * (BM * 128), n0 + 32 * p, m0, (int)blockIdx.z);
                        if (C_lo) tma_store_3d(&maps.c_lo, smem + BN * BM * 4 + p * (BM * 128), n0 + 32 * p, m0, (int)blockIdx.z);
                    }
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // shared memory is read before the CTA goes
                GT_PHASE(8);   // stores read out of shared memory
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(TM_COLS) : "memory");
}

// x_lo = x - trunc_tf32(x) over n floats (external inputs: the caller's states, the weights)
__global__ void __launch_bounds__(256) split_lo_kernel(const float *__restrict__ x, float *__restrict__ lo, long long n)
{
    hrp_pdl_release();
    hrp_pdl_wait();
    long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i + 3 < n && (((uintptr_t)x | (uintptr_t)lo) & 15) == 0) {
        const float4 v = *reinterpret_cast<const float4 *>(x + i);
        float4 l;
        l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
        l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
        l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
        l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
        *reinterpret_cast<float4 *>(lo + i) = l;
    } else {
        for (int j = 0; j < 4 && i + j < n; ++j) lo[i + j] = x[i + j] - __uint_as_float(__float_as_uint(x[i + j]) & 0xFFFFE000u);
    }
}

// ---- tensor maps ----------------------------------------------------------------------------------------------------
// An operand P(row, k) with element strides (srow, sk): K-major when sk == 1 (tensor map dims {K, rows}, box {32, box_rows}),
// MN-major when srow == 1 (dims {rows, K}, box {32, 32}).  TMA needs a 16-byte aligned base and pitch.
bool tma_ok(const float *p, long long srow, long long sk, int rows, int K)
{
    if (((uintptr_t)p & 15) != 0) return false;
    if (K < 8 || rows < 8) return false;
    if (sk == 1) return srow % 4 == 0 && srow >= K;
    if (srow == 1) return sk % 4 == 0 && sk >= rows;
    return false;
}
// cuTensorMapEncodeTiled is a driver entry point: it is looked up through the runtime (cudaGetDriverEntryPoint), so the
// library does not link against libcuda.so.1 and still loads on a machine without a driver (build check, CPU tests)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

int encode(CUtensorMap *map, const float *p, long long srow, long long sk, int rows, int K, int box_rows)
{
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) { hrp_set_error("cuTensorMapEncodeTiled is not available from this driver"); return -2; }
    const bool kmajor = sk == 1;
    cuuint64_t dims[2] = {(cuuint64_t)(kmajor ? K : rows), (cuuint64_t)(kmajor ? rows : K)};
    cuuint64_t strides[1] = {(cuuint64_t)((kmajor ? srow : sk) * 4)};
    cuuint32_t box[2] = {(cuuint32_t)(kmajor ? BK : 32), (cuuint32_t)(kmajor ? box_rows : BK)};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(p), dims, strides, box, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        kmajor ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        hrp_set_error("cuTensorMapEncodeTiled failed: CUresult %d (rows %d, K %d, strides %lld / %lld)", (int)r, rows, K, srow, sk);
        return -2;
    }
    return 0;
}

// output tensor [splits][M][N] (pitch ldc), stored as boxes of [128 rows][32 columns]
int encode_out(CUtensorMap *map, float *p, int M, int N, int ldc, int splits)
{
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) { hrp_set_error("cuTensorMapEncodeTiled is not available from this driver"); return -2; }
    cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)M, (cuuint64_t)splits};
    cuuint64_t strides[2] = {(cuuint64_t)ldc * 4, (cuuint64_t)M * ldc * 4};
    cuuint32_t box[3] = {32u, (cuuint32_t)BM, 1u};
    cuuint32_t estr[3] = {1u, 1u, 1u};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, p, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { hrp_set_error("cuTensorMapEncodeTiled (output) failed: CUresult %d (M %d, N %d, ldc %d)", (int)r, M, N, ldc); return -2; }
    return 0;
}

template <int BN, bool A_MN, bool B_MN>
int launch(const GemmMaps &maps, dim3 grid, cudaStream_t s, int M, int N, int K, int k_chunk, float *C, float *C_lo, int ldc,
           const float *bias, int relu, const float *mask, int ldm, int tma_store)
{
    constexpr int SMEM = gt_stages<BN>() * gt_stage_bytes<BN>() + 1024;
    static bool configured[HRP_MAX_DEVICES] = {false};
    auto kern = tma_gemm_kernel<BN, A_MN, B_MN>;
    int dev = 0;
    HRP_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= HRP_MAX_DEVICES) { hrp_set_error("device index %d not supported", dev); return -1; }
    if (!configured[dev]) {
        HRP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        configured[dev] = true;
    }
    HRP_CUDA_OK(hrp_launch_pdl(kern, grid, dim3(GT_THREADS), (size_t)SMEM, s, maps, M, N, K, k_chunk, C, C_lo, ldc, bias, relu,
                               mask, ldm, tma_store));
    return 0;
}

}  // namespace

// x_lo = x - trunc_tf32(x)
int hrp_split_lo(const float *x, float *lo, long long n, cudaStream_t s)
{
    unsigned grid = (unsigned)((n + 1023) / 1024);
    HRP_CUDA_OK(hrp_launch_pdl(split_lo_kernel, dim3(grid), dim3(256), 0, s, x, lo, n));
    return 0;
}

// whether hrp_tma_gemm can take these operands (alignment / stride rules of a tensor map)
bool hrp_tma_gemm_ok(int M, int N, int K, const float *A, long long sam, long long sak, const float *A_lo, const float *B,
                     long long sbn, long long sbk, const float *B_lo)
{
    return A_lo && B_lo && tma_ok(A, sam, sak, M, K) && tma_ok(A_lo, sam, sak, M, K) && tma_ok(B, sbn, sbk, N, K) &&
           tma_ok(B_lo, sbn, sbk, N, K);
}

// 3xTF32 GEMM on pre-split operands; returns the number of K splits used (partials at C + z*M*ldc), < 0 on error.
// bn_hint: 0 = choose, else 64 or 128.
int hrp_tma_gemm(int M, int N, int K, const float *A, const float *A_lo, long long sam, long long sak, const float *B,
                 const float *B_lo, long long sbn, long long sbk, float *C, float *C_lo, int ldc, const float *bias, int relu,
                 const float *mask, int ldm, int splits, int bn_hint, cudaStream_t s)
{
    int k_chunk = K;
    if (splits > 1) {
        k_chunk = ((K + splits - 1) / splits + BK - 1) / BK * BK;
        splits = (K + k_chunk - 1) / k_chunk;
    } else {
        splits = 1;
    }
    const int mt = (M + BM - 1) / BM;
    int bn = bn_hint;
    if (bn != 64 && bn != 128) bn = (N <= 64 || mt * ((N + 127) / 128) * splits < 100) ? 64 : 128;
    if (C_lo && bn == 128 && gt_stages<128>() * gt_stage_bytes<128>() < 2 * BM * 128 * 4) bn = 64;   // staging must fit
    const bool a_mn = sak != 1, b_mn = sbk != 1;
    GemmMaps maps;
    if (encode(&maps.a, A, sam, sak, M, K, BM) || encode(&maps.a_lo, A_lo, sam, sak, M, K, BM) ||
        encode(&maps.b, B, sbn, sbk, N, K, bn) || encode(&maps.b_lo, B_lo, sbn, sbk, N, K, bn))
        return -2;
    dim3 grid((N + bn - 1) / bn, mt, splits);
    int rc;
    // output through TMA stores when a tensor map can describe it (16-byte aligned base and pitch)
    int tma_store = ldc % 4 == 0 && ((uintptr_t)C & 15) == 0 && (!C_lo || ((uintptr_t)C_lo & 15) == 0) && !getenv("HRP_NO_TMA_STORE");
    if (tma_store) {
        if (encode_out(&maps.c, C, M, N, ldc, splits) || (C_lo && encode_out(&maps.c_lo, C_lo, M, N, ldc, splits))) return -2;
    }
#define HRP_GT_GO(BN_, AM, BMj) launch<BN_, AM, BMj>(maps, grid, s, M, N, K, k_chunk, C, C_lo, ldc, bias, relu, mask, ldm, tma_store)
    if (bn == 64) {
        if (!a_mn && !b_mn) rc = HRP_GT_GO(64, false, false);
        else if (!a_mn && b_mn) rc = HRP_GT_GO(64, false, true);
        else if (a_mn && !b_mn) rc = HRP_GT_GO(64, true, false);
        else rc = HRP_GT_GO(64, true, true);
    } else {
        if (!a_mn && !b_mn) rc = HRP_GT_GO(128, false, false);
        else if (!a_mn && b_mn) rc = HRP_GT_GO(128, false, true);
        else if (a_mn && !b_mn) rc = HRP_GT_GO(128, true, false);
        else rc = HRP_GT_GO(128, true, true);
    }
#undef HRP_GT_GO
    return rc < 0 ? rc : splits;
}

// debugging aids (not part of include/hrp.h): the TMA GEMM on caller-provided pre-split operands, and the phase clocks
// of its CTA (0,0,0)
extern "C" int hrp_debug_tma_gemm(int M, int N, int K, const float *A, const float *A_lo, long long sam, long long sak,
                                  const float *B, const float *B_lo, long long sbn, long long sbk, float *C, float *C_lo, int ldc,
                                  const float *bias, int relu, int splits, int bn_hint, void *stream)
{
    int rc = hrp_tma_gemm(M, N, K, A, A_lo, sam, sak, B, B_lo, sbn, sbk, C, C_lo, ldc, bias, relu, nullptr, 0, splits, bn_hint,
                          (cudaStream_t)stream);
    return rc < 0 ? rc : 0;
}
extern "C" int hrp_debug_split_lo(const float *x, float *lo, long long n, void *stream) { return hrp_split_lo(x, lo, n, (cudaStream_t)stream); }
extern "C" int hrp_debug_tma_gemm_clocks(int on, long long *out16)
{
    HRP_CUDA_OK(cudaDeviceSynchronize());
    if (out16) HRP_CUDA_OK(cudaMemcpyFromSymbol(out16, g_gt_phase, sizeof(long long) * 16));
    HRP_CUDA_OK(cudaMemcpyToSymbol(g_gt_phase_on, &on, sizeof(int)));
    return 0;
}
