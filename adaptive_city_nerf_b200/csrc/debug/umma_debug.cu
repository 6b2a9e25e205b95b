// Diagnostics: one tcgen05 tile GEMM D(128,N) = A(128,K) * W(N,K)^T (fp16 in, fp32 out) that validates, on real
// hardware, the shared-memory / instruction descriptors and the barrier protocol the fused field-MLP kernels
// (field_mma.cu) are built on.  Not on any product path.
#include "field_common.cuh"
#include "field_internal.cuh"
#include "umma.cuh"
#include "../../../include/acn_b200_debug.h"

namespace {

constexpr int TM = 128;           // points per tile = MMA M
constexpr uint32_t A_SBO = 1024;  // A tile: 8 K-chunks of 128 B per 8-row group (K up to 64)
constexpr uint32_t TMEM_COLS = 64;

// Issue one layer: D(128 x N) = A(128 x K) * W(N x K)^T, K a multiple of 16.  Single thread.
__device__ __forceinline__ void issue_layer(uint32_t tmem_d, const uint8_t* a_tile, const uint8_t* w_tile, int N, int K, uint64_t* bar) {
    const uint32_t idesc = umma::make_idesc_f16(128, N);
    const uint32_t a_addr = umma::smem_u32(a_tile), w_addr = umma::smem_u32(w_tile);
    const uint32_t w_sbo = (uint32_t)(K / 8) * 128u;
    for (int ks = 0; ks < K / 16; ++ks) {
        uint64_t da = umma::make_desc(a_addr + ks * 256, 128, A_SBO);
        uint64_t db = umma::make_desc(w_addr + ks * 256, 128, w_sbo);
        umma::mma_f16_ss(tmem_d, da, db, idesc, ks > 0);
    }
    umma::commit(bar);
}

// all threads: publish smem writes to the async proxy, order TMEM reads, then let thread 0 issue
__device__ __forceinline__ void run_layer(uint32_t tmem_d, const uint8_t* a_tile, const uint8_t* w_tile, int N, int K,
                                          uint64_t* bar, uint32_t& phase) {
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    if (threadIdx.x == 0) {
        umma::fence_after_sync();
        issue_layer(tmem_d, a_tile, w_tile, N, K, bar);
    }
    umma::mbar_wait(bar, phase);
    phase ^= 1u;
    umma::fence_after_sync();
}

// D(128,N) = A(128,K) * W(N,K)^T on one CTA: validates the descriptors on real hardware.
__global__ void __launch_bounds__(TM) k_debug_umma(const __half* __restrict__ a, const __half* __restrict__ w, int N, int K,
                                                   float* __restrict__ d)
{
    __shared__ __align__(128) uint8_t a_tile[TM * 64 * 2];
    __shared__ __align__(128) uint8_t w_tile[64 * 64 * 2];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int c = 0; c < K / 8; ++c)
        *reinterpret_cast<uint4*>(a_tile + umma::chunk_off(tid, c, A_SBO)) = *reinterpret_cast<const uint4*>(a + (size_t)tid * K + c * 8);
    const uint32_t w_sbo = (uint32_t)(K / 8) * 128u;
    for (int idx = tid; idx < N * (K / 8); idx += TM) {
        int n = idx / (K / 8), c = idx - n * (K / 8);
        *reinterpret_cast<uint4*>(w_tile + umma::chunk_off(n, c, w_sbo)) = *reinterpret_cast<const uint4*>(w + (size_t)n * K + c * 8);
    }
    if (tid == 0) { umma::mbar_init(&bar, 1); umma::fence_mbar_init(); }
    if (warp == 0) umma::tmem_alloc(&tmem_ptr, TMEM_COLS);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem_base = tmem_ptr;
    uint32_t phase = 0;
    run_layer(tmem_base, a_tile, w_tile, N, K, &bar, phase);
    const uint32_t tmem_row = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int q = 0; q < N / 16; ++q) {
        float v[16];
        umma::ld16(tmem_row + q * 16, v);
        umma::wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) d[(size_t)tid * N + q * 16 + j] = v[j];
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem_base, TMEM_COLS);
}

// Same product with the A operand in TENSOR MEMORY: thread r packs row r of A into 32-bit words (k even in the low
// half) and stores them with tcgen05.st at columns [64, 64 + K/2); the MMA then reads A from TMEM and B from shared
// memory.  Establishes the TMEM operand layout the forward MLP kernel's activation chain relies on.
__global__ void __launch_bounds__(TM) k_debug_umma_ts(const __half* __restrict__ a, const __half* __restrict__ w, int N, int K,
                                                      float* __restrict__ d)
{
    __shared__ __align__(128) uint8_t w_tile[64 * 64 * 2];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t w_sbo = (uint32_t)(K / 8) * 128u;
    for (int idx = tid; idx < N * (K / 8); idx += TM) {
        int n = idx / (K / 8), c = idx - n * (K / 8);
        *reinterpret_cast<uint4*>(w_tile + umma::chunk_off(n, c, w_sbo)) = *reinterpret_cast<const uint4*>(w + (size_t)n * K + c * 8);
    }
    if (tid == 0) { umma::mbar_init(&bar, 1); umma::fence_mbar_init(); }
    if (warp == 0) umma::tmem_alloc(&tmem_ptr, 128);
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem_base = tmem_ptr;
    const uint32_t tmem_row = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int c8 = 0; c8 < K / 16; ++c8) {           // 16 halves = 8 words per store
        uint32_t r[8];
        const uint4 q0 = *reinterpret_cast<const uint4*>(a + (size_t)tid * K + c8 * 16);
        const uint4 q1 = *reinterpret_cast<const uint4*>(a + (size_t)tid * K + c8 * 16 + 8);
        r[0] = q0.x; r[1] = q0.y; r[2] = q0.z; r[3] = q0.w; r[4] = q1.x; r[5] = q1.y; r[6] = q1.z; r[7] = q1.w;
        umma::st8(tmem_row + 64 + c8 * 8, r);
    }
    umma::wait_st();
    umma::fence_before_sync();
    __syncthreads();
    if (tid == 0) {
        umma::fence_after_sync();
        const uint32_t idesc = umma::make_idesc_f16(128, N);
        const uint32_t w_addr = umma::smem_u32(w_tile);
        for (int ks = 0; ks < K / 16; ++ks)
            umma::mma_f16_ts(tmem_base, tmem_base + 64 + ks * 8, umma::make_desc(w_addr + ks * 256, 128, w_sbo), idesc, ks > 0);
        umma::commit(&bar);
    }
    umma::mbar_wait(&bar, 0);
    umma::fence_after_sync();
    for (int q = 0; q < N / 16; ++q) {
        float v[16];
        umma::ld16(tmem_row + q * 16, v);
        umma::wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) d[(size_t)tid * N + q * 16 + j] = v[j];
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem_base, 128);
}

// Dispatch-rate probe: `issuers` threads (one per warp, warps 0..issuers-1 = different SM sub-partitions) each issue
// `nmma` back-to-back MMAs into their own accumulator window, commit and wait, `reps` times; thread 0 reports the
// average SM cycles per round.  mode 0: A and B from shared memory; 1: A from tensor memory.
__global__ void __launch_bounds__(TM) k_umma_rate(int mode, int M, int N, int nmma, int reps, int issuers, long long* __restrict__ out)
{
    __shared__ __align__(128) uint8_t a_tile[TM * 64 * 2];
    __shared__ __align__(128) uint8_t w_tile[64 * 64 * 2];
    __shared__ __align__(8) uint64_t bar[4];
    __shared__ uint32_t tmem_ptr;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < TM * 64 * 2 / 16; i += TM) reinterpret_cast<uint4*>(a_tile)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < 64 * 64 * 2 / 16; i += TM) reinterpret_cast<uint4*>(w_tile)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) { for (int i = 0; i < 4; ++i) umma::mbar_init(&bar[i], 1); umma::fence_mbar_init(); }
    if (warp == 0) umma::tmem_alloc(&tmem_ptr, 512);
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem_base = tmem_ptr;
    if (warp < issuers && lane == 0) {
        const uint32_t idesc = umma::make_idesc_f16(M, N);
        const uint32_t a_addr = umma::smem_u32(a_tile), w_addr = umma::smem_u32(w_tile);
        const uint32_t d = tmem_base + warp * 128, a_t = d + 64;
        uint32_t phase = 0;
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            for (int i = 0; i < nmma; ++i) {
                const int ks = i & 3;
                const uint64_t db = umma::make_desc(w_addr + ks * 256, 128, 1024);
                if (mode == 0) umma::mma_f16_ss(d, umma::make_desc(a_addr + ks * 256, 128, 1024), db, idesc, true);
                else umma::mma_f16_ts(d, a_t + ks * 8, db, idesc, true);
            }
            umma::commit(&bar[warp]);
            umma::mbar_wait(&bar[warp], phase);
            phase ^= 1u;
        }
        long long t1 = clock64();
        if (blockIdx.x == 0) out[warp] = (t1 - t0) / reps;
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem_base, 512);
}

}  // namespace

extern "C" int acn_debug_umma_rate(acn_ctx* ctx, int mode, int M, int N, int nmma, int reps, int issuers, long long* out4, acn_stream stream) {
    ACN_CHECK_CTX(ctx);
    ACN_REQUIRE(out4 && (mode == 0 || mode == 1) && (M == 64 || M == 128) && N >= 8 && N <= 64 && N % 8 == 0 && nmma >= 1 && reps >= 1 &&
                issuers >= 1 && issuers <= 4, ACN_EINVAL, "acn_debug_umma_rate: bad arguments");
    k_umma_rate<<<ctx->sm_count, TM, 0, (cudaStream_t)stream>>>(mode, M, N, nmma, reps, issuers, out4);
    ACN_CHECK_LAUNCH();
    return ACN_OK;
}

extern "C" int acn_debug_umma_gemm_ts(acn_ctx* ctx, const void* a_f16, const void* w_f16, int N, int K, float* d, acn_stream stream) {
    ACN_CHECK_CTX(ctx);
    ACN_REQUIRE(a_f16 && w_f16 && d, ACN_EINVAL, "acn_debug_umma_gemm_ts: null buffer");
    ACN_REQUIRE((N == 16 || N == 32 || N == 64) && (K == 16 || K == 32 || K == 64), ACN_EUNSUPPORTED,
                "acn_debug_umma_gemm_ts: N,K must be in {16,32,64}");
    k_debug_umma_ts<<<1, TM, 0, (cudaStream_t)stream>>>((const __half*)a_f16, (const __half*)w_f16, N, K, d);
    ACN_CHECK_LAUNCH();
    return ACN_OK;
}

extern "C" int acn_debug_umma_gemm(acn_ctx* ctx, const void* a_f16, const void* w_f16, int N, int K, float* d, acn_stream stream) {
    ACN_CHECK_CTX(ctx);
    ACN_REQUIRE(a_f16 && w_f16 && d, ACN_EINVAL, "acn_debug_umma_gemm: null buffer");
    ACN_REQUIRE((N == 16 || N == 32 || N == 64) && (K == 16 || K == 32 || K == 64), ACN_EUNSUPPORTED,
                "acn_debug_umma_gemm: N,K must be in {16,32,64}");
    k_debug_umma<<<1, TM, 0, (cudaStream_t)stream>>>((const __half*)a_f16, (const __half*)w_f16, N, K, d);
    ACN_CHECK_LAUNCH();
    return ACN_OK;
}
