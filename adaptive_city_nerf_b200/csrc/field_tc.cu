// Stage 3, fp16 precision: the expert's field MLPs as ONE tcgen05 kernel (forward).
//
// This is the reference's autocast(fp16) path (models/metamodule/metamodule.py:150-155 under
// torch.autocast: fp16 GEMM operands, fp32 accumulate, fp32 bias + ReLU) -- the one real dense
// contraction on the hot path.  A 128-thread CTA owns a tile of 128 points:
//   * thread t owns point t: it stages the point's fp16 feature row in shared memory (canonical
//     K-major UMMA layout, see umma.cuh), and after every layer reads ITS accumulator row back
//     from TMEM (lane t), applies bias + ReLU in fp32, rounds to fp16 and rewrites the row as the
//     next layer's A operand.  Activations never leave the SM.
//   * one thread issues the layer's tcgen05.mma's (M=128, N=64|16, K=16 per instruction) with
//     all six weight matrices resident in shared memory as B operands, and commits to an mbarrier.
// The six layers of a tile are serially dependent, so tensor-pipe overlap comes from running
// several CTAs per SM (each needs ~46 KB smem and 64 TMEM columns).
#include "field_common.cuh"
#include "field_internal.cuh"
#include "umma.cuh"

namespace {

constexpr int TM = 128;           // points per tile = MMA M
constexpr uint32_t A_SBO = 1024;  // A tile: 8 K-chunks of 128 B per 8-row group (K up to 64)
constexpr uint32_t TMEM_COLS = 64;

struct TcSmem {
    uint8_t *a, *w_t0, *w_t1, *w_hd, *w_c0, *w_c1, *w_c2;
    float *b_t0, *b_t1, *b_hd, *b_c0, *b_c1, *b_c2;
    uint64_t* bar;
    uint32_t* tmem_ptr;
};

__host__ __device__ inline size_t tc_carve(int E, uint8_t* base, TcSmem* s) {
    size_t off = 0;
    auto take = [&](size_t bytes) { uint8_t* p = base ? base + off : nullptr; off += (bytes + 127) & ~(size_t)127; return p; };
    uint8_t* a = take(TM * 64 * 2);
    uint8_t* w_t0 = take(64 * E * 2);
    uint8_t* w_t1 = take(64 * 64 * 2);
    uint8_t* w_hd = take(16 * 64 * 2);
    uint8_t* w_c0 = take(64 * 32 * 2);
    uint8_t* w_c1 = take(64 * 64 * 2);
    uint8_t* w_c2 = take(16 * 64 * 2);
    float* b_t0 = (float*)take(64 * 4); float* b_t1 = (float*)take(64 * 4); float* b_hd = (float*)take(16 * 4);
    float* b_c0 = (float*)take(64 * 4); float* b_c1 = (float*)take(64 * 4); float* b_c2 = (float*)take(16 * 4);
    uint64_t* bar = (uint64_t*)take(8);
    uint32_t* tp = (uint32_t*)take(4);
    if (s) *s = TcSmem{ a, w_t0, w_t1, w_hd, w_c0, w_c1, w_c2, b_t0, b_t1, b_hd, b_c0, b_c1, b_c2, bar, tp };
    return off;
}

// W (rows, K) fp16 canonical tile from an fp32 (n_src, k_src) row-major matrix, zero padded.
// row_map(n) gives the source row (or -1 for a zero row).
template <typename RowSrc>
__device__ void stage_weight(uint8_t* tile, int rows, int K, int k_src, RowSrc src_row) {
    const uint32_t sbo = (uint32_t)(K / 8) * 128u;
    for (int idx = threadIdx.x; idx < rows * (K / 8); idx += blockDim.x) {
        int n = idx / (K / 8), c = idx - n * (K / 8);
        const float* src = src_row(n);
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int k = c * 8 + j;
            v[j] = (src && k < k_src) ? __ldg(src + k) : 0.0f;
        }
        umma::st_chunk(tile, n, c, sbo, v);
    }
}

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

// hidden-layer epilogue: 64 accumulator columns -> bias + ReLU -> fp16 row of the A tile
__device__ __forceinline__ void hidden_epilogue(uint32_t tmem_row, const float* bias, uint8_t* a_tile, int row) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float v[16];
        umma::ld16(tmem_row + q * 16, v);
        umma::wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j] + bias[q * 16 + j], 0.0f);
        umma::st_chunk(a_tile, row, 2 * q, A_SBO, v);
        umma::st_chunk(a_tile, row, 2 * q + 1, A_SBO, v + 8);
    }
}

template <typename EncT>
__global__ void __launch_bounds__(TM) k_field_fwd_tc(
    const EncT* __restrict__ enc, const float* __restrict__ dirs, int dstride, int dgroup, int64_t P, int E, int G,
    acn_field_weights w, float4* __restrict__ rgb_sigma)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    TcSmem s;
    tc_carve(E, smem_raw, &s);
    const int tid = threadIdx.x, warp = tid >> 5;

    // ---- one-time setup: weights -> fp16 canonical tiles, biases, barrier, TMEM ----
    stage_weight(s.w_t0, 64, E, E, [&](int n) { return w.p[0] + (size_t)n * E; });
    stage_weight(s.w_t1, 64, 64, 64, [&](int n) { return w.p[2] + (size_t)n * 64; });
    stage_weight(s.w_hd, 16, 64, 64, [&](int n) { return n < G ? w.p[6] + (size_t)n * 64 : (n == 15 ? w.p[4] : (const float*)nullptr); });
    stage_weight(s.w_c0, 64, 32, G + 16, [&](int n) { return w.p[8] + (size_t)n * (G + 16); });
    stage_weight(s.w_c1, 64, 64, 64, [&](int n) { return w.p[10] + (size_t)n * 64; });
    stage_weight(s.w_c2, 16, 64, 64, [&](int n) { return n < 3 ? w.p[12] + (size_t)n * 64 : (const float*)nullptr; });
    if (tid < 64) {
        s.b_t0[tid] = __ldg(w.p[1] + tid); s.b_t1[tid] = __ldg(w.p[3] + tid);
        s.b_c0[tid] = __ldg(w.p[9] + tid); s.b_c1[tid] = __ldg(w.p[11] + tid);
    }
    if (tid < 16) {
        s.b_hd[tid] = tid < G ? __ldg(w.p[7] + tid) : (tid == 15 ? __ldg(w.p[5]) : 0.0f);
        s.b_c2[tid] = tid < 3 ? __ldg(w.p[13] + tid) : 0.0f;
    }
    if (tid == 0) { umma::mbar_init(s.bar, 1); umma::fence_mbar_init(); }
    if (warp == 0) umma::tmem_alloc(s.tmem_ptr, TMEM_COLS);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem_base = *s.tmem_ptr;
    const uint32_t tmem_row = tmem_base + ((uint32_t)(warp * 32) << 16);  // this warp's 32 lanes
    uint32_t phase = 0;

    const int64_t ntiles = (P + TM - 1) / TM;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t p = tile * TM + tid;
        const bool on = p < P;
        // ---- stage this point's encoding row (fp16) ----
        for (int c = 0; c < E / 8; ++c) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = 0.0f;
            if (on) {
                if constexpr (sizeof(EncT) == 2) {
                    uint4 q = __ldg(reinterpret_cast<const uint4*>(enc + p * E) + c);
                    *reinterpret_cast<uint4*>(s.a + umma::chunk_off(tid, c, A_SBO)) = q;
                    continue;
                } else {
                    const float4* src = reinterpret_cast<const float4*>(enc + p * E + c * 8);
                    float4 a = __ldg(src), b = __ldg(src + 1);
                    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
                }
            }
            umma::st_chunk(s.a, tid, c, A_SBO, v);
        }
        // ---- density trunk ----
        run_layer(tmem_base, s.a, s.w_t0, 64, E, s.bar, phase);
        hidden_epilogue(tmem_row, s.b_t0, s.a, tid);
        run_layer(tmem_base, s.a, s.w_t1, 64, 64, s.bar, phase);
        hidden_epilogue(tmem_row, s.b_t1, s.a, tid);
        // ---- heads: cols 0..G-1 geo, col 15 sigma_raw ----
        run_layer(tmem_base, s.a, s.w_hd, 16, 64, s.bar, phase);
        float sigma;
        {
            float v[16], cin[32];
            umma::ld16(tmem_row, v);
            umma::wait_ld();
            sigma = trunc_exp_f(v[15] + s.b_hd[15]);
#pragma unroll
            for (int j = 0; j < 15; ++j) cin[j] = j < G ? v[j] + s.b_hd[j] : 0.0f;
            float sh[16];
            if (on) {
                const float* dp = dir_of(dirs, dstride, dgroup, p);
                sh16_expert(__ldg(dp), __ldg(dp + 1), __ldg(dp + 2), sh);
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) sh[j] = 0.0f;
            }
            // cin = [geo(G), sh(16), 0...]: geo + zero fill as whole chunks, then the SH block as 16
            // scalar fp16 stores starting at column G (same thread, program order)
#pragma unroll
            for (int j = 15; j < 32; ++j) cin[j] = 0.0f;
#pragma unroll
            for (int c = 0; c < 4; ++c) umma::st_chunk(s.a, tid, c, A_SBO, cin + 8 * c);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                int k = G + j;
                *reinterpret_cast<__half*>(s.a + umma::chunk_off(tid, k >> 3, A_SBO) + (k & 7) * 2) = __float2half_rn(sh[j]);
            }
        }
        // ---- colour MLP ----
        run_layer(tmem_base, s.a, s.w_c0, 64, 32, s.bar, phase);
        hidden_epilogue(tmem_row, s.b_c0, s.a, tid);
        run_layer(tmem_base, s.a, s.w_c1, 64, 64, s.bar, phase);
        hidden_epilogue(tmem_row, s.b_c1, s.a, tid);
        run_layer(tmem_base, s.a, s.w_c2, 16, 64, s.bar, phase);
        {
            float v[16];
            umma::ld16(tmem_row, v);
            umma::wait_ld();
            if (on) rgb_sigma[p] = make_float4(sigmoid_f(v[0] + s.b_c2[0]), sigmoid_f(v[1] + s.b_c2[1]),
                                               sigmoid_f(v[2] + s.b_c2[2]), sigma);
        }
    }
    // ---- teardown ----
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem_base, TMEM_COLS);
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

}  // namespace

int acn_field_fwd_tc(acn_ctx* ctx, const void* enc, int enc_dtype, const float* dirs, int dirs_stride, int dirs_group,
                     int64_t P, int E, int H, int G, int C, const acn_field_weights* w, float* rgb_sigma, cudaStream_t st) {
    ACN_REQUIRE(H == 64 && C == 64, ACN_EUNSUPPORTED, "acn_field_fwd(f16): hidden widths must be 64 (got H=%d, C=%d)", H, C);
    ACN_REQUIRE(E == 16 || E == 32 || E == 48 || E == 64, ACN_EUNSUPPORTED, "acn_field_fwd(f16): encoding width %d not in {16,32,48,64}", E);
    ACN_REQUIRE(G >= 1 && G <= 15, ACN_EUNSUPPORTED, "acn_field_fwd(f16): geo_feat_dim %d outside [1,15]", G);
    ACN_REQUIRE(((uintptr_t)enc & 15) == 0, ACN_EINVAL, "acn_field_fwd(f16): enc must be 16-byte aligned");
    const size_t smem = tc_carve(E, nullptr, nullptr);
    const int64_t ntiles = (P + TM - 1) / TM;
    int ctas_per_sm = (int)((size_t)ctx->max_smem_optin / (smem + 1024));
    if (ctas_per_sm > 8) ctas_per_sm = 8;   // 8 x 64 TMEM columns = 512
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    int64_t grid = (int64_t)ctx->sm_count * ctas_per_sm;
    if (grid > ntiles) grid = ntiles;
    if (enc_dtype == ACN_F16) {
        ACN_CUDA(cudaFuncSetAttribute(k_field_fwd_tc<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_field_fwd_tc<__half><<<(int)grid, TM, smem, st>>>((const __half*)enc, dirs, dirs_stride, dirs_group, P, E, G, *w, (float4*)rgb_sigma);
    } else {
        ACN_CUDA(cudaFuncSetAttribute(k_field_fwd_tc<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_field_fwd_tc<float><<<(int)grid, TM, smem, st>>>((const float*)enc, dirs, dirs_stride, dirs_group, P, E, G, *w, (float4*)rgb_sigma);
    }
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
