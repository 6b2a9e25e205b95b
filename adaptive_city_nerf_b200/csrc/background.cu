// Background head of the container (models/inr/meta_container.py:79-93, :347-382): per RAY direction
//   F.normalize(d) -> SH degree 3 (16 comps) -> Linear(16,Hb) -> ReLU -> Linear(Hb,3) -> Sigmoid,
// forward and backward as one kernel each (the reference runs ~8 small launches and two GEMMs per call; a 1080p frame has
// 2 M rays).  fp32 arithmetic (the reference's autocast path rounds the same products to fp16; this is the fp32 value of
// the same expression).  Thread = ray; the 16*Hb + 4*Hb + 3 weights sit in shared memory (broadcast reads).  The backward
// recomputes the hidden layer and reduces the weight gradients warp -> block (shared memory) -> one atomicAdd per
// weight and block.
#include "field_common.cuh"

namespace {

constexpr int BG_MAXH = 64;

__device__ __forceinline__ void bg_sh(const float* __restrict__ d, float* sh) {
    float x = d[0], y = d[1], z = d[2];
    // F.normalize (eps 1e-12), then SHEncoder's own normalise (models/encodings.py:141, eps 1e-9)
    float n = fmaxf(sqrtf(x * x + y * y + z * z), 1e-12f);
    x = __fdiv_rn(x, n); y = __fdiv_rn(y, n); z = __fdiv_rn(z, n);
    n = fmaxf(sqrtf(x * x + y * y + z * z), 1e-9f);
    x = __fdiv_rn(x, n); y = __fdiv_rn(y, n); z = __fdiv_rn(z, n);
    sh16_poly(x, y, z, sh);
}

struct BgW { const float* w1; const float* b1; const float* w2; const float* b2; };
struct BgG { float* w1; float* b1; float* w2; float* b2; };

// shared layout: w1 (Hb,16) | b1 (Hb) | w2 (3,Hb) | b2 (3)
__device__ __forceinline__ void bg_stage(const BgW& w, int Hb, float* s) {
    for (int i = threadIdx.x; i < Hb * 16; i += blockDim.x) s[i] = __ldg(w.w1 + i);
    for (int i = threadIdx.x; i < Hb; i += blockDim.x) s[Hb * 16 + i] = __ldg(w.b1 + i);
    for (int i = threadIdx.x; i < 3 * Hb; i += blockDim.x) s[Hb * 17 + i] = __ldg(w.w2 + i);
    if (threadIdx.x < 3) s[Hb * 20 + threadIdx.x] = __ldg(w.b2 + threadIdx.x);
    __syncthreads();
}

template <typename OutT>
__global__ void __launch_bounds__(256) k_bg_fwd(const float* __restrict__ dirs, int64_t N, int stride, BgW w, int Hb, OutT* __restrict__ rgb) {
    extern __shared__ float s[];
    bg_stage(w, Hb, s);
    const float *w1 = s, *b1 = s + Hb * 16, *w2 = s + Hb * 17, *b2 = s + Hb * 20;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < N; r += (int64_t)gridDim.x * blockDim.x) {
        float sh[16];
        bg_sh(dirs + r * stride, sh);
        float y0 = b2[0], y1 = b2[1], y2 = b2[2];
        for (int j = 0; j < Hb; ++j) {
            float h = b1[j];
#pragma unroll
            for (int i = 0; i < 16; ++i) h = fmaf(w1[j * 16 + i], sh[i], h);
            h = fmaxf(h, 0.0f);
            y0 = fmaf(w2[j], h, y0); y1 = fmaf(w2[Hb + j], h, y1); y2 = fmaf(w2[2 * Hb + j], h, y2);
        }
        if constexpr (sizeof(OutT) == 4) {
            rgb[3 * r] = sigmoid_f(y0); rgb[3 * r + 1] = sigmoid_f(y1); rgb[3 * r + 2] = sigmoid_f(y2);
        } else {
            rgb[3 * r] = __float2half_rn(sigmoid_f(y0)); rgb[3 * r + 1] = __float2half_rn(sigmoid_f(y1)); rgb[3 * r + 2] = __float2half_rn(sigmoid_f(y2));
        }
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// d_rgb (N,3) fp32 -> weight gradients (accumulated).  Block-local accumulators in shared memory after the weights.
__global__ void __launch_bounds__(256) k_bg_bwd(const float* __restrict__ dirs, int64_t N, int stride, BgW w, int Hb,
                                                const float* __restrict__ d_rgb, BgG g) {
    extern __shared__ float s[];
    bg_stage(w, Hb, s);
    const int NW = Hb * 20 + 3;
    const float *w1 = s, *b1 = s + Hb * 16, *w2 = s + Hb * 17, *b2 = s + Hb * 20;
    float* acc = s + ((NW + 3) & ~3);            // same layout as the weights
    for (int i = threadIdx.x; i < NW; i += blockDim.x) acc[i] = 0.0f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t nwarp_iters = (N + 31) / 32;
    for (int64_t wi = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; wi < nwarp_iters; wi += ((int64_t)gridDim.x * blockDim.x) >> 5) {
        const int64_t r = wi * 32 + lane;
        const bool on = r < N;
        float sh[16];
        const float zero3[3] = { 0.f, 0.f, 1.f };
        bg_sh(on ? dirs + r * stride : zero3, sh);
        float h[BG_MAXH];
        float y0 = b2[0], y1 = b2[1], y2 = b2[2];
#pragma unroll
        for (int j = 0; j < BG_MAXH; ++j) {
            if (j < Hb) {
                float a = b1[j];
#pragma unroll
                for (int i = 0; i < 16; ++i) a = fmaf(w1[j * 16 + i], sh[i], a);
                a = fmaxf(a, 0.0f);
                h[j] = a;
                y0 = fmaf(w2[j], a, y0); y1 = fmaf(w2[Hb + j], a, y1); y2 = fmaf(w2[2 * Hb + j], a, y2);
            }
        }
        float dy[3] = { 0.f, 0.f, 0.f };
        if (on) {
            const float s0 = sigmoid_f(y0), s1 = sigmoid_f(y1), s2 = sigmoid_f(y2);
            dy[0] = d_rgb[3 * r] * s0 * (1.0f - s0); dy[1] = d_rgb[3 * r + 1] * s1 * (1.0f - s1); dy[2] = d_rgb[3 * r + 2] * s2 * (1.0f - s2);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = warp_sum(dy[c]);
            if (lane == 0) atomicAdd(acc + Hb * 20 + c, v);
        }
#pragma unroll
        for (int j = 0; j < BG_MAXH; ++j) {
            if (j < Hb) {
                // dW2[c][j] += dy[c] * h[j];  dh = sum_c W2[c][j] dy[c] masked by h > 0
                const float hj = h[j];
                const float dh = hj > 0.0f ? fmaf(w2[j], dy[0], fmaf(w2[Hb + j], dy[1], w2[2 * Hb + j] * dy[2])) : 0.0f;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float v = warp_sum(dy[c] * hj);
                    if (lane == 0) atomicAdd(acc + Hb * 17 + c * Hb + j, v);
                }
                const float vb = warp_sum(dh);
                if (lane == 0) atomicAdd(acc + Hb * 16 + j, vb);
                // dW1[j][i] += dh * sh[i]: lane i ends up holding the sum for column i (transpose-reduce over the warp)
                float mine = 0.0f;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float v = warp_sum(dh * sh[i]);
                    if (lane == i) mine = v;
                }
                if (lane < 16) atomicAdd(acc + j * 16 + lane, mine);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NW; i += blockDim.x) {
        const float v = acc[i];
        if (v == 0.0f) continue;
        float* dst = i < Hb * 16 ? (g.w1 ? g.w1 + i : nullptr)
                   : i < Hb * 17 ? (g.b1 ? g.b1 + (i - Hb * 16) : nullptr)
                   : i < Hb * 20 ? (g.w2 ? g.w2 + (i - Hb * 17) : nullptr)
                   : (g.b2 ? g.b2 + (i - Hb * 20) : nullptr);
        if (dst) atomicAdd(dst, v);
    }
}

int check_bg(const char* fn, int64_t N, int stride, int Hb, const float* const* w4) {
    ACN_REQUIRE(N >= 0 && stride >= 3, ACN_EINVAL, "%s: bad N / stride", fn);
    ACN_REQUIRE(Hb >= 1 && Hb <= BG_MAXH, ACN_EUNSUPPORTED, "%s: hidden width %d outside [1,%d]", fn, Hb, BG_MAXH);
    for (int i = 0; i < 4; ++i) ACN_REQUIRE(w4[i] != nullptr, ACN_EINVAL, "%s: weight pointer %d is null", fn, i);
    return ACN_OK;
}

}  // namespace

extern "C" int acn_background_fwd(acn_ctx* ctx, const float* dirs, int64_t N, int stride, const float* w1, const float* b1,
                                  const float* w2, const float* b2, int hidden, void* rgb, int rgb_dtype, acn_stream stream) {
    ACN_CHECK_CTX(ctx);
    const float* w4[4] = { w1, b1, w2, b2 };
    int rc = check_bg("acn_background_fwd", N, stride, hidden, w4);
    if (rc) return rc;
    ACN_REQUIRE(rgb_dtype == ACN_F32 || rgb_dtype == ACN_F16, ACN_EINVAL, "acn_background_fwd: bad rgb dtype");
    if (N == 0) return ACN_OK;
    ACN_REQUIRE(dirs && rgb, ACN_EINVAL, "acn_background_fwd: null buffer");
    const BgW w{ w1, b1, w2, b2 };
    const size_t smem = (size_t)(hidden * 20 + 3) * sizeof(float);
    const int grid = acn_grid_1d(N, 256, (int64_t)ctx->sm_count * 8);
    if (rgb_dtype == ACN_F32) k_bg_fwd<float><<<grid, 256, smem, (cudaStream_t)stream>>>(dirs, N, stride, w, hidden, (float*)rgb);
    else k_bg_fwd<__half><<<grid, 256, smem, (cudaStream_t)stream>>>(dirs, N, stride, w, hidden, (__half*)rgb);
    ACN_CHECK_LAUNCH();
    return ACN_OK;
}

extern "C" int acn_background_bwd(acn_ctx* ctx, const float* dirs, int64_t N, int stride, const float* w1, const float* b1,
                                  const float* w2, const float* b2, int hidden, const float* d_rgb, float* g_w1, float* g_b1,
                                  float* g_w2, float* g_b2, acn_stream stream) {
    ACN_CHECK_CTX(ctx);
    const float* w4[4] = { w1, b1, w2, b2 };
    int rc = check_bg("acn_background_bwd", N, stride, hidden, w4);
    if (rc) return rc;
    if (N == 0) return ACN_OK;
    ACN_REQUIRE(dirs && d_rgb, ACN_EINVAL, "acn_background_bwd: null buffer");
    const BgW w{ w1, b1, w2, b2 };
    const BgG g{ g_w1, g_b1, g_w2, g_b2 };
    const int NW = hidden * 20 + 3;
    const size_t smem = (size_t)(((NW + 3) & ~3) + NW) * sizeof(float);
    const int grid = acn_grid_1d(N, 256, (int64_t)ctx->sm_count * 4);
    k_bg_bwd<<<grid, 256, smem, (cudaStream_t)stream>>>(dirs, N, stride, w, hidden, d_rgb, g);
    ACN_CHECK_LAUNCH();
    return ACN_OK;
}
