// Stage 4: alpha compositing (nerfs/ray_rendering.py:114-165 volume_render) and its backward.
//
// One warp per ray.  Samples are processed in chunks of 32 (lane = sample) so the (rgb,sigma)
// float4 loads and the weight stores are fully coalesced; transmittance is a warp-level
// prefix product (5 shuffles per chunk) carried across chunks, i.e. a segmented scan with the
// ray as the segment.  Once the carried transmittance drops below T_EPS the remaining chunks
// are skipped (their weights are written as exact zeros) -- the reference has no early exit,
// the skipped mass is < 1e-8 per ray.  All math fp32, like the reference under autocast.
#include "acn_common.cuh"

#define FULL 0xffffffffu
static constexpr float T_EPS = 1e-8f;

__device__ __forceinline__ float warp_incl_prod(float v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        float o = __shfl_up_sync(FULL, v, d);
        if (lane >= d) v *= o;
    }
    return v;
}

__device__ __forceinline__ float warp_incl_sum_rev(float v, int lane) {  // sum over lanes >= lane
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        float o = __shfl_down_sync(FULL, v, d);
        if (lane + d < 32) v += o;
    }
    return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
    return v;
}

// Interval length of sample s: clamp_min(t[s+1]-t[s], 1e-4), the last one repeated.
__device__ __forceinline__ float interval(const float* __restrict__ t, int s, int S) {
    int a = s < S - 1 ? s : S - 2;
    return fmaxf(__fsub_rn(__ldg(t + a + 1), __ldg(t + a)), 1e-4f);
}

__device__ __forceinline__ float alpha_of(float sigma_raw, float sigma_scale, float dl, float* e_out) {
    float sig = fmaxf(sigma_raw, 0.0f) * sigma_scale;
    float e = expf(-sig * dl);
    if (e_out) *e_out = e;
    const float amax = 0x1.fffffcp-1f;  // fp32(1 - 1e-7)
    return fminf(fmaxf(1.0f - e, 0.0f), amax);
}

__global__ void __launch_bounds__(256) k_composite_fwd(
    const float4* __restrict__ rgb_sigma, const float* __restrict__ t_vals, const float* __restrict__ bg,
    int64_t N, int S, float sigma_scale, float* __restrict__ rgb, float* __restrict__ depth,
    float* __restrict__ weights, float* __restrict__ acc_out)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp0; r < N; r += nwarps) {
        const float4* rs = rgb_sigma + r * S;
        const float* t = t_vals + r * S;
        float* w_out = weights + r * S;
        float carry = 1.0f, cr = 0.f, cg = 0.f, cb = 0.f, dep = 0.f, acc = 0.f;
        int c0 = 0;
        for (; c0 < S; c0 += 32) {
            if (carry < T_EPS) break;
            int s = c0 + lane;
            bool on = s < S;
            float4 v = on ? ld_stream_f4(rs + s) : make_float4(0.f, 0.f, 0.f, 0.f);
            float ts = on ? __ldg(t + s) : 0.0f;
            float a = on ? alpha_of(v.w, sigma_scale, interval(t, s, S), nullptr) : 0.0f;
            float q = on ? (1.0f - a) + 1e-10f : 1.0f;
            float incl = warp_incl_prod(q, lane);
            float excl = __shfl_up_sync(FULL, incl, 1);
            if (lane == 0) excl = 1.0f;
            float w = a * (carry * excl);
            if (on) w_out[s] = w;
            cr += w * clampf(v.x, 0.f, 1.f);
            cg += w * clampf(v.y, 0.f, 1.f);
            cb += w * clampf(v.z, 0.f, 1.f);
            dep += w * ts;
            acc += w;
            carry *= __shfl_sync(FULL, incl, 31);
        }
        for (; c0 < S; c0 += 32) { int s = c0 + lane; if (s < S) w_out[s] = 0.0f; }  // early exit tail
        cr = warp_sum(cr); cg = warp_sum(cg); cb = warp_sum(cb); dep = warp_sum(dep); acc = warp_sum(acc);
        if (lane == 0) {
            if (bg) {
                float rem = 1.0f - acc;
                cr += rem * __ldg(bg + 3 * r); cg += rem * __ldg(bg + 3 * r + 1); cb += rem * __ldg(bg + 3 * r + 2);
            }
            rgb[3 * r] = cr; rgb[3 * r + 1] = cg; rgb[3 * r + 2] = cb;
            depth[r] = dep; acc_out[r] = acc;
        }
    }
}

// Backward.  Pass 1 recomputes alpha / T per sample and parks them in the (not yet final)
// d_rgb_sigma row (each lane later re-reads only what it wrote itself).  Pass 2 walks the
// chunks in reverse with a carried suffix sum of gw_i * w_i:
//   d alpha_s = gw_s T_s - (sum_{i>s} gw_i w_i) / q_s,   q_s = 1 - alpha_s + 1e-10.
__global__ void __launch_bounds__(256) k_composite_bwd(
    const float4* __restrict__ rgb_sigma, const float* __restrict__ t_vals, const float* __restrict__ bg,
    int64_t N, int S, float sigma_scale, const float* __restrict__ g_rgb, const float* __restrict__ g_depth,
    const float* __restrict__ g_weights, const float* __restrict__ g_acc, float4* __restrict__ d_rgb_sigma,
    float* __restrict__ d_bg)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const float amax = 0x1.fffffcp-1f;
    for (int64_t r = warp0; r < N; r += nwarps) {
        const float4* rs = rgb_sigma + r * S;
        const float* t = t_vals + r * S;
        float4* d = d_rgb_sigma + r * S;
        float gr = 0.f, gg = 0.f, gb = 0.f;
        if (g_rgb) { gr = __ldg(g_rgb + 3 * r); gg = __ldg(g_rgb + 3 * r + 1); gb = __ldg(g_rgb + 3 * r + 2); }
        float gd = g_depth ? __ldg(g_depth + r) : 0.0f;
        float ga = g_acc ? __ldg(g_acc + r) : 0.0f;
        if (bg) ga -= gr * __ldg(bg + 3 * r) + gg * __ldg(bg + 3 * r + 1) + gb * __ldg(bg + 3 * r + 2);
        // pass 1
        float carry = 1.0f, acc = 0.0f;
        for (int c0 = 0; c0 < S; c0 += 32) {
            int s = c0 + lane;
            bool on = s < S;
            float sg = on ? rs[s].w : 0.0f;
            float a = on ? alpha_of(sg, sigma_scale, interval(t, s, S), nullptr) : 0.0f;
            float q = on ? (1.0f - a) + 1e-10f : 1.0f;
            float incl = warp_incl_prod(q, lane);
            float excl = __shfl_up_sync(FULL, incl, 1);
            if (lane == 0) excl = 1.0f;
            float T = carry * excl;
            if (on) d[s] = make_float4(a, T, 0.f, 0.f);
            acc += a * T;
            carry *= __shfl_sync(FULL, incl, 31);
        }
        acc = warp_sum(acc);
        if (d_bg && bg && lane < 3) d_bg[3 * r + lane] = (1.0f - acc) * (lane == 0 ? gr : (lane == 1 ? gg : gb));
        // pass 2
        float suffix = 0.0f;
        int last = ((S - 1) / 32) * 32;
        for (int c0 = last; c0 >= 0; c0 -= 32) {
            int s = c0 + lane;
            bool on = s < S;
            float4 v = on ? rs[s] : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 st = on ? d[s] : make_float4(0.f, 1.f, 0.f, 0.f);
            float a = st.x, T = st.y;
            float w = a * T;
            float ts = on ? __ldg(t + s) : 0.0f;
            float gw = gr * clampf(v.x, 0.f, 1.f) + gg * clampf(v.y, 0.f, 1.f) + gb * clampf(v.z, 0.f, 1.f) + gd * ts + ga;
            if (g_weights && on) gw += __ldg(g_weights + r * S + s);
            float x = on ? gw * w : 0.0f;
            float incl = warp_incl_sum_rev(x, lane);
            float after = suffix + (incl - x);
            float q = (1.0f - a) + 1e-10f;
            float da = gw * T - after / q;
            float dl = on ? interval(t, s, S) : 1.0f;
            float e;
            float araw_a = alpha_of(v.w, sigma_scale, dl, &e);
            (void)araw_a;
            float araw = 1.0f - e;
            float ds = (araw >= 0.0f && araw <= amax) ? da * dl * e * sigma_scale : 0.0f;
            if (!(v.w >= 0.0f)) ds = 0.0f;
            if (on) {
                float4 o;
                o.x = (v.x >= 0.f && v.x <= 1.f) ? w * gr : 0.f;
                o.y = (v.y >= 0.f && v.y <= 1.f) ? w * gg : 0.f;
                o.z = (v.z >= 0.f && v.z <= 1.f) ? w * gb : 0.f;
                o.w = ds;
                d[s] = o;
            }
            suffix += __shfl_sync(FULL, incl, 0);
        }
    }
}

// ------------------------------------------------------------------------------------------ C ABI
static int composite_grid(acn_ctx* ctx, int64_t N) {
    int64_t blocks = (N + 7) / 8;                 // 8 warps (rays) per 256-thread block
    int64_t cap = (int64_t)ctx->sm_count * 8 * 4;  // a few waves of persistent warps
    if (blocks > cap) blocks = cap;
    return (int)(blocks < 1 ? 1 : blocks);
}

extern "C" int acn_composite_fwd(acn_ctx* ctx, const float* rgb_sigma, const float* t_vals, const float* bg_or_null,
                                 int64_t N, int S, float sigma_scale, float* rgb, float* depth, float* weights,
                                 float* acc, acn_stream stream) {
    ACN_CHECK_CTX(ctx);
    ACN_REQUIRE(N >= 0, ACN_EINVAL, "acn_composite_fwd: negative N");
    ACN_REQUIRE(S >= 2, ACN_EUNSUPPORTED, "acn_composite_fwd: needs at least 2 samples per ray (got %d)", S);
    if (N == 0) return ACN_OK;
    ACN_REQUIRE(rgb_sigma && t_vals && rgb && depth && weights && acc, ACN_EINVAL, "acn_composite_fwd: null buffer");
    ACN_REQUIRE(((uintptr_t)rgb_sigma & 15) == 0, ACN_EINVAL, "acn_composite_fwd: rgb_sigma must be 16-byte aligned");
    k_composite_fwd<<<composite_grid(ctx, N), 256, 0, (cudaStream_t)stream>>>(
        (const float4*)rgb_sigma, t_vals, bg_or_null, N, S, sigma_scale, rgb, depth, weights, acc);
    ACN_CHECK_LAUNCH();
    return ACN_OK;
}

extern "C" int acn_composite_bwd(acn_ctx* ctx, const float* rgb_sigma, const float* t_vals, const float* bg_or_null,
                                 int64_t N, int S, float sigma_scale, const float* g_rgb, const float* g_depth,
                                 const float* g_weights, const float* g_acc, float* d_rgb_sigma, float* d_bg_or_null,
                                 acn_stream stream) {
    ACN_CHECK_CTX(ctx);
    ACN_REQUIRE(N >= 0, ACN_EINVAL, "acn_composite_bwd: negative N");
    ACN_REQUIRE(S >= 2, ACN_EUNSUPPORTED, "acn_composite_bwd: needs at least 2 samples per ray (got %d)", S);
    if (N == 0) return ACN_OK;
    ACN_REQUIRE(rgb_sigma && t_vals && d_rgb_sigma, ACN_EINVAL, "acn_composite_bwd: null buffer");
    ACN_REQUIRE((((uintptr_t)rgb_sigma | (uintptr_t)d_rgb_sigma) & 15) == 0, ACN_EINVAL, "acn_composite_bwd: misaligned buffers");
    k_composite_bwd<<<composite_grid(ctx, N), 256, 0, (cudaStream_t)stream>>>(
        (const float4*)rgb_sigma, t_vals, bg_or_null, N, S, sigma_scale, g_rgb, g_depth, g_weights, g_acc,
        (float4*)d_rgb_sigma, d_bg_or_null);
    ACN_CHECK_LAUNCH();
    return ACN_OK;
}
