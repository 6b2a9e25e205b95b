// Stage 5: Voronoi routing of samples (models/inr/meta_container.py:97-134) and of rays
// (scripts/create_clusters.py:559-634), plus the device-side dispatch that replaces the
// reference's per-expert nonzero()/index_select()/index_add_() round trips.
//
// Distances follow torch.cdist's matmul formulation rounding step by step, so the
// hard assignment, the soft support set and the ray masks are bit-exact w.r.t. the CPU reference.
#include "acn_common.cuh"

#define FULL 0xffffffffu

template <int DIMS>
__device__ __forceinline__ float cdist_mm(const float* x, const float* __restrict__ c) {
    float xn = __fmul_rn(x[0], x[0]), cn = __fmul_rn(__ldg(c), __ldg(c));
#pragma unroll
    for (int k = 1; k < DIMS; ++k) {
        xn = __fadd_rn(xn, __fmul_rn(x[k], x[k]));
        cn = __fadd_rn(cn, __fmul_rn(__ldg(c + k), __ldg(c + k)));
    }
    float acc = __fmul_rn(__fmul_rn(-2.0f, x[0]), __ldg(c));
#pragma unroll
    for (int k = 1; k < DIMS; ++k) acc = __fmaf_rn(__fmul_rn(-2.0f, x[k]), __ldg(c + k), acc);
    acc = __fadd_rn(xn, acc);
    acc = __fadd_rn(acc, cn);
    return __fsqrt_rn(fmaxf(acc, 0.0f));
}

template <int DIMS>
__global__ void __launch_bounds__(256) k_route_points(
    const float* __restrict__ pts, int64_t P, int stride, const float* __restrict__ cen, int K, float margin,
    float* __restrict__ weights, int32_t* __restrict__ hard, int32_t* __restrict__ counts)
{
    constexpr int OFF = DIMS == 2 ? 1 : 0;
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool on = p < P;
    const int lane = threadIdx.x & 31;
    float x[3] = { 0.f, 0.f, 0.f };
    if (on) {
#pragma unroll
        for (int k = 0; k < DIMS; ++k) x[k] = pts[p * stride + OFF + k];
    }
    if (weights) {
        // pass 1: min distance (after the 1e-6 floor); pass 2: masked 1/d sum; pass 3: normalise
        float mind = __int_as_float(0x7f800000);
        for (int k = 0; k < K; ++k) mind = fminf(mind, fmaxf(cdist_mm<DIMS>(x, cen + 3 * k + OFF), 1e-6f));
        const float thr = __fmul_rn(margin, mind);
        float denom = 0.0f;
        for (int k = 0; k < K; ++k) {
            float d = fmaxf(cdist_mm<DIMS>(x, cen + 3 * k + OFF), 1e-6f);
            float inv = d <= thr ? __fdiv_rn(1.0f, d) : 0.0f;
            denom = __fadd_rn(denom, inv);
        }
        denom = fmaxf(denom, 1e-6f);
        for (int k = 0; k < K; ++k) {
            float d = fmaxf(cdist_mm<DIMS>(x, cen + 3 * k + OFF), 1e-6f);
            bool in = on && d <= thr;
            if (on) weights[p * K + k] = in ? __fdiv_rn(__fdiv_rn(1.0f, d), denom) : 0.0f;
            if (counts) {
                unsigned m = __ballot_sync(FULL, in);
                if (lane == 0 && m) atomicAdd(counts + k, __popc(m));
            }
        }
    } else {
        int best = 0;
        float bd = cdist_mm<DIMS>(x, cen + OFF);
        for (int k = 1; k < K; ++k) {
            float d = cdist_mm<DIMS>(x, cen + 3 * k + OFF);
            if (d < bd) { bd = d; best = k; }   // first minimum wins, like argmin
        }
        if (on) hard[p] = best;
        if (counts) {
            for (int k = 0; k < K; ++k) {
                unsigned m = __ballot_sync(FULL, on && best == k);
                if (lane == 0 && m) atomicAdd(counts + k, __popc(m));
            }
        }
    }
}

// scripts/create_clusters.py:592-632: per ray, S lerp'ed samples, min over samples of
// D / (min_c D + 1e-8), compared with the margin.
template <int DIMS, int MAXK>
__global__ void __launch_bounds__(128) k_route_rays(
    const float* __restrict__ rays8, int64_t N, int S, const float* __restrict__ u_lin,
    const float* __restrict__ cen, int K, float margin, uint8_t* __restrict__ mask)
{
    constexpr int OFF = DIMS == 2 ? 1 : 0;
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= N) return;
    const float4 a = __ldg(reinterpret_cast<const float4*>(rays8 + 8 * r));
    const float4 b = __ldg(reinterpret_cast<const float4*>(rays8 + 8 * r) + 1);
    const float o[3] = { a.x, a.y, a.z }, d[3] = { a.w, b.x, b.y };
    const float near = b.z, far = b.w;
    const float diff = __fsub_rn(far, near);
    float rmin[MAXK];
#pragma unroll
    for (int k = 0; k < MAXK; ++k) rmin[k] = __int_as_float(0x7f800000);
    for (int s = 0; s < S; ++s) {
        float z = __ldg(u_lin + s);
        // torch.lerp on CPU is fused: w<0.5 ? a + w*(b-a) : b - (b-a)*(1-w)
        float t = z < 0.5f ? __fmaf_rn(z, diff, near) : __fmaf_rn(-diff, __fsub_rn(1.0f, z), far);
        float x[3];
#pragma unroll
        for (int k = 0; k < DIMS; ++k) x[k] = __fadd_rn(o[OFF + k], __fmul_rn(d[OFF + k], t));
        float D[MAXK];
        float m = __int_as_float(0x7f800000);
#pragma unroll
        for (int k = 0; k < MAXK; ++k) {
            if (k < K) { D[k] = cdist_mm<DIMS>(x, cen + 3 * k + OFF); m = fminf(m, D[k]); }
        }
        float den = __fadd_rn(m, 1e-8f);
#pragma unroll
        for (int k = 0; k < MAXK; ++k) {
            if (k < K) rmin[k] = fminf(rmin[k], __fdiv_rn(D[k], den));
        }
    }
#pragma unroll
    for (int k = 0; k < MAXK; ++k) {
        if (k < K) mask[r * K + k] = rmin[k] <= margin ? 1 : 0;
    }
}

// Device-side bucket: warp-aggregated cursor claim per expert.
__global__ void __launch_bounds__(256) k_bucket_points(
    const float* __restrict__ id6, int64_t P, const float* __restrict__ weights, const int32_t* __restrict__ hard,
    int K, const int32_t* __restrict__ offsets, int32_t* __restrict__ cursor, int32_t* __restrict__ sel,
    float* __restrict__ xd_out, float* __restrict__ w_out)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool on = p < P;
    const int lane = threadIdx.x & 31;
    const int h = (on && hard) ? hard[p] : -1;
    for (int k = 0; k < K; ++k) {
        float w = weights ? (on ? weights[p * K + k] : 0.0f) : (h == k ? 1.0f : 0.0f);
        bool in = on && w > 0.0f;
        unsigned m = __ballot_sync(FULL, in);
        if (!m) continue;
        int base = 0;
        int leader = __ffs(m) - 1;
        if (lane == leader) base = atomicAdd(cursor + k, __popc(m));
        base = __shfl_sync(FULL, base, leader);
        if (in) {
            int slot = __ldg(offsets + k) + base + __popc(m & ((1u << lane) - 1u));
            sel[slot] = (int32_t)p;
            if (w_out) w_out[slot] = w;
            if (xd_out) {
#pragma unroll
                for (int c = 0; c < 6; ++c) xd_out[(size_t)slot * 6 + c] = id6[p * 6 + c];
            }
        }
    }
}

// Fused dispatch for expert sharding: the same bucketing, but the routed [xyz, dir] row of expert k is stored straight
// into the receive buffer of the GPU that owns k -- row_base[k] is that buffer's address as mapped into this process
// (NVLink peer memory, or local memory for the experts this rank owns), row_off[k] the first row reserved there for
// (this source rank, expert k).  sel / w_out stay local.  The dispatch half of the all-to-all is therefore the
// kernel's own store stream; no staging copy, no NCCL send.
__global__ void __launch_bounds__(256) k_dispatch_points(
    const float* __restrict__ id6, int64_t P, const float* __restrict__ weights, const int32_t* __restrict__ hard,
    int K, const int32_t* __restrict__ offsets, int32_t* __restrict__ cursor, int32_t* __restrict__ sel,
    float* __restrict__ w_out, const unsigned long long* __restrict__ row_base, const int32_t* __restrict__ row_off)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool on = p < P;
    const int lane = threadIdx.x & 31;
    const int h = (on && hard) ? hard[p] : -1;
    float2 r0 = make_float2(0.f, 0.f), r1 = r0, r2 = r0;
    if (on) {
        const float2* src = reinterpret_cast<const float2*>(id6 + p * 6);
        r0 = __ldg(src); r1 = __ldg(src + 1); r2 = __ldg(src + 2);
    }
    for (int k = 0; k < K; ++k) {
        float w = weights ? (on ? weights[p * K + k] : 0.0f) : (h == k ? 1.0f : 0.0f);
        bool in = on && w > 0.0f;
        unsigned m = __ballot_sync(FULL, in);
        if (!m) continue;
        int base = 0;
        int leader = __ffs(m) - 1;
        if (lane == leader) base = atomicAdd(cursor + k, __popc(m));
        base = __shfl_sync(FULL, base, leader);
        if (in) {
            const int idx = base + __popc(m & ((1u << lane) - 1u));
            const int slot = __ldg(offsets + k) + idx;
            sel[slot] = (int32_t)p;
            w_out[slot] = w;
            float2* dst = reinterpret_cast<float2*>(__ldg(row_base + k)) + ((size_t)__ldg(row_off + k) + idx) * 3;
            dst[0] = r0; dst[1] = r1; dst[2] = r2;
        }
    }
}

__global__ void k_blend_add(const float4* __restrict__ y, const float* __restrict__ w, const int32_t* __restrict__ sel,
                            int64_t M, float4* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    float4 v = y[i];
    float ww = w ? w[i] : 1.0f;
    int32_t p = sel[i];
    float4 o = out[p];
    o.x += v.x * ww; o.y += v.y * ww; o.z += v.z * ww; o.w += v.w * ww;
    out[p] = o;
}

__global__ void k_blend_bwd(const float4* __restrict__ d_out, const float* __restrict__ w, const int32_t* __restrict__ sel,
                            int64_t M, float4* __restrict__ d_y) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    float4 g = d_out[sel[i]];
    float ww = w ? w[i] : 1.0f;
    d_y[i] = make_float4(g.x * ww, g.y * ww, g.z * ww, g.w * ww);
}

// ------------------------------------------------------------------------------------------ C ABI
extern "C" int acn_route_points(acn_ctx* ctx, const float* pts, int64_t P, int stride, const float* centroids, int K,
                                int dims, float margin, float* weights_or_null, int32_t* hard_or_null,
                                int32_t* counts_or_null, acn_stream stream) {
    ACN_CHECK_CTX(ctx);
    ACN_REQUIRE(P >= 0 && stride >= 3 && centroids && K >= 1, ACN_EINVAL, "acn_route_points: bad arguments");
    ACN_REQUIRE(dims == 2 || dims == 3, ACN_EINVAL, "acn_route_points: dims must be 2 (cluster_2d) or 3");
    ACN_REQUIRE(margin >= 1.0f, ACN_EINVAL, "acn_route_points: boundary_margin must be >= 1");
    const bool soft = margin > 1.0f;
    ACN_REQUIRE(soft ? weights_or_null != nullptr : hard_or_null != nullptr, ACN_EINVAL,
                "acn_route_points: margin %s needs the %s output", soft ? "> 1" : "== 1", soft ? "weights" : "hard");
    if (P == 0) return ACN_OK;
    ACN_REQUIRE(pts, ACN_EINVAL, "acn_route_points: null points");
    const int grid = acn_grid_1d(P, 256);
    cudaStream_t st = (cudaStream_t)stream;
    float* w = soft ? weights_or_null : nullptr;
    if (dims == 2) k_route_points<2><<<grid, 256, 0, st>>>(pts, P, stride, centroids, K, margin, w, hard_or_null, counts_or_null);
    else k_route_points<3><<<grid, 256, 0, st>>>(pts, P, stride, centroids, K, margin, w, hard_or_null, counts_or_null);
    ACN_CHECK_LAUNCH();
    return ACN_OK;
}

extern "C" int acn_route_rays_voronoi(acn_ctx* ctx, const float* rays8, int64_t N, int S, const float* u_lin,
                                      const float* centroids, int K, int dims, float margin, uint8_t* mask,
                                      acn_stream stream) {
    ACN_CHECK_CTX(ctx);
    ACN_REQUIRE(N >= 0 && S >= 1 && u_lin && centroids, ACN_EINVAL, "acn_route_rays_voronoi: bad arguments");
    ACN_REQUIRE(K >= 1 && K <= 64, ACN_EUNSUPPORTED, "acn_route_rays_voronoi: K=%d outside [1,64]", K);
    ACN_REQUIRE(dims == 2 || dims == 3, ACN_EINVAL, "acn_route_rays_voronoi: dims must be 2 or 3");
    if (N == 0) return ACN_OK;
    ACN_REQUIRE(rays8 && mask, ACN_EINVAL, "acn_route_rays_voronoi: null buffer");
    ACN_REQUIRE(((uintptr_t)rays8 & 15) == 0, ACN_EINVAL, "acn_route_rays_voronoi: rays8 must be 16-byte aligned");
    const int grid = acn_grid_1d(N, 128);
    cudaStream_t st = (cudaStream_t)stream;
#define RR(D, MK) k_route_rays<D, MK><<<grid, 128, 0, st>>>(rays8, N, S, u_lin, centroids, K, margin, mask)
    if (dims == 2) { if (K <= 4) RR(2, 4); else if (K <= 8) RR(2, 8); else if (K <= 16) RR(2, 16); else RR(2, 64); }
    else           { if (K <= 4) RR(3, 4); else if (K <= 8) RR(3, 8); else if (K <= 16) RR(3, 16); else RR(3, 64); }
#undef RR
    ACN_CHECK_LAUNCH();
    return ACN_OK;
}

extern "C" int acn_bucket_points(acn_ctx* ctx, const float* id6, int64_t P, const float* weights_or_null,
                                 const int32_t* hard_or_null, int K, const int32_t* offsets, int32_t* cursor,
                                 int32_t* sel, float* xd_out, float* w_out, acn_stream stream) {
    ACN_CHECK_CTX(ctx);
    ACN_REQUIRE(P >= 0 && K >= 1 && offsets && cursor && sel, ACN_EINVAL, "acn_bucket_points: bad arguments");
    ACN_REQUIRE((weights_or_null != nullptr) != (hard_or_null != nullptr), ACN_EINVAL,
                "acn_bucket_points: give exactly one of weights / hard");
    ACN_REQUIRE(!xd_out || id6, ACN_EINVAL, "acn_bucket_points: xd_out needs id6");
    if (P == 0) return ACN_OK;
    k_bucket_points<<<acn_grid_1d(P, 256), 256, 0, (cudaStream_t)stream>>>(id6, P, weights_or_null, hard_or_null, K, offsets,
                                                                           cursor, sel, xd_out, w_out);
    ACN_CHECK_LAUNCH();
    return ACN_OK;
}

extern "C" int acn_dispatch_points(acn_ctx* ctx, const float* id6, int64_t P, const float* weights_or_null,
                                   const int32_t* hard_or_null, int K, const int32_t* offsets, int32_t* cursor, int32_t* sel,
                                   float* w_out, const uint64_t* row_base, const int32_t* row_off, acn_stream stream) {
    ACN_CHECK_CTX(ctx);
    ACN_REQUIRE(P >= 0 && K >= 1, ACN_EINVAL, "acn_dispatch_points: bad arguments");
    ACN_REQUIRE((weights_or_null != nullptr) != (hard_or_null != nullptr), ACN_EINVAL,
                "acn_dispatch_points: exactly one of weights / hard must be given");
    ACN_REQUIRE(offsets && cursor && sel && w_out && row_base && row_off, ACN_EINVAL, "acn_dispatch_points: null buffer");
    if (P == 0) return ACN_OK;
    ACN_REQUIRE(id6 && ((uintptr_t)id6 & 7) == 0, ACN_EINVAL, "acn_dispatch_points: id6 null or not 8-byte aligned");
    k_dispatch_points<<<acn_grid_1d(P, 256), 256, 0, (cudaStream_t)stream>>>(id6, P, weights_or_null, hard_or_null, K, offsets, cursor,
                                                                             sel, w_out, (const unsigned long long*)row_base, row_off);
    ACN_CHECK_LAUNCH();
    return ACN_OK;
}

extern "C" int acn_blend_add(acn_ctx* ctx, const float* y, const float* w, const int32_t* sel, int64_t M, float* out,
                             acn_stream stream) {
    ACN_CHECK_CTX(ctx);
    ACN_REQUIRE(M >= 0, ACN_EINVAL, "acn_blend_add: negative M");
    if (M == 0) return ACN_OK;
    ACN_REQUIRE(y && sel && out, ACN_EINVAL, "acn_blend_add: null buffer");
    k_blend_add<<<acn_grid_1d(M, 256), 256, 0, (cudaStream_t)stream>>>((const float4*)y, w, sel, M, (float4*)out);
    ACN_CHECK_LAUNCH();
    return ACN_OK;
}

extern "C" int acn_blend_bwd(acn_ctx* ctx, const float* d_out, const float* w, const int32_t* sel, int64_t M, float* d_y,
                             acn_stream stream) {
    ACN_CHECK_CTX(ctx);
    ACN_REQUIRE(M >= 0, ACN_EINVAL, "acn_blend_bwd: negative M");
    if (M == 0) return ACN_OK;
    ACN_REQUIRE(d_out && sel && d_y, ACN_EINVAL, "acn_blend_bwd: null buffer");
    k_blend_bwd<<<acn_grid_1d(M, 256), 256, 0, (cudaStream_t)stream>>>((const float4*)d_out, w, sel, M, (float4*)d_y);
    ACN_CHECK_LAUNCH();
    return ACN_OK;
}
