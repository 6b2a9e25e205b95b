// Stage 1: camera rays, scene-box clipping, stratified sampling.
// Every arithmetic step that feeds a bit-exact output (near/far, sample bins, points) uses the
// round-to-nearest intrinsics so nvcc cannot contract a*b+c into an FMA.
#include "acn_common.cuh"

// torch.minimum / torch.maximum propagate NaN; fminf/fmaxf do not.
__device__ __forceinline__ float min_nan(float a, float b) { return (a != a || b != b) ? __int_as_float(0x7fc00000) : fminf(a, b); }
__device__ __forceinline__ float max_nan(float a, float b) { return (a != a || b != b) ? __int_as_float(0x7fc00000) : fmaxf(a, b); }
__device__ __forceinline__ float clamp_nan(float x, float lo, float hi) { return x != x ? x : fminf(fmaxf(x, lo), hi); }

// nerfs/scene_box.py:82-106
__device__ __forceinline__ void slab_test(const float o[3], const float d[3], const float* __restrict__ aabb6,
                                          float eps, float max_bound, float invalid, float& tmn, float& tmx) {
    float lo[3], hi[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float dd = d[a];
        float rd = fabsf(dd) < eps ? (dd >= 0.0f ? eps : -eps) : dd;
        float inv = __fdiv_rn(1.0f, rd);
        float t0 = __fmul_rn(__fsub_rn(__ldg(aabb6 + a), o[a]), inv);
        float t1 = __fmul_rn(__fsub_rn(__ldg(aabb6 + 3 + a), o[a]), inv);
        lo[a] = min_nan(t0, t1);
        hi[a] = max_nan(t0, t1);
    }
    tmn = max_nan(max_nan(lo[0], lo[1]), lo[2]);
    tmx = min_nan(min_nan(hi[0], hi[1]), hi[2]);
    tmn = clamp_nan(tmn, 0.0f, max_bound);
    tmx = clamp_nan(tmx, 0.0f, max_bound);
    bool inval = tmx <= tmn;
    if (inval) { tmn = invalid; tmx = invalid; }
}

// nerfs/ray_sampling.py:111-136
__global__ void k_ray_directions(int H, int W, float fx, float fy, float cx, float cy, int center, float* __restrict__ dirs) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)H * W) return;
    int j = (int)(idx / W), i = (int)(idx % W);
    float fi = (float)i, fj = (float)j;
    if (center) { fi = __fadd_rn(fi, 0.5f); fj = __fadd_rn(fj, 0.5f); }
    float x = __fdiv_rn(__fsub_rn(fi, cx), fx);
    float y = -__fdiv_rn(__fsub_rn(fj, cy), fy);
    float z = -1.0f;
    float n = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), 1.0f));
    n = fmaxf(n, 1e-12f);
    dirs[3 * idx + 0] = __fdiv_rn(x, n);
    dirs[3 * idx + 1] = __fdiv_rn(y, n);
    dirs[3 * idx + 2] = __fdiv_rn(z, n);
}

__global__ void k_aabb_intersect(const float* __restrict__ o, const float* __restrict__ d, int64_t N, int so, int sd,
                                 const float* __restrict__ aabb6, float eps, float max_bound, float invalid,
                                 float* __restrict__ tmin, float* __restrict__ tmax) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= N) return;
    float oo[3] = { o[r * so], o[r * so + 1], o[r * so + 2] };
    float dd[3] = { d[r * sd], d[r * sd + 1], d[r * sd + 2] };
    float a, b;
    slab_test(oo, dd, aabb6, eps, max_bound, invalid, a, b);
    tmin[r] = a;
    tmax[r] = b;
}

// nerfs/ray_sampling.py:10-24, 50-108: one thread builds one packed ray (two float4 stores).
__global__ void k_get_rays(const float* __restrict__ dirs_cam, int64_t N, const float* __restrict__ c2w,
                           const float* __restrict__ aabb6, float near_c, float far_c, float max_bound,
                           float invalid, float* __restrict__ rays8) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= N) return;
    float dc[3] = { dirs_cam[3 * r], dirs_cam[3 * r + 1], dirs_cam[3 * r + 2] };
    float o[3], d[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        o[j] = __ldg(c2w + 4 * j + 3);
        float acc = __fmul_rn(dc[0], __ldg(c2w + 4 * j + 0));
        acc = __fadd_rn(acc, __fmul_rn(dc[1], __ldg(c2w + 4 * j + 1)));
        acc = __fadd_rn(acc, __fmul_rn(dc[2], __ldg(c2w + 4 * j + 2)));
        d[j] = acc;
    }
    float near = near_c, far = far_c;
    if (aabb6) slab_test(o, d, aabb6, 1e-8f, max_bound, invalid, near, far);
    float4* out = reinterpret_cast<float4*>(rays8 + 8 * r);
    out[0] = make_float4(o[0], o[1], o[2], d[0]);
    out[1] = make_float4(d[1], d[2], near, far);
}

// nerfs/ray_sampling.py:139-176
__global__ void k_clamp_near_far(float* __restrict__ rays8, int64_t N, int has_override, float n_ov, float f_ov,
                                 float eps, float invalid, uint8_t* __restrict__ valid) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= N) return;
    float near = rays8[8 * r + 6], far = rays8[8 * r + 7];
    if (has_override) {
        if (n_ov == n_ov) near = max_nan(near, n_ov);
        if (f_ov == f_ov) far = min_nan(far, f_ov);
    }
    bool v = isfinite(near) && isfinite(far) && (far > __fadd_rn(near, eps));
    valid[r] = v ? 1 : 0;
    if (has_override) {
        rays8[8 * r + 6] = v ? near : invalid;
        rays8[8 * r + 7] = v ? far : invalid;
    }
}

// nerfs/ray_rendering.py:279-286.  base(s) = near*(1-u_s) + far*u_s, each op rounded on its own.
__device__ __forceinline__ float strat_base(float near, float far, float u) {
    return __fadd_rn(__fmul_rn(near, __fsub_rn(1.0f, u)), __fmul_rn(far, u));
}

__device__ __forceinline__ float strat_t(float near, float far, const float* __restrict__ u_lin, int s, int S,
                                         const float* __restrict__ jit_row) {
    float b = strat_base(near, far, __ldg(u_lin + s));
    if (!jit_row) return b;
    float lo = (s == 0) ? b : __fmul_rn(0.5f, __fadd_rn(strat_base(near, far, __ldg(u_lin + s - 1)), b));
    float hi = (s == S - 1) ? b : __fmul_rn(0.5f, __fadd_rn(b, strat_base(near, far, __ldg(u_lin + s + 1))));
    return __fadd_rn(lo, __fmul_rn(__fsub_rn(hi, lo), jit_row[s]));
}

__global__ void k_sample_stratified(const float* __restrict__ rays8, int64_t N, int S, const float* __restrict__ u_lin,
                                    const float* __restrict__ jitter, float* __restrict__ t_vals) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * S) return;
    int64_t r = idx / S;
    int s = (int)(idx - r * S);
    float near = __ldg(rays8 + 8 * r + 6), far = __ldg(rays8 + 8 * r + 7);
    t_vals[idx] = strat_t(near, far, u_lin, s, S, jitter ? jitter + r * S : nullptr);
}

// nerfs/ray_rendering.py:317-319
__global__ void k_points(const float* __restrict__ rays8, int64_t N, int S, const float* __restrict__ t_vals,
                         float* __restrict__ id6) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * S) return;
    int64_t r = idx / S;
    float t = t_vals[idx];
    const float* ry = rays8 + 8 * r;
    float* o = id6 + 6 * idx;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float d = __ldg(ry + 3 + a);
        o[a] = __fadd_rn(__ldg(ry + a), __fmul_rn(d, t));
        o[3 + a] = d;
    }
}

// Are consecutive rays neighbouring pixels of a frame?  One warp looks at the first 65 rays: the gap between ray r and
// r+1 at mid depth against the spacing of the samples along the ray; *flag = 1 when more than half of the usable pairs
// (i.e. the median) fall below threshold.  A performance hint for the gather kernels' warp mapping -- never changes results.
__global__ void k_rays_coherent(const float* __restrict__ rays8, int64_t N, int S, float threshold, int32_t* __restrict__ flag) {
    const int lane = threadIdx.x;
    const int pairs = (int)(N < 65 ? N : 65) - 1;
    int valid = 0, good = 0;
    for (int i = lane; i < pairs; i += 32) {
        const float* a = rays8 + 8 * (int64_t)i;
        const float* b = a + 8;
        const float tm = 0.5f * (a[6] + a[7]);
        const float gx = (b[0] - a[0]) + (b[3] - a[3]) * tm, gy = (b[1] - a[1]) + (b[4] - a[4]) * tm,
                    gz = (b[2] - a[2]) + (b[5] - a[5]) * tm;
        const float gap = sqrtf(gx * gx + gy * gy + gz * gz);
        const float step = fabsf(a[7] - a[6]) / (float)S;
        const bool ok = isfinite(gap) && isfinite(step) && step > 0.0f;
        valid += ok;
        good += ok && gap < threshold * step;
    }
    for (int d = 16; d > 0; d >>= 1) {
        valid += __shfl_xor_sync(0xffffffffu, valid, d);
        good += __shfl_xor_sync(0xffffffffu, good, d);
    }
    if (lane == 0) *flag = (N >= 33 && valid > 0 && 2 * good > valid) ? 1 : 0;
}

// ------------------------------------------------------------------------------------------ C ABI
extern "C" int acn_ray_directions(acn_ctx* ctx, int H, int W, float fx, float fy, float cx, float cy,
                                  int center_pixels, float* dirs, acn_stream stream) {
    ACN_CHECK_CTX(ctx);
    ACN_REQUIRE(H >= 0 && W >= 0 && dirs, ACN_EINVAL, "acn_ray_directions: bad arguments");
    int64_t n = (int64_t)H * W;
    if (n == 0) return ACN_OK;
    k_ray_directions<<<acn_grid_1d(n, 256), 256, 0, (cudaStream_t)stream>>>(H, W, fx, fy, cx, cy, center_pixels, dirs);
    ACN_CHECK_LAUNCH();
    return ACN_OK;
}

extern "C" int acn_aabb_intersect(acn_ctx* ctx, const float* o, const float* d, int64_t N, int stride_o, int stride_d,
                                  const float* aabb6, float eps, float max_bound, float invalid, float* tmin,
                                  float* tmax, acn_stream stream) {
    ACN_CHECK_CTX(ctx);
    ACN_REQUIRE(N >= 0 && stride_o >= 3 && stride_d >= 3 && aabb6, ACN_EINVAL, "acn_aabb_intersect: bad arguments");
    if (N == 0) return ACN_OK;
    ACN_REQUIRE(o && d && tmin && tmax, ACN_EINVAL, "acn_aabb_intersect: null buffer");
    k_aabb_intersect<<<acn_grid_1d(N, 256), 256, 0, (cudaStream_t)stream>>>(o, d, N, stride_o, stride_d, aabb6, eps,
                                                                            max_bound, invalid, tmin, tmax);
    ACN_CHECK_LAUNCH();
    return ACN_OK;
}

extern "C" int acn_get_rays(acn_ctx* ctx, const float* dirs_cam, int64_t N, const float* c2w, const float* aabb6_or_null,
                            float near_c, float far_c, float max_bound, float invalid, float* rays8, acn_stream stream) {
    ACN_CHECK_CTX(ctx);
    ACN_REQUIRE(N >= 0 && c2w, ACN_EINVAL, "acn_get_rays: bad arguments");
    if (N == 0) return ACN_OK;
    ACN_REQUIRE(dirs_cam && rays8, ACN_EINVAL, "acn_get_rays: null buffer");
    ACN_REQUIRE(((uintptr_t)rays8 & 15) == 0, ACN_EINVAL, "acn_get_rays: rays8 must be 16-byte aligned");
    k_get_rays<<<acn_grid_1d(N, 256), 256, 0, (cudaStream_t)stream>>>(dirs_cam, N, c2w, aabb6_or_null, near_c, far_c,
                                                                      max_bound, invalid, rays8);
    ACN_CHECK_LAUNCH();
    return ACN_OK;
}

extern "C" int acn_clamp_near_far(acn_ctx* ctx, float* rays8, int64_t N, int has_override, float n_or_nan, float f_or_nan,
                                  float eps, float invalid, uint8_t* valid, acn_stream stream) {
    ACN_CHECK_CTX(ctx);
    ACN_REQUIRE(N >= 0, ACN_EINVAL, "acn_clamp_near_far: negative N");
    if (N == 0) return ACN_OK;
    ACN_REQUIRE(rays8 && valid, ACN_EINVAL, "acn_clamp_near_far: null buffer");
    k_clamp_near_far<<<acn_grid_1d(N, 256), 256, 0, (cudaStream_t)stream>>>(rays8, N, has_override, n_or_nan, f_or_nan,
                                                                            eps, invalid, valid);
    ACN_CHECK_LAUNCH();
    return ACN_OK;
}

extern "C" int acn_sample_stratified(acn_ctx* ctx, const float* rays8, int64_t N, int S, const float* u_lin,
                                     const float* jitter_or_null, float* t_vals, acn_stream stream) {
    ACN_CHECK_CTX(ctx);
    ACN_REQUIRE(N >= 0 && S >= 1 && u_lin, ACN_EINVAL, "acn_sample_stratified: bad arguments");
    if (N == 0) return ACN_OK;
    ACN_REQUIRE(rays8 && t_vals, ACN_EINVAL, "acn_sample_stratified: null buffer");
    k_sample_stratified<<<acn_grid_1d(N * S, 256), 256, 0, (cudaStream_t)stream>>>(rays8, N, S, u_lin, jitter_or_null, t_vals);
    ACN_CHECK_LAUNCH();
    return ACN_OK;
}

extern "C" int acn_points(acn_ctx* ctx, const float* rays8, int64_t N, int S, const float* t_vals, float* id6,
                          acn_stream stream) {
    ACN_CHECK_CTX(ctx);
    ACN_REQUIRE(N >= 0 && S >= 1, ACN_EINVAL, "acn_points: bad arguments");
    if (N == 0) return ACN_OK;
    ACN_REQUIRE(rays8 && t_vals && id6, ACN_EINVAL, "acn_points: null buffer");
    k_points<<<acn_grid_1d(N * S, 256), 256, 0, (cudaStream_t)stream>>>(rays8, N, S, t_vals, id6);
    ACN_CHECK_LAUNCH();
    return ACN_OK;
}

extern "C" int acn_rays_coherent(acn_ctx* ctx, const float* rays8, int64_t N, int S, float threshold, int32_t* flag,
                                 acn_stream stream) {
    ACN_CHECK_CTX(ctx);
    ACN_REQUIRE(N >= 0 && S >= 1 && flag && (N == 0 || rays8), ACN_EINVAL, "acn_rays_coherent: bad arguments");
    k_rays_coherent<<<1, 32, 0, (cudaStream_t)stream>>>(rays8, N, S, threshold, flag);
    ACN_CHECK_LAUNCH();
    return ACN_OK;
}
