// Ray-batch producer: the "dda" routing of data/task_dataset.py:544-627 TaskDataset._route_and_bin
// (the policy nerf_runner.py:207 selects) -- every training ray is assigned to the cell of the task grid it spends the
// longest parametric length in.  The reference runs it as ~64 x 25 elementwise launches over all rays of an expert plus
// an argsort; here it is one thread per ray (all state in registers), per-cell counts aggregated per block, and the
// bins come from the same bucketing kernel the expert dispatch uses.  Integer output: every fp32 operation is rounded
// exactly where torch rounds it (no FMA contraction, IEEE division), NaN handling of torch.minimum / maximum included.
#include "acn_common.cuh"
#include <math.h>

__device__ __forceinline__ float min_nan(float a, float b) { return (a != a || b != b) ? __int_as_float(0x7fc00000) : (a < b ? a : b); }
__device__ __forceinline__ float max_nan(float a, float b) { return (a != a || b != b) ? __int_as_float(0x7fc00000) : (a > b ? a : b); }
__device__ __forceinline__ float nan_to_big(float x) { return (x != x || isinf(x)) ? 1e30f : x; }   // nan_to_num_(1e30, 1e30, 1e30)

// task_dataset.py:125-152 _aabb_intersect (eps = 1e-12), one ray against one box
__device__ __forceinline__ bool slab(const float* o, const float* d, const float* __restrict__ lo, const float* __restrict__ hi,
                                     float& t_entry, float& t_exit)
{
    float te = 0.f, tx = 0.f;
    bool miss_parallel = false;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float l = __ldg(lo + a), h = __ldg(hi + a);
        const bool parallel = fabsf(d[a]) < 1e-12f;
        const float inv_d = __fdiv_rn(1.0f, d[a]);
        const float t0 = __fmul_rn(__fsub_rn(l, o[a]), inv_d), t1 = __fmul_rn(__fsub_rn(h, o[a]), inv_d);
        const float tmin = min_nan(t0, t1), tmax = max_nan(t0, t1);
        miss_parallel |= parallel && !(o[a] >= l && o[a] <= h);
        te = a == 0 ? tmin : max_nan(te, tmin);
        tx = a == 0 ? tmax : min_nan(tx, tmax);
    }
    t_entry = te; t_exit = tx;
    return (tx >= te) && !miss_parallel;
}

__global__ void __launch_bounds__(256) k_dda_route_rays(
    const float* __restrict__ rays8, int64_t N, const float* __restrict__ aabb6, int nx, int ny, int nz,
    const float* __restrict__ cell3, const float* __restrict__ cell_bounds, const float* __restrict__ tol, int max_steps,
    int32_t* __restrict__ cid_out, float* __restrict__ best_len_out, int32_t* __restrict__ counts)
{
    extern __shared__ int s_cnt[];                   // one counter per cell
    const int C = nx * ny * nz;
    for (int c = threadIdx.x; c < C; c += blockDim.x) s_cnt[c] = 0;
    __syncthreads();
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int chosen = -1;
    float best_len = 0.0f;
    if (r < N) {
        const float4 ra = __ldg(reinterpret_cast<const float4*>(rays8 + 8 * r));
        const float4 rb = __ldg(reinterpret_cast<const float4*>(rays8 + 8 * r) + 1);
        const float o[3] = { ra.x, ra.y, ra.z }, d[3] = { ra.w, rb.x, rb.y };
        const float near = rb.z, far = rb.w;
        float te, tx;
        const bool hit = slab(o, d, aabb6, aabb6 + 3, te, tx);                       // :155-172 _region_segment
        const float t0 = max_nan(max_nan(te, 0.0f), near), t1 = min_nan(tx, far);
        if (hit && __fsub_rn(t1, t0) > 0.0f) {
            float tmax[3], tdelta[3];
            int idx[3], step[3];
            const int ncell[3] = { nx, ny, nz };
            const float tstart = __fadd_rn(t0, 1e-6f);
#pragma unroll
            for (int a = 0; a < 3; ++a) {                                              // :239-297 _dda_transform, _dda_init
                const float cs = __ldg(cell3 + a);
                const float go = __fdiv_rn(__fsub_rn(o[a], __ldg(aabb6 + a)), cs), gd = __fdiv_rn(d[a], cs);
                const float p = __fadd_rn(go, __fmul_rn(gd, tstart));
                const float fl = floorf(p);
                step[a] = gd > 0.0f ? 1 : (gd < 0.0f ? -1 : 0);
                const float nb = step[a] > 0 ? __fadd_rn(fl, 1.0f) : __fsub_rn(ceilf(p), 1.0f);
                const float inv = __fdiv_rn(1.0f, gd);
                tmax[a] = nan_to_big(__fmul_rn(__fsub_rn(nb, p), inv));
                tdelta[a] = nan_to_big(__fmul_rn((float)step[a], inv));
                // floor(p).to(int64).clamp(0, n-1): clamp in float first so that the conversion cannot overflow
                idx[a] = (int)fminf(fmaxf(fl, 0.0f), (float)(ncell[a] - 1));
            }
            const int nyz = ny * nz;
            float t = t0;
            int best_cid = idx[0] * nyz + idx[1] * nz + idx[2];
            for (int it = 0; it < max_steps; ++it) {                                   // :299-351 _dda_maxoverlap
                const float m = min_nan(min_nan(tmax[0], tmax[1]), tmax[2]);
                const float t_next = min_nan(m, t1);
                const float dt = fmaxf(__fsub_rn(t_next, t), 0.0f);
                if (dt > best_len) { best_len = dt; best_cid = idx[0] * nyz + idx[1] * nz + idx[2]; }
                if (t_next >= t1) break;
                const bool adv_x = (tmax[0] <= tmax[1]) && (tmax[0] <= tmax[2]);
                const bool adv_y = !(tmax[0] <= tmax[1]) && (tmax[1] <= tmax[2]);
                // no dynamic register indexing: three predicated updates
                if (adv_x) { idx[0] = min(max(idx[0] + step[0], 0), nx - 1); tmax[0] = __fadd_rn(tmax[0], tdelta[0]); }
                else if (adv_y) { idx[1] = min(max(idx[1] + step[1], 0), ny - 1); tmax[1] = __fadd_rn(tmax[1], tdelta[1]); }
                else { idx[2] = min(max(idx[2] + step[2], 0), nz - 1); tmax[2] = __fadd_rn(tmax[2], tdelta[2]); }
                t = t_next;
            }
            const float* cb = cell_bounds + 6 * (size_t)best_cid;                      // :212-228, 590-603: keep or drop
            float ce, cx;
            const bool chit = slab(o, d, cb, cb + 3, ce, cx);
            const float c0 = max_nan(max_nan(ce, 0.0f), near), c1 = min_nan(cx, far);
            float len = __fsub_rn(c1, c0);
            if (len < 0.0f) len = 0.0f;
            if (!chit) len = 0.0f;
            if (len >= __ldg(tol + best_cid)) chosen = best_cid;
        }
        cid_out[r] = chosen;
        if (best_len_out) best_len_out[r] = best_len;
    }
    if (chosen >= 0 && counts) atomicAdd_block(s_cnt + chosen, 1);
    __syncthreads();
    if (counts)
        for (int c = threadIdx.x; c < C; c += blockDim.x)
            if (s_cnt[c]) atomicAdd(counts + c, s_cnt[c]);
}

extern "C" int acn_dda_route_rays(acn_ctx* ctx, const float* rays8, int64_t N, const float* aabb6, int nx, int ny, int nz,
                                  const float* cell3, const float* cell_bounds, const float* tol, int max_steps,
                                  int32_t* cid_out, float* best_len_or_null, int32_t* counts_or_null, acn_stream stream)
{
    ACN_CHECK_CTX(ctx);
    ACN_REQUIRE(N >= 0 && nx >= 1 && ny >= 1 && nz >= 1 && max_steps >= 0, ACN_EINVAL, "acn_dda_route_rays: bad sizes");
    const int64_t C = (int64_t)nx * ny * nz;
    ACN_REQUIRE(C <= 8192, ACN_EUNSUPPORTED, "acn_dda_route_rays: %lld cells > 8192", (long long)C);
    ACN_REQUIRE(aabb6 && cell3 && cell_bounds && tol, ACN_EINVAL, "acn_dda_route_rays: null grid description");
    if (N == 0) return ACN_OK;
    ACN_REQUIRE(rays8 && cid_out && ((uintptr_t)rays8 & 15) == 0, ACN_EINVAL, "acn_dda_route_rays: rays8 / cid_out null or rays8 misaligned");
    k_dda_route_rays<<<acn_grid_1d(N, 256), 256, (size_t)C * sizeof(int), (cudaStream_t)stream>>>(
        rays8, N, aabb6, nx, ny, nz, cell3, cell_bounds, tol, max_steps, cid_out, best_len_or_null, counts_or_null);
    ACN_CHECK_LAUNCH();
    return ACN_OK;
}
