// What surrounds the render in a training step: the colour-space + MSE loss epilogue
// (nerfs/color_space.py:4-66, nerfs/losses.py:9-32) and the optimizer tail of maml_meta_update / runtime_adapt
// (pipelines/offline_stage/meta_core.py:123-141,181-190; pipelines/online_stage/runtime_adapt.py:262-268):
// GradScaler.unscale_ -> clip_grad_norm_ -> Adam/AdamW step, skipped when a gradient is not finite.
//
// All of it is HBM-bound elementwise work over the 64 MiB hash table (+ 13 715 MLP weights):
//   torch:  unscale (r+w g) + norm (r g) + clip (r+w g) + Adam (r p,g,m,v; w p,m,v) = 12 passes over the table
//   here :  norm (r g) + Adam (r p,g,m,v; w p,m,v)                                  =  8 passes, 3 launches,
// and nothing is read back to the host: the step decision (skip / clip coefficient / bias corrections) is taken by
// a one-block kernel and consumed from device memory.
#include "acn_common.cuh"
#include <math.h>

#define FULL 0xffffffffu

// m, v (and g when it is written back) are read and then overwritten by the same thread: not .nc
__device__ __forceinline__ float4 ld_once_f4(const float4* p) {
    float4 v;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ double block_sum_d(double v, double* sh) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = lane < nw ? sh[lane] : 0.0;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
    }
    return v;  // valid in warp 0
}

// ------------------------------------------------------------------------------------------------ loss epilogue
// torch.clamp propagates NaN; fminf/fmaxf would swallow it.
__device__ __forceinline__ float clamp01_nan(float x) { return x != x ? x : fminf(fmaxf(x, 0.0f), 1.0f); }

// color_space.py:13-19.  Tensor / python-scalar is evaluated by torch as a multiplication with the rounded
// reciprocal, on CPU and on CUDA alike.
__device__ __forceinline__ float srgb_to_linear_f(float x) {
    return x <= 0.04045f ? x * (1.0f / 12.92f) : powf((x + 0.055f) * (1.0f / 1.055f), 2.4f);
}

// Transformed prediction, transformed target and d(pred')/d(pred) for one element (color_space.py:22-66).
template <int CS>
__device__ __forceinline__ void color_pair(float p, float g, float& pp, float& gg, float& dpp) {
    const float g01 = clamp01_nan(g);
    if (CS == ACN_COLOR_LINEAR) {
        pp = clamp01_nan(p);
        dpp = (p >= 0.0f && p <= 1.0f) ? 1.0f : 0.0f;
        gg = clamp01_nan(srgb_to_linear_f(g01));
    } else if (CS == ACN_COLOR_SRGB) {
        const float x = clamp01_nan(p);
        const bool lin = x <= 0.0031308f;
        const float e = (float)(1.0 / 2.4);
        const float y = lin ? 12.92f * x : 1.055f * powf(x, e) - 0.055f;
        // The reference differentiates BOTH branches of torch.where, so its pow branch turns the gradient at
        // pred == 0 into 0 * inf = NaN; here the selected (linear) branch's slope is used instead.
        const float dy = lin ? 12.92f : 1.055f * e * powf(x, e - 1.0f);
        pp = clamp01_nan(y);
        dpp = (p >= 0.0f && p <= 1.0f && y >= 0.0f && y <= 1.0f) ? dy : 0.0f;
        gg = g01;
    } else {
        pp = p;
        dpp = 1.0f;
        gg = g01;
    }
}

template <int CS>
__global__ void __launch_bounds__(256) k_color_mse(const float* __restrict__ pred, const float* __restrict__ gt,
                                                   int64_t n, float inv_n, float* __restrict__ elem,
                                                   float* __restrict__ dpred, double* __restrict__ partial,
                                                   float* __restrict__ loss)
{
    __shared__ double sh[8];
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float pp, gg, dpp;
        color_pair<CS>(pred[i], gt[i], pp, gg, dpp);
        const float d = pp - gg;
        const float se = d * d;
        if (elem) elem[i] = se;
        if (dpred) dpred[i] = 2.0f * d * dpp * inv_n;
        acc += (double)se;
    }
    acc = block_sum_d(acc, sh);
    if (threadIdx.x == 0) {
        if (gridDim.x == 1) { if (loss) *loss = (float)(acc * (double)inv_n); }
        else if (partial) partial[blockIdx.x] = acc;          // reduction="none" passes no workspace
    }
}

// Fixed-order sum of the block partials: the loss does not depend on block scheduling.
__global__ void __launch_bounds__(256) k_sum_partials(const double* __restrict__ partial, int count, float inv_n,
                                                      float* __restrict__ loss)
{
    __shared__ double sh[8];
    double acc = 0.0;
    for (int i = threadIdx.x; i < count; i += blockDim.x) acc += partial[i];
    acc = block_sum_d(acc, sh);
    if (threadIdx.x == 0) *loss = (float)(acc * (double)inv_n);
}

extern "C" int acn_color_mse(acn_ctx* ctx, const float* pred, const float* gt, int64_t n, int color_space,
                             int mean, float* loss_or_null, float* elem_or_null, float* dpred_or_null,
                             double* partial, acn_stream stream_)
{
    ACN_CHECK_CTX(ctx);
    ACN_REQUIRE(n >= 0, ACN_EINVAL, "acn_color_mse: negative size");
    ACN_REQUIRE(color_space >= ACN_COLOR_LINEAR && color_space <= ACN_COLOR_IDENTITY, ACN_EINVAL,
                "acn_color_mse: color_space must be ACN_COLOR_LINEAR, _SRGB or _IDENTITY (got %d)", color_space);
    ACN_REQUIRE(n == 0 || (pred && gt), ACN_EINVAL, "acn_color_mse: null input");
    ACN_REQUIRE(!loss_or_null || partial, ACN_EINVAL, "acn_color_mse: the reduced loss needs the partial workspace");
    cudaStream_t stream = (cudaStream_t)stream_;
    const float inv_n = mean ? (n > 0 ? (float)(1.0 / (double)n) : NAN) : 1.0f;   // mean over nothing = nan, like torch
    const int grid = acn_grid_1d(n, 256 * 8, ACN_LOSS_PARTIALS);
    switch (color_space) {
#define GO(CS) k_color_mse<CS><<<grid, 256, 0, stream>>>(pred, gt, n, inv_n, elem_or_null, dpred_or_null, partial, loss_or_null)
        case ACN_COLOR_LINEAR: GO(ACN_COLOR_LINEAR); break;
        case ACN_COLOR_SRGB: GO(ACN_COLOR_SRGB); break;
        default: GO(ACN_COLOR_IDENTITY); break;
#undef GO
    }
    ACN_CHECK_LAUNCH();
    if (grid > 1 && loss_or_null) {
        k_sum_partials<<<1, 256, 0, stream>>>(partial, grid, inv_n, loss_or_null);
        ACN_CHECK_LAUNCH();
    }
    return ACN_OK;
}

// ------------------------------------------------------------------------------------------------ optimizer tail
struct AdamBatch {
    acn_adam_tensor t[ACN_ADAM_MAX_TENSORS];
    int count;
};

// Sum of squares of the UNSCALED gradients (g / grad_scale) and the number of non-finite elements, added into
// acc[0], acc[1] (double).  Every block walks every tensor grid-stride: the table gradient dominates and the 14
// small tensors cost a few predicated iterations.
__global__ void __launch_bounds__(256) k_grad_sqnorm(const __grid_constant__ AdamBatch b,
                                                     const float* __restrict__ grad_scale, double* __restrict__ acc)
{
    __shared__ double sh[8];
    const float inv = grad_scale ? 1.0f / __ldg(grad_scale) : 1.0f;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int bad = 0;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    for (int k = 0; k < b.count; ++k) {
        const float* g = b.t[k].g;
        const int64_t n = b.t[k].n;
        const float nz0 = s0 + s1 + s2 + s3;               // sums of squares only grow: a change = this tensor has a non-zero
        bool nzt = false;                                  // (squares of tiny values underflow: test the values themselves)
        if (((uintptr_t)g & 15) == 0) {
            const float4* g4 = reinterpret_cast<const float4*>(g);
            const int64_t n4 = n >> 2;
            for (int64_t i = tid; i < n4; i += nth) {
                float4 v = ld_stream_f4(g4 + i);
                nzt |= (v.x != 0.0f) | (v.y != 0.0f) | (v.z != 0.0f) | (v.w != 0.0f);
                v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
                bad |= !isfinite(v.x) | !isfinite(v.y) | !isfinite(v.z) | !isfinite(v.w);
                s0 = fmaf(v.x, v.x, s0); s1 = fmaf(v.y, v.y, s1); s2 = fmaf(v.z, v.z, s2); s3 = fmaf(v.w, v.w, s3);
            }
            for (int64_t i = (n4 << 2) + tid; i < n; i += nth) {
                nzt |= g[i] != 0.0f;
                float v = g[i] * inv;
                bad |= !isfinite(v);
                s0 = fmaf(v, v, s0);
            }
        } else {
            for (int64_t i = tid; i < n; i += nth) {
                nzt |= g[i] != 0.0f;
                float v = g[i] * inv;
                bad |= !isfinite(v);
                s0 = fmaf(v, v, s0);
            }
        }
        (void)nz0;
        if (b.t[k].bias && __syncthreads_or(nzt) && threadIdx.x == 0) b.t[k].bias[2] = 1.0;   // benign race: every writer stores 1
    }
    double s = block_sum_d((double)s0 + (double)s1 + (double)s2 + (double)s3, sh);
    const int any_bad = __syncthreads_or(bad);
    if (threadIdx.x == 0) {
        atomicAdd(acc, s);
        if (any_bad) atomicAdd(acc + 1, 1.0);
    }
}

// Step decision (one thread): GradScaler.step's skip, clip_grad_norm_'s coefficient, Adam's bias corrections.
// state (doubles) = [step, coef, skip, bias_correction1, sqrt(bias_correction2), total_norm, 0, 0]; found_inf_out
// (float, optional) is what GradScaler.update() consumes.  acc is cleared for the next step.
// Per-tensor part of the decision: a tensor that takes part in a step that is not skipped advances ITS step count
// (torch.optim.Adam keeps `step` per parameter and leaves it alone while the parameter's grad is None -- an expert that
// saw no ray) and gets its bias corrections from it, in doubles like torch's host code.
// skip_zero: a tensor whose gradient is identically zero this step is treated like torch treats `grad is None` -- left
// alone, step count included.  That is how the sync-free routed path says "this expert received no row": its autograd
// node cannot return None without reading the row count back to the host, it returns zeros.
__device__ __forceinline__ void adam_advance_tensors(const AdamBatch& b, bool skip, double beta1, double beta2, bool skip_zero) {
    for (int k = threadIdx.x; k < b.count; k += blockDim.x) {
        const acn_adam_tensor& t = b.t[k];
        if (!t.step || !t.bias) continue;
        const bool active = !(skip_zero && t.bias[2] == 0.0);
        t.bias[2] = 0.0;                                   // the non-zero mark of acn_grad_sqnorm is consumed
        t.bias[3] = active ? 1.0 : 0.0;
        double step = (double)*t.step;
        if (!skip && active) { step += 1.0; *t.step = (float)step; }
        t.bias[0] = 1.0 - pow(beta1, step);
        t.bias[1] = sqrt(1.0 - pow(beta2, step));
    }
}

__global__ void __launch_bounds__(64) k_adam_advance(const __grid_constant__ AdamBatch b, const double* __restrict__ state,
                                                     double beta1, double beta2, int skip_zero)
{
    adam_advance_tensors(b, state[2] != 0.0, beta1, beta2, skip_zero != 0);
}

__global__ void __launch_bounds__(64) k_adam_prepare(double* __restrict__ acc, const float* __restrict__ grad_scale,
                               const float* __restrict__ found_inf_in, float max_norm, double beta1, double beta2,
                               double* __restrict__ state, float* __restrict__ found_inf_out,
                               const __grid_constant__ AdamBatch b, int skip_zero)
{
    __shared__ int s_bad;
    if (threadIdx.x == 0) s_bad = (acc[1] > 0.0 || !isfinite(acc[0]) || (found_inf_in && *found_inf_in > 0.0f)) ? 1 : 0;
    __syncthreads();
    adam_advance_tensors(b, s_bad != 0, beta1, beta2, skip_zero != 0);
    if (threadIdx.x != 0) return;
    const double sq = acc[0];
    const bool bad = acc[1] > 0.0 || !isfinite(sq) || (found_inf_in && *found_inf_in > 0.0f);
    acc[0] = 0.0;
    acc[1] = 0.0;
    const float inv = grad_scale ? 1.0f / *grad_scale : 1.0f;
    const float norm = (float)sqrt(sq);
    float clip = 1.0f;
    if (max_norm > 0.0f) clip = fminf(max_norm / (norm + 1e-6f), 1.0f);   // torch clip_grad_norm_: clamp(max/(norm+1e-6), max=1)
    double step = state[0];
    if (!bad) step += 1.0;
    state[0] = step;
    state[1] = (double)(inv * clip);
    state[2] = bad ? 1.0 : 0.0;
    state[3] = 1.0 - pow(beta1, step);
    state[4] = sqrt(1.0 - pow(beta2, step));
    state[5] = (double)norm;
    if (found_inf_out) *found_inf_out = bad ? 1.0f : 0.0f;
}

struct AdamK {   // per-launch constants, rounded to fp32 from the doubles torch.optim.Adam computes on the host
    float coef, omb1, beta2, omb2, eps, bc2_sqrt;
};

template <bool ADAMW, bool WRITE_G>
__device__ __forceinline__ void adam_one(float& p, float& g, float& m, float& v, const AdamK& k, float lr_wd, float wd,
                                         float step_size)
{
    g *= k.coef;
    if (wd != 0.0f) {
        if (ADAMW) p -= lr_wd * p;                    // AdamW: param.mul_(1 - lr * weight_decay)
        else g = fmaf(wd, p, g);                      // Adam:  grad.add(param, alpha = weight_decay)
    }
    m = m + k.omb1 * (g - m);                         // exp_avg.lerp_(grad, 1 - beta1)
    v = k.beta2 * v + k.omb2 * g * g;                 // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value = 1 - beta2)
    const float denom = sqrtf(v) / k.bc2_sqrt + k.eps;
    p -= step_size * (m / denom);                     // param.addcdiv_(exp_avg, denom, value = -lr / bias_correction1)
}

template <bool ADAMW, bool WRITE_G>
__global__ void __launch_bounds__(256) k_adam_apply(const __grid_constant__ AdamBatch b,
                                                    const double* __restrict__ state, double beta1, double beta2, double eps)
{
    if (state[2] != 0.0) return;                      // non-finite gradient somewhere: the whole step is skipped
    AdamK K = {(float)state[1], (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), (float)eps, (float)state[4]};
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    for (int k = 0; k < b.count; ++k) {
        const acn_adam_tensor& t = b.t[k];
        if (t.bias && t.bias[3] == 0.0) continue;                    // identically-zero gradient = "grad is None" (acn_adam_prepare)
        const double bc1 = t.bias ? t.bias[0] : state[3];           // per-tensor step (torch.optim.Adam) or the global one
        K.bc2_sqrt = (float)(t.bias ? t.bias[1] : state[4]);
        const float wd = (float)t.weight_decay, lr_wd = (float)(t.lr * t.weight_decay), step_size = (float)(t.lr / bc1);
        const int64_t n = t.n;
        const bool vec = (((uintptr_t)t.p | (uintptr_t)t.g | (uintptr_t)t.m | (uintptr_t)t.v) & 15) == 0;
        int64_t done = 0;
        if (vec) {
            float4* p4 = reinterpret_cast<float4*>(t.p);
            float4* g4 = reinterpret_cast<float4*>(t.g);
            float4* m4 = reinterpret_cast<float4*>(t.m);
            float4* v4 = reinterpret_cast<float4*>(t.v);
            const int64_t n4 = n >> 2;
            for (int64_t i = tid; i < n4; i += nth) {
                float4 P = p4[i], G = WRITE_G ? ld_once_f4(g4 + i) : ld_stream_f4(g4 + i), M = ld_once_f4(m4 + i), V = ld_once_f4(v4 + i);
                adam_one<ADAMW, WRITE_G>(P.x, G.x, M.x, V.x, K, lr_wd, wd, step_size);
                adam_one<ADAMW, WRITE_G>(P.y, G.y, M.y, V.y, K, lr_wd, wd, step_size);
                adam_one<ADAMW, WRITE_G>(P.z, G.z, M.z, V.z, K, lr_wd, wd, step_size);
                adam_one<ADAMW, WRITE_G>(P.w, G.w, M.w, V.w, K, lr_wd, wd, step_size);
                p4[i] = P;
                st_stream_f4(m4 + i, M);
                st_stream_f4(v4 + i, V);
                if (WRITE_G) st_stream_f4(g4 + i, G);
            }
            done = n4 << 2;
        }
        for (int64_t i = done + tid; i < n; i += nth) {
            float P = t.p[i], G = t.g[i], M = t.m[i], V = t.v[i];
            adam_one<ADAMW, WRITE_G>(P, G, M, V, K, lr_wd, wd, step_size);
            t.p[i] = P; t.m[i] = M; t.v[i] = V;
            if (WRITE_G) t.g[i] = G;
        }
    }
}

static int fill_batch(AdamBatch& b, const acn_adam_tensor* tensors, int count, bool need_state, const char* who) {
    ACN_REQUIRE(count >= 0 && count <= ACN_ADAM_MAX_TENSORS, ACN_EINVAL, "%s: count %d outside [0,%d]", who, count,
                ACN_ADAM_MAX_TENSORS);
    ACN_REQUIRE(count == 0 || tensors, ACN_EINVAL, "%s: null tensor list", who);
    b.count = count;
    for (int k = 0; k < count; ++k) {
        b.t[k] = tensors[k];
        ACN_REQUIRE(b.t[k].n >= 0 && (b.t[k].n == 0 || b.t[k].g), ACN_EINVAL, "%s: tensor %d has no gradient", who, k);
        ACN_REQUIRE((b.t[k].step == nullptr) == (b.t[k].bias == nullptr), ACN_EINVAL,
                    "%s: tensor %d: step and bias must be given together", who, k);
        if (need_state)
            ACN_REQUIRE(b.t[k].n == 0 || (b.t[k].p && b.t[k].m && b.t[k].v), ACN_EINVAL,
                        "%s: tensor %d lacks p / exp_avg / exp_avg_sq", who, k);
    }
    return ACN_OK;
}

static int adam_grid(acn_ctx* ctx, const AdamBatch& b) {
    int64_t total = 0;
    for (int k = 0; k < b.count; ++k) total += b.t[k].n;
    return acn_grid_1d(total, 256 * 16, (int64_t)ctx->sm_count * 8);   // 8 resident 256-thread CTAs per SM
}

extern "C" int acn_grad_sqnorm(acn_ctx* ctx, const acn_adam_tensor* tensors, int count,
                               const float* grad_scale_or_null, double* acc2, acn_stream stream)
{
    ACN_CHECK_CTX(ctx);
    ACN_REQUIRE(acc2 != nullptr, ACN_EINVAL, "acn_grad_sqnorm: null accumulator");
    AdamBatch b;
    int rc = fill_batch(b, tensors, count, false, "acn_grad_sqnorm");
    if (rc) return rc;
    if (count == 0) return ACN_OK;
    k_grad_sqnorm<<<adam_grid(ctx, b), 256, 0, (cudaStream_t)stream>>>(b, grad_scale_or_null, acc2);
    ACN_CHECK_LAUNCH();
    return ACN_OK;
}

extern "C" int acn_adam_prepare(acn_ctx* ctx, double* acc2, const float* grad_scale_or_null,
                                const float* found_inf_or_null, float max_norm, double beta1, double beta2,
                                double* state8, float* found_inf_out_or_null, const acn_adam_tensor* tensors_or_null,
                                int count, int skip_zero_grads, acn_stream stream)
{
    ACN_CHECK_CTX(ctx);
    ACN_REQUIRE(acc2 && state8, ACN_EINVAL, "acn_adam_prepare: null accumulator / state");
    ACN_REQUIRE(beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0, ACN_EINVAL,
                "acn_adam_prepare: betas (%g, %g) outside [0,1)", beta1, beta2);
    AdamBatch b;
    int rc = fill_batch(b, tensors_or_null, tensors_or_null ? count : 0, false, "acn_adam_prepare");
    if (rc) return rc;
    k_adam_prepare<<<1, 64, 0, (cudaStream_t)stream>>>(acc2, grad_scale_or_null, found_inf_or_null, max_norm, beta1,
                                                       beta2, state8, found_inf_out_or_null, b, skip_zero_grads);
    ACN_CHECK_LAUNCH();
    return ACN_OK;
}

extern "C" int acn_adam_advance(acn_ctx* ctx, const acn_adam_tensor* tensors, int count, const double* state8,
                                double beta1, double beta2, int skip_zero_grads, acn_stream stream)
{
    ACN_CHECK_CTX(ctx);
    ACN_REQUIRE(state8 != nullptr, ACN_EINVAL, "acn_adam_advance: null state");
    AdamBatch b;
    int rc = fill_batch(b, tensors, count, false, "acn_adam_advance");
    if (rc) return rc;
    if (count == 0) return ACN_OK;
    k_adam_advance<<<1, 64, 0, (cudaStream_t)stream>>>(b, state8, beta1, beta2, skip_zero_grads);
    ACN_CHECK_LAUNCH();
    return ACN_OK;
}

extern "C" int acn_adam_apply(acn_ctx* ctx, const acn_adam_tensor* tensors, int count, const double* state8,
                              double beta1, double beta2, double eps, int adamw, int write_grads, acn_stream stream_)
{
    ACN_CHECK_CTX(ctx);
    ACN_REQUIRE(state8 != nullptr, ACN_EINVAL, "acn_adam_apply: null state");
    AdamBatch b;
    int rc = fill_batch(b, tensors, count, true, "acn_adam_apply");
    if (rc) return rc;
    if (count == 0) return ACN_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int grid = adam_grid(ctx, b);
    if (adamw) {
        if (write_grads) k_adam_apply<true, true><<<grid, 256, 0, stream>>>(b, state8, beta1, beta2, eps);
        else k_adam_apply<true, false><<<grid, 256, 0, stream>>>(b, state8, beta1, beta2, eps);
    } else {
        if (write_grads) k_adam_apply<false, true><<<grid, 256, 0, stream>>>(b, state8, beta1, beta2, eps);
        else k_adam_apply<false, false><<<grid, 256, 0, stream>>>(b, state8, beta1, beta2, eps);
    }
    ACN_CHECK_LAUNCH();
    return ACN_OK;
}
