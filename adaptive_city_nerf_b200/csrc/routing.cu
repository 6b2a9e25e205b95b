// Stage 5: Voronoi routing of samples (models/inr/meta_container.py:97-134) and of rays
// (scripts/create_clusters.py:559-634), plus the device-side dispatch that replaces the
// reference's per-expert nonzero()/index_select()/index_add_() round trips.
//
// Distances follow torch.cdist's matmul formulation rounding step by step, so the
// hard assignment, the soft support set and the ray masks are bit-exact w.r.t. the CPU reference.
#include "acn_common.cuh"

#define FULL 0xffffffffu
static constexpr int ROUTE_THREADS = 256;    // points per block of the routing / bucketing kernels
static constexpr int ROUTE_WARPS = ROUTE_THREADS / 32;

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

// Per-expert counts are aggregated per block in shared memory (one global atomic per expert per block): with a warp
// per atomic, a 1080p frame issued ~35 M atomics onto the two or three counters a view touches and the kernel ran at
// the L2's same-address atomic rate instead of memory bandwidth.
template <int DIMS, int KT>
__global__ void __launch_bounds__(ROUTE_THREADS) k_route_points(
    const float* __restrict__ pts, int64_t P, int stride, const float* __restrict__ cen, int K, float margin,
    float* __restrict__ weights, int32_t* __restrict__ hard, int32_t* __restrict__ counts)
{
    extern __shared__ int s_cnt[];   // K ints (only when counts != nullptr)
    constexpr int OFF = DIMS == 2 ? 1 : 0;
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool on = p < P;
    const int lane = threadIdx.x & 31;
    if (counts) {
        for (int k = threadIdx.x; k < K; k += blockDim.x) s_cnt[k] = 0;
        __syncthreads();
    }
    float x[3] = { 0.f, 0.f, 0.f };
    if (on) {
#pragma unroll
        for (int k = 0; k < DIMS; ++k) x[k] = pts[p * stride + OFF + k];
    }
    if (weights && KT > 0) {
        // K == KT experts: distances computed once and kept in registers, weights stored as float4s
        float d[KT > 0 ? KT : 1];
        float mind = __int_as_float(0x7f800000);
#pragma unroll
        for (int k = 0; k < KT; ++k) {
            d[k] = fmaxf(cdist_mm<DIMS>(x, cen + 3 * k + OFF), 1e-6f);
            mind = fminf(mind, d[k]);
        }
        const float thr = __fmul_rn(margin, mind);
        float denom = 0.0f;
#pragma unroll
        for (int k = 0; k < KT; ++k) {
            d[k] = d[k] <= thr ? __fdiv_rn(1.0f, d[k]) : 0.0f;      // d[k] > 0, so 1/d > 0 marks "in"
            denom = __fadd_rn(denom, d[k]);
        }
        denom = fmaxf(denom, 1e-6f);
#pragma unroll
        for (int k = 0; k < KT; ++k) {
            const bool in = on && d[k] > 0.0f;
            d[k] = in ? __fdiv_rn(d[k], denom) : 0.0f;
            if (counts) {
                unsigned m = __ballot_sync(FULL, in);
                if (lane == 0 && m) atomicAdd_block(s_cnt + k, __popc(m));
            }
        }
        if (on) {
            float4* dst = reinterpret_cast<float4*>(weights + p * KT);
#pragma unroll
            for (int q = 0; q < KT / 4; ++q) dst[q] = make_float4(d[4 * q], d[4 * q + 1], d[4 * q + 2], d[4 * q + 3]);
        }
    } else if (weights) {
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
                if (lane == 0 && m) atomicAdd_block(s_cnt + k, __popc(m));
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
                if (lane == 0 && m) atomicAdd_block(s_cnt + k, __popc(m));
            }
        }
    }
    if (counts) {
        __syncthreads();
        for (int k = threadIdx.x; k < K; k += blockDim.x)
            if (s_cnt[k]) atomicAdd(counts + k, s_cnt[k]);
    }
}

// float atomic min / max on plain fp32 storage (mixed signs; the target starts at +inf / -inf)
__device__ __forceinline__ void atomic_min_f32(float* addr, float v) {
    if (v >= 0.0f) atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_f32(float* addr, float v) {
    if (v >= 0.0f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// scripts/create_clusters.py:592-632: per ray, S lerp'ed samples, min over samples of
// D / (min_c D + 1e-8), compared with the margin.
// AABB = true additionally streams the per-expert sample boxes of scripts/create_clusters.py:386-556 (update_aabbs:
// mins_out / maxs_out / counts_out): every SAMPLE that passes the same test for expert c extends c's box with its 3-D
// position o + d*t and counts once.  Non-finite samples (rays that miss the scene box carry near = far = inf) fail the
// comparison and contribute nothing, as in the reference.  Thread-local boxes -> warp shuffles -> one atomic per warp.
template <int DIMS, int MAXK, bool AABB>
__global__ void __launch_bounds__(128) k_route_rays(
    const float* __restrict__ rays8, int64_t N, int S, const float* __restrict__ u_lin,
    const float* __restrict__ cen, int K, float margin, uint8_t* __restrict__ mask,
    float* __restrict__ mins, float* __restrict__ maxs, unsigned long long* __restrict__ counts)
{
    constexpr int OFF = DIMS == 2 ? 1 : 0;
    constexpr int BK = AABB ? MAXK : 1;
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool ron = r < N;
    if (!AABB && !ron) return;
    if (!ron) r = N - 1;                                   // AABB: keep the lane alive for the shuffles
    const float4 a = __ldg(reinterpret_cast<const float4*>(rays8 + 8 * r));
    const float4 b = __ldg(reinterpret_cast<const float4*>(rays8 + 8 * r) + 1);
    const float o[3] = { a.x, a.y, a.z }, d[3] = { a.w, b.x, b.y };
    const float near = b.z, far = b.w;
    const float diff = __fsub_rn(far, near);
    const float INF = __int_as_float(0x7f800000);
    float rmin[MAXK];
    float bmin[BK][3], bmax[BK][3];
    unsigned int bcnt[BK];
#pragma unroll
    for (int k = 0; k < MAXK; ++k) rmin[k] = INF;
#pragma unroll
    for (int k = 0; k < BK; ++k) {
        bcnt[k] = 0;
#pragma unroll
        for (int c = 0; c < 3; ++c) { bmin[k][c] = INF; bmax[k][c] = -INF; }
    }
    for (int s = 0; s < S; ++s) {
        float z = __ldg(u_lin + s);
        // torch.lerp on CPU is fused: w<0.5 ? a + w*(b-a) : b - (b-a)*(1-w)
        float t = z < 0.5f ? __fmaf_rn(z, diff, near) : __fmaf_rn(-diff, __fsub_rn(1.0f, z), far);
        float x[3];
#pragma unroll
        for (int k = 0; k < DIMS; ++k) x[k] = __fadd_rn(o[OFF + k], __fmul_rn(d[OFF + k], t));
        float D[MAXK];
        float m = INF;
#pragma unroll
        for (int k = 0; k < MAXK; ++k) {
            if (k < K) { D[k] = cdist_mm<DIMS>(x, cen + 3 * k + OFF); m = fminf(m, D[k]); }
        }
        float den = __fadd_rn(m, 1e-8f);
        float p3[3];
        if constexpr (AABB) {
#pragma unroll
            for (int c = 0; c < 3; ++c) p3[c] = __fadd_rn(o[c], __fmul_rn(d[c], t));
        }
#pragma unroll
        for (int k = 0; k < MAXK; ++k) {
            if (k < K) {
                const float q = __fdiv_rn(D[k], den);
                rmin[k] = fminf(rmin[k], q);
                if constexpr (AABB) {
                    if (ron && q <= margin) {          // NaN fails
                        ++bcnt[k];
#pragma unroll
                        for (int c = 0; c < 3; ++c) { bmin[k][c] = fminf(bmin[k][c], p3[c]); bmax[k][c] = fmaxf(bmax[k][c], p3[c]); }
                    }
                }
            }
        }
    }
    if (ron) {
#pragma unroll
        for (int k = 0; k < MAXK; ++k) {
            if (k < K) mask[r * K + k] = rmin[k] <= margin ? 1 : 0;
        }
    }
    if constexpr (AABB) {
        const int lane = threadIdx.x & 31;
#pragma unroll
        for (int k = 0; k < MAXK; ++k) {
            if (k >= K) break;
            unsigned int n = bcnt[k];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) n += __shfl_xor_sync(0xffffffffu, n, off);
            if (n == 0) continue;                          // warp-uniform
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float lo = bmin[k][c], hi = bmax[k][c];
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, off));
                    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, off));
                }
                if (lane == 0) { atomic_min_f32(mins + 3 * k + c, lo); atomic_max_f32(maxs + 3 * k + c, hi); }
            }
            if (lane == 0 && counts) atomicAdd(counts + k, (unsigned long long)n);
        }
    }
}

// Device-side bucket.  Each block claims ONE contiguous range per expert (one global atomic per expert per block, see
// k_route_points): warps publish their per-expert counts, one thread per expert scans them and claims the range, then
// every lane derives its slot from block base + warp prefix + lane prefix.
template <bool DISPATCH>
__device__ __forceinline__ void bucket_body(
    const float* __restrict__ id6, int64_t P, const float* __restrict__ weights, const int32_t* __restrict__ hard,
    int K, const int32_t* __restrict__ offsets, int32_t* __restrict__ cursor, int32_t* __restrict__ sel,
    float* __restrict__ xd_out, float* __restrict__ w_out, const unsigned long long* __restrict__ row_base,
    const int32_t* __restrict__ row_off)
{
    extern __shared__ int s_base[];   // ROUTE_WARPS * K ints: per-warp count, then per-warp first index within the expert
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool on = p < P;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int h = (on && hard) ? hard[p] : -1;
    for (int k = 0; k < K; ++k) {
        const float w = weights ? (on ? weights[p * K + k] : 0.0f) : (h == k ? 1.0f : 0.0f);
        const unsigned m = __ballot_sync(FULL, on && w > 0.0f);
        if (lane == 0) s_base[warp * K + k] = __popc(m);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        int total = 0;
        for (int wi = 0; wi < ROUTE_WARPS; ++wi) total += s_base[wi * K + k];
        int run = total ? atomicAdd(cursor + k, total) : 0;
        for (int wi = 0; wi < ROUTE_WARPS; ++wi) {
            const int c = s_base[wi * K + k];
            s_base[wi * K + k] = run;
            run += c;
        }
    }
    __syncthreads();
    float2 r0 = make_float2(0.f, 0.f), r1 = r0, r2 = r0;
    if (on && (DISPATCH || xd_out)) {
        const float2* src = reinterpret_cast<const float2*>(id6 + p * 6);
        r0 = __ldg(src); r1 = __ldg(src + 1); r2 = __ldg(src + 2);
    }
    for (int k = 0; k < K; ++k) {
        const float w = weights ? (on ? weights[p * K + k] : 0.0f) : (h == k ? 1.0f : 0.0f);
        const bool in = on && w > 0.0f;
        const unsigned m = __ballot_sync(FULL, in);
        if (!in) continue;
        const int idx = s_base[warp * K + k] + __popc(m & ((1u << lane) - 1u));
        const int slot = __ldg(offsets + k) + idx;
        sel[slot] = (int32_t)p;
        if (w_out) w_out[slot] = w;
        if (DISPATCH) {
            float2* dst = reinterpret_cast<float2*>(__ldg(row_base + k)) + ((size_t)__ldg(row_off + k) + idx) * 3;
            dst[0] = r0; dst[1] = r1; dst[2] = r2;
        } else if (xd_out) {
            float2* dst = reinterpret_cast<float2*>(xd_out) + (size_t)slot * 3;
            dst[0] = r0; dst[1] = r1; dst[2] = r2;
        }
    }
}

__global__ void __launch_bounds__(ROUTE_THREADS) k_bucket_points(
    const float* __restrict__ id6, int64_t P, const float* __restrict__ weights, const int32_t* __restrict__ hard,
    int K, const int32_t* __restrict__ offsets, int32_t* __restrict__ cursor, int32_t* __restrict__ sel,
    float* __restrict__ xd_out, float* __restrict__ w_out)
{
    bucket_body<false>(id6, P, weights, hard, K, offsets, cursor, sel, xd_out, w_out, nullptr, nullptr);
}

// Fused dispatch for expert sharding: the same bucketing, but the routed [xyz, dir] row of expert k is stored straight
// into the receive buffer of the GPU that owns k -- row_base[k] is that buffer's address as mapped into this process
// (NVLink peer memory, or local memory for the experts this rank owns), row_off[k] the first row reserved there for
// (this source rank, expert k).  sel / w_out stay local.  The dispatch half of the all-to-all is therefore the
// kernel's own store stream; no staging copy, no NCCL send.
__global__ void __launch_bounds__(ROUTE_THREADS) k_dispatch_points(
    const float* __restrict__ id6, int64_t P, const float* __restrict__ weights, const int32_t* __restrict__ hard,
    int K, const int32_t* __restrict__ offsets, int32_t* __restrict__ cursor, int32_t* __restrict__ sel,
    float* __restrict__ w_out, const unsigned long long* __restrict__ row_base, const int32_t* __restrict__ row_off)
{
    bucket_body<true>(id6, P, weights, hard, K, offsets, cursor, sel, nullptr, w_out, row_base, row_off);
}

// Routing + bucketing straight from (rays, t): the container's render path without the (P,6) point matrix and the
// (P,K) weight matrix in HBM (a 1080p frame: 3.2 GB + 4.2 GB written and read back).  Two launches of the same body:
// COUNT computes every sample's support set (one bit per expert), stores it (2 bytes per sample) and adds up the
// per-expert row counts -- the one host read that sizes the buckets; BUCKET reads the set back, evaluates distances
// and weights only for the experts in it (one for all but the few % of samples near a cell boundary), claims one
// contiguous range per expert per block and writes sel / w / [xyz, dir] rows.  Same arithmetic as k_points +
// k_route_points (terms outside the set contribute an exact +0 to the normaliser), so rows and weights are
// bit-identical to the unfused path.
// cdist_mm without the final square root (the value torch's cdist clamps at 0 and roots)
template <int DIMS>
__device__ __forceinline__ float cdist_mm_sq(const float* x, const float* __restrict__ c) {
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
    return fmaxf(acc, 0.0f);
}

// Support set of one sample (meta_container.py:97-134): bit k set <=> d_k <= margin * min_j d_j, d = max(cdist, 1e-6)
// (margin > 1), or the argmin (margin == 1).  The IEEE square roots are only taken where they can matter: with
// a = d^2 before rounding, a_k > a_min * margin^2 * (1 + 2e-5) implies d_k > fl(margin * d_min) by a margin three orders
// above the rounding of sqrt / the product, so such experts are out without a root; a lone survivor is the minimum
// itself, which is always in the set (margin >= 1).  Everything else -- ties, boundary samples, tiny or non-finite
// distances -- takes the exact path, so the result is bit-identical to evaluating every distance.
template <int DIMS, int MAXK>
__device__ __forceinline__ unsigned support_bits(const float* pos, const float* __restrict__ cen, int K, float margin) {
    constexpr int OFF = DIMS == 2 ? 1 : 0;
    unsigned bits = 0;
    if (margin > 1.0f) {
        float a[MAXK];
        float amin = __int_as_float(0x7f800000);
#pragma unroll
        for (int k = 0; k < MAXK; ++k) {
            a[k] = 0.0f;
            if (k < K) { a[k] = cdist_mm_sq<DIMS>(pos + OFF, cen + 3 * k + OFF); amin = fminf(amin, a[k]); }
        }
        const float cut = amin * (margin * margin * 1.00002f);
        unsigned cand = 0;
#pragma unroll
        for (int k = 0; k < MAXK; ++k)
            if (k < K && !(a[k] > cut)) cand |= 1u << k;          // NaN stays a candidate and fails the exact test below
        if (__popc(cand) == 1 && amin >= 1e-11f && amin < 1e30f) return cand;
        float mind = __int_as_float(0x7f800000);
#pragma unroll
        for (int k = 0; k < MAXK; ++k)
            if ((cand >> k) & 1u) { a[k] = fmaxf(__fsqrt_rn(a[k]), 1e-6f); mind = fminf(mind, a[k]); }
        // the minimum over all experts is attained among the candidates (amin itself is one)
        if (!(amin >= 1e-11f && amin < 1e30f)) {                 // 1e-6 floor in play, or overflow: every distance, exactly
            cand = 0; mind = __int_as_float(0x7f800000);
#pragma unroll
            for (int k = 0; k < MAXK; ++k)
                if (k < K) { a[k] = fmaxf(cdist_mm<DIMS>(pos + OFF, cen + 3 * k + OFF), 1e-6f); mind = fminf(mind, a[k]); cand |= 1u << k; }
        }
        const float thr = __fmul_rn(margin, mind);
#pragma unroll
        for (int k = 0; k < MAXK; ++k)
            if (((cand >> k) & 1u) && a[k] <= thr) bits |= 1u << k;
    } else {                                                // hard: first minimum wins, like argmin
        int best = 0;
        float bd = cdist_mm<DIMS>(pos + OFF, cen + OFF);
#pragma unroll
        for (int k = 1; k < MAXK; ++k) {
            if (k < K) { float d = cdist_mm<DIMS>(pos + OFF, cen + 3 * k + OFF); if (d < bd) { bd = d; best = k; } }
        }
        bits = 1u << best;
    }
    return bits;
}

template <int DIMS, int MAXK, bool BUCKET>
__global__ void __launch_bounds__(ROUTE_THREADS) k_route_samples(
    const float* __restrict__ rays8, const float* __restrict__ t_vals, int64_t P, int S, const float* __restrict__ cen,
    int K, float margin, int ray_major, const int32_t* __restrict__ ray_major_dev, uint16_t* __restrict__ support,
    int32_t* __restrict__ counts,
    const int32_t* __restrict__ offsets, int32_t* __restrict__ cursor, int32_t* __restrict__ sel, float* __restrict__ xd_out,
    float* __restrict__ w_out, const unsigned long long* __restrict__ row_base, const int32_t* __restrict__ row_off,
    const int32_t* __restrict__ row_limit)
{
    extern __shared__ int s_k[];      // COUNT: K ints; BUCKET: ROUTE_WARPS * K ints
    constexpr int OFF = DIMS == 2 ? 1 : 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // Which sample a thread owns decides the ORDER of the rows inside a bucket, i.e. what the experts' gather kernels see
    // in one warp.  sample-major (training batches of unrelated rays): 32 consecutive samples of one ray, which share the
    // coarse cells.  ray-major (frames: consecutive rays are adjacent pixels): the same sample of 32 adjacent pixels,
    // which share cells up to the levels whose cells are as small as a pixel footprint -- 2-3x fewer distinct lines
    // per gather instruction on a 1080p view.
    int64_t p;
    bool on;
    const bool rm = ray_major_dev ? (__ldg(ray_major_dev) != 0) : (ray_major != 0);
    // ray-major: the block owns a 32-ray x ROUTE_WARPS-sample tile (32 x 8).  t_vals and the support sets are (ray,
    // sample) row-major, so a warp (= one sample of 32 rays) would touch 32 different lines per access; the tile goes
    // through shared memory instead, moved by threads laid out (ray = tid / ROUTE_WARPS, sample = tid % ROUTE_WARPS):
    // 32 contiguous bytes of t per ray.
    __shared__ float s_t[32][ROUTE_WARPS + 1];
    __shared__ uint16_t s_sup[32][ROUTE_WARPS + 2];
    int64_t p_tile = 0;                                     // the element this thread moves for the tile
    bool on_tile = false;
    const int tr = threadIdx.x / ROUTE_WARPS, ts = threadIdx.x % ROUTE_WARPS;
    if (rm) {
        const int sgroups = (S + ROUTE_WARPS - 1) / ROUTE_WARPS;
        const int64_t r0 = (int64_t)(blockIdx.x / sgroups) * 32;
        const int s0 = (blockIdx.x % sgroups) * ROUTE_WARPS;
        const int64_t r = r0 + lane;
        const int si = s0 + warp;
        on = si < S && r * S < P;
        p = r * S + si;
        on_tile = (s0 + ts) < S && (r0 + tr) * S < P;
        p_tile = (r0 + tr) * S + s0 + ts;
        if (on_tile) {
            s_t[tr][ts] = t_vals[p_tile];
            if (BUCKET && support) s_sup[tr][ts] = support[p_tile];
        }
    } else {
        p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        on = p < P;
    }
    if (!BUCKET) {
        for (int k = threadIdx.x; k < K; k += blockDim.x) s_k[k] = 0;
    }
    if (rm || !BUCKET) __syncthreads();
    float pos[3] = { 0.f, 0.f, 0.f }, dir[3] = { 0.f, 0.f, 0.f };
    if (on) {
        const int64_t r = p / S;
        const float4 a = __ldg(reinterpret_cast<const float4*>(rays8 + 8 * r));
        const float4 b = __ldg(reinterpret_cast<const float4*>(rays8 + 8 * r) + 1);
        const float t = rm ? s_t[lane][warp] : t_vals[p];
        dir[0] = a.w; dir[1] = b.x; dir[2] = b.y;
        pos[0] = __fadd_rn(a.x, __fmul_rn(dir[0], t));      // k_points
        pos[1] = __fadd_rn(a.y, __fmul_rn(dir[1], t));
        pos[2] = __fadd_rn(a.z, __fmul_rn(dir[2], t));
    }
    unsigned bits = 0;
    if (on) bits = (BUCKET && support) ? (unsigned)(rm ? s_sup[lane][warp] : support[p]) : support_bits<DIMS, MAXK>(pos, cen, K, margin);
    if (!BUCKET && support) {
        if (rm) {
            s_sup[lane][warp] = (uint16_t)bits;
            __syncthreads();
            if (on_tile) support[p_tile] = s_sup[tr][ts];
        } else if (on) {
            support[p] = (uint16_t)bits;
        }
    }
#pragma unroll
    for (int k = 0; k < MAXK; ++k) {
        if (k < K) {
            const unsigned m = __ballot_sync(FULL, (bits >> k) & 1u);
            if (lane == 0) {
                if (BUCKET) s_k[warp * K + k] = __popc(m);
                else if (m) atomicAdd_block(s_k + k, __popc(m));
            }
        }
    }
    __syncthreads();
    if (!BUCKET) {
        for (int k = threadIdx.x; k < K; k += blockDim.x)
            if (s_k[k]) atomicAdd(counts + k, s_k[k]);
        return;
    }
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        int total = 0;
        for (int wi = 0; wi < ROUTE_WARPS; ++wi) total += s_k[wi * K + k];
        int run = total ? atomicAdd(cursor + k, total) : 0;
        for (int wi = 0; wi < ROUTE_WARPS; ++wi) {
            const int c = s_k[wi * K + k];
            s_k[wi * K + k] = run;
            run += c;
        }
    }
    __syncthreads();
    // blend weights of the experts in the set: w_k = (1/d_k) / max(sum over the set of 1/d, 1e-6), summed in expert order
    float inv[MAXK];
    float denom = 0.0f;
    const bool soft = margin > 1.0f;
#pragma unroll
    for (int k = 0; k < MAXK; ++k) {
        inv[k] = 0.0f;
        if (soft && ((bits >> k) & 1u)) {
            inv[k] = __fdiv_rn(1.0f, fmaxf(cdist_mm<DIMS>(pos + OFF, cen + 3 * k + OFF), 1e-6f));
            denom = __fadd_rn(denom, inv[k]);
        }
    }
    denom = fmaxf(denom, 1e-6f);
#pragma unroll
    for (int k = 0; k < MAXK; ++k) {
        if (k < K) {
            const bool in = (bits >> k) & 1u;
            const unsigned m = __ballot_sync(FULL, in);
            // row_limit (capacity-bounded buckets: the host never read the counts, acn_bucket_plan clamped them to what
            // fits): rows beyond an expert's share are dropped; the plan has raised the overflow flag
            if (in && (!row_limit || s_k[warp * K + k] + __popc(m & ((1u << lane) - 1u)) < __ldg(row_limit + k))) {
                const int idx = s_k[warp * K + k] + __popc(m & ((1u << lane) - 1u));
                const int slot = __ldg(offsets + k) + idx;
                sel[slot] = (int32_t)p;
                w_out[slot] = soft ? __fdiv_rn(inv[k], denom) : 1.0f;
                // expert sharding: the row goes straight into the receive buffer of the GPU that owns expert k (see
                // k_dispatch_points); otherwise into this rank's bucket
                float2* dst = row_base ? reinterpret_cast<float2*>(__ldg(row_base + k)) + ((size_t)__ldg(row_off + k) + idx) * 3
                                       : reinterpret_cast<float2*>(xd_out) + (size_t)slot * 3;
                dst[0] = make_float2(pos[0], pos[1]);
                dst[1] = make_float2(pos[2], dir[0]);
                dst[2] = make_float2(dir[1], dir[2]);
            }
        }
    }
}

// Rows [0, M) of (y, w, sel), or -- `range` (device, 2 int32) given -- rows [range[0], range[1]) with M an upper bound on
// their number (grid-stride).  y_base_off: y is addressed from row (range[0] - y_row0) when y_row0 is given (an owner's
// buffer whose rows start elsewhere than the local bucket's).
__global__ void k_blend_add(const float4* __restrict__ y, const float* __restrict__ w, const int32_t* __restrict__ sel,
                            int64_t M, const int32_t* __restrict__ range, const int32_t* __restrict__ y_row0, float4* __restrict__ out) {
    int64_t r0 = 0, r1 = M;
    if (range) { r0 = __ldg(range); r1 = __ldg(range + 1); }
    const int64_t yoff = y_row0 ? (int64_t)__ldg(y_row0) - r0 : 0;
    for (int64_t i = r0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < r1; i += (int64_t)gridDim.x * blockDim.x) {
        float4 v = y[i + yoff];
        float ww = w ? w[i] : 1.0f;
        int32_t p = sel[i];
        float4 o = out[p];
        o.x += v.x * ww; o.y += v.y * ww; o.z += v.z * ww; o.w += v.w * ww;
        out[p] = o;
    }
}

__global__ void k_blend_bwd(const float4* __restrict__ d_out, const float* __restrict__ w, const int32_t* __restrict__ sel,
                            int64_t M, const int32_t* __restrict__ range, const int32_t* __restrict__ y_row0, float4* __restrict__ d_y) {
    int64_t r0 = 0, r1 = M;
    if (range) { r0 = __ldg(range); r1 = __ldg(range + 1); }
    const int64_t yoff = y_row0 ? (int64_t)__ldg(y_row0) - r0 : 0;
    for (int64_t i = r0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < r1; i += (int64_t)gridDim.x * blockDim.x) {
        float4 g = d_out[sel[i]];
        float ww = w ? w[i] : 1.0f;
        d_y[i + yoff] = make_float4(g.x * ww, g.y * ww, g.z * ww, g.w * ww);
    }
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
    ACN_REQUIRE(K <= 4096, ACN_EUNSUPPORTED, "acn_route_points: K=%d > 4096", K);
    const int grid = acn_grid_1d(P, ROUTE_THREADS);
    const size_t smem = counts_or_null ? (size_t)K * sizeof(int) : 0;
    cudaStream_t st = (cudaStream_t)stream;
    float* w = soft ? weights_or_null : nullptr;
#define RP(D, KT) k_route_points<D, KT><<<grid, ROUTE_THREADS, smem, st>>>(pts, P, stride, centroids, K, margin, w, hard_or_null, counts_or_null)
    const bool vec = w && ((uintptr_t)w & 15) == 0;
    if (dims == 2) { if (vec && K == 8) RP(2, 8); else if (vec && K == 4) RP(2, 4); else RP(2, 0); }
    else           { if (vec && K == 8) RP(3, 8); else if (vec && K == 4) RP(3, 4); else RP(3, 0); }
#undef RP
    ACN_CHECK_LAUNCH();
    return ACN_OK;
}

extern "C" int acn_route_rays_voronoi(acn_ctx* ctx, const float* rays8, int64_t N, int S, const float* u_lin,
                                      const float* centroids, int K, int dims, float margin, uint8_t* mask,
                                      float* mins_or_null, float* maxs_or_null, int64_t* counts_or_null, acn_stream stream) {
    ACN_CHECK_CTX(ctx);
    ACN_REQUIRE(N >= 0 && S >= 1 && u_lin && centroids, ACN_EINVAL, "acn_route_rays_voronoi: bad arguments");
    ACN_REQUIRE(K >= 1 && K <= 64, ACN_EUNSUPPORTED, "acn_route_rays_voronoi: K=%d outside [1,64]", K);
    ACN_REQUIRE(dims == 2 || dims == 3, ACN_EINVAL, "acn_route_rays_voronoi: dims must be 2 or 3");
    const bool aabb = mins_or_null != nullptr;
    ACN_REQUIRE((maxs_or_null != nullptr) == aabb && (aabb || !counts_or_null), ACN_EINVAL,
                "acn_route_rays_voronoi: give mins and maxs together (counts only with them)");
    ACN_REQUIRE(!aabb || K <= 16, ACN_EUNSUPPORTED, "acn_route_rays_voronoi: the per-expert boxes are built for K <= 16 (got %d)", K);
    if (N == 0) return ACN_OK;
    ACN_REQUIRE(rays8 && mask, ACN_EINVAL, "acn_route_rays_voronoi: null buffer");
    ACN_REQUIRE(((uintptr_t)rays8 & 15) == 0, ACN_EINVAL, "acn_route_rays_voronoi: rays8 must be 16-byte aligned");
    const int grid = acn_grid_1d(N, 128);
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* cnt = reinterpret_cast<unsigned long long*>(counts_or_null);
#define RR(D, MK, AB) k_route_rays<D, MK, AB><<<grid, 128, 0, st>>>(rays8, N, S, u_lin, centroids, K, margin, mask, mins_or_null, maxs_or_null, cnt)
    if (aabb) {
        if (dims == 2) { if (K <= 4) RR(2, 4, true); else if (K <= 8) RR(2, 8, true); else RR(2, 16, true); }
        else           { if (K <= 4) RR(3, 4, true); else if (K <= 8) RR(3, 8, true); else RR(3, 16, true); }
    } else {
        if (dims == 2) { if (K <= 4) RR(2, 4, false); else if (K <= 8) RR(2, 8, false); else if (K <= 16) RR(2, 16, false); else RR(2, 64, false); }
        else           { if (K <= 4) RR(3, 4, false); else if (K <= 8) RR(3, 8, false); else if (K <= 16) RR(3, 16, false); else RR(3, 64, false); }
    }
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
    ACN_REQUIRE(!xd_out || (id6 && ((uintptr_t)id6 & 7) == 0 && ((uintptr_t)xd_out & 7) == 0), ACN_EINVAL,
                "acn_bucket_points: xd_out needs id6, both 8-byte aligned");
    ACN_REQUIRE(K <= 512, ACN_EUNSUPPORTED, "acn_bucket_points: K=%d > 512", K);
    if (P == 0) return ACN_OK;
    k_bucket_points<<<acn_grid_1d(P, ROUTE_THREADS), ROUTE_THREADS, (size_t)ROUTE_WARPS * K * sizeof(int), (cudaStream_t)stream>>>(
        id6, P, weights_or_null, hard_or_null, K, offsets, cursor, sel, xd_out, w_out);
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
    ACN_REQUIRE(K <= 512, ACN_EUNSUPPORTED, "acn_dispatch_points: K=%d > 512", K);
    k_dispatch_points<<<acn_grid_1d(P, ROUTE_THREADS), ROUTE_THREADS, (size_t)ROUTE_WARPS * K * sizeof(int), (cudaStream_t)stream>>>(
        id6, P, weights_or_null, hard_or_null, K, offsets, cursor, sel, w_out, (const unsigned long long*)row_base, row_off);
    ACN_CHECK_LAUNCH();
    return ACN_OK;
}

template <bool BUCKET>
static int launch_route_samples(const float* rays8, const float* t_vals, int64_t N, int S, const float* cen, int K, int dims,
                                float margin, int ray_major, const int32_t* ray_major_dev, uint16_t* support, int32_t* counts,
                                const int32_t* offsets, int32_t* cursor, int32_t* sel, float* xd_out, float* w_out,
                                const unsigned long long* row_base, const int32_t* row_off, const int32_t* row_limit, cudaStream_t st) {
    const int64_t P = N * S;
    const int grid = (ray_major || ray_major_dev) ? (int)(((N + 31) / 32) * ((S + ROUTE_WARPS - 1) / ROUTE_WARPS)) : acn_grid_1d(P, ROUTE_THREADS);
    const size_t smem = (size_t)(BUCKET ? ROUTE_WARPS : 1) * K * sizeof(int);
#define RS(D, MK) k_route_samples<D, MK, BUCKET><<<grid, ROUTE_THREADS, smem, st>>>(rays8, t_vals, P, S, cen, K, margin, ray_major, ray_major_dev, \
                                                                                  support, \
                                                                                  counts, offsets, cursor, sel, xd_out, w_out, row_base, row_off, row_limit)
    if (dims == 2) { if (K <= 4) RS(2, 4); else if (K <= 8) RS(2, 8); else RS(2, 16); }
    else           { if (K <= 4) RS(3, 4); else if (K <= 8) RS(3, 8); else RS(3, 16); }
#undef RS
    return 0;
}

static int check_route_samples(const char* who, const float* rays8, const float* t_vals, int64_t N, int S, const float* cen,
                               int K, int dims, float margin) {
    ACN_REQUIRE(N >= 0 && S >= 1 && cen, ACN_EINVAL, "%s: bad arguments", who);
    ACN_REQUIRE(K >= 1 && K <= 16, ACN_EUNSUPPORTED, "%s: K=%d outside [1,16] (use acn_route_points + acn_bucket_points)", who, K);
    ACN_REQUIRE(dims == 2 || dims == 3, ACN_EINVAL, "%s: dims must be 2 (cluster_2d) or 3", who);
    ACN_REQUIRE(margin >= 1.0f, ACN_EINVAL, "%s: boundary_margin must be >= 1", who);
    ACN_REQUIRE(N == 0 || (rays8 && t_vals && ((uintptr_t)rays8 & 15) == 0), ACN_EINVAL, "%s: rays8 / t_vals null or rays8 not 16-byte aligned", who);
    ACN_REQUIRE(N * (int64_t)S < ((int64_t)1 << 31), ACN_EUNSUPPORTED, "%s: more than 2^31 samples per call", who);
    return ACN_OK;
}

extern "C" int acn_route_count_rays(acn_ctx* ctx, const float* rays8, const float* t_vals, int64_t N, int S,
                                    const float* centroids, int K, int dims, float margin, int ray_major,
                                    const int32_t* ray_major_dev_or_null, uint16_t* support_or_null, int32_t* counts,
                                    acn_stream stream) {
    ACN_CHECK_CTX(ctx);
    int rc = check_route_samples("acn_route_count_rays", rays8, t_vals, N, S, centroids, K, dims, margin);
    if (rc) return rc;
    ACN_REQUIRE(counts, ACN_EINVAL, "acn_route_count_rays: null counts");
    if (N == 0) return ACN_OK;
    launch_route_samples<false>(rays8, t_vals, N, S, centroids, K, dims, margin, ray_major, ray_major_dev_or_null, support_or_null, counts, nullptr, nullptr, nullptr,
                                nullptr, nullptr, nullptr, nullptr, nullptr, (cudaStream_t)stream);
    ACN_CHECK_LAUNCH();
    return ACN_OK;
}

extern "C" int acn_route_bucket_rays(acn_ctx* ctx, const float* rays8, const float* t_vals, int64_t N, int S,
                                     const float* centroids, int K, int dims, float margin, int ray_major,
                                     const int32_t* ray_major_dev_or_null, const uint16_t* support_or_null,
                                     const int32_t* offsets, int32_t* cursor, int32_t* sel, float* xd_out, float* w_out,
                                     const uint64_t* row_base_or_null, const int32_t* row_off_or_null,
                                     const int32_t* row_limit_or_null, acn_stream stream) {
    ACN_CHECK_CTX(ctx);
    int rc = check_route_samples("acn_route_bucket_rays", rays8, t_vals, N, S, centroids, K, dims, margin);
    if (rc) return rc;
    ACN_REQUIRE(offsets && cursor, ACN_EINVAL, "acn_route_bucket_rays: null offsets / cursor");
    if (N == 0) return ACN_OK;
    ACN_REQUIRE(sel && w_out, ACN_EINVAL, "acn_route_bucket_rays: null sel / w_out");
    ACN_REQUIRE((row_base_or_null != nullptr) == (row_off_or_null != nullptr), ACN_EINVAL, "acn_route_bucket_rays: row_base and row_off go together");
    ACN_REQUIRE(row_base_or_null || (xd_out && ((uintptr_t)xd_out & 7) == 0), ACN_EINVAL,
                "acn_route_bucket_rays: xd_out (8-byte aligned) or per-expert destination buffers are needed");
    launch_route_samples<true>(rays8, t_vals, N, S, centroids, K, dims, margin, ray_major, ray_major_dev_or_null,
                               const_cast<uint16_t*>(support_or_null), nullptr,
                               offsets, cursor, sel, xd_out, w_out, (const unsigned long long*)row_base_or_null, row_off_or_null,
                               row_limit_or_null, (cudaStream_t)stream);
    ACN_CHECK_LAUNCH();
    return ACN_OK;
}

// Turns the device-side per-expert row counts of a routed batch into the bucket layout WITHOUT the host reading them
// ("dispatch counts stay on device"): seg (K+1) = exclusive scan of the counts clamped to `cap` rows in
// total, limit (K) = rows of each expert that fit, cursor (K) = 0 for the bucket pass, *overflow |= 1 when rows were cut.
__global__ void k_bucket_plan(const int32_t* __restrict__ counts, int K, int64_t cap, int32_t* __restrict__ seg,
                              int32_t* __restrict__ limit, int32_t* __restrict__ cursor, int32_t* __restrict__ overflow) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int64_t run = 0;
    bool cut = false;
    seg[0] = 0;
    for (int k = 0; k < K; ++k) {
        int64_t c = counts[k];
        if (run + c > cap) { c = cap - run; cut = true; }
        limit[k] = (int32_t)c;
        cursor[k] = 0;
        run += c;
        seg[k + 1] = (int32_t)run;
    }
    if (cut && overflow) *overflow = 1;
}

extern "C" int acn_bucket_plan(acn_ctx* ctx, const int32_t* counts, int K, int64_t cap, int32_t* seg, int32_t* limit,
                               int32_t* cursor, int32_t* overflow_or_null, acn_stream stream) {
    ACN_CHECK_CTX(ctx);
    ACN_REQUIRE(counts && seg && limit && cursor && K >= 1 && cap >= 0 && cap < ((int64_t)1 << 31), ACN_EINVAL, "acn_bucket_plan: bad arguments");
    k_bucket_plan<<<1, 32, 0, (cudaStream_t)stream>>>(counts, K, cap, seg, limit, cursor, overflow_or_null);
    ACN_CHECK_LAUNCH();
    return ACN_OK;
}

// Expert sharding without a host read: from the all-gathered per-expert row counts of every rank (world x K, on the
// device) every rank derives, identically, where each (source rank s, expert k) segment lies in the receive buffer of
// k's owner (owner = k % world, buffer order: local expert e = k / world, then source rank) with every buffer clamped to
// cap_peer rows, and takes its own part:
//   seg_local (K+1)  this rank's local bucket layout (sel / w arrays, expert order), clamped to cap_local rows
//   limit (K)        rows of expert k this rank may write (fits both the local arrays and the owner's buffer)
//   row_off (K)      first row of (this rank, expert k) in the owner's buffer
//   seg_recv (m+1)   row ranges of this rank's own experts in ITS buffer (all sources), m = K / world
//   cursor (K) = 0;  *overflow = 1 when anything was cut
__global__ void k_shard_plan(const int32_t* __restrict__ all_counts, int world, int K, int rank, int64_t cap_local, int64_t cap_peer,
                             int32_t* __restrict__ seg_local, int32_t* __restrict__ limit, int32_t* __restrict__ row_off,
                             int32_t* __restrict__ seg_recv, int32_t* __restrict__ cursor, int32_t* __restrict__ overflow) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int m = K / world;
    bool cut = false;
    for (int o = 0; o < world; ++o) {                      // every owner's buffer, filled in (local expert, source) order
        int64_t run = 0;
        for (int e = 0; e < m; ++e) {
            const int k = e * world + o;
            if (o == rank) seg_recv[e] = (int32_t)run;
            for (int s_ = 0; s_ < world; ++s_) {
                int64_t c = all_counts[s_ * K + k];
                if (run + c > cap_peer) { c = cap_peer - run; cut = true; }
                if (s_ == rank) { row_off[k] = (int32_t)run; limit[k] = (int32_t)c; }
                run += c;
            }
        }
        if (o == rank) seg_recv[m] = (int32_t)run;
    }
    int64_t run = 0;
    seg_local[0] = 0;
    for (int k = 0; k < K; ++k) {
        int64_t c = limit[k];
        if (run + c > cap_local) { c = cap_local - run; cut = true; }
        limit[k] = (int32_t)c;
        cursor[k] = 0;
        run += c;
        seg_local[k + 1] = (int32_t)run;
    }
    if (cut && overflow) *overflow = 1;
}

extern "C" int acn_shard_plan(acn_ctx* ctx, const int32_t* all_counts, int world, int K, int rank, int64_t cap_local, int64_t cap_peer,
                              int32_t* seg_local, int32_t* limit, int32_t* row_off, int32_t* seg_recv, int32_t* cursor,
                              int32_t* overflow_or_null, acn_stream stream) {
    ACN_CHECK_CTX(ctx);
    ACN_REQUIRE(all_counts && seg_local && limit && row_off && seg_recv && cursor, ACN_EINVAL, "acn_shard_plan: null buffer");
    ACN_REQUIRE(world >= 1 && K >= 1 && K % world == 0 && rank >= 0 && rank < world, ACN_EINVAL, "acn_shard_plan: K=%d experts over %d ranks (rank %d)", K, world, rank);
    ACN_REQUIRE(cap_local >= 0 && cap_peer >= 0 && cap_local < ((int64_t)1 << 31) && cap_peer < ((int64_t)1 << 31), ACN_EINVAL, "acn_shard_plan: bad capacities");
    k_shard_plan<<<1, 32, 0, (cudaStream_t)stream>>>(all_counts, world, K, rank, cap_local, cap_peer, seg_local, limit, row_off, seg_recv,
                                                     cursor, overflow_or_null);
    ACN_CHECK_LAUNCH();
    return ACN_OK;
}

static int blend_grid(acn_ctx* ctx, int64_t M, bool ranged) {
    const int64_t blocks = (M + 255) / 256, cap = (int64_t)ctx->sm_count * 16;
    return (int)(blocks < 1 ? 1 : (ranged && blocks > cap ? cap : blocks));
}

extern "C" int acn_blend_add(acn_ctx* ctx, const float* y, const float* w, const int32_t* sel, int64_t M,
                             const int32_t* range_or_null, const int32_t* y_row0_or_null, float* out, acn_stream stream) {
    ACN_CHECK_CTX(ctx);
    ACN_REQUIRE(M >= 0, ACN_EINVAL, "acn_blend_add: negative M");
    if (M == 0) return ACN_OK;
    ACN_REQUIRE(y && sel && out, ACN_EINVAL, "acn_blend_add: null buffer");
    k_blend_add<<<blend_grid(ctx, M, range_or_null != nullptr), 256, 0, (cudaStream_t)stream>>>((const float4*)y, w, sel, M, range_or_null,
                                                                                               y_row0_or_null, (float4*)out);
    ACN_CHECK_LAUNCH();
    return ACN_OK;
}

extern "C" int acn_blend_bwd(acn_ctx* ctx, const float* d_out, const float* w, const int32_t* sel, int64_t M,
                             const int32_t* range_or_null, const int32_t* y_row0_or_null, float* d_y, acn_stream stream) {
    ACN_CHECK_CTX(ctx);
    ACN_REQUIRE(M >= 0, ACN_EINVAL, "acn_blend_bwd: negative M");
    if (M == 0) return ACN_OK;
    ACN_REQUIRE(d_out && sel && d_y, ACN_EINVAL, "acn_blend_bwd: null buffer");
    k_blend_bwd<<<blend_grid(ctx, M, range_or_null != nullptr), 256, 0, (cudaStream_t)stream>>>((const float4*)d_out, w, sel, M, range_or_null,
                                                                                               y_row0_or_null, (float4*)d_y);
    ACN_CHECK_LAUNCH();
    return ACN_OK;
}
