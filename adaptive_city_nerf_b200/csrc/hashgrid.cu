// Stage 2: multiresolution hash-grid encode, forward and atomic-scatter backward.
//
// Forward: one thread per point walks all L levels (8 independent gathers per level, levels
// unrolled so ~32 loads are in flight per thread) and emits the point's full feature row, so
// every output line is written by one thread.  Indices are bit-exact with the reference's
// int64 hash; features follow the reference's un-fused lerp order (bit-exact in fp32).
// Backward: one thread per (point, level), level fastest, so the dL/dout reads are fully
// coalesced; the 8 corner updates go out as vector red.global.add.v2.f32 (F = 2).
#include "hashgrid.cuh"

// Where a point's position comes from: an explicit (P,>=3) array, or rays (N,8) + t (N,S), in
// which case p = o + d*t is formed here (un-fused, nerfs/ray_rendering.py:317) and the
// reference's (N*S,6) id6 tensor is never materialised.
struct PosSrc { const float* x; int xs; const float* rays; const float* t; int S; int ray_major; const int32_t* ray_major_dev; };

__device__ __forceinline__ void load_pos(const PosSrc& s, int64_t p, float& px, float& py, float& pz) {
    if (s.rays) {
        const float* ry = s.rays + 8 * (p / s.S);
        const float t = __ldg(s.t + p);
        px = __fadd_rn(__ldg(ry + 0), __fmul_rn(__ldg(ry + 3), t));
        py = __fadd_rn(__ldg(ry + 1), __fmul_rn(__ldg(ry + 4), t));
        pz = __fadd_rn(__ldg(ry + 2), __fmul_rn(__ldg(ry + 5), t));
    } else {
        px = s.x[p * s.xs]; py = s.x[p * s.xs + 1]; pz = s.x[p * s.xs + 2];
    }
}

template <int F>
__device__ __forceinline__ void load_feat(const float* __restrict__ t, uint32_t row, float* f) {
    if constexpr (F == 2) {
        float2 v = __ldg(reinterpret_cast<const float2*>(t) + row);
        f[0] = v.x; f[1] = v.y;
    } else if constexpr (F == 4) {
        float4 v = __ldg(reinterpret_cast<const float4*>(t) + row);
        f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
    } else {
#pragma unroll
        for (int i = 0; i < F; ++i) f[i] = __ldg(t + (size_t)row * F + i);
    }
}

template <typename T> __device__ __forceinline__ T to_out(float v);
template <> __device__ __forceinline__ float to_out<float>(float v) { return v; }
template <> __device__ __forceinline__ __half to_out<__half>(float v) { return __float2half_rn(v); }

template <int F, typename OutT>
__device__ __forceinline__ void hashgrid_encode_point(
    const PosSrc& pos, int64_t p, const float* __restrict__ box6,
    const float* __restrict__ table, int L, int log2T, const int32_t* __restrict__ res, int interp,
    OutT* __restrict__ out, int32_t* __restrict__ idx_out)
{
    float px, py, pz;
    load_pos(pos, p, px, py, pz);
    if (box6) {
        px = world_to_unit1(px, __ldg(box6 + 0), __ldg(box6 + 3));
        py = world_to_unit1(py, __ldg(box6 + 1), __ldg(box6 + 4));
        pz = world_to_unit1(pz, __ldg(box6 + 2), __ldg(box6 + 5));
    }
    const uint32_t mask = (1u << log2T) - 1u;
    OutT* orow = out + (size_t)p * L * F;
#pragma unroll 4
    for (int l = 0; l < L; ++l) {
        const float resf = (float)__ldg(res + l);
        const float* lt = table + ((size_t)l << log2T) * F;
        float o[F];
        if (interp == ACN_INTERP_NEAREST) {
            // models/encodings.py:342-345: torch.round = round half to even
            uint32_t ix = (uint32_t)(int)rintf(__fmul_rn(px, resf));
            uint32_t iy = (uint32_t)(int)rintf(__fmul_rn(py, resf));
            uint32_t iz = (uint32_t)(int)rintf(__fmul_rn(pz, resf));
            uint32_t row = grid_hash(ix, iy, iz, mask);
            if (idx_out) idx_out[((size_t)p * L + l) * 8] = (int32_t)(row + ((uint32_t)l << log2T));
            load_feat<F>(lt, row, o);
        } else {
            GridCell g = grid_cell(px, py, pz, resf, interp);
            float f[8][F];
            // x-neighbours of an even x0 are rows r and r^1 (the hash is x ^ ...): one aligned 16-byte load fetches both.
            // Taken only when the whole warp agrees (no divergence): that is the frame case, where a warp holds one sample
            // of 32 adjacent pixels and shares x0 up to the mid levels.  For unrelated points it was measured neutral
            // (3.96 vs 4.01 ms) and the vote almost never passes.
            bool paired = false;
            if constexpr (F == 2) paired = !idx_out && __all_sync(__activemask(), !(g.x0 & 1u));
            if (paired) {
                if constexpr (F == 2) {
                    const uint32_t yp0 = g.y0 * 2654435761u, yp1 = yp0 + 2654435761u;
                    const uint32_t zp0 = g.z0 * 805459861u, zp1 = zp0 + 805459861u;
#pragma unroll
                    for (int yz = 0; yz < 4; ++yz) {
                        const uint32_t h = ((yz & 2) ? yp1 : yp0) ^ ((yz & 1) ? zp1 : zp0);
                        const uint32_t r0 = (g.x0 ^ h) & mask;
                        const float4 v = __ldg(reinterpret_cast<const float4*>(lt) + (r0 >> 1));
                        const bool odd = r0 & 1u;
                        f[yz][0] = odd ? v.z : v.x; f[yz][1] = odd ? v.w : v.y;
                        f[4 + yz][0] = odd ? v.x : v.z; f[4 + yz][1] = odd ? v.y : v.w;
                    }
                }
            } else {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    uint32_t row = grid_corner_row(g, c, mask);
                    if (idx_out) idx_out[((size_t)p * L + l) * 8 + c] = (int32_t)(row + ((uint32_t)l << log2T));
                    load_feat<F>(lt, row, f[c]);
                }
            }
            if constexpr (sizeof(OutT) == 2) {
                // fp16 output: fused lerps a + w (b - a); the ~1e-7 difference to the reference's un-fused form
                // is three orders below the fp16 rounding of the result
#pragma unroll
                for (int i = 0; i < F; ++i) {
                    float c00 = fmaf(g.wx, f[4][i] - f[0][i], f[0][i]), c01 = fmaf(g.wx, f[5][i] - f[1][i], f[1][i]);
                    float c10 = fmaf(g.wx, f[6][i] - f[2][i], f[2][i]), c11 = fmaf(g.wx, f[7][i] - f[3][i], f[3][i]);
                    float c0 = fmaf(g.wy, c10 - c00, c00), c1 = fmaf(g.wy, c11 - c01, c01);
                    o[i] = fmaf(g.wz, c1 - c0, c0);
                }
            } else {
                float ux = __fsub_rn(1.0f, g.wx), uy = __fsub_rn(1.0f, g.wy), uz = __fsub_rn(1.0f, g.wz);
#pragma unroll
                for (int i = 0; i < F; ++i) {
                    float c00 = lerp_rn(f[0][i], f[4][i], g.wx, ux), c01 = lerp_rn(f[1][i], f[5][i], g.wx, ux);
                    float c10 = lerp_rn(f[2][i], f[6][i], g.wx, ux), c11 = lerp_rn(f[3][i], f[7][i], g.wx, ux);
                    float c0 = lerp_rn(c00, c10, g.wy, uy), c1 = lerp_rn(c01, c11, g.wy, uy);
                    o[i] = lerp_rn(c0, c1, g.wz, uz);
                }
            }
        }
        if constexpr (F == 2 && sizeof(OutT) == 4) {
            reinterpret_cast<float2*>(orow)[l] = make_float2(o[0], o[1]);
        } else if constexpr (F == 2 && sizeof(OutT) == 2) {
            reinterpret_cast<__half2*>(orow)[l] = __floats2half2_rn(o[0], o[1]);
        } else {
#pragma unroll
            for (int i = 0; i < F; ++i) orow[l * F + i] = to_out<OutT>(o[i]);
        }
    }
}

// Thread per point.  Rows: all of [0, P), or -- `range` (device, 2 int32) given -- rows [range[0], range[1]) with P only
// an upper bound on their number (the routed path: an expert's bucket, whose size the host never reads); the grid is
// then capped and strides.
template <int F, typename OutT>
__global__ void __launch_bounds__(256) k_hashgrid_fwd(
    PosSrc pos, int64_t P, const int32_t* __restrict__ range, const float* __restrict__ box6,
    const float* __restrict__ table, int L, int log2T, const int32_t* __restrict__ res, int interp,
    OutT* __restrict__ out, int32_t* __restrict__ idx_out)
{
    if (range) {
        const int64_t r0 = __ldg(range), r1 = __ldg(range + 1);
        for (int64_t p = r0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < r1; p += (int64_t)gridDim.x * blockDim.x)
            hashgrid_encode_point<F, OutT>(pos, p, box6, table, L, log2T, res, interp, out, idx_out);
        return;
    }
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pos.ray_major_dev ? (__ldg(pos.ray_major_dev) != 0) : (pos.ray_major != 0)) {
        // frames (consecutive rays = adjacent pixels): a warp takes ONE sample of 32 neighbouring rays, so its lanes
        // sit in the same or neighbouring cells up to the levels whose cells are as small as a pixel footprint
        const int sgroups = (pos.S + 7) / 8;
        const int64_t r = (int64_t)(blockIdx.x / sgroups) * 32 + (threadIdx.x & 31);
        const int si = (blockIdx.x % sgroups) * 8 + (threadIdx.x >> 5);
        if (si >= pos.S) return;
        p = r * pos.S + si;
    }
    if (p >= P) return;
    hashgrid_encode_point<F, OutT>(pos, p, box6, table, L, log2T, res, interp, out, idx_out);
}

template <int F>
__device__ __forceinline__ void scatter_add(float* __restrict__ t, uint32_t row, const float* g, float w) {
    if constexpr (F == 2) {
        atomicAdd(reinterpret_cast<float2*>(t) + row, make_float2(g[0] * w, g[1] * w));
    } else if constexpr (F == 4) {
        atomicAdd(reinterpret_cast<float4*>(t) + row, make_float4(g[0] * w, g[1] * w, g[2] * w, g[3] * w));
    } else {
#pragma unroll
        for (int i = 0; i < F; ++i) atomicAdd(t + (size_t)row * F + i, g[i] * w);
    }
}

template <typename T> __device__ __forceinline__ float from_in(T v);
template <> __device__ __forceinline__ float from_in<float>(float v) { return v; }
template <> __device__ __forceinline__ float from_in<__half>(__half v) { return __half2float(v); }

template <int F, typename InT>
__device__ __forceinline__ void hashgrid_scatter_point_level(
    const PosSrc& pos, int64_t p, int l, const float* __restrict__ box6, int L, int log2T,
    const int32_t* __restrict__ res, int interp, const InT* __restrict__ dout, float* __restrict__ dtable);

template <int F, typename InT>
__global__ void __launch_bounds__(256) k_hashgrid_bwd(
    PosSrc pos, int64_t P, const int32_t* __restrict__ range, const float* __restrict__ box6, int L, int log2T,
    const int32_t* __restrict__ res, int interp, const InT* __restrict__ dout, float* __restrict__ dtable)
{
    // Blocks are visited in a scrambled order (blockIdx * prime mod gridDim, a bijection because the prime exceeds
    // any grid size): points usually arrive in ray / sample order, and then the blocks that run at the same time hit
    // the same few hundred rows of the coarse levels and serialise in the L2 atomic units (measured: 21.8 ms for a
    // routed batch in bucket order vs 8.6 ms for the same points shuffled).
    const int64_t blk = (int64_t)(((uint64_t)blockIdx.x * 2654435761ull) % (uint64_t)gridDim.x);
    int64_t r0 = 0, Pn = P;
    if (range) { r0 = __ldg(range); Pn = __ldg(range + 1) - r0; }      // rows [range[0], range[1]); P is then only a bound
    for (int64_t idx = blk * blockDim.x + threadIdx.x; idx < Pn * L; idx += (int64_t)gridDim.x * blockDim.x)
        hashgrid_scatter_point_level<F, InT>(pos, r0 + idx / L, (int)(idx % L), box6, L, log2T, res, interp, dout, dtable);
}

template <int F, typename InT>
__device__ __forceinline__ void hashgrid_scatter_point_level(
    const PosSrc& pos, int64_t p, int l, const float* __restrict__ box6, int L, int log2T,
    const int32_t* __restrict__ res, int interp, const InT* __restrict__ dout, float* __restrict__ dtable)
{
    const int64_t idx = p * L + l;
    float g[F];
    bool any = false;
#pragma unroll
    for (int i = 0; i < F; ++i) { g[i] = from_in<InT>(dout[idx * F + i]); any |= (g[i] != 0.0f); }
    if (!any) return;  // zero gradient rows (e.g. fully occluded samples) scatter nothing
    float px, py, pz;
    load_pos(pos, p, px, py, pz);
    if (box6) {
        px = world_to_unit1(px, __ldg(box6 + 0), __ldg(box6 + 3));
        py = world_to_unit1(py, __ldg(box6 + 1), __ldg(box6 + 4));
        pz = world_to_unit1(pz, __ldg(box6 + 2), __ldg(box6 + 5));
    }
    const uint32_t mask = (1u << log2T) - 1u;
    const float resf = (float)__ldg(res + l);
    float* lt = dtable + ((size_t)l << log2T) * F;
    if (interp == ACN_INTERP_NEAREST) {
        uint32_t ix = (uint32_t)(int)rintf(__fmul_rn(px, resf));
        uint32_t iy = (uint32_t)(int)rintf(__fmul_rn(py, resf));
        uint32_t iz = (uint32_t)(int)rintf(__fmul_rn(pz, resf));
        scatter_add<F>(lt, grid_hash(ix, iy, iz, mask), g, 1.0f);
        return;
    }
    GridCell c = grid_cell(px, py, pz, resf, interp);
    float wx[2] = { 1.0f - c.wx, c.wx }, wy[2] = { 1.0f - c.wy, c.wy }, wz[2] = { 1.0f - c.wz, c.wz };
    if constexpr (F == 2) {
        float2 acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float w = wz[k & 1] * wy[(k >> 1) & 1] * wx[(k >> 2) & 1];
            acc[k] = make_float2(g[0] * w, g[1] * w);
        }
        scatter_cell_f2(reinterpret_cast<float2*>(lt), c.x0, c.y0, c.z0, mask, acc);
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float w = wz[k & 1] * wy[(k >> 1) & 1] * wx[(k >> 2) & 1];
            scatter_add<F>(lt, grid_corner_row(c, k, mask), g, w);
        }
    }
}

// Ray-marching backward (Linear / Smoothstep, F = 2): 16 lanes per ray, lane = level, each lane
// walks its ray's S samples IN SEQUENCE and keeps the 8 corner gradients of the cell it is
// currently in in registers; they go out as 8 vector REDs only when the ray leaves the cell.
//   * at coarse levels a ray stays in one cell for several samples -> several times fewer atomics;
//   * every ray starts at its own rotated sample offset, so neighbouring rays are NOT at the same
//     depth at the same time -- in plain sample order all 2^18 rays hammer the same ~2*17^2 table
//     rows of a coarse level together and the L2 atomic units serialise them (measured: level 0
//     alone cost 4.0 ms of a 28.8 ms scatter, 13.0 ms after shuffling the points);
//   * the world->unit divides are computed once per sample by 3 lanes and shared by shuffle;
//   * a half-warp reads one 128-byte dL/denc row per step (coalesced).
template <typename InT>
__global__ void __launch_bounds__(256) k_hashgrid_bwd_march(
    const float* __restrict__ rays8, const float* __restrict__ t_vals, int64_t N, int S,
    const float* __restrict__ box6, int L, int log2T, const int32_t* __restrict__ res, int interp,
    const InT* __restrict__ dout, float* __restrict__ dtable)
{
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t ray = gid >> 4;
    const int l = (int)(gid & 15);
    const bool ray_on = ray < N;
    if (!ray_on) ray = N - 1;                       // keep the lane alive for the shuffles
    const bool on = ray_on && l < L;
    const uint32_t mask = (1u << log2T) - 1u;
    const float resf = (float)__ldg(res + (l < L ? l : 0));
    float2* lt = reinterpret_cast<float2*>(dtable) + ((size_t)(l < L ? l : 0) << log2T);
    // lanes 0..2 of each 16-lane group own one coordinate each
    const int comp = l < 3 ? l : 0;
    const float o_c = __ldg(rays8 + 8 * ray + comp), d_c = __ldg(rays8 + 8 * ray + 3 + comp);
    const float mn_c = box6 ? __ldg(box6 + comp) : 0.0f, ex_c = box6 ? __ldg(box6 + 3 + comp) : 1.0f;
    const float* trow = t_vals + ray * S;
    const int start = (int)(((uint32_t)ray * 2654435761u) >> 8) % S;

    uint32_t cx = 0, cy = 0, cz = 0;
    bool have = false;
    float2 acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = make_float2(0.f, 0.f);

    auto flush = [&]() {
        scatter_cell_f2(lt, cx, cy, cz, mask, acc);
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = make_float2(0.f, 0.f);
    };

    for (int it = 0; it < S; ++it) {
        int s = it + start; if (s >= S) s -= S;
        const float t = __ldg(trow + s);
        float pc = __fadd_rn(o_c, __fmul_rn(d_c, t));
        if (box6) pc = world_to_unit1(pc, mn_c, ex_c);
        const float px = __shfl_sync(0xffffffffu, pc, 0, 16);
        const float py = __shfl_sync(0xffffffffu, pc, 1, 16);
        const float pz = __shfl_sync(0xffffffffu, pc, 2, 16);
        if (!on) continue;
        float2 g;
        if constexpr (sizeof(InT) == 4) g = __ldg(reinterpret_cast<const float2*>(dout) + (ray * S + s) * L + l);
        else g = __half22float2(__ldg(reinterpret_cast<const __half2*>(dout) + (ray * S + s) * L + l));
        if (g.x == 0.0f && g.y == 0.0f) continue;
        GridCell c = grid_cell(px, py, pz, resf, interp);
        if (!(have && c.x0 == cx && c.y0 == cy && c.z0 == cz)) {
            if (have) flush();
            cx = c.x0; cy = c.y0; cz = c.z0; have = true;
        }
        const float wx[2] = { 1.0f - c.wx, c.wx }, wy[2] = { 1.0f - c.wy, c.wy }, wz[2] = { 1.0f - c.wz, c.wz };
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float w = wz[k & 1] * wy[(k >> 1) & 1] * wx[(k >> 2) & 1];
            acc[k].x = fmaf(g.x, w, acc[k].x);
            acc[k].y = fmaf(g.y, w, acc[k].y);
        }
    }
    if (have) flush();
}

// The same run-length scatter for an arbitrary (P,>=3) point list -- the routed path, where an expert's rows are
// bucketed in (block of rays, ray, sample) order, so consecutive rows are mostly consecutive samples of one ray.  A
// 16-lane group walks SEG consecutive rows from a rotated start; nothing depends on ray structure: a row in a different
// cell simply flushes the accumulators.  (The per-(point, level) kernel above stays for F != 2, "Nearest" and as the
// cross-check; measured on the 4-expert routed step: 11.6 ms -> see DESIGN.md.)
template <typename InT>
__global__ void __launch_bounds__(256) k_hashgrid_bwd_march_pts(
    const float* __restrict__ x, int xs, int64_t P, const int32_t* __restrict__ range, int seg,
    const float* __restrict__ box6, int L, int log2T, const int32_t* __restrict__ res, int interp,
    const InT* __restrict__ dout, float* __restrict__ dtable)
{
    if (range) {            // rows [range[0], range[1]) of x / dout; P is then only an upper bound on their number
        const int64_t r0 = __ldg(range);
        P = __ldg(range + 1) - r0;
        x += r0 * xs;
        dout += r0 * L * 2;
    }
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int l = (int)(gid & 15);
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t nseg = (P + seg - 1) / seg;
  for (int64_t wp = gid >> 5; 2 * wp < nseg; wp += nwarps) {      // a warp takes two segments per round (warp-uniform trip count)
    const int64_t grp = 2 * wp + ((gid >> 4) & 1);
    const int64_t row0 = grp * seg;
    const int n = row0 < P ? (int)((P - row0) < seg ? (P - row0) : seg) : 0;   // uniform over the 16-lane group
    const bool on = l < L;
    const uint32_t mask = (1u << log2T) - 1u;
    const float resf = (float)__ldg(res + (l < L ? l : 0));
    float2* lt = reinterpret_cast<float2*>(dtable) + ((size_t)(l < L ? l : 0) << log2T);
    const int comp = l < 3 ? l : 0;
    const float mn_c = box6 ? __ldg(box6 + comp) : 0.0f, ex_c = box6 ? __ldg(box6 + 3 + comp) : 1.0f;
    const int start = n > 0 ? (int)(((uint32_t)grp * 2654435761u) >> 8) % n : 0;

    uint32_t cx = 0, cy = 0, cz = 0;
    bool have = false;
    float2 acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = make_float2(0.f, 0.f);

    // the two 16-lane groups of a warp may own segments of different length: iterate to the longer one so the
    // shuffles below stay convergent
    const int n_other = __shfl_xor_sync(0xffffffffu, n, 16);
    const int n_iter = n > n_other ? n : n_other;
    for (int it = 0; it < n_iter; ++it) {
        const bool live = it < n;
        int sidx = it + start; if (sidx >= n) sidx -= n;
        const int64_t prow = row0 + (live ? sidx : 0);
        float pc = (live || n > 0) ? __ldg(x + (row0 < P ? prow : 0) * xs + comp) : 0.0f;
        if (box6) pc = world_to_unit1(pc, mn_c, ex_c);
        const float px = __shfl_sync(0xffffffffu, pc, 0, 16);
        const float py = __shfl_sync(0xffffffffu, pc, 1, 16);
        const float pz = __shfl_sync(0xffffffffu, pc, 2, 16);
        if (!on || !live) continue;
        float2 g;
        if constexpr (sizeof(InT) == 4) g = __ldg(reinterpret_cast<const float2*>(dout) + prow * L + l);
        else g = __half22float2(__ldg(reinterpret_cast<const __half2*>(dout) + prow * L + l));
        if (g.x == 0.0f && g.y == 0.0f) continue;
        GridCell c = grid_cell(px, py, pz, resf, interp);
        if (!(have && c.x0 == cx && c.y0 == cy && c.z0 == cz)) {
            if (have) {
                scatter_cell_f2(lt, cx, cy, cz, mask, acc);
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] = make_float2(0.f, 0.f);
            }
            cx = c.x0; cy = c.y0; cz = c.z0; have = true;
        }
        const float wx[2] = { 1.0f - c.wx, c.wx }, wy[2] = { 1.0f - c.wy, c.wy }, wz[2] = { 1.0f - c.wz, c.wz };
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float w = wz[k & 1] * wy[(k >> 1) & 1] * wx[(k >> 2) & 1];
            acc[k].x = fmaf(g.x, w, acc[k].x);
            acc[k].y = fmaf(g.y, w, acc[k].y);
        }
    }
    if (have) scatter_cell_f2(lt, cx, cy, cz, mask, acc);
  }
}

// ------------------------------------------------------------------------------------------ C ABI
static int check_grid_args(const char* fn, int64_t P, int xs, int L, int F, int log2T, const void* res, int interp) {
    ACN_REQUIRE(P >= 0 && xs >= 3, ACN_EINVAL, "%s: bad P / x_stride", fn);
    ACN_REQUIRE(L >= 1 && L <= 64 && res, ACN_EINVAL, "%s: bad level count / res table", fn);
    ACN_REQUIRE(log2T >= 1 && log2T <= 24, ACN_EUNSUPPORTED, "%s: log2_hashmap_size %d outside [1,24]", fn, log2T);
    ACN_REQUIRE(F == 1 || F == 2 || F == 4 || F == 8, ACN_EUNSUPPORTED, "%s: features_per_level %d not in {1,2,4,8}", fn, F);
    ACN_REQUIRE(interp >= 0 && interp <= 2, ACN_EINVAL, "%s: bad interpolation mode", fn);
    ACN_REQUIRE(((int64_t)L << log2T) <= ((int64_t)1 << 31), ACN_EUNSUPPORTED, "%s: table rows exceed int32", fn);
    return ACN_OK;
}

#define DISPATCH_F(F_, CALL)                    \
    switch (F_) {                               \
        case 1: { constexpr int FF = 1; CALL; } break; \
        case 2: { constexpr int FF = 2; CALL; } break; \
        case 4: { constexpr int FF = 4; CALL; } break; \
        default: { constexpr int FF = 8; CALL; } break; \
    }

// launches over a device-side row range stride a capped grid (the host does not know how many rows there are)
static int range_grid(acn_ctx* ctx, int64_t blocks) {
    const int64_t cap = (int64_t)ctx->sm_count * 16;
    return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

static int hashgrid_fwd_impl(acn_ctx* ctx, const char* fn, PosSrc pos, int64_t P, const int32_t* range, const float* box6_or_null,
                             const float* table, int L, int F, int log2T, const int32_t* res, int interp,
                             void* out, int out_dtype, int32_t* idx_out_or_null, acn_stream stream) {
    ACN_CHECK_CTX(ctx);
    int rc = check_grid_args(fn, P, pos.rays ? 3 : pos.xs, L, F, log2T, res, interp);
    if (rc) return rc;
    ACN_REQUIRE(out_dtype == ACN_F32 || out_dtype == ACN_F16, ACN_EINVAL, "%s: bad out dtype", fn);
    if (P == 0) return ACN_OK;
    ACN_REQUIRE((pos.x || (pos.rays && pos.t && pos.S >= 1)) && table && out, ACN_EINVAL, "%s: null buffer", fn);
    ACN_REQUIRE(((uintptr_t)table & 15) == 0 && ((uintptr_t)out & 7) == 0, ACN_EINVAL, "%s: misaligned table/out", fn);
    const int block = 256;
    // the ray-major grid covers the sample-major one too, so a launch whose mapping is decided on the device uses it
    ACN_REQUIRE(!range || pos.x, ACN_EINVAL, "%s: a row range needs explicit positions", fn);
    const int grid = range ? range_grid(ctx, (P + block - 1) / block)
                   : (pos.ray_major || pos.ray_major_dev) ? (int)(((P / pos.S + 31) / 32) * ((pos.S + 7) / 8)) : acn_grid_1d(P, block);
    cudaStream_t st = (cudaStream_t)stream;
    if (out_dtype == ACN_F32) {
        DISPATCH_F(F, (k_hashgrid_fwd<FF, float><<<grid, block, 0, st>>>(pos, P, range, box6_or_null, table, L, log2T, res,
                                                                        interp, (float*)out, idx_out_or_null)));
    } else {
        DISPATCH_F(F, (k_hashgrid_fwd<FF, __half><<<grid, block, 0, st>>>(pos, P, range, box6_or_null, table, L, log2T, res,
                                                                         interp, (__half*)out, idx_out_or_null)));
    }
    ACN_CHECK_LAUNCH();
    return ACN_OK;
}

static int hashgrid_bwd_impl(acn_ctx* ctx, const char* fn, PosSrc pos, int64_t P, const int32_t* range, const float* box6_or_null, int L,
                             int F, int log2T, const int32_t* res, int interp, const void* dout, int dout_dtype,
                             float* dtable, acn_stream stream, bool plain = false) {
    ACN_CHECK_CTX(ctx);
    int rc = check_grid_args(fn, P, pos.rays ? 3 : pos.xs, L, F, log2T, res, interp);
    if (rc) return rc;
    ACN_REQUIRE(dout_dtype == ACN_F32 || dout_dtype == ACN_F16, ACN_EINVAL, "%s: bad dout dtype", fn);
    if (P == 0) return ACN_OK;
    ACN_REQUIRE((pos.x || (pos.rays && pos.t && pos.S >= 1)) && dout && dtable, ACN_EINVAL, "%s: null buffer", fn);
    ACN_REQUIRE(((uintptr_t)dtable & 15) == 0, ACN_EINVAL, "%s: misaligned dtable", fn);
    cudaStream_t st = (cudaStream_t)stream;
    if (pos.x && F == 2 && L <= 16 && interp != ACN_INTERP_NEAREST && !plain) {
        ACN_REQUIRE(((uintptr_t)dout & 7) == 0, ACN_EINVAL, "%s: misaligned dout", fn);
        const int seg = 64;
        const int64_t blocks_m = (((P + seg - 1) / seg) * 16 + 255) / 256;
        const int grid_m = range ? range_grid(ctx, blocks_m) : acn_grid_1d(blocks_m * 256, 256);
        if (dout_dtype == ACN_F32)
            k_hashgrid_bwd_march_pts<float><<<grid_m, 256, 0, st>>>(pos.x, pos.xs, P, range, seg, box6_or_null, L, log2T, res, interp,
                                                                   (const float*)dout, dtable);
        else
            k_hashgrid_bwd_march_pts<__half><<<grid_m, 256, 0, st>>>(pos.x, pos.xs, P, range, seg, box6_or_null, L, log2T, res, interp,
                                                                    (const __half*)dout, dtable);
        ACN_CHECK_LAUNCH();
        return ACN_OK;
    }
    const int block = 256;
    ACN_REQUIRE(!range || pos.x, ACN_EINVAL, "%s: a row range needs explicit positions", fn);
    const int grid = range ? range_grid(ctx, (P * L + block - 1) / block) : acn_grid_1d(P * L, block);
    if (dout_dtype == ACN_F32) {
        DISPATCH_F(F, (k_hashgrid_bwd<FF, float><<<grid, block, 0, st>>>(pos, P, range, box6_or_null, L, log2T, res, interp,
                                                                        (const float*)dout, dtable)));
    } else {
        DISPATCH_F(F, (k_hashgrid_bwd<FF, __half><<<grid, block, 0, st>>>(pos, P, range, box6_or_null, L, log2T, res, interp,
                                                                         (const __half*)dout, dtable)));
    }
    ACN_CHECK_LAUNCH();
    return ACN_OK;
}

extern "C" int acn_hashgrid_fwd(acn_ctx* ctx, const float* x, int64_t P, int x_stride, const int32_t* range_or_null,
                                const float* box6_or_null, const float* table, int L, int F, int log2T, const int32_t* res,
                                int interp, void* out, int out_dtype, int32_t* idx_out_or_null, acn_stream stream) {
    PosSrc pos{ x, x_stride, nullptr, nullptr, 0, 0, nullptr };
    return hashgrid_fwd_impl(ctx, "acn_hashgrid_fwd", pos, P, range_or_null, box6_or_null, table, L, F, log2T, res, interp, out,
                             out_dtype, idx_out_or_null, stream);
}

extern "C" int acn_hashgrid_bwd(acn_ctx* ctx, const float* x, int64_t P, int x_stride, const int32_t* range_or_null,
                                const float* box6_or_null, int L, int F, int log2T, const int32_t* res, int interp,
                                const void* dout, int dout_dtype, float* dtable, acn_stream stream) {
    PosSrc pos{ x, x_stride, nullptr, nullptr, 0, 0, nullptr };
    return hashgrid_bwd_impl(ctx, "acn_hashgrid_bwd", pos, P, range_or_null, box6_or_null, L, F, log2T, res, interp, dout, dout_dtype,
                             dtable, stream);
}

extern "C" int acn_hashgrid_fwd_rays(acn_ctx* ctx, const float* rays8, const float* t_vals, int64_t N, int S,
                                     const float* box6_or_null, const float* table, int L, int F, int log2T,
                                     const int32_t* res, int interp, void* out, int out_dtype, int ray_major,
                                     const int32_t* ray_major_dev_or_null, acn_stream stream) {
    ACN_REQUIRE(N >= 0 && S >= 1, ACN_EINVAL, "acn_hashgrid_fwd_rays: bad N / S");
    PosSrc pos{ nullptr, 0, rays8, t_vals, S, ray_major ? 1 : 0, ray_major_dev_or_null };
    return hashgrid_fwd_impl(ctx, "acn_hashgrid_fwd_rays", pos, N * S, nullptr, box6_or_null, table, L, F, log2T, res, interp, out,
                             out_dtype, nullptr, stream);
}

extern "C" int acn_hashgrid_bwd_rays(acn_ctx* ctx, const float* rays8, const float* t_vals, int64_t N, int S,
                                     const float* box6_or_null, int L, int F, int log2T, const int32_t* res, int interp,
                                     const void* dout, int dout_dtype, float* dtable, acn_stream stream) {
    ACN_REQUIRE(N >= 0 && S >= 1, ACN_EINVAL, "acn_hashgrid_bwd_rays: bad N / S");
    if (F == 2 && L <= 16 && interp != ACN_INTERP_NEAREST) {
        ACN_CHECK_CTX(ctx);
        int rc = check_grid_args("acn_hashgrid_bwd_rays", N * S, 3, L, F, log2T, res, interp);
        if (rc) return rc;
        ACN_REQUIRE(dout_dtype == ACN_F32 || dout_dtype == ACN_F16, ACN_EINVAL, "acn_hashgrid_bwd_rays: bad dout dtype");
        if (N == 0) return ACN_OK;
        ACN_REQUIRE(rays8 && t_vals && dout && dtable, ACN_EINVAL, "acn_hashgrid_bwd_rays: null buffer");
        ACN_REQUIRE((((uintptr_t)dtable | (uintptr_t)dout) & 7) == 0, ACN_EINVAL, "acn_hashgrid_bwd_rays: misaligned dtable/dout");
        const int grid = acn_grid_1d(N * 16, 256);
        cudaStream_t st = (cudaStream_t)stream;
        if (dout_dtype == ACN_F32)
            k_hashgrid_bwd_march<float><<<grid, 256, 0, st>>>(rays8, t_vals, N, S, box6_or_null, L, log2T, res, interp,
                                                              (const float*)dout, dtable);
        else
            k_hashgrid_bwd_march<__half><<<grid, 256, 0, st>>>(rays8, t_vals, N, S, box6_or_null, L, log2T, res, interp,
                                                               (const __half*)dout, dtable);
        ACN_CHECK_LAUNCH();
        return ACN_OK;
    }
    PosSrc pos{ nullptr, 0, rays8, t_vals, S, 0, nullptr };
    return hashgrid_bwd_impl(ctx, "acn_hashgrid_bwd_rays", pos, N * S, nullptr, box6_or_null, L, F, log2T, res, interp, dout, dout_dtype,
                             dtable, stream);
}

extern "C" int acn_hashgrid_bwd_plain(acn_ctx* ctx, const float* x, int64_t P, int x_stride, const float* box6_or_null, int L,
                                      int F, int log2T, const int32_t* res, int interp, const void* dout, int dout_dtype,
                                      float* dtable, acn_stream stream) {
    PosSrc pos{ x, x_stride, nullptr, nullptr, 0, 0, nullptr };
    return hashgrid_bwd_impl(ctx, "acn_hashgrid_bwd_plain", pos, P, nullptr, box6_or_null, L, F, log2T, res, interp, dout, dout_dtype, dtable,
                             stream, true);
}
