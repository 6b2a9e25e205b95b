// Device-side core of the multiresolution hash grid (models/encodings.py:308-381, torch branch).
// Shared by the standalone encode kernels (hashgrid.cu) and the fused expert kernels.
#pragma once
#include "acn_common.cuh"

// models/inr/meta_ngp.py:155-158: (x - min) / extent as a true division, then the clamp to
// [fp32(1e-6), 1 - fp32(1e-6)].
__device__ __forceinline__ float world_to_unit1(float x, float mn, float ext) {
    float v = __fdiv_rn(__fsub_rn(x, mn), ext);
    const float eps = 1e-6f;
    const float hi = 0x1.ffffdep-1f;  // fp32(1) - fp32(1e-6)
    return v != v ? v : fminf(fmaxf(v, eps), hi);
}

// models/encodings.py:308-316.  int64 arithmetic modulo 2^log2T == uint32 wrap-around + mask.
__device__ __forceinline__ uint32_t grid_hash(uint32_t ix, uint32_t iy, uint32_t iz, uint32_t mask) {
    return (ix ^ (iy * 2654435761u) ^ (iz * 805459861u)) & mask;
}

struct GridCell {
    uint32_t x0, y0, z0;  // floor corner (two's complement of the int64 floor)
    float wx, wy, wz;     // interpolation weights (after smoothstep if selected)
};

// models/encodings.py:333-369: s = x01 * res (rounded), floor, frac; all separately rounded.
__device__ __forceinline__ GridCell grid_cell(float x, float y, float z, float resf, int interp) {
    GridCell c;
    float sx = __fmul_rn(x, resf), sy = __fmul_rn(y, resf), sz = __fmul_rn(z, resf);
    float fx = floorf(sx), fy = floorf(sy), fz = floorf(sz);
    c.wx = __fsub_rn(sx, fx); c.wy = __fsub_rn(sy, fy); c.wz = __fsub_rn(sz, fz);
    c.x0 = (uint32_t)(int)fx; c.y0 = (uint32_t)(int)fy; c.z0 = (uint32_t)(int)fz;
    if (interp == ACN_INTERP_SMOOTHSTEP) {
        c.wx = __fmul_rn(__fmul_rn(c.wx, c.wx), __fsub_rn(3.0f, __fmul_rn(2.0f, c.wx)));
        c.wy = __fmul_rn(__fmul_rn(c.wy, c.wy), __fsub_rn(3.0f, __fmul_rn(2.0f, c.wy)));
        c.wz = __fmul_rn(__fmul_rn(c.wz, c.wz), __fsub_rn(3.0f, __fmul_rn(2.0f, c.wz)));
    }
    return c;
}

// Corner c in 0..7 has bits (x,y,z) = (c>>2, c>>1, c) & 1 -- the reference's f000..f111 naming.
__device__ __forceinline__ uint32_t grid_corner_row(const GridCell& g, int c, uint32_t mask) {
    return grid_hash(g.x0 + ((c >> 2) & 1), g.y0 + ((c >> 1) & 1), g.z0 + (c & 1), mask);
}

// models/encodings.py:373-379, un-fused: a*(1-w) + b*w.
__device__ __forceinline__ float lerp_rn(float a, float b, float w, float u /* = 1-w */) {
    return __fadd_rn(__fmul_rn(a, u), __fmul_rn(b, w));
}

// Gather + trilinear blend of one level for F = 2 (the configuration every BASELINE config uses).
__device__ __forceinline__ float2 grid_level_f2(const float2* __restrict__ level_table, const GridCell& g, uint32_t mask) {
    float2 f[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) f[c] = __ldg(level_table + grid_corner_row(g, c, mask));
    float ux = __fsub_rn(1.0f, g.wx), uy = __fsub_rn(1.0f, g.wy), uz = __fsub_rn(1.0f, g.wz);
    float2 o;
    {
        float c00 = lerp_rn(f[0].x, f[4].x, g.wx, ux), c01 = lerp_rn(f[1].x, f[5].x, g.wx, ux);
        float c10 = lerp_rn(f[2].x, f[6].x, g.wx, ux), c11 = lerp_rn(f[3].x, f[7].x, g.wx, ux);
        float c0 = lerp_rn(c00, c10, g.wy, uy), c1 = lerp_rn(c01, c11, g.wy, uy);
        o.x = lerp_rn(c0, c1, g.wz, uz);
    }
    {
        float c00 = lerp_rn(f[0].y, f[4].y, g.wx, ux), c01 = lerp_rn(f[1].y, f[5].y, g.wx, ux);
        float c10 = lerp_rn(f[2].y, f[6].y, g.wx, ux), c11 = lerp_rn(f[3].y, f[7].y, g.wx, ux);
        float c0 = lerp_rn(c00, c10, g.wy, uy), c1 = lerp_rn(c01, c11, g.wy, uy);
        o.y = lerp_rn(c0, c1, g.wz, uz);
    }
    return o;
}

// The 8 corner updates of one cell for F = 2 as 16-byte REDs on row PAIRS (the hash is x ^ (y*p1) ^ (z*p2), so the
// x-neighbours of an even x0 are rows r and r^1: one aligned pair, one RED carries both; for an odd x0 they sit in two
// pairs and each RED adds zeros to the other row of its pair -- an 8-byte and a 16-byte RED cost the L2 the same
// request).  Straight-line code: the second RED of a (y,z) corner is predicated on the parity of x0, nothing else
// branches; the earlier version (8-byte REDs for odd x0, zero tests per corner) spent ~150 instructions per cell, and
// the scatter kernels are bound by instruction issue, not by the atomic units (profiles/r02_v3_fused_bwd_ncu_*).
// acc is indexed c = (x<<2)|(y<<1)|z.  lt must be 16-byte aligned.
__device__ __forceinline__ void scatter_cell_f2(float2* __restrict__ lt, uint32_t x0, uint32_t y0, uint32_t z0, uint32_t mask,
                                                const float2* acc) {
    const uint32_t yp0 = y0 * 2654435761u, yp1 = yp0 + 2654435761u;
    const uint32_t zp0 = z0 * 805459861u, zp1 = zp0 + 805459861u;
    const bool odd = (x0 & 1u) != 0u;
    float4* lt4 = reinterpret_cast<float4*>(lt);
#pragma unroll
    for (int yz = 0; yz < 4; ++yz) {
        const uint32_t h = ((yz & 2) ? yp1 : yp0) ^ ((yz & 1) ? zp1 : zp0);
        const uint32_t ra = (x0 ^ h) & mask;
        const float2 a = acc[yz], b = acc[4 + yz];
        const float bx = odd ? 0.0f : b.x, by = odd ? 0.0f : b.y;            // even x0: b rides along in a's pair
        atomicAdd(lt4 + (ra >> 1), (ra & 1u) ? make_float4(bx, by, a.x, a.y) : make_float4(a.x, a.y, bx, by));
        if (odd) {
            const uint32_t rb = ((x0 + 1u) ^ h) & mask;
            atomicAdd(lt4 + (rb >> 1), (rb & 1u) ? make_float4(0.0f, 0.0f, b.x, b.y) : make_float4(b.x, b.y, 0.0f, 0.0f));
        }
    }
}
