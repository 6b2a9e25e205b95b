// Memory-system ceilings for the hash-grid kernels, measured rather than assumed ("measure the L2 gather
// peak with a micro-benchmark before quoting a fraction").  The hash-grid gathers and the gradient scatter touch a
// 64 MiB table that is L2-resident on B200, 8 bytes (one F = 2 row) at a time at pseudo-random rows: HBM bandwidth is
// not what bounds them.  These kernels issue the same access shape with nothing else in the way:
//   mode 0  independent scattered loads (8 or 16 bytes), 8 in flight per thread per iteration
//   mode 1  scattered red.global.add (v2.f32 = 8 bytes, v4.f32 = 16 bytes), fire-and-forget
// tools/l2_peak.py times them with CUDA events at full occupancy and reports accesses/s and GB/s.
#include "acn_common.cuh"
#include "../../../include/acn_b200_debug.h"

namespace {

__device__ __forceinline__ uint32_t mix(uint32_t x) {       // a cheap avalanche (lowbias32)
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

template <int BYTES>
__global__ void __launch_bounds__(256) k_l2_gather(const uint8_t* __restrict__ buf, uint32_t mask, int iters, float* __restrict__ sink) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t s = mix(tid * 2654435761u + 1u);
    float acc = 0.0f;
    for (int it = 0; it < iters; ++it) {
        uint32_t off[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { s = s * 1664525u + 1013904223u; off[j] = mix(s) & mask & ~(uint32_t)(BYTES - 1); }
        if constexpr (BYTES == 8) {
            float2 v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __ldg(reinterpret_cast<const float2*>(buf + off[j]));
#pragma unroll
            for (int j = 0; j < 8; ++j) acc += v[j].x + v[j].y;
        } else {
            float4 v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __ldg(reinterpret_cast<const float4*>(buf + off[j]));
#pragma unroll
            for (int j = 0; j < 8; ++j) acc += v[j].x + v[j].y + v[j].z + v[j].w;
        }
    }
    if (acc == 123.456f) sink[0] = acc;      // never true for the zero-filled buffer; keeps the loads alive
}

template <int BYTES>
__global__ void __launch_bounds__(256) k_l2_red(uint8_t* __restrict__ buf, uint32_t mask, int iters) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t s = mix(tid * 2654435761u + 7u);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s = s * 1664525u + 1013904223u;
            const uint32_t off = mix(s) & mask & ~(uint32_t)(BYTES - 1);
            if constexpr (BYTES == 8) atomicAdd(reinterpret_cast<float2*>(buf + off), make_float2(1.0f, 1.0f));
            else atomicAdd(reinterpret_cast<float4*>(buf + off), make_float4(1.0f, 1.0f, 1.0f, 1.0f));
        }
    }
}

// RED rate as a function of the warps an SM has and of how full their RED instructions are: `active_lanes` of every
// warp issue the 16-byte REDs, the others skip them (predicated), `work` dependent integer operations separate two REDs
// of a thread.  What the fused backward can get out of its 16 resident warps is read off this, not off the
// full-occupancy peak.
__global__ void k_red_probe(uint8_t* __restrict__ buf, uint32_t mask, int iters, int active_lanes, int work) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const bool on = (int)(threadIdx.x & 31) < active_lanes;
    uint32_t s = mix(tid * 2654435761u + 7u);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s = s * 1664525u + 1013904223u;
            for (int k = 0; k < work; ++k) s = s * 1664525u + 1013904223u + (uint32_t)k;
            const uint32_t off = mix(s) & mask & ~15u;
            if (on) atomicAdd(reinterpret_cast<float4*>(buf + off), make_float4(1.0f, 1.0f, 1.0f, 1.0f));
        }
    }
}

}  // namespace

extern "C" int acn_debug_red_probe(acn_ctx* ctx, void* buf, int64_t buf_bytes, int iters, int grid, int block, int active_lanes,
                                   int work, acn_stream stream) {
    ACN_CHECK_CTX(ctx);
    ACN_REQUIRE(buf && buf_bytes >= 4096 && (buf_bytes & (buf_bytes - 1)) == 0 && buf_bytes <= ((int64_t)1 << 32), ACN_EINVAL,
                "acn_debug_red_probe: buf_bytes must be a power of two in [4 KiB, 4 GiB]");
    ACN_REQUIRE(iters >= 1 && grid >= 1 && block >= 32 && block <= 1024 && block % 32 == 0 && active_lanes >= 1 && active_lanes <= 32 && work >= 0,
                ACN_EINVAL, "acn_debug_red_probe: bad arguments");
    k_red_probe<<<grid, block, 0, (cudaStream_t)stream>>>((uint8_t*)buf, (uint32_t)(buf_bytes - 1), iters, active_lanes, work);
    ACN_CHECK_LAUNCH();
    return ACN_OK;
}

extern "C" int acn_debug_l2_probe(acn_ctx* ctx, int mode, int bytes_per_access, void* buf, int64_t buf_bytes, int iters,
                                  int grid, float* sink, acn_stream stream) {
    ACN_CHECK_CTX(ctx);
    ACN_REQUIRE(buf && buf_bytes >= 4096 && (buf_bytes & (buf_bytes - 1)) == 0 && buf_bytes <= ((int64_t)1 << 32), ACN_EINVAL,
                "acn_debug_l2_probe: buf_bytes must be a power of two in [4 KiB, 4 GiB]");
    ACN_REQUIRE((bytes_per_access == 8 || bytes_per_access == 16) && iters >= 1 && grid >= 1 && (mode == 0 || mode == 1), ACN_EINVAL,
                "acn_debug_l2_probe: bad mode / access size / iters / grid");
    ACN_REQUIRE(((uintptr_t)buf & 15) == 0 && (mode == 1 || sink), ACN_EINVAL, "acn_debug_l2_probe: misaligned buffer or null sink");
    const uint32_t mask = (uint32_t)(buf_bytes - 1);
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == 0) {
        if (bytes_per_access == 8) k_l2_gather<8><<<grid, 256, 0, st>>>((const uint8_t*)buf, mask, iters, sink);
        else k_l2_gather<16><<<grid, 256, 0, st>>>((const uint8_t*)buf, mask, iters, sink);
    } else {
        if (bytes_per_access == 8) k_l2_red<8><<<grid, 256, 0, st>>>((uint8_t*)buf, mask, iters);
        else k_l2_red<16><<<grid, 256, 0, st>>>((uint8_t*)buf, mask, iters);
    }
    ACN_CHECK_LAUNCH();
    return ACN_OK;
}
