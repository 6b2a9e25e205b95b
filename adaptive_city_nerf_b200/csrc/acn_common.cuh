// Shared helpers for the libacn_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/acn_b200.h"

struct acn_ctx {
    int device;
    int sm_count;
    int cc_major, cc_minor;
    int64_t l2_bytes;
    int max_smem_optin;
    unsigned int* scratch;      // ACN_SCRATCH_WORDS device words (loss-scale reductions), ring-allocated
    unsigned int scratch_next;  // ring cursor (atomic)
};
constexpr int ACN_SCRATCH_WORDS = 256;

// One device word of context scratch; consecutive calls (possibly on different streams / host threads)
// get different words, so they never share one while in flight.
unsigned int* acn_scratch_word(acn_ctx* ctx);

void acn_set_error(const char* fmt, ...);

#define ACN_REQUIRE(cond, code, ...)                 \
    do {                                             \
        if (!(cond)) {                               \
            acn_set_error(__VA_ARGS__);              \
            return (code);                           \
        }                                            \
    } while (0)

// A context belongs to one device; a call made while another device is current would launch on a stream of that other
// device.  Refuse it instead (the Python shim switches devices before it calls).
#define ACN_CHECK_CTX(ctx)                                                                                              \
    do {                                                                                                                \
        ACN_REQUIRE((ctx) != nullptr, ACN_EINVAL, "%s: null context", __func__);                                        \
        int cur__ = -1;                                                                                                 \
        ACN_REQUIRE(cudaGetDevice(&cur__) == cudaSuccess && cur__ == (ctx)->device, ACN_EINVAL,                         \
                    "%s: the context belongs to device %d but device %d is current", __func__, (ctx)->device, cur__);   \
    } while (0)

// Launch-error check without a device sync (SURVEY 8b "Error convention").
#define ACN_CHECK_LAUNCH()                                                         \
    do {                                                                           \
        cudaError_t e__ = cudaGetLastError();                                      \
        if (e__ != cudaSuccess) {                                                  \
            acn_set_error("%s: CUDA error: %s", __func__, cudaGetErrorString(e__)); \
            return ACN_ECUDA;                                                      \
        }                                                                          \
    } while (0)

#define ACN_CUDA(call)                                                              \
    do {                                                                            \
        cudaError_t e__ = (call);                                                   \
        if (e__ != cudaSuccess) {                                                   \
            acn_set_error("%s: %s failed: %s", __func__, #call, cudaGetErrorString(e__)); \
            return ACN_ECUDA;                                                       \
        }                                                                           \
    } while (0)

static inline int acn_grid_1d(int64_t n, int block, int64_t cap = (int64_t)1 << 30) {
    int64_t g = (n + block - 1) / block;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (int)g;
}

// Streaming (read-once / write-once) global accesses: keep them out of L1 so the hash table
// and the MLP weights own the cache.
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream_f4(float4* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// torch.clamp semantics: NaN propagates (fminf / fmaxf alone would replace it by a bound)
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return x != x ? x : fminf(fmaxf(x, lo), hi); }
