// Diagnostics: a raw tcgen05 harness whose descriptors come straight from the host, used to pin
// down (on real hardware) the operand layouts the backward kernels rely on: MN-major operands,
// mixed fp16/bf16 operands and the M = 64 accumulator layout.  Not on any product path.
#include "acn_common.cuh"
#include "umma.cuh"
#include "../../../include/acn_b200_debug.h"

namespace {

// Stage a (rows, cols) row-major 16-bit matrix as a canonical tile: (r, c) ->
// (r>>3)*rg + (c>>3)*128 + (r&7)*16 + (c&7)*2 with rg = (cols/8)*128.
__device__ void stage16(uint8_t* tile, const uint16_t* src, int rows, int cols) {
    const uint32_t rg = (uint32_t)(cols / 8) * 128u;
    for (int idx = threadIdx.x; idx < rows * (cols / 8); idx += blockDim.x) {
        int r = idx / (cols / 8), c = idx - r * (cols / 8);
        *reinterpret_cast<uint4*>(tile + umma::chunk_off(r, c, rg)) = *reinterpret_cast<const uint4*>(src + (size_t)r * cols + c * 8);
    }
}

__global__ void __launch_bounds__(128) k_umma_raw(
    const uint16_t* __restrict__ a, int rows_a, int cols_a, const uint16_t* __restrict__ b, int rows_b, int cols_b,
    uint32_t idesc, uint32_t a_lbo, uint32_t a_sbo, uint32_t a_step, uint32_t b_lbo, uint32_t b_sbo, uint32_t b_step,
    int ksteps, int ncols, float* __restrict__ out)
{
    extern __shared__ __align__(128) uint8_t sm[];
    uint8_t* a_tile = sm;
    uint8_t* b_tile = sm + 40 * 1024;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 80 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0u;
    __syncthreads();
    stage16(a_tile, a, rows_a, cols_a);
    stage16(b_tile, b, rows_b, cols_b);
    if (tid == 0) { umma::mbar_init(&bar, 1); umma::fence_mbar_init(); }
    if (warp == 0) umma::tmem_alloc(&tmem_ptr, 256);
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem_base = tmem_ptr;
    const uint32_t tmem_row = tmem_base + ((uint32_t)(warp * 32) << 16);
    // zero the accumulator window first (via an MMA-free path: tcgen05.st is not wrapped; use accumulate=false on k=0)
    if (tid == 0) {
        const uint32_t a_addr = umma::smem_u32(a_tile), b_addr = umma::smem_u32(b_tile);
        for (int ks = 0; ks < ksteps; ++ks)
            umma::mma_f16_ss(tmem_base, umma::make_desc(a_addr + ks * a_step, a_lbo, a_sbo),
                             umma::make_desc(b_addr + ks * b_step, b_lbo, b_sbo), idesc, ks > 0);
        umma::commit(&bar);
    }
    umma::mbar_wait(&bar, 0);
    umma::fence_after_sync();
    for (int q = 0; q < ncols / 16; ++q) {
        float v[16];
        umma::ld16(tmem_row + q * 16, v);
        umma::wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) out[(size_t)tid * ncols + q * 16 + j] = v[j];
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem_base, 256);
}

}  // namespace

extern "C" int acn_debug_umma_raw(acn_ctx* ctx, const void* a16, int rows_a, int cols_a, const void* b16, int rows_b,
                                  int cols_b, uint32_t idesc, uint32_t a_lbo, uint32_t a_sbo, uint32_t a_step,
                                  uint32_t b_lbo, uint32_t b_sbo, uint32_t b_step, int ksteps, int ncols, float* out,
                                  acn_stream stream) {
    ACN_CHECK_CTX(ctx);
    ACN_REQUIRE(a16 && b16 && out, ACN_EINVAL, "acn_debug_umma_raw: null buffer");
    ACN_REQUIRE(rows_a > 0 && rows_a * cols_a * 2 <= 40 * 1024 && rows_b > 0 && rows_b * cols_b * 2 <= 40 * 1024 &&
                cols_a % 8 == 0 && cols_b % 8 == 0, ACN_EUNSUPPORTED, "acn_debug_umma_raw: operand too large / misaligned");
    ACN_REQUIRE(ncols % 16 == 0 && ncols >= 16 && ncols <= 256 && ksteps >= 1, ACN_EINVAL, "acn_debug_umma_raw: bad ncols/ksteps");
    const int smem = 80 * 1024;
    ACN_CUDA(cudaFuncSetAttribute(k_umma_raw, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k_umma_raw<<<1, 128, smem, (cudaStream_t)stream>>>((const uint16_t*)a16, rows_a, cols_a, (const uint16_t*)b16, rows_b,
                                                       cols_b, idesc, a_lbo, a_sbo, a_step, b_lbo, b_sbo, b_step, ksteps,
                                                       ncols, out);
    ACN_CHECK_LAUNCH();
    return ACN_OK;
}
