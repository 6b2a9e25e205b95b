/*
 * acn_b200_debug.h -- hardware probes and kernel timelines, built into libacn_b200_debug.so (NOT into the product
 * library libacn_b200.so).  The debug library is a superset build of the product sources (-DACN_DEBUG_BUILD) plus
 * csrc/debug/ (tcgen05 layout / rate probes, the L2 gather / atomic peak micro-benchmarks); it keeps its own contexts.
 * Used by tools/ and by the descriptor self-tests only.
 */
#ifndef ACN_B200_DEBUG_H
#define ACN_B200_DEBUG_H
#include "acn_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

/* One tcgen05 tile GEMM  D(128,N) = A(128,K) * W(N,K)^T  (fp16 in, fp32 out); validates the
 * shared-memory / instruction descriptors the fused MLP kernels are built on.  N in
 * {16,32,64}, K in {16,32,64}. */
int acn_debug_umma_gemm(acn_ctx*, const void* a_f16, const void* w_f16, int N, int K, float* d,
                        acn_stream);

/* Same product with A read from TENSOR MEMORY (written there by tcgen05.st): validates the TMEM operand layout of
 * the forward MLP kernel's activation chain. */
int acn_debug_umma_gemm_ts(acn_ctx*, const void* a_f16, const void* w_f16, int N, int K, float* d, acn_stream);

/* Dispatch-rate probe: `issuers` threads each issue `nmma` back-to-back M x N x 16 MMAs (mode 0: operands in shared
 * memory, 1: A in tensor memory), commit and wait, `reps` times, on every SM; out4[i] = average SM cycles per round of
 * issuer i on CTA 0 (tools/umma_rate.py). */
int acn_debug_umma_rate(acn_ctx*, int mode, int M, int N, int nmma, int reps, int issuers, long long* out4, acn_stream);

/* Raw harness: stages two 16-bit matrices as canonical tiles and issues `ksteps` MMAs with
 * host-supplied descriptors, then dumps TMEM lanes 0..127 x ncols.  Used by tools/umma_probe.py to
 * establish MN-major / mixed-dtype / M=64 layouts on hardware. */
int acn_debug_umma_raw(acn_ctx*, const void* a16, int rows_a, int cols_a, const void* b16, int rows_b,
                       int cols_b, uint32_t idesc, uint32_t a_lbo, uint32_t a_sbo, uint32_t a_step,
                       uint32_t b_lbo, uint32_t b_sbo, uint32_t b_step, int ksteps, int ncols, float* out,
                       acn_stream);

/* Timeline of the fused forward MLP kernel: when `trace` (device, 1024 int64) is non-NULL, CTA 0 of the
 * following acn_field_fwd(ACN_F16) launches logs (SM clock << 8 | tag) pairs from row 0 of its first
 * warpgroup (tools/field_trace.py decodes them).  NULL switches it off. */
int acn_debug_field_trace(acn_ctx*, long long* trace_or_null);


/* Memory-system ceilings for the hash-grid kernels (tools/l2_peak.py): every thread of a full-occupancy grid issues
 * `iters` x 8 independent accesses of `bytes_per_access` (8 or 16) at pseudo-random, aligned offsets inside a buffer of
 * buf_bytes (a power of two; 64 MiB = the T = 2^19 table, L2-resident on B200).
 *   mode 0: ld.global (gathers; sums land in sink so nothing is optimised away)
 *   mode 1: red.global.add.f32 vectors (v2 for 8 bytes, v4 for 16)
 * The caller times the launch with CUDA events; accesses per launch = threads * iters * 8, threads = grid * 256. */
int acn_debug_l2_probe(acn_ctx*, int mode, int bytes_per_access, void* buf, int64_t buf_bytes, int iters, int grid,
                       float* sink, acn_stream);

/* The single-role fused backward (every MLP thread scatters its own d enc columns between its epilogues), superseded in
 * the product by the warp-specialised acn_render_expert_bwd: same arguments, same results to fp32 summation order.  The
 * A/B partner of tools/prof_fused_bwd.py and the cross-check of tests/test_gpu_tc.py. */
int acn_debug_render_expert_bwd_single(acn_ctx*, const float* x_or_null, int x_stride, const float* rays8_or_null,
                                       const float* t_vals_or_null, int64_t P, int S, const int32_t* range_or_null,
                                       const float* box6_or_null, int L, int F,
                                       int log2T, const int32_t* res, int interp, const void* enc_f16, const float* dirs,
                                       int dirs_stride, int dirs_group, int H, int G, int C, const acn_field_weights* w,
                                       const float* d_rgb_sigma, const acn_field_grads* g, float* dtable, acn_stream);

/* 16-byte scattered REDs from `grid` CTAs of `block` threads in which only `active_lanes` lanes of every warp issue
 * them, with `work` dependent integer operations between two REDs of a thread (tools/red_probe.py): the RED rate an SM
 * reaches with few resident warps and partly filled RED instructions.  REDs per launch = grid * block / 32 * active_lanes
 * * iters * 8. */
int acn_debug_red_probe(acn_ctx*, void* buf, int64_t buf_bytes, int iters, int grid, int block, int active_lanes, int work,
                        acn_stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* ACN_B200_DEBUG_H */
