// Stage 3 backward on tcgen05 tensor cores (the autocast path's gradient kernel).
//
// One 128-thread CTA per SM walks 128-point tiles.  Per tile it
//   1. recomputes the forward (6 layer MMAs) keeping every activation tile in shared memory,
//   2. walks the layers in reverse; each step issues, back to back from one thread,
//        wgrad  dW_l  += G_l^T X_l      A = G_l tile viewed MN-major, B = X_l tile viewed MN-major,
//                                       contraction over the tile's 128 points (8 K-steps),
//                                       accumulator PERSISTENT in TMEM across all tiles of the CTA
//        bgrad  db_l  += G_l^T 1        same A, B = a constant tile of ones (N = 16)
//        dgrad  D      = G_l W_l        A = G_l K-major, B = the forward weight tile viewed MN-major
//      then one commit; the epilogue applies the ReLU mask and overwrites the (now dead) activation
//      tile with the next gradient tile in place -- no transposes, no extra gradient buffers.
//   3. after the last tile, each thread reads its TMEM lane of the six weight-gradient accumulators
//      and adds them to the global gradients (one atomicAdd per weight per CTA).
// Operands are bf16 (fp32 range: per-sample gradients of a mean loss are ~1e-9 and would flush to
// zero in fp16 without a loss scaler); accumulation is fp32.  tcgen05 requires A and B of one MMA to
// have the same 16-bit type (mixed fp16/bf16 traps with "illegal instruction" on B200), so the
// recomputed activations and the weights are bf16 here as well.
// Layout facts used below were established on hardware with tools/umma_probe.py: for a canonical
// tile with row-group stride RG, the K-major view is (lbo=128, sbo=RG, +256 B per K-step) and the
// MN-major view is (lbo=RG, sbo=128, +2*RG per K-step); an M=128 MMA whose A has only 64 valid
// columns simply produces garbage in TMEM lanes 64..127.
#include <cuda_bf16.h>
#include "field_common.cuh"
#include "field_internal.cuh"
#include "umma.cuh"

namespace {

constexpr int TM = 128;
constexpr uint32_t COL_D = 0, COL_WC2 = 64, COL_WC1 = 128, COL_WC0 = 192, COL_WHD = 224, COL_WT1 = 288, COL_WT0 = 352;
constexpr uint32_t COL_BC2 = 416, COL_BC1 = 432, COL_BC0 = 448, COL_BHD = 464, COL_BT1 = 480, COL_BT0 = 496;
constexpr uint32_t TMEM_COLS = 512;

struct Tile { uint8_t* p; uint32_t rg; };   // canonical tile + its row-group stride (= cols/8 * 128)

__device__ __forceinline__ uint64_t desc_k(const Tile& t, int ks) { return umma::make_desc(umma::smem_u32(t.p) + ks * 256, 128, t.rg); }
__device__ __forceinline__ uint64_t desc_mn(const Tile& t, int ks) { return umma::make_desc(umma::smem_u32(t.p) + ks * 2 * t.rg, t.rg, 128); }

__host__ __device__ constexpr uint32_t idesc_bf16(int N, bool a_mn, bool b_mn) {
    return umma::make_idesc_f16(128, N, a_mn, b_mn) | (1u << 7) | (1u << 10);   // a_format = b_format = BF16
}

__device__ __forceinline__ uint32_t pack_bf2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void st_chunk_bf(const Tile& t, int r, int c, const float* v8) {
    uint4 q = make_uint4(pack_bf2(v8[0], v8[1]), pack_bf2(v8[2], v8[3]), pack_bf2(v8[4], v8[5]), pack_bf2(v8[6], v8[7]));
    *reinterpret_cast<uint4*>(t.p + umma::chunk_off(r, c, t.rg)) = q;
}
// 8 "activation > 0" flags of chunk (r, c) of a post-ReLU bf16 tile
__device__ __forceinline__ uint32_t mask_chunk(const Tile& t, int r, int c) {
    uint4 q = *reinterpret_cast<const uint4*>(t.p + umma::chunk_off(r, c, t.rg));
    uint32_t w[4] = { q.x, q.y, q.z, q.w }, m = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (w[j] & 0x00007fffu) m |= 1u << (2 * j);
        if (w[j] & 0x7fff0000u) m |= 1u << (2 * j + 1);
    }
    return m;
}

struct BwdSmem {
    Tile xe, h1, h2, cin, c1, c2, drr, ghd, ones;
    Tile w_t0, w_t1, w_hd, w_c0, w_c1, w_c2;
    float *b_t0, *b_t1, *b_hd, *b_c0, *b_c1, *b_c2;
    uint64_t* bar;
    uint32_t* tmem_ptr;
};

__host__ __device__ inline size_t bwd_carve(int E, uint8_t* base, BwdSmem* s) {
    size_t off = 0;
    auto take = [&](size_t bytes) { uint8_t* p = base ? base + off : nullptr; off += (bytes + 127) & ~(size_t)127; return p; };
    auto tile = [&](int rows, int cols) { Tile t; t.p = take((size_t)rows * cols * 2); t.rg = (uint32_t)(cols / 8) * 128u; return t; };
    BwdSmem b;
    // gradient / activation tiles first: an M=128 MN-major read of a 64-column tile runs past its end
    // into whatever follows, which must be mapped shared memory (the values land in ignored lanes)
    b.drr = tile(TM, 16); b.ghd = tile(TM, 16);
    b.c2 = tile(TM, 64); b.c1 = tile(TM, 64); b.h2 = tile(TM, 64); b.h1 = tile(TM, 64);
    b.xe = tile(TM, E); b.cin = tile(TM, 32); b.ones = tile(TM, 16);
    b.w_t0 = tile(64, E); b.w_t1 = tile(64, 64); b.w_hd = tile(16, 64);
    b.w_c0 = tile(64, 32); b.w_c1 = tile(64, 64); b.w_c2 = tile(16, 64);
    b.b_t0 = (float*)take(64 * 4); b.b_t1 = (float*)take(64 * 4); b.b_hd = (float*)take(16 * 4);
    b.b_c0 = (float*)take(64 * 4); b.b_c1 = (float*)take(64 * 4); b.b_c2 = (float*)take(16 * 4);
    b.bar = (uint64_t*)take(8); b.tmem_ptr = (uint32_t*)take(4);
    take(4096);   // slack for the over-reads described above
    if (s) *s = b;
    return off;
}

template <typename RowSrc>
__device__ void stage_weight_bf(const Tile& t, int rows, int K, int k_src, RowSrc src_row) {
    for (int idx = threadIdx.x; idx < rows * (K / 8); idx += blockDim.x) {
        int n = idx / (K / 8), c = idx - n * (K / 8);
        const float* src = src_row(n);
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { int k = c * 8 + j; v[j] = (src && k < k_src) ? __ldg(src + k) : 0.0f; }
        st_chunk_bf(t, n, c, v);
    }
}

// barrier protocol around one group of MMAs issued by thread 0 (see field_tc.cu::run_layer)
template <typename Issue>
__device__ __forceinline__ void mma_phase(uint64_t* bar, uint32_t& phase, Issue issue) {
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    if (threadIdx.x == 0) {
        umma::fence_after_sync();
        issue();
        umma::commit(bar);
    }
    umma::mbar_wait(bar, phase);
    phase ^= 1u;
    umma::fence_after_sync();
}

__device__ __forceinline__ void fwd_mma(uint32_t tmem, const Tile& a, const Tile& w, int N, int K) {
    const uint32_t id = idesc_bf16(N, false, false);
    for (int ks = 0; ks < K / 16; ++ks) umma::mma_f16_ss(tmem + COL_D, desc_k(a, ks), desc_k(w, ks), id, ks > 0);
}
// dW (+)= G^T X over the tile's 128 points; db (+)= G^T 1
__device__ __forceinline__ void wgrad_mma(uint32_t tmem, uint32_t col_w, uint32_t col_b, const Tile& g, const Tile& x, int N,
                                          const Tile& ones, bool first) {
    const uint32_t idw = idesc_bf16(N, true, true), idb = idesc_bf16(16, true, true);
    for (int ks = 0; ks < TM / 16; ++ks) umma::mma_f16_ss(tmem + col_w, desc_mn(g, ks), desc_mn(x, ks), idw, !(first && ks == 0));
    for (int ks = 0; ks < TM / 16; ++ks) umma::mma_f16_ss(tmem + col_b, desc_mn(g, ks), desc_mn(ones, ks), idb, !(first && ks == 0));
}
// D = G W  (G: 128 x Kout K-major; W tile: Kout rows x N cols, viewed MN-major)
__device__ __forceinline__ void dgrad_mma(uint32_t tmem, const Tile& g, const Tile& w, int N, int Kout) {
    const uint32_t id = idesc_bf16(N, false, true);
    for (int ks = 0; ks < Kout / 16; ++ks) umma::mma_f16_ss(tmem + COL_D, desc_k(g, ks), desc_mn(w, ks), id, ks > 0);
}

// forward epilogue: 64 columns -> bias + ReLU -> bf16 row of `dst`
__device__ __forceinline__ void relu_epilogue(uint32_t tmem_row, const float* bias, const Tile& dst, int row) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float v[16];
        umma::ld16(tmem_row + COL_D + q * 16, v);
        umma::wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j] + bias[q * 16 + j], 0.0f);
        st_chunk_bf(dst, row, 2 * q, v);
        st_chunk_bf(dst, row, 2 * q + 1, v + 8);
    }
}
// backward epilogue: D (64 cols) * [act > 0] -> bf16 row written over the activation tile itself
__device__ __forceinline__ void mask_epilogue(uint32_t tmem_row, const Tile& act, int row) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float v[16];
        umma::ld16(tmem_row + COL_D + q * 16, v);
        umma::wait_ld();
        uint32_t m0 = mask_chunk(act, row, 2 * q), m1 = mask_chunk(act, row, 2 * q + 1);
#pragma unroll
        for (int j = 0; j < 8; ++j) { v[j] = (m0 >> j) & 1u ? v[j] : 0.0f; v[8 + j] = (m1 >> j) & 1u ? v[8 + j] : 0.0f; }
        st_chunk_bf(act, row, 2 * q, v);
        st_chunk_bf(act, row, 2 * q + 1, v + 8);
    }
}

template <typename EncT>
__global__ void __launch_bounds__(TM, 1) k_field_bwd_tc(
    const EncT* __restrict__ enc, const float* __restrict__ dirs, int dstride, int dgroup, int64_t P, int E, int G,
    acn_field_weights w, const float4* __restrict__ d_rgb_sigma, acn_field_grads g, float* __restrict__ d_enc)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    BwdSmem s;
    bwd_carve(E, smem_raw, &s);
    const int tid = threadIdx.x, warp = tid >> 5;

    stage_weight_bf(s.w_t0, 64, E, E, [&](int n) { return w.p[0] + (size_t)n * E; });
    stage_weight_bf(s.w_t1, 64, 64, 64, [&](int n) { return w.p[2] + (size_t)n * 64; });
    stage_weight_bf(s.w_hd, 16, 64, 64, [&](int n) { return n < G ? w.p[6] + (size_t)n * 64 : (n == 15 ? w.p[4] : (const float*)nullptr); });
    stage_weight_bf(s.w_c0, 64, 32, G + 16, [&](int n) { return w.p[8] + (size_t)n * (G + 16); });
    stage_weight_bf(s.w_c1, 64, 64, 64, [&](int n) { return w.p[10] + (size_t)n * 64; });
    stage_weight_bf(s.w_c2, 16, 64, 64, [&](int n) { return n < 3 ? w.p[12] + (size_t)n * 64 : (const float*)nullptr; });
    if (tid < 64) {
        s.b_t0[tid] = __ldg(w.p[1] + tid); s.b_t1[tid] = __ldg(w.p[3] + tid);
        s.b_c0[tid] = __ldg(w.p[9] + tid); s.b_c1[tid] = __ldg(w.p[11] + tid);
    }
    if (tid < 16) {
        s.b_hd[tid] = tid < G ? __ldg(w.p[7] + tid) : (tid == 15 ? __ldg(w.p[5]) : 0.0f);
        s.b_c2[tid] = tid < 3 ? __ldg(w.p[13] + tid) : 0.0f;
    }
    {   // constant ones tile (row = this thread), and zero the slack
        float one[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) one[j] = 1.0f;
        st_chunk_bf(s.ones, tid, 0, one);
        st_chunk_bf(s.ones, tid, 1, one);
    }
    if (tid == 0) { umma::mbar_init(s.bar, 1); umma::fence_mbar_init(); }
    if (warp == 0) umma::tmem_alloc(s.tmem_ptr, TMEM_COLS);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = *s.tmem_ptr;
    const uint32_t tmem_row = tmem + ((uint32_t)(warp * 32) << 16);
    uint32_t phase = 0;
    bool first = true;
    const bool want_denc = d_enc != nullptr;

    const int64_t ntiles = (P + TM - 1) / TM;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t p = tile * TM + tid;
        const bool on = p < P;
        // ---------------- forward recompute ----------------
        for (int c = 0; c < E / 8; ++c) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = 0.0f;
            if (on) {
                if constexpr (sizeof(EncT) == 2) {
                    uint4 q = __ldg(reinterpret_cast<const uint4*>(enc + p * E) + c);
                    const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
                    for (int j = 0; j < 4; ++j) { float2 f = __half22float2(h[j]); v[2 * j] = f.x; v[2 * j + 1] = f.y; }
                } else {
                    const float4* src = reinterpret_cast<const float4*>(enc + p * E + c * 8);
                    float4 a = __ldg(src), b = __ldg(src + 1);
                    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
                }
            }
            st_chunk_bf(s.xe, tid, c, v);
        }
        mma_phase(s.bar, phase, [&] { fwd_mma(tmem, s.xe, s.w_t0, 64, E); });
        relu_epilogue(tmem_row, s.b_t0, s.h1, tid);
        mma_phase(s.bar, phase, [&] { fwd_mma(tmem, s.h1, s.w_t1, 64, 64); });
        relu_epilogue(tmem_row, s.b_t1, s.h2, tid);
        mma_phase(s.bar, phase, [&] { fwd_mma(tmem, s.h2, s.w_hd, 16, 64); });
        float sig_raw;
        {
            float v[16], cin[32];
            umma::ld16(tmem_row + COL_D, v);
            umma::wait_ld();
            sig_raw = v[15] + s.b_hd[15];
#pragma unroll
            for (int j = 0; j < 15; ++j) cin[j] = j < G ? v[j] + s.b_hd[j] : 0.0f;
#pragma unroll
            for (int j = 15; j < 32; ++j) cin[j] = 0.0f;
#pragma unroll
            for (int c = 0; c < 4; ++c) st_chunk_bf(s.cin, tid, c, cin + 8 * c);
            float sh[16];
            if (on) {
                const float* dp = dir_of(dirs, dstride, dgroup, p);
                sh16_expert(__ldg(dp), __ldg(dp + 1), __ldg(dp + 2), sh);
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) sh[j] = 0.0f;
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                int k = G + j;
                *reinterpret_cast<__nv_bfloat16*>(s.cin.p + umma::chunk_off(tid, k >> 3, s.cin.rg) + (k & 7) * 2) = __float2bfloat16_rn(sh[j]);
            }
        }
        mma_phase(s.bar, phase, [&] { fwd_mma(tmem, s.cin, s.w_c0, 64, 32); });
        relu_epilogue(tmem_row, s.b_c0, s.c1, tid);
        mma_phase(s.bar, phase, [&] { fwd_mma(tmem, s.c1, s.w_c1, 64, 64); });
        relu_epilogue(tmem_row, s.b_c1, s.c2, tid);
        mma_phase(s.bar, phase, [&] { fwd_mma(tmem, s.c2, s.w_c2, 16, 64); });
        float d_sig;
        {   // output gradients: d rgb_raw = dy * y (1 - y); d sigma_raw = dy * exp(clamp(sigma_raw))
            float v[16], drr[16];
            umma::ld16(tmem_row + COL_D, v);
            umma::wait_ld();
            float4 dy = on ? __ldg(d_rgb_sigma + p) : make_float4(0.f, 0.f, 0.f, 0.f);
            float y0 = sigmoid_f(v[0] + s.b_c2[0]), y1 = sigmoid_f(v[1] + s.b_c2[1]), y2 = sigmoid_f(v[2] + s.b_c2[2]);
#pragma unroll
            for (int j = 0; j < 16; ++j) drr[j] = 0.0f;
            drr[0] = dy.x * y0 * (1.0f - y0); drr[1] = dy.y * y1 * (1.0f - y1); drr[2] = dy.z * y2 * (1.0f - y2);
            d_sig = dy.w * trunc_exp_f(sig_raw);
            st_chunk_bf(s.drr, tid, 0, drr);
            st_chunk_bf(s.drr, tid, 1, drr + 8);
        }
        // ---------------- backward ----------------
        // colour head: dW_c2 = drr^T C2, g1 = (drr W_c2) * [c2 > 0]  -> over C2
        mma_phase(s.bar, phase, [&] {
            wgrad_mma(tmem, COL_WC2, COL_BC2, s.drr, s.c2, 64, s.ones, first);
            dgrad_mma(tmem, s.drr, s.w_c2, 64, 16);
        });
        mask_epilogue(tmem_row, s.c2, tid);
        // colour hidden 2
        mma_phase(s.bar, phase, [&] {
            wgrad_mma(tmem, COL_WC1, COL_BC1, s.c2, s.c1, 64, s.ones, first);
            dgrad_mma(tmem, s.c2, s.w_c1, 64, 64);
        });
        mask_epilogue(tmem_row, s.c1, tid);
        // colour hidden 1: input [geo, sh]; d cin (32 cols) -> heads gradient tile [d geo | 0 | d sigma_raw @15]
        mma_phase(s.bar, phase, [&] {
            wgrad_mma(tmem, COL_WC0, COL_BC0, s.c1, s.cin, 32, s.ones, first);
            dgrad_mma(tmem, s.c1, s.w_c0, 32, 64);
        });
        {
            float v[16];
            umma::ld16(tmem_row + COL_D, v);
            umma::wait_ld();
#pragma unroll
            for (int j = 0; j < 15; ++j) v[j] = j < G ? v[j] : 0.0f;
            v[15] = d_sig;
            st_chunk_bf(s.ghd, tid, 0, v);
            st_chunk_bf(s.ghd, tid, 1, v + 8);
        }
        // heads
        mma_phase(s.bar, phase, [&] {
            wgrad_mma(tmem, COL_WHD, COL_BHD, s.ghd, s.h2, 64, s.ones, first);
            dgrad_mma(tmem, s.ghd, s.w_hd, 64, 16);
        });
        mask_epilogue(tmem_row, s.h2, tid);
        // trunk 2
        mma_phase(s.bar, phase, [&] {
            wgrad_mma(tmem, COL_WT1, COL_BT1, s.h2, s.h1, 64, s.ones, first);
            dgrad_mma(tmem, s.h2, s.w_t1, 64, 64);
        });
        mask_epilogue(tmem_row, s.h1, tid);
        // trunk 1 (+ optional d enc)
        mma_phase(s.bar, phase, [&] {
            wgrad_mma(tmem, COL_WT0, COL_BT0, s.h1, s.xe, E, s.ones, first);
            if (want_denc) dgrad_mma(tmem, s.h1, s.w_t0, E, 64);
        });
        if (want_denc) {
            for (int q = 0; q < E / 16; ++q) {
                float v[16];
                umma::ld16(tmem_row + COL_D + q * 16, v);
                umma::wait_ld();
                if (on) {
                    float4* dst = reinterpret_cast<float4*>(d_enc + p * E + q * 16);
#pragma unroll
                    for (int j = 0; j < 4; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
            }
        }
        first = false;
    }

    // ---------------- flush the TMEM-resident weight gradients ----------------
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    if (!first) {
        // thread t owns TMEM lane t = row t of every accumulator; `dst` is that row's destination (or null)
        auto flush = [&](uint32_t col, int ncols, float* dst, int nvalid) {
            for (int q = 0; q < ncols / 16; ++q) {
                float v[16];
                umma::ld16(tmem_row + col + q * 16, v);     // warp-collective: every lane loads
                umma::wait_ld();
                if (dst) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        int n = q * 16 + j;
                        if (n < nvalid && v[j] != 0.0f) atomicAdd(dst + n, v[j]);
                    }
                }
            }
        };
        auto row = [&](float* base, bool valid, int ld) { return (base && valid) ? base + (size_t)tid * ld : (float*)nullptr; };
        flush(COL_WC2, 64, row(g.p[12], tid < 3, 64), 64);
        flush(COL_WC1, 64, row(g.p[10], tid < 64, 64), 64);
        flush(COL_WC0, 32, row(g.p[8], tid < 64, G + 16), G + 16);
        flush(COL_WHD, 64, tid == 15 ? g.p[4] : row(g.p[6], tid < G, 64), 64);
        flush(COL_WT1, 64, row(g.p[2], tid < 64, 64), 64);
        flush(COL_WT0, E, row(g.p[0], tid < 64, E), E);
        // bias gradients: every column of the G^T 1 accumulators holds the same sum; take column 0
        flush(COL_BC2, 16, row(g.p[13], tid < 3, 1), 1);
        flush(COL_BC1, 16, row(g.p[11], tid < 64, 1), 1);
        flush(COL_BC0, 16, row(g.p[9], tid < 64, 1), 1);
        flush(COL_BHD, 16, tid == 15 ? g.p[5] : row(g.p[7], tid < G, 1), 1);
        flush(COL_BT1, 16, row(g.p[3], tid < 64, 1), 1);
        flush(COL_BT0, 16, row(g.p[1], tid < 64, 1), 1);
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, TMEM_COLS);
}

}  // namespace

int acn_field_bwd_tc(acn_ctx* ctx, const void* enc, int enc_dtype, const float* dirs, int dirs_stride, int dirs_group,
                     int64_t P, int E, int H, int G, int C, const acn_field_weights* w, const float* d_rgb_sigma,
                     const acn_field_grads* g, float* d_enc, cudaStream_t st) {
    ACN_REQUIRE(H == 64 && C == 64, ACN_EUNSUPPORTED, "acn_field_bwd(f16): hidden widths must be 64 (got H=%d, C=%d)", H, C);
    ACN_REQUIRE(E == 16 || E == 32 || E == 48 || E == 64, ACN_EUNSUPPORTED, "acn_field_bwd(f16): encoding width %d not in {16,32,48,64}", E);
    ACN_REQUIRE(G >= 1 && G <= 15, ACN_EUNSUPPORTED, "acn_field_bwd(f16): geo_feat_dim %d outside [1,15]", G);
    ACN_REQUIRE(((uintptr_t)enc & 15) == 0 && ((uintptr_t)d_enc & 15) == 0, ACN_EINVAL, "acn_field_bwd(f16): enc / d_enc must be 16-byte aligned");
    const size_t smem = bwd_carve(E, nullptr, nullptr);
    ACN_REQUIRE((int)smem <= ctx->max_smem_optin, ACN_EUNSUPPORTED, "acn_field_bwd(f16): needs %zu B shared memory", smem);
    const int64_t ntiles = (P + TM - 1) / TM;
    int64_t grid = ctx->sm_count;       // one CTA per SM: the persistent accumulators take all 512 TMEM columns
    if (grid > ntiles) grid = ntiles;
    if (enc_dtype == ACN_F16) {
        ACN_CUDA(cudaFuncSetAttribute(k_field_bwd_tc<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_field_bwd_tc<__half><<<(int)grid, TM, smem, st>>>((const __half*)enc, dirs, dirs_stride, dirs_group, P, E, G, *w,
                                                           (const float4*)d_rgb_sigma, *g, d_enc);
    } else {
        ACN_CUDA(cudaFuncSetAttribute(k_field_bwd_tc<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_field_bwd_tc<float><<<(int)grid, TM, smem, st>>>((const float*)enc, dirs, dirs_stride, dirs_group, P, E, G, *w,
                                                          (const float4*)d_rgb_sigma, *g, d_enc);
    }
    ACN_CHECK_LAUNCH();
    return ACN_OK;
}
