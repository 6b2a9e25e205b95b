// Building blocks shared by the tcgen05 field-MLP kernels (field_mma.cu: acn_field_fwd / acn_field_bwd /
// acn_render_expert_bwd; expert_fwd.cu: acn_render_expert_fwd): tile descriptors, the shared-memory map of the weights,
// MMA issue helpers, TMEM epilogue pieces.  Everything sits in an anonymous namespace (one copy per translation unit).
// Layout facts and the reasons behind them are in the header comment of field_mma.cu.
#pragma once
#include "field_common.cuh"
#include "field_internal.cuh"
#include "hashgrid.cuh"
#include "umma.cuh"
#ifdef ACN_DEBUG_BUILD
#include "../../include/acn_b200_debug.h"
#endif

namespace {

using umma::lds128;
using umma::sts128;

constexpr int TM = 128;                 // points per tile = MMA M

// A canonical tile: shared-space byte address >> 4 (all of shared memory fits the descriptor's 14-bit field) and
// row-group stride (= cols/8 * 128 B).  Descriptors are built as `a16 + immediate` -- ONE uniform add per operand.
// This matters: with umma::make_desc's mask-and-shift per MMA the issuing thread spent ~90 cycles per tcgen05.mma
// on the uniform datapath and the kernels were bound by MMA ISSUE (26 / 140 MMAs per tile), not by the tensor pipe.
struct Tile { uint32_t a16; uint32_t rg; };
__device__ __forceinline__ Tile mk_tile(uint32_t addr, int cols) { return Tile{ addr >> 4, (uint32_t)(cols / 8) * 128u }; }

// smem descriptor (cute::UMMA::SmemDescriptor): [0,14) addr>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version = 1
__device__ __forceinline__ uint64_t desc_k(const Tile& t, int ks) {          // K-major: LBO = 128, SBO = RG, +256 B per K step
    const uint32_t lo = t.a16 + (uint32_t)ks * 16u + (8u << 16);
    const uint32_t hi = (t.rg >> 4) | (1u << 14);
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint64_t desc_mn(const Tile& t, int ks) {         // MN-major: LBO = RG, SBO = 128, +2*RG per K step
    const uint32_t lo = t.a16 + (uint32_t)ks * (t.rg >> 3) + ((t.rg >> 4) << 16);
    const uint32_t hi = 8u | (1u << 14);
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint32_t chunk_addr(const Tile& t, int r, int c) { return (t.a16 << 4) + umma::chunk_off(r, c, t.rg); }

// ---- small PTX helpers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
// two fp32 -> packed halves with ReLU, one instruction
__device__ __forceinline__ uint32_t pack_relu_h2(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
// g * [act > 0] on packed halves
__device__ __forceinline__ uint32_t hmask2(uint32_t g, uint32_t act) {
    uint32_t m, d;
    asm("set.gt.f16x2.f16x2 %0, %1, %2;" : "=r"(m) : "r"(act), "r"(0u));
    asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(g), "r"(m));
    return d;
}
__device__ __forceinline__ float fast_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_rsqrt(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// MUFU-based activations: relative error ~2^-22, far inside the fp16 operand rounding of this path
__device__ __forceinline__ float sigmoid_fast(float x) { return fast_rcp(1.0f + fast_ex2(-1.4426950408889634f * x)); }
__device__ __forceinline__ float trunc_exp_fast(float x) { return fast_ex2(1.4426950408889634f * fminf(fmaxf(x, -88.722839111f), 88.722839111f)); }

// models/inr/meta_ngp.py:166-169 + models/encodings.py:141 (d / max(|d|, 1e-9), twice), then the SH basis.
// The results feed fp16 GEMM operands, so the two normalisations use rsqrt.approx instead of sqrt + divide.
__device__ __forceinline__ void sh16_fast(float x, float y, float z, float* sh) {
#pragma unroll
    for (int rep = 0; rep < 2; ++rep) {
        const float inv = fminf(fast_rsqrt(x * x + y * y + z * z), 1e9f);
        x *= inv; y *= inv; z *= inv;
    }
    sh16_poly(x, y, z, sh);
}

// TMEM -> registers, 32 consecutive columns of this thread's lane
__device__ __forceinline__ void ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ uint4 pack8(const float* v) {
    return make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
}

// ---- shared-memory map of the weights (byte offsets from the weight block) ------------------------------
// B-operand tiles, rows = output unit: t0 (64,E) t1 (64,64) hd (16,64) c0 (64,32) c1 (64,64) c2 (16,64).
// Hidden-layer biases are B operands too: (64,16) tiles whose column 0 holds the bias, multiplied by a constant A
// tile `one` (128,16) whose column 0 is 1 -- one extra K=16 MMA per hidden layer adds the bias inside the fp32
// accumulator, so the epilogue is a single cvt.rn.relu.f16x2.f32 per two activations.  Head / output biases stay
// fp32 (64 B each) and are added in the epilogue.
struct WMap { uint32_t t0, t1, hd, c0, c1, c2, bt_t0, bt_t1, bt_c0, bt_c1, one, b_hd, b_c2, end; };
__host__ __device__ constexpr WMap wmap(int E) {
    WMap m{};
    uint32_t o = 0;
    m.t0 = o; o += 64 * E * 2;
    m.t1 = o; o += 64 * 64 * 2;
    m.hd = o; o += 16 * 64 * 2;
    m.c0 = o; o += 64 * 32 * 2;
    m.c1 = o; o += 64 * 64 * 2;
    m.c2 = o; o += 16 * 64 * 2;
    m.bt_t0 = o; o += 2048; m.bt_t1 = o; o += 2048; m.bt_c0 = o; o += 2048; m.bt_c1 = o; o += 2048;
    m.one = o; o += 4096;
    m.b_hd = o; o += 64; m.b_c2 = o; o += 64;
    m.end = o;
    return m;
}

// Colour-MLP input columns are permuted so that the SH block sits at a fixed position whatever G is:
// A-tile column k' < 16 is SH component k', column 16 + j is geo feature j (j < G), the rest is zero.
// The reference order is [geo(G), sh(16)] (models/inr/meta_ngp.py:186), so tile column k' reads source
// column G + k' (k' < 16) or k' - 16.
__device__ __forceinline__ int cin_src_col(int kp, int G) { return kp < 16 ? G + kp : (kp - 16 < G ? kp - 16 : -1); }

template <int E>
__device__ void stage_weights(const acn_field_weights& w, int G, uint32_t wb) {
    constexpr WMap m = wmap(E);
    auto stage = [&](uint32_t addr, int rows, int K, auto elem) {
        const Tile t = mk_tile(addr, K);
        for (int idx = threadIdx.x; idx < rows * (K / 8); idx += blockDim.x) {
            const int n = idx / (K / 8), c = idx - n * (K / 8);
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = elem(n, c * 8 + j);
            sts128(chunk_addr(t, n, c), pack8(v));
        }
    };
    stage(wb + m.t0, 64, E, [&](int n, int k) { return __ldg(w.p[0] + (size_t)n * E + k); });
    stage(wb + m.t1, 64, 64, [&](int n, int k) { return __ldg(w.p[2] + n * 64 + k); });
    stage(wb + m.hd, 16, 64, [&](int n, int k) { return n < G ? __ldg(w.p[6] + n * 64 + k) : (n == 15 ? __ldg(w.p[4] + k) : 0.0f); });
    stage(wb + m.c0, 64, 32, [&](int n, int k) { const int sc = cin_src_col(k, G); return sc >= 0 ? __ldg(w.p[8] + n * (G + 16) + sc) : 0.0f; });
    stage(wb + m.c1, 64, 64, [&](int n, int k) { return __ldg(w.p[10] + n * 64 + k); });
    stage(wb + m.c2, 16, 64, [&](int n, int k) { return n < 3 ? __ldg(w.p[12] + n * 64 + k) : 0.0f; });
    // hidden biases ride in as (hi, lo) fp16 pairs against two ones columns: hi + lo reproduces the fp32 bias to ~2^-22,
    // i.e. the reference's fp32 bias add (metamodule.py:152-154).  Rounding the bias itself to fp16 cost 0.09 dB of
    // training PSNR after 150 steps against the reference's autocast trace (0.04 dB when done to the reference itself,
    // tools/train_amp_diag.py); the pair costs nothing: same K step.
    auto bias_hl = [&](const float* b, int n, int k) {
        if (k > 1) return 0.0f;
        const float v = __ldg(b + n), hi = __half2float(__float2half_rn(v));
        return k == 0 ? hi : v - hi;
    };
    stage(wb + m.bt_t0, 64, 16, [&](int n, int k) { return bias_hl(w.p[1], n, k); });
    stage(wb + m.bt_t1, 64, 16, [&](int n, int k) { return bias_hl(w.p[3], n, k); });
    stage(wb + m.bt_c0, 64, 16, [&](int n, int k) { return bias_hl(w.p[9], n, k); });
    stage(wb + m.bt_c1, 64, 16, [&](int n, int k) { return bias_hl(w.p[11], n, k); });
    stage(wb + m.one, 128, 16, [&](int, int k) { return k < 2 ? 1.0f : 0.0f; });
    for (int i = threadIdx.x; i < 16; i += blockDim.x) {
        umma::sts_f32(wb + m.b_hd + 4 * i, i < G ? __ldg(w.p[7] + i) : (i == 15 ? __ldg(w.p[5]) : 0.0f));
        umma::sts_f32(wb + m.b_c2 + 4 * i, i < 3 ? __ldg(w.p[13] + i) : 0.0f);
    }
}

// ---- MMA issue (one elected lane, warp-uniform operands) ---------------------------------------------------
__device__ __forceinline__ void mma_fwd(uint32_t d, const Tile& a, const Tile& w, int N, int K) {
    const uint32_t id = umma::make_idesc_f16(128, N, false, false);
    for (int ks = 0; ks < K / 16; ++ks) umma::mma_f16_ss(d, desc_k(a, ks), desc_k(w, ks), id, ks > 0);
}
// hidden layer: D = A W^T + 1 b^T (the bias rides in as one more K step: `one` x `bias tile`)
__device__ __forceinline__ void mma_fwd_bias(uint32_t d, const Tile& a, const Tile& w, const Tile& one, const Tile& bias, int K) {
    const uint32_t id = umma::make_idesc_f16(128, 64, false, false);
    for (int ks = 0; ks < K / 16; ++ks) umma::mma_f16_ss(d, desc_k(a, ks), desc_k(w, ks), id, ks > 0);
    umma::mma_f16_ss(d, desc_k(one, 0), desc_k(bias, 0), id, true);
}
// acc += A^T B over the tile's 128 points: A, B canonical tiles whose ROWS are points (both MN-major views).
// M = 64: the A operand has at most 64 feature columns, and an M=64 MMA fetches half the A bytes of an M=128 one --
// these MMAs are bound by shared-memory operand fetch, not by the tensor pipe.  Accumulator row m lands in TMEM
// lane (m >> 4) * 32 + (m & 15) (measured: tools/umma_probe.py H5b).
__device__ __forceinline__ void mma_over_points(uint32_t acc, const Tile& a, const Tile& b, int N, int M = 64) {
    const uint32_t id = umma::make_idesc_f16(M, N, true, true);
#pragma unroll
    for (int ks = 0; ks < TM / 16; ++ks) umma::mma_f16_ss(acc, desc_mn(a, ks), desc_mn(b, ks), id, true);
}
// D = G W: G (128, Kout) K-major, W tile (Kout rows, N cols) viewed MN-major
__device__ __forceinline__ void mma_dgrad(uint32_t d, const Tile& g, const Tile& w, int N, int Kout) {
    const uint32_t id = umma::make_idesc_f16(128, N, false, true);
    for (int ks = 0; ks < Kout / 16; ++ks) umma::mma_f16_ss(d, desc_k(g, ks), desc_mn(w, ks), id, ks > 0);
}

// Every thread of a group: my shared-memory writes are visible to the tensor core, my TMEM reads are ordered
// before whatever the group's issuer launches next; then meet.
__device__ __forceinline__ void group_sync(uint32_t bar_id, uint32_t nthreads) {
    umma::fence_async_smem();
    umma::fence_before_sync();
    umma::bar_sync(bar_id, nthreads);
}
__device__ __forceinline__ void wait_done(uint32_t bar_addr, uint32_t& phase) {
    umma::mbar_wait_a(bar_addr, phase);
    phase ^= 1u;
    umma::fence_after_sync();
}

// ---- epilogue pieces (thread = row) ------------------------------------------------------------------------
// 32 accumulator columns [col0, col0+32) (bias already inside) -> ReLU -> fp16 -> chunks col0/8.. of `dst`
__device__ __forceinline__ void epi_hidden32(uint32_t tmem_d, int col0, const Tile& dst, int row) {
    float v[32];
    ld32(tmem_d + col0, v);
    umma::wait_ld();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        uint4 o;
        o.x = pack_relu_h2(v[8 * q + 0], v[8 * q + 1]);
        o.y = pack_relu_h2(v[8 * q + 2], v[8 * q + 3]);
        o.z = pack_relu_h2(v[8 * q + 4], v[8 * q + 5]);
        o.w = pack_relu_h2(v[8 * q + 6], v[8 * q + 7]);
        sts128(chunk_addr(dst, row, col0 / 8 + q), o);
    }
}
// heads, geo part: accumulator columns 0..G-1 = geo, 15 = raw sigma -> cin chunks 2,3 ([geo(G) | 0]);
// returns raw sigma (bias added)
__device__ __forceinline__ float epi_heads_geo(uint32_t tmem_d, uint32_t b_hd_addr, int G, const Tile& cin, int row, float pad_last = 0.0f) {
    float v[16];
    umma::ld16(tmem_d, v);
    float b[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint4 u = lds128(b_hd_addr + 16 * q);
        b[4 * q] = __uint_as_float(u.x); b[4 * q + 1] = __uint_as_float(u.y); b[4 * q + 2] = __uint_as_float(u.z); b[4 * q + 3] = __uint_as_float(u.w);
    }
    umma::wait_ld();
    const float sig_raw = v[15] + b[15];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = j < G ? v[j] + b[j] : 0.0f;
    v[15] = pad_last;                      // column 31 of the colour input: always zero weight; the backward sets it to 1
    sts128(chunk_addr(cin, row, 2), pack8(v));
    sts128(chunk_addr(cin, row, 3), pack8(v + 8));
    return sig_raw;
}
// heads, direction part: SH(dir) -> cin chunks 0,1
__device__ __forceinline__ void epi_heads_sh(const float* dir3, const Tile& cin, int row) {
    float sh[16];
    sh16_fast(dir3[0], dir3[1], dir3[2], sh);
    sts128(chunk_addr(cin, row, 0), pack8(sh));
    sts128(chunk_addr(cin, row, 1), pack8(sh + 8));
}
__device__ __forceinline__ void load_b_c2(uint32_t addr, float* b3) {
    const uint4 u = lds128(addr);
    b3[0] = __uint_as_float(u.x); b3[1] = __uint_as_float(u.y); b3[2] = __uint_as_float(u.z);
}

// Optional timeline for tools/field_trace.py: row 0 of warpgroup 0 in CTA 0 logs (SM clock << 8 | tag).
template <bool ON> struct Tracer {
    long long* buf; int n, cap;
    __device__ __forceinline__ void operator()(int tag) {
        if constexpr (ON) { if (buf && n < cap) { buf[n++] = (clock64() << 8) | (long long)(tag & 0xff); } }
    }
};
constexpr int TRACE_CAP = 1024;

// ======================================================================================================
// forward: FWG warpgroups x 1 tile, activations chained through TENSOR MEMORY
// ======================================================================================================
// With both operands in shared memory an M=128, N=64, K=16 MMA fetches 4 KB of A and 2 KB of B per 32 tensor-pipe
// cycles -- 192 B/cycle against a 128 B/cycle shared-memory port that the epilogue's stores also use.  The first
// versions of this kernel were bound exactly there (tools/field_trace.py: a layer's MMAs took ~3x their tensor
// time).  So the activations never touch shared memory: the epilogue writes the next layer's A operand back into
// tensor memory (tcgen05.st, two halves per 32-bit column) and the MMA reads it from there
// (tcgen05.mma [d], [a_tmem], b_desc); only the 2 KB weight slice comes from shared memory.
//
// TMEM map, 128 columns per tile = two halves H0 = [0,64), H1 = [64,128).  An accumulator is read by its owner
// thread and the fp16 activations are written back over its first 32 columns (read cols 0..31 -> write 0..15, read
// 32..63 -> write 16..31); the next layer accumulates into the OTHER half:
//     enc -> H1[0,E/2) | L0: D=H0 | h1 -> H0[0,32) | L1: D=H1 | h2 -> H1[0,32) | heads: D=H0[0,16) | cin -> H0[0,16)
//     | L3: D=H1 | c1 -> H1[0,32) | L4: D=H0 | c2 -> H0[0,32) | out: D=H1[0,16)
constexpr int FWG = 4;
constexpr uint32_t FWD_TMEM_COLS = 512;          // FWG windows of 128 columns
template <int E> struct FwdMap {
    static constexpr uint32_t w = 0;
    static constexpr uint32_t bars = (wmap(E).end + 127u) & ~127u;          // FWG mbarriers
    static constexpr uint32_t tmem_ptr = bars + FWG * 8u;
    static constexpr uint32_t bytes = tmem_ptr + 16u;
};

// TS-mode layer: D = A[tmem] W^T (+ 1 b^T through the shared-memory `one` tile when `bias` is given)
__device__ __forceinline__ void mma_layer_ts(uint32_t d, uint32_t a_tmem, const Tile& w, int N, int K, const Tile* one, const Tile* bias) {
    const uint32_t id = umma::make_idesc_f16(128, N, false, false);
    for (int ks = 0; ks < K / 16; ++ks) umma::mma_f16_ts(d, a_tmem + ks * 8, desc_k(w, ks), id, ks > 0);
    if (bias) umma::mma_f16_ss(d, desc_k(*one, 0), desc_k(*bias, 0), id, true);
}

// hidden-layer epilogue in tensor memory: 64 fp32 accumulator columns at `acc` (bias inside) -> ReLU -> 64 halves
// written back over columns [0,32) of the same window
__device__ __forceinline__ void epi_hidden_tmem(uint32_t acc) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float v[32];
        ld32(acc + h * 32, v);
        umma::wait_ld();
        uint32_t r[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) r[j] = pack_relu_h2(v[2 * j], v[2 * j + 1]);
        umma::st16(acc + h * 16, r);
    }
    umma::wait_st();
}

// ======================================================================================================
// backward: 2 slots x 1 tile, 256 threads per slot (two threads per row)
// ======================================================================================================
constexpr int BNS = 2;
// TMEM columns: per-slot D windows, then the persistent weight-gradient accumulators.  Bias gradients ride inside
// them (see "ones" below) except for the first trunk layer, which keeps a separate G^T 1 accumulator (8 columns).
constexpr uint32_t COL_ACC0 = 128;
constexpr uint32_t COL_WC1 = 128, COL_WT1 = 208, COL_WC0 = 288, COL_WT0 = 320, COL_WC2T = 384, COL_WHDT = 400, COL_BT0 = 416;
constexpr uint32_t COL_ACC1 = 432;
constexpr uint32_t BWD_TMEM_COLS = 512;
constexpr int HW = 72;   // hidden activation tiles are 72 columns wide: 64 units + a chunk [1,0,0,0,0,0,0,0]

// Per-slot tiles (byte offsets inside the slot block).  The four hidden tiles carry a constant ONES column (col 64):
// as the B operand of a weight-gradient MMA (N = 72) it makes column 64 of dW the bias gradient, as the A operand
// of a transposed one (M = 128) it makes lane 64 the bias gradient -- no separate G^T 1 MMAs (48 of 140 per tile).
// The colour-MLP input has a free column (31) that plays the same role.  `dg` holds d rgb_raw (forward end) and
// later the heads gradient: their lifetimes are disjoint.  MN-major M=128 views read up to 2 KB past a tile into
// whatever follows, which must be mapped shared memory.
template <int E> struct SlotMap {
    static constexpr uint32_t dg = 0, c2 = 4096, c1 = c2 + TM * HW * 2, h2 = c1 + TM * HW * 2, h1 = h2 + TM * HW * 2;
    static constexpr uint32_t xe = h1 + TM * HW * 2, cin = xe + TM * E * 2, bytes = cin + TM * 32 * 2;
};
template <int E> struct BwdMap {
    static constexpr uint32_t slots = 0;
    static constexpr uint32_t w = BNS * SlotMap<E>::bytes;
    static constexpr uint32_t bars = (w + wmap(E).end + 127u) & ~127u;     // per slot: done_d, done_w
    static constexpr uint32_t tmem_ptr = bars + BNS * 2 * 8u;
    static constexpr uint32_t res = tmem_ptr + 16u;                         // fused scatter: float resolution per level (16)
    static constexpr uint32_t bytes = res + 64u;
};

// Fused table scatter (acn_render_expert_bwd): instead of storing d_enc, the thread that holds 16 columns (8 levels) of
// a point's encoding gradient forms the point's cell at each of those levels and adds w_corner * g to the 8 corner rows
// of the table gradient -- models/encodings.py:331-381 differentiated w.r.t. the table, the same arithmetic as
// k_hashgrid_bwd.  The 2 GiB (P,E) fp32 d_enc round trip through HBM disappears and the REDs (bound by the per-SM
// atomic issue rate, not by anything the MMAs use) overlap the tensor-core work of the tiles in flight.
struct ScatterArgs {
    const float* x; int xs;                           // explicit positions (P,>=3), or
    const float* rays; const float* t; int S;         // rays (N,8) + t (N,S): p = o + d*t (nerfs/ray_rendering.py:317)
    const float* box6;                                // [min, extent] (world -> unit) or null
    float* dtable; int L; int log2T; const int32_t* res; int interp;
};

// power-of-two loss scale from max |dL/dy|: max * scale in [2^9, 2^10)
__device__ __forceinline__ float grad_scale_from_max(float mx) {
    if (!(mx > 0.0f) || !(mx < 3.0e38f)) return 1.0f;
    int e;
    frexpf(mx, &e);                       // mx = m * 2^e, m in [0.5, 1)
    e = 10 - e;
    e = e < -100 ? -100 : (e > 100 ? 100 : e);
    return ldexpf(1.0f, e);
}

__global__ void __launch_bounds__(256) k_absmax(const float4* __restrict__ x, int64_t n4, unsigned int* __restrict__ out,
                                                const int32_t* __restrict__ range) {
    if (range) { const int64_t r0 = __ldg(range); x += r0; n4 = __ldg(range + 1) - r0; }
    float m = 0.0f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = ld_stream_f4(x + i);
        m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));   // fmaxf drops NaNs
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(out, __float_as_uint(m));
}

// backward epilogue, part 1: 32 columns of D -> fp16 -> * [act > 0] in registers (the act tile is only read)
__device__ __forceinline__ void epi_mask32_load(uint32_t tmem_d, int col0, const Tile& act, int row, uint4* o) {
    float v[32];
    ld32(tmem_d + col0, v);
    uint4 a[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) a[q] = lds128(chunk_addr(act, row, col0 / 8 + q));
    umma::wait_ld();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        o[q].x = hmask2(pack_h2(v[8 * q + 0], v[8 * q + 1]), a[q].x);
        o[q].y = hmask2(pack_h2(v[8 * q + 2], v[8 * q + 3]), a[q].y);
        o[q].z = hmask2(pack_h2(v[8 * q + 4], v[8 * q + 5]), a[q].z);
        o[q].w = hmask2(pack_h2(v[8 * q + 6], v[8 * q + 7]), a[q].w);
    }
}
// part 2 (after the weight-gradient MMAs that read the act tile have completed): overwrite it in place
__device__ __forceinline__ void epi_mask32_store(const Tile& act, int row, int col0, const uint4* o) {
#pragma unroll
    for (int q = 0; q < 4; ++q) sts128(chunk_addr(act, row, col0 / 8 + q), o[q]);
}

int check_dims(const char* fn, int enc_dtype, int E, int H, int G, int C, const void* enc) {
    ACN_REQUIRE(enc_dtype == ACN_F16, ACN_EUNSUPPORTED, "%s(f16): the tensor-core path takes fp16 encodings (cast in the caller)", fn);
    ACN_REQUIRE(H == 64 && C == 64, ACN_EUNSUPPORTED, "%s(f16): hidden widths must be 64 (got H=%d, C=%d)", fn, H, C);
    ACN_REQUIRE(E == 16 || E == 32 || E == 48 || E == 64, ACN_EUNSUPPORTED, "%s(f16): encoding width %d not in {16,32,48,64}", fn, E);
    ACN_REQUIRE(G >= 1 && G <= 15, ACN_EUNSUPPORTED, "%s(f16): geo_feat_dim %d outside [1,15]", fn, G);
    ACN_REQUIRE(((uintptr_t)enc & 15) == 0, ACN_EINVAL, "%s(f16): enc must be 16-byte aligned", fn);
    return ACN_OK;
}


int absmax_word(acn_ctx* ctx, const char* fn, const float* d_rgb_sigma, int64_t P, const int32_t* range, cudaStream_t st, unsigned int** out) {
    unsigned int* slot = acn_scratch_word(ctx);
    ACN_REQUIRE(slot != nullptr, ACN_ECUDA, "%s(f16): no context scratch", fn);
    ACN_CUDA(cudaMemsetAsync(slot, 0, sizeof(unsigned int), st));
    k_absmax<<<acn_grid_1d(P, 256 * 8, (int64_t)ctx->sm_count * 8), 256, 0, st>>>((const float4*)d_rgb_sigma, P, slot, range);
    ACN_CHECK_LAUNCH();
    *out = slot;
    return ACN_OK;
}

}  // namespace
