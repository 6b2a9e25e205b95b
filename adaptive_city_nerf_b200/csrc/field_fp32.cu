// Stage 3, fp32 precision: the expert's density trunk + heads + SH + colour MLP
// (models/inr/meta_ngp.py:171-241) as ONE SIMT kernel per direction.  This is the path the
// reference runs with autocast off (meta-training query loss, pipelines/offline_stage/
// meta_train_step.py:110) and the tight-tolerance (<= 1e-5) parity anchor for the tcgen05
// kernels in field_mma.cu, which serve the autocast(fp16) / TF32-allowed case.
//
// A 256-thread block owns a tile of 64 points.  Weights live in shared memory in nn.Linear
// layout; activations live in shared memory feature-major ([feature][point], row stride LDP),
// so every layer is a block-cooperative 64x64xK register-tiled GEMM (4x4 outputs per thread).
// The backward recomputes the forward into shared memory, then walks the layers in reverse:
// dgrad is the same GEMM on W^T, wgrad contracts over the 64 points into per-thread register
// accumulators that persist across tiles and are flushed with one atomicAdd per weight per block.
#include "field_common.cuh"
#include "field_internal.cuh"

namespace {

constexpr int TP = 64;    // points per tile
constexpr int NT = 256;   // threads per block
constexpr int LDP = 68;   // activation row stride (floats): 16B-aligned rows, spreads banks

struct FieldDims { int E, H, G, C, HM, HMP, CIN, CINP; };

__host__ __device__ inline int roundup4(int v) { return (v + 3) & ~3; }

struct Smem {
    float *w_t0, *b_t0, *w_t1, *b_t1, *w_hd, *b_hd, *w_c0, *b_c0, *w_c1, *b_c1, *w_c2, *b_c2;
    float *xe, *h1, *h2, *cin, *c1, *c2, *raw, *sg, *da, *db;
};

__host__ __device__ inline size_t carve(const FieldDims& d, bool bwd, float* base, Smem* s) {
    size_t off = 0;
    auto take = [&](size_t n) { float* p = base ? base + off : nullptr; off += (n + 3) & ~(size_t)3; return p; };
    float* w_t0 = take(64 * d.E); float* b_t0 = take(64);
    float* w_t1 = take(64 * 64);  float* b_t1 = take(64);
    float* w_hd = take(d.HMP * 64); float* b_hd = take(d.HMP);
    float* w_c0 = take(64 * d.CINP); float* b_c0 = take(64);
    float* w_c1 = take(64 * 64);  float* b_c1 = take(64);
    float* w_c2 = take(4 * 64);   float* b_c2 = take(4);
    float* xe = take(d.E * LDP);  float* h1 = take(64 * LDP); float* h2 = take(64 * LDP);
    float* cin = take(d.CINP * LDP); float* c1 = take(64 * LDP); float* c2 = take(64 * LDP);
    float* raw = take(4 * LDP);   float* sg = take(64);
    float* da = bwd ? take(64 * LDP) : nullptr;
    float* db = bwd ? take(64 * LDP) : nullptr;
    if (s) *s = Smem{ w_t0, b_t0, w_t1, b_t1, w_hd, b_hd, w_c0, b_c0, w_c1, b_c1, w_c2, b_c2,
                      xe, h1, h2, cin, c1, c2, raw, sg, da, db };
    return off * sizeof(float);
}

// ---- weights -> shared memory (zero padded) ------------------------------------------------
__device__ void load_weights(const acn_field_weights& w, const FieldDims& d, const Smem& s) {
    const int tid = threadIdx.x;
    for (int i = tid; i < 64 * d.E; i += NT) s.w_t0[i] = __ldg(w.p[0] + i);
    for (int i = tid; i < 64 * 64; i += NT) { s.w_t1[i] = __ldg(w.p[2] + i); s.w_c1[i] = __ldg(w.p[10] + i); }
    for (int i = tid; i < d.HMP * 64; i += NT) {
        int r = i >> 6, c = i & 63;
        s.w_hd[i] = r < d.G ? __ldg(w.p[6] + r * 64 + c) : (r == d.G ? __ldg(w.p[4] + c) : 0.0f);
    }
    for (int i = tid; i < 64 * d.CINP; i += NT) {
        int r = i / d.CINP, c = i - r * d.CINP;
        s.w_c0[i] = c < d.CIN ? __ldg(w.p[8] + r * d.CIN + c) : 0.0f;
    }
    for (int i = tid; i < 4 * 64; i += NT) s.w_c2[i] = i < 3 * 64 ? __ldg(w.p[12] + i) : 0.0f;
    if (tid < 64) {
        s.b_t0[tid] = __ldg(w.p[1] + tid); s.b_t1[tid] = __ldg(w.p[3] + tid);
        s.b_c0[tid] = __ldg(w.p[9] + tid); s.b_c1[tid] = __ldg(w.p[11] + tid);
    }
    if (tid < d.HMP) s.b_hd[tid] = tid < d.G ? __ldg(w.p[7] + tid) : (tid == d.G ? __ldg(w.p[5]) : 0.0f);
    if (tid < 4) s.b_c2[tid] = tid < 3 ? __ldg(w.p[13] + tid) : 0.0f;
}

// ---- register-tiled GEMMs over a 64-point tile ----------------------------------------------
// acc[j][c] = sum_{k<K} A(m0+j, k) * B[k*LDP + n0 + c]
//   AMODE 0: A(m,k) = A[m*lda + k]  (forward: W[o][i])      AMODE 1: A(m,k) = A[k*lda + m]  (dgrad: W^T)
template <int AMODE>
__device__ __forceinline__ void mm_points(const float* __restrict__ A, int lda, const float* __restrict__ B, int K,
                                          int m0, int n0, float (&acc)[4][4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[j][c] = 0.0f;
    for (int k = 0; k < K; ++k) {
        float a[4];
        if (AMODE == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) a[j] = A[(m0 + j) * lda + k];
        } else {
            float4 v = *reinterpret_cast<const float4*>(A + k * lda + m0);
            a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w;
        }
        float4 b = *reinterpret_cast<const float4*>(B + k * LDP + n0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            acc[j][0] = fmaf(a[j], b.x, acc[j][0]); acc[j][1] = fmaf(a[j], b.y, acc[j][1]);
            acc[j][2] = fmaf(a[j], b.z, acc[j][2]); acc[j][3] = fmaf(a[j], b.w, acc[j][3]);
        }
    }
}

// wgrad: acc[j][c] += sum_{k<TP} D[(m0+j)*LDP + k] * X[(tn + 16c)*LDP + k]   (n = tn + 16c < N)
__device__ __forceinline__ void mm_wgrad(const float* __restrict__ D, const float* __restrict__ X, int N, int m0, int tn,
                                         float (&acc)[4][4]) {
    for (int k = 0; k < TP; k += 4) {
        float4 dv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) dv[j] = *reinterpret_cast<const float4*>(D + (m0 + j) * LDP + k);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            int n = tn + 16 * c;
            if (n < N) {
                float4 xv = *reinterpret_cast<const float4*>(X + n * LDP + k);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    acc[j][c] += dv[j].x * xv.x + dv[j].y * xv.y + dv[j].z * xv.z + dv[j].w * xv.w;
            }
        }
    }
}

// One dense layer on the tile: OUT[m][n] = act(sum_k W[m][k] IN[k][n] + b[m]), m < M (M % 4 == 0).
template <bool RELU>
__device__ __forceinline__ void layer_fwd(const float* W, int lda, const float* b, const float* IN, int K, int M, float* OUT) {
    const int m0 = (threadIdx.x >> 4) * 4, n0 = (threadIdx.x & 15) * 4;
    if (m0 < M) {
        float acc[4][4];
        mm_points<0>(W, lda, IN, K, m0, n0, acc);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float bb = b[m0 + j];
            float4 o = make_float4(acc[j][0] + bb, acc[j][1] + bb, acc[j][2] + bb, acc[j][3] + bb);
            if (RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            *reinterpret_cast<float4*>(OUT + (m0 + j) * LDP + n0) = o;
        }
    }
}

// dgrad: OUT[i][n] = (ACT[i][n] > 0 ? 1 : 0 if MASK) * sum_o W[o][i] DIN[o][n], i < M, o < K
template <bool MASK>
__device__ __forceinline__ void layer_dgrad(const float* W, int lda, const float* DIN, int K, int M, const float* ACT, float* OUT) {
    const int m0 = (threadIdx.x >> 4) * 4, n0 = (threadIdx.x & 15) * 4;
    if (m0 < M) {
        float acc[4][4];
        mm_points<1>(W, lda, DIN, K, m0, n0, acc);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float4 o = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
            if (MASK) {
                float4 a = *reinterpret_cast<const float4*>(ACT + (m0 + j) * LDP + n0);
                o.x = a.x > 0.f ? o.x : 0.f; o.y = a.y > 0.f ? o.y : 0.f; o.z = a.z > 0.f ? o.z : 0.f; o.w = a.w > 0.f ? o.w : 0.f;
            }
            *reinterpret_cast<float4*>(OUT + (m0 + j) * LDP + n0) = o;
        }
    }
}

template <typename EncT> __device__ __forceinline__ float enc_ld(const EncT* p);
template <> __device__ __forceinline__ float enc_ld<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float enc_ld<__half>(const __half* p) { return __half2float(__ldg(p)); }

// Forward of one tile into shared memory (everything the backward needs stays resident).
template <typename EncT>
__device__ void tile_forward(const EncT* __restrict__ enc, const float* __restrict__ dirs, int dstride, int dgroup,
                             int64_t p0, int64_t P, const FieldDims& d, const Smem& s) {
    const int tid = threadIdx.x;
    for (int idx = tid; idx < d.E * TP; idx += NT) {
        int n = idx / d.E, i = idx - n * d.E;
        int64_t p = p0 + n;
        s.xe[i * LDP + n] = p < P ? enc_ld<EncT>(enc + p * d.E + i) : 0.0f;
    }
    __syncthreads();
    layer_fwd<true>(s.w_t0, d.E, s.b_t0, s.xe, d.E, 64, s.h1);
    __syncthreads();
    layer_fwd<true>(s.w_t1, 64, s.b_t1, s.h1, 64, 64, s.h2);
    __syncthreads();
    {   // heads: rows < G -> geo -> CIN; row G -> sigma_raw -> SG
        const int m0 = (tid >> 4) * 4, n0 = (tid & 15) * 4;
        if (m0 < d.HMP) {
            float acc[4][4];
            mm_points<0>(s.w_hd, 64, s.h2, 64, m0, n0, acc);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int m = m0 + j;
                float bb = s.b_hd[m];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float v = acc[j][c] + bb;
                    if (m < d.G) s.cin[m * LDP + n0 + c] = v;
                    else if (m == d.G) s.sg[n0 + c] = v;
                }
            }
        }
        // SH rows G..G+15 and zero padding rows, one point per thread
        if (tid < TP) {
            int64_t p = p0 + tid;
            float sh[16];
            if (p < P) {
                const float* dp = dir_of(dirs, dstride, dgroup, p);
                sh16_expert(__ldg(dp), __ldg(dp + 1), __ldg(dp + 2), sh);
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) sh[j] = 0.0f;
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) s.cin[(d.G + j) * LDP + tid] = sh[j];
            for (int r = d.CIN; r < d.CINP; ++r) s.cin[r * LDP + tid] = 0.0f;
        }
    }
    __syncthreads();
    layer_fwd<true>(s.w_c0, d.CINP, s.b_c0, s.cin, d.CINP, 64, s.c1);
    __syncthreads();
    layer_fwd<true>(s.w_c1, 64, s.b_c1, s.c1, 64, 64, s.c2);
    __syncthreads();
    layer_fwd<false>(s.w_c2, 64, s.b_c2, s.c2, 64, 4, s.raw);
    __syncthreads();
}

template <typename EncT>
__global__ void __launch_bounds__(NT, 1) k_field_fwd_fp32(
    const EncT* __restrict__ enc, const float* __restrict__ dirs, int dstride, int dgroup, int64_t P, FieldDims d,
    acn_field_weights w, float4* __restrict__ rgb_sigma, const int32_t* __restrict__ range)
{
    extern __shared__ __align__(16) float smem[];
    if (range) {            // rows [range[0], range[1]) only: P was just the launch's upper bound
        const int64_t r0 = __ldg(range);
        P = __ldg(range + 1) - r0;
        enc += r0 * d.E; dirs += r0 * dstride; rgb_sigma += r0;
    }
    Smem s;
    carve(d, false, smem, &s);
    load_weights(w, d, s);
    __syncthreads();
    const int64_t ntiles = (P + TP - 1) / TP;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t p0 = tile * TP;
        tile_forward<EncT>(enc, dirs, dstride, dgroup, p0, P, d, s);
        if (threadIdx.x < TP && p0 + threadIdx.x < P) {
            int n = threadIdx.x;
            rgb_sigma[p0 + n] = make_float4(sigmoid_f(s.raw[0 * LDP + n]), sigmoid_f(s.raw[1 * LDP + n]),
                                            sigmoid_f(s.raw[2 * LDP + n]), trunc_exp_f(s.sg[n]));
        }
        __syncthreads();
    }
}

// bias-gradient partial: sum over the tile's points of row `tid` of D
__device__ __forceinline__ float row_sum(const float* D, int row) {
    float a = 0.0f;
#pragma unroll 8
    for (int n = 0; n < TP; ++n) a += D[row * LDP + n];
    return a;
}

template <typename EncT, typename DEncT>
__global__ void __launch_bounds__(NT, 1) k_field_bwd_fp32(
    const EncT* __restrict__ enc, const float* __restrict__ dirs, int dstride, int dgroup, int64_t P, FieldDims d,
    acn_field_weights w, const float4* __restrict__ d_rgb_sigma, acn_field_grads g, DEncT* __restrict__ d_enc,
    const int32_t* __restrict__ range)
{
    extern __shared__ __align__(16) float smem[];
    if (range) {            // rows [range[0], range[1]) only: P was just the launch's upper bound
        const int64_t r0 = __ldg(range);
        P = __ldg(range + 1) - r0;
        enc += r0 * d.E; dirs += r0 * dstride; d_rgb_sigma += r0;
        if (d_enc) d_enc += r0 * d.E;
    }
    Smem s;
    carve(d, true, smem, &s);
    load_weights(w, d, s);
    __syncthreads();
    const int tid = threadIdx.x;
    const int m0 = (tid >> 4) * 4, tn = tid & 15;
    // persistent wgrad accumulators (micro-tile rows m0..m0+3, columns tn + 16c)
    float a_c2[4][4] = {}, a_c1[4][4] = {}, a_c0[4][4] = {}, a_hd[4][4] = {}, a_t1[4][4] = {}, a_t0[4][4] = {};
    float b_c2 = 0.f, b_c1 = 0.f, b_c0 = 0.f, b_hd = 0.f, b_t1 = 0.f, b_t0 = 0.f;
    const int64_t ntiles = (P + TP - 1) / TP;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t p0 = tile * TP;
        tile_forward<EncT>(enc, dirs, dstride, dgroup, p0, P, d, s);
        // output grads -> DA rows 0..3 (rgb raw, row 3 = 0) and SG := d sigma_raw
        if (tid < TP) {
            int n = tid;
            int64_t p = p0 + n;
            float4 dy = p < P ? d_rgb_sigma[p] : make_float4(0.f, 0.f, 0.f, 0.f);
            float y0 = sigmoid_f(s.raw[n]), y1 = sigmoid_f(s.raw[LDP + n]), y2 = sigmoid_f(s.raw[2 * LDP + n]);
            s.da[n] = dy.x * y0 * (1.0f - y0);
            s.da[LDP + n] = dy.y * y1 * (1.0f - y1);
            s.da[2 * LDP + n] = dy.z * y2 * (1.0f - y2);
            s.da[3 * LDP + n] = 0.0f;
            s.sg[n] = dy.w * trunc_exp_f(s.sg[n]);
        }
        __syncthreads();
        // colour head (3 x 64)
        if (m0 < 4) mm_wgrad(s.da, s.c2, 64, m0, tn, a_c2);
        if (tid < 3) b_c2 += row_sum(s.da, tid);
        layer_dgrad<true>(s.w_c2, 64, s.da, 4, 64, s.c2, s.db);
        __syncthreads();
        // colour hidden 2
        mm_wgrad(s.db, s.c1, 64, m0, tn, a_c1);
        if (tid < 64) b_c1 += row_sum(s.db, tid);
        layer_dgrad<true>(s.w_c1, 64, s.db, 64, 64, s.c1, s.da);
        __syncthreads();
        // colour hidden 1: input = [geo, sh]
        mm_wgrad(s.da, s.cin, d.CINP, m0, tn, a_c0);
        if (tid < 64) b_c0 += row_sum(s.da, tid);
        layer_dgrad<false>(s.w_c0, d.CINP, s.da, 64, d.CINP, nullptr, s.db);
        __syncthreads();
        // heads: rows < G carry d geo; row G = d sigma_raw; padding rows zero
        if (tid < TP) {
            s.db[d.G * LDP + tid] = s.sg[tid];
            for (int r = d.G + 1; r < d.HMP; ++r) s.db[r * LDP + tid] = 0.0f;
        }
        __syncthreads();
        if (m0 < d.HMP) mm_wgrad(s.db, s.h2, 64, m0, tn, a_hd);
        if (tid < d.HM) b_hd += row_sum(s.db, tid);
        layer_dgrad<true>(s.w_hd, 64, s.db, d.HMP, 64, s.h2, s.da);
        __syncthreads();
        // trunk 2
        mm_wgrad(s.da, s.h1, 64, m0, tn, a_t1);
        if (tid < 64) b_t1 += row_sum(s.da, tid);
        layer_dgrad<true>(s.w_t1, 64, s.da, 64, 64, s.h1, s.db);
        __syncthreads();
        // trunk 1
        mm_wgrad(s.db, s.xe, d.E, m0, tn, a_t0);
        if (tid < 64) b_t0 += row_sum(s.db, tid);
        if (d_enc) {
            layer_dgrad<false>(s.w_t0, d.E, s.db, 64, d.E, nullptr, s.da);
            __syncthreads();
            for (int idx = tid; idx < d.E * TP; idx += NT) {
                int n = idx / d.E, i = idx - n * d.E;
                int64_t p = p0 + n;
                if (p < P) {
                    float v = s.da[i * LDP + n];
                    if constexpr (sizeof(DEncT) == 2) d_enc[p * d.E + i] = __float2half_rn(v);
                    else d_enc[p * d.E + i] = v;
                }
            }
        }
        __syncthreads();
    }
    // flush: weight (m, n = tn + 16c)
    auto flush = [&](float* gw, float (&acc)[4][4], int M, int N, int ld) {
        if (!gw) return;
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                int m = m0 + j, n = tn + 16 * c;
                if (m < M && n < N && acc[j][c] != 0.0f) atomicAdd(gw + m * ld + n, acc[j][c]);
            }
    };
    flush(g.p[12], a_c2, 3, 64, 64);
    flush(g.p[10], a_c1, 64, 64, 64);
    flush(g.p[8], a_c0, 64, d.CIN, d.CIN);
    flush(g.p[6], a_hd, d.G, 64, 64);
    if (g.p[4]) {   // sigma head = row G of the fused heads matrix
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (m0 + j == d.G) {
#pragma unroll
                for (int c = 0; c < 4; ++c) atomicAdd(g.p[4] + tn + 16 * c, a_hd[j][c]);
            }
    }
    flush(g.p[2], a_t1, 64, 64, 64);
    flush(g.p[0], a_t0, 64, d.E, d.E);
    if (tid < 3 && g.p[13]) atomicAdd(g.p[13] + tid, b_c2);
    if (tid < 64) {
        if (g.p[11]) atomicAdd(g.p[11] + tid, b_c1);
        if (g.p[9]) atomicAdd(g.p[9] + tid, b_c0);
        if (g.p[3]) atomicAdd(g.p[3] + tid, b_t1);
        if (g.p[1]) atomicAdd(g.p[1] + tid, b_t0);
    }
    if (tid < d.G && g.p[7]) atomicAdd(g.p[7] + tid, b_hd);
    if (tid == d.G && g.p[5]) atomicAdd(g.p[5], b_hd);
}

int make_dims(const char* fn, int E, int H, int G, int C, FieldDims* d) {
    ACN_REQUIRE(H == 64 && C == 64, ACN_EUNSUPPORTED, "%s: hidden widths must be 64 (got H=%d, C=%d)", fn, H, C);
    ACN_REQUIRE(E >= 4 && E <= 64 && (E % 4) == 0, ACN_EUNSUPPORTED, "%s: encoding width %d not a multiple of 4 in [4,64]", fn, E);
    ACN_REQUIRE(G >= 1 && G <= 15, ACN_EUNSUPPORTED, "%s: geo_feat_dim %d outside [1,15]", fn, G);
    d->E = E; d->H = H; d->G = G; d->C = C;
    d->HM = G + 1; d->HMP = roundup4(G + 1); d->CIN = G + 16; d->CINP = roundup4(G + 16);
    return ACN_OK;
}

int field_grid(acn_ctx* ctx, int64_t P) {
    int64_t tiles = (P + TP - 1) / TP;
    int64_t g = tiles < ctx->sm_count ? tiles : ctx->sm_count;
    return (int)(g < 1 ? 1 : g);
}

}  // namespace

// ------------------------------------------------------------------------------- launchers
int acn_field_fwd_fp32(acn_ctx* ctx, const void* enc, int enc_dtype, const float* dirs, int dirs_stride, int dirs_group,
                       int64_t P, int E, int H, int G, int C, const acn_field_weights* w, float* rgb_sigma, const int32_t* range,
                       cudaStream_t st) {
    FieldDims d;
    int rc = make_dims("acn_field_fwd", E, H, G, C, &d);
    if (rc) return rc;
    size_t smem = carve(d, false, nullptr, nullptr);
    ACN_REQUIRE((int)smem <= ctx->max_smem_optin, ACN_EUNSUPPORTED, "acn_field_fwd: needs %zu B shared memory", smem);
    if (enc_dtype == ACN_F32) {
        ACN_CUDA(cudaFuncSetAttribute(k_field_fwd_fp32<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_field_fwd_fp32<float><<<field_grid(ctx, P), NT, smem, st>>>((const float*)enc, dirs, dirs_stride, dirs_group, P, d, *w,
                                                                     (float4*)rgb_sigma, range);
    } else {
        ACN_CUDA(cudaFuncSetAttribute(k_field_fwd_fp32<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_field_fwd_fp32<__half><<<field_grid(ctx, P), NT, smem, st>>>((const __half*)enc, dirs, dirs_stride, dirs_group, P, d, *w,
                                                                      (float4*)rgb_sigma, range);
    }
    ACN_CHECK_LAUNCH();
    return ACN_OK;
}

int acn_field_bwd_fp32(acn_ctx* ctx, const void* enc, int enc_dtype, const float* dirs, int dirs_stride, int dirs_group,
                       int64_t P, int E, int H, int G, int C, const acn_field_weights* w, const float* d_rgb_sigma,
                       const acn_field_grads* g, void* d_enc, int d_enc_dtype, const int32_t* range, cudaStream_t st) {
    FieldDims d;
    int rc = make_dims("acn_field_bwd", E, H, G, C, &d);
    if (rc) return rc;
    size_t smem = carve(d, true, nullptr, nullptr);
    ACN_REQUIRE((int)smem <= ctx->max_smem_optin, ACN_EUNSUPPORTED, "acn_field_bwd: needs %zu B shared memory", smem);
    const int grid = field_grid(ctx, P);
#define LAUNCH(ET, DT)                                                                                                   \
    do {                                                                                                                 \
        ACN_CUDA(cudaFuncSetAttribute(k_field_bwd_fp32<ET, DT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        k_field_bwd_fp32<ET, DT><<<grid, NT, smem, st>>>((const ET*)enc, dirs, dirs_stride, dirs_group, P, d, *w,         \
                                                         (const float4*)d_rgb_sigma, *g, (DT*)d_enc, range);             \
    } while (0)
    if (enc_dtype == ACN_F32) { if (d_enc_dtype == ACN_F32) LAUNCH(float, float); else LAUNCH(float, __half); }
    else                      { if (d_enc_dtype == ACN_F32) LAUNCH(__half, float); else LAUNCH(__half, __half); }
#undef LAUNCH
    ACN_CHECK_LAUNCH();
    return ACN_OK;
}
