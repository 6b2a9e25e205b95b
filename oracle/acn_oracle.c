/*
 * acn_oracle.c -- CPU restatement of the adaptive-city-nerf per-ray rendering hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (adaptive_city_nerf_b200/) never imports, links or executes anything in oracle/.
 *
 * Every function cites the reference lines (relative to /root/reference) it restates.
 * Parity status: PINNED -- tests/test_oracle_golden.py checks every function here against
 * outputs of the unmodified reference imported in-process (tests/golden/make_golden.py) and
 * against excerpts of the reference's shipped Voronoi masks.
 *
 * Build: gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off -fno-fast-math (oracle/Makefile).
 * -ffp-contract=off matters: integer-valued outputs (hash indices, sample bins, expert
 * assignment) depend on each fp32 op being rounded separately unless the reference itself
 * fuses (torch.lerp, the cdist matmul form) -- those spots call fmaf() explicitly.
 */
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

static inline float clampf(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }
static inline float f16r(float x) { return (float)(_Float16)x; } /* round-trip through IEEE half */

/* ------------------------------------------------------------------------------------------
 * Stage 1 -- camera rays, scene-box clipping, sampling
 * ---------------------------------------------------------------------------------------- */

/* nerfs/ray_sampling.py:111-136 get_ray_directions: [(i+.5-cx)/fx, -(j+.5-cy)/fy, -1], L2
 * normalised with the norm clamped at 1e-12.  Float-tolerance output (the norm reduction
 * order is a torch implementation detail). */
ORC_API void orc_ray_directions(int H, int W, float fx, float fy, float cx, float cy,
                                int center_pixels, float* dirs)
{
#pragma omp parallel for schedule(static)
    for (int j = 0; j < H; ++j)
        for (int i = 0; i < W; ++i) {
            float fi = (float)i, fj = (float)j;
            if (center_pixels) { fi = fi + 0.5f; fj = fj + 0.5f; }
            float x = (fi - cx) / fx;
            float y = -((fj - cy) / fy);
            float z = -1.0f;
            float n = sqrtf(x * x + y * y + z * z);
            if (n < 1e-12f) n = 1e-12f;
            float* o = dirs + 3 * ((size_t)j * W + i);
            o[0] = x / n; o[1] = y / n; o[2] = z / n;
        }
}

/* nerfs/scene_box.py:45-107 SceneBox.ray_aabb_intersect.  Slab test with an eps-guarded
 * reciprocal (IEEE divide, then (bound - o) * inv), clamp to [0,max_bound], invalid tagging. */
ORC_API void orc_aabb_intersect(const float* o, const float* d, int64_t N, int stride,
                                const float* aabb6, float eps, float max_bound, float invalid,
                                float* tmin_out, float* tmax_out)
{
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < N; ++r) {
        float tmn = -INFINITY, tmx = INFINITY;
        int nan_mn = 0, nan_mx = 0;
        for (int a = 0; a < 3; ++a) {
            float dd = d[r * stride + a], oo = o[r * stride + a];
            float rd = fabsf(dd) < eps ? (dd >= 0.0f ? eps : -eps) : dd;
            float inv = 1.0f / rd;
            float t0 = (aabb6[a] - oo) * inv;
            float t1 = (aabb6[3 + a] - oo) * inv;
            /* torch.minimum / maximum / amax / amin propagate NaN */
            float lo = (t0 != t0 || t1 != t1) ? NAN : fminf(t0, t1);
            float hi = (t0 != t0 || t1 != t1) ? NAN : fmaxf(t0, t1);
            if (lo != lo) nan_mn = 1; else if (lo > tmn) tmn = lo;
            if (hi != hi) nan_mx = 1; else if (hi < tmx) tmx = hi;
        }
        if (nan_mn) tmn = NAN;
        if (nan_mx) tmx = NAN;
        /* clamp(min=0,max=max_bound) keeps NaN */
        if (tmn == tmn) tmn = clampf(tmn, 0.0f, max_bound);
        if (tmx == tmx) tmx = clampf(tmx, 0.0f, max_bound);
        int inval = tmx <= tmn; /* NaN compares false -> stays as is */
        tmin_out[r] = inval ? invalid : tmn;
        tmax_out[r] = inval ? invalid : tmx;
    }
}

/* nerfs/ray_sampling.py:10-24, 50-108 _rays_cam_to_world + get_rays + pack_rays:
 * d_w = d_c @ R^T, o = t, near/far from the scene box (or constants), packed (N,8). */
ORC_API void orc_get_rays(const float* dirs_cam, int64_t N, const float* c2w /*3x4 row-major*/,
                          const float* aabb6_or_null, float near_c, float far_c,
                          float max_bound, float invalid, float* rays8)
{
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < N; ++r) {
        const float* dc = dirs_cam + 3 * r;
        float* o = rays8 + 8 * r;
        for (int j = 0; j < 3; ++j) {
            o[j] = c2w[4 * j + 3];
            float acc = dc[0] * c2w[4 * j + 0];
            acc = acc + dc[1] * c2w[4 * j + 1];
            acc = acc + dc[2] * c2w[4 * j + 2];
            o[3 + j] = acc;
        }
        if (!aabb6_or_null) { o[6] = near_c; o[7] = far_c; }
    }
    if (aabb6_or_null) {
        float* tmn = (float*)malloc(sizeof(float) * (size_t)(N > 0 ? N : 1));
        float* tmx = (float*)malloc(sizeof(float) * (size_t)(N > 0 ? N : 1));
        orc_aabb_intersect(rays8, rays8 + 3, N, 8, aabb6_or_null, 1e-8f, max_bound, invalid, tmn, tmx);
        for (int64_t r = 0; r < N; ++r) { rays8[8 * r + 6] = tmn[r]; rays8[8 * r + 7] = tmx[r]; }
        free(tmn); free(tmx);
    }
}

/* nerfs/ray_sampling.py:139-176 clamp_rays_near_far.  has_override=0 reproduces the
 * `near_far_override is None` branch (validity only, rays untouched). */
ORC_API void orc_clamp_near_far(float* rays8, int64_t N, int has_override, float n_or_nan,
                                float f_or_nan, float eps, float invalid, uint8_t* valid)
{
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < N; ++r) {
        float near = rays8[8 * r + 6], far = rays8[8 * r + 7];
        if (has_override) {
            /* torch.maximum / minimum propagate NaN from either side */
            if (n_or_nan == n_or_nan) near = (near != near) ? near : (near > n_or_nan ? near : n_or_nan);
            if (f_or_nan == f_or_nan) far = (far != far) ? far : (far < f_or_nan ? far : f_or_nan);
        }
        int v = isfinite(near) && isfinite(far) && (far > near + eps);
        valid[r] = (uint8_t)v;
        if (has_override) {
            rays8[8 * r + 6] = v ? near : invalid;
            rays8[8 * r + 7] = v ? far : invalid;
        }
    }
}

/* torch.linspace(0,1,S) on CPU, as used at nerfs/ray_rendering.py:278 and
 * scripts/create_clusters.py:596: step = 1/(S-1); first half start + step*i, second half
 * end - step*(S-1-i), the latter with a single rounding. */
ORC_API void orc_linspace01(int S, float* out)
{
    if (S == 1) { out[0] = 0.0f; return; }
    float step = 1.0f / (float)(S - 1);
    int half = S / 2;
    for (int i = 0; i < S; ++i)
        out[i] = i < half ? step * (float)i : fmaf(-step, (float)(S - 1 - i), 1.0f);
}

/* nerfs/ray_rendering.py:262-287 stratified_t_vals.  t = near*(1-u) + far*u; with jitter:
 * mids, [t0,mids], [mids,tS-1], t = lo + (hi-lo)*rand.  `jitter` is the rand tensor (N,S)
 * or NULL for eval.  Every op separately rounded. */
ORC_API void orc_stratified_t(const float* rays8, int64_t N, int S, const float* u_lin,
                              const float* jitter_or_null, float* t_vals)
{
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < N; ++r) {
        float near = rays8[8 * r + 6], far = rays8[8 * r + 7];
        float* t = t_vals + (size_t)r * S;
        float base[S];
        for (int s = 0; s < S; ++s) {
            float a = near * (1.0f - u_lin[s]);
            float b = far * u_lin[s];
            base[s] = a + b;
        }
        if (!jitter_or_null) { for (int s = 0; s < S; ++s) t[s] = base[s]; continue; }
        const float* u = jitter_or_null + (size_t)r * S;
        for (int s = 0; s < S; ++s) {
            float lo = (s == 0) ? base[0] : 0.5f * (base[s - 1] + base[s]);
            float hi = (s == S - 1) ? base[S - 1] : 0.5f * (base[s] + base[s + 1]);
            float span = hi - lo;
            t[s] = lo + span * u[s];
        }
    }
}

/* nerfs/ray_rendering.py:317 pts = o + d*t (un-fused multiply then add). */
ORC_API void orc_points(const float* rays8, int64_t N, int S, const float* t_vals, float* pts)
{
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < N; ++r)
        for (int s = 0; s < S; ++s) {
            float t = t_vals[(size_t)r * S + s];
            for (int a = 0; a < 3; ++a) {
                float m = rays8[8 * r + 3 + a] * t;
                pts[((size_t)r * S + s) * 3 + a] = rays8[8 * r + a] + m;
            }
        }
}

/* ------------------------------------------------------------------------------------------
 * Stage 2 -- multiresolution hash grid (torch branch of the reference)
 * ---------------------------------------------------------------------------------------- */

/* models/inr/meta_ngp.py:155-158 _world_to_unit: (x - min) / extent (true division), clamp to
 * [fp32(1e-6), 1 - fp32(1e-6)]. */
ORC_API void orc_world_to_unit(const float* x, int64_t P, const float* box_min3,
                               const float* extent3, float* x01)
{
    const float eps = 1e-6f;
    const float hi = 1.0f - eps;
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < P; ++p)
        for (int a = 0; a < 3; ++a) {
            float v = (x[3 * p + a] - box_min3[a]) / extent3[a];
            x01[3 * p + a] = (v != v) ? v : clampf(v, eps, hi);
        }
}

/* models/encodings.py:211-214: level resolutions = floor(min_res * g^l) computed in fp32,
 * g = exp((ln max - ln min)/(L-1)) in double then stored as python float. torch computes
 * `growth ** arange(L, float32)` as powf on fp32 and multiplies by min_res in fp32. */
ORC_API void orc_level_resolutions(int L, int min_res, int max_res, int32_t* res)
{
    double g = (L <= 1) ? 1.0 : exp((log((double)max_res) - log((double)min_res)) / (double)(L - 1));
    for (int l = 0; l < L; ++l) {
        float p = powf((float)g, (float)l);
        float s = (float)min_res * p;
        res[l] = (int32_t)floorf(s);
    }
}

/* models/encodings.py:308-316 _hash: ((ix*1) ^ (iy*2654435761) ^ (iz*805459861)) % 2^log2T in
 * int64 (python-style remainder == mask for a power of two). */
static inline int64_t orc_hash(int64_t ix, int64_t iy, int64_t iz, int log2T)
{
    int64_t h = (ix * 1LL) ^ (iy * 2654435761LL) ^ (iz * 805459861LL);
    return h & (((int64_t)1 << log2T) - 1);
}

/* interp: 0 = Nearest, 1 = Linear, 2 = Smoothstep (models/encodings.py:156, 340-381). */
/* models/encodings.py:331-381 _torch_forward.  x01 (P,3) -> out (P, L*F), level-major.
 * idx_out (optional, (P,L,8) int32, corner order 000,001,010,011,100,101,110,111 with the
 * digits meaning (x,y,z) ceil flags as in the reference's f000..f111 names) receives the
 * absolute table rows (hash + l*T); for Nearest only slot 0 is written. */
ORC_API void orc_hashgrid_fwd(const float* x01, int64_t P, const float* table, int L, int F,
                              int log2T, const int32_t* res, int interp, float* out,
                              int32_t* idx_out)
{
    const int64_t T = (int64_t)1 << log2T;
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < P; ++p) {
        for (int l = 0; l < L; ++l) {
            float rl = (float)res[l];
            float sx = x01[3 * p + 0] * rl, sy = x01[3 * p + 1] * rl, sz = x01[3 * p + 2] * rl;
            float* o = out + ((size_t)p * L + l) * F;
            if (interp == 0) {
                int64_t ix = (int64_t)nearbyintf(sx), iy = (int64_t)nearbyintf(sy), iz = (int64_t)nearbyintf(sz);
                int64_t row = orc_hash(ix, iy, iz, log2T) + (int64_t)l * T;
                if (idx_out) idx_out[((size_t)p * L + l) * 8] = (int32_t)row;
                for (int f = 0; f < F; ++f) o[f] = table[row * F + f];
                continue;
            }
            float fx = floorf(sx), fy = floorf(sy), fz = floorf(sz);
            float wx = sx - fx, wy = sy - fy, wz = sz - fz;
            int64_t x0 = (int64_t)fx, y0 = (int64_t)fy, z0 = (int64_t)fz;
            int64_t rows[8];
            for (int c = 0; c < 8; ++c) {
                int64_t ix = x0 + ((c >> 2) & 1), iy = y0 + ((c >> 1) & 1), iz = z0 + (c & 1);
                rows[c] = orc_hash(ix, iy, iz, log2T) + (int64_t)l * T;
                if (idx_out) idx_out[((size_t)p * L + l) * 8 + c] = (int32_t)rows[c];
            }
            if (interp == 2) {
                wx = wx * wx * (3.0f - 2.0f * wx);
                wy = wy * wy * (3.0f - 2.0f * wy);
                wz = wz * wz * (3.0f - 2.0f * wz);
            }
            float ux = 1.0f - wx, uy = 1.0f - wy, uz = 1.0f - wz;
            for (int f = 0; f < F; ++f) {
                float f000 = table[rows[0] * F + f], f001 = table[rows[1] * F + f];
                float f010 = table[rows[2] * F + f], f011 = table[rows[3] * F + f];
                float f100 = table[rows[4] * F + f], f101 = table[rows[5] * F + f];
                float f110 = table[rows[6] * F + f], f111 = table[rows[7] * F + f];
                float c00 = f000 * ux + f100 * wx;
                float c01 = f001 * ux + f101 * wx;
                float c10 = f010 * ux + f110 * wx;
                float c11 = f011 * ux + f111 * wx;
                float c0 = c00 * uy + c10 * wy;
                float c1 = c01 * uy + c11 * wy;
                o[f] = c0 * uz + c1 * wz;
            }
        }
    }
}

/* Autograd of models/encodings.py:331-381 w.r.t. hash_table: d table[row_c] += dout * w_c with
 * w_c the trilinear corner weight, accumulated like index_put_(accumulate=True).  Serial over
 * points (deterministic order) -- parallel over levels, which never collide. */
ORC_API void orc_hashgrid_bwd(const float* x01, int64_t P, const float* dout, int L, int F,
                              int log2T, const int32_t* res, int interp, float* dtable)
{
    const int64_t T = (int64_t)1 << log2T;
#pragma omp parallel for schedule(static)
    for (int l = 0; l < L; ++l) {
        float rl = (float)res[l];
        for (int64_t p = 0; p < P; ++p) {
            float sx = x01[3 * p + 0] * rl, sy = x01[3 * p + 1] * rl, sz = x01[3 * p + 2] * rl;
            const float* g = dout + ((size_t)p * L + l) * F;
            if (interp == 0) {
                int64_t ix = (int64_t)nearbyintf(sx), iy = (int64_t)nearbyintf(sy), iz = (int64_t)nearbyintf(sz);
                int64_t row = orc_hash(ix, iy, iz, log2T) + (int64_t)l * T;
                for (int f = 0; f < F; ++f) dtable[row * F + f] += g[f];
                continue;
            }
            float fx = floorf(sx), fy = floorf(sy), fz = floorf(sz);
            float wx = sx - fx, wy = sy - fy, wz = sz - fz;
            int64_t x0 = (int64_t)fx, y0 = (int64_t)fy, z0 = (int64_t)fz;
            if (interp == 2) {
                wx = wx * wx * (3.0f - 2.0f * wx);
                wy = wy * wy * (3.0f - 2.0f * wy);
                wz = wz * wz * (3.0f - 2.0f * wz);
            }
            for (int c = 0; c < 8; ++c) {
                int cx = (c >> 2) & 1, cy = (c >> 1) & 1, cz = c & 1;
                int64_t row = orc_hash(x0 + cx, y0 + cy, z0 + cz, log2T) + (int64_t)l * T;
                float w = (cz ? wz : 1.0f - wz);
                w = w * (cy ? wy : 1.0f - wy);
                w = w * (cx ? wx : 1.0f - wx);
                for (int f = 0; f < F; ++f) dtable[row * F + f] += g[f] * w;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Stage 3 -- field MLPs (density trunk + heads, SH, colour MLP)
 * ---------------------------------------------------------------------------------------- */

/* models/encodings.py:27-81 components_from_spherical_harmonics (degree 3, 16 comps) applied
 * after the double normalisation of models/inr/meta_ngp.py:166-169 and encodings.py:141. */
static inline void orc_normalize3(float* v, float eps)
{
    float n = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    if (n < eps) n = eps;
    v[0] = v[0] / n; v[1] = v[1] / n; v[2] = v[2] / n;
}

static void orc_sh16_poly(const float* v, float* sh)
{
    float x = v[0], y = v[1], z = v[2];
    float xx = x * x, yy = y * y, zz = z * z;
    sh[0] = 0.28209479177387814f;
    sh[1] = 0.4886025119029199f * y;
    sh[2] = 0.4886025119029199f * z;
    sh[3] = 0.4886025119029199f * x;
    sh[4] = 1.0925484305920792f * x * y;
    sh[5] = 1.0925484305920792f * y * z;
    sh[6] = 0.9461746957575601f * zz - 0.31539156525251999f;
    sh[7] = 1.0925484305920792f * x * z;
    sh[8] = 0.5462742152960396f * (xx - yy);
    sh[9] = 0.5900435899266435f * y * (3.0f * xx - yy);
    sh[10] = 2.890611442640554f * x * y * z;
    sh[11] = 0.4570457994644658f * y * (5.0f * zz - 1.0f);
    sh[12] = 0.3731763325901154f * z * (5.0f * zz - 3.0f);
    sh[13] = 0.4570457994644658f * x * (5.0f * zz - 1.0f);
    sh[14] = 1.445305721320277f * z * (xx - yy);
    sh[15] = 0.5900435899266435f * x * (xx - 3.0f * yy);
}

/* expert path: _enc_dir normalises (eps 1e-9), then SHEncoder.forward normalises again */
static void orc_sh16_one(const float* d_in, float* sh)
{
    float v[3] = { d_in[0], d_in[1], d_in[2] };
    orc_normalize3(v, 1e-9f);
    orc_normalize3(v, 1e-9f);
    orc_sh16_poly(v, sh);
}

ORC_API void orc_sh16(const float* d, int64_t P, float* out)
{
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < P; ++p) orc_sh16_one(d + 3 * p, out + 16 * p);
}

/* Weight pack for one expert, all fp32 row-major (out,in) like nn.Linear
 * (models/inr/meta_ngp.py:74-99): trunk0 (H,E) trunk1 (H,H) sigma (1,H) geo (G,H)
 * col0 (C,G+16) col1 (C,C) col2 (3,C), each followed by its bias. */
typedef struct {
    const float *w_t0, *b_t0, *w_t1, *b_t1, *w_sig, *b_sig, *w_geo, *b_geo;
    const float *w_c0, *b_c0, *w_c1, *b_c1, *w_c2, *b_c2;
} orc_field_weights;

#define ORC_MAXW 64

/* y = W x + b over `in` inputs; half=1 emulates torch autocast(fp16): inputs and weights are
 * rounded to fp16, products accumulate in fp32, the GEMM result is rounded to fp16, and the
 * fp32 bias is then added in fp32 (models/metamodule/metamodule.py:150-155 under autocast).
 * The loop nest is input-major so the compiler can vectorise across outputs WITHOUT changing
 * any output's summation order (i ascending), i.e. results are identical to the naive nest. */
static inline void orc_linear(const float* restrict W, const float* restrict b, const float* restrict x,
                              int in, int out, int half, float* restrict y)
{
    for (int o = 0; o < out; ++o) y[o] = 0.0f;
    if (half) {
        for (int i = 0; i < in; ++i) {
            float xi = f16r(x[i]);
            for (int o = 0; o < out; ++o) y[o] += f16r(W[o * in + i]) * xi;
        }
        for (int o = 0; o < out; ++o) y[o] = f16r(y[o]);
    } else {
        for (int i = 0; i < in; ++i) {
            float xi = x[i];
#pragma omp simd
            for (int o = 0; o < out; ++o) y[o] += W[o * in + i] * xi;
        }
    }
    for (int o = 0; o < out; ++o) y[o] = y[o] + b[o];
}

/* models/trunc_exp.py:32-61: exp(clamp(x, -88.722839111, 88.722839111)). */
static inline float orc_trunc_exp(float x) { return expf(clampf(x, -88.722839111f, 88.722839111f)); }

/* models/inr/meta_ngp.py:171-241 MetaNGP.forward on already-encoded positions:
 * enc (P,E), dirs (P,3) -> rgb_sigma (P,4).  act (optional) receives per point
 * [h1(H) h2(H) sig_raw(1) geo(G) sh(16) c1(C) c2(C) rgb_raw(3)] for the backward. */
ORC_API void orc_field_fwd(const float* enc, const float* dirs, int64_t P, int E, int H, int G,
                           int C, const orc_field_weights* w, int half, float* rgb_sigma,
                           float* act)
{
    const int stride = 2 * H + 1 + G + 16 + 2 * C + 3;
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < P; ++p) {
        float h1[ORC_MAXW], h2[ORC_MAXW], sg[1], cin[ORC_MAXW], c1[ORC_MAXW], c2[ORC_MAXW], rr[3];
        orc_linear(w->w_t0, w->b_t0, enc + (size_t)p * E, E, H, half, h1);
        for (int i = 0; i < H; ++i) h1[i] = h1[i] > 0.0f ? h1[i] : 0.0f;
        orc_linear(w->w_t1, w->b_t1, h1, H, H, half, h2);
        for (int i = 0; i < H; ++i) h2[i] = h2[i] > 0.0f ? h2[i] : 0.0f;
        orc_linear(w->w_sig, w->b_sig, h2, H, 1, half, sg);
        orc_linear(w->w_geo, w->b_geo, h2, H, G, half, cin);
        orc_sh16_one(dirs + 3 * p, cin + G);
        orc_linear(w->w_c0, w->b_c0, cin, G + 16, C, half, c1);
        for (int i = 0; i < C; ++i) c1[i] = c1[i] > 0.0f ? c1[i] : 0.0f;
        orc_linear(w->w_c1, w->b_c1, c1, C, C, half, c2);
        for (int i = 0; i < C; ++i) c2[i] = c2[i] > 0.0f ? c2[i] : 0.0f;
        orc_linear(w->w_c2, w->b_c2, c2, C, 3, half, rr);
        float* o = rgb_sigma + 4 * p;
        for (int k = 0; k < 3; ++k) o[k] = 1.0f / (1.0f + expf(-rr[k]));
        o[3] = orc_trunc_exp(sg[0]);
        if (act) {
            float* a = act + (size_t)p * stride;
            memcpy(a, h1, sizeof(float) * H); a += H;
            memcpy(a, h2, sizeof(float) * H); a += H;
            *a++ = sg[0];
            memcpy(a, cin, sizeof(float) * (G + 16)); a += G + 16;
            memcpy(a, c1, sizeof(float) * C); a += C;
            memcpy(a, c2, sizeof(float) * C); a += C;
            memcpy(a, rr, sizeof(float) * 3);
        }
    }
}

/* Gradient pack, same shapes as orc_field_weights (accumulated into). */
typedef struct {
    float *w_t0, *b_t0, *w_t1, *b_t1, *w_sig, *b_sig, *w_geo, *b_geo;
    float *w_c0, *b_c0, *w_c1, *b_c1, *w_c2, *b_c2;
} orc_field_grads;

/* Backward of orc_field_fwd (fp32 math): d rgb_sigma (P,4) -> weight grads (+=) and d enc (P,E).
 * sigmoid' = y(1-y); trunc_exp' = exp(clamped x) (models/trunc_exp.py:52-57); ReLU' = (h>0). */
static void orc_field_bwd_range(const float* enc, const float* dirs, int64_t p0, int64_t p1, int E, int H, int G,
                                int C, const orc_field_weights* w, const float* d_rgb_sigma,
                                orc_field_grads* g, float* d_enc)
{
    const int stride = 2 * H + 1 + G + 16 + 2 * C + 3;
    float* act = (float*)malloc(sizeof(float) * (size_t)stride);
    float rs[4];
    for (int64_t p = p0; p < p1; ++p) {
        orc_field_fwd(enc + (size_t)p * E, dirs + 3 * p, 1, E, H, G, C, w, 0, rs, act);
        const float *h1 = act, *h2 = act + H, *cin = act + 2 * H + 1, *c1 = cin + G + 16, *c2 = c1 + C;
        const float* x = enc + (size_t)p * E;
        const float* dy = d_rgb_sigma + 4 * p;
        float d_rr[3], d_c2[ORC_MAXW], d_c1[ORC_MAXW], d_cin[ORC_MAXW], d_h2[ORC_MAXW], d_h1[ORC_MAXW];
        for (int k = 0; k < 3; ++k) d_rr[k] = dy[k] * rs[k] * (1.0f - rs[k]);
        float d_sg = dy[3] * rs[3];
        /* colour head */
        for (int i = 0; i < C; ++i) d_c2[i] = 0.0f;
        for (int o = 0; o < 3; ++o) {
            g->b_c2[o] += d_rr[o];
            _Pragma("omp simd")
            for (int i = 0; i < C; ++i) { g->w_c2[o * C + i] += d_rr[o] * c2[i]; d_c2[i] += d_rr[o] * w->w_c2[o * C + i]; }
        }
        for (int i = 0; i < C; ++i) { if (!(c2[i] > 0.0f)) d_c2[i] = 0.0f; d_c1[i] = 0.0f; }
        for (int o = 0; o < C; ++o) {
            g->b_c1[o] += d_c2[o];
            _Pragma("omp simd")
            for (int i = 0; i < C; ++i) { g->w_c1[o * C + i] += d_c2[o] * c1[i]; d_c1[i] += d_c2[o] * w->w_c1[o * C + i]; }
        }
        for (int i = 0; i < C; ++i) if (!(c1[i] > 0.0f)) d_c1[i] = 0.0f;
        for (int i = 0; i < G + 16; ++i) d_cin[i] = 0.0f;
        for (int o = 0; o < C; ++o) {
            g->b_c0[o] += d_c1[o];
            _Pragma("omp simd")
            for (int i = 0; i < G + 16; ++i) { g->w_c0[o * (G + 16) + i] += d_c1[o] * cin[i]; d_cin[i] += d_c1[o] * w->w_c0[o * (G + 16) + i]; }
        }
        /* heads */
        for (int i = 0; i < H; ++i) d_h2[i] = 0.0f;
        g->b_sig[0] += d_sg;
        for (int i = 0; i < H; ++i) { g->w_sig[i] += d_sg * h2[i]; d_h2[i] += d_sg * w->w_sig[i]; }
        for (int o = 0; o < G; ++o) {
            g->b_geo[o] += d_cin[o];
            _Pragma("omp simd")
            for (int i = 0; i < H; ++i) { g->w_geo[o * H + i] += d_cin[o] * h2[i]; d_h2[i] += d_cin[o] * w->w_geo[o * H + i]; }
        }
        /* trunk */
        for (int i = 0; i < H; ++i) { if (!(h2[i] > 0.0f)) d_h2[i] = 0.0f; d_h1[i] = 0.0f; }
        for (int o = 0; o < H; ++o) {
            g->b_t1[o] += d_h2[o];
            _Pragma("omp simd")
            for (int i = 0; i < H; ++i) { g->w_t1[o * H + i] += d_h2[o] * h1[i]; d_h1[i] += d_h2[o] * w->w_t1[o * H + i]; }
        }
        for (int i = 0; i < H; ++i) if (!(h1[i] > 0.0f)) d_h1[i] = 0.0f;
        float* dx = d_enc ? d_enc + (size_t)p * E : NULL;
        if (dx) for (int i = 0; i < E; ++i) dx[i] = 0.0f;
        for (int o = 0; o < H; ++o) {
            g->b_t0[o] += d_h1[o];
            for (int i = 0; i < E; ++i) { g->w_t0[o * E + i] += d_h1[o] * x[i]; if (dx) dx[i] += d_h1[o] * w->w_t0[o * E + i]; }
        }
    }
    free(act);
}

/* Points are split over OpenMP threads with private gradient buffers (used by the CPU baseline). */
ORC_API void orc_field_bwd(const float* enc, const float* dirs, int64_t P, int E, int H, int G,
                           int C, const orc_field_weights* w, const float* d_rgb_sigma,
                           orc_field_grads* g, float* d_enc)
{
    /* sizes of the 14 tensors, in struct order */
    const size_t sz[14] = { (size_t)H * E, (size_t)H, (size_t)H * H, (size_t)H, (size_t)H, 1, (size_t)G * H, (size_t)G,
                            (size_t)C * (G + 16), (size_t)C, (size_t)C * C, (size_t)C, (size_t)3 * C, 3 };
    size_t total = 0;
    for (int i = 0; i < 14; ++i) total += sz[i];
    float** gp = (float**)g;
#pragma omp parallel
    {
        int nt = omp_get_num_threads(), id = omp_get_thread_num();
        int64_t p0 = P * id / nt, p1 = P * (id + 1) / nt;
        float* buf = (float*)calloc(total, sizeof(float));
        orc_field_grads loc;
        float** lp = (float**)&loc;
        size_t off = 0;
        for (int i = 0; i < 14; ++i) { lp[i] = buf + off; off += sz[i]; }
        orc_field_bwd_range(enc, dirs, p0, p1, E, H, G, C, w, d_rgb_sigma, &loc, d_enc);
#pragma omp critical
        {
            for (int i = 0; i < 14; ++i) for (size_t j = 0; j < sz[i]; ++j) gp[i][j] += lp[i][j];
        }
        free(buf);
    }
}

/* models/inr/meta_container.py:347-382 background_color: SH16(normalize(d)) -> Linear(16,Hb)
 * ReLU -> Linear(Hb,3) -> sigmoid, per ray. */
ORC_API void orc_background(const float* dirs, int64_t N, int Hb, const float* w0, const float* b0,
                            const float* w1, const float* b1, float* rgb)
{
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < N; ++r) {
        float sh[16], h[ORC_MAXW], o[3];
        /* F.normalize (eps 1e-12) then the SH encoder's own normalisation (eps 1e-9) */
        float v[3] = { dirs[3 * r], dirs[3 * r + 1], dirs[3 * r + 2] };
        orc_normalize3(v, 1e-12f);
        orc_normalize3(v, 1e-9f);
        orc_sh16_poly(v, sh);
        orc_linear(w0, b0, sh, 16, Hb, 0, h);
        for (int i = 0; i < Hb; ++i) h[i] = h[i] > 0.0f ? h[i] : 0.0f;
        orc_linear(w1, b1, h, Hb, 3, 0, o);
        for (int k = 0; k < 3; ++k) rgb[3 * r + k] = 1.0f / (1.0f + expf(-o[k]));
    }
}

/* ------------------------------------------------------------------------------------------
 * Stage 4 -- alpha compositing
 * ---------------------------------------------------------------------------------------- */

/* nerfs/ray_rendering.py:114-165 volume_render (raw_rgb = raw_sigma = False). */
ORC_API void orc_composite_fwd(const float* rgb_sigma, const float* t_vals, const float* bg_or_null,
                               int64_t N, int S, float sigma_scale, float* rgb_map,
                               float* depth_map, float* weights, float* acc_map)
{
    const float amax = (float)(1.0 - 1e-7);
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < N; ++r) {
        const float* rs = rgb_sigma + (size_t)r * S * 4;
        const float* t = t_vals + (size_t)r * S;
        float T = 1.0f, cr = 0.0f, cg = 0.0f, cb = 0.0f, dep = 0.0f, acc = 0.0f;
        for (int s = 0; s < S; ++s) {
            float sig = rs[4 * s + 3]; sig = sig > 0.0f ? sig : (sig != sig ? sig : 0.0f);
            if (sigma_scale != 1.0f) sig = sig * sigma_scale;
            float dl;
            if (S == 1) dl = 0.0f; /* degenerate: reference would fail on empty dists */
            else if (s < S - 1) { dl = t[s + 1] - t[s]; if (dl < 1e-4f) dl = 1e-4f; }
            else { dl = t[S - 1] - t[S - 2]; if (dl < 1e-4f) dl = 1e-4f; }
            float a = 1.0f - expf(-sig * dl);
            a = clampf(a, 0.0f, amax);
            float w = a * T;
            weights[(size_t)r * S + s] = w;
            cr += w * clampf(rs[4 * s + 0], 0.0f, 1.0f);
            cg += w * clampf(rs[4 * s + 1], 0.0f, 1.0f);
            cb += w * clampf(rs[4 * s + 2], 0.0f, 1.0f);
            dep += w * t[s];
            acc += w;
            T = T * ((1.0f - a) + 1e-10f);
        }
        if (bg_or_null) {
            float rem = 1.0f - acc;
            cr = cr + rem * bg_or_null[3 * r + 0];
            cg = cg + rem * bg_or_null[3 * r + 1];
            cb = cb + rem * bg_or_null[3 * r + 2];
        }
        rgb_map[3 * r + 0] = cr; rgb_map[3 * r + 1] = cg; rgb_map[3 * r + 2] = cb;
        depth_map[r] = dep; acc_map[r] = acc;
    }
}

/* Autograd of volume_render w.r.t. rgb_sigma (N,S,4) and bg (N,3).  Incoming grads for the
 * four outputs; any may be NULL (= zero).  d_bg may be NULL. */
ORC_API void orc_composite_bwd(const float* rgb_sigma, const float* t_vals, const float* bg_or_null,
                               int64_t N, int S, float sigma_scale, const float* g_rgb,
                               const float* g_depth, const float* g_weights, const float* g_acc,
                               float* d_rgb_sigma, float* d_bg)
{
    const float amax = (float)(1.0 - 1e-7);
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < N; ++r) {
        const float* rs = rgb_sigma + (size_t)r * S * 4;
        const float* t = t_vals + (size_t)r * S;
        float* d = d_rgb_sigma + (size_t)r * S * 4;
        float gr[3] = { 0, 0, 0 };
        if (g_rgb) { gr[0] = g_rgb[3 * r]; gr[1] = g_rgb[3 * r + 1]; gr[2] = g_rgb[3 * r + 2]; }
        float gd = g_depth ? g_depth[r] : 0.0f;
        float ga = g_acc ? g_acc[r] : 0.0f;
        float acc = 0.0f;
        if (bg_or_null) ga -= gr[0] * bg_or_null[3 * r] + gr[1] * bg_or_null[3 * r + 1] + gr[2] * bg_or_null[3 * r + 2];
        /* forward recompute, storing alpha, T, q */
        float* al = (float*)malloc(sizeof(float) * 4 * (size_t)S);
        float *Tt = al + S, *q = al + 2 * S, *dl = al + 3 * S;
        float T = 1.0f;
        for (int s = 0; s < S; ++s) {
            float sig = rs[4 * s + 3]; sig = sig > 0.0f ? sig : 0.0f;
            sig = sig * sigma_scale;
            float dd = (s < S - 1) ? t[s + 1] - t[s] : t[S - 1] - t[S - 2];
            if (dd < 1e-4f) dd = 1e-4f;
            dl[s] = dd;
            float a = clampf(1.0f - expf(-sig * dd), 0.0f, amax);
            al[s] = a; Tt[s] = T; q[s] = (1.0f - a) + 1e-10f;
            acc += a * T;
            T = T * q[s];
        }
        if (d_bg && bg_or_null) for (int k = 0; k < 3; ++k) d_bg[3 * r + k] = (1.0f - acc) * gr[k];
        /* suffix sum of gw_i * w_i for the transmittance chain */
        float suffix = 0.0f;
        for (int s = S - 1; s >= 0; --s) {
            float cr = clampf(rs[4 * s], 0.0f, 1.0f), cg = clampf(rs[4 * s + 1], 0.0f, 1.0f), cb = clampf(rs[4 * s + 2], 0.0f, 1.0f);
            float w = al[s] * Tt[s];
            float gw = gr[0] * cr + gr[1] * cg + gr[2] * cb + gd * t[s] + ga + (g_weights ? g_weights[(size_t)r * S + s] : 0.0f);
            /* d rgb: clamp passes gradient on [0,1] inclusive */
            for (int k = 0; k < 3; ++k) {
                float v = rs[4 * s + k];
                d[4 * s + k] = (v >= 0.0f && v <= 1.0f) ? w * gr[k] : 0.0f;
            }
            /* d alpha = gw*T - (sum_{i>s} gw_i w_i)/q_s */
            float da = gw * Tt[s] - suffix / q[s];
            suffix += gw * w;
            /* alpha clamp passes on [0, amax]; alpha = 1 - exp(-sig*dl) */
            float sig = rs[4 * s + 3]; float sigc = sig > 0.0f ? sig : 0.0f;
            float e = expf(-(sigc * sigma_scale) * dl[s]);
            float araw = 1.0f - e;
            float ds = (araw >= 0.0f && araw <= amax) ? da * dl[s] * e : 0.0f;
            ds = ds * sigma_scale;
            d[4 * s + 3] = (sig >= 0.0f) ? ds : 0.0f;
        }
        free(al);
    }
}

/* ------------------------------------------------------------------------------------------
 * Stage 5 -- Voronoi routing
 * ---------------------------------------------------------------------------------------- */

/* torch.cdist(x, c) on CPU for > 25 rows takes the matmul form; bit-exact recipe for 2-D
 * (SURVEY 7.1): xn = fl(x0*x0)+fl(x1*x1); acc = fl(-2x0*c0); acc = fma(-2x1,c1,acc);
 * acc = fma(xn,1,acc); acc = fma(1,cn,acc); dist = sqrt(max(acc,0)).  The 3-D variant follows
 * the same pattern (tolerance-only: MKL blocks K=5 differently). */
static inline float orc_cdist_mm(const float* x, const float* c, int dim)
{
    float xn = x[0] * x[0], cn = c[0] * c[0];
    for (int k = 1; k < dim; ++k) { xn = xn + x[k] * x[k]; cn = cn + c[k] * c[k]; }
    float acc = (-2.0f * x[0]) * c[0];
    for (int k = 1; k < dim; ++k) acc = fmaf(-2.0f * x[k], c[k], acc);
    acc = fmaf(xn, 1.0f, acc);
    acc = fmaf(1.0f, cn, acc);
    return sqrtf(acc > 0.0f ? acc : 0.0f);
}

/* models/inr/meta_container.py:97-134 _routing.  dims = 2 -> columns (1,2) (cluster_2d) else
 * (0,1,2).  margin > 1: soft weights (P,K); else hard argmin (first minimum) (P,) int32. */
ORC_API void orc_route_points(const float* pts, int64_t P, int stride, const float* centroids,
                              int K, int dims, float margin, float* weights, int32_t* hard)
{
    const int off = dims == 2 ? 1 : 0;
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < P; ++p) {
        float dist[64];
        for (int k = 0; k < K; ++k) dist[k] = orc_cdist_mm(pts + (size_t)p * stride + off, centroids + 3 * k + off, dims);
        if (margin > 1.0f) {
            float mind = INFINITY, denom = 0.0f;
            for (int k = 0; k < K; ++k) { if (dist[k] < 1e-6f) dist[k] = 1e-6f; if (dist[k] < mind) mind = dist[k]; }
            float thr = margin * mind;
            float inv[64];
            for (int k = 0; k < K; ++k) { inv[k] = (1.0f / dist[k]) * (dist[k] <= thr ? 1.0f : 0.0f); denom = denom + inv[k]; }
            if (denom < 1e-6f) denom = 1e-6f;
            for (int k = 0; k < K; ++k) weights[(size_t)p * K + k] = inv[k] / denom;
        } else {
            int best = 0;
            for (int k = 1; k < K; ++k) if (dist[k] < dist[best]) best = k;
            hard[p] = best;
        }
    }
}

/* scripts/create_clusters.py:559-634 compute_voronoi_orig: S uniform samples on [near,far]
 * (torch.lerp -- fused), ray in expert c iff min_s D(x_s,c)/(min_c' D + 1e-8) <= margin.
 * u_lin = linspace(0,1,S).  mask (N,K) uint8. */
ORC_API void orc_route_rays_voronoi(const float* rays8, int64_t N, int S, const float* u_lin,
                                    const float* centroids, int K, int dims, float margin,
                                    uint8_t* mask)
{
    const int off = dims == 2 ? 1 : 0;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t r = 0; r < N; ++r) {
        const float* ry = rays8 + 8 * r;
        float near = ry[6], far = ry[7];
        float diff = far - near;
        float rmin[64];
        for (int k = 0; k < K; ++k) rmin[k] = INFINITY;
        for (int s = 0; s < S; ++s) {
            float z = u_lin[s];
            float t = z < 0.5f ? fmaf(z, diff, near) : fmaf(-diff, 1.0f - z, far);
            float x[3];
            for (int a = 0; a < 3; ++a) { float m = ry[3 + a] * t; x[a] = ry[a] + m; }
            float D[64], m = INFINITY;
            for (int k = 0; k < K; ++k) { D[k] = orc_cdist_mm(x + off, centroids + 3 * k + off, dims); if (D[k] < m) m = D[k]; }
            float den = m + 1e-8f;
            for (int k = 0; k < K; ++k) { float ratio = D[k] / den; if (ratio < rmin[k]) rmin[k] = ratio; }
        }
        for (int k = 0; k < K; ++k) mask[(size_t)r * K + k] = (uint8_t)(rmin[k] <= margin);
    }
}

/* models/inr/meta_container.py:275-343 MetaContainer.forward blend: y = sum_k w_k * y_k over
 * experts with w_k > 0 (soft) or y = y_{hard} (hard).  y_all (K,P,4) precomputed per expert. */
ORC_API void orc_blend(const float* y_all, int64_t P, int K, const float* weights_or_null,
                       const int32_t* hard_or_null, float* out)
{
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < P; ++p)
        for (int c = 0; c < 4; ++c) {
            float acc = 0.0f;
            if (weights_or_null) {
                for (int k = 0; k < K; ++k) {
                    float w = weights_or_null[(size_t)p * K + k];
                    if (w > 0.0f) acc = acc + y_all[((size_t)k * P + p) * 4 + c] * w;
                }
            } else acc = y_all[((size_t)hard_or_null[p] * P + p) * 4 + c];
            out[4 * p + c] = acc;
        }
}

/* ------------------------------------------------------------------------------------------
 * Around the render: loss epilogue and optimizer tail
 * ---------------------------------------------------------------------------------------- */

static inline float orc_clamp01_nan(float x) { return x != x ? x : clampf(x, 0.0f, 1.0f); }

/* nerfs/color_space.py:13-19 srgb_to_linear.  torch evaluates tensor / python-scalar as a multiplication with the
 * fp32-rounded reciprocal. */
static inline float orc_srgb_to_linear(float x)
{
    return x <= 0.04045f ? x * (1.0f / 12.92f) : powf((x + 0.055f) * (1.0f / 1.055f), 2.4f);
}

/* nerfs/color_space.py:22-66 color_space_transformer (cs: 0 linear, 1 srgb, 2 identity) followed by
 * nerfs/losses.py:32 F.mse_loss.  elem (n): squared errors; dpred (n): d(sum of squared errors)/d pred, i.e. NOT yet
 * divided by n; returns the sum of the squared errors (double accumulation).  At pred == 0 in sRGB mode autograd
 * differentiates the unselected pow branch of torch.where (color_space.py:6-10) and returns NaN; this restatement
 * (like the CUDA kernel) reports the selected linear branch's slope there -- the one documented deviation. */
ORC_API double orc_color_mse(const float* pred, const float* gt, int64_t n, int cs, float* elem, float* dpred)
{
    double sum = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        const float p = pred[i], g01 = orc_clamp01_nan(gt[i]);
        float pp, gg, dpp;
        if (cs == 0) {
            pp = orc_clamp01_nan(p);
            dpp = (p >= 0.0f && p <= 1.0f) ? 1.0f : 0.0f;
            gg = orc_clamp01_nan(orc_srgb_to_linear(g01));
        } else if (cs == 1) {
            const float x = orc_clamp01_nan(p), e = (float)(1.0 / 2.4);
            const int lin = x <= 0.0031308f;
            const float y = lin ? 12.92f * x : 1.055f * powf(x, e) - 0.055f;
            const float dy = lin ? 12.92f : 1.055f * e * powf(x, e - 1.0f);
            pp = orc_clamp01_nan(y);
            dpp = (p >= 0.0f && p <= 1.0f && y >= 0.0f && y <= 1.0f) ? dy : 0.0f;
            gg = g01;
        } else {
            pp = p; dpp = 1.0f; gg = g01;
        }
        const float d = pp - gg, se = d * d;
        if (elem) elem[i] = se;
        if (dpred) dpred[i] = 2.0f * d * dpp;
        sum += (double)se;
    }
    return sum;
}

/* One optimizer step as pipelines/offline_stage/meta_core.py:123-141 maml_meta_update runs it:
 *   scaler.unscale_  (g *= 1/scale, any non-finite g -> skip the step),
 *   clip_all_grads   (:181-190 -> torch clip_grad_norm_: coef = min(1, max_norm / (||g||_2 + 1e-6)), g *= coef),
 *   optimizer.step   (torch.optim.Adam, or AdamW when adamw != 0; common/utils.py:16-76 picks them).
 * T tensors p/g/m/v of sizes n[], per-tensor lr/wd (the parameter groups).  step_io: number of steps taken so far
 * (incremented unless skipped).  g is overwritten with the unscaled, clipped gradient.  Returns the total norm;
 * *skipped is set when a gradient was not finite.  Constants are rounded from doubles exactly where torch rounds. */
ORC_API float orc_adam_step(int T, float** p, float** g, float** m, float** v, const int64_t* n, const double* lr,
                            const double* wd, double beta1, double beta2, double eps, int adamw, float grad_scale,
                            float max_norm, double* step_io, int* skipped)
{
    const float inv = grad_scale > 0.0f ? 1.0f / grad_scale : 1.0f;
    int bad = 0;
    double sq = 0.0;
    for (int t = 0; t < T; ++t)
        for (int64_t i = 0; i < n[t]; ++i) {
            g[t][i] = g[t][i] * inv;
            if (!isfinite(g[t][i])) bad = 1;
            sq += (double)g[t][i] * (double)g[t][i];
        }
    const float norm = (float)sqrt(sq);
    *skipped = bad;
    if (bad) return norm;
    if (max_norm > 0.0f) {
        float coef = max_norm / (norm + 1e-6f);
        if (coef > 1.0f) coef = 1.0f;
        for (int t = 0; t < T; ++t)
            for (int64_t i = 0; i < n[t]; ++i) g[t][i] = g[t][i] * coef;
    }
    *step_io += 1.0;
    const double bc1 = 1.0 - pow(beta1, *step_io), bc2 = 1.0 - pow(beta2, *step_io);
    const float omb1 = (float)(1.0 - beta1), b2 = (float)beta2, omb2 = (float)(1.0 - beta2);
    const float bc2s = (float)sqrt(bc2), epsf = (float)eps;
    for (int t = 0; t < T; ++t) {
        const float step_size = (float)(lr[t] / bc1), wdf = (float)wd[t], lr_wd = (float)(lr[t] * wd[t]);
        for (int64_t i = 0; i < n[t]; ++i) {
            float gi = g[t][i], pi = p[t][i];
            if (wdf != 0.0f) {
                if (adamw) pi = pi - lr_wd * pi;
                else gi = fmaf(wdf, pi, gi);
            }
            const float mi = m[t][i] + omb1 * (gi - m[t][i]);
            const float vi = b2 * v[t][i] + omb2 * gi * gi;
            const float denom = sqrtf(vi) / bc2s + epsf;
            p[t][i] = pi - step_size * (mi / denom);
            m[t][i] = mi;
            v[t][i] = vi;
        }
    }
    return norm;
}

/* ------------------------------------------------------------------------------------------
 * Ray-batch producer: DDA max-overlap binning of rays into the task grid
 * ---------------------------------------------------------------------------------------- */

/* torch.minimum / torch.maximum / Tensor.max(dim) propagate NaN; fminf / fmaxf do not. */
static inline float orc_min_nan(float a, float b) { return (a != a || b != b) ? NAN : (a < b ? a : b); }
static inline float orc_max_nan(float a, float b) { return (a != a || b != b) ? NAN : (a > b ? a : b); }
static inline float orc_nan_to_big(float x) { return (x != x || isinf(x)) ? 1e30f : x; }  /* nan_to_num_(nan=posinf=neginf=1e30) */

/* data/task_dataset.py:125-152 _aabb_intersect (eps = 1e-12) for one ray and one box. */
static inline int orc_slab(const float* o, const float* d, const float* lo, const float* hi, float* t_entry, float* t_exit)
{
    float te = -INFINITY, tx = INFINITY;
    int miss_parallel = 0, first = 1;
    for (int a = 0; a < 3; ++a) {
        const int parallel = fabsf(d[a]) < 1e-12f;
        const float inv_d = 1.0f / d[a];
        const float t0 = (lo[a] - o[a]) * inv_d, t1 = (hi[a] - o[a]) * inv_d;
        const float tmin = orc_min_nan(t0, t1), tmax = orc_max_nan(t0, t1);
        if (parallel && !(o[a] >= lo[a] && o[a] <= hi[a])) miss_parallel = 1;
        te = first ? tmin : orc_max_nan(te, tmin);
        tx = first ? tmax : orc_min_nan(tx, tmax);
        first = 0;
    }
    *t_entry = te; *t_exit = tx;
    return (tx >= te) && !miss_parallel;
}

/* The "dda" routing policy of data/task_dataset.py:544-627 _route_and_bin (nerf_runner.py:207 selects it), per ray:
 *   :155-172 _region_segment        clip to the region box and (near, far); invalid rays get cell -1
 *   :239-297 _dda_transform/_dda_init  grid coordinates, first voxel at t0 + 1e-6, tMax / tDelta with nan/inf -> 1e30
 *   :299-351 _dda_maxoverlap        up to max_steps voxel steps, the cell with the longest in-cell parametric length wins
 *                                   (tMax is measured from the entry point but compared with absolute t: kept as is)
 *   :212-228, 590-603               overlap of the ray with the chosen cell must reach tol[cell], else the ray is dropped
 * aabb6 = [lo, hi]; cell3 = clamp((hi-lo)/cells, 1e-12) and cell_bounds (C,2,3), tol (C) are the tensors the reference
 * builds on the host (:174-197, :595-597).  cid_out (N): chosen cell or -1; counts (C) are ADDED to. */
ORC_API void orc_dda_route_rays(const float* rays8, int64_t N, const float* aabb6, const int* cells, const float* cell3,
                                const float* cell_bounds, const float* tol, int max_steps, int32_t* cid_out,
                                float* best_len_out, int64_t* counts)
{
    const int nx = cells[0], ny = cells[1], nz = cells[2], nyz = ny * nz;
    const float* lo = aabb6;
    const float* hi = aabb6 + 3;
    for (int64_t r = 0; r < N; ++r) {
        const float* ray = rays8 + 8 * r;
        const float* o = ray;
        const float* d = ray + 3;
        const float near = ray[6], far = ray[7];
        cid_out[r] = -1;
        if (best_len_out) best_len_out[r] = 0.0f;
        float te, tx;
        const int hit = orc_slab(o, d, lo, hi, &te, &tx);
        float t0 = orc_max_nan(orc_max_nan(te, 0.0f), near);
        float t1 = orc_min_nan(tx, far);
        const float seg = t1 - t0;
        if (!(hit && seg > 0.0f)) continue;
        /* grid transform + init */
        float p[3], gd[3], tmax[3], tdelta[3];
        int64_t idx[3], step[3];
        const int64_t ncell[3] = { nx, ny, nz };
        const float tstart = t0 + 1e-6f;
        for (int a = 0; a < 3; ++a) {
            const float go = (o[a] - lo[a]) / cell3[a];
            gd[a] = d[a] / cell3[a];
            p[a] = go + gd[a] * tstart;
            idx[a] = (int64_t)floorf(p[a]);
            step[a] = gd[a] > 0.0f ? 1 : (gd[a] < 0.0f ? -1 : 0);
            const float nb = step[a] > 0 ? floorf(p[a]) + 1.0f : ceilf(p[a]) - 1.0f;
            const float inv = 1.0f / gd[a];
            tmax[a] = orc_nan_to_big((nb - p[a]) * inv);
            tdelta[a] = orc_nan_to_big((float)step[a] * inv);
            if (idx[a] < 0) idx[a] = 0;
            if (idx[a] > ncell[a] - 1) idx[a] = ncell[a] - 1;
        }
        float t = t0, best_len = 0.0f;
        int64_t best_cid = idx[0] * nyz + idx[1] * nz + idx[2];
        for (int it = 0; it < max_steps; ++it) {
            const float m = orc_min_nan(orc_min_nan(tmax[0], tmax[1]), tmax[2]);
            const float t_next = orc_min_nan(m, t1);
            float dt = t_next - t;
            if (dt < 0.0f) dt = 0.0f;
            const int64_t cid = idx[0] * nyz + idx[1] * nz + idx[2];
            if (dt > best_len) { best_len = dt; best_cid = cid; }
            if (t_next >= t1) break;
            const int adv_x = (tmax[0] <= tmax[1]) && (tmax[0] <= tmax[2]);
            const int adv_y = !(tmax[0] <= tmax[1]) && (tmax[1] <= tmax[2]);
            const int a = adv_x ? 0 : (adv_y ? 1 : 2);
            idx[a] += step[a];
            if (idx[a] < 0) idx[a] = 0;
            if (idx[a] > ncell[a] - 1) idx[a] = ncell[a] - 1;
            tmax[a] = tmax[a] + tdelta[a];
            t = t_next;
        }
        /* overlap with the chosen cell */
        const float* cb = cell_bounds + 6 * best_cid;
        float ce, cx;
        const int chit = orc_slab(o, d, cb, cb + 3, &ce, &cx);
        const float c0 = orc_max_nan(orc_max_nan(ce, 0.0f), near), c1 = orc_min_nan(cx, far);
        float len = c1 - c0;
        if (len < 0.0f) len = 0.0f;                 /* clamp_min(0.0) keeps NaN; NaN >= tol is false either way */
        if (!chit) len = 0.0f;
        if (best_len_out) best_len_out[r] = best_len;
        if (len >= tol[best_cid]) {
            cid_out[r] = (int32_t)best_cid;
            if (counts) counts[best_cid] += 1;
        }
    }
}
