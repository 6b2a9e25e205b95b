// C ABI of stage 3: precision dispatch (fp32 SIMT / fp16 tcgen05) and the SH helper.
#include "field_common.cuh"
#include "field_internal.cuh"

__global__ void k_sh16(const float* __restrict__ dirs, int64_t P, int stride, float* __restrict__ out) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    float sh[16];
    sh16_expert(dirs[p * stride], dirs[p * stride + 1], dirs[p * stride + 2], sh);
    float4* o = reinterpret_cast<float4*>(out + 16 * p);
#pragma unroll
    for (int q = 0; q < 4; ++q) o[q] = make_float4(sh[4 * q], sh[4 * q + 1], sh[4 * q + 2], sh[4 * q + 3]);
}

extern "C" int acn_sh16(acn_ctx* ctx, const float* dirs, int64_t P, int stride, float* out, acn_stream stream) {
    ACN_CHECK_CTX(ctx);
    ACN_REQUIRE(P >= 0 && stride >= 3, ACN_EINVAL, "acn_sh16: bad arguments");
    if (P == 0) return ACN_OK;
    ACN_REQUIRE(dirs && out, ACN_EINVAL, "acn_sh16: null buffer");
    k_sh16<<<acn_grid_1d(P, 256), 256, 0, (cudaStream_t)stream>>>(dirs, P, stride, out);
    ACN_CHECK_LAUNCH();
    return ACN_OK;
}

static int check_field(const char* fn, const void* enc, int enc_dtype, const float* dirs, int dirs_stride, int dirs_group,
                       int64_t P, const acn_field_weights* w, int precision) {
    ACN_REQUIRE(P >= 0, ACN_EINVAL, "%s: negative P", fn);
    ACN_REQUIRE(enc_dtype == ACN_F32 || enc_dtype == ACN_F16, ACN_EINVAL, "%s: bad enc dtype", fn);
    ACN_REQUIRE(precision == ACN_F32 || precision == ACN_F16, ACN_EINVAL, "%s: bad precision", fn);
    ACN_REQUIRE(dirs_stride >= 3 && dirs_group >= 1, ACN_EINVAL, "%s: bad dirs stride/group", fn);
    ACN_REQUIRE(w != nullptr, ACN_EINVAL, "%s: null weights", fn);
    for (int i = 0; i < 14; ++i) ACN_REQUIRE(w->p[i] != nullptr, ACN_EINVAL, "%s: weight pointer %d is null", fn, i);
    if (P > 0) ACN_REQUIRE(enc && dirs, ACN_EINVAL, "%s: null enc/dirs", fn);
    return ACN_OK;
}

extern "C" int acn_field_fwd(acn_ctx* ctx, const void* enc, int enc_dtype, const float* dirs, int dirs_stride, int dirs_group,
                             int64_t P, const int32_t* range_or_null, int E, int H, int G, int C, const acn_field_weights* w,
                             int precision, float* rgb_sigma, acn_stream stream) {
    ACN_CHECK_CTX(ctx);
    int rc = check_field("acn_field_fwd", enc, enc_dtype, dirs, dirs_stride, dirs_group, P, w, precision);
    if (rc) return rc;
    if (P == 0) return ACN_OK;
    ACN_REQUIRE(rgb_sigma && ((uintptr_t)rgb_sigma & 15) == 0, ACN_EINVAL, "acn_field_fwd: rgb_sigma null or misaligned");
    ACN_REQUIRE(!range_or_null || dirs_group == 1, ACN_EINVAL, "acn_field_fwd: a row range needs per-point directions");
    cudaStream_t st = (cudaStream_t)stream;
    if (precision == ACN_F16)
        return acn_field_fwd_tc(ctx, enc, enc_dtype, dirs, dirs_stride, dirs_group, P, E, H, G, C, w, rgb_sigma, range_or_null, st);
    return acn_field_fwd_fp32(ctx, enc, enc_dtype, dirs, dirs_stride, dirs_group, P, E, H, G, C, w, rgb_sigma, range_or_null, st);
}

extern "C" int acn_field_bwd(acn_ctx* ctx, const void* enc, int enc_dtype, const float* dirs, int dirs_stride, int dirs_group,
                             int64_t P, const int32_t* range_or_null, int E, int H, int G, int C, const acn_field_weights* w, int precision,
                             const float* d_rgb_sigma, const acn_field_grads* g, void* d_enc_or_null, int d_enc_dtype,
                             acn_stream stream) {
    ACN_CHECK_CTX(ctx);
    int rc = check_field("acn_field_bwd", enc, enc_dtype, dirs, dirs_stride, dirs_group, P, w, precision);
    if (rc) return rc;
    ACN_REQUIRE(g != nullptr, ACN_EINVAL, "acn_field_bwd: null grads");
    ACN_REQUIRE(d_enc_dtype == ACN_F32 || d_enc_dtype == ACN_F16, ACN_EINVAL, "acn_field_bwd: bad d_enc dtype");
    if (P == 0) return ACN_OK;
    ACN_REQUIRE(d_rgb_sigma && ((uintptr_t)d_rgb_sigma & 15) == 0, ACN_EINVAL, "acn_field_bwd: d_rgb_sigma null or misaligned");
    ACN_REQUIRE(!range_or_null || dirs_group == 1, ACN_EINVAL, "acn_field_bwd: a row range needs per-point directions");
    if (precision == ACN_F16) {
        ACN_REQUIRE(d_enc_dtype == ACN_F32, ACN_EUNSUPPORTED, "acn_field_bwd(f16): d_enc must be fp32");
        return acn_field_bwd_tc(ctx, enc, enc_dtype, dirs, dirs_stride, dirs_group, P, E, H, G, C, w, d_rgb_sigma, g,
                                (float*)d_enc_or_null, range_or_null, (cudaStream_t)stream);
    }
    return acn_field_bwd_fp32(ctx, enc, enc_dtype, dirs, dirs_stride, dirs_group, P, E, H, G, C, w, d_rgb_sigma, g,
                              d_enc_or_null, d_enc_dtype, range_or_null, (cudaStream_t)stream);
}
