// Internal launchers behind acn_field_fwd / acn_field_bwd (dispatch in field_api.cu).
#pragma once
#include "acn_common.cuh"

int acn_field_fwd_fp32(acn_ctx* ctx, const void* enc, int enc_dtype, const float* dirs, int dirs_stride, int dirs_group,
                       int64_t P, int E, int H, int G, int C, const acn_field_weights* w, float* rgb_sigma, const int32_t* range,
                       cudaStream_t st);
int acn_field_bwd_fp32(acn_ctx* ctx, const void* enc, int enc_dtype, const float* dirs, int dirs_stride, int dirs_group,
                       int64_t P, int E, int H, int G, int C, const acn_field_weights* w, const float* d_rgb_sigma,
                       const acn_field_grads* g, void* d_enc, int d_enc_dtype, const int32_t* range, cudaStream_t st);
int acn_field_fwd_tc(acn_ctx* ctx, const void* enc, int enc_dtype, const float* dirs, int dirs_stride, int dirs_group,
                     int64_t P, int E, int H, int G, int C, const acn_field_weights* w, float* rgb_sigma, const int32_t* range,
                       cudaStream_t st);
int acn_field_bwd_tc(acn_ctx* ctx, const void* enc, int enc_dtype, const float* dirs, int dirs_stride, int dirs_group,
                     int64_t P, int E, int H, int G, int C, const acn_field_weights* w, const float* d_rgb_sigma,
                     const acn_field_grads* g, float* d_enc, const int32_t* range, cudaStream_t st);
