// Pieces shared by the fp32 (SIMT) and fp16 (tcgen05) field-MLP kernels.
#pragma once
#include "acn_common.cuh"

// models/encodings.py:27-81 real SH basis, degree <= 3, on a unit vector.
__device__ __forceinline__ void sh16_poly(float x, float y, float z, float* sh) {
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

// models/inr/meta_ngp.py:166-169 then models/encodings.py:141: d / max(|d|, 1e-9), twice.
__device__ __forceinline__ void sh16_expert(float x, float y, float z, float* sh) {
#pragma unroll
    for (int rep = 0; rep < 2; ++rep) {
        float n = fmaxf(sqrtf(x * x + y * y + z * z), 1e-9f);
        x = __fdiv_rn(x, n); y = __fdiv_rn(y, n); z = __fdiv_rn(z, n);
    }
    sh16_poly(x, y, z, sh);
}

// models/trunc_exp.py:32-61
__device__ __forceinline__ float trunc_exp_f(float x) { return expf(fminf(fmaxf(x, -88.722839111f), 88.722839111f)); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

// Direction of point p: one row per `group` consecutive points (group = S when dirs are rays).
__device__ __forceinline__ const float* dir_of(const float* __restrict__ dirs, int stride, int group, int64_t p) {
    return dirs + (group > 1 ? p / group : p) * (int64_t)stride;
}
