// The reference-side binding of INTEGRATION.md, as a translation unit that COMPILES against the reference's own
// headers (not part of the product; nothing here is copied from the reference): drop-in bodies for the operator
// entry points of src/tensor/tensor_mul.h / tensor_opr.h that forward to the C ABI of libinferflow_b200.so.
//
//   nvcc -std=c++14 -DUSE_CUDA -I/root/reference/src -I/root/reference/3rd_party -Iinclude -c integration/reference_shim.cu
//
// A reference maintainer replaces the bodies in tensor_mul.cu / tensor_opr.cu by these (or links this object in
// place of them); element-type ids and block bytes are the reference's, so tensors pass through untouched.
// tests/test_integration_shim.py compiles this file when /root/reference is present.
#include "tensor/tensor_mul.h"
#include "tensor/tensor_opr.h"
#include "sslib/log.h"
#include "inferflow_b200.h"

INFER_FLOW_BEGIN

namespace ifb_shim {

// TensorMul::Gemv_AX (tensor_mul.h:27, tensor_mul.cu:97): y[cy] = A[cy][cx] . x[cx] (+ bias), fp16 out.
// X is half, or Q8_B32T2 when the worker already called TensorOpr::Quantize (GetUseFullQuantGemv).
bool Gemv_AX(DeviceTensor &Y, const DeviceTensor &A, const DeviceTensor &X, const DeviceTensor *bias)
{
    const int cx = A.ne[0], cy = A.Rows();
    if (X.ne[0] != cx || Y.ne[0] != cy || Y.data_type != ElementType::F16) {
        LogError("Gemv_AX: incompatible tensors");
        return false;
    }
    int32_t rc = ifb_gemv((int32_t)A.data_type, Y.data, A.data, (int32_t)X.data_type, X.data,
                          cy, cx, bias != nullptr ? bias->data : nullptr, /*stream=*/nullptr);
    if (rc != 0) {
        LogError("Gemv_AX: %s", ifb_last_error());
        return false;
    }
    return true;
}

// TensorOpr::Quantize (tensor_opr.h:130): B (any block type) <- A (half); bytes identical to the reference kernels
bool Quantize(DeviceTensor &B, const DeviceTensor &A)
{
    if (A.data_type != ElementType::F16 || A.ne[0] != B.ne[0] || A.Rows() != B.Rows()) {
        LogError("Quantize: incompatible tensors");
        return false;
    }
    int32_t rc = ifb_quantize_rows((int32_t)B.data_type, B.data, A.data, A.Rows(), A.ne[0], nullptr);
    if (rc != 0) {
        LogError("Quantize: %s", ifb_last_error());
        return false;
    }
    return true;
}

// TensorOpr::Dequantize (tensor_opr.h:128): B (half) <- A (block type)
bool Dequantize(DeviceTensor &B, const DeviceTensor &A)
{
    int32_t rc = ifb_dequantize_rows((int32_t)A.data_type, B.data, A.data, A.Rows(), A.ne[0], nullptr);
    if (rc != 0) {
        LogError("Dequantize: %s", ifb_last_error());
        return false;
    }
    return true;
}

// TensorOpr::RmsNorm (tensor_opr.h:68): T = S * rsqrt(mean(S^2) + 1e-5) * (multi_base + M)
bool RmsNorm(DeviceTensor &T, const DeviceTensor &S, const DeviceTensor *M, float multi_base)
{
    if (M == nullptr || T.ne[0] != S.ne[0]) {
        LogError("RmsNorm: incompatible tensors");
        return false;
    }
    int32_t rc = ifb_rmsnorm(T.data, S.data, M->data, S.Rows(), S.ne[0], 1e-5f, multi_base, nullptr);
    if (rc != 0) {
        LogError("RmsNorm: %s", ifb_last_error());
        return false;
    }
    return true;
}

// The decode attention of one query token (CalculateProductKQ + SoftMax + RepeatKV + Gemm, inference_worker.cc:
// 1130-1320,1639-1724) over caches held in this library's layout (ifb_kv_cache_bytes / ifb_kv_append)
bool AttentionDecode(DeviceTensor &out, const DeviceTensor &q, const void *kcache, const void *vcache,
                     ElementType kv_type, int heads, int kv_heads, int head_dim, int max_context_len,
                     int prefix_len, float kq_scale, void *workspace)
{
    int32_t rc = ifb_attn_decode((int32_t)kv_type, out.data, q.data, kcache, vcache, heads, kv_heads, head_dim,
                                 max_context_len, prefix_len + 1, kq_scale, workspace, nullptr);
    if (rc != 0) {
        LogError("AttentionDecode: %s", ifb_last_error());
        return false;
    }
    return true;
}

}  // namespace ifb_shim

INFER_FLOW_END
