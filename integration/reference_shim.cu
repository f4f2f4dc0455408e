// The reference-side binding of INTEGRATION.md, as a translation unit that COMPILES against the reference's own
// headers (not part of the product; nothing here is copied from the reference): drop-in bodies for the operator
// entry points of src/tensor/tensor_mul.h / tensor_opr.h that forward to the C ABI of libinferflow_b200.so.
//
//   nvcc -std=c++14 -DUSE_CUDA -I/root/reference/src -I/root/reference/3rd_party -Iinclude -c integration/reference_shim.cu
//
// TensorMul::Gemv_AX and TensorOpr::Quantize below are DEFINITIONS OF THE REFERENCE'S OWN MEMBERS (same signatures):
// oracle/Makefile (target `boundary`) links this object into the reference's CUDA engine in place of the stock
// bodies -- the stock tensor_mul.cu / tensor_opr.cu are compiled with the two names renamed, no source is edited --
// and tests/test_gpu_boundary.py runs that engine beside the stock one on the same config and compares the tokens.
// The other functions (namespace ifb_shim) show the remaining operators of the path in the same style.
// Element-type ids and block bytes are the reference's, so tensors pass through untouched.
// tests/test_integration_shim.py compiles this file when /root/reference is present.
#include "tensor/tensor_mul.h"
#include "tensor/tensor_opr.h"
#include "sslib/log.h"
#include "inferflow_b200.h"

INFER_FLOW_BEGIN

// TensorMul::Gemv_AX (tensor_mul.h:26-28, stock body tensor_mul.cu:97-132): y[cy] = A[cy][cx] . x[cx] (+ bias), fp16
// out. X is half, or Q8_B32T2 when the worker already called TensorOpr::Quantize (GetUseFullQuantGemv,
// inference_worker.cc:2385-2395). The worker passes Alg3; Alg4 (an auxiliary tensor U) has no caller on this path.
bool TensorMul::Gemv_AX(DeviceTensor &Y, const DeviceTensor &A, const DeviceTensor &X,
    VectorMatrixMulAlg alg, DeviceTensor *U, const DeviceTensor *bias)
{
    (void)U;
    const int M = A.Rows(), N = A.Columns();
    if (X.size != N) {
        LogError("A and X are not compatible");
        return false;
    }
    if (Y.size != M) {
        LogError("A and Y are not compatible: (%d, %d, %d) vs. (%d, %d, %d))",
            Y.ne[0], Y.ne[1], Y.ne[2], A.ne[0], A.ne[1], A.ne[2]);
        return false;
    }
    if (alg != VectorMatrixMulAlg::Alg3 || Y.data_type != ElementType::F16) {
        LogError("Gemv_AX: Algorithm %d is not impelemtned.", (int)alg);
        return false;
    }
    int32_t rc = ifb_gemv((int32_t)A.data_type, Y.data, A.data, (int32_t)X.data_type, X.data,
                          M, N, bias != nullptr ? bias->data : nullptr, /*stream=*/nullptr);
    if (rc != 0) {
        LogError("Gemv_AX: %s", ifb_last_error());
        return false;
    }
    return true;
}

// TensorOpr::Quantize (tensor_opr.h:130, stock body tensor_opr.cu:1623-1700): B (a block type) <- A (half); the bytes
// are those of the reference's kernels (tests/test_gpu_vs_reference_cuda.py)
bool TensorOpr::Quantize(DeviceTensor &B, const DeviceTensor &A)
{
    if (A.data_type != ElementType::F16) {
        LogError("Quantize: The data type of A should be F32 or F16");     // F32 sources: none on the decode path
        return false;
    }
    if (A.ne[0] != B.ne[0] || A.Rows() != B.Rows()) {
        LogKeyInfo("Quantize: Tensor B is not compatible with A");
        return false;
    }
    int32_t rc = ifb_quantize_rows((int32_t)B.data_type, B.data, A.data, A.Rows(), A.ne[0], nullptr);
    if (rc != 0) {
        LogError("Quantize: %s", ifb_last_error());
        return false;
    }
    return true;
}

namespace ifb_shim {

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
