// oracle/ref_cuda_shim.cu -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// extern "C" doors into the UNMODIFIED reference CUDA operator layer (TensorMul / TensorOpr /
// DeviceTensor), compiled natively for sm_100a by oracle/Makefile from the sources under
// /root/reference into oracle/_ref/libifref_cuda.so. The GPU tests use it as the second
// checker ("outputs match the reference's own CUDA path on the same inputs") and as the
// "reference CUDA kernels on B200" timing baseline. Every function only builds non-owning
// DeviceTensor views over caller memory and calls the reference's own entry points, in the
// order the reference's GpuInferenceWorker calls them (file:line cited).
#include <cstdint>
#include <cmath>
#include <limits>

#include "tensor/device_tensor.h"
#include "tensor/tensor_mul.h"
#include "tensor/tensor_opr.h"

using namespace inferflow;

static void view(DeviceTensor &t, int dt, const void *p, int n0, int n1 = 0, int n2 = 0) {
    t.SetAutoFree(false);
    t.data_type = (ElementType)dt;
    t.data = const_cast<void *>(p);
    t.SetStructure(n0, n1, n2);
}

extern "C" {

// TensorMul::Gemv_AX(Y, A, X, Alg3) -- tensor_mul.cu:97,766
int ifref_cuda_gemv_ax(int w_dt, void *y_f16, const void *w, int x_dt, const void *x, int cy, int cx) {
    DeviceTensor Y(false), A(false), X(false);
    view(Y, (int)ElementType::F16, y_f16, cy);
    view(A, w_dt, w, cx, cy);
    view(X, x_dt, x, cx);
    return TensorMul::Gemv_AX(Y, A, X, VectorMatrixMulAlg::Alg3) ? 0 : -1;
}

// TensorOpr::Quantize(B, A) -- tensor_opr.cu:1623
int ifref_cuda_quantize(int dt, void *dst, const void *src_f16, int rows, int cx) {
    DeviceTensor A(false), B(false);
    view(A, (int)ElementType::F16, src_f16, cx, rows);
    view(B, dt, dst, cx, rows);
    return TensorOpr::Quantize(B, A) ? 0 : -1;
}

// TensorOpr::Dequantize(B, A)
int ifref_cuda_dequantize(int dt, void *dst_f16, const void *src, int rows, int cx) {
    DeviceTensor A(false), B(false);
    view(A, dt, src, cx, rows);
    view(B, (int)ElementType::F16, dst_f16, cx, rows);
    return TensorOpr::Dequantize(B, A) ? 0 : -1;
}

// TensorOpr::RmsNorm(T, S, M) -- tensor_opr.cu:557
int ifref_cuda_rmsnorm(void *y_f16, const void *x_f16, const void *w_f16, int rows, int cols) {
    DeviceTensor T(false), S(false), M(false);
    view(T, (int)ElementType::F16, y_f16, cols, rows);
    view(S, (int)ElementType::F16, x_f16, cols, rows);
    view(M, (int)ElementType::F16, w_f16, cols);
    return TensorOpr::RmsNorm(T, S, w_f16 ? &M : nullptr) ? 0 : -1;
}

// TensorOpr::PositionEmbedding ROPE on x [tokens][heads][head_dim] -- inference_worker.cc:1587-1620
int ifref_cuda_rope(void *x_f16, int tokens, int heads, int head_dim, int context_len, int order, float theta) {
    DeviceTensor A(false);
    view(A, (int)ElementType::F16, x_f16, head_dim, heads, tokens);
    PosEmbeddingParams p;
    p.heads = heads;
    p.dims = head_dim;
    p.order_type = order;
    p.rope_theta = theta;
    p.context_len = context_len;
    return TensorOpr::PositionEmbedding(A, p, 0, tokens) ? 0 : -1;
}

// SiLU then Mul -- inference_worker.cc:1726-1900 (FeedForward): Activation(SILU) ; Mul
int ifref_cuda_silu_mul(void *out_f16, const void *a_f16, const void *b_f16, int n) {
    DeviceTensor O(false), A(false), B(false), T;
    view(O, (int)ElementType::F16, out_f16, n);
    view(A, (int)ElementType::F16, a_f16, n);
    view(B, (int)ElementType::F16, b_f16, n);
    if (!T.New(ElementType::F16, n)) return -2;
    if (!TensorOpr::Activation(T, A, ActivationFn::SILU)) return -1;
    return TensorOpr::Mul(O, T, B) ? 0 : -1;
}

// TensorOpr::Add
int ifref_cuda_add(void *c_f16, const void *a_f16, const void *b_f16, int n) {
    DeviceTensor Cc(false), A(false), B(false);
    view(Cc, (int)ElementType::F16, c_f16, n);
    view(A, (int)ElementType::F16, a_f16, n);
    view(B, (int)ElementType::F16, b_f16, n);
    return TensorOpr::Add(Cc, A, B) ? 0 : -1;
}

// The reference's attention for ONE query token over a context of `ctx` rows
// (GpuInferenceWorker::CalculateProductKQ + ProcessGpuLayer_Attention,
// inference_worker.cc:1639-1724 and :1130-1320): k_f16 / v_f16 are what GetKRows/GetVRows
// return (fp16 [ctx][kv_dim], dequantized if the cache is Q8_B32T2).
int ifref_cuda_attention_decode(void *out_f16, const void *q_f16, const void *k_f16, const void *v_f16,
                                int ctx, int heads, int kv_heads, int head_dim, float kq_scale) {
    const int group = heads / kv_heads;
    const int kv_dim = kv_heads * head_dim, token_dim = heads * head_dim;
    const float neg_inf = -std::numeric_limits<float>::infinity();
    bool ok = true;

    DeviceTensor sub_q(false);
    view(sub_q, (int)ElementType::F16, q_f16, head_dim, heads, 1);
    TensorOpr::TransposeYZ(sub_q);                                          // :1663

    DeviceTensor k_ctx(false);
    view(k_ctx, (int)ElementType::F16, k_f16, kv_dim, ctx);
    ok = ok && TensorOpr::Reshape(k_ctx, 3, head_dim, kv_heads, ctx);       // :1699
    TensorOpr::TransposeYZ(k_ctx);                                          // :1703
    DeviceTensor k_ex;
    DeviceTensor *k_ptr = &k_ctx;
    if (group > 1) {                                                        // :1705-1710
        ok = ok && k_ex.New(ElementType::F16, head_dim, ctx, heads);
        ok = ok && TensorOpr::RepeatKV(k_ex, k_ctx, group);
        k_ptr = &k_ex;
    }
    const float scale = 1.0f / sqrtf(float(head_dim)) / kq_scale;          // :1666
    DeviceTensor kq;
    ok = ok && kq.New(ElementType::F16, ctx, 1, heads);                     // :1717
    ok = ok && TensorMul::Gemm(kq, sub_q, *k_ptr, scale, 0, true, true, MatrixMulAlg::Alg2);   // :1719
    ok = ok && TensorOpr::SoftMax(kq, ctx - 1, neg_inf, kq_scale);          // :1213 (prefix_len = ctx-1)

    DeviceTensor v_ctx(false);
    view(v_ctx, (int)ElementType::F16, v_f16, kv_dim, ctx);
    DeviceTensor v_trans;
    ok = ok && v_trans.New(ElementType::F16, ctx, kv_dim);                  // :1261
    ok = ok && TensorOpr::Transpose(v_trans, v_ctx);                        // :1263
    ok = ok && TensorOpr::Reshape(v_trans, 3, ctx, head_dim, kv_heads);     // :1264
    DeviceTensor v_ex;
    DeviceTensor *v_ptr = &v_trans;
    if (group > 1) {                                                        // :1269-1275
        ok = ok && v_ex.New(ElementType::F16, ctx, head_dim, heads);
        ok = ok && TensorOpr::RepeatKV(v_ex, v_trans, group);
        v_ptr = &v_ex;
    }
    DeviceTensor kqv;
    ok = ok && kqv.New(ElementType::F16, head_dim, 1, heads);               // :1293
    ok = ok && TensorMul::Gemm(kqv, kq, *v_ptr, 1.0f, 0, true, true, MatrixMulAlg::Alg2);      // :1295
    TensorOpr::TransposeYZ(kqv);                                            // :1304
    DeviceTensor merged(false);
    view(merged, (int)ElementType::F16, out_f16, token_dim, 1);
    ok = ok && TensorOpr::Assign(merged, kqv, 0, 0);                        // :1312
    cudaDeviceSynchronize();
    return ok ? 0 : -1;
}

}  // extern "C"
