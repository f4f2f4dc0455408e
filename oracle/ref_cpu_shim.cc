// oracle/ref_cpu_shim.cc -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// extern "C" doors into the UNMODIFIED reference (compiled by oracle/Makefile from
// /root/reference into oracle/_ref/libifref_cpu.so). Used to (a) pin oracle/oracle.c,
// (b) generate tests/golden/*.npz, (c) run the reference's CPU engine for end-to-end
// token parity and the CPU baseline. Nothing in here re-implements reference logic:
// every function forwards to a reference symbol.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include <map>

#include "common/quantization.h"      // Quantization::QuantizeRow_*, Dequantize*, GetInt4
#include "tensor/host_tensor.h"
#include "tensor/host_tensor_opr.h"   // HostTensorOpr::BuildRowsForMoE
#include "tensor/tensor_common.h"
#include "transformer/inference_engine.h"
#include "transformer/network_builder.h"

using namespace inferflow;
using namespace inferflow::transformer;

extern "C" {

int ifref_sizeof_block(int dt) {
    switch ((ElementType)dt) {
        case ElementType::Q8_B32T1: return (int)sizeof(BlockQ8_B32T1);
        case ElementType::Q8_B32T2: return (int)sizeof(BlockQ8_B32T2);
        case ElementType::Q6_B64T1: return (int)sizeof(BlockQ6_B64T1);
        case ElementType::Q5_B64T1: return (int)sizeof(BlockQ5_B64T1);
        case ElementType::Q5_B32T1: return (int)sizeof(BlockQ5_B32T1);
        case ElementType::Q4_B16: return (int)sizeof(BlockQ4_B16);
        case ElementType::Q4_B32T1A: case ElementType::Q4_B32T1B: return (int)sizeof(BlockQ4_B32T1);
        case ElementType::Q4_B64T1: return (int)sizeof(BlockQ4_B64T1);
        case ElementType::Q3H_B64T1: return (int)sizeof(BlockQ3H_B64T1);
        case ElementType::Q3_B32T1A: case ElementType::Q3_B32T1B: return (int)sizeof(BlockQ3_B32T1);
        case ElementType::Q2_B32T1A: case ElementType::Q2_B32T1B: return (int)sizeof(BlockQ2_B32T1);
        default: return 0;
    }
}

// TensorCommon::ByteCount (tensor_common.cc:162)
long long ifref_byte_count(int dt, long long n) {
    return (long long)TensorCommon::ByteCount((ElementType)dt, (uint64_t)n);
}

// Quantization::QuantizeRow_* with a float source (host flavour; Q8_B32T2 = host algorithm)
int ifref_quantize_row(int dt, uint8_t *dst, const float *src, int n) {
    int big = 1 << 30;
    switch ((ElementType)dt) {
        case ElementType::Q8_B32T1: return Quantization::QuantizeRow_Q8_B32T1((BlockQ8_B32T1 *)dst, big, src, n) ? 0 : -1;
        case ElementType::Q8_B32T2: return Quantization::QuantizeRow_Q8_B32T2((BlockQ8_B32T2 *)dst, big, src, n) ? 0 : -1;
        case ElementType::Q6_B64T1: return Quantization::QuantizeRow_Q6_B64T1((BlockQ6_B64T1 *)dst, big, src, n) ? 0 : -1;
        case ElementType::Q5_B64T1: return Quantization::QuantizeRow_Q5_B64T1((BlockQ5_B64T1 *)dst, big, src, n) ? 0 : -1;
        case ElementType::Q5_B32T1: return Quantization::QuantizeQ5Row((BlockQ5_B32T1 *)dst, src, n) ? 0 : -1;
        case ElementType::Q4_B16: return Quantization::QuantizeRow_Q4B16((BlockQ4_B16 *)dst, big, src, n) ? 0 : -1;
        case ElementType::Q4_B32T1A: return Quantization::QuantizeRow_Q4_B32T1A((BlockQ4_B32T1 *)dst, big, src, n) ? 0 : -1;
        case ElementType::Q4_B32T1B: return Quantization::QuantizeRow_Q4_B32T1B((BlockQ4_B32T1 *)dst, big, src, n) ? 0 : -1;
        case ElementType::Q4_B64T1: return Quantization::QuantizeRow_Q4_B64T1((BlockQ4_B64T1 *)dst, big, src, n) ? 0 : -1;
        case ElementType::Q3H_B64T1: return Quantization::QuantizeRow_Q3H_B64T1((BlockQ3H_B64T1 *)dst, big, src, n) ? 0 : -1;
        case ElementType::Q3_B32T1A: return Quantization::QuantizeRow_Q3_B32T1A((BlockQ3_B32T1 *)dst, big, src, n) ? 0 : -1;
        case ElementType::Q3_B32T1B: return Quantization::QuantizeRow_Q3_B32T1B((BlockQ3_B32T1 *)dst, big, src, n) ? 0 : -1;
        case ElementType::Q2_B32T1A: return Quantization::QuantizeRow_Q2_B32T1A((BlockQ2_B32T1 *)dst, big, src, n) ? 0 : -1;
        case ElementType::Q2_B32T1B: return Quantization::QuantizeRow_Q2_B32T1B((BlockQ2_B32T1 *)dst, big, src, n) ? 0 : -1;
        default: return -1;
    }
}

// Quantization::Dequantize* block by block into float
int ifref_dequantize_row(int dt, float *dst, const uint8_t *src, int n) {
    switch ((ElementType)dt) {
        case ElementType::Q8_B32T1: for (int i = 0; i < n / 32; i++) Quantization::DequantizeQ8_B32T1(dst + 32 * i, (const BlockQ8_B32T1 *)src + i); return 0;
        case ElementType::Q8_B32T2: for (int i = 0; i < n / 32; i++) Quantization::DequantizeQ8_B32T2(dst + 32 * i, (const BlockQ8_B32T2 *)src + i); return 0;
        case ElementType::Q6_B64T1: for (int i = 0; i < n / 64; i++) Quantization::DequantizeQ6_B64T1(dst + 64 * i, (const BlockQ6_B64T1 *)src + i); return 0;
        case ElementType::Q5_B64T1: for (int i = 0; i < n / 64; i++) Quantization::DequantizeQ5_B64T1(dst + 64 * i, (const BlockQ5_B64T1 *)src + i); return 0;
        case ElementType::Q5_B32T1: for (int i = 0; i < n / 32; i++) Quantization::DequantizeQ5Block(dst + 32 * i, (const BlockQ5_B32T1 *)src + i); return 0;
        case ElementType::Q4_B16: for (int i = 0; i < n / 16; i++) Quantization::DequantizeQ4_B16(dst + 16 * i, (const BlockQ4_B16 *)src + i); return 0;
        case ElementType::Q4_B32T1A: case ElementType::Q4_B32T1B:
            for (int i = 0; i < n / 32; i++) Quantization::DequantizeQ4_B32T1(dst + 32 * i, (const BlockQ4_B32T1 *)src + i); return 0;
        case ElementType::Q4_B64T1: for (int i = 0; i < n / 64; i++) Quantization::DequantizeQ4_B64T1(dst + 64 * i, (const BlockQ4_B64T1 *)src + i); return 0;
        case ElementType::Q3H_B64T1: for (int i = 0; i < n / 64; i++) Quantization::DequantizeQ3H_B64T1(dst + 64 * i, (const BlockQ3H_B64T1 *)src + i); return 0;
        case ElementType::Q3_B32T1A: case ElementType::Q3_B32T1B:
            for (int i = 0; i < n / 32; i++) Quantization::DequantizeQ3_B32T1(dst + 32 * i, (const BlockQ3_B32T1 *)src + i); return 0;
        case ElementType::Q2_B32T1A: case ElementType::Q2_B32T1B:
            for (int i = 0; i < n / 32; i++) Quantization::DequantizeQ2_B32T1(dst + 32 * i, (const BlockQ2_B32T1 *)src + i); return 0;
        default: return -1;
    }
}

// Quantization::GetInt4 for the dp4a formats: n/4 packed words per row
int ifref_get_int4_row(int dt, int32_t *dst, const uint8_t *src, int n) {
    switch ((ElementType)dt) {
        case ElementType::Q8_B32T2: for (int i = 0; i < n / 4; i++) dst[i] = Quantization::GetInt4(((const BlockQ8_B32T2 *)src)[i / 8], 4 * (i % 8)); return 0;
        case ElementType::Q6_B64T1: for (int i = 0; i < n / 4; i++) dst[i] = Quantization::GetInt4(((const BlockQ6_B64T1 *)src)[i / 16], 4 * (i % 16)); return 0;
        case ElementType::Q5_B64T1: for (int i = 0; i < n / 4; i++) dst[i] = Quantization::GetInt4(((const BlockQ5_B64T1 *)src)[i / 16], 4 * (i % 16)); return 0;
        case ElementType::Q4_B32T1A: case ElementType::Q4_B32T1B:
            for (int i = 0; i < n / 4; i++) dst[i] = Quantization::GetInt4(((const BlockQ4_B32T1 *)src)[i / 8], 4 * (i % 8)); return 0;
        case ElementType::Q4_B64T1: for (int i = 0; i < n / 4; i++) dst[i] = Quantization::GetInt4(((const BlockQ4_B64T1 *)src)[i / 16], 4 * (i % 16)); return 0;
        case ElementType::Q3H_B64T1: for (int i = 0; i < n / 4; i++) dst[i] = Quantization::GetInt4(((const BlockQ3H_B64T1 *)src)[i / 16], 4 * (i % 16)); return 0;
        default: return -1;
    }
}

// HostTensorOpr::BuildRowsForMoE (host_tensor_opr.cc:190): scores fp16 [rows][experts]
int ifref_moe_topk(int32_t *ids, float *weights, int32_t *cnt, const uint16_t *scores_f16,
                   int rows, int experts, int top_k, int norm_top_k_prob) {
    HostTensor logits;
    if (!logits.New(ElementType::F16, experts, rows)) return -1;
    for (int r = 0; r < rows; r++) memcpy(logits.RowData(r), scores_f16 + (size_t)r * experts, 2 * (size_t)experts);
    std::vector<RowItemForMoe> items;
    if (!HostTensorOpr::BuildRowsForMoE(items, logits, top_k, norm_top_k_prob != 0)) return -1;
    for (int r = 0; r < rows; r++) {
        cnt[r] = items[r].size;
        for (int j = 0; j < top_k; j++) {
            bool ok = j < items[r].size;
            ids[r * top_k + j] = ok ? (int32_t)items[r].arr[j].id : -1;
            weights[r * top_k + j] = ok ? items[r].arr[j].weight : 0.0f;
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------
// The reference engine through its public API (inference_engine.h:39-66,118):
// LoadConfig -> Init -> AddQuery -> (Infer -> CommitInferenceResult)*.
// Greedy decoding; returns generated token ids and (optionally) per-step logits.
// ---------------------------------------------------------------------------
struct IfrefEngine {
    InferenceConfig config;
    InferenceEngine engine;
};

void *ifref_engine_create(const char *ini_path, const char *section, const char *data_root_dir) {
    IfrefEngine *e = new IfrefEngine;
    if (!InferenceEngine::LoadConfig(e->config, ini_path, section, data_root_dir ? data_root_dir : "")) {
        delete e; return nullptr;
    }
    if (!e->engine.Init(e->config)) { delete e; return nullptr; }
    return e;
}

void ifref_engine_destroy(void *h) { delete (IfrefEngine *)h; }

// Run one query given as token ids; greedy; writes up to max_new token ids.
// If logits_out != null and the config has return_output_tensors = true, the last-row
// logits of every step (vocab floats each) are appended. Returns #tokens generated or <0.
int ifref_engine_generate(void *h, const int32_t *prompt, int prompt_len, int max_new,
                          int32_t *out_tokens, float *logits_out, int vocab, double *seconds_decode) {
    IfrefEngine *e = (IfrefEngine *)h;
    std::vector<int> tokens(prompt, prompt + prompt_len);
    SamplingStrategy::QueryOptions opt;
    opt.strategy_id = SamplingStrategyId::Greedy;
    int query_id = e->engine.AddQuery(tokens, opt);
    if (query_id <= 0) return -2;
    int n = 0;
    sslib::TaskMonitor tm;
    double decode_s = 0;
    bool done = false;
    for (int step = 0; step < max_new && !done; step++) {
        InferenceResult res;
        tm.Start();
        if (!e->engine.Infer(res)) { e->engine.RemoveQuery(query_id); return -3; }
        double dt = (double)tm.GetElapsedTime(false) / 1e6;
        if (step > 0) decode_s += dt;
        if (res.items.empty()) break;
        std::map<int, QueryNextToken> commit;
        for (int i = 0; i < (int)res.items.size(); i++) {
            const QueryInferenceResult &item = *res.items[i];
            if (item.query_id != query_id || item.next_tokens.empty()) continue;
            out_tokens[n] = (int32_t)item.next_tokens[0].id;
            if (logits_out != nullptr && item.output_tensor.size > 0) {
                // last row of the per-query output tensor (inference_engine.cc:958-962)
                const HostTensor &t = item.output_tensor;
                int rows = t.Rows();
                if (t.data_type == ElementType::F32) {
                    const float *p = (const float *)t.RowData(rows - 1);
                    for (int k = 0; k < vocab; k++) logits_out[(size_t)n * vocab + k] = p[k];
                } else {
                    const inferflow_fp16 *p = (const inferflow_fp16 *)t.RowData(rows - 1);
                    for (int k = 0; k < vocab; k++) logits_out[(size_t)n * vocab + k] = (float)p[k];
                }
            }
            QueryNextToken nt;
            nt.id = (int)item.next_tokens[0].id;
            nt.is_end = (n + 1 >= max_new);
            commit[query_id] = nt;
            n++;
            if (nt.is_end) done = true;
        }
        e->engine.CommitInferenceResult(commit);
    }
    if (seconds_decode) *seconds_decode = decode_s;
    return n;
}

// ---------------------------------------------------------------------------
// the partitioner's host logic, as the reference computes it (pins csrc/ifb_partition.cu)
// ---------------------------------------------------------------------------
// InferenceEngine::LoadDeviceGroups (inference_engine.cc:1738) on item `devices` of `section` of a REAL .ini file read
// by the reference's own ConfigData. sizes_out[g] = devices in group g, ids_out = all ids in order.
// returns the number of groups, -1 when LoadDeviceGroups refuses the string, -2 when the file cannot be read.
int ifref_load_device_groups(const char *ini_path, const char *section, int32_t *ids_out, int ids_cap,
                             int32_t *sizes_out, int sizes_cap) {
    sslib::ConfigData cfg;
    if (!cfg.Load(ini_path)) return -2;
    std::vector<std::vector<int>> groups;
    if (!InferenceEngine::LoadDeviceGroups(groups, cfg, section, "devices")) return -1;
    int k = 0;
    for (int g = 0; g < (int)groups.size(); g++) {
        if (g < sizes_cap) sizes_out[g] = (int32_t)groups[g].size();
        for (int id : groups[g]) if (k < ids_cap) ids_out[k++] = id;
    }
    return (int)groups.size();
}

// NetworkBuilder::SplitGpuLayers (network_builder.cc:2094) for n_groups groups of one device each.
// ranges_out[3k..3k+2] = {start_layer, end_layer, group}; returns the number of assignments.
int ifref_split_gpu_layers(int start_layer, int end_layer, int n_groups, int32_t *ranges_out, int cap) {
    ModelSpec spec;
    for (int g = 0; g < n_groups; g++) spec.device_groups.push_back(std::vector<int>(1, g));
    std::vector<LayerAssignment> la;
    NetworkBuilder::SplitGpuLayers(la, start_layer, end_layer, spec);
    int k = 0;
    for (const auto &a : la) {
        if (k >= cap) break;
        ranges_out[3 * k] = a.start_layer;
        ranges_out[3 * k + 1] = a.end_layer;
        ranges_out[3 * k + 2] = (int32_t)(a.devices - &spec.device_groups[0]);
        k++;
    }
    return k;
}

}  // extern "C"
