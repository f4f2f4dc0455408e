// The hybrid / tensor / layer partitioner's host logic: the `devices` grammar, the strategy it implies, the layer
// ranges of the device groups, the split rule of every layer tensor, and the place of one process in the plan.
//
// Mirrors (reference, read-only):
//   InferenceEngine::LoadDeviceGroups      src/transformer/inference_engine.cc:1738-1783   `0&1;2&3` -> {{0,1},{2,3}}
//   strategy from the groups               src/transformer/inference_engine.cc:1509-1515   BY_LAYER / BY_TENSOR / HYBRID
//   NetworkBuilder::SplitGpuLayers         src/transformer/network_builder.cc:2094-2117    ceil(L/groups) contiguous layers
//   Build{Attn,Ffn,Moe}TensorIdMap +       src/transformer/network_builder.cc:1373-1413,
//   AddLayerTasks_TensorParallel                                               :1594-1688   BY_COL / BY_ROW / DUP per tensor,
//                                                                                           per-tensor tensor_quant_threshold
// The reference drives all GPUs of a box from ONE process (a worker thread per device, the groups one after the
// other). This runtime is one process per GPU: ifb_rank_plan says which group, which tensor-parallel rank and which
// layers a process (rank r = the r-th device of the string) owns, and which ranks it takes the hidden state from /
// hands it to. Pure host code: no CUDA call in this file.
#include <cctype>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "ifb_common.cuh"
#include "ifb_formats.cuh"
#include "../../include/inferflow_b200.h"

using namespace ifb;

namespace {

std::string trim_ws(const std::string &s) {
    // String::Trim (sslib): blanks, tabs, line ends
    size_t a = 0, b = s.size();
    while (a < b && (s[a] == ' ' || s[a] == '\t' || s[a] == '\r' || s[a] == '\n')) a++;
    while (b > a && (s[b - 1] == ' ' || s[b - 1] == '\t' || s[b - 1] == '\r' || s[b - 1] == '\n')) b--;
    return s.substr(a, b - a);
}

// String::Split(str, tokens, delims): every delimiter separates; empty pieces are kept (the caller skips them)
void split_any(const std::string &s, const char *delims, std::vector<std::string> &out) {
    out.clear();
    size_t a = 0;
    for (;;) {
        const size_t b = s.find_first_of(delims, a);
        out.push_back(s.substr(a, b == std::string::npos ? std::string::npos : b - a));
        if (b == std::string::npos) break;
        a = b + 1;
    }
}

int parse_groups(const char *devices, std::vector<std::vector<int>> &groups) {
    groups.clear();
    std::vector<std::string> tokens, subs;
    split_any(devices ? devices : "", ",;", tokens);
    for (auto &t : tokens) {
        const std::string tok = trim_ws(t);
        if (tok.empty()) continue;
        std::vector<int> g;
        split_any(tok, "&|", subs);
        for (auto &s : subs) {
            const std::string id = trim_ws(s);
            if (!id.empty()) g.push_back(atoi(id.c_str()));     // atoi as the reference: "gpu1" is device 0
        }
        groups.push_back(g);                                    // a group of separators only ("&") stays, empty
    }
    if (groups.size() > 1) {
        const size_t n0 = groups[0].size();
        for (size_t i = 1; i < groups.size(); i++) {
            if (groups[i].size() != n0) {
                set_last_error("All device groups should have the same size: %d vs. %d", (int)n0, (int)groups[i].size());
                return IFB_ERR_ARG;
            }
        }
    }
    return IFB_OK;
}

int strategy_of(const std::vector<std::vector<int>> &groups) {
    // inference_engine.cc:1509-1515: group_size <= 1 -> BY_LAYER; one group -> BY_TENSOR; else HYBRID
    if (groups.empty()) return IFB_STRATEGY_BY_LAYER;
    const int gs = (int)groups[0].size();
    return gs <= 1 ? IFB_STRATEGY_BY_LAYER : (groups.size() == 1 ? IFB_STRATEGY_BY_TENSOR : IFB_STRATEGY_HYBRID);
}

}  // namespace

extern "C" {

int32_t ifb_device_groups_parse(const char *devices, int32_t *ids_out, int32_t capacity, int32_t *n_groups,
                                int32_t *group_size, int32_t *strategy) {
    std::vector<std::vector<int>> groups;
    const int rc = parse_groups(devices, groups);
    if (rc != IFB_OK) return rc;
    const int ng = (int)groups.size(), gs = ng > 0 ? (int)groups[0].size() : 0;
    if (n_groups) *n_groups = ng;
    if (group_size) *group_size = gs;
    if (strategy) *strategy = strategy_of(groups);
    if (ids_out != nullptr) {
        IFB_CHECK_ARG(capacity >= ng * gs, "device_groups_parse: %d ids do not fit a buffer of %d", ng * gs, capacity);
        int k = 0;
        for (auto &g : groups)
            for (int id : g) ids_out[k++] = id;
    }
    return IFB_OK;
}

int32_t ifb_split_layers(int32_t start_layer, int32_t end_layer, int32_t n_groups, int32_t *ranges_out,
                         int32_t *group_of_range, int32_t *n_ranges) {
    IFB_CHECK_ARG(n_ranges != nullptr, "split_layers: null count");
    *n_ranges = 0;
    if (n_groups <= 0) return IFB_OK;
    const int layer_num = end_layer - start_layer;
    const int per_group = (layer_num + n_groups - 1) / n_groups;
    int k = 0;
    for (int g = 0; g < n_groups; g++) {
        const int a = start_layer + g * per_group;
        // as written in the reference: the end is NOT offset by start_layer (the same thing whenever start_layer == 0,
        // the only value its callers pass: GetDeviceAssignments, network_builder.cc:2120-2145, start_gpu_layer = 0)
        const int b = g + 1 == n_groups ? end_layer : (g + 1) * per_group;
        if (b > a) {                                    // a group without layers gets no assignment
            if (ranges_out) { ranges_out[2 * k] = a; ranges_out[2 * k + 1] = b; }
            if (group_of_range) group_of_range[k] = g;
            k++;
        }
    }
    *n_ranges = k;
    return IFB_OK;
}

int32_t ifb_tensor_partition_type(const char *name) {
    // layer tensors by their llama names (the loader's names, model_reader / network_structure): BY_COL splits the
    // OUTPUT rows (ne[1]), BY_ROW the contraction dimension (ne[0], block-aligned), DUP copies
    if (name == nullptr) return -1;
    std::string n(name);
    const size_t dot = n.find_last_of('.');
    if (dot != std::string::npos) n = n.substr(dot + 1);
    for (auto &c : n) c = (char)tolower((unsigned char)c);
    if (n == "wq" || n == "wk" || n == "wv" || n == "qkv" || n == "w1" || n == "w3") return IFB_PART_BY_COL;       // type 2
    if (n == "wo" || n == "w2") return IFB_PART_BY_ROW;                                                              // types 3, 5
    if (n == "wq_b" || n == "wk_b" || n == "wv_b" || n == "qkv_b" || n == "w1_b" || n == "w3_b") return IFB_PART_BY_COL;   // type 1: biases of split outputs
    if (n == "attention_norm" || n == "ffn_norm" || n == "attn_pre_norm" || n == "ffn_pre_norm" || n == "attn_post_norm" ||
        n == "ffn_post_norm" || n == "wo_b" || n == "w2_b" || n == "moe_gate" || n == "gate")
        return IFB_PART_DUP;                                                                                         // type 0
    return -1;
}

int32_t ifb_tensor_device_dtype(int64_t elements, int32_t n_devices, int32_t weight_dtype, int64_t tensor_quant_threshold) {
    // network_builder.cc:1558-1561 (one device) and :1648-1652 (a tensor-parallel group): a matrix whose share per
    // device is below the threshold stays fp16 although the model asks for a quantized type
    if (n_devices < 1) n_devices = 1;
    const bool quant = weight_dtype != DT_F16 && weight_dtype != DT_F32 && weight_dtype != DT_BF16;
    if (quant && elements / n_devices < tensor_quant_threshold) return DT_F16;
    return weight_dtype;
}

int32_t ifb_rank_plan(const char *devices, int32_t rank, int32_t n_layers, int32_t plan_out[IFB_PLAN_FIELDS]) {
    IFB_CHECK_ARG(plan_out != nullptr && n_layers > 0, "rank_plan: null output or no layers");
    std::vector<std::vector<int>> groups;
    int rc = parse_groups(devices, groups);
    if (rc != IFB_OK) return rc;
    const int ng = (int)groups.size(), gs = ng > 0 ? (int)groups[0].size() : 0;
    IFB_CHECK_ARG(ng > 0 && gs > 0, "rank_plan: the device string '%s' names no device", devices ? devices : "");
    IFB_CHECK_ARG(rank >= 0 && rank < ng * gs, "rank_plan: rank %d of %d devices", rank, ng * gs);
    std::vector<int32_t> ranges(2 * ng), gof(ng);
    int32_t nr = 0;
    ifb_split_layers(0, n_layers, ng, ranges.data(), gof.data(), &nr);
    {   // SplitGpuLayers can hand a middle group an end beyond the last layer (5 layers over 4 groups: [4, 6)); the
        // reference's workers stop at end_layer (SetInput(..., end_layer), inference_engine.cc:1203): clamp, drop empties
        int k = 0;
        for (int i = 0; i < nr; i++) {
            const int a = ranges[2 * i], b = ranges[2 * i + 1] < n_layers ? ranges[2 * i + 1] : n_layers;
            if (b > a) { ranges[2 * k] = a; ranges[2 * k + 1] = b; gof[k] = gof[i]; k++; }
        }
        nr = k;
    }
    const int g = rank / gs, t = rank % gs;
    int idx = -1;
    for (int i = 0; i < nr; i++)
        if (gof[i] == g) idx = i;
    plan_out[IFB_PLAN_GROUP] = g;
    plan_out[IFB_PLAN_TP_RANK] = t;
    plan_out[IFB_PLAN_TP_SIZE] = gs;
    plan_out[IFB_PLAN_LAYER_BEGIN] = idx >= 0 ? ranges[2 * idx] : 0;
    plan_out[IFB_PLAN_LAYER_END] = idx >= 0 ? ranges[2 * idx + 1] : 0;       // begin == end: this group has no layers (more groups than layers)
    // the hidden state travels group to group between the ranks of equal tensor-parallel index
    plan_out[IFB_PLAN_PREV_RANK] = (idx > 0) ? gof[idx - 1] * gs + t : -1;
    plan_out[IFB_PLAN_NEXT_RANK] = (idx >= 0 && idx + 1 < nr) ? gof[idx + 1] * gs + t : -1;
    plan_out[IFB_PLAN_DEVICE] = groups[g][t];
    plan_out[IFB_PLAN_STRATEGY] = strategy_of(groups);
    return IFB_OK;
}

}  // extern "C"
