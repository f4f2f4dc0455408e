// Engine-shaped host layer over the decode runtime: the outermost seam of the reference for this path.
//
// Mirrors, for ONE live query and greedy decoding, the calls tools and the service make on
// inferflow::transformer::InferenceEngine (src/transformer/inference_engine.h:39-66,118):
//   LoadConfig(ini, section, data_root_dir) + Init   -> ifb_engine_create
//   AddQuery(tokens)                                  -> ifb_engine_add_query   (>0 query id, 0 busy, <0 error)
//   Infer(result)                                     -> ifb_engine_infer       (next token [+ logits of the step])
//   CommitInferenceResult({query: token, is_end})     -> ifb_engine_commit
//   RemoveQuery                                       -> ifb_engine_remove_query
// and reads the same files: the .ini in the layout of bin/llm_inference*.ini (sections [transformer_engine],
// [model.<id>], [app_env.base]; ${config_dir} / ${data_root_dir} / ${model_name} macros) and the model
// specification json next to the weights (LoadModelSpec, inference_engine.cc:1560-1760; ModelReader::Load,
// model_reader.cc:20-150). Supported model_file_format: "llama2.c" (the reference's tiny models). Tokenizer,
// sampling strategies other than greedy, batching of several queries and the other file formats are out of scope
// (SURVEY.md section 8): the engine takes and returns token ids.
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

#include "ifb_common.cuh"
#include "ifb_formats.cuh"
#include "../../include/inferflow_b200.h"

using namespace ifb;

namespace {

std::string trim(const std::string &s) {
    size_t a = 0, b = s.size();
    while (a < b && (s[a] == ' ' || s[a] == '\t' || s[a] == '\r' || s[a] == '\n')) a++;
    while (b > a && (s[b - 1] == ' ' || s[b - 1] == '\t' || s[b - 1] == '\r' || s[b - 1] == '\n')) b--;
    return s.substr(a, b - a);
}

std::string lower(std::string s) {
    for (auto &c : s) c = (char)tolower((unsigned char)c);
    return s;
}

struct Ini {
    std::map<std::string, std::map<std::string, std::string>> sec;
    std::map<std::string, std::string> macros;

    bool load(const std::string &path) {
        FILE *f = fopen(path.c_str(), "rb");
        if (!f) return false;
        std::string cur;
        char line[8192];
        while (fgets(line, sizeof(line), f)) {
            std::string s = trim(line);
            if (s.empty() || s[0] == ';' || s[0] == '#') continue;
            if (s[0] == '[') {
                size_t e = s.find(']');
                if (e != std::string::npos) cur = lower(trim(s.substr(1, e - 1)));
                continue;
            }
            size_t eq = s.find('=');
            if (eq == std::string::npos) continue;
            // group and item names are case-insensitive (sslib ConfigData, 3rd_party/sslib/config_data.cc:39-121)
            sec[cur][lower(trim(s.substr(0, eq)))] = trim(s.substr(eq + 1));
        }
        fclose(f);
        return true;
    }
    std::string expand(std::string v) const {
        for (int guard = 0; guard < 16; guard++) {
            size_t a = v.find("${");
            if (a == std::string::npos) break;
            size_t b = v.find('}', a);
            if (b == std::string::npos) break;
            auto it = macros.find(v.substr(a + 2, b - a - 2));
            v = v.substr(0, a) + (it != macros.end() ? it->second : std::string()) + v.substr(b + 1);
        }
        return v;
    }
    bool get(const std::string &s, const std::string &k, std::string &out) const {
        auto i = sec.find(lower(s));
        if (i == sec.end()) return false;
        auto j = i->second.find(lower(k));
        if (j == i->second.end()) return false;
        out = expand(j->second);
        return true;
    }
};

// the few model-spec fields this path needs, read from the json text by key (model_spec_*.json is flat enough)
bool json_value_pos(const std::string &t, const std::string &key, size_t &pos) {
    const std::string k = "\"" + key + "\"";
    size_t a = t.find(k);
    if (a == std::string::npos) return false;
    a = t.find(':', a + k.size());
    if (a == std::string::npos) return false;
    a++;
    while (a < t.size() && (t[a] == ' ' || t[a] == '\t' || t[a] == '\r' || t[a] == '\n')) a++;
    pos = a;
    return a < t.size();
}
bool json_string(const std::string &t, const std::string &key, std::string &out) {
    size_t p;
    if (!json_value_pos(t, key, p)) return false;
    if (t[p] == '[') {   // first element of an array of strings
        p = t.find_first_of("\"]", p + 1);
        if (p == std::string::npos || t[p] == ']') return false;
    }
    if (t[p] != '"') return false;
    size_t e = t.find('"', p + 1);
    if (e == std::string::npos) return false;
    out = t.substr(p + 1, e - p - 1);
    return true;
}
bool json_int(const std::string &t, const std::string &key, int &out) {
    size_t p;
    if (!json_value_pos(t, key, p)) return false;
    char *end = nullptr;
    long v = strtol(t.c_str() + p, &end, 10);
    if (end == t.c_str() + p) return false;
    out = (int)v;
    return true;
}

struct Resolved {
    std::string model_path, model_id;
    int weight_dtype = DT_F16, kv_dtype = DT_F16, max_ctx = 0, rope_order = 0, return_logits = 0;
    int tensor_quant_threshold = 2000 * 2000;   // model.h:137
    std::string devices;                        // device groups as written ("0", "0;1", "0&1;2&3")
};

int resolve_config(const char *ini_path, const char *section, const char *data_root_dir, Resolved &r) {
    Ini ini;
    std::string path(ini_path);
    if (!ini.load(path)) { set_last_error("engine: cannot read the config file %s", ini_path); return IFB_ERR_IO; }
    size_t slash = path.find_last_of('/');
    ini.macros["config_dir"] = slash == std::string::npos ? std::string("./") : path.substr(0, slash + 1);
    std::string root = data_root_dir ? data_root_dir : "";
    if (root.empty()) ini.get("app_env.base", "data_root_dir", root);
    ini.macros["data_root_dir"] = root;
    const std::string sec = (section && *section) ? section : "transformer_engine";
    std::string models;
    if (!ini.get(sec, "models", models) || trim(models).empty()) { set_last_error("engine: [%s] has no 'models' item", sec.c_str()); return IFB_ERR_IO; }
    r.model_id = trim(models.substr(0, models.find_first_of(",;")));
    ini.macros["model_name"] = r.model_id;
    const std::string msec = "model." + r.model_id;
    std::string dir, spec_file, s;
    if (!ini.get(msec, "model_dir", dir) || !ini.get(msec, "model_specification_file", spec_file)) {
        set_last_error("engine: [%s] needs model_dir and model_specification_file", msec.c_str());
        return IFB_ERR_IO;
    }
    if (!dir.empty() && dir.back() != '/') dir += '/';
    if (ini.get(msec, "device_weight_data_type", s) && !s.empty()) {
        if (lower(s) == "fp16") s = "f16";
        const int dt = ifb_dtype_from_name(s.c_str());
        if (dt < 0 || dt == DT_F32) { set_last_error("engine: invalid device_weight_data_type '%s' for model %s", s.c_str(), r.model_id.c_str()); return IFB_ERR_TYPE; }
        r.weight_dtype = dt;
    }
    if (ini.get(msec, "device_kv_cache_data_type", s) && !s.empty()) {
        if (lower(s) == "fp16") s = "f16";
        if (lower(s) == "fp32") s = "f32";
        const int dt = ifb_dtype_from_name(s.c_str());
        if (dt < 0) { set_last_error("engine: invalid device_kv_cache_data_type '%s' for model %s", s.c_str(), r.model_id.c_str()); return IFB_ERR_TYPE; }
        // element size >= 2 -> F16, else Q8_B32T2 (inference_engine.cc:1702-1704)
        r.kv_dtype = (dt == DT_F16 || dt == DT_F32 || dt == DT_BF16) ? DT_F16 : DT_Q8_B32T2;
    }
    if (ini.get(msec, "max_context_len", s)) r.max_ctx = atoi(s.c_str());
    if (ini.get(msec, "tensor_quant_threshold", s)) r.tensor_quant_threshold = atoi(s.c_str());
    if (ini.get(sec, "return_output_tensors", s)) r.return_logits = (lower(s) == "true" || s == "1") ? 1 : 0;
    // engine-wide device groups; a model section may name its own (inference_engine.cc:1490-1515)
    if (ini.get(sec, "devices", s)) r.devices = s;
    if (ini.get(msec, "devices", s) && !trim(s).empty()) r.devices = s;

    // the model specification
    const std::string spec_path = dir + spec_file;
    FILE *f = fopen(spec_path.c_str(), "rb");
    if (!f) { set_last_error("engine: failed to open the model-spec file %s", spec_path.c_str()); return IFB_ERR_IO; }
    std::string text;
    char buf[4096];
    size_t n;
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) text.append(buf, n);
    fclose(f);
    std::string fmt, file;
    if (!json_string(text, "model_file_format", fmt) || !json_string(text, "model_files", file)) {
        set_last_error("engine: %s lacks model_file_format / model_files", spec_path.c_str());
        return IFB_ERR_IO;
    }
    if (lower(fmt) != "llama2.c") {
        set_last_error("engine: model_file_format '%s' is not read by this runtime (llama2.c only)", fmt.c_str());
        return IFB_ERR_TYPE;
    }
    int order = 0;
    if (json_int(text, "qk_column_order", order)) r.rope_order = order;
    r.model_path = dir + file;
    return IFB_OK;
}

}  // namespace

// one pipeline stage of the engine: a contiguous layer range on one device (BY_LAYER, SplitGpuLayers)
struct Stage {
    ifb_model_t *model = nullptr;
    int device = 0, layer_begin = 0, layer_end = 0;
};

struct ifb_engine {
    ifb_model_t *model = nullptr;      // the last stage's model (the whole model when there is one stage): logits, next token
    std::vector<Stage> stages;
    int d_model = 0;
    Resolved cfg;
    int vocab = 0, max_ctx = 0;
    int pos = 0;                       // tokens already in the KV cache
    std::vector<int32_t> to_feed;      // tokens the next Infer has to run through the model
    bool active = false;
    int query_id = 0;
    std::vector<uint16_t> logits_h;
};

extern "C" {

// What ifb_engine_create would load, without touching the GPU: resolved weight file, element types, context length.
// info_out: {weight_dtype, kv_dtype, max_context_len, qk_column_order, return_output_tensors}
int32_t ifb_engine_resolve(const char *ini_path, const char *section, const char *data_root_dir, char *model_path_out,
                           int32_t path_capacity, int32_t info_out[5]) {
    IFB_CHECK_ARG(ini_path != nullptr, "engine_resolve: null config path");
    Resolved r;
    int rc = resolve_config(ini_path, section, data_root_dir, r);
    if (rc != IFB_OK) return rc;
    int32_t hv[7];
    int32_t shared = 1;
    rc = ifb_llama2c_probe(r.model_path.c_str(), hv, &shared);
    if (rc != IFB_OK) return rc;
    // load-time policy of the reference: a matrix below tensor_quant_threshold elements stays fp16, decided PER TENSOR
    // (network_builder.cc:1558-1561; ifb_tensor_device_dtype). This runtime holds one element type for all layer
    // matrices: the policy is followed when it gives every matrix the same type (the reference's tiny models: all below
    // the threshold; 7B and larger: all above). A model that splits (grouped-query models whose wk / wv alone fall below
    // it) is refused by name rather than loaded under a different policy.
    const int64_t d = hv[0], f = hv[1], kvd = (int64_t)hv[4] * (hv[0] / hv[3]);
    const struct { const char *name; int64_t elements; } mats[] = {
        {"wq / wo", d * d}, {"wk / wv", kvd * d}, {"w1 / w3 / w2", f * d}};
    int first = ifb_tensor_device_dtype(mats[0].elements, 1, r.weight_dtype, r.tensor_quant_threshold);
    for (const auto &mt : mats) {
        const int dt = ifb_tensor_device_dtype(mt.elements, 1, r.weight_dtype, r.tensor_quant_threshold);
        if (dt != first) {
            set_last_error("engine: tensor_quant_threshold = %d keeps %s (%lld elements) in another element type than wq (%lld): "
                           "per-tensor weight types are not supported by this runtime; set tensor_quant_threshold so that all "
                           "layer matrices fall on one side", r.tensor_quant_threshold, mt.name, (long long)mt.elements,
                           (long long)mats[0].elements);
            return IFB_ERR_TYPE;
        }
    }
    r.weight_dtype = first;
    if (r.max_ctx <= 0) r.max_ctx = hv[6];            // inference_engine.cc:112-113
    else if (r.max_ctx <= 16) r.max_ctx = 4096;        // ModelSpec::DEFAULT_MAX_CONTEXT_LEN
    if (model_path_out != nullptr && path_capacity > 0) {
        IFB_CHECK_ARG((int)r.model_path.size() < path_capacity, "engine_resolve: path buffer too small");
        memcpy(model_path_out, r.model_path.c_str(), r.model_path.size() + 1);
    }
    if (info_out) { info_out[0] = r.weight_dtype; info_out[1] = r.kv_dtype; info_out[2] = r.max_ctx; info_out[3] = r.rope_order; info_out[4] = r.return_logits; }
    return IFB_OK;
}

int32_t ifb_engine_create(ifb_engine_t **out, const char *ini_path, const char *section, const char *data_root_dir) {
    IFB_CHECK_ARG(out != nullptr && ini_path != nullptr, "engine_create: null argument");
    char path[4096];
    int32_t info[5];
    int rc = ifb_engine_resolve(ini_path, section, data_root_dir, path, sizeof(path), info);
    if (rc != IFB_OK) return rc;
    ifb_engine *e = new ifb_engine();
    e->cfg.model_path = path;
    e->cfg.weight_dtype = info[0]; e->cfg.kv_dtype = info[1]; e->cfg.max_ctx = info[2]; e->cfg.rope_order = info[3];
    e->cfg.return_logits = info[4];
    int32_t hv[7];
    ifb_llama2c_probe(path, hv, nullptr);
    {
        // `devices`: groups of one device each = layer parallelism, the groups run one after the other as the reference's
        // workers do (Infer_LayerParallelism, inference_engine.cc:1185-1220). Groups of several devices (tensor / hybrid
        // parallelism) are one PROCESS per GPU in this runtime: ifb_rank_plan + ifb_model_create + ifb_model_tp_*.
        Resolved r;
        resolve_config(ini_path, section, data_root_dir, r);
        int32_t ids[64], ng = 0, gs = 0, strat = 0;
        rc = ifb_device_groups_parse(r.devices.c_str(), ids, 64, &ng, &gs, &strat);
        if (rc != IFB_OK) { delete e; return rc; }
        if (ng == 0) { ids[0] = 0; ng = 1; gs = 1; }          // no `devices` item: the current device... id 0
        if (gs > 1) {
            set_last_error("engine: devices = '%s' asks for tensor-parallel groups of %d; this runtime runs one process per GPU for "
                           "those (ifb_rank_plan, ifb_model_create with tp_rank/tp_size, ifb_model_tp_export/import)", r.devices.c_str(), gs);
            delete e;
            return IFB_ERR_ARG;
        }
        int32_t ranges[128], gof[64], nr = 0;
        ifb_split_layers(0, hv[2], ng, ranges, gof, &nr);
        int ndev = 0;
        cudaGetDeviceCount(&ndev);
        for (int i = 0; i < nr; i++) {
            Stage st;
            st.layer_begin = ranges[2 * i];
            st.layer_end = ranges[2 * i + 1] < hv[2] ? ranges[2 * i + 1] : hv[2];
            if (st.layer_end <= st.layer_begin) continue;
            st.device = ids[gof[i]];
            if (st.device < 0 || st.device >= ndev) {
                set_last_error("engine: devices = '%s' names device %d; this process sees %d", r.devices.c_str(), st.device, ndev);
                ifb_engine_destroy(e);
                return IFB_ERR_ARG;
            }
            e->stages.push_back(st);
        }
        if (e->stages.size() == 1) { e->stages[0].layer_begin = 0; e->stages[0].layer_end = 0; }   // the whole model
        for (auto &st : e->stages) {
            cudaSetDevice(st.device);
            rc = ifb_model_load_llama2c_stage(&st.model, path, info[0], info[1], info[2], info[3], st.layer_begin, st.layer_end, nullptr);
            if (rc != IFB_OK) { ifb_engine_destroy(e); return rc; }
        }
        e->model = e->stages.back().model;
    }
    e->d_model = hv[0];
    e->vocab = hv[5];
    e->max_ctx = info[2];
    e->logits_h.resize((size_t)e->vocab);
    *out = e;
    return IFB_OK;
}

void ifb_engine_destroy(ifb_engine_t *e) {
    if (!e) return;
    for (auto &st : e->stages)
        if (st.model) { cudaSetDevice(st.device); ifb_model_destroy(st.model); }
    delete e;
}

int32_t ifb_engine_vocab_size(const ifb_engine_t *e) { return e ? e->vocab : 0; }

// AddQuery(tokens, greedy): > 0 = query id, 0 = busy (one live query in this engine), < 0 = error
int32_t ifb_engine_add_query(ifb_engine_t *e, const int32_t *tokens, int32_t n) {
    IFB_CHECK_ARG(e != nullptr && tokens != nullptr && n > 0, "add_query: empty query");
    if (e->active) return 0;
    IFB_CHECK_ARG(n < e->max_ctx, "add_query: %d tokens do not fit max_context_len %d", n, e->max_ctx);
    for (int i = 0; i < n; i++)
        IFB_CHECK_ARG(tokens[i] >= 0 && tokens[i] < e->vocab, "add_query: token id %d outside the vocabulary", tokens[i]);
    for (auto &st : e->stages) {
        cudaSetDevice(st.device);
        int rc = ifb_model_kv_reset(st.model, 0, 0, 0, nullptr);
        if (rc != IFB_OK) return rc;
    }
    e->to_feed.assign(tokens, tokens + n);
    e->pos = 0;
    e->active = true;
    return ++e->query_id;
}

// Infer: one step for the live query -- the whole prompt on the first call, the committed token afterwards.
// next_token: the greedy choice; logits_f32 (optional, vocab floats): the step's output tensor.
int32_t ifb_engine_infer(ifb_engine_t *e, int32_t *next_token, float *logits_f32) {
    IFB_CHECK_ARG(e != nullptr && next_token != nullptr, "infer: null argument");
    if (!e->active || e->to_feed.empty()) { set_last_error("infer: no query is waiting for a step"); return IFB_ERR_STATE; }
    if (e->pos + (int)e->to_feed.size() > e->max_ctx) { set_last_error("infer: the context is full (%d tokens)", e->max_ctx); return IFB_ERR_STATE; }
    if (e->stages.size() > 1) {
        // layer parallelism: a token passes the stages in order; the hidden state travels from one stage's buffer to the
        // next one's (GetOutput -> SetInput of the next worker, inference_engine.cc:1196-1212)
        while (!e->to_feed.empty()) {
            for (size_t k = 0; k < e->stages.size(); k++) {
                Stage &st = e->stages[k];
                cudaSetDevice(st.device);
                if (k > 0) {
                    Stage &pv = e->stages[k - 1];
                    cudaError_t ce = pv.device == st.device
                        ? cudaMemcpy(ifb_model_hidden_ptr(st.model, 0), ifb_model_hidden_ptr(pv.model, 1), (size_t)e->d_model * 2, cudaMemcpyDeviceToDevice)
                        : cudaMemcpyPeer(ifb_model_hidden_ptr(st.model, 0), st.device, ifb_model_hidden_ptr(pv.model, 1), pv.device, (size_t)e->d_model * 2);
                    if (ce != cudaSuccess) { set_last_error("infer: hidden-state hand-off failed: %s", cudaGetErrorString(ce)); return IFB_ERR_CUDA; }
                }
                int rc = ifb_model_decode(st.model, e->to_feed.front(), nullptr, e->pos, nullptr);
                if (rc == IFB_OK) rc = ifb_device_synchronize();      // the groups run one after the other
                if (rc != IFB_OK) return rc;   // a retry repeats this token on every stage: each rewrites the same cache rows
            }
            e->to_feed.erase(e->to_feed.begin());
            e->pos++;
        }
    } else if (e->to_feed.size() > 1 && ifb_model_prefill_supported(e->model)) {
        // the prompt as ONE step with q_tokens = n, as the reference runs it (tensor-core GEMMs + prefill attention)
        int rc = ifb_model_prefill(e->model, e->to_feed.data(), (int32_t)e->to_feed.size(), e->pos, nullptr);
        if (rc != IFB_OK) return rc;        // nothing was consumed: pos and to_feed unchanged, a retry repeats the same step
        e->pos += (int)e->to_feed.size();
        e->to_feed.clear();
    } else {
        while (!e->to_feed.empty()) {
            int rc = ifb_model_decode(e->model, e->to_feed.front(), nullptr, e->pos, nullptr);
            if (rc != IFB_OK) return rc;    // tokens already fed are gone from to_feed: a retry continues where this stopped
            e->to_feed.erase(e->to_feed.begin());
            e->pos++;
        }
    }
    int rc = ifb_device_synchronize();
    if (rc != IFB_OK) return rc;
    rc = ifb_memcpy_d2h(next_token, ifb_model_next_token(e->model), 4);
    if (rc != IFB_OK) return rc;
    if (logits_f32 != nullptr) {
        rc = ifb_memcpy_d2h(e->logits_h.data(), ifb_model_logits(e->model), (int64_t)e->vocab * 2);
        if (rc != IFB_OK) return rc;
        for (int i = 0; i < e->vocab; i++) {
            __half h = __ushort_as_half(e->logits_h[i]);
            logits_f32[i] = __half2float(h);
        }
    }
    return IFB_OK;
}

// CommitInferenceResult: the token the caller chose for the query (greedy callers pass Infer's token back)
int32_t ifb_engine_commit(ifb_engine_t *e, int32_t token, int32_t is_end) {
    IFB_CHECK_ARG(e != nullptr, "commit: null engine");
    if (!e->active) { set_last_error("commit: no live query"); return IFB_ERR_STATE; }
    IFB_CHECK_ARG(token >= 0 && token < e->vocab, "commit: token id %d outside the vocabulary", token);
    if (is_end) { e->active = false; e->to_feed.clear(); return IFB_OK; }
    e->to_feed.assign(1, token);
    return IFB_OK;
}

int32_t ifb_engine_remove_query(ifb_engine_t *e, int32_t query_id) {
    IFB_CHECK_ARG(e != nullptr, "remove_query: null engine");
    if (!e->active || query_id != e->query_id) { set_last_error("remove_query: unknown query %d", query_id); return IFB_ERR_ARG; }
    e->active = false;
    e->to_feed.clear();
    return IFB_OK;
}

}  // extern "C"
