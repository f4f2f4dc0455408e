// Persistent decode-step kernel: ONE launch per token (ifb_step.cu). Internal interface between the kernel and the
// model runtime (ifb_model.cu); the public surface stays include/inferflow_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace ifb {

// op kinds of one pre-norm llama layer, in execution order, + the lm_head
enum { OPK_QKV = 0, OPK_ATT = 1, OPK_WO = 2, OPK_GLU = 3, OPK_W2 = 4, OPK_HEAD = 5, OPK_COUNT = 6 };
enum { STEP_PART_GEMV = 1, STEP_PART_ATT = 2, STEP_PART_ALL = 3 };

// error codes the kernel leaves in StepParams::err[0] before it traps
enum { STEP_ERR_NONE = 0, STEP_ERR_RING = 1, STEP_ERR_DEP = 2, STEP_ERR_PEER = 3, STEP_ERR_LOGITS = 4, STEP_ERR_CTX = 5 };

// shape of one op kind (the same for every layer)
struct StepOpShape {
    int units;        // work units per instance (ATT: derived on the device from the context length)
    int tasks;        // tasks per unit: a task is what ONE warp does (2 adjacent rows / one w1,w3 row pair / 32 positions)
    int R;            // rows (row pairs for GLU) per unit
    int cx, nblk, row_bytes;
    int unit_bytes;   // bytes of ONE matrix per unit (R * row_bytes); GLU units carry two matrices
    int seg1, seg2;   // QKV: first unit of the wk / wv segment
    int vec4;         // Q4_B32T1 "lane owns 4 blocks" path (cx % 128 == 0)
    int nxb;          // staged activation blocks (cx/32; padded to a multiple of 128 for vec4)
};

// geometry of one op kind for the warp-private pipelines (decode_step_kernel<Fmt, true>): a task = rt adjacent rows,
// streamed as nc K-chunks of kc_blk blocks (kc_bytes per row) into the warp's own two-slot ring
struct StepPrivShape {
    int rt;           // rows per task (2; 4 when two rows are not a multiple of the 16-byte bulk-copy granularity)
    int nc;           // K-chunks per row
    int kc_blk;       // weight blocks per chunk and row (fp16 weights: halves)
    int kc_bytes;     // bytes per chunk and row
    int ntasks;       // tasks of one instance of the op (ATT: derived on the device from the context length)
};

struct StepLayer {
    const uint8_t *wq, *wk, *wv, *wo, *w1, *w3, *w2;
    const __half *attn_norm, *ffn_norm;
    uint8_t *kcache, *vcache;
    const void *pad_;            // 96 bytes: the table is copied to shared memory with 16-byte accesses
};
static_assert(sizeof(StepLayer) == 96, "StepLayer must stay a multiple of 16 bytes");

struct StepParams {
    StepOpShape op[OPK_COUNT];
    const StepLayer *layers;     // device array [n_layers]
    int n_layers, n_ops;         // n_ops = 5 * n_layers + 1
    int d, q_dim, kv_dim, ffn, heads, kv_heads, hd, group, vocab, vocab_l;   // this rank's shard sizes
    int max_ctx, max_ctx_pad, kv_q8, rope_order;
    int w_dtype, head_f16;
    float eps, softmax_scale, rope_theta;
    const __half *tok_emb, *out_norm;
    const uint8_t *lm_head;
    __half *qkv, *attn_out, *hbuf, *ybuf, *logits, *x_out;
    uint2 *qkvw, *attw, *hbufw;  // q|k|v of the layer in hand, the attention output and h = silu(w1 x)*(w3 x) as flag-in-data words {half2, flag}
    float *partials;             // attention partials [heads][max_tiles][hd + 4]: m, l, -, -, acc[hd]
    int max_tiles;
    int32_t *scalars;            // [0] position [1] token [2] next token [3] step sequence number
    int *cursors;                // [n_ops] unit cursors               } one allocation, zeroed before every launch
    unsigned int *done;          // [n_ops] + merge_done[n_layers] + 1 }
    unsigned long long *amax;    // argmax key                         }
    // shared-memory layout (bytes from the start of dynamic shared memory)
    int nslots, slot_bytes, x_off, resid_off, sp_off, rope_off, lay_off, norm_off, smem_bytes;
    int att_pos;                 // positions per ATT unit (64 for a Q8_B32T2 cache, 32 for fp16)
    int parts, advance;
    int debug;                   // measurement aid (IFB_STEP_DEBUG): bit 0 = the tasks skip their arithmetic (pure streaming + boundaries)
    int claim_batch;             // most units one cursor claim takes (guided self-scheduling)
    int max_inflight;            // bulk copies outstanding per SM (0: as many as the ring has slots)
    // warp-private pipelines: every consumer warp streams its own, statically assigned tasks through two private slots
    int priv;                    // 1: this layout (no producer warp, no unit cursors)
    int pslot;                   // bytes per private slot
    StepPrivShape pv[OPK_COUNT];
    // tensor parallelism (size 1: unused) -- exchange regions as in ifb_model.cu
    int tp_size, tp_rank;
    int local_ll;                // one GPU: wo / w2 outputs as flag-in-data words in this GPU's own exchange region
    void *peer[8];
    long long x_recv, x_recv_parity, x_recv_rank, x_logits;
    unsigned long long *trace;   // optional [n_ops][grid][4] globaltimer stamps
    int *err;                    // pinned host memory: [0] code [1] op [2] detail [3] cta
    unsigned long long timeout_ns;
};

// Fills op[], the shared-memory layout, att_pos, max_tiles from the dims / dtypes already set in P.
// Returns IFB_OK, or IFB_ERR_ARG (message in ifb_last_error) when this model cannot run as one persistent kernel
// (the caller then keeps the multi-launch path).
int step_build(StepParams &P, int kv_dtype, int head_dtype, int sm_count_);
int step_launch(const StepParams &P, int grid, cudaStream_t st);
int64_t step_counter_bytes(const StepParams &P);   // cursors + done + amax

}  // namespace ifb
