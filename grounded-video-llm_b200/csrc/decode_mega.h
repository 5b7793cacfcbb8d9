// Single-kernel decode step (decode_mega.cu): plan structures shared with model_lm.cu.
#pragma once
#include "decode.h"

namespace gvl {

constexpr int MEGA_ROWS = 8;        // weight rows per GEMV unit (= the n extent of one mma.m16n8k16)
constexpr int MEGA_SEG = 512;       // max k-elements of one ring item row (1 KB)

struct MegaOp {                 // one weight-streaming GEMV phase (M = 1)
    const __nv_bfloat16* W;     // PACKED copy (decode_mega_pack): [unit][sel][seg] items of 8 rows x seg_len k, each item
                                // [32-k chunk][row][32 k]; source [n_rows, K] row-major (gate_up: 128 gate / 128 up rows
                                // interleaved per 256-row block)
    int K, nseg, seg_len;       // K = nseg * seg_len, seg_len <= MEGA_SEG, seg_len % 64 == 0
    int n_rows;                 // weight rows of the source (rows beyond are zero in the packed copy)
    int n_out;                  // output columns (act == 3: n_rows / 2)
    int units;                  // ceil(n_out / 8)
    int act;                    // 0 none, 3 SwiGLU
    int out_f32;
    int argmax;                 // 1: fold the greedy pick into this phase (lm_head)
    int x_kind;                 // 0: vector in global memory (+ optional RMSNorm), 1: merge of the attention partials
    int from_embed;             // bit 0: x is the embedding row of the current token, bit 1: so is the residual (layer 0)
    const __nv_bfloat16* x;     // input vector [K] (global)
    const __nv_bfloat16* norm_w;
    float eps;
    const __nv_bfloat16* bias;
    const __nv_bfloat16* residual;
    void* out;
};

constexpr int MEGA_MAX_LAYERS = 48;
constexpr int MEGA_TRACE_STRIDE = 1024;  // clock64 marks per CTA per step: 3 + 15 * layers + 3 (layers <= 48 -> 726)
constexpr int MEGA_TRACE_OCC_OFF = 768;  // + 4 GEMV phases of the last layer x 8 warps: ring slots already landed at phase start;
                                         // + 32: cycles the warp waited for ring items in that phase; + 64: items it consumed;
                                         // + 96: 3 staging steps x 4 marks (stride 8); + 128: 4 GEMV phases x 4 marks (stride 8)
constexpr int MEGA_TRACE_ATT_OFF = 960;  // + 8 warps x 8 marks inside the attention phase of the last layer;
                                         // [ATT_OFF - 2], [ATT_OFF - 1]: %globaltimer (ns) at kernel start / end

struct MegaPlan {
    int n_layers, dim, heads, kv_heads, head_dim, vocab, max_ctx;
    int x_bytes;                // activation staging area (also the attention-phase scratch)
    int part_items;             // capacity of the per-item partial-sum buffer (8 floats per item)
    int att_maxp;               // split-KV partials per head in att_ws
    float scale;
    MegaOp ops[4 * MEGA_MAX_LAYERS + 1];
    const __nv_bfloat16* embed;
    const __nv_bfloat16 *rope_cos, *rope_sin;
    __nv_bfloat16* kv;          // [L][2][KVH][max_ctx][hd]
    __nv_bfloat16 *x, *qkv, *mid;
    float* logits;
    float* att_ws;              // [heads][att_maxp][hd + 4] split-KV partials (m, l, -, -, o[hd])
    unsigned long long* amax;   // packed (orderable logit, ~index) of the greedy pick; zero between steps
    unsigned* grid_bar;
    DecodeState* st;
    long long* trace;           // optional [gridDim.x][MEGA_TRACE_STRIDE] phase timestamps of the LAST step (bring-up / profiling)
};

// fills nseg / seg_len / n_out / units of an op from K, n_rows and act; false when the shape is not supported
bool decode_mega_shape(MegaOp* op);
// packed copy of one weight matrix (elements; device-side repack of W [n_rows, ldw])
size_t decode_mega_packed_elems(const MegaOp* op);
int decode_mega_pack(const MegaOp* op, const __nv_bfloat16* W, int ldw, __nv_bfloat16* dst, cudaStream_t s);
// after all ops are set: picks x_bytes / part_items / att_maxp; false when the step does not fit
bool decode_mega_finalize(MegaPlan* plan);
size_t decode_mega_att_ws_bytes(const MegaPlan* plan);
// host view of the attention-phase split (att_split in decode_mega.cu): warps per head, tokens per warp, partial records per head
void decode_mega_attention_split(int ctx, int heads, int n_ctas, int* warps_per_head, int* tokens_per_warp, int* max_partials);
int decode_mega_launch(const MegaPlan* plan_host, const MegaPlan* plan_dev, int n_steps, long long* tokens_out, float* logits_out,
                       long long eos_id, long long pad_id, cudaStream_t s);

}  // namespace gvl
