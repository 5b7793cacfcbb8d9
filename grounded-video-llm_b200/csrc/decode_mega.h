// Plan (device-resident) of the single-kernel decode step, see decode_mega.cu.
#pragma once
#include "decode.h"

namespace gvl {

struct MegaOp {                 // one weight-streaming GEMV phase (M = 1)
    const __nv_bfloat16* W;     // [N, K] row-major (gate_up: interleaved per 256-row block)
    int ldw, K, nseg, seg_len;  // K = nseg * seg_len, seg_len <= 4096
    int units;                  // output columns (act == 3: N / 2)
    int act;                    // 0 none, 3 SwiGLU
    int out_f32;
    const __nv_bfloat16* x;     // input vector [K] (global)
    const __nv_bfloat16* norm_w;
    float eps;
    const __nv_bfloat16* bias;
    const __nv_bfloat16* residual;   // may alias out
    void* out;
};

constexpr int MEGA_MAX_LAYERS = 48;
constexpr int MEGA_TRACE_STRIDE = 768;   // clock64 marks per CTA per step: 1 + 15 * layers + 3

struct MegaPlan {
    int n_layers, dim, heads, kv_heads, head_dim, vocab, max_ctx;
    float scale;
    MegaOp ops[4 * MEGA_MAX_LAYERS + 1];
    const __nv_bfloat16* embed;
    const __nv_bfloat16 *rope_cos, *rope_sin;
    __nv_bfloat16* kv;          // [L][2][KVH][max_ctx][hd]
    __nv_bfloat16 *x, *qkv, *attn_out, *mid;
    float* logits;
    float* att_ws;              // [heads][max_ctx/128][hd+2]
    int* att_counters;          // [heads], zero between launches
    unsigned* grid_bar;
    DecodeState* st;
    long long* trace;           // optional [gridDim.x][MEGA_TRACE_STRIDE] phase timestamps of the LAST step (bring-up / profiling)
};

size_t decode_mega_smem();
int decode_mega_launch(const MegaPlan* plan_dev, unsigned* grid_bar, long long* tokens_out, float* logits_out,
                       long long eos_id, long long pad_id, cudaStream_t s);

}  // namespace gvl
