/* gvl.h -- C ABI of libgvl.so: the B200 (sm_100a) forward path of Grounded-VideoLLM.
 *
 * The reference (WHB139426/Grounded-Video-LLM) has no FFI layer: its operator API on this path is
 * nn.Module.forward. Each entry point below names the reference function(s) it
 * replaces (file:line relative to the reference repo). Plain pointers and sizes only; all pointers
 * are DEVICE pointers unless stated otherwise; `stream` is a cudaStream_t passed as void*.
 * Every function returns 0 (GVL_OK) or a negative gvl_status; nothing throws across the ABI.
 * CUDA errors are sticky: the Python mirror raises RuntimeError.
 */
#ifndef GVL_H_
#define GVL_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    GVL_STATUS_OK = 0,
    GVL_STATUS_ERR_ARG = -1,    /* bad shape / unsupported combination  (reference: ValueError / assert) */
    GVL_STATUS_ERR_ALIGN = -2,  /* pointer or leading dimension not 16-byte aligned */
    GVL_STATUS_ERR_CUDA = -3,   /* launch failed (cudaGetLastError) */
    GVL_STATUS_ERR_DRIVER = -4, /* cuTensorMapEncodeTiled unavailable / failed */
    GVL_STATUS_ERR_NOMEM = -5,
    GVL_STATUS_ERR_STATE = -6
} gvl_status;

/* activation / residual selectors of the fused GEMM epilogue */
enum { GVL_ACT_NONE = 0, GVL_ACT_GELU_ERF = 1, GVL_ACT_QUICK_GELU = 2, GVL_ACT_SWIGLU = 3 };
enum { GVL_RES_NONE = 0, GVL_RES_BF16 = 1, GVL_RES_F32 = 2 };

const char* gvl_version(void);
/* number of kernels this library has launched since load (bench.py "gpu_launches") */
long long gvl_launch_count(void);

/* ------------------------------------------------------------------ operator level
 * out = epilogue(A[M,K] @ W[N,K]^T), bf16 operands, fp32 accumulation in TMEM (tcgen05.mma).
 * Replaces torch.nn.functional.linear under bf16 autocast at every call site listed in SURVEY.md 2.1
 * "Dense linears" (modeling_clip.py:244-247,336-337; internvideo2.py:549-551,624-627;
 * llava_next_video.py:31-32,46-47; modeling_phi3.py:453-454,513-514) and the patch-embed convs
 * (modeling_clip.py:185; internvideo2.py:722) through gvl_im2col_patch14.
 *   bias     bf16 [N] or NULL         gamma  fp32 [N] LayerScale (internvideo2.py:451-466) or NULL
 *   act      GVL_ACT_*  (SWIGLU: W rows interleaved gate/up per 256-row block, output has N/2 columns)
 *   res      GVL_RES_*  residual added after activation/LayerScale; out_f32 selects fp32 output
 *   bn_hint  0 = auto, 128 or 256 = force the N tile                                                */
int gvl_gemm_bf16(const void* A, int lda, const void* W, int ldw, void* out, int ldo, int M, int N, int K,
                  const void* bias, const float* gamma, const void* residual, int ldr, int act, int res,
                  int out_f32, int bn_hint, void* stream);

/* softmax(Q K^T * scale) V, bf16 in/out, fp32 softmax; strides in elements.
 * Replaces CLIPAttention bmm/softmax/bmm (modeling_clip.py:274-314, round_scores=1),
 * FlashAttention.forward (internvideo2.py:493-538), _flash_attention_forward (modeling_phi3.py:778-876,
 * modeling_llama.py:537-594; causal=1, bottom-right aligned). GQA via kv_heads < heads.            */
int gvl_attention(const void* q, const void* k, const void* v, void* o,
                  const long long* q_strides /*[3] batch,token,head*/, const long long* k_strides,
                  const long long* v_strides, const long long* o_strides, int batch, int heads, int kv_heads,
                  int sq, int skv, int head_dim, float scale, int causal, int round_scores, int o_dim /*0 = head_dim*/,
                  void* stream);

/* nn.LayerNorm on an fp32 stream, output rounded to bf16 (modeling_clip.py:351-353, 824-826). */
int gvl_layernorm_f32(const float* x, const float* w, const float* b, void* y_bf16, int rows, int cols, float eps,
                      void* stream);
/* RMSNorm.forward (internvideo2.py:437-448) / Phi3RMSNorm (modeling_phi3.py:310-324) / LlamaRMSNorm
 * (modeling_llama.py:74-88); bf16 in/out, ldx/ldy row strides in elements.                        */
int gvl_rmsnorm_bf16(const void* x, long long ldx, const void* w, void* y, long long ldy, int rows, int cols,
                     float eps, void* stream);
/* q_norm / k_norm over the flattened (heads*head_dim) q and k rows, in place in qkv[rows,3*dim]
 * (internvideo2.py:590-598).                                                                      */
int gvl_iv2_qk_rmsnorm(void* qkv, const void* wq, const void* wk, int rows, int dim, float eps, void* stream);

/* im2col for the stride-14 patch convs; pix [n_img, chans, frames, hw, hw] (fp32 or bf16),
 * out bf16 [n_img*frames*(hw/14)^2, kpad], k = c*196+ky*14+kx.                                    */
int gvl_im2col_patch14(const void* pix, int pix_is_f32, void* out, int n_img, int chans, int frames, int hw,
                       int kpad, void* stream);
/* CLIPVisionEmbeddings tail: cat(cls, patches)+pos -> fp32 (modeling_clip.py:187-190). */
int gvl_clip_assemble(const void* patch_bf16, const float* cls, const float* pos, float* x, int n_img,
                      int n_patch, int dim, void* stream);
/* PretrainInternVideo2.forward prologue: cat(cls, patches)+pos_embed -> bf16 (internvideo2.py:975-1005). */
int gvl_iv2_assemble(const void* patch, const void* cls, const void* pos, void* x, int n_seg, int n_patch, int dim,
                     void* stream);
/* reshape_hd_patches_2x2merge_phi3 + add_image_newline_phi3 (llava_next_video.py:454-489). */
int gvl_hd_merge_newline(const float* hs, const float* sub_gn, void* out_bf16, int n_img, void* stream);
/* AdaptiveAvgPool3d([T,4,4]) of the temporal stream (llava_next_video.py:544-549). */
int gvl_iv2_pool(const void* x, void* out, int n_seg, int frames, int dim, void* stream);
/* AdaptiveAvgPool3d([segs,8,8]) of the spatial stream, Llama variant (llava_next_video.py:509-517). */
int gvl_clip_pool3(const float* hs, void* out_bf16, int n_img, void* stream);
/* per-segment stream concat [spatial a_rows | temporal b_rows | newline 1] (llava_next_video.py:563-564). */
int gvl_visual_concat(const void* a, int a_rows, const void* b, int b_rows, const void* newline, void* out, int n_seg,
                      int dim, void* stream);
/* nn.LayerNorm with fp32 output: CLIP pre_layrnorm (modeling_clip.py:851). x and y must not alias. */
int gvl_layernorm_f32_out_f32(const float* x, const float* w, const float* b, float* y, int rows, int cols, float eps,
                              void* stream);
/* prepare_multimodal_inputs (llava_next_video.py:568-596); ids int64 with the -200 sentinel at img_pos. */
int gvl_embed_splice(const long long* ids, int t_text, int img_pos, const void* table, const void* visual,
                     int n_vis, void* out, int dim, int vis_last, void* stream);
/* apply_rotary_pos_emb + KV-cache append (modeling_phi3.py:413-445, :721; modeling_llama.py:173-204, :451). */
int gvl_rope_qkv_cache(const void* qkv, void* q_out, void* k_cache, void* v_cache, const void* cos_bf16,
                       const void* sin_bf16, const int* positions, int tokens, int heads, int kv_heads,
                       int head_dim, int pos0, int max_ctx, void* stream);

/* ------------------------------------------------------------------ frame preprocessing
 * frame_transform (mm_utils/utils.py:153-183; call sites inference.py:69-88): ToPILImage -> Resize(size, BICUBIC) ->
 * CenterCrop(size) -> ToTensor -> Normalize on uint8 frames [n,3,h,w] -> float32 [n,3,size,size], bit-exact with
 * Pillow's 8-bit bicubic resampling (fixed-point two-pass convolution) and torchvision's float32 arithmetic.
 * new_h / new_w = torchvision's shortest-edge resize result, crop_top / crop_left = its center-crop offsets (both are
 * computed by the caller with the reference's exact integer rules, gvl/preprocess.py). mean3 / std3 are HOST pointers.
 * workspace: gvl_frame_transform_workspace bytes of device memory (the 8-bit intermediate image; 0 when w == new_w). */
size_t gvl_frame_transform_workspace(int n, int h, int w, int new_h, int new_w);
int gvl_frame_transform(const unsigned char* frames, int n, int h, int w, int new_h, int new_w, int crop_top,
                        int crop_left, int size, const float* mean3, const float* std3, float* out, void* workspace,
                        size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------ decode-side operators
 * y[M,N] = x[M,K] @ W[N,K]^T for M <= 8 (weight-streaming, HBM-bound). Optional fused input RMSNorm
 * (norm_w != NULL), bias, SwiGLU (same interleaved W as gvl_gemm_bf16), bf16 residual, fp32 output. */
int gvl_gemv_bf16(const void* x, int ldx, const void* W, int ldw, void* out, int ldo, int M, int N, int K,
                  const void* norm_w, float norm_eps, const void* bias, const void* residual, int ldr, int act,
                  int out_f32, void* stream);
/* single-query attention against the KV cache, split along the context; ctx_len read from device. */
int gvl_decode_attention(const void* q, const void* k_cache, const void* v_cache, void* o, float* workspace,
                         const int* ctx_len_dev, int heads, int kv_heads, int head_dim, int max_ctx, float scale,
                         void* stream);
size_t gvl_decode_attention_workspace(int heads, int head_dim, int max_ctx);
/* greedy next token: argmax over fp32 logits[n] -> out_token[0] (first max index, torch.argmax order). */
int gvl_argmax_f32(const float* logits, int n, long long* out_token, void* stream);

/* ------------------------------------------------------------------ model level (stage entry points)
 * Weight tables are plain structs of device pointers, filled by the host mirror from the reference
 * modules' state_dicts (gvl/weights.py). The library borrows them; the caller keeps them alive.    */
typedef struct {
    const void *ln1_w, *ln1_b;   /* fp32 [D]   */
    const void *qkv_w, *qkv_b;   /* bf16 [3D,D] (q rows pre-scaled by head_dim^-0.5), bf16 [3D] */
    const void *out_w, *out_b;   /* bf16 [D,D], [D] */
    const void *ln2_w, *ln2_b;   /* fp32 [D] */
    const void *fc1_w, *fc1_b;   /* bf16 [F,D], [F] */
    const void *fc2_w, *fc2_b;   /* bf16 [D,F], [D] */
} gvl_clip_layer;

typedef struct {
    int n_layers;                /* layers to run (23 for hidden_states[-2], llava_next_video.py:505) */
    int dim, heads, ffn, n_patch, image, kpad;
    const void* patch_w;         /* bf16 [D, kpad] */
    const void* cls;             /* fp32 [D] */
    const void* pos;             /* fp32 [n_patch+1, D] */
    const void *pre_ln_w, *pre_ln_b; /* fp32 [D] */
    const gvl_clip_layer* layers;    /* host array */
} gvl_clip_weights;

/* CLIPVisionModel.forward(...).hidden_states[-2] (modeling_clip.py:830-872, 578-657):
 * pix fp32 [n_img,3,336,336] -> hs fp32 [n_img,577,1024]. workspace from gvl_clip_workspace().    */
size_t gvl_clip_workspace(const gvl_clip_weights* w, int n_img);
int gvl_clip_encode(const gvl_clip_weights* w, const float* pix, float* hs, int n_img, void* workspace,
                    size_t ws_bytes, void* stream);

typedef struct {
    const void* norm1_w;         /* bf16 [D] */
    const void* qkv_w;           /* bf16 [3D,D], no bias */
    const void *q_norm_w, *k_norm_w; /* bf16 [D] */
    const void *proj_w, *proj_b; /* bf16 [D,D], [D] */
    const void* ls1;             /* fp32 [D] */
    const void* norm2_w;
    const void *fc1_w, *fc1_b;   /* bf16 [F,D], [F] */
    const void *fc2_w, *fc2_b;   /* bf16 [D,F], [D] */
    const void* ls2;             /* fp32 [D] */
} gvl_iv2_block;

typedef struct {
    int n_blocks;                /* blocks to run (39: x_vis_return_idx=-2, internvideo2.py:1028-1030) */
    int dim, heads, ffn, frames, kpad;
    int head_dim_pad;            /* head_dim (88) rounded up to a multiple of 32 (96): qkv_w is [3*heads*head_dim_pad, D]
                                    with zero rows at the pad positions, q/k norm weights are [heads*head_dim_pad] with
                                    zeros there, so q/k/v come out of the GEMM already padded for the tcgen05 attention */
    const void *patch_w, *patch_b; /* bf16 [D,kpad], [D] */
    const void* cls;             /* bf16 [D] */
    const void* pos;             /* bf16 [1+frames*256, D] */
    const gvl_iv2_block* blocks; /* host array */
} gvl_iv2_weights;

/* PretrainInternVideo2.forward(x, None, False, x_vis_return_idx=-2, x_vis_only=True)
 * (internvideo2.py:970-1040): pix fp32 [n_seg,3,T,224,224] -> x bf16 [n_seg,1+T*256,1408].       */
size_t gvl_iv2_workspace(const gvl_iv2_weights* w, int n_seg);
int gvl_iv2_encode(const gvl_iv2_weights* w, const float* pix, void* x_out, int n_seg, void* workspace,
                   size_t ws_bytes, void* stream);

typedef struct {
    const void* in_norm_w;       /* bf16 [D] */
    const void* qkv_w;           /* bf16 [(H+2KVH)*hd, D] */
    const void* o_w;             /* bf16 [D, H*hd] */
    const void* post_norm_w;     /* bf16 [D] */
    const void* gate_up_w;       /* bf16 [2F, D], rows interleaved gate/up per 256-row block */
    const void* down_w;          /* bf16 [D, F] */
} gvl_lm_layer;

typedef struct {
    int n_layers, dim, heads, kv_heads, head_dim, ffn, vocab, max_ctx;
    float rms_eps;
    const void* final_norm_w;    /* bf16 [D] */
    const void* lm_head_w;       /* bf16 [vocab, D] */
    const void* lm_head_b;       /* bf16 [vocab] or NULL (reset_embeddings adds a bias, llava_next_video.py:263) */
    const void* embed;           /* bf16 [vocab, D] */
    const void *rope_cos, *rope_sin; /* bf16 [max_ctx, head_dim] */
    const gvl_lm_layer* layers;  /* host array */
} gvl_lm_weights;

typedef struct gvl_lm gvl_lm;    /* opaque: KV cache, workspaces, decode CUDA graph */

/* Phi3ForCausalLM / LlamaForCausalLM forward + greedy generate (modeling_phi3.py:1249-1383, 1466-1551;
 * modeling_llama.py:934-1044, 1165-1256; HF GenerationMixin greedy loop, llava_next_video.py:655-661). */
int gvl_lm_create(const gvl_lm_weights* w, gvl_lm** out);
/* flags: GVL_LM_NO_SINGLE_KERNEL = do not build the packed weight copy of the single-kernel decode step (+7.4 GB for Phi-3.5) for this
 * object -- for the additional per-sequence objects of a batch, which decode through gvl_lm_decode_batch.                              */
#define GVL_LM_NO_SINGLE_KERNEL 1
int gvl_lm_create_ex(const gvl_lm_weights* w, int flags, gvl_lm** out);
void gvl_lm_destroy(gvl_lm* lm);
/* prefill from inputs_embeds bf16 [S,D]; writes fp32 logits of the LAST position to logits_out[vocab]
 * (may be NULL), optionally all-position hidden states for tests via hidden_out bf16 [S,D] (may be NULL),
 * and leaves the greedy next token in the decode state.                                            */
int gvl_lm_prefill(gvl_lm* lm, const void* embeds, int S, float* logits_out, void* hidden_out, void* stream);
/* n greedy decode steps; tokens_out: device int64 [n] (token t = argmax after consuming token t-1);
 * logits_out: optional device fp32 [n, vocab]. eos_id < 0 disables early stop (bench mode);
 * after EOS the remaining slots are filled with pad_id (HF generate semantics).                    */
int gvl_lm_decode(gvl_lm* lm, int n_steps, long long* tokens_out, float* logits_out, long long eos_id,
                  long long pad_id, void* stream);
/* replay decode steps from a captured CUDA graph (default 1) or launch them one by one (0; needed while profiling) */
int gvl_lm_set_graph(gvl_lm* lm, int on);

/* ------------------------------------------------------------------ measurement hooks (bench.py roofline)
 * When enabled, every launch of a kernel family is bracketed by CUDA events on its launching stream.
 * kind: 0 tcgen05 GEMM (work = FLOPs), 1 prefill attention (FLOPs), 2 decode GEMV (weight bytes).        */
int gvl_profile_enable(int on);
int gvl_profile_collect(int kind, double* total_ms, double* total_work, long long* launches);

/* Bring-up / profiling: with GVL_MEGA_TRACE=1 in the environment at gvl_lm_create time the single-kernel decode step
 * records clock64() per CTA at every phase boundary of the LAST step; this copies [n_ctas][stride] int64 marks to the
 * host (synchronises the device). Returns GVL_ERR_STATE when tracing is off.                                          */
int gvl_lm_mega_trace(gvl_lm* lm, long long* host_out, int max_ctas, int* n_ctas, int* stride);

/* Batched greedy decode: n_seq (1..4) sequences, each prefilled on its OWN gvl_lm object created over the SAME weight table (own KV
 * cache / decode state / RoPE table), advance n_steps together; every weight matrix is streamed once per step for all of them.
 * tokens_out: device int64 [n_seq][n_steps]; logits_out: optional device fp32 [n_seq][n_steps][vocab]. Per-sequence semantics are
 * those of gvl_lm_decode (llava_next_video.py:655-661 with a batch of prompts). Runs the per-op kernel chain as one CUDA graph.  */
int gvl_lm_decode_batch(gvl_lm* const* lms, int n_seq, int n_steps, long long* tokens_out, float* logits_out, long long eos_id,
                        long long pad_id, void* stream);

/* Host-side view of how the single-kernel decode step splits its attention phase over the grid (no GPU needed; tests and capacity
 * planning): n_ctas x 8 consumer warps, every head gets warps_per_head of them, warp i of a head works on tokens
 * [i * tokens_per_warp, min(ctx, (i + 1) * tokens_per_warp)); max_partials = the most CTAs any head spreads over (one partial record
 * per (head, CTA) in the merge workspace), capacity = the records per head the workspace holds. Phi3FlashAttention2 / Llama attention
 * with q_len = 1 (modeling_phi3.py:629-775, modeling_llama.py:537-594).                                                              */
int gvl_lm_attention_split(int ctx, int heads, int n_ctas, int* warps_per_head, int* tokens_per_warp, int* max_partials, int* capacity);

/* which decode step this object runs: 1 = the single persistent kernel (decode_mega.cu), 0 = the per-op chain (CUDA graph).
 * The single kernel is the default whenever the shape fits it; GVL_DECODE_MEGA=0 in the environment at create time selects the chain. */
int gvl_lm_decode_kind(const gvl_lm* lm);

/* first generated token (argmax of the prefill logits), device int64 */
const long long* gvl_lm_first_token(gvl_lm* lm);
/* Sampling decode (HF generate with do_sample=True, llava_next_video.py:655-661 / inference.py:170-176): the caller picks the
 * next token itself from the logits of the previous step (temperature / top-k / top-p / multinomial happen outside, on the
 * device) and hands it back before the next gvl_lm_decode(lm, 1, ...) call. token_dev: ONE device int64. */
int gvl_lm_set_next_token(gvl_lm* lm, const long long* token_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GVL_H_ */
