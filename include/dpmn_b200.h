/*
 * dpmn_b200 -- C ABI of the B200-native DPMN hot path (PGRM stack + Complementation Modulation Module).
 *
 * The reference (jdfxzzy/DPMN) is pure PyTorch: it has no FFI/plugin layer, its boundary for this path
 * is the nn.Module ctor + forward signature + state_dict schema.  This header is the
 * C-ABI a binding sits on; each entry point names the reference interface it replaces:
 *
 *   dpmn_pgrm_forward          <- PGRM.forward(x_q, x_kv, residual_list)            model/pgrm.py:546-565
 *   dpmn_pgrm_forward_probe    <- same, exposing WindowAttention / SwinTransformerBlock outputs
 *                                                                                   model/pgrm.py:184-271,315-331
 *   dpmn_window_attn_forward   <- WindowAttention.forward, per-group core           model/pgrm.py:197-268
 *   dpmn_cmm_forward           <- ComplementationModulationModule.forward(x1, x2)   model/cmm.py:120-161
 *   dpmn_gemm_nt               <- nn.Linear / 1x1 conv contraction (F.linear)       model/pgrm.py:30,37,39,82,188,194
 *   dpmn_pgrm_backward         <- autograd backward of PGRM.forward                 interfaces/super_resolution.py:245-275
 *   dpmn_cmm_backward          <- autograd backward of the CMM forward              interfaces/super_resolution.py:245-275
 *
 * Conventions
 *   - every function returns int: 0 = ok, <0 = argument error (DPMN_E_*), >0 = a cudaError_t value
 *   - all pointers are DEVICE pointers; tensors are fp32 in the layouts stated per argument (NCHW images,
 *     exactly as the reference passes them); image tensors may carry a batch stride (channel-slice views
 *     such as cascade[:, :3], interfaces/super_resolution.py:196)
 *   - the library never allocates or frees device memory and never synchronises: the caller passes a
 *     workspace of at least dpmn_*_workspace_bytes() and a stream (cudaStream_t as void*)
 *   - weights are read in place in the reference's own state_dict layout; nothing is cached across calls
 *   - one exception to "the caller's stream only": dpmn_cmm_backward and the fp32-structured dpmn_cmm_forward run the
 *     CMM's second encoder branch (independent of the first until the SE gate, cmm.py:121-133) on a library-owned
 *     non-blocking side stream per device, forked from and joined to the caller's stream by events inside the call.  The
 *     caller still sees plain stream semantics (the call's work is ordered between what the caller enqueued before and
 *     after it; legal under stream capture).  DPMN_CMM_FORK=0 in the environment keeps everything on the caller's stream.
 */
#ifndef DPMN_B200_H
#define DPMN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DPMN_MAX_GROUPS 4   /* window groups per attention (reference default 3: windows 2,4,8) */
#define DPMN_MAX_MIX 8      /* weight_list_0 .. weight_list_iter (reference uses iter <= 5) */
#define DPMN_MAX_BLOCKS 2   /* BasicLayer depth is hard-wired to 2 (pgrm.py:506) */

enum {
  DPMN_OK = 0,
  DPMN_E_ARG = -1,        /* null pointer / inconsistent sizes */
  DPMN_E_UNSUPPORTED = -2,/* configuration the reference itself cannot run, or outside this build */
  DPMN_E_WORKSPACE = -3,  /* workspace too small */
  DPMN_E_DEVICE = -4      /* not an sm_100 device / driver entry point missing */
};

/* arithmetic of the GEMM / attention contractions.  Residual stream, LayerNorm statistics, softmax
 * and all accumulation are fp32 in every mode. */
enum {
  DPMN_PREC_F32 = 0,      /* fp32 FFMA everywhere: matches the reference to 1e-5 rel */
  DPMN_PREC_F16 = 1,      /* fp16 operands on tcgen05 tensor cores, fp32 accumulate: 1e-3 rel */
  DPMN_PREC_BF16 = 2      /* bf16 operands on tcgen05 tensor cores, fp32 accumulate */
};

/* one SwinTransformerBlock (pgrm.py:290-331); names follow the state_dict keys under
 * layers.0.blocks.<b>. ; every pointer is the fp32 tensor in torch's own layout. */
typedef struct dpmn_block_weights {
  const float *norm1_q_w, *norm1_q_b, *norm1_kv_w, *norm1_kv_b;       /* (C) */
  const float *rpb_table[DPMN_MAX_GROUPS];  /* attn.relative_position_bias_table_g ((2ws-1)^2, heads/G) */
  const float *q_w, *q_b;                   /* attn.q   (C, C), (C) */
  const float *kv_w, *kv_b;                 /* attn.kv  (2C, C), (2C) */
  const float *sk_proj_w, *sk_proj_b;       /* attn.sknet.proj       (C, C) */
  const float *sk_fc1_w, *sk_fc1_b;         /* attn.sknet.fc1        (C/G/2, C) */
  const float *sk_fc2_w, *sk_fc2_b;         /* attn.sknet.fc2        (C, C/G/2) */
  const float *sk_head_w, *sk_head_b;       /* attn.sknet.proj_head  (C, C/G) */
  const float *norm2_w, *norm2_b;           /* (C) */
  const float *fc1_w, *fc1_b;               /* mlp.fc1 (hid, C) */
  const float *fc2_w, *fc2_b;               /* mlp.fc2 (C, hid) */
  const float *dw_w, *dw_b;                 /* mlp.depthwise_conv (hid, 1, 3, 3) */
  const float *pw_w, *pw_b;                 /* mlp.pointwise_conv (hid, hid, 1, 1) */
} dpmn_block_weights;

/* one PGRM (pgrm.py:460-565) */
typedef struct dpmn_pgrm_desc {
  int32_t batch;            /* B */
  int32_t img_h, img_w;     /* 32, 128 */
  int32_t patch;            /* patch_size[iter] (2) */
  int32_t q_chans;          /* channels of x_q: 2 (prior_fusion runs, pgrm.py:547-548) or 3 */
  int32_t embed_dim;        /* C (96) */
  int32_t num_heads;        /* 6 */
  int32_t n_groups;         /* len(window_size[iter]) */
  int32_t window[DPMN_MAX_GROUPS];   /* configured windows (2,4,8); clamp/shift rules applied inside */
  int32_t mlp_hidden;       /* int(C * mlp_ratio) (384) */
  int32_t hidden_size;      /* output channels (3) */
  int32_t precision;        /* DPMN_PREC_* */
  int32_t n_mix;            /* number of terms of the final affine mix = max(1, len(residual_list)) */
  int64_t x_q_batch_stride; /* elements between images of x_q / x_kv; 0 = dense (chans*img_h*img_w) */
  int64_t x_kv_batch_stride;
  const float *prior_fusion_w, *prior_fusion_b;   /* (3,2,3,3),(3) or NULL when q_chans == 3 */
  const float *pe_w, *pe_b;                       /* patch_embed.proj (C,3,p,p),(C) */
  const float *pe_norm_w, *pe_norm_b;             /* patch_embed.norm (C) */
  dpmn_block_weights blocks[DPMN_MAX_BLOCKS];
  const float *head0_w, *head0_b;                 /* conv_before_upsample.0 (hs*p*p, C, 3, 3) */
  const float *head1_w, *head1_b;                 /* conv_before_upsample.1 (hs*p*p, hs*p*p, 3, 3) */
  /* out = head * mix_weight[0] + sum_{i=1}^{n_mix-1} mix_input[i] * mix_weight[i]   (pgrm.py:562-564;
   * mix_input[0] is never read: the reference skips residual_list[0]) */
  const float *mix_weight[DPMN_MAX_MIX];          /* weight_list_i (1, hs, img_h, img_w) */
  const float *mix_input[DPMN_MAX_MIX];           /* residual_list[i] (B, hs, img_h, img_w) */
  int64_t mix_input_batch_stride[DPMN_MAX_MIX];   /* 0 = dense */
  /* Tensor-core modes stage 16-bit copies of the weights.  prepared == NULL: staged in the workspace on every
   * call.  Otherwise a caller-owned device buffer of dpmn_pgrm_prepared_bytes(): the call stages into it when
   * prepared_valid == 0 and reuses it when prepared_valid == 1 (the caller re-validates after the weights change). */
  void *prepared;
  int32_t prepared_valid;
  int32_t flags;                             /* DPMN_PGRM_* bits, 0 by default */
  /* Train-mode stochastic regularisers (module.train(), pgrm.py:24,32,40 Mlp Dropout; :180,248 attn_drop; :310,329-330
   * DropPath; :494,554-555 pos_drop).  All rates 0 (the default) = eval semantics.  With a non-zero rate the forward
   * runs the fp32 training sequence whatever `precision` says and needs dpmn_pgrm_backward_workspace_bytes() of
   * workspace; dpmn_pgrm_backward given the same rates and seed regenerates the same masks.  Masks are a pure function
   * of (seed, site, element index) -- NOT torch's Philox stream (the reference's stream cannot be reproduced). */
  float drop_rate;                           /* drop_rate[iter]: pos_drop and both Mlp dropouts */
  float attn_drop_rate;                      /* attn_drop_rate[iter] */
  float drop_path_rate[DPMN_MAX_BLOCKS];     /* dpr slice of this PGRM (pgrm.py:499,512) */
  uint64_t seed;                             /* fresh per training forward */
} dpmn_pgrm_desc;

/* dpmn_pgrm_backward only: `workspace` is the buffer a dpmn_pgrm_forward call with non-zero drop rates (i.e. the fp32
 * training sequence), the same descriptor, seed and inputs has just filled, untouched since -- the backward then
 * skips its internal forward recompute. */
#define DPMN_PGRM_WORKSPACE_HOLDS_FORWARD 1

/* ---- Complementation Modulation Module (cmm.py:80-161) ------------------------------------------- */
typedef struct dpmn_bn {          /* nn.BatchNorm2d (cmm.py:12), eps 1e-5, momentum 0.1 */
  const float *w, *b;             /* (ch) */
  float *running_mean, *running_var;   /* (ch); written only when training && update_running_stats */
} dpmn_bn;

typedef struct dpmn_cmm_stage {   /* EncodeBlock (cmm.py:38-55) / DecodeBlock (cmm.py:58-77) */
  const float *conv_a_w, *conv_a_b;    /* index 1: enc conv4x4 s2 d2 p3 (ci,ci) | dec convT3x3 (cin,co) */
  dpmn_bn bn_a;                        /* index 2 */
  const float *conv_b_w, *conv_b_b;    /* index 4: enc conv3x3 (co,ci)        | dec convT4x4 s2 (co,co) */
  dpmn_bn bn_b;                        /* index 5 */
} dpmn_cmm_stage;

typedef struct dpmn_cmm_desc {
  int32_t batch;                  /* B */
  int32_t img_h, img_w;           /* 32, 128 (must be divisible by 32) */
  int32_t c_img, cnum;            /* 3, 64 */
  int32_t precision;              /* DPMN_PREC_* */
  int32_t training;               /* 1: BatchNorm uses batch statistics (module.train()) */
  int32_t update_running_stats;   /* 1: also apply the momentum update to running_mean/var */
  int64_t x1_batch_stride, x2_batch_stride;   /* 0 = dense */
  const float *en1_w[2], *en1_b[2];           /* en_1_{1,2}  conv3x3 (cnum, c_img) */
  dpmn_cmm_stage enc[2][4];                   /* en_{2..5}_{1,2} */
  const float *en6_w[2], *en6_b[2];           /* en_6_{1,2}.1 conv4x4 s2 p1 */
  const float *fc1_w, *fc1_b, *fc2_w, *fc2_b; /* SE gate (4cnum,16cnum), (16cnum,4cnum) */
  const float *de6_w, *de6_b;                 /* de_6.1 convT4x4 s2 p1 (16cnum, 8cnum) */
  dpmn_bn de6_bn;                             /* de_6.2 */
  dpmn_cmm_stage dec[4];                      /* de_5, de_4, de_3, de_2 */
  const float *de1_w, *de1_b;                 /* de_1.1 convT3x3 (3cnum, c_img) */
  void *prepared;                             /* as in dpmn_pgrm_desc: staged tap-major 16-bit weights + folded BN */
  int32_t prepared_valid;
  int32_t flags;                              /* DPMN_CMM_* bits, 0 by default */
  /* Optional final blend of the eval / test call sites (interfaces/super_resolution.py:449,705):
   *   out = blend_alpha * CMM(x1, x2) + (1 - blend_alpha) * blend_input      (blend_input = images_lr_psn[:, :3])
   * fused into the kernel that writes `out`.  blend_input == NULL (default): plain CMM output.  (B, c_img, H, W) fp32,
   * batch stride in elements (0 = dense; the call sites pass a channel-slice view of a 4-channel tensor). */
  const float *blend_input;
  int64_t blend_input_batch_stride;
  float blend_alpha;
  int32_t reserved_;
} dpmn_cmm_desc;

/* dpmn_cmm_backward only: `workspace` is the buffer a dpmn_cmm_forward call with precision DPMN_PREC_F32 or with
 * training == 1 (i.e. the fp32-structured forward that keeps raw conv outputs and BatchNorm affines; in the 16-bit modes
 * its convs run on tcgen05 through a 16-bit im2col), the same descriptor and the same x1 / x2 has just filled, and nothing
 * has written to it since -- the backward then skips its internal forward recompute.  The buffer must have
 * dpmn_cmm_backward_workspace_bytes(). */
#define DPMN_CMM_WORKSPACE_HOLDS_FORWARD 1

const char *dpmn_version(void);

/* 0 if the current device can run this library (compute capability 10.x), else DPMN_E_DEVICE. */
int dpmn_check_device(void);

/* sizeof() of the descriptor structs as compiled, so a binding can verify its own struct layout:
 * which = 0 dpmn_block_weights, 1 dpmn_pgrm_desc, 2 dpmn_bn, 3 dpmn_cmm_stage, 4 dpmn_cmm_desc,
 * 5 dpmn_block_grads, 6 dpmn_pgrm_grads, 7 dpmn_cmm_grads. */
size_t dpmn_abi_sizeof(int32_t which);

/* number of kernels launched by this library in this process so far (bench.py's gpu_launches). */
uint64_t dpmn_launch_count(void);

/* Per-launch device timing for the roofline report (bench.py).  While enabled, every kernel launch of
 * this library is bracketed by CUDA events on its stream; dpmn_profile_collect synchronises on them and
 * returns up to `cap` records (kernel-class tag, kernels in the record, milliseconds), then clears. */
int dpmn_profile_enable(int32_t on);
int32_t dpmn_profile_collect(int32_t *tags, int32_t *n_kernels, float *ms, int32_t cap);
const char *dpmn_profile_tag_name(int32_t tag);

size_t dpmn_pgrm_workspace_bytes(const dpmn_pgrm_desc *d);
size_t dpmn_pgrm_prepared_bytes(const dpmn_pgrm_desc *d);   /* 0 in the fp32 mode */

/* PGRM.forward.  x_q (B, q_chans, img_h, img_w), x_kv (B, 3, img_h, img_w), out (B, hidden_size, img_h, img_w). */
int dpmn_pgrm_forward(const dpmn_pgrm_desc *d, const float *x_q, const float *x_kv, float *out,
                      void *workspace, size_t workspace_bytes, void *stream);

/* Debug/probe variant: additionally copies per-block tensors out (any pointer may be NULL):
 * attn_core[b] (B, L, C) the window-major attention output that enters SKConv (fp32),
 * block_out[b] (B, L, C) the x_kv stream after block b. */
int dpmn_pgrm_forward_probe(const dpmn_pgrm_desc *d, const float *x_q, const float *x_kv, float *out,
                            void *workspace, size_t workspace_bytes, void *stream,
                            float *const attn_core[DPMN_MAX_BLOCKS], float *const block_out[DPMN_MAX_BLOCKS]);

/* Stand-alone windowed attention core of one block (the roofline-sweep kernel).
 * q (B, L, C), kv (B, L, 2C) are the PROJECTED tensors in token order (fp32, or fp16/bf16 when
 * precision != F32: then q/kv/out are 16-bit); out (B, L, C) in the reference's window-major row order.
 * grid_h * grid_w = L.  window[g]/shift[g] are the EFFECTIVE window and cyclic shift of group g. */
int dpmn_window_attn_forward(const void *q, const void *kv, void *out, const float *const rpb_table[DPMN_MAX_GROUPS],
                             int32_t batch, int32_t grid_h, int32_t grid_w, int32_t embed_dim, int32_t num_heads,
                             int32_t n_groups, const int32_t window[DPMN_MAX_GROUPS],
                             const int32_t shift[DPMN_MAX_GROUPS], int32_t precision,
                             void *workspace, size_t workspace_bytes, void *stream);
size_t dpmn_window_attn_workspace_bytes(int32_t batch, int32_t tokens, int32_t embed_dim, int32_t precision);

/* The same attention core on WINDOW-MAJOR operands -- the layout the fused pipeline feeds it (the q / kv projection
 * epilogue applies roll + window_partition, pgrm.py:209-225): qw, kw, vw are [n_groups][batch*L][embed_dim/n_groups]
 * 16-bit, row p of group g = token window_row_to_token(p) of that group's (window, shift).  This is the tcgen05
 * kernel (attn2_tc.cu; attn_tc.cu with DPMN_ATTN_V1=1) itself: precision must be F16 or BF16, windows in {2,4,8},
 * head_dim 16 or 32; anything else returns DPMN_E_UNSUPPORTED.  out (batch*L, embed_dim) window-major rows. */
int dpmn_window_attn_forward_windowed(const void *qw, const void *kw, const void *vw, void *out,
                                      const float *const rpb_table[DPMN_MAX_GROUPS], int32_t batch, int32_t grid_h,
                                      int32_t grid_w, int32_t embed_dim, int32_t num_heads, int32_t n_groups,
                                      const int32_t window[DPMN_MAX_GROUPS], const int32_t shift[DPMN_MAX_GROUPS],
                                      int32_t precision, void *stream);
/* Train-mode form (WindowAttention.attn_drop, model/pgrm.py:180,248): each softmax probability is kept with probability
 * 1 - attn_drop and scaled by 1/(1 - attn_drop), inside the same kernel.  The mask of element (b, g, head-in-group, window-major
 * row p, key m) is dpmn_mask_hash(seed, site, ((((b*G + g)*heads_per_group + head)*L + p)*N + m)) (see dpmn_mask_hash). */
int dpmn_window_attn_forward_windowed_train(const void *qw, const void *kw, const void *vw, void *out,
                                            const float *const rpb_table[DPMN_MAX_GROUPS], int32_t batch, int32_t grid_h,
                                            int32_t grid_w, int32_t embed_dim, int32_t num_heads, int32_t n_groups,
                                            const int32_t window[DPMN_MAX_GROUPS], const int32_t shift[DPMN_MAX_GROUPS],
                                            int32_t precision, float attn_drop, uint64_t seed, uint32_t site, void *stream);

/* Backward of the same core on tcgen05 (attn2_bwd_tc.cu): d_out16 (batch*L, embed_dim) 16-bit, the gradient of `out` in its
 * window-major row order; writes dq (batch*L, embed_dim) and dkv (batch*L, 2*embed_dim) fp32 in TOKEN order (the gradients of
 * the projected q / kv tensors, roll + window_partition undone) and ACCUMULATES d_rpb_table[g] ((2ws-1)^2, heads/G) fp32.
 * Scores are recomputed from qw / kw; attn_drop / seed / site must be the forward's.  Windows in {2,4,8}. */
int dpmn_window_attn_backward_windowed(const void *qw, const void *kw, const void *vw, const void *d_out16, float *dq, float *dkv,
                                       const float *const rpb_table[DPMN_MAX_GROUPS], float *const d_rpb_table[DPMN_MAX_GROUPS],
                                       int32_t batch, int32_t grid_h, int32_t grid_w, int32_t embed_dim, int32_t num_heads,
                                       int32_t n_groups, const int32_t window[DPMN_MAX_GROUPS], const int32_t shift[DPMN_MAX_GROUPS],
                                       int32_t precision, float attn_drop, uint64_t seed, uint32_t site, void *stream);

size_t dpmn_cmm_workspace_bytes(const dpmn_cmm_desc *d);
size_t dpmn_cmm_prepared_bytes(const dpmn_cmm_desc *d);     /* 0 in the fp32 mode */

/* ComplementationModulationModule.forward.  x1, x2 (B, c_img, img_h, img_w) -> out (B, c_img, img_h, img_w). */
int dpmn_cmm_forward(const dpmn_cmm_desc *d, const float *x1, const float *x2, float *out,
                     void *workspace, size_t workspace_bytes, void *stream);

/* Debug: copy one internal activation buffer of the last tensor-core dpmn_cmm_forward out of `workspace`
 * (tests localise a parity failure to a layer with it).  which: l = 1..5 encoder outputs after LeakyReLU
 * [2][B][H_l][W_l][C_l]; 10+l (l = 2..5) EncodeBlock intermediates; 20+l (l = 1..5) ReLU'd decoder concat
 * inputs [B][H_l][W_l][C]; 30+l (l = 2..5) DecodeBlock intermediates; 40 en_6 outputs (fp32); 41 gated bottleneck.
 * All NHWC, 16-bit in the precision of the descriptor unless noted. */
size_t dpmn_cmm_debug_bytes(const dpmn_cmm_desc *d, int32_t which);
int dpmn_cmm_debug_copy(const dpmn_cmm_desc *d, void *workspace, int32_t which, void *dst, size_t dst_bytes, void *stream);

/* C (M, N) = A (M, K) * B (N, K)^T + bias (N), fp32 in/out; `precision` picks FFMA or tcgen05 operands.
 * The contraction every nn.Linear / 1x1 conv of the path reduces to; exported for unit tests. */
int dpmn_gemm_nt(const float *A, const float *B, const float *bias, float *C, int32_t M, int32_t N, int32_t K,
                 int32_t precision, void *workspace, size_t workspace_bytes, void *stream);
size_t dpmn_gemm_nt_workspace_bytes(int32_t M, int32_t N, int32_t K, int32_t precision);


/* ---- backward (training) --------------------------------------------------------------------------
 * The reference trains through torch autograd (interfaces/super_resolution.py:245-275: loss.backward(),
 * per-module clip_grad_norm_, Adam).  These entry points replace autograd's backward of PGRM.forward
 * (model/pgrm.py:546-565) and ComplementationModulationModule.forward (model/cmm.py:120-161):
 * given d loss / d out they produce d loss / d (every parameter, x_kv, residual_list[i] | x1, x2).
 *
 *   - stateless: the call recomputes the forward internally (fp32), so nothing has to be kept alive between
 *     dpmn_*_forward and dpmn_*_backward; `desc` and the inputs are the ones given to the forward
 *   - parameter gradients ACCUMULATE (+=, atomics) into the caller's buffers -- zero them (or the flat gradient
 *     bucket they live in) once per step; input gradients are WRITTEN, dense (B, ch, img_h, img_w)
 *   - a NULL gradient pointer for an input skips it; parameter-gradient pointers must all be non-NULL
 *     (prior_fusion_* only when q_chans == 2, mix_weight[i] for i < n_mix)
 *   - arithmetic is fp32 in every precision mode of the descriptor (the 16-bit modes only affect forward)
 *   - eval-mode semantics of Dropout / DropPath (rates 0): the stochastic train-mode paths are not implemented
 */
typedef struct dpmn_block_grads {        /* mirrors dpmn_block_weights field by field */
  float *norm1_q_w, *norm1_q_b, *norm1_kv_w, *norm1_kv_b;
  float *rpb_table[DPMN_MAX_GROUPS];
  float *q_w, *q_b, *kv_w, *kv_b;
  float *sk_proj_w, *sk_proj_b, *sk_fc1_w, *sk_fc1_b, *sk_fc2_w, *sk_fc2_b, *sk_head_w, *sk_head_b;
  float *norm2_w, *norm2_b;
  float *fc1_w, *fc1_b, *fc2_w, *fc2_b, *dw_w, *dw_b, *pw_w, *pw_b;
} dpmn_block_grads;

typedef struct dpmn_pgrm_grads {
  float *prior_fusion_w, *prior_fusion_b;
  float *pe_w, *pe_b, *pe_norm_w, *pe_norm_b;
  dpmn_block_grads blocks[DPMN_MAX_BLOCKS];
  float *head0_w, *head0_b, *head1_w, *head1_b;
  float *mix_weight[DPMN_MAX_MIX];       /* d weight_list_i, i < n_mix */
  float *x_kv;                           /* d x_kv (B, 3, img_h, img_w) or NULL */
  float *mix_input[DPMN_MAX_MIX];        /* d residual_list[i] (B, hs, img_h, img_w) or NULL; [0] is never touched */
} dpmn_pgrm_grads;

/* The hash behind the train-mode masks (host function, no GPU needed): element `idx` of site `site` is kept iff
 * (dpmn_mask_hash(seed, site, idx) >> 8) * 2^-24 >= rate, and then scaled by 1/(1-rate).  Sites: 1 pos_drop(x_q tokens),
 * 2 pos_drop(x_kv tokens), 16*(block+1) + {0 attn_drop, 1 Mlp drop after GELU, 2 Mlp drop after fc2, 3 DropPath of the
 * attention branch, 4 DropPath of the Mlp branch}.  Indices: flat (B,L,C) / (B,L,hidden) element index; DropPath: image
 * index; attn_drop: ((((b*G + g)*heads_per_group + head)*L + window_major_row)*N + key). */
uint32_t dpmn_mask_hash(uint64_t seed, uint32_t site, uint64_t idx);

size_t dpmn_pgrm_backward_workspace_bytes(const dpmn_pgrm_desc *d);
int dpmn_pgrm_backward(const dpmn_pgrm_desc *d, const float *x_q, const float *x_kv, const float *d_out,
                       const dpmn_pgrm_grads *grads, void *workspace, size_t workspace_bytes, void *stream);

typedef struct dpmn_bn_grads { float *w, *b; } dpmn_bn_grads;
typedef struct dpmn_cmm_stage_grads {    /* mirrors dpmn_cmm_stage */
  float *conv_a_w, *conv_a_b; dpmn_bn_grads bn_a;
  float *conv_b_w, *conv_b_b; dpmn_bn_grads bn_b;
} dpmn_cmm_stage_grads;
typedef struct dpmn_cmm_grads {          /* mirrors the parameters of dpmn_cmm_desc; all pointers required */
  float *en1_w[2], *en1_b[2];
  dpmn_cmm_stage_grads enc[2][4];
  float *en6_w[2], *en6_b[2];
  float *fc1_w, *fc1_b, *fc2_w, *fc2_b;
  float *de6_w, *de6_b; dpmn_bn_grads de6_bn;
  dpmn_cmm_stage_grads dec[4];
  float *de1_w, *de1_b;
  float *x1, *x2;                        /* d x1, d x2 (B, c_img, img_h, img_w) or NULL */
} dpmn_cmm_grads;

/* Backward of ComplementationModulationModule.forward (cmm.py:120-161).  d->training selects batch-statistics
 * BatchNorm backward (module.train()) or the running-statistics affine (eval()); running stats are never
 * updated by this call.  x1 / x2 must be dense. */
size_t dpmn_cmm_backward_workspace_bytes(const dpmn_cmm_desc *d);
int dpmn_cmm_backward(const dpmn_cmm_desc *d, const float *x1, const float *x2, const float *d_out,
                      const dpmn_cmm_grads *grads, void *workspace, size_t workspace_bytes, void *stream);

/* ---- the callers' steps on either side of the hot path ------------------------------------------
 * dpmn_image_loss   <- ImageLoss(gradient=True, loss_weight=[w_mse, w_gp]).forward + its autograd backward
 *                      loss/image_loss.py:15-43; call sites interfaces/super_resolution.py:212,239,267
 *   loss[0] += scale * (w_mse * MSE(out, target) + w_gp * L1(gradient_map(out[:, :3]), gradient_map(target[:, :3])));
 *   d_out (dense (B, C, H, W), or NULL) = scale * d ImageLoss / d out.  out / target (B, C, H, W) fp32 with batch
 *   strides in elements (0 = dense).  `loss` is accumulated: zero it once for a sum of several terms.
 * dpmn_to_mask      <- toMask(img) per image, utils/util.py:27-35 (ToPILImage -> convert('L') -> threshold at the mean
 *                      -> 0 / 255 inverted -> ToTensor -> repeat 3 channels); call site super_resolution.py:220-226
 *   img (B, 3, H, W) fp32 in [0, 1] (batch stride in elements, 0 = dense) -> mask (B, 3, H, W) fp32 in {0, 1}, bit-exact. */
int dpmn_image_loss(const float *out, int64_t out_batch_stride, const float *target, int64_t target_batch_stride,
                    int32_t batch, int32_t chans, int32_t img_h, int32_t img_w, float w_mse, float w_gp, float scale,
                    float *loss, float *d_out, void *stream);
int dpmn_to_mask(const float *img, int64_t img_batch_stride, float *mask, int32_t batch, int32_t img_h, int32_t img_w,
                 void *stream);


/* ---- recogniser-input resizes (second half) ----------------------------------------------------
 * dpmn_crnn_input      <- TextBase.parse_crnn_data(imgs), interfaces/base.py:419-425; call sites super_resolution.py:159,165
 *   out (B, 1, out_h, out_w) = luma(F.interpolate(img, (out_h, out_w), mode='bicubic')), luma = 0.299 R + 0.587 G + 0.114 B.
 *   The reference passes (32, 100).  fp32, within 5e-6 of torch's CPU result (summation order).
 * dpmn_visionlan_input <- TextBase.parse_visionlan_data(img) for every image of a batch, interfaces/base.py:473-478; call
 *   site super_resolution.py:177-178: ToPILImage -> cv2.resize(.., (out_w, out_h)) -> ToTensor.  The reference passes
 *   (256, 64).  out (B, 3, out_h, out_w); integer arithmetic, bit-exact against OpenCV's uint8 INTER_LINEAR.
 * img (B, 3, H, W) fp32, batch stride in elements (0 = dense). */
int dpmn_crnn_input(const float *img, int64_t img_batch_stride, float *out, int32_t batch, int32_t img_h, int32_t img_w,
                    int32_t out_h, int32_t out_w, void *stream);
int dpmn_visionlan_input(const float *img, int64_t img_batch_stride, float *out, int32_t batch, int32_t img_h, int32_t img_w,
                         int32_t out_h, int32_t out_w, void *stream);

/* ---- DistillModule ----------------------------------------------------------------------------
 * dpmn_distill_forward  <- DistillModule.forward(x_deep, x_shallow), model/distill_module.py:18-31; call sites
 *                          interfaces/super_resolution.py:245-263 (4 instances per training step)
 *   a = ReLU(bn_1(conv_cat_feature(cat[x_deep, x_shallow])));  s = ReLU(bn_2(conv_feature(x_shallow)));
 *   loss[0] += mean |a - s| (accumulated: zero it once);  feature (B, 3, H, W) dense = a.  Either output may be NULL.
 *   x_deep / x_shallow (B, 3, H, W) fp32, batch strides in elements (0 = dense).  training = 1: batch statistics
 *   (+ the momentum update of the running ones when update_running_stats); 0: running statistics.
 * dpmn_distill_backward <- its autograd backward.  d_loss: DEVICE pointer to the scalar gradient of `loss` (NULL = 0);
 *   d_feature (B, 3, H, W) dense gradient of `feature` (NULL = 0).  Parameter gradients accumulate into `grads`;
 *   grads->x_deep / x_shallow (dense, or NULL) are written.  Stateless by default (recomputes the convs); with
 *   DPMN_DISTILL_WORKSPACE_HOLDS_FORWARD in d->flags the workspace is the one dpmn_distill_forward has just filled for
 *   the same descriptor and inputs.  Running statistics are never updated by the backward. */
#define DPMN_DISTILL_WORKSPACE_HOLDS_FORWARD 1
typedef struct dpmn_distill_desc {
  int32_t batch, img_h, img_w;
  int32_t training, update_running_stats, flags;
  float bn_eps, bn_momentum;                      /* nn.BatchNorm2d defaults: 1e-5, 0.1 */
  int64_t deep_batch_stride, shallow_batch_stride;
  const float *conv_cat_w, *conv_cat_b;           /* conv_cat_feature (3, 6, 3, 3), (3) */
  dpmn_bn bn_1;
  const float *conv_w, *conv_b;                   /* conv_feature (3, 3, 3, 3), (3) */
  dpmn_bn bn_2;
} dpmn_distill_desc;
typedef struct dpmn_distill_grads {
  float *conv_cat_w, *conv_cat_b; dpmn_bn_grads bn_1;
  float *conv_w, *conv_b; dpmn_bn_grads bn_2;
  float *x_deep, *x_shallow;
} dpmn_distill_grads;
size_t dpmn_distill_workspace_bytes(const dpmn_distill_desc *d);
int dpmn_distill_forward(const dpmn_distill_desc *d, const float *x_deep, const float *x_shallow, float *loss,
                         float *feature, void *workspace, size_t workspace_bytes, void *stream);
int dpmn_distill_backward(const dpmn_distill_desc *d, const float *x_deep, const float *x_shallow, const float *d_loss,
                          const float *d_feature, const dpmn_distill_grads *grads, void *workspace, size_t workspace_bytes,
                          void *stream);

/* ---- the tail of one data-parallel training step (SURVEY.md 8b proposal, 8e) --------------------------------------------
 * dpmn_allreduce_bucket <- the gradient reduction of nn.DataParallel(model, device_ids=range(ngpu)), interfaces/base.py:160-162
 *                          (replicate + scatter + gather onto GPU 0 every iteration in the reference); here: ONE
 *                          ncclAllReduce(sum) in place over `count` elements of the flat gradient bucket, on `stream`.
 *   comm: from dpmn_nccl_comm_init (rank 0 makes the 128-byte id with dpmn_nccl_unique_id and ships it to the other ranks
 *   through any side channel -- dpmn_b200.dist uses the torch.distributed store).  dtype: DPMN_PREC_F32 / F16 / BF16.
 *   libnccl is resolved with dlopen at the first call (the copy already loaded in the process first); without it every
 *   function here returns DPMN_E_DEVICE and dpmn_nccl_available() is 0.
 * dpmn_clip_adam_step   <- torch.nn.utils.clip_grad_norm_(module.parameters(), 0.25) per module, super_resolution.py:270-275,
 *                          + optimizer_G.step() = Adam(lr, betas=(0.5, 0.999)), interfaces/base.py:208-221, over FLAT fp32
 *                          buffers (parameters, gradients, exp_avg, exp_avg_sq: same length, 16-byte aligned) in two launches.
 *   segment_offsets: HOST array of n_segments + 1 element offsets (module boundaries inside the flat buffers; offsets[0] = 0,
 *   offsets[n] = total; n_segments <= 32): the 2-norm is taken and clipped per segment, as the reference does per module.
 *   grad_scale multiplies every gradient first (1 / world after a summing all-reduce).  max_norm <= 0: no clipping.
 *   step: 1-based Adam step count (bias correction).  workspace: dpmn_clip_adam_workspace_bytes(n_segments) device bytes. */
int dpmn_nccl_available(void);
int dpmn_nccl_version(void);                                  /* ncclGetVersion(), 0 when unavailable */
int dpmn_nccl_unique_id(void *id128);                         /* HOST buffer of 128 bytes */
int dpmn_nccl_comm_init(const void *id128, int32_t world, int32_t rank, void **comm_out);   /* current device */
int dpmn_nccl_comm_destroy(void *comm);
int dpmn_allreduce_bucket(void *comm, void *bucket, size_t count, int32_t dtype, void *stream);
size_t dpmn_clip_adam_workspace_bytes(int32_t n_segments);
int dpmn_clip_adam_step(float *params, const float *grads, float *exp_avg, float *exp_avg_sq, const int64_t *segment_offsets,
                        int32_t n_segments, float grad_scale, float max_norm, float lr, float beta1, float beta2, float eps,
                        int64_t step, void *workspace, size_t workspace_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DPMN_B200_H */
