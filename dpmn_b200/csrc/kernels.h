// Internal launcher interface between api.cu and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dpmn {

enum DType : int { DT_F32 = 0, DT_F16 = 1, DT_BF16 = 2 };
inline size_t dtype_size(DType t) { return t == DT_F32 ? 4 : 2; }

// ---- SIMT kernels (pgrm_simt.cu) ------------------------------------------------------------------

// prior_fusion (optional, in_ch == 2) + PatchEmbed conv(k=s=patch) + LayerNorm -> tokens (B, L, C) fp32.
// x is NCHW with an arbitrary batch stride (elements); channel/row strides are dense.
// Optionally also writes up to two further LayerNorms of every token (the consumers' norm1_q / norm1_kv) as
// (B*L, C) tensors of `type`; `tokens` may then be nullptr (the query stream is only ever read through norm1_q).
struct PatchEmbedLn {
  int count = 0; int type = 0;
  void* out[2] = {}; const float* w[2] = {}; const float* b[2] = {};
};
int launch_patch_embed(const float* x, long long x_bs, int in_ch, const float* fuse_w, const float* fuse_b,
                       const float* pe_w, const float* pe_b, const float* ln_w, const float* ln_b, float* tokens,
                       int B, int img_h, int img_w, int patch, int C, cudaStream_t st,
                       const PatchEmbedLn* extra = nullptr);

// LayerNorm over the last dim of (rows, C) fp32 -> out (rows, C) of `out_type`.
int launch_layernorm(const float* x, const float* w, const float* b, void* out, DType out_type, int rows, int C,
                     cudaStream_t st);
int launch_layernorm_dual(const float* x, const float* w, const float* b, float* out, void* out16, DType type16, int rows, int C,
                          cudaStream_t st);

// C[z][m, n] = epi( sum_k A[z][m, k] * B[z][n, k] ), everything fp32 (the exact-arithmetic mode).
struct GemmSimtArgs {
  const float* A = nullptr; long long a_bs = 0; int lda = 0;
  const float* Bm = nullptr; long long b_bs = 0; int ldb = 0;
  float* C = nullptr; long long c_bs = 0; int ldc = 0;
  int M = 0, N = 0, K = 0, batch = 1;
  const float* bias = nullptr; long long bias_bs = 0; int bias_mode = 0;   // 0 none, 1 per n, 2 per m
  int act = 0;                               // 0 none, 1 GELU(erf)
  const float* residual = nullptr;           // fp32, same indexing as C (may alias C)
  float* colsum = nullptr;                   // if set: no C store; colsum[z][m_tile][n] = sum_rows act(value)
};
int launch_gemm_simt(const GemmSimtArgs& a, cudaStream_t st);
static constexpr int kSimtTileM = 64;

// tcgen05 GEMM (gemm_tc.cu): same contraction with 16-bit K-contiguous operands (fp16 / bf16), fp32
// accumulation in TMEM.  a_bs / b_bs == 0 with batch > 1 means the operand is shared by all batch entries.
// colsum layout: colsum[z][m_tile(128 rows) * 4 + quarter][n].
struct GemmTcArgs {
  const void* A = nullptr; long long a_bs = 0; int lda = 0;
  const void* Bm = nullptr; long long b_bs = 0; int ldb = 0;
  DType op_type = DT_F16;
  void* C = nullptr; long long c_bs = 0; int ldc = 0; DType out_type = DT_F32;
  int M = 0, N = 0, K = 0, batch = 1;
  const float* bias = nullptr; long long bias_bs = 0; int bias_mode = 0;
  int act = 0;
  const float* residual = nullptr;
  float* colsum = nullptr;
  // out16 (fp32 C only): a second copy of C in the operand type at the same indices (z * c_bs + m * ldc + n) -- the training
  // path keeps fp32 activations for the backward and hands the 16-bit copy to the next GEMM without a convert pass
  void* out16 = nullptr;
  // Second output from the finished row (needs N == the N tile <= 128): ln_mode 1 = LayerNorm(ln_w, ln_b) of the
  // row, 2 = plain copy; written as (batch*M, N) of ln_type.
  int ln_mode = 0; const float* ln_w = nullptr; const float* ln_b = nullptr; void* ln_out = nullptr; DType ln_type = DT_F16;
  // Window scatter (q / kv projections feeding attn_tc.cu): instead of C, every 32-column chunk of a row is
  // written to scatter_dst[n / scatter_C] as [group][window-major row][cg] (roll + window_partition,
  // pgrm.py:209-225, folded into this epilogue).  Rows are tokens of (B, H*W); out_type is the 16-bit type.
  int scatter = 0, scatter_C = 0, scatter_G = 0, scatter_H = 0, scatter_W = 0;
  int scatter_ws[4] = {}, scatter_shift[4] = {};
  void* scatter_dst[2] = {};
};
int launch_gemm_tc(const GemmTcArgs& a, cudaStream_t st);
int launch_gemm_tc_dual(const GemmTcArgs& a, const GemmTcArgs& b, cudaStream_t st);   // two window-scatter projections, one launch
bool gemm_res_ln_supported(const GemmTcArgs& a);
int launch_gemm_res_ln(const GemmTcArgs& a, cudaStream_t st);   // gemm_res_tc.cu
bool gemm_tc_can_fuse_row_output(int N);   // ln_mode != 0 is available for these N (fp32 C only)
static constexpr int kTcTileM = 128;

// Windowed attention core, SIMT (any precision of q/kv/out storage).  q (B,L,*) and kv (B,L,*) in token
// order with row strides q_ld / kv_ld (elements); K channels start at column 0 of a kv row, V channels
// at column v_off.  out (B,L,C) window-major rows per group (quirk 1, pgrm.py:249,263).
struct AttnArgs {
  const void* q = nullptr; const void* kv = nullptr; void* out = nullptr; DType io_type = DT_F32;
  int q_ld = 0, kv_ld = 0, v_off = 0, out_ld = 0;
  const float* table[4] = {nullptr, nullptr, nullptr, nullptr};   // ((2*win-1)^2, heads_per_group) each
  int B = 0, H = 0, W = 0, C = 0, n_groups = 0, heads_per_group = 0;
  int window[4] = {0, 0, 0, 0}, shift[4] = {0, 0, 0, 0};   // EFFECTIVE windows / shifts
  float p_drop = 0.f; unsigned long long seed = 0; uint32_t site = 0;   // attn_drop on P (train mode), pgrm.py:248
};
int launch_window_attn_simt(const AttnArgs& a, cudaStream_t st);

// tcgen05 window attention (attn_tc.cu) on window-major q / k / v: [G][B*L][C/G] 16-bit each.
struct AttnTcArgs {
  const void *qw = nullptr, *kw = nullptr, *vw = nullptr; void* out = nullptr; DType io_type = DT_F16;
  const float* table[4] = {nullptr, nullptr, nullptr, nullptr};
  int B = 0, H = 0, W = 0, C = 0, n_groups = 0, heads_per_group = 0;
  int window[4] = {0, 0, 0, 0}, shift[4] = {0, 0, 0, 0};
  float p_drop = 0.f; unsigned long long seed = 0; uint32_t site = 0;   // attn_drop (train mode), attn2_tc.cu only
};
bool attn_tc_supported(const AttnTcArgs& a);
int launch_window_attn_tc(const AttnTcArgs& a, cudaStream_t st);      // dispatches to attn2_tc.cu unless DPMN_ATTN_V1=1
bool attn2_tc_supported(const AttnTcArgs& a);                           // attn2_tc.cu: M = 64 tiles, two CTAs per SM
int launch_window_attn2_tc(const AttnTcArgs& a, cudaStream_t st);

// SK gate (pgrm.py:84-95 folded): from per-tile column sums of GELU(proj(x)) build, per image,
// Wb = Wp + Wh * diag(softmax_G(fc2(GELU(fc1(mean))))) (C x C) and bias_b = bp + bh.
int launch_sk_gate(const float* colsum, int tiles_per_image, int L, const float* wp, const float* bp,
                   const float* w1, const float* b1, const float* w2, const float* b2, const float* wh,
                   const float* bh, void* wb_out, DType wb_type, float* bias_out, int B, int C, int G,
                   cudaStream_t st);

// Depthwise 3x3 on the raw (B, hid, side, side) view of h (B, L, hid) + bias + GELU, written transposed
// as dt (B, L(pixel), hid(channel)).
int launch_dwconv(const void* h, void* dt, DType io_type, const float* w, const float* b, int B, int L, int hid,
                  cudaStream_t st);

// Mlp front half in one tcgen05 kernel (mlp_fused_a.cu): dt = GELU(dw3x3(raw view of GELU(x16 * w16^T + fc1_b))) written
// pixel-major (B, L, hid), 16-bit.  x16 (B*L, C), w16 (hid, C) are 16-bit of type t.  Only the production geometry.
bool mlp_fc1_dw_supported(int C, int hid, int L);
int launch_mlp_fc1_dw(const void* x16, const void* w16, const float* fc1_b, const float* dw_w, const float* dw_b, void* dt, int B,
                      DType t, cudaStream_t st);

// PatchUnEmbed + conv3x3 (C -> hp) on token-major x (B, gh, gw, C) -> t1 (B, gh, gw, hp) fp32.
int launch_head_conv1(const float* x, const float* w, const float* b, float* t1, int B, int gh, int gw, int C,
                      int hp, cudaStream_t st);
// conv3x3 (hp -> hp) + LeakyReLU(0.01) + PixelShuffle(patch) + affine mix -> out (B, hs, img_h, img_w).
struct MixArgs {
  int n_mix = 1;
  const float* w[8] = {};
  const float* in[8] = {};
  long long in_bs[8] = {};   // batch strides (elements) of the residual inputs
};
int launch_head_conv2_mix(const float* t1, const float* w, const float* b, float* out, int B, int gh, int gw,
                          int hs, int patch, const MixArgs& mix, cudaStream_t st);

// fp32 -> 16-bit conversion of n elements (weights staging for the tensor-core path).
int launch_convert(const float* src, void* dst, DType dst_type, long long n, cudaStream_t st);

// Several fp32 -> 16-bit conversions in ONE launch (all weights of a PGRM for the tensor-core modes).
struct ConvertBatch {
  int count = 0;
  const float* src[16] = {};
  void* dst[16] = {};
  long long n[16] = {};
};
int launch_convert_batch(const ConvertBatch& cb, DType dst_type, cudaStream_t st);
struct TransposeBatch {            // dst[i] (C[i], R[i]) 16-bit = transpose of src[i] (R[i], C[i]) fp32
  int count = 0;
  const float* src[16] = {};
  void* dst[16] = {};
  int R[16] = {}, C[16] = {};
};
int launch_transpose_convert_batch(const TransposeBatch& tb, DType t, cudaStream_t st);
// 16-bit -> fp32 (probe outputs).
int launch_widen(const void* src, DType src_type, float* dst, long long n, cudaStream_t st);

// ---- backward building blocks (bwd_common.cu) -----------------------------------------------------------

// C[z][m*ldc + n] (op)= alpha * sum_k A[z][m*sam + k*sak] * B[z][k*sbk + n*sbn], fp32, any strides.
// mode 0: store, 1: += (non-atomic), 2: atomicAdd (allows ksplit > 1 and c_bs == 0, i.e. a sum over the batch).
struct GemmGenArgs {
  const float* A = nullptr; long long sam = 0, sak = 0, a_bs = 0;
  const float* B = nullptr; long long sbk = 0, sbn = 0, b_bs = 0;
  float* C = nullptr; long long ldc = 0, c_bs = 0;
  int M = 0, N = 0, K = 0, batch = 1, ksplit = 1, mode = 0;
  float alpha = 1.0f;
};
int launch_gemm_gen(const GemmGenArgs& a, cudaStream_t st);
int launch_colsum(const float* x, long long ld, int rows, int N, float* out, cudaStream_t st);   // out[n] += sum_rows
int launch_rowsum(const float* x, int rows, int len, int mod, float* out, cudaStream_t st);      // out[row % mod] += sum_j
int launch_gelu_bwd(float* g, const float* x, long long n, cudaStream_t st);                      // g *= gelu'(x)
int launch_add(const float* a, const float* b, float* y, long long n, cudaStream_t st);
int launch_transpose_convert(const float* src, void* dst, DType t, int batch, int R, int Cc, cudaStream_t st,
                             void* copy16 = nullptr);
// one pass: 16-bit copy (R, Cc), 16-bit transpose (Cc, R) and column sums (+=) of an fp32 gradient matrix; any output may be null
int launch_stage_grad(const float* src, void* copy16, void* trans16, float* colsum, DType t, int R, int Cc, cudaStream_t st);   // (R,Cc) fp32 -> (Cc,R) 16-bit
int launch_reduce_partials(const float* partial, float* dst, int S, long long n, cudaStream_t st);           // dst += sum_s partial[s]
int launch_layernorm_bwd(const float* dy, const float* x, const float* w, float* dx, int accumulate, float* dw,
                         float* db, int rows, int C, cudaStream_t st);

// ---- PGRM backward kernels (pgrm_bwd.cu) ------------------------------------------------------------------
struct AttnBwdArgs {
  const float* q = nullptr; const float* kv = nullptr;     // (B, L, C), (B, L, 2C) token order
  const float* d_attn = nullptr;                           // (B, L, C) window-major rows
  float* dq = nullptr; float* dkv = nullptr;               // token order, written
  const float* table[4] = {}; float* d_table[4] = {};      // accumulated
  int B = 0, H = 0, W = 0, C = 0, n_groups = 0, heads_per_group = 0;
  int window[4] = {}, shift[4] = {};
  float p_drop = 0.f; unsigned long long seed = 0; uint32_t site = 0;   // the forward's attn_drop masks
  int round16 = 0;                                         // 1 fp16 / 2 bf16: the forward was the tcgen05 kernel (16-bit q, k, v, P)
};
int launch_window_attn_bwd(const AttnBwdArgs& a, cudaStream_t st);
// tcgen05 backward of the attention core (attn2_bwd_tc.cu): window-major 16-bit qw / kw / vw [G][rows][C/G] as the forward
// read them, d_out16 (rows, C) 16-bit window-major rows; dq (rows, C), dkv (rows, 2C) fp32 in TOKEN order (written),
// d_table[g] accumulated.  Windows 2 / 4 / 8, head_dim 16 / 32.
struct AttnBwdTcArgs {
  const void *qw = nullptr, *kw = nullptr, *vw = nullptr, *d_out16 = nullptr; DType io_type = DT_F16;
  float* dq = nullptr; float* dkv = nullptr;
  const float* table[4] = {}; float* d_table[4] = {};
  int B = 0, H = 0, W = 0, C = 0, n_groups = 0, heads_per_group = 0;
  int window[4] = {}, shift[4] = {};
  float p_drop = 0.f; unsigned long long seed = 0; uint32_t site = 0;
};
bool attn_bwd_tc_supported(const AttnBwdTcArgs& a);
int launch_window_attn_bwd_tc(const AttnBwdTcArgs& a, cudaStream_t st);
// fp32 token-order q (rows, C) / kv (rows, 2C) -> window-major 16-bit [G][rows][C/G] operands of the tcgen05 attention
int launch_window_scatter16(const float* q, const float* kv, void* qw, void* kw, void* vw, DType t, int B, int H, int W, int C,
                            int G, const int* ws, const int* shift, cudaStream_t st);

// S = mean_L GELU(F); zpre = fc1 S + b1; a = softmax_G(fc2 GELU(zpre) + b2); Xs = sum_m a_m * A_m
int launch_sk_train_fwd(const float* F, const float* A, const float* w1, const float* b1, const float* w2,
                        const float* b2, float* S, float* zpre, float* a, float* Xs, int B, int L, int C, int G,
                        cudaStream_t st);
struct SkBwdArgs {
  float* F = nullptr;                 // in: proj output; out: dF (in place)
  const float* d_out = nullptr;       // (rows, C)
  const float* dXs = nullptr;         // (rows, cg)
  const float* A = nullptr;           // attention output (rows, C)
  const float* a = nullptr; const float* zpre = nullptr; const float* S = nullptr;
  const float* w1 = nullptr; const float* w2 = nullptr;
  float *dw1 = nullptr, *db1 = nullptr, *dw2 = nullptr, *db2 = nullptr;   // accumulated
  float *da = nullptr, *dS = nullptr;                                     // scratch (B, C)
  float* dA = nullptr;                // (rows, C) written: dXs * a
  int B = 0, L = 0, C = 0, G = 0;
};
int launch_sk_bwd(const SkBwdArgs& s, cudaStream_t st);

// p_drop / seed / site: the Mlp's Dropout after GELU(fc1) (pgrm.py:31-32), applied to the conv INPUT on load
int launch_dwconv_train_fwd(const float* h1pre, float* dtpre, float* dt, const float* w, const float* b, int B, int L,
                            int hid, float p_drop, unsigned long long seed, uint32_t site, cudaStream_t st, void* dt16 = nullptr,
                            DType type16 = DT_F16);
int launch_dwconv_bwd(const float* d_dt, const float* dtpre, const float* h1pre, const float* w, float* d_h1pre,
                      float* dw, float* db, int B, int L, int hid, float p_drop, unsigned long long seed, uint32_t site,
                      cudaStream_t st);
// x[i] *= drop_scale(p, seed, site, i)                                  (pos_drop, pgrm.py:554-555, and its backward)
int launch_dropout(float* x, long long n, float p, unsigned long long seed, uint32_t site, cudaStream_t st);
// dst[i] = (base ? base[i] : 0) + y[i] * drop_scale(p_drop, site_drop, i) * drop_scale(p_path, site_path, i / per_image)
// -- a residual branch with its Dropout and per-sample DropPath (pgrm.py:40,329-330); with base == nullptr it is the
// backward of the branch (gradient of y from the gradient of the sum).
int launch_branch_combine(float* dst, const float* base, const float* y, long long n, long long per_image, float p_drop,
                          uint32_t site_drop, float p_path, uint32_t site_path, unsigned long long seed, cudaStream_t st);

int launch_patch_embed_bwd(const float* x, long long x_bs, int in_ch, const float* fuse_w, const float* fuse_b,
                           const float* pe_w, const float* pe_b, const float* ln_w, const float* d_tok, float* d_pe_w,
                           float* d_pe_b, float* d_ln_w, float* d_ln_b, float* dx3, int B, int img_h, int img_w,
                           int patch, int C, cudaStream_t st);
int launch_prior_fusion_wgrad(const float* dx3, const float* xq, long long xq_bs, float* dw, float* db, int B,
                              int img_h, int img_w, cudaStream_t st);

struct MixBwdArgs {
  int n_mix = 1;
  const float* w[8] = {};       // weight_list_i
  const float* in[8] = {};      // residual_list[i]
  long long in_bs[8] = {};
  float* d_w[8] = {};           // accumulated (may be nullptr for i >= 1)
  float* d_in[8] = {};          // written, dense (B, hs, H, W) (may be nullptr)
};
struct HeadBwdArgs {
  const float* tokens = nullptr;      // final kv stream (B, L, C)
  const float* t1 = nullptr;          // conv1 output (B, gh, gw, 16)
  const float *w1 = nullptr, *w2 = nullptr, *b2 = nullptr;
  const float* d_out = nullptr;       // (B, hs, img_h, img_w)
  float *dt2 = nullptr, *dt1 = nullptr;    // scratch (B, gh, gw, 16); dt1 zero-filled by the caller
  float* d_tokens = nullptr;          // (B, L, C) written
  float *d_w1 = nullptr, *d_b1 = nullptr, *d_w2 = nullptr, *d_b2 = nullptr;   // accumulated
  MixBwdArgs mix;
  int B = 0, gh = 0, gw = 0, C = 0, hs = 0, patch = 0;
};
int launch_head_bwd(const HeadBwdArgs& h, cudaStream_t st);

// ---- CMM SIMT kernels (cmm_simt.cu) ------------------------------------------------------------------

// Implicit-GEMM convolution / transposed convolution on NCHW fp32.  The input is the channel-wise
// concatenation of up to 3 tensors; each may carry a per-channel affine (its producer's BatchNorm,
// cmm.py:12) that is applied on load, followed by the consumer's leading activation (cmm.py:41,61).
struct ConvArgs {
  int n_seg = 1;
  const float* in[3] = {};         // (B, seg_ch[i], H, W)
  int seg_ch[3] = {};
  const float* in_scale[3] = {};   // per channel of the segment, or nullptr (identity)
  const float* in_shift[3] = {};
  int in_act = 0;                  // 0 none, 1 LeakyReLU(0.2), 2 ReLU
  const float* w = nullptr;        // conv: (Cout, Cin, k, k); transposed: (Cin, Cout, k, k)
  const float* bias = nullptr;     // (Cout)
  float* out = nullptr;            // (B, Cout, Ho, Wo)
  int B = 0, Cin = 0, H = 0, W = 0, Cout = 0, Ho = 0, Wo = 0;
  int k = 3, stride = 1, pad = 1, dil = 1;
  int transposed = 0;
  int ksplit = 1;                  // set by launch_conv_simt: K split over blockIdx.z with atomic accumulation
};
int launch_conv_simt(const ConvArgs& a, cudaStream_t st);

// BatchNorm2d as a per-channel affine (scale, shift) of a raw conv output x (B, C, HW):
// training: biased batch statistics of x (and, if run_mean != nullptr, the momentum-0.1 running update
// with the unbiased variance, as nn.BatchNorm2d does); eval: running statistics.
constexpr int BN_STATS_MAX_SLICES = 16;
inline size_t bn_stats_scratch_doubles(int C) { return (size_t)C * BN_STATS_MAX_SLICES * 2; }
int launch_bn_affine(const float* x, int B, int C, int HW, const float* w, const float* b, float* run_mean,
                     float* run_var, int training, int update_running, float eps, float* scale, float* shift,
                     cudaStream_t st, double* stats_scratch = nullptr);

// SE gate on the concatenated bottleneck (cmm.py:135-147): z (B, 2*Cb, hw) from two (B, Cb, hw) halves.
int launch_se_gate(const float* z1, const float* z2, float* z, const float* fc1_w, const float* fc1_b,
                   const float* fc2_w, const float* fc2_b, int B, int Cb, int hw, int hidden, cudaStream_t st,
                   float* h_scratch = nullptr);
// pooled mean (optional output) + relu(fc1) of the SE gate over B x hidden / 8 CTAs (shared by the forward and the backward)
int launch_se_pool_fc1(const float* z1, const float* z2, const float* fc1_w, const float* fc1_b, int B, int Cb, int hw,
                       int hidden, float* h_out, int h_stride, float* g0_out, int g0_stride, cudaStream_t st);

// ---- neighbours of the hot path (loss_mask.cu) ------------------------------------------------------------------
// loss[0] += scale * ImageLoss(out, target); d_out (dense (B,C,H,W), may be nullptr) = scale * d ImageLoss / d out
int launch_image_loss(const float* out, long long out_bs, const float* tgt, long long tgt_bs, int B, int C, int H, int W,
                      float w_mse, float w_gp, float scale, float* loss, float* d_out, cudaStream_t st);
// toMask (utils/util.py:27-35) per image: (B,3,H,W) in [0,1] -> inverted binary luma mask on 3 channels, dense
int launch_to_mask(const float* img, long long img_bs, float* mask, int B, int H, int W, cudaStream_t st);

// ---- recogniser-input resizes (resize.cu; interfaces/base.py:419-425,473-478) ---------------------------------------
// (B,3,H,W) -> (B,1,OH,OW): bicubic (ATen semantics) + luma.   (B,3,H,W) -> (B,3,OH,OW): uint8 truncation + OpenCV's
// fixed-point bilinear + / 255, bit-exact.  Batch stride of img in elements; outputs dense.
int launch_crnn_input(const float* img, long long img_bs, float* out, int B, int H, int W, int OH, int OW, cudaStream_t st);
int launch_visionlan_input(const float* img, long long img_bs, float* out, int B, int H, int W, int OH, int OW, cudaStream_t st);

// ---- DistillModule (distill.cu; model/distill_module.py:4-31) --------------------------------------------------------
struct DistillParams {
  const float *conv_cat_w, *conv_cat_b, *bn1_w, *bn1_b; float *bn1_mean, *bn1_var;
  const float *conv_w, *conv_b, *bn2_w, *bn2_b; float *bn2_mean, *bn2_var;
};
struct DistillGrads { float *conv_cat_w, *conv_cat_b, *bn1_w, *bn1_b, *conv_w, *conv_b, *bn2_w, *bn2_b, *x_deep, *x_shallow; };
size_t distill_workspace_floats(int B, int H, int W);
// conv outputs + BatchNorm statistics into ws; feature (B,3,H,W) = a; loss[0] += mean |a - s| (either may be nullptr)
int launch_distill_forward(const float* xd, long long d_bs, const float* xs, long long s_bs, const DistillParams& w, int B,
                           int H, int W, int training, int update_running, float eps, float momentum, float* feature,
                           float* loss, float* ws, cudaStream_t st);
// ws as the forward left it.  Parameter gradients accumulate; x_deep / x_shallow gradients are written (dense).
int launch_distill_backward(const float* xd, long long d_bs, const float* xs, long long s_bs, const DistillParams& w, int B,
                            int H, int W, int training, const float* d_loss, const float* d_feature, const DistillGrads& g,
                            float* ws, cudaStream_t st);

// ---- CMM training-path convs on the tcgen05 GEMM via a 16-bit im2col (cmm_im2col.cu) -----------------------------
struct ConvTcScratch {
  DType t = DT_F16;
  void* col = nullptr; size_t col_bytes = 0;      // im2col matrix
  void* w16 = nullptr; size_t w16_bytes = 0;      // staged (Cout, Kp) weights
  void* dy16 = nullptr; size_t dy16_bytes = 0;    // (Cout, Ntot) output gradient
  float* part = nullptr; size_t part_bytes = 0;   // split-K partials of the weight gradient
  void* act = nullptr; size_t act_bytes = 0;      // activated input as NHWC 16-bit (fast gather path, Cin % 8 == 0)
};
bool conv_tc_im2col_ok(const ConvArgs& a);
int launch_conv_tc_im2col(const ConvArgs& a, const ConvTcScratch& s, cudaStream_t st);               // same contract as launch_conv_simt
bool conv_wgrad_tc_im2col_ok(const ConvArgs& a, const ConvTcScratch& s);
int launch_conv_wgrad_tc_im2col(const ConvArgs& a, const float* dy, float* dw, const ConvTcScratch& s, cudaStream_t st);

// ---- CMM backward kernels (cmm_bwd.cu) ------------------------------------------------------------------
// weight gradient of the conv described by `a` (the FORWARD ConvArgs): dw += dy (x) gather(inputs); dy (B, Cout, Ho, Wo)
int launch_conv_wgrad_simt(const ConvArgs& a, const float* dy, float* dw, cudaStream_t st);
struct BnBwdArgs {
  const float* y = nullptr;            // raw conv output (B, C, HW)
  int B = 0, C = 0, HW = 0;
  const float *gamma = nullptr, *beta = nullptr, *run_mean = nullptr, *run_var = nullptr;   // gamma == nullptr: no BatchNorm
  int training = 0; float eps = 1e-5f;
  int n_src = 1;
  const float* du[2] = {};             // gradient w.r.t. each consumer's (activated) input; channel offset folded in
  long long du_bs[2] = {};             // batch strides (elements)
  int act[2] = {};                     // the consumer's activation: 0 none, 1 LeakyReLU(0.2), 2 ReLU
  float* dy = nullptr;                 // (B, C, HW) written
  float *dgamma = nullptr, *dbeta = nullptr, *dbias = nullptr;   // accumulated
  float* scratch = nullptr;            // >= 4*C floats: enables the multi-CTA-per-channel path for large B*HW
};
int launch_bn_act_bwd(const BnBwdArgs& a, cudaStream_t st);
int launch_chan_sum(const float* x, int B, int C, int HW, float* out, cudaStream_t st);
int launch_se_gate_bwd(const float* z1, const float* z2, const float* du, const float* fc1_w, const float* fc1_b,
                       const float* fc2_w, const float* fc2_b, float* dz1, float* dz2, float* d_fc1_w, float* d_fc1_b,
                       float* d_fc2_w, float* d_fc2_b, int B, int Cb, int hw, int hidden, cudaStream_t st,
                       float* scratch = nullptr);   // scratch: B * (4 Cb + 2 hidden) floats -> no atomics on the weight gradients

// ---- CMM tensor-core kernels (conv_tc.cu, cmm_tc.cu) ----------------------------------------------------

struct ConvTap { int8_t map, dy, dx, pad_; int16_t wslice, pad2_; };   // A-map id, pixel shift, weight slice
struct ConvTcSrc { const void* base = nullptr; long long sx = 0, sy = 0, sb = 0, sg = 0; };   // element strides
struct ConvTcDest {
  void* ptr = nullptr; int type = 0;      // DType of the destination
  long long g_stride = 0;                 // elements between groups
  int ld = 0;                             // elements between pixels (channel count of the destination buffer)
  int ch_off = 0, ch_g_off = 0;           // channel offset = ch_off + g * ch_g_off
  int act = 0;                            // 0 none, 1 LeakyReLU(0.2), 2 ReLU
};
// Implicit-GEMM conv on NHWC 16-bit activations; see conv_tc.cu.  Weights are [G][n_wslices][Cout][Cin] 16-bit.
struct ConvTcArgs {
  DType op_type = DT_F16;
  int n_src = 1; ConvTcSrc src[4];        // the (sub-)grids the taps index; all of extent (Wm, Hm, B, G)
  int Cin = 0, Cout = 0, B = 0, G = 1, P = 1;
  int Hm = 0, Wm = 0;
  int n_taps = 0; ConvTap taps[4][16] = {};
  const void* w = nullptr; int n_wslices = 0;
  int os = 1, Ho = 0, Wo = 0;
  const float* scale = nullptr; const float* shift = nullptr;   // [G][Cout]
  ConvTcDest dst[2];
  // Split-K scratch (optional): layers whose (pixel tile, Cout tile) grid cannot fill the machine split their taps x Cin
  // reduction over several CTAs, which write fp32 partial sums [slice][G][B*Ho*Wo][Cout] here; a second kernel adds the
  // slices in order and applies the epilogue.  nullptr / too small: the layer runs unsplit.
  float* splitk_ws = nullptr; size_t splitk_floats = 0;
};
int launch_conv_tc(const ConvTcArgs& a, cudaStream_t st);

// Weight staging for the tensor-core CMM: fp32 conv (Cout,Cin,k,k) / convT (Cin,Cout,k,k) -> 16-bit [k*k][rows][Cin]
// (rows >= Cout, extra rows zero), and conv bias + eval-BatchNorm folded into per-channel (scale, shift).
struct PrepSeg { const float* src; void* dst; int Cout, Cin, kk, transposed, rows; };
struct PrepBatch { int count = 0; PrepSeg seg[48]; };
int launch_prep_weights(const PrepBatch& pb, DType t, cudaStream_t st);
struct FoldSeg { const float *bias, *bn_w, *bn_b, *bn_rm, *bn_rv; float *scale, *shift; int C; };
struct FoldBatch { int count = 0; FoldSeg seg[48]; };
int launch_fold_bn(const FoldBatch& fb, float eps, cudaStream_t st);

// en_1 (cmm.py:84,95): conv3x3 c_img -> cnum on fp32 NCHW inputs of both branches; writes LeakyReLU(0.2) as
// e1 [2][B][H][W][cnum] and ReLU into the level-1 concat buffer (B,H,W,3*cnum) at channel cnum*(1+branch).
int launch_cmm_en1(const float* x1, const float* x2, const float* w1, const float* b1, const float* w2, const float* b2,
                   void* e1, void* cat1, DType t, int B, int H, int W, int c_img, int cnum, cudaStream_t st);
// SE gate (cmm.py:135-147) on fp32 NHWC halves z6 [2][B][hw][Cb] -> ReLU(z*g + z) as 16-bit (B, hw, 2*Cb).
int launch_se_gate_nhwc(const float* z6, void* zg, DType t, const float* fc1_w, const float* fc1_b, const float* fc2_w,
                        const float* fc2_b, float* pooled, float* hid_buf, int B, int Cb, int hw, int hidden,
                        cudaStream_t st);
// de_1 tail (cmm.py:113-116): out[b,co,y,x] = bias[co] + sum_taps P[b, y+1-ky, x+1-kx][(ky*3+kx)*c_img + co],
// P (B*H*W, ldp) fp32 from the tap-in-N GEMM.
// Optional blend (super_resolution.py:449,705): out = alpha * conv + (1 - alpha) * blend[b] (blend == nullptr: none).
int launch_de1_gather(const float* P, int ldp, const float* bias, float* out, int B, int H, int W, int c_img,
                      cudaStream_t st, const float* blend = nullptr, long long blend_bs = 0, float alpha = 1.f);
// out = alpha * out + (1 - alpha) * blend   (the fp32-structured path's form of the same blend)
int launch_alpha_blend(float* out, const float* blend, long long blend_bs, float alpha, int B, long long per_image, cudaStream_t st);

}  // namespace dpmn
