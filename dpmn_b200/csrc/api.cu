// C-ABI of libdpmn_b200 (include/dpmn_b200.h): argument checking, workspace carving and the launch
// sequences of PGRM.forward (pgrm.py:546-565) and ComplementationModulationModule.forward (cmm.py:120-161).
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/dpmn_b200.h"
#include "common.cuh"
#include "kernels.h"

using namespace dpmn;

namespace {

std::atomic<uint64_t> g_launches{0};

// Optional per-launch timing (bench.py's roofline leg): CUDA events on the launching stream around every
// kernel launch, tagged by kernel class.  Off by default; never active inside a timed throughput region.
enum KTag { T_PATCH_EMBED = 0, T_LAYERNORM, T_GEMM, T_WINDOW_ATTN, T_SK_GATE, T_DWCONV, T_HEAD, T_CONV, T_BN,
            T_SE_GATE, T_CONVERT, T_GEMM_TC, T_CONV_TC, T_PREP, T_CONV_STEM, T_ATTN_TC,
            T_BWD_GEMM, T_BWD_GEMM_TC, T_BWD_MISC, T_BWD_LN, T_BWD_SK, T_BWD_ATTN, T_BWD_DWCONV, T_BWD_HEAD, T_BWD_EMBED, T_BWD_CONV, T_BWD_BN, T_LOSS, T_DISTILL,
            T_MLP_A, T_OPTIM,
            T_COUNT };
const char* const kTagNames[T_COUNT] = {"patch_embed", "layernorm", "gemm", "window_attn", "sk_gate",
                                        "dwconv", "head", "conv", "bn_affine", "se_gate", "convert", "gemm_tc", "conv_tc", "weight_prep", "conv_stem", "window_attn_tc",
                                        "bwd_gemm", "bwd_gemm_tc", "bwd_misc", "bwd_layernorm", "bwd_sk_gate", "bwd_window_attn", "bwd_dwconv", "bwd_head",
                                        "bwd_patch_embed", "bwd_conv", "bwd_batchnorm", "loss_mask", "distill",
                                        "mlp_fc1_dw_tc", "optimizer"};
struct ProfRec { int tag; int n; cudaEvent_t e0, e1; };
bool g_prof = false;
std::mutex g_prof_mu;
std::vector<ProfRec> g_recs;

struct ProfScope {
  bool on; ProfRec r; cudaStream_t st;
  ProfScope(int tag, int n, cudaStream_t s) : on(g_prof), st(s) {
    if (!on) return;
    r.tag = tag; r.n = n;
    cudaEventCreate(&r.e0); cudaEventCreate(&r.e1);
    cudaEventRecord(r.e0, st);
  }
  ~ProfScope() {
    if (!on) return;
    cudaEventRecord(r.e1, st);
    std::lock_guard<std::mutex> g(g_prof_mu);
    g_recs.push_back(r);
  }
};

#define DPMN_RUN(tag, expr, n_kernels)       \
  do {                                       \
    int _rc;                                 \
    {                                        \
      ProfScope _ps(tag, n_kernels, st);     \
      _rc = (expr);                          \
    }                                        \
    if (_rc != 0) return _rc;                \
    g_launches.fetch_add(n_kernels);         \
  } while (0)

// The CMM's two encoder branches (cmm.py:121-133) are independent chains of small-grid kernels until the SE gate: in the
// fp32-structured forward and in the backward the second branch runs on a library-owned side stream per device, forked
// from / joined to the caller's stream by events, so the caller still sees plain stream semantics (everything enqueued by
// a call is ordered before whatever the caller enqueues next on its stream; works under stream capture as well).  Each
// branch has its own scratch.  DPMN_CMM_FORK=0: one stream.
struct BranchFork {
  cudaStream_t main = nullptr, aux = nullptr;
  cudaEvent_t ev = nullptr;
  bool on = false;
  static cudaStream_t aux_for_device(int slot) {
    static std::mutex mu;
    static cudaStream_t streams[64][3] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || slot < 0 || slot >= 3) return nullptr;
    std::lock_guard<std::mutex> g(mu);
    if (!streams[dev][slot] && cudaStreamCreateWithFlags(&streams[dev][slot], cudaStreamNonBlocking) != cudaSuccess)
      streams[dev][slot] = nullptr;
    return streams[dev][slot];
  }
  // the side stream waits for everything enqueued on `m` so far; false: stay on one stream.  slot 0: the CMM's side stream;
  // 1 / 2: the PGRM backward's (two cascades run their backwards concurrently on two caller streams: one side stream each,
  // picked by the caller stream's handle)
  bool begin(cudaStream_t m, int slot = 0) {
    static const bool enabled = !(getenv("DPMN_CMM_FORK") && atoi(getenv("DPMN_CMM_FORK")) == 0);
    main = m;
    if (!enabled) return false;
    aux = aux_for_device(slot);
    if (!aux || cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) { ev = nullptr; return false; }
    if (cudaEventRecord(ev, main) != cudaSuccess || cudaStreamWaitEvent(aux, ev, 0) != cudaSuccess) {
      cudaEventDestroy(ev); ev = nullptr;
      return false;
    }
    on = true;
    return true;
  }
  // the side stream additionally waits for everything enqueued on the caller's stream up to now
  cudaError_t aux_wait_main() {
    if (!on) return cudaSuccess;
    cudaEvent_t e2 = nullptr;
    cudaError_t e = cudaEventCreateWithFlags(&e2, cudaEventDisableTiming);
    if (e != cudaSuccess) return e;
    e = cudaEventRecord(e2, main);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(aux, e2, 0);
    cudaEventDestroy(e2);
    return e;
  }
  // the caller's stream waits for the side stream
  cudaError_t end() {
    if (!on) return cudaSuccess;
    on = false;
    cudaError_t e = cudaEventRecord(ev, aux);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(main, ev, 0);
    cudaEventDestroy(ev);
    ev = nullptr;
    return e;
  }
  ~BranchFork() { end(); }     // error returns inside the forked region still join
};

struct Bump {
  char* base;
  size_t off = 0, cap;
  Bump(void* p, size_t c) : base((char*)p), cap(c) {}
  template <typename T>
  T* take(size_t n) {
    off = (off + 255) & ~(size_t)255;
    T* r = reinterpret_cast<T*>(base + off);
    off += n * sizeof(T);
    return r;
  }
  bool ok() const { return off <= cap; }
};

// One contraction of the PGRM, dispatched on the arithmetic mode: fp32 FFMA tiles or tcgen05 (16-bit operands).
struct GemmCall {
  const void* A = nullptr; long long a_bs = 0; int lda = 0;
  const void* Bm = nullptr; long long b_bs = 0; int ldb = 0;
  void* C = nullptr; long long c_bs = 0; int ldc = 0; DType out_type = DT_F32;
  int M = 0, N = 0, K = 0, batch = 1;
  const float* bias = nullptr; long long bias_bs = 0; int bias_mode = 0;
  int act = 0;
  const float* residual = nullptr;
  float* colsum = nullptr;
  int ln_mode = 0; const float* ln_w = nullptr; const float* ln_b = nullptr; void* ln_out = nullptr; DType ln_type = DT_F16;
};

int run_gemm(int precision, const GemmCall& c, cudaStream_t st) {
  if (precision == DPMN_PREC_F32) {
    GemmSimtArgs g;
    g.A = (const float*)c.A; g.a_bs = c.a_bs; g.lda = c.lda; g.Bm = (const float*)c.Bm; g.b_bs = c.b_bs; g.ldb = c.ldb;
    g.C = (float*)c.C; g.c_bs = c.c_bs; g.ldc = c.ldc; g.M = c.M; g.N = c.N; g.K = c.K; g.batch = c.batch;
    g.bias = c.bias; g.bias_bs = c.bias_bs; g.bias_mode = c.bias_mode; g.act = c.act; g.residual = c.residual;
    g.colsum = c.colsum;
    return launch_gemm_simt(g, st);
  }
  GemmTcArgs g;
  g.A = c.A; g.a_bs = c.a_bs; g.lda = c.lda; g.Bm = c.Bm; g.b_bs = c.b_bs; g.ldb = c.ldb; g.op_type = (DType)precision;
  g.C = c.C; g.c_bs = c.c_bs; g.ldc = c.ldc; g.out_type = c.out_type; g.M = c.M; g.N = c.N; g.K = c.K; g.batch = c.batch;
  g.bias = c.bias; g.bias_bs = c.bias_bs; g.bias_mode = c.bias_mode; g.act = c.act; g.residual = c.residual;
  g.colsum = c.colsum;
  g.ln_mode = c.ln_mode; g.ln_w = c.ln_w; g.ln_b = c.ln_b; g.ln_out = c.ln_out; g.ln_type = c.ln_type;
  return launch_gemm_tc(g, st);
}

// weights of one block as the contraction kernels read them: the fp32 tensors themselves (F32 mode) or
// 16-bit copies staged in the workspace (tensor-core modes)
struct BlockOperands { const void *q_w, *kv_w, *sk_proj_w, *fc1_w, *fc2_w, *pw_w; };

struct PgrmWs {
  float *tq, *tkv, *colsum, *bias_b, *t1;
  void *ln, *q, *kv, *attn, *wb, *h, *dt;      // activation-typed (fp32 or 16-bit)
  void *lnq[2], *ln2;                          // fused-LayerNorm outputs (tensor-core modes)
  void* prep;                                  // staged weights when the caller passes no `prepared` buffer
  size_t bytes;
};

struct PgrmPrep {                              // staged weights of the tensor-core modes
  void* w16[DPMN_MAX_BLOCKS][6];               // q, kv, sk_proj, fc1, fc2, pw
  void* head_w;                                // conv_before_upsample.0 as [9][16][C] (rows 12..15 zero)
  float* head_shift;                           // its bias padded to 16
  size_t bytes;
};

PgrmPrep carve_pgrm_prep(const dpmn_pgrm_desc* d, void* base) {
  const size_t C = d->embed_dim, hid = d->mlp_hidden;
  Bump b(base, (size_t)-1);
  PgrmPrep w;
  memset(&w, 0, sizeof(w));
  const size_t n[6] = {C * C, 2 * C * C, C * C, hid * C, C * hid, hid * hid};
  for (int blk = 0; blk < DPMN_MAX_BLOCKS; ++blk)
    for (int i = 0; i < 6; ++i) w.w16[blk][i] = b.take<char>(n[i] * 2);
  w.head_w = b.take<char>(9 * 16 * C * 2);
  w.head_shift = b.take<float>(16);
  w.bytes = b.off + 256;
  return w;
}

PgrmWs carve_pgrm(const dpmn_pgrm_desc* d, void* ws) {
  const size_t B = d->batch, C = d->embed_dim, hid = d->mlp_hidden;
  const size_t L = (size_t)(d->img_h / d->patch) * (d->img_w / d->patch);
  const size_t es = d->precision == DPMN_PREC_F32 ? 4 : 2;
  const size_t col_tiles = d->precision == DPMN_PREC_F32 ? (L + kSimtTileM - 1) / kSimtTileM
                                                          : ((L + kTcTileM - 1) / kTcTileM) * 4;
  Bump b(ws, (size_t)-1);
  PgrmWs w;
  memset(&w, 0, sizeof(w));
  w.tq = b.take<float>(B * L * C);
  w.tkv = b.take<float>(B * L * C);
  w.ln = b.take<char>(B * L * C * es);
  w.q = b.take<char>(B * L * C * es);
  w.kv = b.take<char>(B * L * 2 * C * es);
  w.attn = b.take<char>(B * L * C * es);
  w.colsum = b.take<float>(B * col_tiles * C);
  w.wb = b.take<char>(B * C * C * es);
  w.bias_b = b.take<float>(B * C);
  w.h = b.take<char>(B * L * hid * es);
  w.dt = b.take<char>(B * L * hid * es);
  w.t1 = b.take<float>(B * L * 16);
  if (d->precision != DPMN_PREC_F32) {
    w.lnq[0] = b.take<char>(B * L * C * 2); w.lnq[1] = b.take<char>(B * L * C * 2); w.ln2 = b.take<char>(B * L * C * 2);
    if (d->prepared == nullptr) w.prep = b.take<char>(carve_pgrm_prep(d, nullptr).bytes);
  }
  w.bytes = b.off + 256;
  return w;
}

int check_pgrm(const dpmn_pgrm_desc* d) {
  if (!d) return DPMN_E_ARG;
  if (d->batch < 1 || d->patch < 1 || d->img_h % d->patch || d->img_w % d->patch) return DPMN_E_ARG;
  if (d->n_groups < 1 || d->n_groups > DPMN_MAX_GROUPS) return DPMN_E_ARG;
  if (d->embed_dim % d->n_groups || d->num_heads % d->n_groups) return DPMN_E_ARG;
  if ((d->embed_dim / d->n_groups) % (d->num_heads / d->n_groups)) return DPMN_E_ARG;
  if (d->q_chans != 2 && d->q_chans != 3) return DPMN_E_ARG;
  if (d->q_chans == 2 && (!d->prior_fusion_w || !d->prior_fusion_b)) return DPMN_E_ARG;
  if (d->n_mix < 1 || d->n_mix > DPMN_MAX_MIX) return DPMN_E_ARG;
  if (d->precision < DPMN_PREC_F32 || d->precision > DPMN_PREC_BF16) return DPMN_E_ARG;
  const int H = d->img_h / d->patch, W = d->img_w / d->patch;
  const int mn = H < W ? H : W;
  for (int g = 0; g < d->n_groups; ++g) {
    const int ws = d->window[g];
    // a window larger than min(H, W) is clamped by the reference but its bias table keeps the configured
    // size, so the reference's own forward raises (pgrm.py:127-151,234-236)
    if (ws < 1 || ws > mn) return DPMN_E_UNSUPPORTED;
    if (H % ws || W % ws) return DPMN_E_UNSUPPORTED;   // the reference's pad path cannot run
  }
  if (d->embed_dim % 32 || d->mlp_hidden % 32) return DPMN_E_UNSUPPORTED;
  return 0;
}

int pgrm_forward_impl(const dpmn_pgrm_desc* d, const float* x_q, const float* x_kv, float* out, void* workspace,
                      size_t workspace_bytes, void* stream, float* const* attn_core, float* const* block_out) {
  int rc = check_pgrm(d);
  if (rc) return rc;
  if (!x_q || !x_kv || !out || !workspace) return DPMN_E_ARG;
  const PgrmWs w = carve_pgrm(d, workspace);
  if (workspace_bytes < w.bytes) return DPMN_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;

  const int prec = d->precision;
  const DType at = (DType)prec;                       // storage type of the intermediate activations
  const int gtag = prec == DPMN_PREC_F32 ? T_GEMM : T_GEMM_TC;
  const int B = d->batch, C = d->embed_dim, hid = d->mlp_hidden, G = d->n_groups;
  const int H = d->img_h / d->patch, W = d->img_w / d->patch, L = H * W;
  const int rows = B * L;
  const long long plane = (long long)d->img_h * d->img_w;
  const long long xq_bs = d->x_q_batch_stride ? d->x_q_batch_stride : (long long)d->q_chans * plane;
  const long long xkv_bs = d->x_kv_batch_stride ? d->x_kv_batch_stride : 3LL * plane;

  // operands of the contractions: fp32 weights in place, or staged 16-bit copies
  BlockOperands ops[DPMN_MAX_BLOCKS];
  PgrmPrep pw;
  memset(&pw, 0, sizeof(pw));
  const int hp = d->hidden_size * d->patch * d->patch;
  const bool tc_head = prec != DPMN_PREC_F32 && hp <= 16 && C % 16 == 0;
  if (prec == DPMN_PREC_F32) {
    for (int blk = 0; blk < DPMN_MAX_BLOCKS; ++blk) {
      const dpmn_block_weights& bw = d->blocks[blk];
      ops[blk] = BlockOperands{bw.q_w, bw.kv_w, bw.sk_proj_w, bw.fc1_w, bw.fc2_w, bw.pw_w};
    }
  } else {
    pw = carve_pgrm_prep(d, d->prepared ? d->prepared : w.prep);
    ConvertBatch cb;
    const long long n[6] = {(long long)C * C, 2LL * C * C, (long long)C * C, (long long)hid * C, (long long)C * hid,
                            (long long)hid * hid};
    for (int blk = 0; blk < DPMN_MAX_BLOCKS; ++blk) {
      const dpmn_block_weights& bw = d->blocks[blk];
      const float* src[6] = {bw.q_w, bw.kv_w, bw.sk_proj_w, bw.fc1_w, bw.fc2_w, bw.pw_w};
      for (int i = 0; i < 6; ++i) {
        cb.src[cb.count] = src[i]; cb.dst[cb.count] = pw.w16[blk][i]; cb.n[cb.count] = n[i]; ++cb.count;
      }
      ops[blk] = BlockOperands{pw.w16[blk][0], pw.w16[blk][1], pw.w16[blk][2], pw.w16[blk][3], pw.w16[blk][4], pw.w16[blk][5]};
    }
    if (!(d->prepared && d->prepared_valid)) {
      DPMN_RUN(T_CONVERT, launch_convert_batch(cb, at, st), 1);
      if (tc_head) {
        PrepBatch pb;
        pb.seg[pb.count++] = PrepSeg{d->head0_w, pw.head_w, hp, C, 9, 0, 16};
        DPMN_RUN(T_PREP, launch_prep_weights(pb, at, st), 1);
        DPMN_CUDA_TRY(cudaMemsetAsync(pw.head_shift, 0, 16 * sizeof(float), st));
        DPMN_CUDA_TRY(cudaMemcpyAsync(pw.head_shift, d->head0_b, hp * sizeof(float), cudaMemcpyDeviceToDevice, st));
      }
    }
  }

  // LayerNorm fusion (tensor-core modes, C in {32,64,96,128}): every LayerNorm of the block is produced by the
  // kernel that produces its input -- the patch embed (norm1_q of both blocks, norm1_kv of block 0) or the
  // residual GEMM epilogue (norm2; norm1_kv of the next block) -- instead of by a pass of its own.
  const bool fuse_ln = prec != DPMN_PREC_F32 && gemm_tc_can_fuse_row_output(C);

  // K0: both streams share the patch-embed weights (pgrm.py:549-550)
  if (fuse_ln) {
    PatchEmbedLn eq;
    eq.count = 2; eq.type = at;
    for (int blk = 0; blk < 2; ++blk) { eq.out[blk] = w.lnq[blk]; eq.w[blk] = d->blocks[blk].norm1_q_w; eq.b[blk] = d->blocks[blk].norm1_q_b; }
    DPMN_RUN(T_PATCH_EMBED, launch_patch_embed(x_q, xq_bs, d->q_chans, d->q_chans == 2 ? d->prior_fusion_w : nullptr,
                                d->prior_fusion_b, d->pe_w, d->pe_b, d->pe_norm_w, d->pe_norm_b, nullptr, B, d->img_h,
                                d->img_w, d->patch, C, st, &eq), 1);
    PatchEmbedLn ek;
    ek.count = 1; ek.type = at; ek.out[0] = w.ln; ek.w[0] = d->blocks[0].norm1_kv_w; ek.b[0] = d->blocks[0].norm1_kv_b;
    DPMN_RUN(T_PATCH_EMBED, launch_patch_embed(x_kv, xkv_bs, 3, nullptr, nullptr, d->pe_w, d->pe_b, d->pe_norm_w, d->pe_norm_b,
                                w.tkv, B, d->img_h, d->img_w, d->patch, C, st, &ek), 1);
  } else {
  DPMN_RUN(T_PATCH_EMBED, launch_patch_embed(x_q, xq_bs, d->q_chans, d->q_chans == 2 ? d->prior_fusion_w : nullptr,
                              d->prior_fusion_b, d->pe_w, d->pe_b, d->pe_norm_w, d->pe_norm_b, w.tq, B, d->img_h,
                              d->img_w, d->patch, C, st), 1);
  DPMN_RUN(T_PATCH_EMBED, launch_patch_embed(x_kv, xkv_bs, 3, nullptr, nullptr, d->pe_w, d->pe_b, d->pe_norm_w, d->pe_norm_b,
                              w.tkv, B, d->img_h, d->img_w, d->patch, C, st), 1);
  }

  for (int blk = 0; blk < DPMN_MAX_BLOCKS; ++blk) {
    const dpmn_block_weights& bw = d->blocks[blk];
    const BlockOperands& op = ops[blk];
    // ---- K1 + K2: LayerNorms, q / kv projections, windowed attention core (pgrm.py:322-323,188,194,197-268)
    AttnArgs a;
    a.q = w.q; a.kv = w.kv; a.out = w.attn; a.io_type = at;
    a.q_ld = C; a.kv_ld = 2 * C; a.v_off = C; a.out_ld = C;
    a.B = B; a.H = H; a.W = W; a.C = C; a.n_groups = G; a.heads_per_group = d->num_heads / G;
    {
      const int mn = H < W ? H : W;
      for (int g = 0; g < G; ++g) {
        a.table[g] = bw.rpb_table[g];
        a.window[g] = d->window[g];
        a.shift[g] = (blk % 2 == 0 || mn <= d->window[g]) ? 0 : d->window[g] / 2;   // pgrm.py:148-150,362
      }
    }
    AttnTcArgs ta;
    ta.qw = w.q; ta.kw = w.kv; ta.vw = (const char*)w.kv + (size_t)rows * C * 2; ta.out = w.attn; ta.io_type = at;
    ta.B = B; ta.H = H; ta.W = W; ta.C = C; ta.n_groups = G; ta.heads_per_group = d->num_heads / G;
    for (int g = 0; g < G; ++g) { ta.table[g] = a.table[g]; ta.window[g] = a.window[g]; ta.shift[g] = a.shift[g]; }
    const bool tc_attn = prec != DPMN_PREC_F32 && attn_tc_supported(ta) && (C / G) % 32 == 0;   // the projection epilogue scatters 32-column chunks per group

    const void* q_in = fuse_ln ? w.lnq[blk] : w.ln;
    auto scatter_proj = [&](const void* A, const void* Wt, const float* bias, int N, void* dst0, void* dst1) {
      GemmTcArgs g;   // projection + roll + window_partition: rows land window-major per group
      g.A = A; g.lda = C; g.Bm = Wt; g.ldb = C; g.op_type = at; g.out_type = at;
      g.M = rows; g.N = N; g.K = C; g.bias = bias; g.bias_mode = 1;
      g.scatter = 1; g.scatter_C = C; g.scatter_G = G; g.scatter_H = H; g.scatter_W = W;
      for (int i = 0; i < G; ++i) { g.scatter_ws[i] = a.window[i]; g.scatter_shift[i] = a.shift[i]; }
      g.scatter_dst[0] = dst0; g.scatter_dst[1] = dst1;
      return g;
    };
    static const bool dual_on = !(getenv("DPMN_DUAL_QKV") && atoi(getenv("DPMN_DUAL_QKV")) == 0);
    if (tc_attn && fuse_ln && dual_on) {
      // both LayerNorm outputs exist already (patch embed / previous residual GEMM): q and kv projections in ONE launch
      const GemmTcArgs gq = scatter_proj(q_in, op.q_w, bw.q_b, C, w.q, nullptr);
      const GemmTcArgs gkv = scatter_proj(w.ln, op.kv_w, bw.kv_b, 2 * C, const_cast<void*>(ta.kw), const_cast<void*>(ta.vw));
      DPMN_RUN(T_GEMM_TC, launch_gemm_tc_dual(gq, gkv, st), 1);
      DPMN_RUN(T_ATTN_TC, launch_window_attn_tc(ta, st), 1);
    } else {
    if (!fuse_ln) DPMN_RUN(T_LAYERNORM, launch_layernorm(w.tq, bw.norm1_q_w, bw.norm1_q_b, w.ln, at, rows, C, st), 1);
    if (!tc_attn) {
      GemmCall g;
      g.A = q_in; g.lda = C; g.Bm = op.q_w; g.ldb = C; g.C = w.q; g.ldc = C; g.out_type = at;
      g.M = rows; g.N = C; g.K = C; g.bias = bw.q_b; g.bias_mode = 1;
      DPMN_RUN(gtag, run_gemm(prec, g, st), 1);
    } else {
      const GemmTcArgs g = scatter_proj(q_in, op.q_w, bw.q_b, C, w.q, nullptr);
      DPMN_RUN(T_GEMM_TC, launch_gemm_tc(g, st), 1);
    }
    if (!fuse_ln) DPMN_RUN(T_LAYERNORM, launch_layernorm(w.tkv, bw.norm1_kv_w, bw.norm1_kv_b, w.ln, at, rows, C, st), 1);
    if (!tc_attn) {
      GemmCall g;
      g.A = w.ln; g.lda = C; g.Bm = op.kv_w; g.ldb = C; g.C = w.kv; g.ldc = 2 * C; g.out_type = at;
      g.M = rows; g.N = 2 * C; g.K = C; g.bias = bw.kv_b; g.bias_mode = 1;
      DPMN_RUN(gtag, run_gemm(prec, g, st), 1);
      DPMN_RUN(T_WINDOW_ATTN, launch_window_attn_simt(a, st), G);
    } else {
      const GemmTcArgs g = scatter_proj(w.ln, op.kv_w, bw.kv_b, 2 * C, const_cast<void*>(ta.kw), const_cast<void*>(ta.vw));
      DPMN_RUN(T_GEMM_TC, launch_gemm_tc(g, st), 1);
      DPMN_RUN(T_ATTN_TC, launch_window_attn_tc(ta, st), 1);
    }
    }
    if (attn_core && attn_core[blk]) {
      if (prec == DPMN_PREC_F32) {
        DPMN_CUDA_TRY(cudaMemcpyAsync(attn_core[blk], w.attn, (size_t)rows * C * sizeof(float),
                                      cudaMemcpyDeviceToDevice, st));
      } else {
        DPMN_RUN(T_CONVERT, launch_widen(w.attn, at, attn_core[blk], (long long)rows * C, st), 1);
      }
    }
    // ---- K3: SK gate (pgrm.py:79-96): pass 1 pooled GELU(proj), fold, pass 2 per-image GEMM + residual
    const int tiles = prec == DPMN_PREC_F32 ? (L + kSimtTileM - 1) / kSimtTileM : ((L + kTcTileM - 1) / kTcTileM) * 4;
    {
      GemmCall g;
      g.A = w.attn; g.a_bs = (long long)L * C; g.lda = C; g.Bm = op.sk_proj_w; g.ldb = C;
      g.M = L; g.N = C; g.K = C; g.batch = B; g.bias = bw.sk_proj_b; g.bias_mode = 1; g.act = 1;
      g.colsum = w.colsum;
      DPMN_RUN(gtag, run_gemm(prec, g, st), 1);
    }
    DPMN_RUN(T_SK_GATE, launch_sk_gate(w.colsum, tiles, L, bw.sk_proj_w, bw.sk_proj_b, bw.sk_fc1_w, bw.sk_fc1_b, bw.sk_fc2_w,
                            bw.sk_fc2_b, bw.sk_head_w, bw.sk_head_b, w.wb, at, w.bias_b, B, C, G, st), 1);
    {
      GemmCall g;   // x_kv = shortcut + attn (pgrm.py:329), in place on the fp32 kv stream
      g.A = w.attn; g.a_bs = (long long)L * C; g.lda = C; g.Bm = w.wb; g.b_bs = (long long)C * C; g.ldb = C;
      g.C = w.tkv; g.c_bs = (long long)L * C; g.ldc = C; g.out_type = DT_F32; g.M = L; g.N = C; g.K = C; g.batch = B;
      g.bias = w.bias_b; g.bias_bs = C; g.bias_mode = 1; g.residual = w.tkv;
      if (fuse_ln) { g.ln_mode = 1; g.ln_w = bw.norm2_w; g.ln_b = bw.norm2_b; g.ln_out = w.ln2; g.ln_type = at; }
      DPMN_RUN(gtag, run_gemm(prec, g, st), 1);
    }
    // ---- K4: norm2 + Mlp (pgrm.py:330, 29-41)
    if (!fuse_ln) DPMN_RUN(T_LAYERNORM, launch_layernorm(w.tkv, bw.norm2_w, bw.norm2_b, w.ln, at, rows, C, st), 1);
    static const bool fused_a_on = !(getenv("DPMN_FUSED_MLP_A") && atoi(getenv("DPMN_FUSED_MLP_A")) == 0);
    if (prec != DPMN_PREC_F32 && fused_a_on && mlp_fc1_dw_supported(C, hid, L)) {
      // fc1 + GELU + depthwise 3x3 + GELU in one kernel: the hidden tensor never leaves the SM between the two
      DPMN_RUN(T_MLP_A, launch_mlp_fc1_dw(fuse_ln ? w.ln2 : w.ln, op.fc1_w, bw.fc1_b, bw.dw_w, bw.dw_b, w.dt, B, at, st), 1);
    } else {
      {
        GemmCall g;   // fc1 + GELU
        g.A = fuse_ln ? w.ln2 : w.ln; g.lda = C; g.Bm = op.fc1_w; g.ldb = C; g.C = w.h; g.ldc = hid; g.out_type = at;
        g.M = rows; g.N = hid; g.K = C; g.bias = bw.fc1_b; g.bias_mode = 1; g.act = 1;
        DPMN_RUN(gtag, run_gemm(prec, g, st), 1);
      }
      DPMN_RUN(T_DWCONV, launch_dwconv(w.h, w.dt, at, bw.dw_w, bw.dw_b, B, L, hid, st), 1);
    }
    {
      GemmCall g;   // pointwise conv: per image (hid x hid) * (hid x L), written (hid, L) = the raw view
      g.A = op.pw_w; g.lda = hid; g.Bm = w.dt; g.b_bs = (long long)L * hid; g.ldb = hid;
      g.C = w.h; g.c_bs = (long long)L * hid; g.ldc = L; g.out_type = at; g.M = hid; g.N = L; g.K = hid; g.batch = B;
      g.bias = bw.pw_b; g.bias_mode = 2;
      DPMN_RUN(gtag, run_gemm(prec, g, st), 1);
    }
    {
      GemmCall g;   // fc2 on the raw (L, hid) view + residual, in place on the fp32 kv stream
      g.A = w.h; g.lda = hid; g.Bm = op.fc2_w; g.ldb = hid; g.C = w.tkv; g.ldc = C; g.out_type = DT_F32;
      g.M = rows; g.N = C; g.K = hid; g.bias = bw.fc2_b; g.bias_mode = 1; g.residual = w.tkv;
      if (fuse_ln) {
        if (blk + 1 < DPMN_MAX_BLOCKS) {   // the next block's norm1_kv
          g.ln_mode = 1; g.ln_w = d->blocks[blk + 1].norm1_kv_w; g.ln_b = d->blocks[blk + 1].norm1_kv_b;
        } else {
          g.ln_mode = 2;                   // 16-bit copy of the final stream for the head conv
        }
        g.ln_out = w.ln; g.ln_type = at;
      }
      DPMN_RUN(gtag, run_gemm(prec, g, st), 1);
    }
    if (block_out && block_out[blk])
      DPMN_CUDA_TRY(cudaMemcpyAsync(block_out[blk], w.tkv, (size_t)rows * C * sizeof(float),
                                    cudaMemcpyDeviceToDevice, st));
  }

  // ---- K5: head (pgrm.py:559-564)
  if (tc_head) {
    // conv3x3 C -> hp as a tcgen05 implicit GEMM on a 16-bit copy of the token stream (tokens are NHWC already)
    if (!fuse_ln) DPMN_RUN(T_CONVERT, launch_convert(w.tkv, w.ln, at, (long long)rows * C, st), 1);
    ConvTcArgs a;
    a.op_type = at; a.n_src = 1;
    a.src[0].base = w.ln; a.src[0].sx = C; a.src[0].sy = (long long)W * C; a.src[0].sb = (long long)L * C;
    a.Cin = C; a.Cout = 16; a.B = B; a.G = 1; a.P = 1; a.Hm = H; a.Wm = W;
    a.n_taps = 9;
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) {
        ConvTap tp{}; tp.map = 0; tp.dy = (int8_t)(ky - 1); tp.dx = (int8_t)(kx - 1); tp.wslice = (int16_t)(ky * 3 + kx);
        a.taps[0][ky * 3 + kx] = tp;
      }
    a.w = pw.head_w; a.n_wslices = 9; a.os = 1; a.Ho = H; a.Wo = W;
    a.scale = nullptr; a.shift = pw.head_shift;
    a.dst[0].ptr = w.t1; a.dst[0].type = DT_F32; a.dst[0].ld = 16;
    DPMN_RUN(T_CONV_TC, launch_conv_tc(a, st), 1);
  } else {
    DPMN_RUN(T_HEAD, launch_head_conv1(w.tkv, d->head0_w, d->head0_b, w.t1, B, H, W, C, hp, st), 1);
  }
  MixArgs mix;
  mix.n_mix = d->n_mix;
  for (int i = 0; i < d->n_mix; ++i) {
    if (!d->mix_weight[i] || (i > 0 && !d->mix_input[i])) return DPMN_E_ARG;
    mix.w[i] = d->mix_weight[i];
    mix.in[i] = d->mix_input[i];
    mix.in_bs[i] = d->mix_input_batch_stride[i] ? d->mix_input_batch_stride[i] : (long long)d->hidden_size * plane;
  }
  DPMN_RUN(T_HEAD, launch_head_conv2_mix(w.t1, d->head1_w, d->head1_b, out, B, H, W, d->hidden_size, d->patch, mix, st), 1);
  return 0;
}

// ---------------------------------------------------------------------------------------------------
struct CmmWs {
  float* o[2][6];        // encoder outputs o1..o6 per branch (raw conv outputs)
  float* mid[2][4];      // EncodeBlock intermediates
  float* o_sc[2][6];     // BN affine of o2..o5 (index 1..4), nullptr otherwise
  float* o_sh[2][6];
  float* mid_sc[2][4];
  float* mid_sh[2][4];
  float* z;              // gated bottleneck (B, 16c, h/32, w/32)
  float* se_h;           // SE gate hidden activations (B, 4c)
  double* bn_stats[2];   // batch-statistics partial sums (launch_bn_affine), one per stream of the forked encoder section
  float* d6; float *d6_sc, *d6_sh;
  float* dmid[4]; float *dmid_sc[4], *dmid_sh[4];
  float* dout[4]; float *dout_sc[4], *dout_sh[4];
  ConvTcScratch tc;      // 16-bit modes: the convs of this fp32-structured path run on tcgen05 through an im2col
  ConvTcScratch tc2;     // the second encoder branch's scratch (it runs on the side stream, BranchFork)
  size_t bytes;
};

CmmWs carve_cmm(const dpmn_cmm_desc* d, void* ws) {
  const size_t B = d->batch, c = d->cnum, H = d->img_h, W = d->img_w;
  const size_t ch_o[6] = {c, 2 * c, 4 * c, 8 * c, 8 * c, 8 * c};
  Bump b(ws, (size_t)-1);
  CmmWs w;
  memset(&w, 0, sizeof(w));
  for (int br = 0; br < 2; ++br) {
    for (int l = 0; l < 6; ++l) {
      const size_t hw = (H >> l) * (W >> l);
      w.o[br][l] = b.take<float>(B * ch_o[l] * hw);
      if (l >= 1 && l <= 4) {
        w.o_sc[br][l] = b.take<float>(ch_o[l]);
        w.o_sh[br][l] = b.take<float>(ch_o[l]);
      }
    }
    for (int l = 0; l < 4; ++l) {   // EncodeBlock l+2: conv_a keeps ch_o[l] channels at half resolution
      const size_t hw = (H >> (l + 1)) * (W >> (l + 1));
      w.mid[br][l] = b.take<float>(B * ch_o[l] * hw);
      w.mid_sc[br][l] = b.take<float>(ch_o[l]);
      w.mid_sh[br][l] = b.take<float>(ch_o[l]);
    }
  }
  w.z = b.take<float>(B * 16 * c * (H >> 5) * (W >> 5));
  w.se_h = b.take<float>(B * 4 * c);
  for (int i = 0; i < 2; ++i) w.bn_stats[i] = b.take<double>(bn_stats_scratch_doubles((int)(16 * c)));
  w.d6 = b.take<float>(B * 8 * c * (H >> 4) * (W >> 4));
  w.d6_sc = b.take<float>(8 * c);
  w.d6_sh = b.take<float>(8 * c);
  const size_t ch_d[4] = {8 * c, 4 * c, 2 * c, c};   // de_5, de_4, de_3, de_2 output channels
  for (int i = 0; i < 4; ++i) {
    const size_t hw_in = (H >> (4 - i)) * (W >> (4 - i));
    w.dmid[i] = b.take<float>(B * ch_d[i] * hw_in);
    w.dmid_sc[i] = b.take<float>(ch_d[i]);
    w.dmid_sh[i] = b.take<float>(ch_d[i]);
    w.dout[i] = b.take<float>(B * ch_d[i] * hw_in * 4);
    w.dout_sc[i] = b.take<float>(ch_d[i]);
    w.dout_sh[i] = b.take<float>(ch_d[i]);
  }
  if (d->precision != DPMN_PREC_F32) {
    const size_t kmax = 27 * c > 16 * c ? 27 * c : 16 * c;              // de_1 (3c x 3x3) and the 4x4 convs at full resolution
    w.tc.t = (DType)d->precision;
    w.tc.col_bytes = B * H * W * ((kmax + 15) / 16 * 16) * 2;
    w.tc.col = b.take<char>(w.tc.col_bytes);
    w.tc.w16_bytes = (size_t)16 * c * 16 * 8 * c * 2 + 4096;             // de_6: (8c) x (16c * 16)
    w.tc.w16 = b.take<char>(w.tc.w16_bytes);
    w.tc.dy16_bytes = B * H * W * 3 * c * 2;                             // widest gradient: de_1's concat input
    w.tc.dy16 = b.take<char>(w.tc.dy16_bytes);
    w.tc.part_bytes = (size_t)128 << 20;
    w.tc.part = b.take<float>(w.tc.part_bytes / 4);
    w.tc.act_bytes = B * H * W * 3 * c * 2;                              // widest conv input: de_1's concat at full resolution
    w.tc.act = b.take<char>(w.tc.act_bytes);
    // second branch: same sizes (its data-gradient im2col of EncodeBlock 1 is as wide as the decoder's)
    w.tc2 = w.tc;
    w.tc2.col = b.take<char>(w.tc2.col_bytes);
    w.tc2.w16 = b.take<char>(w.tc2.w16_bytes);
    w.tc2.dy16 = b.take<char>(w.tc2.dy16_bytes);
    w.tc2.part = b.take<float>(w.tc2.part_bytes / 4);
    w.tc2.act = b.take<char>(w.tc2.act_bytes);
  }
  w.bytes = b.off + 256;
  return w;
}

int bn_affine(const dpmn_cmm_desc* d, const dpmn_bn& bn, const float* x, int ch, int hw, float* sc, float* sh,
              cudaStream_t st, double* stats_scratch) {
  if (!bn.w || !bn.b || !bn.running_mean || !bn.running_var) return DPMN_E_ARG;
  return launch_bn_affine(x, d->batch, ch, hw, bn.w, bn.b, bn.running_mean, bn.running_var, d->training,
                          d->training && d->update_running_stats, 1e-5f, sc, sh, st, stats_scratch);
}


// ---------------------------------------------------------------------------------------------------
// Tensor-core CMM (fp16 / bf16 operands): NHWC 16-bit activations, implicit-GEMM convs on tcgen05.
struct CmmTcWs {
  void* e[6];        // e[l], l = 1..5: LeakyReLU'd encoder outputs [2][B][H_l][W_l][C_l]
  void* mid[6];      // mid[l], l = 2..5: LeakyReLU'd EncodeBlock intermediates [2][B][H_l][W_l][C_{l-1}]
  void* cat[6];      // cat[l], l = 1..5: ReLU'd decoder input [B][H_l][W_l][Cd_l + 2 C_l]
  float* z6;         // en_6 outputs [2][B][hw][8c] fp32
  void* zg;          // ReLU(SE-gated bottleneck) [B][hw][16c]
  void* dmid[6];     // dmid[l], l = 5..2: ReLU'd DecodeBlock intermediates [B][H_l][W_l][Co_l]
  float* P;          // de_1 tap products [B*H*W][32]
  float *pooled, *se_hidden;   // SE gate scratch (B, 16c), (B, 4c)
  void *w_enc_a[4], *w_enc_b[4], *w_en6, *w_de6, *w_dec_a[4], *w_dec_b[4], *w_de1;
  float *s_enc_a[4], *t_enc_a[4], *s_enc_b[4], *t_enc_b[4], *t_en6, *s_en6, *s_de6, *t_de6;
  float *s_dec_a[4], *t_dec_a[4], *s_dec_b[4], *t_dec_b[4];
  size_t bytes;
};

struct CmmDims {
  int B, c, H, W, ci;
  int Hl(int l) const { return H >> (l - 1); }
  int Wl(int l) const { return W >> (l - 1); }
  int Cl(int l) const { return l == 1 ? c : l == 2 ? 2 * c : l == 3 ? 4 * c : 8 * c; }   // encoder channels, l = 1..6
  int Co(int l) const { return l >= 5 ? 8 * c : l == 4 ? 4 * c : l == 3 ? 2 * c : c; }   // decoder out channels, l = 6..2
  int Cd(int l) const { return Co(l + 1); }                                               // decoder channels entering cat[l]
  int Ccat(int l) const { return Cd(l) + 2 * Cl(l); }
};

size_t carve_cmm_tc_prep(const CmmDims& D, void* base, CmmTcWs& w) {
  Bump b(base, (size_t)-1);
  for (int l = 1; l <= 4; ++l) {
    w.w_enc_a[l - 1] = b.take<uint16_t>((size_t)2 * 16 * D.Cl(l) * D.Cl(l));
    w.w_enc_b[l - 1] = b.take<uint16_t>((size_t)2 * 9 * D.Cl(l + 1) * D.Cl(l));
    w.s_enc_a[l - 1] = b.take<float>(2 * D.Cl(l)); w.t_enc_a[l - 1] = b.take<float>(2 * D.Cl(l));
    w.s_enc_b[l - 1] = b.take<float>(2 * D.Cl(l + 1)); w.t_enc_b[l - 1] = b.take<float>(2 * D.Cl(l + 1));
  }
  w.w_en6 = b.take<uint16_t>((size_t)2 * 16 * 8 * D.c * 8 * D.c);
  w.s_en6 = b.take<float>(2 * 8 * D.c); w.t_en6 = b.take<float>(2 * 8 * D.c);
  w.w_de6 = b.take<uint16_t>((size_t)16 * 8 * D.c * 16 * D.c);
  w.s_de6 = b.take<float>(8 * D.c); w.t_de6 = b.take<float>(8 * D.c);
  for (int i = 0; i < 4; ++i) {
    const int l = 5 - i;
    w.w_dec_a[i] = b.take<uint16_t>((size_t)9 * D.Co(l) * D.Ccat(l));
    w.w_dec_b[i] = b.take<uint16_t>((size_t)16 * D.Co(l) * D.Co(l));
    w.s_dec_a[i] = b.take<float>(D.Co(l)); w.t_dec_a[i] = b.take<float>(D.Co(l));
    w.s_dec_b[i] = b.take<float>(D.Co(l)); w.t_dec_b[i] = b.take<float>(D.Co(l));
  }
  w.w_de1 = b.take<uint16_t>((size_t)32 * 3 * D.c);
  return b.off + 256;
}

CmmTcWs carve_cmm_tc(const dpmn_cmm_desc* d, void* ws) {
  const CmmDims D{d->batch, d->cnum, d->img_h, d->img_w, d->c_img};
  const size_t B = D.B;
  Bump b(ws, (size_t)-1);
  CmmTcWs w;
  memset(&w, 0, sizeof(w));
  for (int l = 1; l <= 5; ++l) {
    const size_t px = (size_t)D.Hl(l) * D.Wl(l);
    w.e[l] = b.take<uint16_t>(2 * B * px * D.Cl(l));
    if (l >= 2) w.mid[l] = b.take<uint16_t>(2 * B * px * D.Cl(l - 1));
    w.cat[l] = b.take<uint16_t>(B * px * D.Ccat(l));
    if (l >= 2) w.dmid[l] = b.take<uint16_t>(B * px * D.Co(l));
  }
  const size_t hw6 = (size_t)(D.H >> 5) * (D.W >> 5);
  w.z6 = b.take<float>(2 * B * hw6 * 8 * D.c);
  w.zg = b.take<uint16_t>(B * hw6 * 16 * D.c);
  w.P = b.take<float>(B * (size_t)D.H * D.W * 32);
  w.pooled = b.take<float>(B * 16 * D.c);
  w.se_hidden = b.take<float>(B * 4 * D.c);
  if (d->prepared != nullptr) {
    carve_cmm_tc_prep(D, d->prepared, w);
  } else {
    char* base = b.take<char>(0);
    const size_t n = carve_cmm_tc_prep(D, base, w);
    b.take<char>(n);
  }
  w.bytes = b.off + 256;
  return w;
}

// transposed-conv 4x4 stride 2 pad 1: output parity -> the two kernel taps that hit it and their input shift
void convt4_taps(ConvTap (&taps)[4][16]) {
  // oy = 2*iy - 1 + ky:  oy even -> ky 1 (iy = a), ky 3 (iy = a-1);  oy odd -> ky 0 (iy = a+1), ky 2 (iy = a)
  const int kk[2][2] = {{1, 3}, {0, 2}};
  const int dd[2][2] = {{0, -1}, {1, 0}};
  for (int py = 0; py < 2; ++py)
    for (int px = 0; px < 2; ++px) {
      int n = 0;
      for (int a = 0; a < 2; ++a)
        for (int bq = 0; bq < 2; ++bq) {
          ConvTap t{};
          t.map = 0; t.dy = (int8_t)dd[py][a]; t.dx = (int8_t)dd[px][bq];
          t.wslice = (int16_t)(kk[py][a] * 4 + kk[px][bq]);
          taps[py * 2 + px][n++] = t;
        }
    }
}

int cmm_forward_tc(const dpmn_cmm_desc* d, const float* x1, const float* x2, float* out, void* workspace,
                   size_t workspace_bytes, cudaStream_t st) {
  if (d->training) return DPMN_E_UNSUPPORTED;   // batch-statistics BatchNorm runs in the fp32 mode only (this build)
  if (d->cnum % 16 || d->c_img > 3) return DPMN_E_UNSUPPORTED;   // 64-channel k-blocks, whole UMMA k-steps, 9*c_img <= 32
  const CmmTcWs w = carve_cmm_tc(d, workspace);
  // split-K scratch of the deep layers: the de_1 tap-product buffer, which is only written by the last GEMM of the forward
  const size_t sk_floats = (size_t)d->batch * d->img_h * d->img_w * 32;
  if (workspace_bytes < w.bytes) return DPMN_E_WORKSPACE;
  const CmmDims D{d->batch, d->cnum, d->img_h, d->img_w, d->c_img};
  const DType t = (DType)d->precision;
  const int B = D.B, c = D.c, H = D.H, W = D.W;

  // ---- stage weights (16-bit, tap-major) and fold bias + eval BatchNorm into (scale, shift)
  if (!(d->prepared && d->prepared_valid)) {
    PrepBatch pb; FoldBatch fb;
    auto prep = [&](const float* src, void* dst, int Cout, int Cin, int kk, int tr) {
      pb.seg[pb.count++] = PrepSeg{src, dst, Cout, Cin, kk, tr, Cout};
    };
    auto fold = [&](const float* bias, const dpmn_bn* bn, float* sc, float* sh, int C) {
      fb.seg[fb.count++] = FoldSeg{bias, bn ? bn->w : nullptr, bn ? bn->b : nullptr, bn ? bn->running_mean : nullptr,
                                   bn ? bn->running_var : nullptr, sc, sh, C};
    };
    for (int l = 1; l <= 4; ++l) {
      const int Ci = D.Cl(l), Cn = D.Cl(l + 1);
      for (int g = 0; g < 2; ++g) {
        const dpmn_cmm_stage& s = d->enc[g][l - 1];
        prep(s.conv_a_w, (uint16_t*)w.w_enc_a[l - 1] + (size_t)g * 16 * Ci * Ci, Ci, Ci, 16, 0);
        prep(s.conv_b_w, (uint16_t*)w.w_enc_b[l - 1] + (size_t)g * 9 * Cn * Ci, Cn, Ci, 9, 0);
        fold(s.conv_a_b, &s.bn_a, w.s_enc_a[l - 1] + g * Ci, w.t_enc_a[l - 1] + g * Ci, Ci);
        fold(s.conv_b_b, &s.bn_b, w.s_enc_b[l - 1] + g * Cn, w.t_enc_b[l - 1] + g * Cn, Cn);
      }
    }
    for (int g = 0; g < 2; ++g) {
      prep(d->en6_w[g], (uint16_t*)w.w_en6 + (size_t)g * 16 * 64 * c * c, 8 * c, 8 * c, 16, 0);
      fold(d->en6_b[g], nullptr, w.s_en6 + g * 8 * c, w.t_en6 + g * 8 * c, 8 * c);
    }
    prep(d->de6_w, w.w_de6, 8 * c, 16 * c, 16, 1);
    fold(d->de6_b, &d->de6_bn, w.s_de6, w.t_de6, 8 * c);
    for (int i = 0; i < 4; ++i) {
      const int l = 5 - i;
      const dpmn_cmm_stage& s = d->dec[i];
      prep(s.conv_a_w, w.w_dec_a[i], D.Co(l), D.Ccat(l), 9, 1);
      prep(s.conv_b_w, w.w_dec_b[i], D.Co(l), D.Co(l), 16, 1);
      fold(s.conv_a_b, &s.bn_a, w.s_dec_a[i], w.t_dec_a[i], D.Co(l));
      fold(s.conv_b_b, &s.bn_b, w.s_dec_b[i], w.t_dec_b[i], D.Co(l));
    }
    DPMN_CUDA_TRY(cudaMemsetAsync(w.w_de1, 0, (size_t)32 * 3 * c * 2, st));
    prep(d->de1_w, w.w_de1, D.ci, 3 * c, 9, 1);
    DPMN_RUN(T_PREP, launch_prep_weights(pb, t, st), 1);
    DPMN_RUN(T_PREP, launch_fold_bn(fb, 1e-5f, st), 1);
  }

  // ---- en_1 (stem, 3 channels: SIMT)
  DPMN_RUN(T_CONV_STEM, launch_cmm_en1(x1, x2, d->en1_w[0], d->en1_b[0], d->en1_w[1], d->en1_b[1], w.e[1], w.cat[1], t, B, H,
                                       W, D.ci, c, st), 1);

  // ---- encoders, both branches per launch (G = 2)
  for (int l = 1; l <= 4; ++l) {
    const int Hi = D.Hl(l), Wi = D.Wl(l), Ci = D.Cl(l), Cn = D.Cl(l + 1);
    const int Ho = Hi / 2, Wo = Wi / 2;
    {
      ConvTcArgs a;   // LeakyReLU -> conv4x4 s2 d2 p3 -> BN -> (LeakyReLU)          cmm.py:41-45
      a.splitk_ws = w.P; a.splitk_floats = sk_floats; a.op_type = t; a.n_src = 1;
      a.src[0].base = (const uint16_t*)w.e[l] + ((size_t)Wi + 1) * Ci;   // odd rows / odd columns sub-grid
      a.src[0].sx = 2LL * Ci; a.src[0].sy = 2LL * Wi * Ci; a.src[0].sb = (long long)Hi * Wi * Ci;
      a.src[0].sg = (long long)B * Hi * Wi * Ci;
      a.Cin = Ci; a.Cout = Ci; a.B = B; a.G = 2; a.P = 1; a.Hm = Ho; a.Wm = Wo;
      a.n_taps = 16;
      for (int ky = 0; ky < 4; ++ky)
        for (int kx = 0; kx < 4; ++kx) {
          ConvTap tp{}; tp.map = 0; tp.dy = (int8_t)(ky - 2); tp.dx = (int8_t)(kx - 2); tp.wslice = (int16_t)(ky * 4 + kx);
          a.taps[0][ky * 4 + kx] = tp;
        }
      a.w = w.w_enc_a[l - 1]; a.n_wslices = 16; a.os = 1; a.Ho = Ho; a.Wo = Wo;
      a.scale = w.s_enc_a[l - 1]; a.shift = w.t_enc_a[l - 1];
      a.dst[0].ptr = w.mid[l + 1]; a.dst[0].type = t; a.dst[0].g_stride = (long long)B * Ho * Wo * Ci; a.dst[0].ld = Ci;
      a.dst[0].act = 1;
      DPMN_RUN(T_CONV_TC, launch_conv_tc(a, st), 1);
    }
    {
      ConvTcArgs a;   // conv3x3 -> BN; stored LeakyReLU'd for the next stage and ReLU'd for the decoder skip
      a.splitk_ws = w.P; a.splitk_floats = sk_floats; a.op_type = t; a.n_src = 1;
      a.src[0].base = w.mid[l + 1];
      a.src[0].sx = Ci; a.src[0].sy = (long long)Wo * Ci; a.src[0].sb = (long long)Ho * Wo * Ci;
      a.src[0].sg = (long long)B * Ho * Wo * Ci;
      a.Cin = Ci; a.Cout = Cn; a.B = B; a.G = 2; a.P = 1; a.Hm = Ho; a.Wm = Wo;
      a.n_taps = 9;
      for (int ky = 0; ky < 3; ++ky)
        for (int kx = 0; kx < 3; ++kx) {
          ConvTap tp{}; tp.map = 0; tp.dy = (int8_t)(ky - 1); tp.dx = (int8_t)(kx - 1); tp.wslice = (int16_t)(ky * 3 + kx);
          a.taps[0][ky * 3 + kx] = tp;
        }
      a.w = w.w_enc_b[l - 1]; a.n_wslices = 9; a.os = 1; a.Ho = Ho; a.Wo = Wo;
      a.scale = w.s_enc_b[l - 1]; a.shift = w.t_enc_b[l - 1];
      a.dst[0].ptr = w.e[l + 1]; a.dst[0].type = t; a.dst[0].g_stride = (long long)B * Ho * Wo * Cn; a.dst[0].ld = Cn;
      a.dst[0].act = 1;
      a.dst[1].ptr = w.cat[l + 1]; a.dst[1].type = t; a.dst[1].g_stride = 0; a.dst[1].ld = D.Ccat(l + 1);
      a.dst[1].ch_off = D.Cd(l + 1); a.dst[1].ch_g_off = Cn; a.dst[1].act = 2;
      DPMN_RUN(T_CONV_TC, launch_conv_tc(a, st), 1);
    }
  }
  const int H5 = D.Hl(5), W5 = D.Wl(5), C5 = 8 * c, H6 = H5 / 2, W6 = W5 / 2;
  {
    ConvTcArgs a;   // en_6: LeakyReLU -> conv4x4 s2 p1 (cmm.py:91-93): each tap reads one parity sub-grid
    a.splitk_ws = w.P; a.splitk_floats = sk_floats; a.op_type = t; a.n_src = 4;
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px) {
        ConvTcSrc& sc = a.src[py * 2 + px];
        sc.base = (const uint16_t*)w.e[5] + ((size_t)py * W5 + px) * C5;
        sc.sx = 2LL * C5; sc.sy = 2LL * W5 * C5; sc.sb = (long long)H5 * W5 * C5; sc.sg = (long long)B * H5 * W5 * C5;
      }
    a.Cin = C5; a.Cout = C5; a.B = B; a.G = 2; a.P = 1; a.Hm = H6; a.Wm = W6;
    a.n_taps = 16;
    const int par[4] = {1, 0, 1, 0}, sh[4] = {-1, 0, 0, 1};   // iy = 2*oy - 1 + ky
    for (int ky = 0; ky < 4; ++ky)
      for (int kx = 0; kx < 4; ++kx) {
        ConvTap tp{}; tp.map = (int8_t)(par[ky] * 2 + par[kx]); tp.dy = (int8_t)sh[ky]; tp.dx = (int8_t)sh[kx];
        tp.wslice = (int16_t)(ky * 4 + kx);
        a.taps[0][ky * 4 + kx] = tp;
      }
    a.w = w.w_en6; a.n_wslices = 16; a.os = 1; a.Ho = H6; a.Wo = W6;
    a.scale = nullptr; a.shift = w.t_en6;
    a.dst[0].ptr = w.z6; a.dst[0].type = DT_F32; a.dst[0].g_stride = (long long)B * H6 * W6 * C5; a.dst[0].ld = C5;
    DPMN_RUN(T_CONV_TC, launch_conv_tc(a, st), 1);
  }
  DPMN_RUN(T_SE_GATE, launch_se_gate_nhwc(w.z6, w.zg, t, d->fc1_w, d->fc1_b, d->fc2_w, d->fc2_b, w.pooled, w.se_hidden, B, C5,
                                          H6 * W6, 4 * c, st), 3);

  // ---- decoder
  auto convt4 = [&](const void* src, int Hi, int Wi, int Ci, const void* wt, int Co, const float* sc, const float* sh,
                    void* dst, int dst_ld) -> int {
    ConvTcArgs a;   // ReLU'd input -> convT4x4 s2 p1 -> BN -> ReLU into channel 0.. of the next concat buffer
    a.splitk_ws = w.P; a.splitk_floats = sk_floats; a.op_type = t; a.n_src = 1;
    a.src[0].base = src; a.src[0].sx = Ci; a.src[0].sy = (long long)Wi * Ci; a.src[0].sb = (long long)Hi * Wi * Ci;
    a.Cin = Ci; a.Cout = Co; a.B = B; a.G = 1; a.P = 4; a.Hm = Hi; a.Wm = Wi;
    a.n_taps = 4;
    convt4_taps(a.taps);
    a.w = wt; a.n_wslices = 16; a.os = 2; a.Ho = 2 * Hi; a.Wo = 2 * Wi;
    a.scale = sc; a.shift = sh;
    a.dst[0].ptr = dst; a.dst[0].type = t; a.dst[0].ld = dst_ld; a.dst[0].act = 2;
    return launch_conv_tc(a, st);
  };
  DPMN_RUN(T_CONV_TC, convt4(w.zg, H6, W6, 16 * c, w.w_de6, 8 * c, w.s_de6, w.t_de6, w.cat[5], D.Ccat(5)), 1);
  for (int i = 0; i < 4; ++i) {
    const int l = 5 - i;
    const int Hi = D.Hl(l), Wi = D.Wl(l), Cc = D.Ccat(l), Co = D.Co(l);
    {
      ConvTcArgs a;   // convT3x3 s1 p1 on the concat buffer = conv with mirrored taps (cmm.py:62)
      a.splitk_ws = w.P; a.splitk_floats = sk_floats; a.op_type = t; a.n_src = 1;
      a.src[0].base = w.cat[l]; a.src[0].sx = Cc; a.src[0].sy = (long long)Wi * Cc; a.src[0].sb = (long long)Hi * Wi * Cc;
      a.Cin = Cc; a.Cout = Co; a.B = B; a.G = 1; a.P = 1; a.Hm = Hi; a.Wm = Wi;
      a.n_taps = 9;
      for (int ky = 0; ky < 3; ++ky)
        for (int kx = 0; kx < 3; ++kx) {
          ConvTap tp{}; tp.map = 0; tp.dy = (int8_t)(1 - ky); tp.dx = (int8_t)(1 - kx); tp.wslice = (int16_t)(ky * 3 + kx);
          a.taps[0][ky * 3 + kx] = tp;
        }
      a.w = w.w_dec_a[i]; a.n_wslices = 9; a.os = 1; a.Ho = Hi; a.Wo = Wi;
      a.scale = w.s_dec_a[i]; a.shift = w.t_dec_a[i];
      a.dst[0].ptr = w.dmid[l]; a.dst[0].type = t; a.dst[0].ld = Co; a.dst[0].act = 2;
      DPMN_RUN(T_CONV_TC, launch_conv_tc(a, st), 1);
    }
    DPMN_RUN(T_CONV_TC, convt4(w.dmid[l], Hi, Wi, Co, w.w_dec_b[i], Co, w.s_dec_b[i], w.t_dec_b[i], w.cat[l - 1],
                               D.Ccat(l - 1)), 1);
  }
  // ---- de_1: 3 output channels -> the 9 taps ride in N (one pass over the concat buffer), then a gather
  {
    GemmTcArgs g;
    g.A = w.cat[1]; g.lda = 3 * c; g.Bm = w.w_de1; g.ldb = 3 * c; g.op_type = t;
    g.C = w.P; g.ldc = 32; g.out_type = DT_F32; g.M = B * H * W; g.N = 32; g.K = 3 * c;
    DPMN_RUN(T_GEMM_TC, launch_gemm_tc(g, st), 1);
    const long long blend_bs = d->blend_input_batch_stride ? d->blend_input_batch_stride : (long long)D.ci * H * W;
    DPMN_RUN(T_CONV_STEM, launch_de1_gather(w.P, 32, d->de1_b, out, B, H, W, D.ci, st, d->blend_input, blend_bs, d->blend_alpha), 1);
  }
  return 0;
}

// defined in api_bwd.inc (same unnamed namespace): the fp32 training sequence with Dropout / DropPath
int pgrm_stochastic_forward(const dpmn_pgrm_desc* d, const float* x_q, const float* x_kv, float* out, void* workspace,
                            size_t workspace_bytes, void* stream);
bool pgrm_is_stochastic(const dpmn_pgrm_desc* d);
size_t pgrm_bwd_workspace(const dpmn_pgrm_desc* d);

}  // namespace

extern "C" {

const char* dpmn_version(void) { return "dpmn_b200 0.1 (sm_100a)"; }

int dpmn_check_device(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return DPMN_E_DEVICE;
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return DPMN_E_DEVICE;
  return major == 10 ? 0 : DPMN_E_DEVICE;
}

uint64_t dpmn_launch_count(void) { return g_launches.load(); }

int dpmn_profile_enable(int32_t on) {
  std::lock_guard<std::mutex> g(g_prof_mu);
  for (auto& r : g_recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  g_recs.clear();
  g_prof = on != 0;
  return 0;
}

int32_t dpmn_profile_collect(int32_t* tags, int32_t* n_kernels, float* ms, int32_t cap) {
  std::lock_guard<std::mutex> g(g_prof_mu);
  int32_t n = 0;
  for (auto& r : g_recs) {
    if (cudaEventSynchronize(r.e1) != cudaSuccess) return -1;
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.e0, r.e1) != cudaSuccess) return -1;
    if (n < cap) { tags[n] = r.tag; n_kernels[n] = r.n; ms[n] = t; ++n; }
    cudaEventDestroy(r.e0); cudaEventDestroy(r.e1);
  }
  g_recs.clear();
  return n;
}

const char* dpmn_profile_tag_name(int32_t tag) { return (tag >= 0 && tag < T_COUNT) ? kTagNames[tag] : ""; }

size_t dpmn_abi_sizeof(int32_t which) {
  switch (which) {
    case 0: return sizeof(dpmn_block_weights);
    case 1: return sizeof(dpmn_pgrm_desc);
    case 2: return sizeof(dpmn_bn);
    case 3: return sizeof(dpmn_cmm_stage);
    case 4: return sizeof(dpmn_cmm_desc);
    case 5: return sizeof(dpmn_block_grads);
    case 6: return sizeof(dpmn_pgrm_grads);
    case 7: return sizeof(dpmn_cmm_grads);
    case 8: return sizeof(dpmn_distill_desc);
    case 9: return sizeof(dpmn_distill_grads);
  }
  return 0;
}

size_t dpmn_pgrm_workspace_bytes(const dpmn_pgrm_desc* d) {
  if (check_pgrm(d)) return 0;
  const size_t n = carve_pgrm(d, nullptr).bytes;
  if (pgrm_is_stochastic(d)) { const size_t t = pgrm_bwd_workspace(d); return t > n ? t : n; }
  return n;
}

size_t dpmn_pgrm_prepared_bytes(const dpmn_pgrm_desc* d) {
  if (check_pgrm(d) || d->precision == DPMN_PREC_F32) return 0;
  return carve_pgrm_prep(d, nullptr).bytes;
}

int dpmn_pgrm_forward(const dpmn_pgrm_desc* d, const float* x_q, const float* x_kv, float* out, void* workspace,
                      size_t workspace_bytes, void* stream) {
  if (d && pgrm_is_stochastic(d)) return pgrm_stochastic_forward(d, x_q, x_kv, out, workspace, workspace_bytes, stream);
  return pgrm_forward_impl(d, x_q, x_kv, out, workspace, workspace_bytes, stream, nullptr, nullptr);
}

int dpmn_pgrm_forward_probe(const dpmn_pgrm_desc* d, const float* x_q, const float* x_kv, float* out,
                            void* workspace, size_t workspace_bytes, void* stream,
                            float* const attn_core[DPMN_MAX_BLOCKS], float* const block_out[DPMN_MAX_BLOCKS]) {
  return pgrm_forward_impl(d, x_q, x_kv, out, workspace, workspace_bytes, stream, attn_core, block_out);
}

size_t dpmn_window_attn_workspace_bytes(int32_t, int32_t, int32_t, int32_t) { return 256; }

int dpmn_window_attn_forward(const void* q, const void* kv, void* out, const float* const rpb_table[DPMN_MAX_GROUPS],
                             int32_t batch, int32_t grid_h, int32_t grid_w, int32_t embed_dim, int32_t num_heads,
                             int32_t n_groups, const int32_t window[DPMN_MAX_GROUPS],
                             const int32_t shift[DPMN_MAX_GROUPS], int32_t precision, void* workspace,
                             size_t workspace_bytes, void* stream) {
  (void)workspace; (void)workspace_bytes;
  if (!q || !kv || !out || !rpb_table || !window || !shift) return DPMN_E_ARG;
  if (n_groups < 1 || n_groups > DPMN_MAX_GROUPS || batch < 1) return DPMN_E_ARG;
  if (embed_dim % n_groups || num_heads % n_groups) return DPMN_E_ARG;
  if (precision < DPMN_PREC_F32 || precision > DPMN_PREC_BF16) return DPMN_E_ARG;
  AttnArgs a;
  a.q = q; a.kv = kv; a.out = out; a.io_type = (DType)precision;
  a.q_ld = embed_dim; a.kv_ld = 2 * embed_dim; a.v_off = embed_dim; a.out_ld = embed_dim;
  a.B = batch; a.H = grid_h; a.W = grid_w; a.C = embed_dim; a.n_groups = n_groups;
  a.heads_per_group = num_heads / n_groups;
  for (int g = 0; g < n_groups; ++g) {
    if (!rpb_table[g]) return DPMN_E_ARG;
    if (shift[g] < 0 || shift[g] >= window[g]) return DPMN_E_ARG;
    a.table[g] = rpb_table[g];
    a.window[g] = window[g];
    a.shift[g] = shift[g];
  }
  cudaStream_t st = (cudaStream_t)stream;
  DPMN_RUN(T_WINDOW_ATTN, launch_window_attn_simt(a, st), n_groups);
  return 0;
}

int dpmn_window_attn_forward_windowed(const void* qw, const void* kw, const void* vw, void* out,
                                      const float* const rpb_table[DPMN_MAX_GROUPS], int32_t batch, int32_t grid_h,
                                      int32_t grid_w, int32_t embed_dim, int32_t num_heads, int32_t n_groups,
                                      const int32_t window[DPMN_MAX_GROUPS], const int32_t shift[DPMN_MAX_GROUPS],
                                      int32_t precision, void* stream) {
  return dpmn_window_attn_forward_windowed_train(qw, kw, vw, out, rpb_table, batch, grid_h, grid_w, embed_dim, num_heads, n_groups,
                                                 window, shift, precision, 0.f, 0, 0, stream);
}

int dpmn_window_attn_forward_windowed_train(const void* qw, const void* kw, const void* vw, void* out,
                                            const float* const rpb_table[DPMN_MAX_GROUPS], int32_t batch, int32_t grid_h,
                                            int32_t grid_w, int32_t embed_dim, int32_t num_heads, int32_t n_groups,
                                            const int32_t window[DPMN_MAX_GROUPS], const int32_t shift[DPMN_MAX_GROUPS],
                                            int32_t precision, float attn_drop, uint64_t seed, uint32_t site, void* stream) {
  if (!(attn_drop >= 0.f && attn_drop < 1.f)) return DPMN_E_ARG;
  if (!qw || !kw || !vw || !out || !rpb_table || !window || !shift) return DPMN_E_ARG;
  if (n_groups < 1 || n_groups > DPMN_MAX_GROUPS || batch < 1) return DPMN_E_ARG;
  if (embed_dim % n_groups || num_heads % n_groups) return DPMN_E_ARG;
  if (precision != DPMN_PREC_F16 && precision != DPMN_PREC_BF16) return DPMN_E_UNSUPPORTED;
  AttnTcArgs a;
  a.qw = qw; a.kw = kw; a.vw = vw; a.out = out; a.io_type = (DType)precision;
  a.B = batch; a.H = grid_h; a.W = grid_w; a.C = embed_dim; a.n_groups = n_groups; a.heads_per_group = num_heads / n_groups;
  for (int g = 0; g < n_groups; ++g) {
    if (!rpb_table[g]) return DPMN_E_ARG;
    if (shift[g] < 0 || shift[g] >= window[g]) return DPMN_E_ARG;
    a.table[g] = rpb_table[g]; a.window[g] = window[g]; a.shift[g] = shift[g];
  }
  a.p_drop = attn_drop; a.seed = seed; a.site = site;
  if (!attn_tc_supported(a)) return DPMN_E_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  DPMN_RUN(T_ATTN_TC, launch_window_attn_tc(a, st), 1);
  return 0;
}

int dpmn_window_attn_backward_windowed(const void* qw, const void* kw, const void* vw, const void* d_out16, float* dq, float* dkv,
                                       const float* const rpb_table[DPMN_MAX_GROUPS], float* const d_rpb_table[DPMN_MAX_GROUPS],
                                       int32_t batch, int32_t grid_h, int32_t grid_w, int32_t embed_dim, int32_t num_heads,
                                       int32_t n_groups, const int32_t window[DPMN_MAX_GROUPS], const int32_t shift[DPMN_MAX_GROUPS],
                                       int32_t precision, float attn_drop, uint64_t seed, uint32_t site, void* stream) {
  if (!qw || !kw || !vw || !d_out16 || !dq || !dkv || !rpb_table || !d_rpb_table || !window || !shift) return DPMN_E_ARG;
  if (n_groups < 1 || n_groups > DPMN_MAX_GROUPS || batch < 1) return DPMN_E_ARG;
  if (embed_dim % n_groups || num_heads % n_groups) return DPMN_E_ARG;
  if (!(attn_drop >= 0.f && attn_drop < 1.f)) return DPMN_E_ARG;
  if (precision != DPMN_PREC_F16 && precision != DPMN_PREC_BF16) return DPMN_E_UNSUPPORTED;
  AttnBwdTcArgs a;
  a.qw = qw; a.kw = kw; a.vw = vw; a.d_out16 = d_out16; a.dq = dq; a.dkv = dkv; a.io_type = (DType)precision;
  a.B = batch; a.H = grid_h; a.W = grid_w; a.C = embed_dim; a.n_groups = n_groups; a.heads_per_group = num_heads / n_groups;
  a.p_drop = attn_drop; a.seed = seed; a.site = site;
  for (int g = 0; g < n_groups; ++g) {
    if (!rpb_table[g] || !d_rpb_table[g]) return DPMN_E_ARG;
    if (shift[g] < 0 || shift[g] >= window[g]) return DPMN_E_ARG;
    a.table[g] = rpb_table[g]; a.d_table[g] = d_rpb_table[g]; a.window[g] = window[g]; a.shift[g] = shift[g];
  }
  if (!attn_bwd_tc_supported(a)) return DPMN_E_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  DPMN_RUN(T_BWD_ATTN, launch_window_attn_bwd_tc(a, st), 1);
  return 0;
}

size_t dpmn_gemm_nt_workspace_bytes(int32_t M, int32_t N, int32_t K, int32_t precision) {
  if (precision == DPMN_PREC_F32) return 256;
  return ((size_t)M * K + (size_t)N * K) * 2 + 1024;
}

int dpmn_gemm_nt(const float* A, const float* B, const float* bias, float* C, int32_t M, int32_t N, int32_t K,
                 int32_t precision, void* workspace, size_t workspace_bytes, void* stream) {
  if (!A || !B || !C || M < 1 || N < 1 || K < 1) return DPMN_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (precision == DPMN_PREC_F32) {
    GemmSimtArgs g;
    g.A = A; g.lda = K; g.Bm = B; g.ldb = K; g.C = C; g.ldc = N; g.M = M; g.N = N; g.K = K;
    g.bias = bias; g.bias_mode = bias ? 1 : 0;
    DPMN_RUN(T_GEMM, launch_gemm_simt(g, st), 1);
    return 0;
  }
  if (precision != DPMN_PREC_F16 && precision != DPMN_PREC_BF16) return DPMN_E_ARG;
  if (!workspace || workspace_bytes < dpmn_gemm_nt_workspace_bytes(M, N, K, precision)) return DPMN_E_WORKSPACE;
  const DType t = (DType)precision;
  Bump b(workspace, workspace_bytes);
  uint16_t* a16 = b.take<uint16_t>((size_t)M * K);
  uint16_t* b16 = b.take<uint16_t>((size_t)N * K);
  DPMN_RUN(T_CONVERT, launch_convert(A, a16, t, (long long)M * K, st), 1);
  DPMN_RUN(T_CONVERT, launch_convert(B, b16, t, (long long)N * K, st), 1);
  GemmTcArgs g;
  g.A = a16; g.lda = K; g.Bm = b16; g.ldb = K; g.op_type = t; g.C = C; g.ldc = N; g.out_type = DT_F32;
  g.M = M; g.N = N; g.K = K; g.bias = bias; g.bias_mode = bias ? 1 : 0;
  DPMN_RUN(T_GEMM_TC, launch_gemm_tc(g, st), 1);
  return 0;
}

static int check_cmm(const dpmn_cmm_desc* d) {
  if (!d) return DPMN_E_ARG;
  if (d->batch < 1 || d->c_img < 1 || d->cnum < 1) return DPMN_E_ARG;
  if (d->img_h % 32 || d->img_w % 32 || d->img_h < 32 || d->img_w < 32) return DPMN_E_UNSUPPORTED;
  if (d->precision < DPMN_PREC_F32 || d->precision > DPMN_PREC_BF16) return DPMN_E_ARG;
  return 0;
}

size_t dpmn_cmm_workspace_bytes(const dpmn_cmm_desc* d) {
  if (check_cmm(d)) return 0;
  if (d->precision != DPMN_PREC_F32 && !d->training) return carve_cmm_tc(d, nullptr).bytes;
  return carve_cmm(d, nullptr).bytes;
}

size_t dpmn_cmm_prepared_bytes(const dpmn_cmm_desc* d) {
  if (check_cmm(d) || d->precision == DPMN_PREC_F32) return 0;
  const CmmDims D{d->batch, d->cnum, d->img_h, d->img_w, d->c_img};
  CmmTcWs w;
  return carve_cmm_tc_prep(D, nullptr, w);
}

size_t dpmn_cmm_debug_bytes(const dpmn_cmm_desc* d, int32_t which) {
  if (check_cmm(d) || d->precision == DPMN_PREC_F32) return 0;
  const CmmDims D{d->batch, d->cnum, d->img_h, d->img_w, d->c_img};
  const size_t B = D.B;
  if (which >= 1 && which <= 5) return 2 * B * D.Hl(which) * D.Wl(which) * D.Cl(which) * 2;
  if (which >= 12 && which <= 15) { const int l = which - 10; return 2 * B * D.Hl(l) * D.Wl(l) * D.Cl(l - 1) * 2; }
  if (which >= 21 && which <= 25) { const int l = which - 20; return B * D.Hl(l) * D.Wl(l) * D.Ccat(l) * 2; }
  if (which >= 32 && which <= 35) { const int l = which - 30; return B * D.Hl(l) * D.Wl(l) * D.Co(l) * 2; }
  if (which == 40) return 2 * B * (size_t)(D.H >> 5) * (D.W >> 5) * 8 * D.c * 4;
  if (which == 41) return B * (size_t)(D.H >> 5) * (D.W >> 5) * 16 * D.c * 2;
  return 0;
}

int dpmn_cmm_debug_copy(const dpmn_cmm_desc* d, void* workspace, int32_t which, void* dst, size_t dst_bytes, void* stream) {
  const size_t n = dpmn_cmm_debug_bytes(d, which);
  if (n == 0 || !workspace || !dst || dst_bytes < n) return DPMN_E_ARG;
  const CmmTcWs w = carve_cmm_tc(d, workspace);
  const void* src = nullptr;
  if (which >= 1 && which <= 5) src = w.e[which];
  else if (which >= 12 && which <= 15) src = w.mid[which - 10];
  else if (which >= 21 && which <= 25) src = w.cat[which - 20];
  else if (which >= 32 && which <= 35) src = w.dmid[which - 30];
  else if (which == 40) src = w.z6;
  else if (which == 41) src = w.zg;
  DPMN_CUDA_TRY(cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}

// A conv of the fp32-structured CMM path: fp32 FFMA tiles, or (16-bit modes) the tcgen05 GEMM through a 16-bit im2col.
static int cmm_conv_any(const ConvArgs& a, const ConvTcScratch& tc, bool use_tc, cudaStream_t st) {
  if (use_tc && conv_tc_im2col_ok(a)) {
    const int rc = launch_conv_tc_im2col(a, tc, st);
    if (rc != -3) return rc;           // -3: scratch too small for this layer -> SIMT
  }
  return launch_conv_simt(a, st);
}

// The fp32-structured forward (raw conv outputs + BatchNorm affines kept in the workspace: what the backward reads).
// It is the fp32 mode, and in the 16-bit modes the train()-mode forward (batch-statistics BatchNorm) and the backward's
// recompute, with the convs on tensor cores.
static int cmm_forward_struct(const dpmn_cmm_desc* d, const float* x1, const float* x2, float* out, void* workspace,
                              size_t workspace_bytes, void* stream) {
  int rc = check_cmm(d);
  if (rc) return rc;
  if (!x1 || !x2 || !out || !workspace) return DPMN_E_ARG;
  const CmmWs w = carve_cmm(d, workspace);
  const bool use_tc = d->precision != DPMN_PREC_F32;
  if (workspace_bytes < w.bytes) return DPMN_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const int B = d->batch, c = d->cnum, H = d->img_h, W = d->img_w;
  const int ch_o[6] = {c, 2 * c, 4 * c, 8 * c, 8 * c, 8 * c};
  const float* xin[2] = {x1, x2};

  // ---- encoders (cmm.py:121-133): branch 2 on the side stream with its own scratch (BranchFork)
  BranchFork fork;
  const cudaStream_t st_main = st;
  const bool forked = fork.begin(st_main);
  for (int br = 0; br < 2; ++br) {
    st = (br == 1 && forked) ? fork.aux : st_main;                       // DPMN_RUN and the launches below use `st`
    const ConvTcScratch& tcs = (br == 1 && forked) ? w.tc2 : w.tc;
    double* bnst = w.bn_stats[(br == 1 && forked) ? 1 : 0];
    {
      ConvArgs a;   // en_1: conv3x3 c_img -> cnum, no activation before, no BN after
      a.n_seg = 1; a.in[0] = xin[br]; a.seg_ch[0] = d->c_img;
      const long long bs = br == 0 ? d->x1_batch_stride : d->x2_batch_stride;
      if (bs != 0 && bs != (long long)d->c_img * H * W) return DPMN_E_UNSUPPORTED;
      a.w = d->en1_w[br]; a.bias = d->en1_b[br]; a.out = w.o[br][0];
      a.B = B; a.Cin = d->c_img; a.H = H; a.W = W; a.Cout = c; a.Ho = H; a.Wo = W; a.k = 3; a.stride = 1; a.pad = 1;
      DPMN_RUN(use_tc ? T_CONV_TC : T_CONV, cmm_conv_any(a, tcs, use_tc, st), 1);
    }
    for (int l = 0; l < 4; ++l) {
      const dpmn_cmm_stage& s = d->enc[br][l];
      const int hi = H >> l, wi = W >> l, ho = hi / 2, wo = wi / 2;
      {
        ConvArgs a;   // LeakyReLU(0.2) -> conv4x4 stride 2 dilation 2 pad 3 (cmm.py:41-44)
        a.n_seg = 1; a.in[0] = w.o[br][l]; a.seg_ch[0] = ch_o[l];
        a.in_scale[0] = w.o_sc[br][l]; a.in_shift[0] = w.o_sh[br][l]; a.in_act = 1;
        a.w = s.conv_a_w; a.bias = s.conv_a_b; a.out = w.mid[br][l];
        a.B = B; a.Cin = ch_o[l]; a.H = hi; a.W = wi; a.Cout = ch_o[l]; a.Ho = ho; a.Wo = wo;
        a.k = 4; a.stride = 2; a.pad = 3; a.dil = 2;
        DPMN_RUN(use_tc ? T_CONV_TC : T_CONV, cmm_conv_any(a, tcs, use_tc, st), 1);
      }
      DPMN_RUN(T_BN, bn_affine(d, s.bn_a, w.mid[br][l], ch_o[l], ho * wo, w.mid_sc[br][l], w.mid_sh[br][l], st, bnst), 1);
      {
        ConvArgs a;   // LeakyReLU -> conv3x3 (cmm.py:46-49)
        a.n_seg = 1; a.in[0] = w.mid[br][l]; a.seg_ch[0] = ch_o[l];
        a.in_scale[0] = w.mid_sc[br][l]; a.in_shift[0] = w.mid_sh[br][l]; a.in_act = 1;
        a.w = s.conv_b_w; a.bias = s.conv_b_b; a.out = w.o[br][l + 1];
        a.B = B; a.Cin = ch_o[l]; a.H = ho; a.W = wo; a.Cout = ch_o[l + 1]; a.Ho = ho; a.Wo = wo;
        a.k = 3; a.stride = 1; a.pad = 1;
        DPMN_RUN(use_tc ? T_CONV_TC : T_CONV, cmm_conv_any(a, tcs, use_tc, st), 1);
      }
      DPMN_RUN(T_BN, bn_affine(d, s.bn_b, w.o[br][l + 1], ch_o[l + 1], ho * wo, w.o_sc[br][l + 1], w.o_sh[br][l + 1], st, bnst), 1);
    }
    {
      ConvArgs a;   // en_6: LeakyReLU -> conv4x4 stride 2 pad 1 (cmm.py:91-93)
      const int hi = H >> 4, wi = W >> 4;
      a.n_seg = 1; a.in[0] = w.o[br][4]; a.seg_ch[0] = ch_o[4];
      a.in_scale[0] = w.o_sc[br][4]; a.in_shift[0] = w.o_sh[br][4]; a.in_act = 1;
      a.w = d->en6_w[br]; a.bias = d->en6_b[br]; a.out = w.o[br][5];
      a.B = B; a.Cin = ch_o[4]; a.H = hi; a.W = wi; a.Cout = ch_o[5]; a.Ho = hi / 2; a.Wo = wi / 2;
      a.k = 4; a.stride = 2; a.pad = 1;
      DPMN_RUN(use_tc ? T_CONV_TC : T_CONV, cmm_conv_any(a, tcs, use_tc, st), 1);
    }
  }
  st = st_main;
  DPMN_CUDA_TRY(fork.end());
  // ---- SE gate (cmm.py:135-147)
  const int hb = H >> 5, wb = W >> 5;
  DPMN_RUN(T_SE_GATE, launch_se_gate(w.o[0][5], w.o[1][5], w.z, d->fc1_w, d->fc1_b, d->fc2_w, d->fc2_b, B, 8 * c, hb * wb,
                          4 * c, st, w.se_h), 2);
  // ---- decoder (cmm.py:149-159)
  {
    ConvArgs a;   // de_6: ReLU -> convT4x4 stride 2 pad 1 -> BN
    a.n_seg = 1; a.in[0] = w.z; a.seg_ch[0] = 16 * c; a.in_act = 2;
    a.w = d->de6_w; a.bias = d->de6_b; a.out = w.d6;
    a.B = B; a.Cin = 16 * c; a.H = hb; a.W = wb; a.Cout = 8 * c; a.Ho = 2 * hb; a.Wo = 2 * wb;
    a.k = 4; a.stride = 2; a.pad = 1; a.transposed = 1;
    DPMN_RUN(use_tc ? T_CONV_TC : T_CONV, cmm_conv_any(a, w.tc, use_tc, st), 1);
    DPMN_RUN(T_BN, bn_affine(d, d->de6_bn, w.d6, 8 * c, 4 * hb * wb, w.d6_sc, w.d6_sh, st, w.bn_stats[0]), 1);
  }
  const float* dprev = w.d6;
  const float *dprev_sc = w.d6_sc, *dprev_sh = w.d6_sh;
  int dprev_ch = 8 * c;
  const int ch_d[4] = {8 * c, 4 * c, 2 * c, c};
  for (int i = 0; i < 4; ++i) {
    const dpmn_cmm_stage& s = d->dec[i];
    const int lvl = 4 - i;                 // skip connection index into o[][]: o5, o4, o3, o2
    const int hi = H >> lvl, wi = W >> lvl;
    {
      ConvArgs a;   // ReLU -> convT3x3 on cat[dec, enc1 skip, enc2 skip] (cmm.py:61-64,150-157)
      a.n_seg = 3;
      a.in[0] = dprev; a.seg_ch[0] = dprev_ch; a.in_scale[0] = dprev_sc; a.in_shift[0] = dprev_sh;
      for (int br = 0; br < 2; ++br) {
        a.in[1 + br] = w.o[br][lvl]; a.seg_ch[1 + br] = ch_o[lvl];
        a.in_scale[1 + br] = w.o_sc[br][lvl]; a.in_shift[1 + br] = w.o_sh[br][lvl];
      }
      a.in_act = 2;
      a.w = s.conv_a_w; a.bias = s.conv_a_b; a.out = w.dmid[i];
      a.B = B; a.Cin = dprev_ch + 2 * ch_o[lvl]; a.H = hi; a.W = wi; a.Cout = ch_d[i]; a.Ho = hi; a.Wo = wi;
      a.k = 3; a.stride = 1; a.pad = 1; a.transposed = 1;
      DPMN_RUN(use_tc ? T_CONV_TC : T_CONV, cmm_conv_any(a, w.tc, use_tc, st), 1);
    }
    DPMN_RUN(T_BN, bn_affine(d, s.bn_a, w.dmid[i], ch_d[i], hi * wi, w.dmid_sc[i], w.dmid_sh[i], st, w.bn_stats[0]), 1);
    {
      ConvArgs a;   // ReLU -> convT4x4 stride 2 pad 1 (cmm.py:66-69)
      a.n_seg = 1; a.in[0] = w.dmid[i]; a.seg_ch[0] = ch_d[i];
      a.in_scale[0] = w.dmid_sc[i]; a.in_shift[0] = w.dmid_sh[i]; a.in_act = 2;
      a.w = s.conv_b_w; a.bias = s.conv_b_b; a.out = w.dout[i];
      a.B = B; a.Cin = ch_d[i]; a.H = hi; a.W = wi; a.Cout = ch_d[i]; a.Ho = 2 * hi; a.Wo = 2 * wi;
      a.k = 4; a.stride = 2; a.pad = 1; a.transposed = 1;
      DPMN_RUN(use_tc ? T_CONV_TC : T_CONV, cmm_conv_any(a, w.tc, use_tc, st), 1);
    }
    DPMN_RUN(T_BN, bn_affine(d, s.bn_b, w.dout[i], ch_d[i], 4 * hi * wi, w.dout_sc[i], w.dout_sh[i], st, w.bn_stats[0]), 1);
    dprev = w.dout[i]; dprev_sc = w.dout_sc[i]; dprev_sh = w.dout_sh[i]; dprev_ch = ch_d[i];
  }
  {
    ConvArgs a;   // de_1: ReLU -> convT3x3 3*cnum -> c_img (cmm.py:113-116,158-159)
    a.n_seg = 3;
    a.in[0] = dprev; a.seg_ch[0] = c; a.in_scale[0] = dprev_sc; a.in_shift[0] = dprev_sh;
    a.in[1] = w.o[0][0]; a.seg_ch[1] = c;
    a.in[2] = w.o[1][0]; a.seg_ch[2] = c;
    a.in_act = 2;
    a.w = d->de1_w; a.bias = d->de1_b; a.out = out;
    a.B = B; a.Cin = 3 * c; a.H = H; a.W = W; a.Cout = d->c_img; a.Ho = H; a.Wo = W;
    a.k = 3; a.stride = 1; a.pad = 1; a.transposed = 1;
    DPMN_RUN(use_tc ? T_CONV_TC : T_CONV, cmm_conv_any(a, w.tc, use_tc, st), 1);
  }
  if (d->blend_input != nullptr) {   // alpha * CMM + (1 - alpha) * PSN image (super_resolution.py:449,705)
    const long long per = (long long)d->c_img * H * W;
    DPMN_RUN(T_CONV_STEM, launch_alpha_blend(out, d->blend_input, d->blend_input_batch_stride ? d->blend_input_batch_stride : per,
                                             d->blend_alpha, B, per, st), 1);
  }
  return 0;
}


int dpmn_cmm_forward(const dpmn_cmm_desc* d, const float* x1, const float* x2, float* out, void* workspace,
                     size_t workspace_bytes, void* stream) {
  int rc = check_cmm(d);
  if (rc) return rc;
  if (!x1 || !x2 || !out || !workspace) return DPMN_E_ARG;
  if (d->precision != DPMN_PREC_F32 && !d->training) {        // eval in a 16-bit mode: the fused NHWC tensor-core path
    for (int br = 0; br < 2; ++br) {
      const long long bs = br == 0 ? d->x1_batch_stride : d->x2_batch_stride;
      if (bs != 0 && bs != (long long)d->c_img * d->img_h * d->img_w) return DPMN_E_UNSUPPORTED;
    }
    return cmm_forward_tc(d, x1, x2, out, workspace, workspace_bytes, (cudaStream_t)stream);
  }
  return cmm_forward_struct(d, x1, x2, out, workspace, workspace_bytes, stream);
}

}  // extern "C"

#include "api_bwd.inc"
