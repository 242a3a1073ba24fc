// C-ABI implementation (include/bimodal_b200.h): the scorer runtime that drives one decoder pass
// over all candidates with a shared prefix KV, plus thin wrappers for the stand-alone kernels.
// The scorer owns its activation workspace and the per-layer prefix K/V (device memory from
// cudaMalloc); weights are borrowed pointers (the HF model's parameters or packed copies made by
// the Python host).
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/bimodal_b200.h"
#include "gemm_host.cuh"
#include "kernels.h"

using namespace bma;

namespace {

thread_local std::string g_last_error;
unsigned long long g_launches = 0;

int fail(cudaError_t e, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = std::string(buf) + ": " + cudaGetErrorString(e);
  return static_cast<int>(e == cudaSuccess ? cudaErrorUnknown : e);
}
// BMA_CHECK: any CUDA call; BMA_LAUNCH: a kernel launch of this library (counted).
#define BMA_CHECK(expr, what)                             \
  do {                                                    \
    cudaError_t e__ = (expr);                             \
    if (e__ != cudaSuccess) return fail(e__, "%s", what); \
  } while (0)
#define BMA_LAUNCH(expr, what)                            \
  do {                                                    \
    cudaError_t e__ = (expr);                             \
    ++g_launches;                                         \
    if (e__ != cudaSuccess) return fail(e__, "%s", what); \
  } while (0)

// Kernel classes of the scoring pass, for the optional per-class CUDA-event profile.
enum : int { PC_EMBED = 0, PC_NORM, PC_GEMM_QKV, PC_ATTN, PC_GEMM_O, PC_GEMM_GLU, PC_GEMM_DOWN,
             PC_GEMM_CE, PC_FINAL, PC_COUNT };

inline cudaStream_t as_stream(void* s) { return static_cast<cudaStream_t>(s); }
inline const __nv_bfloat16* bf(const void* p) { return static_cast<const __nv_bfloat16*>(p); }
inline __nv_bfloat16* bf(void* p) { return static_cast<__nv_bfloat16*>(p); }

constexpr int kCeBlockN = 256;

// K-block stagger of the token-gradient GEMM (BMA_TOKGRAD_STAGGER overrides, 0 = off; A/B runs)
inline int token_grad_stagger() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("BMA_TOKGRAD_STAGGER");
    v = e != nullptr ? atoi(e) : 7;
  }
  return v;
}

}  // namespace

struct bma_comm;   // comm.cu: NCCL communicator owned by this library (one rank per process)

struct bma_scorer {
  bma_model_desc desc;
  std::vector<bma_layer_weights> layers;
  const __nv_bfloat16* norm_final = nullptr;
  const __nv_bfloat16* lm_head = nullptr;
  const __nv_bfloat16* embed = nullptr;
  long long ld_embed = 0;
  int q_dim = 0, kv_dim = 0, qkv_dim = 0;
  int qk_dim = 0;                        // row stride of the Q|K buffers (V is stored transposed)
  long long ld_vt = 0, ld_vtp = 0;       // row strides (in keys) of the transposed V buffers
  int prefix_len = 0;
  // fused prefill: the prefix embeddings handed to bma_scorer_prefill_deferred wait here until the
  // next score call runs them inside its first chunk (or an explicit prefill consumes them)
  const __nv_bfloat16* pending_prefix = nullptr;
  int pending_P = 0;
  // multi-rank (one process per GPU): prefix rows are split over the ranks of `comm`, their K/V
  // all-gathered per layer; the emulated group (tests) uses vrank/vsize with local copies instead
  bma_comm* comm = nullptr;
  int vrank = 0, vsize = 1;
  __nv_bfloat16 *xchg_send = nullptr, *xchg_recv = nullptr;
  size_t xchg_elems = 0;
  // peer-memory exchange (NVLink): every rank's prefix buffers + flag words mapped into this process
  // (cudaIpc handles between the ranks of a box, plain pointers between the scorers of a test group)
  int peer_n = 0, peer_rank = 0;
  bool peer_ipc = false;
  void* peer_qk[8] = {};
  void* peer_vt[8] = {};
  void* peer_flags[8] = {};
  unsigned long long* flags_dev = nullptr;     // [8] one word per sending rank
  unsigned int* done_dev = nullptr;            // block counter of the scatter kernel
  unsigned long long epoch = 0;                // exchanges issued so far (same on every rank)
  // workspace
  __nv_bfloat16 *x = nullptr, *xn = nullptr, *qk = nullptr, *vt = nullptr, *attn = nullptr, *act = nullptr;
  __nv_bfloat16* prefix_qk = nullptr;    // [layers][max_prefix][qk_dim]   Q|K rows of the prefix
  __nv_bfloat16* prefix_vt = nullptr;    // [layers][kv_dim][ld_vtp]       V^T of the prefix
  float2* ce_part = nullptr;
  int* ce_arg = nullptr;
  __nv_bfloat16 *x_c = nullptr, *attn_c = nullptr;   // the scored rows of the last layer, made contiguous
  float* tgt_logit = nullptr;
  // ragged-row plan (suffix-head sharing)
  int2 *row_info = nullptr, *row_meta = nullptr;
  int* cand_vstart = nullptr;
  int2* cand_tmp = nullptr;
  int dedupe = 1;                        // identical candidates share one set of rows (bma_scorer_set_dedupe)
  int early_signal = 0;                  // split prefill over peers: the QKV GEMM raises the flags itself, publishing tiles first
  int fused_publish = 1;                 // prefix / base K/V leave through the QKV GEMM's epilogue (bma_scorer_set_fused_publish)
  int* total_rows_dev = nullptr;         // live row count of the current chunk (read by every kernel of the pass)
  int* total_rows_host = nullptr;        // pinned
  unsigned long long* rows_accum_dev = nullptr;   // rows run by the ragged chunks of the last score call
  long long* base_dev = nullptr;         // staging for the *_host entry
  long long last_rows = 0;               // rows of the uniform (non-ragged) chunks of the last score call
  long long* ids_dev = nullptr;          // staging for the *_host entry
  float* loss_dev = nullptr;
  unsigned long long* key_dev = nullptr;
  int* flag_dev = nullptr;
  size_t ws_bytes = 0;
  int ce_rows_cap = 0;
  // optional profile: event pairs around every launch, summed per class on read
  bool profiling = false;
  std::vector<cudaEvent_t> ev_start, ev_stop;
  std::vector<int> ev_class;
  std::vector<double> ev_flops;
  size_t ev_used = 0;
};

int comm_allgather_bytes(bma_comm* c, const void* send, void* recv, size_t bytes_per_rank, cudaStream_t st);
int comm_rank(const bma_comm* c);
int comm_size(const bma_comm* c);

extern "C" int bma_scorer_clear_peers(bma_scorer* s);

namespace {
struct ProfScope {
  bma_scorer* s; cudaStream_t st; size_t slot; bool on;
  ProfScope(bma_scorer* s_, int cls, double flops, cudaStream_t st_) : s(s_), st(st_), slot(0), on(false) {
    if (!s->profiling) return;
    if (s->ev_used >= s->ev_start.size()) {       // grow instead of silently dropping launches
      const size_t n = s->ev_start.size() + 4096;
      const size_t old = s->ev_start.size();
      s->ev_start.resize(n); s->ev_stop.resize(n); s->ev_class.resize(n); s->ev_flops.resize(n);
      for (size_t i = old; i < n; ++i) { cudaEventCreate(&s->ev_start[i]); cudaEventCreate(&s->ev_stop[i]); }
    }
    on = true;
    slot = s->ev_used++;
    s->ev_class[slot] = cls;
    s->ev_flops[slot] = flops;
    cudaEventRecord(s->ev_start[slot], st);
  }
  ~ProfScope() { if (on) cudaEventRecord(s->ev_stop[slot], st); }
};
inline double gemm_flops(const GemmParams& p) { return 2.0 * p.M * static_cast<double>(p.N) * p.K; }
}  // namespace

namespace {

template <typename T>
bool dev_alloc(T** p, size_t count, size_t* total) {
  if (cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T)) != cudaSuccess) return false;
  // zero once: a ragged pass bounds its TMA maps by the row CAPACITY, so tiles may touch rows past
  // the live count; those must hold finite values (0 * NaN would poison P.V in the attention)
  if (cudaMemset(*p, 0, count * sizeof(T)) != cudaSuccess) return false;
  *total += count * sizeof(T);
  return true;
}

// CE partials: [vocab tiles][rows] (max, sum-exp) + [rows] target logits, for `rows` scored rows of
// one chunk. Grow-only; 262k-vocab models need 8 KB per scored row, so this is sized from B*T, not
// from the tail-row capacity.
int ensure_ce_rows(bma_scorer* s, int rows) {
  if (rows <= s->ce_rows_cap) return 0;
  const int n_ce_tiles = (s->desc.vocab + kCeBlockN - 1) / kCeBlockN;
  BMA_CHECK(cudaDeviceSynchronize(), "ensure_ce_rows: sync");
  cudaFree(s->ce_part); cudaFree(s->ce_arg); cudaFree(s->tgt_logit); cudaFree(s->x_c); cudaFree(s->attn_c);
  s->ce_part = nullptr; s->ce_arg = nullptr; s->tgt_logit = nullptr; s->x_c = nullptr; s->attn_c = nullptr;
  s->ce_rows_cap = 0;
  size_t bytes = 0;
  if (!dev_alloc(&s->ce_part, static_cast<size_t>(n_ce_tiles) * rows, &bytes) ||
      !dev_alloc(&s->ce_arg, static_cast<size_t>(n_ce_tiles) * rows, &bytes) ||
      !dev_alloc(&s->tgt_logit, static_cast<size_t>(rows), &bytes) ||
      !dev_alloc(&s->x_c, static_cast<size_t>(rows) * s->desc.d_model, &bytes) ||
      !dev_alloc(&s->attn_c, static_cast<size_t>(rows) * s->q_dim, &bytes))
    return fail(cudaGetLastError(), "ensure_ce_rows: cudaMalloc");
  s->ws_bytes += bytes;
  s->ce_rows_cap = rows;
  return 0;
}

// One decoder pass over `rows` rows. Uniform layout: sequences of `seq` rows at positions
// pos_base + (r % seq). prefill == true: rows are the prefix itself (their q/k/v land in
// prefix_qkv and the last layer stops after its K/V are produced).
//   Llama layer:  x += o(attn(norm_attn x));                x += down(glu(norm_mlp x))
//   Gemma layer:  x += norm_post_attn(o(attn(norm_attn x))); x += norm_post_mlp(down(glu(norm_mlp x)))
// In the Gemma form the residual adds cannot ride in the GEMM epilogue (a row norm sits in
// between), so each (post-norm + add + next pre-norm) triple is one fused row kernel.
// Ragged layout (n_share > 0; row_info / row_meta filled by plan_rows): rows are
//   [ pre_n prefix positions (fused prefill) | the base suffix's n_share.. rows | candidates' own rows ].
// Every layer publishes the K/V of the prefix rows and of the base's first n_share rows into the
// per-layer prefix buffers before its attention runs (`publish`): a local copy on one rank, a pack +
// all-gather + unpack when the prefix positions are split over ranks. `rows` is then an upper bound
// and `rows_dev` the live count every kernel reads on the device -- no host round trip per chunk.
// scored != null (candidate pass): only the T rows per candidate that predict the target are needed
// after the LAST layer's attention -- no later layer reads the other rows -- so that layer's o_proj,
// norms and MLP run on those rows alone, gathered into x_c / attn_c (they stay the residual stream
// for the final norm and the LM head).
struct ScoredRows { int n, T, ce_row0; const int* vstart; };

struct Pass {
  int rows = 0;                    // rows (upper bound when rows_dev != null)
  const int* rows_dev = nullptr;
  int seq = 1, pos_base = 0;
  bool prefill = false;
  int n_share = 0;                 // > 0: ragged rows
  bool publish = false;            // this chunk carries prefix rows / the base: publish their K/V per layer
  int pre0 = 0, pre_n = 0;         // fused prefill: prefix positions [pre0, pre0 + pre_n) are rows [0, pre_n)
  int pre_chunk = 0;               // prefix positions per rank (split prefill), 0 = not split
  const ScoredRows* scored = nullptr;
  // running state of the pass
  int cur_rows = 0;
  const int* cur_rows_dev = nullptr;
  __nv_bfloat16* xres = nullptr;
  const __nv_bfloat16* attn_in = nullptr;
};

inline void pass_start(bma_scorer* s, Pass& ps) {
  ps.cur_rows = ps.rows; ps.cur_rows_dev = ps.rows_dev;
  ps.xres = s->x; ps.attn_in = s->attn;
}
inline __nv_bfloat16* layer_pre_qk(bma_scorer* s, int l) {
  return s->prefix_qk + (size_t)l * s->desc.max_prefix * s->qk_dim;
}
inline __nv_bfloat16* layer_pre_vt(bma_scorer* s, int l) {
  return s->prefix_vt + (size_t)l * s->kv_dim * s->ld_vtp;
}

// Fused publish (default): the QKV GEMM's epilogue stores the K/V of the prefix rows and of the base rows
// straight into the per-layer prefix buffers (of every rank when the prefix is split over NVLink peers),
// so the per-layer exchange is part of the GEMM. BMA_FUSED_PUBLISH=0 selects the separate copy / scatter
// kernels (A/B runs); the NCCL all-gather exchange always uses pack + unpack.
inline int fused_publish_default() {
  const char* e = getenv("BMA_FUSED_PUBLISH");
  return (e != nullptr && atoi(e) == 0) ? 0 : 1;
}
inline bool publish_in_epilogue(const bma_scorer* s, const Pass& ps) {
  if (!ps.publish || ps.prefill || !s->fused_publish) return false;
  if (ps.pre_chunk > 0 && s->peer_n <= 1) return false;      // split prefill through NCCL: pack / all-gather / unpack
  return true;
}

// Early signal (GemmParams::pub_flag): only with the fused publish, over peer memory, K/V column tiles aligned.
inline bool early_signal(const bma_scorer* s, const Pass& ps) {
  return publish_in_epilogue(s, ps) && ps.pre_chunk > 0 && s->peer_n > 1 && s->early_signal != 0 &&
         s->q_dim % 256 == 0 && ps.pre_n + ps.n_share > 0;      // (a rank that publishes nothing signals by kernel)
}

// pre-norm + fused QKV projection (+ q/k-norm + rotary, V transposed)
int layer_front(bma_scorer* s, int l, Pass& ps, cudaStream_t st) {
  const bma_model_desc& D = s->desc;
  const int d = D.d_model;
  const bma_layer_weights& W = s->layers[l];
  const bool sandwich = D.sandwich_norm != 0;
  __nv_bfloat16* qk = ps.prefill ? layer_pre_qk(s, l) : s->qk;
  __nv_bfloat16* vt = ps.prefill ? layer_pre_vt(s, l) : s->vt;
  const long long ldvt = ps.prefill ? s->ld_vtp : s->ld_vt;
  if (!sandwich || l == 0) {     // sandwich layers > 0 got xn from the previous layer's add_norm
    ProfScope pr(s, PC_NORM, 0, st);
    BMA_LAUNCH(launch_rmsnorm(s->x, d, bf(W.norm_attn), s->xn, d, ps.cur_rows, d, D.rms_eps, D.norm_mode,
                              0, 0, 0, nullptr, st, ps.cur_rows_dev), "rmsnorm(attn)");
  }
  GemmParams p{};
  p.M = ps.cur_rows; p.M_dev = ps.cur_rows_dev; p.N = s->qkv_dim; p.K = d; p.group_m = group_m_for(p.K);
  p.C = qk; p.ldc = s->qk_dim;
  p.Vt = vt; p.ldvt = ldvt;
  p.rope_cos = W.rope_cos; p.rope_sin = W.rope_sin;
  p.rope_cols = s->qk_dim;
  p.rows_per_seq = ps.seq; p.pos_base = ps.pos_base;
  if (ps.n_share > 0) p.row_info = s->row_info;
  if (D.qk_norm) {
    p.q_norm_w = bf(W.q_norm); p.k_norm_w = bf(W.k_norm);
    p.q_cols = s->q_dim; p.qk_norm_eps = D.rms_eps;
  }
  if (publish_in_epilogue(s, ps)) {
    // the prefix rows' and the base rows' K/V leave through this GEMM's epilogue: into this rank's per-layer
    // prefix buffers and, when the prefix positions are split over the ranks of the box, into every peer's
    const bool peers = ps.pre_chunk > 0 && s->peer_n > 1;
    p.pub_rows = ps.pre_n + ps.n_share; p.pub_pre_n = ps.pre_n;
    p.pub_key0 = ps.pre0; p.pub_base_key = s->prefix_len;
    p.pub_n = peers ? s->peer_n : 1; p.pub_self = peers ? s->peer_rank : 0;
    p.pub_kcol0 = s->q_dim; p.pub_ldk = s->qk_dim; p.pub_ldvt = s->ld_vtp;
    const size_t qk_off = (size_t)l * D.max_prefix * s->qk_dim + s->q_dim;
    const size_t vt_off = (size_t)l * s->kv_dim * s->ld_vtp;
    for (int q = 0; q < p.pub_n; ++q) {
      p.pub_k[q] = (peers ? static_cast<__nv_bfloat16*>(s->peer_qk[q]) : s->prefix_qk) + qk_off;
      p.pub_vt[q] = (peers ? static_cast<__nv_bfloat16*>(s->peer_vt[q]) : s->prefix_vt) + vt_off;
    }
    if (early_signal(s, ps)) {         // this exchange's flag is raised by the GEMM itself (its publishing tiles run first)
      ++s->epoch;
      for (int q = 0; q < p.pub_n; ++q) p.pub_flag[q] = static_cast<unsigned long long*>(s->peer_flags[q]);
      p.pub_epoch = s->epoch;
      p.pub_done = s->done_dev;
    }
  }
  ProfScope pr(s, PC_GEMM_QKV, gemm_flops(p), st);
  if (D.head_dim == 128) {
    BMA_LAUNCH((launch_gemm_big<EPI_ROPE, 128>(s->xn, d, bf(W.w_qkv), d, p, st)), "gemm(qkv+rope)");
  } else if (D.head_dim == 64) {
    BMA_LAUNCH((launch_gemm_big<EPI_ROPE, 64>(s->xn, d, bf(W.w_qkv), d, p, st)), "gemm(qkv+rope)");
  } else {
    BMA_LAUNCH((launch_gemm_big<EPI_ROPE, 256>(s->xn, d, bf(W.w_qkv), d, p, st)), "gemm(qkv+rope)");
  }
  return 0;
}

// K rows / V^T columns [0, n) of (k_src, vt_src) -> keys [at, at + n) of (k_dst, vt_dst)
int copy_kv(bma_scorer* s, const __nv_bfloat16* k_src, long long ldk, const __nv_bfloat16* vt_src, long long ldvt,
            __nv_bfloat16* k_dst, long long ldkd, __nv_bfloat16* vt_dst, long long ldvtd, int n, int at,
            cudaStream_t st) {
  if (n <= 0) return 0;
  BMA_LAUNCH(launch_append_prefix_kv(k_src, ldk, vt_src, ldvt, k_dst, ldkd, vt_dst, ldvtd, n, s->kv_dim, at, st),
             "append_prefix_kv");
  return 0;
}

// Split prefill, step 1: this rank's prefix-position K/V of layer l -> its contiguous send block
// [pre_chunk x kv_dim K rows | kv_dim x pre_chunk V^T].
int layer_pack(bma_scorer* s, int l, const Pass& ps, cudaStream_t st) {
  (void)l;
  ProfScope pr(s, PC_EMBED, 0, st);
  return copy_kv(s, s->qk + s->q_dim, s->qk_dim, s->vt, s->ld_vt, s->xchg_send, s->kv_dim,
                 s->xchg_send + (size_t)ps.pre_chunk * s->kv_dim, ps.pre_chunk, ps.pre_n, 0, st);
}
// The base suffix's first n_share rows (they follow the prefix rows) -> prefix keys [P, P + n_share).
int layer_publish_base(bma_scorer* s, int l, const Pass& ps, int P, cudaStream_t st) {
  ProfScope pr(s, PC_EMBED, 0, st);
  return copy_kv(s, s->qk + (size_t)ps.pre_n * s->qk_dim + s->q_dim, s->qk_dim, s->vt + ps.pre_n, s->ld_vt,
                 layer_pre_qk(s, l) + s->q_dim, s->qk_dim, layer_pre_vt(s, l), s->ld_vtp, ps.n_share, P, st);
}
// Everything a single rank publishes in one copy: prefix rows [0, pre_n) (pre0 == 0) + the base rows.
int layer_publish_local(bma_scorer* s, int l, const Pass& ps, int P, cudaStream_t st) {
  ProfScope pr(s, PC_EMBED, 0, st);
  return copy_kv(s, s->qk + s->q_dim, s->qk_dim, s->vt, s->ld_vt, layer_pre_qk(s, l) + s->q_dim, s->qk_dim,
                 layer_pre_vt(s, l), s->ld_vtp, ps.pre_n + ps.n_share, P - ps.pre_n, st);
}

// attention, o_proj, MLP. Returns 1 when the pass is complete after this layer's front half
// (prefill of the last layer).
int layer_back(bma_scorer* s, int l, Pass& ps, cudaStream_t st) {
  const bma_model_desc& D = s->desc;
  const int d = D.d_model;
  const bma_layer_weights& W = s->layers[l];
  const bool sandwich = D.sandwich_norm != 0;
  const bool ragged = ps.n_share > 0;
  const int P = ps.prefill ? 0 : s->prefix_len;
  __nv_bfloat16* pre_qk = layer_pre_qk(s, l);
  __nv_bfloat16* pre_vt = layer_pre_vt(s, l);
  __nv_bfloat16* qk = ps.prefill ? pre_qk : s->qk;
  __nv_bfloat16* vt = ps.prefill ? pre_vt : s->vt;
  const long long ldvt = ps.prefill ? s->ld_vtp : s->ld_vt;
  {
    AttnParams a{};
    a.q = qk; a.ldq = s->qk_dim;
    a.k_tail = qk + s->q_dim; a.ldk = s->qk_dim;
    a.vt_tail = vt; a.ldvt = ldvt;
    a.k_pre = pre_qk + s->q_dim; a.ldkp = s->qk_dim;
    a.vt_pre = pre_vt; a.ldvtp = s->ld_vtp;
    a.out = s->attn; a.ldo = s->q_dim;
    a.M = ps.cur_rows; a.M_dev = ps.cur_rows_dev; a.S = ps.seq; a.P = P + ps.n_share;
    if (ragged) { a.row_meta = s->row_meta; a.P_shared = P; a.pre_rows = ps.pre_n; }
    a.kv_group = D.n_heads / D.n_kv_heads;
    a.head_dim = D.head_dim;
    a.scale_log2 = D.attn_scale * 1.4426950408889634f;
    // QK^T and PV over the visible keys: ~4 * rows * (P + seq/2) * q_dim
    ProfScope pr(s, PC_ATTN, 4.0 * ps.cur_rows * (P + 0.5 * (ps.seq + ps.n_share)) * s->q_dim, st);
    BMA_LAUNCH(launch_prefix_tail_attention(a, D.n_heads, st), "attention");
  }
  const ScoredRows* scored = ps.scored;
  if (scored != nullptr && l == D.n_layers - 1) {
    ProfScope pr(s, PC_EMBED, 0, st);
    BMA_LAUNCH(launch_gather_rows(s->attn, s->q_dim, s->attn_c, s->q_dim, scored->n, s->q_dim, scored->T, ps.seq,
                                  scored->ce_row0, scored->vstart, st), "gather(attn rows)");
    BMA_LAUNCH(launch_gather_rows(s->x, d, s->x_c, d, scored->n, d, scored->T, ps.seq, scored->ce_row0,
                                  scored->vstart, st), "gather(residual rows)");
    ps.cur_rows = scored->n; ps.cur_rows_dev = nullptr;
    ps.xres = s->x_c;
    ps.attn_in = s->attn_c;
  }
  const int rows = ps.cur_rows;
  const int* rows_dev = ps.cur_rows_dev;
  __nv_bfloat16* xres = ps.xres;
  {
    GemmParams p{};
    p.M = rows; p.M_dev = rows_dev; p.N = d; p.K = s->q_dim; p.group_m = group_m_for(p.K);
    if (sandwich) { p.C = s->xn; p.ldc = d; }                       // y: normed before the add
    else { p.C = xres; p.ldc = d; p.R = xres; p.ldr = d; }
    ProfScope pr(s, PC_GEMM_O, gemm_flops(p), st);
    BMA_LAUNCH((launch_gemm_big<EPI_STORE>(ps.attn_in, s->q_dim, bf(W.w_o), s->q_dim, p, st)),
               "gemm(o_proj)");
  }
  {
    ProfScope pr(s, PC_NORM, 0, st);
    if (sandwich) {
      BMA_LAUNCH(launch_add_norm(s->xn, d, bf(W.norm_post_attn), xres, d, bf(W.norm_mlp), s->xn, d,
                                 rows, d, D.rms_eps, D.norm_mode, st, rows_dev), "add_norm(post_attn, pre_mlp)");
    } else {
      BMA_LAUNCH(launch_rmsnorm(xres, d, bf(W.norm_mlp), s->xn, d, rows, d, D.rms_eps, D.norm_mode,
                                0, 0, 0, nullptr, st, rows_dev), "rmsnorm(mlp)");
    }
  }
  {
    GemmParams p{};
    p.M = rows; p.M_dev = rows_dev; p.N = 2 * D.d_ff; p.K = d; p.group_m = group_m_for(p.K);
    p.C = s->act; p.ldc = D.d_ff; p.act = D.act;
    ProfScope pr(s, PC_GEMM_GLU, gemm_flops(p), st);
    BMA_LAUNCH((launch_gemm_big<EPI_GLU>(s->xn, d, bf(W.w_gate_up), d, p, st)), "gemm(gate_up)");
  }
  {
    GemmParams p{};
    p.M = rows; p.M_dev = rows_dev; p.N = d; p.K = D.d_ff; p.group_m = group_m_for(p.K);
    if (sandwich) { p.C = s->xn; p.ldc = d; }
    else { p.C = xres; p.ldc = d; p.R = xres; p.ldr = d; }
    ProfScope pr(s, PC_GEMM_DOWN, gemm_flops(p), st);
    BMA_LAUNCH((launch_gemm_big<EPI_STORE>(s->act, D.d_ff, bf(W.w_down), D.d_ff, p, st)),
               "gemm(down)");
  }
  if (sandwich) {
    const __nv_bfloat16* next = (l + 1 < D.n_layers) ? bf(s->layers[l + 1].norm_attn) : nullptr;
    ProfScope pr(s, PC_NORM, 0, st);
    BMA_LAUNCH(launch_add_norm(s->xn, d, bf(W.norm_post_mlp), xres, d, next, s->xn, d, rows, d,
                               D.rms_eps, D.norm_mode, st, rows_dev), "add_norm(post_mlp, next input)");
  }
  return 0;
}

// Peer-memory exchange of layer l, send side: this rank's prefix-position K/V go straight into every
// rank's layer-l prefix buffers (its own too), then its flag is raised on every rank.
int layer_scatter_peers(bma_scorer* s, int l, const Pass& ps, cudaStream_t st) {
  void* pk[8]; void* pv[8];
  const size_t qk_off = ((size_t)l * s->desc.max_prefix * s->qk_dim + s->q_dim) * sizeof(__nv_bfloat16);
  const size_t vt_off = (size_t)l * s->kv_dim * s->ld_vtp * sizeof(__nv_bfloat16);
  for (int q = 0; q < s->peer_n; ++q) {
    pk[q] = static_cast<char*>(s->peer_qk[q]) + qk_off;
    pv[q] = static_cast<char*>(s->peer_vt[q]) + vt_off;
  }
  ++s->epoch;
  ProfScope pr(s, PC_EMBED, 0, st);
  BMA_LAUNCH(launch_prefix_scatter(s->qk + s->q_dim, s->qk_dim, s->vt, s->ld_vt, pk, pv, s->peer_flags, s->peer_n,
                                   s->peer_rank, s->qk_dim, s->ld_vtp, ps.pre_n, ps.pre0, s->kv_dim, s->done_dev,
                                   s->epoch, st), "prefix_scatter");
  return 0;
}
// ... receive side: wait until every rank's flag has reached this exchange, then append the base rows locally.
int layer_wait_peers(bma_scorer* s, int l, const Pass& ps, int P, cudaStream_t st) {
  {
    ProfScope pr(s, PC_EMBED, 0, st);
    BMA_LAUNCH(launch_prefix_wait(s->flags_dev, s->peer_n, s->epoch, st, s->flags_dev + 8), "prefix_wait");
  }
  if (publish_in_epilogue(s, ps)) return 0;          // the base rows went out with the GEMM
  return layer_publish_base(s, l, ps, P, st);
}
// Fused publish: the QKV GEMM has already stored this rank's slice into every rank's buffers; what is left
// of the exchange is ONE launch that raises this rank's flag on every peer (stream order + a system-scope
// fence put the GEMM's peer stores before it) and waits for the other ranks' flags.
int layer_signal_peers(bma_scorer* s, bool wait, cudaStream_t st, bool count_wait = true) {
  ++s->epoch;
  ProfScope pr(s, PC_EMBED, 0, st);
  BMA_LAUNCH(launch_prefix_signal_wait(s->peer_flags, s->flags_dev, s->peer_n, s->peer_rank, s->epoch, 1, wait ? 1 : 0,
                                       st, count_wait ? s->flags_dev + 8 : nullptr), "prefix_signal_wait");
  return 0;
}
int layer_exchange_peers(bma_scorer* s, int l, const Pass& ps, int P, cudaStream_t st) {
  if (early_signal(s, ps)) {           // signalled from inside the QKV GEMM: only the wait is left
    ProfScope pr(s, PC_EMBED, 0, st);
    BMA_LAUNCH(launch_prefix_signal_wait(s->peer_flags, s->flags_dev, s->peer_n, s->peer_rank, s->epoch, 0, 1, st,
                                         s->flags_dev + 8), "prefix_wait(early signal)");
    return 0;
  }
  if (publish_in_epilogue(s, ps)) return layer_signal_peers(s, true, st);
  int rc = layer_scatter_peers(s, l, ps, st);
  if (rc != 0) return rc;
  return layer_wait_peers(s, l, ps, P, st);
}
// Barrier over the ranks (no data): nobody writes into a rank's prefix buffers while that rank may still
// be reading them for an earlier pass.
int peers_barrier(bma_scorer* s, cudaStream_t st) {
  // (not counted as exchange wait: this one absorbs the ranks' START skew -- host jitter, the previous step's tail --
  // which says nothing about their speed)
  if (s->fused_publish) return layer_signal_peers(s, true, st, false);
  Pass none;
  none.pre_n = 0; none.pre0 = 0;
  int rc = layer_scatter_peers(s, 0, none, st);
  if (rc != 0) return rc;
  BMA_LAUNCH(launch_prefix_wait(s->flags_dev, s->peer_n, s->epoch, st, nullptr), "prefix_wait");
  return 0;
}

// A whole pass on ONE scorer (no other scorer takes part in the prefix exchange except through `comm`).
int run_layers(bma_scorer* s, Pass& ps, cudaStream_t st) {
  const bma_model_desc& D = s->desc;
  pass_start(s, ps);
  const int P = s->prefix_len;
  for (int l = 0; l < D.n_layers; ++l) {
    int rc = layer_front(s, l, ps, st);
    if (rc != 0) return rc;
    if (ps.prefill && l == D.n_layers - 1) break;   // only this layer's K/V are needed
    if (ps.publish) {
      if (ps.pre_chunk > 0 && s->peer_n > 1) {       // split prefill over NVLink peer memory
        rc = layer_exchange_peers(s, l, ps, P, st);
      } else if (ps.pre_chunk > 0) {                 // split prefill through an NCCL all-gather
        rc = layer_pack(s, l, ps, st);
        if (rc != 0) return rc;
        const size_t blk = 2 * (size_t)ps.pre_chunk * s->kv_dim;
        {
          ProfScope pr(s, PC_EMBED, 0, st);
          rc = comm_allgather_bytes(s->comm, s->xchg_send, s->xchg_recv, blk * sizeof(__nv_bfloat16), st);
          if (rc != 0) return fail(cudaErrorUnknown, "prefix K/V all-gather (NCCL)");
        }
        {                                            // every rank's slice + the base rows in ONE launch
          ProfScope pr(s, PC_EMBED, 0, st);
          BMA_LAUNCH(launch_unpack_prefix_kv(s->xchg_recv, static_cast<long long>(blk), ps.pre_chunk,
                                             comm_size(s->comm), P, s->qk + s->q_dim, s->qk_dim, s->vt, s->ld_vt,
                                             ps.pre_n, ps.n_share, layer_pre_qk(s, l) + s->q_dim, s->qk_dim,
                                             layer_pre_vt(s, l), s->ld_vtp, s->kv_dim, st), "unpack_prefix_kv");
        }
      } else if (!publish_in_epilogue(s, ps)) {
        rc = layer_publish_local(s, l, ps, P, st);
      }
      if (rc != 0) return rc;
    }
    rc = layer_back(s, l, ps, st);
    if (rc != 0) return rc;
  }
  return 0;
}

}  // namespace

extern "C" {

int bma_version(void) { return 200; }
void bma_set_last_error(const char* msg) { g_last_error = msg != nullptr ? msg : ""; }
void bma_count_launches(unsigned long long n) { g_launches += n; }
const char* bma_last_error(void) { return g_last_error.c_str(); }
unsigned long long bma_launch_count(void) { return g_launches; }

int bma_watchdog_code(unsigned int* code_out) {
  // host-mapped word: stays readable after a device trap has killed the context
  unsigned int* w = watchdog_host_word();
  if (code_out) *code_out = (w != nullptr) ? *w : 0u;
  return 0;
}

int bma_pgd_step(float* image, const float* grad, const float* orig, long long n, float step,
                 float eps, float* norm_out, const float* mean_host, const float* std_host,
                 int channels, long long plane, void* stream) {
  BMA_LAUNCH(launch_pgd_step(image, grad, orig, norm_out, n, step, eps, mean_host, std_host,
                             channels, plane, as_stream(stream)), "pgd_step");
  return 0;
}

int bma_token_grad(const void* grad_embeds, long long ld_g, const void* W, long long ld_w,
                   void* out, long long ld_out, int n_tok, int d, int V, void* stream) {
  if (n_tok <= 0 || d <= 0 || V <= 0) return fail(cudaErrorInvalidValue, "token_grad: empty shape");
  // W is the M operand (streamed once); the few gradient rows are the N operand. Suffixes longer
  // than one N tile are processed in slabs of 128 positions.
  for (int t0 = 0; t0 < n_tok; t0 += 128) {
    const int nt = n_tok - t0 < 128 ? n_tok - t0 : 128;
    GemmParams p{};
    p.M = V; p.N = nt; p.K = d; p.group_m = 1 << 20;
    p.k_stagger = token_grad_stagger();
    p.C = bf(out) + (long long)t0 * ld_out; p.ldc = ld_out;
    const __nv_bfloat16* g = bf(grad_embeds) + (long long)t0 * ld_g;
    if (nt <= 32) {
      BMA_LAUNCH((launch_gemm<32, EPI_STORE_T>(bf(W), ld_w, g, ld_g, p, as_stream(stream))),
                 "gemm(token_grad)");
    } else {
      BMA_LAUNCH((launch_gemm<128, EPI_STORE_T>(bf(W), ld_w, g, ld_g, p, as_stream(stream))),
                 "gemm(token_grad)");
    }
  }
  return 0;
}

size_t bma_sampler_workspace_bytes(int n_tok, int V, int topk) {
  return sampler_workspace_bytes(n_tok, V, topk);
}

int bma_sample_ids(const long long* ids, const void* grad, int grad_dtype, long long row_stride,
                   int n_tok, int V, const uint32_t* mask_bits, int search_width, int topk,
                   int n_replace, unsigned long long seed, unsigned long long offset,
                   void* workspace, long long* topk_ids_out, long long* new_ids,
                   unsigned long long* offset_advance, void* stream) {
  cudaError_t e = launch_sample_ids(ids, grad, grad_dtype, row_stride, n_tok, V, mask_bits,
                                    search_width, topk, n_replace, seed, offset, workspace,
                                    topk_ids_out, new_ids, offset_advance, as_stream(stream));
  g_launches += 3;
  if (e != cudaSuccess) return fail(e, "sample_ids");
  return 0;
}

bma_scorer* bma_scorer_create(const bma_model_desc* desc) {
  if (desc == nullptr) return nullptr;
  const bma_model_desc& D = *desc;
  if (D.head_dim != 64 && D.head_dim != 128 && D.head_dim != 256) {
    g_last_error = "scorer: head_dim must be 64, 128 or 256";
    return nullptr;
  }
  if (D.n_layers <= 0 || D.d_model % 64 != 0 || D.d_ff % 128 != 0 || D.n_heads % D.n_kv_heads != 0 ||
      D.max_rows <= 0 || D.max_prefix < 0 || D.vocab % 8 != 0) {
    g_last_error = "scorer: unsupported model shape";
    return nullptr;
  }
  bma_scorer* s = new bma_scorer();
  s->desc = D;
  s->fused_publish = fused_publish_default();
  {
    const char* e = getenv("BMA_EARLY_SIGNAL");       // off by default: no gain measured at N = 2 (bench.py: early_signal_ab)
    s->early_signal = (e != nullptr && atoi(e) != 0) ? 1 : 0;
  }
  s->layers.resize(D.n_layers);
  s->q_dim = D.n_heads * D.head_dim;
  s->kv_dim = D.n_kv_heads * D.head_dim;
  s->qkv_dim = s->q_dim + 2 * s->kv_dim;
  s->qk_dim = s->q_dim + s->kv_dim;
  const size_t R = static_cast<size_t>(D.max_rows > D.max_prefix ? D.max_rows : D.max_prefix);
  const int n_ce_tiles = (D.vocab + kCeBlockN - 1) / kCeBlockN;
  bool ok = true;
  ok = ok && dev_alloc(&s->x, R * D.d_model, &s->ws_bytes);
  ok = ok && dev_alloc(&s->xn, R * D.d_model, &s->ws_bytes);
  s->ld_vt = static_cast<long long>((R + 7) / 8 * 8);
  s->ld_vtp = static_cast<long long>(((D.max_prefix > 0 ? D.max_prefix : 1) + 7) / 8 * 8);
  ok = ok && dev_alloc(&s->qk, R * s->qk_dim, &s->ws_bytes);
  ok = ok && dev_alloc(&s->vt, static_cast<size_t>(s->kv_dim) * s->ld_vt, &s->ws_bytes);
  ok = ok && dev_alloc(&s->attn, R * s->q_dim, &s->ws_bytes);
  ok = ok && dev_alloc(&s->act, R * D.d_ff, &s->ws_bytes);
  ok = ok && dev_alloc(&s->prefix_qk,
                       static_cast<size_t>(D.n_layers) * (D.max_prefix > 0 ? D.max_prefix : 1) * s->qk_dim,
                       &s->ws_bytes);
  ok = ok && dev_alloc(&s->prefix_vt, static_cast<size_t>(D.n_layers) * s->kv_dim * s->ld_vtp, &s->ws_bytes);
  (void)n_ce_tiles;   // the CE partial buffers are sized on first use (score knows T): ensure_ce_rows
  ok = ok && dev_alloc(&s->row_info, static_cast<size_t>(D.max_rows), &s->ws_bytes);
  ok = ok && dev_alloc(&s->row_meta, static_cast<size_t>(D.max_rows), &s->ws_bytes);
  ok = ok && dev_alloc(&s->cand_vstart, static_cast<size_t>(D.max_rows) + 1, &s->ws_bytes);
  ok = ok && dev_alloc(&s->cand_tmp, static_cast<size_t>(D.max_rows) + 1, &s->ws_bytes);
  ok = ok && dev_alloc(&s->total_rows_dev, 1, &s->ws_bytes);
  ok = ok && dev_alloc(&s->rows_accum_dev, 1, &s->ws_bytes);
  ok = ok && dev_alloc(&s->flags_dev, 16, &s->ws_bytes);   // [0, 8) flags, [8] accumulated exchange wait (ns)
  ok = ok && dev_alloc(&s->done_dev, 1, &s->ws_bytes);
  ok = ok && dev_alloc(&s->base_dev, 4096, &s->ws_bytes);
  ok = ok && cudaHostAlloc(reinterpret_cast<void**>(&s->total_rows_host), sizeof(int), cudaHostAllocDefault) == cudaSuccess;
  ok = ok && dev_alloc(&s->ids_dev, static_cast<size_t>(D.max_rows), &s->ws_bytes);
  ok = ok && dev_alloc(&s->loss_dev, static_cast<size_t>(D.max_rows), &s->ws_bytes);
  ok = ok && dev_alloc(&s->key_dev, 1, &s->ws_bytes);
  ok = ok && dev_alloc(&s->flag_dev, 1, &s->ws_bytes);
  if (!ok) {
    g_last_error = std::string("scorer: cudaMalloc failed: ") + cudaGetErrorString(cudaGetLastError());
    bma_scorer_destroy(s);
    return nullptr;
  }
  return s;
}

void bma_scorer_destroy(bma_scorer* s) {
  if (s == nullptr) return;
  bma_scorer_clear_peers(s);
  cudaFree(s->flags_dev); cudaFree(s->done_dev);
  cudaFree(s->x); cudaFree(s->xn); cudaFree(s->qk); cudaFree(s->vt); cudaFree(s->attn); cudaFree(s->act);
  cudaFree(s->prefix_qk); cudaFree(s->prefix_vt); cudaFree(s->ce_part); cudaFree(s->ce_arg); cudaFree(s->tgt_logit);
  cudaFree(s->xchg_send); cudaFree(s->xchg_recv); cudaFree(s->rows_accum_dev);
  cudaFree(s->x_c); cudaFree(s->attn_c);
  cudaFree(s->row_info); cudaFree(s->row_meta); cudaFree(s->cand_vstart); cudaFree(s->cand_tmp); cudaFree(s->total_rows_dev);
  cudaFree(s->base_dev);
  if (s->total_rows_host != nullptr) cudaFreeHost(s->total_rows_host);
  cudaFree(s->ids_dev); cudaFree(s->loss_dev); cudaFree(s->key_dev); cudaFree(s->flag_dev);
  for (cudaEvent_t e : s->ev_start) cudaEventDestroy(e);
  for (cudaEvent_t e : s->ev_stop) cudaEventDestroy(e);
  delete s;
}

size_t bma_scorer_workspace_bytes(const bma_scorer* s) { return s ? s->ws_bytes : 0; }

int bma_scorer_profile(bma_scorer* s, int enable, int max_events) {
  if (s == nullptr) return fail(cudaErrorInvalidValue, "profile: null scorer");
  s->profiling = enable != 0;
  s->ev_used = 0;
  if (enable && static_cast<int>(s->ev_start.size()) < max_events) {
    const size_t old = s->ev_start.size();
    s->ev_start.resize(max_events); s->ev_stop.resize(max_events);
    s->ev_class.resize(max_events); s->ev_flops.resize(max_events);
    for (size_t i = old; i < static_cast<size_t>(max_events); ++i) {
      BMA_CHECK(cudaEventCreate(&s->ev_start[i]), "profile: event create");
      BMA_CHECK(cudaEventCreate(&s->ev_stop[i]), "profile: event create");
    }
  }
  return 0;
}

int bma_scorer_profile_read(bma_scorer* s, double* ms, double* flops, long long* launches, int n_classes) {
  if (s == nullptr || n_classes < PC_COUNT) return fail(cudaErrorInvalidValue, "profile_read: bad arguments");
  for (int c = 0; c < n_classes; ++c) { ms[c] = 0; flops[c] = 0; launches[c] = 0; }
  for (size_t i = 0; i < s->ev_used; ++i) {
    BMA_CHECK(cudaEventSynchronize(s->ev_stop[i]), "profile_read: sync");
    float t = 0;
    BMA_CHECK(cudaEventElapsedTime(&t, s->ev_start[i], s->ev_stop[i]), "profile_read: elapsed");
    ms[s->ev_class[i]] += t;
    flops[s->ev_class[i]] += s->ev_flops[i];
    launches[s->ev_class[i]] += 1;
  }
  s->ev_used = 0;
  return 0;
}

int bma_scorer_set_layer(bma_scorer* s, int layer, const bma_layer_weights* w) {
  if (s == nullptr || w == nullptr || layer < 0 || layer >= s->desc.n_layers)
    return fail(cudaErrorInvalidValue, "set_layer: bad arguments");
  if (s->desc.qk_norm && (w->q_norm == nullptr || w->k_norm == nullptr))
    return fail(cudaErrorInvalidValue, "set_layer: qk_norm needs q_norm and k_norm weights");
  if (s->desc.sandwich_norm && (w->norm_post_attn == nullptr || w->norm_post_mlp == nullptr))
    return fail(cudaErrorInvalidValue, "set_layer: sandwich_norm needs the two post-norm weights");
  s->layers[layer] = *w;
  return 0;
}

int bma_scorer_set_head(bma_scorer* s, const void* norm_final, const void* lm_head,
                        const void* embed_table, long long ld_embed) {
  if (s == nullptr) return fail(cudaErrorInvalidValue, "set_head: null scorer");
  s->norm_final = bf(norm_final);
  s->lm_head = bf(lm_head);
  s->embed = bf(embed_table);
  s->ld_embed = ld_embed;
  return 0;
}

int bma_scorer_prefill(bma_scorer* s, const void* prefix_embeds, int P, void* stream) {
  if (s == nullptr) return fail(cudaErrorInvalidValue, "prefill: null scorer");
  if (P < 0 || P > s->desc.max_prefix) return fail(cudaErrorInvalidValue, "prefill: P exceeds max_prefix");
  s->pending_prefix = nullptr; s->pending_P = 0;
  s->prefix_len = P;
  if (P == 0) return 0;
  cudaStream_t st = as_stream(stream);
  BMA_CHECK(cudaMemcpyAsync(s->x, prefix_embeds, (size_t)P * s->desc.d_model * sizeof(__nv_bfloat16),
                            cudaMemcpyDeviceToDevice, st), "prefill: copy embeds");
  Pass ps;
  ps.rows = P; ps.seq = P; ps.pos_base = 0; ps.prefill = true;
  return run_layers(s, ps, st);
}

int bma_scorer_prefill_deferred(bma_scorer* s, const void* prefix_embeds, int P) {
  if (s == nullptr) return fail(cudaErrorInvalidValue, "prefill_deferred: null scorer");
  if (P < 0 || P > s->desc.max_prefix) return fail(cudaErrorInvalidValue, "prefill_deferred: P exceeds max_prefix");
  s->prefix_len = P;
  s->pending_prefix = P > 0 ? bf(prefix_embeds) : nullptr;
  s->pending_P = P > 0 ? P : 0;
  return 0;
}

int bma_scorer_reset_key(bma_scorer* s, unsigned long long* key, int* all_match, void* stream) {
  (void)s;
  BMA_LAUNCH(launch_init_key(key, all_match, as_stream(stream)), "init_key");
  return 0;
}

long long bma_scorer_last_rows(const bma_scorer* s) {
  if (s == nullptr) return 0;
  unsigned long long ragged = 0;     // synchronising read: the ragged chunks counted their rows on the device
  if (cudaMemcpy(&ragged, s->rows_accum_dev, sizeof(ragged), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return s->last_rows + static_cast<long long>(ragged);
}

int bma_scorer_set_early_signal(bma_scorer* s, int enable) {
  if (s == nullptr) return fail(cudaErrorInvalidValue, "set_early_signal: null scorer");
  s->early_signal = enable != 0;
  return 0;
}

int bma_scorer_set_fused_publish(bma_scorer* s, int enable) {
  if (s == nullptr) return fail(cudaErrorInvalidValue, "set_fused_publish: null scorer");
  s->fused_publish = enable != 0;
  return 0;
}

int bma_scorer_set_dedupe(bma_scorer* s, int enable) {
  if (s == nullptr) return fail(cudaErrorInvalidValue, "set_dedupe: null scorer");
  s->dedupe = enable != 0;
  return 0;
}

}  // extern "C"

namespace {

struct ScoreArgs {
  const long long* ids; int ids_stride; int B; const int* slot; const void* fixed; int S;
  const long long* target_ids; int T; int ce_row0; float* loss; unsigned long long* key; int* all_match;
  unsigned int idx_base; const long long* base_ids; int n_id_rows;
};

// Ragged-mode geometry of one score call on one scorer.
struct ScorePlan {
  int n_share = 0;
  bool ragged = false, fused = false;
  int chunk = 0;            // candidates per chunk
  int pre0 = 0, pre_n = 0, pre_chunk = 0;
};

int plan_score(bma_scorer* s, const ScoreArgs& a, ScorePlan& sp, int rank, int size) {
  const bma_model_desc& D = s->desc;
  // Suffix-head sharing: rows [0, n_share) of the tail can coincide with the base's rows; n_share is
  // one past the last id-dependent tail position (the host knows its slot table). Needs room for the
  // base (one extra sequence per chunk, n_share extra prefix keys); shared rows precede the scored rows.
  int n_share = 0;
  if (a.base_ids != nullptr) {
    n_share = a.n_id_rows < 0 ? 0 : a.n_id_rows;
    if (n_share > a.ce_row0) n_share = a.ce_row0;      // scored rows are always the candidate's own
    if (s->prefix_len + n_share > D.max_prefix || 2 * a.S > D.max_rows) n_share = 0;
  }
  sp.n_share = n_share;
  sp.ragged = n_share > 0;
  sp.fused = sp.ragged && s->pending_prefix != nullptr;
  if (sp.fused) {
    const int P = s->pending_P;
    if (size > 1) {
      sp.pre_chunk = ((P + size - 1) / size + 7) & ~7;    // 8-aligned key offsets: 16-byte stores into the V^T rows
      sp.pre0 = rank * sp.pre_chunk < P ? rank * sp.pre_chunk : P;
      sp.pre_n = P - sp.pre0 < sp.pre_chunk ? P - sp.pre0 : sp.pre_chunk;
    } else {
      sp.pre0 = 0; sp.pre_n = P;
    }
    if (sp.pre_n + 2 * a.S > D.max_rows) { sp.fused = false; sp.pre0 = sp.pre_n = sp.pre_chunk = 0; }
  }
  sp.chunk = sp.ragged ? (D.max_rows - sp.pre_n) / a.S - 1 : D.max_rows / a.S;
  if (sp.ragged && sp.chunk > 8190) sp.chunk = 8190;              // plan_rows handles 8 entries per thread
  if (sp.chunk < 1) return fail(cudaErrorInvalidValue, "score: workspace too small for one candidate");
  return 0;
}

int ensure_xchg(bma_scorer* s, int pre_chunk, int size) {
  const size_t need = 2 * (size_t)pre_chunk * s->kv_dim * size;     // receive side: one block per rank
  if (need <= s->xchg_elems) return 0;
  BMA_CHECK(cudaDeviceSynchronize(), "ensure_xchg: sync");
  cudaFree(s->xchg_send); cudaFree(s->xchg_recv);
  s->xchg_send = s->xchg_recv = nullptr; s->xchg_elems = 0;
  size_t bytes = 0;
  if (!dev_alloc(&s->xchg_send, need / size, &bytes) || !dev_alloc(&s->xchg_recv, need, &bytes))
    return fail(cudaGetLastError(), "ensure_xchg: cudaMalloc");
  s->ws_bytes += bytes;
  s->xchg_elems = need;
  return 0;
}

// Rows + embeddings of chunk [b0, b0 + nb) (nb may be 0 on a rank that only contributes prefix rows).
int chunk_begin(bma_scorer* s, const ScoreArgs& a, const ScorePlan& sp, int b0, int nb, bool first, Pass& ps,
                ScoredRows& scored, cudaStream_t st) {
  const bma_model_desc& D = s->desc;
  const int d = D.d_model;
  const long long* ids_c = a.ids + (long long)b0 * a.ids_stride;
  const bool fused = sp.fused && first;
  ps = Pass{};
  ps.seq = a.S; ps.pos_base = s->prefix_len;
  ps.rows = nb * a.S;
  if (sp.ragged) {
    ps.n_share = sp.n_share;
    ps.publish = true;          // every chunk carries the base as its first sequence
    if (fused) { ps.pre0 = sp.pre0; ps.pre_n = sp.pre_n; ps.pre_chunk = sp.pre_chunk; }
    ProfScope pr(s, PC_EMBED, 0, st);
    BMA_LAUNCH(launch_plan_rows(ids_c, a.ids_stride, a.base_ids, a.slot, a.S, sp.n_share, nb, 1, s->prefix_len,
                                s->dedupe, ps.pre0, ps.pre_n, s->row_info, s->row_meta, s->cand_vstart,
                                s->cand_tmp, s->total_rows_dev, s->rows_accum_dev, st), "plan_rows");
    ps.rows = ps.pre_n + (nb + 1) * a.S;               // upper bound; the live count stays on the device
    ps.rows_dev = s->total_rows_dev;
    if (s->profiling) {
      // the per-class profile wants exact FLOPs: fetch the row count (profiling passes only)
      BMA_CHECK(cudaMemcpyAsync(s->total_rows_host, s->total_rows_dev, sizeof(int), cudaMemcpyDeviceToHost, st),
                "score: row count D2H");
      BMA_CHECK(cudaStreamSynchronize(st), "score: sync");
      ps.rows = *s->total_rows_host;
      if (ps.rows <= 0 || ps.rows > D.max_rows) return fail(cudaErrorInvalidValue, "score: bad ragged row count");
    }
  } else {
    s->last_rows += ps.rows;
  }
  {
    ProfScope pr(s, PC_EMBED, 0, st);
    BMA_LAUNCH(launch_build_tail_embeds(ids_c, a.ids_stride, a.slot, bf(a.fixed), s->embed, s->ld_embed, D.vocab,
                                        D.embed_scale, s->x, ps.rows, a.S, d, sp.ragged ? s->row_info : nullptr,
                                        a.base_ids, fused ? s->pending_prefix : nullptr, s->prefix_len,
                                        ps.rows_dev, st), "build_tail_embeds");
  }
  scored = ScoredRows{nb * a.T, a.T, a.ce_row0, sp.ragged ? s->cand_vstart : nullptr};
  ps.scored = nb > 0 ? &scored : nullptr;
  return 0;
}

// final norm + LM head fused with CE + per-candidate finalise of chunk [b0, b0 + nb)
int chunk_end(bma_scorer* s, const ScoreArgs& a, int b0, int nb, cudaStream_t st) {
  if (nb <= 0) return 0;
  const bma_model_desc& D = s->desc;
  const int d = D.d_model;
  const int n_ce_tiles = (D.vocab + kCeBlockN - 1) / kCeBlockN;
  const int ce_rows = nb * a.T;
  {                                            // final norm: the scored rows are contiguous in x_c
    ProfScope pr(s, PC_NORM, 0, st);
    BMA_LAUNCH(launch_rmsnorm(s->x_c, d, s->norm_final, s->xn, d, ce_rows, d, D.rms_eps, D.norm_mode,
                              0, 0, 0, nullptr, st), "rmsnorm(final)");
  }
  {
    GemmParams p{};
    p.M = ce_rows; p.N = D.vocab; p.K = d; p.group_m = group_m_for(p.K);
    p.target_ids = a.target_ids; p.T = a.T;
    p.ce_part = s->ce_part; p.tgt_logit = s->tgt_logit;
    p.ce_arg = a.all_match != nullptr ? s->ce_arg : nullptr;
    ProfScope pr(s, PC_GEMM_CE, gemm_flops(p), st);
    BMA_LAUNCH((launch_gemm_big<EPI_CE>(s->xn, d, s->lm_head, d, p, st)), "gemm(lm_head+ce)");
  }
  {
    ProfScope pr(s, PC_FINAL, 0, st);
    BMA_LAUNCH(launch_ce_finalize(s->ce_part, s->tgt_logit, a.all_match != nullptr ? s->ce_arg : nullptr,
                                  a.target_ids, n_ce_tiles, nb, a.T, a.loss + b0, a.key,
                                  a.all_match, a.idx_base + static_cast<unsigned int>(b0), st),
               "ce_finalize");
  }
  return 0;
}

int check_score_args(bma_scorer* s, const ScoreArgs& a) {
  if (s == nullptr) return fail(cudaErrorInvalidValue, "score: null scorer");
  const bma_model_desc& D = s->desc;
  if (a.S <= 0 || a.T <= 0 || a.ce_row0 < 0 || a.ce_row0 + a.T > a.S)
    return fail(cudaErrorInvalidValue, "score: target rows [ce_row0, ce_row0+T) must lie inside the tail");
  if (a.S > D.max_rows) return fail(cudaErrorInvalidValue, "score: S exceeds max_rows");
  if (s->lm_head == nullptr || s->norm_final == nullptr || s->embed == nullptr)
    return fail(cudaErrorInvalidValue, "score: head weights not set");
  return 0;
}

// The score call of ONE scorer (rank `rank` of `size` when a communicator is attached).
int score_one(bma_scorer* s, const ScoreArgs& a, cudaStream_t st, bool reset_counters = true) {
  int rc = check_score_args(s, a);
  if (rc != 0) return rc;
  const int size = s->comm != nullptr ? comm_size(s->comm) : 1;
  const int rank = s->comm != nullptr ? comm_rank(s->comm) : 0;
  if (s->peer_n > 1 && (s->peer_n != size || s->peer_rank != rank))
    return fail(cudaErrorInvalidValue, "score: peer set does not match the communicator");
  if (a.B <= 0 && !(size > 1 && s->pending_prefix != nullptr && a.base_ids != nullptr)) return 0;
  if (reset_counters) {
    s->last_rows = 0;
    BMA_CHECK(cudaMemsetAsync(s->rows_accum_dev, 0, sizeof(unsigned long long), st), "score: reset row counter");
  }
  ScorePlan sp;
  rc = plan_score(s, a, sp, rank, size);
  if (rc != 0) return rc;
  if (s->pending_prefix != nullptr && !sp.fused) {      // no ragged pass to ride in: prefill now
    rc = bma_scorer_prefill(s, s->pending_prefix, s->pending_P, st);
    if (rc != 0) return rc;
  }
  if (sp.fused && sp.pre_chunk > 0 && s->peer_n > 1) {
    rc = peers_barrier(s, st);       // every rank is done reading its prefix buffers from earlier passes
    if (rc != 0) return rc;
  } else if (sp.fused && sp.pre_chunk > 0) {
    rc = ensure_xchg(s, sp.pre_chunk, size);
    if (rc != 0) return rc;
  }
  {
    const int most = a.B < sp.chunk ? a.B : sp.chunk;
    rc = ensure_ce_rows(s, (most > 0 ? most : 1) * a.T);
    if (rc != 0) return rc;
  }
  bool first = true;
  for (int b0 = 0; b0 < a.B || first; b0 += sp.chunk) {
    const int nb = a.B - b0 < sp.chunk ? a.B - b0 : sp.chunk;
    Pass ps; ScoredRows scored;
    rc = chunk_begin(s, a, sp, b0, nb, first, ps, scored, st);
    if (rc != 0) return rc;
    rc = run_layers(s, ps, st);
    if (rc != 0) return rc;
    rc = chunk_end(s, a, b0, nb, st);
    if (rc != 0) return rc;
    first = false;
  }
  if (sp.fused) { s->pending_prefix = nullptr; s->pending_P = 0; }
  return 0;
}

}  // namespace

extern "C" {

int bma_scorer_score(bma_scorer* s, const long long* ids, int ids_stride, int B, const int* slot,
                     const void* fixed, int S, const long long* target_ids, int T, int ce_row0,
                     float* loss, unsigned long long* key, int* all_match, unsigned int idx_base,
                     const long long* base_ids, int n_id_rows, void* stream) {
  const ScoreArgs a{ids, ids_stride, B, slot, fixed, S, target_ids, T, ce_row0, loss, key, all_match,
                    idx_base, base_ids, n_id_rows};
  return score_one(s, a, as_stream(stream));
}

// Test / single-process form of the multi-rank path: `n` scorers (same model, same deferred prefix)
// act as the ranks of one box on ONE device. Rank r scores the contiguous shard r of the B candidates
// (first ranks take the extras, like bimodalattack_b200.dist.shard_bounds) and computes the r-th slice
// of the prefix positions; the per-layer K/V exchange is a set of local copies instead of an NCCL
// all-gather. Everything else is the code path of a real rank.
int bma_scorer_group_score(bma_scorer** group, int n, const long long* ids, int ids_stride, int B,
                           const int* slot, const void* fixed, int S, const long long* target_ids, int T,
                           int ce_row0, float* loss, unsigned long long* key, int* all_match,
                           const long long* base_ids, int n_id_rows, int use_peers, void* stream) {
  if (group == nullptr || n <= 0) return fail(cudaErrorInvalidValue, "group_score: empty group");
  cudaStream_t st = as_stream(stream);
  std::vector<ScoreArgs> args(n);
  std::vector<ScorePlan> plans(n);
  std::vector<Pass> pass(n);
  std::vector<ScoredRows> scored(n);
  std::vector<int> nb0(n);
  const int base = B / n, extra = B % n;
  bool fused = true;
  for (int r = 0; r < n; ++r) {
    bma_scorer* s = group[r];
    const int lo = r * base + (r < extra ? r : extra), cnt = base + (r < extra ? 1 : 0);
    args[r] = ScoreArgs{ids + (long long)lo * ids_stride, ids_stride, cnt, slot, fixed, S, target_ids, T, ce_row0,
                        loss + lo, key, all_match, static_cast<unsigned int>(lo), base_ids, n_id_rows};
    int rc = check_score_args(s, args[r]);
    if (rc != 0) return rc;
    s->last_rows = 0;
    BMA_CHECK(cudaMemsetAsync(s->rows_accum_dev, 0, sizeof(unsigned long long), st), "group_score: reset");
    rc = plan_score(s, args[r], plans[r], r, n);
    if (rc != 0) return rc;
    fused = fused && plans[r].fused && plans[r].pre_chunk == plans[0].pre_chunk;
  }
  if (!fused || n == 1) {                     // nothing to exchange: every scorer runs on its own
    for (int r = 0; r < n; ++r) {
      if (n > 1 && group[r]->pending_prefix != nullptr) {
        int rc = bma_scorer_prefill(group[r], group[r]->pending_prefix, group[r]->pending_P, stream);
        if (rc != 0) return rc;
      }
      int rc = score_one(group[r], args[r], st);
      if (rc != 0) return rc;
    }
    return 0;
  }
  const int P = group[0]->pending_P;
  const bool via_peers = n <= 8 && use_peers;
  for (int r = 0; r < n; ++r) {
    bma_scorer* s = group[r];
    int rc = 0;
    if (via_peers) {                           // the scorers of the group are each other's "peers": plain pointers
      bma_scorer_clear_peers(s);
      BMA_CHECK(cudaMemsetAsync(s->flags_dev, 0, 8 * sizeof(unsigned long long), st), "group_score: reset flags");
      BMA_CHECK(cudaMemsetAsync(s->done_dev, 0, sizeof(unsigned int), st), "group_score: reset counter");
      s->epoch = 0;
      s->peer_n = n; s->peer_rank = r; s->peer_ipc = false;
      for (int q = 0; q < n; ++q) {
        s->peer_qk[q] = group[q]->prefix_qk; s->peer_vt[q] = group[q]->prefix_vt; s->peer_flags[q] = group[q]->flags_dev;
      }
    } else {
      rc = ensure_xchg(s, plans[r].pre_chunk, n);
      if (rc != 0) return rc;
    }
    const int most = args[r].B < plans[r].chunk ? args[r].B : plans[r].chunk;
    rc = ensure_ce_rows(s, (most > 0 ? most : 1) * T);
    if (rc != 0) return rc;
    nb0[r] = most > 0 ? most : 0;
    rc = chunk_begin(s, args[r], plans[r], 0, nb0[r], true, pass[r], scored[r], st);
    if (rc != 0) return rc;
    pass_start(s, pass[r]);
  }
  const int L = group[0]->desc.n_layers;
  for (int l = 0; l < L; ++l) {
    for (int r = 0; r < n; ++r) {
      int rc = layer_front(group[r], l, pass[r], st);
      if (rc != 0) return rc;
      if (!via_peers) rc = layer_pack(group[r], l, pass[r], st);
      else if (early_signal(group[r], pass[r])) rc = 0;       // the GEMM raised the flags
      else if (publish_in_epilogue(group[r], pass[r])) rc = layer_signal_peers(group[r], false, st);
      else rc = layer_scatter_peers(group[r], l, pass[r], st);
      if (rc != 0) return rc;
    }
    for (int r = 0; r < n; ++r) {
      bma_scorer* s = group[r];
      int rc = 0;
      if (via_peers) {
        rc = layer_wait_peers(s, l, pass[r], P, st);
        if (rc != 0) return rc;
      } else {
        const size_t blk = 2 * (size_t)pass[r].pre_chunk * s->kv_dim;
        for (int q = 0; q < n; ++q)             // the all-gather, as local copies into this rank's receive buffer
          BMA_CHECK(cudaMemcpyAsync(s->xchg_recv + q * blk, group[q]->xchg_send, blk * sizeof(__nv_bfloat16),
                                    cudaMemcpyDeviceToDevice, st), "group_score: gather copy");
        BMA_LAUNCH(launch_unpack_prefix_kv(s->xchg_recv, static_cast<long long>(blk), pass[r].pre_chunk, n, P,
                                           s->qk + s->q_dim, s->qk_dim, s->vt, s->ld_vt, pass[r].pre_n,
                                           pass[r].n_share, layer_pre_qk(s, l) + s->q_dim, s->qk_dim,
                                           layer_pre_vt(s, l), s->ld_vtp, s->kv_dim, st), "unpack_prefix_kv");
      }
      rc = layer_back(group[r], l, pass[r], st);
      if (rc != 0) return rc;
    }
  }
  if (via_peers)
    for (int r = 0; r < n; ++r) bma_scorer_clear_peers(group[r]);
  for (int r = 0; r < n; ++r) {
    bma_scorer* s = group[r];
    int rc = chunk_end(s, args[r], 0, nb0[r], st);
    if (rc != 0) return rc;
    s->pending_prefix = nullptr; s->pending_P = 0;
    if (args[r].B > nb0[r]) {                 // remaining chunks of this rank need no exchange
      ScoreArgs rest = args[r];
      rest.ids += (long long)nb0[r] * ids_stride; rest.B -= nb0[r]; rest.loss += nb0[r];
      rest.idx_base += static_cast<unsigned int>(nb0[r]);
      rc = score_one(s, rest, st, false);
      if (rc != 0) return rc;
    }
  }
  return 0;
}

int bma_scorer_score_host(bma_scorer* s, const long long* ids_host, int ids_stride, int B,
                          const int* slot, const void* fixed, int S, const long long* target_ids,
                          int T, int ce_row0, float* loss_host, unsigned long long* key_host,
                          unsigned int idx_base, const long long* base_ids_host, int n_id_rows, void* stream) {
  if (s == nullptr) return fail(cudaErrorInvalidValue, "score_host: null scorer");
  if ((long long)B * ids_stride > s->desc.max_rows || ids_stride > 4096)
    return fail(cudaErrorInvalidValue, "score_host: ids exceed staging capacity");
  cudaStream_t st = as_stream(stream);
  BMA_CHECK(cudaMemcpyAsync(s->ids_dev, ids_host, sizeof(long long) * B * ids_stride,
                            cudaMemcpyHostToDevice, st), "score_host: H2D ids");
  if (base_ids_host != nullptr)
    BMA_CHECK(cudaMemcpyAsync(s->base_dev, base_ids_host, sizeof(long long) * ids_stride,
                              cudaMemcpyHostToDevice, st), "score_host: H2D base ids");
  int rc = bma_scorer_reset_key(s, s->key_dev, s->flag_dev, stream);
  if (rc != 0) return rc;
  rc = bma_scorer_score(s, s->ids_dev, ids_stride, B, slot, fixed, S, target_ids, T, ce_row0,
                        s->loss_dev, s->key_dev, nullptr, idx_base,
                        base_ids_host != nullptr ? s->base_dev : nullptr, n_id_rows, stream);
  if (rc != 0) return rc;
  if (s->comm != nullptr) {                    // multi-rank: the winner key is reduced before it leaves the device
    rc = bma_comm_allreduce_min_packed(s->comm, s->key_dev, 1, stream);
    if (rc != 0) return rc;
  }
  BMA_CHECK(cudaMemcpyAsync(loss_host, s->loss_dev, sizeof(float) * B, cudaMemcpyDeviceToHost, st),
            "score_host: D2H loss");
  BMA_CHECK(cudaMemcpyAsync(key_host, s->key_dev, sizeof(unsigned long long),
                            cudaMemcpyDeviceToHost, st), "score_host: D2H key");
  BMA_CHECK(cudaStreamSynchronize(st), "score_host: sync");
  return 0;
}

// ---- peer memory (NVLink) ------------------------------------------------------------------------
// `count` data-free exchanges back to back on `stream` (collective over the ranks): the latency floor of one
// per-layer exchange, measured by the caller with events around the call.
int bma_scorer_peers_barrier(bma_scorer* s, int count, void* stream) {
  if (s == nullptr || count < 0) return fail(cudaErrorInvalidValue, "peers_barrier: bad arguments");
  if (s->peer_n <= 1) return 0;
  for (int i = 0; i < count; ++i) {
    const int rc = peers_barrier(s, static_cast<cudaStream_t>(stream));
    if (rc != 0) return rc;
  }
  return 0;
}

// Nanoseconds this scorer's stream has spent in the per-layer prefix exchange waiting for other ranks' flags
// since the last reset (synchronises the device: call it between steps).
int bma_scorer_exchange_wait_ns(bma_scorer* s, unsigned long long* out, int reset) {
  if (s == nullptr || out == nullptr) return fail(cudaErrorInvalidValue, "exchange_wait_ns: null argument");
  BMA_CHECK(cudaMemcpy(out, s->flags_dev + 8, sizeof(unsigned long long), cudaMemcpyDeviceToHost), "exchange_wait_ns: read");
  if (reset) BMA_CHECK(cudaMemset(s->flags_dev + 8, 0, sizeof(unsigned long long)), "exchange_wait_ns: reset");
  return 0;
}

int bma_scorer_clear_peers(bma_scorer* s) {
  if (s == nullptr) return 0;
  if (s->peer_ipc) {
    for (int q = 0; q < s->peer_n; ++q) {
      if (q == s->peer_rank) continue;
      if (s->peer_qk[q] != nullptr) cudaIpcCloseMemHandle(s->peer_qk[q]);
      if (s->peer_vt[q] != nullptr) cudaIpcCloseMemHandle(s->peer_vt[q]);
      if (s->peer_flags[q] != nullptr) cudaIpcCloseMemHandle(s->peer_flags[q]);
    }
  }
  for (int q = 0; q < 8; ++q) { s->peer_qk[q] = s->peer_vt[q] = s->peer_flags[q] = nullptr; }
  s->peer_n = 0; s->peer_ipc = false;
  return 0;
}

int bma_scorer_export_ipc(bma_scorer* s, void* out192) {
  if (s == nullptr || out192 == nullptr) return fail(cudaErrorInvalidValue, "export_ipc: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  cudaIpcMemHandle_t h[3];
  // flags and epoch restart here, BEFORE the handles travel: no peer can signal into this scorer earlier
  bma_scorer_clear_peers(s);
  BMA_CHECK(cudaMemset(s->flags_dev, 0, 8 * sizeof(unsigned long long)), "export_ipc: reset flags");
  BMA_CHECK(cudaMemset(s->done_dev, 0, sizeof(unsigned int)), "export_ipc: reset counter");
  s->epoch = 0;
  BMA_CHECK(cudaIpcGetMemHandle(&h[0], s->prefix_qk), "export_ipc: prefix K");
  BMA_CHECK(cudaIpcGetMemHandle(&h[1], s->prefix_vt), "export_ipc: prefix V^T");
  BMA_CHECK(cudaIpcGetMemHandle(&h[2], s->flags_dev), "export_ipc: flags");
  memcpy(out192, h, sizeof(h));
  return 0;
}

int bma_scorer_import_ipc(bma_scorer* s, const void* handles, int size, int rank) {
  if (s == nullptr || handles == nullptr || size < 2 || size > 8 || rank < 0 || rank >= size)
    return fail(cudaErrorInvalidValue, "import_ipc: bad arguments (2..8 ranks)");
  const cudaIpcMemHandle_t* h = static_cast<const cudaIpcMemHandle_t*>(handles);
  s->peer_n = size; s->peer_rank = rank; s->peer_ipc = true;
  for (int q = 0; q < size; ++q) {
    if (q == rank) {
      s->peer_qk[q] = s->prefix_qk; s->peer_vt[q] = s->prefix_vt; s->peer_flags[q] = s->flags_dev;
      continue;
    }
    BMA_CHECK(cudaIpcOpenMemHandle(&s->peer_qk[q], h[3 * q + 0], cudaIpcMemLazyEnablePeerAccess), "import_ipc: open prefix K");
    BMA_CHECK(cudaIpcOpenMemHandle(&s->peer_vt[q], h[3 * q + 1], cudaIpcMemLazyEnablePeerAccess), "import_ipc: open prefix V^T");
    BMA_CHECK(cudaIpcOpenMemHandle(&s->peer_flags[q], h[3 * q + 2], cudaIpcMemLazyEnablePeerAccess), "import_ipc: open flags");
  }
  return 0;
}

int bma_scorer_set_comm(bma_scorer* s, bma_comm* comm) {
  if (s == nullptr) return fail(cudaErrorInvalidValue, "set_comm: null scorer");
  s->comm = comm;
  return 0;
}

int bma_gemm_bf16(const void* A, long long lda, const void* B, long long ldb, void* C,
                  long long ldc, const void* residual, long long ldr, int M, int N, int K,
                  int block_n, void* stream) {
  GemmParams p{};
  p.M = M; p.N = N; p.K = K; p.group_m = 16;
  p.C = bf(C); p.ldc = ldc; p.R = bf(residual); p.ldr = ldr;
  if (block_n == 256) {
    BMA_LAUNCH((launch_gemm_big<EPI_STORE>(bf(A), lda, bf(B), ldb, p, as_stream(stream))), "gemm");
  } else if (block_n == 128) {
    BMA_LAUNCH((launch_gemm<128, EPI_STORE>(bf(A), lda, bf(B), ldb, p, as_stream(stream))), "gemm");
  } else {
    return fail(cudaErrorInvalidValue, "gemm: block_n must be 128 or 256");
  }
  return 0;
}

int bma_rmsnorm(const void* x, long long ldx, const void* w, void* y, long long ldy, int rows,
                int d, float eps, int mode, void* stream) {
  BMA_LAUNCH(launch_rmsnorm(bf(x), ldx, bf(w), bf(y), ldy, rows, d, eps, mode, 0, 0, 0, nullptr,
                            as_stream(stream)), "rmsnorm");
  return 0;
}

int bma_add_norm(const void* y, long long ldy, const void* w_post, void* x, long long ldx,
                 const void* w_next, void* xn, long long ldxn, int rows, int d, float eps, int mode,
                 void* stream) {
  BMA_LAUNCH(launch_add_norm(bf(y), ldy, bf(w_post), bf(x), ldx, bf(w_next), bf(xn), ldxn, rows, d, eps,
                             mode, as_stream(stream)), "add_norm");
  return 0;
}

int bma_attention(const void* q, long long ldq, const void* k_tail, long long ldk,
                  const void* vt_tail, long long ldvt, const void* k_pre, long long ldkp,
                  const void* vt_pre, long long ldvtp, void* out, long long ldo, int M, int S, int P,
                  int n_heads, int kv_group, int head_dim, float scale, void* stream) {
  AttnParams a{};
  a.head_dim = head_dim;
  a.q = bf(q); a.ldq = ldq;
  a.k_tail = bf(k_tail); a.ldk = ldk;
  a.vt_tail = bf(vt_tail); a.ldvt = ldvt;
  a.k_pre = bf(k_pre); a.ldkp = ldkp;
  a.vt_pre = bf(vt_pre); a.ldvtp = ldvtp;
  a.out = bf(out); a.ldo = ldo;
  a.M = M; a.S = S; a.P = P; a.kv_group = kv_group;
  a.no_tail = S == 0 ? 1 : 0;          // S = 0: the rows attend to the P prefix keys only
  a.scale_log2 = scale * 1.4426950408889634f;
  BMA_LAUNCH(launch_prefix_tail_attention(a, n_heads, as_stream(stream)), "attention");
  return 0;
}

}  // extern "C"
