/* bimodal_b200 -- C ABI of the B200-native GCG+PGD step engine.
 *
 * Drop-in boundary for the per-step hot path of Bimo99B9/BimodalAttack (pure Python; file:line
 * below refer to the reference's bimodalattack/bimodal_attack.py). Every entry point takes plain
 * pointers and sizes: device pointers unless the name ends in `_host`, and a CUDA stream passed as
 * `void*` (cudaStream_t; NULL = legacy default stream). All functions return 0 on success or a
 * non-zero cudaError_t value; bma_last_error() gives a text for the most recent failure on the
 * calling thread. Nothing here falls back to the CPU.
 *
 * The Python host (bimodalattack_b200/native.py) binds these with ctypes; INTEGRATION.md shows the
 * stub a maintainer of the reference would add.
 */
#ifndef BIMODAL_B200_H
#define BIMODAL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BMA_DTYPE_BF16 0
#define BMA_DTYPE_F16 1
#define BMA_DTYPE_F32 2

/* ---- library ---------------------------------------------------------------------------- */
int bma_version(void);
const char* bma_last_error(void);
/* Code recorded by the in-kernel pipeline watchdog (0 = none). */
int bma_watchdog_code(unsigned int* code_out);

/* ---- (4) PGD image step ------------------------------------------------------------------
 * Replaces BimodalAttack.perform_pgd_step (:1030-1037):
 *   image <- clamp(clamp(image - step*sign(grad), orig-eps, orig+eps), 0, 1),  step = alpha*eps.
 * fp32, in place, bit-exact with the torch sequence. If norm_out != NULL also writes
 * (image - mean[c]) / std[c] (torchvision Normalize, reference :526/:703/:971); mean/std are HOST
 * arrays of `channels` floats and `plane` = H*W. */
int bma_pgd_step(float* image, const float* grad, const float* orig, long long n, float step,
                 float eps, float* norm_out, const float* mean_host, const float* std_host,
                 int channels, long long plane, void* stream);

/* ---- (1) GCG token gradient --------------------------------------------------------------
 * Replaces the backward of `optim_ids_onehot @ embedding_layer.weight` (:968, :1016-1025):
 *   out[n_tok, V] = grad_embeds[n_tok, d] * W[V, d]^T     (bf16 in, fp32 accumulate, bf16 out)
 * W is streamed exactly once through TMA as the tcgen05 A operand. */
int bma_token_grad(const void* grad_embeds, long long ld_g, const void* W, long long ld_w,
                   void* out, long long ld_out, int n_tok, int d, int V, void* stream);

/* ---- (2) candidate sampling --------------------------------------------------------------
 * Replaces sample_ids_from_grad (:130-163). grad is [n_tok, V] of `grad_dtype`; mask_bits is a
 * V-bit set (uint32 words, bit i of word w <-> token 32*w+i, 1 = not allowed) or NULL; (seed,
 * offset) is torch's CUDA Philox state; *offset_advance (host) receives how far torch's generator
 * must be advanced afterwards. topk_ids_out ([n_tok, topk] int64) may be NULL. */
size_t bma_sampler_workspace_bytes(int n_tok, int V, int topk);
int bma_sample_ids(const long long* ids, const void* grad, int grad_dtype, long long row_stride,
                   int n_tok, int V, const uint32_t* mask_bits, int search_width, int topk,
                   int n_replace, unsigned long long seed, unsigned long long offset,
                   void* workspace, long long* topk_ids_out, long long* new_ids,
                   unsigned long long* offset_advance, void* stream);

/* ---- (3) candidate scoring ---------------------------------------------------------------
 * Replaces _build_input_embeds* (:1112-1276) + _compute_candidates_loss_original (:1278-1310) +
 * loss.min()/argmin() (:550-551, :584-585, :636-637) for a Llama/Gemma-style decoder. */
typedef struct {
  int n_layers, d_model, n_heads, n_kv_heads, head_dim, d_ff, vocab;
  float rms_eps;
  int norm_mode;      /* 0: Llama  w*bf16(x*rstd);  1: Gemma  x*rstd*(1+w) */
  int act;            /* 0: SiLU, 1: GELU(tanh) */
  float embed_scale;  /* multiplies embedding rows of sampled ids (Gemma: sqrt(d)), else 1 */
  float attn_scale;   /* softmax scale, e.g. head_dim^-0.5 */
  int max_rows;       /* capacity in tail rows (candidates per pass = max_rows / S) */
  int max_prefix;     /* capacity in prefix rows */
  int qk_norm;        /* 1: per-head RMSNorm (norm_mode) on Q and K before the rotation (Gemma-3) */
  int sandwich_norm;  /* 1: Gemma-style layer  x += norm_post_attn(attn(norm_attn(x)));
                            x += norm_post_mlp(mlp(norm_mlp(x)))   (4 norms per layer)         */
} bma_model_desc;

typedef struct {
  const void* w_qkv;      /* [(H+2*Hkv)*hd, d] rows: Q heads | K heads | V heads            */
  const void* w_o;        /* [d, H*hd]                                                      */
  const void* w_gate_up;  /* [2*d_ff, d]; per 256-row block: 128 gate rows then 128 up rows */
  const void* w_down;     /* [d, d_ff]                                                      */
  const void* norm_attn;  /* [d] */
  const void* norm_mlp;   /* [d] */
  const float* rope_cos;  /* [max_pos, hd/2] fp32, values already rounded to the model dtype;
                             per layer, so local/global layers can use different tables        */
  const float* rope_sin;
  const void* norm_post_attn;  /* [d]   sandwich_norm only */
  const void* norm_post_mlp;   /* [d]   sandwich_norm only */
  const void* q_norm;          /* [hd]  qk_norm only */
  const void* k_norm;          /* [hd]  qk_norm only */
} bma_layer_weights;

typedef struct bma_scorer bma_scorer;

bma_scorer* bma_scorer_create(const bma_model_desc* desc);
void bma_scorer_destroy(bma_scorer* s);
int bma_scorer_set_layer(bma_scorer* s, int layer, const bma_layer_weights* w);
int bma_scorer_set_head(bma_scorer* s, const void* norm_final, const void* lm_head,
                        const void* embed_table, long long ld_embed);
size_t bma_scorer_workspace_bytes(const bma_scorer* s);

/* Run the shared prefix (image tokens + prompt) once; its K/V stay resident in the scorer. */
int bma_scorer_prefill(bma_scorer* s, const void* prefix_embeds, int P, void* stream);
/* Fused prefill: only RECORD the prefix ([P, d] bf16, must stay valid until it has been consumed).
 * The next bma_scorer_score call that carries base_ids runs the P prefix positions as extra rows of
 * its first chunk -- one row space [prefix | base suffix | candidates] per layer, every weight matrix
 * streamed once instead of twice and half the launches -- and publishes their K/V to the per-layer
 * prefix buffers before that layer's attention. With a communicator attached (bma_scorer_set_comm)
 * the prefix positions are split over the ranks and each layer's K/V slices are all-gathered. A score
 * call without base_ids (or bma_scorer_prefill) consumes the recorded prefix with a plain prefill. */
int bma_scorer_prefill_deferred(bma_scorer* s, const void* prefix_embeds, int P);

/* Score B candidates. ids is [B, ids_stride] int64; tail position s takes the embedding of
 * ids[b, slot[s]] when slot[s] >= 0, else row s of fixed ([S, d] bf16). The loss of candidate b is
 * the mean CE of target_ids[0..T) against the logits of tail rows ce_row0 .. ce_row0+T-1 (the
 * reference's `logits[..., L-T-1:-1, :]`, :1289-1298). The row of the LAST target token predicts
 * nothing that is scored and no other row attends to it, so callers pass S = (tail length - 1)
 * and ce_row0 = S - T. Outputs: loss[B]
 * fp32; *key = min over candidates of (orderable(loss) << 31 | idx_base + b) merged into the
 * existing value (call bma_scorer_reset_key first); *all_match != 0 if some candidate's argmax
 * equals the target everywhere (early_stop, :1300-1306).
 *
 * base_ids (device, [ids_stride] int64, or NULL) is the suffix the candidates were sampled from
 * (the reference's optim_ids, :930-939). When given, tail rows a candidate shares with the base --
 * everything before its first replaced position -- are computed once instead of once per
 * candidate: the base is run as one extra sequence, its K/V are appended to the shared prefix and
 * each candidate computes only the rows from its first difference on. Any base gives the same
 * losses; one unrelated to the candidates merely saves nothing. n_id_rows is one past the last tail
 * position whose slot is >= 0 (the host knows its slot table); the ragged row count stays on the
 * device (every kernel of the pass reads it there), so the call never synchronises the stream. */
int bma_scorer_reset_key(bma_scorer* s, unsigned long long* key, int* all_match, void* stream);
int bma_scorer_score(bma_scorer* s, const long long* ids, int ids_stride, int B, const int* slot,
                     const void* fixed, int S, const long long* target_ids, int T, int ce_row0,
                     float* loss, unsigned long long* key, int* all_match, unsigned int idx_base,
                     const long long* base_ids, int n_id_rows, void* stream);
/* Tail rows the last bma_scorer_score call ran through the decoder (B*S without base_ids). Reads a
 * device counter: synchronises. */
long long bma_scorer_last_rows(const bma_scorer* s);
/* With base_ids: candidates identical to an earlier candidate of the chunk (or to the base) reuse its
 * rows instead of recomputing them (default on); their loss is the representative's. */
int bma_scorer_set_dedupe(bma_scorer* s, int enable);
/* With base_ids / a deferred prefix: the K/V of the prefix rows and of the base rows are stored into the
 * per-layer prefix buffers by the QKV projection's epilogue -- and, when the prefix is split over the
 * ranks of a box, into every peer's buffers over NVLink, so the per-layer all-gather is part of the
 * GEMM (default on; BMA_FUSED_PUBLISH=0 or enable = 0 selects the separate copy / scatter kernels). */
int bma_scorer_set_fused_publish(bma_scorer* s, int enable);

/* Same call with HOST buffers for ids / loss / key (pinned recommended): copies in, scores,
 * copies out and synchronises the stream. This is the end-to-end entry the benchmark times. */
int bma_scorer_score_host(bma_scorer* s, const long long* ids_host, int ids_stride, int B,
                          const int* slot, const void* fixed, int S, const long long* target_ids,
                          int T, int ce_row0, float* loss_host, unsigned long long* key_host,
                          unsigned int idx_base, const long long* base_ids_host, int n_id_rows,
                          void* stream);

/* ---- multi-GPU ------------------------------------------------------
 * One process per GPU. Candidates are sharded over the ranks (the host passes each rank its slice
 * and idx_base); the library owns an NCCL communicator for the two exchanges of the path: the
 * per-layer all-gather of the split prefix K/V (fused prefill) and ONE all-reduce(MIN) over the packed
 * winner key. NCCL is resolved at run time from the libnccl.so.2 already loaded in the process.
 * bma_comm_unique_id fills 128 bytes on rank 0; the host hands them to every rank (the reference
 * has no multi-GPU path; its winner selection is loss.min()/argmin(), :550-551). */
typedef struct bma_comm bma_comm;
int bma_comm_unique_id(void* out128);
bma_comm* bma_comm_create(const void* id128, int size, int rank);
void bma_comm_destroy(bma_comm* c);
int bma_comm_rank(const bma_comm* c);
int bma_comm_size(const bma_comm* c);
const char* bma_comm_last_error(void);
/* In-place MIN over n packed keys on `stream`. */
int bma_comm_allreduce_min_packed(bma_comm* c, unsigned long long* keys, int n, void* stream);
/* Winner key (MIN) and, when all_match != NULL, the early-stop flag (MAX) in one NCCL launch. */
int bma_comm_reduce_winner(bma_comm* c, unsigned long long* key, int* all_match, void* stream);
/* Attach (or detach with NULL) the communicator the scorer splits its fused prefill over. */
int bma_scorer_set_comm(bma_scorer* s, bma_comm* comm);
/* Single-process form of the multi-rank path for tests: the n scorers act as the ranks of one box on
 * one device (contiguous candidate shards, split prefix, the all-gather done as local copies). Each
 * scorer must hold the same deferred prefix. loss is [B]; key / all_match are merged like
 * bma_scorer_score does. */
int bma_scorer_group_score(bma_scorer** group, int n, const long long* ids, int ids_stride, int B,
                           const int* slot, const void* fixed, int S, const long long* target_ids, int T,
                           int ce_row0, float* loss, unsigned long long* key, int* all_match,
                           const long long* base_ids, int n_id_rows, int use_peers, void* stream);
/* Peer-memory form of the split prefill (NVLink 5 / NVSwitch): every rank maps the other ranks' per-layer
 * prefix buffers and flag words (cudaIpc handles, exchanged by the host) and each layer's prefix K/V slice
 * is STORED straight into all of them by one kernel that then raises this rank's flag everywhere; a
 * bounded spin on the local flags replaces the collective. Without peers the exchange is an NCCL
 * all-gather (bma_scorer_set_comm). export: 3 x 64 handle bytes of this scorer; import: the handles of all
 * `size` ranks ([size][3][64], 2..8 ranks of one box) -- both collective over the ranks, repeated whenever a
 * scorer is re-created. group_score's use_peers runs the same kernels between the scorers of one process. */
/* Split prefill over peer memory: 1 (BMA_EARLY_SIGNAL=1) = the QKV GEMM walks its publishing tiles first and raises
 * this rank's flag itself, so the other ranks' wait has the rest of that GEMM as slack against rank-to-rank jitter;
 * 0 (default) = one signal + wait launch after the GEMM. Same results either way. */
int bma_scorer_set_early_signal(bma_scorer* s, int enable);
int bma_scorer_export_ipc(bma_scorer* s, void* out192);
int bma_scorer_import_ipc(bma_scorer* s, const void* handles, int size, int rank);
int bma_scorer_clear_peers(bma_scorer* s);
/* `count` data-free exchanges (flag raise on every peer + wait for all flags) back to back on `stream`; collective
 * over the ranks, no-op without peers. Timed by the caller: the latency floor of one per-layer exchange. */
int bma_scorer_peers_barrier(bma_scorer* s, int count, void* stream);
/* Nanoseconds spent in the per-layer prefix exchange waiting for other ranks' flags since the last reset
 * (device-synchronising read). The rank whose figure stays near zero is the one the others wait for. */
int bma_scorer_exchange_wait_ns(bma_scorer* s, unsigned long long* out_ns, int reset);

/* Optional per-class CUDA-event profile of the scoring pass (events on the launching stream).
 * Classes: 0 tail-embed, 1 rmsnorm, 2 gemm qkv+rope, 3 attention, 4 gemm o_proj, 5 gemm gate/up+GLU,
 * 6 gemm down, 7 gemm lm_head+CE, 8 CE finalise. profile_read sums the elapsed ms, the algorithmic
 * FLOPs and the launch count per class since the last read (arrays of >= 9 entries). */
#define BMA_PROFILE_CLASSES 9
int bma_scorer_profile(bma_scorer* s, int enable, int max_events);
int bma_scorer_profile_read(bma_scorer* s, double* ms, double* flops, long long* launches, int n_classes);

/* Kernel launches issued by this library since load (the benchmark's gpu_launches counter). */
unsigned long long bma_launch_count(void);

/* ---- kernel-level entry points (used by the parity tests) --------------------------------- */
int bma_gemm_bf16(const void* A, long long lda, const void* B, long long ldb, void* C,
                  long long ldc, const void* residual, long long ldr, int M, int N, int K,
                  int block_n, void* stream);
int bma_rmsnorm(const void* x, long long ldx, const void* w, void* y, long long ldy, int rows,
                int d, float eps, int mode, void* stream);
/* x <- bf16(x + norm(y; w_post)); xn <- norm(x; w_next) when w_next != NULL (xn may alias y). */
int bma_add_norm(const void* y, long long ldy, const void* w_post, void* x, long long ldx,
                 const void* w_next, void* xn, long long ldxn, int rows, int d, float eps, int mode,
                 void* stream);
/* Shared-prefix + causal-tail attention, head_dim 64, 128 or 256. q/k_* are row-major
 * [rows, heads*head_dim]; the value operands are TRANSPOSED: vt_tail is [kv_heads*head_dim, M]
 * (row stride ldvt), vt_pre [kv_heads*head_dim, P]. S = 0: no causal tail at all -- every row attends to
 * the P prefix keys only (plain non-causal attention of one sequence over itself when its own K / V^T are
 * handed in as the prefix: the vision tower of the scoring phase); k_tail / vt_tail must still be valid
 * pointers (they are mapped, never read). */
int bma_attention(const void* q, long long ldq, const void* k_tail, long long ldk,
                  const void* vt_tail, long long ldvt, const void* k_pre, long long ldkp,
                  const void* vt_pre, long long ldvtp, void* out, long long ldo, int M, int S, int P,
                  int n_heads, int kv_group, int head_dim, float scale, void* stream);

/* ---- gradient step (reference BimodalAttack.compute_gradient, :953-1028) ---------------------
 * Batch-1 forward + INPUT-gradient backward through the decoder and the CLIP tower on this library's
 * kernels. The host (bimodalattack_b200/gradient.py) strings the calls below together on one stream
 * (the whole step is captured into a CUDA graph). Contractions: the tcgen05 GEMM with a selectable
 * tile and epilogue; per-head attention products use its batched mode. Everything else: row kernels
 * with the rounding points of the bf16 autograd graph the reference runs. */
typedef struct {
  const void* A; long long lda;     /* [M, K] bf16, K contiguous                                   */
  const void* B; long long ldb;     /* [N, K] bf16, K contiguous:  C = A * B^T                      */
  void* C; float* C32; long long ldc; /* bf16 output, or raw fp32 accumulators when C32 != NULL     */
  const void* R; long long ldr;     /* optional bf16 residual added after the bf16 rounding (N % 8 == 0) */
  const void* bias;                 /* optional [N] bf16, added in fp32 before the rounding         */
  int M, N, K;
  int block_n, ctas;                /* tile: 64/128/160/256 columns; ctas 2 = CTA pair (block_n 256) */
  int batch;                        /* > 1: problem b reads A at (a_row b, a_col b), B at (b_row b,  */
  int a_row, a_col, b_row, b_col;   /*      b_col b) inside the same matrices, writes C + c_batch b  */
  long long c_batch;
  int transpose_out;                /* out[n, m] (N <= 128): the big operand streams once as A       */
  int causal;                       /* products of ONE causal sequence (single-CTA tiles): 1 skip tiles above the
                                       diagonal (S = Q K^T, dP = dO V^T), 2 K over keys <= the tile's last row
                                       (O = P V, dQ = dS K), 3 K over queries >= the tile's first row
                                       (dK = dS^T Q, dV = P^T dO); 0 = plain                                  */
} bma_gemm_desc;
int bma_gemm_ex(const bma_gemm_desc* g, void* stream);
/* Fused QKV projection of ONE sequence (rows at positions pos0..): Q|K rotated (HF rounding points),
 * V stored row-major into qkv (ldc spans Q|K|V) AND transposed into vt[(kv_dim), ldvt]. */
int bma_qkv_rope(const void* xn, long long ldx, const void* w_qkv, void* qkv, long long ldc, void* vt, long long ldvt,
                 const float* rope_cos, const float* rope_sin, int rows, int d, int q_dim, int kv_dim, int head_dim,
                 int pos0, void* stream);
/* P = softmax(scale S) per (head, row) of one sequence, causal or full; pad columns zeroed. */
int bma_softmax_rows(const float* S, long long lds, void* P, long long ldp, int H, int L, float scale, int causal,
                     void* stream);
/* dS = P (dP - sum(P dP)) scale */
int bma_softmax_bwd_rows(const void* P, long long ldp, const float* dP, long long ldd, void* dS, long long lds, int H,
                         int L, float scale, int causal, void* stream);
int bma_transpose_bf16(const void* src, long long lds, long long src_batch, void* dst, long long ldd,
                       long long dst_batch, int rows, int cols, int batch, void* stream);
/* dx = dres + d(rmsnorm)/dx (mode as bma_model_desc.norm_mode); dres may be NULL */
int bma_rmsnorm_bwd(const void* dy, long long lddy, const void* x, long long ldx, const void* w, const void* dres,
                    long long ldr, void* dx, long long lddx, int rows, int d, float eps, int mode, void* stream);
int bma_layernorm(const void* x, long long ldx, const void* w, const void* b, void* y, long long ldy, int rows, int d,
                  float eps, void* stream);
int bma_layernorm_bwd(const void* dy, long long lddy, const void* x, long long ldx, const void* w, const void* dres,
                      long long ldr, void* dx, long long lddx, int rows, int d, float eps, void* stream);
int bma_rope_bwd(void* g, long long ld, const float* rope_cos, const float* rope_sin, int rows, int n_heads,
                 int head_dim, int pos0, void* stream);
/* GLU on the packed gate|up layout of w_gate_up (per 256 columns: 128 gate, 128 up) and its backward */
int bma_glu(const void* gu, long long ldg, void* out, long long ldo, int rows, int d_ff, int act, void* stream);
int bma_glu_bwd(const void* gu, long long ldg, const void* dact, long long lda, void* dgu, long long ldd, int rows,
                int d_ff, int act, void* stream);
/* pointwise activation (kind 0 quick_gelu, 1 erf-GELU): out = f(x), or out = dy f'(x) when dy != NULL */
int bma_act(const void* x, const void* dy, void* out, long long n, int kind, void* stream);
/* dlogits = (softmax(logits) - onehot(target)) / T over T rows (F.cross_entropy mean, :1010-1012) */
int bma_ce_grad(const void* logits, long long ldl, const long long* target, void* dlogits, long long ldd,
                float* loss_rows, int T, int V, void* stream);
/* patch embedding as a GEMM: im2col with the Normalize folded in, and its adjoint back to fp32 pixels */
int bma_patchify(const float* img, void* patches, long long ldp, int C, int H, int W, int patch,
                 const float* mean_host, const float* std_host, void* stream);
int bma_unpatchify_grad(const void* dpatches, long long ldp, float* grad_img, int C, int H, int W, int patch,
                        const float* std_host, void* stream);
/* out[r] = a[r] + b[r % b_rows] */
int bma_add_rows(const void* a, long long lda, const void* b, long long ldb, int b_rows, void* out, long long ldo,
                 int rows, int d, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BIMODAL_B200_H */
