// Launch entry points shared between the kernel translation units and the C-ABI layer.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace bma {

struct AttnParams {
  const __nv_bfloat16* q;       long long ldq;     // [M, ...] query rows (head h at column h*head_dim)
  const __nv_bfloat16* k_tail;  long long ldk;     // [M, ...] tail keys   (kv head g at g*head_dim)
  const __nv_bfloat16* vt_tail; long long ldvt;    // [Hkv*HD, M] tail values, TRANSPOSED (row = dim)
  const __nv_bfloat16* k_pre;   long long ldkp;    // [P, ...] prefix keys
  const __nv_bfloat16* vt_pre;  long long ldvtp;   // [Hkv*HD, P] prefix values, transposed
  __nv_bfloat16* out;           long long ldo;     // [M, H*HD]
  int M, S, P;
  const int* M_dev;             // optional live row count on the device (<= M; M bounds grid and TMA maps)
  // ragged rows (suffix-head sharing): row_meta[r] = (first row of r's candidate, prefix keys
  // visible to r); all rows see prefix keys < P_shared. null = uniform S-row candidates.
  const int2* row_meta;
  int P_shared;
  int pre_rows;                 // leading rows that are prefix positions themselves (fused prefill)
  int no_tail;                  // 1: rows attend to the P prefix keys only (no causal tail at all): plain
                                // non-causal attention of one sequence over itself (vision tower) when K/V^T
                                // of all its rows are handed in as the "prefix"
  int kv_group;                 // query heads per kv head
  int head_dim;                 // 128 or 256
  float scale_log2;             // softmax scale * log2(e)
};

cudaError_t launch_prefix_tail_attention(const AttnParams& p, int n_heads, cudaStream_t stream);
// head_dim 128: split-key kernel, the two softmax threads of a row never synchronise inside an item
cudaError_t launch_attention_split(const AttnParams& p, int n_heads, cudaStream_t stream);
// head_dim 128: 64-key blocks, small enough for TWO resident CTAs per SM (the default)
cudaError_t launch_attention_occ(const AttnParams& p, int n_heads, cudaStream_t stream);

cudaError_t launch_pgd_step(float* x, const float* g, const float* x0, float* norm_out,
                            long long n, float step, float eps, const float* mean,
                            const float* std_, int channels, long long plane, cudaStream_t stream);

size_t sampler_workspace_bytes(int n_tok, int V, int topk);
cudaError_t launch_sample_ids(const long long* ids, const void* grad, int grad_dtype,
                              long long row_stride, int n_tok, int V, const uint32_t* mask_bits,
                              int search_width, int topk, int n_replace, unsigned long long seed,
                              unsigned long long offset, void* workspace, long long* topk_ids_out,
                              long long* new_ids, unsigned long long* offset_advance,
                              cudaStream_t stream);

cudaError_t launch_build_tail_embeds(const long long* ids, int ids_stride, const int* slot,
                                     const __nv_bfloat16* fixed, const __nv_bfloat16* W,
                                     long long ldw, int V, float scale, __nv_bfloat16* out, int rows,
                                     int S, int d, const int2* row_info, const long long* base_ids,
                                     const __nv_bfloat16* prefix_embeds, int P, const int* rows_dev,
                                     cudaStream_t stream);
cudaError_t launch_plan_rows(const long long* ids, int ids_stride, const long long* base_ids,
                             const int* slot, int S, int n_share, int B, int include_base, int P, int dedupe,
                             int pre0, int pre_n, int2* row_info, int2* row_meta, int* cand_vstart, int2* cand_tmp,
                             int* total_rows, unsigned long long* rows_accum, cudaStream_t stream);
cudaError_t launch_append_prefix_kv(const __nv_bfloat16* k_tail, long long ldk, const __nv_bfloat16* vt_tail,
                                    long long ldvt, __nv_bfloat16* k_pre, long long ldkp,
                                    __nv_bfloat16* vt_pre, long long ldvtp, int n, int kv_dim, int P,
                                    cudaStream_t stream);
cudaError_t launch_unpack_prefix_kv(const __nv_bfloat16* recv, long long blk, int chunk, int n_ranks, int P,
                                    const __nv_bfloat16* k_tail, long long ldk, const __nv_bfloat16* vt_tail,
                                    long long ldvt, int base_row0, int n_base, __nv_bfloat16* k_pre, long long ldkp,
                                    __nv_bfloat16* vt_pre, long long ldvtp, int kv_dim, cudaStream_t stream);
cudaError_t launch_prefix_scatter(const __nv_bfloat16* k_tail, long long ldk, const __nv_bfloat16* vt_tail, long long ldvt,
                                  void* const* peer_k, void* const* peer_vt, void* const* peer_flags, int n_ranks,
                                  int my_rank, long long ldkp, long long ldvtp, int n, int key0, int kv_dim,
                                  unsigned int* done_counter, unsigned long long epoch, cudaStream_t stream);
cudaError_t launch_prefix_wait(unsigned long long* flags, int n_ranks, unsigned long long epoch, cudaStream_t stream,
                               unsigned long long* wait_ns = nullptr);
cudaError_t launch_prefix_signal_wait(void* const* peer_flags, unsigned long long* my_flags, int n_ranks, int my_rank,
                                      unsigned long long epoch, int signal, int wait, cudaStream_t stream,
                                      unsigned long long* wait_ns = nullptr);
cudaError_t launch_rmsnorm(const __nv_bfloat16* x, long long ldx, const __nv_bfloat16* w,
                           __nv_bfloat16* y, long long ldy, int rows, int d, float eps, int mode,
                           int gather_T, int gather_S, int gather_off, const int* gather_vstart,
                           cudaStream_t stream, const int* rows_dev = nullptr);
cudaError_t launch_gather_rows(const __nv_bfloat16* src, long long lds, __nv_bfloat16* dst, long long ldd,
                               int rows, int width, int T, int S, int off, const int* vstart,
                               cudaStream_t stream);
cudaError_t launch_add_norm(const __nv_bfloat16* y, long long ldy, const __nv_bfloat16* w_post,
                            __nv_bfloat16* x, long long ldx, const __nv_bfloat16* w_next,
                            __nv_bfloat16* xn, long long ldxn, int rows, int d, float eps, int mode,
                            cudaStream_t stream, const int* rows_dev = nullptr);
cudaError_t launch_init_key(unsigned long long* key, int* flag, cudaStream_t stream);
cudaError_t launch_ce_finalize(const float2* part, const float* tgt_logit, const int* arg,
                               const long long* target_ids, int n_tiles, int B,
                               int T, float* loss, unsigned long long* key, int* all_match_flag,
                               unsigned int idx_base, cudaStream_t stream);

// ---- native gradient step (phase_a.cu) ---------------------------------------------------------
cudaError_t launch_softmax_rows(const float* S, long long lds, __nv_bfloat16* P, long long ldp, int H, int L,
                                float scale, int causal, cudaStream_t st);
cudaError_t launch_softmax_bwd_rows(const __nv_bfloat16* P, long long ldp, const float* dP, long long ldd,
                                    __nv_bfloat16* dS, long long lds, int H, int L, float scale, int causal,
                                    cudaStream_t st);
cudaError_t launch_transpose_bf16(const __nv_bfloat16* src, long long lds, long long src_batch, __nv_bfloat16* dst,
                                  long long ldd, long long dst_batch, int rows, int cols, int batch, cudaStream_t st);
cudaError_t launch_rmsnorm_bwd(const __nv_bfloat16* dy, long long lddy, const __nv_bfloat16* x, long long ldx,
                               const __nv_bfloat16* w, const __nv_bfloat16* dres, long long ldr, __nv_bfloat16* dx,
                               long long lddx, int rows, int d, float eps, int mode, cudaStream_t st);
cudaError_t launch_layernorm(const __nv_bfloat16* x, long long ldx, const __nv_bfloat16* w, const __nv_bfloat16* b,
                             __nv_bfloat16* y, long long ldy, int rows, int d, float eps, cudaStream_t st);
cudaError_t launch_layernorm_bwd(const __nv_bfloat16* dy, long long lddy, const __nv_bfloat16* x, long long ldx,
                                 const __nv_bfloat16* w, const __nv_bfloat16* dres, long long ldr, __nv_bfloat16* dx,
                                 long long lddx, int rows, int d, float eps, cudaStream_t st);
cudaError_t launch_rope_bwd(__nv_bfloat16* g, long long ld, const float* cosT, const float* sinT, int rows,
                            int n_heads, int hd, int pos0, cudaStream_t st);
cudaError_t launch_glu(const __nv_bfloat16* gu, long long ldg, __nv_bfloat16* out, long long ldo, int rows, int d_ff,
                       int act, cudaStream_t st);
cudaError_t launch_glu_bwd(const __nv_bfloat16* gu, long long ldg, const __nv_bfloat16* dact, long long lda,
                           __nv_bfloat16* dgu, long long ldd, int rows, int d_ff, int act, cudaStream_t st);
cudaError_t launch_act(const __nv_bfloat16* x, const __nv_bfloat16* dy, __nv_bfloat16* out, long long n, int kind,
                       cudaStream_t st);
cudaError_t launch_ce_grad(const __nv_bfloat16* logits, long long ldl, const long long* target, __nv_bfloat16* dlogits,
                           long long ldd, float* loss_rows, int T, int V, cudaStream_t st);
cudaError_t launch_patchify(const float* img, __nv_bfloat16* patches, long long ldp, int C, int H, int W, int p,
                            const float* mean, const float* sd, cudaStream_t st);
cudaError_t launch_unpatchify_grad(const __nv_bfloat16* dpatches, long long ldp, float* gimg, int C, int H, int W,
                                   int p, const float* sd, cudaStream_t st);
cudaError_t launch_add_rows(const __nv_bfloat16* a, long long lda, const __nv_bfloat16* b, long long ldb, int b_rows,
                            __nv_bfloat16* out, long long ldo, int rows, int d, cudaStream_t st);

}  // namespace bma
