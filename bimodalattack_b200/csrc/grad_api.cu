// C-ABI entry points of the native gradient step (include/bimodal_b200.h, "gradient step" section):
// the generalised tcgen05 GEMM (tile shape, batched per-head operands, fp32 / bias / residual /
// transposed epilogues), the fused QKV + rotary projection for one sequence, and the row kernels
// of phase_a.cu. The Python host (bimodalattack_b200/gradient.py) strings them together on the current
// stream, so the whole step is captured into one CUDA graph after warm-up.
#include <cstdarg>
#include <cstdio>
#include <string>

#include "../../include/bimodal_b200.h"
#include "gemm_host.cuh"
#include "kernels.h"

using namespace bma;

extern "C" void bma_set_last_error(const char* msg);       // engine.cu
extern "C" void bma_count_launches(unsigned long long n);  // engine.cu

namespace {
inline cudaStream_t as_stream(void* s) { return static_cast<cudaStream_t>(s); }
inline const __nv_bfloat16* bf(const void* p) { return static_cast<const __nv_bfloat16*>(p); }
inline __nv_bfloat16* bf(void* p) { return static_cast<__nv_bfloat16*>(p); }

int done(cudaError_t e, const char* what, int launches = 1) {
  bma_count_launches(static_cast<unsigned long long>(launches));
  if (e == cudaSuccess) return 0;
  bma_set_last_error((std::string(what) + ": " + cudaGetErrorString(e)).c_str());
  return static_cast<int>(e);
}
}  // namespace

extern "C" {

int bma_gemm_ex(const bma_gemm_desc* g, void* stream) {
  if (g == nullptr) return done(cudaErrorInvalidValue, "gemm_ex: null descriptor", 0);
  GemmParams p{};
  p.M = g->M; p.N = g->N; p.K = g->K;
  p.group_m = g->batch > 1 ? (1 << 20) : group_m_for(g->K);
  p.C = bf(g->C); p.C32 = g->C32; p.ldc = g->ldc;
  p.R = bf(g->R); p.ldr = g->ldr;
  p.bias = bf(g->bias);
  p.batch = g->batch; p.a_row = g->a_row; p.a_col = g->a_col; p.b_row = g->b_row; p.b_col = g->b_col;
  p.c_batch = g->c_batch;
  if (g->causal != 0 && (g->ctas == 2 || g->transpose_out || g->causal < 0 || g->causal > 3))
    return done(cudaErrorInvalidValue, "gemm_ex: causal modes 1..3 need single-CTA tiles", 0);
  p.causal = g->causal;
  cudaStream_t st = as_stream(stream);
  const __nv_bfloat16 *A = bf(g->A), *B = bf(g->B);
  // rows of the output are written with 16-byte stores
  if (!g->transpose_out && (g->C32 != nullptr ? (g->ldc % 4 != 0 || (reinterpret_cast<uintptr_t>(g->C32) & 15))
                                              : (g->ldc % 8 != 0 || (reinterpret_cast<uintptr_t>(g->C) & 15))))
    return done(cudaErrorInvalidValue, "gemm_ex: output rows must be 16-byte aligned (ldc % 8 bf16, % 4 fp32)", 0);
  if (g->R != nullptr && (g->ldr % 8 != 0 || g->N % 8 != 0))
    return done(cudaErrorInvalidValue, "gemm_ex: residual needs N % 8 == 0 and ldr % 8 == 0", 0);
  if (g->transpose_out) {            // out[n, m]: the big operand streams once as A (M = its rows)
    if (g->batch > 1 || g->C32 != nullptr) return done(cudaErrorInvalidValue, "gemm_ex: transpose_out is plain bf16", 0);
    if (g->N <= 32) return done(launch_gemm<32, EPI_STORE_T>(A, g->lda, B, g->ldb, p, st), "gemm_ex(T,32)");
    if (g->N <= 128) return done(launch_gemm<128, EPI_STORE_T>(A, g->lda, B, g->ldb, p, st), "gemm_ex(T,128)");
    return done(cudaErrorInvalidValue, "gemm_ex: transpose_out needs N <= 128", 0);
  }
  if (g->block_n == 256 && g->ctas == 2) return done(launch_gemm<256, EPI_STORE, 128, 2>(A, g->lda, B, g->ldb, p, st), "gemm_ex(256x2)");
  if (g->block_n == 256) return done(launch_gemm<256, EPI_STORE, 128, 1>(A, g->lda, B, g->ldb, p, st), "gemm_ex(256)");
  if (g->block_n == 160) return done(launch_gemm<160, EPI_STORE, 128, 1>(A, g->lda, B, g->ldb, p, st), "gemm_ex(160)");
  if (g->block_n == 128) return done(launch_gemm<128, EPI_STORE, 128, 1>(A, g->lda, B, g->ldb, p, st), "gemm_ex(128)");
  if (g->block_n == 64) return done(launch_gemm<64, EPI_STORE, 128, 1>(A, g->lda, B, g->ldb, p, st), "gemm_ex(64)");
  return done(cudaErrorInvalidValue, "gemm_ex: block_n must be 64, 128, 160 or 256", 0);
}

int bma_qkv_rope(const void* xn, long long ldx, const void* w_qkv, void* qkv, long long ldc, void* vt, long long ldvt,
                 const float* rope_cos, const float* rope_sin, int rows, int d, int q_dim, int kv_dim, int head_dim,
                 int pos0, void* stream) {
  GemmParams p{};
  p.M = rows; p.N = q_dim + 2 * kv_dim; p.K = d; p.group_m = group_m_for(d);
  p.C = bf(qkv); p.ldc = ldc;
  p.Vt = bf(vt); p.ldvt = ldvt; p.v_rowmajor = 1;
  p.rope_cos = rope_cos; p.rope_sin = rope_sin; p.rope_cols = q_dim + kv_dim;
  p.rows_per_seq = rows > 0 ? rows : 1; p.pos_base = pos0;
  cudaStream_t st = as_stream(stream);
  if (head_dim == 128) return done(launch_gemm<256, EPI_ROPE, 128, 1>(bf(xn), ldx, bf(w_qkv), d, p, st), "qkv_rope");
  if (head_dim == 64) return done(launch_gemm<256, EPI_ROPE, 64, 1>(bf(xn), ldx, bf(w_qkv), d, p, st), "qkv_rope");
  return done(cudaErrorInvalidValue, "qkv_rope: head_dim must be 64 or 128", 0);
}

int bma_softmax_rows(const float* S, long long lds, void* P, long long ldp, int H, int L, float scale, int causal,
                     void* stream) {
  return done(launch_softmax_rows(S, lds, bf(P), ldp, H, L, scale, causal, as_stream(stream)), "softmax_rows");
}
int bma_softmax_bwd_rows(const void* P, long long ldp, const float* dP, long long ldd, void* dS, long long lds, int H,
                         int L, float scale, int causal, void* stream) {
  return done(launch_softmax_bwd_rows(bf(P), ldp, dP, ldd, bf(dS), lds, H, L, scale, causal, as_stream(stream)),
              "softmax_bwd_rows");
}
int bma_transpose_bf16(const void* src, long long lds, long long src_batch, void* dst, long long ldd,
                       long long dst_batch, int rows, int cols, int batch, void* stream) {
  return done(launch_transpose_bf16(bf(src), lds, src_batch, bf(dst), ldd, dst_batch, rows, cols, batch,
                                    as_stream(stream)), "transpose_bf16");
}
int bma_rmsnorm_bwd(const void* dy, long long lddy, const void* x, long long ldx, const void* w, const void* dres,
                    long long ldr, void* dx, long long lddx, int rows, int d, float eps, int mode, void* stream) {
  return done(launch_rmsnorm_bwd(bf(dy), lddy, bf(x), ldx, bf(w), bf(dres), ldr, bf(dx), lddx, rows, d, eps, mode,
                                 as_stream(stream)), "rmsnorm_bwd");
}
int bma_layernorm(const void* x, long long ldx, const void* w, const void* b, void* y, long long ldy, int rows, int d,
                  float eps, void* stream) {
  return done(launch_layernorm(bf(x), ldx, bf(w), bf(b), bf(y), ldy, rows, d, eps, as_stream(stream)), "layernorm");
}
int bma_layernorm_bwd(const void* dy, long long lddy, const void* x, long long ldx, const void* w, const void* dres,
                      long long ldr, void* dx, long long lddx, int rows, int d, float eps, void* stream) {
  return done(launch_layernorm_bwd(bf(dy), lddy, bf(x), ldx, bf(w), bf(dres), ldr, bf(dx), lddx, rows, d, eps,
                                   as_stream(stream)), "layernorm_bwd");
}
int bma_rope_bwd(void* g, long long ld, const float* rope_cos, const float* rope_sin, int rows, int n_heads,
                 int head_dim, int pos0, void* stream) {
  return done(launch_rope_bwd(bf(g), ld, rope_cos, rope_sin, rows, n_heads, head_dim, pos0, as_stream(stream)),
              "rope_bwd");
}
int bma_glu(const void* gu, long long ldg, void* out, long long ldo, int rows, int d_ff, int act, void* stream) {
  return done(launch_glu(bf(gu), ldg, bf(out), ldo, rows, d_ff, act, as_stream(stream)), "glu");
}
int bma_glu_bwd(const void* gu, long long ldg, const void* dact, long long lda, void* dgu, long long ldd, int rows,
                int d_ff, int act, void* stream) {
  return done(launch_glu_bwd(bf(gu), ldg, bf(dact), lda, bf(dgu), ldd, rows, d_ff, act, as_stream(stream)), "glu_bwd");
}
int bma_act(const void* x, const void* dy, void* out, long long n, int kind, void* stream) {
  return done(launch_act(bf(x), bf(dy), bf(out), n, kind, as_stream(stream)), "act");
}
int bma_ce_grad(const void* logits, long long ldl, const long long* target, void* dlogits, long long ldd,
                float* loss_rows, int T, int V, void* stream) {
  return done(launch_ce_grad(bf(logits), ldl, target, bf(dlogits), ldd, loss_rows, T, V, as_stream(stream)), "ce_grad");
}
int bma_patchify(const float* img, void* patches, long long ldp, int C, int H, int W, int patch,
                 const float* mean_host, const float* std_host, void* stream) {
  return done(launch_patchify(img, bf(patches), ldp, C, H, W, patch, mean_host, std_host, as_stream(stream)), "patchify");
}
int bma_unpatchify_grad(const void* dpatches, long long ldp, float* grad_img, int C, int H, int W, int patch,
                        const float* std_host, void* stream) {
  return done(launch_unpatchify_grad(bf(dpatches), ldp, grad_img, C, H, W, patch, std_host, as_stream(stream)),
              "unpatchify_grad");
}
int bma_add_rows(const void* a, long long lda, const void* b, long long ldb, int b_rows, void* out, long long ldo,
                 int rows, int d, void* stream) {
  return done(launch_add_rows(bf(a), lda, bf(b), ldb, b_rows, bf(out), ldo, rows, d, as_stream(stream)), "add_rows");
}

}  // extern "C"
