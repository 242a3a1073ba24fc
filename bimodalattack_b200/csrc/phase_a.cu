// Row / element kernels of the native gradient step (reference BimodalAttack.compute_gradient,
// bimodal_attack.py:953-1028): the batch-1 forward + INPUT-gradient backward through the decoder and
// the CLIP tower. The contractions run on the tcgen05 GEMM (gemm_tcgen05.cuh, incl. its batched mode
// for the per-head attention products); this file holds what sits between them. All of it is
// HBM/latency-bound row work: one warp per row, 16-byte accesses, fp32 arithmetic with the rounding
// points of the bf16 autograd graph the reference runs (every tensor that autograd materialises in
// bf16 is rounded here too).
//
//   softmax_rows / softmax_bwd_rows   attention probabilities of one sequence (causal or full) from
//                                     fp32 scores, and dS = P * (dP - sum(P dP)) * scale
//   transpose_bf16                    batched 2-D transpose (operands of the K-major GEMM)
//   rmsnorm_bwd / layernorm(+bwd)     input gradients of the norms (no weight gradients: nothing
//                                     upstream of the inputs is trained)
//   rope_bwd                          rotation by -theta on dQ | dK
//   glu / glu_bwd                     SwiGLU / GeGLU on the gate|up-interleaved layout of the packed weights
//   act_fwd / act_bwd                 quick_gelu (CLIP MLP) and erf-GELU (projector)
//   ce_grad                           dlogits = (softmax(logits) - onehot) / T for the T target rows
#include "common.cuh"
#include "kernels.h"

namespace bma {

namespace {

__device__ __forceinline__ float ldbf(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void unpack8(const uint4& q, float (&f)[8]) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) { f[2 * j] = bf16_lo(w[j]); f[2 * j + 1] = bf16_hi(w[j]); }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}
inline int warp_blocks(long long rows) { return static_cast<int>((rows * 32 + 255) / 256); }

// Row reductions with TPR threads per row: 32 (one warp per row, 8 rows per block) for tall matrices,
// 256 (one block per row) when there are too few rows to fill the machine with warps (the batch-1
// gradient step: a few hundred rows).
template <int TPR>
__device__ __forceinline__ float row_sum(float v, float* red) {
  v = warp_sum(v);
  if constexpr (TPR == 32) {
    return v;
  } else {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();                       // `red` may still be read from the previous reduction
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = 0.0f;
#pragma unroll
    for (int i = 0; i < TPR / 32; ++i) t += red[i];
    return t;
  }
}
inline bool few_rows(int rows) { return rows < 4096; }

}  // namespace

// ------------------------------------------------------------------------------------------------
// P[h, r, :] = softmax(scale * S[h, r, 0:valid]) with valid = r + 1 (causal) or L; columns >= valid
// (up to ldp) are written as zeros so the K-major GEMMs that consume P need no masking.
__global__ void __launch_bounds__(256)
softmax_rows_kernel(const float* __restrict__ S, long long lds, __nv_bfloat16* __restrict__ P, long long ldp,
                    int H, int L, float scale_log2, int causal) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= (long long)H * L) return;
  const int r = static_cast<int>(row % L);
  const int valid = causal ? r + 1 : L;
  const float* s = S + row * lds;
  __nv_bfloat16* p = P + row * ldp;
  float m = -INFINITY;
  for (int j = lane; j < valid; j += 32) m = fmaxf(m, s[j]);
  m = warp_max(m) * scale_log2;
  float sum = 0.0f;
  for (int j = lane; j < valid; j += 32) sum += exp2f(fmaf(s[j], scale_log2, -m));
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  for (int j = lane; j < ldp; j += 32)
    p[j] = __float2bfloat16_rn(j < valid ? exp2f(fmaf(s[j], scale_log2, -m)) * inv : 0.0f);
}

cudaError_t launch_softmax_rows(const float* S, long long lds, __nv_bfloat16* P, long long ldp, int H, int L,
                                float scale, int causal, cudaStream_t st) {
  if (H <= 0 || L <= 0) return cudaSuccess;
  softmax_rows_kernel<<<warp_blocks((long long)H * L), 256, 0, st>>>(S, lds, P, ldp, H, L,
                                                                     scale * 1.4426950408889634f, causal);
  return cudaGetLastError();
}

// dS[h, r, j] = P * (dP - delta) * scale, delta = sum_j P dP over the valid columns; zeros elsewhere.
__global__ void __launch_bounds__(256)
softmax_bwd_rows_kernel(const __nv_bfloat16* __restrict__ P, long long ldp, const float* __restrict__ dP,
                        long long ldd, __nv_bfloat16* __restrict__ dS, long long lds, int H, int L, float scale,
                        int causal) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= (long long)H * L) return;
  const int r = static_cast<int>(row % L);
  const int valid = causal ? r + 1 : L;
  const __nv_bfloat16* p = P + row * ldp;
  const float* dp = dP + row * ldd;
  __nv_bfloat16* ds = dS + row * lds;
  float delta = 0.0f;
  for (int j = lane; j < valid; j += 32) delta = fmaf(ldbf(p + j), dp[j], delta);
  delta = warp_sum(delta);
  for (int j = lane; j < lds; j += 32)
    ds[j] = __float2bfloat16_rn(j < valid ? ldbf(p + j) * (dp[j] - delta) * scale : 0.0f);
}

cudaError_t launch_softmax_bwd_rows(const __nv_bfloat16* P, long long ldp, const float* dP, long long ldd,
                                    __nv_bfloat16* dS, long long lds, int H, int L, float scale, int causal,
                                    cudaStream_t st) {
  if (H <= 0 || L <= 0) return cudaSuccess;
  softmax_bwd_rows_kernel<<<warp_blocks((long long)H * L), 256, 0, st>>>(P, ldp, dP, ldd, dS, lds, H, L, scale, causal);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// dst[b][c][r] = src[b][r][c] for r < rows, c < cols (32 x 32 tiles through shared memory).
__global__ void __launch_bounds__(256)
transpose_bf16_kernel(const __nv_bfloat16* __restrict__ src, long long lds, long long src_batch,
                      __nv_bfloat16* __restrict__ dst, long long ldd, long long dst_batch, int rows, int cols) {
  __shared__ __nv_bfloat16 tile[32][34];
  const int b = blockIdx.z;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 32 x 8
  const __nv_bfloat16* s = src + (long long)b * src_batch;
  __nv_bfloat16* d = dst + (long long)b * dst_batch;
#pragma unroll
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    tile[i][tx] = (r < rows && c < cols) ? s[(long long)r * lds + c] : __float2bfloat16_rn(0.0f);
  }
  __syncthreads();
#pragma unroll
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (c < cols && r < rows) d[(long long)c * ldd + r] = tile[tx][i];
  }
}

cudaError_t launch_transpose_bf16(const __nv_bfloat16* src, long long lds, long long src_batch, __nv_bfloat16* dst,
                                  long long ldd, long long dst_batch, int rows, int cols, int batch, cudaStream_t st) {
  if (rows <= 0 || cols <= 0 || batch <= 0) return cudaSuccess;
  dim3 grid((cols + 31) / 32, (rows + 31) / 32, batch);
  transpose_bf16_kernel<<<grid, 256, 0, st>>>(src, lds, src_batch, dst, ldd, dst_batch, rows, cols);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// RMSNorm input gradient, one warp per row (d % 8 == 0), HF rounding points:
//   forward   h = bf16(x * rstd);  y = w * h  (mode 0, Llama)    |   y = bf16(x * rstd * (1 + w))  (mode 1, Gemma)
//   backward  dh = bf16(dy * w)  [mode 1: dy * (1 + w), fp32];  dx = bf16(rstd * (dh - xhat * mean(dh * xhat)))
//   out       dx_out = bf16(dres + dx)   (the residual branch's gradient, when given)
template <int TPR>
__global__ void __launch_bounds__(256)
rmsnorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy, long long lddy, const __nv_bfloat16* __restrict__ x,
                   long long ldx, const __nv_bfloat16* __restrict__ w, const __nv_bfloat16* dres, long long ldr,
                   __nv_bfloat16* dx, long long lddx, int rows, int d, float eps, int mode) {
  __shared__ float red[8];
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) / TPR;
  const int lane = threadIdx.x % TPR;
  if (r >= rows) return;      // whole rows (warps / blocks) leave together
  const uint4* xv = reinterpret_cast<const uint4*>(x + (long long)r * ldx);
  const uint4* gv = reinterpret_cast<const uint4*>(dy + (long long)r * lddy);
  const uint4* wv = reinterpret_cast<const uint4*>(w);
  const int nvec = d >> 3;
  float ss = 0.0f;
  for (int v = lane; v < nvec; v += TPR) {
    float a[8];
    unpack8(xv[v], a);
#pragma unroll
    for (int j = 0; j < 8; ++j) ss = fmaf(a[j], a[j], ss);
  }
  ss = row_sum<TPR>(ss, red);
  const float rstd = rsqrtf(ss / static_cast<float>(d) + eps);
  float dot = 0.0f;                       // sum(dh * xhat)
  for (int v = lane; v < nvec; v += TPR) {
    float a[8], g[8], ww[8];
    unpack8(xv[v], a); unpack8(gv[v], g); unpack8(wv[v], ww);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float dh = mode == 0 ? bf16_round(g[j] * ww[j]) : g[j] * (1.0f + ww[j]);
      dot = fmaf(dh, a[j] * rstd, dot);
    }
  }
  dot = row_sum<TPR>(dot, red) / static_cast<float>(d);
  const uint4* rv = dres != nullptr ? reinterpret_cast<const uint4*>(dres + (long long)r * ldr) : nullptr;
  uint4* ov = reinterpret_cast<uint4*>(dx + (long long)r * lddx);
  for (int v = lane; v < nvec; v += TPR) {
    float a[8], g[8], ww[8], o[8];
    unpack8(xv[v], a); unpack8(gv[v], g); unpack8(wv[v], ww);
    float rs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (rv != nullptr) unpack8(rv[v], rs);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float dh = mode == 0 ? bf16_round(g[j] * ww[j]) : g[j] * (1.0f + ww[j]);
      const float gx = bf16_round(rstd * (dh - a[j] * rstd * dot));
      o[j] = rv != nullptr ? rs[j] + gx : gx;
    }
    ov[v] = pack8(o);
  }
}

cudaError_t launch_rmsnorm_bwd(const __nv_bfloat16* dy, long long lddy, const __nv_bfloat16* x, long long ldx,
                               const __nv_bfloat16* w, const __nv_bfloat16* dres, long long ldr, __nv_bfloat16* dx,
                               long long lddx, int rows, int d, float eps, int mode, cudaStream_t st) {
  if (rows <= 0) return cudaSuccess;
  if ((d & 7) || (lddy & 7) || (ldx & 7) || (lddx & 7) || (dres != nullptr && (ldr & 7))) return cudaErrorInvalidValue;
  if (few_rows(rows)) rmsnorm_bwd_kernel<256><<<rows, 256, 0, st>>>(dy, lddy, x, ldx, w, dres, ldr, dx, lddx, rows, d, eps, mode);
  else rmsnorm_bwd_kernel<32><<<warp_blocks(rows), 256, 0, st>>>(dy, lddy, x, ldx, w, dres, ldr, dx, lddx, rows, d, eps, mode);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// LayerNorm (CLIP): y = bf16((x - mean) * rstd * w + b), statistics in fp32; optional residual-free.
template <int TPR>
__global__ void __launch_bounds__(256)
layernorm_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, const __nv_bfloat16* __restrict__ w,
                 const __nv_bfloat16* __restrict__ b, __nv_bfloat16* __restrict__ y, long long ldy, int rows, int d,
                 float eps) {
  __shared__ float red[8];
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) / TPR;
  const int lane = threadIdx.x % TPR;
  if (r >= rows) return;      // whole rows (warps / blocks) leave together
  const uint4* xv = reinterpret_cast<const uint4*>(x + (long long)r * ldx);
  const int nvec = d >> 3;
  float s1 = 0.0f;
  for (int v = lane; v < nvec; v += TPR) {
    float a[8];
    unpack8(xv[v], a);
#pragma unroll
    for (int j = 0; j < 8; ++j) s1 += a[j];
  }
  const float mean = row_sum<TPR>(s1, red) / static_cast<float>(d);
  float s2 = 0.0f;
  for (int v = lane; v < nvec; v += TPR) {
    float a[8];
    unpack8(xv[v], a);
#pragma unroll
    for (int j = 0; j < 8; ++j) { const float c = a[j] - mean; s2 = fmaf(c, c, s2); }
  }
  const float rstd = rsqrtf(row_sum<TPR>(s2, red) / static_cast<float>(d) + eps);
  const uint4* wv = reinterpret_cast<const uint4*>(w);
  const uint4* bv = reinterpret_cast<const uint4*>(b);
  uint4* yv = reinterpret_cast<uint4*>(y + (long long)r * ldy);
  for (int v = lane; v < nvec; v += TPR) {
    float a[8], ww[8], bb[8], o[8];
    unpack8(xv[v], a); unpack8(wv[v], ww); unpack8(bv[v], bb);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = (a[j] - mean) * rstd * ww[j] + bb[j];
    yv[v] = pack8(o);
  }
}

cudaError_t launch_layernorm(const __nv_bfloat16* x, long long ldx, const __nv_bfloat16* w, const __nv_bfloat16* b,
                             __nv_bfloat16* y, long long ldy, int rows, int d, float eps, cudaStream_t st) {
  if (rows <= 0) return cudaSuccess;
  if ((d & 7) || (ldx & 7) || (ldy & 7)) return cudaErrorInvalidValue;
  if (few_rows(rows)) layernorm_kernel<256><<<rows, 256, 0, st>>>(x, ldx, w, b, y, ldy, rows, d, eps);
  else layernorm_kernel<32><<<warp_blocks(rows), 256, 0, st>>>(x, ldx, w, b, y, ldy, rows, d, eps);
  return cudaGetLastError();
}

// dx = bf16(dres + rstd * (g - mean(g) - xhat * mean(g * xhat))), g = dy * w
template <int TPR>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy, long long lddy, const __nv_bfloat16* __restrict__ x,
                     long long ldx, const __nv_bfloat16* __restrict__ w, const __nv_bfloat16* dres, long long ldr,
                     __nv_bfloat16* dx, long long lddx, int rows, int d, float eps) {
  __shared__ float red[8];
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) / TPR;
  const int lane = threadIdx.x % TPR;
  if (r >= rows) return;      // whole rows (warps / blocks) leave together
  const uint4* xv = reinterpret_cast<const uint4*>(x + (long long)r * ldx);
  const uint4* gv = reinterpret_cast<const uint4*>(dy + (long long)r * lddy);
  const uint4* wv = reinterpret_cast<const uint4*>(w);
  const int nvec = d >> 3;
  float s1 = 0.0f;
  for (int v = lane; v < nvec; v += TPR) {
    float a[8];
    unpack8(xv[v], a);
#pragma unroll
    for (int j = 0; j < 8; ++j) s1 += a[j];
  }
  const float mean = row_sum<TPR>(s1, red) / static_cast<float>(d);
  float s2 = 0.0f;
  for (int v = lane; v < nvec; v += TPR) {
    float a[8];
    unpack8(xv[v], a);
#pragma unroll
    for (int j = 0; j < 8; ++j) { const float c = a[j] - mean; s2 = fmaf(c, c, s2); }
  }
  const float rstd = rsqrtf(row_sum<TPR>(s2, red) / static_cast<float>(d) + eps);
  float sg = 0.0f, sgx = 0.0f;
  for (int v = lane; v < nvec; v += TPR) {
    float a[8], g[8], ww[8];
    unpack8(xv[v], a); unpack8(gv[v], g); unpack8(wv[v], ww);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float gg = g[j] * ww[j];
      sg += gg;
      sgx = fmaf(gg, (a[j] - mean) * rstd, sgx);
    }
  }
  sg = row_sum<TPR>(sg, red) / static_cast<float>(d);
  sgx = row_sum<TPR>(sgx, red) / static_cast<float>(d);
  const uint4* rv = dres != nullptr ? reinterpret_cast<const uint4*>(dres + (long long)r * ldr) : nullptr;
  uint4* ov = reinterpret_cast<uint4*>(dx + (long long)r * lddx);
  for (int v = lane; v < nvec; v += TPR) {
    float a[8], g[8], ww[8], o[8];
    unpack8(xv[v], a); unpack8(gv[v], g); unpack8(wv[v], ww);
    float rs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (rv != nullptr) unpack8(rv[v], rs);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float gx = bf16_round(rstd * (g[j] * ww[j] - sg - (a[j] - mean) * rstd * sgx));
      o[j] = rv != nullptr ? rs[j] + gx : gx;
    }
    ov[v] = pack8(o);
  }
}

cudaError_t launch_layernorm_bwd(const __nv_bfloat16* dy, long long lddy, const __nv_bfloat16* x, long long ldx,
                                 const __nv_bfloat16* w, const __nv_bfloat16* dres, long long ldr, __nv_bfloat16* dx,
                                 long long lddx, int rows, int d, float eps, cudaStream_t st) {
  if (rows <= 0) return cudaSuccess;
  if ((d & 7) || (lddy & 7) || (ldx & 7) || (lddx & 7) || (dres != nullptr && (ldr & 7))) return cudaErrorInvalidValue;
  if (few_rows(rows)) layernorm_bwd_kernel<256><<<rows, 256, 0, st>>>(dy, lddy, x, ldx, w, dres, ldr, dx, lddx, rows, d, eps);
  else layernorm_bwd_kernel<32><<<warp_blocks(rows), 256, 0, st>>>(dy, lddy, x, ldx, w, dres, ldr, dx, lddx, rows, d, eps);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Rotary backward, in place on [rows, n_heads * hd]: the forward is y1 = bf16(x1 c) + bf16(-x2 s),
// y2 = bf16(x2 c) + bf16(x1 s) (HF apply_rotary_pos_emb in bf16), so
//   dx1 = bf16(d1 c) + bf16(d2 s),   dx2 = bf16(d2 c) + bf16(-(d1 s));   position of row r is pos0 + r.
__global__ void __launch_bounds__(256)
rope_bwd_kernel(__nv_bfloat16* g, long long ld, const float* __restrict__ cosT, const float* __restrict__ sinT,
                int rows, int n_heads, int hd, int pos0) {
  const int half = hd >> 1;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // (row, head, j < half)
  const long long total = (long long)rows * n_heads * half;
  if (i >= total) return;
  const int j = static_cast<int>(i % half);
  const long long t = i / half;
  const int h = static_cast<int>(t % n_heads);
  const int r = static_cast<int>(t / n_heads);
  const float c = cosT[(long long)(pos0 + r) * half + j], s = sinT[(long long)(pos0 + r) * half + j];
  __nv_bfloat16* p = g + (long long)r * ld + h * hd + j;
  const float d1 = ldbf(p), d2 = ldbf(p + half);
  p[0] = __float2bfloat16_rn(bf16_round(d1 * c) + bf16_round(d2 * s));
  p[half] = __float2bfloat16_rn(bf16_round(d2 * c) + bf16_round(-(d1 * s)));
}

cudaError_t launch_rope_bwd(__nv_bfloat16* g, long long ld, const float* cosT, const float* sinT, int rows,
                            int n_heads, int hd, int pos0, cudaStream_t st) {
  const long long total = (long long)rows * n_heads * (hd >> 1);
  if (total <= 0) return cudaSuccess;
  rope_bwd_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, st>>>(g, ld, cosT, sinT, rows, n_heads, hd, pos0);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// GLU on the packed gate|up layout (per 256 columns of `gu`: 128 gate values, then the 128 matching
// up values -- the row order of the packed w_gate_up): out[r, 128 b + j] = bf16(act(g)) * u.
__device__ __forceinline__ float glu_act(float g, int act) {
  if (act == 0) return g / (1.0f + __expf(-g));
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  return 0.5f * g * (1.0f + tanhf(k0 * (g + k1 * g * g * g)));
}
__device__ __forceinline__ float glu_act_grad(float g, int act) {
  if (act == 0) {
    const float sg = 1.0f / (1.0f + __expf(-g));
    return sg * (1.0f + g * (1.0f - sg));
  }
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  const float u = k0 * (g + k1 * g * g * g), t = tanhf(u);
  return 0.5f * (1.0f + t) + 0.5f * g * (1.0f - t * t) * k0 * (1.0f + 3.0f * k1 * g * g);
}

__global__ void __launch_bounds__(256)
glu_kernel(const __nv_bfloat16* __restrict__ gu, long long ldg, __nv_bfloat16* __restrict__ out, long long ldo,
           int rows, int d_ff, int act) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // one thread per 8 outputs
  const int per_row = d_ff >> 3;
  if (i >= (long long)rows * per_row) return;
  const int r = static_cast<int>(i / per_row), c = static_cast<int>(i % per_row) * 8;
  const int blk = c >> 7, j = c & 127;
  const __nv_bfloat16* src = gu + (long long)r * ldg + blk * 256 + j;
  float g[8], u[8], o[8];
  unpack8(*reinterpret_cast<const uint4*>(src), g);
  unpack8(*reinterpret_cast<const uint4*>(src + 128), u);
#pragma unroll
  for (int k = 0; k < 8; ++k) o[k] = bf16_round(glu_act(g[k], act)) * u[k];
  *reinterpret_cast<uint4*>(out + (long long)r * ldo + c) = pack8(o);
}

cudaError_t launch_glu(const __nv_bfloat16* gu, long long ldg, __nv_bfloat16* out, long long ldo, int rows, int d_ff,
                       int act, cudaStream_t st) {
  if (rows <= 0) return cudaSuccess;
  if ((d_ff & 127) || (ldg & 7) || (ldo & 7)) return cudaErrorInvalidValue;
  const long long n = (long long)rows * (d_ff >> 3);
  glu_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, st>>>(gu, ldg, out, ldo, rows, d_ff, act);
  return cudaGetLastError();
}

// d_gu (same packed layout): a = bf16(act(g)); du = bf16(dact * a); da = bf16(dact * u); dg = bf16(da * act'(g))
__global__ void __launch_bounds__(256)
glu_bwd_kernel(const __nv_bfloat16* __restrict__ gu, long long ldg, const __nv_bfloat16* __restrict__ dact,
               long long lda, __nv_bfloat16* __restrict__ dgu, long long ldd, int rows, int d_ff, int act) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int per_row = d_ff >> 3;
  if (i >= (long long)rows * per_row) return;
  const int r = static_cast<int>(i / per_row), c = static_cast<int>(i % per_row) * 8;
  const int blk = c >> 7, j = c & 127;
  const __nv_bfloat16* src = gu + (long long)r * ldg + blk * 256 + j;
  __nv_bfloat16* dst = dgu + (long long)r * ldd + blk * 256 + j;
  float g[8], u[8], da[8], og[8], ou[8];
  unpack8(*reinterpret_cast<const uint4*>(src), g);
  unpack8(*reinterpret_cast<const uint4*>(src + 128), u);
  unpack8(*reinterpret_cast<const uint4*>(dact + (long long)r * lda + c), da);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    ou[k] = da[k] * bf16_round(glu_act(g[k], act));
    og[k] = bf16_round(da[k] * u[k]) * glu_act_grad(g[k], act);
  }
  *reinterpret_cast<uint4*>(dst) = pack8(og);
  *reinterpret_cast<uint4*>(dst + 128) = pack8(ou);
}

cudaError_t launch_glu_bwd(const __nv_bfloat16* gu, long long ldg, const __nv_bfloat16* dact, long long lda,
                           __nv_bfloat16* dgu, long long ldd, int rows, int d_ff, int act, cudaStream_t st) {
  if (rows <= 0) return cudaSuccess;
  if ((d_ff & 127) || (ldg & 7) || (lda & 7) || (ldd & 7)) return cudaErrorInvalidValue;
  const long long n = (long long)rows * (d_ff >> 3);
  glu_bwd_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, st>>>(gu, ldg, dact, lda, dgu, ldd, rows, d_ff, act);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Pointwise activations of the vision side: kind 0 quick_gelu x * sigmoid(1.702 x) (CLIP MLP),
// kind 1 erf-GELU (LLaVA projector). fwd: y = bf16(f(x)); bwd: dx = bf16(dy * f'(x)), in place allowed.
__device__ __forceinline__ float act_f(float x, int kind) {
  if (kind == 0) return x / (1.0f + __expf(-1.702f * x));
  return 0.5f * x * (1.0f + erff(x * 0.7071067811865476f));
}
__device__ __forceinline__ float act_df(float x, int kind) {
  if (kind == 0) {
    const float s = 1.0f / (1.0f + __expf(-1.702f * x));
    return s * (1.0f + 1.702f * x * (1.0f - s));
  }
  return 0.5f * (1.0f + erff(x * 0.7071067811865476f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}

__global__ void __launch_bounds__(256)
act_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ out,
           long long n8, int kind) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  float a[8], o[8];
  unpack8(reinterpret_cast<const uint4*>(x)[i], a);
  if (dy == nullptr) {
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = act_f(a[k], kind);
  } else {
    float g[8];
    unpack8(reinterpret_cast<const uint4*>(dy)[i], g);
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = g[k] * act_df(a[k], kind);
  }
  reinterpret_cast<uint4*>(out)[i] = pack8(o);
}

cudaError_t launch_act(const __nv_bfloat16* x, const __nv_bfloat16* dy, __nv_bfloat16* out, long long n, int kind,
                       cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  if (n & 7) return cudaErrorInvalidValue;
  act_kernel<<<static_cast<int>((n / 8 + 255) / 256), 256, 0, st>>>(x, dy, out, n / 8, kind);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Cross-entropy over the T target rows (reference :1010-1012, F.cross_entropy mean): one block per
// row. dlogits[t, :] = bf16((softmax(logits[t, :]) - onehot(target[t])) / T); loss_rows[t] = lse - logit[target].
__global__ void __launch_bounds__(1024)
ce_grad_kernel(const __nv_bfloat16* __restrict__ logits, long long ldl, const long long* __restrict__ target,
               __nv_bfloat16* __restrict__ dlogits, long long ldd, float* __restrict__ loss_rows, int T, int V) {
  __shared__ float red[32];
  __shared__ float bcast;
  const int t = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const __nv_bfloat16* row = logits + (long long)t * ldl;
  float m = -INFINITY;
  for (int j = tid; j < V; j += blockDim.x) m = fmaxf(m, ldbf(row + j));
  m = warp_max(m);
  if (lane == 0) red[warp] = m;
  __syncthreads();
  if (tid == 0) {
    float v = red[0];
    for (int i = 1; i < nw; ++i) v = fmaxf(v, red[i]);
    bcast = v;
  }
  __syncthreads();
  m = bcast;
  float s = 0.0f;
  for (int j = tid; j < V; j += blockDim.x) s += __expf(ldbf(row + j) - m);
  s = warp_sum(s);
  __syncthreads();
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (tid == 0) {
    float v = 0.0f;
    for (int i = 0; i < nw; ++i) v += red[i];
    bcast = v;
  }
  __syncthreads();
  s = bcast;
  const int tgt = static_cast<int>(target[t]);
  const float inv = 1.0f / s, invT = 1.0f / static_cast<float>(T);
  __nv_bfloat16* out = dlogits + (long long)t * ldd;
  for (int j = tid; j < V; j += blockDim.x) {
    const float p = __expf(ldbf(row + j) - m) * inv;
    out[j] = __float2bfloat16_rn((p - (j == tgt ? 1.0f : 0.0f)) * invT);
  }
  if (tid == 0 && loss_rows != nullptr) loss_rows[t] = (m + logf(s)) - ldbf(row + tgt);
}

cudaError_t launch_ce_grad(const __nv_bfloat16* logits, long long ldl, const long long* target, __nv_bfloat16* dlogits,
                           long long ldd, float* loss_rows, int T, int V, cudaStream_t st) {
  if (T <= 0 || V <= 0) return cudaSuccess;
  ce_grad_kernel<<<T, 1024, 0, st>>>(logits, ldl, target, dlogits, ldd, loss_rows, T, V);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Patch embedding as a GEMM (CLIP conv with stride == kernel, no bias): im2col of the fp32 image with
// the torchvision Normalize folded in, and its adjoint (input gradient back to fp32 pixels).
//   patches[(gy * gw + gx), c * p * p + py * p + px] = bf16((img[c, gy p + py, gx p + px] - mean[c]) / std[c])
__global__ void __launch_bounds__(256)
patchify_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ patches, long long ldp, int C, int H, int W,
                int p, float m0, float m1, float m2, float s0, float s1, float s2) {
  const int gw = W / p, gh = H / p, K = C * p * p;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)gh * gw * K) return;
  const int k = static_cast<int>(i % K), pt = static_cast<int>(i / K);
  const int c = k / (p * p), py = (k / p) % p, px = k % p;
  const int gy = pt / gw, gx = pt % gw;
  const float mean = c == 0 ? m0 : c == 1 ? m1 : m2, sd = c == 0 ? s0 : c == 1 ? s1 : s2;
  const float v = (img[((long long)c * H + gy * p + py) * W + gx * p + px] - mean) / sd;
  patches[(long long)pt * ldp + k] = __float2bfloat16_rn(v);
}

cudaError_t launch_patchify(const float* img, __nv_bfloat16* patches, long long ldp, int C, int H, int W, int p,
                            const float* mean, const float* sd, cudaStream_t st) {
  if (C != 3 || p <= 0) return cudaErrorInvalidValue;
  const long long n = (long long)(H / p) * (W / p) * C * p * p;
  patchify_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, st>>>(img, patches, ldp, C, H, W, p, mean[0], mean[1],
                                                                     mean[2], sd[0], sd[1], sd[2]);
  return cudaGetLastError();
}

// grad_img[c, y, x] = float(dpatches[pt, k]) / std[c]  (pixels outside the patch grid get 0)
__global__ void __launch_bounds__(256)
unpatchify_grad_kernel(const __nv_bfloat16* __restrict__ dpatches, long long ldp, float* __restrict__ gimg, int C,
                       int H, int W, int p, float s0, float s1, float s2) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)C * H * W) return;
  const int x = static_cast<int>(i % W), y = static_cast<int>((i / W) % H), c = static_cast<int>(i / ((long long)W * H));
  const int gw = W / p, gh = H / p;
  const int gy = y / p, gx = x / p;
  float v = 0.0f;
  if (gy < gh && gx < gw) {
    const int k = c * p * p + (y - gy * p) * p + (x - gx * p);
    v = ldbf(dpatches + (long long)(gy * gw + gx) * ldp + k) / (c == 0 ? s0 : c == 1 ? s1 : s2);
  }
  gimg[i] = v;
}

cudaError_t launch_unpatchify_grad(const __nv_bfloat16* dpatches, long long ldp, float* gimg, int C, int H, int W,
                                   int p, const float* sd, cudaStream_t st) {
  if (C != 3 || p <= 0) return cudaErrorInvalidValue;
  const long long n = (long long)C * H * W;
  unpatchify_grad_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, st>>>(dpatches, ldp, gimg, C, H, W, p, sd[0],
                                                                            sd[1], sd[2]);
  return cudaGetLastError();
}

// out[r, :] = bf16(a[r, :] + b[r % b_rows, :])   (class/position embeddings, bias rows, gradient sums)
__global__ void __launch_bounds__(256)
add_rows_kernel(const __nv_bfloat16* __restrict__ a, long long lda, const __nv_bfloat16* __restrict__ b, long long ldb,
                int b_rows, __nv_bfloat16* __restrict__ out, long long ldo, int rows, int d) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int per_row = d >> 3;
  if (i >= (long long)rows * per_row) return;
  const int r = static_cast<int>(i / per_row), c = static_cast<int>(i % per_row) * 8;
  float x[8], y[8], o[8];
  unpack8(*reinterpret_cast<const uint4*>(a + (long long)r * lda + c), x);
  unpack8(*reinterpret_cast<const uint4*>(b + (long long)(r % b_rows) * ldb + c), y);
#pragma unroll
  for (int k = 0; k < 8; ++k) o[k] = x[k] + y[k];
  *reinterpret_cast<uint4*>(out + (long long)r * ldo + c) = pack8(o);
}

cudaError_t launch_add_rows(const __nv_bfloat16* a, long long lda, const __nv_bfloat16* b, long long ldb, int b_rows,
                            __nv_bfloat16* out, long long ldo, int rows, int d, cudaStream_t st) {
  if (rows <= 0) return cudaSuccess;
  if ((d & 7) || (lda & 7) || (ldb & 7) || (ldo & 7) || b_rows <= 0) return cudaErrorInvalidValue;
  const long long n = (long long)rows * (d >> 3);
  add_rows_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, st>>>(a, lda, b, ldb, b_rows, out, ldo, rows, d);
  return cudaGetLastError();
}

}  // namespace bma
