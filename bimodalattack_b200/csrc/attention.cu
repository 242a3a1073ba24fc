// Shared-prefix + causal-tail attention for batched candidate scoring.
//
// All search_width candidates share one prefix (image tokens + prompt) whose K/V were computed
// once; each candidate only owns a short tail of S rows. Query rows of all candidates are
// flattened to [B*S] and tiled 128 rows per CTA, so one CTA covers several candidates:
//   * prefix keys [0, P): visible to every query row, read once per CTA (L2 resident);
//   * tail keys: row rk is visible to query row rq iff both belong to the same candidate and
//     rk <= rq (block-diagonal causal mask over the flattened rows).
// The same kernel with P = 0, S = P_len computes the prefix itself (plain causal attention).
//
// Flash-attention style: online softmax in fp32, QK^T and PV on mma.sync m16n8k16 bf16 with
// ldmatrix fragments from XOR-swizzled shared memory, K/V blocks double-buffered with cp.async.
// (Attention is ~2.4% of the scoring FLOPs; the GEMMs that dominate are on tcgen05.)
#include "common.cuh"
#include "kernels.h"

namespace bma {

constexpr int ATT_BM = 128;      // query rows per CTA
constexpr int ATT_BN = 64;       // keys per block
constexpr int ATT_HD = 128;      // head dim
constexpr int ATT_THREADS = 256;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t d = smem_u32(smem_dst);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                        uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                          uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// element (row, 16-byte chunk) of a [rows][128] bf16 tile with XOR swizzle on the chunk index
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) {
  return static_cast<uint32_t>(row * 256 + ((chunk ^ (row & 7)) << 4));
}

__global__ void __launch_bounds__(ATT_THREADS, 1) prefix_tail_attention_kernel(const AttnParams p) {
  extern __shared__ __align__(128) uint8_t att_smem[];
  uint8_t* sQ = att_smem;                              // 128 x 256 B
  uint8_t* sK = sQ + ATT_BM * 256;                     // 2 x (64 x 256 B)
  uint8_t* sV = sK + 2 * ATT_BN * 256;                 // 2 x (64 x 256 B)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int head = blockIdx.y;
  const int kvh = head / p.kv_group;
  const int r0 = blockIdx.x * ATT_BM;
  const int r1 = min(p.M, r0 + ATT_BM);

  const int kt0 = (r0 / p.S) * p.S;                    // first tail key row any query here can see
  const int n_pre = (p.P + ATT_BN - 1) / ATT_BN;
  const int n_tail = (r1 - kt0 + ATT_BN - 1) / ATT_BN;
  const int n_blocks = n_pre + n_tail;

  // ---- async loads -------------------------------------------------------------------------
  auto load_q = [&]() {
    for (int c = tid; c < ATT_BM * 16; c += ATT_THREADS) {
      const int row = c >> 4, ch = c & 15;
      const int gr = r0 + row;
      const bool ok = gr < p.M;
      const __nv_bfloat16* src = p.q + (long long)(ok ? gr : 0) * p.ldq + head * ATT_HD + ch * 8;
      cp_async16(sQ + tile_off(row, ch), src, ok);
    }
  };
  auto load_kv = [&](int blk, int buf) {
    const bool pre = blk < n_pre;
    const int base = pre ? blk * ATT_BN : kt0 + (blk - n_pre) * ATT_BN;
    const int limit = pre ? p.P : r1;
    const __nv_bfloat16* kb = pre ? p.k_pre : p.k_tail;
    const __nv_bfloat16* vb = pre ? p.v_pre : p.v_tail;
    const long long ldk = pre ? p.ldkp : p.ldk;
    const long long ldv = pre ? p.ldvp : p.ldv;
    uint8_t* dk = sK + buf * (ATT_BN * 256);
    uint8_t* dv = sV + buf * (ATT_BN * 256);
    for (int c = tid; c < ATT_BN * 16; c += ATT_THREADS) {
      const int row = c >> 4, ch = c & 15;
      const int gr = base + row;
      const bool ok = gr < limit;
      const long long rr = ok ? gr : 0;
      cp_async16(dk + tile_off(row, ch), kb + rr * ldk + kvh * ATT_HD + ch * 8, ok);
      cp_async16(dv + tile_off(row, ch), vb + rr * ldv + kvh * ATT_HD + ch * 8, ok);
    }
  };

  load_q();
  load_kv(0, 0);
  cp_async_commit();

  // per-thread query rows and their visibility window in tail-row space
  const int rq0 = r0 + warp * 16 + g;
  const int rq1 = rq0 + 8;
  const int cs0 = (rq0 / p.S) * p.S;
  const int cs1 = (rq1 / p.S) * p.S;

  float o_acc[16][4];
#pragma unroll
  for (int i = 0; i < 16; ++i) o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.0f;
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.0f, 0.0f};
  uint32_t qf[8][4];

  for (int blk = 0; blk < n_blocks; ++blk) {
    const int buf = blk & 1;
    if (blk + 1 < n_blocks) {
      load_kv(blk + 1, buf ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();

    if (blk == 0) {
      // Q fragments: 16 rows x 128 dims per warp, kept in registers for the whole kernel
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const int row = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int ch = ks * 2 + (lane >> 4);
        ldsm_x4(smem_u32(sQ + tile_off(row, ch)), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
      }
    }

    const uint8_t* bk = sK + buf * (ATT_BN * 256);
    const uint8_t* bv = sV + buf * (ATT_BN * 256);

    // ---- S = Q K^T (16 x 64 per warp) ------------------------------------------------------
    float s_acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s_acc[i][0] = s_acc[i][1] = s_acc[i][2] = s_acc[i][3] = 0.0f;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {       // pairs of 8-key tiles
        uint32_t b0, b1, b2, b3;
        const int key = np * 16 + (lane & 7) + (lane >> 4) * 8;
        const int ch = ks * 2 + ((lane >> 3) & 1);
        ldsm_x4(smem_u32(bk + tile_off(key, ch)), b0, b1, b2, b3);
        mma_bf16(s_acc[np * 2], qf[ks], b0, b1);
        mma_bf16(s_acc[np * 2 + 1], qf[ks], b2, b3);
      }
    }

    // ---- mask + online softmax ---------------------------------------------------------------
    const bool pre = blk < n_pre;
    const int kbase = pre ? blk * ATT_BN : kt0 + (blk - n_pre) * ATT_BN;
    float bm[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = kbase + nt * 8 + t * 2 + (e & 1);
        const int hi = e >> 1;
        const int rq = hi ? rq1 : rq0;
        const int cs = hi ? cs1 : cs0;
        const bool vis = pre ? (key < p.P) : (key >= cs && key <= rq && key < r1);
        const float v = vis ? s_acc[nt][e] * p.scale_log2 : -INFINITY;
        s_acc[nt][e] = v;
        bm[hi] = fmaxf(bm[hi], v);
      }
    }
    float corr[2], m_use[2];
#pragma unroll
    for (int hi = 0; hi < 2; ++hi) {
      bm[hi] = fmaxf(bm[hi], __shfl_xor_sync(0xffffffffu, bm[hi], 1));
      bm[hi] = fmaxf(bm[hi], __shfl_xor_sync(0xffffffffu, bm[hi], 2));
      const float m_new = fmaxf(m_run[hi], bm[hi]);
      m_use[hi] = (m_new == -INFINITY) ? 0.0f : m_new;
      corr[hi] = exp2f(m_run[hi] - m_use[hi]);     // m_run = -inf -> 0
      m_run[hi] = m_new;
      l_run[hi] *= corr[hi];
    }
    float bs[2] = {0.0f, 0.0f};
    uint32_t pf[8][2];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p0 = exp2f(s_acc[nt][0] - m_use[0]);
      const float p1 = exp2f(s_acc[nt][1] - m_use[0]);
      const float p2 = exp2f(s_acc[nt][2] - m_use[1]);
      const float p3 = exp2f(s_acc[nt][3] - m_use[1]);
      bs[0] += p0 + p1;
      bs[1] += p2 + p3;
      pf[nt][0] = pack_bf16x2(p0, p1);
      pf[nt][1] = pack_bf16x2(p2, p3);
    }
    l_run[0] += bs[0];
    l_run[1] += bs[1];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      o_acc[i][0] *= corr[0];
      o_acc[i][1] *= corr[0];
      o_acc[i][2] *= corr[1];
      o_acc[i][3] *= corr[1];
    }

    // ---- O += P V ------------------------------------------------------------------------------
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {          // 16 keys per step
      const uint32_t a[4] = {pf[kk * 2][0], pf[kk * 2][1], pf[kk * 2 + 1][0], pf[kk * 2 + 1][1]};
#pragma unroll
      for (int dp = 0; dp < 8; ++dp) {        // pairs of 8-dim tiles
        uint32_t b0, b1, b2, b3;
        const int key = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int ch = dp * 2 + (lane >> 4);
        ldsm_x4_t(smem_u32(bv + tile_off(key, ch)), b0, b1, b2, b3);
        mma_bf16(o_acc[dp * 2], a, b0, b1);
        mma_bf16(o_acc[dp * 2 + 1], a, b2, b3);
      }
    }
    __syncthreads();   // everyone is done with this K/V buffer before it is refilled
  }

  // ---- normalise, stage through this warp's (now unused) Q rows, 16-byte coalesced stores ------
#pragma unroll
  for (int hi = 0; hi < 2; ++hi) {
    l_run[hi] += __shfl_xor_sync(0xffffffffu, l_run[hi], 1);
    l_run[hi] += __shfl_xor_sync(0xffffffffu, l_run[hi], 2);
  }
  const float inv0 = l_run[0] > 0.0f ? 1.0f / l_run[0] : 0.0f;
  const float inv1 = l_run[1] > 0.0f ? 1.0f / l_run[1] : 0.0f;
  __syncwarp();
#pragma unroll
  for (int nt = 0; nt < 16; ++nt) {
    const int row_lo = warp * 16 + g;
    const int ch = nt;                       // 8 dims = one 16-byte chunk
    uint32_t* d0 = reinterpret_cast<uint32_t*>(sQ + tile_off(row_lo, ch)) + t;
    uint32_t* d1 = reinterpret_cast<uint32_t*>(sQ + tile_off(row_lo + 8, ch)) + t;
    *d0 = pack_bf16x2(o_acc[nt][0] * inv0, o_acc[nt][1] * inv0);
    *d1 = pack_bf16x2(o_acc[nt][2] * inv1, o_acc[nt][3] * inv1);
  }
  __syncwarp();
  for (int c = lane; c < 16 * 16; c += 32) {
    const int row = warp * 16 + (c >> 4), ch = c & 15;
    const int gr = r0 + row;
    if (gr < p.M) {
      const uint4 v = *reinterpret_cast<const uint4*>(sQ + tile_off(row, ch));
      *reinterpret_cast<uint4*>(p.out + (long long)gr * p.ldo + head * ATT_HD + ch * 8) = v;
    }
  }
}

cudaError_t launch_prefix_tail_attention(const AttnParams& p, int n_heads, cudaStream_t stream) {
  if (p.M <= 0) return cudaSuccess;
  if (p.S <= 0 || p.P < 0 || p.kv_group <= 0) return cudaErrorInvalidValue;
  constexpr int smem = ATT_BM * 256 + 4 * ATT_BN * 256;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(prefix_tail_attention_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  dim3 grid((p.M + ATT_BM - 1) / ATT_BM, n_heads);
  prefix_tail_attention_kernel<<<grid, ATT_THREADS, smem, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace bma
