// Shared-prefix + causal-tail attention for head_dim 128, "two CTAs per SM" variant.
//
// Same problem, masks and TMA layouts as attention_split.cu / attention_tc.cu. ncu showed the
// one-CTA-per-SM kernels latency-bound (tensor pipe 25 %, no unit saturated): per key block the chain
// S = Q K^T -> tcgen05.ld -> softmax -> tcgen05.st P -> O += P V is long and a single CTA has only one
// or two blocks in flight. Instead of a deeper hand-built pipeline this kernel is made SMALL enough
// for two CTAs to be resident per SM, so the hardware interleaves two independent chains:
//   * 64-key blocks (BN = 64): K tile 16 KB, V^T tile 16 KB, two stages each; ONE Q tile (32 KB)
//     -> 97 KB of shared memory per CTA;
//   * tensor memory: S_0 S_1 (64 columns each) + O (128 columns) = 256 of the 512 columns;
//   * 256 threads, <= 128 registers: ONE softmax thread per query row (4 warps); the row's 64 scores
//     are read from tensor memory twice (maximum, then exponentials) 32 columns at a time instead of
//     being held in registers; no exchange between threads at all;
//   * P never touches shared memory: bf16 pairs go back over the S columns with tcgen05.st and the
//     P V MMAs read their A operand from tensor memory (as in attention_split.cu).
//   warp 0  TMA producer   warp 1  MMA issuer (flat stream over work items; S_g is issued before the
//   PV of block g-1)   warp 2  TMEM allocator   warps 4-7  softmax + output
#include <cstdlib>

#include "common.cuh"
#include "gemm_host.cuh"
#include "kernels.h"

namespace bma {

namespace occ {

constexpr int BM = 128, BN = 64, HD = 128, THREADS = 256;
constexpr int QPANEL = BM * 128;              // [128 rows][64 dims] bf16, 128-byte swizzle
constexpr int QTILE = 2 * QPANEL;             // 32 KB
constexpr int KPANEL = BN * 128;              // [64 keys][64 dims]
constexpr int KTILE = 2 * KPANEL;             // 16 KB
constexpr int VTILE = HD * 128;               // [128 dims][64 keys]: one panel, 16 KB
constexpr int SMEM = QTILE + 2 * KTILE + 2 * VTILE + 256 /*barriers*/ + 1024 /*alignment slack*/;
constexpr int TMEM_COLS = 256;                // S_0 | S_1 | O

struct Params {
  __nv_bfloat16* out; long long ldo;
  int M, S, P;
  const int* M_dev;
  const int2* row_meta;
  int P_shared;
  int pre_rows;
  int no_tail;                // rows attend to the prefix keys only
  int n_heads, kv_group;
  float scale_log2;
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

struct Block { int pre; int key0; int n_mma; };

__device__ __forceinline__ int tail_start(int r0, int S, const int2* row_meta) {
  return (row_meta != nullptr ? row_meta[r0].x : (r0 / S) * S) & ~7;
}
__device__ __forceinline__ Block block_of(int j, int n_pre, int P, int kt0, int r1) {
  Block b;
  b.pre = j < n_pre;
  b.key0 = b.pre ? j * BN : kt0 + (j - n_pre) * BN;
  const int nvalid = min(BN, (b.pre ? P : r1) - b.key0);
  b.n_mma = (nvalid + 15) & ~15;              // keys the MMAs cover (UMMA N / K granularity 16)
  return b;
}

// One query row's pass over a 64-key block of S (tensor memory, 32 columns at a time). MASKED = false is
// the hot path (every key of the block is visible to every row of the warp: most prefix blocks) and does
// no per-element tests at all; MASKED = true keeps the row's visible column range [vlo, vhi).
template <bool MASKED>
__device__ __forceinline__ float block_row_max_impl(uint32_t ts, int n_mma, int vlo, int vhi) {
  float raw = -INFINITY;
#pragma unroll
  for (int c0 = 0; c0 < BN; c0 += 32) {
    if (!MASKED || c0 < n_mma) {
      uint32_t r[32];
      tmem_ld_32x32(ts + c0, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float v = __uint_as_float(r[i]);
        if (MASKED && !(c0 + i >= vlo && c0 + i < vhi)) v = -INFINITY;
        raw = fmaxf(raw, v);
      }
    }
  }
  return raw;
}
// P = exp2(s * scale - m) as packed bf16 pairs over the S columns just read (P_g occupies columns [0, 32)
// of S_st: chunk 0 -> columns 0..15, chunk 1 -> 16..31); returns the row sum.
template <bool MASKED>
__device__ __forceinline__ float block_row_exp_impl(uint32_t ts, int n_mma, int vlo, int vhi, float scale_log2, float m_use) {
  float rs = 0.0f;
#pragma unroll
  for (int c0 = 0; c0 < BN; c0 += 32) {
    if (!MASKED || c0 < n_mma) {
      uint32_t r[32], pk[16];
      tmem_ld_32x32(ts + c0, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        float v0 = __uint_as_float(r[i]), v1 = __uint_as_float(r[i + 1]);
        if (MASKED) {
          if (!(c0 + i >= vlo && c0 + i < vhi)) v0 = -INFINITY;
          if (!(c0 + i + 1 >= vlo && c0 + i + 1 < vhi)) v1 = -INFINITY;
        }
        const float p0 = ex2(fmaf(v0, scale_log2, -m_use));         // -inf -> 0
        const float p1 = ex2(fmaf(v1, scale_log2, -m_use));
        rs += p0 + p1;
        pk[i >> 1] = pack_bf16x2(p0, p1);
      }
      tmem_st_32x16(ts + (c0 >> 1), pk);
    }
  }
  return rs;
}

// The masked variants are real calls (rare path): inlined next to the hot path the compiler merges the
// two bodies into one predicated loop and every block pays for the per-element tests again.
__device__ __noinline__ float block_row_max_masked(uint32_t ts, int n_mma, int vlo, int vhi) {
  return block_row_max_impl<true>(ts, n_mma, vlo, vhi);
}
__device__ __noinline__ float block_row_exp_masked(uint32_t ts, int n_mma, int vlo, int vhi, float scale_log2, float m_use) {
  return block_row_exp_impl<true>(ts, n_mma, vlo, vhi, scale_log2, m_use);
}

__global__ void __launch_bounds__(THREADS, 2)
attention_occ_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kt,
                     const __grid_constant__ CUtensorMap tm_kp, const __grid_constant__ CUtensorMap tm_vt,
                     const __grid_constant__ CUtensorMap tm_vp, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;                          // [tile]
  uint8_t* sK = sQ + QTILE;                    // [2][tile]
  uint8_t* sV = sK + 2 * KTILE;                // [2][tile]   V^T: [128 dims][64 keys]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * VTILE);
  uint64_t* q_full = bars;            // [1]
  uint64_t* q_empty = bars + 1;       // [1]
  uint64_t* k_full = bars + 2;        // [2]
  uint64_t* k_empty = bars + 4;       // [2]
  uint64_t* v_full = bars + 6;        // [2]
  uint64_t* v_empty = bars + 8;       // [2]
  uint64_t* s_full = bars + 10;       // [2]
  uint64_t* p_full = bars + 12;       // [2]  per S/P buffer: the 4 softmax warps have written P_g
  uint64_t* p_empty = bars + 14;      // [2]  per S/P buffer: PV_g has retired
  uint64_t* o_empty = bars + 16;      // [1]  4 warps: the item's output has left TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int M = p.M_dev != nullptr ? min(p.M, __ldg(p.M_dev)) : p.M;   // live row count (ragged pass)
  const int m_tiles = (M + BM - 1) / BM;
  const int n_items = m_tiles * p.n_heads;
  const int n_pre = (p.P + BN - 1) / BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_kt); tma_prefetch_desc(&tm_kp);
    tma_prefetch_desc(&tm_vt); tma_prefetch_desc(&tm_vp);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);  mbar_init(q_empty, 1);
    mbar_init(o_empty, 4);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&p_full[i], 4);  mbar_init(&p_empty[i], 1);
      mbar_init(&k_full[i], 1);  mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);  mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base;             // S_st at + 64 st
  const uint32_t tmem_o = tmem_base + 128;       // O: 128 columns

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------------ TMA producer
      uint32_t it = 0, item_n = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_n) {
        const int head = item / m_tiles, tile = item - head * m_tiles;
        const int kvh = head / p.kv_group;
        const int r0 = tile * BM, r1 = min(M, r0 + BM);
        const int kt0 = tail_start(r0, p.S, p.row_meta);
        const int nb = n_pre + (p.no_tail ? 0 : (r1 - kt0 + BN - 1) / BN);
        mbar_wait(q_empty, (item_n & 1) ^ 1, 0x800);
        mbar_arrive_expect_tx(q_full, QTILE);
        tma_load_2d(sQ, &tm_q, q_full, head * HD, r0);
        tma_load_2d(sQ + QPANEL, &tm_q, q_full, head * HD + 64, r0);
        auto load_k = [&](int j) {
          const Block b = block_of(j, n_pre, p.P, kt0, r1);
          const uint32_t g = it + j;
          const int st = g & 1;
          mbar_wait(&k_empty[st], ((g >> 1) & 1) ^ 1, 0x810 + st);
          mbar_arrive_expect_tx(&k_full[st], KTILE);
          const CUtensorMap* mk = b.pre ? &tm_kp : &tm_kt;
          tma_load_2d(sK + st * KTILE, mk, &k_full[st], kvh * HD, b.key0);
          tma_load_2d(sK + st * KTILE + KPANEL, mk, &k_full[st], kvh * HD + 64, b.key0);
        };
        auto load_v = [&](int j) {
          const Block b = block_of(j, n_pre, p.P, kt0, r1);
          const uint32_t g = it + j;
          const int st = g & 1;
          mbar_wait(&v_empty[st], ((g >> 1) & 1) ^ 1, 0x820 + st);
          mbar_arrive_expect_tx(&v_full[st], VTILE);
          tma_load_2d(sV + st * VTILE, b.pre ? &tm_vp : &tm_vt, &v_full[st], b.key0, kvh * HD);
        };
        load_k(0);                               // K_{j+1} is requested before V_j
        for (int j = 0; j < nb; ++j) {
          if (j + 1 < nb) load_k(j + 1);
          load_v(j);
        }
        it += nb;
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------------------------------------------ MMA issuer
      uint32_t it = 0, item_n = 0;
      // O (+)= P_g V_g; `first`: this PV overwrites O (first block of its item)
      auto issue_pv = [&](uint32_t g, int n_mma, bool first, uint32_t item_par) {
        const int st = g & 1;
        mbar_wait(&v_full[st], (g >> 1) & 1, 0x860 + st);
        mbar_wait(&p_full[st], (g >> 1) & 1, 0x862);
        if (first) mbar_wait(o_empty, item_par ^ 1, 0x863);      // the previous item's output was read out
        tc_fence_after();
        constexpr uint32_t idesc = umma_idesc_bf16(BM, HD);
        const uint32_t pa = tmem_s + st * BN;                    // P_g: bf16 pairs, 8 columns per 16 keys
        const uint32_t vb = smem_u32(sV + st * VTILE);
        for (int kk = 0; kk < n_mma / 16; ++kk)
          umma_bf16_ts(tmem_o, pa + kk * 8, umma_desc_k_sw128(vb + kk * 32), idesc, !(first && kk == 0));
        umma_commit(&p_empty[st]);
        umma_commit(&v_empty[st]);
      };
      bool have_pending = false;
      uint32_t pend_g = 0, pend_item_par = 0;
      int pend_n = 0;
      bool pend_first = false;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_n) {
        const int head = item / m_tiles, tile = item - head * m_tiles;
        const int r0 = tile * BM, r1 = min(M, r0 + BM);
        const int kt0 = tail_start(r0, p.S, p.row_meta);
        const int nb = n_pre + (p.no_tail ? 0 : (r1 - kt0 + BN - 1) / BN);
        (void)head;
        mbar_wait(q_full, item_n & 1, 0x840);
        for (int j = 0; j < nb; ++j, ++it) {
          const Block b = block_of(j, n_pre, p.P, kt0, r1);
          const int st = it & 1;
          mbar_wait(&k_full[st], (it >> 1) & 1, 0x850 + st);     // (S_st is free: the PV of block it-2 precedes this MMA)
          tc_fence_after();
          const uint32_t idesc = umma_idesc_bf16(BM, b.n_mma);
          const uint32_t qa = smem_u32(sQ), kb = smem_u32(sK + st * KTILE);
#pragma unroll
          for (int ks = 0; ks < HD / 16; ++ks)
            umma_bf16_ss(tmem_s + st * BN, umma_desc_k_sw128(qa + (ks >> 2) * QPANEL + (ks & 3) * 32),
                         umma_desc_k_sw128(kb + (ks >> 2) * KPANEL + (ks & 3) * 32), idesc, ks != 0);
          umma_commit(&s_full[st]);
          umma_commit(&k_empty[st]);
          if (j == nb - 1) umma_commit(q_empty);
          if (have_pending) issue_pv(pend_g, pend_n, pend_first, pend_item_par);
          have_pending = true;
          pend_g = it; pend_n = b.n_mma; pend_first = (j == 0); pend_item_par = item_n & 1;
        }
      }
      if (have_pending) issue_pv(pend_g, pend_n, pend_first, pend_item_par);
    }
  } else if (warp >= 4) {
    // -------------------------------------------------------------------- softmax: one thread per query row
    const int sub = warp & 3;
    const int row_in_tile = sub * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(sub * 32) << 16;
    constexpr float kRaise = 8.0f;               // raise the reference max only beyond 2^8: P stays <= 256
    uint32_t it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int head = item / m_tiles, tile = item - head * m_tiles;
      const int r0 = tile * BM, r1 = min(M, r0 + BM);
      const int kt0 = tail_start(r0, p.S, p.row_meta);
      const int nb = n_pre + (p.no_tail ? 0 : (r1 - kt0 + BN - 1) / BN);
      const int rq = r0 + row_in_tile;
      int cs = (rq / p.S) * p.S, plim = p.P;
      if (p.row_meta != nullptr && rq < M) {
        const int2 rm = p.row_meta[rq];
        cs = rm.x; plim = rm.y;
      }
      const int p_shared = r0 < p.pre_rows ? 0 : p.P_shared;
      float m_ref = -INFINITY, l_run = 0.0f;

      for (int j = 0; j < nb; ++j, ++it) {
        const Block b = block_of(j, n_pre, p.P, kt0, r1);
        const int st = it & 1;
        const uint32_t ts = tmem_s + lane_off + st * BN;
        // this row's visible columns of the block: prefix keys below its limit, or its own candidate's
        // tail keys up to itself
        int vlo = 0, vhi = BN;
        if (b.pre) {
          if (b.key0 + BN > p_shared) vhi = min(max(plim - b.key0, 0), BN);
        } else {
          vlo = max(cs - b.key0, 0);
          vhi = min(max(min(rq + 1, r1) - b.key0, 0), BN);
        }
        const bool all_visible = __all_sync(0xffffffffu, vlo == 0 && vhi == BN);
        mbar_wait(&s_full[st], (it >> 1) & 1, 0x910 + st);
        tc_fence_after();
        // pass 1: row maximum over the visible keys
        const float raw = all_visible ? block_row_max_impl<false>(ts, b.n_mma, vlo, vhi)
                                      : block_row_max_masked(ts, b.n_mma, vlo, vhi);
        const float bm = raw * p.scale_log2;
        const bool raise = bm > m_ref + kRaise;
        const float m_new = raise ? bm : m_ref;
        const float m_use = (m_new == -INFINITY) ? 0.0f : m_new;
        const float factor = raise ? ex2(m_ref - m_use) : 1.0f;
        // No wait on the previous PV here: P_g has its own buffer and barrier (s_full of block g implies that
        // PV_{g-2}, the last reader of this buffer, has retired -- the tensor pipe completes in order), so the
        // softmax of block g overlaps PV_{g-1}. Only the rare rescale needs O quiescent.
        const bool fix = raise && (m_ref != -INFINITY) && (j > 0);
        if (__any_sync(0xffffffffu, fix)) {                // rare: rescale this row of O in TMEM
          mbar_wait(&p_empty[(it - 1) & 1], ((it - 1) >> 1) & 1, 0x920);    // PV_{g-1} has retired
          tc_fence_after();
          const float f = fix ? factor : 1.0f;
#pragma unroll
          for (int c0 = 0; c0 < HD; c0 += 32) {
            uint32_t r[32];
            tmem_ld_32x32(tmem_o + lane_off + c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * f);
            tmem_st_32x32(tmem_o + lane_off + c0, r);
          }
          tmem_st_wait();
        }
        // pass 2: exponentials, P back into tensor memory
        const float rs = all_visible ? block_row_exp_impl<false>(ts, b.n_mma, vlo, vhi, p.scale_log2, m_use)
                                     : block_row_exp_masked(ts, b.n_mma, vlo, vhi, p.scale_log2, m_use);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[st]);
        l_run = l_run * factor + rs;
        m_ref = m_new;
      }

      // ---- end of the item: normalise and store this thread's output row
      {
        mbar_wait(&p_empty[(it - 1) & 1], ((it - 1) >> 1) & 1, 0x940);   // the item's last PV has retired
        tc_fence_after();
        const float inv = l_run > 0.0f ? 1.0f / l_run : 0.0f;
        __nv_bfloat16* dst = p.out + (long long)rq * p.ldo + head * HD;
#pragma unroll
        for (int c0 = 0; c0 < HD; c0 += 32) {
          uint32_t a[32];
          tmem_ld_32x32(tmem_o + lane_off + c0, a);
          tmem_ld_wait();
          if (rq < M) {
#pragma unroll
            for (int v4 = 0; v4 < 4; ++v4) {
              reinterpret_cast<uint4*>(dst + c0)[v4] = make_uint4(
                  pack_bf16x2(__uint_as_float(a[v4 * 8 + 0]) * inv, __uint_as_float(a[v4 * 8 + 1]) * inv),
                  pack_bf16x2(__uint_as_float(a[v4 * 8 + 2]) * inv, __uint_as_float(a[v4 * 8 + 3]) * inv),
                  pack_bf16x2(__uint_as_float(a[v4 * 8 + 4]) * inv, __uint_as_float(a[v4 * 8 + 5]) * inv),
                  pack_bf16x2(__uint_as_float(a[v4 * 8 + 6]) * inv, __uint_as_float(a[v4 * 8 + 7]) * inv));
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_empty);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace occ

cudaError_t launch_attention_occ(const AttnParams& a, int n_heads, cudaStream_t stream) {
  using namespace occ;
  const int kv_heads = n_heads / a.kv_group;
  CUtensorMap tq, tkt, tkp, tvt, tvp;
  if (!make_tmap_bf16(&tq, a.q, a.M, (long long)n_heads * HD, a.ldq, BM)) return cudaErrorInvalidValue;
  if (!make_tmap_bf16(&tkt, a.k_tail, a.M, (long long)kv_heads * HD, a.ldk, BN)) return cudaErrorInvalidValue;
  if (!make_tmap_bf16(&tvt, a.vt_tail, (long long)kv_heads * HD, a.M, a.ldvt, HD)) return cudaErrorInvalidValue;
  if (a.P > 0) {
    if (!make_tmap_bf16(&tkp, a.k_pre, a.P, (long long)kv_heads * HD, a.ldkp, BN)) return cudaErrorInvalidValue;
    if (!make_tmap_bf16(&tvp, a.vt_pre, (long long)kv_heads * HD, a.P, a.ldvtp, HD)) return cudaErrorInvalidValue;
  } else {
    tkp = tkt;
    tvp = tvt;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attention_occ_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  Params p{};
  p.out = a.out; p.ldo = a.ldo;
  p.M = a.M; p.S = a.S > 0 ? a.S : 1; p.P = a.P; p.M_dev = a.M_dev;
  p.no_tail = a.no_tail;
  p.row_meta = a.row_meta;
  p.P_shared = a.row_meta != nullptr ? a.P_shared : a.P;
  p.pre_rows = a.row_meta != nullptr ? a.pre_rows : 0;
  p.n_heads = n_heads; p.kv_group = a.kv_group;
  p.scale_log2 = a.scale_log2;
  watchdog_attach();
  const int m_tiles = (a.M + BM - 1) / BM;
  long long grid = (long long)m_tiles * n_heads;
  if (grid > 2ll * num_sms()) grid = 2ll * num_sms();
  attention_occ_kernel<<<static_cast<int>(grid), THREADS, SMEM, stream>>>(tq, tkt, tkp, tvt, tvp, p);
  return cudaGetLastError();
}

}  // namespace bma
