// Shared-prefix + causal-tail attention for head_dim 128, "split-key" variant: the two softmax
// threads of a query row never talk to each other inside a work item.
//
// Same problem, tiling, masks and TMA layouts as attention_tc.cu (which stays the kernel for
// head_dim 256, where TMEM has no room for a second accumulator). In attention_tc.cu the two
// threads of a row (low / high half of a block's 128 key columns) must agree on one running
// maximum, i.e. swap their partial maxima through smem behind a named barrier in EVERY key block --
// ncu showed that barrier and its smem round trip as the top stall. Here each half keeps its own
// running (max, sum) and its own accumulator in TMEM:
//     O_lo += P[:, 0:64] V[0:64, :]      O_hi += P[:, 64:128] V[64:128, :]      (two K = 64 MMAs)
// so a block needs no communication at all; the halves are merged once per work item,
//     out = (O_lo f_lo + O_hi f_hi) / (l_lo f_lo + l_hi f_hi),   f_h = 2^(m_h - max(m_lo, m_hi)),
// each thread producing 64 of the 128 output columns. TMEM: S_0 S_1 O_lo O_hi = 4 x 128 columns.
//
// P never touches shared memory: a thread packs its 64 probabilities to bf16 and writes them with
// tcgen05.st over the S columns it has just read (P_lo at columns 0..31, P_hi at 64..95 of S_st), and
// the PV MMAs take their A operand from tensor memory (tcgen05.mma [d], [a_tmem], b_desc). With both
// operands in smem a key block costs 224 KB of smem traffic (TMA 64 + QK 64 + PV 64 + P 32) against
// 128 B/clk -- 1750 cycles, more than the 1024 cycles of MMA or MUFU work; this way it is 160 KB.
// The S buffer needs no "empty" barrier any more: S_{g+2} is issued after the PVs of block g, the
// tensor pipe executes in order, and those PVs wait for the softmax threads that read S_g. What
// keeps the four warps of a half within one block of each other is the p_empty wait before P is
// written (see the softmax loop).
//   warp 0  TMA producer: Q (double-buffered) per item; K and V^T blocks in separate 2-stage rings
//   warp 1  MMA issuer, one flat stream over items: S_g before the PVs of block g-1
//   warp 2  TMEM allocator
//   warps 4-7 / 8-11  softmax, low / high key half (warp w and w+4 share a TMEM lane quadrant)
#include <cstdlib>

#include "common.cuh"
#include "gemm_host.cuh"
#include "kernels.h"

namespace bma {

namespace spl {

constexpr int BM = 128, BN = 128, HD = 128, THREADS = 384, CH = 64;
constexpr int PANEL = 128 * 128;              // [128 rows][64 bf16], 128-byte swizzle
constexpr int TILE = 2 * PANEL;               // 128 x 128 bf16
constexpr int SMEM = 2 * TILE /*Q*/ + 2 * TILE /*K*/ + 2 * TILE /*V^T*/ + 256 /*barriers*/ +
                     2 * BM * 8 /*(max, sum) swap*/;

struct Params {
  __nv_bfloat16* out; long long ldo;
  int M, S, P;
  const int* M_dev;
  const int2* row_meta;
  int P_shared;
  int pre_rows;              // leading rows that are shared-prefix positions (fused prefill): they see fewer keys
  int n_heads, kv_group;
  float scale_log2;
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
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
// keys of the block that fall into half h (0: columns 0..63, 1: columns 64..127)
__device__ __forceinline__ int keys_of(int n_mma, int h) { return h == 0 ? min(n_mma, CH) : max(n_mma - CH, 0); }

__global__ void __launch_bounds__(THREADS, 1)
attention_split_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kt,
                       const __grid_constant__ CUtensorMap tm_kp, const __grid_constant__ CUtensorMap tm_vt,
                       const __grid_constant__ CUtensorMap tm_vp, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sQ = smem;                          // [2][tile]
  uint8_t* sK = sQ + 2 * TILE;                 // [2][tile]
  uint8_t* sV = sK + 2 * TILE;                 // [2][tile]   V^T: [128 dims][keys]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * TILE);
  uint64_t* q_full = bars;            // [2]
  uint64_t* q_empty = bars + 2;       // [2]
  uint64_t* k_full = bars + 4;        // [2]
  uint64_t* k_empty = bars + 6;       // [2]
  uint64_t* v_full = bars + 8;        // [2]
  uint64_t* v_empty = bars + 10;      // [2]
  uint64_t* s_full = bars + 12;       // [2]
  uint64_t* p_full = bars + 16;       // [2]  per key half: its 4 warps
  uint64_t* p_empty = bars + 18;      // [2]  per key half
  uint64_t* o_empty = bars + 20;      // [1]  8 warps: the item's output has left TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);
  float2* swap = reinterpret_cast<float2*>(bars + 32);     // [2 halves][128 rows] (max, sum)

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
    mbar_init(o_empty, 8);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);  mbar_init(&q_empty[i], 1);
      mbar_init(&k_full[i], 1);  mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);  mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);  mbar_init(&p_empty[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base;             // S_st at + 128 st
  const uint32_t tmem_o = tmem_base + 256;       // O_h  at + 128 h

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------------ TMA producer
      uint32_t it = 0, item_n = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_n) {
        const int head = item / m_tiles, tile = item - head * m_tiles;
        const int kvh = head / p.kv_group;
        const int r0 = tile * BM, r1 = min(M, r0 + BM);
        const int kt0 = tail_start(r0, p.S, p.row_meta);
        const int nb = n_pre + (r1 - kt0 + BN - 1) / BN;
        {
          const int qb = item_n & 1;
          mbar_wait(&q_empty[qb], ((item_n >> 1) & 1) ^ 1, 0x500 + qb);
          mbar_arrive_expect_tx(&q_full[qb], TILE);
          tma_load_2d(sQ + qb * TILE, &tm_q, &q_full[qb], head * HD, r0);
          tma_load_2d(sQ + qb * TILE + PANEL, &tm_q, &q_full[qb], head * HD + 64, r0);
        }
        auto load_k = [&](int j) {
          const Block b = block_of(j, n_pre, p.P, kt0, r1);
          const uint32_t g = it + j;
          const int st = g & 1;
          mbar_wait(&k_empty[st], ((g >> 1) & 1) ^ 1, 0x510 + st);
          mbar_arrive_expect_tx(&k_full[st], TILE);
          const CUtensorMap* mk = b.pre ? &tm_kp : &tm_kt;
          tma_load_2d(sK + st * TILE, mk, &k_full[st], kvh * HD, b.key0);
          tma_load_2d(sK + st * TILE + PANEL, mk, &k_full[st], kvh * HD + 64, b.key0);
        };
        auto load_v = [&](int j) {
          const Block b = block_of(j, n_pre, p.P, kt0, r1);
          const uint32_t g = it + j;
          const int st = g & 1;
          mbar_wait(&v_empty[st], ((g >> 1) & 1) ^ 1, 0x520 + st);
          mbar_arrive_expect_tx(&v_full[st], TILE);
          const CUtensorMap* mv = b.pre ? &tm_vp : &tm_vt;
          tma_load_2d(sV + st * TILE, mv, &v_full[st], b.key0, kvh * HD);
          tma_load_2d(sV + st * TILE + PANEL, mv, &v_full[st], b.key0 + 64, kvh * HD);
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
      uint32_t cnt[2] = {0, 0};                  // PVs issued so far per key half (barrier phases)
      // O_h (+)= P[:, half h] V[half h, :] for block g; `first` bit h: this PV overwrites O_h
      auto issue_pv = [&](uint32_t g, int n_mma, int first, uint32_t item_par) {
        const int st = g & 1;
        mbar_wait(&v_full[st], (g >> 1) & 1, 0x660 + st);
        constexpr uint32_t idesc = umma_idesc_bf16(BM, HD);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int keys = keys_of(n_mma, h);
          if (keys == 0) continue;
          mbar_wait(&p_full[h], cnt[h] & 1, 0x620 + h);
          const bool fst = (first >> h) & 1;
          if (fst) mbar_wait(o_empty, item_par ^ 1, 0x630);      // the previous item's output was read out
          tc_fence_after();
          const uint32_t pa = tmem_s + st * 128 + h * CH;       // P_h: bf16 pairs, 8 columns per 16 keys
          const uint32_t vb = smem_u32(sV + st * TILE + h * PANEL);
          for (int kk = 0; kk < keys / 16; ++kk)
            umma_bf16_ts(tmem_o + h * 128, pa + kk * 8, umma_desc_k_sw128(vb + kk * 32), idesc, !(fst && kk == 0));
          umma_commit(&p_empty[h]);
          ++cnt[h];
        }
        umma_commit(&v_empty[st]);
      };
      bool have_pending = false;
      uint32_t pend_g = 0, pend_item_par = 0;
      int pend_n = 0, pend_first = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_n) {
        const int head = item / m_tiles, tile = item - head * m_tiles;
        const int r0 = tile * BM, r1 = min(M, r0 + BM);
        const int kt0 = tail_start(r0, p.S, p.row_meta);
        const int nb = n_pre + (r1 - kt0 + BN - 1) / BN;
        const int qb = item_n & 1;
        mbar_wait(&q_full[qb], (item_n >> 1) & 1, 0x600 + qb);
        int started = 0;                         // bit h: O_h already has a PV in this item
        for (int j = 0; j < nb; ++j, ++it) {
          const Block b = block_of(j, n_pre, p.P, kt0, r1);
          const int st = it & 1;
          const uint32_t par = (it >> 1) & 1;
          mbar_wait(&k_full[st], par, 0x640 + st);          // (S_st is free: the PVs of block it-2 precede this MMA)
          tc_fence_after();
          const uint32_t idesc = umma_idesc_bf16(BM, b.n_mma);
          const uint32_t qa = smem_u32(sQ + qb * TILE), kb = smem_u32(sK + st * TILE);
#pragma unroll
          for (int ks = 0; ks < HD / 16; ++ks) {
            const uint32_t off = (ks >> 2) * PANEL + (ks & 3) * 32;
            umma_bf16_ss(tmem_s + st * 128, umma_desc_k_sw128(qa + off), umma_desc_k_sw128(kb + off), idesc, ks != 0);
          }
          umma_commit(&s_full[st]);
          umma_commit(&k_empty[st]);
          if (j == nb - 1) umma_commit(&q_empty[qb]);
          if (have_pending) issue_pv(pend_g, pend_n, pend_first, pend_item_par);
          have_pending = true;
          const int part = (keys_of(b.n_mma, 0) > 0 ? 1 : 0) | (keys_of(b.n_mma, 1) > 0 ? 2 : 0);
          pend_g = it; pend_n = b.n_mma; pend_first = part & ~started; pend_item_par = item_n & 1;
          started |= part;
        }
      }
      if (have_pending) issue_pv(pend_g, pend_n, pend_first, pend_item_par);
    }
  } else if (warp >= 4) {
    // -------------------------------------------------------------------- softmax
    const int sub = warp & 3;
    const int h = (warp - 4) >> 2;               // key half: columns [64 h, 64 h + 64) of every block
    const int row_in_tile = sub * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(sub * 32) << 16;
    const uint32_t to_mine = tmem_o + lane_off + h * 128;
    constexpr float kRaise = 8.0f;               // raise the reference max only beyond 2^8: P stays <= 256
    uint32_t it = 0, item_n = 0;
    uint32_t cnt_mine = 0, cnt_other = 0;        // per-half block counters, as in the MMA warp (scalars:
                                                 // an array indexed by h would live in local memory)
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_n) {
      const int head = item / m_tiles, tile = item - head * m_tiles;
      const int r0 = tile * BM, r1 = min(M, r0 + BM);
      const int kt0 = tail_start(r0, p.S, p.row_meta);
      const int nb = n_pre + (r1 - kt0 + BN - 1) / BN;
      const int rq = r0 + row_in_tile;
      int cs = (rq / p.S) * p.S, plim = p.P;
      if (p.row_meta != nullptr && rq < M) {
        const int2 rm = p.row_meta[rq];
        cs = rm.x; plim = rm.y;
      }
      float m_ref = -INFINITY, l_run = 0.0f;
      int mine = 0;                              // blocks of this item in which this half took part
      bool other_used = false;                   // the other half's accumulator was written in this item

      for (int j = 0; j < nb; ++j, ++it) {
        const Block b = block_of(j, n_pre, p.P, kt0, r1);
        const int st = it & 1;
        const uint32_t par = (it >> 1) & 1;
        const int keys = keys_of(b.n_mma, h);
        const int other_keys = keys_of(b.n_mma, h ^ 1);
        if (other_keys > 0) { other_used = true; ++cnt_other; }
        mbar_wait(&s_full[st], par, 0x710 + st);
        tc_fence_after();
        float sv[CH];
#pragma unroll
        for (int c0 = 0; c0 < CH; c0 += 32) {
          uint32_t r[32];
          if (c0 < keys) {
            tmem_ld_32x32(tmem_s + lane_off + st * 128 + h * CH + c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) sv[c0 + i] = __uint_as_float(r[i]);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) sv[c0 + i] = -INFINITY;
          }
        }
        if (keys == 0) continue;                           // this half has no keys in the block
        const int kbase = b.key0 + h * CH;
        if (!b.pre || b.key0 + BN > (r0 < p.pre_rows ? 0 : p.P_shared)) {
#pragma unroll
          for (int i = 0; i < CH; ++i) {
            const int key = kbase + i;
            const bool vis = b.pre ? (key < plim) : (key >= cs && key <= rq && key < r1);
            if (!vis) sv[i] = -INFINITY;
          }
        }
        float raw = -INFINITY;
#pragma unroll
        for (int i = 0; i < CH; ++i) raw = fmaxf(raw, sv[i]);
        const float bm = raw * p.scale_log2;
        const bool raise = bm > m_ref + kRaise;
        const float m_new = raise ? bm : m_ref;
        const float m_use = (m_new == -INFINITY) ? 0.0f : m_new;
        const float factor = raise ? ex2(m_ref - m_use) : 1.0f;
        // Back-pressure: this half's previous PV has retired, i.e. ALL FOUR warps of the half had
        // delivered their part of the previous P. Without it a fast warp could arrive on p_full for two
        // consecutive blocks before a slow one has arrived for the first (the two S/P buffers would let
        // it), completing a phase with a row quarter of P missing and desynchronising the barrier.
        // (It also makes O_h quiescent for the rescale below.) Usually satisfied long before.
        mbar_wait(&p_empty[h], (cnt_mine & 1) ^ 1, 0x720 + h);
        const bool fix = raise && (m_ref != -INFINITY) && (mine > 0);
        if (__any_sync(0xffffffffu, fix)) {                // rare: rescale this row of O_h in TMEM
          tc_fence_after();
          const float f = fix ? factor : 1.0f;
#pragma unroll
          for (int c0 = 0; c0 < HD; c0 += 32) {
            uint32_t r[32];
            tmem_ld_32x32(to_mine + c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * f);
            tmem_st_32x32(to_mine + c0, r);
          }
          tmem_st_wait();
        }
        // P = exp2(s - m) as packed bf16 pairs over this half's own S columns (already in registers)
        float rs = 0.0f;
        uint32_t pk[32];
#pragma unroll
        for (int i = 0; i < CH; i += 2) {
          const float p0 = ex2(fmaf(sv[i], p.scale_log2, -m_use));         // -inf -> 0
          const float p1 = ex2(fmaf(sv[i + 1], p.scale_log2, -m_use));
          rs += p0 + p1;
          pk[i >> 1] = pack_bf16x2(p0, p1);
        }
        tmem_st_32x32(tmem_s + lane_off + st * 128 + h * CH, pk);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[h]);
        l_run = l_run * factor + rs;
        m_ref = m_new;
        ++mine;
        ++cnt_mine;
      }

      // ---- end of the item: swap (max, sum) with the row's other thread, merge 64 output columns
      {
        swap[h * BM + row_in_tile] = make_float2(m_ref, l_run);
        asm volatile("bar.sync %0, 64;" ::"r"(1 + sub) : "memory");
        const float2 other = swap[(h ^ 1) * BM + row_in_tile];
        const float m0 = h == 0 ? m_ref : other.x, l0 = h == 0 ? l_run : other.y;
        const float m1 = h == 0 ? other.x : m_ref, l1 = h == 0 ? other.y : l_run;
        const float m_all = fmaxf(m0, m1);
        const float f0 = l0 > 0.0f ? ex2(m0 - m_all) : 0.0f;
        const float f1 = l1 > 0.0f ? ex2(m1 - m_all) : 0.0f;
        const float l_all = l0 * f0 + l1 * f1;
        const float inv = l_all > 0.0f ? 1.0f / l_all : 0.0f;
        const float s0 = f0 * inv, s1 = f1 * inv;
        // the last PV of each accumulator that was written in this item has retired
        const bool used_mine = mine > 0;
        if (used_mine) mbar_wait(&p_empty[h], (cnt_mine - 1) & 1, 0x740);
        if (other_used) mbar_wait(&p_empty[h ^ 1], (cnt_other - 1) & 1, 0x741);
        tc_fence_after();
        const bool use0 = h == 0 ? used_mine : other_used, use1 = h == 0 ? other_used : used_mine;
        __nv_bfloat16* dst = p.out + (long long)rq * p.ldo + head * HD + h * 64;
        const uint32_t col0 = static_cast<uint32_t>(h * 64);
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 32) {
          float o[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = 0.0f;
          if (use0) {
            uint32_t a[32];
            tmem_ld_32x32(tmem_o + lane_off + col0 + c0, a);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __uint_as_float(a[i]) * s0;
          }
          if (use1) {
            uint32_t a[32];
            tmem_ld_32x32(tmem_o + lane_off + 128 + col0 + c0, a);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = fmaf(__uint_as_float(a[i]), s1, o[i]);
          }
          if (rq < M) {
#pragma unroll
            for (int v4 = 0; v4 < 4; ++v4) {
              reinterpret_cast<uint4*>(dst + c0)[v4] = make_uint4(
                  pack_bf16x2(o[v4 * 8 + 0], o[v4 * 8 + 1]), pack_bf16x2(o[v4 * 8 + 2], o[v4 * 8 + 3]),
                  pack_bf16x2(o[v4 * 8 + 4], o[v4 * 8 + 5]), pack_bf16x2(o[v4 * 8 + 6], o[v4 * 8 + 7]));
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_empty);
        asm volatile("bar.sync %0, 64;" ::"r"(1 + sub) : "memory");     // `swap` may be rewritten from here on
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace spl

cudaError_t launch_attention_split(const AttnParams& a, int n_heads, cudaStream_t stream) {
  using namespace spl;
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
    cudaError_t e = cudaFuncSetAttribute(attention_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  Params p{};
  p.out = a.out; p.ldo = a.ldo;
  p.M = a.M; p.S = a.S; p.P = a.P; p.M_dev = a.M_dev;
  p.row_meta = a.row_meta;
  p.P_shared = a.row_meta != nullptr ? a.P_shared : a.P;
  p.pre_rows = a.row_meta != nullptr ? a.pre_rows : 0;
  p.n_heads = n_heads; p.kv_group = a.kv_group;
  p.scale_log2 = a.scale_log2;
  watchdog_attach();
  const int m_tiles = (a.M + BM - 1) / BM;
  int grid = m_tiles * n_heads;
  if (grid > num_sms()) grid = num_sms();
  attention_split_kernel<<<grid, THREADS, SMEM, stream>>>(tq, tkt, tkp, tvt, tvp, p);
  return cudaGetLastError();
}

}  // namespace bma
