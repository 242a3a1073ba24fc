// Shared-prefix + causal-tail attention on tcgen05, persistent and warp-specialised.
//
// All search_width candidates share one prefix (image tokens + prompt) whose K/V were computed
// once; each candidate owns a short tail of S rows. Query rows of all candidates are flattened to
// [B*S] and tiled 128 rows per work item (one head), so one item covers several candidates:
//   * prefix keys [0, P): visible to every query row;
//   * tail keys: row rk is visible to query row rq iff both belong to the same candidate and
//     rk <= rq (block-diagonal causal mask over the flattened rows).
// With P = 0 and S = P_len the same kernel computes the prefix itself (plain causal attention).
//
// This file is the kernel for head_dim 256 (Gemma-3; 64-key blocks) and the A/B fallback for
// head_dim 128 (BMA_ATTN=tc); head_dim 128 normally runs attention_occ.cu, BMA_ATTN=split attention_split.cu (per-half
// accumulators, P in tensor memory).
//
//   warp 0      TMA producer: Q tile per item (double-buffered for head_dim 128: the next item's Q is
//               fetched while this one runs), then K blocks and V^T blocks into two 2-stage rings
//   warp 1      MMA issuer:   S_j = Q K_j^T  and  O += P_j V_j   (tcgen05.mma, operands in smem,
//                             accumulators in TMEM: 2 x S[128x128] + O[128x128])
//   warp 2      TMEM allocator
//   warps 4..11 softmax: TWO threads per query row (warps w and w+4 share a TMEM lane quadrant),
//               each pulls its half of the S_j row from TMEM into registers in
//               one pass, applies mask + online softmax and writes P_j (bf16) to smem in the UMMA
//               K-major 128B-swizzled layout. O stays in TMEM across all key blocks; the running
//               maximum is only raised when a block exceeds it by more than 2^8 (so P <= 256), and
//               only then is the O row rescaled in TMEM (tcgen05.ld / .st) -- rare after the first
//               block. O is read once per work item, normalised and stored.
// Both MMAs are K-major x K-major because the QKV projection writes V transposed ([dim][row]).
// Key blocks have 128 keys except the last prefix / tail block, which is rounded up to 16 (the
// UMMA N granularity), so the masked-out work stays small.
#include <cstdlib>

#include "common.cuh"
#include "gemm_host.cuh"
#include "kernels.h"

namespace bma {

constexpr int ATC_BM = 128;        // query rows per work item
constexpr int ATC_THREADS = 384;     // 4 control warps + 8 softmax warps (two threads per query row)

// Tile geometry for head dim HD and (at most) BN keys per block. Every operand is stored as
// panels of [rows][64 bf16] with the 128-byte swizzle; a tile is a run of such panels.
//   HD 128: BN 128 (LLaVA / Llama);  HD 256: BN 64 (Gemma-3), which fits the same 227 KB of smem
template <int HD, int BN>
struct AtcCfg {
  static constexpr int kQPanel = ATC_BM * 128;          // [128 rows][64 dims]
  static constexpr int kKPanel = BN * 128;              // [BN keys][64 dims]
  static constexpr int kVPanel = HD * 128;              // [HD dims][64 keys]   (V^T)
  static constexpr int kPPanel = ATC_BM * 128;          // [128 rows][64 keys]
  static constexpr int kQTile = (HD / 64) * kQPanel;
  static constexpr int kKTile = (HD / 64) * kKPanel;
  static constexpr int kVTile = (BN / 64) * kVPanel;
  static constexpr int kPTile = (BN / 64) * kPPanel;
  // head_dim 128: TWO Q tiles (the next item's Q is fetched while this item runs) and ONE P tile
  // (P_j is written late in the softmax of block j, long after PV_{j-1} has drained the tile);
  // head_dim 256: one Q tile, there is no room for a second one
  static constexpr int kQBuf = (HD <= 128) ? 2 : 1;
  static constexpr int kXchg = 2 * 2 * ATC_BM * 4;       // [2 rotating buffers][2 halves][128 rows] floats
  static constexpr int kSmem = kQBuf * kQTile + 2 * (kKTile + kVTile) + kPTile + 256 /*barriers*/ + kXchg;
  static constexpr int kTmemS = 0;                      // S_0 at +0, S_1 at +BN
  static constexpr int kTmemO = 2 * BN;                 // O: HD columns
  static_assert(2 * BN + HD <= 512, "TMEM budget");
  static_assert(HD % 64 == 0 && BN % 64 == 0 && HD <= 256, "panel geometry / UMMA N limit");
};

struct AtcParams {
  __nv_bfloat16* out; long long ldo;
  int M, S, P;
  const int* M_dev;
  const int2* row_meta;      // ragged rows: (first row of the row's candidate, visible prefix keys)
  int P_shared;              // prefix keys every row sees (== P without row_meta)
  int pre_rows;              // leading rows that are shared-prefix positions (fused prefill): they see fewer keys
  int no_tail;               // rows attend to the prefix keys only
  int n_heads, kv_group;
  int q_col0, k_col0;        // column of head 0 inside the Q / K tensor maps
  float scale_log2;
  int debug;                 // bring-up switches (env BMA_ATC_DEBUG): 1 no QK mma, 2 no PV mma,
                             // 4 no S tmem loads, 8 no P stores, 16 no final O tmem loads
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct BlockInfo { int pre; int key0; int nvalid; int n_mma; };


// First tail key row any query row of the tile starting at r0 can see (start of that row's
// candidate), rounded DOWN to a multiple of 8 rows: the V^T tiles are fetched with the key index
// as the innermost TMA coordinate, whose byte address has to stay 16-byte aligned (an unaligned
// coordinate faults with "illegal instruction"). The up-to-7 extra keys are masked like any other
// key that belongs to an earlier candidate.
__device__ __forceinline__ int tail_key_start(int r0, int S, const int2* row_meta) {
  return (row_meta != nullptr ? row_meta[r0].x : (r0 / S) * S) & ~7;
}

template <int BN>
__device__ __forceinline__ BlockInfo block_info(int j, int n_pre, int P, int kt0, int r1) {
  BlockInfo b;
  b.pre = j < n_pre;
  b.key0 = b.pre ? j * BN : kt0 + (j - n_pre) * BN;
  const int limit = b.pre ? P : r1;
  b.nvalid = min(BN, limit - b.key0);
  b.n_mma = (b.nvalid + 15) & ~15;
  return b;
}

template <int HD, int BN>
__global__ void __launch_bounds__(ATC_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kt,
                    const __grid_constant__ CUtensorMap tm_kp, const __grid_constant__ CUtensorMap tm_vt,
                    const __grid_constant__ CUtensorMap tm_vp, const AtcParams p) {
  // The swizzled tiles need 1024-byte alignment. The whole 227 KB budget is used, so there is no
  // slack for rounding the base up: the dynamic segment is declared 1024-aligned (it is the only
  // shared allocation of the kernel) and checked.
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  using Cfg = AtcCfg<HD, BN>;
  constexpr int QB = Cfg::kQBuf;
  uint8_t* sQ = smem;                          // [QB][tile]
  uint8_t* sK = sQ + QB * Cfg::kQTile;         // [2 stages][tile]
  uint8_t* sV = sK + 2 * Cfg::kKTile;          // [2 stages][tile]  (V^T: [HD dims][keys])
  uint8_t* sP = sV + 2 * Cfg::kVTile;          // [tile]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + Cfg::kPTile);
  uint64_t* q_full = bars + 22;       // [2]
  uint64_t* q_empty = bars + 24;      // [2]
  uint64_t* k_full = bars + 2;        // [2]  K and V^T blocks ride in SEPARATE rings: a K stage is
  uint64_t* k_empty = bars + 4;       // [2]  released as soon as S_j = Q K_j^T has been computed, a V
  uint64_t* v_full = bars + 16;       // [2]  stage when O += P_j V_j has, so the loads of block j+2
  uint64_t* v_empty = bars + 18;      // [2]  are not serialised behind the softmax of block j
  uint64_t* s_full = bars + 6;        // [2]
  uint64_t* s_empty = bars + 8;       // [2]
  uint64_t* p_full = bars + 10;       // [1]
  uint64_t* p_empty = bars + 12;      // [1]
  uint64_t* o_empty = bars + 14;      // [1]  softmax has read the finished O of the item
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
  float* xchg = reinterpret_cast<float*>(bars + 32);     // [2 rotating][2 halves][128 rows] row max / row sum swaps

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int M = p.M_dev != nullptr ? min(p.M, __ldg(p.M_dev)) : p.M;   // live row count (ragged pass)
  const int m_tiles = (M + ATC_BM - 1) / ATC_BM;
  const int n_items = m_tiles * p.n_heads;
  const int n_pre = (p.P + BN - 1) / BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_kt); tma_prefetch_desc(&tm_kp);
    tma_prefetch_desc(&tm_vt); tma_prefetch_desc(&tm_vp);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(o_empty, 8);
    mbar_init(p_full, 8); mbar_init(p_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);  mbar_init(&q_empty[i], 1);
      mbar_init(&k_full[i], 1);  mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);  mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);  mbar_init(&s_empty[i], 8);
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
  const uint32_t tmem_s = tmem_base + Cfg::kTmemS;   // S_0 at +0, S_1 at +BN
  const uint32_t tmem_o = tmem_base + Cfg::kTmemO;   // O: HD columns

  // Every role walks the same sequence of (item, block); `it` counts blocks globally so that ring
  // stage = it & 1 and barrier parity = (it >> 1) & 1 stay in step across work items.
  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0, item_n = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_n) {
        const int head = item / m_tiles, tile = item - head * m_tiles;
        const int kvh = head / p.kv_group;
        const int r0 = tile * ATC_BM, r1 = min(M, r0 + ATC_BM);
        const int kt0 = tail_key_start(r0, p.S, p.row_meta);
        const int nb = n_pre + (p.no_tail ? 0 : (r1 - kt0 + BN - 1) / BN);
        {
          const int qb = item_n % QB;
          const uint32_t qpar = (item_n / QB) & 1;
          mbar_wait(&q_empty[qb], qpar ^ 1, 0x500 + qb);
          mbar_arrive_expect_tx(&q_full[qb], Cfg::kQTile);
#pragma unroll
          for (int c = 0; c < HD / 64; ++c)
            tma_load_2d(sQ + qb * Cfg::kQTile + c * Cfg::kQPanel, &tm_q, &q_full[qb],
                        p.q_col0 + head * HD + c * 64, r0);
        }
        // issue order K_0, K_1, V_0, K_2, V_1, ...: K_{j+1} is requested before V_j, so every load
        // has about two softmax periods of slack before its MMA
        auto load_k = [&](int j) {
          const BlockInfo b = block_info<BN>(j, n_pre, p.P, kt0, r1);
          const uint32_t g = it + j;
          const int st = g & 1;
          mbar_wait(&k_empty[st], ((g >> 1) & 1) ^ 1, 0x510 + st);
          mbar_arrive_expect_tx(&k_full[st], Cfg::kKTile);
          const CUtensorMap* mk = b.pre ? &tm_kp : &tm_kt;
          uint8_t* dk = sK + st * Cfg::kKTile;
#pragma unroll
          for (int c = 0; c < HD / 64; ++c)
            tma_load_2d(dk + c * Cfg::kKPanel, mk, &k_full[st], p.k_col0 + kvh * HD + c * 64, b.key0);
        };
        auto load_v = [&](int j) {
          const BlockInfo b = block_info<BN>(j, n_pre, p.P, kt0, r1);
          const uint32_t g = it + j;
          const int st = g & 1;
          mbar_wait(&v_empty[st], ((g >> 1) & 1) ^ 1, 0x520 + st);
          mbar_arrive_expect_tx(&v_full[st], Cfg::kVTile);
          const CUtensorMap* mv = b.pre ? &tm_vp : &tm_vt;
          uint8_t* dv = sV + st * Cfg::kVTile;
#pragma unroll
          for (int c = 0; c < BN / 64; ++c)
            tma_load_2d(dv + c * Cfg::kVPanel, mv, &v_full[st], b.key0 + c * 64, kvh * HD);
        };
        load_k(0);
        for (int j = 0; j < nb; ++j) {
          if (j + 1 < nb) load_k(j + 1);
          load_v(j);
        }
        it += nb;
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t it = 0, item_n = 0;
      // O (+)= P_g V_g for the block with global index `g`; the first block of an item overwrites O
      auto issue_pv = [&](uint32_t g, int n_mma, bool first, uint32_t item_par) {
        const int st = g & 1;
        mbar_wait(&v_full[st], (g >> 1) & 1, 0x660 + st);
        mbar_wait(p_full, g & 1, 0x620);
        if (first) mbar_wait(o_empty, item_par ^ 1, 0x630);     // previous item's O has been read out
        tc_fence_after();
        const uint32_t idesc = umma_idesc_bf16(ATC_BM, HD);
        const uint32_t pa = smem_u32(sP), vb = smem_u32(sV + st * Cfg::kVTile);
        for (int kk = 0; kk < n_mma / 16 && !(p.debug & 2); ++kk) {
          const uint32_t off_p = (kk >> 2) * Cfg::kPPanel + (kk & 3) * 32;
          const uint32_t off_v = (kk >> 2) * Cfg::kVPanel + (kk & 3) * 32;
          umma_bf16_ss(tmem_o, umma_desc_k_sw128(pa + off_p), umma_desc_k_sw128(vb + off_v), idesc,
                       !(first && kk == 0));
        }
        umma_commit(&v_empty[st]);
        umma_commit(p_empty);
      };
      // One flat stream of blocks across work items: S_g is issued BEFORE the PV of block g-1, also
      // when g is the first block of the next item, so the softmax warps always find their next S
      // ready (the PV of an item's last block is issued after the next item's first QK).
      bool have_pending = false;
      uint32_t pend_g = 0, pend_item_par = 0;
      int pend_n = 0;
      bool pend_first = false;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_n) {
        const int head = item / m_tiles, tile = item - head * m_tiles;
        const int r0 = tile * ATC_BM, r1 = min(M, r0 + ATC_BM);
        const int kt0 = tail_key_start(r0, p.S, p.row_meta);
        const int nb = n_pre + (p.no_tail ? 0 : (r1 - kt0 + BN - 1) / BN);
        const int qb = item_n % QB;
        if (QB == 1 && have_pending) {                  // a single Q tile cannot be prefetched: do not hold
          issue_pv(pend_g, pend_n, pend_first, pend_item_par);   // the last PV back behind the Q reload
          have_pending = false;
        }
        mbar_wait(&q_full[qb], (item_n / QB) & 1, 0x600 + qb);
        for (int j = 0; j < nb; ++j, ++it) {
          const BlockInfo b = block_info<BN>(j, n_pre, p.P, kt0, r1);
          const int st = it & 1;
          const uint32_t par = (it >> 1) & 1;
          mbar_wait(&k_full[st], par, 0x640 + st);
          mbar_wait(&s_empty[st], par ^ 1, 0x650 + st);
          tc_fence_after();
          const uint32_t idesc = umma_idesc_bf16(ATC_BM, b.n_mma);
          const uint32_t qa = smem_u32(sQ + qb * Cfg::kQTile), kb = smem_u32(sK + st * Cfg::kKTile);
#pragma unroll
          for (int ks = 0; ks < HD / 16 && !(p.debug & 1); ++ks) {
            const uint32_t off_q = (ks >> 2) * Cfg::kQPanel + (ks & 3) * 32;
            const uint32_t off_k = (ks >> 2) * Cfg::kKPanel + (ks & 3) * 32;
            umma_bf16_ss(tmem_s + st * BN, umma_desc_k_sw128(qa + off_q), umma_desc_k_sw128(kb + off_k), idesc,
                         ks != 0);
          }
          umma_commit(&s_full[st]);
          umma_commit(&k_empty[st]);                    // K_j is free once S_j is complete
          if (j == nb - 1) umma_commit(&q_empty[qb]);   // Q tile no longer needed after this MMA
          if (have_pending) issue_pv(pend_g, pend_n, pend_first, pend_item_par);
          have_pending = true;
          pend_g = it; pend_n = b.n_mma; pend_first = (j == 0); pend_item_par = item_n & 1;
        }
      }
      if (have_pending) issue_pv(pend_g, pend_n, pend_first, pend_item_par);
    }
  } else if (warp >= 4) {
    // Two threads per query row: warps 4-7 take the low half of every block's key columns (and of
    // O's columns), warps 8-11 the high half; warp w may touch TMEM lanes 32*(w%4)..+31, so a
    // row's two threads sit in warps w and w+4. They exchange their partial row maxima through
    // smem (one named barrier per block) and then agree on every decision; the partial row sums
    // are only combined at the end of the item.
    constexpr int CH = BN / 2;                 // key columns per thread and block
    constexpr int OH = HD / 2;                 // O columns per thread
    const int sub = warp & 3;
    const int half = (warp - 4) >> 2;
    const int row_in_tile = sub * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(sub * 32) << 16;
    const uint32_t to = tmem_o + lane_off + half * OH;
    constexpr float kRaise = 8.0f;         // raise the reference max only beyond 2^8: P stays <= 256
    uint32_t it = 0, item_n = 0, xn = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_n) {
      const int head = item / m_tiles, tile = item - head * m_tiles;
      const int r0 = tile * ATC_BM, r1 = min(M, r0 + ATC_BM);
      const int kt0 = tail_key_start(r0, p.S, p.row_meta);
      const int nb = n_pre + (p.no_tail ? 0 : (r1 - kt0 + BN - 1) / BN);
      const int rq = r0 + row_in_tile;
      int cs = (rq / p.S) * p.S, plim = p.P;
      if (p.row_meta != nullptr && rq < M) {
        const int2 rm = p.row_meta[rq];
        cs = rm.x; plim = rm.y;
      }
      float m_ref = -INFINITY, l_run = 0.0f;

      for (int j = 0; j < nb; ++j, ++it) {
        const BlockInfo b = block_info<BN>(j, n_pre, p.P, kt0, r1);
        const int st = it & 1;
        const uint32_t par = (it >> 1) & 1;
        const uint32_t ts = tmem_s + lane_off + st * BN + half * CH;
        const int kbase = b.key0 + half * CH;              // key of this thread's column 0
        mbar_wait(&s_full[st], par, 0x710 + st);
        tc_fence_after();
        // ---- this thread's half of the S row into registers (columns >= n_mma were not computed)
        float sv[CH];
#pragma unroll
        for (int c0 = 0; c0 < CH; c0 += 32) {
          uint32_t r[32];
          if (half * CH + c0 < b.n_mma && !(p.debug & 4)) {
            tmem_ld_32x32(ts + c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) sv[c0 + i] = __uint_as_float(r[i]);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) sv[c0 + i] = -INFINITY;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[st]);          // S_j is in registers: the buffer is free
        if (!b.pre || b.key0 + BN > (r0 < p.pre_rows ? 0 : p.P_shared)) {      // masked blocks: prefix keys past P_shared and the tail
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
        // ---- row maximum over both halves: the two threads of a row swap their partial maxima
        //      through one of two rotating smem buffers (one named barrier per swap; a buffer is
        //      rewritten two swaps later, i.e. after a barrier that follows every read of it)
        {
          float* xb = xchg + (xn & 1) * 2 * ATC_BM;
          ++xn;
          xb[half * ATC_BM + row_in_tile] = raw;
          asm volatile("bar.sync %0, 64;" ::"r"(1 + sub) : "memory");   // only this row quadrant's two warps
          raw = fmaxf(raw, xb[(half ^ 1) * ATC_BM + row_in_tile]);
        }
        const float bm = raw * p.scale_log2;               // scale > 0; -inf stays -inf
        const bool raise = bm > m_ref + kRaise;            // also true for the first finite block
        const float m_new = raise ? bm : m_ref;
        const float m_use = (m_new == -INFINITY) ? 0.0f : m_new;
        const float factor = raise ? ex2(m_ref - m_use) : 1.0f;     // m_ref = -inf -> 0
        // ---- rare: rescale this thread's half of the O row in TMEM (others scale by 1)
        const bool fix = raise && (m_ref != -INFINITY) && (j > 0);
        if (__any_sync(0xffffffffu, fix)) {
          mbar_wait(p_empty, (it - 1) & 1, 0x730);        // PV_{j-1} must have retired
          tc_fence_after();
          const float f = fix ? factor : 1.0f;
#pragma unroll
          for (int c0 = 0; c0 < OH; c0 += 32) {
            uint32_t r[32];
            tmem_ld_32x32(to + c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * f);
            tmem_st_32x32(to + c0, r);
          }
          tmem_st_wait();
          tc_fence_before();
        }
        // ---- P = exp2(s - m) as bf16 into the swizzled K-major P tile; partial row sum
        mbar_wait(p_empty, (it & 1) ^ 1, 0x720);           // the P tile is free: PV of block it-1 retired
        uint8_t* prow = sP + row_in_tile * 128;
        float rs = 0.0f;
#pragma unroll
        for (int c0 = 0; c0 < CH; c0 += 32) {
          const int col = half * CH + c0;                  // column inside the block
          if (col < b.n_mma) {
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float p0 = ex2(fmaf(sv[c0 + i], p.scale_log2, -m_use));       // -inf -> 0
              const float p1 = ex2(fmaf(sv[c0 + i + 1], p.scale_log2, -m_use));
              rs += p0 + p1;
              pk[i >> 1] = pack_bf16x2(p0, p1);
            }
            uint8_t* panel = prow + (col >> 6) * Cfg::kPPanel;
            const int q0 = (col & 63) >> 3;               // first 16-byte chunk of this 32-key group
#pragma unroll
            for (int q = 0; q < 4 && !(p.debug & 8); ++q) {
              const int phys = (q0 + q) ^ (row_in_tile & 7);
              *reinterpret_cast<uint4*>(panel + phys * 16) =
                  make_uint4(pk[q * 4 + 0], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
            }
          }
        }
        fence_proxy_async_smem();                         // P visible to the tensor-core (async) proxy
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
        l_run = l_run * factor + rs;
        m_ref = m_new;
      }

      // ---- finished item: combine the two partial row sums, wait for the last PV, read this
      //      thread's half of O once, normalise, store
      {
        float* xb = xchg + (xn & 1) * 2 * ATC_BM;
        ++xn;
        xb[half * ATC_BM + row_in_tile] = l_run;
        asm volatile("bar.sync %0, 64;" ::"r"(1 + sub) : "memory");
        const float l_all = l_run + xb[(half ^ 1) * ATC_BM + row_in_tile];
        mbar_wait(p_empty, (it - 1) & 1, 0x740);          // the item's last PV has retired
        tc_fence_after();
        const float inv = l_all > 0.0f ? 1.0f / l_all : 0.0f;
        __nv_bfloat16* dst = p.out + (long long)rq * p.ldo + head * HD + half * OH;
#pragma unroll
        for (int c0 = 0; c0 < OH; c0 += 32) {
          uint32_t r[32];
          if (p.debug & 16) {
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = 0;
          } else {
            tmem_ld_32x32(to + c0, r);
            tmem_ld_wait();
          }
          if (rq < M) {
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              reinterpret_cast<uint4*>(dst + c0)[v] = make_uint4(
                  pack_bf16x2(__uint_as_float(r[v * 8 + 0]) * inv, __uint_as_float(r[v * 8 + 1]) * inv),
                  pack_bf16x2(__uint_as_float(r[v * 8 + 2]) * inv, __uint_as_float(r[v * 8 + 3]) * inv),
                  pack_bf16x2(__uint_as_float(r[v * 8 + 4]) * inv, __uint_as_float(r[v * 8 + 5]) * inv),
                  pack_bf16x2(__uint_as_float(r[v * 8 + 6]) * inv, __uint_as_float(r[v * 8 + 7]) * inv));
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
    tmem_dealloc(tmem_base, 512);
  }
}

template <int HD, int BN>
static cudaError_t launch_atc(const AttnParams& a, int n_heads, cudaStream_t stream) {
  using Cfg = AtcCfg<HD, BN>;
  const int kv_heads = n_heads / a.kv_group;
  CUtensorMap tq, tkt, tkp, tvt, tvp;
  if (!make_tmap_bf16(&tq, a.q, a.M, (long long)n_heads * HD, a.ldq, ATC_BM)) return cudaErrorInvalidValue;
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
    cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel<HD, BN>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  AtcParams p{};
  p.out = a.out; p.ldo = a.ldo;
  p.M = a.M; p.S = a.S > 0 ? a.S : 1; p.P = a.P; p.M_dev = a.M_dev;
  p.no_tail = a.no_tail;
  p.row_meta = a.row_meta;
  p.P_shared = a.row_meta != nullptr ? a.P_shared : a.P;
  p.pre_rows = a.row_meta != nullptr ? a.pre_rows : 0;
  p.n_heads = n_heads; p.kv_group = a.kv_group;
  p.q_col0 = 0; p.k_col0 = 0;
  p.scale_log2 = a.scale_log2;
  const char* dbg = getenv("BMA_ATC_DEBUG");
  p.debug = dbg != nullptr ? atoi(dbg) : 0;
  watchdog_attach();
  const int m_tiles = (a.M + ATC_BM - 1) / ATC_BM;
  int grid = m_tiles * n_heads;
  if (grid > num_sms()) grid = num_sms();
  attention_tc_kernel<HD, BN><<<grid, ATC_THREADS, Cfg::kSmem, stream>>>(tq, tkt, tkp, tvt, tvp, p);
  return cudaGetLastError();
}

cudaError_t launch_prefix_tail_attention(const AttnParams& a, int n_heads, cudaStream_t stream) {
  if (a.M <= 0) return cudaSuccess;
  if (a.P < 0 || a.kv_group <= 0) return cudaErrorInvalidValue;
  if (a.no_tail ? (a.P <= 0 || a.row_meta != nullptr) : a.S <= 0) return cudaErrorInvalidValue;
  if (a.head_dim == 128) {
    // default: two-CTAs-per-SM kernel (attention_occ.cu). BMA_ATTN=split selects the split-key kernel
    // (attention_split.cu), BMA_ATTN=tc this file's kernel (A/B runs).
    static int which = -1;
    if (which < 0) {
      const char* e = getenv("BMA_ATTN");
      which = e == nullptr ? 0 : (e[0] == 's' ? 1 : (e[0] == 't' ? 2 : 0));
    }
    if (which == 0 || a.no_tail) return launch_attention_occ(a, n_heads, stream);
    return which == 1 ? launch_attention_split(a, n_heads, stream) : launch_atc<128, 128>(a, n_heads, stream);
  }
  if (a.head_dim == 256) return launch_atc<256, 64>(a, n_heads, stream);
  if (a.head_dim == 64) return launch_atc<64, 128>(a, n_heads, stream);
  return cudaErrorInvalidValue;
}

}  // namespace bma
