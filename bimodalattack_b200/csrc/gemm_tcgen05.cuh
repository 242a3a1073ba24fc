// Persistent, warp-specialised bf16 GEMM for sm_100a:  C[M,N] = A[M,K] * B[N,K]^T  (both K-major).
//
//   warp 0        TMA producer   (cp.async.bulk.tensor, SWIZZLE_128B tiles, mbarrier ring)
//   warp 1        MMA issuer     (one elected lane issues tcgen05.mma, accumulators in TMEM)
//   warp 2        TMEM allocator
//   warps 4..7    epilogue       (tcgen05.ld -> registers -> fused epilogue -> global)
//
// Two TMEM accumulator buffers let the epilogue of tile i overlap the main loop of tile i+1.
// The epilogue is a template parameter; the variants are the fused ops of the candidate-scoring
// path:
//   EPI_STORE    bf16 store, optional residual add              (O-proj, down-proj)
//   EPI_ROPE     bf16 store with rotary embedding on Q,K heads  (fused QKV projection)
//   EPI_GLU      act(gate) * up, gate/up rows interleaved per tile (SwiGLU / GeGLU)
//   EPI_CE       per-row online log-sum-exp partials + target-logit pick (LM head; logits
//                are never written to memory)
//   EPI_STORE_T  transposed store  out[n, m]                    (token gradient, W streamed as A)
#pragma once
#include "common.cuh"

namespace bma {

enum : int { EPI_STORE = 0, EPI_ROPE = 1, EPI_GLU = 2, EPI_CE = 3, EPI_STORE_T = 4 };

struct GemmParams {
  int M, N, K;
  const int* M_dev;           // optional: the live row count on the device (<= M). M then only bounds
                              // the grid and the TMA maps, so a ragged pass needs no host round trip
  // Batched mode (batch > 1): `batch` independent M x N x K problems inside the SAME A / B matrices.
  // Problem b reads A at (row a_row * b, column a_col * b) and B at (b_row * b, b_col * b) -- heads
  // laid side by side along the columns, or stacked along the rows -- and writes C + c_batch * b.
  int batch;
  int a_row, a_col, b_row, b_col;
  long long c_batch;
  int k_stagger;              // > 0: tile t walks its K blocks starting at (t * k_stagger) % k_blocks, so CTAs that
                              // stream different rows of a power-of-two-pitch matrix do not all touch the same
                              // HBM channels at the same time (token gradient: 8 KB pitch, tiles 1 MB apart)
  int causal;                 // (single-CTA tiles) products of ONE causal sequence, rows = cols = positions:
                              //   1: skip tiles that lie entirely above the diagonal   (S = Q K^T, dP = dO V^T)
                              //   2: K runs over keys <= the tile's last row           (O = P V,  dQ = dS K)
                              //   3: K runs over queries >= the tile's first row       (dK = dS^T Q, dV = P^T dO)
  float* C32;                 // EPI_STORE: raw fp32 accumulators instead of bf16 (attention scores)
  const __nv_bfloat16* bias;  // EPI_STORE: [N] added in fp32 before the bf16 rounding (nn.Linear bias)
  int group_m;                // rasterisation: m-blocks per L2 supergroup
  // output
  __nv_bfloat16* C;           // EPI_STORE/ROPE/GLU/STORE_T
  long long ldc;
  const __nv_bfloat16* R;     // optional residual (EPI_STORE), same shape as C
  long long ldr;
  // rotary (EPI_ROPE): tables [pos][head_dim/2] fp32 (values already rounded to the model dtype)
  const float* rope_cos;
  const float* rope_sin;
  int rope_cols;              // leading output columns that are Q/K heads (multiple of 128)
  int rows_per_seq;           // position of row r is pos_base + (r % rows_per_seq), or
  int pos_base;               //   pos_base + row_info[r].y when row_info != null (ragged rows)
  const int2* row_info;
  __nv_bfloat16* Vt;          // EPI_ROPE: columns >= rope_cols (the V heads) go here TRANSPOSED,
  long long ldvt;             //   Vt[(col - rope_cols) * ldvt + row], for the K-major PV operand
  int v_rowmajor;             // EPI_ROPE: ALSO store the V heads row-major into C (ldc then spans Q|K|V)
  // optional per-head RMSNorm on Q and K before the rotation (Gemma-3 q_norm / k_norm):
  // y = bf16(x * rstd * (1 + w)), w is [head_dim]; columns < q_cols are Q heads
  const __nv_bfloat16* q_norm_w;
  const __nv_bfloat16* k_norm_w;
  int q_cols;
  float qk_norm_eps;
  // GLU
  int act;                    // 0 = SiLU, 1 = GELU(tanh)
  // cross-entropy (EPI_CE)
  const long long* target_ids;  // [T]; row r is scored against target_ids[r % T]
  int T;
  float2* ce_part;            // [n_tiles][M] (running max, sum of exp)
  float* tgt_logit;           // [M]
  int* ce_arg;                // [n_tiles][M] first column attaining the tile maximum (torch.argmax order)
  // EPI_ROPE, fused K/V publish (pub_rows > 0): the leading rows of the pass ALSO store their K heads and V
  // heads into per-layer prefix buffers -- rows [0, pub_pre_n) are prefix positions pub_key0 + r and go to
  // all pub_n targets (the ranks of the box incl. this one: peer-mapped pointers, st.global over NVLink, so
  // the per-layer all-gather of the split prefill rides in this GEMM's epilogue tile by tile); rows
  // [pub_pre_n, pub_rows) are the base suffix's rows and go to target pub_self only, at key pub_base_key + ...
  int pub_rows, pub_pre_n, pub_key0, pub_base_key, pub_n, pub_self;
  int pub_kcol0;              // first K-head column of the projection (= q_dim)
  long long pub_ldk, pub_ldvt;
  __nv_bfloat16* pub_k[8];    // [key][pub_ldk], column 0 = first K-head column
  __nv_bfloat16* pub_vt[8];   // [dim][pub_ldvt]
  // Early signal (pub_flag[0] != nullptr; pub_kcol0 a multiple of BN): the tiles that publish -- the row blocks with
  // prefix / base rows x the K and V column tiles -- are walked FIRST, and the epilogue warp that stores the last of
  // them raises this rank's flag (word pub_self := pub_epoch) on all pub_n ranks from inside the GEMM. The other
  // ranks' wait, launched after this GEMM, then has the rest of the GEMM as slack against rank-to-rank jitter.
  unsigned long long* pub_flag[8];
  unsigned long long pub_epoch;
  unsigned int* pub_done;     // device counter, 0 between launches
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 256;

// CTAS = 2: a CTA pair computes a 256 x BN tile with one cta_group::2 MMA stream; each CTA stages
// its own 128 rows of A and HALF of the B tile, so a stage is 32 KB instead of 48 KB at BN = 256
// and the ring is 7 deep instead of 4 (3500 vs 2050 cycles of operand buffering at the MMA's
// consumption rate) while L2->smem traffic per flop drops by a third.
template <int BN, int CTAS = 1>
struct GemmCfg {
  static constexpr int kStageBytesA = GEMM_BM * GEMM_BK * 2;
  static constexpr int kStageBytesB = (BN / CTAS) * GEMM_BK * 2;
  static constexpr int kStageBytes = kStageBytesA + kStageBytesB;
  static constexpr int kStages = CTAS == 2 ? (BN >= 256 ? 7 : 9) : (BN >= 256) ? 4 : (BN >= 128 ? 6 : (BN >= 64 ? 8 : 10));
  static constexpr int kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
};

__device__ __forceinline__ float act_silu(float x) { return x / (1.0f + __expf(-x)); }
__device__ __forceinline__ float act_gelu_tanh(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float u = k0 * (x + k1 * x * x * x);
  return 0.5f * x * (1.0f + tanhf(u));
}

__device__ __forceinline__ void store_bf16x32(__nv_bfloat16* dst, const uint32_t (&p)[16],
                                              int valid_cols) {
  // dst is 16-byte aligned; valid_cols counts the columns (of 32) inside the matrix.
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    if (v * 8 + 8 <= valid_cols) {
      uint4 q = make_uint4(p[v * 4 + 0], p[v * 4 + 1], p[v * 4 + 2], p[v * 4 + 3]);
      *reinterpret_cast<uint4*>(dst + v * 8) = q;
    } else if (v * 8 < valid_cols) {              // ragged last vector (N not a multiple of 8)
      uint16_t* d16 = reinterpret_cast<uint16_t*>(dst + v * 8);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (v * 8 + j < valid_cols)
          d16[j] = static_cast<uint16_t>((j & 1) ? (p[v * 4 + (j >> 1)] >> 16) : (p[v * 4 + (j >> 1)] & 0xffffu));
    }
  }
}

template <int BN, int EPI, int HD = 128, int CTAS = 1>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmap_a,
                    const __grid_constant__ CUtensorMap tmap_b, const GemmParams p) {
  using Cfg = GemmCfg<BN, CTAS>;
  static_assert(CTAS == 1 || CTAS == 2, "one CTA or a CTA pair");
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * Cfg::kStageBytesA;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* full_bar = bars;                    // [kStages]
  uint64_t* empty_bar = bars + kStages;         // [kStages]
  uint64_t* tmem_full = bars + 2 * kStages;     // [2]
  uint64_t* tmem_empty = bars + 2 * kStages + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // Work units: CTAS == 1 -> 128 x BN tiles, one per CTA; CTAS == 2 -> 256 x BN tiles, one per
  // CTA pair (this CTA owns rows [128 rank, 128 rank + 128) of it). Units are walked persistently.
  const uint32_t cta_rank = CTAS == 2 ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  const int unit0 = blockIdx.x / CTAS, unit_stride = gridDim.x / CTAS;
  const int M = p.M_dev != nullptr ? min(p.M, __ldg(p.M_dev)) : p.M;
  const int m_tiles = ((M + GEMM_BM - 1) / GEMM_BM + CTAS - 1) / CTAS;    // in units of CTAS*128 rows
  const int n_tiles = (p.N + BN - 1) / BN;
  const int tiles_per_batch = m_tiles * n_tiles;
  const int num_tiles = tiles_per_batch * (p.batch > 1 ? p.batch : 1);
  const int k_blocks = (p.K + GEMM_BK - 1) / GEMM_BK;
  const int group_m = max(1, (p.group_m > 0 ? p.group_m : 16) / CTAS);

  // publish-first order (see GemmParams::pub_flag): virtual tiles [0, n_first) are the publishing tiles, the rest is
  // the normal order with those tiles skipped (tile_active)
  const bool pub_first = EPI == EPI_ROPE && p.pub_flag[0] != nullptr && p.pub_rows > 0 && p.batch <= 1;
  const int pub_mu = pub_first ? (p.pub_rows + GEMM_BM * CTAS - 1) / (GEMM_BM * CTAS) : 0;   // row units that publish
  const int pub_n0 = pub_first ? p.pub_kcol0 / BN : 0;                                        // first K/V column tile
  const int n_first = pub_first ? pub_mu * (n_tiles - pub_n0) : 0;
  const int num_virtual = num_tiles + n_first;
  auto tile_coords = [&](int tile, int& m_blk, int& n_blk) -> int {
    if constexpr (EPI == EPI_ROPE) {
      if (tile < n_first) {
        m_blk = (tile % pub_mu) * CTAS + static_cast<int>(cta_rank);
        n_blk = pub_n0 + tile / pub_mu;
        return 0;
      }
      tile -= n_first;
    }
    const int b = tile / tiles_per_batch;
    tile -= b * tiles_per_batch;
    const int per_group = group_m * n_tiles;
    const int g = tile / per_group;
    const int first_m = g * group_m;
    const int rows_in_group = min(group_m, m_tiles - first_m);
    const int r = tile - g * per_group;
    m_blk = (first_m + r % rows_in_group) * CTAS + static_cast<int>(cta_rank);   // this CTA's 128-row block
    n_blk = r / rows_in_group;
    return b;
  };

  // causal products: which tiles exist and which K blocks they need (same answer in every warp role)
  auto tile_active = [&](int tile, int m_blk, int n_blk) {
    if (pub_first && tile >= n_first && m_blk / CTAS < pub_mu && n_blk >= pub_n0) return false;   // done in the first segment
    return p.causal != 1 || n_blk * BN <= m_blk * GEMM_BM + GEMM_BM - 1;
  };
  auto k_first = [&](int m_blk) { return p.causal == 3 ? min(k_blocks, (m_blk * GEMM_BM) / GEMM_BK) : 0; };
  auto k_last = [&](int m_blk) {
    return p.causal == 2 ? min(k_blocks, ((m_blk + 1) * GEMM_BM + GEMM_BK - 1) / GEMM_BK) : k_blocks;
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4 * CTAS);     // the epilogue warps of BOTH CTAs release the leader
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (CTAS == 2) {
      tmem_alloc_pair(tmem_slot, Cfg::kTmemCols);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, Cfg::kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (CTAS == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = unit0; tile < num_virtual; tile += unit_stride) {
        int m_blk, n_blk;
        const int bi = tile_coords(tile, m_blk, n_blk);
        const int a_r0 = bi * p.a_row + m_blk * GEMM_BM, a_c0 = bi * p.a_col;
        const int b_r0 = bi * p.b_row + n_blk * BN, b_c0 = bi * p.b_col;
        if (!tile_active(tile, m_blk, n_blk)) continue;
        const int kf = k_first(m_blk), kl = k_last(m_blk);
        const int kb0 = p.k_stagger > 0 ? static_cast<int>(((long long)tile * p.k_stagger) % k_blocks) : 0;
        for (int kk = kf; kk < kl; ++kk) {
          const int kb = kk + kb0 < k_blocks ? kk + kb0 : kk + kb0 - k_blocks;
          mbar_wait(&empty_bar[stage], phase ^ 1, 0x100 + stage);
          if constexpr (CTAS == 2) {
            // both CTAs' bytes are credited to the LEADER's full barrier, which the MMA issuer waits on
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
            const uint32_t bar = mapa_shared(smem_u32(&full_bar[stage]), 0);
            tma_load_2d_pair(smem_a + stage * Cfg::kStageBytesA, &tmap_a, bar, a_c0 + kb * GEMM_BK, a_r0);
            tma_load_2d_pair(smem_b + stage * Cfg::kStageBytesB, &tmap_b, bar, b_c0 + kb * GEMM_BK,
                             b_r0 + static_cast<int>(cta_rank) * (BN / 2));
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
            tma_load_2d(smem_a + stage * Cfg::kStageBytesA, &tmap_a, &full_bar[stage],
                        a_c0 + kb * GEMM_BK, a_r0);
            tma_load_2d(smem_b + stage * Cfg::kStageBytesB, &tmap_b, &full_bar[stage],
                        b_c0 + kb * GEMM_BK, b_r0);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM * CTAS, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = unit0; tile < num_virtual; tile += unit_stride) {
        int m_blk, n_blk;
        tile_coords(tile, m_blk, n_blk);
        if (!tile_active(tile, m_blk, n_blk)) continue;
        const int kf = k_first(m_blk), kl = k_last(m_blk);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1, 0x200 + acc);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kf; kb < kl; ++kb) {
          mbar_wait(&full_bar[stage], phase, 0x300 + stage);
          tc_fence_after();
          const uint64_t da = umma_desc_k_sw128(smem_u32(smem_a + stage * Cfg::kStageBytesA));
          const uint64_t db = umma_desc_k_sw128(smem_u32(smem_b + stage * Cfg::kStageBytesB));
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            // advance 16 elements (32 B) along K inside the 128-B swizzle atom: +2 in 16-B units
            if constexpr (CTAS == 2) umma_bf16_ss_pair(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb != kf) || k != 0);
            else umma_bf16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb != kf) || k != 0);
          }
          // frees the smem slot (in both CTAs of a pair) once these MMAs retire
          if constexpr (CTAS == 2) umma_commit_pair(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        // accumulator complete -> epilogue (of both CTAs)
        if constexpr (CTAS == 2) umma_commit_pair(&tmem_full[acc]); else umma_commit(&tmem_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue
    const int sub = warp & 3;                 // TMEM sub-partition owned by this warp
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = unit0; tile < num_virtual; tile += unit_stride) {
      int m_blk, n_blk;
      const int bi = tile_coords(tile, m_blk, n_blk);
      if (!tile_active(tile, m_blk, n_blk)) continue;
      mbar_wait(&tmem_full[acc], acc_phase, 0x400 + acc);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(sub * 32) << 16) + acc * BN;
      const int row = m_blk * GEMM_BM + sub * 32 + lane;
      const bool row_ok = row < M;
      const int n0 = n_blk * BN;
      (void)bi;

      if constexpr (EPI == EPI_STORE) {
        const long long c_off = (long long)bi * p.c_batch;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t r[32];
          tmem_ld_32x32(t_row + c0, r);
          tmem_ld_wait();
          const int col = n0 + c0;
          const int valid = min(32, p.N - col);
          if (row_ok && valid > 0) {
            if (p.C32 != nullptr) {                  // raw fp32 accumulators (16-byte aligned rows)
              float* dst = p.C32 + c_off + (long long)row * p.ldc + col;
#pragma unroll
              for (int v = 0; v < 8; ++v) {
                if (v * 4 + 4 <= valid) {
                  *reinterpret_cast<uint4*>(dst + v * 4) = make_uint4(r[v * 4], r[v * 4 + 1], r[v * 4 + 2], r[v * 4 + 3]);
                } else {
#pragma unroll
                  for (int j = 0; j < 4; ++j)
                    if (v * 4 + j < valid) dst[v * 4 + j] = __uint_as_float(r[v * 4 + j]);
                }
              }
              continue;
            }
            if (p.bias != nullptr) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < valid) r[j] = __float_as_uint(__uint_as_float(r[j]) + __bfloat162float(p.bias[col + j]));
            }
            uint32_t o[16];
            if (p.R != nullptr) {
              const uint4* rp = reinterpret_cast<const uint4*>(p.R + c_off + (long long)row * p.ldr + col);
#pragma unroll
              for (int v = 0; v < 4; ++v) {
                uint4 q = make_uint4(0, 0, 0, 0);
                if (v * 8 + 8 <= valid) q = rp[v];
                const uint32_t qq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  // the GEMM result is rounded to bf16 first (as a separate Linear would), then
                  // added to the bf16 residual.
                  float a0 = bf16_round(__uint_as_float(r[v * 8 + 2 * j])) + bf16_lo(qq[j]);
                  float a1 = bf16_round(__uint_as_float(r[v * 8 + 2 * j + 1])) + bf16_hi(qq[j]);
                  o[v * 4 + j] = pack_bf16x2(a0, a1);
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                o[j] = pack_bf16x2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
            }
            store_bf16x32(p.C + c_off + (long long)row * p.ldc + col, o, valid);
          }
        }
      } else if constexpr (EPI == EPI_ROPE) {
        static_assert(BN % HD == 0 && HD % 64 == 0, "a tile holds whole heads");
        constexpr int HALF = HD / 2;
        const int pos = p.pos_base + (!row_ok ? 0 : p.row_info != nullptr ? p.row_info[row].y
                                                                          : (row % p.rows_per_seq));
        const float4* cosp = reinterpret_cast<const float4*>(p.rope_cos + (long long)pos * HALF);
        const float4* sinp = reinterpret_cast<const float4*>(p.rope_sin + (long long)pos * HALF);
#pragma unroll 1
        for (int hb = 0; hb < BN; hb += HD) {
          const int col_h = n0 + hb;
          if (col_h >= p.N) break;
          if (col_h < p.rope_cols) {
            // optional q_norm / k_norm: RMS over the head's HD columns of the bf16 projection
            const __nv_bfloat16* nw = nullptr;
            float rstd = 1.0f;
            if (p.q_norm_w != nullptr) {
              nw = col_h < p.q_cols ? p.q_norm_w : p.k_norm_w;
              float ss = 0.0f;
#pragma unroll 1
              for (int c0 = 0; c0 < HD; c0 += 32) {
                uint32_t r[32];
                tmem_ld_32x32(t_row + hb + c0, r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  const float a = bf16_round(__uint_as_float(r[j]));
                  ss = fmaf(a, a, ss);
                }
              }
              rstd = rsqrtf(ss * (1.0f / HD) + p.qk_norm_eps);
            }
#pragma unroll 1
            for (int c0 = 0; c0 < HALF; c0 += 32) {
              uint32_t x1[32], x2[32];
              tmem_ld_32x32(t_row + hb + c0, x1);
              tmem_ld_32x32(t_row + hb + HALF + c0, x2);
              tmem_ld_wait();
              if (row_ok) {
                uint32_t o1[16], o2[16];
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                  const float4 c = cosp[(c0 >> 2) + j4];
                  const float4 s = sinp[(c0 >> 2) + j4];
                  const float cc[4] = {c.x, c.y, c.z, c.w};
                  const float ss[4] = {s.x, s.y, s.z, s.w};
                  float y1[4], y2[4];
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    // projection output is bf16 in the reference before the rotation is applied
                    float a = bf16_round(__uint_as_float(x1[j4 * 4 + j]));
                    float b = bf16_round(__uint_as_float(x2[j4 * 4 + j]));
                    if (nw != nullptr) {
                      const int cn = c0 + j4 * 4 + j;
                      a = bf16_round((a * rstd) * (1.0f + __bfloat162float(nw[cn])));
                      b = bf16_round((b * rstd) * (1.0f + __bfloat162float(nw[HALF + cn])));
                    }
                    y1[j] = bf16_round(a * cc[j]) + bf16_round(-b * ss[j]);
                    y2[j] = bf16_round(b * cc[j]) + bf16_round(a * ss[j]);
                  }
                  o1[j4 * 2 + 0] = pack_bf16x2(y1[0], y1[1]);
                  o1[j4 * 2 + 1] = pack_bf16x2(y1[2], y1[3]);
                  o2[j4 * 2 + 0] = pack_bf16x2(y2[0], y2[1]);
                  o2[j4 * 2 + 1] = pack_bf16x2(y2[2], y2[3]);
                }
                __nv_bfloat16* dst = p.C + (long long)row * p.ldc + col_h + c0;
                store_bf16x32(dst, o1, 32);
                store_bf16x32(dst + HALF, o2, 32);
                if (row < p.pub_rows && col_h >= p.pub_kcol0) {       // K head of a prefix / base row
                  const bool pre = row < p.pub_pre_n;
                  const long long off = (long long)(pre ? p.pub_key0 + row : p.pub_base_key + row - p.pub_pre_n) *
                                            p.pub_ldk + (col_h - p.pub_kcol0) + c0;
#pragma unroll 1
                  for (int q = pre ? 0 : p.pub_self; q < (pre ? p.pub_n : p.pub_self + 1); ++q) {
                    store_bf16x32(p.pub_k[q] + off, o1, 32);
                    store_bf16x32(p.pub_k[q] + off + HALF, o2, 32);
                  }
                }
              }
            }
          } else {
#pragma unroll 1
            for (int c0 = 0; c0 < HD; c0 += 32) {
              uint32_t r[32];
              tmem_ld_32x32(t_row + hb + c0, r);
              tmem_ld_wait();
              if (row_ok) {
                // V head: transposed store; the 32 lanes of a warp hold 32 consecutive rows, so each
                // store instruction writes one 64-byte segment of a Vt row
                __nv_bfloat16* dst = p.Vt + (long long)(col_h + c0 - p.rope_cols) * p.ldvt + row;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  if (col_h + c0 + j < p.N) dst[(long long)j * p.ldvt] = __float2bfloat16_rn(__uint_as_float(r[j]));
                }
                if (row < p.pub_rows) {                               // V head of a prefix / base row
                  const bool pre = row < p.pub_pre_n;
                  const long long off = (long long)(col_h + c0 - p.rope_cols) * p.pub_ldvt +
                                        (pre ? p.pub_key0 + row : p.pub_base_key + row - p.pub_pre_n);
#pragma unroll 1
                  for (int q = pre ? 0 : p.pub_self; q < (pre ? p.pub_n : p.pub_self + 1); ++q) {
                    __nv_bfloat16* pd = p.pub_vt[q] + off;
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                      if (col_h + c0 + j < p.N) pd[(long long)j * p.pub_ldvt] = __float2bfloat16_rn(__uint_as_float(r[j]));
                  }
                }
                if (p.v_rowmajor) {
                  uint32_t o[16];
#pragma unroll
                  for (int j = 0; j < 16; ++j)
                    o[j] = pack_bf16x2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
                  store_bf16x32(p.C + (long long)row * p.ldc + col_h + c0, o, min(32, p.N - (col_h + c0)));
                }
              }
            }
          }
        }
      } else if constexpr (EPI == EPI_GLU) {
        // tile columns [0, BN/2) are gate rows j0.., columns [BN/2, BN) the matching up rows
        constexpr int H = BN / 2;
        const int out_n = p.N / 2;
        const int j0 = n_blk * H;
#pragma unroll 1
        for (int c0 = 0; c0 < H; c0 += 32) {
          uint32_t g[32], u[32];
          tmem_ld_32x32(t_row + c0, g);
          tmem_ld_32x32(t_row + H + c0, u);
          tmem_ld_wait();
          const int col = j0 + c0;
          const int valid = min(32, out_n - col);
          if (row_ok && valid > 0) {
            uint32_t o[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float y[2];
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const float gv = bf16_round(__uint_as_float(g[2 * j + e]));
                const float uv = bf16_round(__uint_as_float(u[2 * j + e]));
                const float a = bf16_round(p.act == 0 ? act_silu(gv) : act_gelu_tanh(gv));
                y[e] = a * uv;
              }
              o[j] = pack_bf16x2(y[0], y[1]);
            }
            store_bf16x32(p.C + (long long)row * p.ldc + col, o, valid);
          }
        }
      } else if constexpr (EPI == EPI_CE) {
        const int tcol = row_ok ? static_cast<int>(p.target_ids[row % p.T]) : -1;
        float run_m = -INFINITY, run_s = 0.0f;
        int run_arg = n0;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t r[32];
          tmem_ld_32x32(t_row + c0, r);
          tmem_ld_wait();
          const int col = n0 + c0;
          const int valid = min(32, p.N - col);
          if (valid <= 0) break;
          float v[32];
          float cm = -INFINITY;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            v[j] = (j < valid) ? bf16_round(__uint_as_float(r[j])) : -INFINITY;
            cm = fmaxf(cm, v[j]);
          }
          if (cm > run_m) {                       // strictly: an earlier column keeps a tie
            int first = 0;
#pragma unroll
            for (int j = 31; j >= 0; --j)
              if (v[j] == cm) first = j;
            run_arg = col + first;
          }
          if (tcol >= col && tcol < col + valid) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col + j == tcol) p.tgt_logit[row] = v[j];
          }
          const float nm = fmaxf(run_m, cm);
          float s = 0.0f;
#pragma unroll
          for (int j = 0; j < 32; ++j) s += __expf(v[j] - nm);
          run_s = run_s * __expf(run_m - nm) + s;
          run_m = nm;
        }
        if (row_ok) {
          p.ce_part[(long long)n_blk * p.M + row] = make_float2(run_m, run_s);
          if (p.ce_arg != nullptr) p.ce_arg[(long long)n_blk * p.M + row] = run_arg;
        }
      } else if constexpr (EPI == EPI_STORE_T) {
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          if (n0 + c0 >= p.N) break;
          uint32_t r[32];
          tmem_ld_32x32(t_row + c0, r);
          tmem_ld_wait();
          if (row_ok) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int n = n0 + c0 + j;
              if (n < p.N)
                p.C[(long long)n * p.ldc + row] = __float2bfloat16_rn(__uint_as_float(r[j]));
            }
          }
        }
      }

      if constexpr (EPI == EPI_ROPE) {
        if (pub_first && tile < n_first) {
          // every epilogue warp of every publishing tile counts in; the last one raises the flags. Each lane fences
          // its own stores (into the peers' memory) at system scope before its warp's increment.
          __threadfence_system();
          __syncwarp();
          if (lane == 0) {
            const unsigned int expect = static_cast<unsigned int>(n_first) * 4u * CTAS;
            if (atomicAdd(p.pub_done, 1u) + 1u == expect) {
              atomicExch(p.pub_done, 0u);
              __threadfence_system();
              for (int q = 0; q < p.pub_n; ++q)
                atomicExch_system(p.pub_flag[q] + p.pub_self, p.pub_epoch);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CTAS == 2) mbar_arrive_cluster(mapa_shared(smem_u32(&tmem_empty[acc]), 0));
        else mbar_arrive(&tmem_empty[acc]);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  if constexpr (CTAS == 2) cluster_sync_all(); else __syncthreads();   // the peer's smem/TMEM stay alive
  if (warp == 2) {
    tc_fence_after();
    if constexpr (CTAS == 2) tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
    else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace bma
