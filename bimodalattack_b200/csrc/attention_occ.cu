// Shared-prefix + causal-tail attention for head_dim 128, "two CTAs per SM" variant (the default).
//
// Same problem, masks and TMA layouts as attention_split.cu / attention_tc.cu. ncu showed the
// one-CTA-per-SM kernels latency-bound (tensor pipe 25 %, no unit saturated): per key block the chain
// S = Q K^T -> tcgen05.ld -> softmax -> tcgen05.st P -> O += P V is long and a single CTA has only one
// or two blocks in flight. Instead of a deeper hand-built pipeline this kernel is made SMALL enough
// for two CTAs to be resident per SM, so the hardware interleaves two independent chains:
//   * 64-key blocks (BN = 64): K tile 16 KB, V^T tile 16 KB, two stages each; ONE Q tile (32 KB); a
//     16 KB staging tile for the output's TMA store -> 113 KB of shared memory per CTA;
//   * tensor memory: S_0 S_1 (64 columns each) + O (128 columns) = 256 of the 512 columns;
//   * 256 threads, <= 128 registers: ONE softmax thread per query row (4 warps), nothing is exchanged
//     between threads. A block whose 64 keys are visible to all 32 rows of the warp (most prefix
//     blocks) keeps its 64 scores in registers between the maximum and the exponentials; the other
//     blocks (partial last prefix block, base-suffix boundary, causal tail: ~4 of 13 per item) test a
//     64-bit visibility mask of the row with predicates (ptxas: R2P + predicated FMNMX / FSEL), 32
//     columns at a time, as real calls so the hot path carries no mask code;
//   * P never touches shared memory: bf16 pairs go back over the S columns with tcgen05.st and the
//     P V MMAs read their A operand from tensor memory (as in attention_split.cu);
//   * an item's output rows are written AFTER the first key block of the next item (its last PV and
//     the slowest warp of its last masked block get a block of slack), through a swizzled staging
//     tile and one TMA store per [32 rows][64 columns] box.
//   warp 0  TMA producer   warp 1  MMA issuer (flat stream over work items; S_g is issued before the
//   PV of block g-1)   warp 2  TMEM allocator   warps 4-7  softmax + output
//
// Measured (tools/ab_attention.py, same process, sustained clocks, 15 872 rows x 616 prefix keys x 32
// heads): 324 us for the round-2 two-pass version -> 262 us. Tried on the way and dropped (numbers in
// DESIGN.md 3.2): Q as the TMEM A operand with a single S buffer (366 us: the S -> softmax -> PV chain
// of a CTA becomes strictly serial), two threads per row / eight softmax warps (347 us), back-off in
// the control warps' barrier polls (no change). Scale-and-subtract / row sums use packed fp32 pairs (FFMA2 /
// FADD2, bit-identical, ~3 %; BMA_OCC_PACKED=0 = scalar). BMA_OCC_TIMING=1 prints per-phase cycle counts of the
// softmax warps of CTA 0, BMA_OCC_KNOCK removes one cost at a time (timing only, wrong results).
#include <cstdio>
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
constexpr int OSTAGE = 32 * 128;              // per softmax warp: [32 rows][64 columns] bf16 staged for the TMA store
// Two CTAs per SM: 2 x (SMEM + 1 KB reserved) must fit the SM's 228 KB, which leaves 768 bytes of alignment
// slack (the dynamic window starts 1024-aligned in practice; the kernel traps if the slack does not suffice).
constexpr int SMEM_USED = QTILE + 2 * KTILE + 2 * VTILE + 4 * OSTAGE + 256 /*barriers*/;
constexpr int SMEM = SMEM_USED + 768;
static_assert(2 * (SMEM + 1024) <= 228 * 1024, "two CTAs per SM");
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
  int pv_first;               // MMA issuer: last PV of an item before the wait for the next Q tile (A/B switch)
  int defer;                  // write an item's output after the first block of the next item (A/B switch)
  unsigned long long* dbg;    // TIMING instantiation only
  int knock;                  // TIMING only: knock-out experiments (wrong results, timing only), BMA_OCC_KNOCK bits:
                              // 1 no ex2, 2 no maximum, 4 no QK MMAs, 8 no PV MMAs, 16 no K/V loads, 32 no output
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// Packed fp32 pairs (Blackwell FFMA2 / FADD2): {d0, d1} = {a0, a1} * b + c and {d0, d1} += {a0, a1}; per lane the same
// IEEE result as the scalar instructions, half the issue slots.
__device__ __forceinline__ void fma2(float& d0, float& d1, float a0, float a1, float b, float c) {
  uint64_t d, a, bb, cc;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
  asm("mov.b64 %0, {%1, %1};" : "=l"(cc) : "f"(c));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(bb), "l"(cc));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
}
__device__ __forceinline__ void add2(float& d0, float& d1, float a0, float a1) {
  uint64_t d, a;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(d0), "f"(d1));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(d), "l"(a));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// TMA store of one [rows][64 columns] bf16 box (128-byte swizzle) from shared memory; bulk-group completion.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

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

// bit c set <=> key column c of the block is visible to this row (columns [vlo, vhi))
__device__ __forceinline__ uint64_t window_mask(int vlo, int vhi) {
  if (vhi <= vlo) return 0ull;
  return (vhi >= 64 ? ~0ull : ((1ull << vhi) - 1ull)) & ~((1ull << vlo) - 1ull);
}
// maximum over the visible ones of 32 score columns (bits = their visibility); skipped when no row of the warp sees any
__device__ __noinline__ float masked_max32(uint32_t taddr, uint32_t bits) {
  if (__all_sync(0xffffffffu, bits == 0u)) return -INFINITY;
  uint32_t r[32];
  tmem_ld_32x32(taddr, r);
  tmem_ld_wait();
  float raw = -INFINITY;
#pragma unroll
  for (int i = 0; i < 32; ++i)
    if ((bits >> i) & 1u) raw = fmaxf(raw, __uint_as_float(r[i]));
  return raw;
}
// P over 32 score columns with the visibility bits applied, stored as 16 packed columns at tp (which only overlaps
// score columns that have been read already); returns the row sum
__device__ __noinline__ float masked_exp32_store(uint32_t ts, uint32_t tp, uint32_t bits, float scale_log2, float m_use) {
  uint32_t pk[16];
  float rs = 0.0f;
  if (__all_sync(0xffffffffu, bits == 0u)) {
#pragma unroll
    for (int i = 0; i < 16; ++i) pk[i] = 0u;
  } else {
    uint32_t r[32];
    tmem_ld_32x32(ts, r);
    tmem_ld_wait();
    float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      float x0, x1;
      fma2(x0, x1, __uint_as_float(r[i]), __uint_as_float(r[i + 1]), scale_log2, -m_use);
      x0 = ((bits >> i) & 1u) ? x0 : -INFINITY;
      x1 = ((bits >> (i + 1)) & 1u) ? x1 : -INFINITY;
      const float p0 = ex2(x0), p1 = ex2(x1);               // -inf -> 0
      add2(s0, s1, p0, p1);
      pk[i >> 1] = pack_bf16x2(p0, p1);
    }
    rs = s0 + s1;
  }
  tmem_st_32x16(tp, pk);
  return rs;
}
// Rare path: multiply one row of the O accumulator (128 fp32 columns in tensor memory) by f.
__device__ __noinline__ void rescale_o_row(uint32_t to, float f) {
#pragma unroll
  for (int c0 = 0; c0 < HD; c0 += 32) {
    uint32_t r[32];
    tmem_ld_32x32(to + c0, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * f);
    tmem_st_32x32(to + c0, r);
  }
  tmem_st_wait();
}

// TIMING: the softmax warps of CTA 0 accumulate clock64() deltas per phase in registers and write them to
// p.dbg at the end (tools/bench_attention.py, BMA_OCC_TIMING=1); p.knock then also applies.
template <bool TIMING, bool PACKED>
__global__ void __launch_bounds__(THREADS, 2)
attention_occ_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kt,
                     const __grid_constant__ CUtensorMap tm_kp, const __grid_constant__ CUtensorMap tm_vt,
                     const __grid_constant__ CUtensorMap tm_vp, const __grid_constant__ CUtensorMap tm_o,
                     const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;                          // [tile]
  uint8_t* sK = sQ + QTILE;                    // [2][tile]
  uint8_t* sV = sK + 2 * KTILE;                // [2][tile]   V^T: [128 dims][64 keys]
  uint8_t* sO = sV + 2 * VTILE;                // [4 softmax warps][32 rows][64 columns], 128-byte swizzle
  uint64_t* bars = reinterpret_cast<uint64_t*>(sO + 4 * OSTAGE);
  if (smem + SMEM_USED > smem_raw + SMEM) __trap();
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
    tma_prefetch_desc(&tm_vt); tma_prefetch_desc(&tm_vp); tma_prefetch_desc(&tm_o);
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
          if (TIMING && (p.knock & 16)) { mbar_arrive(&k_full[st]); return; }
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
          if (TIMING && (p.knock & 16)) { mbar_arrive(&v_full[st]); return; }
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
        if (!(TIMING && (p.knock & 8)))
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
        const int tile = item % m_tiles;
        const int r0 = tile * BM, r1 = min(M, r0 + BM);
        const int kt0 = tail_start(r0, p.S, p.row_meta);
        const int nb = n_pre + (p.no_tail ? 0 : (r1 - kt0 + BN - 1) / BN);
        // the Q tile is single-buffered: its load only starts when the previous item's last S has completed, so
        // the previous item's last PV is issued before waiting for it, not behind it
        if (have_pending && p.pv_first && !mbar_try_wait(q_full, item_n & 1)) {
          issue_pv(pend_g, pend_n, pend_first, pend_item_par);
          have_pending = false;
        }
        mbar_wait(q_full, item_n & 1, 0x840);
        for (int j = 0; j < nb; ++j, ++it) {
          const Block b = block_of(j, n_pre, p.P, kt0, r1);
          const int st = it & 1;
          mbar_wait(&k_full[st], (it >> 1) & 1, 0x850 + st);     // (S_st is free: the PV of block it-2 precedes this MMA)
          tc_fence_after();
          const uint32_t idesc = umma_idesc_bf16(BM, b.n_mma);
          const uint32_t qa = smem_u32(sQ), kb = smem_u32(sK + st * KTILE);
          if (!(TIMING && (p.knock & 4))) {
#pragma unroll
            for (int ks = 0; ks < HD / 16; ++ks)
              umma_bf16_ss(tmem_s + st * BN, umma_desc_k_sw128(qa + (ks >> 2) * QPANEL + (ks & 3) * 32),
                           umma_desc_k_sw128(kb + (ks >> 2) * KPANEL + (ks & 3) * 32), idesc, ks != 0);
          }
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
    uint32_t acc[14];                            // TIMING: cycle counts per phase, kept in registers
#pragma unroll
    for (int i = 0; i < 14; ++i) acc[i] = 0;

    // ---- end of an item: normalise and store this thread's output row (e_it = the item's last block + 1).
    // The warp's 32 rows go out as two [32 rows][64 columns] boxes: each thread writes its row's 128 bytes into the
    // staging buffer (16-byte chunk c of row r at chunk position c ^ (r & 7): the 128-byte swizzle of the tensor
    // map, bank-conflict free), one TMA store per box. Rows >= M of the tensor map are clipped by the TMA unit;
    // rows between the live count and M receive finite values nobody reads.
    auto write_output = [&](float e_l, int e_head, int e_r0, uint32_t e_it) {
      long long t0 = 0, t1 = 0;
      if (TIMING) t0 = clock64();
      mbar_wait(&p_empty[(e_it - 1) & 1], ((e_it - 1) >> 1) & 1, 0x940);   // the item's last PV has retired
      tc_fence_after();
      if (TIMING) t1 = clock64();
      const float inv = e_l > 0.0f ? 1.0f / e_l : 0.0f;
      uint8_t* so = sO + sub * OSTAGE;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (lane == 0) tma_store_wait_read();            // the previous box has been read out of the buffer
        __syncwarp();
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 32) {
          uint32_t a[32];
          tmem_ld_32x32(tmem_o + lane_off + h * 64 + c0, a);
          tmem_ld_wait();
#pragma unroll
          for (int v4 = 0; v4 < 4; ++v4) {
            const int chunk = (c0 >> 3) + v4;
            *reinterpret_cast<uint4*>(so + lane * 128 + ((chunk ^ (lane & 7)) << 4)) = make_uint4(
                pack_bf16x2(__uint_as_float(a[v4 * 8 + 0]) * inv, __uint_as_float(a[v4 * 8 + 1]) * inv),
                pack_bf16x2(__uint_as_float(a[v4 * 8 + 2]) * inv, __uint_as_float(a[v4 * 8 + 3]) * inv),
                pack_bf16x2(__uint_as_float(a[v4 * 8 + 4]) * inv, __uint_as_float(a[v4 * 8 + 5]) * inv),
                pack_bf16x2(__uint_as_float(a[v4 * 8 + 6]) * inv, __uint_as_float(a[v4 * 8 + 7]) * inv));
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0 && !(TIMING && (p.knock & 32))) {
          tma_store_2d(&tm_o, so, e_head * HD + h * 64, e_r0 + sub * 32);
          tma_store_commit();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
      if (TIMING) {
        acc[6] += 1;                                                        // items
        acc[7] += static_cast<uint32_t>(t1 - t0);                           // wait for the last PV
        acc[8] += static_cast<uint32_t>(clock64() - t1);                    // output rows
      }
    };
    float prev_l = 0.0f;
    int prev_head = 0, prev_r0 = 0;
    uint32_t prev_it = 0;
    bool have_prev = false;

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
        long long t0 = 0, t1 = 0, t2 = 0, t3 = 0;
        if (TIMING) t0 = clock64();
        mbar_wait(&s_full[st], (it >> 1) & 1, 0x910 + st);
        tc_fence_after();
        if (TIMING) t1 = clock64();

        // The reference maximum is only raised when the block exceeds it by more than 2^8. No wait on the
        // previous PV: P_g has its own buffer and barrier (s_full of block g implies that PV_{g-2}, the last
        // reader of this buffer, has retired -- the tensor pipe completes in order), so the softmax of block g
        // overlaps PV_{g-1}. Only the rare rescale needs O quiescent.
        float m_new, m_use, factor;
        auto update_max = [&](float raw) -> bool {
          const float bm = raw * p.scale_log2;
          const bool raise = bm > m_ref + kRaise;
          m_new = raise ? bm : m_ref;
          m_use = (m_new == -INFINITY) ? 0.0f : m_new;
          factor = raise ? ex2(m_ref - m_use) : 1.0f;
          const bool fix = raise && (m_ref != -INFINITY) && (j > 0);
          const bool any_fix = __any_sync(0xffffffffu, fix);
          if (any_fix) {                                   // rare: rescale this row of O in TMEM
            mbar_wait(&p_empty[(it - 1) & 1], ((it - 1) >> 1) & 1, 0x920);    // PV_{g-1} has retired
            tc_fence_after();
            rescale_o_row(tmem_o + lane_off, fix ? factor : 1.0f);
          }
          return any_fix;
        };

        float rs;
        if (all_visible) {
          // ---- every key of the block is visible to every row of the warp: the 64 scores stay in registers
          uint32_t ra[32], rb[32];
          tmem_ld_32x32(ts, ra);
          tmem_ld_32x32(ts + 32, rb);
          tmem_ld_wait();
          float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
          if (TIMING && (p.knock & 2)) {
            mx0 = 0.0f;
          } else {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {              // four independent chains (FMNMX3: two new values each)
              mx0 = fmaxf(mx0, fmaxf(__uint_as_float(ra[i]), __uint_as_float(ra[i + 1])));
              mx1 = fmaxf(mx1, fmaxf(__uint_as_float(ra[i + 2]), __uint_as_float(ra[i + 3])));
              mx2 = fmaxf(mx2, fmaxf(__uint_as_float(rb[i]), __uint_as_float(rb[i + 1])));
              mx3 = fmaxf(mx3, fmaxf(__uint_as_float(rb[i + 2]), __uint_as_float(rb[i + 3])));
            }
          }
          if (update_max(fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)))) {   // the rescale used the registers: read S again
            tmem_ld_32x32(ts, ra);
            tmem_ld_32x32(ts + 32, rb);
            tmem_ld_wait();
          }
          if (TIMING) t2 = clock64();
          float s0 = 0.0f, s1 = 0.0f;
          uint32_t pk[16];
          const bool no_ex2 = TIMING && (p.knock & 1);
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float x0, x1;
            if (PACKED) {
              fma2(x0, x1, __uint_as_float(ra[i]), __uint_as_float(ra[i + 1]), p.scale_log2, -m_use);
            } else {
              x0 = fmaf(__uint_as_float(ra[i]), p.scale_log2, -m_use);
              x1 = fmaf(__uint_as_float(ra[i + 1]), p.scale_log2, -m_use);
            }
            const float p0 = no_ex2 ? x0 : ex2(x0), p1 = no_ex2 ? x1 : ex2(x1);
            if (PACKED) add2(s0, s1, p0, p1); else { s0 += p0; s1 += p1; }
            pk[i >> 1] = pack_bf16x2(p0, p1);
          }
          tmem_st_32x16(ts, pk);                           // P_g: columns [0, 32) of S_st, bf16 pairs
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float x0, x1;
            if (PACKED) {
              fma2(x0, x1, __uint_as_float(rb[i]), __uint_as_float(rb[i + 1]), p.scale_log2, -m_use);
            } else {
              x0 = fmaf(__uint_as_float(rb[i]), p.scale_log2, -m_use);
              x1 = fmaf(__uint_as_float(rb[i + 1]), p.scale_log2, -m_use);
            }
            const float p0 = no_ex2 ? x0 : ex2(x0), p1 = no_ex2 ? x1 : ex2(x1);
            if (PACKED) add2(s0, s1, p0, p1); else { s0 += p0; s1 += p1; }
            pk[i >> 1] = pack_bf16x2(p0, p1);
          }
          tmem_st_32x16(ts + 16, pk);
          rs = s0 + s1;
        } else {
          // ---- blocks with hidden keys: the row's visible columns as a 64-bit mask, 32 columns at a time; a half
          // no row of the warp sees is skipped (P = 0). Columns the MMA did not cover (>= n_mma) are never visible.
          const uint64_t vis = window_mask(vlo, vhi);
          const uint32_t v_lo = static_cast<uint32_t>(vis), v_hi = static_cast<uint32_t>(vis >> 32);
          update_max(fmaxf(masked_max32(ts, v_lo), masked_max32(ts + 32, v_hi)));
          if (TIMING) t2 = clock64();
          rs = masked_exp32_store(ts, ts, v_lo, p.scale_log2, m_use);
          rs += masked_exp32_store(ts + 32, ts + 16, v_hi, p.scale_log2, m_use);
        }
        if (TIMING) t3 = clock64();
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[st]);
        l_run = l_run * factor + rs;
        m_ref = m_new;
        long long t4 = 0;
        if (TIMING) t4 = clock64();
        if (j == 0 && have_prev) {                         // the previous item's output, one block late
          write_output(prev_l, prev_head, prev_r0, prev_it);
          have_prev = false;
        }
        if (TIMING) {
          acc[0] += 1;                                                      // blocks
          if (j == 0) acc[10] += static_cast<uint32_t>(t1 - t0);            // wait for S, first block of an item
          acc[1] += static_cast<uint32_t>(t1 - t0);                         // wait for S
          acc[2] += static_cast<uint32_t>(t2 - t1);                         // S load + maximum
          acc[3] += static_cast<uint32_t>(t3 - t2);                         // exponentials + P store issue
          acc[4] += static_cast<uint32_t>(t4 - t3);                         // store wait + arrive
          acc[5] += all_visible ? 0 : 1;
          if (!all_visible) { acc[12] += static_cast<uint32_t>(t2 - t1); acc[13] += static_cast<uint32_t>(t3 - t2); }
        }
      }

      if (p.defer) {
        prev_l = l_run; prev_head = head; prev_r0 = r0; prev_it = it; have_prev = true;
      } else {
        write_output(l_run, head, r0, it);
      }
    }
    if (have_prev) write_output(prev_l, prev_head, prev_r0, prev_it);
    if (TIMING && blockIdx.x == 0 && lane == 0) {
#pragma unroll
      for (int i = 0; i < 14; ++i) p.dbg[sub * 16 + i] = acc[i];
    }
    if (lane == 0) tma_store_wait_read();                  // shared memory stays valid until the stores have read it
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
  CUtensorMap tq, tkt, tkp, tvt, tvp, to;
  if (!make_tmap_bf16(&to, a.out, a.M, (long long)n_heads * HD, a.ldo, 32)) return cudaErrorInvalidValue;
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
  // A/B switches (tools/ab_attention.py): BMA_OCC_DEFER, BMA_OCC_PVFIRST, BMA_OCC_PACKED (default 1), BMA_OCC_TIMING = 1 (per-phase
  // cycle counts of the softmax warps of CTA 0 on stderr; BMA_OCC_KNOCK then removes costs). Read once;
  // BMA_OCC_RELOAD = 1 re-reads them at every launch (in-process A/B).
  static int inited = 0, timing = 0, defer = 1, pv_first = 1, knock = 0, reload = 0, packed = 1;
  static unsigned long long* dbg = nullptr;
  if (!inited || reload) {
    const char* e = getenv("BMA_OCC_TIMING");
    timing = (e != nullptr && e[0] == '1');
    e = getenv("BMA_OCC_DEFER");
    defer = (e != nullptr && e[0] == '0') ? 0 : 1;
    e = getenv("BMA_OCC_PVFIRST");
    pv_first = (e != nullptr && e[0] == '0') ? 0 : 1;
    e = getenv("BMA_OCC_PACKED");
    packed = (e != nullptr && e[0] == '0') ? 0 : 1;
    e = getenv("BMA_OCC_KNOCK");
    knock = e != nullptr ? atoi(e) : 0;
    e = getenv("BMA_OCC_RELOAD");
    reload = (e != nullptr && e[0] == '1');
    if (!inited) {
      cudaError_t e1 = cudaFuncSetAttribute(attention_occ_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
      if (e1 == cudaSuccess) e1 = cudaFuncSetAttribute(attention_occ_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
      if (e1 == cudaSuccess) e1 = cudaFuncSetAttribute(attention_occ_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
      if (e1 != cudaSuccess) return e1;
      inited = 1;
    }
  }
  if (timing && dbg == nullptr && cudaMalloc(&dbg, 64 * sizeof(unsigned long long)) != cudaSuccess) return cudaErrorMemoryAllocation;
  Params p{};
  p.out = a.out; p.ldo = a.ldo;
  p.M = a.M; p.S = a.S > 0 ? a.S : 1; p.P = a.P; p.M_dev = a.M_dev;
  p.no_tail = a.no_tail;
  p.row_meta = a.row_meta;
  p.P_shared = a.row_meta != nullptr ? a.P_shared : a.P;
  p.pre_rows = a.row_meta != nullptr ? a.pre_rows : 0;
  p.n_heads = n_heads; p.kv_group = a.kv_group;
  p.scale_log2 = a.scale_log2;
  p.defer = defer; p.pv_first = pv_first;
  watchdog_attach();
  const int m_tiles = (a.M + BM - 1) / BM;
  long long grid = (long long)m_tiles * n_heads;
  if (grid > 2ll * num_sms()) grid = 2ll * num_sms();
  const int g = static_cast<int>(grid);
  if (!timing) {
    if (packed) attention_occ_kernel<false, true><<<g, THREADS, SMEM, stream>>>(tq, tkt, tkp, tvt, tvp, to, p);
    else attention_occ_kernel<false, false><<<g, THREADS, SMEM, stream>>>(tq, tkt, tkp, tvt, tvp, to, p);
    return cudaGetLastError();
  }
  p.dbg = dbg;
  p.knock = knock;
  cudaMemsetAsync(dbg, 0, 64 * sizeof(unsigned long long), stream);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, stream);
  attention_occ_kernel<true, true><<<g, THREADS, SMEM, stream>>>(tq, tkt, tkp, tvt, tvp, to, p);
  cudaEventRecord(e1, stream);
  unsigned long long hh[64];
  float ms = 0.0f;
  if (cudaStreamSynchronize(stream) == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (cudaMemcpy(hh, dbg, sizeof(hh), cudaMemcpyDeviceToHost) == cudaSuccess) {
    fprintf(stderr, "occ timing, knock %d: %.1f us\n", p.knock, ms * 1e3);
    for (int w = 0; w < 4; ++w) {
      const unsigned long long* h = hh + w * 16;
      if (h[0] == 0 || h[6] == 0) continue;
      const double nb = (double)h[0], ni = (double)h[6], nm = (double)(h[5] ? h[5] : 1);
      fprintf(stderr, "  warp %d: %llu blocks (%llu masked), %llu items; cycles/block: wait_s %.0f load+max %.0f exp %.0f "
              "st_wait+arrive %.0f; cycles/item: wait_s(first block) %.0f wait_pv %.0f output %.0f; masked blocks: "
              "max %.0f exp %.0f\n", 4 + w, h[0], h[5], h[6], h[1] / nb, h[2] / nb, h[3] / nb, h[4] / nb,
              h[10] / ni, h[7] / ni, h[8] / ni, h[12] / nm, h[13] / nm);
    }
  }
  return cudaGetLastError();
}

}  // namespace bma
