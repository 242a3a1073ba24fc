// Memory-bound helpers of the candidate-scoring path:
//   build_tail_embeds   per-candidate tail = embedding rows of the sampled suffix ids + the fixed
//                       (after / target / ...) rows shared by all candidates; replaces the
//                       reference's sw-fold repeat+cat of the whole prompt (bimodal_attack.py:1217-1225)
//   rmsnorm             RMSNorm with HF's rounding points, optional row gather (target rows only)
//   ce_finalize         combine the LM-head GEMM's per-tile (max, sum-exp) partials into per-token
//                       CE, mean per candidate (bimodal_attack.py:1293-1298), packed argmin key
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace bma {

// ------------------------------------------------------------------------------------------------
// Suffix-head sharing. A candidate equals the current suffix
// ("base") up to its first replaced position j, so its tail rows [0, j) are the base's rows: they
// are computed once (the base rides along as an extra candidate) and their K/V are appended to the
// shared prefix. Candidate c then owns only rows [j_c, S): the flattened row space is ragged.
//   row_info[r] = (candidate index in ids, or -1 for the base; tail position s)
//   row_meta[r] = (first flattened row of the candidate; number of prefix keys the row may see,
//                  = P + j_c)
//   cand_vstart[c] = first row of candidate c minus j_c  (row of tail position s is vstart + s)
// Candidates that are IDENTICAL to an earlier candidate of the chunk (or to the base) own no rows at
// all: their cand_vstart points at the representative's rows, so the LM head scores the same hidden
// states again and their loss is the representative's (the reference scores duplicates redundantly;
// at sw 512 about 5 % of the draws repeat a (position, token) pair).
// One block; candidates are spread over the threads, rows are numbered by a block-wide scan.
__global__ void __launch_bounds__(1024)
plan_rows_kernel(const long long* __restrict__ ids, int ids_stride, const long long* __restrict__ base_ids,
                 const int* __restrict__ slot, int S, int n_share, int B, int include_base, int P, int dedupe,
                 int pre0, int pre_n,
                 int2* __restrict__ row_info, int2* __restrict__ row_meta, int* __restrict__ cand_vstart,
                 int2* __restrict__ cand_tmp, int* __restrict__ total_rows,
                 unsigned long long* __restrict__ rows_accum) {
  __shared__ int part[1024];
  const int t = threadIdx.x, nt = blockDim.x;
  // Fused prefill: rows [0, pre_n) are the shared-prefix positions [pre0, pre0 + pre_n) themselves
  // (candidate -2, tail position relative to P so the rotary position pos_base + y comes out right).
  // Their K/V reach the prefix buffers before this layer's attention, so row i sees prefix keys
  // [0, i) there and itself as its only tail key.
  for (int i = t; i < pre_n; i += nt) {
    row_info[i] = make_int2(-2, pre0 + i - P);
    row_meta[i] = make_int2(i, pre0 + i);
  }
  const int nbase = include_base ? 1 : 0;
  const int n = B + nbase;                               // entry 0 is the base when included
  const int per = (n + nt - 1) / nt;
  const int e0 = min(n, t * per), e1 = min(n, e0 + per);
  // (first differing tail position, token there) of every entry
  for (int e = e0; e < e1; ++e) {
    int jd = n_share, tk = -1;
    if (e < nbase) {
      jd = 0;
    } else {
      const long long* row = ids + (long long)(e - nbase) * ids_stride;
      for (int s = 0; s < n_share; ++s) {
        const int sl = slot[s];
        if (sl >= 0 && row[sl] != base_ids[sl]) { jd = s; tk = static_cast<int>(row[sl]); break; }
      }
    }
    cand_tmp[e] = make_int2(jd, tk);
  }
  __syncthreads();
  // representative of every entry (itself unless an identical earlier entry exists)
  int local = 0;
  for (int e = e0; e < e1; ++e) {
    const int2 me = cand_tmp[e];
    int rep = e;
    if (dedupe && e >= nbase) {
      const long long* row = ids + (long long)(e - nbase) * ids_stride;
      auto same_from = [&](const long long* other, int s0) {
        for (int s = s0; s < S; ++s) {
          const int sl = slot[s];
          if (sl >= 0 && row[sl] != other[sl]) return false;
        }
        return true;
      };
      if (me.x == n_share && nbase == 1 && same_from(base_ids, n_share)) {
        rep = 0;                                         // the candidate IS the base
      } else if (me.x < n_share) {
        for (int o = nbase; o < e; ++o) {
          const int2 ot = cand_tmp[o];
          if (ot.x == me.x && ot.y == me.y &&
              same_from(ids + (long long)(o - nbase) * ids_stride, me.x + 1)) { rep = o; break; }
        }
      }
    }
    cand_vstart[e] = rep;                                // staged: entry -> representative entry
    local += rep == e ? S - me.x : 0;
  }
  part[t] = local;
  __syncthreads();
  for (int off = 1; off < nt; off <<= 1) {               // inclusive Hillis-Steele scan
    const int v = t >= off ? part[t - off] : 0;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  int row = pre_n + part[t] - local;
  if (t == nt - 1) {
    *total_rows = pre_n + part[t];
    if (rows_accum != nullptr) *rows_accum += static_cast<unsigned long long>(part[t]);   // tail rows only
  }
  for (int e = e0; e < e1; ++e) {
    if (cand_vstart[e] != e) continue;                   // duplicates are resolved below
    const int jd = cand_tmp[e].x;
    const int c = e - nbase;                             // -1 = base
    cand_tmp[e].y = row - jd;                            // virtual start of this representative
    for (int s = jd; s < S; ++s) {
      row_info[row + s - jd] = make_int2(c, s);
      row_meta[row + s - jd] = make_int2(row, P + jd);
    }
    row += S - jd;
  }
  __syncthreads();
  // cand_vstart is indexed by CANDIDATE (entry - nbase) for the CE gather; entries were staged in
  // place, so resolve through a register and re-synchronise before the shifted write
  int vs[8];
  for (int e = e0, k = 0; e < e1 && k < 8; ++e, ++k) vs[k] = cand_tmp[cand_vstart[e]].y;
  __syncthreads();
  for (int e = e0, k = 0; e < e1 && k < 8; ++e, ++k)
    if (e >= nbase) cand_vstart[e - nbase] = vs[k];
}

cudaError_t launch_plan_rows(const long long* ids, int ids_stride, const long long* base_ids,
                             const int* slot, int S, int n_share, int B, int include_base, int P, int dedupe,
                             int pre0, int pre_n,
                             int2* row_info, int2* row_meta, int* cand_vstart, int2* cand_tmp,
                             int* total_rows, unsigned long long* rows_accum, cudaStream_t stream) {
  if (B + 1 > 8 * 1024) return cudaErrorInvalidValue;    // 8 entries per thread at most
  plan_rows_kernel<<<1, 1024, 0, stream>>>(ids, ids_stride, base_ids, slot, S, n_share, B, include_base, P,
                                           dedupe, pre0, pre_n, row_info, row_meta, cand_vstart, cand_tmp,
                                           total_rows, rows_accum);
  return cudaGetLastError();
}

// Appends the base's K rows / V^T columns [0, n) of one layer to the shared-prefix buffers at key
// index P.. (K: row-major [key][kv_dim] slices with row strides ldk / ldkp; V^T: [kv_dim][key]).
__global__ void __launch_bounds__(256)
append_prefix_kv_kernel(const __nv_bfloat16* __restrict__ k_tail, long long ldk,
                        const __nv_bfloat16* __restrict__ vt_tail, long long ldvt,
                        __nv_bfloat16* __restrict__ k_pre, long long ldkp,
                        __nv_bfloat16* __restrict__ vt_pre, long long ldvtp, int n, int kv_dim, int P, int vec_v) {
  // 16-byte units: K rows are contiguous along the columns, V^T rows along the keys (vector stores when
  // the destination key index, both leading dimensions and both base pointers are 8-aligned: vec_v)
  const int kvec = kv_dim >> 3, nvec = (n + 7) >> 3;
  const long long k_units = (long long)n * kvec;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < k_units) {
    const int key = static_cast<int>(i / kvec), c = static_cast<int>(i - (long long)key * kvec) * 8;
    *reinterpret_cast<uint4*>(k_pre + (long long)(P + key) * ldkp + c) =
        *reinterpret_cast<const uint4*>(k_tail + (long long)key * ldk + c);
  } else if (i < k_units + (long long)kv_dim * nvec) {
    const long long j = i - k_units;
    const int dim = static_cast<int>(j / nvec), kk = static_cast<int>(j - (long long)dim * nvec) * 8;
    const __nv_bfloat16* src = vt_tail + (long long)dim * ldvt + kk;
    __nv_bfloat16* dst = vt_pre + (long long)dim * ldvtp + P + kk;
    if (vec_v && kk + 8 <= n) {
      *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(src);
    } else {
      for (int e = 0; e < 8 && kk + e < n; ++e) dst[e] = src[e];
    }
  }
}

cudaError_t launch_append_prefix_kv(const __nv_bfloat16* k_tail, long long ldk, const __nv_bfloat16* vt_tail,
                                    long long ldvt, __nv_bfloat16* k_pre, long long ldkp,
                                    __nv_bfloat16* vt_pre, long long ldvtp, int n, int kv_dim, int P,
                                    cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  if ((kv_dim & 7) || (ldk & 7) || (ldkp & 7) || (reinterpret_cast<uintptr_t>(k_tail) & 15) ||
      (reinterpret_cast<uintptr_t>(k_pre) & 15))
    return cudaErrorInvalidValue;
  const int vec_v = ((P & 7) == 0 && (ldvt & 7) == 0 && (ldvtp & 7) == 0 && (reinterpret_cast<uintptr_t>(vt_tail) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(vt_pre) & 15) == 0) ? 1 : 0;
  const long long total = (long long)n * (kv_dim >> 3) + (long long)kv_dim * ((n + 7) >> 3);
  append_prefix_kv_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, stream>>>(k_tail, ldk, vt_tail, ldvt, k_pre, ldkp,
                                                                                      vt_pre, ldvtp, n, kv_dim, P, vec_v);
  return cudaGetLastError();
}

// Split prefill: ONE launch scatters every rank's gathered block ([chunk x kv_dim K rows | kv_dim x chunk
// V^T], `blk` elements apart in `recv`) into prefix keys [q chunk, min(P, (q+1) chunk)) and appends the
// base suffix's n_base K/V rows (tail rows base_row0.. of this rank) at key index P.
__global__ void __launch_bounds__(256)
unpack_prefix_kv_kernel(const __nv_bfloat16* __restrict__ recv, long long blk, int chunk, int n_ranks, int P,
                        const __nv_bfloat16* __restrict__ k_tail, long long ldk,
                        const __nv_bfloat16* __restrict__ vt_tail, long long ldvt, int base_row0, int n_base,
                        __nv_bfloat16* __restrict__ k_pre, long long ldkp,
                        __nv_bfloat16* __restrict__ vt_pre, long long ldvtp, int kv_dim) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n_keys = (long long)P + n_base;
  if (i >= n_keys * kv_dim) return;
  {                                                   // K: consecutive threads -> consecutive columns
    const int key = static_cast<int>(i / kv_dim), c = static_cast<int>(i - (long long)key * kv_dim);
    __nv_bfloat16 v;
    if (key < P) {
      const int q = key / chunk, r = key - q * chunk;
      v = recv[q * blk + (long long)r * kv_dim + c];
    } else {
      v = k_tail[(long long)(base_row0 + key - P) * ldk + c];
    }
    k_pre[(long long)key * ldkp + c] = v;
  }
  {                                                   // V^T: consecutive threads -> consecutive keys
    const int dim = static_cast<int>(i / n_keys), key = static_cast<int>(i - (long long)dim * n_keys);
    __nv_bfloat16 v;
    if (key < P) {
      const int q = key / chunk, r = key - q * chunk;
      v = recv[q * blk + (long long)chunk * kv_dim + (long long)dim * chunk + r];
    } else {
      v = vt_tail[(long long)dim * ldvt + base_row0 + key - P];
    }
    vt_pre[(long long)dim * ldvtp + key] = v;
  }
  (void)n_ranks;
}

cudaError_t launch_unpack_prefix_kv(const __nv_bfloat16* recv, long long blk, int chunk, int n_ranks, int P,
                                    const __nv_bfloat16* k_tail, long long ldk, const __nv_bfloat16* vt_tail,
                                    long long ldvt, int base_row0, int n_base, __nv_bfloat16* k_pre, long long ldkp,
                                    __nv_bfloat16* vt_pre, long long ldvtp, int kv_dim, cudaStream_t stream) {
  const long long total = ((long long)P + n_base) * kv_dim;
  if (total <= 0) return cudaSuccess;
  unpack_prefix_kv_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, stream>>>(
      recv, blk, chunk, n_ranks, P, k_tail, ldk, vt_tail, ldvt, base_row0, n_base, k_pre, ldkp, vt_pre, ldvtp, kv_dim);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Split prefill over NVLink peer memory. Every rank owns the K / V^T of `n` prefix positions starting at
// key `key0` (tail rows [0, n) of its pass) and STORES them straight into the per-layer prefix buffers
// of every rank of the box -- its own included -- through peer-mapped pointers (st.global over
// NVLink 5 / NVSwitch); no staging buffer, no collective. The last block to finish fences at system
// scope and raises this rank's flag on every peer; prefix_wait_kernel then spins (bounded, then traps)
// until all ranks' flags have reached the epoch of this exchange.
struct PeerPtrs {
  __nv_bfloat16* k[8];              // per rank: K columns of this layer's prefix buffer  ([key][ldkp])
  __nv_bfloat16* vt[8];             // per rank: V^T of this layer's prefix buffer        ([dim][ldvtp])
  unsigned long long* flags[8];     // per rank: its flag array, one word per sending rank
};

__global__ void __launch_bounds__(256)
prefix_scatter_kernel(const __nv_bfloat16* __restrict__ k_tail, long long ldk, const __nv_bfloat16* __restrict__ vt_tail,
                      long long ldvt, PeerPtrs peers, int n_ranks, int my_rank, long long ldkp, long long ldvtp,
                      int n, int key0, int kv_dim, unsigned int* done_counter, unsigned long long epoch) {
  // 16-byte units: K rows are contiguous along the columns; V^T rows along the keys (8 keys per unit,
  // vector stores when the destination key index is 8-aligned -- the host rounds the per-rank slice to 8)
  const int kvec = kv_dim >> 3, nvec = (n + 7) >> 3;
  const long long k_units = (long long)n * kvec, v_units = (long long)kv_dim * nvec;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < k_units) {
    const int key = static_cast<int>(i / kvec), c = static_cast<int>(i - (long long)key * kvec) * 8;
    const uint4 v = *reinterpret_cast<const uint4*>(k_tail + (long long)key * ldk + c);
    for (int q = 0; q < n_ranks; ++q)
      *reinterpret_cast<uint4*>(peers.k[q] + (long long)(key0 + key) * ldkp + c) = v;
  } else if (i < k_units + v_units) {
    const long long j = i - k_units;
    const int dim = static_cast<int>(j / nvec), kk = static_cast<int>(j - (long long)dim * nvec) * 8;
    const __nv_bfloat16* src = vt_tail + (long long)dim * ldvt + kk;
    if (kk + 8 <= n && ((key0 | ldvtp | ldvt) & 7) == 0) {
      const uint4 v = *reinterpret_cast<const uint4*>(src);
      for (int q = 0; q < n_ranks; ++q)
        *reinterpret_cast<uint4*>(peers.vt[q] + (long long)dim * ldvtp + key0 + kk) = v;
    } else {
      for (int e = 0; e < 8 && kk + e < n; ++e) {
        const __nv_bfloat16 v = src[e];
        for (int q = 0; q < n_ranks; ++q) peers.vt[q][(long long)dim * ldvtp + key0 + kk + e] = v;
      }
    }
  }
  // publish: every block's stores are fenced system-wide before it counts itself done; the last one signals
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(done_counter, 1u);
    if (prev == gridDim.x - 1) {
      *done_counter = 0;                                     // ready for the next exchange
      __threadfence_system();
      for (int q = 0; q < n_ranks; ++q)
        atomicExch_system(peers.flags[q] + my_rank, epoch);
    }
  }
}

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// wait_ns (optional): nanoseconds this rank spent waiting for the other ranks' flags, accumulated over the
// exchanges (bma_scorer_exchange_wait_ns): a rank that never waits is the one the others wait for.
__global__ void prefix_wait_kernel(volatile unsigned long long* flags, int n_ranks, unsigned long long epoch,
                                   unsigned long long* wait_ns) {
  const int q = threadIdx.x;
  const unsigned long long t0 = (wait_ns != nullptr && q == 0) ? global_ns() : 0ull;
  if (q < n_ranks) {
    unsigned int spins = 0;
    while (flags[q] < epoch) {
      if (++spins > (1u << 27)) {                            // ~seconds: a rank never arrived
        if (g_watchdog_host != nullptr) { *g_watchdog_host = 0xE000u | q; __threadfence_system(); }
        __trap();
      }
    }
  }
  __syncwarp();
  if (wait_ns != nullptr && q == 0) atomicAdd(wait_ns, global_ns() - t0);
  __threadfence_system();
}

// Fused publish (the QKV GEMM's epilogue has stored this rank's K/V slice into every rank's buffers): raise
// this rank's flag on every peer, then wait for all ranks' flags. The GEMM precedes this launch in stream
// order, so its peer stores are performed before the system-scope fence below, and the flag after it.
struct PeerFlags { unsigned long long* f[8]; };
__global__ void prefix_signal_wait_kernel(PeerFlags peers, volatile unsigned long long* flags, int n_ranks, int my_rank,
                                          unsigned long long epoch, int signal, int wait, unsigned long long* wait_ns) {
  const int q = threadIdx.x;
  unsigned long long t0 = 0ull;
  if (q < n_ranks) {
    if (signal) {
      __threadfence_system();
      atomicExch_system(peers.f[q] + my_rank, epoch);
    }
    if (wait_ns != nullptr && q == 0) t0 = global_ns();
    if (wait) {
      unsigned int spins = 0;
      while (flags[q] < epoch) {
        if (++spins > (1u << 27)) {                          // ~seconds: a rank never arrived
          if (g_watchdog_host != nullptr) { *g_watchdog_host = 0xE100u | q; __threadfence_system(); }
          __trap();
        }
      }
    }
  }
  __syncwarp();
  if (wait && wait_ns != nullptr && q == 0) atomicAdd(wait_ns, global_ns() - t0);
  __threadfence_system();
}

cudaError_t launch_prefix_signal_wait(void* const* peer_flags, unsigned long long* my_flags, int n_ranks, int my_rank,
                                      unsigned long long epoch, int signal, int wait, cudaStream_t stream,
                                      unsigned long long* wait_ns) {
  if (n_ranks < 1 || n_ranks > 8) return cudaErrorInvalidValue;
  PeerFlags pf{};
  for (int q = 0; q < n_ranks; ++q) pf.f[q] = static_cast<unsigned long long*>(peer_flags[q]);
  prefix_signal_wait_kernel<<<1, 32, 0, stream>>>(pf, my_flags, n_ranks, my_rank, epoch, signal, wait, wait_ns);
  return cudaGetLastError();
}

cudaError_t launch_prefix_scatter(const __nv_bfloat16* k_tail, long long ldk, const __nv_bfloat16* vt_tail, long long ldvt,
                                  void* const* peer_k, void* const* peer_vt, void* const* peer_flags, int n_ranks,
                                  int my_rank, long long ldkp, long long ldvtp, int n, int key0, int kv_dim,
                                  unsigned int* done_counter, unsigned long long epoch, cudaStream_t stream) {
  if (n_ranks < 1 || n_ranks > 8) return cudaErrorInvalidValue;
  PeerPtrs pp{};
  for (int q = 0; q < n_ranks; ++q) {
    pp.k[q] = static_cast<__nv_bfloat16*>(peer_k[q]);
    pp.vt[q] = static_cast<__nv_bfloat16*>(peer_vt[q]);
    pp.flags[q] = static_cast<unsigned long long*>(peer_flags[q]);
  }
  if ((kv_dim & 7) || (ldk & 7) || (ldkp & 7)) return cudaErrorInvalidValue;
  const int nn = n > 0 ? n : 0;
  const long long total = (long long)nn * (kv_dim >> 3) + (long long)kv_dim * ((nn + 7) >> 3);
  const int blocks = total > 0 ? static_cast<int>((total + 255) / 256) : 1;   // a rank without positions still signals
  prefix_scatter_kernel<<<blocks, 256, 0, stream>>>(k_tail, ldk, vt_tail, ldvt, pp, n_ranks, my_rank, ldkp, ldvtp,
                                                    n > 0 ? n : 0, key0, kv_dim, done_counter, epoch);
  return cudaGetLastError();
}

cudaError_t launch_prefix_wait(unsigned long long* flags, int n_ranks, unsigned long long epoch, cudaStream_t stream,
                               unsigned long long* wait_ns) {
  prefix_wait_kernel<<<1, 32, 0, stream>>>(flags, n_ranks, epoch, wait_ns);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// out[row, :] = slot[s] >= 0 ? round_bf16(W[ids[b, slot[s]], :] * scale) : fixed[s, :]
// Uniform layout (row_info == null): row = b*S + s. Ragged layout: (b, s) = row_info[row], b = -1
// reads the base ids.
__global__ void __launch_bounds__(256)
build_tail_embeds_kernel(const long long* __restrict__ ids, int ids_stride,
                         const int* __restrict__ slot, const __nv_bfloat16* __restrict__ fixed,
                         const __nv_bfloat16* __restrict__ W, long long ldw, int V, float scale,
                         __nv_bfloat16* __restrict__ out, int rows, int S, int d,
                         const int2* __restrict__ row_info, const long long* __restrict__ base_ids,
                         const __nv_bfloat16* __restrict__ prefix_embeds, int P,
                         const int* __restrict__ rows_dev) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows || (rows_dev != nullptr && warp >= *rows_dev)) return;
  int b, s;
  if (row_info != nullptr) {
    const int2 ri = row_info[warp];
    b = ri.x; s = ri.y;
  } else {
    b = warp / S; s = warp - b * S;
  }
  const int nvec = d >> 3;
  uint4* dst = reinterpret_cast<uint4*>(out + (long long)warp * d);
  if (b == -2) {                                   // a shared-prefix position (fused prefill)
    const uint4* src = reinterpret_cast<const uint4*>(prefix_embeds + (long long)(P + s) * d);
    for (int v = lane; v < nvec; v += 32) dst[v] = src[v];
    return;
  }
  const int sl = slot[s];
  if (sl < 0) {
    const uint4* src = reinterpret_cast<const uint4*>(fixed + (long long)s * d);
    for (int v = lane; v < nvec; v += 32) dst[v] = src[v];
    return;
  }
  long long tok = b < 0 ? base_ids[sl] : ids[(long long)b * ids_stride + sl];
  if (tok < 0 || tok >= V) tok = 0;   // never read outside the table; host validates ids
  const uint4* src = reinterpret_cast<const uint4*>(W + tok * ldw);
  if (scale == 1.0f) {
    for (int v = lane; v < nvec; v += 32) dst[v] = src[v];
  } else {
    for (int v = lane; v < nvec; v += 32) {
      const uint4 q = src[v];
      const uint32_t in[4] = {q.x, q.y, q.z, q.w};
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = pack_bf16x2(bf16_lo(in[j]) * scale, bf16_hi(in[j]) * scale);
      dst[v] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

cudaError_t launch_build_tail_embeds(const long long* ids, int ids_stride, const int* slot,
                                     const __nv_bfloat16* fixed, const __nv_bfloat16* W,
                                     long long ldw, int V, float scale, __nv_bfloat16* out, int rows,
                                     int S, int d, const int2* row_info, const long long* base_ids,
                                     const __nv_bfloat16* prefix_embeds, int P, const int* rows_dev,
                                     cudaStream_t stream) {
  if (rows <= 0 || S <= 0) return cudaSuccess;
  if ((d & 7) != 0) return cudaErrorInvalidValue;
  const int blocks = static_cast<int>(((long long)rows * 32 + 255) / 256);
  build_tail_embeds_kernel<<<blocks, 256, 0, stream>>>(ids, ids_stride, slot, fixed, W, ldw, V,
                                                       scale, out, rows, S, d, row_info, base_ids,
                                                       prefix_embeds, P, rows_dev);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
constexpr int kNormRegVecs = 16;     // 16-byte vectors a lane keeps in registers: rows of up to 4096 elements

// RMSNorm, one warp per output row. mode 0 (Llama): y = bf16(w * bf16(x * rstd));
// mode 1 (Gemma): y = bf16(x * rstd * (1 + w)). Row gather: when gather_T > 0 output row r reads
// input row start(r / gather_T) + gather_off + (r % gather_T), where start(c) = c * gather_S for
// the uniform layout and gather_vstart[c] for the ragged one.
template <int MINB>
__global__ void __launch_bounds__(128, MINB)
rmsnorm_kernel(const __nv_bfloat16* __restrict__ x, long long ldx,
               const __nv_bfloat16* __restrict__ w, __nv_bfloat16* __restrict__ y, long long ldy,
               int rows, int d, float eps, int mode, int gather_T, int gather_S, int gather_off,
               const int* __restrict__ gather_vstart, const int* __restrict__ rows_dev) {
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= rows || (rows_dev != nullptr && r >= *rows_dev)) return;
  long long src_row = r;
  if (gather_T > 0) {
    const int c = r / gather_T;
    const long long start = gather_vstart != nullptr ? gather_vstart[c] : (long long)c * gather_S;
    src_row = start + gather_off + (r - c * gather_T);
  }
  const uint4* src = reinterpret_cast<const uint4*>(x + src_row * ldx);
  const uint4* wv = reinterpret_cast<const uint4*>(w);
  uint4* dst = reinterpret_cast<uint4*>(y + (long long)r * ldy);
  const int nvec = d >> 3;
  if (nvec <= 32 * kNormRegVecs) {
    // Row in registers (d <= 4096): x is read ONCE, all of a lane's 16-byte loads are in flight together.
    uint4 q[kNormRegVecs];
#pragma unroll
    for (int i = 0; i < kNormRegVecs; ++i) {
      const int v = lane + 32 * i;
      q[i] = v < nvec ? src[v] : make_uint4(0u, 0u, 0u, 0u);
    }
    float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
    for (int i = 0; i < kNormRegVecs; ++i) {
      const uint32_t in[4] = {q[i].x, q[i].y, q[i].z, q[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float a = bf16_lo(in[j]), b = bf16_hi(in[j]);
        s0 = fmaf(a, a, s0);
        s1 = fmaf(b, b, s1);
      }
    }
    const float rstd = rsqrtf(warp_sum(s0 + s1) / static_cast<float>(d) + eps);
#pragma unroll
    for (int i = 0; i < kNormRegVecs; ++i) {
      const int v = lane + 32 * i;
      if (v < nvec) {
        const uint4 ww = wv[v];
        const uint32_t in[4] = {q[i].x, q[i].y, q[i].z, q[i].w};
        const uint32_t wi[4] = {ww.x, ww.y, ww.z, ww.w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float a = bf16_lo(in[j]) * rstd, b = bf16_hi(in[j]) * rstd;
          if (mode == 0) {
            // HF rounds x * rstd to bf16 before the weight multiply: both values through ONE packed conversion
            // (F2FP, ALU pipe) instead of two scalar F2F round trips (XU pipe, quarter rate)
            const uint32_t pr = pack_bf16x2(a, b);
            a = bf16_lo(pr) * bf16_lo(wi[j]);
            b = bf16_hi(pr) * bf16_hi(wi[j]);
          } else {
            a = a * (1.0f + bf16_lo(wi[j]));
            b = b * (1.0f + bf16_hi(wi[j]));
          }
          o[j] = pack_bf16x2(a, b);
        }
        dst[v] = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
    return;
  }
  float ss = 0.0f;
  for (int v = lane; v < nvec; v += 32) {
    const uint4 q = src[v];
    const uint32_t in[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = bf16_lo(in[j]), b = bf16_hi(in[j]);
      ss += a * a + b * b;
    }
  }
  ss = warp_sum(ss);
  const float rstd = rsqrtf(ss / static_cast<float>(d) + eps);
  for (int v = lane; v < nvec; v += 32) {
    const uint4 q = src[v];
    const uint4 ww = wv[v];
    const uint32_t in[4] = {q.x, q.y, q.z, q.w};
    const uint32_t wi[4] = {ww.x, ww.y, ww.z, ww.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float a = bf16_lo(in[j]) * rstd, b = bf16_hi(in[j]) * rstd;
      if (mode == 0) {
        a = bf16_round(a) * bf16_lo(wi[j]);
        b = bf16_round(b) * bf16_hi(wi[j]);
      } else {
        a = a * (1.0f + bf16_lo(wi[j]));
        b = b * (1.0f + bf16_hi(wi[j]));
      }
      o[j] = pack_bf16x2(a, b);
    }
    dst[v] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

cudaError_t launch_rmsnorm(const __nv_bfloat16* x, long long ldx, const __nv_bfloat16* w,
                           __nv_bfloat16* y, long long ldy, int rows, int d, float eps, int mode,
                           int gather_T, int gather_S, int gather_off, const int* gather_vstart,
                           cudaStream_t stream, const int* rows_dev) {
  if (rows <= 0) return cudaSuccess;
  if ((d & 7) != 0 || (ldx & 7) != 0 || (ldy & 7) != 0) return cudaErrorInvalidValue;
  const int blocks = static_cast<int>(((long long)rows * 32 + 127) / 128);      // 4 rows per block
  // BMA_NORM_MINB = 3 (150 registers, 12 resident warps per SM) | 4 (128 registers, 16 warps); A/B switch,
  // re-read per launch when BMA_NORM_RELOAD = 1
  static int minb = 0, reload = 0;
  if (minb == 0 || reload) {
    const char* e = getenv("BMA_NORM_MINB");
    minb = (e != nullptr && e[0] == '3') ? 3 : 4;
    e = getenv("BMA_NORM_RELOAD");
    reload = (e != nullptr && e[0] == '1');
  }
  if (minb == 3)
    rmsnorm_kernel<3><<<blocks, 128, 0, stream>>>(x, ldx, w, y, ldy, rows, d, eps, mode, gather_T,
                                                  gather_S, gather_off, gather_vstart, rows_dev);
  else
    rmsnorm_kernel<4><<<blocks, 128, 0, stream>>>(x, ldx, w, y, ldy, rows, d, eps, mode, gather_T,
                                                  gather_S, gather_off, gather_vstart, rows_dev);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Row gather: dst[r, :] = src[start(r / T) + off + r % T, :] with start(c) = c * S (uniform tails)
// or vstart[c] (ragged tails) -- the T scored rows of every candidate, made contiguous.
__global__ void __launch_bounds__(256)
gather_rows_kernel(const __nv_bfloat16* __restrict__ src, long long lds, __nv_bfloat16* __restrict__ dst,
                   long long ldd, int rows, int width, int T, int S, int off, const int* __restrict__ vstart) {
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const int c = r / T;
  const long long start = vstart != nullptr ? vstart[c] : (long long)c * S;
  const uint4* s4 = reinterpret_cast<const uint4*>(src + (start + off + (r - c * T)) * lds);
  uint4* d4 = reinterpret_cast<uint4*>(dst + (long long)r * ldd);
  for (int v = lane; v < (width >> 3); v += 32) d4[v] = s4[v];
}

cudaError_t launch_gather_rows(const __nv_bfloat16* src, long long lds, __nv_bfloat16* dst, long long ldd,
                               int rows, int width, int T, int S, int off, const int* vstart,
                               cudaStream_t stream) {
  if (rows <= 0) return cudaSuccess;
  if ((width & 7) != 0 || (lds & 7) != 0 || (ldd & 7) != 0 || T <= 0) return cudaErrorInvalidValue;
  const int blocks = static_cast<int>(((long long)rows * 32 + 255) / 256);
  gather_rows_kernel<<<blocks, 256, 0, stream>>>(src, lds, dst, ldd, rows, width, T, S, off, vstart);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Sandwich-norm step of a Gemma-style layer, one warp per row:
//   x  <- bf16(x + norm(y; w_post))            (post_attention / post_feedforward norm + residual)
//   xn <- norm(x; w_next)   if w_next != null  (the next pre-norm: pre_feedforward / input norm)
// y and xn may alias (each lane rewrites only elements it has already consumed).
__device__ __forceinline__ float norm_apply(float v, float rstd, float w, int mode) {
  return mode == 0 ? bf16_round(v * rstd) * w : (v * rstd) * (1.0f + w);
}

__global__ void __launch_bounds__(256)
add_norm_kernel(const __nv_bfloat16* y, long long ldy, const __nv_bfloat16* __restrict__ w_post,
                __nv_bfloat16* x, long long ldx, const __nv_bfloat16* __restrict__ w_next,
                __nv_bfloat16* xn, long long ldxn, int rows, int d, float eps, int mode,
                const int* __restrict__ rows_dev) {
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= rows || (rows_dev != nullptr && r >= *rows_dev)) return;
  const uint4* yv = reinterpret_cast<const uint4*>(y + (long long)r * ldy);
  uint4* xv = reinterpret_cast<uint4*>(x + (long long)r * ldx);
  const int nvec = d >> 3;
  float ss = 0.0f;
  for (int v = lane; v < nvec; v += 32) {
    const uint4 q = yv[v];
    const uint32_t in[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = bf16_lo(in[j]), b = bf16_hi(in[j]);
      ss += a * a + b * b;
    }
  }
  ss = warp_sum(ss);
  const float rstd_y = rsqrtf(ss / static_cast<float>(d) + eps);
  const uint4* wp = reinterpret_cast<const uint4*>(w_post);
  float sx = 0.0f;
  for (int v = lane; v < nvec; v += 32) {
    const uint4 q = yv[v];
    const uint4 xx = xv[v];
    const uint4 ww = wp[v];
    const uint32_t in[4] = {q.x, q.y, q.z, q.w};
    const uint32_t xi[4] = {xx.x, xx.y, xx.z, xx.w};
    const uint32_t wi[4] = {ww.x, ww.y, ww.z, ww.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = bf16_round(bf16_lo(xi[j]) + bf16_round(norm_apply(bf16_lo(in[j]), rstd_y, bf16_lo(wi[j]), mode)));
      const float b = bf16_round(bf16_hi(xi[j]) + bf16_round(norm_apply(bf16_hi(in[j]), rstd_y, bf16_hi(wi[j]), mode)));
      sx += a * a + b * b;
      o[j] = pack_bf16x2(a, b);
    }
    xv[v] = make_uint4(o[0], o[1], o[2], o[3]);
  }
  if (w_next == nullptr) return;
  sx = warp_sum(sx);
  const float rstd_x = rsqrtf(sx / static_cast<float>(d) + eps);
  const uint4* wn = reinterpret_cast<const uint4*>(w_next);
  uint4* dst = reinterpret_cast<uint4*>(xn + (long long)r * ldxn);
  for (int v = lane; v < nvec; v += 32) {
    const uint4 xx = xv[v];                     // this lane's own stores above
    const uint4 ww = wn[v];
    const uint32_t xi[4] = {xx.x, xx.y, xx.z, xx.w};
    const uint32_t wi[4] = {ww.x, ww.y, ww.z, ww.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      o[j] = pack_bf16x2(norm_apply(bf16_lo(xi[j]), rstd_x, bf16_lo(wi[j]), mode),
                         norm_apply(bf16_hi(xi[j]), rstd_x, bf16_hi(wi[j]), mode));
    dst[v] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

cudaError_t launch_add_norm(const __nv_bfloat16* y, long long ldy, const __nv_bfloat16* w_post,
                            __nv_bfloat16* x, long long ldx, const __nv_bfloat16* w_next,
                            __nv_bfloat16* xn, long long ldxn, int rows, int d, float eps, int mode,
                            cudaStream_t stream, const int* rows_dev) {
  if (rows <= 0) return cudaSuccess;
  if ((d & 7) != 0 || (ldx & 7) != 0 || (ldy & 7) != 0 || (ldxn & 7) != 0) return cudaErrorInvalidValue;
  const int blocks = static_cast<int>(((long long)rows * 32 + 255) / 256);
  add_norm_kernel<<<blocks, 256, 0, stream>>>(y, ldy, w_post, x, ldx, w_next, xn, ldxn, rows, d, eps, mode,
                                              rows_dev);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long pack_loss_key(float loss, unsigned int idx) {
  uint32_t u = __float_as_uint(loss);
  uint32_t ord = (loss != loss) ? 0xFFFFFFFFu : ((u & 0x80000000u) ? ~u : (u | 0x80000000u));
  return (static_cast<unsigned long long>(ord) << 31) | static_cast<unsigned long long>(idx & 0x7FFFFFFFu);
}

__global__ void init_key_kernel(unsigned long long* key, int* flag) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    if (key != nullptr) *key = 0x7FFFFFFFFFFFFFFFull;
    if (flag != nullptr) *flag = 0;
  }
}

// One block per candidate, one warp per target row at a time: the lanes stride over the n_tiles
// (max, sum-exp) partials of the row (part is [n_tiles][M]) and merge them with a fixed-order
// shuffle tree, so the result does not depend on scheduling. Thread 0 then averages the T token
// losses in order and publishes the candidate's loss and packed key.
__device__ __forceinline__ void lse_merge(float& m, float& s, float om, float os) {
  const float nm = fmaxf(m, om);
  if (nm == -INFINITY) return;
  s = s * __expf(m - nm) + os * __expf(om - nm);
  m = nm;
}

__global__ void __launch_bounds__(256)
ce_finalize_kernel(const float2* __restrict__ part, const float* __restrict__ tgt_logit,
                   const int* __restrict__ arg, const long long* __restrict__ target_ids,
                   int n_tiles, int B, int T, float* __restrict__ loss,
                   unsigned long long* __restrict__ key, int* __restrict__ all_match_flag,
                   unsigned int idx_base) {
  extern __shared__ float ce_row[];              // [T] token losses, then [T] match flags
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
  const long long M = (long long)B * T;
  for (int tt = warp; tt < T; tt += n_warps) {
    const long long row = (long long)b * T + tt;
    float m = -INFINITY, s = 0.0f;
    float am = -INFINITY;                        // running maximum and the FIRST column attaining it:
    int ai = 0x7fffffff;                         // torch.argmax's answer (reference :1300-1306)
    for (int nt = lane; nt < n_tiles; nt += 32) {
      const float2 ps = part[(long long)nt * M + row];
      lse_merge(m, s, ps.x, ps.y);
      if (arg != nullptr && ps.x > am) { am = ps.x; ai = arg[(long long)nt * M + row]; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, m, o);
      const float os = __shfl_xor_sync(0xffffffffu, s, o);
      lse_merge(m, s, om, os);
      const float oam = __shfl_xor_sync(0xffffffffu, am, o);
      const int oai = __shfl_xor_sync(0xffffffffu, ai, o);
      if (oam > am || (oam == am && oai < ai)) { am = oam; ai = oai; }
    }
    if (lane == 0) {
      const float tl = tgt_logit[row];
      ce_row[tt] = (m + logf(s)) - tl;
      const bool hit = arg != nullptr ? (ai == static_cast<int>(target_ids[tt])) : (tl >= m);
      ce_row[T + tt] = hit ? 1.0f : 0.0f;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float acc = 0.0f;
    bool all_match = true;
    for (int tt = 0; tt < T; ++tt) {
      acc += ce_row[tt];
      all_match = all_match && (ce_row[T + tt] != 0.0f);
    }
    const float l = acc / static_cast<float>(T);
    loss[b] = l;
    if (key != nullptr) atomicMin(key, pack_loss_key(l, idx_base + static_cast<unsigned int>(b)));
    if (all_match_flag != nullptr && all_match) atomicOr(all_match_flag, 1);
  }
}

cudaError_t launch_init_key(unsigned long long* key, int* flag, cudaStream_t stream) {
  init_key_kernel<<<1, 32, 0, stream>>>(key, flag);
  return cudaGetLastError();
}

cudaError_t launch_ce_finalize(const float2* part, const float* tgt_logit, const int* arg,
                               const long long* target_ids, int n_tiles, int B,
                               int T, float* loss, unsigned long long* key, int* all_match_flag,
                               unsigned int idx_base, cudaStream_t stream) {
  if (B <= 0) return cudaSuccess;
  ce_finalize_kernel<<<B, 256, 2 * T * sizeof(float), stream>>>(part, tgt_logit, arg, target_ids, n_tiles, B, T,
                                                                 loss, key, all_match_flag, idx_base);
  return cudaGetLastError();
}

}  // namespace bma
