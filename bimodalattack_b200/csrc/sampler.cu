// On-device GCG candidate sampler (reference bimodal_attack.py:130-163).
//
//   1. masked per-position top-k of -grad over the vocabulary        (topk_chunk + topk_merge)
//   2. n_replace distinct random positions per candidate + a random pick among the top-k at each
//      position, written straight into the candidate id matrix        (sample_replace)
//
// Step 2 consumes torch's CUDA Philox stream exactly as `torch.rand((sw, L))` followed by
// `torch.randint(0, topk, (sw, n_replace, 1))` would (ATen DistributionTemplates.h grid-stride
// kernel: 256-thread blocks, thread t <- curand_init(seed, t, offset), unroll 4), so that with the
// same (seed, offset) the candidate ids are bit-identical to the reference run on CUDA.
//
// Top-k keys: every element becomes a unique 64-bit key  (orderable(-g) << 32) | (~index), so a
// larger key means "earlier in torch.topk(-grad)" and equal gradients resolve to the lower token
// id. The not-allowed mask (reference: grad[:, not_allowed_ids] = inf) is a V-bit set folded into
// the read; the caller's gradient is not modified.
#include <cstdlib>

#include <cuda_fp16.h>
#include <curand_kernel.h>

#include "common.cuh"
#include "kernels.h"

namespace bma {

constexpr int TOPK_THREADS = 512;

__device__ __forceinline__ uint32_t orderable_neg(float g) {
  // key of f = -g such that unsigned order == float order; NaN sorts above everything (torch.topk)
  if (g != g) return 0xFFFFFFFFu;
  const uint32_t u = __float_as_uint(-g);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ uint64_t make_key(float g, uint32_t idx, const uint32_t* mask) {
  if (mask != nullptr && ((mask[idx >> 5] >> (idx & 31)) & 1u)) g = INFINITY;
  return (static_cast<uint64_t>(orderable_neg(g)) << 32) | static_cast<uint64_t>(0xFFFFFFFFu - idx);
}

// Histogram increment with warp aggregation: gradient rows put almost every key into two or three
// exponent buckets, so plain shared-memory atomics serialise; lanes that hit the same bin elect one
// leader that adds their count.
__device__ __forceinline__ void hist_add(uint32_t* hist, uint32_t bin) {
  const uint32_t peers = __match_any_sync(__activemask(), bin);
  if ((threadIdx.x & 31) == static_cast<uint32_t>(__ffs(peers) - 1)) atomicAdd(&hist[bin], __popc(peers));
}

// Select the k largest of n unique keys held in shared memory. On return *kth_out is a threshold
// such that exactly k keys are >= it. hist/ctl are shared scratch (256 + 4 words).
__device__ void radix_select_topk(const uint64_t* keys, int n, int k, uint32_t* hist,
                                  uint64_t* ctl /* [0]=prefix [1]=mask [2]=remaining [3]=done */) {
  if (threadIdx.x == 0) {
    ctl[0] = 0;
    ctl[1] = 0;
    ctl[2] = static_cast<uint64_t>(k);
    ctl[3] = 0;
  }
  __syncthreads();
  for (int pass = 0; pass < 8; ++pass) {
    const int shift = 56 - 8 * pass;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const uint64_t prefix = ctl[0], pmask = ctl[1];
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const uint64_t key = keys[i];
      if ((key & pmask) == prefix) hist_add(hist, static_cast<uint32_t>(key >> shift) & 255u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t remaining = static_cast<uint32_t>(ctl[2]);
      uint32_t above = 0;
      int d = 255;
      for (; d > 0; --d) {
        if (above + hist[d] >= remaining) break;
        above += hist[d];
      }
      remaining -= above;
      ctl[0] = prefix | (static_cast<uint64_t>(d) << shift);
      ctl[1] = pmask | (static_cast<uint64_t>(255) << shift);
      ctl[2] = remaining;
      // every key in this bucket is wanted: the prefix (low bits zero) is already the threshold
      if (hist[d] == remaining) ctl[3] = 1;
    }
    __syncthreads();
    if (ctl[3] != 0) break;
  }
}

template <typename T>
__device__ __forceinline__ float load_as_float(const T* p, long long i);
template <>
__device__ __forceinline__ float load_as_float<float>(const float* p, long long i) { return p[i]; }
template <>
__device__ __forceinline__ float load_as_float<__nv_bfloat16>(const __nv_bfloat16* p, long long i) {
  return __bfloat162float(p[i]);
}
template <>
__device__ __forceinline__ float load_as_float<__half>(const __half* p, long long i) {
  return __half2float(p[i]);
}

// grid = (chunks, rows). Each CTA keeps its chunk's keys in shared memory, selects the chunk-local
// top-k and appends them (unordered) to cand[row][chunk][k]; unused slots hold key 0.
template <typename T>
__global__ void __launch_bounds__(TOPK_THREADS)
topk_chunk_kernel(const T* __restrict__ grad, long long row_stride, int V, int chunk_len, int k,
                  const uint32_t* __restrict__ mask, uint64_t* __restrict__ cand) {
  extern __shared__ uint64_t sm_keys[];
  __shared__ uint32_t hist[256];
  __shared__ uint64_t ctl[4];
  __shared__ uint32_t out_count;
  const int row = blockIdx.y, chunk = blockIdx.x, chunks = gridDim.x;
  const int begin = chunk * chunk_len;
  const int n = min(chunk_len, V - begin);
  const T* src = grad + static_cast<long long>(row) * row_stride + begin;
  uint64_t* dst = cand + (static_cast<long long>(row) * chunks + chunk) * k;

  constexpr int VEC = 16 / sizeof(T);
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  if (vec_ok) {
    const int nv = n / VEC;
    for (int v = threadIdx.x; v < nv; v += blockDim.x) {
      const uint4 q = reinterpret_cast<const uint4*>(src)[v];
      const T* e = reinterpret_cast<const T*>(&q);
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        const int i = v * VEC + j;
        sm_keys[i] = make_key(load_as_float<T>(e, j), static_cast<uint32_t>(begin + i), mask);
      }
    }
    for (int i = nv * VEC + threadIdx.x; i < n; i += blockDim.x)
      sm_keys[i] = make_key(load_as_float<T>(src, i), static_cast<uint32_t>(begin + i), mask);
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x)
      sm_keys[i] = make_key(load_as_float<T>(src, i), static_cast<uint32_t>(begin + i), mask);
  }
  if (threadIdx.x == 0) out_count = 0;
  __syncthreads();

  if (n <= k) {
    for (int i = threadIdx.x; i < k; i += blockDim.x) dst[i] = (i < n) ? sm_keys[i] : 0ull;
    return;
  }
  radix_select_topk(sm_keys, n, k, hist, ctl);
  const uint64_t kth = ctl[0];
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const uint64_t key = sm_keys[i];
    if (key >= kth) {
      const uint32_t pos = atomicAdd(&out_count, 1u);
      if (pos < static_cast<uint32_t>(k)) dst[pos] = key;
    }
  }
}

// grid = rows. Merge the per-chunk candidates, keep the k largest, sort them descending (bitonic,
// shared memory) and emit token ids as int64 in torch.topk order.
__global__ void __launch_bounds__(TOPK_THREADS)
topk_merge_kernel(const uint64_t* __restrict__ cand, int n_cand, int k, int k_pow2,
                  long long* __restrict__ topk_ids) {
  extern __shared__ uint64_t sm_keys[];          // [n_cand] then [k_pow2]
  uint64_t* sel = sm_keys + n_cand;
  __shared__ uint32_t hist[256];
  __shared__ uint64_t ctl[4];
  __shared__ uint32_t out_count;
  const int row = blockIdx.x;
  const uint64_t* src = cand + static_cast<long long>(row) * n_cand;
  for (int i = threadIdx.x; i < n_cand; i += blockDim.x) sm_keys[i] = src[i];
  for (int i = threadIdx.x; i < k_pow2; i += blockDim.x) sel[i] = 0ull;
  if (threadIdx.x == 0) out_count = 0;
  __syncthreads();
  if (n_cand <= k) {
    for (int i = threadIdx.x; i < n_cand; i += blockDim.x) sel[i] = sm_keys[i];
  } else {
    radix_select_topk(sm_keys, n_cand, k, hist, ctl);
    const uint64_t kth = ctl[0];
    for (int i = threadIdx.x; i < n_cand; i += blockDim.x) {
      const uint64_t key = sm_keys[i];
      // padding keys (0) can only pass when kth == 0, i.e. fewer than k real entries exist
      if (key >= kth && key != 0ull) {
        const uint32_t pos = atomicAdd(&out_count, 1u);
        if (pos < static_cast<uint32_t>(k)) sel[pos] = key;
      }
    }
  }
  __syncthreads();
  // bitonic sort, descending
  for (int size = 2; size <= k_pow2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < k_pow2; i += blockDim.x) {
        const int j = i ^ stride;
        if (j > i) {
          const uint64_t a = sel[i], b = sel[j];
          const bool desc = ((i & size) == 0);
          if (desc ? (a < b) : (a > b)) {
            sel[i] = b;
            sel[j] = a;
          }
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < k; i += blockDim.x)
    topk_ids[static_cast<long long>(row) * k + i] =
        static_cast<long long>(0xFFFFFFFFu - static_cast<uint32_t>(sel[i] & 0xFFFFFFFFull));
}

// ------------------------------------------------------------------------------------------------
// One-launch top-k for 16-bit gradients (what the product feeds: the bf16 token gradient) over
// vocabularies of at most 65 536 tokens: value and index both fit 16 bits, so a row's keys are 32-bit
//   (orderable16(-g) << 16) | (0xFFFF - index)
// and the whole row (V * 4 B <= 256 KB... 128 KB at V = 32 064) sits in ONE CTA's shared memory: read the
// row once with 16-byte loads, 8-bit radix select (<= 4 passes), bitonic sort of the k winners, int64
// ids out in torch.topk order. Same ordering as the 64-bit two-level path (bf16/fp16 -> fp32 is an
// order-preserving injection; ties resolve to the lower token id).
__device__ __forceinline__ uint32_t orderable16_neg(uint16_t bits, bool is_nan) {
  if (is_nan) return 0xFFFFu;
  const uint16_t u = bits ^ 0x8000u;                       // -g
  return (u & 0x8000u) ? static_cast<uint16_t>(~u) : static_cast<uint16_t>(u | 0x8000u);
}
template <typename T>
__device__ __forceinline__ uint32_t make_key16(uint16_t bits, uint32_t idx, const uint32_t* mask);
template <>
__device__ __forceinline__ uint32_t make_key16<__nv_bfloat16>(uint16_t bits, uint32_t idx, const uint32_t* mask) {
  if (mask != nullptr && ((mask[idx >> 5] >> (idx & 31)) & 1u)) bits = 0x7F80u;        // +inf
  const bool nan = (bits & 0x7FFFu) > 0x7F80u;
  return (orderable16_neg(bits, nan) << 16) | (0xFFFFu - idx);
}
template <>
__device__ __forceinline__ uint32_t make_key16<__half>(uint16_t bits, uint32_t idx, const uint32_t* mask) {
  if (mask != nullptr && ((mask[idx >> 5] >> (idx & 31)) & 1u)) bits = 0x7C00u;        // +inf
  const bool nan = (bits & 0x7FFFu) > 0x7C00u;
  return (orderable16_neg(bits, nan) << 16) | (0xFFFFu - idx);
}

template <typename T>
__global__ void __launch_bounds__(1024)
topk_row16_kernel(const T* __restrict__ grad, long long row_stride, int V, int k, int k_pow2,
                  const uint32_t* __restrict__ mask, long long* __restrict__ topk_ids) {
  extern __shared__ uint32_t sm32[];              // [V] keys, then [k_pow2] winners
  uint32_t* keys = sm32;
  uint32_t* sel = sm32 + V;
  __shared__ uint32_t hist[256];
  __shared__ uint32_t ctl[4];                     // prefix, mask, remaining, done
  __shared__ uint32_t out_count;
  const int row = blockIdx.x;
  const uint16_t* src = reinterpret_cast<const uint16_t*>(grad + static_cast<long long>(row) * row_stride);
  if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    const int nv = V >> 3;
    for (int v = threadIdx.x; v < nv; v += blockDim.x) {
      const uint4 q = reinterpret_cast<const uint4*>(src)[v];
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t i = static_cast<uint32_t>(v * 8 + 2 * j);
        keys[i] = make_key16<T>(static_cast<uint16_t>(w[j] & 0xFFFFu), i, mask);
        keys[i + 1] = make_key16<T>(static_cast<uint16_t>(w[j] >> 16), i + 1, mask);
      }
    }
    for (int i = (nv << 3) + threadIdx.x; i < V; i += blockDim.x) keys[i] = make_key16<T>(src[i], i, mask);
  } else {
    for (int i = threadIdx.x; i < V; i += blockDim.x) keys[i] = make_key16<T>(src[i], i, mask);
  }
  for (int i = threadIdx.x; i < k_pow2; i += blockDim.x) sel[i] = 0u;
  if (threadIdx.x == 0) { ctl[0] = 0; ctl[1] = 0; ctl[2] = static_cast<uint32_t>(k); ctl[3] = 0; out_count = 0; }
  __syncthreads();
  if (V > k) {
    for (int pass = 0; pass < 4; ++pass) {
      const int shift = 24 - 8 * pass;
      for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
      __syncthreads();
      const uint32_t prefix = ctl[0], pmask = ctl[1];
      for (int i = threadIdx.x; i < V; i += blockDim.x) {
        const uint32_t key = keys[i];
        if ((key & pmask) == prefix) hist_add(hist, (key >> shift) & 255u);
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        uint32_t remaining = ctl[2], above = 0;
        int d = 255;
        for (; d > 0; --d) {
          if (above + hist[d] >= remaining) break;
          above += hist[d];
        }
        remaining -= above;
        ctl[0] = prefix | (static_cast<uint32_t>(d) << shift);
        ctl[1] = pmask | (255u << shift);
        ctl[2] = remaining;
        if (hist[d] == remaining) ctl[3] = 1;     // the whole bucket is wanted: prefix (low bits 0) is the threshold
      }
      __syncthreads();
      if (ctl[3] != 0) break;
    }
  }
  const uint32_t kth = V > k ? ctl[0] : 0u;
  for (int i = threadIdx.x; i < V; i += blockDim.x) {
    const uint32_t key = keys[i];
    if (key >= kth) {
      const uint32_t pos = atomicAdd(&out_count, 1u);
      if (pos < static_cast<uint32_t>(k)) sel[pos] = key;
    }
  }
  __syncthreads();
  for (int size = 2; size <= k_pow2; size <<= 1) {            // bitonic sort, descending
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < k_pow2; i += blockDim.x) {
        const int j = i ^ stride;
        if (j > i) {
          const uint32_t a = sel[i], b = sel[j];
          const bool desc = ((i & size) == 0);
          if (desc ? (a < b) : (a > b)) { sel[i] = b; sel[j] = a; }
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < k; i += blockDim.x)
    topk_ids[static_cast<long long>(row) * k + i] = static_cast<long long>(0xFFFFu - (sel[i] & 0xFFFFu));
}

// torch's grid for an n-element distribution kernel: min(sms * (max_threads_per_sm / 256),
// ceil(n / 256)) blocks of 256 threads; element e is produced by thread e % T on its
// (e / T)-th value, values coming four at a time from one Philox counter step.
struct PhiloxPlan {
  unsigned long long seed;
  unsigned long long offset;   // in 32-bit outputs, multiple of 4
  unsigned int total_threads;  // T
};

__device__ __forceinline__ uint32_t philox_raw(const PhiloxPlan& pl, unsigned long long e) {
  const unsigned long long t = e % pl.total_threads;
  const unsigned long long kth = e / pl.total_threads;
  curandStatePhilox4_32_10_t st;
  curand_init(pl.seed, t, pl.offset + 4ull * (kth >> 2), &st);
  const uint4 r = curand4(&st);
  const uint32_t c = static_cast<uint32_t>(kth & 3);
  return c == 0 ? r.x : (c == 1 ? r.y : (c == 2 ? r.z : r.w));
}
__device__ __forceinline__ float philox_uniform(const PhiloxPlan& pl, unsigned long long e) {
  // curand_uniform4's mapping, then torch's (0,1] -> [0,1) flip
  const float u = _curand_uniform(philox_raw(pl, e));
  return u == 1.0f ? 0.0f : u;
}

// One warp per candidate.
__global__ void __launch_bounds__(256)
sample_replace_kernel(const long long* __restrict__ ids, const long long* __restrict__ topk_ids,
                      long long* __restrict__ new_ids, int search_width, int n_tok, int topk,
                      int n_replace, PhiloxPlan rand_plan, PhiloxPlan int_plan) {
  extern __shared__ float sm_u[];   // [warps][n_tok]
  const int warp_in_block = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + warp_in_block;
  if (b >= search_width) return;
  float* u = sm_u + static_cast<long long>(warp_in_block) * n_tok;
  for (int p = lane; p < n_tok; p += 32) {
    u[p] = philox_uniform(rand_plan, static_cast<unsigned long long>(b) * n_tok + p);
    new_ids[static_cast<long long>(b) * n_tok + p] = ids[p];
  }
  __syncwarp();
  for (int j = 0; j < n_replace; ++j) {
    // argsort(rand)[j]: smallest remaining value, lowest index on ties
    float best = INFINITY;
    int best_p = 0x7fffffff;
    for (int p = lane; p < n_tok; p += 32) {
      const float v = u[p];
      if (v < best || (v == best && p < best_p)) {
        best = v;
        best_p = p;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int op = __shfl_xor_sync(0xffffffffu, best_p, o);
      if (ov < best || (ov == best && op < best_p)) {
        best = ov;
        best_p = op;
      }
    }
    if (lane == 0) {
      u[best_p] = INFINITY;   // uniforms are < 1, so INFINITY marks "taken"
      const uint32_t r = philox_raw(int_plan, static_cast<unsigned long long>(b) * n_replace + j);
      const int pick = static_cast<int>(r % static_cast<uint32_t>(topk));
      new_ids[static_cast<long long>(b) * n_tok + best_p] =
          topk_ids[static_cast<long long>(best_p) * topk + pick];
    }
    __syncwarp();
  }
}

static unsigned int torch_grid_threads(long long n, int sms, int max_threads_per_sm) {
  long long blocks = (n + 255) / 256;
  const long long cap = static_cast<long long>(sms) * (max_threads_per_sm / 256);
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<unsigned int>(blocks * 256);
}
static unsigned long long torch_offset_advance(long long n, unsigned int total_threads) {
  return static_cast<unsigned long long>((n - 1) / (static_cast<long long>(total_threads) * 4) + 1) * 4ull;
}

// Chunk length of the first top-k level: at most 32 chunks per row, at least 4096 elements each.
static int topk_chunk_len(int V) {
  int len = ((V + 31) / 32 + 1023) / 1024 * 1024;
  return len < 4096 ? 4096 : len;
}

// BMA_TOPK_ONE_LAUNCH=0 keeps the two-level path for 16-bit gradients too (A/B runs)
static bool topk_one_launch_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("BMA_TOPK_ONE_LAUNCH");
    v = (e != nullptr && atoi(e) == 0) ? 0 : 1;
  }
  return v != 0;
}

size_t sampler_workspace_bytes(int n_tok, int V, int topk) {
  const int chunk_len = topk_chunk_len(V);
  const int chunks = (V + chunk_len - 1) / chunk_len;
  return static_cast<size_t>(n_tok) * chunks * topk * sizeof(uint64_t) +
         static_cast<size_t>(n_tok) * topk * sizeof(long long);
}

// grad: [n_tok, V] (dtype: 0 bf16, 1 fp16, 2 fp32), row stride in elements.
cudaError_t launch_sample_ids(const long long* ids, const void* grad, int grad_dtype,
                              long long row_stride, int n_tok, int V, const uint32_t* mask_bits,
                              int search_width, int topk, int n_replace, unsigned long long seed,
                              unsigned long long offset, void* workspace, long long* topk_ids_out,
                              long long* new_ids, unsigned long long* offset_advance,
                              cudaStream_t stream) {
  if (n_tok <= 0 || V <= 0 || search_width <= 0) return cudaErrorInvalidValue;
  if (topk <= 0 || topk > V || topk > 4096) return cudaErrorInvalidValue;
  if (n_replace <= 0 || n_replace > n_tok) return cudaErrorInvalidValue;
  static int sms = 0, mtps = 0;                     // one device per process
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&mtps, cudaDevAttrMaxThreadsPerMultiProcessor, dev);
    if (sms <= 0) sms = 148;
    if (mtps <= 0) mtps = 2048;
  }

  const int chunk_len = topk_chunk_len(V);
  const int chunks = (V + chunk_len - 1) / chunk_len;
  uint64_t* cand = static_cast<uint64_t*>(workspace);
  long long* topk_ids = topk_ids_out != nullptr
                            ? topk_ids_out
                            : reinterpret_cast<long long*>(cand + static_cast<size_t>(n_tok) * chunks * topk);

  cudaError_t e;
  int k_pow2 = 1;
  while (k_pow2 < topk) k_pow2 <<= 1;
  // 16-bit gradients over <= 65 536 tokens: the whole row fits one CTA's shared memory as 32-bit keys
  const size_t smem16 = (static_cast<size_t>(V) + k_pow2) * sizeof(uint32_t);
  const bool one_launch = (grad_dtype == 0 || grad_dtype == 1) && V <= 65536 && smem16 <= 220 * 1024 &&
                          topk_one_launch_enabled();
  if (one_launch) {
    static size_t attr16[2] = {0, 0};
    if (attr16[grad_dtype] < smem16) {
      e = grad_dtype == 0 ? cudaFuncSetAttribute(topk_row16_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem16))
                          : cudaFuncSetAttribute(topk_row16_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem16));
      if (e != cudaSuccess) return e;
      attr16[grad_dtype] = smem16;
    }
    if (grad_dtype == 0)
      topk_row16_kernel<__nv_bfloat16><<<n_tok, 1024, smem16, stream>>>(static_cast<const __nv_bfloat16*>(grad), row_stride, V,
                                                                         topk, k_pow2, mask_bits, topk_ids);
    else
      topk_row16_kernel<__half><<<n_tok, 1024, smem16, stream>>>(static_cast<const __half*>(grad), row_stride, V, topk, k_pow2,
                                                                  mask_bits, topk_ids);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  } else {
  const size_t smem1 = static_cast<size_t>(chunk_len) * sizeof(uint64_t);
  dim3 grid1(chunks, n_tok);
#define BMA_LAUNCH_CHUNK(T)                                                                       \
  do {                                                                                            \
    e = cudaFuncSetAttribute(topk_chunk_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                             static_cast<int>(smem1));                                            \
    if (e != cudaSuccess) return e;                                                               \
    topk_chunk_kernel<T><<<grid1, TOPK_THREADS, smem1, stream>>>(                                 \
        static_cast<const T*>(grad), row_stride, V, chunk_len, topk, mask_bits, cand);            \
  } while (0)
  if (grad_dtype == 0) BMA_LAUNCH_CHUNK(__nv_bfloat16);
  else if (grad_dtype == 1) BMA_LAUNCH_CHUNK(__half);
  else if (grad_dtype == 2) BMA_LAUNCH_CHUNK(float);
  else return cudaErrorInvalidValue;
#undef BMA_LAUNCH_CHUNK
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;

  const int n_cand = chunks * topk;
  const size_t smem2 = static_cast<size_t>(n_cand + k_pow2) * sizeof(uint64_t);
  e = cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           static_cast<int>(smem2));
  if (e != cudaSuccess) return e;
  topk_merge_kernel<<<n_tok, TOPK_THREADS, smem2, stream>>>(cand, n_cand, topk, k_pow2, topk_ids);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  }

  const long long n1 = static_cast<long long>(search_width) * n_tok;
  const long long n2 = static_cast<long long>(search_width) * n_replace;
  PhiloxPlan p1{seed, offset, torch_grid_threads(n1, sms, mtps)};
  const unsigned long long adv1 = torch_offset_advance(n1, p1.total_threads);
  PhiloxPlan p2{seed, offset + adv1, torch_grid_threads(n2, sms, mtps)};
  const unsigned long long adv2 = torch_offset_advance(n2, p2.total_threads);
  if (offset_advance != nullptr) *offset_advance = adv1 + adv2;

  const int warps = 8;
  const size_t smem3 = static_cast<size_t>(warps) * n_tok * sizeof(float);
  if (smem3 > 48 * 1024) return cudaErrorInvalidValue;
  const int blocks = (search_width + warps - 1) / warps;
  sample_replace_kernel<<<blocks, warps * 32, smem3, stream>>>(ids, topk_ids, new_ids, search_width,
                                                               n_tok, topk, n_replace, p1, p2);
  return cudaGetLastError();
}

}  // namespace bma
