// Host-side launch helpers for the tcgen05 GEMM: TMA tensor-map construction (driver entry point
// resolved at run time, so the library does not link libcuda) and a small descriptor cache.
#pragma once
#include <cudaTypedefs.h>

#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "gemm_tcgen05.cuh"

namespace bma {

inline int g_num_sms = 0;

inline int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

// Points this translation unit's device-side watchdog at the process-wide host-mapped word.
static inline void watchdog_attach() {
  static bool done = false;
  if (done) return;
  unsigned int* host = watchdog_host_word();
  unsigned int* dev = nullptr;
  if (host != nullptr && cudaHostGetDevicePointer(reinterpret_cast<void**>(&dev), host, 0) == cudaSuccess)
    cudaMemcpyToSymbol(g_watchdog_host, &dev, sizeof(dev));
  done = true;
}

inline PFN_cuTensorMapEncodeTiled_v12000 tensor_map_encoder() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || ptr == nullptr) {
      fprintf(stderr, "[bimodal_b200] cuTensorMapEncodeTiled unavailable (%s)\n",
              cudaGetErrorString(e));
      return nullptr;
    }
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  }
  return fn;
}

struct TmapKey {
  const void* ptr;
  long long rows, cols, ld;
  int box_rows;
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    h ^= static_cast<size_t>(k.rows) * 0x9E3779B97F4A7C15ull;
    h ^= static_cast<size_t>(k.cols) * 0xC2B2AE3D27D4EB4Full;
    h ^= static_cast<size_t>(k.ld) * 0x165667B19E3779F9ull;
    h ^= static_cast<size_t>(k.box_rows) << 7;
    return h;
  }
};

// Row-major [rows, cols] bf16 matrix with leading dimension ld (elements); box = box_rows x 64,
// 128-byte swizzle, zero fill outside the matrix.
inline bool make_tmap_bf16(CUtensorMap* out, const void* ptr, long long rows, long long cols,
                           long long ld, int box_rows) {
  static std::mutex mu;
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  TmapKey key{ptr, rows, cols, ld, box_rows};
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return true;
    }
  }
  auto enc = tensor_map_encoder();
  if (enc == nullptr) return false;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || ((ld * 2) & 15) != 0) {
    fprintf(stderr, "[bimodal_b200] TMA operand must be 16-byte aligned (ptr=%p ld=%lld)\n", ptr, ld);
    return false;
  }
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(GEMM_BK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "[bimodal_b200] cuTensorMapEncodeTiled failed: %d (rows=%lld cols=%lld ld=%lld)\n",
            static_cast<int>(r), rows, cols, ld);
    return false;
  }
  std::lock_guard<std::mutex> g(mu);
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, *out);
  return true;
}

template <int BN, int EPI, int HD = 128, int CTAS = 1>
inline cudaError_t launch_gemm(const __nv_bfloat16* A, long long lda, const __nv_bfloat16* B,
                               long long ldb, GemmParams p, cudaStream_t stream,
                               int max_ctas = 0) {
  using Cfg = GemmCfg<BN, CTAS>;
  if (p.M <= 0 || p.N <= 0 || p.K <= 0) return cudaSuccess;
  watchdog_attach();
  CUtensorMap ta, tb;
  const int nb = p.batch > 1 ? p.batch : 1;
  if (nb > 1 && ((p.a_col != 0 || p.b_col != 0) && p.K % GEMM_BK != 0)) return cudaErrorInvalidValue;
  if (!make_tmap_bf16(&ta, A, p.M + (long long)(nb - 1) * p.a_row, p.K + (long long)(nb - 1) * p.a_col, lda, GEMM_BM))
    return cudaErrorInvalidValue;
  if (!make_tmap_bf16(&tb, B, p.N + (long long)(nb - 1) * p.b_row, p.K + (long long)(nb - 1) * p.b_col, ldb, BN / CTAS))
    return cudaErrorInvalidValue;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tn_kernel<BN, EPI, HD, CTAS>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int m_tiles = ((p.M + GEMM_BM - 1) / GEMM_BM + CTAS - 1) / CTAS;
  const int n_tiles = (p.N + BN - 1) / BN;
  long long grid_ll = (long long)m_tiles * n_tiles * CTAS * nb;
  int grid = grid_ll > (1 << 30) ? (1 << 30) : static_cast<int>(grid_ll);
  const int cap = (max_ctas > 0 ? max_ctas : num_sms()) / CTAS * CTAS;
  if (grid > cap) grid = cap;
  if constexpr (CTAS == 1) {
    gemm_bf16_tn_kernel<BN, EPI, HD, 1><<<grid, GEMM_THREADS, Cfg::kSmemBytes, stream>>>(ta, tb, p);
    return cudaGetLastError();
  } else {
    // CTA pairs: clusters of 2 along x (both CTAs of a cluster share a TPC)
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CTAS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, gemm_bf16_tn_kernel<BN, EPI, HD, CTAS>, ta, tb, p);
  }
}

// L2 rasterisation: the M supergroup's A rows (group_m x 128 x K bf16) stay L2-resident while the B
// (weight) tiles stream once per supergroup, so a larger group means fewer weight re-reads from
// HBM. ~48 MB of A per supergroup leaves the rest of the 126 MB L2 to B and the output.
// BMA_GROUP_M overrides (experiments).
inline int group_m_for(int K) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("BMA_GROUP_M");
    forced = e != nullptr ? atoi(e) : 0;
  }
  if (forced > 0) return forced;
  int g = static_cast<int>((48ll << 20) / (128ll * K * 2));
  g &= ~1;
  return g < 8 ? 8 : (g > 64 ? 64 : g);
}

// The scorer's big GEMMs (BN = 256) run on CTA pairs unless BMA_GEMM_CTAS=1 is set (A/B switch).
inline int gemm_ctas() {
  static int v = 0;
  if (v == 0) {
    const char* e = getenv("BMA_GEMM_CTAS");
    v = (e != nullptr && atoi(e) == 1) ? 1 : 2;
  }
  return v;
}

template <int EPI, int HD = 128>
inline cudaError_t launch_gemm_big(const __nv_bfloat16* A, long long lda, const __nv_bfloat16* B,
                                   long long ldb, GemmParams p, cudaStream_t stream) {
  if (gemm_ctas() == 2) return launch_gemm<256, EPI, HD, 2>(A, lda, B, ldb, p, stream);
  return launch_gemm<256, EPI, HD, 1>(A, lda, B, ldb, p, stream);
}

}  // namespace bma
