// Fused l-inf PGD image step (reference bimodal_attack.py:1030-1037):
//   x <- clamp(clamp(x - step * sign(g), x0 - eps, x0 + eps), 0, 1)      with step = alpha * eps
// One pass over four fp32 streams (x, g, x0 in; x out), float4 vectorised. Optionally also writes
// the per-channel normalised image (x - mean) / std that the vision tower consumes (K15).
// Arithmetic order follows the reference op by op so the result is bit-identical in fp32:
// sign(g) is (0 < g) - (g < 0) (so sign(+-0) = sign(NaN) = 0), the product step*sign is exact,
// and the two clamps use min(max(.)) with torch.clamp's NaN propagation.
#include "common.cuh"
#include "kernels.h"

namespace bma {

__device__ __forceinline__ float clamp_like_torch(float v, float lo, float hi) {
  // torch.clamp(v, lo, hi) == min(max(v, lo), hi) with NaN in v propagated.
  if (v != v) return v;
  return fminf(fmaxf(v, lo), hi);
}

__device__ __forceinline__ float pgd_one(float x, float g, float x0, float step, float eps) {
  const float sgn = static_cast<float>(0.0f < g) - static_cast<float>(g < 0.0f);
  float v = __fsub_rn(x, __fmul_rn(step, sgn));
  v = clamp_like_torch(v, __fsub_rn(x0, eps), __fadd_rn(x0, eps));
  v = clamp_like_torch(v, 0.0f, 1.0f);
  return v;
}

struct NormArgs {
  float mean[4];
  float std_[4];
  int channels;        // 0 = no normalised output
  long long plane;     // H*W
};

__global__ void pgd_step_kernel(float* __restrict__ x, const float* __restrict__ g,
                                const float* __restrict__ x0, float* __restrict__ norm_out,
                                long long n, float step, float eps, NormArgs na) {
  const long long nvec = n >> 2;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
       i += stride) {
    const float4 xv = reinterpret_cast<const float4*>(x)[i];
    const float4 gv = reinterpret_cast<const float4*>(g)[i];
    const float4 ov = reinterpret_cast<const float4*>(x0)[i];
    float4 r;
    r.x = pgd_one(xv.x, gv.x, ov.x, step, eps);
    r.y = pgd_one(xv.y, gv.y, ov.y, step, eps);
    r.z = pgd_one(xv.z, gv.z, ov.z, step, eps);
    r.w = pgd_one(xv.w, gv.w, ov.w, step, eps);
    reinterpret_cast<float4*>(x)[i] = r;
    if (norm_out != nullptr) {
      const float rr[4] = {r.x, r.y, r.z, r.w};
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = static_cast<int>(((i * 4 + j) / na.plane) % na.channels);
        o[j] = __fdiv_rn(__fsub_rn(rr[j], na.mean[c]), na.std_[c]);
      }
      reinterpret_cast<float4*>(norm_out)[i] = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
  // scalar tail (n not a multiple of 4)
  const long long tail0 = nvec << 2;
  for (long long i = tail0 + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += stride) {
    const float r = pgd_one(x[i], g[i], x0[i], step, eps);
    x[i] = r;
    if (norm_out != nullptr) {
      const int c = static_cast<int>((i / na.plane) % na.channels);
      norm_out[i] = __fdiv_rn(__fsub_rn(r, na.mean[c]), na.std_[c]);
    }
  }
}

cudaError_t launch_pgd_step(float* x, const float* g, const float* x0, float* norm_out,
                            long long n, float step, float eps, const float* mean,
                            const float* std_, int channels, long long plane,
                            cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  NormArgs na{};
  na.channels = 0;
  na.plane = plane > 0 ? plane : 1;
  if (norm_out != nullptr) {
    if (channels <= 0 || channels > 4 || mean == nullptr || std_ == nullptr)
      return cudaErrorInvalidValue;
    na.channels = channels;
    for (int c = 0; c < channels; ++c) {
      na.mean[c] = mean[c];
      na.std_[c] = std_[c];
    }
  }
  const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(g) |
                         reinterpret_cast<uintptr_t>(x0) |
                         reinterpret_cast<uintptr_t>(norm_out)) & 15) == 0;
  if (!aligned) return cudaErrorMisalignedAddress;
  const int threads = 256;
  long long want = ((n >> 2) + threads - 1) / threads;
  int sms = 148;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long cap = static_cast<long long>(sms) * 8;
  int blocks = static_cast<int>(want < 1 ? 1 : (want > cap ? cap : want));
  pgd_step_kernel<<<blocks, threads, 0, stream>>>(x, g, x0, norm_out, n, step, eps, na);
  return cudaGetLastError();
}

}  // namespace bma
