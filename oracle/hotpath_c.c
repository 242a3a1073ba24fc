/* ORACLE -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * Plain-C restatement of the integer / fp32 pieces of the reference's hot path, used by the tests
 * as a second, torch-free checker:
 *   - PGD step                       bimodalattack/bimodal_attack.py:1030-1037
 *   - candidate sampling             bimodal_attack.py:130-163, given the random draws
 *   - torch's CUDA random streams    the draws themselves: torch.rand / torch.randint on CUDA are
 *     Philox4x32-10 (cuRAND device API, pinned by torch==2.6.0+cu126 in the reference's
 *     requirements.txt:1; algorithm: Salmon et al., "Parallel random numbers: as easy as 1, 2, 3",
 *     SC'11) consumed by ATen's grid-stride distribution kernel (DistributionTemplates.h).
 *   - mean cross-entropy of a logits block   bimodal_attack.py:1293-1298
 * Pinned by tests/test_oracle_c.py against torch on CPU and against vectors captured from a B200
 * (tests/golden/cuda_philox.json).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------- PGD (:1033-1035) */
static float clampf_nan(float v, float lo, float hi) {
  if (v != v) return v;
  float m = v < lo ? lo : v;
  return m > hi ? hi : m;
}
void oracle_pgd_step(float* x, const float* g, const float* x0, long n, float step, float eps) {
  for (long i = 0; i < n; ++i) {
    float sgn = (float)(0.0f < g[i]) - (float)(g[i] < 0.0f);   /* torch.sign: 0 for +-0 and NaN */
    float v = x[i] - step * sgn;
    v = clampf_nan(v, x0[i] - eps, x0[i] + eps);
    x[i] = clampf_nan(v, 0.0f, 1.0f);
  }
}

/* ---------------------------------------------------------------- Philox4x32-10 */
static void philox_round(uint32_t c[4], const uint32_t k[2]) {
  const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
  const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
  const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
  const uint32_t n1 = (uint32_t)p1;
  const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
  const uint32_t n3 = (uint32_t)p0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
/* One counter step of curand_init(seed, subsequence, offset) ... curand4(): offset counts 32-bit
 * outputs and must be a multiple of 4 (torch always advances by multiples of 4). */
void oracle_philox4(uint64_t seed, uint64_t subsequence, uint64_t offset, uint32_t out[4]) {
  uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  const uint64_t ctr = offset / 4;
  uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)subsequence, (uint32_t)(subsequence >> 32)};
  for (int r = 0; r < 9; ++r) {
    philox_round(c, k);
    k[0] += 0x9E3779B9u;
    k[1] += 0xBB67AE85u;
  }
  philox_round(c, k);
  memcpy(out, c, sizeof(c));
}

/* torch's launch shape for an n-element distribution kernel (calc_execution_policy). */
static uint64_t grid_threads(long n, int sms, int max_threads_per_sm) {
  uint64_t blocks = ((uint64_t)n + 255) / 256;
  const uint64_t cap = (uint64_t)sms * (uint64_t)(max_threads_per_sm / 256);
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return blocks * 256;
}
uint64_t oracle_offset_advance(long n, int sms, int max_threads_per_sm) {
  const uint64_t t = grid_threads(n, sms, max_threads_per_sm);
  return (((uint64_t)n - 1) / (t * 4) + 1) * 4;
}
static uint32_t raw_draw(uint64_t seed, uint64_t offset, uint64_t total_threads, uint64_t e) {
  const uint64_t t = e % total_threads, kth = e / total_threads;
  uint32_t r[4];
  oracle_philox4(seed, t, offset + 4 * (kth >> 2), r);
  return r[kth & 3];
}
/* torch.rand(n, device="cuda") with generator state (seed, offset) */
void oracle_cuda_rand(uint64_t seed, uint64_t offset, long n, int sms, int mtps, float* out) {
  const uint64_t T = grid_threads(n, sms, mtps);
  for (long e = 0; e < n; ++e) {
    const uint32_t x = raw_draw(seed, offset, T, (uint64_t)e);
    float u = (float)x * 2.3283064365386963e-10f + 1.1641532182693481e-10f;   /* curand_uniform */
    out[e] = (u == 1.0f) ? 0.0f : u;                                           /* (0,1] -> [0,1) */
  }
}
/* torch.randint(0, range, (n,), device="cuda"), range < 2^32 */
void oracle_cuda_randint(uint64_t seed, uint64_t offset, long n, int sms, int mtps, uint32_t range, int64_t* out) {
  const uint64_t T = grid_threads(n, sms, mtps);
  for (long e = 0; e < n; ++e) out[e] = (int64_t)(raw_draw(seed, offset, T, (uint64_t)e) % range);
}

/* ---------------------------------------------------------------- sampling (:142-162) */
typedef struct { float v; int32_t i; } vi_t;
static int by_value_then_index(const void* a, const void* b) {
  const vi_t* x = (const vi_t*)a; const vi_t* y = (const vi_t*)b;
  if (x->v < y->v) return -1;
  if (x->v > y->v) return 1;
  return (x->i > y->i) - (x->i < y->i);
}
/* grad [L][V] fp32 (not modified), banned: V flags or NULL, uniforms [sw][L], picks [sw][n_replace]
 * -> new_ids [sw][L]. Ties in grad and in the uniforms resolve to the lower index. */
int oracle_sample_ids(const int64_t* ids, const float* grad, int L, int V, const uint8_t* banned,
                      int sw, int topk, int n_replace, const float* uniforms, const int64_t* picks,
                      int64_t* new_ids) {
  vi_t* row = (vi_t*)malloc(sizeof(vi_t) * (size_t)V);
  int64_t* top = (int64_t*)malloc(sizeof(int64_t) * (size_t)L * topk);
  vi_t* pos = (vi_t*)malloc(sizeof(vi_t) * (size_t)L);
  if (!row || !top || !pos) return -1;
  for (int l = 0; l < L; ++l) {                      /* topk of -grad == smallest grad first */
    for (int v = 0; v < V; ++v) {
      row[v].v = (banned && banned[v]) ? INFINITY : grad[(size_t)l * V + v];
      row[v].i = v;
    }
    qsort(row, (size_t)V, sizeof(vi_t), by_value_then_index);
    for (int k = 0; k < topk; ++k) top[(size_t)l * topk + k] = row[k].i;
  }
  for (int b = 0; b < sw; ++b) {
    for (int l = 0; l < L; ++l) {
      new_ids[(size_t)b * L + l] = ids[l];
      pos[l].v = uniforms[(size_t)b * L + l];
      pos[l].i = l;
    }
    qsort(pos, (size_t)L, sizeof(vi_t), by_value_then_index);      /* argsort(rand) */
    for (int j = 0; j < n_replace; ++j) {
      const int p = pos[j].i;
      new_ids[(size_t)b * L + p] = top[(size_t)p * topk + picks[(size_t)b * n_replace + j]];
    }
  }
  free(row); free(top); free(pos);
  return 0;
}

/* ---------------------------------------------------------------- mean CE (:1293-1298) */
float oracle_mean_ce(const float* logits, int T, int V, const int64_t* target) {
  double acc = 0.0;
  for (int t = 0; t < T; ++t) {
    const float* r = logits + (size_t)t * V;
    float m = r[0];
    for (int v = 1; v < V; ++v) m = r[v] > m ? r[v] : m;
    double s = 0.0;
    for (int v = 0; v < V; ++v) s += exp((double)r[v] - m);
    acc += (m + log(s)) - r[target[t]];
  }
  return (float)(acc / T);
}
