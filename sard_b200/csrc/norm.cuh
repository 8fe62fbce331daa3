// RunningNorm pieces shared between norm.cu and the fused tower prologue (tower_tc.cu).
#pragma once
#include "common.cuh"

struct Mom { double n, mean, m2; };
__device__ __forceinline__ Mom mom_merge(Mom a, Mom b) {
  if (b.n == 0.0) return a;
  if (a.n == 0.0) return b;
  const double n = a.n + b.n, d = b.mean - a.mean;
  Mom r;
  r.n = n;
  r.mean = a.mean + d * (b.n / n);
  r.m2 = a.m2 + b.m2 + d * d * (a.n * b.n / n);
  return r;
}

// Running update of one RunningNorm from n_records moment records {n, mean[D], M2[D]} (khrylib/rl/core/running_norm.py:30-42),
// by ONE block; also refreshes inv = 1 / (std + 1e-8).
__device__ __forceinline__ void norm_update_body(const double* __restrict__ records, int n_records, int D,
                                                 float* __restrict__ mean, float* __restrict__ var, float* __restrict__ std,
                                                 long long* __restrict__ count, float* __restrict__ inv) {
  const long long n_old = *count;
  const int stride = 1 + 2 * D;
  double m_total = 0.0;
  for (int k = 0; k < n_records; ++k) m_total += records[(size_t)k * stride];
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    Mom acc = {0.0, 0.0, 0.0};
    for (int k = 0; k < n_records; ++k) {
      Mom b;
      b.n = records[(size_t)k * stride];
      b.mean = records[(size_t)k * stride + 1 + c];
      b.m2 = records[(size_t)k * stride + 1 + D + c];
      acc = mom_merge(acc, b);
    }
    if (acc.n > 0.0) {
      const double mean_x = acc.mean, var_x = acc.m2 / acc.n;   // biased (unbiased=False)
      const double w = (double)n_old / (acc.n + (double)n_old);
      const double mu = (double)mean[c], v = (double)var[c];
      const double dm = mean_x - mu;
      const double v_new = w * v + (1.0 - w) * var_x + w * (1.0 - w) * dm * dm;
      const double mu_new = w * mu + (1.0 - w) * mean_x;
      const float sd = (float)sqrt(v_new);
      var[c] = (float)v_new;
      mean[c] = (float)mu_new;
      std[c] = sd;
      inv[c] = 1.0f / (sd + 1e-8f);
    } else {
      inv[c] = 1.0f / (std[c] + 1e-8f);
    }
  }
  if (threadIdx.x == 0) *count = n_old + (long long)(m_total + 0.5);
}
