// RunningNorm (khrylib/rl/core/running_norm.py:22-42): batch moments, running update, apply.
//
// Moments are accumulated as SHIFTED sums per 128-row chunk (shift = the chunk's first row), which
// makes constant columns (the subgroup embedding, the all-zero sim columns of the design stages)
// come out with variance exactly 0 like the reference's fp64 var_mean, and are merged with the
// pairwise (Chan) formula in fp64 in a fixed order -> deterministic and rank-count independent.
#include "common.cuh"
#include "norm.cuh"

#define NORM_CHUNK 128

// one block per 128-row chunk; threadIdx.y splits the chunk into 4 sub-chunks of 32 rows that are merged in
// sub-chunk order (pairwise formula), so the partial layout stays one {mean, M2} pair per chunk and column
#define NORM_SUB 4
__global__ void norm_moments_kernel(const float* __restrict__ x, int ldx, int n, int D,
                                    float* __restrict__ partials) {
  pdl_sync();
  extern __shared__ float sm[];                    // [NORM_SUB][3][blockDim.x]  (count, mean, M2)
  const int chunk = blockIdx.x;
  const int sub = threadIdx.y;
  const int rows_per_sub = NORM_CHUNK / NORM_SUB;
  const int r0 = min(n, chunk * NORM_CHUNK + sub * rows_per_sub), r1 = min(n, r0 + rows_per_sub);
  const int W = blockDim.x;
  for (int c0 = 0; c0 < D; c0 += W) {
    const int c = c0 + threadIdx.x;
    float cnt = (float)(r1 - r0), mean = 0.f, m2 = 0.f;
    if (c < D && r1 > r0) {
      const float shift = x[(size_t)r0 * ldx + c];
      float s1 = 0.f, s2 = 0.f;
      for (int r = r0; r < r1; ++r) {
        const float d = x[(size_t)r * ldx + c] - shift;
        s1 += d;
        s2 = fmaf(d, d, s2);
      }
      const float m = s1 / cnt;
      mean = shift + m;
      m2 = fmaxf(s2 - s1 * m, 0.f);
    }
    sm[(sub * 3 + 0) * W + threadIdx.x] = cnt;
    sm[(sub * 3 + 1) * W + threadIdx.x] = mean;
    sm[(sub * 3 + 2) * W + threadIdx.x] = m2;
    __syncthreads();
    if (sub == 0 && c < D) {
      float na = sm[0 * W + threadIdx.x], ma = sm[1 * W + threadIdx.x], qa = sm[2 * W + threadIdx.x];
#pragma unroll
      for (int j = 1; j < NORM_SUB; ++j) {
        const float nb = sm[(j * 3 + 0) * W + threadIdx.x];
        if (nb > 0.f) {
          const float mb = sm[(j * 3 + 1) * W + threadIdx.x], qb = sm[(j * 3 + 2) * W + threadIdx.x];
          const float nn = na + nb, d = mb - ma;
          ma = ma + d * (nb / nn);
          qa = qa + qb + d * d * (na * nb / nn);
          na = nn;
        }
      }
      partials[((size_t)chunk * 2 + 0) * D + c] = ma;
      partials[((size_t)chunk * 2 + 1) * D + c] = qa;
    }
    __syncthreads();
  }
}


// one warp per column: lanes take chunks round-robin, then a fixed shuffle tree
__global__ void norm_merge_chunks_kernel(const float* __restrict__ partials, int n, int D, int nchunks,
                                         double* __restrict__ record) {
  pdl_sync();
  const int col = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (col >= D) return;
  Mom acc = {0.0, 0.0, 0.0};
  for (int k = lane; k < nchunks; k += 32) {
    Mom b;
    b.n = (double)(min(n, (k + 1) * NORM_CHUNK) - k * NORM_CHUNK);
    b.mean = (double)partials[((size_t)k * 2 + 0) * D + col];
    b.m2 = (double)partials[((size_t)k * 2 + 1) * D + col];
    acc = mom_merge(acc, b);
  }
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    Mom other;
    other.n = __shfl_xor_sync(0xffffffffu, acc.n, o);
    other.mean = __shfl_xor_sync(0xffffffffu, acc.mean, o);
    other.m2 = __shfl_xor_sync(0xffffffffu, acc.m2, o);
    // keep the merge order identical on both partners: lower lane's data is always "a"
    acc = (lane & o) ? mom_merge(other, acc) : mom_merge(acc, other);
  }
  if (lane == 0) {
    record[1 + col] = acc.mean;
    record[1 + D + col] = acc.m2;
    if (col == 0) record[0] = (double)n;
  }
}

extern "C" int sard_norm_moments(const float* x, int ldx, int n, int D, float* partials, double* record,
                                 sard_stream_t stream) {
  SARD_REQUIRE(D > 0 && n >= 0, "norm dims");
  cudaStream_t st = as_stream(stream);
  const int nchunks = ceil_div(n, NORM_CHUNK);
  if (n > 0) {
    int threads = ((D + 31) / 32) * 32;
    if (threads > 256) threads = 256;
    dim3 block(threads, NORM_SUB);
    sard_prof_bytes(4.0 * n * (double)D);
    sard_launch(norm_moments_kernel, nchunks, block, NORM_SUB * 3 * threads * sizeof(float), st, x, ldx, n, D, partials);
    SARD_LAUNCH_OK("norm_moments_kernel");
  }
  sard_launch(norm_merge_chunks_kernel, ceil_div((long long)D * 32, 256), 256, 0, st, partials, n, D, nchunks, record);
  SARD_LAUNCH_OK("norm_merge_chunks_kernel");
  return 0;
}

__global__ void norm_update_kernel(const double* __restrict__ records, int n_records, int D,
                                   float* __restrict__ mean, float* __restrict__ var, float* __restrict__ std,
                                   long long* __restrict__ count, float* __restrict__ inv) {
  pdl_sync();
  norm_update_body(records, n_records, D, mean, var, std, count, inv);
}

extern "C" int sard_norm_update(const double* records, int n_records, int D, float* mean, float* var, float* std,
                                long long* count, float* inv, sard_stream_t stream) {
  SARD_REQUIRE(n_records >= 1, "n_records");
  sard_launch(norm_update_kernel, 1, 256, 0, as_stream(stream), records, n_records, D, mean, var, std, count, inv);
  SARD_LAUNCH_OK("norm_update_kernel");
  return 0;
}

__global__ void norm_prepare_eval_kernel(const float* __restrict__ std, int D, float* __restrict__ inv) {
  pdl_sync();
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < D; c += gridDim.x * blockDim.x)
    inv[c] = 1.0f / (std[c] + 1e-8f);
}

extern "C" int sard_norm_prepare_eval(const float* std, int D, float* inv, sard_stream_t stream) {
  sard_launch(norm_prepare_eval_kernel, ceil_div(D, 256), 256, 0, as_stream(stream), std, D, inv);
  SARD_LAUNCH_OK("norm_prepare_eval_kernel");
  return 0;
}

__global__ void norm_apply_kernel(float* __restrict__ x, int ldx, int n, int D, const float* __restrict__ mean,
                                  const float* __restrict__ inv, const long long* __restrict__ count, float clip) {
  pdl_sync();
  if (*count <= 0) return;
  const long long total = (long long)n * D;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(idx / D), c = (int)(idx - (long long)r * D);
    float* p = x + (size_t)r * ldx + c;
    float v = (*p - __ldg(mean + c)) * __ldg(inv + c);
    *p = fminf(fmaxf(v, -clip), clip);
  }
}

extern "C" int sard_norm_apply(float* x, int ldx, int n, int D, const float* mean, const float* inv,
                               const long long* count, float clip, sard_stream_t stream) {
  if (n <= 0) return 0;
  sard_prof_bytes(8.0 * n * (double)D);
  sard_launch(norm_apply_kernel, grid_for((long long)n * D, 256), 256, 0, as_stream(stream), x, ldx, n, D, mean, inv, count, clip);
  SARD_LAUNCH_OK("norm_apply_kernel");
  return 0;
}
