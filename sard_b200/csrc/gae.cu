// GAE as a reverse segmented scan (khrylib/rl/core/common.py:5-25).
//   delta_i = r_i + gamma * V_{i+1} * m_i - V_i          (V_T := 0)
//   A_i     = delta_i + gamma * tau * m_i * A_{i+1}      (A_T := 0; m_i = 0 at an episode end resets)
// The recurrence is affine in A_{i+1}, so a block of elements composes to one (a, b) pair and the
// whole batch is a three-phase suffix scan.  Arithmetic is fp64 like the reference's tensors; the
// fp32 entry point only rounds inputs/outputs.  Normalisation uses the unbiased std (torch.std).
#include "common.cuh"

#define GAE_PER_THREAD 8
#define GAE_THREADS 256
#define GAE_PER_BLOCK (GAE_PER_THREAD * GAE_THREADS)

struct Aff { double a, b; };                       // x -> b + a * x
__device__ __forceinline__ Aff aff_compose(Aff left, Aff right) {   // left applied after right
  Aff r; r.a = left.a * right.a; r.b = left.b + left.a * right.b; return r;
}

template <typename T>
__device__ __forceinline__ Aff thread_chunk(const T* __restrict__ r, const T* __restrict__ m, const T* __restrict__ v,
                                            int s, int e, int Tn, double gamma, double gt) {
  Aff f; f.a = 1.0; f.b = 0.0;
  for (int i = e - 1; i >= s; --i) {
    const double mi = (double)m[i];
    const double vn = (i + 1 < Tn) ? (double)v[i + 1] : 0.0;
    const double delta = (double)r[i] + gamma * vn * mi - (double)v[i];
    const double c = gt * mi;
    f.b = delta + c * f.b;
    f.a = c * f.a;
  }
  return f;
}

// inclusive suffix scan of the per-thread maps within the block; returns the map of threads [t, end)
__device__ __forceinline__ Aff block_suffix_scan(Aff mine, Aff* sh) {
  const int t = threadIdx.x;
  sh[t] = mine;
  __syncthreads();
  for (int o = 1; o < GAE_THREADS; o <<= 1) {
    Aff cur = sh[t];
    Aff nxt; nxt.a = 1.0; nxt.b = 0.0;
    if (t + o < GAE_THREADS) nxt = sh[t + o];
    __syncthreads();
    sh[t] = aff_compose(cur, nxt);
    __syncthreads();
  }
  return sh[t];
}

template <typename T>
__global__ void gae_block_maps(const T* r, const T* m, const T* v, int Tn, double gamma, double gt, double* blk) {
  pdl_sync();
  __shared__ Aff sh[GAE_THREADS];
  const int s = blockIdx.x * GAE_PER_BLOCK + threadIdx.x * GAE_PER_THREAD;
  const int e = min(Tn, s + GAE_PER_THREAD);
  Aff f = thread_chunk(r, m, v, min(s, Tn), max(e, min(s, Tn)), Tn, gamma, gt);
  Aff all = block_suffix_scan(f, sh);
  if (threadIdx.x == 0) { blk[2 * blockIdx.x] = all.a; blk[2 * blockIdx.x + 1] = all.b; }
}

__global__ void gae_block_carry(int nblk, const double* blk, double* carry) {
  pdl_sync();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double c = 0.0;
  for (int b = nblk - 1; b >= 0; --b) {
    carry[b] = c;
    c = blk[2 * b + 1] + blk[2 * b] * c;
  }
}

template <typename T>
__global__ void gae_apply(const T* r, const T* m, const T* v, int Tn, double gamma, double gt, const double* carry,
                          T* adv, T* ret, double* blk_sum) {
  pdl_sync();
  __shared__ Aff sh[GAE_THREADS];
  __shared__ double red[GAE_THREADS];
  const int s0 = blockIdx.x * GAE_PER_BLOCK + threadIdx.x * GAE_PER_THREAD;
  const int s = min(s0, Tn), e = max(min(Tn, s0 + GAE_PER_THREAD), s);
  Aff f = thread_chunk(r, m, v, s, e, Tn, gamma, gt);
  block_suffix_scan(f, sh);
  // incoming advantage = everything to the right of this thread's chunk applied to the block carry
  Aff right; right.a = 1.0; right.b = 0.0;
  if (threadIdx.x + 1 < GAE_THREADS) right = sh[threadIdx.x + 1];
  double a_next = right.b + right.a * carry[blockIdx.x];
  double local = 0.0;
  for (int i = e - 1; i >= s; --i) {
    const double mi = (double)m[i];
    const double vn = (i + 1 < Tn) ? (double)v[i + 1] : 0.0;
    const double delta = (double)r[i] + gamma * vn * mi - (double)v[i];
    a_next = delta + gt * mi * a_next;
    adv[i] = (T)a_next;
    ret[i] = (T)((double)v[i] + a_next);
    local += a_next;
  }
  red[threadIdx.x] = local;
  __syncthreads();
  for (int o = GAE_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) blk_sum[blockIdx.x] = red[0];
}

template <typename T>
__global__ void gae_var_partial(const T* adv, int Tn, int nblk, const double* blk_sum, double* blk_sq) {
  pdl_sync();
  __shared__ double red[GAE_THREADS];
  __shared__ double mean_s;
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int b = 0; b < nblk; ++b) s += blk_sum[b];
    mean_s = s / (double)Tn;
  }
  __syncthreads();
  const double mean = mean_s;
  double local = 0.0;
  const int s0 = blockIdx.x * GAE_PER_BLOCK;
  for (int i = s0 + threadIdx.x; i < min(Tn, s0 + GAE_PER_BLOCK); i += GAE_THREADS) {
    const double d = (double)adv[i] - mean;
    local += d * d;
  }
  red[threadIdx.x] = local;
  __syncthreads();
  for (int o = GAE_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) blk_sq[blockIdx.x] = red[0];
}

template <typename T>
__global__ void gae_normalise(T* adv, int Tn, int nblk, const double* blk_sum, const double* blk_sq) {
  pdl_sync();
  __shared__ double mean_s, inv_std_s;
  if (threadIdx.x == 0) {
    double s = 0.0, q = 0.0;
    for (int b = 0; b < nblk; ++b) { s += blk_sum[b]; q += blk_sq[b]; }
    mean_s = s / (double)Tn;
    inv_std_s = 1.0 / sqrt(q / (double)(Tn - 1));
  }
  __syncthreads();
  const int s0 = blockIdx.x * GAE_PER_BLOCK;
  for (int i = s0 + threadIdx.x; i < min(Tn, s0 + GAE_PER_BLOCK); i += GAE_THREADS)
    adv[i] = (T)(((double)adv[i] - mean_s) * inv_std_s);
}

template <typename T>
static int gae_impl(const T* rewards, const T* masks, const T* values, int Tn, double gamma, double tau,
                    T* advantages, T* returns, double* scratch, sard_stream_t stream) {
  if (Tn <= 0) return 0;
  cudaStream_t st = as_stream(stream);
  const int nblk = ceil_div(Tn, GAE_PER_BLOCK);
  double* blk = scratch;                 // [2*nblk]
  double* carry = scratch + 2 * nblk;    // [nblk]
  double* blk_sum = carry + nblk;        // [nblk]  (blk is dead after gae_block_carry and is reused for blk_sq)
  double* blk_sq = blk;
  const double gt = gamma * tau;
  sard_launch(gae_block_maps<T>, nblk, GAE_THREADS, 0, st, rewards, masks, values, Tn, gamma, gt, blk);
  SARD_LAUNCH_OK("gae_block_maps");
  sard_launch(gae_block_carry, 1, 32, 0, st, nblk, blk, carry);
  SARD_LAUNCH_OK("gae_block_carry");
  sard_launch(gae_apply<T>, nblk, GAE_THREADS, 0, st, rewards, masks, values, Tn, gamma, gt, carry, advantages, returns, blk_sum);
  SARD_LAUNCH_OK("gae_apply");
  sard_launch(gae_var_partial<T>, nblk, GAE_THREADS, 0, st, advantages, Tn, nblk, blk_sum, blk_sq);
  SARD_LAUNCH_OK("gae_var_partial");
  sard_launch(gae_normalise<T>, nblk, GAE_THREADS, 0, st, advantages, Tn, nblk, blk_sum, blk_sq);
  SARD_LAUNCH_OK("gae_normalise");
  return 0;
}

extern "C" int sard_gae(const float* rewards, const float* masks, const float* values, int T, double gamma, double tau,
                        float* advantages, float* returns, double* scratch, sard_stream_t stream) {
  return gae_impl<float>(rewards, masks, values, T, gamma, tau, advantages, returns, scratch, stream);
}
extern "C" int sard_gae_f64(const double* rewards, const double* masks, const double* values, int T, double gamma,
                            double tau, double* advantages, double* returns, double* scratch, sard_stream_t stream) {
  return gae_impl<double>(rewards, masks, values, T, gamma, tau, advantages, returns, scratch, stream);
}
