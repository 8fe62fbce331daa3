// Fused ObsEncoder chain (design_opt/models/jsmlp.py:7-20 + process_hfield_obj, policy.py:204-213):
//   e1 = relu(x W1^T + b1),  ext[:, 64k:64k+64] = relu(e1 W2^T + b2)      for the encoders k with a small input
// and its backward, one launch each instead of 6 + 15 GEMM launches per net and minibatch.  In most SARD
// envs the hfield / obj / goal inputs are 1..12 wide, so these layers are launch-latency, not FLOP, bound.
// Encoders with a wide input (hfield 1410 in locomotion_vt / escape) stay on the generic GEMM path.
#include "common.cuh"
#include "encoder.cuh"

#define ENC_H 64
#define ENC_KMAX 32
#define ENC_ROWS 32
#define ENC_BROWS 32

// W2s[j][i] = w2[j * 64 + i]: four 16-byte loads per thread in flight, then the (padded, scalar) stores
__device__ __forceinline__ void fill_w2(float (*W2s)[ENC_H + 1], const float* __restrict__ w2, int tid) {
  float4 v[4];
  if ((reinterpret_cast<uintptr_t>(w2) & 15) == 0) {
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = __ldg(reinterpret_cast<const float4*>(w2) + u * 256 + tid);
  } else {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float* q = w2 + (u * 256 + tid) * 4;
      v[u] = make_float4(__ldg(q), __ldg(q + 1), __ldg(q + 2), __ldg(q + 3));
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int e = (u * 256 + tid) * 4, r = e / ENC_H, c = e % ENC_H;
    W2s[r][c] = v[u].x; W2s[r][c + 1] = v[u].y; W2s[r][c + 2] = v[u].z; W2s[r][c + 3] = v[u].w;
  }
}

__global__ void __launch_bounds__(256) enc_fwd_kernel(const EncFwdP p) {
  pdl_sync();
  __shared__ float W2s[ENC_H][ENC_H + 1];
  __shared__ float h1[ENC_ROWS][ENC_H + 1];
  __shared__ float xs[ENC_ROWS][ENC_KMAX];
  const int k = p.enc_list[blockIdx.y];
  const int K = p.K[k];
  const int r0 = blockIdx.x * ENC_ROWS;
  const int tid = threadIdx.x;
  // every global load of the fill is issued before the first store (a rolled load -> store loop pays the memory latency
  // once per iteration: ncu showed 41 % of this kernel's stall samples on that one store)
  fill_w2(W2s, p.w2[k], tid);
  {
    const float* src = p.src[k];
    for (int i0 = 0; i0 < ENC_ROWS * K; i0 += 4 * 256) {
      float v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * 256 + tid, r = i / K, c = i - r * K;
        v[u] = (i < ENC_ROWS * K && r0 + r < p.B) ? __ldg(src + (size_t)__ldg(p.ids + r0 + r) * K + c) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * 256 + tid, r = i / K, c = i - r * K;
        if (i < ENC_ROWS * K) xs[r][c] = v[u];
      }
    }
  }
  __syncthreads();
  {
    const int j = tid & 63, rg = tid >> 6;                 // 8 rows per thread
    const float bj = __ldg(p.b1[k] + j);
    float w[ENC_KMAX];
#pragma unroll
    for (int c = 0; c < ENC_KMAX; ++c) w[c] = c < K ? __ldg(p.w1[k] + (size_t)j * K + c) : 0.f;
#pragma unroll 1
    for (int rr = 0; rr < 8; ++rr) {
      const int r = rg * 8 + rr;
      float s = bj;
#pragma unroll
      for (int c = 0; c < ENC_KMAX; ++c) s = fmaf(xs[r][c < K ? c : 0], w[c], s);
      s = s > 0.f ? s : 0.f;
      h1[r][j] = s;
      if (r0 + r < p.B) p.e1[k][(size_t)(r0 + r) * ENC_H + j] = s;
    }
  }
  __syncthreads();
  {
    const int j = tid & 63, rg = tid >> 6;
    const float bj = __ldg(p.b2[k] + j);
    float acc[8];
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) acc[rr] = bj;
#pragma unroll 4
    for (int i = 0; i < ENC_H; ++i) {
      const float wv = W2s[j][i];
#pragma unroll
      for (int rr = 0; rr < 8; ++rr) acc[rr] = fmaf(h1[rg * 8 + rr][i], wv, acc[rr]);
    }
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) {
      const int r = r0 + rg * 8 + rr;
      if (r < p.B) p.ext[(size_t)r * p.ld_ext + k * ENC_H + j] = acc[rr] > 0.f ? acc[rr] : 0.f;
    }
  }
}

// per-chunk partial layout: dW2[64*64] | db2[64] | dW1[64*KMAX] | db1[64]
#define ENC_PART (ENC_H * ENC_H + ENC_H + ENC_H * ENC_KMAX + ENC_H)

__global__ void __launch_bounds__(256) enc_bwd_kernel(const EncBwdP p) {
  pdl_sync();
  __shared__ float dz[ENC_BROWS][ENC_H + 1];      // dext tile (relu' already applied by the producer)
  __shared__ float es[ENC_BROWS][ENC_H + 1];      // e1 tile, later de1
  __shared__ float W2s[ENC_H][ENC_H + 1];
  __shared__ float xs[ENC_BROWS][ENC_KMAX];
  const int k = p.enc_list[blockIdx.y];
  const int K = p.K[k];
  const int r0 = blockIdx.x * ENC_BROWS;
  const int tid = threadIdx.x;
  {
    // all loads of the tile fills first, then the stores (see enc_fwd_kernel)
    const float* e1k = p.e1[k];
    const float* src = p.src[k];
    float4 a[2], b[2];
    float xv[4];
    const bool al = ((reinterpret_cast<uintptr_t>(p.dext) | reinterpret_cast<uintptr_t>(e1k)) & 15) == 0 && (p.ld_dext & 3) == 0;
#pragma unroll
    for (int u = 0; u < 2; ++u) {                            // 32 rows x 16 float4 = 512 float4 per array
      const int i = u * 256 + tid, r = i >> 4, c4 = i & 15;
      const bool ok = r0 + r < p.B;
      const float* qa = p.dext + (size_t)(r0 + r) * p.ld_dext + k * ENC_H + c4 * 4;
      const float* qb = e1k + (size_t)(r0 + r) * ENC_H + c4 * 4;
      a[u] = b[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok && al) { a[u] = *reinterpret_cast<const float4*>(qa); b[u] = *reinterpret_cast<const float4*>(qb); }
      else if (ok) { a[u] = make_float4(qa[0], qa[1], qa[2], qa[3]); b[u] = make_float4(qb[0], qb[1], qb[2], qb[3]); }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {                            // 32 rows x 32 columns
      const int i = u * 256 + tid, r = i / ENC_KMAX, c = i % ENC_KMAX;
      xv[u] = (r0 + r < p.B && c < K) ? __ldg(src + (size_t)__ldg(p.ids + r0 + r) * K + c) : 0.f;
    }
    fill_w2(W2s, p.w2[k], tid);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int i = u * 256 + tid, r = i >> 4, c = (i & 15) * 4;
      dz[r][c] = a[u].x; dz[r][c + 1] = a[u].y; dz[r][c + 2] = a[u].z; dz[r][c + 3] = a[u].w;
      es[r][c] = b[u].x; es[r][c + 1] = b[u].y; es[r][c + 2] = b[u].z; es[r][c + 3] = b[u].w;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = u * 256 + tid;
      xs[i / ENC_KMAX][i % ENC_KMAX] = xv[u];
    }
  }
  __syncthreads();
  float* part = p.partials + ((size_t)blockIdx.x * p.n_enc + blockIdx.y) * ENC_PART;
  const int j = tid & 63, g4 = tid >> 6;
  {   // dW2[j][i] = sum_r dz[r][j] e1[r][i], i in [16 g4, 16 g4 + 16);  db2[j] = sum_r dz[r][j]
    float acc[16], bs = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;
#pragma unroll 2
    for (int r = 0; r < ENC_BROWS; ++r) {
      const float d = dz[r][j];
      bs += d;
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = fmaf(d, es[r][g4 * 16 + i], acc[i]);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) part[j * ENC_H + g4 * 16 + i] = acc[i];
    if (g4 == 0) part[ENC_H * ENC_H + j] = bs;
  }
  __syncthreads();
  {   // de1[r][i] = relu'(e1[r][i]) * sum_j dz[r][j] W2[j][i]; thread: i = tid&63, rows 8 g4 .. +8
    const int i = j;
    float acc[8];
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) acc[rr] = 0.f;
#pragma unroll 4
    for (int jj = 0; jj < ENC_H; ++jj) {
      const float wv = W2s[jj][i];
#pragma unroll
      for (int rr = 0; rr < 8; ++rr) acc[rr] = fmaf(dz[g4 * 8 + rr][jj], wv, acc[rr]);
    }
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) {
      const int r = g4 * 8 + rr;
      es[r][i] = es[r][i] > 0.f ? acc[rr] : 0.f;      // (r, i) is owned by exactly this thread; dW2 pass is behind the barrier above
    }
  }
  __syncthreads();
  {   // dW1[j][c] = sum_r de1[r][j] x[r][c], c in [8 g4, 8 g4 + 8);  db1[j] = sum_r de1[r][j]
    float acc[8], bs = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = 0.f;
#pragma unroll 2
    for (int r = 0; r < ENC_BROWS; ++r) {
      const float d = es[r][j];
      bs += d;
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[c] = fmaf(d, xs[r][g4 * 8 + c], acc[c]);
    }
    float* pw1 = part + ENC_H * ENC_H + ENC_H;
#pragma unroll
    for (int c = 0; c < 8; ++c) pw1[j * ENC_KMAX + g4 * 8 + c] = acc[c];
    if (g4 == 0) pw1[ENC_H * ENC_KMAX + j] = bs;
  }
}

__global__ void enc_bwd_reduce_kernel(const EncBwdP p, int n_chunks) {
  pdl_sync();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ENC_PART) return;
  const int k = p.enc_list[blockIdx.y];
  const int K = p.K[k];
  float s = 0.f;
  for (int c = 0; c < n_chunks; ++c) s += p.partials[((size_t)c * p.n_enc + blockIdx.y) * ENC_PART + e];
  if (e < ENC_H * ENC_H) p.gw2[k][e] += s;
  else if (e < ENC_H * ENC_H + ENC_H) p.gb2[k][e - ENC_H * ENC_H] += s;
  else if (e < ENC_H * ENC_H + ENC_H + ENC_H * ENC_KMAX) {
    const int q = e - (ENC_H * ENC_H + ENC_H), jj = q / ENC_KMAX, c = q % ENC_KMAX;
    if (c < K) p.gw1[k][(size_t)jj * K + c] += s;
  } else p.gb1[k][e - (ENC_H * ENC_H + ENC_H + ENC_H * ENC_KMAX)] += s;
}

// ---- host entry points used by the schedules (engine.cu) ----
int sard_enc_fused_ok(int K, int h1, int h2) { return K <= ENC_KMAX && h1 == ENC_H && h2 == ENC_H; }
long long sard_enc_partials_floats(int B, int n_enc) { return (long long)ceil_div(B, ENC_BROWS) * n_enc * ENC_PART; }

int sard_enc_fwd_launch(const EncFwdP& p, cudaStream_t st) {
  if (p.n_enc <= 0 || p.B <= 0) return 0;
  dim3 grid(ceil_div(p.B, ENC_ROWS), p.n_enc);
  sard_launch(enc_fwd_kernel, grid, 256, 0, st, p);
  SARD_LAUNCH_OK("enc_fwd_kernel");
  return 0;
}

int sard_enc_bwd_launch(const EncBwdP& p, cudaStream_t st) {
  if (p.n_enc <= 0 || p.B <= 0) return 0;
  const int n_chunks = ceil_div(p.B, ENC_BROWS);
  dim3 grid(n_chunks, p.n_enc);
  sard_launch(enc_bwd_kernel, grid, 256, 0, st, p);
  SARD_LAUNCH_OK("enc_bwd_kernel");
  dim3 rgrid(ceil_div(ENC_PART, 256), p.n_enc);
  sard_launch(enc_bwd_reduce_kernel, rgrid, 256, 0, st, p, n_chunks);
  SARD_LAUNCH_OK("enc_bwd_reduce_kernel");
  return 0;
}

// ---------------------------------------------------------------------------------------- wide first layer
// First ObsEncoder layer with a WIDE input (height field: K = 1410 in locomotion_vt / escape, jsmlp.py:7-20):
//   e1[b, j] = relu(b1[j] + sum_k x[ids[b], k] W1[j, k]),   M = B (2048) rows, N = 64, K = 1410.
// K is neither a multiple of 16 nor are the rows 16-byte aligned, so the aligned GEMM tiles do not apply, and the
// generic tile kernel sees 16-32 CTAs for the whole product (86 us).  Here the reduction is split over CTAs:
// grid (B / 32 row tiles, WIDE_SPLITS k-ranges), partial sums in split order (deterministic), a finishing kernel adds
// bias + ReLU.  The weight gradient splits the other way: grid (k tiles of 64, row chunks of 256).
#define WIDE_H 64
#define WIDE_KT 64
#define WIDE_ROWS 32
#define WIDE_SPLITS 6
#define WIDE_WG_ROWS 256

#define WIDE_LD 68               // shared-memory row stride: 16-byte aligned rows, float4 reads along k without bank conflicts
__global__ void __launch_bounds__(256, 2) enc_wide_fwd_kernel(const float* __restrict__ src, int K, const int* __restrict__ ids, int B,
                                                              const float* __restrict__ w1, float* __restrict__ partials) {
  pdl_sync();
  __shared__ __align__(16) float xs[WIDE_ROWS][WIDE_LD];
  __shared__ __align__(16) float ws[WIDE_H][WIDE_LD];
  const int r0 = blockIdx.x * WIDE_ROWS, tid = threadIdx.x;
  const int per = ((K + WIDE_SPLITS - 1) / WIDE_SPLITS + WIDE_KT - 1) / WIDE_KT * WIDE_KT;     // k-range of a split, multiple of the k tile
  const int k_lo = blockIdx.y * per, k_hi = min(K, k_lo + per);
  const int tr = tid >> 4, tc = tid & 15;                // rows 2 tr, 2 tr + 1; output columns tc + 16 a
  float acc[2][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int a = 0; a < 4; ++a) acc[i][a] = 0.f;
  // this thread's 8 input rows (fixed over the k loop): the arena row is looked up once
  const float* xrow[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int r = (u * 256 + tid) >> 6;
    xrow[u] = (r0 + r < B) ? src + (size_t)__ldg(ids + r0 + r) * K : nullptr;
  }
  const int cc = tid & 63;
  float xv[8], wv[16];
  auto fetch = [&](int k0) {                             // 32 x 64 input tile and 64 x 64 weight tile, consecutive threads along k
#pragma unroll
    for (int u = 0; u < 8; ++u) xv[u] = (xrow[u] && k0 + cc < k_hi) ? __ldg(xrow[u] + k0 + cc) : 0.f;
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int r = (u * 256 + tid) >> 6;
      wv[u] = (k0 + cc < k_hi) ? __ldg(w1 + (size_t)r * K + k0 + cc) : 0.f;
    }
  };
  if (k_lo < k_hi) fetch(k_lo);
  for (int k0 = k_lo; k0 < k_hi; k0 += WIDE_KT) {
    __syncthreads();                                     // the previous tile has been consumed
#pragma unroll
    for (int u = 0; u < 8; ++u) { const int e = u * 256 + tid; xs[e >> 6][e & 63] = xv[u]; }
#pragma unroll
    for (int u = 0; u < 16; ++u) { const int e = u * 256 + tid; ws[e >> 6][e & 63] = wv[u]; }
    __syncthreads();
    if (k0 + WIDE_KT < k_hi) fetch(k0 + WIDE_KT);        // the next tile's loads fly under this tile's math
#pragma unroll 4
    for (int c = 0; c < WIDE_KT; c += 4) {
      float4 x[2], w[4];
#pragma unroll
      for (int i = 0; i < 2; ++i) x[i] = *reinterpret_cast<const float4*>(&xs[2 * tr + i][c]);
#pragma unroll
      for (int a = 0; a < 4; ++a) w[a] = *reinterpret_cast<const float4*>(&ws[tc + 16 * a][c]);
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          acc[i][a] = fmaf(x[i].x, w[a].x, acc[i][a]); acc[i][a] = fmaf(x[i].y, w[a].y, acc[i][a]);
          acc[i][a] = fmaf(x[i].z, w[a].z, acc[i][a]); acc[i][a] = fmaf(x[i].w, w[a].w, acc[i][a]);
        }
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int r = r0 + 2 * tr + i;
    if (r < B) {
#pragma unroll
      for (int a = 0; a < 4; ++a) partials[((size_t)blockIdx.y * B + r) * WIDE_H + tc + 16 * a] = acc[i][a];
    }
  }
}

__global__ void __launch_bounds__(256) enc_wide_finish_kernel(const float* __restrict__ partials, int B, const float* __restrict__ b1,
                                                              float* __restrict__ e1) {
  pdl_sync();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B * WIDE_H) return;
  float s = __ldg(b1 + (e & 63));
#pragma unroll
  for (int sp = 0; sp < WIDE_SPLITS; ++sp) s += partials[(size_t)sp * B * WIDE_H + e];
  e1[e] = s > 0.f ? s : 0.f;
}

// dW1[j, k] partial over one chunk of rows: de1 is the gradient w.r.t. the layer's pre-activation (relu' applied)
__global__ void __launch_bounds__(256) enc_wide_wgrad_kernel(const float* __restrict__ src, int K, const int* __restrict__ ids, int B,
                                                             const float* __restrict__ de1, float* __restrict__ partials) {
  pdl_sync();
  __shared__ __align__(16) float ds[WIDE_ROWS][WIDE_LD];
  __shared__ __align__(16) float xs[WIDE_ROWS][WIDE_LD];
  const int k0 = blockIdx.x * WIDE_KT, chunk = blockIdx.y, tid = threadIdx.x;
  const int r_lo = chunk * WIDE_WG_ROWS, r_hi = min(B, r_lo + WIDE_WG_ROWS);
  const int tj = (tid >> 4) * 4, tk = (tid & 15) * 4;   // 4 x 4 outputs per thread: rows j = tj.., columns k = tk..
  float acc[4][4], bsum = 0.f;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  float dv[8], xv[8];
  auto fetch = [&](int r0) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int e = u * 256 + tid, r = e >> 6, c = e & 63;
      const bool ok = r0 + r < r_hi;
      dv[u] = ok ? de1[(size_t)(r0 + r) * WIDE_H + c] : 0.f;
      xv[u] = (ok && k0 + c < K) ? __ldg(src + (size_t)__ldg(ids + r0 + r) * K + k0 + c) : 0.f;
    }
  };
  if (r_lo < r_hi) fetch(r_lo);
  for (int r0 = r_lo; r0 < r_hi; r0 += WIDE_ROWS) {
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 8; ++u) { const int e = u * 256 + tid; ds[e >> 6][e & 63] = dv[u]; xs[e >> 6][e & 63] = xv[u]; }
    __syncthreads();
    if (r0 + WIDE_ROWS < r_hi) fetch(r0 + WIDE_ROWS);    // next step's loads under this step's math
#pragma unroll 8
    for (int r = 0; r < WIDE_ROWS; ++r) {
      const float4 d4 = *reinterpret_cast<const float4*>(&ds[r][tj]), x4 = *reinterpret_cast<const float4*>(&xs[r][tk]);
      const float d[4] = {d4.x, d4.y, d4.z, d4.w}, x[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(d[a], x[b], acc[a][b]);
    }
    if (blockIdx.x == 0 && tid < WIDE_H) {               // bias gradient: column sums of de1
#pragma unroll 8
      for (int r = 0; r < WIDE_ROWS; ++r) bsum += ds[r][tid];
    }
  }
  const int ldp = K + 1;                                 // partial row: K weight columns + the bias column
  float* part = partials + (size_t)chunk * WIDE_H * ldp;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
      if (k0 + tk + b < K) part[(size_t)(tj + a) * ldp + k0 + tk + b] = acc[a][b];
  if (blockIdx.x == 0 && tid < WIDE_H) part[(size_t)tid * ldp + K] = bsum;
}

__global__ void __launch_bounds__(256) enc_wide_wgrad_reduce_kernel(const float* __restrict__ partials, int n_chunks, int K,
                                                                    float* __restrict__ gw1, float* __restrict__ gb1) {
  pdl_sync();
  const int ldp = K + 1;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= WIDE_H * ldp) return;
  float s = 0.f;
  for (int c = 0; c < n_chunks; ++c) s += partials[(size_t)c * WIDE_H * ldp + e];      // chunk order: deterministic
  const int jrow = e / ldp, k = e - jrow * ldp;
  if (k < K) gw1[(size_t)jrow * K + k] += s;
  else gb1[jrow] += s;
}

int sard_enc_wide_ok(int K, int h1) { return K > 128 && h1 == WIDE_H; }
long long sard_enc_wide_partials_floats(int B, int K) {
  const long long fwd = (long long)WIDE_SPLITS * B * WIDE_H, bwd = (long long)ceil_div(B, WIDE_WG_ROWS) * WIDE_H * (K + 1);
  return fwd > bwd ? fwd : bwd;
}
int sard_enc_wide_fwd(const float* src, int K, const int* ids, int B, const float* w1, const float* b1, float* e1, float* partials,
                      cudaStream_t st) {
  if (B <= 0) return 0;
  sard_prof_bytes(4.0 * ((double)B * K + 64.0 * K + 64.0 * B));
  sard_launch(enc_wide_fwd_kernel, dim3(ceil_div(B, WIDE_ROWS), WIDE_SPLITS), 256, 0, st, src, K, ids, B, w1, partials);
  SARD_LAUNCH_OK("enc_wide_fwd_kernel");
  sard_launch(enc_wide_finish_kernel, ceil_div((long long)B * WIDE_H, 256), 256, 0, st, (const float*)partials, B, b1, e1);
  SARD_LAUNCH_OK("enc_wide_finish_kernel");
  return 0;
}
int sard_enc_wide_wgrad(const float* src, int K, const int* ids, int B, const float* de1, float* gw1, float* gb1, float* partials,
                        cudaStream_t st) {
  if (B <= 0) return 0;
  const int n_chunks = ceil_div(B, WIDE_WG_ROWS);
  sard_prof_bytes(4.0 * ((double)B * K + 64.0 * K + 64.0 * B));
  sard_launch(enc_wide_wgrad_kernel, dim3(ceil_div(K, WIDE_KT), n_chunks), 256, 0, st, src, K, ids, B, de1, partials);
  SARD_LAUNCH_OK("enc_wide_wgrad_kernel");
  sard_launch(enc_wide_wgrad_reduce_kernel, ceil_div((long long)WIDE_H * (K + 1), 256), 256, 0, st, (const float*)partials, n_chunks, K,
              gw1, gb1);
  SARD_LAUNCH_OK("enc_wide_wgrad_reduce_kernel");
  return 0;
}
