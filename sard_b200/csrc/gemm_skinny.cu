// Skinny linear layers: the value head (448 -> 1) and the last JSMLP layer (128 -> 1..8 outputs per body
// index).  With at most 8 outputs a tiled GEMM wastes 7/8 of its tile and, worse, spends ~20 us of fixed
// latency walking k-blocks; these are bandwidth-shaped, so each is written as a streaming kernel:
//   forward        one warp per row: lanes hold the row as float4, N dot products, shuffle reduce
//   backward-data  one thread per (row, 4 output columns): R <= 8 multiply-adds
//   backward-weight  (row-lane, float4 column) threads stream a chunk once, N1 sums in registers,
//                    row-lanes reduced through shared memory in lane order (deterministic)
// -1 from the sard_*_skinny entry points means "shape not covered, use the tiled kernels".
#include "common.cuh"

#define SKINNY_ROWS 16
static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ------------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(256) skinny_fwd_kernel(const SardGemm g) {
  pdl_sync();
  int row_start, row_end, group = 0;
  if (g.tiles) {
    if ((int)blockIdx.x >= *g.num_tiles) return;
    const int4 t = reinterpret_cast<const int4*>(g.tiles)[blockIdx.x];
    row_start = t.x; row_end = t.y; group = t.z;
  } else {
    row_start = blockIdx.x * 128;
    row_end = min(g.M, row_start + 128);
    if (row_start >= g.M) return;
  }
  // blockIdx.y: 16-row slice of the 128-row tile (the head has only M/128 = 16 tiles: slices fill the SMs)
  row_start += blockIdx.y * SKINNY_ROWS;
  row_end = min(row_end, row_start + SKINNY_ROWS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* W = g.B0 + (size_t)group * g.b_group_stride;
  const float* bias = g.bias ? g.bias + (size_t)group * g.bias_group_stride : nullptr;
  const int R4 = g.R >> 2;
  for (int row = row_start + warp; row < row_end; row += 8) {
    const float4* a = reinterpret_cast<const float4*>(g.A + (size_t)(g.a_rows ? g.a_rows[row] : row) * g.lda);
    float acc[8];
#pragma unroll
    for (int n = 0; n < 8; ++n) acc[n] = 0.f;
    for (int c = lane; c < R4; c += 32) {
      const float4 x = __ldg(a + c);
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        if (n < g.N) {
          const float4 w = __ldg(reinterpret_cast<const float4*>(W + (size_t)n * g.ldb0) + c);
          acc[n] = fmaf(x.x, w.x, fmaf(x.y, w.y, fmaf(x.z, w.z, fmaf(x.w, w.w, acc[n]))));
        }
      }
    }
    float mine = 0.f;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      if (n < g.N) {
        const float s = warp_sum(acc[n]);
        if (lane == n) mine = s;
      }
    }
    if (lane < g.N) {
      const long long yrow = g.y_rows ? g.y_rows[row] : row;
      float v = mine;
      if (bias) v += __ldg(bias + lane);
      v = epi_forward(v, g.act, g.post, g.Ypre ? g.Ypre + yrow * g.ldy + lane : nullptr);
      g.Y[yrow * g.ldy + lane] = v;
    }
  }
}

int sard_gemm_skinny_fwd(const SardGemm* g, cudaStream_t st) {
  if (g->N > 8 || g->B1 || g->R % 4 || g->lda % 4 || g->ldb0 % 4 || g->b_group_stride % 4 || !al16(g->A) || !al16(g->B0))
    return -1;
  const int tiles = g->tiles ? g->max_tiles : ceil_div(g->M, 128);
  if (tiles <= 0) return 0;
  sard_launch(skinny_fwd_kernel, dim3(tiles, 128 / SKINNY_ROWS), 256, 0, st, *g);
  SARD_LAUNCH_OK("skinny_fwd_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------------ backward-data
__global__ void __launch_bounds__(256) skinny_bwd_kernel(const SardGemm g) {
  pdl_sync();
  int row_start, row_end, group = 0;
  if (g.tiles) {
    if ((int)blockIdx.x >= *g.num_tiles) return;
    const int4 t = reinterpret_cast<const int4*>(g.tiles)[blockIdx.x];
    row_start = t.x; row_end = t.y; group = t.z;
  } else {
    row_start = blockIdx.x * 128;
    row_end = min(g.M, row_start + 128);
    if (row_start >= g.M) return;
  }
  row_start += blockIdx.y * SKINNY_ROWS;
  row_end = min(row_end, row_start + SKINNY_ROWS);
  if (row_start >= row_end) return;
  const float* W = g.B0 + (size_t)group * g.b_group_stride;
  const int N4 = g.N >> 2;
  const int total = (row_end - row_start) * N4;
  for (int idx = threadIdx.x; idx < total; idx += 256) {
    const int m = idx / N4, c = (idx - m * N4) * 4;
    const int row = row_start + m;
    const float* dz = g.A + (size_t)(g.a_rows ? g.a_rows[row] : row) * g.lda;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (r < g.R) {
        const float d = __ldg(dz + r);
        const float4 w = __ldg(reinterpret_cast<const float4*>(W + (size_t)r * g.ldb0 + c));
        v.x = fmaf(d, w.x, v.x); v.y = fmaf(d, w.y, v.y); v.z = fmaf(d, w.z, v.z); v.w = fmaf(d, w.w, v.w);
      }
    }
    const long long yrow = g.y_rows ? g.y_rows[row] : row;
    if (g.Dact) {
      const float4 y = *reinterpret_cast<const float4*>(g.Dact + yrow * g.ldd + c);
      v.x *= act_grad_from_output(y.x, c + 0 < g.dsplit ? g.dact0 : g.dact1);
      v.y *= act_grad_from_output(y.y, c + 1 < g.dsplit ? g.dact0 : g.dact1);
      v.z *= act_grad_from_output(y.z, c + 2 < g.dsplit ? g.dact0 : g.dact1);
      v.w *= act_grad_from_output(y.w, c + 3 < g.dsplit ? g.dact0 : g.dact1);
    }
    *reinterpret_cast<float4*>(g.Y + yrow * g.ldy + c) = v;
  }
}

int sard_gemm_skinny_bwd(const SardGemm* g, cudaStream_t st) {
  if (g->R > 8 || g->B1 || g->N % 4 || g->ldy % 4 || g->ldb0 % 4 || g->b_group_stride % 4 || !al16(g->Y) || !al16(g->B0))
    return -1;
  if (g->Dact && (g->ldd % 4 || !al16(g->Dact))) return -1;
  const int tiles = g->tiles ? g->max_tiles : ceil_div(g->M, 128);
  if (tiles <= 0) return 0;
  sard_launch(skinny_bwd_kernel, dim3(tiles, 128 / SKINNY_ROWS), 256, 0, st, *g);
  SARD_LAUNCH_OK("skinny_bwd_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------------ backward-weight
// block (32, 8): threadIdx.x = float4 column inside a 128-column slab, threadIdx.y = row lane
__global__ void __launch_bounds__(256) skinny_wgrad_kernel(const SardWGrad w) {
  pdl_sync();
  extern __shared__ float red[];                    // [8 row lanes][N1][128]
  int rs, re;
  if (w.chunks) {
    if ((int)blockIdx.x >= *w.num_chunks) return;
    const int4 t = reinterpret_cast<const int4*>(w.chunks)[blockIdx.x];
    rs = t.x; re = t.y;
  } else {
    rs = blockIdx.x * w.chunk_rows;
    re = min(w.M, rs + w.chunk_rows);
    if (rs >= w.M) return;
  }
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int c = blockIdx.y * 128 + tx * 4;
  const bool col_ok = c < w.N2;                     // N2 % 4 == 0
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
  float bsum = 0.f;
  const bool do_bias = blockIdx.y == 0 && tx < w.N1;
#pragma unroll 2
  for (int row = rs + ty; row < re; row += 8) {
    const float* pz = w.P + (size_t)(w.p_rows ? w.p_rows[row] : row) * w.ldp;
    const float4 x = col_ok ? __ldg(reinterpret_cast<const float4*>(w.Q + (size_t)(w.q_rows ? w.q_rows[row] : row) * w.ldq + c))
                            : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < w.N1) {
        const float d = __ldg(pz + i);
        acc[i][0] = fmaf(d, x.x, acc[i][0]); acc[i][1] = fmaf(d, x.y, acc[i][1]);
        acc[i][2] = fmaf(d, x.z, acc[i][2]); acc[i][3] = fmaf(d, x.w, acc[i][3]);
      }
    }
    if (do_bias) bsum += __ldg(pz + tx);
  }
  // reduce the 8 row lanes in lane order
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (i < w.N1) *reinterpret_cast<float4*>(red + ((size_t)(ty * w.N1 + i) * 128 + tx * 4)) =
        make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  }
  __shared__ float bred[8][8];
  if (do_bias) bred[ty][tx] = bsum;
  __syncthreads();
  const int ldpart = SARD_WG_LDPART(w.N2);
  float* part = w.partials + (size_t)blockIdx.x * w.N1 * ldpart;
  for (int e = ty * 32 + tx; e < w.N1 * 128; e += 256) {
    const int i = e >> 7, cc = e & 127;
    if (blockIdx.y * 128 + cc < w.N2) {
      float s = 0.f;
#pragma unroll
      for (int l = 0; l < 8; ++l) s += red[(size_t)(l * w.N1 + i) * 128 + cc];
      part[(size_t)i * ldpart + blockIdx.y * 128 + cc] = s;
    }
  }
  if (do_bias && ty == 0) {
    float s = 0.f;
#pragma unroll
    for (int l = 0; l < 8; ++l) s += bred[l][tx];
    part[(size_t)tx * ldpart + w.N2] = s;
  }
}

int sard_wgrad_skinny(const SardWGrad* w, int nchunks, cudaStream_t st) {
  if (w->N1 > 8 || w->N2 % 4 || w->ldq % 4 || !al16(w->Q)) return -1;
  dim3 grid(nchunks, ceil_div(w->N2, 128));
  sard_launch(skinny_wgrad_kernel, grid, dim3(32, 8), sizeof(float) * 8 * w->N1 * 128, st, *w);
  SARD_LAUNCH_OK("skinny_wgrad_kernel");
  return 0;
}
