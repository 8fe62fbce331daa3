// Linear layers of the SARD nets: fp32 SIMT tiled GEMMs with fused prologues/epilogues.
//
//   forward        Y  = post(act(A W^T + b))        A [M,R] rows optionally gathered, W [N,R]
//   backward-data  dA = (dY W) * act'(saved output)  same weight storage read transposed
//   backward-weight dW += dY^T A, db += colsum(dY)   per-chunk partial products + ordered reduction
//
// Three things make these specific to the path rather than library GEMMs: (1) the weight operand can
// be the concatenation [lin_l | lin_r] of a GraphConv layer without materialising it, (2) row tiles
// can carry a body index selecting the weight slab of an IndexLinear (grouped GEMM over rows sorted
// by index), (3) rows can be gathered/scattered through index vectors (root readout, un-sorting).
// fp32 FFMA is used because the parity target is rtol 1e-4 against an fp64 oracle; see DESIGN.md
// for the tcgen05 3xTF32 plan.
#include "common.cuh"
#include <stdlib.h>

// Shared epilogue of both engines, executed from a ROLLED loop over a tile staged in shared memory:
// one copy of the bias / activation / sin / act' / scatter code per kernel (keeps the SASS small enough
// for the instruction cache: the first version unrolled it 32x and every GEMM kernel exceeded 50 KB)
// and coalesced global stores (consecutive threads -> consecutive output columns).
template <bool NT>
__device__ __forceinline__ void gemm_epilogue_store(const SardGemm& g, float v, long long yrow, int gc, const float* bias) {
  if (NT) {
    if (bias) v += __ldg(bias + gc);
    v = epi_forward(v, g.act, g.post, g.Ypre ? g.Ypre + yrow * g.ldy + gc : nullptr);
  } else if (g.Dact) {
    v *= act_grad_from_output(g.Dact[yrow * g.ldd + gc], gc < g.dsplit ? g.dact0 : g.dact1);
  }
  g.Y[yrow * g.ldy + gc] = v;
}

template <int BM, int BN, int BK, int TM, int TN, bool NT>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
gemm_kernel(const SardGemm g) {
  pdl_sync();
  constexpr int NTH = (BM / TM) * (BN / TN);
  constexpr int EA = BM * BK / NTH;
  constexpr int EB = BN * BK / NTH;
  static_assert(BM * BK % NTH == 0 && BN * BK % NTH == 0, "tile/thread mismatch");
  static_assert(TM % 4 == 0, "TM");
  constexpr int AS = BK * (BM + 4), BS = BK * (BN + 4);          // floats per buffer
  constexpr int TILE_FLOATS = 2 * (AS + BS);
  constexpr int C_FLOATS = BM * (BN + 1);
  __shared__ __align__(16) float smem[TILE_FLOATS > C_FLOATS ? TILE_FLOATS : C_FLOATS];
  float* As = smem;                 // [2][BK][BM+4]
  float* Bs = smem + 2 * AS;        // [2][BK][BN+4]

  int row_start, row_end, group = 0;
  if (g.tiles) {
    if ((int)blockIdx.x >= *g.num_tiles) return;
    const int4 t = reinterpret_cast<const int4*>(g.tiles)[blockIdx.x];
    row_start = t.x; row_end = t.y; group = t.z;
  } else {
    row_start = blockIdx.x * BM;
    row_end = min(g.M, row_start + BM);
    if (row_start >= g.M) return;
  }
  const int n0 = blockIdx.y * BN;
  const int tid = threadIdx.x;
  const float* B0 = g.B0 + (size_t)group * g.b_group_stride;
  const float* B1 = g.B1 ? g.B1 + (size_t)group * g.b_group_stride : nullptr;

  // per-thread A rows are fixed across k-blocks: resolve the gather once
  long long a_off[EA];
#pragma unroll
  for (int i = 0; i < EA; ++i) {
    const int grow = row_start + (tid + i * NTH) / BK;
    a_off[i] = grow < row_end ? (long long)(g.a_rows ? g.a_rows[grow] : grow) * g.lda : -1;
  }
  const int ty = tid / (BN / TN), tx = tid % (BN / TN);
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float ra[EA], rb[EB];
  const int nkb = (g.R + BK - 1) / BK;
  int buf = 0;
  // software pipeline with a single call site for load / compute / store: iteration kb loads tile kb into
  // registers, computes tile kb-1 from smem[buf^1], then stores tile kb into smem[buf]
#pragma unroll 1
  for (int kb = 0; kb <= nkb; ++kb) {
    if (kb < nkb) {
      const int k0 = kb * BK;
#pragma unroll
      for (int i = 0; i < EA; ++i) {
        const int kk = k0 + (tid + i * NTH) % BK;
        ra[i] = (a_off[i] >= 0 && kk < g.R) ? __ldg(g.A + a_off[i] + kk) : 0.f;
      }
#pragma unroll
      for (int i = 0; i < EB; ++i) {
        const int idx = tid + i * NTH;
        const int r = NT ? idx % BK : idx / BN, c = NT ? idx / BK : idx % BN;
        const int kk = k0 + r, gc = n0 + c;
        float v = 0.f;
        if (kk < g.R && gc < g.N) {
          if (NT) v = kk < g.b_split ? __ldg(B0 + (size_t)gc * g.ldb0 + kk) : __ldg(B1 + (size_t)gc * g.ldb1 + (kk - g.b_split));
          else    v = gc < g.b_split ? __ldg(B0 + (size_t)kk * g.ldb0 + gc) : __ldg(B1 + (size_t)kk * g.ldb1 + (gc - g.b_split));
        }
        rb[i] = v;
      }
    }
    if (kb > 0) {
      const float* Ac = As + (buf ^ 1) * AS + ty * TM;
      const float* Bc = Bs + (buf ^ 1) * BS + tx * TN;
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        float a[TM], b[TN];
#pragma unroll
        for (int i = 0; i < TM; i += 4) {
          const float4 v = *reinterpret_cast<const float4*>(Ac + k * (BM + 4) + i);
          a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
        }
        if constexpr (TN % 4 == 0) {
#pragma unroll
          for (int j = 0; j < TN; j += 4) {
            const float4 v = *reinterpret_cast<const float4*>(Bc + k * (BN + 4) + j);
            b[j] = v.x; b[j + 1] = v.y; b[j + 2] = v.z; b[j + 3] = v.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < TN; ++j) b[j] = Bc[k * (BN + 4) + j];
        }
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
    }
    if (kb < nkb) {
#pragma unroll
      for (int i = 0; i < EA; ++i) {
        const int idx = tid + i * NTH;
        As[buf * AS + (idx % BK) * (BM + 4) + idx / BK] = ra[i];
      }
#pragma unroll
      for (int i = 0; i < EB; ++i) {
        const int idx = tid + i * NTH;
        if (NT) Bs[buf * BS + (idx % BK) * (BN + 4) + idx / BK] = rb[i];
        else    Bs[buf * BS + (idx / BN) * (BN + 4) + idx % BN] = rb[i];
      }
    }
    __syncthreads();
    buf ^= 1;
  }

  // stage the accumulators in shared memory (the operand tiles are dead), then a rolled, coalesced epilogue
  float* Cs = smem;
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) Cs[(ty * TM + i) * (BN + 1) + tx * TN + j] = acc[i][j];
  __syncthreads();
  const float* bias = g.bias ? g.bias + (size_t)group * g.bias_group_stride : nullptr;
  const int rows = row_end - row_start;
#pragma unroll 1
  for (int idx = tid; idx < rows * BN; idx += NTH) {
    const int m = idx / BN, c = idx % BN;
    const int gc = n0 + c;
    if (gc >= g.N) continue;
    const int grow = row_start + m;
    const long long yrow = g.y_rows ? g.y_rows[grow] : grow;
    gemm_epilogue_store<NT>(g, Cs[m * (BN + 1) + c], yrow, gc, bias);
  }
}

int sard_umma_gemm(const SardGemm* g, bool nt, cudaStream_t st);   // umma.cu; -1 = shape not handled
int sard_umma_wgrad(const SardWGrad* w, int nchunks, cudaStream_t st);
int sard_gemm_fast(const SardGemm* g, bool nt, cudaStream_t st);      // gemm_fast.cu; -1 = not covered
int sard_wgrad_fast(const SardWGrad* w, int nchunks, cudaStream_t st);
int sard_gemm_mma(const SardGemm* g, bool nt, cudaStream_t st);       // gemm_mma.cu; -1 = not covered
int sard_gemm_skinny_fwd(const SardGemm* g, cudaStream_t st);        // gemm_skinny.cu; -1 = not covered
int sard_gemm_skinny_bwd(const SardGemm* g, cudaStream_t st);
int g_sard_gemm_engine = 4;   // 0 fp32 FFMA (SIMT) | 1,2 per-layer tcgen05 | 3 mma.sync | 4 (default) fused tcgen05 JSMLP + SIMT elsewhere
extern "C" int sard_set_gemm_engine(int engine) { g_sard_gemm_engine = engine; return 0; }
extern "C" int sard_get_gemm_engine(void) { return g_sard_gemm_engine; }

template <bool NT>
static int launch_gemm(const SardGemm* g, sard_stream_t stream) {
  if (g->M <= 0 || g->N <= 0) return 0;
  SARD_REQUIRE(g->R > 0, "gemm: empty reduction");
  SARD_REQUIRE(g->B1 != nullptr || g->b_split >= (NT ? g->R : g->N), "gemm: b_split without second source");
  cudaStream_t st = as_stream(stream);
  const double groups = g->b_group_stride ? 13.0 : 1.0;   // weight slabs touched (live indices, nominal)
  ProfScope prof(st, 2.0 * g->M * (double)g->N * g->R,
                 4.0 * ((double)g->M * g->R + (double)g->M * g->N + groups * (double)g->N * g->R));
  if (g->agg_cols > 0) return sard_gemm_fast(g, NT, st);      // fused neighbour sum: one implementation
  if (g_sard_gemm_engine == 1 || g_sard_gemm_engine == 2) {
    const int r = sard_umma_gemm(g, NT, st);
    if (r >= 0) return r;
  }
  if (g_sard_gemm_engine == 3) {
    const int r = sard_gemm_mma(g, NT, st);
    if (r >= 0) return r;
  }
  if (NT ? g->N <= 8 : g->R <= 8) {                   // value head, last JSMLP layer: streaming kernels
    const int r = NT ? sard_gemm_skinny_fwd(g, st) : sard_gemm_skinny_bwd(g, st);
    if (r >= 0) return r;
  }
  {
    const int r = sard_gemm_fast(g, NT, st);
    if (r >= 0) return r;
  }
  if (g->N <= 16) {
    constexpr int BM = 128, BN = 16;
    const int tiles = g->tiles ? g->max_tiles : ceil_div(g->M, BM);
    if (tiles <= 0) return 0;
    dim3 grid(tiles, ceil_div(g->N, BN));
    sard_launch(gemm_kernel<BM, BN, 16, 4, 2, NT>, grid, 256, 0, st, *g);
  } else if (!g->tiles && (long long)ceil_div(g->M, 128) * ceil_div(g->N, 64) < 2 * 148) {
    // few large tiles would leave most of the 148 SMs idle: use 64x64 tiles (4x the CTAs)
    constexpr int BM = 64, BN = 64;
    dim3 grid(ceil_div(g->M, BM), ceil_div(g->N, BN));
    sard_launch(gemm_kernel<BM, BN, 16, 4, 4, NT>, grid, 256, 0, st, *g);
  } else {
    constexpr int BM = 128, BN = 64;
    const int tiles = g->tiles ? g->max_tiles : ceil_div(g->M, BM);
    if (tiles <= 0) return 0;
    dim3 grid(tiles, ceil_div(g->N, BN));
    sard_launch(gemm_kernel<BM, BN, 16, 8, 4, NT>, grid, 256, 0, st, *g);
  }
  SARD_LAUNCH_OK(NT ? "gemm_kernel<fwd>" : "gemm_kernel<bwd_data>");
  return 0;
}

extern "C" int sard_linear_fwd(const SardGemm* g, sard_stream_t stream) { return launch_gemm<true>(g, stream); }
extern "C" int sard_linear_bwd_data(const SardGemm* g, sard_stream_t stream) { return launch_gemm<false>(g, stream); }

// ------------------------------------------------------------------------------------------
// Weight gradient: partial[chunk][c1][c2] = sum_{rows of chunk} P[row,c1] Q[row,c2]  (+ colsum(P) at c2==N2)
// ------------------------------------------------------------------------------------------
template <int B1, int B2, int BKM, int T1, int T2>
__global__ void __launch_bounds__((B1 / T1) * (B2 / T2))
wgrad_kernel(const SardWGrad w) {
  pdl_sync();
  constexpr int NTH = (B1 / T1) * (B2 / T2);
  constexpr int EP = B1 * BKM / NTH, EQ = B2 * BKM / NTH;
  __shared__ __align__(16) float Ps[2][BKM][B1 + 4];
  __shared__ __align__(16) float Qs[2][BKM][B2 + 4];
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
  const int c10 = blockIdx.y * B1, c20 = blockIdx.z * B2;
  const int tid = threadIdx.x;
  const int t1 = tid / (B2 / T2), t2 = tid % (B2 / T2);
  float acc[T1][T2], bsum[T1];
#pragma unroll
  for (int i = 0; i < T1; ++i) {
    bsum[i] = 0.f;
#pragma unroll
    for (int j = 0; j < T2; ++j) acc[i][j] = 0.f;
  }
  float rp[EP], rq[EQ];
  auto load_tile = [&](int m0) {
#pragma unroll
    for (int i = 0; i < EP; ++i) {
      const int idx = tid + i * NTH, c = idx % B1, r = idx / B1;
      const int row = m0 + r, gc = c10 + c;
      float v = 0.f;
      if (row < re && gc < w.N1) v = __ldg(w.P + (size_t)(w.p_rows ? w.p_rows[row] : row) * w.ldp + gc);
      rp[i] = v;
    }
#pragma unroll
    for (int i = 0; i < EQ; ++i) {
      const int idx = tid + i * NTH, c = idx % B2, r = idx / B2;
      const int row = m0 + r, gc = c20 + c;
      float v = 0.f;
      if (row < re && gc < w.N2) v = __ldg(w.Q + (size_t)(w.q_rows ? w.q_rows[row] : row) * w.ldq + gc);
      rq[i] = v;
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int i = 0; i < EP; ++i) { const int idx = tid + i * NTH; Ps[buf][idx / B1][idx % B1] = rp[i]; }
#pragma unroll
    for (int i = 0; i < EQ; ++i) { const int idx = tid + i * NTH; Qs[buf][idx / B2][idx % B2] = rq[i]; }
  };
  const int nsteps = (re - rs + BKM - 1) / BKM;
  int buf = 0;
  load_tile(rs);
  store_tile(0);
  __syncthreads();
  for (int s = 0; s < nsteps; ++s) {
    if (s + 1 < nsteps) load_tile(rs + (s + 1) * BKM);
#pragma unroll
    for (int k = 0; k < BKM; ++k) {
      float p[T1], q[T2];
#pragma unroll
      for (int i = 0; i < T1; i += 4) {
        const float4 v = *reinterpret_cast<const float4*>(&Ps[buf][k][t1 * T1 + i]);
        p[i] = v.x; p[i + 1] = v.y; p[i + 2] = v.z; p[i + 3] = v.w;
      }
#pragma unroll
      for (int j = 0; j < T2; j += 4) {
        const float4 v = *reinterpret_cast<const float4*>(&Qs[buf][k][t2 * T2 + j]);
        q[j] = v.x; q[j + 1] = v.y; q[j + 2] = v.z; q[j + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < T1; ++i) {
        bsum[i] += p[i];
#pragma unroll
        for (int j = 0; j < T2; ++j) acc[i][j] = fmaf(p[i], q[j], acc[i][j]);
      }
    }
    if (s + 1 < nsteps) store_tile(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }
  const int ldpart = SARD_WG_LDPART(w.N2);
  float* part = w.partials + (size_t)blockIdx.x * w.N1 * ldpart;
#pragma unroll
  for (int i = 0; i < T1; ++i) {
    const int c1 = c10 + t1 * T1 + i;
    if (c1 >= w.N1) continue;
#pragma unroll
    for (int j = 0; j < T2; ++j) {
      const int c2 = c20 + t2 * T2 + j;
      if (c2 < w.N2) part[(size_t)c1 * ldpart + c2] = acc[i][j];
    }
    if (blockIdx.z == 0 && t2 == 0) part[(size_t)c1 * ldpart + w.N2] = bsum[i];
  }
}

// Ordered reduction of the per-chunk partial products.  A block owns 32 consecutive output elements;
// its 8 warps each sum every 8th chunk (8 independent load streams instead of one serial chain), then
// the 8 slice sums are added in slice order -> deterministic, ~8x shorter latency chain.
#define WR_SLICES 8
__global__ void __launch_bounds__(32 * WR_SLICES) wgrad_reduce_kernel(const SardWGrad w, int nchunks_plain) {
  pdl_sync();
  __shared__ float4 part[WR_SLICES][32];
  const int ldpart = SARD_WG_LDPART(w.N2);             // multiple of 4: a float4 never straddles two rows
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int total4 = w.N1 * ldpart / 4;
  int k0, k1, slot = 0;
  if (w.grp_chunk_ptr) {
    const int u = blockIdx.y;
    k0 = w.grp_chunk_ptr[u]; k1 = w.grp_chunk_ptr[u + 1];
    if (k0 >= k1) return;
    slot = w.slot_of_group ? w.slot_of_group[u] : u;
    if (slot < 0) return;
  } else {
    k0 = 0; k1 = nchunks_plain;
  }
  const size_t stride4 = (size_t)w.N1 * ldpart / 4;
  const float4* P4 = reinterpret_cast<const float4*>(w.partials);
  // grid-stride over blocks of 32 float4 (grouped launches keep the grid small: most groups are empty)
  for (int eb = blockIdx.x; eb * 32 < total4; eb += gridDim.x) {
    const int e4 = eb * 32 + lane;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (e4 < total4) {
      const float4* p = P4 + e4;
      int k = k0 + slice;
      for (; k + 3 * WR_SLICES < k1; k += 4 * WR_SLICES) {
        const float4 a0 = p[(size_t)k * stride4], a1 = p[(size_t)(k + WR_SLICES) * stride4];
        const float4 a2 = p[(size_t)(k + 2 * WR_SLICES) * stride4], a3 = p[(size_t)(k + 3 * WR_SLICES) * stride4];
        s.x += a0.x; s.y += a0.y; s.z += a0.z; s.w += a0.w;
        s.x += a1.x; s.y += a1.y; s.z += a1.z; s.w += a1.w;
        s.x += a2.x; s.y += a2.y; s.z += a2.z; s.w += a2.w;
        s.x += a3.x; s.y += a3.y; s.z += a3.z; s.w += a3.w;
      }
      for (; k < k1; k += WR_SLICES) {
        const float4 a = p[(size_t)k * stride4];
        s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
      }
    }
    __syncthreads();                    // previous iteration's readers of `part` are done
    part[slice][lane] = s;
    __syncthreads();
    if (slice == 0 && e4 < total4) {
#pragma unroll
      for (int j = 1; j < WR_SLICES; ++j) {
        const float4 a = part[j][lane];
        s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
      }
      const int c1 = (e4 * 4) / ldpart, c2 = e4 * 4 - c1 * ldpart;
      const float v[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = c2 + i;
        if (c >= w.N2) {
          if (c == w.N2 && w.db) w.db[(size_t)slot * w.b_slot_stride + c1] += v[i];
        } else if (c < w.w_split) {
          w.dW0[(size_t)slot * w.w_slot_stride + (size_t)c1 * w.ldw0 + c] += v[i];
        } else {
          w.dW1[(size_t)slot * w.w_slot_stride + (size_t)c1 * w.ldw1 + (c - w.w_split)] += v[i];
        }
      }
    }
  }
}

// Rows per chunk for an ungrouped weight gradient: enough chunks to put ~1 CTA on every SM, at least
// 16 rows each (the partial products are reduced in chunk order afterwards).
extern "C" int sard_wgrad_chunk_rows(int M, int N1, int N2) {
  const bool fast = N1 % 64 == 0 && (N2 % 128 == 0 || (N2 >= 16 && N2 < 128));   // wgrad_fast_kernel tiles: 64 x 128 (masked below 128)
  if (N1 <= 8) {                                             // wgrad_skinny_kernel: rows are streamed once per 128 columns
    long long rows = ((long long)M * ceil_div(N2, 128) + 591) / 592;
    rows = (rows + 15) / 16 * 16;
    return (int)(rows < 32 ? 32 : (rows > 1024 ? 1024 : rows));
  }
  const long long tiles = fast ? (long long)(N1 / 64) * ((N2 + 127) / 128)
                               : (long long)ceil_div(N1, N1 <= 16 ? 16 : 64) * ceil_div(N2, 64);
  static const int waves = getenv("SARD_WG_WAVES") ? atoi(getenv("SARD_WG_WAVES")) : 1;   // tuning knob: 1 wave (195.5 -> 193.2 ms, fewer partials)
  long long chunks = ((long long)waves * 148 + tiles - 1) / tiles;
  if (chunks < 1) chunks = 1;
  long long rows = (M + chunks - 1) / chunks;
  rows = (rows + 15) / 16 * 16;
  if (rows < 16) rows = 16;
  if (rows > 1024) rows = 1024;
  return (int)rows;
}

extern "C" int sard_linear_bwd_weight(const SardWGrad* w_in, sard_stream_t stream) {
  if (w_in->M <= 0 || w_in->N1 <= 0 || w_in->N2 <= 0) return 0;
  SardWGrad wl = *w_in;
  SardWGrad* w = &wl;
  SARD_REQUIRE(w->dW1 != nullptr || w->w_split >= w->N2, "wgrad: w_split without second destination");
  SARD_REQUIRE(w->partials != nullptr, "wgrad: partials scratch");
  cudaStream_t st = as_stream(stream);
  int nchunks;
  if (w->chunks) {
    nchunks = w->max_chunks;
  } else {
    if (w->chunk_rows <= 0) w->chunk_rows = sard_wgrad_chunk_rows(w->M, w->N1, w->N2);
    nchunks = ceil_div(w->M, w->chunk_rows);
  }
  if (nchunks <= 0) return 0;
  const double wgroups = w->grp_chunk_ptr ? 13.0 : 1.0;
  ProfScope prof(st, 2.0 * w->M * (double)w->N1 * w->N2,
                 4.0 * ((double)w->M * w->N1 + (double)w->M * w->N2 + wgroups * (double)w->N1 * w->N2));
  int umma_rc = -1;
  if (g_sard_gemm_engine == 1 || g_sard_gemm_engine == 2) {
    umma_rc = sard_umma_wgrad(w, nchunks, st);
    if (umma_rc > 0) return umma_rc;
  }
  if (umma_rc != 0) {
    umma_rc = sard_wgrad_fast(w, nchunks, st);          // aligned shapes: 8x8 register tiles, 16-byte loads
    if (umma_rc > 0) return umma_rc;
  }
  if (umma_rc == 0) {
    // partial products already launched (tensor cores or the fast SIMT kernel)
  } else if (w->N1 <= 16) {
    dim3 grid(nchunks, 1, ceil_div(w->N2, 64));
    sard_launch(wgrad_kernel<16, 64, 16, 4, 4>, grid, 64, 0, st, *w);
  } else {
    dim3 grid(nchunks, ceil_div(w->N1, 64), ceil_div(w->N2, 64));
    sard_launch(wgrad_kernel<64, 64, 16, 4, 4>, grid, 256, 0, st, *w);
  }
  if (umma_rc != 0) SARD_LAUNCH_OK("wgrad_kernel");
  int rblocks = ceil_div((long long)w->N1 * SARD_WG_LDPART(w->N2) / 4, 32);
  if (w->grp_chunk_ptr && rblocks > 64) rblocks = 64;
  dim3 rgrid(rblocks, w->grp_chunk_ptr ? w->n_groups : 1);
  sard_launch(wgrad_reduce_kernel, rgrid, 32 * WR_SLICES, 0, st, *w, w->chunks ? 0 : nchunks);
  SARD_LAUNCH_OK("wgrad_reduce_kernel");
  return 0;
}
