// Fast paths of the fp32 SIMT GEMMs for the aligned shapes that dominate the update (GraphConv, JSMLP,
// critic MLP): 16-byte global loads with per-thread base pointers (no per-element index math or
// bounds branches: out-of-range rows are clamped to a valid row and discarded in the epilogue), 8x8 /
// 8x4 register tiles, register-prefetch double buffering, shared-memory-staged rolled epilogue.
// ncu on the generic kernels showed them issue-bound with only 37-46 % of issued instructions being
// FFMA; these variants are ~90 % FFMA in the main loop.  Anything unaligned falls back to gemm.cu.
#include "common.cuh"
#include <stdlib.h>

// ------------------------------------------------------------------------------------------ forward / backward-data
// arrival counters of the split-K output tiles: one row per stream slot (two split-K GEMMs on different streams must not
// share counters), zero at load, every launch leaves its counters at zero again
#define SPLITK_SLOTS 8
#define SPLITK_MAX_TILES 1024
__device__ unsigned g_splitk_cnt[SPLITK_SLOTS][SPLITK_MAX_TILES];

template <int BN, int TN, bool NT, bool AGG, int MINB, bool SPLITK>
__global__ void __launch_bounds__(256, MINB) gemm_fast_kernel(const SardGemm g, unsigned* splitk_counters) {
  pdl_sync();
  constexpr int BM = 128, BK = 16, TM = 8;
  constexpr int AS = BK * (BM + 4), BS = BK * (BN + 4);
  constexpr int A_LD = BM * BK / 4 / 256;                  // float4 loads per thread: 2
  constexpr int B_LD = BN * BK / 4 / 256;                  // 1 (BN=64) or 2 (BN=128)
  extern __shared__ __align__(16) float smem[];
  float* As = smem;
  float* Bs = smem + 2 * AS;

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

  const float* a_ptr[A_LD];
  // AGG: fused GraphConv neighbour sum (see SardGemm.agg_*): up to 4 neighbour rows per thread-row are kept
  // as pointers, longer lists (rare) are walked from memory; the sum is taken in edge order like
  // csr_aggregate_kernel, so the forward is bit-identical to the unfused path
  const float* nb_ptr[AGG ? A_LD : 1][4];
  int nb_e0[AGG ? A_LD : 1], nb_e1[AGG ? A_LD : 1], nb_gs[AGG ? A_LD : 1];
  float* agg_dst[AGG ? A_LD : 1];
#pragma unroll
  for (int i = 0; i < A_LD; ++i) {
    const int idx = tid + i * 256;
    const int grow = min(row_start + (idx >> 2), row_end - 1);         // clamp: extra rows are never stored
    if (AGG) {
      a_ptr[i] = g.A + (size_t)grow * g.lda + (idx & 3) * 4;
      const int a = __ldg(g.agg_node + grow);
      const int e0 = __ldg(g.agg_ptr + a), e1 = __ldg(g.agg_ptr + a + 1);
      const int gs = __ldg(g.agg_cum + __ldg(g.agg_graph + grow)) - g.agg_node_base;
      nb_e0[i] = e0 + 4; nb_e1[i] = e1; nb_gs[i] = gs;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        nb_ptr[i][q] = e0 + q < e1 ? g.A + (size_t)(gs + __ldg(g.agg_nbr + e0 + q)) * g.lda + (idx & 3) * 4 : nullptr;
      agg_dst[i] = (g.agg_out && blockIdx.y == 0 && row_start + (idx >> 2) < row_end)
                       ? g.agg_out + (size_t)grow * g.ld_agg_out + (idx & 3) * 4 : nullptr;
    } else {
      a_ptr[i] = g.A + (size_t)(g.a_rows ? g.a_rows[grow] : grow) * g.lda + (idx & 3) * 4;
    }
  }
  const int ty = tid >> 4, tx = tid & 15;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float4 ra[A_LD], rb[B_LD];
  // SPLITK: the gridDim.z CTAs of an output tile each walk a slice of the k-blocks and store raw partial tiles to
  // g.splitk_ws; splitk_finish_kernel adds them in slice order and applies the epilogue.  For the critic MLP, whose
  // 2048-row operands give only 16-128 CTAs with 16-32 k-blocks each (65-150 us latency chains on idle SMs).
  const int nkb_all = g.R / BK;
  const int kb_first = SPLITK ? nkb_all * (int)blockIdx.z / (int)gridDim.z : 0;
  const int nkb = SPLITK ? nkb_all * ((int)blockIdx.z + 1) / (int)gridDim.z : nkb_all;
  int buf = 0;
#pragma unroll 1
  for (int kb = kb_first; kb <= nkb; ++kb) {
    if (kb < nkb) {
      const int k0 = kb * BK;
      if (AGG && k0 < g.agg_cols) {
#pragma unroll
        for (int i = 0; i < A_LD; ++i) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (nb_ptr[i][q]) {
              const float4 t = __ldg(reinterpret_cast<const float4*>(nb_ptr[i][q] + k0));
              v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
            }
          }
          for (int e = nb_e0[i]; e < nb_e1[i]; ++e) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(
                g.A + (size_t)(nb_gs[i] + __ldg(g.agg_nbr + e)) * g.lda + ((tid + i * 256) & 3) * 4 + k0));
            v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
          }
          if (agg_dst[i]) *reinterpret_cast<float4*>(agg_dst[i] + k0) = v;
          ra[i] = v;
        }
      } else {
        const int ka = AGG ? k0 - g.agg_cols : k0;
#pragma unroll
        for (int i = 0; i < A_LD; ++i) ra[i] = __ldg(reinterpret_cast<const float4*>(a_ptr[i] + ka));
      }
#pragma unroll
      for (int i = 0; i < B_LD; ++i) {
        const int idx = tid + i * 256;
        if (NT) {            // W[n][k]: 4 consecutive k of output column n
          const int n = idx >> 2, kk = k0 + (idx & 3) * 4;
          rb[i] = kk < g.b_split ? __ldg(reinterpret_cast<const float4*>(B0 + (size_t)(n0 + n) * g.ldb0 + kk))
                                 : __ldg(reinterpret_cast<const float4*>(B1 + (size_t)(n0 + n) * g.ldb1 + (kk - g.b_split)));
        } else {             // W[k][c]: 4 consecutive output columns of reduction row k
          const int r = idx / (BN / 4), c = (idx % (BN / 4)) * 4;
          if (AGG && g.b_ksplit > 0)       // weights split on the reduction index: [W_l ; W_r]
            rb[i] = k0 + r < g.b_ksplit ? __ldg(reinterpret_cast<const float4*>(B0 + (size_t)(k0 + r) * g.ldb0 + n0 + c))
                                        : __ldg(reinterpret_cast<const float4*>(B1 + (size_t)(k0 + r - g.b_ksplit) * g.ldb1 + n0 + c));
          else
            rb[i] = n0 < g.b_split ? __ldg(reinterpret_cast<const float4*>(B0 + (size_t)(k0 + r) * g.ldb0 + n0 + c))
                                   : __ldg(reinterpret_cast<const float4*>(B1 + (size_t)(k0 + r) * g.ldb1 + (n0 - g.b_split) + c));
        }
      }
    }
    if (kb > kb_first) {
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
#pragma unroll
        for (int j = 0; j < TN; j += 4) {
          const float4 v = *reinterpret_cast<const float4*>(Bc + k * (BN + 4) + j);
          b[j] = v.x; b[j + 1] = v.y; b[j + 2] = v.z; b[j + 3] = v.w;
        }
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
    }
    if (kb < nkb) {
#pragma unroll
      for (int i = 0; i < A_LD; ++i) {
        const int idx = tid + i * 256;
        float* d = As + buf * AS + ((idx & 3) * 4) * (BM + 4) + (idx >> 2);
        d[0] = ra[i].x; d[BM + 4] = ra[i].y; d[2 * (BM + 4)] = ra[i].z; d[3 * (BM + 4)] = ra[i].w;
      }
#pragma unroll
      for (int i = 0; i < B_LD; ++i) {
        const int idx = tid + i * 256;
        if (NT) {
          float* d = Bs + buf * BS + ((idx & 3) * 4) * (BN + 4) + (idx >> 2);
          d[0] = rb[i].x; d[BN + 4] = rb[i].y; d[2 * (BN + 4)] = rb[i].z; d[3 * (BN + 4)] = rb[i].w;
        } else {
          *reinterpret_cast<float4*>(Bs + buf * BS + (idx / (BN / 4)) * (BN + 4) + (idx % (BN / 4)) * 4) = rb[i];
        }
      }
    }
    __syncthreads();
    buf ^= 1;
  }

  if (SPLITK) {
    float* ws = g.splitk_ws + (size_t)blockIdx.z * g.M * g.N;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int grow = row_start + ty * TM + i;
      if (grow < row_end) {
#pragma unroll
        for (int j = 0; j < TN; j += 4)
          *reinterpret_cast<float4*>(ws + (size_t)grow * g.N + n0 + tx * TN + j) =
              make_float4(acc[i][j], acc[i][j + 1], acc[i][j + 2], acc[i][j + 3]);
      }
    }
    // The slice that arrives LAST at this output tile finishes it: partial tiles added in slice order (so the result does
    // not depend on which slice that was) + the epilogue.  One launch instead of two on the critic's chain.
    if (!splitk_counters) return;                            // finished by splitk_finish_kernel
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    unsigned* cnt = splitk_counters + blockIdx.y * gridDim.x + blockIdx.x;
    if (tid == 0) {
      const unsigned t = atomicAdd(cnt, 1u);
      s_last = (t == gridDim.z - 1) ? 1 : 0;
      if (s_last) *cnt = 0;                                   // ready for the next launch on this stream
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int rows = row_end - row_start;
    const size_t slice = (size_t)g.M * g.N;
    for (int idx = tid; idx < rows * (BN / 4); idx += 256) {
      const int m = idx / (BN / 4), c = (idx % (BN / 4)) * 4;
      const int grow = row_start + m, gc0 = n0 + c;
      const float* pp = g.splitk_ws + (size_t)grow * g.N + gc0;
      float4 v = __ldcg(reinterpret_cast<const float4*>(pp));
      for (int z = 1; z < (int)gridDim.z; ++z) {
        const float4 t = __ldcg(reinterpret_cast<const float4*>(pp + z * slice));
        v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
      }
      const long long yrow = g.y_rows ? g.y_rows[grow] : grow;
      float o[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int gc = gc0 + i;
        float x = o[i];
        if (NT) {
          if (g.bias) x += __ldg(g.bias + gc);
          x = epi_forward(x, g.act, g.post, g.Ypre ? g.Ypre + yrow * g.ldy + gc : nullptr);
        } else if (g.Dact) {
          x *= act_grad_from_output(g.Dact[yrow * g.ldd + gc], gc < g.dsplit ? g.dact0 : g.dact1);
        }
        g.Y[yrow * g.ldy + gc] = x;
      }
    }
    return;
  }
  float* Cs = smem;                                          // [BM][BN+1]; the operand tiles are dead
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) Cs[(ty * TM + i) * (BN + 1) + tx * TN + j] = acc[i][j];
  __syncthreads();
  const float* bias = g.bias ? g.bias + (size_t)group * g.bias_group_stride : nullptr;
  const int rows = row_end - row_start;
#pragma unroll 1
  for (int idx = tid; idx < rows * BN; idx += 256) {
    const int m = idx / BN, c = idx % BN;
    const int gc = n0 + c;
    const int grow = row_start + m;
    const long long yrow = g.y_rows ? g.y_rows[grow] : grow;
    float v = Cs[m * (BN + 1) + c];
    if (NT) {
      if (bias) v += __ldg(bias + gc);
      v = epi_forward(v, g.act, g.post, g.Ypre ? g.Ypre + yrow * g.ldy + gc : nullptr);
    } else if (g.Dact) {
      v *= act_grad_from_output(g.Dact[yrow * g.ldd + gc], gc < g.dsplit ? g.dact0 : g.dact1);
    }
    g.Y[yrow * g.ldy + gc] = v;
  }
}

// Tile-shape knobs.  Defaults are the LOW-REGISTER shapes: with six streams sharing the SMs the register file
// decides how many CTAs are resident, and resident warps are what hides the latency these kernels are bound by
// (measured on the bench workload: 128-wide GEMM tiles 220.7 ms, 64-wide 209.2 ms per update).
static const int g_wide = getenv("SARD_GEMM_WIDE") ? atoi(getenv("SARD_GEMM_WIDE")) : 0;
static const int g_wg_wide = getenv("SARD_WG_WIDE") ? atoi(getenv("SARD_WG_WIDE")) : 0;
static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int BN, int TN, bool NT, bool AGG, int MINB>
static int launch_fast_mb(const SardGemm* g, cudaStream_t st) {
  constexpr int tile_f = 2 * 16 * (128 + 4) + 2 * 16 * (BN + 4), c_f = 128 * (BN + 1);
  constexpr size_t smem = sizeof(float) * (tile_f > c_f ? tile_f : c_f);
  static bool configured[SARD_MAX_DEVICES] = {};
  if (sard_first_use_on_device(configured)) {
    SARD_CUDA(cudaFuncSetAttribute(gemm_fast_kernel<BN, TN, NT, AGG, MINB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  const int tiles = g->tiles ? g->max_tiles : ceil_div(g->M, 128);
  if (tiles <= 0) return 0;
  dim3 grid(tiles, g->N / BN);
  sard_launch(gemm_fast_kernel<BN, TN, NT, AGG, MINB, false>, grid, 256, smem, st, *g, (unsigned*)nullptr);
  SARD_LAUNCH_OK(NT ? "gemm_fast_kernel<fwd>" : "gemm_fast_kernel<bwd_data>");
  return 0;
}

template <bool NT>
__global__ void __launch_bounds__(256) splitk_finish_kernel(const SardGemm g, int S) {
  pdl_sync();
  const int N4 = g.N >> 2;
  const long long total = (long long)g.M * N4;
  const size_t slice = (size_t)g.M * g.N;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(idx / N4), c = (int)(idx - (long long)m * N4) * 4;
    const float* p = g.splitk_ws + (size_t)m * g.N + c;
    float4 v = *reinterpret_cast<const float4*>(p);
    for (int z = 1; z < S; ++z) {
      const float4 t = *reinterpret_cast<const float4*>(p + z * slice);
      v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
    }
    const long long yrow = g.y_rows ? g.y_rows[m] : m;
    float o[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int gc = c + i;
      float x = o[i];
      if (NT) {
        if (g.bias) x += __ldg(g.bias + gc);
        x = epi_forward(x, g.act, g.post, g.Ypre ? g.Ypre + yrow * g.ldy + gc : nullptr);
      } else if (g.Dact) {
        x *= act_grad_from_output(g.Dact[yrow * g.ldd + gc], gc < g.dsplit ? g.dact0 : g.dact1);
      }
      g.Y[yrow * g.ldy + gc] = x;
    }
  }
}

// split-K launch (ungrouped, no fused aggregate); S slices
template <bool NT>
static int launch_splitk(const SardGemm* g, int S, cudaStream_t st) {
  constexpr int BN = 64;
  constexpr int tile_f = 2 * 16 * (128 + 4) + 2 * 16 * (BN + 4), c_f = 128 * (BN + 1);
  constexpr size_t smem = sizeof(float) * (tile_f > c_f ? tile_f : c_f);
  static bool configured[SARD_MAX_DEVICES] = {};
  if (sard_first_use_on_device(configured)) {
    SARD_CUDA(cudaFuncSetAttribute(gemm_fast_kernel<BN, 4, NT, false, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  dim3 grid(ceil_div(g->M, 128), g->N / BN, S);
  // finishing a tile inside the GEMM launch (last-arriving slice) saves a launch but was measured slower on the update
  // (112.1 / 112.8 vs 110.0 / 110.2 ms, same box): the fence + arrival atomics and the serial tail cost more
  static const int fused = getenv("SARD_SPLITK_FUSED") ? atoi(getenv("SARD_SPLITK_FUSED")) : 0;
  if (fused && (long long)grid.x * grid.y <= SPLITK_MAX_TILES) {
    // the last-arriving slice of a tile finishes it in the same launch
    unsigned* base = nullptr;
    SARD_CUDA(cudaGetSymbolAddress(reinterpret_cast<void**>(&base), g_splitk_cnt));
    const unsigned slot = (unsigned)((reinterpret_cast<uintptr_t>(st) >> 4) * 2654435761u >> 28) & (SPLITK_SLOTS - 1);
    sard_launch(gemm_fast_kernel<BN, 4, NT, false, 3, true>, grid, 256, smem, st, *g, base + (size_t)slot * SPLITK_MAX_TILES);
    SARD_LAUNCH_OK("gemm_fast_kernel<splitk>");
    return 0;
  }
  sard_launch(gemm_fast_kernel<BN, 4, NT, false, 3, true>, grid, 256, smem, st, *g, (unsigned*)nullptr);
  SARD_LAUNCH_OK("gemm_fast_kernel<splitk>");
  sard_launch(splitk_finish_kernel<NT>, grid_for((long long)g->M * (g->N / 4), 256), 256, 0, st, *g, S);
  SARD_LAUNCH_OK("splitk_finish_kernel");
  return 0;
}
static const int g_splitk = getenv("SARD_GEMM_SPLITK") ? atoi(getenv("SARD_GEMM_SPLITK")) : 1;
// slices for a caller that provides a split-K workspace: largest power of two <= min(8, k-blocks / 4) keeping <= 296 CTAs
static int splitk_slices(const SardGemm* g) {
  if (!g_splitk || !g->splitk_ws || g->tiles || g->agg_cols > 0) return 1;
  const long long ctas = (long long)ceil_div(g->M, 128) * (g->N / 64);
  const int nkb = g->R / 16;
  int S = 1;
  while (S * 2 <= 8 && S * 2 <= nkb / 4 && ctas * S * 2 <= 296 && (long long)S * 2 * g->M * g->N <= g->splitk_floats) S *= 2;
  static const int s_min = getenv("SARD_GEMM_SPLITK_MIN") ? atoi(getenv("SARD_GEMM_SPLITK_MIN")) : 4;     // two slices do not pay for the finishing launch
  return S >= s_min ? S : 1;
}
static const int g_wg_minb = getenv("SARD_WG_MINB") ? atoi(getenv("SARD_WG_MINB")) : 3;     // 3 CTAs/SM (80 registers, 32 B of spill): 195.5 -> 191.2 ms
static const int g_minb = getenv("SARD_GEMM_MINB") ? atoi(getenv("SARD_GEMM_MINB")) : 0;
template <int BN, int TN, bool NT, bool AGG = false>
static int launch_fast(const SardGemm* g, cudaStream_t st) {
  // minimum resident CTAs per SM (caps registers): 64-wide tiles 3 (80 registers), fused-aggregate 2, 128-wide 1
  constexpr int DEF = BN == 64 ? (AGG ? 2 : 3) : 1;
  if (BN == 64 && g_minb == 4) return launch_fast_mb<BN, TN, NT, AGG, (BN == 64 ? 4 : 1)>(g, st);
  if (BN == 64 && AGG && g_minb == 3) return launch_fast_mb<BN, TN, NT, AGG, (BN == 64 ? 3 : 1)>(g, st);
  return launch_fast_mb<BN, TN, NT, AGG, DEF>(g, st);
}

// -1: the shape / alignment is not covered, use the generic kernel
extern "C" int sard_gemm_agg_supported(int h_in, int h_out) { return h_in > 0 && h_out > 0 && h_in % 64 == 0 && h_out % 64 == 0; }

// fused neighbour-sum GEMM (SardGemm.agg_cols > 0): only this kernel family implements it
static int sard_gemm_fast_agg(const SardGemm* g, bool nt, cudaStream_t st) {
  SARD_REQUIRE(!g->a_rows && !g->tiles && g->agg_ptr && g->agg_nbr && g->agg_node && g->agg_graph && g->agg_cum, "fused aggregate: lists");
  SARD_REQUIRE(g->R % 16 == 0 && g->agg_cols % 16 == 0 && g->agg_cols < g->R && g->N % 64 == 0, "fused aggregate: shape");
  SARD_REQUIRE(g->lda % 4 == 0 && al16(g->A) && g->ldb0 % 4 == 0 && al16(g->B0) && g->B1 && g->ldb1 % 4 == 0 && al16(g->B1),
               "fused aggregate: alignment");
  SARD_REQUIRE(!g->agg_out || (g->ld_agg_out % 4 == 0 && al16(g->agg_out)), "fused aggregate: agg_out alignment");
  if (nt) {
    SARD_REQUIRE(g->b_split % 16 == 0, "fused aggregate: b_split");
    return launch_fast<64, 4, true, true>(g, st);
  }
  SARD_REQUIRE(g->b_ksplit > 0 && g->b_ksplit % 16 == 0, "fused aggregate: b_ksplit");
  return launch_fast<64, 4, false, true>(g, st);
}

int sard_gemm_fast(const SardGemm* g, bool nt, cudaStream_t st) {
  if (g->agg_cols > 0) return sard_gemm_fast_agg(g, nt, st);
  if (g->R < 16 || g->R % 16 || g->N % 64 || g->lda % 4 || !al16(g->A) || g->ldb0 % 4 || !al16(g->B0)) return -1;
  if (g->b_group_stride % 4) return -1;
  const bool split = g->B1 != nullptr;
  if (split && (g->ldb1 % 4 || !al16(g->B1))) return -1;
  if (nt) {
    if (split && g->b_split % 16) return -1;
    if (!split && g->b_split < g->R) return -1;
    { const int S = splitk_slices(g); if (S > 1) return launch_splitk<true>(g, S, st); }
    if (g_wide && g->N % 128 == 0 && (long long)ceil_div(g->M, 128) * (g->N / 128) >= 120) return launch_fast<128, 8, true>(g, st);
    return launch_fast<64, 4, true>(g, st);
  }
  if (split) {
    if (g->b_split % 64) return -1;
    return launch_fast<64, 4, false>(g, st);               // every 64-column block lies in one weight matrix
  }
  if (g->b_split < g->N) return -1;
  { const int S = splitk_slices(g); if (S > 1) return launch_splitk<false>(g, S, st); }
  if (g_wide && g->N % 128 == 0 && (long long)ceil_div(g->M, 128) * (g->N / 128) >= 120) return launch_fast<128, 8, false>(g, st);
  return launch_fast<64, 4, false>(g, st);
}

// ------------------------------------------------------------------------------------------ weight gradient
// MASK: N2 < 128 and not a multiple of 4 allowed (in_fc: N2 = obs width): loads beyond r4(N2) are skipped, columns
// >= N2 are zeroed before they reach shared memory, and stores stop at N2 (column N2 of a partial row is the bias slot)
template <int BN1, int T2, int MINB, bool MASK>
__global__ void __launch_bounds__(BN1 / 8 * (128 / T2), MINB) wgrad_fast_kernel(const SardWGrad w) {
  pdl_sync();
  constexpr int BN2 = 128, BKM = 16, C2 = BN2 / T2, NTH = BN1 / 8 * C2;   // 8 x T2 register tile: (BN1/8) x (128/T2) threads
  constexpr int PS = BKM * (BN1 + 4), QS = BKM * (BN2 + 4);
  constexpr int P_LD = BKM * BN1 / 4 / NTH, Q_LD = BKM * BN2 / 4 / NTH;
  __shared__ __align__(16) float Ps[2 * PS];
  __shared__ __align__(16) float Qs[2 * QS];
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
  const int c10 = blockIdx.y * BN1, c20 = blockIdx.z * BN2;
  const int tid = threadIdx.x;
  const int t1 = tid / C2, t2 = tid % C2;
  float acc[8][T2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < T2; ++j) acc[i][j] = 0.f;
  float bsum = 0.f;
  const bool do_bias = blockIdx.z == 0 && tid < BN1;

  float4 rp[P_LD], rq[Q_LD];
  const int nsteps = (re - rs + BKM - 1) / BKM;
  int buf = 0;
#pragma unroll 1
  for (int s = 0; s <= nsteps; ++s) {
    if (s < nsteps) {
      const int m0 = rs + s * BKM;
#pragma unroll
      for (int i = 0; i < P_LD; ++i) {
        const int idx = tid + i * NTH;
        const int row = m0 + idx / (BN1 / 4), c = (idx % (BN1 / 4)) * 4;
        rp[i] = row < re ? __ldg(reinterpret_cast<const float4*>(w.P + (size_t)(w.p_rows ? w.p_rows[row] : row) * w.ldp + c10 + c))
                         : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int i = 0; i < Q_LD; ++i) {
        const int idx = tid + i * NTH;
        const int row = m0 + idx / (BN2 / 4), c = (idx % (BN2 / 4)) * 4;
        rq[i] = (row < re && (!MASK || c20 + c < w.N2))
                    ? __ldg(reinterpret_cast<const float4*>(w.Q + (size_t)(w.q_rows ? w.q_rows[row] : row) * w.ldq + c20 + c))
                    : make_float4(0.f, 0.f, 0.f, 0.f);
        if (MASK) {
          const int left = w.N2 - (c20 + c);               // valid components of this float4
          if (left < 4) { if (left < 2) rq[i].y = 0.f; if (left < 3) rq[i].z = 0.f; rq[i].w = 0.f; }
        }
      }
    }
    if (s > 0) {
      const float* Pc = Ps + (buf ^ 1) * PS;
      const float* Qc = Qs + (buf ^ 1) * QS;
#pragma unroll
      for (int k = 0; k < BKM; ++k) {
        float p[8], q[T2];
#pragma unroll
        for (int i = 0; i < 8; i += 4) {
          const float4 v = *reinterpret_cast<const float4*>(Pc + k * (BN1 + 4) + t1 * 8 + i);
          p[i] = v.x; p[i + 1] = v.y; p[i + 2] = v.z; p[i + 3] = v.w;
        }
#pragma unroll
        for (int j = 0; j < T2; j += 4) {
          const float4 v = *reinterpret_cast<const float4*>(Qc + k * (BN2 + 4) + t2 * T2 + j);
          q[j] = v.x; q[j + 1] = v.y; q[j + 2] = v.z; q[j + 3] = v.w;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < T2; ++j) acc[i][j] = fmaf(p[i], q[j], acc[i][j]);
      }
      if (do_bias) {
#pragma unroll
        for (int k = 0; k < BKM; ++k) bsum += Pc[k * (BN1 + 4) + tid];
      }
    }
    if (s < nsteps) {
#pragma unroll
      for (int i = 0; i < P_LD; ++i) {
        const int idx = tid + i * NTH;
        *reinterpret_cast<float4*>(Ps + buf * PS + (idx / (BN1 / 4)) * (BN1 + 4) + (idx % (BN1 / 4)) * 4) = rp[i];
      }
#pragma unroll
      for (int i = 0; i < Q_LD; ++i) {
        const int idx = tid + i * NTH;
        *reinterpret_cast<float4*>(Qs + buf * QS + (idx / (BN2 / 4)) * (BN2 + 4) + (idx % (BN2 / 4)) * 4) = rq[i];
      }
    }
    __syncthreads();
    buf ^= 1;
  }
  const int ldpart = SARD_WG_LDPART(w.N2);
  float* part = w.partials + (size_t)blockIdx.x * w.N1 * ldpart;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float* pp = part + (size_t)(c10 + t1 * 8 + i) * ldpart + c20 + t2 * T2;     // 16-byte aligned: ldpart % 4 == 0
#pragma unroll
    for (int j = 0; j < T2; j += 4) {
      const int left = MASK ? w.N2 - (c20 + t2 * T2 + j) : 4;
      if (left >= 4) {
        *reinterpret_cast<float4*>(pp + j) = make_float4(acc[i][j], acc[i][j + 1], acc[i][j + 2], acc[i][j + 3]);
      } else {
        if (left > 0) pp[j] = acc[i][j];
        if (left > 1) pp[j + 1] = acc[i][j + 1];
        if (left > 2) pp[j + 2] = acc[i][j + 2];
      }
    }
  }
  if (do_bias) part[(size_t)(c10 + tid) * ldpart + w.N2] = bsum;
}

// partial products for aligned shapes; -1 = not covered (caller uses the generic kernel)
int sard_wgrad_skinny(const SardWGrad* w, int nchunks, cudaStream_t st);
int sard_wgrad_fast(const SardWGrad* w, int nchunks, cudaStream_t st) {
  if (w->N1 <= 8) return sard_wgrad_skinny(w, nchunks, st);        // gemm_skinny.cu
  if (w->N1 % 64 || w->ldp % 4 || w->ldq % 4 || !al16(w->P) || !al16(w->Q)) return -1;
  if (w->N2 < 128 && w->N2 >= 16 && w->ldq >= ((w->N2 + 3) & ~3)) {      // in_fc: one masked 128-column tile
    dim3 grid(nchunks, w->N1 / 64, 1);
    sard_launch(wgrad_fast_kernel<64, 4, 2, true>, grid, 256, 0, st, *w);
    SARD_LAUNCH_OK("wgrad_fast_kernel<mask>");
    return 0;
  }
  if (w->N2 % 128) return -1;
  if (g_wg_wide == 2 && w->N1 % 128 == 0) {             // 128 x 128 tile, 8 x 8 per thread, 130 registers
    dim3 grid(nchunks, w->N1 / 128, w->N2 / 128);
    sard_launch(wgrad_fast_kernel<128, 8, 1, false>, grid, 256, 0, st, *w);
  } else if (g_wg_wide == 1) {                          // 64 x 128 tile, 8 x 8 per thread, 128 threads, 163 registers
    dim3 grid(nchunks, w->N1 / 64, w->N2 / 128);
    sard_launch(wgrad_fast_kernel<64, 8, 1, false>, grid, 128, 0, st, *w);
  } else {                                              // 64 x 128 tile, 8 x 4 per thread, 256 threads (default: most warps per SM)
    dim3 grid(nchunks, w->N1 / 64, w->N2 / 128);
    if (g_wg_minb == 3) sard_launch(wgrad_fast_kernel<64, 4, 3, false>, grid, 256, 0, st, *w);
    else sard_launch(wgrad_fast_kernel<64, 4, 2, false>, grid, 256, 0, st, *w);
  }
  SARD_LAUNCH_OK("wgrad_fast_kernel");
  return 0;
}
