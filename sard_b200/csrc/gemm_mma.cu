// Warp-level tensor-core variant of the aligned GEMMs: mma.sync m16n8k8 TF32 with the 3xTF32 split
// (x = hi + lo, x*y ~ hi*hi + hi*lo + lo*hi, fp32 accumulate), which keeps ~fp32 accuracy (needed for the
// rtol 1e-4 parity target) at ~4x fewer issued instructions than the FFMA kernels.  Unlike the tcgen05
// engines (umma.cu) a CTA here is as small as the FFMA one (256 threads, ~50 KB of shared memory, no tensor
// memory), so it co-runs with the other streams' kernels -- the regime this path lives in.
// Global->shared staging, tile table, split weights and the rolled epilogue are those of gemm_fast.cu.
#include "common.cuh"

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  lo = __float_as_uint(x - __uint_as_float(hi));            // the MMA truncates the low mantissa bits of lo
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// CTA tile 128 x BN, 8 warps as 4 (rows) x 2 (columns): warp tile 32 x BN/2 = 2 x (BN/16) MMA tiles
template <int BN, bool NT>
__global__ void __launch_bounds__(256) gemm_mma_kernel(const SardGemm g) {
  pdl_sync();
  constexpr int BM = 128, BK = 16;
  constexpr int LDA_S = BM + 8, LDB_S = BN + 8;            // +8: fragment loads (4 k x 8 m per warp) hit 32 distinct banks
  constexpr int AS = BK * LDA_S, BS = BK * LDB_S;
  constexpr int A_LD = BM * BK / 4 / 256;
  constexpr int B_LD = BN * BK / 4 / 256;
  constexpr int NTW = BN / 16;                              // n8 tiles per warp
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
#pragma unroll
  for (int i = 0; i < A_LD; ++i) {
    const int idx = tid + i * 256;
    const int grow = min(row_start + (idx >> 2), row_end - 1);
    a_ptr[i] = g.A + (size_t)(g.a_rows ? g.a_rows[grow] : grow) * g.lda + (idx & 3) * 4;
  }
  const int warp = tid >> 5, lane = tid & 31;
  const int wm = (warp >> 1) * 32, wn = (warp & 1) * (BN / 2);
  const int gq = lane >> 2, tq = lane & 3;
  float acc[2][NTW][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < NTW; ++j)
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[i][j][r] = 0.f;

  float4 ra[A_LD], rb[B_LD];
  const int nkb = g.R / BK;
  int buf = 0;
#pragma unroll 1
  for (int kb = 0; kb <= nkb; ++kb) {
    if (kb < nkb) {
      const int k0 = kb * BK;
#pragma unroll
      for (int i = 0; i < A_LD; ++i) ra[i] = __ldg(reinterpret_cast<const float4*>(a_ptr[i] + k0));
#pragma unroll
      for (int i = 0; i < B_LD; ++i) {
        const int idx = tid + i * 256;
        if (NT) {
          const int n = idx >> 2, kk = k0 + (idx & 3) * 4;
          rb[i] = kk < g.b_split ? __ldg(reinterpret_cast<const float4*>(B0 + (size_t)(n0 + n) * g.ldb0 + kk))
                                 : __ldg(reinterpret_cast<const float4*>(B1 + (size_t)(n0 + n) * g.ldb1 + (kk - g.b_split)));
        } else {
          const int r = idx / (BN / 4), c = (idx % (BN / 4)) * 4;
          rb[i] = n0 < g.b_split ? __ldg(reinterpret_cast<const float4*>(B0 + (size_t)(k0 + r) * g.ldb0 + n0 + c))
                                 : __ldg(reinterpret_cast<const float4*>(B1 + (size_t)(k0 + r) * g.ldb1 + (n0 - g.b_split) + c));
        }
      }
    }
    if (kb > 0) {
      const float* Ac = As + (buf ^ 1) * AS;
      const float* Bc = Bs + (buf ^ 1) * BS;
#pragma unroll
      for (int k8 = 0; k8 < BK; k8 += 8) {
        uint32_t ah[2][4], al[2][4];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const float* ap = Ac + (k8 + tq) * LDA_S + wm + i * 16 + gq;
          split_tf32(ap[0], ah[i][0], al[i][0]);
          split_tf32(ap[8], ah[i][1], al[i][1]);
          split_tf32(ap[4 * LDA_S], ah[i][2], al[i][2]);
          split_tf32(ap[4 * LDA_S + 8], ah[i][3], al[i][3]);
        }
        // 4 column tiles at a time: 8 independent accumulators between two MMAs on the same one (HMMA latency)
#pragma unroll
        for (int j0 = 0; j0 < NTW; j0 += 4) {
          uint32_t bh[4][2], bl[4][2];
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const float* bp = Bc + (k8 + tq) * LDB_S + wn + (j0 + jj) * 8 + gq;
            split_tf32(bp[0], bh[jj][0], bl[jj][0]);
            split_tf32(bp[4 * LDB_S], bh[jj][1], bl[jj][1]);
          }
#pragma unroll
          for (int jj = 0; jj < 4; ++jj)
#pragma unroll
            for (int i = 0; i < 2; ++i) mma_tf32(acc[i][j0 + jj], al[i], bh[jj]);      // small terms first
#pragma unroll
          for (int jj = 0; jj < 4; ++jj)
#pragma unroll
            for (int i = 0; i < 2; ++i) mma_tf32(acc[i][j0 + jj], ah[i], bl[jj]);
#pragma unroll
          for (int jj = 0; jj < 4; ++jj)
#pragma unroll
            for (int i = 0; i < 2; ++i) mma_tf32(acc[i][j0 + jj], ah[i], bh[jj]);
        }
      }
    }
    if (kb < nkb) {
#pragma unroll
      for (int i = 0; i < A_LD; ++i) {
        const int idx = tid + i * 256;
        float* d = As + buf * AS + ((idx & 3) * 4) * LDA_S + (idx >> 2);
        d[0] = ra[i].x; d[LDA_S] = ra[i].y; d[2 * LDA_S] = ra[i].z; d[3 * LDA_S] = ra[i].w;
      }
#pragma unroll
      for (int i = 0; i < B_LD; ++i) {
        const int idx = tid + i * 256;
        if (NT) {
          float* d = Bs + buf * BS + ((idx & 3) * 4) * LDB_S + (idx >> 2);
          d[0] = rb[i].x; d[LDB_S] = rb[i].y; d[2 * LDB_S] = rb[i].z; d[3 * LDB_S] = rb[i].w;
        } else {
          *reinterpret_cast<float4*>(Bs + buf * BS + (idx / (BN / 4)) * LDB_S + (idx % (BN / 4)) * 4) = rb[i];
        }
      }
    }
    __syncthreads();
    buf ^= 1;
  }

  float* Cs = smem;                                          // [BM][BN+1]
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < NTW; ++j) {
      const int m = wm + i * 16 + gq, c = wn + j * 8 + 2 * tq;
      Cs[m * (BN + 1) + c] = acc[i][j][0];
      Cs[m * (BN + 1) + c + 1] = acc[i][j][1];
      Cs[(m + 8) * (BN + 1) + c] = acc[i][j][2];
      Cs[(m + 8) * (BN + 1) + c + 1] = acc[i][j][3];
    }
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

template <int BN, bool NT>
static int launch_mma(const SardGemm* g, cudaStream_t st) {
  constexpr int tile_f = 2 * 16 * (128 + 8) + 2 * 16 * (BN + 8), c_f = 128 * (BN + 1);
  constexpr size_t smem = sizeof(float) * (tile_f > c_f ? tile_f : c_f);
  static bool configured[SARD_MAX_DEVICES] = {};
  if (sard_first_use_on_device(configured)) {
    SARD_CUDA(cudaFuncSetAttribute(gemm_mma_kernel<BN, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  const int tiles = g->tiles ? g->max_tiles : ceil_div(g->M, 128);
  if (tiles <= 0) return 0;
  dim3 grid(tiles, g->N / BN);
  sard_launch(gemm_mma_kernel<BN, NT>, grid, 256, smem, st, *g);
  SARD_LAUNCH_OK(NT ? "gemm_mma_kernel<fwd>" : "gemm_mma_kernel<bwd_data>");
  return 0;
}

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// -1: the shape / alignment is not covered (same conditions as sard_gemm_fast)
int sard_gemm_mma(const SardGemm* g, bool nt, cudaStream_t st) {
  if (g->R < 16 || g->R % 16 || g->N % 64 || g->lda % 4 || !al16(g->A) || g->ldb0 % 4 || !al16(g->B0)) return -1;
  if (g->b_group_stride % 4) return -1;
  const bool split = g->B1 != nullptr;
  if (split && (g->ldb1 % 4 || !al16(g->B1))) return -1;
  if (nt) {
    if (split && g->b_split % 16) return -1;
    if (!split && g->b_split < g->R) return -1;
    if (g->N % 128 == 0 && (long long)ceil_div(g->M, 128) * (g->N / 128) >= 120) return launch_mma<128, true>(g, st);
    return launch_mma<64, true>(g, st);
  }
  if (split) {
    if (g->b_split % 64) return -1;
    return launch_mma<64, false>(g, st);
  }
  if (g->b_split < g->N) return -1;
  if (g->N % 128 == 0 && (long long)ceil_div(g->M, 128) * (g->N / 128) >= 120) return launch_mma<128, false>(g, st);
  return launch_mma<64, false>(g, st);
}
