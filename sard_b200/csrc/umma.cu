// tcgen05 (5th-gen tensor core) engine for the linear layers: error-compensated 3xTF32.
//
//   D[128 x Nt] (fp32, TMEM) += A_hi B_hi^T + A_lo B_hi^T + A_hi B_lo^T        per 8-wide k-step
//
// A and B tiles are read from global memory by all threads through the same operand accessors as the
// SIMT engine (row gather, [lin_l|lin_r] split weights, per-body-index weight slabs, transposed
// weights for backward-data), split into hi = top 19 bits and lo = v - hi, and written to shared
// memory in the canonical K-major SWIZZLE_128B layout that the UMMA shared-memory descriptors
// describe (8-row x 128-byte atoms, 16-byte chunk index XOR row&7).  One elected thread issues
// tcgen05.mma.kind::tf32 (M=128, N=Nt, K=8); completion is tracked with tcgen05.commit on
// mbarriers, two smem stages overlap the fill of chunk k+1 with the MMAs of chunk k; the epilogue
// reads the accumulator with tcgen05.ld (32 lanes x 32 columns per warp) and applies the same
// bias / activation / sin / act' / row-scatter epilogue as the SIMT engine.
// hi*hi + lo*hi + hi*lo keeps ~2^-21 relative error, inside the rtol 1e-4 parity target that a
// single TF32 pass (2^-11) would miss.
#include "common.cuh"

#define UMMA_BM 128
#define UMMA_KC 32            // fp32 elements per 128-byte swizzle row
#define UMMA_THREADS 256

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// K-major SWIZZLE_128B descriptor: start address, LBO = 1 (unused for swizzled K-major), SBO = 1024 B
// between 8-row groups, version 1 (sm_100), layout type 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// store one 16-byte chunk (4 consecutive k) of tile row `row` as hi and lo parts
__device__ __forceinline__ void store_split(float* hi_tile, float* lo_tile, int row, int chunk, float4 v) {
  const int off = row * UMMA_KC + ((chunk ^ (row & 7)) << 2);
  float4 h, l;
  h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); l.x = v.x - h.x;
  h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); l.y = v.y - h.y;
  h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); l.z = v.z - h.z;
  h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); l.w = v.w - h.w;
  *reinterpret_cast<float4*>(hi_tile + off) = h;
  *reinterpret_cast<float4*>(lo_tile + off) = l;
}

template <int NT_TILE, bool NT>
__global__ void __launch_bounds__(UMMA_THREADS, 1) umma_gemm_kernel(const SardGemm g) {
  pdl_sync();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int A_TILE = UMMA_BM * UMMA_KC;                // floats
  constexpr int B_TILE = NT_TILE * UMMA_KC;
  constexpr int STAGE = 2 * A_TILE + 2 * B_TILE;           // A_hi, A_lo, B_hi, B_lo
  // 1024-byte aligned base (dynamic smem is only guaranteed 16-byte aligned)
  float* base = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar_stage[2];
  __shared__ uint64_t bar_done;
  __shared__ uint32_t tmem_slot;

  int row_start, row_end, group = 0;
  if (g.tiles) {
    const int nt = *g.num_tiles;
    if ((int)blockIdx.x >= nt) return;
    const int4 t = reinterpret_cast<const int4*>(g.tiles)[blockIdx.x];
    row_start = t.x; row_end = t.y; group = t.z;
  } else {
    row_start = blockIdx.x * UMMA_BM;
    row_end = min(g.M, row_start + UMMA_BM);
    if (row_start >= g.M) return;
  }
  const int n0 = blockIdx.y * NT_TILE;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* B0 = g.B0 + (size_t)group * g.b_group_stride;
  const float* B1 = g.B1 ? g.B1 + (size_t)group * g.b_group_stride : nullptr;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"((uint32_t)NT_TILE) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    mbar_init(&bar_stage[0], 1);
    mbar_init(&bar_stage[1], 1);
    mbar_init(&bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_slot;

  // A rows handled by this thread: chunk-fastest mapping (8 threads cover one 128-byte row)
  constexpr int A_CH = UMMA_BM * 8 / UMMA_THREADS;          // 4 chunks per thread
  long long a_off[A_CH];
#pragma unroll
  for (int i = 0; i < A_CH; ++i) {
    const int m = (tid + i * UMMA_THREADS) >> 3;
    const int grow = row_start + m;
    a_off[i] = grow < row_end ? (long long)(g.a_rows ? g.a_rows[grow] : grow) * g.lda : -1;
  }
  const bool a_vec = ((g.lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.A) & 15) == 0);
  const bool b_vec = NT && ((g.ldb0 & 3) == 0) && ((reinterpret_cast<uintptr_t>(B0) & 15) == 0) && ((g.b_split & 3) == 0) &&
                     (!B1 || (((g.ldb1 & 3) == 0) && ((reinterpret_cast<uintptr_t>(B1) & 15) == 0)));
  // instruction descriptor: D fp32, A/B tf32, both K-major, N, M
  constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NT_TILE >> 3) << 17) | ((uint32_t)(UMMA_BM >> 4) << 24);

  const int nkc = (g.R + UMMA_KC - 1) / UMMA_KC;
  for (int kc = 0; kc < nkc; ++kc) {
    const int s = kc & 1;
    float* a_hi = base + s * STAGE;
    float* a_lo = a_hi + A_TILE;
    float* b_hi = a_lo + A_TILE;
    float* b_lo = b_hi + B_TILE;
    if (kc >= 2) mbar_wait(&bar_stage[s], ((kc >> 1) - 1) & 1);   // MMAs of chunk kc-2 have drained this stage
    const int k0 = kc * UMMA_KC;
    // ---- A tile ----
#pragma unroll
    for (int i = 0; i < A_CH; ++i) {
      const int idx = tid + i * UMMA_THREADS;
      const int m = idx >> 3, c = idx & 7;
      const int kk = k0 + c * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a_off[i] >= 0 && kk < g.R) {
        const float* p = g.A + a_off[i] + kk;
        if (a_vec && kk + 3 < g.R) v = __ldg(reinterpret_cast<const float4*>(p));
        else {
          v.x = __ldg(p);
          if (kk + 1 < g.R) v.y = __ldg(p + 1);
          if (kk + 2 < g.R) v.z = __ldg(p + 2);
          if (kk + 3 < g.R) v.w = __ldg(p + 3);
        }
      }
      store_split(a_hi, a_lo, m, c, v);
    }
    // ---- B tile: B(n, k) ----
    constexpr int B_CH = NT_TILE * 8 / UMMA_THREADS;         // chunks per thread
#pragma unroll
    for (int i = 0; i < B_CH; ++i) {
      const int idx = tid + i * UMMA_THREADS;
      int n, c;
      if (NT) { n = idx >> 3; c = idx & 7; }                 // source contiguous along k
      else    { n = idx % NT_TILE; c = idx / NT_TILE; }      // source contiguous along n
      const int gc = n0 + n, kk = k0 + c * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gc < g.N && kk < g.R) {
        if (NT) {
          if (b_vec && kk + 3 < g.R) {
            v = kk < g.b_split ? __ldg(reinterpret_cast<const float4*>(B0 + (size_t)gc * g.ldb0 + kk))
                               : __ldg(reinterpret_cast<const float4*>(B1 + (size_t)gc * g.ldb1 + (kk - g.b_split)));
          } else {
            float e[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int k = kk + j;
              if (k < g.R) e[j] = k < g.b_split ? __ldg(B0 + (size_t)gc * g.ldb0 + k) : __ldg(B1 + (size_t)gc * g.ldb1 + (k - g.b_split));
            }
            v = make_float4(e[0], e[1], e[2], e[3]);
          }
        } else {
          float e[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int k = kk + j;
            if (k < g.R) e[j] = gc < g.b_split ? __ldg(B0 + (size_t)k * g.ldb0 + gc) : __ldg(B1 + (size_t)k * g.ldb1 + (gc - g.b_split));
          }
          v = make_float4(e[0], e[1], e[2], e[3]);
        }
      }
      store_split(b_hi, b_lo, n, c, v);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> async proxy (UMMA)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t sa_hi = smem_u32(a_hi), sa_lo = smem_u32(a_lo), sb_hi = smem_u32(b_hi), sb_lo = smem_u32(b_lo);
      const int ksteps = min(4, (g.R - k0 + 7) >> 3);
      for (int k = 0; k < ksteps; ++k) {
        const uint32_t adv = k * 32;                               // 8 tf32 = 32 bytes inside the swizzle row
        const uint64_t da_hi = umma_desc(sa_hi + adv), da_lo = umma_desc(sa_lo + adv);
        const uint64_t db_hi = umma_desc(sb_hi + adv), db_lo = umma_desc(sb_lo + adv);
        umma_tf32(tmem_d, da_hi, db_hi, idesc, (kc | k) ? 1u : 0u);
        umma_tf32(tmem_d, da_lo, db_hi, idesc, 1u);
        umma_tf32(tmem_d, da_hi, db_lo, idesc, 1u);
      }
      umma_commit(&bar_stage[s]);
      if (kc == nkc - 1) umma_commit(&bar_done);
    }
  }
  mbar_wait(&bar_done, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // ---- epilogue: warp w reads TMEM lanes 32*(w%4).., warps w and w+4 split the columns ----
  const int lane_grp = warp & 3;
  const int grow = row_start + lane_grp * 32 + lane;
  const bool row_ok = grow < row_end;
  const long long yrow = row_ok ? (g.y_rows ? g.y_rows[grow] : grow) : 0;
  const float* bias = g.bias ? g.bias + (size_t)group * g.bias_group_stride : nullptr;
  constexpr int NCHUNK = NT_TILE / 32;
  for (int cc = (warp >> 2); cc < NCHUNK; cc += 2) {
    float v[32];
    tmem_ld32(tmem_d + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(cc * 32), v);
    if (row_ok) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int gc = n0 + cc * 32 + j;
        if (gc >= g.N) break;
        float x = v[j];
        if (NT) {
          if (bias) x += __ldg(bias + gc);
          x = epi_forward(x, g.act, g.post, g.Ypre ? g.Ypre + yrow * g.ldy + gc : nullptr);
        } else if (g.Dact) {
          x *= act_grad_from_output(g.Dact[yrow * g.ldd + gc], gc < g.dsplit ? g.dact0 : g.dact1);
        }
        g.Y[yrow * g.ldy + gc] = x;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)NT_TILE) : "memory");
  }
}

template <int NT_TILE, bool NT>
static int launch_umma_t(const SardGemm* g, cudaStream_t st) {
  constexpr size_t smem = 2 * (2 * UMMA_BM * UMMA_KC + 2 * NT_TILE * UMMA_KC) * sizeof(float) + 1024;
  static bool configured[SARD_MAX_DEVICES] = {};
  if (sard_first_use_on_device(configured)) {
    SARD_CUDA(cudaFuncSetAttribute(umma_gemm_kernel<NT_TILE, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  const int tiles = g->tiles ? g->max_tiles : ceil_div(g->M, UMMA_BM);
  if (tiles <= 0) return 0;
  dim3 grid(tiles, ceil_div(g->N, NT_TILE));
  sard_launch(umma_gemm_kernel<NT_TILE, NT>, grid, UMMA_THREADS, smem, st, *g);
  SARD_LAUNCH_OK(NT ? "umma_gemm_kernel<fwd>" : "umma_gemm_kernel<bwd_data>");
  return 0;
}


// ------------------------------------------------------------------------------------------
// v2: A operand in tensor memory.  In SS mode every tf32 MMA re-reads A (4 KB) and B (4 KB) from shared
// memory, which pins the tensor pipe at the 128 B/clk smem limit (measured: 4x the 64-cycle floor).
// Here each thread owns one row of the A tile: it loads that row's 16 k-values straight from global
// memory, splits them and writes hi/lo into TMEM columns with tcgen05.st; only B goes through shared
// memory.  Global loads of chunk k+1 are issued before the barrier of chunk k (register prefetch).
// TMEM map (256 columns): [0,128) accumulator, [128,256) two stages of {A_hi[32], A_lo[32]}.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
}
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}

template <int NT_TILE, bool NT>
__global__ void __launch_bounds__(UMMA_THREADS, 1) umma_gemm_ts_kernel(const SardGemm g) {
  pdl_sync();
  // one pipeline step covers KSTEP = 64 k-values: two 128-byte swizzle sub-tiles of B in smem, 64 TMEM
  // columns each for A_hi and A_lo.  Each thread then reads exactly one 128-byte line of its A row per step.
  constexpr int KSUB = 2, KSTEP = KSUB * UMMA_KC;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int B_TILE = NT_TILE * UMMA_KC;                // floats per sub-tile
  constexpr int STAGE = 2 * KSUB * B_TILE;                 // B_hi[KSUB], B_lo[KSUB]
  constexpr uint32_t TMEM_COLS = 512;                      // [0,128) D, then 2 stages x {A_hi[64], A_lo[64]}
  float* base = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar_stage[2];
  __shared__ uint64_t bar_done;
  __shared__ uint32_t tmem_slot;

  int row_start, row_end, group = 0;
  if (g.tiles) {
    const int nt = *g.num_tiles;
    if ((int)blockIdx.x >= nt) return;
    const int4 t = reinterpret_cast<const int4*>(g.tiles)[blockIdx.x];
    row_start = t.x; row_end = t.y; group = t.z;
  } else {
    row_start = blockIdx.x * UMMA_BM;
    row_end = min(g.M, row_start + UMMA_BM);
    if (row_start >= g.M) return;
  }
  const int n0 = blockIdx.y * NT_TILE;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int lane_grp = warp & 3, half = warp >> 2;
  const float* B0 = g.B0 + (size_t)group * g.b_group_stride;
  const float* B1 = g.B1 ? g.B1 + (size_t)group * g.b_group_stride : nullptr;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    mbar_init(&bar_stage[0], 1);
    mbar_init(&bar_stage[1], 1);
    mbar_init(&bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_slot;
  const uint32_t tmem_lane = (uint32_t)(lane_grp * 32) << 16;

  // this thread's A row (= TMEM lane) and its 32-wide k slice inside each 64-wide step
  const int my_row = row_start + lane_grp * 32 + lane;
  const long long a_row_off = my_row < row_end ? (long long)(g.a_rows ? g.a_rows[my_row] : my_row) * g.lda : -1;
  const bool a_vec = ((g.lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.A) & 15) == 0);
  const bool b_vec = NT && ((g.ldb0 & 3) == 0) && ((reinterpret_cast<uintptr_t>(B0) & 15) == 0) && ((g.b_split & 3) == 0) &&
                     (!B1 || (((g.ldb1 & 3) == 0) && ((reinterpret_cast<uintptr_t>(B1) & 15) == 0)));
  constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NT_TILE >> 3) << 17) | ((uint32_t)(UMMA_BM >> 4) << 24);
  constexpr int B_CH = NT_TILE * (KSTEP / 4) / UMMA_THREADS;     // 16-byte chunks per thread and step

  float ra[32];
  float4 rb[B_CH];
  auto b_coord = [&](int idx, int& n, int& c) {
    if (NT) { n = idx / (KSTEP / 4); c = idx % (KSTEP / 4); }      // source contiguous along k
    else    { n = idx % NT_TILE; c = idx / NT_TILE; }              // source contiguous along n
  };
  auto load_regs = [&](int step) {
    const int k0 = step * KSTEP;
    const int ka = k0 + half * 32;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int kk = ka + q * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a_row_off >= 0 && kk < g.R) {
        const float* p = g.A + a_row_off + kk;
        if (a_vec && kk + 3 < g.R) v = __ldg(reinterpret_cast<const float4*>(p));
        else {
          v.x = __ldg(p);
          if (kk + 1 < g.R) v.y = __ldg(p + 1);
          if (kk + 2 < g.R) v.z = __ldg(p + 2);
          if (kk + 3 < g.R) v.w = __ldg(p + 3);
        }
      }
      ra[q * 4 + 0] = v.x; ra[q * 4 + 1] = v.y; ra[q * 4 + 2] = v.z; ra[q * 4 + 3] = v.w;
    }
#pragma unroll
    for (int i = 0; i < B_CH; ++i) {
      int n, c;
      b_coord(tid + i * UMMA_THREADS, n, c);
      const int gc = n0 + n, kk = k0 + c * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gc < g.N && kk < g.R) {
        if (NT && b_vec && kk + 3 < g.R) {
          v = kk < g.b_split ? __ldg(reinterpret_cast<const float4*>(B0 + (size_t)gc * g.ldb0 + kk))
                             : __ldg(reinterpret_cast<const float4*>(B1 + (size_t)gc * g.ldb1 + (kk - g.b_split)));
        } else {
          float e[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int k = kk + j;
            if (k < g.R) {
              if (NT) e[j] = k < g.b_split ? __ldg(B0 + (size_t)gc * g.ldb0 + k) : __ldg(B1 + (size_t)gc * g.ldb1 + (k - g.b_split));
              else    e[j] = gc < g.b_split ? __ldg(B0 + (size_t)k * g.ldb0 + gc) : __ldg(B1 + (size_t)k * g.ldb1 + (gc - g.b_split));
            }
          }
          v = make_float4(e[0], e[1], e[2], e[3]);
        }
      }
      rb[i] = v;
    }
  };

  const int nsteps = (g.R + KSTEP - 1) / KSTEP;
  load_regs(0);
  for (int st = 0; st < nsteps; ++st) {
    const int s = st & 1;
    float* b_hi = base + s * STAGE;                        // sub-tile j at b_hi + j*B_TILE
    float* b_lo = b_hi + KSUB * B_TILE;
    const uint32_t ta_hi = tmem_d + 128u + (uint32_t)s * 128u;      // A_hi columns [0,64), A_lo at +64
    if (st >= 2) {
      mbar_wait(&bar_stage[s], ((st >> 1) - 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
#pragma unroll
    for (int h16 = 0; h16 < 2; ++h16) {
      float hi[16], lo[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        hi[j] = __uint_as_float(__float_as_uint(ra[h16 * 16 + j]) & 0xffffe000u);
        lo[j] = ra[h16 * 16 + j] - hi[j];
      }
      tmem_st16(ta_hi + tmem_lane + (uint32_t)(half * 32 + h16 * 16), hi);
      tmem_st16(ta_hi + 64u + tmem_lane + (uint32_t)(half * 32 + h16 * 16), lo);
    }
#pragma unroll
    for (int i = 0; i < B_CH; ++i) {
      int n, c;
      b_coord(tid + i * UMMA_THREADS, n, c);
      const int sub = c >> 3;
      store_split(b_hi + sub * B_TILE, b_lo + sub * B_TILE, n, c & 7, rb[i]);
    }
    if (st + 1 < nsteps) load_regs(st + 1);                 // in flight across the barrier and the MMA issue
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t sb_hi = smem_u32(b_hi), sb_lo = smem_u32(b_lo);
      const int k0 = st * KSTEP;
      const int ksteps = min(KSTEP / 8, (g.R - k0 + 7) >> 3);
      for (int k = 0; k < ksteps; ++k) {
        const uint32_t boff = (uint32_t)((k >> 2) * B_TILE * 4 + (k & 3) * 32);
        const uint64_t db_hi = umma_desc(sb_hi + boff), db_lo = umma_desc(sb_lo + boff);
        const uint32_t a_hi = ta_hi + (uint32_t)(k * 8), a_lo = ta_hi + 64u + (uint32_t)(k * 8);
        umma_tf32_ts(tmem_d, a_hi, db_hi, idesc, (st | k) ? 1u : 0u);
        umma_tf32_ts(tmem_d, a_lo, db_hi, idesc, 1u);
        umma_tf32_ts(tmem_d, a_hi, db_lo, idesc, 1u);
      }
      umma_commit(&bar_stage[s]);
      if (st == nsteps - 1) umma_commit(&bar_done);
    }
  }
  mbar_wait(&bar_done, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  const int grow = my_row;
  const bool row_ok = grow < row_end;
  const long long yrow = row_ok ? (g.y_rows ? g.y_rows[grow] : grow) : 0;
  const float* bias = g.bias ? g.bias + (size_t)group * g.bias_group_stride : nullptr;
  const bool y_vec = ((g.ldy & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.Y) & 15) == 0) && ((n0 & 3) == 0);
  constexpr int NCHUNK = NT_TILE / 32;
  for (int cc = half; cc < NCHUNK; cc += 2) {
    float v[32];
    tmem_ld32(tmem_d + tmem_lane + (uint32_t)(cc * 32), v);
    if (!row_ok) continue;
    const int c0 = n0 + cc * 32;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int gc = c0 + j;
      float x = v[j];
      if (gc < g.N) {
        if (NT) {
          if (bias) x += __ldg(bias + gc);
          x = epi_forward(x, g.act, g.post, g.Ypre ? g.Ypre + yrow * g.ldy + gc : nullptr);
        } else if (g.Dact) {
          x *= act_grad_from_output(g.Dact[yrow * g.ldd + gc], gc < g.dsplit ? g.dact0 : g.dact1);
        }
      }
      v[j] = x;
    }
    float* yp = g.Y + yrow * g.ldy + c0;
    if (y_vec && c0 + 32 <= g.N) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(yp + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) if (c0 + j < g.N) yp[j] = v[j];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(TMEM_COLS) : "memory");
  }
}

template <int NT_TILE, bool NT>
static int launch_umma_ts(const SardGemm* g, cudaStream_t st) {
  constexpr size_t smem = 2 * (2 * 2 * NT_TILE * UMMA_KC) * sizeof(float) + 1024;
  static bool configured[SARD_MAX_DEVICES] = {};
  if (sard_first_use_on_device(configured)) {
    SARD_CUDA(cudaFuncSetAttribute(umma_gemm_ts_kernel<NT_TILE, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  const int tiles = g->tiles ? g->max_tiles : ceil_div(g->M, UMMA_BM);
  if (tiles <= 0) return 0;
  dim3 grid(tiles, ceil_div(g->N, NT_TILE));
  sard_launch(umma_gemm_ts_kernel<NT_TILE, NT>, grid, UMMA_THREADS, smem, st, *g);
  SARD_LAUNCH_OK(NT ? "umma_gemm_ts_kernel<fwd>" : "umma_gemm_ts_kernel<bwd_data>");
  return 0;
}

extern int g_sard_gemm_engine;


// ------------------------------------------------------------------------------------------
// Weight gradient on the tensor cores:  partial[chunk][n1][n2] = sum_{rows of chunk} dZ[row][n1] * X[row][n2].
// The reduction runs over rows, so the UMMA "M" dimension is n1 (<= 128 per tile) and "K" is the row index:
//   A(m = n1, k = row) = dZ[row][n1]   -> every thread owns one n1 (TMEM lane), loads 16 rows per 32-row step
//                                          (coalesced across the warp along n1) and stores hi/lo with tcgen05.st
//   B(n = n2, k = row) = X[row][n2]    -> consecutive threads take consecutive n2 (coalesced), 4 rows per 16-byte
//                                          chunk, hi/lo written K-major SWIZZLE_128B (conflict-free: the XOR
//                                          swizzle spreads the 8 rows of an atom over all banks)
// TMEM map: [0,256) accumulator (N2 tile up to 256), [256,512) two stages of {A_hi[32], A_lo[32]} (+pad).
// The bias gradient (column sums of dZ) is accumulated in registers from the same A values.
// ------------------------------------------------------------------------------------------
template <int NB>
__global__ void __launch_bounds__(UMMA_THREADS, 1) umma_wgrad_kernel(const SardWGrad w) {
  pdl_sync();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int B_TILE = NB * UMMA_KC;                     // floats
  constexpr int STAGE = 2 * B_TILE;
  constexpr uint32_t TMEM_COLS = 512;
  float* base = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar_stage[2];
  __shared__ uint64_t bar_done;
  __shared__ uint32_t tmem_slot;
  __shared__ float bsum_s[UMMA_BM];

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
  const int c10 = blockIdx.y * UMMA_BM, c20 = blockIdx.z * NB;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int lane_grp = warp & 3, half = warp >> 2;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    mbar_init(&bar_stage[0], 1);
    mbar_init(&bar_stage[1], 1);
    mbar_init(&bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_slot;
  const uint32_t tmem_lane = (uint32_t)(lane_grp * 32) << 16;

  const int n1 = c10 + lane_grp * 32 + lane;               // this thread's output row (TMEM lane)
  const bool n1_ok = n1 < w.N1;
  constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(UMMA_BM >> 4) << 24);
  constexpr int B_CH = NB * 8 / UMMA_THREADS;

  float ra[16];
  float4 rb[B_CH];
  float bsum = 0.f;
  auto load_regs = [&](int step) {
    const int r0 = rs + step * UMMA_KC;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int r = r0 + half * 16 + i;
      float v = 0.f;
      if (n1_ok && r < re) v = __ldg(w.P + (size_t)(w.p_rows ? w.p_rows[r] : r) * w.ldp + n1);
      ra[i] = v;
    }
#pragma unroll
    for (int i = 0; i < B_CH; ++i) {
      const int idx = tid + i * UMMA_THREADS;
      const int n = idx % NB, c = idx / NB;
      const int n2 = c20 + n;
      float e[4] = {0.f, 0.f, 0.f, 0.f};
      if (n2 < w.N2) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = r0 + c * 4 + j;
          if (r < re) e[j] = __ldg(w.Q + (size_t)(w.q_rows ? w.q_rows[r] : r) * w.ldq + n2);
        }
      }
      rb[i] = make_float4(e[0], e[1], e[2], e[3]);
    }
  };

  const int nsteps = (re - rs + UMMA_KC - 1) / UMMA_KC;
  load_regs(0);
  for (int st = 0; st < nsteps; ++st) {
    const int s = st & 1;
    float* b_hi = base + s * STAGE;
    float* b_lo = b_hi + B_TILE;
    const uint32_t ta_hi = tmem_d + 256u + (uint32_t)s * 128u;     // A_hi columns [0,32), A_lo at +64
    if (st >= 2) {
      mbar_wait(&bar_stage[s], ((st >> 1) - 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    {
      float hi[16], lo[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        bsum += ra[j];
        hi[j] = __uint_as_float(__float_as_uint(ra[j]) & 0xffffe000u);
        lo[j] = ra[j] - hi[j];
      }
      tmem_st16(ta_hi + tmem_lane + (uint32_t)(half * 16), hi);
      tmem_st16(ta_hi + 64u + tmem_lane + (uint32_t)(half * 16), lo);
    }
#pragma unroll
    for (int i = 0; i < B_CH; ++i) {
      const int idx = tid + i * UMMA_THREADS;
      store_split(b_hi, b_lo, idx % NB, idx / NB, rb[i]);
    }
    if (st + 1 < nsteps) load_regs(st + 1);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t sb_hi = smem_u32(b_hi), sb_lo = smem_u32(b_lo);
      const int ksteps = min(4, (re - rs - st * UMMA_KC + 7) >> 3);
      for (int k = 0; k < ksteps; ++k) {
        const uint64_t db_hi = umma_desc(sb_hi + k * 32), db_lo = umma_desc(sb_lo + k * 32);
        const uint32_t a_hi = ta_hi + (uint32_t)(k * 8), a_lo = ta_hi + 64u + (uint32_t)(k * 8);
        umma_tf32_ts(tmem_d, a_hi, db_hi, idesc, (st | k) ? 1u : 0u);
        umma_tf32_ts(tmem_d, a_lo, db_hi, idesc, 1u);
        umma_tf32_ts(tmem_d, a_hi, db_lo, idesc, 1u);
      }
      umma_commit(&bar_stage[s]);
      if (st == nsteps - 1) umma_commit(&bar_done);
    }
  }
  // bias gradient: the two halves of a lane group hold the two 16-row halves of every step
  if (half == 1) bsum_s[lane_grp * 32 + lane] = bsum;
  mbar_wait(&bar_done, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  __syncthreads();

  const int ldpart = SARD_WG_LDPART(w.N2);
  float* part = w.partials + (size_t)blockIdx.x * w.N1 * ldpart;
  constexpr int NCHUNK = NB / 32;
  for (int cc = half; cc < NCHUNK; cc += 2) {
    float v[32];
    tmem_ld32(tmem_d + tmem_lane + (uint32_t)(cc * 32), v);
    if (!n1_ok) continue;
    float* pp = part + (size_t)n1 * ldpart + c20 + cc * 32;
#pragma unroll
    for (int j = 0; j < 32; ++j) if (c20 + cc * 32 + j < w.N2) pp[j] = v[j];
  }
  if (half == 0 && n1_ok && blockIdx.z == 0) part[(size_t)n1 * ldpart + w.N2] = bsum + bsum_s[lane_grp * 32 + lane];
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(TMEM_COLS) : "memory");
  }
}

template <int NB>
static int launch_umma_wgrad(const SardWGrad* w, int nchunks, cudaStream_t st) {
  constexpr size_t smem = 2 * (2 * NB * UMMA_KC) * sizeof(float) + 1024;
  static bool configured[SARD_MAX_DEVICES] = {};
  if (sard_first_use_on_device(configured)) {
    SARD_CUDA(cudaFuncSetAttribute(umma_wgrad_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  dim3 grid(nchunks, ceil_div(w->N1, UMMA_BM), ceil_div(w->N2, NB));
  sard_launch(umma_wgrad_kernel<NB>, grid, UMMA_THREADS, smem, st, *w);
  SARD_LAUNCH_OK("umma_wgrad_kernel");
  return 0;
}

// partial products of the weight gradient on the tensor cores; -1 = shape not handled (caller uses the SIMT kernel)
int sard_umma_wgrad(const SardWGrad* w, int nchunks, cudaStream_t st) {
  if (w->N1 < 32 || w->N2 < 8) return -1;
  if (w->N2 <= 64) return launch_umma_wgrad<64>(w, nchunks, st);
  if (w->N2 <= 128) return launch_umma_wgrad<128>(w, nchunks, st);
  return launch_umma_wgrad<256>(w, nchunks, st);
}

// returns -1 when the shape is not handled by this engine (caller falls through to the SIMT kernels)
int sard_umma_gemm(const SardGemm* g, bool nt, cudaStream_t st) {
  // long reductions stay on the FFMA engine: tensor-core fp32 accumulation truncates, and the bias grows
  // with the number of accumulated k-steps (measured 5e-5 relative at R = 1410)
  if (g->N <= 16 || g->R < 8 || g->R > 512) return -1;
  if (g_sard_gemm_engine >= 2) {     // A operand in tensor memory
    if (g->N <= 64) return nt ? launch_umma_ts<64, true>(g, st) : launch_umma_ts<64, false>(g, st);
    return nt ? launch_umma_ts<128, true>(g, st) : launch_umma_ts<128, false>(g, st);
  }
  if (g->N <= 64) return nt ? launch_umma_t<64, true>(g, st) : launch_umma_t<64, false>(g, st);
  return nt ? launch_umma_t<128, true>(g, st) : launch_umma_t<128, false>(g, st);
}
