// Fused GNN tower on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), error-compensated 3xTF32.
//
// Replaces, for one run of whole graphs, RunningNorm.forward (khrylib/rl/core/running_norm.py:22-42) + GNNSimple.forward
// (design_opt/models/gnn.py:54-68): x -> norm -> in_fc -> L x (GraphConv -> act), with GraphConv the torch-geometric 1.6.1
// operator  y_i = W_l sum_{j->i} x_j + b + W_r x_i.
//
//   tower_fwd_tc_kernel   ONE launch for the whole tower.  Graphs are packed whole into 128-row tiles (one CTA each),
//                         so every neighbour sum is tile-local.  A GraphConv layer is ONE GEMM  x [W_l | W_r]^T  (N = 128,
//                         K = 64; the sum over neighbours commutes with W_l) whose first 64 output columns are
//                         neighbour-summed in the epilogue through shared memory.  The activations never leave the SM
//                         between layers: epilogue -> (bias, act, hi/lo split) -> TENSOR MEMORY = the next GEMM's A
//                         operand.  RunningNorm's normalise + clamp is applied while the in_fc operand is built.
//   tower_bwd_tc_kernel   the adjoint chain, same structure:  dh_{l-1} = A^T(dz_l W_l) + dz_l W_r,  dz_{l-1} = dh act'.
//                         It also writes the operands of the weight gradients as hi|lo pair arrays: [A^T dz_l | dz_l]
//                         (the gradient of W_l is (A^T dz_l)^T h_{l-1}: aggregating dz instead of h).
//   tower_wgrad_tc_kernel all weight and bias gradients of the tower for one chunk of rows: MN-major operands by TMA,
//                         seven accumulators (448 columns) in tensor memory; bias gradients are products with ones.
//   tower_wgrad_reduce_kernel  adds the per-chunk partial products in chunk order (deterministic) into the gradient buffer.
#include "tc.cuh"
#include "tower_tc.cuh"
#include "norm.cuh"
#include <string.h>

#define TT_THREADS 320          // warp 0: TMA producer, warp 1: MMA issuer, warps 2-9: row workers (2 per 32-row quarter)
#define TT_WBYTES 65536         // weight operand of the current layer: hi | lo
#define TT_SBYTES 32768         // staging [128 rows][64 fp32] for the neighbour sums, 16-byte pieces XOR (row & 7)
#define TT_SMEM (TT_WBYTES + TT_SBYTES + 1024)
#ifndef TT_MINB
#define TT_MINB 2
#endif
#define TT_WS (64 * 256)        // floats per pre-split weight block

// ------------------------------------------------------------------------------------------ tile table
__global__ void tower_tiles_kernel(const int* __restrict__ cum, int node_base, int B, int span, int num_tiles, int* __restrict__ tile_first) {
  pdl_sync();
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= B) return;
  const int t = (cum[g] - node_base) / span;
  if (g == 0 || (cum[g - 1] - node_base) / span != t) tile_first[t] = g;
}
__global__ void tower_rows_kernel(const int* __restrict__ cum, int node_base, int n, int span, const int* __restrict__ tile_first,
                                  const int* __restrict__ nd_graph, const int* __restrict__ nd_arena,
                                  const int* __restrict__ in_ptr, const int* __restrict__ in_nbr,
                                  const int* __restrict__ out_ptr, const int* __restrict__ out_nbr,
                                  uint2* __restrict__ tab_in, uint2* __restrict__ tab_out);
int sard_tower_tc_tiles(const SardArena* arena, const SardSel* sel, const TowerTcBufs& tb, const int* nd_graph, const int* nd_arena,
                        sard_stream_t stream) {
  SARD_CUDA(cudaMemsetAsync(tb.tile_first, 0xff, sizeof(int) * (size_t)(tb.num_tiles + 1), as_stream(stream)));
  sard_launch(tower_tiles_kernel, ceil_div(sel->B, 256), 256, 0, as_stream(stream), sel->cum, sel->node_base, sel->B, tb.span, tb.num_tiles,
              tb.tile_first);
  SARD_LAUNCH_OK("tower_tiles_kernel");
  sard_launch(tower_rows_kernel, ceil_div(sel->n, 256), 256, 0, as_stream(stream), sel->cum, sel->node_base, sel->n, tb.span,
              (const int*)tb.tile_first, nd_graph, nd_arena, arena->in_ptr, arena->in_nbr, arena->out_ptr, arena->out_nbr,
              reinterpret_cast<uint2*>(tb.tab_in), reinterpret_cast<uint2*>(tb.tab_out));
  SARD_LAUNCH_OK("tower_rows_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------------ weights
// wsplit: block 0 = in_fc W [64][hi 128 | lo 128] (k >= D zero); blocks 1..L = forward GraphConv operands
// [W_l ; W_r] as [128][hi 64 | lo 64]; blocks 1+L..2L = adjoint operands [W_l^T ; W_r^T] likewise.
struct TtSplitParams { const float* P; long long in_w, wl[TT_MAXL], wr[TT_MAXL]; int D, L; float* out; TowerNormStep norm; };
__global__ void __launch_bounds__(256) tower_presplit_kernel(const TtSplitParams p) {
  pdl_sync();
  const int blk = blockIdx.y;
  if (blk == 1 + 2 * p.L) {
    // the RunningNorm step of this run rides along as one more block row (one launch fewer on the chain): running
    // update from the moment records (train) or inv = 1 / (std + 1e-8) (eval); the forward kernel follows in stream order
    if (blockIdx.x != 0) return;
    const TowerNormStep& q = p.norm;
    if (q.mode == 1) norm_update_body(q.records, q.n_records, p.D, q.mean, q.var, q.std, q.count, q.inv);
    else if (q.mode == 2)
      for (int c = threadIdx.x; c < p.D; c += blockDim.x) q.inv[c] = 1.0f / (q.std[c] + 1e-8f);
    return;
  }
  float* o = p.out + (size_t)blk * TT_WS;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < TT_WS / 2; e += gridDim.x * blockDim.x) {
    float v;
    int row, col, width;
    if (blk == 0) {
      row = e >> 7; col = e & 127; width = 128;
      v = col < p.D ? p.P[p.in_w + (size_t)row * p.D + col] : 0.f;
    } else {
      const bool adj = blk > p.L;
      const int l = (blk - 1) % p.L;
      row = e >> 6; col = e & 63; width = 64;
      const float* W = p.P + (row < 64 ? p.wl[l] : p.wr[l]);
      const int rr = row & 63;
      v = adj ? W[(size_t)col * 64 + rr] : W[(size_t)rr * 64 + col];
    }
    const float h = tc::tf32_hi(v);
    o[(size_t)row * 2 * width + col] = h;
    o[(size_t)row * 2 * width + width + col] = v - h;
  }
}
long long sard_tower_tc_wsplit_floats(int L) { return (long long)(1 + 2 * L) * TT_WS; }
int sard_tower_tc_presplit(const SardTower& tw, const float* P, float* wsplit, const TowerNormStep& norm, sard_stream_t stream) {
  TtSplitParams p;
  p.P = P; p.in_w = tw.in_w; p.D = tw.D; p.L = tw.L; p.out = wsplit; p.norm = norm;
  for (int l = 0; l < tw.L; ++l) { p.wl[l] = tw.wl[l]; p.wr[l] = tw.wr[l]; }
  // one element per thread (no rolled load -> store loop); + one block row for the RunningNorm step when there is one
  sard_launch(tower_presplit_kernel, dim3(32, 1 + 2 * tw.L + (norm.mode ? 1 : 0)), 256, 0, as_stream(stream), p);
  SARD_LAUNCH_OK("tower_presplit_kernel");
  return 0;
}

bool sard_tower_tc_supported(const SardTower& tw, int max_nodes) {
  if (tw.L < 1 || tw.L > TT_MAXL || tw.D < 1 || tw.D > 127 || max_nodes < 1 || max_nodes > TT_MAX_NODES) return false;
  for (int l = 0; l <= tw.L; ++l)
    if (tw.h[l] != TT_H) return false;
  return true;
}

__device__ long long g_tt_dbg[96];
extern "C" int sard_tt_debug_read(long long* out, int n) {
  SARD_CUDA(cudaMemcpyFromSymbol(out, g_tt_dbg, sizeof(long long) * (n < 96 ? n : 96)));
  return 0;
}
#define TT_STAMP(i) do { if (blockIdx.x == 0 && threadIdx.x == 64) g_tt_dbg[i] = clock64(); } while (0)
// ------------------------------------------------------------------------------------------ per-row neighbour table
// One 8-byte word per row and direction: byte 0 = degree (255: more than 6 neighbours, the kernels then walk the arena
// lists), bytes 1..6 = TILE-LOCAL rows of the neighbours.  Built once per run in the gather phase, read by the forward
// (by-destination lists) and the adjoint (by-source lists) with one coalesced load per thread.
__global__ void tower_rows_kernel(const int* __restrict__ cum, int node_base, int n, int span, const int* __restrict__ tile_first,
                                  const int* __restrict__ nd_graph, const int* __restrict__ nd_arena,
                                  const int* __restrict__ in_ptr, const int* __restrict__ in_nbr,
                                  const int* __restrict__ out_ptr, const int* __restrict__ out_nbr,
                                  uint2* __restrict__ tab_in, uint2* __restrict__ tab_out) {
  pdl_sync();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int g = nd_graph[r], a = nd_arena[r];
  const int g_row = cum[g] - node_base;
  const int row_lo = cum[tile_first[g_row / span]] - node_base;
  const int gb = g_row - row_lo;
  for (int dir = 0; dir < 2; ++dir) {
    const int* ptr = dir ? out_ptr : in_ptr;
    const int* nbr = dir ? out_nbr : in_nbr;
    const int e0 = ptr[a], deg = ptr[a + 1] - e0;
    unsigned long long w = deg > 6 ? 255ull : (unsigned long long)deg;
    if (deg <= 6)
      for (int i = 0; i < deg; ++i) w |= (unsigned long long)(gb + nbr[e0 + i]) << (8 * (i + 1));
    (dir ? tab_out : tab_in)[r] = make_uint2((unsigned)w, (unsigned)(w >> 32));
  }
}

// ------------------------------------------------------------------------------------------ shared device pieces
namespace {
// One row worker: tile row `trow` (= TMEM lane), 32 of the 64 columns of every block (c0 = 0 or 32).  The warp (q, c0)
// owns the 4 KB of the staging area that holds ITS rows' ITS columns: rows 32q..32q+31, bytes [4*c0, 4*c0 + 128) of
// each 256-byte row.  Other warps only read it between the two named barriers of a neighbour sum; outside of those
// it doubles as the warp's private transpose scratch for coalesced stores.
struct RowCtx {
  int trow, c0, q, lane; bool valid;
  int deg, e0, gb; unsigned long long nbw;
  uint32_t lane_off;
  int nv; long long row0;
  uint8_t* own;                  // S + 32q*256 + 4*c0
};

__device__ __forceinline__ float4* stage_ptr(const uint8_t* S, int row, int c0, int j) {
  return reinterpret_cast<float4*>(const_cast<uint8_t*>(S) + row * 256 + c0 * 4 + ((j ^ (row & 7)) << 4));
}
__device__ __forceinline__ void stage_write(uint8_t* S, const RowCtx& c, const float* v) {
#pragma unroll
  for (int j = 0; j < 8; ++j) *stage_ptr(S, c.trow, c.c0, j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}
__device__ __forceinline__ void stage_add_row(const uint8_t* S, int row, int c0, float* acc) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 v = *stage_ptr(S, row, c0, j);
    acc[4 * j] += v.x; acc[4 * j + 1] += v.y; acc[4 * j + 2] += v.z; acc[4 * j + 3] += v.w;
  }
}
// acc[0..32) += sum over the neighbours of this row of their staged 32 columns
__device__ __forceinline__ void stage_gather(const uint8_t* S, const RowCtx& c, const int* __restrict__ nbr, float* acc) {
  if (c.deg <= 6) {
#pragma unroll
    for (int i = 0; i < 6; ++i)
      if (i < c.deg) stage_add_row(S, (int)((c.nbw >> (8 * (i + 1))) & 0xff), c.c0, acc);
  } else {
    for (int i = 0; i < c.deg; ++i) stage_add_row(S, c.gb + __ldg(nbr + c.e0 + i), c.c0, acc);
  }
}
// split 32 values into the A operand in tensor memory: columns c0.. of the hi block [0,64) and of the lo block [64,128)
__device__ __forceinline__ void tmem_put32(uint32_t tmem, const RowCtx& c, const float* v) {
#pragma unroll
  for (int j = 0; j < 4; ++j) tc::tmem_st8_split(tmem + c.lane_off + c.c0 + 8 * j, tmem + c.lane_off + 64 + c.c0 + 8 * j, v + 8 * j);
  tc::tmem_st_wait();
}
// Coalesced store of the warp's 32 x 32 block as it sits in the warp's own staging region (lane = row layout): every
// instruction writes four whole 128-byte rows.  MODE 0: values as they are, 1: tf32 hi parts, 2: lo parts.
template <int MODE>
__device__ __forceinline__ void store_from_own(const RowCtx& c, float* dst0, size_t ld) {
  const int pc = c.lane & 7;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int row = 4 * k + (c.lane >> 3);
    float4 t = *reinterpret_cast<const float4*>(c.own + row * 256 + ((pc ^ (row & 7)) << 4));
    if (MODE != 0) {
      float4 h, l;
      tc::split4(t, h, l);
      t = MODE == 1 ? h : l;
    }
    if (row < c.nv) *reinterpret_cast<float4*>(dst0 + (size_t)row * ld + pc * 4) = t;
  }
}
__device__ __forceinline__ void own_write(const RowCtx& c, const float* v) {
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(c.own + c.lane * 256 + ((j ^ (c.lane & 7)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}
// pair store of the block in the own region: hi at column col, lo at column col + lo_off
__device__ __forceinline__ void pair_store_own(const RowCtx& c, float* base, size_t ld, int col, int lo_off) {
  float* d = base + (size_t)c.row0 * ld + col;
  store_from_own<1>(c, d, ld);
  store_from_own<2>(c, d + lo_off, ld);
}
__device__ __forceinline__ void row_setup(RowCtx& c, uint8_t* S, int warp, int lane, int row_lo, int nrows, const uint2* __restrict__ tab,
                                          const int* __restrict__ nd_graph, const int* __restrict__ nd_arena, const int* __restrict__ cum,
                                          int node_base, const int* __restrict__ nptr) {
  c.q = warp & 3;
  c.c0 = ((warp - 2) >> 2) * 32;
  c.lane = lane;
  c.trow = c.q * 32 + lane;
  c.valid = c.trow < nrows;
  c.lane_off = (uint32_t)(c.q * 32) << 16;
  c.row0 = row_lo + c.q * 32; c.nv = min(max(nrows - c.q * 32, 0), 32);
  c.own = S + c.q * 32 * 256 + c.c0 * 4;
  c.deg = 0; c.e0 = 0; c.gb = 0; c.nbw = 0;
  if (c.valid) {
    const int r = row_lo + c.trow;
    const uint2 w = __ldg(tab + r);
    c.nbw = (unsigned long long)w.x | ((unsigned long long)w.y << 32);
    c.deg = (int)(w.x & 0xff);
    if (c.deg == 255) {                                  // rare: walk the arena lists
      const int g = __ldg(nd_graph + r), a = __ldg(nd_arena + r);
      c.gb = __ldg(cum + g) - node_base - row_lo;
      c.e0 = __ldg(nptr + a);
      c.deg = __ldg(nptr + a + 1) - c.e0;
    }
  }
}
__device__ __forceinline__ void act32(float* v, int act) {
  if (act == SARD_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  } else if (act == SARD_ACT_TANH) {
    // SFU form (tc::tanh_fast, absolute error < 5e-7) as in the fused JSMLP: libdevice tanhf is ~40 instructions with
    // divergent branches, and 32 of them per row and layer were 60 % of a GraphConv layer's epilogue
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = tc::tanh_fast(v[j]);
  } else if (act == SARD_ACT_SIGMOID) {
#pragma unroll 4
    for (int j = 0; j < 32; ++j) v[j] = 1.f / (1.f + expf(-v[j]));
  }
}
}  // namespace

// ------------------------------------------------------------------------------------------ forward
struct TfParams {
  CUtensorMap tm_win, tm_wl;
  const int* tile_first; const int* cum; int node_base, B;
  const int* nd_graph; const int* nd_arena; const int* nptr; const int* nnbr; const uint2* tab;   // neighbours by destination
  const float* x; int ldx; int D;
  const float* mean; const float* inv; const long long* count; float clip;
  const float* b_in; const float* bl[TT_MAXL]; int L, act;
  float* hout; float* XH; float* PH[TT_MAXL]; unsigned* mask[TT_MAXL];
};

__global__ void __launch_bounds__(TT_THREADS, TT_MINB) tower_fwd_tc_kernel(const __grid_constant__ TfParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sW = tc::align1024(smem_raw);
  uint8_t* S = sW + TT_WBYTES;
  __shared__ uint64_t bar_wfull, bar_wfree, bar_mma, bar_a, bar_half;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float s_mean[128], s_inv[128], s_bias[1 + TT_MAXL][64];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 1) tc::tmem_alloc(&tmem_slot, 256);
  if (tid == 0) {
    tc::mbar_init(&bar_wfull, 1); tc::mbar_init(&bar_wfree, 1); tc::mbar_init(&bar_mma, 1); tc::mbar_init(&bar_half, 1);
    tc::mbar_init(&bar_a, 256);
    tc::mbar_fence_init();
    tc::tma_prefetch_desc(&p.tm_win); tc::tma_prefetch_desc(&p.tm_wl);
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t AH = tmem, AL = tmem + 64, DD = tmem + 128;
  pdl_sync();

  // a window without a graph start (only the last one can be) is an empty tile: tile_first = -1
  const int g0 = p.tile_first[blockIdx.x];
  const bool active = g0 >= 0;
  int g1 = active ? p.tile_first[blockIdx.x + 1] : 0;
  if (g1 < 0) g1 = p.B;
  const int row_lo = active ? p.cum[g0] - p.node_base : 0, nrows = active ? p.cum[g1] - p.node_base - row_lo : 0;
  const int Dp = (p.D + 7) & ~7;
  const int nhalf = Dp > 64 ? 2 : 1;

  if (!active) {
  } else if (warp == 0) {
    if (lane == 0) {
      tc::mbar_expect_tx(&bar_wfull, TT_WBYTES);
      for (int c = 0; c < 8; ++c) tc::tma_load_2d(sW + c * 8192, &p.tm_win, c * 32, 0, &bar_wfull);
      for (int l = 1; l <= p.L; ++l) {
        tc::mbar_wait(&bar_wfree, (l - 1) & 1);
        tc::mbar_expect_tx(&bar_wfull, TT_WBYTES);
        for (int c = 0; c < 4; ++c) tc::tma_load_2d(sW + c * 16384, &p.tm_wl, c * 32, (l - 1) * 128, &bar_wfull);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t ID64 = tc::idesc_tf32(128, 64, 0, 0), ID128 = tc::idesc_tf32(128, 128, 0, 0);
      const uint32_t sw = tc::smem_u32(sW);
      int an = 0;
      tc::mbar_wait(&bar_wfull, 0);
      for (int h = 0; h < nhalf; ++h) {
        tc::mbar_wait(&bar_a, an & 1); ++an;
        tc::tc_fence_after();
        const int nk = (h == 0 ? (Dp < 64 ? Dp : 64) : Dp - 64) >> 3;
        for (int ks = 0; ks < nk; ++ks) {
          const int kk = h * 64 + ks * 8;
          const uint32_t off = (uint32_t)(kk >> 5) * 8192 + (uint32_t)(kk & 31) * 4;
          const uint64_t b_hi = tc::kdesc(sw + off), b_lo = tc::kdesc(sw + 32768 + off);
          tc::mma_ts(DD, AH + ks * 8, b_hi, ID64, (h | ks) ? 1u : 0u);
          tc::mma_ts(DD, AL + ks * 8, b_hi, ID64, 1u);
          tc::mma_ts(DD, AH + ks * 8, b_lo, ID64, 1u);
        }
        if (h + 1 < nhalf) tc::mma_commit(&bar_half);
      }
      tc::mma_commit(&bar_wfree);
      tc::mma_commit(&bar_mma);
      for (int l = 1; l <= p.L; ++l) {
        tc::mbar_wait(&bar_wfull, l & 1);
        tc::mbar_wait(&bar_a, an & 1); ++an;
        tc::tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t off = (uint32_t)(ks >> 2) * 16384 + (uint32_t)(ks & 3) * 32;
          const uint64_t b_hi = tc::kdesc(sw + off), b_lo = tc::kdesc(sw + 32768 + off);
          tc::mma_ts(DD, AH + ks * 8, b_hi, ID128, ks ? 1u : 0u);
          tc::mma_ts(DD, AL + ks * 8, b_hi, ID128, 1u);
          tc::mma_ts(DD, AH + ks * 8, b_lo, ID128, 1u);
        }
        tc::mma_commit(&bar_wfree);
        tc::mma_commit(&bar_mma);
      }
    }
  } else {
    RowCtx c;
    TT_STAMP(32);
    row_setup(c, S, warp, lane, row_lo, nrows, p.tab, p.nd_graph, p.nd_arena, p.cum, p.node_base, p.nptr);
    const bool normalise = *p.count > 0;
    const long long r = row_lo + c.trow;
    // RunningNorm statistics and every bias go to shared memory once (read back as broadcast float4)
    {
      const int wt = tid - 64;
      if (wt < 128) {
        s_mean[wt] = (normalise && wt < p.D) ? __ldg(p.mean + wt) : 0.f;
        s_inv[wt] = (normalise && wt < p.D) ? __ldg(p.inv + wt) : 1.f;
      } else {
        const int b = wt - 128, l = b >> 6, j = b & 63;        // l = 0: in_fc, 1: GraphConv layer 1
        s_bias[l][j] = l == 0 ? __ldg(p.b_in + j) : (p.L >= 1 && p.bl[0] ? __ldg(p.bl[0] + j) : 0.f);
        if (TT_MAXL >= 2) s_bias[2 + l][j] = (p.L >= 2 + l && p.bl[1 + l]) ? __ldg(p.bl[1 + l] + j) : 0.f;
      }
      tc::named_sync(3, 256);
    }
    float v[32];
    // ---- in_fc operand: normalised + clamped input row, 64 columns of K per pass.  The warp's 32 x 32 block of x is
    // read with whole-row (128-byte) transactions into the warp's own staging region and picked up one row per lane.
    for (int h = 0; h < 2; ++h) {
      // D < 64: the second 64 columns of XH are all zero (the column of ones sits at D) and are not stored at all -- the
      // weight-gradient kernel's tensor map ends at column 64 and TMA zero-fills the rest of its box
      if (h >= nhalf && (!p.XH || p.D < 64)) break;
      const int k0 = h * 64 + c.c0;
      {
        const int pc = lane & 7, kcol = k0 + pc * 4;
        float4 t4[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int row = 4 * k + (lane >> 3);
          t4[k] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (row < c.nv && kcol < p.ldx) t4[k] = *reinterpret_cast<const float4*>(p.x + (size_t)(c.row0 + row) * p.ldx + kcol);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int row = 4 * k + (lane >> 3);
          *reinterpret_cast<float4*>(c.own + row * 256 + ((pc ^ (row & 7)) << 4)) = t4[k];
        }
        __syncwarp();
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = k0 + 4 * j;
        const float4 xv = *reinterpret_cast<const float4*>(c.own + lane * 256 + ((j ^ (lane & 7)) << 4));
        const float4 mu = *reinterpret_cast<const float4*>(s_mean + (k & 127)), iv = *reinterpret_cast<const float4*>(s_inv + (k & 127));
        float t4[4] = {(xv.x - mu.x) * iv.x, (xv.y - mu.y) * iv.y, (xv.z - mu.z) * iv.z, (xv.w - mu.w) * iv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float t = t4[i];
          if (normalise) t = fminf(fmaxf(t, -p.clip), p.clip);
          v[4 * j + i] = (k + i < p.D) ? t : 0.f;
        }
      }
      __syncwarp();
      if (h < nhalf) {
        if (h == 1) { tc::mbar_wait(&bar_half, 0); tc::tc_fence_after(); }
        tmem_put32(tmem, c, v);
        tc::tc_fence_before();
        tc::mbar_arrive(&bar_a);
      }
      if (p.XH) {                                        // weight-gradient operand, with the column of ones at D
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = (k0 + j == p.D) ? 1.f : v[j];
        own_write(c, v);
        __syncwarp();
        store_from_own<0>(c, p.XH + (size_t)c.row0 * 128 + k0, 128);
        __syncwarp();
      }
    }
    // ---- in_fc epilogue: h_0 = x W^T + b (no activation)
    int mm = 0;
    TT_STAMP(33);
    tc::mbar_wait(&bar_mma, mm & 1); ++mm;
    tc::tc_fence_after();
    TT_STAMP(34);
    tc::tmem_ld32(DD + c.lane_off + c.c0, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b4 = *reinterpret_cast<const float4*>(&s_bias[0][c.c0 + 4 * j]);
      v[4 * j] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
    }
    tmem_put32(tmem, c, v);
    tc::tc_fence_before();
    tc::mbar_arrive(&bar_a);
    if (p.PH[0]) {                                       // under the next GEMM
      own_write(c, v);
      __syncwarp();
      store_from_own<0>(c, p.PH[0] + (size_t)c.row0 * TT_H + c.c0, TT_H);
      __syncwarp();
    }
    // ---- GraphConv layers
#pragma unroll 1
    for (int l = 1; l <= p.L; ++l) {
      TT_STAMP(35 + 5 * (l - 1));
      tc::mbar_wait(&bar_mma, mm & 1); ++mm;
      tc::tc_fence_after();
      TT_STAMP(36 + 5 * (l - 1));
      tc::tmem_ld32(DD + c.lane_off + c.c0, v);          // (x W_l^T) of this row: to be summed over the neighbours
      stage_write(S, c, v);
      tc::tmem_ld32(DD + c.lane_off + 64 + c.c0, v);     // x W_r^T
      tc::named_sync(1, 256);
#pragma unroll
      for (int j = 0; j < 8; ++j) {                       // zero when the layer has no bias
        const float4 b4 = *reinterpret_cast<const float4*>(&s_bias[l][c.c0 + 4 * j]);
        v[4 * j] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
      }
      TT_STAMP(37 + 5 * (l - 1));
      stage_gather(S, c, p.nnbr, v);
      act32(v, p.act);
      TT_STAMP(38 + 5 * (l - 1));
      if (l < p.L) {
        tmem_put32(tmem, c, v);
        tc::tc_fence_before();
        tc::mbar_arrive(&bar_a);
        if (p.mask[l] && c.valid) {                      // act' for the adjoint: one bit per element when it is a ReLU
          unsigned m = 0;
#pragma unroll
          for (int j = 0; j < 32; ++j) m |= (v[j] > 0.f ? 1u : 0u) << j;
          p.mask[l][r * 2 + (c.c0 >> 5)] = m;
        }
      }
      tc::named_sync(2, 256);                             // every gather is done: the own region is private again
      TT_STAMP(39 + 5 * (l - 1));
      if (l < p.L ? p.PH[l] != nullptr : true) {
        own_write(c, v);
        __syncwarp();
        if (l < p.L) store_from_own<0>(c, p.PH[l] + (size_t)c.row0 * TT_H + c.c0, TT_H);    // under the next GEMM
        else store_from_own<0>(c, p.hout + (size_t)c.row0 * TT_H + c.c0, TT_H);
        __syncwarp();
      }
    }
  }
  TT_STAMP(50);
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem, 256);
}

// ------------------------------------------------------------------------------------------ backward (adjoint chain)
struct TbParams {
  CUtensorMap tm_wl;
  const int* tile_first; const int* cum; int node_base, B;
  const int* nd_graph; const int* nd_arena; const int* nptr; const int* nnbr; const uint2* tab;   // neighbours by source
  const float* dzL;                  // [n, 64]
  const float* PH[TT_MAXL];          // h_l pairs (act' of layers 1..L-1 when the activation is not a ReLU)
  const unsigned* mask[TT_MAXL];     // ReLU: h_l > 0, one bit per element
  float* PA[TT_MAXL]; float* PZ0;
  int L, act;
};

__global__ void __launch_bounds__(TT_THREADS, TT_MINB) tower_bwd_tc_kernel(const __grid_constant__ TbParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sW = tc::align1024(smem_raw);
  uint8_t* S = sW + TT_WBYTES;
  __shared__ uint64_t bar_wfull, bar_wfree, bar_mma, bar_a;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 1) tc::tmem_alloc(&tmem_slot, 256);
  if (tid == 0) {
    tc::mbar_init(&bar_wfull, 1); tc::mbar_init(&bar_wfree, 1); tc::mbar_init(&bar_mma, 1);
    tc::mbar_init(&bar_a, 256);
    tc::mbar_fence_init();
    tc::tma_prefetch_desc(&p.tm_wl);
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t AH = tmem, AL = tmem + 64, DD = tmem + 128;
  pdl_sync();

  const int g0 = p.tile_first[blockIdx.x];
  const bool active = g0 >= 0;
  int g1 = active ? p.tile_first[blockIdx.x + 1] : 0;
  if (g1 < 0) g1 = p.B;
  const int row_lo = active ? p.cum[g0] - p.node_base : 0, nrows = active ? p.cum[g1] - p.node_base - row_lo : 0;

  if (!active) {
  } else if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < p.L; ++i) {                     // layers L, L-1, .., 1
        const int l = p.L - i;
        if (i > 0) tc::mbar_wait(&bar_wfree, (i - 1) & 1);
        tc::mbar_expect_tx(&bar_wfull, TT_WBYTES);
        for (int c = 0; c < 4; ++c) tc::tma_load_2d(sW + c * 16384, &p.tm_wl, c * 32, (l - 1) * 128, &bar_wfull);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t ID128 = tc::idesc_tf32(128, 128, 0, 0);
      const uint32_t sw = tc::smem_u32(sW);
      for (int i = 0; i < p.L; ++i) {
        tc::mbar_wait(&bar_wfull, i & 1);
        tc::mbar_wait(&bar_a, i & 1);
        tc::tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t off = (uint32_t)(ks >> 2) * 16384 + (uint32_t)(ks & 3) * 32;
          const uint64_t b_hi = tc::kdesc(sw + off), b_lo = tc::kdesc(sw + 32768 + off);
          tc::mma_ts(DD, AH + ks * 8, b_hi, ID128, ks ? 1u : 0u);
          tc::mma_ts(DD, AL + ks * 8, b_hi, ID128, 1u);
          tc::mma_ts(DD, AH + ks * 8, b_lo, ID128, 1u);
        }
        tc::mma_commit(&bar_wfree);
        tc::mma_commit(&bar_mma);
      }
    }
  } else {
    RowCtx c;
    TT_STAMP(0);
    row_setup(c, S, warp, lane, row_lo, nrows, p.tab, p.nd_graph, p.nd_arena, p.cum, p.node_base, p.nptr);
    TT_STAMP(1);
    const long long r = row_lo + c.trow;
    float v[32];
    // ---- dz_L: the warp's 32 x 32 block read with whole 128-byte row pieces through its own staging region
    {
      const int pc = lane & 7;
      float4 t4[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int row = 4 * k + (lane >> 3);
        t4[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < c.nv) t4[k] = *reinterpret_cast<const float4*>(p.dzL + (size_t)(c.row0 + row) * TT_H + c.c0 + pc * 4);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int row = 4 * k + (lane >> 3);
        *reinterpret_cast<float4*>(c.own + row * 256 + ((pc ^ (row & 7)) << 4)) = t4[k];
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 t = *reinterpret_cast<const float4*>(c.own + lane * 256 + ((j ^ (lane & 7)) << 4));
        v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
      }
      __syncwarp();
    }
#pragma unroll 1
    for (int i = 0; i < p.L; ++i) {
      const int l = p.L - i;                              // v = dz_l of this row
      TT_STAMP(2 + 6 * i);
      stage_write(S, c, v);
      tmem_put32(tmem, c, v);
      tc::tc_fence_before();
      tc::mbar_arrive(&bar_a);                            // the GEMM runs under the neighbour sum and the stores below
      tc::named_sync(1, 256);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = 0.f;
      stage_gather(S, c, p.nnbr, v);                      // A^T dz_l
      TT_STAMP(3 + 6 * i);
      tc::named_sync(2, 256);                             // every gather is done: the own region (= dz_l) is private again
      store_from_own<0>(c, p.PA[l - 1] + (size_t)c.row0 * 128 + 64 + c.c0, 128);
      __syncwarp();
      own_write(c, v);
      __syncwarp();
      store_from_own<0>(c, p.PA[l - 1] + (size_t)c.row0 * 128 + c.c0, 128);
      __syncwarp();
      // ---- dz_l [W_l | W_r]: first half is summed over the out-neighbours, second half is the self term
      TT_STAMP(4 + 6 * i);
      tc::mbar_wait(&bar_mma, i & 1);
      tc::tc_fence_after();
      TT_STAMP(5 + 6 * i);
      tc::tmem_ld32(DD + c.lane_off + c.c0, v);
      stage_write(S, c, v);
      tc::tmem_ld32(DD + c.lane_off + 64 + c.c0, v);
      tc::named_sync(1, 256);
      TT_STAMP(6 + 6 * i);
      stage_gather(S, c, p.nnbr, v);                      // dh_{l-1}
      TT_STAMP(7 + 6 * i);
      if (l - 1 >= 1) {
        if (c.valid) {
          if (p.mask[l - 1]) {
            const unsigned m = __ldg(p.mask[l - 1] + r * 2 + (c.c0 >> 5));
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = (m >> j) & 1u ? v[j] : 0.f;
          } else {
            const float* hp = p.PH[l - 1] + (size_t)r * TT_H + c.c0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 hh = __ldg(reinterpret_cast<const float4*>(hp + 4 * j));
              v[4 * j] *= act_grad_from_output(hh.x, p.act); v[4 * j + 1] *= act_grad_from_output(hh.y, p.act);
              v[4 * j + 2] *= act_grad_from_output(hh.z, p.act); v[4 * j + 3] *= act_grad_from_output(hh.w, p.act);
            }
          }
        }
        tc::named_sync(2, 256);                           // the staged d_agg rows have been read by everyone
      } else {
        tc::named_sync(2, 256);
        own_write(c, v);
        __syncwarp();
        store_from_own<0>(c, p.PZ0 + (size_t)c.row0 * TT_H + c.c0, TT_H);
      }
    }
  }
  TT_STAMP(30);
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem, 256);
}

// ------------------------------------------------------------------------------------------ weight gradients
#define TWG_ROWS 8
#ifndef TWG_STAGES
#define TWG_STAGES 4
#endif
#define TWG_OFF_A(l) ((l) * 12288)             // layer l = 1..3 at slot l-1: [A^T dz | dz] 8 slabs (hi 4 | lo 4) then h_{l-1} 4 slabs (hi 2 | lo 2)
#define TWG_OFF_X 36864                        // x pair 8 slabs, then dz_0 pair 4 slabs
#define TWG_STAGE_BYTES 49152
#define TWG_SMEM (TWG_STAGES * TWG_STAGE_BYTES + 1024 + 1024)
#define TWG_COLS 448                           // dW_in^T 64 | [dW_l ; dW_r] 64 x 3 | bias rows 64 x 3
#define TWG_THREADS 320

struct TwParams {
  CUtensorMap tm_pa[TT_MAXL], tm_ph[TT_MAXL], tm_xh, tm_z0;
  int n, chunk_rows, L;
  float* partials;
};

__global__ void __launch_bounds__(TWG_THREADS, 1) tower_wgrad_tc_kernel(const __grid_constant__ TwParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = tc::align1024(smem_raw);
  uint8_t* ones = base + TWG_STAGES * TWG_STAGE_BYTES;
  __shared__ uint64_t bar_full[TWG_STAGES], bar_ready[TWG_STAGES], bar_empty[TWG_STAGES], bar_done;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 1) tc::tmem_alloc(&tmem_slot, 512);
  if (tid == 0) {
    for (int s = 0; s < TWG_STAGES; ++s) { tc::mbar_init(&bar_full[s], 1); tc::mbar_init(&bar_ready[s], TWG_THREADS - 64); tc::mbar_init(&bar_empty[s], 1); }
    tc::mbar_init(&bar_done, 1);
    tc::mbar_fence_init();
  }
  if (tid >= 64) reinterpret_cast<float*>(ones)[tid - 64] = 1.0f;
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  TT_STAMP(64);
  pdl_sync();
  TT_STAMP(65);

  const int row_start = blockIdx.x * p.chunk_rows;
  const int row_end = min(p.n, row_start + p.chunk_rows);
  const int nks = (row_end - row_start + TWG_ROWS - 1) / TWG_ROWS;      // rows past n are zero-filled by TMA
  const uint32_t stage_bytes = ((uint32_t)p.L * 12288 + 12288) / 2;    // TMA brings the fp32 operands; their lo halves are made on the SM
  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < nks; ++it) {
        const int s = it % TWG_STAGES;
        tc::mbar_wait(&bar_empty[s], ((it / TWG_STAGES) & 1) ^ 1);
        uint8_t* st = base + s * TWG_STAGE_BYTES;
        const int row = row_start + it * TWG_ROWS;
        tc::mbar_expect_tx(&bar_full[s], stage_bytes);
        for (int l = 1; l <= p.L; ++l) {
          tc::tma_load_3d(st + TWG_OFF_A(l - 1), &p.tm_pa[l - 1], 0, row, 0, &bar_full[s]);
          tc::tma_load_3d(st + TWG_OFF_A(l - 1) + 8192, &p.tm_ph[l - 1], 0, row, 0, &bar_full[s]);
        }
        tc::tma_load_3d(st + TWG_OFF_X, &p.tm_xh, 0, row, 0, &bar_full[s]);
        tc::tma_load_3d(st + TWG_OFF_X + 8192, &p.tm_z0, 0, row, 0, &bar_full[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t ID_W = tc::idesc_tf32(128, 64, 1, 1), ID_B = tc::idesc_tf32(128, 64, 0, 1);
      constexpr uint32_t SL = TWG_ROWS * 128;
      const uint64_t one_k = tc::smem_desc_sw128(tc::smem_u32(ones), 16, 0);
      for (int it = 0; it < nks; ++it) {
        const int s = it % TWG_STAGES;
        tc::mbar_wait(&bar_ready[s], (it / TWG_STAGES) & 1);
        tc::tc_fence_after();
        const uint32_t sa = tc::smem_u32(base + s * TWG_STAGE_BYTES);
        const uint32_t acc = it ? 1u : 0u;
        auto mn = [&](uint32_t off) { return tc::smem_desc_mn32(sa + off, SL, 512); };
        {   // in_fc: x^T dz_0 (row D of the result = bias gradient through the column of ones)
          const uint64_t xh = mn(TWG_OFF_X), xl = mn(TWG_OFF_X + 4 * SL), zh = mn(TWG_OFF_X + 8192), zl = mn(TWG_OFF_X + 8192 + 2 * SL);
          tc::mma_ss(tmem, xh, zh, ID_W, acc); tc::mma_ss(tmem, xl, zh, ID_W, 1u); tc::mma_ss(tmem, xh, zl, ID_W, 1u);
        }
        for (int l = 1; l <= p.L; ++l) {
          const uint32_t o = TWG_OFF_A(l - 1);
          const uint64_t ah = mn(o), al = mn(o + 4 * SL), hh = mn(o + 8192), hl = mn(o + 8192 + 2 * SL);
          const uint64_t zh = mn(o + 2 * SL), zl = mn(o + 6 * SL);             // the dz half of [A^T dz | dz]
          const uint32_t DW = tmem + 64 * l, DB = tmem + 256 + 64 * (l - 1);
          tc::mma_ss(DW, ah, hh, ID_W, acc); tc::mma_ss(DW, al, hh, ID_W, 1u); tc::mma_ss(DW, ah, hl, ID_W, 1u);
          tc::mma_ss(DB, one_k, zh, ID_B, acc); tc::mma_ss(DB, one_k, zl, ID_B, 1u);
        }
        tc::mma_commit(&bar_empty[s]);
      }
      tc::mma_commit(&bar_done);
    }
  } else {
    // ---- operand split: the operands are stored once, as fp32 (half the HBM traffic of hi|lo pair arrays: this kernel
    // was DRAM-bound on them); TMA lands them in the "hi" slabs of a stage and these eight warps rewrite every value v in
    // place as hi = tf32(v) and put lo = v - hi into the matching "lo" slab (same swizzled position: the split is
    // element-wise, so the layout never has to be decoded), then hand the stage to the MMA warp
    {
      const int st_t = tid - 64;                          // 0 .. 255
      const int n4 = (p.L + 1) * 384;                     // float4 per stage: per layer block 256 ([A^T dz | dz] or x) + 128 (h or dz_0)
      for (int it = 0; it < nks; ++it) {
        const int s = it % TWG_STAGES;
        tc::mbar_wait(&bar_full[s], (it / TWG_STAGES) & 1);
        uint8_t* stg = base + s * TWG_STAGE_BYTES;
        for (int i = st_t; i < n4; i += TWG_THREADS - 64) {
          const int blk = i / 384, j = i - blk * 384;
          uint8_t* src = stg + (blk < p.L ? TWG_OFF_A(blk) : TWG_OFF_X) + (j < 256 ? j * 16 : 8192 + (j - 256) * 16);
          const float4 v = *reinterpret_cast<const float4*>(src);
          float4 h, l;
          tc::split4(v, h, l);
          *reinterpret_cast<float4*>(src) = h;
          *reinterpret_cast<float4*>(src + (j < 256 ? 4096 : 2048)) = l;
        }
        tc::fence_async_smem();                           // generic-proxy writes -> visible to the tensor core's reads
        tc::mbar_arrive(&bar_ready[s]);
      }
    }
    // two warps per TMEM lane quarter, seven 32-column groups each; the idle pipeline stages are the transpose scratch
    const int q = warp & 3, half = (warp - 2) >> 2;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    uint8_t* scr = base + (warp - 2) * 4096;
    float* dst = p.partials + ((size_t)blockIdx.x * 128 + q * 32) * TWG_COLS;
    TT_STAMP(66);
    tc::mbar_wait(&bar_done, 0);
    tc::tc_fence_after();
    TT_STAMP(67);
#pragma unroll 1
    for (int g = half * 7; g < half * 7 + 7; ++g)
      tc::tmem_block_to_global(tmem + lane_off + g * 32, scr, lane, dst + g * 32, TWG_COLS);
  }
  TT_STAMP(68);
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem, 512);
}

// A block owns 32 float4 of the [128 x 448] gradient block; its 8 warps each add every 8th chunk, the 8 slice sums are
// added in slice order (deterministic), then routed into the flat gradient buffer.
struct TwReduceParams {
  const float* partials; int num_chunks, D, L, has_bias;
  float* G; long long in_w, in_b, wl[TT_MAXL], wr[TT_MAXL], bl[TT_MAXL];
};
__global__ void __launch_bounds__(256) tower_wgrad_reduce_kernel(const __grid_constant__ TwReduceParams p) {
  pdl_sync();
  __shared__ float4 part[8][32];
  constexpr int PER = 128 * TWG_COLS / 4;
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int e4 = blockIdx.x * 32 + lane;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (e4 < PER) {
    const float4* P4 = reinterpret_cast<const float4*>(p.partials) + e4;
    int k = slice;
    for (; k + 24 < p.num_chunks; k += 32) {
      const float4 a0 = P4[(size_t)k * PER], a1 = P4[(size_t)(k + 8) * PER], a2 = P4[(size_t)(k + 16) * PER], a3 = P4[(size_t)(k + 24) * PER];
      s.x += a0.x; s.y += a0.y; s.z += a0.z; s.w += a0.w;
      s.x += a1.x; s.y += a1.y; s.z += a1.z; s.w += a1.w;
      s.x += a2.x; s.y += a2.y; s.z += a2.z; s.w += a2.w;
      s.x += a3.x; s.y += a3.y; s.z += a3.z; s.w += a3.w;
    }
    for (; k < p.num_chunks; k += 8) {
      const float4 a = P4[(size_t)k * PER];
      s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
    }
  }
  part[slice][lane] = s;
  __syncthreads();
  if (slice != 0 || e4 >= PER) return;
#pragma unroll
  for (int j = 1; j < 8; ++j) {
    const float4 a = part[j][lane];
    s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
  }
  const int m = e4 / (TWG_COLS / 4), c = (e4 - m * (TWG_COLS / 4)) * 4;
  const float v[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int cc = c + i;
    if (cc < 64) {
      if (m < p.D) p.G[p.in_w + (size_t)cc * p.D + m] += v[i];
      else if (m == p.D) p.G[p.in_b + cc] += v[i];
    } else if (cc < 256) {
      const int l = cc >> 6, nn = cc & 63;
      if (l <= p.L) {
        if (m < 64) p.G[p.wl[l - 1] + (size_t)m * 64 + nn] += v[i];
        else p.G[p.wr[l - 1] + (size_t)(m - 64) * 64 + nn] += v[i];
      }
    } else {
      const int l = ((cc - 256) >> 6) + 1, nn = (cc - 256) & 63;
      if (m == 0 && l <= p.L && p.has_bias) p.G[p.bl[l - 1] + nn] += v[i];
    }
  }
}

// ------------------------------------------------------------------------------------------ host
long long sard_tower_tc_partials_floats(int num_chunks) { return (long long)(num_chunks > 0 ? num_chunks : 1) * 128 * TWG_COLS; }
int sard_tower_tc_chunk_rows(int n) {
  // 100 row chunks (not one per SM): fewer partial blocks to write and reduce, and SMs left for the other net's kernels
  // (measured 74: 109.7, 100: 109.1, 148: 110.8, 296: 117.9 ms / update)
  static const int target = getenv("SARD_TOWER_WG_CHUNKS") ? atoi(getenv("SARD_TOWER_WG_CHUNKS")) : 100;
  int rows = (n + target - 1) / target;
  rows = (rows + 7) & ~7;
  return rows < 64 ? 64 : rows;
}

int sard_tower_tc_forward(const SardTower& tw, const float* P, const SardArena* arena, const SardSel* sel, const TowerTcBufs& tb,
                          const float* x, int ldx, const int* nd_graph, const int* nd_arena, float* hout, bool train,
                          sard_stream_t stream) {
  cudaStream_t st = as_stream(stream);
  TfParams p;
  memset(&p, 0, sizeof(p));
  SARD_TRY(sard_tmap_2d(&p.tm_win, tb.wsplit, 64, 256, 256, 64, 32));
  SARD_TRY(sard_tmap_2d(&p.tm_wl, tb.wsplit + TT_WS, (long long)tw.L * 128, 128, 128, 128, 32));
  p.tile_first = tb.tile_first; p.cum = sel->cum; p.node_base = sel->node_base; p.B = sel->B;
  p.nd_graph = nd_graph; p.nd_arena = nd_arena; p.nptr = arena->in_ptr; p.nnbr = arena->in_nbr; p.tab = reinterpret_cast<const uint2*>(tb.tab_in);
  p.x = x; p.ldx = ldx; p.D = tw.D; p.mean = tw.mean; p.inv = tw.inv; p.count = tw.count; p.clip = 5.0f;
  p.b_in = P + tw.in_b; p.L = tw.L; p.act = tw.act;
  for (int l = 0; l < tw.L; ++l) {
    p.bl[l] = tw.has_bias ? P + tw.bl[l] : nullptr; p.PH[l] = train ? tb.PH[l] : nullptr;
    p.mask[l] = train && tw.act == SARD_ACT_RELU ? tb.mask[l] : nullptr;
  }
  p.hout = hout; p.XH = train ? tb.XH : nullptr;
  static bool configured[SARD_MAX_DEVICES] = {};
  if (sard_first_use_on_device(configured))
    SARD_CUDA(cudaFuncSetAttribute(tower_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TT_SMEM));
  ProfScope prof(st, 2.0 * sel->n * ((double)tw.D * TT_H + (double)tw.L * 2 * TT_H * TT_H),
                 4.0 * sel->n * ((double)tw.D + TT_H));
  sard_launch(tower_fwd_tc_kernel, dim3(tb.num_tiles), dim3(TT_THREADS), (size_t)TT_SMEM, st, p);
  SARD_LAUNCH_OK("tower_fwd_tc_kernel");
  return 0;
}

int sard_tower_tc_backward(const SardTower& tw, const SardArena* arena, const SardSel* sel, const TowerTcBufs& tb, const float* dzL,
                           const int* nd_graph, const int* nd_arena, sard_stream_t stream) {
  cudaStream_t st = as_stream(stream);
  TbParams p;
  memset(&p, 0, sizeof(p));
  SARD_TRY(sard_tmap_2d(&p.tm_wl, tb.wsplit + (size_t)(1 + tw.L) * TT_WS, (long long)tw.L * 128, 128, 128, 128, 32));
  p.tile_first = tb.tile_first; p.cum = sel->cum; p.node_base = sel->node_base; p.B = sel->B;
  p.nd_graph = nd_graph; p.nd_arena = nd_arena; p.nptr = arena->out_ptr; p.nnbr = arena->out_nbr; p.tab = reinterpret_cast<const uint2*>(tb.tab_out);
  p.dzL = dzL; p.L = tw.L; p.act = tw.act; p.PZ0 = tb.PZ0;
  for (int l = 0; l < tw.L; ++l) { p.PH[l] = tb.PH[l]; p.PA[l] = tb.PA[l]; p.mask[l] = tw.act == SARD_ACT_RELU ? tb.mask[l] : nullptr; }
  static bool configured[SARD_MAX_DEVICES] = {};
  if (sard_first_use_on_device(configured))
    SARD_CUDA(cudaFuncSetAttribute(tower_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TT_SMEM));
  ProfScope prof(st, 2.0 * sel->n * ((double)tw.L * 2 * TT_H * TT_H), 4.0 * sel->n * (double)(2 * TT_H));
  sard_launch(tower_bwd_tc_kernel, dim3(tb.num_tiles), dim3(TT_THREADS), (size_t)TT_SMEM, st, p);
  SARD_LAUNCH_OK("tower_bwd_tc_kernel");
  return 0;
}

int sard_tower_tc_wgrad(const SardTower& tw, const SardSel* sel, const TowerTcBufs& tb, float* G, sard_stream_t stream) {
  cudaStream_t st = as_stream(stream);
  TwParams p;
  memset(&p, 0, sizeof(p));
  const int n = sel->n;
  for (int l = 0; l < tw.L; ++l) {
    SARD_TRY(sard_tmap_slab(&p.tm_pa[l], tb.PA[l], n, 128, 128, TWG_ROWS, 4));
    SARD_TRY(sard_tmap_slab(&p.tm_ph[l], tb.PH[l], n, 64, 64, TWG_ROWS, 2));
  }
  // D < 64: only the first two 32-column slabs of XH exist (forward kernel); the out-of-bounds slabs of the box are
  // zero-filled by TMA (and still count towards the stage's transaction bytes)
  SARD_TRY(sard_tmap_slab(&p.tm_xh, tb.XH, n, tw.D < 64 ? 64 : 128, 128, TWG_ROWS, 4));
  SARD_TRY(sard_tmap_slab(&p.tm_z0, tb.PZ0, n, 64, 64, TWG_ROWS, 2));
  p.n = n; p.chunk_rows = tb.chunk_rows; p.L = tw.L; p.partials = tb.partials;
  static bool configured[SARD_MAX_DEVICES] = {};
  if (sard_first_use_on_device(configured))
    SARD_CUDA(cudaFuncSetAttribute(tower_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TWG_SMEM));
  ProfScope prof(st, 2.0 * n * ((double)tw.D * TT_H + (double)tw.L * 2 * TT_H * TT_H), 4.0 * n * ((double)tw.D + (double)tw.L * 3 * TT_H + TT_H));
  sard_launch(tower_wgrad_tc_kernel, dim3(tb.num_chunks), dim3(TWG_THREADS), (size_t)TWG_SMEM, st, p);
  SARD_LAUNCH_OK("tower_wgrad_tc_kernel");
  TwReduceParams r;
  memset(&r, 0, sizeof(r));
  r.partials = tb.partials; r.num_chunks = tb.num_chunks; r.D = tw.D; r.L = tw.L; r.has_bias = tw.has_bias; r.G = G;
  r.in_w = tw.in_w; r.in_b = tw.in_b;
  for (int l = 0; l < tw.L; ++l) { r.wl[l] = tw.wl[l]; r.wr[l] = tw.wr[l]; r.bl[l] = tw.bl[l]; }
  sard_launch(tower_wgrad_reduce_kernel, dim3(ceil_div(128 * TWG_COLS / 4, 32)), dim3(256), 0, st, r);
  SARD_LAUNCH_OK("tower_wgrad_reduce_kernel");
  return 0;
}
