// Packing kernels: neighbour lists, orbit representatives, run gather, stable sort by body index.
// All integer layouts here are bit-exact restatements of what the reference builds with numpy /
// torch indexing (transform2act_policy.py:183-202, 421-436; jsmlp.py:33-39).
#include "common.cuh"

// ------------------------------------------------------------------------------------------
// Neighbour lists per graph.  One thread owns one graph and walks its edges in order, so each
// node's list keeps the edge order (the order torch's CPU index_add_ sums in).
// ------------------------------------------------------------------------------------------
__global__ void csr_kernel(const int* __restrict__ node_ptr, const int* __restrict__ edge_ptr,
                           const int* __restrict__ key, const int* __restrict__ val, int T,
                           int* __restrict__ ptr, int* __restrict__ nbr) {
  pdl_sync();
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= T) return;
  const int gn = node_ptr[g], N = node_ptr[g + 1] - gn;
  const int e0 = edge_ptr[g], E = edge_ptr[g + 1] - e0;
  for (int v = 0; v < N; ++v) ptr[gn + v] = 0;
  for (int e = 0; e < E; ++e) ptr[gn + key[e0 + e]] += 1;
  int run = e0;
  for (int v = 0; v < N; ++v) { int c = ptr[gn + v]; ptr[gn + v] = run; run += c; }
  for (int e = 0; e < E; ++e) { int k = key[e0 + e]; nbr[ptr[gn + k]++] = val[e0 + e]; }
  for (int v = N - 1; v >= 1; --v) ptr[gn + v] = ptr[gn + v - 1];
  if (N > 0) ptr[gn] = e0;
  if (g == T - 1) ptr[gn + N] = e0 + E;
}

extern "C" int sard_build_csr(const int* node_ptr, const int* edge_ptr, const int* esrc, const int* edst, int T,
                              int* in_ptr, int* in_nbr, int* out_ptr, int* out_nbr, sard_stream_t stream) {
  if (T <= 0) return 0;
  const int threads = 128, blocks = ceil_div(T, threads);
  sard_launch(csr_kernel, blocks, threads, 0, as_stream(stream), node_ptr, edge_ptr, edst, esrc, T, in_ptr, in_nbr);
  SARD_LAUNCH_OK("csr_kernel(in)");
  sard_launch(csr_kernel, blocks, threads, 0, as_stream(stream), node_ptr, edge_ptr, esrc, edst, T, out_ptr, out_nbr);
  SARD_LAUNCH_OK("csr_kernel(out)");
  return 0;
}

// ------------------------------------------------------------------------------------------
// Orbit representatives.
// ------------------------------------------------------------------------------------------
__global__ void orbit_rep_kernel(const int* __restrict__ node_ptr, const int* __restrict__ body, int T,
                                 const int* __restrict__ leg_rep, int base, int* __restrict__ rep_local,
                                 int* __restrict__ err_flag) {
  pdl_sync();
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= T) return;
  const int gn = node_ptr[g], N = node_ptr[g + 1] - gn;
  for (int i = 0; i < N; ++i) {
    const int b = body[gn + i];
    int rep = i;
    if (b > 0) {
      const int leg = b % base;
      const int target = b - leg + leg_rep[leg];
      if (target != b) {
        rep = -1;
        for (int j = 0; j < N; ++j)
          if (body[gn + j] == target) { rep = j; break; }
        if (rep < 0) { rep = i; *err_flag = 1; }
      }
    }
    rep_local[gn + i] = rep;
  }
}

extern "C" int sard_orbit_rep(const int* node_ptr, const int* body_ind, int T, const int* leg_rep, int index_base,
                              int* rep_local, int* err_flag, sard_stream_t stream) {
  if (T <= 0) return 0;
  SARD_REQUIRE(index_base >= 2, "index_base");
  sard_launch(orbit_rep_kernel, ceil_div(T, 128), 128, 0, as_stream(stream), node_ptr, body_ind, T, leg_rep, index_base,
                                                                   rep_local, err_flag);
  SARD_LAUNCH_OK("orbit_rep_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------------
// Run gather: one thread per 16 bytes of the output block x[n, ldx] (ldx is a multiple of 4).  The thread finds its graph
// by bisection over the run's node prefix `cum` (B+1 ints, L1-resident), then follows ids -> node_ptr -> obs once.  With
// ~150 k threads per launch every load of the launch is in flight at the same time: the kernel costs one dependent chain
// of DRAM latencies instead of one chain per graph and lane (one warp per graph: 20 us for 11.7 MB = 0.57 TB/s).
// Column program of a row: [obs[:lead] | obs[d_obs-tail:] | consts | onehot(stage)] and zeros up to ldx.
// ------------------------------------------------------------------------------------------
struct GatherP {
  const float* obs; int d_obs; const int* node_ptr; const int* stage; const int* body_ind;
  const int* ids; const int* cum; int B; int node_base; int n;
  int lead, tail, nconst, onehot; const float* consts;
  float* x; int ldx; int* nd_graph; int* nd_arena; int* body; int* root_row;
};

__global__ void __launch_bounds__(256) gather_kernel(const GatherP p) {
  pdl_sync();
  const int ld4 = p.ldx >> 2;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)p.n * ld4) return;
  const int row = (int)(t / ld4), c0 = ((int)(t - (long long)row * ld4)) << 2;
  // graph of this row: the last g with cum[g] - node_base <= row
  const int key = row + p.node_base;
  int lo = 0, hi = p.B - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(p.cum + mid) <= key) lo = mid; else hi = mid - 1;
  }
  const int g = lo;
  const int r0 = __ldg(p.cum + g) - p.node_base;
  const int gid = __ldg(p.ids + g);
  const int an = __ldg(p.node_ptr + gid) + (row - r0);
  const float* src = p.obs + (size_t)an * p.d_obs;
  const int c_tail = p.lead, c_const = p.lead + p.tail, c_hot = c_const + p.nconst;
  const int c_end = c_hot + (p.onehot ? 3 : 0);
  float4 v;
  const int tail_shift = p.d_obs - p.tail - p.lead;             // source column = c + tail_shift inside the tail region
  const bool al = ((p.d_obs & 3) == 0) && ((((uintptr_t)p.obs) & 15) == 0);
  if (al && c0 + 4 <= c_tail) {
    v = __ldg(reinterpret_cast<const float4*>(src + c0));
  } else if (al && c0 >= c_tail && c0 + 4 <= c_const && ((tail_shift & 3) == 0)) {
    v = __ldg(reinterpret_cast<const float4*>(src + c0 + tail_shift));
  } else {
    float e[4];
    int st = -1;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = c0 + u;
      float w = 0.f;
      if (c < c_tail) w = __ldg(src + c);
      else if (c < c_const) w = __ldg(src + c + tail_shift);
      else if (c < c_hot) w = __ldg(p.consts + (c - c_const));
      else if (c < c_end) {
        if (st < 0) st = __ldg(p.stage + gid);
        w = (c - c_hot) == st ? 1.f : 0.f;
      }
      e[u] = w;
    }
    v = make_float4(e[0], e[1], e[2], e[3]);
  }
  *reinterpret_cast<float4*>(p.x + (size_t)row * p.ldx + c0) = v;
  if (c0 == 0) {
    p.nd_graph[row] = g;
    p.nd_arena[row] = an;
    p.body[row] = __ldg(p.body_ind + an);
    if (row == r0 && p.root_row) p.root_row[g] = r0;
  }
}

extern "C" int sard_gather_nodes(const SardArena* arena, const SardSel* sel, const SardGatherOut* out,
                                 sard_stream_t stream) {
  if (sel->B <= 0) return 0;
  const int width = out->lead + out->tail + out->nconst + (out->onehot ? 3 : 0);
  SARD_REQUIRE(width <= out->ldx, "gather: ldx too small");
  SARD_REQUIRE(out->lead + out->tail <= arena->d_obs, "gather: column program exceeds d_obs");
  GatherP p;
  p.obs = arena->obs; p.d_obs = arena->d_obs; p.node_ptr = arena->node_ptr; p.stage = arena->stage;
  p.body_ind = arena->body_ind; p.ids = sel->ids; p.cum = sel->cum; p.B = sel->B; p.node_base = sel->node_base;
  p.lead = out->lead; p.tail = out->tail; p.nconst = out->nconst; p.onehot = out->onehot; p.consts = out->consts;
  p.x = out->x; p.ldx = out->ldx; p.nd_graph = out->nd_graph; p.nd_arena = out->nd_arena; p.body = out->body;
  p.root_row = out->root_row;
  p.n = sel->n;
  SARD_REQUIRE(sel->n > 0 && (out->ldx & 3) == 0 && ((uintptr_t)out->x & 15) == 0, "gather: x rows must be 16-byte aligned");
  const int threads = 256;
  // algorithmic bytes: obs columns read + x row written + three index arrays written per node
  sard_prof_bytes(4.0 * sel->n * ((double)(out->lead + out->tail) + out->ldx + 3.0));
  sard_launch(gather_kernel, ceil_div((long long)sel->n * (out->ldx / 4), threads), threads, 0, as_stream(stream), p);
  SARD_LAUNCH_OK("gather_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------------
// Stable counting sort by body index (3 kernels) + tile / chunk tables.
// ------------------------------------------------------------------------------------------
#define SORT_BLOCK 256

__global__ void sort_hist_kernel(const int* __restrict__ body, int n, int max_index, int* __restrict__ blockhist,
                                 int* __restrict__ counts) {
  pdl_sync();
  extern __shared__ int hist[];
  for (int u = threadIdx.x; u < max_index; u += blockDim.x) hist[u] = 0;
  __syncthreads();
  const int r = blockIdx.x * SORT_BLOCK + threadIdx.x;
  if (r < n) {
    int u = body[r];
    if (u < 0 || u >= max_index) { counts[2] = 1; u = min(max(u, 0), max_index - 1); }
    atomicAdd(&hist[u], 1);
  }
  __syncthreads();
  for (int u = threadIdx.x; u < max_index; u += blockDim.x) blockhist[(size_t)blockIdx.x * max_index + u] = hist[u];
}

// one block of 1024 threads, thread u owns body index u: per-index totals, then block-wide exclusive
// scans give the row / tile / chunk offsets and every thread writes its own tile and chunk entries
__global__ void __launch_bounds__(1024) sort_scan_kernel(int nb, int n, int max_index, int tile_rows, int chunk_rows, int pad,
                                 int* __restrict__ blockhist, int* __restrict__ cnt, int* __restrict__ grp_ptr,
                                 int4* __restrict__ tiles, int4* __restrict__ chunks, int* __restrict__ grp_chunk_ptr,
                                 int* __restrict__ counts) {
  pdl_sync();
  __shared__ int sc[3][1024];
  const int u = threadIdx.x;
  int total = 0;
  if (u < max_index) {
    for (int b = 0; b < nb; ++b) {
      const int t = blockhist[(size_t)b * max_index + u];
      blockhist[(size_t)b * max_index + u] = total;
      total += t;
    }
    cnt[u] = total;
  }
  const int my_t = (total + tile_rows - 1) / tile_rows, my_c = (total + chunk_rows - 1) / chunk_rows;
  // pad: every index starts on a tile boundary (sorted positions beyond its rows stay unused: srow = -1)
  const int span = pad ? my_t * tile_rows : total;
  sc[0][u] = span; sc[1][u] = my_t; sc[2][u] = my_c;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {          // inclusive Hillis-Steele scan of the three counters
    int v0 = 0, v1 = 0, v2 = 0;
    if (u >= o) { v0 = sc[0][u - o]; v1 = sc[1][u - o]; v2 = sc[2][u - o]; }
    __syncthreads();
    sc[0][u] += v0; sc[1][u] += v1; sc[2][u] += v2;
    __syncthreads();
  }
  if (u < max_index) {
    const int run = sc[0][u] - span, t0 = sc[1][u] - my_t, c0 = sc[2][u] - my_c;
    grp_ptr[u] = run;
    grp_chunk_ptr[u] = c0;
    for (int k = 0; k < my_t; ++k) tiles[t0 + k] = make_int4(run + k * tile_rows, run + min((k + 1) * tile_rows, total), u, 0);
    for (int k = 0; k < my_c; ++k) chunks[c0 + k] = make_int4(run + k * chunk_rows, run + min((k + 1) * chunk_rows, total), u, 0);
    if (u == max_index - 1) {
      grp_ptr[max_index] = sc[0][u];
      grp_chunk_ptr[max_index] = sc[2][u];
      counts[0] = sc[1][u];
      counts[1] = sc[2][u];
      counts[3] = sc[0][u];                      // sorted positions in use (== n without padding)
    }
  }
}

__global__ void sort_scatter_kernel(const int* __restrict__ body, int n, int max_index,
                                    const int* __restrict__ blockhist, const int* __restrict__ grp_ptr,
                                    int* __restrict__ srow, int* __restrict__ spos) {
  pdl_sync();
  extern __shared__ int warp_cnt[];   // [SORT_BLOCK/32][max_index]
  const int nwarp = SORT_BLOCK / 32;
  for (int i = threadIdx.x; i < nwarp * max_index; i += blockDim.x) warp_cnt[i] = 0;
  __syncthreads();
  const int r = blockIdx.x * SORT_BLOCK + threadIdx.x;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int u = -1;
  if (r < n) u = min(max(body[r], 0), max_index - 1);
  const unsigned peers = __match_any_sync(0xffffffffu, u);
  const int rank = __popc(peers & ((1u << lane) - 1u));
  if (u >= 0 && rank == 0) warp_cnt[w * max_index + u] = __popc(peers);
  __syncthreads();
  if (u >= 0) {
    int off = 0;
    for (int ww = 0; ww < w; ++ww) off += warp_cnt[ww * max_index + u];
    const int pos = grp_ptr[u] + blockhist[(size_t)blockIdx.x * max_index + u] + off + rank;
    srow[pos] = r;
    spos[r] = pos;
  }
}

extern "C" int sard_sort_by_index(const int* body, int n, int max_index, int tile_rows, int chunk_rows,
                                  int* srow, int* spos, int* grp_ptr, int* tiles, int* chunks, int* grp_chunk_ptr,
                                  int* counts, int* scratch, sard_stream_t stream) {
  return sard_sort_by_index_padded(body, n, max_index, tile_rows, chunk_rows, 0, srow, spos, grp_ptr, tiles, chunks, grp_chunk_ptr,
                                   counts, scratch, stream);
}

extern "C" int sard_sort_by_index_padded(const int* body, int n, int max_index, int tile_rows, int chunk_rows, int srow_capacity,
                                         int* srow, int* spos, int* grp_ptr, int* tiles, int* chunks, int* grp_chunk_ptr,
                                         int* counts, int* scratch, sard_stream_t stream) {
  SARD_REQUIRE(max_index > 0 && max_index <= 1024, "max_index");
  SARD_REQUIRE(tile_rows > 0 && chunk_rows > 0, "tile/chunk rows");
  const int pad = srow_capacity > 0;
  SARD_REQUIRE(!pad || chunk_rows % tile_rows == 0, "padded sort: chunk_rows must be a multiple of tile_rows");
  cudaStream_t st = as_stream(stream);
  SARD_CUDA(cudaMemsetAsync(counts, 0, 4 * sizeof(int), st));
  if (pad) SARD_CUDA(cudaMemsetAsync(srow, 0xff, sizeof(int) * (size_t)srow_capacity, st));
  const int nb = n > 0 ? ceil_div(n, SORT_BLOCK) : 0;
  int* blockhist = scratch;
  int* cnt = scratch + (size_t)nb * max_index;
  if (nb > 0) {
    sard_launch(sort_hist_kernel, nb, SORT_BLOCK, max_index * sizeof(int), st, body, n, max_index, blockhist, counts);
    SARD_LAUNCH_OK("sort_hist_kernel");
  }
  sard_launch(sort_scan_kernel, 1, 1024, 0, st, nb, n, max_index, tile_rows, chunk_rows, pad, blockhist, cnt, grp_ptr,
                                      reinterpret_cast<int4*>(tiles), reinterpret_cast<int4*>(chunks),
                                      grp_chunk_ptr, counts);
  SARD_LAUNCH_OK("sort_scan_kernel");
  if (nb > 0) {
    sard_launch(sort_scatter_kernel, nb, SORT_BLOCK, (SORT_BLOCK / 32) * max_index * sizeof(int), st, 
        body, n, max_index, blockhist, grp_ptr, srow, spos);
    SARD_LAUNCH_OK("sort_scatter_kernel");
  }
  return 0;
}
