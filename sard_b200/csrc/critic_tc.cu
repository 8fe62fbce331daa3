// Critic MLP forward on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), error-compensated 3xTF32, both layers in
// ONE launch.  Replaces, on the 2048 root rows of a minibatch (transform2act_critic.py:104-107; the reference runs the MLP
// on every node and keeps the roots), three dependent FFMA launches (64 -> 512 GEMM, split-K 512 -> 256 GEMM, finish):
//
//   h1 = act(x W1^T + b1)   [B, 512]      x = the tower output of each graph's root node (gathered by root_row)
//   h2 = act(h1 W2^T + b2)  [B, 256]      written into the first 256 columns of the value head's input `cat`
//
// Grid (B / 128 row tiles, 4 column slices of layer 2).  Layer 1 is walked in eight 64-column blocks; block j is, at the
// same time, K-block j of layer 2: its epilogue (bias, activation, hi/lo split) writes straight into tensor memory as the
// A operand of the layer-2 product, so h1 never makes an HBM round trip inside the kernel (it is stored once, by slice 0,
// for the backward pass).  Every CTA recomputes layer 1 (20 % of the flops) in exchange for 4 x the CTAs on layer 2.
//   warp 0: TMA producer -- the sixteen 32 KB weight blocks (hi | lo, pre-split per optimiser step) in consumption order
//   warp 1: MMA issuer   -- D1 = x W1_j^T (A: x in TMEM), D2 += h1_j W2_{s,j}^T (A: h1_j in TMEM), accumulators in TMEM
//   warps 2-9: one row per thread (two threads per row, 32 columns each): TMEM -> bias / act / split -> TMEM, stores
#include "tc.cuh"
#include "critic_tc.cuh"
#include <string.h>

#define CM_THREADS 320
#define CM_STAGES 4
#define CM_BLOCK_BYTES 32768         // one weight block: [64 rows][hi 64 | lo 64] fp32 as four 32-column SWIZZLE_128B chunks
#define CM_SCRATCH 32768             // eight warps x 4 KB transpose scratch for whole-row stores
#define CM_SMEM (CM_STAGES * CM_BLOCK_BYTES + CM_SCRATCH + 1024)
#define CM_IN 64
#define CM_H1 512
#define CM_H2 256
#define CM_NB (CM_H1 / 64)           // layer-1 column blocks = layer-2 K blocks
#define CM_SLICES (CM_H2 / 64)
#define CM_FWD_FLOATS ((size_t)CM_H1 * 128 + (size_t)CM_SLICES * CM_NB * 64 * 128)      // forward blocks; the W2^T blocks follow
#define CM_BWD1_OFF (CM_FWD_FLOATS + (size_t)CM_NB * CM_SLICES * 64 * 128)               // ... then the W1^T blocks

struct CmParams {
  CUtensorMap tm_w1, tm_w2;
  const float* hout; const int* root_row; int B;
  const float* b1; const float* b2; int act;
  float* H1;                          // [B, 512] or NULL
  float* cat; int ldcat;              // h2 -> cat[:, 0:256]
};

// ------------------------------------------------------------------------------------------ weights
// ws: W1 blocks [512][hi 64 | lo 64], then W2 blocks [(slice * 8 + kblock) * 64 + row][hi 64 | lo 64]
struct CmSplitParams { const float* W1; const float* W2; float* out; };
__global__ void __launch_bounds__(256) critic_presplit_kernel(const CmSplitParams p) {
  pdl_sync();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  constexpr int N1 = CM_H1 * CM_IN, N2 = CM_H2 * CM_H1;
  if (e < N1) {
    const int r = e >> 6, k = e & 63;
    const float v = p.W1[e], h = tc::tf32_hi(v);
    p.out[(size_t)r * 128 + k] = h;
    p.out[(size_t)r * 128 + 64 + k] = v - h;
    // backward-data operand of layer 1: block (r >> 6) row k holds W1[64 (r >> 6) + .][k], i.e. W1^T with K = the 512 outputs
    float* t = p.out + CM_BWD1_OFF + ((size_t)(r >> 6) * 64 + k) * 128;
    t[r & 63] = h;
    t[64 + (r & 63)] = v - h;
  } else if (e < N1 + N2) {
    const int q = e - N1, n = q / CM_H1, k = q - n * CM_H1;           // W2[n][k]
    const int s = n >> 6, rr = n & 63, j = k >> 6, kk = k & 63;
    const float v = p.W2[q], h = tc::tf32_hi(v);
    float* o = p.out + (size_t)CM_H1 * 128 + ((size_t)(s * CM_NB + j) * 64 + rr) * 128;
    o[kk] = h;
    o[64 + kk] = v - h;
    // backward-data operand of layer 2: block (j, s) row (k & 63) holds W2[64 s + .][k], i.e. W2^T with K = the 256 outputs
    float* t = p.out + CM_FWD_FLOATS + ((size_t)(j * CM_SLICES + s) * 64 + kk) * 128;
    t[rr] = h;
    t[64 + rr] = v - h;
  }
}

long long sard_critic_tc_wsplit_floats(void) { return (long long)CM_BWD1_OFF + (long long)CM_NB * 64 * 128; }
bool sard_critic_tc_supported(int in0, int h1, int h2, int n_mlp) { return n_mlp == 2 && in0 == CM_IN && h1 == CM_H1 && h2 == CM_H2; }

int sard_critic_tc_presplit(const float* W1, const float* W2, float* wsplit, sard_stream_t stream) {
  CmSplitParams p;
  p.W1 = W1; p.W2 = W2; p.out = wsplit;
  sard_launch(critic_presplit_kernel, ceil_div(CM_H1 * CM_IN + CM_H2 * CM_H1, 256), 256, 0, as_stream(stream), p);
  SARD_LAUNCH_OK("critic_presplit_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------------ forward
namespace {
__device__ __forceinline__ void cm_act32(float* v, const float* bias, int act) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 b4 = *reinterpret_cast<const float4*>(bias + 4 * j);
    v[4 * j] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
  }
  if (act == SARD_ACT_TANH) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = tc::tanh_fast(v[j]);
  } else if (act == SARD_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  } else if (act == SARD_ACT_SIGMOID) {
#pragma unroll 4
    for (int j = 0; j < 32; ++j) v[j] = 1.f / (1.f + expf(-v[j]));
  }
}
// the warp's 32 rows x 32 columns (lane = row) -> global rows of `ld` floats, whole 128-byte row pieces per instruction
__device__ __forceinline__ void cm_store_block(uint8_t* scr, int lane, const float* v, float* dst0, size_t ld, int nv) {
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(scr + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  __syncwarp();
  const int pc = lane & 7;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int row = 4 * k + (lane >> 3);
    if (row < nv)
      *reinterpret_cast<float4*>(dst0 + (size_t)row * ld + pc * 4) = *reinterpret_cast<const float4*>(scr + row * 128 + ((pc ^ (row & 7)) << 4));
  }
  __syncwarp();
}
__device__ __forceinline__ void cm_put32(uint32_t t_hi, uint32_t t_lo, const float* v) {
#pragma unroll
  for (int j = 0; j < 4; ++j) tc::tmem_st8_split(t_hi + 8 * j, t_lo + 8 * j, v + 8 * j);
  tc::tmem_st_wait();
}
}  // namespace

__global__ void __launch_bounds__(CM_THREADS, 1) critic_mlp_fwd_tc_kernel(const __grid_constant__ CmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = tc::align1024(smem_raw);
  uint8_t* scratch = ring + CM_STAGES * CM_BLOCK_BYTES;
  __shared__ uint64_t bar_full[CM_STAGES], bar_empty[CM_STAGES], bar_a1, bar_d1full, bar_d1free, bar_a2full, bar_a2free, bar_d2full;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float s_b1[CM_H1], s_b2[64];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int slice = blockIdx.y;
  if (warp == 1) tc::tmem_alloc(&tmem_slot, 512);
  if (tid == 0) {
    for (int s = 0; s < CM_STAGES; ++s) { tc::mbar_init(&bar_full[s], 1); tc::mbar_init(&bar_empty[s], 1); }
    tc::mbar_init(&bar_a1, 256); tc::mbar_init(&bar_d1full, 1); tc::mbar_init(&bar_d1free, 256);
    tc::mbar_init(&bar_a2full, 256); tc::mbar_init(&bar_a2free, 1); tc::mbar_init(&bar_d2full, 1);
    tc::mbar_fence_init();
    tc::tma_prefetch_desc(&p.tm_w1); tc::tma_prefetch_desc(&p.tm_w2);
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t A1H = tmem, A1L = tmem + 64, D1 = tmem + 128, A2H = tmem + 192, A2L = tmem + 256, D2 = tmem + 320;
  pdl_sync();

  if (warp == 0) {
    if (lane == 0) {
      // consumption order: W1_0, then for j: [W1_{j+1}], W2_{slice,j}
      for (int it = 0; it < 2 * CM_NB; ++it) {
        const int s = it % CM_STAGES;
        tc::mbar_wait(&bar_empty[s], ((it / CM_STAGES) & 1) ^ 1);
        uint8_t* st = ring + s * CM_BLOCK_BYTES;
        // it = 0: W1_0; odd it = 2j+1 (j < 7): W1_{j+1}; even it = 2j+2: W2_j; the last (it = 15): W2_7
        bool is_w1; int blk;
        if (it == 0) { is_w1 = true; blk = 0; }
        else if (it == 2 * CM_NB - 1) { is_w1 = false; blk = CM_NB - 1; }
        else if (it & 1) { is_w1 = true; blk = (it + 1) >> 1; }
        else { is_w1 = false; blk = (it >> 1) - 1; }
        tc::mbar_expect_tx(&bar_full[s], CM_BLOCK_BYTES);
        const CUtensorMap* tm = is_w1 ? &p.tm_w1 : &p.tm_w2;
        const int row0 = is_w1 ? blk * 64 : (slice * CM_NB + blk) * 64;
        for (int c = 0; c < 4; ++c) tc::tma_load_2d(st + c * 8192, tm, c * 32, row0, &bar_full[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t ID64 = tc::idesc_tf32(128, 64, 0, 0);
      int it = 0;
      auto mma_block = [&](uint32_t D, uint32_t AH, uint32_t AL, bool first) {
        const int s = it % CM_STAGES;
        tc::mbar_wait(&bar_full[s], (it / CM_STAGES) & 1);
        tc::tc_fence_after();
        const uint32_t sw = tc::smem_u32(ring + s * CM_BLOCK_BYTES);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t off = (uint32_t)(ks >> 2) * 8192 + (uint32_t)(ks & 3) * 32;
          const uint64_t b_hi = tc::kdesc(sw + off), b_lo = tc::kdesc(sw + 16384 + off);
          tc::mma_ts(D, AH + ks * 8, b_hi, ID64, (first && ks == 0) ? 0u : 1u);
          tc::mma_ts(D, AL + ks * 8, b_hi, ID64, 1u);
          tc::mma_ts(D, AH + ks * 8, b_lo, ID64, 1u);
        }
        tc::mma_commit(&bar_empty[s]);
        ++it;
      };
      tc::mbar_wait(&bar_a1, 0);
      mma_block(D1, A1H, A1L, true);                         // layer 1, block 0
      tc::mma_commit(&bar_d1full);
      for (int j = 0; j < CM_NB; ++j) {
        if (j + 1 < CM_NB) {
          tc::mbar_wait(&bar_d1free, j & 1);                  // the row workers have read block j out of D1
          mma_block(D1, A1H, A1L, true);                      // layer 1, block j + 1 (runs under block j's epilogue)
          tc::mma_commit(&bar_d1full);
        }
        tc::mbar_wait(&bar_a2full, j & 1);                    // act(block j) sits in tensor memory as hi | lo
        mma_block(D2, A2H, A2L, j == 0);                      // layer 2, K block j
        tc::mma_commit(&bar_a2free);
      }
      tc::mma_commit(&bar_d2full);
    }
  } else {
    const int q = warp & 3, half = (warp - 2) >> 2, c0 = half * 32;
    const int trow = q * 32 + lane;
    const int b = blockIdx.x * 128 + trow;
    const bool valid = b < p.B;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const int row0 = blockIdx.x * 128 + q * 32, nv = min(max(p.B - row0, 0), 32);
    uint8_t* scr = scratch + (warp - 2) * 4096;
    // biases to shared memory (read back as broadcast float4)
    {
      const int wt = tid - 64;
      s_b1[wt] = __ldg(p.b1 + wt); s_b1[256 + wt] = __ldg(p.b1 + 256 + wt);
      if (wt < 64) s_b2[wt] = __ldg(p.b2 + slice * 64 + wt);
    }
    float v[32];
    // ---- x: the tower output of this graph's root node
    {
      const float* xr = valid ? p.hout + (size_t)__ldg(p.root_row + b) * CM_IN + c0 : nullptr;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 t = valid ? __ldg(reinterpret_cast<const float4*>(xr + 4 * j)) : make_float4(0.f, 0.f, 0.f, 0.f);
        v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
      }
      cm_put32(A1H + lane_off + c0, A1L + lane_off + c0, v);
      tc::tc_fence_before();
      tc::mbar_arrive(&bar_a1);
    }
    tc::named_sync(1, 256);                                   // the biases are in shared memory
#pragma unroll 1
    for (int j = 0; j < CM_NB; ++j) {
      tc::mbar_wait(&bar_d1full, j & 1);
      tc::tc_fence_after();
      tc::tmem_ld32(D1 + lane_off + c0, v);
      tc::tc_fence_before();
      tc::mbar_arrive(&bar_d1free);
      cm_act32(v, s_b1 + j * 64 + c0, p.act);
      if (j > 0) { tc::mbar_wait(&bar_a2free, (j - 1) & 1); tc::tc_fence_after(); }     // layer 2 has consumed block j - 1
      cm_put32(A2H + lane_off + c0, A2L + lane_off + c0, v);
      tc::tc_fence_before();
      tc::mbar_arrive(&bar_a2full);
      if (p.H1 && slice == 0) cm_store_block(scr, lane, v, p.H1 + (size_t)row0 * CM_H1 + j * 64 + c0, CM_H1, nv);   // under the MMAs
    }
    tc::mbar_wait(&bar_d2full, 0);
    tc::tc_fence_after();
    tc::tmem_ld32(D2 + lane_off + c0, v);
    cm_act32(v, s_b2 + c0, p.act);
    cm_store_block(scr, lane, v, p.cat + (size_t)row0 * p.ldcat + slice * 64 + c0, (size_t)p.ldcat, nv);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem, 512);
}

int sard_critic_tc_forward(const float* wsplit, const float* hout, const int* root_row, int B, const float* b1, const float* b2, int act,
                           float* H1, float* cat, int ldcat, sard_stream_t stream) {
  if (B <= 0) return 0;
  SARD_REQUIRE((ldcat & 3) == 0 && (reinterpret_cast<uintptr_t>(cat) & 15) == 0 && (!H1 || (reinterpret_cast<uintptr_t>(H1) & 15) == 0),
               "critic_tc: cat / H1 alignment");
  cudaStream_t st = as_stream(stream);
  CmParams p;
  memset(&p, 0, sizeof(p));
  SARD_TRY(sard_tmap_2d(&p.tm_w1, wsplit, CM_H1, 128, 128, 64, 32));
  SARD_TRY(sard_tmap_2d(&p.tm_w2, wsplit + (size_t)CM_H1 * 128, (long long)CM_SLICES * CM_NB * 64, 128, 128, 64, 32));
  p.hout = hout; p.root_row = root_row; p.B = B; p.b1 = b1; p.b2 = b2; p.act = act; p.H1 = H1; p.cat = cat; p.ldcat = ldcat;
  static bool configured[SARD_MAX_DEVICES] = {};
  if (sard_first_use_on_device(configured))
    SARD_CUDA(cudaFuncSetAttribute(critic_mlp_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CM_SMEM));
  ProfScope prof(st, 2.0 * B * ((double)CM_IN * CM_H1 + (double)CM_H1 * CM_H2), 4.0 * B * ((double)CM_IN + CM_H1 + CM_H2));
  sard_launch(critic_mlp_fwd_tc_kernel, dim3(ceil_div(B, 128), CM_SLICES), dim3(CM_THREADS), (size_t)CM_SMEM, st, p);
  SARD_LAUNCH_OK("critic_mlp_fwd_tc_kernel");
  return 0;
}


// ------------------------------------------------------------------------------------------ backward-data
// One kernel, two uses (a product  out[:, 64 jb ..] = (A W^T-block) * act'(Y)  with A read in 64-wide K chunks):
//   layer 2:  dz1 = (dz2 W2) * act'(h1)            [B, 512]  K = 256 (4 chunks), 8 column blocks, rows in place
//   layer 1:  dz_tower[root] = (dz1 W1) * act'(hout[root])   [B, 64]   K = 512 (8 chunks), 1 column block, rows scattered
// Grid (B / 128 row tiles, column blocks).  The row workers read their row's chunk of A, split it and put it into one of
// two A buffers in tensor memory while the MMAs of the previous chunk run; the W^T blocks arrive through a TMA ring.
// Replaces two split-K FFMA GEMMs + their finishing kernels on the critic's chain.
struct CbParams {
  CUtensorMap tm_wt;
  const float* A; int lda;            // [B, 64 nk]
  const float* Y; int ldy;            // activations whose derivative multiplies the result
  float* out; int ldo;
  const int* rows;                    // NULL: row b of Y / out; else row rows[b] (root rows of the tower)
  int B, act, nk;
};

__global__ void __launch_bounds__(CM_THREADS, 1) critic_mlp_bwd_tc_kernel(const __grid_constant__ CbParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = tc::align1024(smem_raw);
  uint8_t* scratch = ring + CM_STAGES * CM_BLOCK_BYTES;
  __shared__ uint64_t bar_full[CM_STAGES], bar_empty[CM_STAGES], bar_afull[2], bar_afree[2], bar_dfull;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int jb = blockIdx.y;
  if (warp == 1) tc::tmem_alloc(&tmem_slot, 512);
  if (tid == 0) {
    for (int s = 0; s < CM_STAGES; ++s) { tc::mbar_init(&bar_full[s], 1); tc::mbar_init(&bar_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { tc::mbar_init(&bar_afull[s], 256); tc::mbar_init(&bar_afree[s], 1); }
    tc::mbar_init(&bar_dfull, 1);
    tc::mbar_fence_init();
    tc::tma_prefetch_desc(&p.tm_wt);
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t D = tmem + 256;
  pdl_sync();

  if (warp == 0) {
    if (lane == 0) {
      for (int kc = 0; kc < p.nk; ++kc) {
        const int s = kc % CM_STAGES;
        tc::mbar_wait(&bar_empty[s], ((kc / CM_STAGES) & 1) ^ 1);
        uint8_t* st = ring + s * CM_BLOCK_BYTES;
        tc::mbar_expect_tx(&bar_full[s], CM_BLOCK_BYTES);
        for (int c = 0; c < 4; ++c) tc::tma_load_2d(st + c * 8192, &p.tm_wt, c * 32, (jb * p.nk + kc) * 64, &bar_full[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t ID64 = tc::idesc_tf32(128, 64, 0, 0);
      for (int kc = 0; kc < p.nk; ++kc) {
        const int buf = kc & 1, s = kc % CM_STAGES;
        const uint32_t AH = tmem + buf * 128, AL = AH + 64;
        tc::mbar_wait(&bar_full[s], (kc / CM_STAGES) & 1);
        tc::mbar_wait(&bar_afull[buf], (kc >> 1) & 1);
        tc::tc_fence_after();
        const uint32_t sw = tc::smem_u32(ring + s * CM_BLOCK_BYTES);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t off = (uint32_t)(ks >> 2) * 8192 + (uint32_t)(ks & 3) * 32;
          const uint64_t b_hi = tc::kdesc(sw + off), b_lo = tc::kdesc(sw + 16384 + off);
          tc::mma_ts(D, AH + ks * 8, b_hi, ID64, (kc | ks) ? 1u : 0u);
          tc::mma_ts(D, AL + ks * 8, b_hi, ID64, 1u);
          tc::mma_ts(D, AH + ks * 8, b_lo, ID64, 1u);
        }
        tc::mma_commit(&bar_empty[s]);
        tc::mma_commit(&bar_afree[buf]);
      }
      tc::mma_commit(&bar_dfull);
    }
  } else {
    const int q = warp & 3, half = (warp - 2) >> 2, c0 = half * 32;
    const int trow = q * 32 + lane;
    const int b = blockIdx.x * 128 + trow;
    const bool valid = b < p.B;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const int row0 = blockIdx.x * 128 + q * 32, nv = min(max(p.B - row0, 0), 32);
    uint8_t* scr = scratch + (warp - 2) * 4096;
    float v[32];
    float4 nx[8];                                            // the next chunk of this row, in flight under the current one
    const float* arow = p.A + (size_t)(valid ? b : 0) * p.lda + c0;
#pragma unroll
    for (int j = 0; j < 8; ++j) nx[j] = valid ? *reinterpret_cast<const float4*>(arow + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
    for (int kc = 0; kc < p.nk; ++kc) {
      const int buf = kc & 1;
#pragma unroll
      for (int j = 0; j < 8; ++j) { v[4 * j] = nx[j].x; v[4 * j + 1] = nx[j].y; v[4 * j + 2] = nx[j].z; v[4 * j + 3] = nx[j].w; }
      if (kc + 1 < p.nk) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          nx[j] = valid ? *reinterpret_cast<const float4*>(arow + (kc + 1) * 64 + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (kc >= 2) { tc::mbar_wait(&bar_afree[buf], ((kc >> 1) - 1) & 1); tc::tc_fence_after(); }   // chunk kc - 2 has been consumed
      cm_put32(tmem + buf * 128 + lane_off + c0, tmem + buf * 128 + 64 + lane_off + c0, v);
      tc::tc_fence_before();
      tc::mbar_arrive(&bar_afull[buf]);
    }
    // the activations whose derivative scales this column block, loaded under the MMAs
    const long long yrow = valid ? (p.rows ? (long long)__ldg(p.rows + b) : (long long)b) : 0;
    float y[32];
    {
      const float* hr = p.Y + (size_t)yrow * p.ldy + jb * 64 + c0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 t = valid ? *reinterpret_cast<const float4*>(hr + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
        y[4 * j] = t.x; y[4 * j + 1] = t.y; y[4 * j + 2] = t.z; y[4 * j + 3] = t.w;
      }
    }
    tc::mbar_wait(&bar_dfull, 0);
    tc::tc_fence_after();
    tc::tmem_ld32(D + lane_off + c0, v);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= act_grad_from_output(y[j], p.act);
    if (p.rows) {                                            // scattered rows: every thread writes its own 128 bytes
      if (valid) {
        float* o = p.out + (size_t)yrow * p.ldo + jb * 64 + c0;
#pragma unroll
        for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(o + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
    } else {
      cm_store_block(scr, lane, v, p.out + (size_t)row0 * p.ldo + jb * 64 + c0, (size_t)p.ldo, nv);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem, 512);
}

static int launch_bwd(const CbParams& p, int col_blocks, double flops, double bytes, const char* what, cudaStream_t st) {
  static bool configured[SARD_MAX_DEVICES] = {};
  if (sard_first_use_on_device(configured))
    SARD_CUDA(cudaFuncSetAttribute(critic_mlp_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CM_SMEM));
  ProfScope prof(st, flops, bytes);
  sard_launch(critic_mlp_bwd_tc_kernel, dim3(ceil_div(p.B, 128), col_blocks), dim3(CM_THREADS), (size_t)CM_SMEM, st, p);
  SARD_LAUNCH_OK(what);
  return 0;
}

int sard_critic_tc_bwd_layer2(const float* wsplit, const float* dz2, int ld_dz2, const float* H1, int B, int act, float* dZ1,
                              sard_stream_t stream) {
  if (B <= 0) return 0;
  SARD_REQUIRE((ld_dz2 & 3) == 0 && ((reinterpret_cast<uintptr_t>(dz2) | reinterpret_cast<uintptr_t>(H1) | reinterpret_cast<uintptr_t>(dZ1)) & 15) == 0,
               "critic_tc backward: alignment");
  CbParams p;
  memset(&p, 0, sizeof(p));
  SARD_TRY(sard_tmap_2d(&p.tm_wt, wsplit + CM_FWD_FLOATS, (long long)CM_NB * CM_SLICES * 64, 128, 128, 64, 32));
  p.A = dz2; p.lda = ld_dz2; p.Y = H1; p.ldy = CM_H1; p.out = dZ1; p.ldo = CM_H1; p.rows = nullptr; p.B = B; p.act = act; p.nk = CM_SLICES;
  return launch_bwd(p, CM_NB, 2.0 * B * (double)CM_H1 * CM_H2, 4.0 * B * ((double)CM_H2 + 2.0 * CM_H1), "critic_mlp_bwd_tc_kernel<layer2>",
                    as_stream(stream));
}

int sard_critic_tc_bwd_layer1(const float* wsplit, const float* dZ1, const float* hout, const int* root_row, int B, int act, float* dz_tower,
                              sard_stream_t stream) {
  if (B <= 0) return 0;
  SARD_REQUIRE(((reinterpret_cast<uintptr_t>(dZ1) | reinterpret_cast<uintptr_t>(hout) | reinterpret_cast<uintptr_t>(dz_tower)) & 15) == 0,
               "critic_tc backward: alignment");
  CbParams p;
  memset(&p, 0, sizeof(p));
  SARD_TRY(sard_tmap_2d(&p.tm_wt, wsplit + CM_BWD1_OFF, (long long)CM_NB * 64, 128, 128, 64, 32));
  p.A = dZ1; p.lda = CM_H1; p.Y = hout; p.ldy = CM_IN; p.out = dz_tower; p.ldo = CM_IN; p.rows = root_row; p.B = B; p.act = act; p.nk = CM_NB;
  return launch_bwd(p, 1, 2.0 * B * (double)CM_H1 * CM_IN, 4.0 * B * ((double)CM_H1 + 2.0 * CM_IN), "critic_mlp_bwd_tc_kernel<layer1>",
                    as_stream(stream));
}
