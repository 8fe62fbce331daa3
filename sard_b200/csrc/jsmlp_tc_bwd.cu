// Backward of the fused JSMLP on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), 3xTF32.
//
// For one stage of Transform2ActPolicy this is the autograd of JSMLP.forward (design_opt/models/jsmlp.py:32-40, 74-79)
//     out = IL3(tanh(IL2(tanh(IL1(x, ind)), ind)), ind)
// in TWO launches (+ one ordered reduction):
//   jsmlp_bwd_tc_kernel    dZ2 = (dOut W3[u]) * act'(H2);  dZ1 = (dZ2 W2[u]) * act'(H1);  dX = dZ1 W1[u]
//                          one CTA per 128-row tile of one body index u; same structure as the forward kernel: a TMA
//                          producer warp, one MMA-issuing thread, 16 epilogue warps; dZ2 / dZ1 go from the epilogue
//                          straight back into TENSOR MEMORY as the next GEMM's A operand, and out to HBM once as hi|lo
//                          pair arrays (TMA stores) for the weight gradients.
//   jsmlp_wgrad_tc_kernel  dW1[u] = dZ1^T Xs, dW2[u] = dZ2^T H1, dW3[u] = dOut^T H2 and the three bias gradients for one
//                          chunk of rows of one body index: the reduction runs over ROWS, so both operands are read
//                          MN-major from the same [row][feature] pair arrays every other kernel uses (3-D "slab" TMA
//                          boxes -> canonical SWIZZLE_128B MN-major tiles); all six accumulators (480 columns) stay in
//                          tensor memory for the whole chunk; bias gradients are products with a tile of ones.
//   jsmlp_wgrad_reduce_kernel  adds the per-chunk partial products of every body index in chunk order (deterministic)
//                          into the compact gradient tables.
// Rows are in the PADDED body-index order of sard_sort_by_index_padded: every tile is whole, padding rows carry
// dOut = 0 and therefore dZ = 0, so they add nothing to any gradient.
#include "tc.cuh"
#include <string.h>

#define JB_EPI_WARPS 16
#define JB_THREADS (64 + 32 * JB_EPI_WARPS)
#define JB_STAGES 3
#define JB_STAGE_BYTES 65536        // [A_hi | A_lo | W_hi | W_lo], 16 KB each (128 rows x 32 fp32, SWIZZLE_128B)
#define JB_SUB 16384
#define JB_HID 128
#define JB_IN 256
#define JB_SMEM (JB_STAGES * JB_STAGE_BYTES + 32768 + 1024)

// ------------------------------------------------------------------------------------------ operand preparation
// dOs_hl[pos][0:32] = hi, [32:64] = lo of the output gradient of the row at sorted position pos (0 for padding rows and
// for columns >= d)
__global__ void __launch_bounds__(256) dout_sorted_split_kernel(const float* __restrict__ dout, int ld, int d,
                                                                const int* __restrict__ srow, int rows_cap,
                                                                float* __restrict__ dst) {
  pdl_sync();
  const int total = rows_cap * 8;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int pos = idx >> 3, c4 = idx & 7;
    const int row = __ldg(srow + pos);
    float v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = c4 * 4 + i;
      v[i] = (row >= 0 && c < d) ? __ldg(dout + (size_t)row * ld + c) : 0.f;
    }
    float4 h, l;
    tc::split4(make_float4(v[0], v[1], v[2], v[3]), h, l);
    *reinterpret_cast<float4*>(dst + (size_t)pos * 64 + c4 * 4) = h;
    *reinterpret_cast<float4*>(dst + (size_t)pos * 64 + 32 + c4 * 4) = l;
  }
}

// xs_hl[pos] = [hi(h[row] | ext[graph(row)]) | lo(...)] for the row at sorted position pos, zeros at padding positions:
// the repeat_interleave + cat of transform2act_policy.py:204-213 written directly in the operand format of the fused
// kernels (sorted, padded, split)
__global__ void __launch_bounds__(256) concat_sorted_split_kernel(const float* __restrict__ h, int ld_h, int Fh4,
                                                                  const float* __restrict__ ext, int ld_ext, int Fe4,
                                                                  const int* __restrict__ srow, const int* __restrict__ nd_graph,
                                                                  int rows_cap, float* __restrict__ dst) {
  pdl_sync();
  const int W4 = Fh4 + Fe4;
  const long long total = (long long)rows_cap * W4;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int pos = (int)(idx / W4), c4 = (int)(idx - (long long)pos * W4);
    const int row = __ldg(srow + pos);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row >= 0) {
      if (c4 < Fh4) v = *reinterpret_cast<const float4*>(h + (size_t)row * ld_h + c4 * 4);
      else v = *reinterpret_cast<const float4*>(ext + (size_t)__ldg(nd_graph + row) * ld_ext + (c4 - Fh4) * 4);
    }
    float4 hi, lo;
    tc::split4(v, hi, lo);
    float* d = dst + (size_t)pos * 8 * W4;
    *reinterpret_cast<float4*>(d + c4 * 4) = hi;
    *reinterpret_cast<float4*>(d + 4 * W4 + c4 * 4) = lo;
  }
}

extern "C" int sard_concat_sorted_split(const float* h, int ld_h, int Fh, const float* ext, int ld_ext, int Fe, const int* srow,
                                        const int* nd_graph, int rows_cap, float* xs_hl, sard_stream_t stream) {
  if (rows_cap <= 0) return 0;
  SARD_REQUIRE(Fh % 4 == 0 && Fe % 4 == 0 && ld_h % 4 == 0 && ld_ext % 4 == 0, "concat_sorted_split: multiples of 4");
  sard_prof_bytes(4.0 * rows_cap * (3.0 * (Fh + Fe) + 1.0));      // one row read, hi | lo pair row written
  sard_launch(concat_sorted_split_kernel, grid_for((long long)rows_cap * ((Fh + Fe) / 4), 256, 16), 256, 0, as_stream(stream), h, ld_h,
              Fh / 4, ext, ld_ext, Fe / 4, srow, nd_graph, rows_cap, xs_hl);
  SARD_LAUNCH_OK("concat_sorted_split_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------------ backward data
struct JsBwdParams {
  CUtensorMap tm_do, tm_w3t, tm_w2t, tm_w1t;              // loads (128-row x 32-column boxes)
  CUtensorMap tm_dz2, tm_dz1, tm_dx;                      // stores (32 x 32 boxes)
  const int4* tiles; const int* num_tiles; const int* slot_of_index;
  const float* H1hl; const float* H2hl;                   // [rows_cap, 256] hi | lo
  int out_dim, act;
};

__global__ void __launch_bounds__(JB_THREADS, 1) jsmlp_bwd_tc_kernel(const __grid_constant__ JsBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = tc::align1024(smem_raw);
  __shared__ uint64_t bar_full[JB_STAGES], bar_empty[JB_STAGES], bar_d[3], bar_a[2][4];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 1) tc::tmem_alloc(&tmem_slot, 512);
  if (tid == 0) {
    for (int s = 0; s < JB_STAGES; ++s) { tc::mbar_init(&bar_full[s], 1); tc::mbar_init(&bar_empty[s], 1); }
    for (int l = 0; l < 3; ++l) tc::mbar_init(&bar_d[l], 1);
    for (int l = 0; l < 2; ++l)
      for (int c = 0; c < 4; ++c) tc::mbar_init(&bar_a[l][c], 32 * JB_EPI_WARPS);
    tc::mbar_fence_init();
    tc::tma_prefetch_desc(&p.tm_do); tc::tma_prefetch_desc(&p.tm_w3t); tc::tma_prefetch_desc(&p.tm_w2t); tc::tma_prefetch_desc(&p.tm_w1t);
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  pdl_sync();

  const bool active = (int)blockIdx.x < *p.num_tiles;
  if (active) {
    const int4 t = p.tiles[blockIdx.x];
    const int row_start = t.x, u = t.z;
    const int slot = p.slot_of_index[u];
    const int nks3 = (p.out_dim + 7) >> 3;               // 8-wide k-steps of the first GEMM (K = out_dim, zero padded)
    constexpr uint32_t ID128 = tc::idesc_tf32(128, 128, 0, 0);
    const uint32_t D1 = tmem, AH = tmem + 128, AL = tmem + 256, D2 = tmem + 384;

    if (warp == 0) {
      if (lane == 0) {
        int it = 0;
        {                                       // dOut tile + W3^T slab
          const int s = 0;
          tc::mbar_wait(&bar_empty[s], 1);
          uint8_t* st = base;
          tc::mbar_expect_tx(&bar_full[s], 4 * JB_SUB);
          tc::tma_load_2d(st, &p.tm_do, 0, row_start, &bar_full[s]);
          tc::tma_load_2d(st + JB_SUB, &p.tm_do, 32, row_start, &bar_full[s]);
          tc::tma_load_2d(st + 2 * JB_SUB, &p.tm_w3t, 0, slot * JB_HID, &bar_full[s]);
          tc::tma_load_2d(st + 3 * JB_SUB, &p.tm_w3t, 32, slot * JB_HID, &bar_full[s]);
          ++it;
        }
        for (int kc = 0; kc < 4; ++kc, ++it) {  // W2^T k-chunks
          const int s = it % JB_STAGES;
          tc::mbar_wait(&bar_empty[s], ((it / JB_STAGES) & 1) ^ 1);
          uint8_t* st = base + s * JB_STAGE_BYTES;
          tc::mbar_expect_tx(&bar_full[s], 2 * JB_SUB);
          tc::tma_load_2d(st + 2 * JB_SUB, &p.tm_w2t, kc * 32, slot * JB_HID, &bar_full[s]);
          tc::tma_load_2d(st + 3 * JB_SUB, &p.tm_w2t, JB_HID + kc * 32, slot * JB_HID, &bar_full[s]);
        }
        for (int half = 0; half < 2; ++half)    // W1^T: the 256 input features as two 128-row halves
          for (int kc = 0; kc < 4; ++kc, ++it) {
            const int s = it % JB_STAGES;
            tc::mbar_wait(&bar_empty[s], ((it / JB_STAGES) & 1) ^ 1);
            uint8_t* st = base + s * JB_STAGE_BYTES;
            tc::mbar_expect_tx(&bar_full[s], 2 * JB_SUB);
            tc::tma_load_2d(st + 2 * JB_SUB, &p.tm_w1t, kc * 32, slot * JB_IN + half * 128, &bar_full[s]);
            tc::tma_load_2d(st + 3 * JB_SUB, &p.tm_w1t, JB_HID + kc * 32, slot * JB_IN + half * 128, &bar_full[s]);
          }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        int it = 0;
        {   // ---- dOut W3: A and B from shared memory
          tc::mbar_wait(&bar_full[0], 0);
          tc::tc_fence_after();
          const uint32_t sa = tc::smem_u32(base);
          for (int k = 0; k < nks3; ++k) {
            const uint64_t a_hi = tc::kdesc(sa + k * 32), a_lo = tc::kdesc(sa + JB_SUB + k * 32);
            const uint64_t b_hi = tc::kdesc(sa + 2 * JB_SUB + k * 32), b_lo = tc::kdesc(sa + 3 * JB_SUB + k * 32);
            tc::mma_ss(D1, a_hi, b_hi, ID128, k ? 1u : 0u);
            tc::mma_ss(D1, a_lo, b_hi, ID128, 1u);
            tc::mma_ss(D1, a_hi, b_lo, ID128, 1u);
          }
          tc::mma_commit(&bar_empty[0]);
          tc::mma_commit(&bar_d[0]);
          ++it;
        }
        // ---- dZ2 W2: A = dZ2 from tensor memory, chunk by chunk as the epilogue produces it
        for (int kc = 0; kc < 4; ++kc, ++it) {
          const int s = it % JB_STAGES;
          tc::mbar_wait(&bar_a[0][kc], 0);
          tc::mbar_wait(&bar_full[s], (it / JB_STAGES) & 1);
          tc::tc_fence_after();
          const uint32_t sa = tc::smem_u32(base + s * JB_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t b_hi = tc::kdesc(sa + 2 * JB_SUB + k * 32), b_lo = tc::kdesc(sa + 3 * JB_SUB + k * 32);
            const uint32_t col = kc * 32 + k * 8;
            tc::mma_ts(D2, AH + col, b_hi, ID128, (kc | k) ? 1u : 0u);
            tc::mma_ts(D2, AL + col, b_hi, ID128, 1u);
            tc::mma_ts(D2, AH + col, b_lo, ID128, 1u);
          }
          tc::mma_commit(&bar_empty[s]);
        }
        tc::mma_commit(&bar_d[1]);
        // ---- dZ1 W1: columns 0..127 of dX into D1 (trailing the epilogue), then 128..255 into D2 once the epilogue
        //      has read all of it
        for (int half = 0; half < 2; ++half) {
          const uint32_t D = half == 0 ? D1 : D2;
          for (int kc = 0; kc < 4; ++kc, ++it) {
            const int s = it % JB_STAGES;
            tc::mbar_wait(&bar_a[1][half == 0 ? kc : 3], 0);
            tc::mbar_wait(&bar_full[s], (it / JB_STAGES) & 1);
            tc::tc_fence_after();
            const uint32_t sa = tc::smem_u32(base + s * JB_STAGE_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t b_hi = tc::kdesc(sa + 2 * JB_SUB + k * 32), b_lo = tc::kdesc(sa + 3 * JB_SUB + k * 32);
              const uint32_t col = kc * 32 + k * 8;
              tc::mma_ts(D, AH + col, b_hi, ID128, (kc | k) ? 1u : 0u);
              tc::mma_ts(D, AL + col, b_hi, ID128, 1u);
              tc::mma_ts(D, AH + col, b_lo, ID128, 1u);
            }
            tc::mma_commit(&bar_empty[s]);
          }
        }
        tc::mma_commit(&bar_d[2]);
      }
    } else {
      // ---- epilogue: thread <-> tile row <-> TMEM lane; the four warps of a lane quarter share every 32-column chunk
      const int q = warp & 3;
      const int sub = (warp - 2) >> 2;                   // 0..3: which 8 columns of every chunk
      const int trow = q * 32 + lane;
      const int r = row_start + trow;
      const uint32_t lane_off = (uint32_t)(q * 32) << 16;
#pragma unroll 1
      for (int l = 0; l < 2; ++l) {
        const float* Hsrc = (l == 0 ? p.H2hl : p.H1hl) + (size_t)r * 2 * JB_HID;
        const uint32_t D = l == 0 ? D1 : D2;
        float tp[4][8];                                  // act'(h) of this thread's 32 columns, loaded under the MMAs
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float* hp = Hsrc + c * 32 + sub * 8;
          const float4 h0 = __ldg(reinterpret_cast<const float4*>(hp)), h1 = __ldg(reinterpret_cast<const float4*>(hp + 4));
          const float4 l0 = __ldg(reinterpret_cast<const float4*>(hp + JB_HID)), l1 = __ldg(reinterpret_cast<const float4*>(hp + JB_HID + 4));
          tp[c][0] = act_grad_from_output(h0.x + l0.x, p.act); tp[c][1] = act_grad_from_output(h0.y + l0.y, p.act);
          tp[c][2] = act_grad_from_output(h0.z + l0.z, p.act); tp[c][3] = act_grad_from_output(h0.w + l0.w, p.act);
          tp[c][4] = act_grad_from_output(h1.x + l1.x, p.act); tp[c][5] = act_grad_from_output(h1.y + l1.y, p.act);
          tp[c][6] = act_grad_from_output(h1.z + l1.z, p.act); tp[c][7] = act_grad_from_output(h1.w + l1.w, p.act);
        }
        if (l == 1) {                                    // layer 0's stores have read the staging slabs
          if (lane == 0) tc::tma_store_wait_read();
          tc::named_sync(1 + q, 128);
        }
        tc::mbar_wait(&bar_d[l], 0);
        tc::tc_fence_after();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int c0 = c * 32 + sub * 8;
          float v[8];
          tc::tmem_ld8(D + lane_off + c0, v);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] *= tp[c][j];
          tc::tmem_st8_split(AH + lane_off + c0, AL + lane_off + c0, v);
          tc::tmem_st_wait();
          tc::tc_fence_before();
          tc::mbar_arrive(&bar_a[l][c]);
          float4 h0, l0, h1, l1;
          tc::split4(make_float4(v[0], v[1], v[2], v[3]), h0, l0);
          tc::split4(make_float4(v[4], v[5], v[6], v[7]), h1, l1);
          uint8_t* stg = base + (c < 3 ? c * JB_STAGE_BYTES : JB_STAGES * JB_STAGE_BYTES) + q * 4096;     // [hi 16 KB | lo 16 KB]
          const uint32_t o0 = tc::sw128_off(lane, sub * 2), o1 = tc::sw128_off(lane, sub * 2 + 1);
          *reinterpret_cast<float4*>(stg + o0) = h0;
          *reinterpret_cast<float4*>(stg + o1) = h1;
          *reinterpret_cast<float4*>(stg + JB_SUB + o0) = l0;
          *reinterpret_cast<float4*>(stg + JB_SUB + o1) = l1;
          tc::fence_async_smem();
          tc::named_sync(1 + q, 128);
          if (sub == c && lane == 0) {
            const CUtensorMap* tm = l == 0 ? &p.tm_dz2 : &p.tm_dz1;
            tc::tma_store_2d(tm, c * 32, row_start + q * 32, stg);
            tc::tma_store_2d(tm, JB_HID + c * 32, row_start + q * 32, stg + JB_SUB);
            tc::tma_store_commit();
          }
        }
      }
      // ---- dX: 8 chunks of 32 columns, raw fp32, staged through the (now idle) ring and written by TMA
      if (lane == 0) tc::tma_store_wait_read();
      tc::named_sync(1 + q, 128);
      tc::mbar_wait(&bar_d[2], 0);
      tc::tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < 8; ++cc) {
        const uint32_t D = cc < 4 ? D1 : D2;
        float v[8];
        tc::tmem_ld8(D + lane_off + (cc & 3) * 32 + sub * 8, v);
        const int blk = cc >> 1;
        uint8_t* stg = base + (blk < 3 ? blk * JB_STAGE_BYTES : JB_STAGES * JB_STAGE_BYTES) + (cc & 1) * JB_SUB + q * 4096;
        *reinterpret_cast<float4*>(stg + tc::sw128_off(lane, sub * 2)) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(stg + tc::sw128_off(lane, sub * 2 + 1)) = make_float4(v[4], v[5], v[6], v[7]);
        tc::fence_async_smem();
        tc::named_sync(1 + q, 128);
        if (sub == (cc & 3) && lane == 0) {
          tc::tma_store_2d(&p.tm_dx, cc * 32, row_start + q * 32, stg);
          tc::tma_store_commit();
        }
      }
      if (lane == 0) tc::tma_store_wait_all();
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------ weight gradients
#define JW_ROWS 8                   // rows (= reduction steps) per pipeline stage: one 8-deep tf32 MMA
#ifndef JW_STAGES
#define JW_STAGES 4
#endif
#define JW_OFF_DZ1 0                // [8 slabs: 4 hi | 4 lo] x 8 rows x 128 B
#define JW_OFF_XS 8192              // [16 slabs: 8 hi | 8 lo]
#define JW_OFF_DZ2 24576
#define JW_OFF_H1 32768
#define JW_OFF_H2 40960
#define JW_OFF_DO 49152             // [2 slabs: hi | lo]
#define JW_STAGE_BYTES 51200
#define JW_SMEM (JW_STAGES * JW_STAGE_BYTES + 1024 + 1024)
#define JW_EPI_WARPS 8
#define JW_THREADS (64 + 32 * JW_EPI_WARPS)
#define JW_COLS 480                 // dW1 256 | dW2 128 | dW3^T 32 | db1 16 | db2 16 | db3 32

struct JsWgParams {
  CUtensorMap tm_dz1, tm_xs, tm_dz2, tm_h1, tm_h2, tm_do;      // slab maps
  const int4* chunks; const int* num_chunks;
  float* partials;                                              // [chunk][128][JW_COLS]
};

__device__ float g_jw_dbg[64];
extern "C" int sard_jw_debug_read(float* out, int n) {
  SARD_CUDA(cudaMemcpyFromSymbol(out, g_jw_dbg, sizeof(float) * (n < 64 ? n : 64)));
  return 0;
}
__global__ void __launch_bounds__(JW_THREADS, 1) jsmlp_wgrad_tc_kernel(const __grid_constant__ JsWgParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = tc::align1024(smem_raw);
  uint8_t* ones = base + JW_STAGES * JW_STAGE_BYTES;             // 1 KB of 1.0f: an all-ones operand atom
  __shared__ uint64_t bar_full[JW_STAGES], bar_empty[JW_STAGES], bar_done;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 1) tc::tmem_alloc(&tmem_slot, 512);
  if (tid == 0) {
    for (int s = 0; s < JW_STAGES; ++s) { tc::mbar_init(&bar_full[s], 1); tc::mbar_init(&bar_empty[s], 1); }
    tc::mbar_init(&bar_done, 1);
    tc::mbar_fence_init();
    tc::tma_prefetch_desc(&p.tm_dz1); tc::tma_prefetch_desc(&p.tm_xs); tc::tma_prefetch_desc(&p.tm_dz2);
    tc::tma_prefetch_desc(&p.tm_h1); tc::tma_prefetch_desc(&p.tm_h2); tc::tma_prefetch_desc(&p.tm_do);
  }
  if (tid >= 64 && tid < 64 + 256) reinterpret_cast<float*>(ones)[tid - 64] = 1.0f;
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  pdl_sync();

  const bool active = (int)blockIdx.x < *p.num_chunks;
  if (active) {
    const int4 t = p.chunks[blockIdx.x];
    const int row_start = t.x;
    const int nks = (t.y - t.x + JW_ROWS - 1) / JW_ROWS;        // rows past row_end inside the last step are padding (dZ = 0)
    if (warp == 0) {
      if (lane == 0) {
        for (int it = 0; it < nks; ++it) {
          const int s = it % JW_STAGES;
          tc::mbar_wait(&bar_empty[s], ((it / JW_STAGES) & 1) ^ 1);
          uint8_t* st = base + s * JW_STAGE_BYTES;
          const int row = row_start + it * JW_ROWS;
          tc::mbar_expect_tx(&bar_full[s], JW_STAGE_BYTES);
          tc::tma_load_3d(st + JW_OFF_DZ1, &p.tm_dz1, 0, row, 0, &bar_full[s]);
          tc::tma_load_3d(st + JW_OFF_XS, &p.tm_xs, 0, row, 0, &bar_full[s]);
          tc::tma_load_3d(st + JW_OFF_DZ2, &p.tm_dz2, 0, row, 0, &bar_full[s]);
          tc::tma_load_3d(st + JW_OFF_H1, &p.tm_h1, 0, row, 0, &bar_full[s]);
          tc::tma_load_3d(st + JW_OFF_H2, &p.tm_h2, 0, row, 0, &bar_full[s]);
          tc::tma_load_3d(st + JW_OFF_DO, &p.tm_do, 0, row, 0, &bar_full[s]);
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        constexpr uint32_t ID_W1 = tc::idesc_tf32(128, 256, 1, 1), ID_W2 = tc::idesc_tf32(128, 128, 1, 1), ID_W3 = tc::idesc_tf32(128, 32, 1, 1);
        constexpr uint32_t ID_B = tc::idesc_tf32(128, 16, 1, 0), ID_B3 = tc::idesc_tf32(128, 32, 0, 1);
        constexpr uint32_t SL = JW_ROWS * 128;                    // bytes between 32-feature slabs of one operand
        const uint32_t DW1 = tmem, DW2 = tmem + 256, DW3 = tmem + 384, DB1 = tmem + 416, DB2 = tmem + 432, DB3 = tmem + 448;
        // an operand of ones: every 8-row group aliases the same 1 KB atom (stride 0)
        const uint64_t one_k = tc::smem_desc_sw128(tc::smem_u32(ones), 16, 0);
        for (int it = 0; it < nks; ++it) {
          const int s = it % JW_STAGES;
          tc::mbar_wait(&bar_full[s], (it / JW_STAGES) & 1);
          tc::tc_fence_after();
          const uint32_t sa = tc::smem_u32(base + s * JW_STAGE_BYTES);
          if (blockIdx.x == 0 && it == 0) {
            const float* f = reinterpret_cast<const float*>(base + s * JW_STAGE_BYTES);
            g_jw_dbg[0] = (float)nks; g_jw_dbg[1] = (float)*p.num_chunks;
            for (int i = 0; i < 8; ++i) g_jw_dbg[2 + i] = f[i];                         // dz1 hi row 0
            for (int i = 0; i < 8; ++i) g_jw_dbg[10 + i] = f[JW_OFF_XS / 4 + i];        // xs hi row 0
            for (int i = 0; i < 8; ++i) g_jw_dbg[18 + i] = f[JW_OFF_DO / 4 + i];        // dO hi row 0
            g_jw_dbg[26] = reinterpret_cast<const float*>(ones)[5];
          }
          const uint32_t acc = it ? 1u : 0u;
          auto mn = [&](uint32_t off) { return tc::smem_desc_mn32(sa + off, SL, 512); };
          const uint64_t z1h = mn(JW_OFF_DZ1), z1l = mn(JW_OFF_DZ1 + 4 * SL), xh = mn(JW_OFF_XS), xl = mn(JW_OFF_XS + 8 * SL);
          const uint64_t z2h = mn(JW_OFF_DZ2), z2l = mn(JW_OFF_DZ2 + 4 * SL), h1h = mn(JW_OFF_H1), h1l = mn(JW_OFF_H1 + 4 * SL);
          const uint64_t h2h = mn(JW_OFF_H2), h2l = mn(JW_OFF_H2 + 4 * SL), doh = mn(JW_OFF_DO), dol = mn(JW_OFF_DO + SL);
          tc::mma_ss(DW1, z1h, xh, ID_W1, acc); tc::mma_ss(DW1, z1l, xh, ID_W1, 1u); tc::mma_ss(DW1, z1h, xl, ID_W1, 1u);
          tc::mma_ss(DW2, z2h, h1h, ID_W2, acc); tc::mma_ss(DW2, z2l, h1h, ID_W2, 1u); tc::mma_ss(DW2, z2h, h1l, ID_W2, 1u);
          tc::mma_ss(DW3, h2h, doh, ID_W3, acc); tc::mma_ss(DW3, h2l, doh, ID_W3, 1u); tc::mma_ss(DW3, h2h, dol, ID_W3, 1u);
          tc::mma_ss(DB1, z1h, one_k, ID_B, acc); tc::mma_ss(DB1, z1l, one_k, ID_B, 1u);
          tc::mma_ss(DB2, z2h, one_k, ID_B, acc); tc::mma_ss(DB2, z2l, one_k, ID_B, 1u);
          tc::mma_ss(DB3, one_k, doh, ID_B3, acc); tc::mma_ss(DB3, one_k, dol, ID_B3, 1u);
          tc::mma_commit(&bar_empty[s]);
        }
        tc::mma_commit(&bar_done);
      }
    } else {
      // ---- epilogue: two warps per TMEM lane quarter, 32-column groups 0..7 / 8..14; the pipeline stages are idle by
      //      now and serve as the warps' transpose scratch (coalesced partial rows)
      const int q = warp & 3, half = (warp - 2) >> 2;
      const uint32_t lane_off = (uint32_t)(q * 32) << 16;
      uint8_t* scr = base + (warp - 2) * 4096;
      float* dst = p.partials + ((size_t)blockIdx.x * 128 + q * 32) * JW_COLS;
      tc::mbar_wait(&bar_done, 0);
      tc::tc_fence_after();
#pragma unroll 1
      for (int g = half * 8; g < (half ? 15 : 8); ++g)
        tc::tmem_block_to_global(tmem + lane_off + g * 32, scr, lane, dst + g * 32, JW_COLS);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem, 512);
}

// One float4 of one body index's gradient block per thread: the chunks of that index are added in chunk order.
struct JsWgReduceParams {
  const float* partials; const int* grp_chunk_ptr; const int* slot_of_index; int max_index, out_dim;
  float* gW1; float* gW2; float* gW3; float* gb1; float* gb2; float* gb3;
};
__global__ void __launch_bounds__(256) jsmlp_wgrad_reduce_kernel(const JsWgReduceParams p) {
  pdl_sync();
  __shared__ int s_u;
  const int slot = blockIdx.y;
  if (threadIdx.x == 0) s_u = -1;
  __syncthreads();
  for (int u = threadIdx.x; u < p.max_index; u += blockDim.x)
    if (p.slot_of_index[u] == slot) s_u = u;
  __syncthreads();
  const int u = s_u;
  if (u < 0) return;
  const int k0 = p.grp_chunk_ptr[u], k1 = p.grp_chunk_ptr[u + 1];
  if (k0 >= k1) return;
  constexpr int PER = 128 * JW_COLS / 4;
  const int e4 = blockIdx.x * blockDim.x + threadIdx.x;
  if (e4 >= PER) return;
  const float4* P4 = reinterpret_cast<const float4*>(p.partials) + e4;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  int k = k0;
  for (; k + 3 < k1; k += 4) {
    const float4 a0 = P4[(size_t)k * PER], a1 = P4[(size_t)(k + 1) * PER], a2 = P4[(size_t)(k + 2) * PER], a3 = P4[(size_t)(k + 3) * PER];
    s.x += a0.x; s.y += a0.y; s.z += a0.z; s.w += a0.w;
    s.x += a1.x; s.y += a1.y; s.z += a1.z; s.w += a1.w;
    s.x += a2.x; s.y += a2.y; s.z += a2.z; s.w += a2.w;
    s.x += a3.x; s.y += a3.y; s.z += a3.z; s.w += a3.w;
  }
  for (; k < k1; ++k) {
    const float4 a = P4[(size_t)k * PER];
    s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
  }
  const int m = e4 / (JW_COLS / 4), c = (e4 - m * (JW_COLS / 4)) * 4;
  if (c < 256) {
    float4* d = reinterpret_cast<float4*>(p.gW1 + ((size_t)slot * JB_HID + m) * JB_IN + c);
    float4 o = *d; o.x += s.x; o.y += s.y; o.z += s.z; o.w += s.w; *d = o;
  } else if (c < 384) {
    float4* d = reinterpret_cast<float4*>(p.gW2 + ((size_t)slot * JB_HID + m) * JB_HID + (c - 256));
    float4 o = *d; o.x += s.x; o.y += s.y; o.z += s.z; o.w += s.w; *d = o;
  } else if (c < 416) {
    const float v[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int o = c - 384 + i;
      if (o < p.out_dim) p.gW3[((size_t)slot * p.out_dim + o) * JB_HID + m] += v[i];
    }
  } else if (c == 416) {
    p.gb1[(size_t)slot * JB_HID + m] += s.x;
  } else if (c == 432) {
    p.gb2[(size_t)slot * JB_HID + m] += s.x;
  } else if (c >= 448 && m == 0) {
    const float v[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int o = c - 448 + i;
      if (o < p.out_dim) p.gb3[(size_t)slot * p.out_dim + o] += v[i];
    }
  }
}

// ------------------------------------------------------------------------------------------ host
extern "C" int sard_jsmlp_tc_bwd_supported(int in_dim, int h1, int h2, int out_dim, int n_il) {
  return n_il == 3 && h1 == JB_HID && h2 == JB_HID && in_dim == JB_IN && out_dim >= 1 && out_dim <= 16;
}
extern "C" long long sard_jsmlp_tc_partials_floats(int max_chunks) { return (long long)(max_chunks > 0 ? max_chunks : 1) * 128 * JW_COLS; }

extern "C" int sard_jsmlp_tc_backward(const SardJsmlpTc* a, const SardJsmlpTcBwd* b, sard_stream_t stream) {
  if (a->n <= 0 || a->max_tiles <= 0) return 0;
  SARD_REQUIRE(sard_jsmlp_tc_bwd_supported(a->in_dim, a->hid, a->hid, a->out_dim, 3), "jsmlp_tc backward: unsupported shape");
  SARD_REQUIRE(a->n % 128 == 0, "jsmlp_tc backward: the row capacity must be a multiple of 128 (padded sort)");
  SARD_REQUIRE(a->Hhl[0] && a->Hhl[1] && b->dOs_hl && b->dZhl[0] && b->dZhl[1] && b->dxs && b->ld_dxs == JB_IN, "jsmlp_tc backward: buffers");
  const JsLayout L = sard_js_layout(a->n, a->in_dim, a->out_dim, a->n_slots);
  SARD_REQUIRE(L.total <= a->ws_floats, "jsmlp_tc: workspace too small");
  cudaStream_t st = as_stream(stream);
  if (b->dout) {            // NULL: the loss kernel already wrote dOs_hl (SardPolicyLoss.dout_sorted_hl)
    sard_launch(dout_sorted_split_kernel, grid_for((long long)a->n * 8, 256, 8), 256, 0, st, b->dout, b->ld_dout, a->out_dim, a->srow, a->n,
                b->dOs_hl);
    SARD_LAUNCH_OK("dout_sorted_split_kernel");
  }
  JsBwdParams p;
  memset(&p, 0, sizeof(p));
  SARD_TRY(sard_tmap_2d(&p.tm_do, b->dOs_hl, a->n, 64, 64, 128, 32));
  SARD_TRY(sard_tmap_2d(&p.tm_w3t, a->ws + L.wthl[2], (long long)a->n_slots * JB_HID, 64, 64, 128, 32));
  SARD_TRY(sard_tmap_2d(&p.tm_w2t, a->ws + L.wthl[1], (long long)a->n_slots * JB_HID, 2 * JB_HID, 2 * JB_HID, 128, 32));
  SARD_TRY(sard_tmap_2d(&p.tm_w1t, a->ws + L.wthl[0], (long long)a->n_slots * JB_IN, 2 * JB_HID, 2 * JB_HID, 128, 32));
  SARD_TRY(sard_tmap_2d(&p.tm_dz2, b->dZhl[1], a->n, 2 * JB_HID, 2 * JB_HID, 32, 32));
  SARD_TRY(sard_tmap_2d(&p.tm_dz1, b->dZhl[0], a->n, 2 * JB_HID, 2 * JB_HID, 32, 32));
  SARD_TRY(sard_tmap_2d(&p.tm_dx, b->dxs, a->n, JB_IN, JB_IN, 32, 32));
  p.tiles = reinterpret_cast<const int4*>(a->tiles); p.num_tiles = a->num_tiles; p.slot_of_index = a->slot_of_index;
  p.H1hl = a->Hhl[0]; p.H2hl = a->Hhl[1]; p.out_dim = a->out_dim; p.act = a->act;
  static bool configured[SARD_MAX_DEVICES] = {};
  if (sard_first_use_on_device(configured))
    SARD_CUDA(cudaFuncSetAttribute(jsmlp_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)JB_SMEM));
  ProfScope prof(st, 2.0 * a->n * ((double)JB_IN * JB_HID + (double)JB_HID * JB_HID + (double)JB_HID * a->out_dim),
                 4.0 * ((double)a->n * (JB_IN + 4 * JB_HID + a->out_dim) +
                        13.0 * ((double)JB_IN * JB_HID + (double)JB_HID * JB_HID + (double)JB_HID * a->out_dim)));
  sard_launch(jsmlp_bwd_tc_kernel, dim3(a->max_tiles), dim3(JB_THREADS), (size_t)JB_SMEM, st, p);
  SARD_LAUNCH_OK("jsmlp_bwd_tc_kernel");
  return 0;
}

extern "C" int sard_jsmlp_tc_wgrad(const SardJsmlpTc* a, const SardJsmlpTcBwd* b, sard_stream_t stream) {
  if (a->n <= 0 || b->max_chunks <= 0) return 0;
  SARD_REQUIRE(sard_jsmlp_tc_bwd_supported(a->in_dim, a->hid, a->hid, a->out_dim, 3), "jsmlp_tc wgrad: unsupported shape");
  SARD_REQUIRE(a->n % 128 == 0 && b->partials && b->chunks && b->grp_chunk_ptr, "jsmlp_tc wgrad: buffers");
  const JsLayout L = sard_js_layout(a->n, a->in_dim, a->out_dim, a->n_slots);
  cudaStream_t st = as_stream(stream);
  JsWgParams p;
  memset(&p, 0, sizeof(p));
  SARD_TRY(sard_tmap_slab(&p.tm_dz1, b->dZhl[0], a->n, 2 * JB_HID, 2 * JB_HID, JW_ROWS, 8));
  SARD_TRY(sard_tmap_slab(&p.tm_xs, a->ws + L.xs_hl, a->n, 2 * JB_IN, 2 * JB_IN, JW_ROWS, 16));
  SARD_TRY(sard_tmap_slab(&p.tm_dz2, b->dZhl[1], a->n, 2 * JB_HID, 2 * JB_HID, JW_ROWS, 8));
  SARD_TRY(sard_tmap_slab(&p.tm_h1, a->Hhl[0], a->n, 2 * JB_HID, 2 * JB_HID, JW_ROWS, 8));
  SARD_TRY(sard_tmap_slab(&p.tm_h2, a->Hhl[1], a->n, 2 * JB_HID, 2 * JB_HID, JW_ROWS, 8));
  SARD_TRY(sard_tmap_slab(&p.tm_do, b->dOs_hl, a->n, 64, 64, JW_ROWS, 2));
  p.chunks = reinterpret_cast<const int4*>(b->chunks); p.num_chunks = b->num_chunks; p.partials = b->partials;
  static bool configured[SARD_MAX_DEVICES] = {};
  if (sard_first_use_on_device(configured))
    SARD_CUDA(cudaFuncSetAttribute(jsmlp_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)JW_SMEM));
  {
    ProfScope prof(st, 2.0 * a->n * ((double)JB_IN * JB_HID + (double)JB_HID * JB_HID + (double)JB_HID * a->out_dim),
                   4.0 * ((double)a->n * (JB_IN + 4 * JB_HID + a->out_dim) +
                          13.0 * ((double)JB_IN * JB_HID + (double)JB_HID * JB_HID + (double)JB_HID * a->out_dim)));
    sard_launch(jsmlp_wgrad_tc_kernel, dim3(b->max_chunks), dim3(JW_THREADS), (size_t)JW_SMEM, st, p);
    SARD_LAUNCH_OK("jsmlp_wgrad_tc_kernel");
    JsWgReduceParams r;
    r.partials = b->partials; r.grp_chunk_ptr = b->grp_chunk_ptr; r.slot_of_index = a->slot_of_index; r.max_index = a->max_index;
    r.out_dim = a->out_dim;
    r.gW1 = b->gW[0]; r.gW2 = b->gW[1]; r.gW3 = b->gW[2]; r.gb1 = b->gb[0]; r.gb2 = b->gb[1]; r.gb3 = b->gb[2];
    sard_launch(jsmlp_wgrad_reduce_kernel, dim3(ceil_div(128 * JW_COLS / 4, 256), a->n_slots), dim3(256), 0, st, r);
    SARD_LAUNCH_OK("jsmlp_wgrad_reduce_kernel");
  }
  return 0;
}

// ------------------------------------------------------------------------------------------ descriptor probe (tools/tc_probe.py umma)
// One tf32 MMA (M = 128, K = 8) with operands written by the threads in the canonical SWIZZLE_128B layouts:
//   MN-major operand X[k][f]: slab f/32, row k, 32 floats (16-byte pieces XOR k);  K-major operand X[f][k]: row f, 32-float row
__global__ void __launch_bounds__(128, 1) umma_probe_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D,
                                                            int N, int a_mn, int b_mn, uint32_t lbo, uint32_t sbo) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = tc::align1024(smem_raw);
  uint8_t* sA = base;                 // 16 KB
  uint8_t* sB = base + 16384;         // 32 KB
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) tc::tmem_alloc(&tmem_slot, 256);
  if (tid == 0) { tc::mbar_init(&bar, 1); tc::mbar_fence_init(); }
  for (int i = tid; i < 49152 / 4; i += 128) reinterpret_cast<float*>(base)[i] = 0.f;
  __syncthreads();
  // A is given as A[k][m] (8 x 128), B as B[k][n] (8 x N)
  for (int e = tid; e < 8 * 128; e += 128) {
    const int k = e / 128, m = e % 128;
    const float v = A[e];
    if (a_mn) *reinterpret_cast<float*>(sA + (m / 32) * 1024 + tc::sw128b32_off(k, (m % 32) / 8) + (m % 8) * 4) = v;
    else *reinterpret_cast<float*>(sA + tc::sw128_off(m, k / 4) + (k % 4) * 4) = v;
  }
  for (int e = tid; e < 8 * N; e += 128) {
    const int k = e / N, n = e % N;
    const float v = B[e];
    if (b_mn) *reinterpret_cast<float*>(sB + (n / 32) * 1024 + tc::sw128b32_off(k, (n % 32) / 8) + (n % 8) * 4) = v;
    else *reinterpret_cast<float*>(sB + tc::sw128_off(n, k / 4) + (k % 4) * 4) = v;
  }
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint64_t ad = a_mn ? tc::smem_desc_mn32(tc::smem_u32(sA), lbo, sbo) : tc::kdesc(tc::smem_u32(sA));
    const uint64_t bd = b_mn ? tc::smem_desc_mn32(tc::smem_u32(sB), lbo, sbo) : tc::kdesc(tc::smem_u32(sB));
    const uint32_t id = tc::idesc_tf32(128, N, a_mn, b_mn);
    tc::mma_ss(tmem, ad, bd, id, 0u);
    tc::mma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::tc_fence_after();
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
  for (int c = 0; c < N; c += 8) {
    float v[8];
    tc::tmem_ld8(tmem + lane_off + c, v);
    for (int j = 0; j < 8; ++j) D[(size_t)(warp * 32 + lane) * N + c + j] = v[j];
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 256);
}
extern "C" int sard_umma_probe(const float* A, const float* B, float* D, int N, int a_mn, int b_mn, int lbo, int sbo, sard_stream_t stream) {
  static bool configured[SARD_MAX_DEVICES] = {};
  if (sard_first_use_on_device(configured))
    SARD_CUDA(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 50176 + 1024));
  sard_launch(umma_probe_kernel, dim3(1), dim3(128), (size_t)(50176 + 1024), as_stream(stream), A, B, D, N, a_mn, b_mn, (uint32_t)lbo, (uint32_t)sbo);
  SARD_LAUNCH_OK("umma_probe_kernel");
  return 0;
}
