// Fused JSMLP on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), error-compensated 3xTF32.
//
// Replaces, for one stage of Transform2ActPolicy, the three IndexLinear layers of JSMLP.forward
// (design_opt/models/jsmlp.py:32-40, 74-79):  out = IL3(tanh(IL2(tanh(IL1(x, ind)), ind)), ind)
// with ONE launch forward.  Rows arrive sorted by body index (sard_sort_by_index), cut into tiles of <= 128 rows that
// share one body index = one weight slab.  Per tile (one CTA, 192 threads):
//   warp 0      TMA producer: x tile k-chunks + the slab's weight k-chunks (pre-split hi|lo pair arrays) into a
//               3-stage shared-memory ring, mbarrier expect_tx / complete_tx
//   warp 1      MMA issuer: one thread issues tcgen05.mma.kind::tf32 (M=128, N=128|16, K=8), three per k-step
//               (hi*hi + lo*hi + hi*lo); layer 1 reads A from shared memory, layers 2 and 3 read A = the previous
//               layer's activations straight from TENSOR MEMORY (the hidden activations never leave the SM on their
//               way to the next layer); tcgen05.commit releases ring slots and publishes accumulators
//   warps 2-5   epilogue: thread i owns tile row i = TMEM lane i: tcgen05.ld the accumulator, + bias, tanh, write the
//               activation to HBM once (the backward needs it), split hi/lo and tcgen05.st it back as the next layer's
//               A operand, 32 columns at a time so layer l+1's MMAs start while layer l's epilogue is still running
// TMEM map (512 columns): [0,128) D1 (later D3), [128,256) A_hi, [256,384) A_lo, [384,512) D2.
#include "tc.cuh"
#include <mutex>
#include <string.h>

// ------------------------------------------------------------------------------------------ host: tensor maps
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

int sard_tmap_2d(CUtensorMap* out, const float* base, long long rows, long long cols, long long ld, int box_rows, int box_cols) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return sard_set_error("cuTensorMapEncodeTiled is not available from the driver");
  SARD_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld & 3) == 0, "tensor map: base / pitch must be 16-byte aligned");
  SARD_REQUIRE(box_cols * 4 <= 128 && box_rows <= 256 && rows > 0 && cols > 0, "tensor map: box");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return sard_set_error("cuTensorMapEncodeTiled(2d rows=%lld cols=%lld ld=%lld box=%dx%d) failed: %d", rows, cols, ld, box_rows, box_cols, (int)r);
  return 0;
}

int sard_tmap_slab(CUtensorMap* out, const float* base, long long rows, long long cols, long long ld, int box_rows, int slabs) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return sard_set_error("cuTensorMapEncodeTiled is not available from the driver");
  SARD_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld & 3) == 0 && cols % 32 == 0, "slab tensor map: alignment");
  cuuint64_t dims[3] = {32, (cuuint64_t)rows, (cuuint64_t)(cols / 32)};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 4, 128};
  cuuint32_t box[3] = {32, (cuuint32_t)box_rows, (cuuint32_t)slabs};
  cuuint32_t estr[3] = {1, 1, 1};
  // MN-major tf32 operands: the one layout tcgen05 reads them in (tc::smem_desc_mn32)
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return sard_set_error("cuTensorMapEncodeTiled(slab rows=%lld cols=%lld ld=%lld box=%dx%d) failed: %d", rows, cols, ld, box_rows, slabs, (int)r);
  return 0;
}

// ------------------------------------------------------------------------------------------ operand preparation
// Pair copies of the live weight slabs of one IndexLinear table W[max_index, out, in]:
//   Whl [slot][outp][hi(in) | lo(in)]     forward B operand  (K-major: k = in);   rows >= out are zero
//   Wthl[slot][in  ][hi(kp) | lo(kp)]     backward-data B operand (K-major: k = out); k >= out is zero
// One block per (body index, 64-element stripe); indices without a slot are skipped.
struct JsPresplitParams {
  const float* W[3]; int out[3], in[3], outp[3], kp[3];
  const int* slot_of_index; int max_index;
  float* Whl[3]; float* Wthl[3];
};
// One block per (stripe, slot, table): the body index of the slot is found by scanning slot_of_index.
__global__ void __launch_bounds__(256) jsmlp_presplit_kernel(const __grid_constant__ JsPresplitParams p) {
  pdl_sync();
  __shared__ int s_u;
  const int slot = blockIdx.y, j = blockIdx.z;
  if (threadIdx.x == 0) s_u = -1;
  __syncthreads();
  for (int u = threadIdx.x; u < p.max_index; u += blockDim.x)
    if (p.slot_of_index[u] == slot) s_u = u;
  __syncthreads();
  const int u = s_u;
  if (u < 0) return;
  const int out = p.out[j], in = p.in[j], outp = p.outp[j], kp = p.kp[j];
  const float* Wu = p.W[j] + (size_t)u * out * in;
  float* F = p.Whl[j] + (size_t)slot * outp * 2 * in;
  float* T = p.Wthl[j] + (size_t)slot * in * 2 * kp;
  // 32 x 32 tiles (o, i): rows of W are read and rows of F written 128 bytes at a time, the transposed copy goes through
  // shared memory so its rows are written whole too; the four loads of a thread are in flight together
  __shared__ float th[32][33], tl[32][33];
  const int tiles_i = in >> 5, n_tiles = ((outp + 31) >> 5) * tiles_i;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int o0 = (t / tiles_i) << 5, i0 = (t % tiles_i) << 5;
    float v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int o = o0 + w + 8 * q;
      v[q] = o < out ? __ldg(Wu + (size_t)o * in + i0 + lane) : 0.f;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int ol = w + 8 * q, o = o0 + ol;
      const float h = tc::tf32_hi(v[q]), l = v[q] - h;
      if (o < outp) {
        F[(size_t)o * 2 * in + i0 + lane] = h;
        F[(size_t)o * 2 * in + in + i0 + lane] = l;
      }
      th[ol][lane] = h; tl[ol][lane] = l;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int il = w + 8 * q, o = o0 + lane;
      if (o < kp && o < outp) {
        T[(size_t)(i0 + il) * 2 * kp + o] = th[lane][il];
        T[(size_t)(i0 + il) * 2 * kp + kp + o] = tl[lane][il];
      }
    }
    __syncthreads();
  }
  if (kp > outp) {                         // zero the k padding beyond the padded rows
    const int extra = (kp - outp) * in;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < extra; e += gridDim.x * blockDim.x) {
      const int o = outp + e / in, i = e % in;
      T[(size_t)i * 2 * kp + o] = 0.f;
      T[(size_t)i * 2 * kp + kp + o] = 0.f;
    }
  }
}

// pair copy of a row-major activation matrix: dst[r][0:C] = hi, dst[r][C:2C] = lo  (C % 4 == 0)
__global__ void __launch_bounds__(256) split_rows_kernel(const float* __restrict__ src, int ld_src, int n, int C4,
                                                         float* __restrict__ dst) {
  pdl_sync();
  const long long total = (long long)n * C4;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(idx / C4), c4 = (int)(idx - (long long)r * C4);
    const float4 v = *reinterpret_cast<const float4*>(src + (size_t)r * ld_src + c4 * 4);
    float4 h, l;
    tc::split4(v, h, l);
    float* d = dst + (size_t)r * 8 * C4;
    *reinterpret_cast<float4*>(d + c4 * 4) = h;
    *reinterpret_cast<float4*>(d + 4 * C4 + c4 * 4) = l;
  }
}

// ------------------------------------------------------------------------------------------ forward
#define JS_EPI_WARPS 16              // epilogue warps: 4 per TMEM lane quarter, one 32-column chunk each
#define JS_THREADS (64 + 32 * JS_EPI_WARPS)
#define JS_STAGES 3
#define JS_STAGE_BYTES 65536        // A_hi | A_lo | W_hi | W_lo, 16 KB each (128 rows x 32 fp32, SWIZZLE_128B)
#define JS_SUB 16384
#define JS_HID 128
#define JS_STAGING3 (JS_STAGES * JS_STAGE_BYTES)    // 32 KB store staging of column chunk 3 (chunks 0-2 reuse the ring's A regions)
#define JS_SMEM (JS_STAGES * JS_STAGE_BYTES + 32768 + 1024)

struct JsFwdParams {
  CUtensorMap tm_x, tm_w1, tm_w2, tm_w3;
  CUtensorMap tm_h1, tm_h2;                               // H1hl / H2hl [n, 256], 32 x 32 store boxes
  const int4* tiles; const int* num_tiles; const int* slot_of_index;
  const float* b1; const float* b2; const float* b3;     // bias tables [max_index, out]
  float* H1; float* H2; int ldh;                          // hidden activations, raw fp32 [n, 128] (sorted rows)
  float* H1hl; float* H2hl;                               // optional pair copies [n, 256]
  float* out; float* pre; int ldo; const int* srow;      // final layer, scattered to row srow[p]
  int in_dim, out_dim, act, post, padded;
};

// 8 activations; the branch on the activation kind is taken once per call, not once per element
__device__ __forceinline__ void js_act8(float* v, const float* b, int act) {
  if (act == SARD_ACT_TANH) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = tc::tanh_fast(v[j] + b[j]);
  } else if (act == SARD_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j] + b[j], 0.f);
  } else if (act == SARD_ACT_SIGMOID) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 1.f / (1.f + __expf(-(v[j] + b[j])));
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] += b[j];
  }
}

// cycle stamps of block 0 (debug / tuning): sard_tc_debug_read copies them out
__device__ long long g_js_dbg[32];
extern "C" int sard_tc_debug_read(long long* out, int n) {
  SARD_CUDA(cudaMemcpyFromSymbol(out, g_js_dbg, sizeof(long long) * (n < 32 ? n : 32)));
  return 0;
}
#define JS_STAMP(i) do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) g_js_dbg[i] = clock64(); } while (0)

__global__ void __launch_bounds__(JS_THREADS, 1) jsmlp_fwd_tc_kernel(const __grid_constant__ JsFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = tc::align1024(smem_raw);
  __shared__ uint64_t bar_full[JS_STAGES], bar_empty[JS_STAGES], bar_d[3], bar_a[2][4];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 1) tc::tmem_alloc(&tmem_slot, 512);
  if (tid == 0) {
    for (int s = 0; s < JS_STAGES; ++s) { tc::mbar_init(&bar_full[s], 1); tc::mbar_init(&bar_empty[s], 1); }
    for (int l = 0; l < 3; ++l) tc::mbar_init(&bar_d[l], 1);
    for (int l = 0; l < 2; ++l)
      for (int c = 0; c < 4; ++c) tc::mbar_init(&bar_a[l][c], 32 * JS_EPI_WARPS);
    tc::mbar_fence_init();
    tc::tma_prefetch_desc(&p.tm_x); tc::tma_prefetch_desc(&p.tm_w1); tc::tma_prefetch_desc(&p.tm_w2); tc::tma_prefetch_desc(&p.tm_w3);
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) JS_STAMP(0);
  pdl_sync();                                  // everything above overlaps the predecessor's tail
  if (tid == 0) JS_STAMP(1);

  const bool active = (int)blockIdx.x < *p.num_tiles;
  if (active) {
    const int4 t = p.tiles[blockIdx.x];
    const int row_start = t.x, row_end = t.y, u = t.z;
    const int slot = p.slot_of_index[u];
    const int nk1 = p.in_dim >> 5;             // 32-wide k-chunks of layer 1
    constexpr uint32_t ID128 = tc::idesc_tf32(128, 128, 0, 0), ID16 = tc::idesc_tf32(128, 16, 0, 0);
    const uint32_t D1 = tmem, AH = tmem + 128, AL = tmem + 256, D2 = tmem + 384;

    if (warp == 0) {
      if (lane == 0) {
        int it = 0;
        for (int kc = 0; kc < nk1; ++kc, ++it) {
          const int s = it % JS_STAGES;
          tc::mbar_wait(&bar_empty[s], ((it / JS_STAGES) & 1) ^ 1);
          uint8_t* st = base + s * JS_STAGE_BYTES;
          tc::mbar_expect_tx(&bar_full[s], 4 * JS_SUB);
          tc::tma_load_2d(st, &p.tm_x, kc * 32, row_start, &bar_full[s]);
          tc::tma_load_2d(st + JS_SUB, &p.tm_x, p.in_dim + kc * 32, row_start, &bar_full[s]);
          tc::tma_load_2d(st + 2 * JS_SUB, &p.tm_w1, kc * 32, slot * JS_HID, &bar_full[s]);
          tc::tma_load_2d(st + 3 * JS_SUB, &p.tm_w1, p.in_dim + kc * 32, slot * JS_HID, &bar_full[s]);
        }
        for (int kc = 0; kc < JS_HID / 32; ++kc, ++it) {
          const int s = it % JS_STAGES;
          tc::mbar_wait(&bar_empty[s], ((it / JS_STAGES) & 1) ^ 1);
          uint8_t* st = base + s * JS_STAGE_BYTES;
          tc::mbar_expect_tx(&bar_full[s], 2 * JS_SUB);
          tc::tma_load_2d(st + 2 * JS_SUB, &p.tm_w2, kc * 32, slot * JS_HID, &bar_full[s]);
          tc::tma_load_2d(st + 3 * JS_SUB, &p.tm_w2, JS_HID + kc * 32, slot * JS_HID, &bar_full[s]);
        }
        {                                       // layer 3: the whole [16 x 128] slab (hi and lo) in one stage
          const int s = it % JS_STAGES;
          tc::mbar_wait(&bar_empty[s], ((it / JS_STAGES) & 1) ^ 1);
          uint8_t* st = base + s * JS_STAGE_BYTES;
          tc::mbar_expect_tx(&bar_full[s], 8 * 2048);
          for (int kc = 0; kc < 4; ++kc) {
            tc::tma_load_2d(st + 2 * JS_SUB + kc * 2048, &p.tm_w3, kc * 32, slot * 16, &bar_full[s]);
            tc::tma_load_2d(st + 3 * JS_SUB + kc * 2048, &p.tm_w3, JS_HID + kc * 32, slot * 16, &bar_full[s]);
          }
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        int it = 0;
        // ---- layer 1: A and B from shared memory
        for (int kc = 0; kc < nk1; ++kc, ++it) {
          const int s = it % JS_STAGES;
          tc::mbar_wait(&bar_full[s], (it / JS_STAGES) & 1);
          tc::tc_fence_after();
          const uint32_t sa = tc::smem_u32(base + s * JS_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t a_hi = tc::kdesc(sa + k * 32), a_lo = tc::kdesc(sa + JS_SUB + k * 32);
            const uint64_t b_hi = tc::kdesc(sa + 2 * JS_SUB + k * 32), b_lo = tc::kdesc(sa + 3 * JS_SUB + k * 32);
            tc::mma_ss(D1, a_hi, b_hi, ID128, (kc | k) ? 1u : 0u);
            tc::mma_ss(D1, a_lo, b_hi, ID128, 1u);
            tc::mma_ss(D1, a_hi, b_lo, ID128, 1u);
          }
          tc::mma_commit(&bar_empty[s]);
        }
        tc::mma_commit(&bar_d[0]);
        JS_STAMP(8);
        // ---- layer 2: A = tanh(layer 1) from tensor memory, chunk by chunk as the epilogue produces it
        for (int kc = 0; kc < 4; ++kc, ++it) {
          const int s = it % JS_STAGES;
          tc::mbar_wait(&bar_a[0][kc], 0);
          tc::mbar_wait(&bar_full[s], (it / JS_STAGES) & 1);
          tc::tc_fence_after();
          const uint32_t sa = tc::smem_u32(base + s * JS_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t b_hi = tc::kdesc(sa + 2 * JS_SUB + k * 32), b_lo = tc::kdesc(sa + 3 * JS_SUB + k * 32);
            const uint32_t col = kc * 32 + k * 8;
            tc::mma_ts(D2, AH + col, b_hi, ID128, (kc | k) ? 1u : 0u);
            tc::mma_ts(D2, AL + col, b_hi, ID128, 1u);
            tc::mma_ts(D2, AH + col, b_lo, ID128, 1u);
          }
          tc::mma_commit(&bar_empty[s]);
        }
        tc::mma_commit(&bar_d[1]);
        JS_STAMP(9);
        // ---- layer 3 (N = 16): accumulator reuses D1's columns
        {
          const int s = it % JS_STAGES;
          tc::mbar_wait(&bar_full[s], (it / JS_STAGES) & 1);
          const uint32_t sa = tc::smem_u32(base + s * JS_STAGE_BYTES);
          for (int kc = 0; kc < 4; ++kc) {
            tc::mbar_wait(&bar_a[1][kc], 0);
            tc::tc_fence_after();
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t b_hi = tc::kdesc(sa + 2 * JS_SUB + kc * 2048 + k * 32), b_lo = tc::kdesc(sa + 3 * JS_SUB + kc * 2048 + k * 32);
              const uint32_t col = kc * 32 + k * 8;
              tc::mma_ts(D1, AH + col, b_hi, ID16, (kc | k) ? 1u : 0u);
              tc::mma_ts(D1, AL + col, b_hi, ID16, 1u);
              tc::mma_ts(D1, AH + col, b_lo, ID16, 1u);
            }
          }
          tc::mma_commit(&bar_empty[s]);
          tc::mma_commit(&bar_d[2]);
        }
      }
    } else {
      // ---- epilogue: thread <-> tile row <-> TMEM lane (a warp may only touch lanes 32*(warp%4) ..).  The four warps
      //      of a lane quarter work on the SAME 32-column chunk, 8 columns each, so the chunks of a layer complete one
      //      after the other and the next layer's MMAs (which wait per chunk) trail the epilogue instead of following it
      const int q = warp & 3;
      const int sub = (warp - 2) >> 2;                   // 0..3: which 8 columns of every chunk
      const int trow = q * 32 + lane;
      const int r = row_start + trow;
      const bool valid = r < row_end;
      const uint32_t lane_off = (uint32_t)(q * 32) << 16;
      const bool slab_full = p.padded || row_start + q * 32 + 32 <= row_end;   // this warp's 32 rows all belong to the tile
#pragma unroll 1
      for (int l = 0; l < 2; ++l) {
        const float* bias = (l == 0 ? p.b1 : p.b2) + (size_t)u * JS_HID;
        float* H = l == 0 ? p.H1 : p.H2;
        float* Hhl = l == 0 ? p.H1hl : p.H2hl;
        const uint32_t D = l == 0 ? D1 : D2;
        float bv[4][8];                                  // this thread's 32 bias values (8 per chunk), loaded under the MMAs
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c * 32 + sub * 8));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + c * 32 + sub * 8 + 4));
          bv[c][0] = b0.x; bv[c][1] = b0.y; bv[c][2] = b0.z; bv[c][3] = b0.w;
          bv[c][4] = b1.x; bv[c][5] = b1.y; bv[c][6] = b1.z; bv[c][7] = b1.w;
        }
        if (l == 1 && Hhl) {                             // layer 1's stores have read the staging slabs
          if (lane == 0) tc::tma_store_wait_read();
          tc::named_sync(1 + q, 128);
        }
        tc::mbar_wait(&bar_d[l], 0);
        tc::tc_fence_after();
        if (warp == 2) JS_STAMP(2 + 2 * l);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int c0 = c * 32 + sub * 8;
          float v[8];
          tc::tmem_ld8(D + lane_off + c0, v);
          js_act8(v, bv[c], p.act);
          // hand the activation to the next layer first (its MMAs wait on this), then write it out for the backward
          tc::tmem_st8_split(AH + lane_off + c0, AL + lane_off + c0, v);
          tc::tmem_st_wait();
          tc::tc_fence_before();
          tc::mbar_arrive(&bar_a[l][c]);
          if (H && valid) {                               // raw copy: bring-up / SIMT consumers only (scattered stores, slow)
            float* hp = H + (size_t)r * p.ldh + c0;
            *reinterpret_cast<float4*>(hp) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(hp + 4) = make_float4(v[4], v[5], v[6], v[7]);
          }
          if (Hhl) {
            float4 h0, l0, h1, l1;
            tc::split4(make_float4(v[0], v[1], v[2], v[3]), h0, l0);
            tc::split4(make_float4(v[4], v[5], v[6], v[7]), h1, l1);
            // pair copy for the backward / weight gradient: staged in shared memory in the swizzled box layout
            // (conflict-free: 8 lanes cover all 32 banks) and written out by TMA, one 32 x 32 box per quarter and chunk
            uint8_t* stg = base + (c < 3 ? c * JS_STAGE_BYTES : JS_STAGING3) + q * 4096;     // [hi 16 KB | lo 16 KB]
            if (slab_full) {
              const uint32_t o0 = tc::sw128_off(lane, sub * 2), o1 = tc::sw128_off(lane, sub * 2 + 1);
              *reinterpret_cast<float4*>(stg + o0) = h0;
              *reinterpret_cast<float4*>(stg + o1) = h1;
              *reinterpret_cast<float4*>(stg + JS_SUB + o0) = l0;
              *reinterpret_cast<float4*>(stg + JS_SUB + o1) = l1;
              tc::fence_async_smem();
              tc::named_sync(1 + q, 128);
              if (sub == c && lane == 0) {
                const CUtensorMap* tm = l == 0 ? &p.tm_h1 : &p.tm_h2;
                tc::tma_store_2d(tm, c * 32, row_start + q * 32, stg);
                tc::tma_store_2d(tm, JS_HID + c * 32, row_start + q * 32, stg + JS_SUB);
                tc::tma_store_commit();
              }
            } else if (valid) {                           // ragged last tile of a body index: predicated row stores
              float* hp = Hhl + (size_t)r * 2 * JS_HID + c0;
              *reinterpret_cast<float4*>(hp) = h0;
              *reinterpret_cast<float4*>(hp + 4) = h1;
              *reinterpret_cast<float4*>(hp + JS_HID) = l0;
              *reinterpret_cast<float4*>(hp + JS_HID + 4) = l1;
            }
          }
        }
        if (warp == 2) JS_STAMP(3 + 2 * l);
      }
      if (sub == 0) {
        tc::mbar_wait(&bar_d[2], 0);
        tc::tc_fence_after();
        if (warp == 2) JS_STAMP(6);
        float v[16];
        tc::tmem_ld16(D1 + lane_off, v);
        if (valid) {
          const float* bias = p.b3 + (size_t)u * p.out_dim;
          const long long yrow = p.srow ? p.srow[r] : r;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (j < p.out_dim) {
              const float x = v[j] + __ldg(bias + j);
              if (p.pre) p.pre[yrow * p.ldo + j] = x;
              p.out[yrow * p.ldo + j] = p.post == SARD_POST_SIN_HALF_PI ? sinpif(0.5f * x) : x;
            }
          }
        }
        if (warp == 2) JS_STAMP(7);
      }
      if (lane == 0) tc::tma_store_wait_all();
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------ host
// floats; every block 256-byte aligned
JsLayout sard_js_layout(int n, int in_dim, int out_dim, int n_slots) {
  JsLayout L;
  const int in[3] = {in_dim, JS_HID, JS_HID};
  L.outp[0] = JS_HID; L.outp[1] = JS_HID; L.outp[2] = 16;
  L.kp[0] = JS_HID; L.kp[1] = JS_HID; L.kp[2] = 32;
  long long off = 0;
  auto take = [&](long long v) { const long long o = off; off += (v + 63) & ~63LL; return o; };
  for (int j = 0; j < 3; ++j) L.whl[j] = take((long long)n_slots * L.outp[j] * 2 * in[j]);
  for (int j = 0; j < 3; ++j) L.wthl[j] = take((long long)n_slots * in[j] * 2 * L.kp[j]);
  L.xs_hl = take((long long)n * 2 * in_dim);
  L.total = off + 64;
  (void)out_dim;
  return L;
}
static inline JsLayout js_layout(int n, int in_dim, int out_dim, int n_slots) { return sard_js_layout(n, in_dim, out_dim, n_slots); }

extern "C" int sard_jsmlp_tc_supported(int in_dim, int h1, int h2, int out_dim, int n_il) {
  return n_il == 3 && h1 == JS_HID && h2 == JS_HID && in_dim % 32 == 0 && in_dim >= 32 && in_dim <= 1024 && out_dim >= 1 && out_dim <= 16;
}
extern "C" long long sard_jsmlp_tc_workspace(int n, int in_dim, int out_dim, int n_slots) {
  return js_layout(n > 0 ? n : 1, in_dim, out_dim, n_slots > 0 ? n_slots : 1).total;
}

// weight pair copies of the three tables (live slots only) into ws
extern "C" int sard_jsmlp_tc_presplit(const SardJsmlpTc* a, sard_stream_t stream) {
  SARD_REQUIRE(sard_jsmlp_tc_supported(a->in_dim, a->hid, a->hid, a->out_dim, 3), "jsmlp_tc: unsupported shape");
  const JsLayout L = js_layout(a->n, a->in_dim, a->out_dim, a->n_slots);
  SARD_REQUIRE(L.total <= a->ws_floats, "jsmlp_tc: workspace too small");
  SARD_REQUIRE((reinterpret_cast<uintptr_t>(a->ws) & 255) == 0, "jsmlp_tc: workspace must be 256-byte aligned");
  cudaStream_t st = as_stream(stream);
  const int in[3] = {a->in_dim, JS_HID, JS_HID}, out[3] = {JS_HID, JS_HID, a->out_dim};
  JsPresplitParams p;
  for (int j = 0; j < 3; ++j) {
    p.W[j] = a->W[j]; p.out[j] = out[j]; p.in[j] = in[j]; p.outp[j] = L.outp[j]; p.kp[j] = L.kp[j];
    p.Whl[j] = a->ws + L.whl[j]; p.Wthl[j] = a->ws + L.wthl[j];
  }
  p.slot_of_index = a->slot_of_index; p.max_index = a->max_index;
  sard_launch(jsmlp_presplit_kernel, dim3(32, a->n_slots, 3), 256, 0, st, p);
  SARD_LAUNCH_OK("jsmlp_presplit_kernel");
  return 0;
}

// pair copy of the sorted input rows xs[n, in_dim] into ws (callers that build xs themselves write it directly)
extern "C" int sard_jsmlp_tc_split_input(const SardJsmlpTc* a, sard_stream_t stream) {
  if (a->n <= 0) return 0;
  const JsLayout L = js_layout(a->n, a->in_dim, a->out_dim, a->n_slots);
  SARD_REQUIRE(a->xs && a->ld_xs % 4 == 0, "jsmlp_tc: xs");
  sard_launch(split_rows_kernel, grid_for((long long)a->n * (a->in_dim / 4), 256, 16), 256, 0, as_stream(stream), a->xs, a->ld_xs,
              a->n, a->in_dim / 4, a->ws + L.xs_hl);
  SARD_LAUNCH_OK("split_rows_kernel");
  return 0;
}
extern "C" float* sard_jsmlp_tc_xs_hl(const SardJsmlpTc* a) {
  return a->ws + js_layout(a->n, a->in_dim, a->out_dim, a->n_slots).xs_hl;
}

extern "C" int sard_jsmlp_tc_forward(const SardJsmlpTc* a, sard_stream_t stream) {
  if (a->n <= 0 || a->max_tiles <= 0) return 0;
  SARD_REQUIRE(sard_jsmlp_tc_supported(a->in_dim, a->hid, a->hid, a->out_dim, 3), "jsmlp_tc: unsupported shape");
  const JsLayout L = js_layout(a->n, a->in_dim, a->out_dim, a->n_slots);
  SARD_REQUIRE(L.total <= a->ws_floats, "jsmlp_tc: workspace too small");
  cudaStream_t st = as_stream(stream);
  JsFwdParams p;
  memset(&p, 0, sizeof(p));
  SARD_TRY(sard_tmap_2d(&p.tm_x, a->ws + L.xs_hl, a->n, 2 * a->in_dim, 2 * a->in_dim, 128, 32));
  SARD_TRY(sard_tmap_2d(&p.tm_w1, a->ws + L.whl[0], (long long)a->n_slots * JS_HID, 2 * a->in_dim, 2 * a->in_dim, 128, 32));
  SARD_TRY(sard_tmap_2d(&p.tm_w2, a->ws + L.whl[1], (long long)a->n_slots * JS_HID, 2 * JS_HID, 2 * JS_HID, 128, 32));
  SARD_TRY(sard_tmap_2d(&p.tm_w3, a->ws + L.whl[2], (long long)a->n_slots * 16, 2 * JS_HID, 2 * JS_HID, 16, 32));
  p.tiles = reinterpret_cast<const int4*>(a->tiles); p.num_tiles = a->num_tiles; p.slot_of_index = a->slot_of_index;
  p.b1 = a->b[0]; p.b2 = a->b[1]; p.b3 = a->b[2];
  p.H1 = a->H[0]; p.H2 = a->H[1]; p.ldh = a->ldh; p.H1hl = a->Hhl[0]; p.H2hl = a->Hhl[1];
  p.out = a->out; p.pre = a->pre; p.ldo = a->ldo; p.srow = a->srow;
  p.in_dim = a->in_dim; p.out_dim = a->out_dim; p.act = a->act; p.post = a->post; p.padded = a->padded;
  SARD_REQUIRE(!a->padded || a->n % 128 == 0, "jsmlp_tc: a padded row space has a multiple of 128 rows");
  if (a->Hhl[0]) SARD_TRY(sard_tmap_2d(&p.tm_h1, a->Hhl[0], a->n, 2 * JS_HID, 2 * JS_HID, 32, 32));
  if (a->Hhl[1]) SARD_TRY(sard_tmap_2d(&p.tm_h2, a->Hhl[1], a->n, 2 * JS_HID, 2 * JS_HID, 32, 32));
  constexpr size_t smem = JS_SMEM;
  static bool configured[SARD_MAX_DEVICES] = {};
  if (sard_first_use_on_device(configured))
    SARD_CUDA(cudaFuncSetAttribute(jsmlp_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ProfScope prof(st, 2.0 * a->n * ((double)a->in_dim * JS_HID + (double)JS_HID * JS_HID + (double)JS_HID * a->out_dim),
                 4.0 * ((double)a->n * (a->in_dim + 2 * JS_HID + a->out_dim) +
                        13.0 * ((double)a->in_dim * JS_HID + (double)JS_HID * JS_HID + (double)JS_HID * a->out_dim)));
  sard_launch(jsmlp_fwd_tc_kernel, dim3(a->max_tiles), dim3(JS_THREADS), smem, st, p);
  SARD_LAUNCH_OK("jsmlp_fwd_tc_kernel");
  return 0;
}
