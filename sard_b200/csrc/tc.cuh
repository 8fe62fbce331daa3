// Blackwell (sm_100a) building blocks shared by the fused tensor-core kernels:
//   mbarrier producer/consumer primitives, TMA (cp.async.bulk.tensor) tile loads, tcgen05 tensor-memory
//   allocation / MMA / load / store, UMMA shared-memory and instruction descriptors, and the host-side
//   tensor-map encoder (driver entry point looked up at run time, no link-time libcuda dependency).
//
// Numerics: every tensor-core product in this library is error-compensated 3xTF32.  A value x is stored as the
// pair (hi, lo) with hi = x with the 13 low mantissa bits cleared (exactly a TF32 number) and lo = x - hi
// (exact in fp32); a product a*b is accumulated as a_hi*b_hi + a_lo*b_hi + a_hi*b_lo in fp32 (TMEM), which keeps
// ~2^-21 relative error per product -- inside the rtol 1e-4 forward / 1e-3 gradient parity targets that a single
// TF32 pass (2^-11) would miss.  Operands are pre-split by the kernel that produces them ("pair arrays"
// [rows][hi(C) | lo(C)]), so the consumers are fed by TMA alone.
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// first 1024-byte boundary of the dynamic shared memory, computed as pointer + offset so that the compiler keeps the
// shared state space (LDS / STS instead of generic loads and stores)
__device__ __forceinline__ uint8_t* align1024(uint8_t* smem) { return smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u); }

// ---------------------------------------------------------------------------------------- split
__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xffffe000u); }
__device__ __forceinline__ void split4(const float4 v, float4& h, float4& l) {
  h.x = tf32_hi(v.x); l.x = v.x - h.x;
  h.y = tf32_hi(v.y); l.y = v.y - h.y;
  h.z = tf32_hi(v.z); l.z = v.z - h.z;
  h.w = tf32_hi(v.w); l.w = v.w - h.w;
}

// tanh(x) = 1 - 2 / (1 + 2^(x * 2/ln 2)) on the special-function unit (ex2.approx, rcp.approx): absolute error
// < 5e-7 over the whole range (saturates correctly at +-1), 5 instructions instead of tanhf's ~40.  The fused
// epilogues are bound by this math (one thread per tile row), and the activations feed 3xTF32 products whose own
// error is of the same order.
__device__ __forceinline__ float tanh_fast(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.8853900817779268f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
  return fmaf(-2.0f, r, 1.0f);
}

// ---------------------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded wait: a protocol bug must surface as a trapped kernel (an error the host sees), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  long long t0 = 0;
  for (uint32_t spin = 0;; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"      // the thread is suspended up to %3 ns per try
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity), "r"(20000u)
        : "memory");
    if (done) return;
    if ((spin & 0x3ff) == 0x3ff) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000LL) {           // ~2 s at 1.9 GHz
        printf("sard_b200: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, addr, parity);
        __trap();
      }
    }
  }
}

// ---------------------------------------------------------------------------------------- fences
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---------------------------------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}

// shared -> global tile store (bulk async group of the issuing thread)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, int c0, int c1, const void* src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(smem_u32(src)) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }   // sources reusable
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }        // writes complete
// byte offset of 16-byte piece j (0..7) of row r inside a SWIZZLE_128B slab of 128-byte rows (slab base 1024-aligned)
__device__ __forceinline__ uint32_t sw128_off(int r, int j) { return (uint32_t)(r * 128 + ((j ^ (r & 7)) << 4)); }

// ---------------------------------------------------------------------------------------- tensor memory
// whole warp; `cols` a power of two in [32, 512]; the base address lands in *slot (shared memory)
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 consecutive columns: thread `lane` of the warp receives its lane's 32 columns
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
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// One warp moves its 32 TMEM lanes x 32 columns (at taddr) to 32 global rows of `ld` floats starting at dst0: the block
// is transposed through a 4 KB warp-private shared-memory scratch so that every store instruction writes four whole
// 128-byte rows (a thread storing its own TMEM lane directly would touch 32 different rows per instruction).
__device__ __forceinline__ void tmem_block_to_global(uint32_t taddr, uint8_t* scr, int lane, float* dst0, size_t ld) {
  float v[32];
  tmem_ld32(taddr, v);
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(scr + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  __syncwarp();
  const int pc = lane & 7;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int row = 4 * k + (lane >> 3);
    *reinterpret_cast<float4*>(dst0 + (size_t)row * ld + pc * 4) = *reinterpret_cast<const float4*>(scr + row * 128 + ((pc ^ (row & 7)) << 4));
  }
  __syncwarp();
}
// named barrier among `threads` threads (ids 1..15; 0 is __syncthreads)
__device__ __forceinline__ void named_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
      ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
      : "memory");
}
// split 8 values and store hi / lo into two tensor-memory column ranges of this thread's lane
__device__ __forceinline__ void tmem_st8_split(uint32_t t_hi, uint32_t t_lo, const float* v) {
  float h[8], l[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { h[j] = tf32_hi(v[j]); l[j] = v[j] - h[j]; }
  tmem_st8(t_hi, h);
  tmem_st8(t_lo, l);
}

// ---------------------------------------------------------------------------------------- UMMA
// Instruction descriptor, kind::tf32, fp32 accumulator (cute/arch/mma_sm100_desc.hpp InstrDescriptor):
//   [4,6) c_format=1 (F32)  [7,10) a_format=2 (TF32)  [10,13) b_format=2  [15] a_major  [16] b_major (0 = K, 1 = MN)
//   [17,23) N>>3  [24,29) M>>4
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
// Shared-memory matrix descriptor, SWIZZLE_128B (layout type 2), sm_100 version bit.
//   K-major operand  (rows of 32 fp32 = 128 B along K): LBO ignored (1), SBO = bytes between 8-row groups (1024).
//   MN-major operand (rows of 32 fp32 = 128 B along M/N, 8 K-rows per swizzle atom): LBO = bytes between
//   32-element slabs along M/N, SBO = bytes between 8-row K groups.
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// MN-major tf32 operands exist in ONE shared-memory layout: SWIZZLE_128B_BASE32B (layout type 1; TMA swizzle mode
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): rows of 32 fp32 (128 B) along M/N, atoms of 4 K-rows, the 32-byte piece index
// XORed with (row & 3).  LBO = bytes between 32-element slabs along M/N, SBO = bytes between 4-row K groups.
__device__ __forceinline__ uint64_t smem_desc_mn32(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}
// byte offset of 32-byte piece j (0..3) of row r inside a SWIZZLE_128B_BASE32B slab of 128-byte rows
__device__ __forceinline__ uint32_t sw128b32_off(int r, int j) { return (uint32_t)(r * 128 + ((j ^ (r & 3)) << 5)); }
__device__ __forceinline__ uint64_t kdesc(uint32_t smem_addr) { return smem_desc_sw128(smem_addr, 16, 1024); }

// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void mma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// the mbarrier receives one arrival when every MMA issued so far by this thread has completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

}  // namespace tc

// ---------------------------------------------------------------------------------------- host: tensor maps
// fp32 tensor maps with SWIZZLE_128B.  rank 2: dims {cols, rows}, row pitch ld floats, box {box_cols (<= 32), box_rows}.
// rank 3 ("slab" maps for MN-major operands): dims {32, rows, cols/32} with strides {ld, 32} floats and box
// {32, box_rows, slabs}: shared memory receives [slab][row][32 floats] in the MN-major canonical layout of tf32 operands
// (SWIZZLE_128B_BASE32B: 32-byte pieces XOR row & 3).
int sard_tmap_2d(CUtensorMap* out, const float* base, long long rows, long long cols, long long ld, int box_rows, int box_cols);
int sard_tmap_slab(CUtensorMap* out, const float* base, long long rows, long long cols, long long ld, int box_rows, int slabs);

// scratch layout of the fused JSMLP kernels (floats): pair copies of the live weight slabs, forward (whl) and
// transposed (wthl), and of the sorted input rows
struct JsLayout { int outp[3], kp[3]; long long whl[3], wthl[3], xs_hl, total; };
JsLayout sard_js_layout(int n, int in_dim, int out_dim, int n_slots);
