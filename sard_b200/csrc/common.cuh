// Shared helpers for the sard_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/sard_b200.h"

int sard_set_error(const char* fmt, ...);
extern long long g_sard_launches;          // kernels launched by this library (bench.py's gpu_launches)
extern thread_local int g_sard_site;       // call-site label of the next linear / aggregate launch (profiling)

// Optional per-call-site timing with CUDA events on the launching stream (sard_prof_*): a no-op
// unless the site's bit is enabled, so the timed path pays one branch.
struct ProfScope {
  int slot;
  cudaStream_t st;
  ProfScope(cudaStream_t stream, double flops, double bytes);
  ~ProfScope();
};

#define SARD_CUDA(call)                                                                      \
  do {                                                                                       \
    cudaError_t e__ = (call);                                                                \
    if (e__ != cudaSuccess)                                                                  \
      return sard_set_error("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
  } while (0)

// Per-kernel timing of EVERY launch (sard_prof_kernels_enable): events around each launch on its stream, keyed by the
// name SARD_LAUNCH_OK reports; sard_prof_bytes() attaches the algorithmic bytes of the next launch (HBM-class kernels).
extern int g_sard_prof_all;
void sard_prof_pre(cudaStream_t st);
void sard_prof_post(cudaStream_t st);
void sard_prof_label(const char* name);
void sard_prof_bytes(double bytes);

#define SARD_LAUNCH_OK(name)                                                                 \
  do {                                                                                       \
    __atomic_fetch_add(&g_sard_launches, 1LL, __ATOMIC_RELAXED);   /* two host threads issue */ \
    if (g_sard_prof_all) sard_prof_label(name);                                              \
    cudaError_t e__ = cudaGetLastError();                                                    \
    if (e__ != cudaSuccess)                                                                  \
      return sard_set_error("launch %s failed: %s", name, cudaGetErrorString(e__));          \
  } while (0)

// Programmatic dependent launch (sm_90+): every kernel of this library is launched with the programmatic
// stream-serialization attribute and opens with pdl_sync() = griddepcontrol.wait, so the launch of kernel
// k+1 is processed while kernel k drains and its CTAs start the moment k's last CTA exits (the implicit
// trigger).  Measured on the 50 k-transition update (two streams): 319 -> 296 ms.  An EARLY trigger
// (griddepcontrol.launch_dependents at kernel entry) was measured too and is worse (376 ms): the parked
// CTAs of k+1 hold shared memory and registers that the other stream's kernels need.  A LATE trigger (after the
// GEMM kernels' main loop) was neutral to slightly worse (195.9 -> 197.3 ms).
// Launched without the attribute (sard_set_pdl(0)) the wait is a no-op.
extern int g_sard_pdl;
__device__ __forceinline__ void pdl_sync() {
#ifdef SARD_PDL_EARLY_TRIGGER
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
template <typename... P, typename... A>
static inline void sard_launch(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = g_sard_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (g_sard_prof_all) sard_prof_pre(st);
  cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(args)...);   // errors surface through SARD_LAUNCH_OK
  if (g_sard_prof_all) sard_prof_post(st);
}

// cudaFuncSetAttribute is per device: remember per (kernel instantiation, device) whether it was applied
#define SARD_MAX_DEVICES 32
static inline bool sard_first_use_on_device(bool (&done)[SARD_MAX_DEVICES]) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= SARD_MAX_DEVICES) return true;
  if (done[dev]) return false;
  done[dev] = true;
  return true;
}

#define SARD_REQUIRE(cond, msg)                                                              \
  do {                                                                                       \
    if (!(cond)) return sard_set_error("%s:%d requirement failed: %s (%s)", __FILE__, __LINE__, #cond, msg); \
  } while (0)

#define SARD_TRY(call)                                                                       \
  do {                                                                                       \
    int r__ = (call);                                                                        \
    if (r__ != 0) return r__;                                                                \
  } while (0)

static inline cudaStream_t as_stream(sard_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// Grid sizing: a multiple of the SM count for grid-stride kernels (148 SMs on B200).
static inline int grid_for(long long work_items, int per_block, int max_waves = 8) {
  long long blocks = (work_items + per_block - 1) / per_block;
  long long cap = 148LL * max_waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Transcendental epilogues live in out-of-line functions: inlining tanhf/expf/sinf into 32-element
// unrolled epilogues made every GEMM kernel > 64 KB of SASS, i.e. instruction-cache-miss bound.
static __device__ __noinline__ float epi_slow(float v, int act, int post) {
  if (act == SARD_ACT_TANH) v = tanhf(v);
  else if (act == SARD_ACT_SIGMOID) v = 1.f / (1.f + expf(-v));
  else if (act == SARD_ACT_RELU) v = v > 0.f ? v : 0.f;
  if (post == SARD_POST_SIN_HALF_PI) v = sinpif(0.5f * v);       // sin(pi/2 * v), exact argument reduction
  return v;
}
// forward epilogue after the bias: activation; *pre receives the value before `post`
__device__ __forceinline__ float epi_forward(float v, int act, int post, float* pre) {
  if (act <= SARD_ACT_RELU && post == SARD_POST_NONE) {
    v = (act == SARD_ACT_RELU && v < 0.f) ? 0.f : v;
    if (pre) *pre = v;
    return v;
  }
  if (pre) {                       // rare (attribute head): keep the pre-sin value
    const float a = epi_slow(v, act, SARD_POST_NONE);
    *pre = a;
    return post == SARD_POST_SIN_HALF_PI ? epi_slow(a, SARD_ACT_NONE, post) : a;
  }
  return epi_slow(v, act, post);
}
static __device__ __noinline__ float act_grad_slow(float y, int act) {
  return act == SARD_ACT_TANH ? 1.f - y * y : y * (1.f - y);
}

__device__ __forceinline__ float act_apply(float v, int act) {
  switch (act) {
    case SARD_ACT_RELU: return v > 0.f ? v : 0.f;
    case SARD_ACT_TANH: return tanhf(v);
    case SARD_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    default: return v;
  }
}
// derivative of the activation expressed through its OUTPUT y
__device__ __forceinline__ float act_grad_from_output(float y, int act) {
  if (act == SARD_ACT_NONE) return 1.f;
  if (act == SARD_ACT_RELU) return y > 0.f ? 1.f : 0.f;
  return act == SARD_ACT_TANH ? 1.f - y * y : y * (1.f - y);
}
