// Message passing and segment kernels: deterministic CSR gather/reduce (GraphConv aggregate and its
// adjoint), the sorted concat feeding the JSMLP, and its adjoint.  HBM/L2-bound: every thread moves
// one float4; a node's neighbour list is walked in edge order (no atomics).
#include "common.cuh"

struct AggP {
  const float* in; int ld_in; float* out; int ld_out; const float* base; int ld_base;
  const float* yprev; int ld_yprev; int dact;
  int n, F4; const int* nd_graph; const int* nd_arena; const int* cum; int node_base;
  const int* nbr_ptr; const int* nbr;
};

__global__ void csr_aggregate_kernel(const AggP p) {
  pdl_sync();
  const long long total = (long long)p.n * p.F4;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx / p.F4), c = (int)(idx - (long long)i * p.F4) * 4;
    const int g = __ldg(p.nd_graph + i), a = __ldg(p.nd_arena + i);
    const int gstart = __ldg(p.cum + g) - p.node_base;
    const int e0 = __ldg(p.nbr_ptr + a), e1 = __ldg(p.nbr_ptr + a + 1);
    float4 acc = p.base ? *reinterpret_cast<const float4*>(p.base + (size_t)i * p.ld_base + c)
                        : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e = e0; e < e1; ++e) {
      const int j = gstart + __ldg(p.nbr + e);
      const float4 v = *reinterpret_cast<const float4*>(p.in + (size_t)j * p.ld_in + c);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    if (p.yprev) {
      const float4 y = *reinterpret_cast<const float4*>(p.yprev + (size_t)i * p.ld_yprev + c);
      acc.x *= act_grad_from_output(y.x, p.dact); acc.y *= act_grad_from_output(y.y, p.dact);
      acc.z *= act_grad_from_output(y.z, p.dact); acc.w *= act_grad_from_output(y.w, p.dact);
    }
    *reinterpret_cast<float4*>(p.out + (size_t)i * p.ld_out + c) = acc;
  }
}

extern "C" int sard_csr_aggregate(const float* in, int ld_in, float* out, int ld_out, const float* base, int ld_base,
                                  const float* yprev, int ld_yprev, int dact,
                                  int n, int F, const int* nd_graph, const int* nd_arena, const int* cum, int node_base,
                                  const int* nbr_ptr, const int* nbr, sard_stream_t stream) {
  if (n <= 0) return 0;
  SARD_REQUIRE(F % 4 == 0 && ld_in % 4 == 0 && ld_out % 4 == 0, "aggregate: feature width / strides must be multiples of 4");
  SARD_REQUIRE(!base || ld_base % 4 == 0, "aggregate: ld_base");
  SARD_REQUIRE(!yprev || ld_yprev % 4 == 0, "aggregate: ld_yprev");
  SARD_REQUIRE(((uintptr_t)in % 16 == 0) && ((uintptr_t)out % 16 == 0), "aggregate: 16-byte alignment");
  AggP p{in, ld_in, out, ld_out, base, ld_base, yprev, ld_yprev, dact, n, F / 4, nd_graph, nd_arena, cum, node_base, nbr_ptr, nbr};
  // algorithmic bytes: read the feature rows once, write them once, plus the index arrays (2 n F 4 + (n+1+e) 4,
  // e = 2(n - graphs) for trees ~ 2n)
  ProfScope prof(as_stream(stream), (double)n * F, 4.0 * ((2.0 + (base ? 1.0 : 0.0) + (yprev ? 1.0 : 0.0)) * (double)n * F + 5.0 * n));
  sard_launch(csr_aggregate_kernel, grid_for((long long)n * (F / 4), 256, 16), 256, 0, as_stream(stream), p);
  SARD_LAUNCH_OK("csr_aggregate_kernel");
  return 0;
}

// xs[p, :] = [ h[srow[p], 0:Fh] | ext[graph(srow[p]), 0:Fe] ]
__global__ void concat_sorted_kernel(const float* __restrict__ h, int ld_h, int Fh4, const float* __restrict__ ext,
                                     int ld_ext, int Fe4, const int* __restrict__ srow,
                                     const int* __restrict__ nd_graph, int n, float* __restrict__ xs, int ld_xs) {
  pdl_sync();
  const int W4 = Fh4 + Fe4;
  const long long total = (long long)n * W4;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int pos = (int)(idx / W4), c4 = (int)(idx - (long long)pos * W4);
    const int row = __ldg(srow + pos);
    float4 v;
    if (c4 < Fh4) v = *reinterpret_cast<const float4*>(h + (size_t)row * ld_h + c4 * 4);
    else v = *reinterpret_cast<const float4*>(ext + (size_t)__ldg(nd_graph + row) * ld_ext + (c4 - Fh4) * 4);
    *reinterpret_cast<float4*>(xs + (size_t)pos * ld_xs + c4 * 4) = v;
  }
}

extern "C" int sard_concat_sorted(const float* h, int ld_h, int Fh, const float* ext, int ld_ext, int Fe,
                                  const int* srow, const int* nd_graph, int n, float* xs, int ld_xs,
                                  sard_stream_t stream) {
  if (n <= 0) return 0;
  SARD_REQUIRE(Fh % 4 == 0 && Fe % 4 == 0 && ld_h % 4 == 0 && ld_ext % 4 == 0 && ld_xs % 4 == 0, "concat: multiples of 4");
  sard_launch(concat_sorted_kernel, grid_for((long long)n * ((Fh + Fe) / 4), 256, 16), 256, 0, as_stream(stream), 
      h, ld_h, Fh / 4, ext, ld_ext, Fe / 4, srow, nd_graph, n, xs, ld_xs);
  SARD_LAUNCH_OK("concat_sorted_kernel");
  return 0;
}

// (Fh+Fe)/4 threads per graph, one float4 column each: per-node rows go back to unsorted order, the ext
// part is summed over the graph's nodes in node order (deterministic segment sum)
__global__ void unsort_grad_kernel(const float* __restrict__ dxs, int ld_dxs, const int* __restrict__ spos,
                                   const int* __restrict__ cum, int node_base, int B,
                                   const float* __restrict__ h, int ld_h, int Fh, int h_act,
                                   float* __restrict__ dh, int ld_dh,
                                   const float* __restrict__ ext, int ld_ext, int Fe,
                                   float* __restrict__ dext, int ld_dext) {
  pdl_sync();
  const int W4 = (Fh + Fe) >> 2;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int g = (int)(idx / W4), c = (int)(idx - (long long)g * W4) * 4;
  if (g >= B) return;
  const int r0 = cum[g] - node_base, r1 = cum[g + 1] - node_base;
  // four nodes at a time: the position loads, then the row loads, are in flight together (a rolled loop pays two
  // dependent memory latencies per node)
  if (c < Fh) {
    for (int rb = r0; rb < r1; rb += 4) {
      int pos[4];
      float4 v[4], y[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) pos[u] = rb + u < r1 ? __ldg(spos + rb + u) : 0;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (rb + u < r1) {
          v[u] = *reinterpret_cast<const float4*>(dxs + (size_t)pos[u] * ld_dxs + c);
          y[u] = *reinterpret_cast<const float4*>(h + (size_t)(rb + u) * ld_h + c);
        }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (rb + u < r1) {
          v[u].x *= act_grad_from_output(y[u].x, h_act); v[u].y *= act_grad_from_output(y[u].y, h_act);
          v[u].z *= act_grad_from_output(y[u].z, h_act); v[u].w *= act_grad_from_output(y[u].w, h_act);
          *reinterpret_cast<float4*>(dh + (size_t)(rb + u) * ld_dh + c) = v[u];
        }
    }
  } else {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int rb = r0; rb < r1; rb += 4) {
      int pos[4];
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) pos[u] = rb + u < r1 ? __ldg(spos + rb + u) : 0;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        v[u] = rb + u < r1 ? *reinterpret_cast<const float4*>(dxs + (size_t)pos[u] * ld_dxs + c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 4; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }   // node order kept
    }
    const int ce = c - Fh;
    const float4 e = *reinterpret_cast<const float4*>(ext + (size_t)g * ld_ext + ce);
    acc.x = e.x > 0.f ? acc.x : 0.f; acc.y = e.y > 0.f ? acc.y : 0.f;
    acc.z = e.z > 0.f ? acc.z : 0.f; acc.w = e.w > 0.f ? acc.w : 0.f;
    *reinterpret_cast<float4*>(dext + (size_t)g * ld_dext + ce) = acc;
  }
}

extern "C" int sard_unsort_grad(const float* dxs, int ld_dxs, const int* spos, const int* cum, int node_base, int B,
                                const float* h, int ld_h, int Fh, int h_act, float* dh, int ld_dh,
                                const float* ext, int ld_ext, int Fe, float* dext, int ld_dext,
                                sard_stream_t stream) {
  if (B <= 0) return 0;
  SARD_REQUIRE(Fh % 4 == 0 && Fe % 4 == 0 && ld_dxs % 4 == 0 && ld_h % 4 == 0 && ld_dh % 4 == 0 && ld_ext % 4 == 0 && ld_dext % 4 == 0,
               "unsort_grad: widths / strides must be multiples of 4");
  sard_launch(unsort_grad_kernel, ceil_div((long long)B * ((Fh + Fe) / 4), 256), 256, 0, as_stream(stream), 
      dxs, ld_dxs, spos, cum, node_base, B, h, ld_h, Fh, h_act, dh, ld_dh, ext, ld_ext, Fe, dext, ld_dext);
  SARD_LAUNCH_OK("unsort_grad_kernel");
  return 0;
}
