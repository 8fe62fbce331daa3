// Fused policy/value losses: per-node log-prob, per-graph segment sum, ratio, clipped surrogate and
// the backward of all of it in one kernel (one warp per graph).  The orbit tie of
// transform_policy_output (transform2act_policy.py:421-473) is folded in: a tied column of node i is
// read from its representative's row (forward) and the gradients of all members are summed into the
// representative in node order (backward).  Per-graph statistics go to gstat and are reduced in row
// order by sard_loss_reduce, so results do not depend on scheduling.
#include "common.cuh"

#define MAXD 8
#define LOG_2PI 1.8378770664093453f

__device__ __forceinline__ float ppo_dlp(float lp, float fixed, float adv, float eps, float inv_B, float* surr_out) {
  const float ratio = expf(lp - fixed);
  const float s1 = ratio * adv;
  const float rc = fminf(fmaxf(ratio, 1.f - eps), 1.f + eps);
  const float s2 = rc * adv;
  *surr_out = fminf(s1, s2);
  // d min(s1,s2)/d ratio with torch's tie rule (half to each branch, the clamp passes inside its range)
  float dr;
  if (ratio >= 1.f - eps && ratio <= 1.f + eps) dr = adv;
  else dr = (s1 < s2) ? adv : 0.f;
  return -inv_B * dr * ratio;      // d(-mean surr)/d lp
}


// Second copy of a row's output gradient in the fused JSMLP backward's operand format (SardPolicyLoss.dout_sorted_hl):
// 32 columns of TF32-exact high parts | 32 columns of fp32 remainders at the row's position in the padded body-index
// order; columns >= d are zero.  Saves the separate formatting launch between the loss and the backward kernel.
__device__ __forceinline__ void store_sorted_hl(float* __restrict__ dst, const float* v /* MAXD */) {
  float4* p = reinterpret_cast<float4*>(dst);
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    float4 h, l;
    h.x = __uint_as_float(__float_as_uint(v[4 * q + 0]) & 0xffffe000u); l.x = v[4 * q + 0] - h.x;
    h.y = __uint_as_float(__float_as_uint(v[4 * q + 1]) & 0xffffe000u); l.y = v[4 * q + 1] - h.y;
    h.z = __uint_as_float(__float_as_uint(v[4 * q + 2]) & 0xffffe000u); l.z = v[4 * q + 2] - h.z;
    h.w = __uint_as_float(__float_as_uint(v[4 * q + 3]) & 0xffffe000u); l.w = v[4 * q + 3] - h.w;
    p[q] = h; p[8 + q] = l;
  }
#pragma unroll
  for (int q = 2; q < 8; ++q) { p[q] = z; p[8 + q] = z; }
}
// blocks past the per-graph ones: zero rows at the padding positions of the sorted order (one position per thread: a
// single dependent load per thread instead of a strided loop of them)
__device__ __forceinline__ void zero_sorted_pads(const SardPolicyLoss& a, int first_block) {
  const int pos = ((int)blockIdx.x - first_block) * (int)blockDim.x + (int)threadIdx.x;
  if (pos >= a.rows_cap || __ldg(a.srow + pos) >= 0) return;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  float4* d = reinterpret_cast<float4*>(a.dout_sorted_hl + (size_t)pos * 64);
#pragma unroll
  for (int q = 0; q < 16; ++q) d[q] = z;
}

__global__ void gauss_loss_kernel(const SardPolicyLoss a) {
  pdl_sync();
  const int graph_blocks = (a.B * 32 + (int)blockDim.x - 1) / (int)blockDim.x;
  if ((int)blockIdx.x >= graph_blocks) { zero_sorted_pads(a, graph_blocks); return; }
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (g >= a.B) return;
  const int gid = a.ids[g];
  const int r0 = a.cum[g] - a.node_base, N = a.cum[g + 1] - a.cum[g];
  const int d = a.d;
  float ls[MAXD], ivar[MAXD], S[MAXD];
#pragma unroll
  for (int c = 0; c < MAXD; ++c) {
    ls[c] = c < d ? __ldg(a.log_std + c) : 0.f;
    ivar[c] = expf(-2.f * ls[c]);
    S[c] = 0.f;
  }
  const bool tie = a.rep_local != nullptr && a.tie_c1 > a.tie_c0;
  float lp = 0.f;
  for (int i = lane; i < N; i += 32) {
    const int row = r0 + i, an = a.nd_arena[row];
    const int rrow = tie ? r0 + a.rep_local[an] : row;
#pragma unroll
    for (int c = 0; c < MAXD; ++c) {
      if (c >= d) break;
      const int src = (tie && c >= a.tie_c0 && c < a.tie_c1) ? rrow : row;
      const float mu = a.out[(size_t)src * a.ld_out + c];
      const float x = __ldg(a.actions + (size_t)an * SARD_ACTION_DIM + a.act_col + c);
      const float z2 = (x - mu) * (x - mu) * ivar[c];
      lp += -0.5f * z2 - ls[c] - 0.5f * LOG_2PI;
      S[c] += z2 - 1.f;
      if (a.tied_out) a.tied_out[(size_t)row * a.ld_out + c] = mu;
    }
  }
  lp = warp_sum(lp);
  if (lane == 0) {
    if (a.log_prob_out) a.log_prob_out[gid] = lp;
    if (a.log_prob_run) a.log_prob_run[g] = lp;
  }
  if (!a.advantages) return;                    // eval mode
  const float fixed = a.fixed_log_probs[gid];
  float surr;
  const float coef = ppo_dlp(lp, fixed, a.advantages[gid], a.clip_eps, a.inv_B, &surr);
  const int W = 4 + d;
#pragma unroll
  for (int c = 0; c < MAXD; ++c) {
    if (c >= d) break;
    const float s = warp_sum(S[c]);
    if (lane == 0) a.gstat[(size_t)g * W + 4 + c] = coef * s;
  }
  if (lane == 0) {
    a.gstat[(size_t)g * W + 0] = surr;
    a.gstat[(size_t)g * W + 1] = fabsf(fixed - lp);
    a.gstat[(size_t)g * W + 2] = 0.f;
    a.gstat[(size_t)g * W + 3] = 0.f;
  }
  if (!a.dout && !a.dout_sorted_hl) return;
  for (int j = lane; j < N; j += 32) {
    const int row = r0 + j, an = a.nd_arena[row];
    float gv[MAXD];
#pragma unroll
    for (int c = 0; c < MAXD; ++c) gv[c] = 0.f;
#pragma unroll
    for (int c = 0; c < MAXD; ++c) {
      if (c >= d) break;
      const float mu = a.out[(size_t)row * a.ld_out + c];
      float dmu;
      if (tie && c >= a.tie_c0 && c < a.tie_c1) {
        dmu = 0.f;
        for (int i = 0; i < N; ++i) {
          const int ai = a.nd_arena[r0 + i];
          if (a.rep_local[ai] == j)
            dmu += coef * (__ldg(a.actions + (size_t)ai * SARD_ACTION_DIM + a.act_col + c) - mu) * ivar[c];
        }
      } else {
        dmu = coef * (__ldg(a.actions + (size_t)an * SARD_ACTION_DIM + a.act_col + c) - mu) * ivar[c];
      }
      if (a.pre) dmu *= 1.57079632679489662f * cospif(0.5f * a.pre[(size_t)row * a.ld_out + c]);
      if (a.dout) a.dout[(size_t)row * a.ld_dout + c] = dmu;
      gv[c] = dmu;
    }
    if (a.dout_sorted_hl) store_sorted_hl(a.dout_sorted_hl + (size_t)__ldg(a.spos + row) * 64, gv);
  }
}

extern "C" int sard_gauss_loss(const SardPolicyLoss* a, sard_stream_t stream) {
  if (a->B <= 0) return 0;
  SARD_REQUIRE(a->d >= 1 && a->d <= MAXD, "gauss loss: action dims");
  SARD_REQUIRE(a->log_std != nullptr, "gauss loss: log_std");
  SARD_REQUIRE(!a->dout_sorted_hl || (a->spos && a->srow && a->rows_cap > 0 && a->advantages), "gauss loss: sorted gradient copy");
  sard_launch(gauss_loss_kernel, ceil_div((long long)a->B * 32, 128) + (a->dout_sorted_hl ? ceil_div(a->rows_cap, 128) : 0), 128, 0,
              as_stream(stream), *a);
  SARD_LAUNCH_OK("gauss_loss_kernel");
  return 0;
}

__global__ void cat_loss_kernel(const SardPolicyLoss a) {
  pdl_sync();
  const int graph_blocks = (a.B * 32 + (int)blockDim.x - 1) / (int)blockDim.x;
  if ((int)blockIdx.x >= graph_blocks) { zero_sorted_pads(a, graph_blocks); return; }
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (g >= a.B) return;
  const int gid = a.ids[g];
  const int r0 = a.cum[g] - a.node_base, N = a.cum[g + 1] - a.cum[g];
  const int d = a.d;
  const bool tie = a.rep_local != nullptr && a.tie_c1 > a.tie_c0;
  float lp = 0.f, ent = 0.f;
  for (int i = lane; i < N; i += 32) {
    const int row = r0 + i, an = a.nd_arena[row];
    const int rrow = tie ? r0 + a.rep_local[an] : row;
    float l[MAXD], mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < MAXD; ++c) {
      if (c >= d) break;
      const int src = (tie && c >= a.tie_c0 && c < a.tie_c1) ? rrow : row;
      l[c] = a.out[(size_t)src * a.ld_out + c];
      mx = fmaxf(mx, l[c]);
      if (a.tied_out) a.tied_out[(size_t)row * a.ld_out + c] = l[c];
    }
    float se = 0.f;
#pragma unroll
    for (int c = 0; c < MAXD; ++c) if (c < d) se += expf(l[c] - mx);
    const float lse = mx + logf(se);
    const int act = (int)__ldg(a.actions + (size_t)an * SARD_ACTION_DIM + a.act_col);
#pragma unroll
    for (int c = 0; c < MAXD; ++c) {
      if (c >= d) break;
      const float logp = l[c] - lse;
      if (c == act) lp += logp;
      ent -= expf(logp) * logp;
    }
  }
  lp = warp_sum(lp);
  ent = warp_sum(ent);
  if (lane == 0) {
    if (a.log_prob_out) a.log_prob_out[gid] = lp;
    if (a.log_prob_run) a.log_prob_run[g] = lp;
  }
  if (!a.advantages) return;
  const float fixed = a.fixed_log_probs[gid];
  float surr;
  const float coef = ppo_dlp(lp, fixed, a.advantages[gid], a.clip_eps, a.inv_B, &surr);
  const int W = 4 + d;
  if (lane == 0) {
    a.gstat[(size_t)g * W + 0] = surr;
    a.gstat[(size_t)g * W + 1] = fabsf(fixed - lp);
    a.gstat[(size_t)g * W + 2] = ent;
    a.gstat[(size_t)g * W + 3] = 0.f;
    for (int c = 0; c < d; ++c) a.gstat[(size_t)g * W + 4 + c] = 0.f;
  }
  if (!a.dout && !a.dout_sorted_hl) return;
  for (int j = lane; j < N; j += 32) {
    const int row = r0 + j, an = a.nd_arena[row];
    float l[MAXD], mx = -INFINITY, se = 0.f;
    float gv[MAXD];
#pragma unroll
    for (int c = 0; c < MAXD; ++c) gv[c] = 0.f;
#pragma unroll
    for (int c = 0; c < MAXD; ++c) {
      if (c >= d) break;
      l[c] = a.out[(size_t)row * a.ld_out + c];
      mx = fmaxf(mx, l[c]);
    }
#pragma unroll
    for (int c = 0; c < MAXD; ++c) if (c < d) se += expf(l[c] - mx);
    const float lse = mx + logf(se);
    // columns tied to this row collect every member of the orbit; untied columns only the row itself
    float cnt_t = 0.f, hot_t[MAXD];
#pragma unroll
    for (int c = 0; c < MAXD; ++c) hot_t[c] = 0.f;
    if (tie) {
      for (int i = 0; i < N; ++i) {
        const int ai = a.nd_arena[r0 + i];
        if (a.rep_local[ai] == j) {
          cnt_t += 1.f;
          const int act = (int)__ldg(a.actions + (size_t)ai * SARD_ACTION_DIM + a.act_col);
#pragma unroll
          for (int c = 0; c < MAXD; ++c) if (c == act) hot_t[c] += 1.f;
        }
      }
    }
    const int act_j = (int)__ldg(a.actions + (size_t)an * SARD_ACTION_DIM + a.act_col);
    // when the row is itself tied to another representative its own logits are unused for tied columns
    const bool self_rep = !tie || a.rep_local[an] == j;
#pragma unroll
    for (int c = 0; c < MAXD; ++c) {
      if (c >= d) break;
      const float p = expf(l[c] - lse);
      float gval;
      if (tie && c >= a.tie_c0 && c < a.tie_c1) gval = coef * (hot_t[c] - cnt_t * p);
      else gval = coef * ((c == act_j ? 1.f : 0.f) - p);
      (void)self_rep;
      if (a.dout) a.dout[(size_t)row * a.ld_dout + c] = gval;
      gv[c] = gval;
    }
    if (a.dout_sorted_hl) store_sorted_hl(a.dout_sorted_hl + (size_t)__ldg(a.spos + row) * 64, gv);
  }
}

extern "C" int sard_cat_loss(const SardPolicyLoss* a, sard_stream_t stream) {
  if (a->B <= 0) return 0;
  SARD_REQUIRE(a->d >= 2 && a->d <= MAXD, "cat loss: number of actions");
  SARD_REQUIRE(!a->dout_sorted_hl || (a->spos && a->srow && a->rows_cap > 0 && a->advantages), "cat loss: sorted gradient copy");
  sard_launch(cat_loss_kernel, ceil_div((long long)a->B * 32, 128) + (a->dout_sorted_hl ? ceil_div(a->rows_cap, 128) : 0), 128, 0,
              as_stream(stream), *a);
  SARD_LAUNCH_OK("cat_loss_kernel");
  return 0;
}

// one block per column, rows strided over threads, fixed-order tree: deterministic
__global__ void loss_reduce_kernel(const float* __restrict__ gstat, int B, int W, float* __restrict__ stats,
                                   float* __restrict__ dlogstd, int d, const float* __restrict__ log_std) {
  pdl_sync();
  __shared__ float red[256];
  const int c = blockIdx.x;
  float s = 0.f;
  for (int b = threadIdx.x; b < B; b += 256) s += gstat[(size_t)b * W + c];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (c < 3) stats[c] += red[0];
    else if (c >= 4 && dlogstd && c - 4 < d) dlogstd[c - 4] += red[0];
    if (c == 0 && log_std) {
      float e = 0.f;
      for (int j = 0; j < d; ++j) e += 0.5f + 0.5f * LOG_2PI + log_std[j];
      stats[4] += e / (float)d;
    }
  }
}

extern "C" int sard_loss_reduce(const float* gstat, int B, int W, float* stats, float* dlogstd, int d,
                                const float* log_std_for_entropy, sard_stream_t stream) {
  if (W <= 0) return 0;
  sard_launch(loss_reduce_kernel, W, 256, 0, as_stream(stream), gstat, B, W, stats, dlogstd, d, log_std_for_entropy);
  SARD_LAUNCH_OK("loss_reduce_kernel");
  return 0;
}

__global__ void value_loss_kernel(const float* __restrict__ v, const int* __restrict__ ids, int B,
                                  const float* __restrict__ returns, float inv_B, float* __restrict__ dv,
                                  float* __restrict__ sq_err, float* __restrict__ values_out) {
  pdl_sync();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int gid = ids[b];
  const float val = v[b];
  if (values_out) values_out[gid] = val;
  if (returns) {
    const float e = val - returns[gid];
    if (dv) dv[b] = 2.f * e * inv_B;
    if (sq_err) sq_err[b] = e * e;
  }
}

extern "C" int sard_value_loss(const float* v, const int* ids, int B, const float* returns, float inv_B,
                               float* dv, float* sq_err, float* values_out, sard_stream_t stream) {
  if (B <= 0) return 0;
  sard_launch(value_loss_kernel, ceil_div(B, 256), 256, 0, as_stream(stream), v, ids, B, returns, inv_B, dv, sq_err, values_out);
  SARD_LAUNCH_OK("value_loss_kernel");
  return 0;
}


// Value head, loss and the head's backward-data in one launch (one warp per root row):
//   v = cat . w + b                                   (transform2act_critic.py:107-108, value_head)
//   e = v - returns[gid],  dv = 2 e / B_global,  sq_err = e^2          (agent_pg.py:21-22, MSE)
//   dcat[c] = dv * w[c] * act'(cat[c])     act' = act0 on the first dsplit columns (MLP output), ReLU on the rest (ObsEncoders)
// Replaces three dependent launches (skinny forward, value_loss, skinny backward) on the critic's chain.
__global__ void __launch_bounds__(256) value_head_kernel(const float* __restrict__ cat, int ldcat, const float* __restrict__ w,
                                                         const float* __restrict__ bias, int B, int K, int dsplit, int act0,
                                                         const int* __restrict__ ids, const float* __restrict__ returns, float inv_B,
                                                         float* __restrict__ v_out, float* __restrict__ dv_out, float* __restrict__ sq_err,
                                                         float* __restrict__ values_out, float* __restrict__ dcat, int ld_dcat) {
  pdl_sync();
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b >= B) return;
  const float* row = cat + (size_t)b * ldcat;
  float s = 0.f;
  for (int c = lane; c < K; c += 32) s = fmaf(row[c], __ldg(w + c), s);
  s = warp_sum(s) + __ldg(bias);
  const int gid = ids[b];
  if (lane == 0) {
    if (v_out) v_out[b] = s;
    if (values_out) values_out[gid] = s;
  }
  if (!returns) return;
  const float e = s - returns[gid];
  const float dv = 2.f * e * inv_B;
  if (lane == 0) {
    if (dv_out) dv_out[b] = dv;
    if (sq_err) sq_err[b] = e * e;
  }
  if (!dcat) return;
  float* drow = dcat + (size_t)b * ld_dcat;
  for (int c = lane; c < K; c += 32) {
    const float y = row[c];
    const float g = c < dsplit ? act_grad_from_output(y, act0) : (y > 0.f ? 1.f : 0.f);
    drow[c] = dv * __ldg(w + c) * g;
  }
}

extern "C" int sard_value_head(const float* cat, int ldcat, const float* w, const float* bias, int B, int K, int dsplit, int act0,
                               const int* ids, const float* returns, float inv_B, float* v, float* dv, float* sq_err,
                               float* values_out, float* dcat, int ld_dcat, sard_stream_t stream) {
  if (B <= 0) return 0;
  SARD_REQUIRE(K > 0 && ldcat >= K && (!dcat || ld_dcat >= K), "value head: widths");
  sard_prof_bytes(4.0 * B * (double)K * (dcat ? 3.0 : 1.0));
  sard_launch(value_head_kernel, ceil_div((long long)B * 32, 256), 256, 0, as_stream(stream), cat, ldcat, w, bias, B, K, dsplit, act0,
              ids, returns, inv_B, v, dv, sq_err, values_out, dcat, ld_dcat);
  SARD_LAUNCH_OK("value_head_kernel");
  return 0;
}
