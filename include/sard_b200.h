/*
 * sard_b200.h -- C-ABI of the B200-native SARD PPO-update hot path (libsard_b200.so).
 *
 * The reference (drdh/SARD) is pure Python and has no FFI of its own: the boundary it offers is
 * the Python call surface of design_opt/agents/transform2act_agent.py, design_opt/models/*.py and
 * khrylib/rl/core/*.py.  Every entry point below replaces the tensor work
 * behind one of those calls; the comment above each one cites the reference lines it replaces
 * (paths relative to the reference root).  INTEGRATION.md shows the ctypes binding a reference
 * maintainer would add.
 *
 * Conventions: plain C; every pointer is caller-owned DEVICE memory unless the comment says
 * "host"; sizes are host integers; `stream` is a cudaStream_t passed as void*; functions only
 * enqueue work on `stream` (no device allocation, no synchronisation); the return value is 0 on
 * success or a non-zero code, with sard_last_error_string() describing the failure.  fp32
 * arithmetic unless stated; integer layouts are int32.  All kernels are deterministic (no
 * floating-point atomics).
 *
 * Process-global state, all of it host side: the scheduling switches set through sard_set_* below
 * (GEMM engine, programmatic dependent launch, side-stream overlap, single-launch Adam; defaults
 * from SARD_* environment variables at load), the side streams / events the library creates once
 * per (device, launching stream) for sard_value_run / sard_policy_run, the per-device kernel
 * attribute flags, the profiling buffers of sard_prof_*, the launch counter and the thread-local
 * error string.  None of it changes results.
 */
#ifndef SARD_B200_H
#define SARD_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef void* sard_stream_t;

#define SARD_ABI_VERSION 1
#define SARD_MAX_LAYERS 8
#define SARD_MAX_MLP 4
#define SARD_MAX_IL 4
#define SARD_ACTION_DIM 8          /* [control(1) | attr(6) | skel(1)], transform2act_policy.py:66 */
/* row stride of one weight-gradient partial: N2 products + the bias column, padded to 16 bytes */
#define SARD_WG_LDPART(N2) (((N2) + 4) & ~3)
#define SARD_EXT_DIM 64            /* ObsEncoder out_dim, jsmlp.py:8 */

enum { SARD_ACT_NONE = 0, SARD_ACT_RELU = 1, SARD_ACT_TANH = 2, SARD_ACT_SIGMOID = 3 };
enum { SARD_POST_NONE = 0, SARD_POST_SIN_HALF_PI = 1 };  /* transform2act_policy.py:290 */
enum { SARD_STAGE_SKEL = 0, SARD_STAGE_ATTR = 1, SARD_STAGE_CONTROL = 2 };

int sard_abi_version(void);
const char* sard_last_error_string(void);
/* sizeof() of the ABI structs (SardJsmlpTc last, the others in declaration order), for binding self-checks (host memory, no GPU). */
int sard_struct_sizes(int* out, int capacity);

/* Kernels launched by the library since load (monotonic; host counter). */
long long sard_launch_count(void);

/* Programmatic dependent launch for every kernel of the library (default 1): a kernel's CTAs may be
 * scheduled while its predecessor on the stream drains; they wait for its completion before touching
 * memory, so results are unchanged.  0 = plain stream-ordered launches. */
int sard_set_pdl(int enabled);
int sard_get_pdl(void);

/* Weight gradients of sard_value_run / sard_policy_run on a low-priority side stream owned by the library
 * (default 1): forked after each layer's output gradient, joined before the call returns, so the
 * caller's stream semantics are unchanged.  0 = everything on the caller's stream. */
int sard_set_wgrad_overlap(int enabled);
/* sard_adam_update without a gradient norm as one launch (gate + bias corrections + Adam) instead of two; bit-identical,
 * off by default (measured neutral to slightly slower on the bench workload). */
int sard_set_adam_single_launch(int enabled);
int sard_get_wgrad_overlap(void);

/* Per-call-site kernel timing for the roofline numbers (CUDA events on the launching stream).
 * Sites label the GEMM / aggregate launches of the network schedules: site = 3*kind + pass with
 * kind in {0 in_fc, 1 graphconv, 2 index-linear, 3 obs-encoder, 4 critic mlp, 5 value head, 6 csr aggregate}
 * and pass in {0 forward, 1 backward-data / adjoint, 2 backward-weight}.
 * sard_prof_enable(mask, capacity): time launches whose site bit is set in mask (0 disables), keeping
 * at most `capacity` event pairs.  sard_prof_collect synchronises the recorded events and returns,
 * per site, {launches, total ms, algorithmic flops, algorithmic bytes} in rows[site*4 .. site*4+3]
 * (host doubles, n_sites rows), then clears the records. */
#define SARD_PROF_SITES 24
int sard_prof_enable(unsigned mask, int capacity);
int sard_prof_collect(double* rows, int n_sites);
/* Raw timeline of the recorded launches: rows[i*4..] = {site, stream ordinal, start ms, end ms} relative to the
 * first record; returns the number of rows written (<= capacity).  Call before sard_prof_collect. */
int sard_prof_timeline(double* rows, int capacity);
/* Every launch of the library by kernel name (diagnostic: events around each launch suppress the programmatic
 * overlap of consecutive kernels).  sard_prof_kernels_enable(capacity > 0) starts recording, (0) stops.
 * sard_prof_kernels_collect: names = newline-separated kernel names in id order; rows[i*6..] = {name id, stream
 * ordinal, start ms, end ms, algorithmic bytes, algorithmic flops (0 when the wrapper states none)}; returns the row
 * count, clears. */
int sard_prof_kernels_enable(int capacity);
int sard_prof_kernels_collect(char* names, int names_len, double* rows, int capacity);

/* =====================================================================================
 * Packing: graph indexing layouts (bit-exact integer work)
 * ===================================================================================== */

/* Per-node neighbour lists for every graph of a rollout arena, by destination (in_*) and by
 * source (out_*), neighbours stored as node ids LOCAL to the graph, in edge order.
 * Replaces the edge concatenation + offsetting of Transform2ActPolicy.batch_data
 * (design_opt/models/transform2act_policy.py:183-202) and the scatter inside
 * torch_geometric GraphConv.propagate (called at design_opt/models/gnn.py:63).
 * node_ptr,edge_ptr: [T+1]; esrc,edst: [e_total] local ids; in_ptr,out_ptr: [n_total+1];
 * in_nbr,out_nbr: [e_total]. */
int sard_build_csr(const int* node_ptr, const int* edge_ptr, const int* esrc, const int* edst, int T,
                   int* in_ptr, int* in_nbr, int* out_ptr, int* out_nbr, sard_stream_t stream);

/* Orbit representative (graph-local node id) of every node under the current design subgroup.
 * Replaces the index bookkeeping of Transform2ActPolicy.transform_policy_output
 * (transform2act_policy.py:421-436): level = #digits of body_ind in base `index_base`, leg = last
 * digit, representative leg = leg_rep[leg] (LearnedGroup.get_cur_orbits, group_action.py:322-325).
 * rep_local[i] = local id of the node of the same graph whose body index has the same leading
 * digits and the representative leg; *err_flag is set to 1 if some representative is missing
 * (the reference's reshape would raise). leg_rep: [index_base]. */
int sard_orbit_rep(const int* node_ptr, const int* body_ind, int T, const int* leg_rep, int index_base,
                   int* rep_local, int* err_flag, sard_stream_t stream);

/* Rollout arena resident in HBM for one update_params call (tensorfy, transform2act_agent.py:20-24,
 * :278-282, packed once instead of 5*T small copies). */
typedef struct SardArena {
  int T, n_total, e_total, d_obs;
  const float* obs;            /* [n_total, d_obs] node observations */
  const float* actions;        /* [n_total, SARD_ACTION_DIM] */
  const int* node_ptr;         /* [T+1] */
  const int* stage;            /* [T] 0 skel / 1 attr / 2 execution */
  const int* body_ind;         /* [n_total] */
  const int *in_ptr, *in_nbr, *out_ptr, *out_nbr;   /* sard_build_csr */
  const int* rep_local;        /* [n_total] sard_orbit_rep */
  const float* ext[3];         /* hfield [T,ext_dim[0]], obj [T,ext_dim[1]], goal [T,ext_dim[2]] */
  int ext_dim[3];
  float* values;               /* [T] critic pre-pass output (agent :283-291) */
  float* returns;              /* [T] */
  float* advantages;           /* [T] */
  float* fixed_log_probs;      /* [T] (agent :316-325) */
  int max_nodes;               /* largest graph of the arena (0 = unknown: the fused tower kernels are not used) */
} SardArena;

/* A run of graphs of the current ordering (one minibatch, or one stage of it). */
typedef struct {
  const int* ids;              /* [B] arena graph ids, in order */
  const int* cum;              /* [B+1] running node count of the ordering; graph k starts at cum[k]-node_base */
  int B;                       /* graphs in the run */
  int n;                       /* nodes in the run (host-known: cum[B]-cum[0]) */
  int node_base;               /* cum[0] */
} SardSel;

/* Gather one run into contiguous node-major buffers.  Replaces torch.cat(obs) + column selection +
 * embed / design-flag concatenation (transform2act_policy.py:188,266-269,310-313;
 * transform2act_critic.py:63,80-88).
 * x[i,:] = [obs[:,0:lead] | obs[:,d_obs-tail:d_obs] | consts[0:nconst] | onehot3(stage) if onehot];
 * columns beyond the written ones up to ldx are zeroed. */
typedef struct {
  int lead, tail, nconst, onehot;
  const float* consts;         /* [nconst] subgroup embedding (LearnedGroup.get_cur_subgroup_embeds) */
  float* x; int ldx;           /* [n, ldx] */
  int* nd_graph;               /* [n] run-local graph index of each node */
  int* nd_arena;               /* [n] arena node id of each node */
  int* body;                   /* [n] body index of each node */
  int* root_row;               /* [B] row of each graph's first node (critic readout, critic.py:104-107) */
} SardGatherOut;
int sard_gather_nodes(const SardArena* arena, const SardSel* sel, const SardGatherOut* out, sard_stream_t stream);

/* Stable counting sort of the rows of a run by body index, plus the row-tile tables the grouped
 * IndexLinear GEMMs consume.  Replaces `ind.unique()` + boolean masks of IndexLinear.forward
 * (design_opt/models/jsmlp.py:33-39).
 * srow[p] = row at sorted position p; spos[row] = sorted position; grp_ptr[u]..grp_ptr[u+1] = rows of
 * index u; tiles[t] = {row_start,row_end,index,0} with <= tile_rows rows (same for chunks with
 * chunk_rows); grp_chunk_ptr[u] = first chunk of index u.  counts[0]=#tiles, counts[1]=#chunks.
 * scratch: int[(ceil(n/256)+2)*max_index]. */
int sard_sort_by_index(const int* body, int n, int max_index, int tile_rows, int chunk_rows,
                       int* srow, int* spos, int* grp_ptr, int* tiles, int* chunks, int* grp_chunk_ptr,
                       int* counts, int* scratch, sard_stream_t stream);
/* As above with every index starting on a tile boundary (what the fused tensor-core JSMLP kernels consume: all
 * tiles are whole, the unused positions behind an index's rows are padding).  srow has srow_capacity >=
 * n + (number of present indices) * (tile_rows - 1) entries and is -1 at padding positions; counts[3] = positions
 * in use; chunk_rows must be a multiple of tile_rows; row_end of a tile / chunk stays the end of the real rows. */
int sard_sort_by_index_padded(const int* body, int n, int max_index, int tile_rows, int chunk_rows, int srow_capacity,
                              int* srow, int* spos, int* grp_ptr, int* tiles, int* chunks, int* grp_chunk_ptr,
                              int* counts, int* scratch, sard_stream_t stream);

/* =====================================================================================
 * RunningNorm  (khrylib/rl/core/running_norm.py:22-42)
 * ===================================================================================== */

/* Batch moments of x[n,D] as one record {count, mean[D], M2[D]} of doubles (shifted sums per
 * 128-row chunk merged pairwise in fixed order; deterministic).  `partials`: float[2*D*ceil(n/128)]. */
int sard_norm_moments(const float* x, int ldx, int n, int D, float* partials, double* record, sard_stream_t stream);
/* Merge `n_records` records (ranks) in order, then apply RunningNorm.update (:22-29): biased batch
 * variance, weight w = n_old/(m+n_old).  Writes mean,var,std [D] fp32, *count (int64) += m and
 * inv[d] = 1/(std[d]+1e-8). */
int sard_norm_update(const double* records, int n_records, int D, float* mean, float* var, float* std,
                     long long* count, float* inv, sard_stream_t stream);
/* inv[d] = 1/(std[d]+1e-8) from the stored statistics (eval mode). */
int sard_norm_prepare_eval(const float* std, int D, float* inv, sard_stream_t stream);
/* x = clamp((x-mean)*inv, -clip, clip) in place when *count > 0 (:34-41). */
int sard_norm_apply(float* x, int ldx, int n, int D, const float* mean, const float* inv,
                    const long long* count, float clip, sard_stream_t stream);

/* =====================================================================================
 * Message passing and segment ops
 * ===================================================================================== */

/* Deterministic CSR gather/reduce: out[i,:] = (base? base[i,:] : 0) + sum_{j in nbr(i)} in[gstart(i)+j,:],
 * optionally multiplied by act'(yprev[i,:]).  Forward aggregate of GraphConv (aggr='add', gnn.py:63)
 * with the by-destination lists, its adjoint with the by-source lists.  F % 4 == 0, lds % 4 == 0. */
int sard_csr_aggregate(const float* in, int ld_in, float* out, int ld_out, const float* base, int ld_base,
                       const float* yprev, int ld_yprev, int dact,
                       int n, int F, const int* nd_graph, const int* nd_arena, const int* cum, int node_base,
                       const int* nbr_ptr, const int* nbr, sard_stream_t stream);

/* xs[p,:] = [h[srow[p], 0:Fh] | ext[nd_graph[srow[p]], 0:Fe]]: the torch.cat([x, x_ext]) +
 * repeat_interleave of transform2act_policy.py:204-213,242-243, in body-index-sorted row order. */
int sard_concat_sorted(const float* h, int ld_h, int Fh, const float* ext, int ld_ext, int Fe,
                       const int* srow, const int* nd_graph, int n, float* xs, int ld_xs, sard_stream_t stream);

/* Adjoint of sard_concat_sorted: dh[row,:] = dxs[spos[row],0:Fh] * act'(h[row,:]);
 * dext[g,:] = relu'(ext[g,:]) * sum_{rows of g} dxs[spos[row],Fh:Fh+Fe]. */
int sard_unsort_grad(const float* dxs, int ld_dxs, const int* spos, const int* cum, int node_base, int B,
                     const float* h, int ld_h, int Fh, int h_act, float* dh, int ld_dh,
                     const float* ext, int ld_ext, int Fe, float* dext, int ld_dext, sard_stream_t stream);

/* =====================================================================================
 * Linear layers (dense, GraphConv-concatenated, and per-body-index grouped)
 * ===================================================================================== */
typedef struct {
  /* A operand [M, R]: row m read at A[(a_rows? a_rows[m] : m) * lda + r] */
  const float* A; int lda; const int* a_rows;
  /* weights: forward  W[c, r] = r < b_split ? B0[c*ldb0 + r] : B1[c*ldb1 + r-b_split]
   *          backward-data uses the same storage transposed:  W[r, c] with the split on c */
  const float* B0; int ldb0; const float* B1; int ldb1; int b_split;
  long long b_group_stride;    /* floats between the weight slabs of consecutive body indices (0: dense) */
  const float* bias; int bias_group_stride;
  float* Y; int ldy; const int* y_rows;      /* output row m written at (y_rows? y_rows[m] : m) */
  float* Ypre;                 /* optional copy of the value before `post` (same layout as Y) */
  const float* Dact; int ldd;  /* backward-data: Y *= act'(Dact[yrow, c]); dact0 for c < dsplit, dact1 after */
  int act, post, dact0, dact1, dsplit;
  int M, N, R;                 /* rows, output columns, reduction length */
  const int* tiles;            /* optional int4 tiles {row_start,row_end,index,0} from sard_sort_by_index */
  const int* num_tiles;        /* device count of tiles */
  int max_tiles;               /* host upper bound (grid size) */
  /* Optional fused GraphConv neighbour sum (aggr='add', gnn.py:63).  With agg_cols > 0 the A operand is virtual:
   *   A[m, r] = r < agg_cols ? sum_{j in nbr(m)} S[j, r] : S[m, r - agg_cols],   S = A with leading dimension lda,
   * nbr(m) = rows gstart(m) + agg_nbr[e], e in [agg_ptr[a], agg_ptr[a+1]) (edge order), a = agg_node[m] the arena
   * node of row m, gstart(m) = agg_cum[agg_graph[m]] - agg_node_base the first row of m's graph.  agg_out
   * (optional) receives the sums, [M, agg_cols] with leading dimension ld_agg_out.  Forward: S = x, weights
   * [lin_l | lin_r].  Backward-data: S = dz with the by-source lists and b_ksplit > 0, which splits the weights on
   * the reduction index:  W[r, c] = r < b_ksplit ? B0[r*ldb0 + c] : B1[(r - b_ksplit)*ldb1 + c].
   * Needs R, agg_cols multiples of 16, N a multiple of 64, a_rows == NULL (sard_gemm_agg_supported). */
  int agg_cols; const int* agg_ptr; const int* agg_nbr; const int* agg_node; const int* agg_graph; const int* agg_cum;
  int agg_node_base; float* agg_out; int ld_agg_out; int b_ksplit;
  /* Optional split-K scratch (ungrouped, unfused launches with few rows and a long reduction, i.e. the critic
   * MLP): when given, the reduction may be cut into S <= 8 slices whose raw partial tiles [S, M, N] go here and
   * are summed in slice order by a finishing kernel that applies the epilogue.  splitk_floats = capacity. */
  float* splitk_ws; long long splitk_floats;
} SardGemm;

/* Engine behind sard_linear_fwd / sard_linear_bwd_data: 0 = fp32 FFMA SIMT tiles, 1 = tcgen05 tensor cores
 * with error-compensated 3xTF32 (hi/lo split, fp32 accumulation in TMEM) for N > 16 and R >= 8. */
int sard_set_gemm_engine(int engine);
int sard_get_gemm_engine(void);

/* Y = post(act(A W^T + bias)).  Replaces nn.Linear / GraphConv lin_l+lin_r / IndexLinear.forward
 * (gnn.py:57,63; jsmlp.py:15-20,32-40; mlp.py:22-25; critic.py:103). */
int sard_linear_fwd(const SardGemm* g, sard_stream_t stream);
/* 1 when a GraphConv layer h_in -> h_out can use the fused neighbour sum (both directions). */
int sard_gemm_agg_supported(int h_in, int h_out);
/* dA = (dY W) * act'(Dact)   (autograd of the same call sites). */
int sard_linear_bwd_data(const SardGemm* g, sard_stream_t stream);

typedef struct {
  const float* P; int ldp; const int* p_rows;    /* dZ [M, N1] */
  const float* Q; int ldq; const int* q_rows;    /* layer input [M, N2] */
  int M, N1, N2;
  const int* chunks; const int* num_chunks; int max_chunks;   /* optional int4 {row_start,row_end,index,0} */
  int chunk_rows;                                /* rows per chunk when chunks == NULL (0: automatic) */
  const int* grp_chunk_ptr; int n_groups;        /* grouped: chunks of index u are [ptr[u], ptr[u+1]) */
  const int* slot_of_group;                      /* grouped: destination slot of each index (-1: skip) */
  float* dW0; int ldw0; float* dW1; int ldw1; int w_split;   /* dW[c1, c2] destination, split on c2 */
  long long w_slot_stride;                       /* floats between slots (grouped) */
  float* db; int b_slot_stride;                  /* optional bias grad [N1] (per slot when grouped) */
  float* partials;                               /* scratch float[max_chunks * N1 * SARD_WG_LDPART(N2)] */
} SardWGrad;
/* dW += dZ^T A, db += colsum(dZ): two kernels (per-chunk partial products, ordered reduction).
 * Ungrouped calls may pass chunk_rows = 0 to let the library pick sard_wgrad_chunk_rows(M,N1,N2)
 * (the partials scratch must then hold ceil(M/that) * N1 * SARD_WG_LDPART(N2) floats). */
int sard_wgrad_chunk_rows(int M, int N1, int N2);
int sard_linear_bwd_weight(const SardWGrad* w, sard_stream_t stream);

/* =====================================================================================
 * Fused JSMLP on the tcgen05 tensor cores (sm_100a): the three IndexLinear layers of JSMLP.forward
 * (design_opt/models/jsmlp.py:32-40, 74-79) in one launch per direction, 3xTF32 error-compensated,
 * TMA-fed, hidden activations handed from layer to layer through tensor memory.
 * Rows are in body-index-sorted order (sard_sort_by_index with tile_rows = 128).
 * ===================================================================================== */
typedef struct {
  int n, in_dim, hid, out_dim, act, post, max_index, n_slots;
  const float* xs; int ld_xs;            /* [n, in_dim] sorted input rows (sard_jsmlp_tc_split_input reads it) */
  const float* W[3]; const float* b[3];  /* IndexLinear tables W[max_index, out, in], b[max_index, out] */
  const int* slot_of_index;              /* [max_index] compact slot of each live body index, -1 otherwise */
  const int* tiles; const int* num_tiles; int max_tiles;    /* sard_sort_by_index */
  const int* srow;                       /* [n] destination row of sorted position p for `out` / `pre` (NULL: p) */
  float* H[2]; int ldh;                  /* hidden activations tanh(IL1), tanh(IL2), raw fp32 [n, hid], sorted rows (may be NULL) */
  float* Hhl[2];                         /* optional [n, 2*hid] pair copies (hi | lo) for the tensor-core weight gradients */
  float* out; float* pre; int ldo;       /* IL3 output after `post`, optional value before it */
  float* ws; long long ws_floats;        /* scratch: pair copies of the live weight slabs and of xs (256-byte aligned) */
  int padded;                            /* 1: rows follow sard_sort_by_index_padded (n = row capacity, a multiple of 128; srow -1 at padding) */
} SardJsmlpTc;
/* 1 when the fused kernels cover JSMLP(in_dim -> h1 -> h2 -> out_dim) with n_il layers. */
int sard_jsmlp_tc_supported(int in_dim, int h1, int h2, int out_dim, int n_il);
long long sard_jsmlp_tc_workspace(int n, int in_dim, int out_dim, int n_slots);
/* hi/lo pair copies of the live weight slabs (call after every optimiser step, before forward / backward). */
int sard_jsmlp_tc_presplit(const SardJsmlpTc* a, sard_stream_t stream);
/* hi/lo pair copy of xs into the scratch; sard_jsmlp_tc_xs_hl returns where it lives ([n, 2*in_dim]). */
int sard_jsmlp_tc_split_input(const SardJsmlpTc* a, sard_stream_t stream);
float* sard_jsmlp_tc_xs_hl(const SardJsmlpTc* a);
int sard_jsmlp_tc_forward(const SardJsmlpTc* a, sard_stream_t stream);
/* Backward of the fused JSMLP (padded row space only; in_dim 256, hid 128).  All pair arrays are [n, hi | lo]. */
typedef struct {
  const float* dout; int ld_dout;        /* [rows, ld] gradient w.r.t. IL3's output before `post`, ORIGINAL row order (srow maps);
                                          * NULL: dOs_hl was already filled by the loss kernel (SardPolicyLoss.dout_sorted_hl) */
  float* dOs_hl;                         /* [n, 64] scratch: dout at sorted positions, 32 columns hi | 32 lo */
  float* dZhl[2];                        /* [n, 2*hid] out: gradients w.r.t. the pre-activations of IL1 ([0]) and IL2 ([1]) */
  float* dxs; int ld_dxs;                /* [n, in_dim] out: gradient w.r.t. the sorted input rows (raw fp32) */
  const int* chunks; const int* num_chunks; int max_chunks; const int* grp_chunk_ptr;   /* sard_sort_by_index_padded */
  float* partials;                       /* scratch float[sard_jsmlp_tc_partials_floats(max_chunks)] */
  float* gW[3]; float* gb[3];            /* compact gradients, accumulated: gW[j][slot][out][in], gb[j][slot][out] */
} SardJsmlpTcBwd;
int sard_jsmlp_tc_bwd_supported(int in_dim, int h1, int h2, int out_dim, int n_il);
long long sard_jsmlp_tc_partials_floats(int max_chunks);
/* dZ2, dZ1 (pair arrays) and dxs from dout: one launch (+ the small re-ordering kernel for dout). */
int sard_jsmlp_tc_backward(const SardJsmlpTc* a, const SardJsmlpTcBwd* b, sard_stream_t stream);
/* dW / db of the three tables for the live body indices: one launch over row chunks + the ordered reduction. */
int sard_jsmlp_tc_wgrad(const SardJsmlpTc* a, const SardJsmlpTcBwd* b, sard_stream_t stream);
/* [h[row] | ext[graph(row)]] of the row at every sorted position, split hi | lo, zeros at padding positions
 * (transform2act_policy.py:204-213 in the operand format of the fused kernels): xs_hl [rows_cap, 2*(Fh+Fe)]. */
int sard_concat_sorted_split(const float* h, int ld_h, int Fh, const float* ext, int ld_ext, int Fe, const int* srow,
                             const int* nd_graph, int rows_cap, float* xs_hl, sard_stream_t stream);
/* Tuning aid: clock64() stamps written by block 0 of the last fused launch (host array of n <= 32 values). */
int sard_tc_debug_read(long long* out, int n);

/* =====================================================================================
 * Losses (fused forward + backward)
 * ===================================================================================== */
typedef struct {
  const float* out; int ld_out;        /* JSMLP output per node: mean [n,d] (after sin) or logits [n,3] */
  const float* pre;                    /* attr: value before sin (same ld), else NULL */
  const float* log_std;                /* [d] gaussian stages */
  int d;                               /* action dims of the stage: 1 / 6 / 3 logits */
  int act_col;                         /* first action column of the stage inside the 8-wide action row */
  int tie_c0, tie_c1;                  /* orbit-tied output columns [c0,c1) (transform2act_policy.py:466-473) */
  const int* nd_arena; const int* cum; int node_base; const int* ids; int B;
  const float* actions;                /* arena [n_total, 8] */
  const int* rep_local;                /* arena, NULL when untied */
  const float* advantages;             /* arena [T] (NULL in eval mode) */
  const float* fixed_log_probs;        /* arena [T] (NULL in eval mode) */
  float* log_prob_out;                 /* arena [T]: per-graph log-prob written at ids[k] (eval mode), may be NULL */
  float* log_prob_run;                 /* [B] per-graph log-prob in run order, may be NULL */
  float* tied_out;                     /* optional [n, ld_out]: outputs after orbit tying (forward() API) */
  float* dout; int ld_dout;            /* [n, d] gradient w.r.t. the JSMLP output before sin / tie (train), may be NULL */
  float* gstat;                        /* [B, 4+d]: surr, |dlp|, cat-entropy sum, 0, dlogstd[d] per graph (train) */
  float clip_eps, inv_B;               /* 1/global minibatch size */
  /* optional: the gradient written a second time in the fused JSMLP backward's operand format (sard_jsmlp_tc_backward with
   * dout == NULL then skips its own formatting launch): dout_sorted_hl[spos[row]] = [hi(32) | lo(32)] (columns >= d zero) and
   * zero rows at the padding positions (srow[pos] < 0) of the padded body-index sort, pos < rows_cap */
  float* dout_sorted_hl; const int* spos; const int* srow; int rows_cap; int pad_;
} SardPolicyLoss;
/* DiagGaussian.log_prob / entropy (khrylib/rl/core/distributions.py:21-22), per-graph reduction
 * (transform2act_policy.py:390-405), ratio/clip/min (transform2act_agent.py:392-396) and the
 * backward of all of it, with the orbit tie folded in. */
int sard_gauss_loss(const SardPolicyLoss* a, sard_stream_t stream);
/* Categorical.log_prob / entropy (distributions.py:46-50; policy.py:407-414) likewise. */
int sard_cat_loss(const SardPolicyLoss* a, sard_stream_t stream);
/* Ordered column sums of gstat[B, W]: stats[k] += sum_b gstat[b,k] for k<3, dlogstd[j] += sum_b gstat[b,4+j]
 * (j<d, when dlogstd != NULL).  When log_std_for_entropy != NULL the stage's mean Gaussian entropy
 * mean_j(0.5 + 0.5 log 2pi + log_std[j]) is added to stats[4] (policy.py:396,405). */
int sard_loss_reduce(const float* gstat, int B, int W, float* stats, float* dlogstd, int d,
                     const float* log_std_for_entropy, sard_stream_t stream);
/* Value head + MSE loss + the head's backward-data, fused (transform2act_critic.py:107-108, agent_pg.py:21-22):
 * v[b] = cat[b,:K] . w + *bias; values_out[ids[b]] = v (optional); with `returns`: e = v - returns[ids[b]],
 * dv[b] = 2 e inv_B, sq_err[b] = e^2 and (optional) dcat[b,c] = dv w[c] act'(cat[b,c]) with act' of `act0` on the first
 * dsplit columns and ReLU' on the rest. */
int sard_value_head(const float* cat, int ldcat, const float* w, const float* bias, int B, int K, int dsplit, int act0,
                    const int* ids, const float* returns, float inv_B, float* v, float* dv, float* sq_err,
                    float* values_out, float* dcat, int ld_dcat, sard_stream_t stream);

/* Critic loss mean((v-R)^2) and dv (khrylib/rl/agents/agent_pg.py:21-22): dv[b] = 2 (v[b]-returns[ids[b]]) inv_B;
 * partial[b] = (v-R)^2; eval mode (returns==NULL) writes values_out[ids[b]] = v[b]. */
int sard_value_loss(const float* v, const int* ids, int B, const float* returns, float inv_B,
                    float* dv, float* sq_err, float* values_out, sard_stream_t stream);

/* =====================================================================================
 * GAE  (khrylib/rl/core/common.py:5-25)
 * ===================================================================================== */
/* Reverse segmented scan in fp64: delta_i = r_i + gamma V_{i+1} m_i - V_i, A_i = delta_i + gamma tau m_i A_{i+1};
 * returns = V + A; advantages = (A - mean) / std (unbiased).  fp32 in/out.  scratch: double[4*ceil(T/2048)+8]. */
int sard_gae(const float* rewards, const float* masks, const float* values, int T, double gamma, double tau,
             float* advantages, float* returns, double* scratch, sard_stream_t stream);
/* fp64 in/out variant for callers that keep the reference's torch.float64 tensors. */
int sard_gae_f64(const double* rewards, const double* masks, const double* values, int T, double gamma, double tau,
                 double* advantages, double* returns, double* scratch, sard_stream_t stream);

/* =====================================================================================
 * Optimiser: gradient-norm clip + Adam over live parameter segments
 * (khrylib/rl/agents/agent_ppo.py:53-56; torch.optim.Adam at transform2act_agent.py:143,148)
 * ===================================================================================== */
typedef struct {
  float* p; float* m; float* v; const float* g;   /* parameter, exp_avg, exp_avg_sq, gradient */
  long long n;                                    /* elements */
  int group;                                      /* parameter group (own Adam step counter) */
  int pad;
} SardAdamSeg;

/* Device control block shared by prepare/apply: see sard_adam_prepare. */
#define SARD_CTL_FLOATS 64
#define SARD_CTL_INTS 32
#define SARD_MAX_GROUPS 8
/* OR into active_mask: zero_grad() semantics of torch < 2.0, the reference's pin (README.md:14, torch 1.8): gradients
 * of a stage absent from a minibatch are zero TENSORS, not None, so once a parameter group has taken one applied step
 * Adam keeps stepping it from momentum (g = 0, step counter advancing) on every later applied step.  Without the
 * flag an absent group is skipped (torch >= 2.0, set_to_none=True).  State: ctl_i[2] = groups seen so far,
 * ctl_i[3] = groups updated by the current step. */
#define SARD_ADAM_STICKY (1 << 30)

/* sumsq partials of g[n] (ordered two-stage reduction): out[0] = sum g^2.  scratch: float[1024]. */
int sard_sqnorm(const float* g, long long n, float* scratch, float* out, sard_stream_t stream);

/* Decide the step on device (no host sync):
 *   surr_loss = -stats[0]*inv_B, approx_kl = stats[1]*inv_B,
 *   entropy = stats[4]*inv_world + stats[2]*inv_cat_nodes   (stats is SUM-reduced over ranks)
 *   valid = !use_gate || !(approx_kl > 0.1 || surr_loss > 10)            (transform2act_agent.py:404-407)
 *   scale = (max_norm > 0 && clip armed) ? min(1, max_norm/(sqrt(*sqnorm)+1e-6)) : 1   (clip_grad_norm_)
 *   first_only: the clip is disarmed after the first applied step -- the reference passes a generator
 *   to clip_grad_norm_ (agent :48), exhausted by its first use (agent_ppo.py:53-56);
 *   for every group with its active_mask bit set: step += valid, bias corrections in fp64.
 * ctl_f[0]=scale, [1]=valid, [8+g]=1-beta1^step_g, [16+g]=sqrt(1-beta2^step_g); ctl_i[0]=valid,
 * ctl_i[1]=clip armed flag, ctl_i[8+g]=step_g.
 * log_row (may be NULL): {surr_loss, entropy, approx_kl, valid, grad_norm, stats[5]*inv_B}. */
int sard_adam_prepare(const float* stats, float inv_B, float inv_world, float inv_cat_nodes,
                      const float* sqnorm, float max_norm, int first_only, int use_gate, int active_mask,
                      double beta1, double beta2, float* ctl_f, int* ctl_i, float* log_row, sard_stream_t stream);
/* Adam (torch defaults: no amsgrad, weight_decay 0) over `n_seg` segments, skipped when !valid or the
 * segment's group is inactive.  max_seg_len bounds the grid. */
int sard_adam_apply(const SardAdamSeg* segs, int n_seg, long long max_seg_len, float lr, float beta1, float beta2,
                    float eps, int active_mask, const float* ctl_f, const int* ctl_i, sard_stream_t stream);

/* The whole optimiser tail in one call, for update loops that keep the gradient bucket resident: optional
 * sum of squares of g_sq[n_sq] (NULL: no clipping; scratch float[1024]), sard_adam_prepare, sard_adam_apply, and
 * the bucket is left ZEROED for the next step (stats[0..15] and every segment's gradients, also when the step
 * is gated off), which replaces the memset before the next backward.  Same arithmetic as the three calls. */
int sard_adam_update(float* stats, float inv_B, float inv_world, float inv_cat_nodes, const float* g_sq, long long n_sq,
                     float* sq_scratch, float max_norm, int first_only, int use_gate, int active_mask, double beta1,
                     double beta2, float* ctl_f, int* ctl_i, float* log_row, const SardAdamSeg* segs, int n_seg,
                     long long max_seg_len, float lr, float eps, sard_stream_t stream);

/* Data-parallel gradient SUM over NVLink peer memory (the reference is single-process: this is the
 * collective the transition-sharded update adds).  bufs_dev / pads_dev: DEVICE arrays of `world` pointers to every
 * rank's gradient bucket and signal pad (symmetric allocations mapped into this process; the pad needs
 * sard_allreduce_p2p_pad_bytes() zero-initialised bytes).  Reduces bucket[0:n_floats) in place on every rank:
 * two-shot (rank r sums slice r in rank order with peer loads, then stores it to every rank), deterministic,
 * bit-identical on all ranks; one kernel on `stream`, two in-kernel barriers whose expected values are
 * first_barrier_value and first_barrier_value + 1 (the caller advances it by 2 per call).  Every rank must call it
 * with the same arguments in the same order. */
int sard_allreduce_p2p_pad_bytes(void);
int sard_allreduce_p2p(const void* bufs_dev, const void* pads_dev, int rank, int world, long long n_floats,
                       unsigned first_barrier_value, sard_stream_t stream);

/* All-gather of one small record (RunningNorm moments of a run, record_doubles = 1 + 2D) over peer memory: rank r's record
 * is stored into every rank's symmetric board at chunk_offset_doubles + r * record_doubles, between two in-kernel barriers
 * ("every rank is ready to receive" / "every record has landed").  boards_dev / pads_dev: DEVICE arrays of `world`
 * pointers to every rank's board and signal pad (sard_allgather_p2p_pad_bytes() bytes, zero-initialised, used by this
 * chain of all-gathers only); first_barrier_value: 1, 3, 5, ... per call on that pad.  Replaces the NCCL all-gather of the
 * prefetched gather phase (the reference is single-process: SURVEY.md 8e). */
int sard_allgather_p2p_pad_bytes(void);
int sard_allgather_p2p(const void* boards_dev, const void* pads_dev, int rank, int world, long long chunk_offset_doubles,
                       const double* local_record, int record_doubles, unsigned first_barrier_value, sard_stream_t stream);

/* =====================================================================================
 * Network schedules: the whole forward/backward of one net over one run, launched from C++
 * (one host call per minibatch instead of ~100)
 * ===================================================================================== */
typedef struct { int out, in, act; long long w, b; } SardDense;   /* offsets (floats) into flat param/grad buffers */

typedef struct {                       /* RunningNorm + GNNSimple (gnn.py:7-68) */
  int D, L, act, has_bias;
  int h[SARD_MAX_LAYERS + 1];          /* h[0]: in_fc width, h[l]: width after GraphConv l */
  long long in_w, in_b;
  long long wl[SARD_MAX_LAYERS], bl[SARD_MAX_LAYERS], wr[SARD_MAX_LAYERS];
  float *mean, *var, *std, *inv;       /* RunningNorm buffers [D] + scratch inv[D] */
  long long* count;                    /* RunningNorm.n */
} SardTower;

typedef struct {                       /* Transform2ActValue (transform2act_critic.py:13-108) */
  SardTower tower;
  SardDense enc[3][2];                 /* hfield/obj/goal ObsEncoders */
  int n_mlp; SardDense mlp[SARD_MAX_MLP];
  SardDense head;
  int nconst; const float* consts; int onehot;
  float* P; float* G;                  /* flat parameters / gradients */
} SardValueNet;

typedef struct { int out, in; float* W; float* b; long long gW, gb; } SardIndexLinear;

typedef struct {                       /* one stage of Transform2ActPolicy (transform2act_policy.py:233-335) */
  int stage;
  SardTower tower;
  int lead, tail;
  int n_il; SardIndexLinear il[SARD_MAX_IL]; int il_act;
  int out_dim, post;
  long long log_std;                   /* offset in flat params, -1 for the categorical stage */
  int tie_c0, tie_c1;
  int max_index;
  int n_live;                          /* number of body indices with a slot (bounds the grouped tile / chunk grids) */
  const int* slot_of_index;            /* [max_index] compact gradient slot per body index */
  float* GT;                           /* compact table gradients of this stage */
} SardPolicyStage;

typedef struct {
  SardPolicyStage stage[3];
  SardDense enc[3][2];
  int nconst; const float* consts;
  float* P; float* G;
} SardPolicyNet;

enum { SARD_PHASE_GATHER = 1, SARD_PHASE_COMPUTE = 2, SARD_PHASE_ALL = 3 };

/* Workspace floats needed for a run of n nodes / B graphs. */
long long sard_value_workspace(const SardValueNet* net, int n, int B);
long long sard_policy_workspace(const SardPolicyNet* net, int stage, int n, int B);

/* Critic forward over one run.  train: 0 = eval, 1 = RunningNorm update + loss + backward,
 * 2 = RunningNorm update + forward only (a train-mode forward without a loss).
 * phase GATHER: gather + local RunningNorm moments into moments_local [1+2D doubles] (train only);
 * phase COMPUTE: merge `n_records` records from moments_all, normalise, forward; train: MSE loss
 * against arena->returns with 1/B_global scaling, backward into net->G (accumulated), sum of squared
 * errors added to *sq_err_sum; eval: arena->values[ids] = V.
 * Replaces Transform2ActValue.forward (critic.py:78-108) and AgentPG.update_value's
 * forward/backward (agent_pg.py:18-24). */
int sard_value_run(const SardValueNet* net, const SardArena* arena, const SardSel* sel, int train, int phase,
                   float inv_B_global, double* moments_local, const double* moments_all, int n_records,
                   float* sq_err_sum, float* ws, long long ws_floats, sard_stream_t stream);

/* One policy stage forward (+ PPO loss and backward when train != 0) over one run of that stage.
 * eval: arena->fixed_log_probs[ids] = log pi(a|s) (get_log_prob, policy.py:383-419);
 * train: stats[0..2] += {sum surr, sum |dlp|, sum cat entropy}; grads into net->G and stage GT.
 * outputs (optional, [n, out_dim]): the tied means/logits for the forward() API. */
int sard_policy_run(const SardPolicyNet* net, int stage, const SardArena* arena, const SardSel* sel, int train,
                    int phase, float clip_eps, float inv_B_global, double* moments_local,
                    const double* moments_all, int n_records, float* stats, float* outputs,
                    float* log_prob_run, float* ws, long long ws_floats, sard_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SARD_B200_H */
