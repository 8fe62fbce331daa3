// Network schedules: one host call launches the whole forward(+loss+backward) of a net over one run.
// The schedules mirror the reference modules layer by layer:
//   tower   = RunningNorm -> in_fc -> L x (GraphConv -> act)          (running_norm.py, gnn.py:54-68)
//   critic  = tower -> MLP(tanh) on root rows -> cat ext -> value_head  (transform2act_critic.py:78-108)
//   stage   = tower -> cat ext -> JSMLP by body index -> [sin] -> [orbit tie] -> loss
//                                                                      (transform2act_policy.py:233-335)
// All buffers come from a caller-provided workspace carved deterministically (bump allocation), so
// the GATHER and COMPUTE phases of a data-parallel step can be separate calls.
#include "common.cuh"
#include "encoder.cuh"
#include "tower_tc.cuh"
#include "critic_tc.cuh"
#include <string.h>
#include <stdlib.h>
#include <mutex>
#include <unordered_map>

int g_sard_wgrad_overlap = 1;
extern int g_sard_gemm_engine;
static const int g_sard_tower_tc = getenv("SARD_TOWER_TC") ? atoi(getenv("SARD_TOWER_TC")) : 1;   // fused tcgen05 tower (tower_tc.cu)
static const int g_sard_fuse_dout = getenv("SARD_FUSE_DOUT") ? atoi(getenv("SARD_FUSE_DOUT")) : 1;   // loss kernel writes the JSMLP backward's operand
static const int g_sard_side_reduce = getenv("SARD_SIDE_REDUCE") ? atoi(getenv("SARD_SIDE_REDUCE")) : 1; // loss statistics reduced off the chain
static const int g_sard_fuse_agg = getenv("SARD_FUSE_AGG") ? atoi(getenv("SARD_FUSE_AGG")) : 1;   // GraphConv neighbour sum inside the GEMM loader
extern "C" int sard_set_wgrad_overlap(int enabled) { g_sard_wgrad_overlap = enabled ? 1 : 0; return 0; }
extern "C" int sard_get_wgrad_overlap(void) { return g_sard_wgrad_overlap; }

thread_local char g_sard_err[1024] = "";
int sard_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_sard_err, sizeof(g_sard_err), fmt, ap);
  va_end(ap);
  return 1;
}
extern "C" const char* sard_last_error_string(void) { return g_sard_err; }
extern "C" int sard_abi_version(void) { return SARD_ABI_VERSION; }
extern "C" int sard_struct_sizes(int* out, int capacity) {
  const int sizes[] = {(int)sizeof(SardArena), (int)sizeof(SardSel), (int)sizeof(SardGatherOut), (int)sizeof(SardGemm),
                       (int)sizeof(SardWGrad), (int)sizeof(SardPolicyLoss), (int)sizeof(SardAdamSeg), (int)sizeof(SardDense),
                       (int)sizeof(SardTower), (int)sizeof(SardValueNet), (int)sizeof(SardIndexLinear),
                       (int)sizeof(SardPolicyStage), (int)sizeof(SardPolicyNet), (int)sizeof(SardJsmlpTc), (int)sizeof(SardJsmlpTcBwd)};
  const int n = (int)(sizeof(sizes) / sizeof(sizes[0]));
  for (int i = 0; i < n && i < capacity; ++i) out[i] = sizes[i];
  return n;
}

namespace {

// profiling call-site labels: 3*kind + pass (see sard_prof_enable)
enum { K_INFC = 0, K_GCONV = 1, K_IL = 2, K_ENC = 3, K_MLP = 4, K_HEAD = 5, K_AGG = 6 };
inline void site(int kind, int pass) { g_sard_site = 3 * kind + pass; }

// Weight gradients leave the critical path: in every backward pass the chain dz_l -> bwd_data -> dz_{l-1}
// is what the next layer waits for, while dW_l (partials + ordered reduce) is only needed by the optimiser
// step.  Each launching stream owns two low-priority side streams (one for the tower's weight gradients,
// one for the head / MLP / JSMLP / encoder ones: a single side stream lagged 100-150 us behind the main
// chain at the end of every minibatch); a weight gradient is forked onto its side stream with an event
// recorded right after dz_l was produced, and the run joins both before returning.  A side stream
// serialises its weight gradients, so each has ONE partials scratch; dz buffers are per layer (no
// ping-pong) so the main stream never overwrites one that is still read.
#define N_SIDE 3
struct SideCtx {
  cudaStream_t main, side[N_SIDE];   // [0]: tower weight gradients, [1]: head / MLP / JSMLP weight gradients, [2]: ObsEncoders
  cudaEvent_t ev[8];
  int k;
  bool forked[N_SIDE];
};
std::mutex g_side_mu;
std::unordered_map<unsigned long long, SideCtx*> g_side;    // key: launching stream handle ^ device (the default stream's handle is 0 on every device)

SideCtx* side_for(cudaStream_t main) {
  if (!g_sard_wgrad_overlap) return nullptr;
  std::lock_guard<std::mutex> lock(g_side_mu);
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long key = (unsigned long long)reinterpret_cast<uintptr_t>(main) ^ ((unsigned long long)(dev + 1) << 56);
  auto it = g_side.find(key);
  if (it != g_side.end()) { for (int i = 0; i < N_SIDE; ++i) it->second->forked[i] = false; return it->second; }
  SideCtx* c = new SideCtx();
  c->main = main; c->k = 0;
  for (int i = 0; i < N_SIDE; ++i) c->forked[i] = false;
  // side streams run `SARD_SIDE_PRIO_DELTA` priority steps below their launching stream (default 1; the least priority of
  // the device at most): the net whose chain is the critical path is given a high-priority launching stream by the agent
  int lo = 0, hi = 0, mp = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  if (cudaStreamGetPriority(main, &mp) != cudaSuccess) { cudaGetLastError(); mp = lo; }
  static const int delta = getenv("SARD_SIDE_PRIO_DELTA") ? atoi(getenv("SARD_SIDE_PRIO_DELTA")) : 1;
  int sp = mp + delta;
  if (sp > lo) sp = lo;
  if (sp < hi) sp = hi;
  for (int i = 0; i < N_SIDE; ++i)
    if (cudaStreamCreateWithPriority(&c->side[i], cudaStreamNonBlocking, sp) != cudaSuccess) { delete c; return nullptr; }
  for (int i = 0; i < 8; ++i) cudaEventCreateWithFlags(&c->ev[i], cudaEventDisableTiming);
  g_side[key] = c;
  return c;
}
// stream for a weight gradient whose inputs were just produced on `st`
sard_stream_t wg_stream(SideCtx* c, sard_stream_t st, int which) {
  if (!c) return st;
  cudaEvent_t e = c->ev[c->k++ & 7];
  cudaEventRecord(e, c->main);
  cudaStreamWaitEvent(c->side[which], e, 0);
  c->forked[which] = true;
  return reinterpret_cast<sard_stream_t>(c->side[which]);
}
int side_join(SideCtx* c) {
  if (!c) return 0;
  for (int i = 0; i < N_SIDE; ++i) {
    if (!c->forked[i]) continue;
    cudaEvent_t e = c->ev[c->k++ & 7];
    SARD_CUDA(cudaEventRecord(e, c->side[i]));
    SARD_CUDA(cudaStreamWaitEvent(c->main, e, 0));
    c->forked[i] = false;
  }
  return 0;
}

struct Bump {
  float* base; long long off;
  float* f(long long n) { float* p = base ? base + off : nullptr; off += (n + 3) & ~3LL; return p; }
  // 256-byte aligned block (TMA sources / destinations); alignment is relative to the ADDRESS, so the sizing pass
  // (base == nullptr) reserves the worst case
  float* a256(long long n) {
    if (!base) { off += 64 + ((n + 63) & ~63LL); return nullptr; }
    const uintptr_t cur = reinterpret_cast<uintptr_t>(base + off);
    const long long pad = (long long)(((cur + 255) & ~uintptr_t(255)) - cur) / 4;
    float* p = base + off + pad;
    off += 64 + ((n + 63) & ~63LL);
    return p;
  }
  int* i(long long n) { return reinterpret_cast<int*>(f(n)); }
};

inline int r4(int v) { return (v + 3) & ~3; }
inline long long wg_partials(int M, int N1, int N2, int /*unused*/) {
  if (M <= 0) return 0;
  return (long long)ceil_div(M, sard_wgrad_chunk_rows(M, N1, N2)) * N1 * SARD_WG_LDPART(N2);
}

const int WG_CHUNK = getenv("SARD_WG_CHUNK") ? atoi(getenv("SARD_WG_CHUNK")) : 256;     // rows per grouped weight-gradient chunk
constexpr int IL_TILE = 128;      // rows per grouped-GEMM tile (= BM of gemm_kernel)

struct TowerBufs {
  float* x; int ldx;
  int *nd_graph, *nd_arena, *body, *root_row;
  float* partials;
  float* xa[SARD_MAX_LAYERS];
  float* hout;
  float *dz[SARD_MAX_LAYERS + 1], *dxa;      // dz[l]: gradient w.r.t. the pre-activation of layer l (0 = in_fc)
  int maxh;
  bool tc_shape;                // widths served by the fused tensor-core tower: its scratch is carved
  bool tc_on;                   // ... and this run's graphs are small enough (set per run from arena->max_nodes)
  TowerTcBufs tc;
  TowerNormStep norm;           // RunningNorm step of this run, launched with the weight pre-split when tc_on
};

// the fused tower serves the default engine when the widths fit (every layer 64 wide, D <= 127, <= 3 layers)
bool tower_tc_shape(const SardTower& tw) { return g_sard_tower_tc && g_sard_gemm_engine == 4 && sard_tower_tc_supported(tw, 1); }
// per run: graphs must fit whole into a 128-row tile with room to pack (max_nodes <= 64)
void tower_tc_select(TowerBufs& t, const SardTower& tw, const SardArena* arena, int n) {
  t.tc_on = t.tc_shape && sard_tower_tc_supported(tw, arena->max_nodes);
  if (!t.tc_on) return;
  t.tc.span = sard_tower_tc_span(arena->max_nodes);
  t.tc.num_tiles = ceil_div(n, t.tc.span);
}

void carve_tower(Bump& b, const SardTower& tw, int n, int B, bool train, TowerBufs& t) {
  t.ldx = r4(tw.D);
  t.x = b.f((long long)n * t.ldx);
  t.nd_graph = b.i(n); t.nd_arena = b.i(n); t.body = b.i(n); t.root_row = b.i(B);
  t.partials = b.f(2LL * tw.D * ceil_div(n > 0 ? n : 1, 128));
  t.maxh = 0;
  for (int l = 0; l <= tw.L; ++l) t.maxh = tw.h[l] > t.maxh ? tw.h[l] : t.maxh;
  for (int l = 0; l < tw.L; ++l) t.xa[l] = b.f((long long)n * 2 * tw.h[l]);
  t.hout = b.f((long long)n * tw.h[tw.L]);
  if (train) {
    for (int l = 0; l <= tw.L; ++l) t.dz[l] = b.f((long long)n * tw.h[l]);
    t.dxa = b.f((long long)n * 2 * t.maxh);
  } else {
    for (int l = 0; l <= tw.L; ++l) t.dz[l] = nullptr;
    t.dxa = nullptr;
  }
  t.tc_shape = tower_tc_shape(tw);
  t.tc_on = false;
  memset(&t.tc, 0, sizeof(t.tc));
  memset(&t.norm, 0, sizeof(t.norm));
  if (t.tc_shape) {
    const int nn = n > 0 ? n : 1;
    t.tc.wsplit = b.a256(sard_tower_tc_wsplit_floats(tw.L));
    t.tc.tile_first = b.i(ceil_div(nn, sard_tower_tc_span(TT_MAX_NODES)) + 2);
    t.tc.tab_in = reinterpret_cast<unsigned long long*>(b.a256(2LL * nn));
    t.tc.tab_out = reinterpret_cast<unsigned long long*>(b.a256(2LL * nn));
    t.tc.chunk_rows = sard_tower_tc_chunk_rows(nn);
    t.tc.num_chunks = ceil_div(nn, t.tc.chunk_rows);
    if (train) {
      t.tc.XH = b.a256((long long)nn * 128);
      for (int l = 0; l < tw.L; ++l) {
        t.tc.PH[l] = b.a256((long long)nn * 64); t.tc.PA[l] = b.a256((long long)nn * 128);
        t.tc.mask[l] = reinterpret_cast<unsigned*>(b.a256(2LL * nn));
      }
      t.tc.PZ0 = b.a256((long long)nn * 64);
      t.tc.partials = b.a256(sard_tower_tc_partials_floats(t.tc.num_chunks));
    }
  }
}

long long tower_wg_need(const SardTower& tw, int n) {
  long long need = wg_partials(n, tw.h[0], tw.D, WG_CHUNK);
  for (int l = 1; l <= tw.L; ++l) {
    long long v = wg_partials(n, tw.h[l], 2 * tw.h[l - 1], WG_CHUNK);
    need = v > need ? v : need;
  }
  return need;
}

SardGemm gemm0() { SardGemm g; memset(&g, 0, sizeof(g)); return g; }
SardWGrad wgrad0() { SardWGrad w; memset(&w, 0, sizeof(w)); w.chunk_rows = 0; return w; }   // 0: auto chunking

int tower_gather(const SardTower& tw, const SardArena* arena, const SardSel* sel, int lead, int tail, int nconst,
                 const float* consts, int onehot, TowerBufs& t, sard_stream_t st) {
  SardGatherOut o;
  o.lead = lead; o.tail = tail; o.nconst = nconst; o.onehot = onehot; o.consts = consts;
  o.x = t.x; o.ldx = t.ldx; o.nd_graph = t.nd_graph; o.nd_arena = t.nd_arena; o.body = t.body; o.root_row = t.root_row;
  SARD_REQUIRE(lead + tail + nconst + (onehot ? 3 : 0) == tw.D, "tower input width != column program");
  SARD_TRY(sard_gather_nodes(arena, sel, &o, st));
  tower_tc_select(t, tw, arena, sel->n);
  if (t.tc_on) SARD_TRY(sard_tower_tc_tiles(arena, sel, t.tc, t.nd_graph, t.nd_arena, st));
  return 0;
}

// RunningNorm step of a compute phase: with the fused tower it rides in the pre-split launch, otherwise its own kernel
int tower_norm_step(const SardTower& tw, const SardArena* arena, int n, TowerBufs& t, bool upd, const double* recs, int n_records,
                    sard_stream_t st) {
  tower_tc_select(t, tw, arena, n);
  if (t.tc_on) {
    t.norm.mode = upd ? 1 : 2; t.norm.n_records = n_records; t.norm.records = recs;
    t.norm.mean = tw.mean; t.norm.var = tw.var; t.norm.std = tw.std; t.norm.inv = tw.inv; t.norm.count = tw.count;
    return 0;
  }
  if (upd) return sard_norm_update(recs, n_records, tw.D, tw.mean, tw.var, tw.std, tw.count, tw.inv, st);
  return sard_norm_prepare_eval(tw.std, tw.D, tw.inv, st);
}

// normalise + in_fc + GraphConv layers
int tower_forward(const SardTower& tw, const float* P, const SardArena* arena, const SardSel* sel, TowerBufs& t,
                  sard_stream_t st) {
  const int n = sel->n;
  tower_tc_select(t, tw, arena, n);
  if (t.tc_on) {
    // normalise + in_fc + every GraphConv layer in one tcgen05 launch (tower_tc.cu); t.x stays un-normalised
    SARD_TRY(sard_tower_tc_presplit(tw, P, t.tc.wsplit, t.norm, st));
    site(K_GCONV, 0);
    return sard_tower_tc_forward(tw, P, arena, sel, t.tc, t.x, t.ldx, t.nd_graph, t.nd_arena, t.hout, t.dz[tw.L] != nullptr, st);
  }
  SARD_TRY(sard_norm_apply(t.x, t.ldx, n, tw.D, tw.mean, tw.inv, tw.count, 5.0f, st));
  {
    SardGemm g = gemm0();
    g.A = t.x; g.lda = t.ldx; g.B0 = P + tw.in_w; g.ldb0 = tw.D; g.b_split = tw.D; g.bias = P + tw.in_b;
    g.M = n; g.N = tw.h[0]; g.R = tw.D; g.act = SARD_ACT_NONE;
    if (tw.L > 0) { g.Y = t.xa[0] + tw.h[0]; g.ldy = 2 * tw.h[0]; } else { g.Y = t.hout; g.ldy = tw.h[0]; }
    site(K_INFC, 0);
    SARD_TRY(sard_linear_fwd(&g, st));
  }
  for (int l = 1; l <= tw.L; ++l) {
    const int hi = tw.h[l - 1], ho = tw.h[l];
    float* xin = t.xa[l - 1];
    const bool fuse = g_sard_fuse_agg && sard_gemm_agg_supported(hi, ho);
    SardGemm g = gemm0();
    if (fuse) {
      // the neighbour sum is taken in the GEMM's A-loader (and written to xin[:, 0:hi] for the weight gradient)
      g.A = xin + hi; g.agg_cols = hi; g.agg_ptr = arena->in_ptr; g.agg_nbr = arena->in_nbr; g.agg_node = t.nd_arena;
      g.agg_graph = t.nd_graph; g.agg_cum = sel->cum; g.agg_node_base = sel->node_base; g.agg_out = xin; g.ld_agg_out = 2 * hi;
    } else {
      site(K_AGG, 0);
      SARD_TRY(sard_csr_aggregate(xin + hi, 2 * hi, xin, 2 * hi, nullptr, 0, nullptr, 0, 0, n, hi, t.nd_graph, t.nd_arena,
                                  sel->cum, sel->node_base, arena->in_ptr, arena->in_nbr, st));
      g.A = xin;
    }
    g.lda = 2 * hi; g.B0 = P + tw.wl[l - 1]; g.ldb0 = hi; g.B1 = P + tw.wr[l - 1]; g.ldb1 = hi; g.b_split = hi;
    g.bias = tw.has_bias ? P + tw.bl[l - 1] : nullptr;
    g.M = n; g.N = ho; g.R = 2 * hi; g.act = tw.act;
    if (l < tw.L) { g.Y = t.xa[l] + ho; g.ldy = 2 * ho; } else { g.Y = t.hout; g.ldy = ho; }
    site(K_GCONV, 0);
    SARD_TRY(sard_linear_fwd(&g, st));
  }
  return 0;
}

// dz_top: [n, h[L]] gradient w.r.t. the pre-activation of the last layer (act' already applied), in t.dz[L]
int tower_backward(const SardTower& tw, const float* P, float* G, const SardArena* arena, const SardSel* sel,
                   TowerBufs& t, float* wg_scratch, SideCtx* side, sard_stream_t st) {
  const int n = sel->n;
  if (t.tc_on) {
    (void)P; (void)wg_scratch;
    site(K_GCONV, 1);
    SARD_TRY(sard_tower_tc_backward(tw, arena, sel, t.tc, t.dz[tw.L], t.nd_graph, t.nd_arena, st));
    // the adjoint chain is the last thing on the launching stream and the optimiser waits for these weight gradients:
    // they stay on it (a fork + join through a side stream costs two event hops, ~15 us, and buys no overlap here)
    site(K_GCONV, 2);
    static const int wg_side = getenv("SARD_TOWER_WG_SIDE") ? atoi(getenv("SARD_TOWER_WG_SIDE")) : 0;
    SARD_TRY(sard_tower_tc_wgrad(tw, sel, t.tc, G, wg_side ? wg_stream(side, st, 0) : st));
    return side_join(side);
  }
  for (int l = tw.L; l >= 1; --l) {
    const int hi = tw.h[l - 1], ho = tw.h[l];
    float* dz = t.dz[l];
    float* xin = t.xa[l - 1];
    SardWGrad w = wgrad0();
    w.P = dz; w.ldp = ho; w.Q = xin; w.ldq = 2 * hi; w.M = n; w.N1 = ho; w.N2 = 2 * hi;
    w.dW0 = G + tw.wl[l - 1]; w.ldw0 = hi; w.dW1 = G + tw.wr[l - 1]; w.ldw1 = hi; w.w_split = hi;
    w.db = tw.has_bias ? G + tw.bl[l - 1] : nullptr; w.partials = wg_scratch;
    site(K_GCONV, 2);
    SARD_TRY(sard_linear_bwd_weight(&w, wg_stream(side, st, 0)));
    // dz_{l-1} = (dz_l W_r + A^T (dz_l W_l)) * act'(y_{l-1});  layer 0 (in_fc) has no activation
    const float* yprev = (l - 1 >= 1) ? xin + hi : nullptr;
    SardGemm g = gemm0();
    g.A = dz; g.lda = ho; g.B0 = P + tw.wl[l - 1]; g.ldb0 = hi; g.B1 = P + tw.wr[l - 1]; g.ldb1 = hi;
    g.M = n;
    if (g_sard_fuse_agg && sard_gemm_agg_supported(hi, ho)) {
      // one GEMM over the virtual operand [A^T dz_l | dz_l] (K = 2 ho) with the weights stacked [W_l ; W_r]
      g.agg_cols = ho; g.agg_ptr = arena->out_ptr; g.agg_nbr = arena->out_nbr; g.agg_node = t.nd_arena; g.agg_graph = t.nd_graph;
      g.agg_cum = sel->cum; g.agg_node_base = sel->node_base; g.b_ksplit = ho; g.b_split = hi;
      g.N = hi; g.R = 2 * ho; g.Y = t.dz[l - 1]; g.ldy = hi; g.dsplit = hi;
      if (yprev) { g.Dact = yprev; g.ldd = 2 * hi; g.dact0 = tw.act; g.dact1 = tw.act; }
      site(K_GCONV, 1);
      SARD_TRY(sard_linear_bwd_data(&g, st));
    } else {
      g.b_split = hi; g.N = 2 * hi; g.R = ho; g.Y = t.dxa; g.ldy = 2 * hi;
      site(K_GCONV, 1);
      SARD_TRY(sard_linear_bwd_data(&g, st));
      site(K_AGG, 1);
      SARD_TRY(sard_csr_aggregate(t.dxa, 2 * hi, t.dz[l - 1], hi, t.dxa + hi, 2 * hi, yprev, 2 * hi, tw.act, n, hi,
                                  t.nd_graph, t.nd_arena, sel->cum, sel->node_base, arena->out_ptr, arena->out_nbr, st));
    }
  }
  SardWGrad w = wgrad0();
  w.P = t.dz[0]; w.ldp = tw.h[0]; w.Q = t.x; w.ldq = t.ldx; w.M = n; w.N1 = tw.h[0]; w.N2 = tw.D;
  w.dW0 = G + tw.in_w; w.ldw0 = tw.D; w.w_split = tw.D; w.db = G + tw.in_b; w.partials = wg_scratch;
  site(K_INFC, 2);
  SARD_TRY(sard_linear_bwd_weight(&w, wg_stream(side, st, 0)));
  return side_join(side);
}

// ext[B, 3*64] written at `ext` with leading dimension ld_ext; e1[k]: [B,64] hidden activations
int encoders_forward(const SardDense enc[3][2], const float* P, const SardArena* arena, const SardSel* sel,
                     float* const e1[3], float* ext, int ld_ext, float* wide_scratch, sard_stream_t st) {
  site(K_ENC, 0);
  EncFwdP fp;
  memset(&fp, 0, sizeof(fp));
  fp.ids = sel->ids; fp.B = sel->B; fp.ext = ext; fp.ld_ext = ld_ext;
  for (int k = 0; k < 3; ++k) {
    SARD_REQUIRE(enc[k][0].in == arena->ext_dim[k], "encoder input dim != arena ext dim");
    if (sard_enc_fused_ok(enc[k][0].in, enc[k][0].out, enc[k][1].out)) {
      fp.src[k] = arena->ext[k]; fp.K[k] = enc[k][0].in; fp.w1[k] = P + enc[k][0].w; fp.b1[k] = P + enc[k][0].b;
      fp.w2[k] = P + enc[k][1].w; fp.b2[k] = P + enc[k][1].b; fp.e1[k] = e1[k];
      fp.enc_list[fp.n_enc++] = k;
    }
  }
  SARD_TRY(sard_enc_fwd_launch(fp, as_stream(st)));
  for (int k = 0; k < 3; ++k) {
    if (sard_enc_fused_ok(enc[k][0].in, enc[k][0].out, enc[k][1].out)) continue;
    SARD_REQUIRE(enc[k][0].in == arena->ext_dim[k], "encoder input dim != arena ext dim");
    if (wide_scratch && sard_enc_wide_ok(enc[k][0].in, enc[k][0].out)) {
      SARD_TRY(sard_enc_wide_fwd(arena->ext[k], enc[k][0].in, sel->ids, sel->B, P + enc[k][0].w, P + enc[k][0].b, e1[k], wide_scratch,
                                 as_stream(st)));
    } else {
      SardGemm g = gemm0();
      g.A = arena->ext[k]; g.lda = arena->ext_dim[k]; g.a_rows = sel->ids;
      g.B0 = P + enc[k][0].w; g.ldb0 = enc[k][0].in; g.b_split = enc[k][0].in; g.bias = P + enc[k][0].b;
      g.M = sel->B; g.N = enc[k][0].out; g.R = enc[k][0].in; g.act = SARD_ACT_RELU; g.Y = e1[k]; g.ldy = enc[k][0].out;
      SARD_TRY(sard_linear_fwd(&g, st));
    }
    SardGemm h = gemm0();
    h.A = e1[k]; h.lda = enc[k][0].out; h.B0 = P + enc[k][1].w; h.ldb0 = enc[k][1].in; h.b_split = enc[k][1].in;
    h.bias = P + enc[k][1].b; h.M = sel->B; h.N = enc[k][1].out; h.R = enc[k][1].in; h.act = SARD_ACT_RELU;
    h.Y = ext + k * SARD_EXT_DIM; h.ldy = ld_ext;
    SARD_TRY(sard_linear_fwd(&h, st));
  }
  return 0;
}

// dext: [B, 3*64] gradient w.r.t. the pre-activation of the encoders' second layer (relu' applied)
int encoders_backward(const SardDense enc[3][2], const float* P, float* G, const SardArena* arena, const SardSel* sel,
                      float* const e1[3], const float* dext, int ld_dext, float* de1, float* wg_scratch,
                      sard_stream_t st) {
  site(K_ENC, 2);
  EncBwdP bp;
  memset(&bp, 0, sizeof(bp));
  bp.ids = sel->ids; bp.B = sel->B; bp.dext = dext; bp.ld_dext = ld_dext; bp.partials = wg_scratch;
  for (int k = 0; k < 3; ++k) {
    if (!sard_enc_fused_ok(enc[k][0].in, enc[k][0].out, enc[k][1].out)) continue;
    bp.src[k] = arena->ext[k]; bp.K[k] = enc[k][0].in; bp.w2[k] = P + enc[k][1].w; bp.e1[k] = e1[k];
    bp.gw1[k] = G + enc[k][0].w; bp.gb1[k] = G + enc[k][0].b; bp.gw2[k] = G + enc[k][1].w; bp.gb2[k] = G + enc[k][1].b;
    bp.enc_list[bp.n_enc++] = k;
  }
  SARD_TRY(sard_enc_bwd_launch(bp, as_stream(st)));
  for (int k = 0; k < 3; ++k) {
    if (sard_enc_fused_ok(enc[k][0].in, enc[k][0].out, enc[k][1].out)) continue;
    const float* dz2 = dext + k * SARD_EXT_DIM;
    SardWGrad w = wgrad0();
    w.P = dz2; w.ldp = ld_dext; w.Q = e1[k]; w.ldq = enc[k][0].out; w.M = sel->B; w.N1 = enc[k][1].out; w.N2 = enc[k][1].in;
    w.dW0 = G + enc[k][1].w; w.ldw0 = enc[k][1].in; w.w_split = enc[k][1].in; w.db = G + enc[k][1].b; w.partials = wg_scratch;
    site(K_ENC, 2);
    SARD_TRY(sard_linear_bwd_weight(&w, st));
    SardGemm g = gemm0();
    g.A = dz2; g.lda = ld_dext; g.B0 = P + enc[k][1].w; g.ldb0 = enc[k][1].in; g.b_split = enc[k][1].in;
    g.M = sel->B; g.N = enc[k][1].in; g.R = enc[k][1].out; g.Y = de1; g.ldy = enc[k][1].in;
    g.Dact = e1[k]; g.ldd = enc[k][0].out; g.dact0 = SARD_ACT_RELU; g.dact1 = SARD_ACT_RELU; g.dsplit = g.N;
    site(K_ENC, 1);
    SARD_TRY(sard_linear_bwd_data(&g, st));
    site(K_ENC, 2);
    if (sard_enc_wide_ok(enc[k][0].in, enc[k][0].out)) {
      SARD_TRY(sard_enc_wide_wgrad(arena->ext[k], enc[k][0].in, sel->ids, sel->B, de1, G + enc[k][0].w, G + enc[k][0].b, wg_scratch,
                                   as_stream(st)));
      continue;
    }
    SardWGrad w1 = wgrad0();
    w1.P = de1; w1.ldp = enc[k][0].out; w1.Q = arena->ext[k]; w1.ldq = arena->ext_dim[k]; w1.q_rows = sel->ids;
    w1.M = sel->B; w1.N1 = enc[k][0].out; w1.N2 = enc[k][0].in;
    w1.dW0 = G + enc[k][0].w; w1.ldw0 = enc[k][0].in; w1.w_split = enc[k][0].in; w1.db = G + enc[k][0].b; w1.partials = wg_scratch;
    SARD_TRY(sard_linear_bwd_weight(&w1, st));
  }
  return 0;
}

// forward scratch of the wide-input first layers (split-K partials); 0 when no encoder is wide
long long enc_wide_fwd_need(const SardDense enc[3][2], int B) {
  long long need = 0;
  for (int k = 0; k < 3; ++k)
    if (sard_enc_wide_ok(enc[k][0].in, enc[k][0].out)) {
      const long long v = sard_enc_wide_partials_floats(B, enc[k][0].in);
      need = v > need ? v : need;
    }
  return need;
}

long long enc_wg_need(const SardDense enc[3][2], int B) {
  long long need = sard_enc_partials_floats(B, 3);
  for (int k = 0; k < 3; ++k)
    if (sard_enc_wide_ok(enc[k][0].in, enc[k][0].out)) {
      const long long v = sard_enc_wide_partials_floats(B, enc[k][0].in);
      need = v > need ? v : need;
    }
  for (int k = 0; k < 3; ++k)
    for (int j = 0; j < 2; ++j) {
      long long v = wg_partials(B, enc[k][j].out, enc[k][j].in, WG_CHUNK);
      need = v > need ? v : need;
    }
  return need;
}

// ------------------------------------------------------------------------------------------ critic
struct ValueBufs {
  TowerBufs t;
  double* rec;
  float* e1[3]; float* cat; int ldcat; float* m[SARD_MAX_MLP]; float* v; float* dv; float* sq;
  float* dcat; float* dm[SARD_MAX_MLP]; float* de1; float* wg; float* wg2; float* wg3; float* encw; float* sk; long long sk_floats;
  float* cmw;                    // pair copies of the MLP weights for the fused tensor-core forward (critic_tc.cu), or NULL
};

// the fused tensor-core MLP forward serves the default engine when the shapes are derl.yml's
static const int g_sard_critic_tc = getenv("SARD_CRITIC_TC") ? atoi(getenv("SARD_CRITIC_TC")) : 1;
bool critic_uses_tc(const SardValueNet& net) {
  return g_sard_critic_tc && g_sard_gemm_engine == 4 && net.n_mlp == 2 &&
         sard_critic_tc_supported(net.mlp[0].in, net.mlp[0].out, net.mlp[1].out, net.n_mlp) && net.mlp[1].in == net.mlp[0].out &&
         net.mlp[0].act == net.mlp[1].act && net.tower.h[net.tower.L] == net.mlp[0].in;
}

void carve_value(Bump& b, const SardValueNet& net, int n, int B, bool train, ValueBufs& vb) {
  carve_tower(b, net.tower, n, B, train, vb.t);
  vb.rec = reinterpret_cast<double*>(b.f(2 * (1 + 2 * net.tower.D) + 2));
  for (int k = 0; k < 3; ++k) vb.e1[k] = b.f((long long)B * net.enc[k][0].out);
  const int m_last = net.n_mlp > 0 ? net.mlp[net.n_mlp - 1].out : net.tower.h[net.tower.L];
  vb.ldcat = m_last + 3 * SARD_EXT_DIM;
  vb.cat = b.f((long long)B * vb.ldcat);
  int maxm = 0;
  for (int j = 0; j < net.n_mlp; ++j) {
    vb.m[j] = (j < net.n_mlp - 1) ? b.f((long long)B * net.mlp[j].out) : vb.cat;
    maxm = net.mlp[j].out > maxm ? net.mlp[j].out : maxm;
  }
  vb.v = b.f(B); vb.dv = b.f(B); vb.sq = b.f(B);
  vb.sk_floats = 1024LL * B;                                 // split-K partial tiles of the MLP GEMMs: S*N <= 1024 columns
  vb.sk = b.f(vb.sk_floats);
  { const long long w = enc_wide_fwd_need(net.enc, B); vb.encw = w > 0 ? b.f(w) : nullptr; }
  vb.cmw = critic_uses_tc(net) ? b.a256(sard_critic_tc_wsplit_floats()) : nullptr;
  if (train) {
    vb.dcat = b.f((long long)B * vb.ldcat);
    for (int j = 0; j < net.n_mlp; ++j) vb.dm[j] = b.f((long long)B * (maxm > 0 ? maxm : 4));
    vb.de1 = b.f((long long)B * SARD_EXT_DIM);
    vb.wg = b.f(tower_wg_need(net.tower, n));               // side stream 0
    long long need = 4, v;                                   // side stream 1
    for (int j = 0; j < net.n_mlp; ++j) { v = wg_partials(B, net.mlp[j].out, net.mlp[j].in, WG_CHUNK); need = v > need ? v : need; }
    v = wg_partials(B, 1, vb.ldcat, WG_CHUNK); need = v > need ? v : need;
    vb.wg2 = b.f(need);
    vb.wg3 = b.f(enc_wg_need(net.enc, B));                   // side stream 2
  } else {
    vb.dcat = vb.de1 = vb.wg = vb.wg2 = vb.wg3 = nullptr;
    for (int j = 0; j < net.n_mlp; ++j) vb.dm[j] = nullptr;
  }
}

}  // namespace

extern "C" long long sard_value_workspace(const SardValueNet* net, int n, int B) {
  Bump b{nullptr, 0};
  ValueBufs vb;
  carve_value(b, *net, n, B, true, vb);
  return b.off + 64;
}

extern "C" int sard_value_run(const SardValueNet* net, const SardArena* arena, const SardSel* sel, int train, int phase,
                              float inv_B_global, double* moments_local, const double* moments_all, int n_records,
                              float* sq_err_sum, float* ws, long long ws_floats, sard_stream_t st) {
  if (sel->B <= 0 || sel->n <= 0) return 0;
  const SardTower& tw = net->tower;
  SARD_REQUIRE(tw.L >= 0 && tw.L <= SARD_MAX_LAYERS && net->n_mlp <= SARD_MAX_MLP, "layer counts");
  // train: 0 = eval, 1 = RunningNorm update + loss + backward, 2 = RunningNorm update + forward only
  const bool upd = train != 0, bwd = train == 1;
  Bump b{ws, 0};
  ValueBufs vb;
  carve_value(b, *net, sel->n, sel->B, bwd, vb);
  SARD_REQUIRE(b.off <= ws_floats, "value workspace too small");
  const float* P = net->P;
  float* G = net->G;
  const int n = sel->n, B = sel->B;
  if (phase & SARD_PHASE_GATHER) {
    SARD_TRY(tower_gather(tw, arena, sel, arena->d_obs, 0, net->nconst, net->consts, net->onehot, vb.t, st));
    if (upd) {
      double* rec = moments_local ? moments_local : vb.rec;
      SARD_TRY(sard_norm_moments(vb.t.x, vb.t.ldx, n, tw.D, vb.t.partials, rec, st));
    }
  }
  if (!(phase & SARD_PHASE_COMPUTE)) return 0;
  {
    const double* recs = moments_all ? moments_all : (moments_local ? moments_local : vb.rec);
    SARD_TRY(tower_norm_step(tw, arena, n, vb.t, upd, recs, moments_all ? n_records : 1, st));
  }
  // the ObsEncoders only need the arena: they run beside the tower and are joined before the value head
  SideCtx* side = side_for(as_stream(st));
  SARD_TRY(encoders_forward(net->enc, P, arena, sel, vb.e1, vb.cat + (vb.ldcat - 3 * SARD_EXT_DIM), vb.ldcat, vb.encw,
                            wg_stream(side, st, 2)));
  if (vb.cmw) SARD_TRY(sard_critic_tc_presplit(P + net->mlp[0].w, P + net->mlp[1].w, vb.cmw, wg_stream(side, st, 1)));
  SARD_TRY(tower_forward(tw, P, arena, sel, vb.t, st));
  // MLP on root rows only: the reference evaluates it on every node and then keeps the roots (critic.py:104-107)
  const int hL = tw.h[tw.L];
  if (vb.cmw) {
    // both layers in one tcgen05 launch (critic_tc.cu); h1 is kept for the backward pass only
    SARD_TRY(side_join(side));
    site(K_MLP, 0);
    SARD_TRY(sard_critic_tc_forward(vb.cmw, vb.t.hout, vb.t.root_row, B, P + net->mlp[0].b, P + net->mlp[1].b, net->mlp[0].act,
                                    bwd ? vb.m[0] : nullptr, vb.cat, vb.ldcat, st));
  } else
  for (int j = 0; j < net->n_mlp; ++j) {
    SardGemm g = gemm0();
    if (j == 0) { g.A = vb.t.hout; g.lda = hL; g.a_rows = vb.t.root_row; }
    else { g.A = vb.m[j - 1]; g.lda = net->mlp[j - 1].out; }
    g.B0 = P + net->mlp[j].w; g.ldb0 = net->mlp[j].in; g.b_split = net->mlp[j].in; g.bias = P + net->mlp[j].b;
    g.M = B; g.N = net->mlp[j].out; g.R = net->mlp[j].in; g.act = net->mlp[j].act;
    g.Y = vb.m[j]; g.ldy = (j < net->n_mlp - 1) ? net->mlp[j].out : vb.ldcat;
    g.splitk_ws = vb.sk; g.splitk_floats = vb.sk_floats;
    site(K_MLP, 0);
    SARD_TRY(sard_linear_fwd(&g, st));
  }
  SARD_REQUIRE(net->n_mlp > 0, "critic without mlp is not supported");
  SARD_TRY(side_join(side));
  SARD_REQUIRE(net->head.in == vb.ldcat && net->head.out == 1, "value head width");
  // value head + loss + the head's backward-data in one launch
  site(K_HEAD, 0);
  SARD_TRY(sard_value_head(vb.cat, vb.ldcat, P + net->head.w, P + net->head.b, B, vb.ldcat, vb.ldcat - 3 * SARD_EXT_DIM,
                           net->mlp[net->n_mlp - 1].act, sel->ids, bwd ? arena->returns : nullptr, inv_B_global, vb.v,
                           bwd ? vb.dv : nullptr, bwd ? vb.sq : nullptr, bwd ? nullptr : arena->values, bwd ? vb.dcat : nullptr,
                           vb.ldcat, st));
  if (!bwd) return 0;
  // the summed squared error only feeds the log: off the launching stream
  if (sq_err_sum) SARD_TRY(sard_loss_reduce(vb.sq, B, 1, sq_err_sum, nullptr, 0, nullptr, wg_stream(side, st, 1)));
  {   // head weight gradient
    SardWGrad w = wgrad0();
    w.P = vb.dv; w.ldp = 1; w.Q = vb.cat; w.ldq = vb.ldcat; w.M = B; w.N1 = 1; w.N2 = vb.ldcat;
    w.dW0 = G + net->head.w; w.ldw0 = vb.ldcat; w.w_split = vb.ldcat; w.db = G + net->head.b; w.partials = vb.wg2;
    site(K_HEAD, 2);
    SARD_TRY(sard_linear_bwd_weight(&w, wg_stream(side, st, 1)));
  }
  // the encoders' backward ends in weight gradients only: all of it runs beside the MLP / tower backward
  SARD_TRY(encoders_backward(net->enc, P, G, arena, sel, vb.e1, vb.dcat + (vb.ldcat - 3 * SARD_EXT_DIM), vb.ldcat,
                             vb.de1, vb.wg3, wg_stream(side, st, 2)));
  // MLP backward (dz of layer j lives in dzj with leading dimension ldz)
  const float* dzj = vb.dcat; int ldz = vb.ldcat;
  SARD_CUDA(cudaMemsetAsync(vb.t.dz[tw.L], 0, sizeof(float) * (size_t)n * hL, as_stream(st)));
  for (int j = net->n_mlp - 1; j >= 0; --j) {
    SardWGrad w = wgrad0();
    w.P = dzj; w.ldp = ldz; w.M = B; w.N1 = net->mlp[j].out; w.N2 = net->mlp[j].in;
    if (j == 0) { w.Q = vb.t.hout; w.ldq = hL; w.q_rows = vb.t.root_row; } else { w.Q = vb.m[j - 1]; w.ldq = net->mlp[j - 1].out; }
    w.dW0 = G + net->mlp[j].w; w.ldw0 = net->mlp[j].in; w.w_split = net->mlp[j].in; w.db = G + net->mlp[j].b; w.partials = vb.wg2;
    site(K_MLP, 2);
    SARD_TRY(sard_linear_bwd_weight(&w, wg_stream(side, st, 1)));
    SardGemm g = gemm0();
    g.A = dzj; g.lda = ldz; g.B0 = P + net->mlp[j].w; g.ldb0 = net->mlp[j].in; g.b_split = net->mlp[j].in;
    g.M = B; g.N = net->mlp[j].in; g.R = net->mlp[j].out; g.dsplit = g.N;
    if (j == 0) {
      // scatter into the root rows of the tower gradient; other rows stay zero
      g.Y = vb.t.dz[tw.L]; g.ldy = hL; g.y_rows = vb.t.root_row;
      g.Dact = vb.t.hout; g.ldd = hL; g.dact0 = tw.L > 0 ? tw.act : SARD_ACT_NONE; g.dact1 = g.dact0;
    } else {
      g.Y = vb.dm[j]; g.ldy = net->mlp[j - 1].out;
      g.Dact = vb.m[j - 1]; g.ldd = net->mlp[j - 1].out; g.dact0 = net->mlp[j - 1].act; g.dact1 = g.dact0;
    }
    g.splitk_ws = vb.sk; g.splitk_floats = vb.sk_floats;
    site(K_MLP, 1);
    if (j == 1 && vb.cmw && ldz % 4 == 0) {
      // dz1 = (dz2 W2) * act'(h1) in one tcgen05 launch (critic_tc.cu)
      SARD_TRY(sard_critic_tc_bwd_layer2(vb.cmw, dzj, ldz, vb.m[0], B, net->mlp[0].act, vb.dm[1], st));
    } else if (j == 0 && vb.cmw && dzj == vb.dm[1]) {
      // dz_tower[root rows] = (dz1 W1) * act'(hout[root rows]) likewise
      SARD_TRY(sard_critic_tc_bwd_layer1(vb.cmw, dzj, vb.t.hout, vb.t.root_row, B, tw.L > 0 ? tw.act : SARD_ACT_NONE, vb.t.dz[tw.L], st));
    } else
    SARD_TRY(sard_linear_bwd_data(&g, st));
    if (j > 0) { dzj = vb.dm[j]; ldz = net->mlp[j - 1].out; }
  }
  return tower_backward(tw, P, G, arena, sel, vb.t, vb.wg, side, st);
}

// ------------------------------------------------------------------------------------------ policy stage
namespace {
struct StageBufs {
  TowerBufs t;
  double* rec;
  float* e1[3]; float* ext;
  int *srow, *spos, *grp_ptr, *tiles, *chunks, *grp_chunk_ptr, *counts, *sort_scratch;
  float* xs; int ldxs; float* H[SARD_MAX_IL]; float* out; float* pre; int ldo;
  float* dout; float* gstat; float* dH[SARD_MAX_IL]; float* dxs; float* dext; float* de1; float* wg; float* wg2; float* wg3; float* encw;
  int max_tiles, max_chunks;
  bool tc;                       // fused tcgen05 JSMLP (jsmlp_tc.cu)
  bool tcb;                      // ... in the padded row space, with the fused backward (jsmlp_tc_bwd.cu)
  int rows_cap;                  // sorted positions (padded: every body index starts on a 128-row tile boundary)
  float* tcws; long long tcws_floats; float* Hhl[2];
  float* dOs_hl; float* dZhl[2]; float* tc_partials;
};

// the fused tensor-core JSMLP serves engine 4 (the default) when the stage's shapes are covered
bool stage_uses_tc(const SardPolicyStage& s) {
  return g_sard_gemm_engine == 4 && s.n_il == 3 &&
         sard_jsmlp_tc_supported(s.il[0].in, s.il[0].out, s.il[1].out, s.out_dim, s.n_il) && s.il[1].in == s.il[0].out &&
         s.il[2].in == s.il[1].out;
}
SardJsmlpTc jsmlp_tc_args(const SardPolicyStage& s, const StageBufs& sb, int n, bool bwd) {
  SardJsmlpTc a;
  memset(&a, 0, sizeof(a));
  a.n = sb.tcb ? sb.rows_cap : n; a.padded = sb.tcb ? 1 : 0; a.in_dim = s.il[0].in; a.hid = s.il[0].out; a.out_dim = s.out_dim; a.act = s.il_act; a.post = s.post;
  a.max_index = s.max_index; a.n_slots = s.n_live;
  a.xs = sb.xs; a.ld_xs = sb.ldxs;
  for (int j = 0; j < 3; ++j) { a.W[j] = s.il[j].W; a.b[j] = s.il[j].b; }
  a.slot_of_index = s.slot_of_index;
  a.tiles = sb.tiles; a.num_tiles = sb.counts; a.max_tiles = sb.max_tiles;
  a.srow = sb.srow;
  // the hidden activations only leave the SM when a backward pass follows
  a.H[0] = bwd && !sb.tcb ? sb.H[0] : nullptr; a.H[1] = bwd && !sb.tcb ? sb.H[1] : nullptr; a.ldh = s.il[0].out;
  a.Hhl[0] = bwd && sb.tcb ? sb.Hhl[0] : nullptr; a.Hhl[1] = bwd && sb.tcb ? sb.Hhl[1] : nullptr;
  a.out = sb.out; a.pre = s.post != SARD_POST_NONE ? sb.pre : nullptr; a.ldo = sb.ldo;
  a.ws = sb.tcws; a.ws_floats = sb.tcws_floats;
  return a;
}

void carve_stage(Bump& b, const SardPolicyNet& net, const SardPolicyStage& s, int n, int B, bool train, StageBufs& sb) {
  carve_tower(b, s.tower, n, B, train, sb.t);
  sb.rec = reinterpret_cast<double*>(b.f(2 * (1 + 2 * s.tower.D) + 2));
  for (int k = 0; k < 3; ++k) sb.e1[k] = b.f((long long)B * net.enc[k][0].out);
  sb.ext = b.f((long long)B * 3 * SARD_EXT_DIM);
  { const long long w = enc_wide_fwd_need(net.enc, B); sb.encw = w > 0 ? b.f(w) : nullptr; }
  // every present body index adds at most one partial tile / chunk
  const int live = s.n_live > 0 && s.n_live < s.max_index ? s.n_live : s.max_index;
  sb.tc = stage_uses_tc(s);
  sb.tcb = sb.tc && WG_CHUNK % IL_TILE == 0 &&
           sard_jsmlp_tc_bwd_supported(s.il[0].in, s.il[0].out, s.il[1].out, s.out_dim, s.n_il);
  sb.rows_cap = sb.tcb ? (n + IL_TILE - 1) / IL_TILE * IL_TILE + live * IL_TILE : n;
  sb.max_tiles = sb.tcb ? sb.rows_cap / IL_TILE : n / IL_TILE + live;
  sb.max_chunks = n / WG_CHUNK + live;
  sb.srow = b.i(sb.rows_cap); sb.spos = b.i(n); sb.grp_ptr = b.i(s.max_index + 1);
  sb.tiles = b.i(4LL * sb.max_tiles); sb.chunks = b.i(4LL * sb.max_chunks);
  sb.grp_chunk_ptr = b.i(s.max_index + 1); sb.counts = b.i(4);
  sb.sort_scratch = b.i((long long)(ceil_div(n > 0 ? n : 1, 256) + 2) * s.max_index);
  const int hL = s.tower.h[s.tower.L];
  sb.ldxs = hL + 3 * SARD_EXT_DIM;
  sb.xs = sb.tcb ? nullptr : b.f((long long)n * sb.ldxs);
  int maxo = 4;
  for (int j = 0; j < s.n_il - 1; ++j) { sb.H[j] = sb.tcb ? nullptr : b.f((long long)n * s.il[j].out); maxo = s.il[j].out > maxo ? s.il[j].out : maxo; }
  sb.ldo = r4(s.out_dim);
  sb.out = b.f((long long)n * sb.ldo);
  sb.pre = b.f((long long)n * sb.ldo);
  sb.tcws = nullptr; sb.tcws_floats = 0; sb.Hhl[0] = sb.Hhl[1] = nullptr;
  sb.dOs_hl = sb.dZhl[0] = sb.dZhl[1] = sb.tc_partials = nullptr;
  if (sb.tc) {
    sb.tcws_floats = sard_jsmlp_tc_workspace(sb.rows_cap, s.il[0].in, s.out_dim, s.n_live);
    sb.tcws = b.a256(sb.tcws_floats);
    if (train && sb.tcb) {
      for (int j = 0; j < 2; ++j) sb.Hhl[j] = b.a256((long long)sb.rows_cap * 2 * s.il[j].out);
      for (int j = 0; j < 2; ++j) sb.dZhl[j] = b.a256((long long)sb.rows_cap * 2 * s.il[j].out);
      sb.dOs_hl = b.a256((long long)sb.rows_cap * 64);
      sb.tc_partials = b.a256(sard_jsmlp_tc_partials_floats(sb.max_chunks));
    }
  }
  if (train) {
    sb.dout = b.f((long long)n * sb.ldo);
    sb.gstat = b.f((long long)B * (4 + s.out_dim));
    for (int j = 0; j < s.n_il; ++j) sb.dH[j] = j > 0 && !sb.tcb ? b.f((long long)n * s.il[j - 1].out) : nullptr;
    sb.dxs = sb.tcb ? b.a256((long long)sb.rows_cap * sb.ldxs) : b.f((long long)n * sb.ldxs);
    sb.dext = b.f((long long)B * 3 * SARD_EXT_DIM);
    sb.de1 = b.f((long long)B * SARD_EXT_DIM);
    sb.wg = b.f(tower_wg_need(s.tower, n));                 // side stream 0
    long long need = 4, v;                                   // side stream 1
    if (!sb.tcb)
      for (int j = 0; j < s.n_il; ++j) { v = (long long)sb.max_chunks * s.il[j].out * SARD_WG_LDPART(s.il[j].in); need = v > need ? v : need; }
    sb.wg2 = b.f(need);
    sb.wg3 = b.f(enc_wg_need(net.enc, B));                   // side stream 2
  } else {
    sb.dout = sb.gstat = sb.dxs = sb.dext = sb.de1 = sb.wg = sb.wg2 = sb.wg3 = nullptr;
    for (int j = 0; j < s.n_il; ++j) sb.dH[j] = nullptr;
  }
}
}  // namespace

extern "C" long long sard_policy_workspace(const SardPolicyNet* net, int stage, int n, int B) {
  Bump b{nullptr, 0};
  StageBufs sb;
  carve_stage(b, *net, net->stage[stage], n, B, true, sb);
  return b.off + 64;
}

extern "C" int sard_policy_run(const SardPolicyNet* net, int stage, const SardArena* arena, const SardSel* sel, int train,
                               int phase, float clip_eps, float inv_B_global, double* moments_local,
                               const double* moments_all, int n_records, float* stats, float* outputs,
                               float* log_prob_run, float* ws, long long ws_floats, sard_stream_t st) {
  if (sel->B <= 0 || sel->n <= 0) return 0;
  SARD_REQUIRE(stage >= 0 && stage < 3, "stage");
  const SardPolicyStage& s = net->stage[stage];
  const SardTower& tw = s.tower;
  SARD_REQUIRE(s.n_il >= 1 && s.n_il <= SARD_MAX_IL, "index-linear layer count");
  const bool upd = train != 0, bwd = train == 1;
  Bump b{ws, 0};
  StageBufs sb;
  carve_stage(b, *net, s, sel->n, sel->B, bwd, sb);
  SARD_REQUIRE(b.off <= ws_floats, "policy workspace too small");
  const float* P = net->P;
  float* G = net->G;
  const int n = sel->n, B = sel->B;
  if (phase & SARD_PHASE_GATHER) {
    const int nconst = (stage == SARD_STAGE_CONTROL) ? 0 : net->nconst;
    SARD_TRY(tower_gather(tw, arena, sel, s.lead, s.tail, nconst, net->consts, 0, sb.t, st));
    if (upd) {
      double* rec = moments_local ? moments_local : sb.rec;
      SARD_TRY(sard_norm_moments(sb.t.x, sb.t.ldx, n, tw.D, sb.t.partials, rec, st));
    }
    // the body-index sort only needs the gathered run: it belongs to the (prefetched) gather phase
    SARD_TRY(sard_sort_by_index_padded(sb.t.body, n, s.max_index, IL_TILE, WG_CHUNK, sb.tcb ? sb.rows_cap : 0, sb.srow, sb.spos,
                                       sb.grp_ptr, sb.tiles, sb.chunks, sb.grp_chunk_ptr, sb.counts, sb.sort_scratch, st));
  }
  if (!(phase & SARD_PHASE_COMPUTE)) return 0;
  {
    const double* recs = moments_all ? moments_all : (moments_local ? moments_local : sb.rec);
    SARD_TRY(tower_norm_step(tw, arena, n, sb.t, upd, recs, moments_all ? n_records : 1, st));
  }
  // the ObsEncoders only need the arena and the hi/lo pair copies of the live weight slabs only the weights (they change
  // with every optimiser step): both run beside the tower
  SideCtx* side = side_for(as_stream(st));
  SARD_TRY(encoders_forward(net->enc, P, arena, sel, sb.e1, sb.ext, 3 * SARD_EXT_DIM, sb.encw, wg_stream(side, st, 2)));
  if (sb.tc) {
    const SardJsmlpTc a = jsmlp_tc_args(s, sb, n, bwd);
    SARD_TRY(sard_jsmlp_tc_presplit(&a, wg_stream(side, st, 1)));
  }
  SARD_TRY(tower_forward(tw, P, arena, sel, sb.t, st));
  SARD_TRY(side_join(side));
  const int hL = tw.h[tw.L];
  SARD_REQUIRE(s.il[0].in == sb.ldxs, "first index-linear input width");
  if (sb.tcb) {
    // cat + repeat_interleave written straight into the fused kernels' operand format (sorted, padded, hi | lo)
    const SardJsmlpTc a = jsmlp_tc_args(s, sb, n, bwd);
    SARD_TRY(sard_concat_sorted_split(sb.t.hout, hL, hL, sb.ext, 3 * SARD_EXT_DIM, 3 * SARD_EXT_DIM, sb.srow, sb.t.nd_graph,
                                      sb.rows_cap, sard_jsmlp_tc_xs_hl(&a), st));
    site(K_IL, 0);
    SARD_TRY(sard_jsmlp_tc_forward(&a, st));
  } else {
    SARD_TRY(sard_concat_sorted(sb.t.hout, hL, hL, sb.ext, 3 * SARD_EXT_DIM, 3 * SARD_EXT_DIM, sb.srow, sb.t.nd_graph, n,
                                sb.xs, sb.ldxs, st));
  }
  if (sb.tcb) {
  } else if (sb.tc) {
    // the three IndexLinear layers in one tcgen05 launch (jsmlp_tc.cu)
    const SardJsmlpTc a = jsmlp_tc_args(s, sb, n, bwd);
    SARD_TRY(sard_jsmlp_tc_split_input(&a, st));
    site(K_IL, 0);
    SARD_TRY(sard_jsmlp_tc_forward(&a, st));
  } else
  for (int j = 0; j < s.n_il; ++j) {
    const bool last = j == s.n_il - 1;
    SardGemm g = gemm0();
    g.A = j == 0 ? sb.xs : sb.H[j - 1]; g.lda = j == 0 ? sb.ldxs : s.il[j - 1].out;
    g.B0 = s.il[j].W; g.ldb0 = s.il[j].in; g.b_split = s.il[j].in; g.b_group_stride = (long long)s.il[j].out * s.il[j].in;
    g.bias = s.il[j].b; g.bias_group_stride = s.il[j].out;
    g.M = n; g.N = s.il[j].out; g.R = s.il[j].in;
    g.tiles = sb.tiles; g.num_tiles = sb.counts; g.max_tiles = sb.max_tiles;
    if (last) {
      g.act = SARD_ACT_NONE; g.post = s.post; g.Y = sb.out; g.ldy = sb.ldo; g.y_rows = sb.srow;
      g.Ypre = s.post != SARD_POST_NONE ? sb.pre : nullptr;
      SARD_REQUIRE(s.il[j].out == s.out_dim, "last index-linear width");
    } else {
      g.act = s.il_act; g.Y = sb.H[j]; g.ldy = s.il[j].out;
    }
    site(K_IL, 0);
    SARD_TRY(sard_linear_fwd(&g, st));
  }
  // loss (eval: log-probs only)
  SardPolicyLoss L;
  memset(&L, 0, sizeof(L));
  L.out = sb.out; L.ld_out = sb.ldo; L.pre = s.post != SARD_POST_NONE ? sb.pre : nullptr;
  L.log_std = s.log_std >= 0 ? P + s.log_std : nullptr;
  L.d = s.out_dim;
  L.act_col = stage == SARD_STAGE_CONTROL ? 0 : (stage == SARD_STAGE_ATTR ? 1 : SARD_ACTION_DIM - 1);
  L.tie_c0 = s.tie_c0; L.tie_c1 = s.tie_c1;
  L.nd_arena = sb.t.nd_arena; L.cum = sel->cum; L.node_base = sel->node_base; L.ids = sel->ids; L.B = B;
  L.actions = arena->actions; L.rep_local = (s.tie_c1 > s.tie_c0) ? arena->rep_local : nullptr;
  L.clip_eps = clip_eps; L.inv_B = inv_B_global;
  L.tied_out = outputs; L.log_prob_run = log_prob_run;
  if (bwd) {
    L.advantages = arena->advantages; L.fixed_log_probs = arena->fixed_log_probs;
    L.dout = sb.dout; L.ld_dout = sb.ldo; L.gstat = sb.gstat;
    if (sb.tcb && g_sard_fuse_dout) {
      // the fused JSMLP backward reads the gradient in its own operand format: the loss kernel writes it there directly
      L.dout = nullptr;
      L.dout_sorted_hl = sb.dOs_hl; L.spos = sb.spos; L.srow = sb.srow; L.rows_cap = sb.rows_cap;
    }
  } else if (!upd) {
    L.log_prob_out = arena->fixed_log_probs;
  }
  if (outputs) SARD_REQUIRE(sb.ldo == s.out_dim || true, "outputs use the padded leading dimension");
  SARD_TRY(stage == SARD_STAGE_SKEL ? sard_cat_loss(&L, st) : sard_gauss_loss(&L, st));
  if (!bwd) return 0;
  // the column sums (loss statistics, log_std gradients) are only read by the optimiser step: off the chain, on the side
  // stream the JSMLP weight gradients follow on
  SARD_TRY(sard_loss_reduce(sb.gstat, B, 4 + s.out_dim, stats, s.log_std >= 0 ? G + s.log_std : nullptr, s.out_dim,
                            s.log_std >= 0 ? P + s.log_std : nullptr, g_sard_side_reduce ? wg_stream(side, st, 1) : st));

  // ---- backward through the JSMLP (sorted row space) ----
  const float* dz = sb.dout; int ldz = sb.ldo; const int* dz_rows = sb.srow;
  if (sb.tcb) {
    // dZ2, dZ1, dX in one tcgen05 launch; the weight gradients of the three tables in one more (+ ordered reduction)
    const SardJsmlpTc a = jsmlp_tc_args(s, sb, n, bwd);
    SardJsmlpTcBwd bb;
    memset(&bb, 0, sizeof(bb));
    bb.dout = g_sard_fuse_dout ? nullptr : sb.dout; bb.ld_dout = sb.ldo; bb.dOs_hl = sb.dOs_hl; bb.dZhl[0] = sb.dZhl[0]; bb.dZhl[1] = sb.dZhl[1];
    bb.dxs = sb.dxs; bb.ld_dxs = sb.ldxs;
    bb.chunks = sb.chunks; bb.num_chunks = sb.counts + 1; bb.max_chunks = sb.max_chunks; bb.grp_chunk_ptr = sb.grp_chunk_ptr;
    bb.partials = sb.tc_partials;
    for (int j = 0; j < 3; ++j) { bb.gW[j] = s.GT + s.il[j].gW; bb.gb[j] = s.GT + s.il[j].gb; }
    site(K_IL, 1);
    SARD_TRY(sard_jsmlp_tc_backward(&a, &bb, st));
    site(K_IL, 2);
    SARD_TRY(sard_jsmlp_tc_wgrad(&a, &bb, wg_stream(side, st, 1)));
  } else
  for (int j = s.n_il - 1; j >= 0; --j) {
    const float* Aj = j == 0 ? sb.xs : sb.H[j - 1];
    const int lda = j == 0 ? sb.ldxs : s.il[j - 1].out;
    SardWGrad w = wgrad0();
    w.P = dz; w.ldp = ldz; w.p_rows = dz_rows; w.Q = Aj; w.ldq = lda; w.M = n; w.N1 = s.il[j].out; w.N2 = s.il[j].in;
    w.chunks = sb.chunks; w.num_chunks = sb.counts + 1; w.max_chunks = sb.max_chunks;
    w.grp_chunk_ptr = sb.grp_chunk_ptr; w.n_groups = s.max_index; w.slot_of_group = s.slot_of_index;
    w.dW0 = s.GT + s.il[j].gW; w.ldw0 = s.il[j].in; w.w_split = s.il[j].in; w.w_slot_stride = (long long)s.il[j].out * s.il[j].in;
    w.db = s.GT + s.il[j].gb; w.b_slot_stride = s.il[j].out; w.partials = sb.wg2;
    site(K_IL, 2);
    SARD_TRY(sard_linear_bwd_weight(&w, wg_stream(side, st, 1)));
    SardGemm g = gemm0();
    g.A = dz; g.lda = ldz; g.a_rows = dz_rows;
    g.B0 = s.il[j].W; g.ldb0 = s.il[j].in; g.b_split = s.il[j].in; g.b_group_stride = (long long)s.il[j].out * s.il[j].in;
    g.M = n; g.N = s.il[j].in; g.R = s.il[j].out; g.dsplit = g.N;
    g.tiles = sb.tiles; g.num_tiles = sb.counts; g.max_tiles = sb.max_tiles;
    if (j == 0) { g.Y = sb.dxs; g.ldy = sb.ldxs; }
    else { g.Y = sb.dH[j]; g.ldy = s.il[j - 1].out; g.Dact = sb.H[j - 1]; g.ldd = s.il[j - 1].out; g.dact0 = s.il_act; g.dact1 = s.il_act; }
    site(K_IL, 1);
    SARD_TRY(sard_linear_bwd_data(&g, st));
    if (j > 0) { dz = sb.dH[j]; ldz = s.il[j - 1].out; dz_rows = nullptr; }
  }
  SARD_TRY(sard_unsort_grad(sb.dxs, sb.ldxs, sb.spos, sel->cum, sel->node_base, B, sb.t.hout, hL, hL,
                            tw.L > 0 ? tw.act : SARD_ACT_NONE, sb.t.dz[tw.L], hL, sb.ext, 3 * SARD_EXT_DIM, 3 * SARD_EXT_DIM,
                            sb.dext, 3 * SARD_EXT_DIM, st));
  SARD_TRY(encoders_backward(net->enc, P, G, arena, sel, sb.e1, sb.dext, 3 * SARD_EXT_DIM, sb.de1, sb.wg3, wg_stream(side, st, 2)));
  return tower_backward(tw, P, G, arena, sel, sb.t, sb.wg, side, st);
}
