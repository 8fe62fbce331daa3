// Fused GNN tower (RunningNorm -> in_fc -> L x GraphConv) on the tcgen05 tensor cores: internal interface of
// tower_tc.cu, used by the network schedules in engine.cu.
#pragma once
#include "common.cuh"

#define TT_H 64                 // width of in_fc and of every GraphConv layer the fused kernels serve
#define TT_MAXL 3
#define TT_TILE 128
#define TT_MAX_NODES 64         // graphs are packed whole into 128-row tiles

struct TowerTcBufs {
  float* wsplit;                // hi|lo pair copies of the weights (sard_tower_tc_wsplit_floats)
  int* tile_first;              // [num_tiles + 1] first graph of every row tile
  unsigned long long *tab_in, *tab_out;   // [n] per-row neighbour words (degree + tile-local neighbour rows), by destination / by source
  unsigned* mask[TT_MAXL];      // [n, 2] ReLU masks of h_l (train)
  int num_tiles, span;          // tile t holds the graphs whose first row lies in [span*t, span*(t+1))
  int chunk_rows, num_chunks;   // weight-gradient row chunks
  // training only: operands of the weight gradients, fp32 (tower_wgrad_tc_kernel splits them into hi | lo on the SM)
  float* XH;                    // [n, 128]  normalised input (+ a column of ones at D: the in_fc bias gradient)
  float* PH[TT_MAXL];           // [n, 64]   h_l, l = 0..L-1
  float* PA[TT_MAXL];           // [n, 128]  [A^T dz_l | dz_l] of GraphConv layer l = 1..L at index l-1
  float* PZ0;                   // [n, 64]   dz_0 (gradient w.r.t. in_fc's output)
  float* partials;              // [num_chunks, 128, 448]
};

// RunningNorm step fused into the pre-split launch: mode 0 none, 1 running update from `records`, 2 eval (inv from std)
struct TowerNormStep {
  int mode, n_records;
  const double* records;
  float *mean, *var, *std, *inv;
  long long* count;
};

bool sard_tower_tc_supported(const SardTower& tw, int max_nodes);
inline int sard_tower_tc_span(int max_nodes) { return TT_TILE - (max_nodes > 1 ? max_nodes - 1 : 0); }
long long sard_tower_tc_wsplit_floats(int L);
long long sard_tower_tc_partials_floats(int num_chunks);
int sard_tower_tc_chunk_rows(int n);
// tile table of a run (gather phase): tile_first[t] for t <= num_tiles
int sard_tower_tc_tiles(const SardArena* arena, const SardSel* sel, const TowerTcBufs& tb, const int* nd_graph, const int* nd_arena,
                        sard_stream_t stream);
int sard_tower_tc_presplit(const SardTower& tw, const float* P, float* wsplit, const TowerNormStep& norm, sard_stream_t stream);
// x: gathered, un-normalised node features [n, ldx]; hout: [n, 64] output of the last layer
int sard_tower_tc_forward(const SardTower& tw, const float* P, const SardArena* arena, const SardSel* sel, const TowerTcBufs& tb,
                          const float* x, int ldx, const int* nd_graph, const int* nd_arena, float* hout, bool train,
                          sard_stream_t stream);
// dzL: [n, 64] gradient w.r.t. the pre-activation of the last layer
int sard_tower_tc_backward(const SardTower& tw, const SardArena* arena, const SardSel* sel, const TowerTcBufs& tb, const float* dzL,
                           const int* nd_graph, const int* nd_arena, sard_stream_t stream);
int sard_tower_tc_wgrad(const SardTower& tw, const SardSel* sel, const TowerTcBufs& tb, float* G, sard_stream_t stream);
