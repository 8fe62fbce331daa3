// Fused critic MLP forward on the tcgen05 tensor cores: internal interface of critic_tc.cu, used by engine.cu.
#pragma once
#include "common.cuh"

bool sard_critic_tc_supported(int in0, int h1, int h2, int n_mlp);      // 64 -> 512 -> 256 (derl.yml value_specs.mlp)
long long sard_critic_tc_wsplit_floats(void);
// hi | lo pair copies of W1 [512, 64] and W2 [256, 512] in the kernel's block order (once per optimiser step)
int sard_critic_tc_presplit(const float* W1, const float* W2, float* wsplit, sard_stream_t stream);
// h1 = act(hout[root_row] W1^T + b1) -> H1 [B, 512] (optional); h2 = act(h1 W2^T + b2) -> cat[:, 0:256]
int sard_critic_tc_forward(const float* wsplit, const float* hout, const int* root_row, int B, const float* b1, const float* b2, int act,
                           float* H1, float* cat, int ldcat, sard_stream_t stream);
// dZ1 [B, 512] = (dz2 [B, 256] W2) * act'(H1): backward-data of layer 2 fused with the derivative of layer 1's activation
int sard_critic_tc_bwd_layer2(const float* wsplit, const float* dz2, int ld_dz2, const float* H1, int B, int act, float* dZ1,
                              sard_stream_t stream);
// dz_tower[root_row[b]] [64] = (dZ1 [B, 512] W1) * act'(hout[root_row[b]]): backward-data of layer 1, scattered into the root rows
int sard_critic_tc_bwd_layer1(const float* wsplit, const float* dZ1, const float* hout, const int* root_row, int B, int act, float* dz_tower,
                              sard_stream_t stream);
