// Parameter blocks of the fused ObsEncoder kernels (encoder.cu), shared with the schedules (engine.cu).
#pragma once
#include <cuda_runtime.h>

struct EncFwdP {
  const float* src[3]; int K[3]; const int* ids; int B;
  const float* w1[3]; const float* b1[3]; const float* w2[3]; const float* b2[3];
  float* e1[3]; float* ext; int ld_ext; int enc_list[3]; int n_enc;
};


struct EncBwdP {
  const float* src[3]; int K[3]; const int* ids; int B;
  const float* w2[3]; const float* e1[3]; const float* dext; int ld_dext;
  float* partials;                     // [n_chunks][n_enc][ENC_PART]
  int enc_list[3]; int n_enc;
  float* gw1[3]; float* gb1[3]; float* gw2[3]; float* gb2[3];
};

int sard_enc_fused_ok(int K, int h1, int h2);
long long sard_enc_partials_floats(int B, int n_enc);
int sard_enc_fwd_launch(const EncFwdP& p, cudaStream_t st);
int sard_enc_bwd_launch(const EncBwdP& p, cudaStream_t st);

// first layer with a wide input (K > 128, e.g. the 1410-sample height field): split-K kernels, encoder.cu
int sard_enc_wide_ok(int K, int h1);
long long sard_enc_wide_partials_floats(int B, int K);
int sard_enc_wide_fwd(const float* src, int K, const int* ids, int B, const float* w1, const float* b1, float* e1, float* partials,
                      cudaStream_t st);
int sard_enc_wide_wgrad(const float* src, int K, const int* ids, int B, const float* de1, float* gw1, float* gb1, float* partials,
                        cudaStream_t st);
