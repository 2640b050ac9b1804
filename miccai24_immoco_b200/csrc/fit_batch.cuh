// Instance dimension of the fit loop's small kernels ("within a GPU, batch B
// instances").  The row / column passes, the gradient-entropy kernel, the fixed-point finalize and the MLP
// forward kernels of ONE 320x320 slice are latency-bound (160 - 1600 CTAs, 8 - 30 us each); with
// blockIdx.z (MLP forward: blockIdx.y) = instance the same launch carries B independent slices of the same
// shape, each with its own buffers, masks and loss slots.  The struct is passed by value as a
// __grid_constant__ kernel parameter; B = 1 is the ordinary single-slice fit.
#pragma once
#include <stdint.h>

#include "immoco_b200.h"

constexpr int kMaxFitBatch = 8;

struct FitBatchInst {
  const float2* image;      // (H, W) complex image prior
  const float2* disp;       // (M, H, W, 2) displacements (tanh output)
  float2* c_tmp;            // row-pass result, input of the column pass
  const float2* k_in;
  float2* k_out;
  float2* d_c;
  float2* d_image;
  float2* d_disp;
  long long* fx;            // deterministic mode: fixed-point image-cotangent plane (else nullptr)
  uint32_t* dmax;           // deterministic mode: this iteration's max |d_c| word (else nullptr)
  double* loss_dc;          // loss accumulators of this iteration (atomic path) ...
  double* loss_ge;
  double* slots_dc;         // ... or per-CTA slots (fit loop)
  double* slots_ge;
  immoco_lines lines;
};

struct FitBatch {
  int n;
  FitBatchInst inst[kMaxFitBatch];
};

struct MlpFwdBatch {
  int n;
  const float2* enc[kMaxFitBatch];
  const float* w1[kMaxFitBatch];
  const float* w2[kMaxFitBatch];
  float2* out[kMaxFitBatch];
};
