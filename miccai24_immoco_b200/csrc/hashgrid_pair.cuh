// Lane-pair index arithmetic of the hash-grid kernels, shared by hashgrid.cu and the fused MLP-backward scatter
// (mlp_tc.cu): two adjacent lanes take the two dim-0 corners of one point.
#pragma once
#include "common.cuh"

// Index arithmetic specialised per level kind (uniform per CTA item, so the branch is free):
//   HASH : idx = (q0 ^ q1*P1 ^ q2*P2) & (entries-1)      entries a power of two
//   DENSE: idx = (q0 + q1*res + q2*res^2) & (entries-1)  entries a power of two
//   ANY  : grid_index() with its general modulo (odd resolutions, non-power-of-two tables)
// The per-dimension terms are computed once per point; a corner costs one or two XOR/ADDs and an AND.
enum { kIdxHash = 0, kIdxDense = 1, kIdxAny = 2 };

template <int D, int MODE>
struct PairTerms {
  uint32_t t[D][2];     // t[d][bit]: contribution of corner bit `bit` of dimension d (d >= 1)
  uint32_t q0, mask, entries, res, hashed, swz;
  uint32_t cell[D];
  __device__ __forceinline__ void init(const uint32_t (&c)[D], int half, uint32_t entries_, uint32_t res_,
                                       uint32_t hashed_, uint32_t swz_) {
    entries = entries_; res = res_; hashed = hashed_; mask = entries_ - 1u; swz = swz_;
    q0 = c[0] + (uint32_t)half;
    uint32_t mul = 1u;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      cell[d] = c[d];
      if (MODE == kIdxHash) mul = (d == 0) ? 1u : (d == 1 ? 2654435761u : 805459861u);
      t[d][0] = c[d] * mul;
      t[d][1] = (c[d] + 1u) * mul;
      if (MODE == kIdxDense) mul *= res_;
    }
  }
  // corner bits of dimensions 1.. in `c` (bit d-1 = dimension d); dimension 0 is this lane's half
  __device__ __forceinline__ uint32_t index(int c) const {
    if (MODE == kIdxAny) {
      uint32_t q[D];
      q[0] = q0;
#pragma unroll
      for (int d = 1; d < D; ++d) q[d] = cell[d] + (uint32_t)((c >> (d - 1)) & 1);
      return grid_index<D>(q, hashed, entries, res, swz);
    }
    uint32_t idx = q0;      // prime 1 / stride 1
#pragma unroll
    for (int d = 1; d < D; ++d) {
      const uint32_t term = t[d][(c >> (d - 1)) & 1];
      idx = (MODE == kIdxHash) ? (idx ^ term) : (idx + term);
    }
    idx &= mask;
    return (MODE == kIdxHash) ? grid_swizzle(idx, swz) : idx;
  }
};

template <int D>
__device__ __forceinline__ float pair_weight(const float (&frac)[D], float w0, int c) {
  float w = w0;
#pragma unroll
  for (int d = 1; d < D; ++d) w = w * (((c >> (d - 1)) & 1) ? frac[d] : 1.0f - frac[d]);   // (w0 * w1) * w2
  return w;
}

