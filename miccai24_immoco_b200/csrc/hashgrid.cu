// Multiresolution hash-grid encoding, forward gather and backward scatter.
// Replaces tiny-cuda-nn's kernel_grid / kernel_grid_backward used through
// tcnn.NetworkWithInputEncoding at src/models/immoco.py:60-65.
//
// Layout: table rows are float2 (n_features_per_level == 2); enc[level][point] float2 planes, so
// a warp handling 32 consecutive points of one level reads/writes 256 contiguous bytes.
// Grid = (point tiles, levels): CTAs of one level are adjacent in launch order, so one level's
// table (<= 4 MB) is the L2 working set at any time.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

template <int D>
__global__ void __launch_bounds__(kThreads)
hashgrid_fwd_kernel(const __grid_constant__ immoco_grid_desc g, const float* __restrict__ coords,
                    const float2* __restrict__ table, float2* __restrict__ enc, int n) {
  const int level = blockIdx.y;
  const float scale = g.scale[level];
  const uint32_t res = g.resolution[level];
  const uint32_t entries = g.entries[level];
  const uint32_t hashed = g.hashed[level];
  const float2* __restrict__ tab = table + g.offset[level];

  for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
    uint32_t cell[D];
    float frac[D];
#pragma unroll
    for (int d = 0; d < D; ++d) grid_pos(__ldg(coords + (size_t)i * D + d), scale, cell[d], frac[d]);

    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < (1 << D); ++c) {
      uint32_t q[D];
      float w = 1.0f;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const int bit = (c >> d) & 1;
        q[d] = cell[d] + (uint32_t)bit;
        // same association order as the oracle: w = w0 * w1 * w2
        w = (d == 0) ? (bit ? frac[0] : 1.0f - frac[0]) : w * (bit ? frac[d] : 1.0f - frac[d]);
      }
      const float2 v = __ldg(tab + grid_index<D>(q, hashed, entries, res));
      acc.x = fmaf(w, v.x, acc.x);
      acc.y = fmaf(w, v.y, acc.y);
    }
    enc[(size_t)level * n + i] = acc;
  }
}

template <int D>
__global__ void __launch_bounds__(kThreads)
hashgrid_bwd_kernel(const __grid_constant__ immoco_grid_desc g, const float* __restrict__ coords,
                    const float2* __restrict__ d_enc, float2* __restrict__ grad_table, int n) {
  const int level = blockIdx.y;
  const float scale = g.scale[level];
  const uint32_t res = g.resolution[level];
  const uint32_t entries = g.entries[level];
  const uint32_t hashed = g.hashed[level];
  float2* __restrict__ gtab = grad_table + g.offset[level];

  for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
    const float2 go = __ldg(d_enc + (size_t)level * n + i);
    if (go.x == 0.0f && go.y == 0.0f) continue;  // adding +-0 is a no-op for the scatter
    uint32_t cell[D];
    float frac[D];
#pragma unroll
    for (int d = 0; d < D; ++d) grid_pos(__ldg(coords + (size_t)i * D + d), scale, cell[d], frac[d]);
#pragma unroll
    for (int c = 0; c < (1 << D); ++c) {
      uint32_t q[D];
      float w = 1.0f;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const int bit = (c >> d) & 1;
        q[d] = cell[d] + (uint32_t)bit;
        w = (d == 0) ? (bit ? frac[0] : 1.0f - frac[0]) : w * (bit ? frac[d] : 1.0f - frac[d]);
      }
      // one 64-bit vector reduction per corner (RED.ADD.F32x2 on sm_90+)
      atomicAdd(gtab + grid_index<D>(q, hashed, entries, res), make_float2(w * go.x, w * go.y));
    }
  }
}

int check(const immoco_grid_desc* g, int64_t n) {
  if (!g || n < 0 || n > (int64_t)0x7fffffff / 2) return IMMOCO_ERR_BAD_ARG;
  if (g->n_levels < 1 || g->n_levels > IMMOCO_MAX_LEVELS) return IMMOCO_ERR_BAD_ARG;
  if (g->n_dims != 2 && g->n_dims != 3) return IMMOCO_ERR_UNSUPPORTED;
  return 0;
}

}  // namespace

extern "C" int immoco_hashgrid_fwd(const immoco_grid_desc* grid, const float* coords,
                                   const float* table, float* enc, int64_t n_points, void* stream) {
  if (int e = check(grid, n_points)) return e;
  if (n_points == 0) return 0;
  const int n = (int)n_points;
  dim3 g((unsigned)ceil_div64(n, kThreads), (unsigned)grid->n_levels);
  cudaStream_t s = (cudaStream_t)stream;
  if (grid->n_dims == 2)
    hashgrid_fwd_kernel<2><<<g, kThreads, 0, s>>>(*grid, coords, (const float2*)table, (float2*)enc, n);
  else
    hashgrid_fwd_kernel<3><<<g, kThreads, 0, s>>>(*grid, coords, (const float2*)table, (float2*)enc, n);
  IMMOCO_LAUNCH_CHECK();
  return 0;
}

extern "C" int immoco_hashgrid_bwd(const immoco_grid_desc* grid, const float* coords,
                                   const float* d_enc, float* grad_table, int64_t n_points,
                                   void* stream) {
  if (int e = check(grid, n_points)) return e;
  if (n_points == 0) return 0;
  const int n = (int)n_points;
  dim3 g((unsigned)ceil_div64(n, kThreads), (unsigned)grid->n_levels);
  cudaStream_t s = (cudaStream_t)stream;
  if (grid->n_dims == 2)
    hashgrid_bwd_kernel<2><<<g, kThreads, 0, s>>>(*grid, coords, (const float2*)d_enc,
                                                  (float2*)grad_table, n);
  else
    hashgrid_bwd_kernel<3><<<g, kThreads, 0, s>>>(*grid, coords, (const float2*)d_enc,
                                                  (float2*)grad_table, n);
  IMMOCO_LAUNCH_CHECK();
  return 0;
}
