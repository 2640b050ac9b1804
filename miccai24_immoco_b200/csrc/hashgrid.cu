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
                    const float2* __restrict__ table, float2* __restrict__ enc, int n, int level0) {
  const int level = level0 + blockIdx.y;
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

// Backward scatter.  Hashed (fine) levels: every corner of every point is its own table row, one
// RED.ADD.F32x2 each.  Dense (coarse) levels: the 32 consecutive points of a warp fall into a few
// cells, so equal-row runs of lanes are summed with shuffles first and only the run heads issue the
// reduction (level 0 of the 3-D grid: ~10 points per cell -> ~9x fewer same-address atomics, which
// otherwise serialise in L2: 86 us -> see profiles/round1_hashgrid_levels.txt).
template <int D>
__global__ void __launch_bounds__(kThreads)
hashgrid_bwd_kernel(const __grid_constant__ immoco_grid_desc g, const float* __restrict__ coords,
                    const float2* __restrict__ d_enc, float2* __restrict__ grad_table, int n, int level0) {
  const int level = level0 + blockIdx.y;
  const float scale = g.scale[level];
  const uint32_t res = g.resolution[level];
  const uint32_t entries = g.entries[level];
  const uint32_t hashed = g.hashed[level];
  float2* __restrict__ gtab = grad_table + g.offset[level];
  const unsigned lane = threadIdx.x & 31u;

  for (int base = blockIdx.x * kThreads; base < n; base += gridDim.x * kThreads) {
    const int i = base + threadIdx.x;
    const bool valid = i < n;
    float2 go = make_float2(0.f, 0.f);
    uint32_t cell[D];
    float frac[D];
    if (valid) {
      go = __ldg(d_enc + (size_t)level * n + i);
#pragma unroll
      for (int d = 0; d < D; ++d) grid_pos(__ldg(coords + (size_t)i * D + d), scale, cell[d], frac[d]);
    } else {
#pragma unroll
      for (int d = 0; d < D; ++d) { cell[d] = 0; frac[d] = 0.f; }
    }
    const bool live = valid && !(go.x == 0.0f && go.y == 0.0f);   // adding +-0 is a no-op
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
      const uint32_t idx = grid_index<D>(q, hashed, entries, res);
      float vx = w * go.x, vy = w * go.y;
      if (hashed) {
        // one 64-bit vector reduction per corner (RED.ADD.F32x2)
        if (live) atomicAdd(gtab + idx, make_float2(vx, vy));
      } else {
        // contiguous runs of equal rows within the warp -> one reduction per run
        const uint32_t key = live ? idx : 0xFFFFFFFFu;
        const uint32_t prev = __shfl_up_sync(0xffffffffu, key, 1);
        const bool head = (lane == 0) || (prev != key);
        const unsigned heads = __ballot_sync(0xffffffffu, head);
        const unsigned after = (lane == 31) ? 0u : (heads & ~((2u << lane) - 1u));
        const int run_end = after ? (__ffs(after) - 2) : 31;
#pragma unroll
        for (int ofs = 1; ofs < 32; ofs <<= 1) {
          const float ox = __shfl_down_sync(0xffffffffu, vx, ofs);
          const float oy = __shfl_down_sync(0xffffffffu, vy, ofs);
          if ((int)lane + ofs <= run_end) { vx += ox; vy += oy; }
        }
        if (head && live) atomicAdd(gtab + idx, make_float2(vx, vy));
      }
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

static int run_fwd(const immoco_grid_desc* grid, const float* coords, const float* table, float* enc,
                   int64_t n_points, int l0, int l1, void* stream) {
  if (int e = check(grid, n_points)) return e;
  if (l0 < 0 || l1 > grid->n_levels || l0 > l1) return IMMOCO_ERR_BAD_ARG;
  if (n_points == 0 || l0 == l1) return 0;
  const int n = (int)n_points;
  dim3 g((unsigned)ceil_div64(n, kThreads), (unsigned)(l1 - l0));
  cudaStream_t s = (cudaStream_t)stream;
  if (grid->n_dims == 2)
    hashgrid_fwd_kernel<2><<<g, kThreads, 0, s>>>(*grid, coords, (const float2*)table, (float2*)enc, n, l0);
  else
    hashgrid_fwd_kernel<3><<<g, kThreads, 0, s>>>(*grid, coords, (const float2*)table, (float2*)enc, n, l0);
  IMMOCO_LAUNCH_CHECK();
  return 0;
}

static int run_bwd(const immoco_grid_desc* grid, const float* coords, const float* d_enc, float* grad_table,
                   int64_t n_points, int l0, int l1, void* stream) {
  if (int e = check(grid, n_points)) return e;
  if (l0 < 0 || l1 > grid->n_levels || l0 > l1) return IMMOCO_ERR_BAD_ARG;
  if (n_points == 0 || l0 == l1) return 0;
  const int n = (int)n_points;
  dim3 g((unsigned)ceil_div64(n, kThreads), (unsigned)(l1 - l0));
  cudaStream_t s = (cudaStream_t)stream;
  if (grid->n_dims == 2)
    hashgrid_bwd_kernel<2><<<g, kThreads, 0, s>>>(*grid, coords, (const float2*)d_enc, (float2*)grad_table, n, l0);
  else
    hashgrid_bwd_kernel<3><<<g, kThreads, 0, s>>>(*grid, coords, (const float2*)d_enc, (float2*)grad_table, n, l0);
  IMMOCO_LAUNCH_CHECK();
  return 0;
}

extern "C" int immoco_hashgrid_fwd(const immoco_grid_desc* grid, const float* coords,
                                   const float* table, float* enc, int64_t n_points, void* stream) {
  return run_fwd(grid, coords, table, enc, n_points, 0, grid ? grid->n_levels : 0, stream);
}

extern "C" int immoco_hashgrid_bwd(const immoco_grid_desc* grid, const float* coords,
                                   const float* d_enc, float* grad_table, int64_t n_points,
                                   void* stream) {
  return run_bwd(grid, coords, d_enc, grad_table, n_points, 0, grid ? grid->n_levels : 0, stream);
}

// level-range variants (profiling / per-level checks): levels [level_begin, level_end)
extern "C" int immoco_hashgrid_fwd_levels(const immoco_grid_desc* grid, const float* coords, const float* table,
                                          float* enc, int64_t n_points, int32_t level_begin, int32_t level_end,
                                          void* stream) {
  return run_fwd(grid, coords, table, enc, n_points, level_begin, level_end, stream);
}
extern "C" int immoco_hashgrid_bwd_levels(const immoco_grid_desc* grid, const float* coords, const float* d_enc,
                                          float* grad_table, int64_t n_points, int32_t level_begin,
                                          int32_t level_end, void* stream) {
  return run_bwd(grid, coords, d_enc, grad_table, n_points, level_begin, level_end, stream);
}
