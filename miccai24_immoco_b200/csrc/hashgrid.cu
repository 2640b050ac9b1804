// Multiresolution hash-grid encoding, forward gather and backward scatter.
// Replaces tiny-cuda-nn's kernel_grid / kernel_grid_backward used through
// tcnn.NetworkWithInputEncoding at src/models/immoco.py:60-65.
//
// Layout: table rows are float2 (n_features_per_level == 2); enc[level][point] float2 planes, so
// a warp handling 32 consecutive points of one level reads/writes 256 contiguous bytes.
// Grid = (point tiles, levels): CTAs of one level are adjacent in launch order, so one level's
// table (<= 4 MB) is the L2 working set at any time.
#include "common.cuh"
#include "hashgrid_pair.cuh"

namespace {

// threads per CTA of every hash-grid kernel (iteration on B200: 128 -> 621 us, 256 -> 623 us, 512 -> 648 us)
#ifndef IMMOCO_HG_THREADS
#define IMMOCO_HG_THREADS 256
#endif
constexpr int kThreads = IMMOCO_HG_THREADS;

// table-row gather of the forward kernels.  Rows are re-used ~6x per pass but by unrelated pixels, so L1
// never hits beyond the lane pair itself: IMMOCO_HG_LOAD selects the cache policy of the gather
// (0: ld.global.nc, 1: ld.global.cg -- L2 only, 2: ld.global.nc.L1::no_allocate)
#ifndef IMMOCO_HG_LOAD
#define IMMOCO_HG_LOAD 0
#endif
__device__ __forceinline__ float2 load_row(const float2* p) {
#if IMMOCO_HG_LOAD == 1
  return __ldcg(p);
#elif IMMOCO_HG_LOAD == 2
  float2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
#else
  return __ldg(p);
#endif
}

template <int D>
__global__ void __launch_bounds__(kThreads)
hashgrid_fwd_kernel(const __grid_constant__ immoco_grid_desc g, const float* __restrict__ coords,
                    const float2* __restrict__ table, float2* __restrict__ enc, int n, int level0) {
  pdl_wait();
  const int level = level0 + blockIdx.y;
  const float scale = g.scale[level];
  const uint32_t res = g.resolution[level];
  const uint32_t entries = g.entries[level];
  const uint32_t hashed = g.hashed[level];
  const uint32_t swz = g.swizzle[level];
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
      const float2 v = __ldg(tab + grid_index<D>(q, hashed, entries, res, swz));
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
  pdl_wait();
  const int level = level0 + blockIdx.y;
  const float scale = g.scale[level];
  const uint32_t res = g.resolution[level];
  const uint32_t entries = g.entries[level];
  const uint32_t hashed = g.hashed[level];
  const uint32_t swz = g.swizzle[level];
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
      const uint32_t idx = grid_index<D>(q, hashed, entries, res, swz);
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
        // log-step segmented sum; stops as soon as no run of the warp extends past the current offset
        // (runs are ~10 / 5 / 2.5 lanes long at the three dense levels of the 3-D grid: 4 / 3 / 2 rounds
        // instead of 5)
#pragma unroll
        for (int ofs = 1; ofs < 32; ofs <<= 1) {
          const bool take = (int)lane + ofs <= run_end;
          if (!__any_sync(0xffffffffu, take)) break;
          const float ox = __shfl_down_sync(0xffffffffu, vx, ofs);
          const float oy = __shfl_down_sync(0xffffffffu, vy, ofs);
          if (take) { vx += ox; vy += oy; }
        }
        if (head && live) atomicAdd(gtab + idx, make_float2(vx, vy));
      }
    }
  }
}


// ---- lane-pair kernels -----------------------------------------------------------------------
// The first input dimension carries hash prime 1, so the two corners q0 / q0+1 of that dimension sit
// in rows idx and idx ^ (2^k - 1): the same 128-byte line 15 times out of 16 (the same 32-byte sector
// half of the time).  Two ADJACENT LANES therefore take the two dim-0 corners of one point: every
// gather / reduction instruction of a warp touches 16 lines instead of 32, which is what the L1TEX
// tag stage (1 line per cycle per SM) and the L2 atomic units are paced by.  The partial sums of a
// lane pair are combined with one shuffle.
constexpr int kPairPoints = kThreads / 2;

// IMMOCO_HG_FWD_PTS points per lane pair and item (the gathers of all of them are issued before the first
// use: 4 x PTS rows in flight per thread)
#ifndef IMMOCO_HG_FWD_PTS
#define IMMOCO_HG_FWD_PTS 2      // B200, 3-D grid at C2: 1 -> 130 us, 2 -> 125 us, 4 -> 128 us (tools/hg_levels.py)
#endif
constexpr int kFwdPts = IMMOCO_HG_FWD_PTS;

template <int D, int MODE>
__device__ __forceinline__ void fwd_pair_item(const float* __restrict__ coords, const float2* __restrict__ tab,
                                              float2* __restrict__ enc_level, int n, int base, float scale,
                                              uint32_t res, uint32_t entries, uint32_t hashed, uint32_t swz) {
  const int half = threadIdx.x & 1;
  constexpr int NC = 1 << (D - 1);
  float2 v[kFwdPts][NC];
  float frac[kFwdPts][D];
#pragma unroll
  for (int p = 0; p < kFwdPts; ++p) {
    const int i = base + p * kPairPoints + (threadIdx.x >> 1);
    uint32_t cell[D];
#pragma unroll
    for (int d = 0; d < D; ++d) { cell[d] = 0; frac[p][d] = 0.f; }
    if (i < n) {
#pragma unroll
      for (int d = 0; d < D; ++d) grid_pos(__ldg(coords + (size_t)i * D + d), scale, cell[d], frac[p][d]);
    }
    PairTerms<D, MODE> pt;
    pt.init(cell, half, entries, res, hashed, swz);
#pragma unroll
    for (int c = 0; c < NC; ++c) v[p][c] = (i < n) ? load_row(tab + pt.index(c)) : make_float2(0.f, 0.f);   // all gathers in flight
  }
#pragma unroll
  for (int p = 0; p < kFwdPts; ++p) {
    const int i = base + p * kPairPoints + (threadIdx.x >> 1);
    const float w0 = half ? frac[p][0] : 1.0f - frac[p][0];
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const float w = pair_weight<D>(frac[p], w0, c);
      acc.x = fmaf(w, v[p][c].x, acc.x);
      acc.y = fmaf(w, v[p][c].y, acc.y);
    }
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 1);
    acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 1);
    if (i < n && half == 0) enc_level[i] = acc;
  }
}

// resident CTAs per SM the forward kernel is compiled for: 6 -> 40 registers with two points per lane pair
// (5 CTAs at the unconstrained 44).  Iteration on B200: unconstrained 625 us, 6: 620.5, 7 / 8 (32 registers,
// 24 bytes of spills): 619-622 (gpurun_out/r104)
#ifndef IMMOCO_HG_FWD_MIN_CTAS
#define IMMOCO_HG_FWD_MIN_CTAS (1536 / IMMOCO_HG_THREADS)
#endif
template <int D>
__global__ void __launch_bounds__(kThreads, IMMOCO_HG_FWD_MIN_CTAS)
hashgrid_fwd_pair_kernel(const __grid_constant__ immoco_grid_desc g, const float* __restrict__ coords,
                         const float2* __restrict__ table, float2* __restrict__ enc, int n, int level0,
                         int n_levels, int tiles) {
  pdl_wait();
  // CTAs walk (level, tile) items in level-major order: one level's table (<= 4 MB) is the L2 working
  // set at any time; with a capped grid (immoco_set_hashgrid_ctas_per_sm) the CTAs are persistent
  for (int item = blockIdx.x; item < n_levels * tiles; item += gridDim.x) {
    const int level = level0 + item / tiles;
    const int base = (item % tiles) * (kPairPoints * kFwdPts);
    const float scale = g.scale[level];
    const uint32_t res = g.resolution[level];
    const uint32_t entries = g.entries[level];
    const uint32_t hashed = g.hashed[level];
    const float2* __restrict__ tab = table + g.offset[level];
    float2* __restrict__ out = enc + (size_t)level * n;
    const bool pow2 = (entries & (entries - 1u)) == 0u;
    const uint32_t swz = g.swizzle[level];
    if (pow2 && hashed) fwd_pair_item<D, kIdxHash>(coords, tab, out, n, base, scale, res, entries, hashed, swz);
    else if (pow2) fwd_pair_item<D, kIdxDense>(coords, tab, out, n, base, scale, res, entries, hashed, 0u);
    else fwd_pair_item<D, kIdxAny>(coords, tab, out, n, base, scale, res, entries, hashed, 0u);
  }
}

#ifndef IMMOCO_HG_BWD_PTS
#define IMMOCO_HG_BWD_PTS 1
#endif
constexpr int kBwdPts = IMMOCO_HG_BWD_PTS;

template <int D, int MODE>
__device__ __forceinline__ void bwd_pair_item(const float* __restrict__ coords, const float2* __restrict__ d_enc_level,
                                              float2* __restrict__ gtab, int n, int base, float scale,
                                              uint32_t res, uint32_t entries, uint32_t swz) {
  const int half = threadIdx.x & 1;
  float2 go[kBwdPts];
  float x[kBwdPts][D];
#pragma unroll
  for (int p = 0; p < kBwdPts; ++p) {          // all loads of the item before the first reduction
    const int i = base + p * kPairPoints + (threadIdx.x >> 1);
    go[p] = make_float2(0.f, 0.f);
#pragma unroll
    for (int d = 0; d < D; ++d) x[p][d] = 0.f;
    if (i < n) {
      go[p] = __ldg(d_enc_level + i);
#pragma unroll
      for (int d = 0; d < D; ++d) x[p][d] = __ldg(coords + (size_t)i * D + d);
    }
  }
#pragma unroll
  for (int p = 0; p < kBwdPts; ++p) {
    uint32_t cell[D];
    float frac[D];
#pragma unroll
    for (int d = 0; d < D; ++d) grid_pos(x[p][d], scale, cell[d], frac[d]);
    const bool live = !(go[p].x == 0.0f && go[p].y == 0.0f);      // adding +-0 is a no-op (also covers i >= n)
    PairTerms<D, MODE> pt;
    pt.init(cell, half, entries, res, 1u, swz);
    const float w0 = half ? frac[0] : 1.0f - frac[0];
    // The two lanes' hash indices differ by X = cell ^ (cell + 1) (hash prime 1 on dimension 0); the row layout
    // S is linear, so their rows differ by S(X).  S(X) == 1 (reference layout: even base cell): the rows are
    // idx and idx ^ 1, one aligned 16-byte slot -> ONE 128-bit reduction (RED.ADD.F32x4) issued by the even
    // lane instead of two 64-bit ones.  Both lanes of a pair hold the same point: the predicate is pair-uniform.
    const bool merge = (MODE == kIdxHash) && (grid_swizzle((cell[0] ^ (cell[0] + 1u)) & (entries - 1u), swz) == 1u);
#pragma unroll
    for (int c = 0; c < (1 << (D - 1)); ++c) {
      const float w = pair_weight<D>(frac, w0, c);
      const uint32_t idx = pt.index(c);
      const float vx = w * go[p].x, vy = w * go[p].y;
      const float ox = __shfl_xor_sync(0xffffffffu, vx, 1);
      const float oy = __shfl_xor_sync(0xffffffffu, vy, 1);
      if (!live) continue;
      if (merge) {
        if (half == 0) {
          const float4 v = (idx & 1u) ? make_float4(ox, oy, vx, vy) : make_float4(vx, vy, ox, oy);
          atomicAdd(reinterpret_cast<float4*>(gtab + (idx & ~1u)), v);
        }
      } else {
        atomicAdd(gtab + idx, make_float2(vx, vy));
      }
    }
  }
}

// hashed levels only (dense levels keep the run-aggregating kernel above)
template <int D>
__global__ void __launch_bounds__(kThreads)
hashgrid_bwd_pair_kernel(const __grid_constant__ immoco_grid_desc g, const float* __restrict__ coords,
                         const float2* __restrict__ d_enc, float2* __restrict__ grad_table, int n, int level0,
                         int n_levels, int tiles) {
  pdl_wait();
  for (int item = blockIdx.x; item < n_levels * tiles; item += gridDim.x) {
    const int level = level0 + item / tiles;
    const int base = (item % tiles) * (kPairPoints * kBwdPts);
    const float scale = g.scale[level];
    const uint32_t res = g.resolution[level];
    const uint32_t entries = g.entries[level];
    float2* __restrict__ gtab = grad_table + g.offset[level];
    const float2* __restrict__ go = d_enc + (size_t)level * n;
    if ((entries & (entries - 1u)) == 0u) bwd_pair_item<D, kIdxHash>(coords, go, gtab, n, base, scale, res, entries, g.swizzle[level]);
    else bwd_pair_item<D, kIdxAny>(coords, go, gtab, n, base, scale, res, entries, 0u);
  }
}

int check(const immoco_grid_desc* g, int64_t n) {
  if (!g || n < 0 || n > (int64_t)0x7fffffff / 2) return IMMOCO_ERR_BAD_ARG;
  if (g->n_levels < 1 || g->n_levels > IMMOCO_MAX_LEVELS) return IMMOCO_ERR_BAD_ARG;
  if (g->n_dims != 2 && g->n_dims != 3) return IMMOCO_ERR_UNSUPPORTED;
  return 0;
}

}  // namespace

static int g_pair = 1;   // 1: lane-pair kernels (product path); 0: one thread per point (A/B check)
static int g_ctas_per_sm = 0;   // > 0: persistent pair kernels with this many 256-thread CTAs per SM; 0: one CTA per item
// backward only (overrides g_ctas_per_sm when > 0): the scatter is paced by the L2 atomic units, not by
// occupancy, so it can run as a thin persistent grid beside an SM-bound kernel of the other branch
static int g_bwd_ctas_per_sm = 0;
extern "C" int immoco_set_hashgrid_bwd_ctas_per_sm(int32_t ctas) {
  if (ctas < 0 || ctas > 64) return IMMOCO_ERR_BAD_ARG;
  g_bwd_ctas_per_sm = ctas;
  return 0;
}
extern "C" int immoco_set_hashgrid_impl(int32_t pair) { g_pair = pair ? 1 : 0; return 0; }
extern "C" int immoco_set_hashgrid_ctas_per_sm(int32_t ctas) {
  if (ctas < 0 || ctas > 64) return IMMOCO_ERR_BAD_ARG;
  g_ctas_per_sm = ctas;
  return 0;
}

static int run_fwd(const immoco_grid_desc* grid, const float* coords, const float* table, float* enc,
                   int64_t n_points, int l0, int l1, void* stream) {
  if (int e = check(grid, n_points)) return e;
  if (l0 < 0 || l1 > grid->n_levels || l0 > l1) return IMMOCO_ERR_BAD_ARG;
  if (n_points == 0 || l0 == l1) return 0;
  const int n = (int)n_points;
  cudaStream_t s = (cudaStream_t)stream;
  if (g_pair) {
    const int tiles = (int)ceil_div64(n, kPairPoints * kFwdPts);
    const int64_t items = (int64_t)tiles * (l1 - l0);
    const int64_t cap = g_ctas_per_sm > 0 ? (int64_t)IMMOCO_NUM_SMS * g_ctas_per_sm : items;
    const unsigned g = (unsigned)(items < cap ? items : cap);
    if (grid->n_dims == 2)
      immoco_launch(hashgrid_fwd_pair_kernel<2>, dim3(g), dim3(kThreads), 0, s, *grid, coords, (const float2*)table, (float2*)enc, n, l0, l1 - l0, tiles);
    else
      immoco_launch(hashgrid_fwd_pair_kernel<3>, dim3(g), dim3(kThreads), 0, s, *grid, coords, (const float2*)table, (float2*)enc, n, l0, l1 - l0, tiles);
  } else {
    dim3 g((unsigned)ceil_div64(n, kThreads), (unsigned)(l1 - l0));
    if (grid->n_dims == 2)
      immoco_launch(hashgrid_fwd_kernel<2>, dim3(g), dim3(kThreads), 0, s, *grid, coords, (const float2*)table, (float2*)enc, n, l0);
    else
      immoco_launch(hashgrid_fwd_kernel<3>, dim3(g), dim3(kThreads), 0, s, *grid, coords, (const float2*)table, (float2*)enc, n, l0);
  }
  IMMOCO_LAUNCH_CHECK();
  return 0;
}

static int run_bwd(const immoco_grid_desc* grid, const float* coords, const float* d_enc, float* grad_table,
                   int64_t n_points, int l0, int l1, void* stream) {
  if (int e = check(grid, n_points)) return e;
  if (l0 < 0 || l1 > grid->n_levels || l0 > l1) return IMMOCO_ERR_BAD_ARG;
  if (n_points == 0 || l0 == l1) return 0;
  const int n = (int)n_points;
  cudaStream_t s = (cudaStream_t)stream;
  // consecutive levels of one kind (dense: run-aggregating kernel, hashed: lane-pair kernel) per launch
  for (int a = l0; a < l1;) {
    int b = a + 1;
    const bool pair = g_pair && grid->hashed[a];
    while (b < l1 && (g_pair && grid->hashed[b]) == pair) ++b;
    const int per_sm = g_bwd_ctas_per_sm > 0 ? g_bwd_ctas_per_sm : g_ctas_per_sm;
    dim3 g((unsigned)ceil_div64(n, kThreads), (unsigned)(b - a));
    if (per_sm > 0) {       // the dense-level kernel strides over the point tiles
      const unsigned gx = (unsigned)((IMMOCO_NUM_SMS * per_sm + (b - a) - 1) / (b - a));
      if (gx < g.x) g.x = gx;
    }
    if (pair) {
      const int tiles = (int)ceil_div64(n, kPairPoints * kBwdPts);
      const int64_t items = (int64_t)tiles * (b - a);
      const int64_t cap = per_sm > 0 ? (int64_t)IMMOCO_NUM_SMS * per_sm : items;
      const unsigned gp = (unsigned)(items < cap ? items : cap);
      if (grid->n_dims == 2)
        immoco_launch(hashgrid_bwd_pair_kernel<2>, dim3(gp), dim3(kThreads), 0, s, *grid, coords, (const float2*)d_enc, (float2*)grad_table, n, a, b - a, tiles);
      else
        immoco_launch(hashgrid_bwd_pair_kernel<3>, dim3(gp), dim3(kThreads), 0, s, *grid, coords, (const float2*)d_enc, (float2*)grad_table, n, a, b - a, tiles);
    } else {
      if (grid->n_dims == 2)
        immoco_launch(hashgrid_bwd_kernel<2>, dim3(g), dim3(kThreads), 0, s, *grid, coords, (const float2*)d_enc, (float2*)grad_table, n, a);
      else
        immoco_launch(hashgrid_bwd_kernel<3>, dim3(g), dim3(kThreads), 0, s, *grid, coords, (const float2*)d_enc, (float2*)grad_table, n, a);
    }
    IMMOCO_LAUNCH_CHECK();
    a = b;
  }
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

// the levels immoco_mlp_bwd_scatter leaves to the feature planes: every level that is not a power-of-two hashed
// level (dense levels: run-aggregating kernel)
extern "C" int immoco_hashgrid_bwd_dense_levels(const immoco_grid_desc* grid, const float* coords, const float* d_enc,
                                                float* grad_table, int64_t n_points, void* stream) {
  if (int e = check(grid, n_points)) return e;
  for (int a = 0; a < grid->n_levels;) {
    const uint32_t ent = grid->entries[a];
    const bool fused = grid->hashed[a] != 0u && (ent & (ent - 1u)) == 0u;
    if (fused) { ++a; continue; }
    int b = a + 1;
    while (b < grid->n_levels) {
      const uint32_t eb = grid->entries[b];
      if (grid->hashed[b] != 0u && (eb & (eb - 1u)) == 0u) break;
      ++b;
    }
    if (int e = run_bwd(grid, coords, d_enc, grad_table, n_points, a, b, stream)) return e;
    a = b;
  }
  return 0;
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
