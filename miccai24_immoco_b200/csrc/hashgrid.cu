// Multiresolution hash-grid encoding, forward gather and backward scatter.
// Replaces tiny-cuda-nn's kernel_grid / kernel_grid_backward used through
// tcnn.NetworkWithInputEncoding at src/models/immoco.py:60-65.
//
// Layout: table rows are float2 (n_features_per_level == 2); enc[level][point] float2 planes, so
// a warp handling 32 consecutive points of one level reads/writes 256 contiguous bytes.
// Grid = (point tiles, levels): CTAs of one level are adjacent in launch order, so one level's
// table (<= 4 MB) is the L2 working set at any time.
//
// Four families of kernels:
//   hashgrid_fwd_kernel / hashgrid_bwd_kernel         one thread per (point, level): checkers, dense-level scatter
//   hashgrid_*_pair_kernel                            two lanes = the two dim-0 corners of a point (Gray/exchange layout)
//   hashgrid_*_bundle_kernel                          2 M lanes = all groups x both dim-0 corners of one pixel of the grouped
//                                                     3-D grid, rows under a general linear layout (chunk tables)
//   hashgrid_*_taps_kernel (+ hashgrid_tap_rows_kernel)  2-D grid whose hashed levels are stored ranked by first touch
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

// Cache policy of the backward kernels (build-variant experiments, tools/l2_policy_ab.py):
//  IMMOCO_HG_BWD_STREAM 1: the cotangent planes are read once -> streaming (evict-first) loads;
//  IMMOCO_HG_RED_EVICT_LAST 1: the reductions carry an L2 evict-last policy, so the gradient table is still
//  L2-resident when Adam reads it.
#ifndef IMMOCO_HG_BWD_STREAM
#define IMMOCO_HG_BWD_STREAM 0
#endif
#ifndef IMMOCO_HG_RED_EVICT_LAST
#define IMMOCO_HG_RED_EVICT_LAST 0
#endif
__device__ __forceinline__ float2 load_cotangent(const float2* p) {
#if IMMOCO_HG_BWD_STREAM
  return __ldcs(p);
#else
  return __ldg(p);
#endif
}
__device__ __forceinline__ void red_add(float2* p, float2 v) {
#if IMMOCO_HG_RED_EVICT_LAST
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  asm volatile("red.global.add.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(p), "f"(v.x), "f"(v.y), "l"(pol) : "memory");
#else
  atomicAdd(p, v);
#endif
}
__device__ __forceinline__ void red_add(float4* p, float4 v) {
#if IMMOCO_HG_RED_EVICT_LAST
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  asm volatile("red.global.add.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w), "l"(pol)
               : "memory");
#else
  atomicAdd(p, v);
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
  const uint32_t* __restrict__ lut = grid_level_lut(g, level);       // general linear layout: tables read from global memory
  const uint32_t swz = lut ? 0u : g.swizzle[level];
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
      uint32_t row = grid_index<D>(q, hashed, entries, res, swz);
      if (lut) row = grid_lut_row(lut, row);
      const float2 v = __ldg(tab + row);
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
  const uint32_t* __restrict__ lut = grid_level_lut(g, level);
  const uint32_t swz = lut ? 0u : g.swizzle[level];
  float2* __restrict__ gtab = grad_table + g.offset[level];
  const unsigned lane = threadIdx.x & 31u;

  for (int base = blockIdx.x * kThreads; base < n; base += gridDim.x * kThreads) {
    const int i = base + threadIdx.x;
    const bool valid = i < n;
    float2 go = make_float2(0.f, 0.f);
    uint32_t cell[D];
    float frac[D];
    if (valid) {
      go = load_cotangent(d_enc + (size_t)level * n + i);
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
      uint32_t idx = grid_index<D>(q, hashed, entries, res, swz);
      if (lut) idx = grid_lut_row(lut, idx);
      float vx = w * go.x, vy = w * go.y;
      if (hashed) {
        // one 64-bit vector reduction per corner (RED.ADD.F32x2)
        if (live) red_add(gtab + idx, make_float2(vx, vy));
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
        if (head && live) red_add(gtab + idx, make_float2(vx, vy));
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
      go[p] = load_cotangent(d_enc_level + i);
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
          red_add(reinterpret_cast<float4*>(gtab + (idx & ~1u)), v);
        }
      } else {
        red_add(gtab + idx, make_float2(vx, vy));
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

// ---- compact ("tap-indexed") storage of a grid's hashed levels ----------------------------------------------
// The coordinates of a fit are constant, so the table rows a point touches can be looked up once
// (hashgrid_tap_rows_kernel) and the hashed levels re-stored in ANY row order: the host ranks the rows by their
// first touch in (level, point, corner) order (immoco.py:GridTaps).  What that buys on the 2-D image grid, whose
// hashed levels are under-subscribed (102400 pixels x 4 corners into 2^19 rows: 53 % of the rows are ever touched):
//  * rows nobody touches keep g = m = v = 0 for ever -- they are stored behind the touched ones and Adam / the
//    gradient memset never visit them (C2: 5.59 M -> 3.11 M rows);
//  * the rows a pixel touches first are consecutive, in pixel order: ~70 % of the gathers / reductions of a warp
//    fall into a few adjacent lines instead of one line each.
// The kernels read 16 bytes of row indices per (point, level) instead of hashing.
template <int D>
__global__ void __launch_bounds__(kThreads)
hashgrid_tap_rows_kernel(const __grid_constant__ immoco_grid_desc g, const float* __restrict__ coords,
                         uint32_t* __restrict__ rows, int n, int level0) {
  const int level = level0 + blockIdx.y;
  const float scale = g.scale[level];
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
    uint32_t cell[D];
    float frac[D];
#pragma unroll
    for (int d = 0; d < D; ++d) grid_pos(__ldg(coords + (size_t)i * D + d), scale, cell[d], frac[d]);
#pragma unroll
    for (int c = 0; c < (1 << D); ++c) {
      uint32_t q[D];
#pragma unroll
      for (int d = 0; d < D; ++d) q[d] = cell[d] + (uint32_t)((c >> d) & 1);
      const uint32_t* lut = grid_level_lut(g, level);
      const uint32_t row = grid_index<D>(q, g.hashed[level], g.entries[level], g.resolution[level], lut ? 0u : g.swizzle[level]);
      rows[((size_t)blockIdx.y * n + i) * (1 << D) + c] = lut ? grid_lut_row(lut, row) : row;
    }
  }
}

// points per thread and item of the tap-indexed kernels (all index and row loads of an item are issued before
// the first use)
#ifndef IMMOCO_HG_TAP_PTS
#define IMMOCO_HG_TAP_PTS 2
#endif
constexpr int kTapPts = IMMOCO_HG_TAP_PTS;

// 2-D, one thread per point.  Same products and the same summation order as the lane-pair kernel (each dim-0
// half: fma chain over the dim-1 corners; then half 0 + half 1), so the features are bit-identical to it.
__device__ __forceinline__ void fwd_tap_item(const float2* __restrict__ coords, const float2* __restrict__ table,
                                             const uint4* __restrict__ taps_level, float2* __restrict__ enc_level,
                                             int n, int base, float scale) {
  float2 v[kTapPts][4];
  float2 x[kTapPts];
#pragma unroll
  for (int p = 0; p < kTapPts; ++p) {
    const int i = base + p * kThreads + threadIdx.x;
    x[p] = make_float2(0.f, 0.f);
    uint4 t = make_uint4(0u, 0u, 0u, 0u);
    if (i < n) { x[p] = __ldg(coords + i); t = __ldg(taps_level + i); }
    if (i < n) {
      v[p][0] = load_row(table + t.x); v[p][1] = load_row(table + t.y);
      v[p][2] = load_row(table + t.z); v[p][3] = load_row(table + t.w);
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c) v[p][c] = make_float2(0.f, 0.f);
    }
  }
#pragma unroll
  for (int p = 0; p < kTapPts; ++p) {
    const int i = base + p * kThreads + threadIdx.x;
    uint32_t cell;
    float f0, f1;
    grid_pos(x[p].x, scale, cell, f0);
    grid_pos(x[p].y, scale, cell, f1);
    float2 half[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float w0 = h ? f0 : 1.0f - f0;
      const float wa = w0 * (1.0f - f1), wb = w0 * f1;          // pair_weight: dim-1 corner 0, 1
      half[h].x = fmaf(wb, v[p][2 + h].x, fmaf(wa, v[p][h].x, 0.f));
      half[h].y = fmaf(wb, v[p][2 + h].y, fmaf(wa, v[p][h].y, 0.f));
    }
    if (i < n) enc_level[i] = make_float2(half[0].x + half[1].x, half[0].y + half[1].y);
  }
}

// Forward of a grid with tap-indexed levels in ONE launch: items of the levels below `first` run the lane-pair
// code, the others the indexed code; level-major order as in hashgrid_fwd_pair_kernel.
__global__ void __launch_bounds__(kThreads, IMMOCO_HG_FWD_MIN_CTAS)
hashgrid_fwd_taps_kernel(const __grid_constant__ immoco_grid_desc g, const float* __restrict__ coords,
                         const float2* __restrict__ table, const uint4* __restrict__ taps, float2* __restrict__ enc,
                         int n, int first, int tiles_pair, int tiles_tap) {
  pdl_wait();
  const int items_pair = first * tiles_pair;
  const int items = items_pair + (g.n_levels - first) * tiles_tap;
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    if (item < items_pair) {
      const int level = item / tiles_pair;
      const int base = (item % tiles_pair) * (kPairPoints * kFwdPts);
      const uint32_t entries = g.entries[level], hashed = g.hashed[level], res = g.resolution[level];
      const float2* __restrict__ tab = table + g.offset[level];
      float2* __restrict__ out = enc + (size_t)level * n;
      const bool pow2 = (entries & (entries - 1u)) == 0u;
      if (pow2 && hashed) fwd_pair_item<2, kIdxHash>(coords, tab, out, n, base, g.scale[level], res, entries, hashed, g.swizzle[level]);
      else if (pow2) fwd_pair_item<2, kIdxDense>(coords, tab, out, n, base, g.scale[level], res, entries, hashed, 0u);
      else fwd_pair_item<2, kIdxAny>(coords, tab, out, n, base, g.scale[level], res, entries, hashed, 0u);
    } else {
      const int k = (item - items_pair) / tiles_tap;
      const int level = first + k;
      const int base = ((item - items_pair) % tiles_tap) * (kThreads * kTapPts);
      fwd_tap_item(reinterpret_cast<const float2*>(coords), table, taps + (size_t)k * n, enc + (size_t)level * n, n,
                   base, g.scale[level]);
    }
  }
}

// Backward of the tap-indexed levels: one thread per point, four reductions; two rows that ended up in one
// aligned 16-byte slot (consecutive first-touch ranks) go out as ONE RED.ADD.F32x4.
__global__ void __launch_bounds__(kThreads)
hashgrid_bwd_taps_kernel(const __grid_constant__ immoco_grid_desc g, const float* __restrict__ coords,
                         const float2* __restrict__ d_enc, const uint4* __restrict__ taps,
                         float2* __restrict__ grad_table, int n, int first, int tiles) {
  pdl_wait();
  const int items = (g.n_levels - first) * tiles;
  const float2* __restrict__ xy = reinterpret_cast<const float2*>(coords);
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int k = item / tiles;
    const int level = first + k;
    const int i = (item % tiles) * kThreads + threadIdx.x;
    if (i >= n) continue;
    const float2 go = load_cotangent(d_enc + (size_t)level * n + i);
    const float2 x = __ldg(xy + i);
    const uint4 t = __ldg(taps + (size_t)k * n + i);
    if (go.x == 0.0f && go.y == 0.0f) continue;                  // adding +-0 is a no-op
    uint32_t cell;
    float f0, f1;
    grid_pos(x.x, g.scale[level], cell, f0);
    grid_pos(x.y, g.scale[level], cell, f1);
    const float wl = 1.0f - f0, wa = 1.0f - f1;
    const float w[4] = {wl * wa, f0 * wa, wl * f1, f0 * f1};     // corner c: bit 0 = dim 0, bit 1 = dim 1
    const uint32_t r[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int c = 0; c < 4; c += 2) {
      const float ax = w[c] * go.x, ay = w[c] * go.y, bx = w[c + 1] * go.x, by = w[c + 1] * go.y;
      if (r[c + 1] == r[c] + 1u && (r[c] & 1u) == 0u) {
        red_add(reinterpret_cast<float4*>(grad_table + r[c]), make_float4(ax, ay, bx, by));
      } else {
        red_add(grad_table + r[c], make_float2(ax, ay));
        red_add(grad_table + r[c + 1], make_float2(bx, by));
      }
    }
  }
}

// ---- grouped ("bundle") kernels of the 3-D grid -----------------------------------------------------------------
// Coordinates (t_g, y_p, x_p): the 2 M lanes of a bundle take the M groups x 2 dim-0 corners of ONE pixel.  The hash
// is XOR-linear and so is the row layout S (Gray/exchange word or chunk tables), hence
//     row(corner) = S(q0) ^ S((c1 + b1) * P1) ^ S((c2 + b2) * P2):
// S(q0) is a per-lane constant of the item (the group's cell), the four pixel terms are evaluated once per BUNDLE
// (lanes 0..3, three shared-memory look-ups each) and broadcast.  With encoding.py:linear_layout the 2 M rows of a
// pixel corner lie in one or two 128-byte lines, so a gather / reduction instruction of a warp touches 8 lines
// (M = 4) instead of the lane-pair kernels' 16: the L1TEX tag stage (one line per cycle and SM) is what paces the
// gathers.  Same products and summation order as the lane-pair kernels: bit-identical features.
// pixel passes per item: an item pays the table staging (a global load + two barriers) and the per-lane group
// constants once.  B200, 3-D grid at C2 (tools/grouped_ab.py, gpurun_out/r305 / r306): forward 2 -> 108 us,
// 4 -> 90 us, 8 -> 84 us; resident CTAs per SM the forward kernel is compiled for: 4 (64 registers) -> 98 us,
// 6 (40) -> 90 us, 8 (32, spills) -> 106 us
#ifndef IMMOCO_HG_BUNDLE_ITERS
#define IMMOCO_HG_BUNDLE_ITERS 8
#endif
#ifndef IMMOCO_HG_BUNDLE_ITERS_BWD
#define IMMOCO_HG_BUNDLE_ITERS_BWD 4
#endif
constexpr int kBundleIters = IMMOCO_HG_BUNDLE_ITERS;
constexpr int kBundleItersBwd = IMMOCO_HG_BUNDLE_ITERS_BWD;
#ifndef IMMOCO_HG_BUNDLE_MIN_CTAS
#define IMMOCO_HG_BUNDLE_MIN_CTAS 6
#endif

// lanes of a bundle: the next power of two >= 2 M, at least 4 (the four pixel terms are evaluated by lanes 0..3)
__host__ __device__ __forceinline__ int bundle_lanes(int m) {
  int l = 4;
  while (l < 2 * m) l <<= 1;
  return l;
}

struct BundleLevel {
  uint32_t mask, swz, res, entries, hashed;
  bool lin;           // hashed power-of-two level: XOR-linear index + linear layout
  bool lut;
};

__device__ __forceinline__ uint32_t bundle_layout(const BundleLevel& lv, const uint32_t* lut_s, uint32_t x) {
  return lv.lut ? grid_lut_row(lut_s, x) : grid_swizzle(x, lv.swz);
}

// rows of the four (dim-1, dim-2) corners of this lane's (group, dim-0 corner) for one pixel
__device__ __forceinline__ void bundle_rows(const BundleLevel& lv, const uint32_t* lut_s, uint32_t q0, uint32_t s0,
                                            uint32_t c1, uint32_t c2, int lb, unsigned lead, uint32_t (&row)[4]) {
  if (lv.lin) {
    // lane lb & 3 of the bundle evaluates one of the four pixel terms
    const uint32_t term = (lb & 2) ? (c2 + (uint32_t)(lb & 1)) * 805459861u : (c1 + (uint32_t)(lb & 1)) * 2654435761u;
    const uint32_t s = lb < 4 ? bundle_layout(lv, lut_s, term & lv.mask) : 0u;     // fewer active lanes: fewer bank conflicts
    const uint32_t a0 = __shfl_sync(0xffffffffu, s, lead), a1 = __shfl_sync(0xffffffffu, s, lead + 1);
    const uint32_t b0 = __shfl_sync(0xffffffffu, s, lead + 2), b1 = __shfl_sync(0xffffffffu, s, lead + 3);
    row[0] = s0 ^ a0 ^ b0; row[1] = s0 ^ a1 ^ b0; row[2] = s0 ^ a0 ^ b1; row[3] = s0 ^ a1 ^ b1;
  } else {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const uint32_t q[3] = {q0, c1 + (uint32_t)(c & 1), c2 + (uint32_t)(c >> 1)};
      row[c] = grid_index<3>(q, lv.hashed, lv.entries, lv.res, 0u);
    }
  }
}

__device__ __forceinline__ BundleLevel bundle_level(const immoco_grid_desc& g, int level) {
  BundleLevel lv;
  lv.entries = g.entries[level]; lv.mask = lv.entries - 1u; lv.res = g.resolution[level]; lv.hashed = g.hashed[level];
  lv.lin = lv.hashed != 0u && (lv.entries & lv.mask) == 0u;
  lv.lut = lv.lin && g.swizzle[level] == IMMOCO_LAYOUT_LUT && g.layout_lut != nullptr;
  lv.swz = (lv.lin && !lv.lut) ? g.swizzle[level] : 0u;
  return lv;
}

__global__ void __launch_bounds__(kThreads, IMMOCO_HG_BUNDLE_MIN_CTAS)
hashgrid_fwd_bundle_kernel(const __grid_constant__ immoco_grid_desc g, const float* __restrict__ coords,
                           const float2* __restrict__ table, float2* __restrict__ enc, int P, int M, int level0,
                           int n_levels, int tiles) {
  __shared__ uint32_t lut_s[256];
  pdl_wait();
  const int lanes = bundle_lanes(M), bundles = kThreads / lanes;
  const int lb = threadIdx.x % lanes, bundle = threadIdx.x / lanes;
  const bool active = lb < 2 * M;               // group counts that are no power of two leave the bundle's last lanes idle
  const int grp = active ? lb >> 1 : 0, half = lb & 1;
  const unsigned lead = (threadIdx.x & 31u) & ~(unsigned)(lanes - 1);
  const size_t n = (size_t)P * M;
  const float t_g = __ldg(coords + (size_t)grp * P * 3);
  for (int item = blockIdx.x; item < n_levels * tiles; item += gridDim.x) {
    const int level = level0 + item / tiles;
    const int pix0 = (item % tiles) * (bundles * kBundleIters);
    const BundleLevel lv = bundle_level(g, level);
    const float scale = g.scale[level];
    __syncthreads();
    if (lv.lut) lut_s[threadIdx.x] = __ldg(g.layout_lut + 256 * level + threadIdx.x);
    __syncthreads();
    uint32_t c0;
    float f0;
    grid_pos(t_g, scale, c0, f0);
    const uint32_t q0 = c0 + (uint32_t)half;
    const uint32_t s0 = lv.lin ? bundle_layout(lv, lut_s, q0 & lv.mask) : 0u;
    const float w0 = half ? f0 : 1.0f - f0;
    const float2* __restrict__ tab = table + g.offset[level];
    float2* __restrict__ out = enc + (size_t)level * n + (size_t)grp * P;
#pragma unroll
    for (int it = 0; it < kBundleIters; it += 2) {
      float2 v[2][4];
      float fr[2][3];
      int pix[2];
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        pix[p] = pix0 + (it + p) * bundles + bundle;
        const bool ok = pix[p] < P;
        const float y = ok ? __ldg(coords + (size_t)pix[p] * 3 + 1) : 0.f;
        const float x = ok ? __ldg(coords + (size_t)pix[p] * 3 + 2) : 0.f;
        uint32_t c1, c2, row[4];
        grid_pos(y, scale, c1, fr[p][1]);
        grid_pos(x, scale, c2, fr[p][2]);
        fr[p][0] = f0;
        bundle_rows(lv, lut_s, q0, s0, c1, c2, lb, lead, row);
#pragma unroll
        for (int c = 0; c < 4; ++c) v[p][c] = (ok && active) ? load_row(tab + row[c]) : make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float w = pair_weight<3>(fr[p], w0, c);
          acc.x = fmaf(w, v[p][c].x, acc.x);
          acc.y = fmaf(w, v[p][c].y, acc.y);
        }
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 1);
        acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 1);
        if (pix[p] < P && half == 0 && active) out[pix[p]] = acc;
      }
    }
  }
}

// hashed power-of-two levels only (the dense levels keep the run-aggregating kernel)
__global__ void __launch_bounds__(kThreads)
hashgrid_bwd_bundle_kernel(const __grid_constant__ immoco_grid_desc g, const float* __restrict__ coords,
                           const float2* __restrict__ d_enc, float2* __restrict__ grad_table, int P, int M,
                           int level0, int n_levels, int tiles) {
  __shared__ uint32_t lut_s[256];
  pdl_wait();
  const int lanes = bundle_lanes(M), bundles = kThreads / lanes;
  const int lb = threadIdx.x % lanes, bundle = threadIdx.x / lanes;
  const bool active = lb < 2 * M;               // group counts that are no power of two leave the bundle's last lanes idle
  const int grp = active ? lb >> 1 : 0, half = lb & 1;
  const unsigned lead = (threadIdx.x & 31u) & ~(unsigned)(lanes - 1);
  const size_t n = (size_t)P * M;
  const float t_g = __ldg(coords + (size_t)grp * P * 3);
  for (int item = blockIdx.x; item < n_levels * tiles; item += gridDim.x) {
    const int level = level0 + item / tiles;
    const int pix0 = (item % tiles) * (bundles * kBundleItersBwd);
    const BundleLevel lv = bundle_level(g, level);
    const float scale = g.scale[level];
    __syncthreads();
    if (lv.lut) lut_s[threadIdx.x] = __ldg(g.layout_lut + 256 * level + threadIdx.x);
    __syncthreads();
    uint32_t c0;
    float f0;
    grid_pos(t_g, scale, c0, f0);
    const uint32_t q0 = c0 + (uint32_t)half;
    const uint32_t s0 = bundle_layout(lv, lut_s, q0 & lv.mask);
    const float w0 = half ? f0 : 1.0f - f0;
    // the pair's rows differ by S(cell ^ (cell + 1)); == 1: one aligned 16-byte slot -> ONE RED.ADD.F32x4 by the even lane
    const bool merge = (s0 ^ __shfl_xor_sync(0xffffffffu, s0, 1)) == 1u;
    float2* __restrict__ gtab = grad_table + g.offset[level];
    const float2* __restrict__ go_l = d_enc + (size_t)level * n + (size_t)grp * P;
#pragma unroll 2
    for (int it = 0; it < kBundleItersBwd; ++it) {
      const int pix = pix0 + it * bundles + bundle;
      const bool ok = pix < P && active;
      float2 go = make_float2(0.f, 0.f);
      float y = 0.f, x = 0.f;
      if (ok) {
        go = load_cotangent(go_l + pix);
        y = __ldg(coords + (size_t)pix * 3 + 1);
        x = __ldg(coords + (size_t)pix * 3 + 2);
      }
      float fr[3];
      uint32_t c1, c2, row[4];
      grid_pos(y, scale, c1, fr[1]);
      grid_pos(x, scale, c2, fr[2]);
      fr[0] = f0;
      bundle_rows(lv, lut_s, q0, s0, c1, c2, lb, lead, row);
      const bool live = !(go.x == 0.0f && go.y == 0.0f);       // adding +-0 is a no-op (also covers pix >= P)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float w = pair_weight<3>(fr, w0, c);
        const float vx = w * go.x, vy = w * go.y;
        const float ox = __shfl_xor_sync(0xffffffffu, vx, 1);
        const float oy = __shfl_xor_sync(0xffffffffu, vy, 1);
        if (!live) continue;
        if (merge) {
          if (half == 0) {
            const float4 v = (row[c] & 1u) ? make_float4(ox, oy, vx, vy) : make_float4(vx, vy, ox, oy);
            red_add(reinterpret_cast<float4*>(gtab + (row[c] & ~1u)), v);
          }
        } else {
          red_add(gtab + row[c], make_float2(vx, vy));
        }
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
  if (g_pair && !grid_has_lut(*grid)) {
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
    const bool use_pair = g_pair && !grid_has_lut(*grid);
    const bool pair = use_pair && grid->hashed[a];
    while (b < l1 && (use_pair && grid->hashed[b]) == pair) ++b;
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

// ---- tap-indexed levels (see hashgrid_tap_rows_kernel) --------------------------------------------------------
static int check_taps(const immoco_grid_desc* grid, const immoco_grid_taps* taps, int64_t n) {
  if (int e = check(grid, n)) return e;
  if (!taps || !taps->rows || taps->first_level < 0 || taps->first_level > grid->n_levels || taps->n_points != n)
    return IMMOCO_ERR_BAD_ARG;
  if (grid->n_dims != 2) return IMMOCO_ERR_UNSUPPORTED;
  if (((uintptr_t)taps->rows & 15) != 0) return IMMOCO_ERR_BAD_ARG;
  return 0;
}

extern "C" int immoco_hashgrid_tap_rows(const immoco_grid_desc* grid, const float* coords, int64_t n_points,
                                        int32_t level_begin, int32_t level_end, uint32_t* rows, void* stream) {
  if (int e = check(grid, n_points)) return e;
  if (!coords || !rows || level_begin < 0 || level_end > grid->n_levels || level_begin > level_end) return IMMOCO_ERR_BAD_ARG;
  if (n_points == 0 || level_begin == level_end) return 0;
  const int n = (int)n_points;
  dim3 g((unsigned)ceil_div64(n, kThreads), (unsigned)(level_end - level_begin));
  if (grid->n_dims == 2)
    hashgrid_tap_rows_kernel<2><<<g, kThreads, 0, (cudaStream_t)stream>>>(*grid, coords, rows, n, level_begin);
  else
    hashgrid_tap_rows_kernel<3><<<g, kThreads, 0, (cudaStream_t)stream>>>(*grid, coords, rows, n, level_begin);
  IMMOCO_LAUNCH_CHECK();
  return 0;
}

extern "C" int immoco_hashgrid_fwd_taps(const immoco_grid_desc* grid, const immoco_grid_taps* taps, const float* coords,
                                        const float* table, float* enc, int64_t n_points, void* stream) {
  if (int e = check_taps(grid, taps, n_points)) return e;
  if (n_points == 0) return 0;
  const int n = (int)n_points;
  const int first = taps->first_level;
  const int tiles_pair = (int)ceil_div64(n, kPairPoints * kFwdPts);
  const int tiles_tap = (int)ceil_div64(n, kThreads * kTapPts);
  const int64_t items = (int64_t)first * tiles_pair + (int64_t)(grid->n_levels - first) * tiles_tap;
  const int64_t cap = g_ctas_per_sm > 0 ? (int64_t)IMMOCO_NUM_SMS * g_ctas_per_sm : items;
  const unsigned g = (unsigned)(items < cap ? items : cap);
  immoco_launch(hashgrid_fwd_taps_kernel, dim3(g), dim3(kThreads), 0, (cudaStream_t)stream, *grid, coords,
                (const float2*)table, (const uint4*)taps->rows, (float2*)enc, n, first, tiles_pair, tiles_tap);
  IMMOCO_LAUNCH_CHECK();
  return 0;
}

extern "C" int immoco_hashgrid_bwd_taps(const immoco_grid_desc* grid, const immoco_grid_taps* taps, const float* coords,
                                        const float* d_enc, float* grad_table, int64_t n_points, void* stream) {
  if (int e = check_taps(grid, taps, n_points)) return e;
  if (n_points == 0) return 0;
  const int first = taps->first_level;
  if (int e = run_bwd(grid, coords, d_enc, grad_table, n_points, 0, first, stream)) return e;
  if (first == grid->n_levels) return 0;
  const int n = (int)n_points;
  const int tiles = (int)ceil_div64(n, kThreads);
  const int64_t items = (int64_t)tiles * (grid->n_levels - first);
  const int per_sm = g_bwd_ctas_per_sm > 0 ? g_bwd_ctas_per_sm : g_ctas_per_sm;
  const int64_t cap = per_sm > 0 ? (int64_t)IMMOCO_NUM_SMS * per_sm : items;
  const unsigned g = (unsigned)(items < cap ? items : cap);
  immoco_launch(hashgrid_bwd_taps_kernel, dim3(g), dim3(kThreads), 0, (cudaStream_t)stream, *grid, coords,
                (const float2*)d_enc, (const uint4*)taps->rows, (float2*)grad_table, n, first, tiles);
  IMMOCO_LAUNCH_CHECK();
  return 0;
}

// ---- grouped 3-D entry points (see hashgrid_fwd_bundle_kernel) ------------------------------------------------
static int check_grouped(const immoco_grid_desc* grid, int64_t n_pixels, int32_t n_groups) {
  if (!grid || n_pixels < 0 || n_groups < 1) return IMMOCO_ERR_BAD_ARG;
  if (int e = check(grid, n_pixels * n_groups)) return e;
  if (grid->n_dims != 3 || kThreads != 256) return IMMOCO_ERR_UNSUPPORTED;
  if (n_groups < 2 || n_groups > 16) return IMMOCO_ERR_UNSUPPORTED;
  if (grid_has_lut(*grid) && !grid->layout_lut) return IMMOCO_ERR_BAD_ARG;
  return 0;
}

extern "C" int immoco_hashgrid_fwd_grouped(const immoco_grid_desc* grid, const float* coords, const float* table,
                                           float* enc, int64_t n_pixels, int32_t n_groups, void* stream) {
  if (int e = check_grouped(grid, n_pixels, n_groups)) return e;
  if (n_pixels == 0) return 0;
  const int per_item = (kThreads / bundle_lanes(n_groups)) * kBundleIters;
  const int tiles = (int)ceil_div64(n_pixels, per_item);
  const int64_t items = (int64_t)tiles * grid->n_levels;
  const int64_t cap = g_ctas_per_sm > 0 ? (int64_t)IMMOCO_NUM_SMS * g_ctas_per_sm : items;
  const unsigned g = (unsigned)(items < cap ? items : cap);
  immoco_launch(hashgrid_fwd_bundle_kernel, dim3(g), dim3(kThreads), 0, (cudaStream_t)stream, *grid, coords,
                (const float2*)table, (float2*)enc, (int)n_pixels, (int)n_groups, 0, grid->n_levels, tiles);
  IMMOCO_LAUNCH_CHECK();
  return 0;
}

extern "C" int immoco_hashgrid_bwd_grouped(const immoco_grid_desc* grid, const float* coords, const float* d_enc,
                                           float* grad_table, int64_t n_pixels, int32_t n_groups, void* stream) {
  if (int e = check_grouped(grid, n_pixels, n_groups)) return e;
  if (n_pixels == 0) return 0;
  const int64_t n = n_pixels * n_groups;
  cudaStream_t s = (cudaStream_t)stream;
  const int per_item = (kThreads / bundle_lanes(n_groups)) * kBundleItersBwd;
  const int tiles = (int)ceil_div64(n_pixels, per_item);
  const int per_sm = g_bwd_ctas_per_sm > 0 ? g_bwd_ctas_per_sm : g_ctas_per_sm;
  // consecutive levels of one kind per launch: hashed power-of-two levels -> bundle kernel, the others -> the
  // run-aggregating kernel (one thread per point)
  for (int a = 0; a < grid->n_levels;) {
    auto linear = [&](int l) { return grid->hashed[l] != 0u && (grid->entries[l] & (grid->entries[l] - 1u)) == 0u; };
    int b = a + 1;
    const bool lin = linear(a);
    while (b < grid->n_levels && linear(b) == lin) ++b;
    if (lin) {
      const int64_t items = (int64_t)tiles * (b - a);
      const int64_t cap = per_sm > 0 ? (int64_t)IMMOCO_NUM_SMS * per_sm : items;
      const unsigned gp = (unsigned)(items < cap ? items : cap);
      immoco_launch(hashgrid_bwd_bundle_kernel, dim3(gp), dim3(kThreads), 0, s, *grid, coords, (const float2*)d_enc,
                    (float2*)grad_table, (int)n_pixels, (int)n_groups, a, b - a, tiles);
    } else {
      dim3 gd((unsigned)ceil_div64(n, kThreads), (unsigned)(b - a));
      if (per_sm > 0) {
        const unsigned gx = (unsigned)((IMMOCO_NUM_SMS * per_sm + (b - a) - 1) / (b - a));
        if (gx < gd.x) gd.x = gx;
      }
      immoco_launch(hashgrid_bwd_kernel<3>, gd, dim3(kThreads), 0, s, *grid, coords, (const float2*)d_enc,
                    (float2*)grad_table, (int)n, a);
    }
    IMMOCO_LAUNCH_CHECK();
    a = b;
  }
  return 0;
}
