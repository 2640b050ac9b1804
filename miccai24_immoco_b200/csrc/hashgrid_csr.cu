// Deterministic hash-grid backward: row-sorted tap list (CSR) built once per coordinate set, then a
// gather with a fixed summation order per table row -- optionally with the Adam update of that row fused in.
// Replaces tiny-cuda-nn's kernel_grid_backward (atomic scatter) + the table part of torch.optim.Adam.step
// on the IM-MoCo fit path (src/models/immoco.py:149-154,164-175): the fit's coordinates are registered
// buffers (immoco.py:72-80), i.e. WHICH point touches WHICH row with WHICH weight never changes.
//
//   build : per level  keys = (row << 32 | point * 2^D + corner)  -> radix sort (CUB, set-up only)
//           -> taps[k] = {point, weight}, row_ptr[row] = lower bound of the row in the sorted keys
//   gather: g[row] = sum_k w_k * d_enc[plane_k] with a FIXED summation tree per row, no atomics; rows without taps
//           are skipped (their gradient / moments stay zero for ever).  Coarse levels (hundreds of taps per row):
//           a warp per row, lane-strided.  All other levels: a warp owns 32 consecutive rows and walks THEIR taps
//           32 at a time -- lane = tap (coalesced record loads, every lane busy whatever the rows' tap counts: a
//           thread per row idled half the warp, the counts are Poisson, mean 6, warp maximum ~13), a segmented
//           scan over the lanes sums the taps of each row inside the step, and the row's owner lane (lane = row)
//           picks its partial sum up by shuffle and accumulates it step after step.
//   A tap record is 8 bytes: {plane index (level * n + point) | (row & 31) << 27, weight}.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr uint32_t kRidShift = 27;                   // tap.x = plane index (27 bits) | row-in-warp (5 bits)
constexpr uint32_t kPlaneMask = (1u << kRidShift) - 1u;

template <int D>
__global__ void __launch_bounds__(kThreads)
csr_emit_keys_kernel(const __grid_constant__ immoco_grid_desc g, const float* __restrict__ coords, int n, int level,
                     unsigned long long* __restrict__ keys) {
  const float scale = g.scale[level];
  const uint32_t res = g.resolution[level], entries = g.entries[level], hashed = g.hashed[level];
  const uint32_t swz = g.swizzle[level];
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
      const uint32_t row = grid_index<D>(q, hashed, entries, res, swz);
      const uint32_t tap = (uint32_t)i * (1u << D) + (uint32_t)c;
      keys[(size_t)tap] = ((unsigned long long)row << 32) | tap;
    }
  }
}

// sorted keys of one level -> {point, weight} records (weight = (w0 * w1) * w2, the forward kernels' order)
template <int D>
__global__ void __launch_bounds__(kThreads)
csr_fill_taps_kernel(const __grid_constant__ immoco_grid_desc g, const float* __restrict__ coords, int n, int level,
                     const unsigned long long* __restrict__ keys, uint32_t n_taps, uint2* __restrict__ taps) {
  const float scale = g.scale[level];
  for (uint32_t k = blockIdx.x * kThreads + threadIdx.x; k < n_taps; k += gridDim.x * kThreads) {
    const uint32_t tap = (uint32_t)(keys[k] & 0xffffffffull);
    const uint32_t row = (uint32_t)(keys[k] >> 32);
    const uint32_t point = tap >> D, c = tap & ((1u << D) - 1u);
    float w = 1.0f;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      uint32_t cell;
      float frac;
      grid_pos(__ldg(coords + (size_t)point * D + d), scale, cell, frac);
      const float wd = ((c >> d) & 1u) ? frac : 1.0f - frac;
      w = (d == 0) ? wd : w * wd;
    }
    taps[k] = make_uint2(((uint32_t)level * (uint32_t)n + point) | ((row & 31u) << kRidShift), __float_as_uint(w));
  }
}

// row_ptr[r] = base + (number of sorted keys whose row is < r), r = 0 .. entries (inclusive when `last`)
__global__ void __launch_bounds__(kThreads)
csr_row_ptr_kernel(const unsigned long long* __restrict__ keys, uint32_t n_taps, uint32_t entries, uint32_t base,
                   int last, uint32_t* __restrict__ row_ptr) {
  const uint32_t count = entries + (last ? 1u : 0u);
  for (uint32_t r = blockIdx.x * kThreads + threadIdx.x; r < count; r += gridDim.x * kThreads) {
    const unsigned long long want = (unsigned long long)r << 32;
    uint32_t lo = 0, hi = n_taps;           // first k with keys[k] >= want
    while (lo < hi) {
      const uint32_t mid = lo + ((hi - lo) >> 1);
      if (keys[mid] < want) lo = mid + 1;
      else hi = mid;
    }
    row_ptr[r] = base + lo;
  }
}

struct GatherPlan {
  uint32_t cta_begin[IMMOCO_MAX_LEVELS + 1];
  uint32_t wide[IMMOCO_MAX_LEVELS];       // 1: one warp per row (coarse levels), 0: one thread per row
};

struct AdamConsts {
  float omb1, b2, omb2, step_size, bc2_sqrt, eps;
};

// parameter / moment loads of a row are issued BEFORE its gather loop (ADAM) so that they overlap it
struct RowState {
  float2 p, m, v;
};
template <bool ADAM>
__device__ __forceinline__ RowState load_row_state(uint32_t row, const float2* __restrict__ table,
                                                   const float2* __restrict__ m, const float2* __restrict__ v) {
  RowState st;
  if (ADAM) { st.p = table[row]; st.m = m[row]; st.v = v[row]; }
  return st;
}

template <bool ADAM>
__device__ __forceinline__ void finish_row(uint32_t row, float gx, float gy, const RowState& st,
                                           float2* __restrict__ grad, float2* __restrict__ table,
                                           float2* __restrict__ m, float2* __restrict__ v, const AdamConsts& a) {
  if (ADAM) {
    float2 p = st.p, mm = st.m, vv = st.v;
    adam_update(p.x, mm.x, vv.x, gx, a.omb1, a.b2, a.omb2, a.step_size, a.bc2_sqrt, a.eps);
    adam_update(p.y, mm.y, vv.y, gy, a.omb1, a.b2, a.omb2, a.step_size, a.bc2_sqrt, a.eps);
    table[row] = p;
    m[row] = mm;
    v[row] = vv;
    if (grad) grad[row] = make_float2(gx, gy);
  } else {
    grad[row] = make_float2(gx, gy);
  }
}

// The tap list is read once per pass and is 4 x larger than everything the gathers hit (420 MB against the 52 MB
// of cotangent planes and the tables).  Streaming it past L2 (evict-first) so that it does not push the planes
// out was tried and is slower; the default policy stays.
#ifndef IMMOCO_CSR_STREAM_TAPS
#define IMMOCO_CSR_STREAM_TAPS 0      // measured: streaming the taps is SLOWER (serial 329 vs 301 us, iteration 1079 vs ~1010 us)
#endif
__device__ __forceinline__ uint2 ld_tap(const uint2* p) {
#if IMMOCO_CSR_STREAM_TAPS
  return __ldcs(p);
#else
  return __ldg(p);
#endif
}

template <bool ADAM>
__global__ void __launch_bounds__(kThreads)
hashgrid_bwd_gather_kernel(const __grid_constant__ immoco_grid_desc g, const __grid_constant__ GatherPlan plan,
                           const uint32_t* __restrict__ row_ptr, const uint2* __restrict__ taps,
                           const float2* __restrict__ d_enc, int n, float2* __restrict__ grad,
                           float2* __restrict__ table, float2* __restrict__ m, float2* __restrict__ v,
                           const AdamConsts a) {
  pdl_wait();
  int level = 0;
  while (level + 1 < g.n_levels && blockIdx.x >= plan.cta_begin[level + 1]) ++level;
  const uint32_t cta = blockIdx.x - plan.cta_begin[level];
  const uint32_t entries = g.entries[level];
  const uint32_t off = g.offset[level];
  if (plan.wide[level]) {
    const uint32_t r = cta * (kThreads / 32) + (threadIdx.x >> 5);
    if (r >= entries) return;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t s = __ldg(row_ptr + off + r), e = __ldg(row_ptr + off + r + 1);
    if (s == e) return;
    RowState st;
    if (lane == 0) st = load_row_state<ADAM>(off + r, table, m, v);
    float ax = 0.f, ay = 0.f;
    for (uint32_t k = s + lane; k < e; k += 64) {       // lane-strided, two taps in flight per lane
      const uint2 t0 = ld_tap(taps + k);
      const bool two = k + 32 < e;
      const uint2 t1 = two ? ld_tap(taps + k + 32) : make_uint2(0u, 0u);
      const float2 d0 = __ldg(d_enc + (t0.x & kPlaneMask));
      const float2 d1 = two ? __ldg(d_enc + (t1.x & kPlaneMask)) : make_float2(0.f, 0.f);
      ax = fmaf(__uint_as_float(t0.y), d0.x, ax);
      ay = fmaf(__uint_as_float(t0.y), d0.y, ay);
      ax = fmaf(__uint_as_float(t1.y), d1.x, ax);
      ay = fmaf(__uint_as_float(t1.y), d1.y, ay);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {                   // fixed butterfly: same order every run
      ax += __shfl_xor_sync(0xffffffffu, ax, o);
      ay += __shfl_xor_sync(0xffffffffu, ay, o);
    }
    if (lane == 0) finish_row<ADAM>(off + r, ax, ay, st, grad, table, m, v, a);
  } else {
    // ---- balanced path: this warp owns rows [r0, r0 + 32) of the level; lane = row (owner role) AND lane = tap
    //      position inside a 32-tap step (worker role)
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t r = cta * kThreads + threadIdx.x;
    const bool row_ok = r < entries;
    const unsigned rows_mask = __ballot_sync(0xffffffffu, row_ok);
    if (rows_mask == 0u) return;
    uint32_t s = 0u, e = 0u;
    if (row_ok) { s = __ldg(row_ptr + off + r); e = __ldg(row_ptr + off + r + 1); }
    const uint32_t S = __shfl_sync(0xffffffffu, s, 0);
    const uint32_t E = __shfl_sync(0xffffffffu, e, 31 - __clz(rows_mask));
    if (S == E) return;                                   // none of the 32 rows is touched
    RowState st;
    if (row_ok && s != e) st = load_row_state<ADAM>(off + r, table, m, v);
    float ax = 0.f, ay = 0.f;
    // one 32-tap step: worker lanes hold contribution c of tap K0 + lane; segmented inclusive scan over the lanes
    // (taps are sorted by row, so a row's taps are a run of lanes); the owner of a row reads the scan value at the
    // last lane of its run inside this step
    auto step = [&](uint32_t K0, uint2 t, float2 d) {
      const bool valid = K0 + lane < E;
      const float w = valid ? __uint_as_float(t.y) : 0.f;
      const uint32_t rid = valid ? (t.x >> kRidShift) : 32u;
      float cx = w * d.x, cy = w * d.y;
      const uint32_t rid_prev = __shfl_up_sync(0xffffffffu, rid, 1);
      const unsigned heads = __ballot_sync(0xffffffffu, lane == 0u || rid_prev != rid);
      const int run_start = 31 - __clz(heads & (0xffffffffu >> (31u - lane)));
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float vx = __shfl_up_sync(0xffffffffu, cx, o);
        const float vy = __shfl_up_sync(0xffffffffu, cy, o);
        if ((int)lane - o >= run_start) { cx += vx; cy += vy; }
      }
      const bool has = row_ok && s < K0 + 32u && e > K0 && s != e;
      const uint32_t last = (e < K0 + 32u ? e : K0 + 32u) - 1u - K0;
      const float px = __shfl_sync(0xffffffffu, cx, has ? last : 0u);
      const float py = __shfl_sync(0xffffffffu, cy, has ? last : 0u);
      if (has) { ax += px; ay += py; }
    };
    for (uint32_t K0 = S; K0 < E; K0 += 64u) {            // two steps per trip: both gathers in flight
      const uint32_t ka = K0 + lane, kb = K0 + 32u + lane;
      const uint2 ta = ka < E ? ld_tap(taps + ka) : make_uint2(0u, 0u);
      const uint2 tb = kb < E ? ld_tap(taps + kb) : make_uint2(0u, 0u);
      const float2 da = ka < E ? __ldg(d_enc + (ta.x & kPlaneMask)) : make_float2(0.f, 0.f);
      const float2 db = kb < E ? __ldg(d_enc + (tb.x & kPlaneMask)) : make_float2(0.f, 0.f);
      step(K0, ta, da);
      if (K0 + 32u < E) step(K0 + 32u, tb, db);           // warp-uniform
    }
    if (row_ok && s != e) finish_row<ADAM>(off + r, ax, ay, st, grad, table, m, v, a);
  }
}

int check(const immoco_grid_desc* g, int64_t n) {
  if (!g || n < 0) return IMMOCO_ERR_BAD_ARG;
  if (g->n_levels < 1 || g->n_levels > IMMOCO_MAX_LEVELS) return IMMOCO_ERR_BAD_ARG;
  if (g->n_dims != 2 && g->n_dims != 3) return IMMOCO_ERR_UNSUPPORTED;
  // tap ids (point * 2^D + corner) and tap offsets (level * n * 2^D + k) are 32-bit
  if ((n << g->n_dims) * (int64_t)g->n_levels >= ((int64_t)1 << 32)) return IMMOCO_ERR_UNSUPPORTED;
  // plane indices (level * n + point) share a word with 5 row-in-warp bits
  if (n * (int64_t)g->n_levels >= ((int64_t)1 << 27)) return IMMOCO_ERR_UNSUPPORTED;
  return 0;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int row_bits(uint32_t entries) {
  int b = 0;
  while (((uint64_t)1 << b) < (uint64_t)entries) ++b;
  return b;
}

int run_gather(const immoco_grid_desc* grid, const immoco_grid_csr* csr, const float* d_enc, float* grad_table,
               float* table, float* exp_avg, float* exp_avg_sq, const AdamConsts& a, bool adam, void* stream) {
  if (!csr || !csr->row_ptr || !csr->taps || !d_enc) return IMMOCO_ERR_BAD_ARG;
  if (int e = check(grid, csr->n_points)) return e;
  if (csr->n_points == 0) return 0;
  if (csr->n_taps != (csr->n_points << grid->n_dims) * (int64_t)grid->n_levels) return IMMOCO_ERR_BAD_ARG;
  GatherPlan plan;
  const int64_t taps_per_level = csr->n_points << grid->n_dims;
  uint32_t ctas = 0;
  for (int l = 0; l < grid->n_levels; ++l) {
    plan.cta_begin[l] = ctas;
    const uint32_t entries = grid->entries[l];
    plan.wide[l] = (taps_per_level >= 8 * (int64_t)entries) ? 1u : 0u;
    const uint32_t rows_per_cta = plan.wide[l] ? kThreads / 32 : kThreads;
    ctas += (entries + rows_per_cta - 1) / rows_per_cta;
  }
  for (int l = grid->n_levels; l <= IMMOCO_MAX_LEVELS; ++l) plan.cta_begin[l] = ctas;
  cudaStream_t s = (cudaStream_t)stream;
  if (adam)
    immoco_launch(hashgrid_bwd_gather_kernel<true>, dim3(ctas), dim3(kThreads), 0, s, *grid, plan, csr->row_ptr,
                  (const uint2*)csr->taps, (const float2*)d_enc, (int)csr->n_points, (float2*)grad_table,
                  (float2*)table, (float2*)exp_avg, (float2*)exp_avg_sq, a);
  else
    immoco_launch(hashgrid_bwd_gather_kernel<false>, dim3(ctas), dim3(kThreads), 0, s, *grid, plan, csr->row_ptr,
                  (const uint2*)csr->taps, (const float2*)d_enc, (int)csr->n_points, (float2*)grad_table,
                  (float2*)nullptr, (float2*)nullptr, (float2*)nullptr, a);
  IMMOCO_LAUNCH_CHECK();
  return 0;
}

}  // namespace

extern "C" int64_t immoco_hashgrid_csr_workspace_bytes(const immoco_grid_desc* grid, int64_t n_points) {
  if (check(grid, n_points)) return -1;
  const int64_t n_level = n_points << grid->n_dims;
  if (n_level == 0) return 256;
  size_t temp = 0;
  if (cub::DeviceRadixSort::SortKeys(nullptr, temp, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                     n_level, 0, 64) != cudaSuccess)
    return -1;
  return (int64_t)(2 * align_up((size_t)n_level * 8, 256) + align_up(temp, 256) + 256);
}

extern "C" int immoco_hashgrid_csr_build(const immoco_grid_desc* grid, const float* coords, int64_t n_points,
                                         uint32_t* row_ptr, void* taps, void* workspace, int64_t workspace_bytes,
                                         void* stream) {
  if (int e = check(grid, n_points)) return e;
  if (grid_has_lut(*grid)) return IMMOCO_ERR_UNSUPPORTED;      // chunk-table layouts: grouped scatter kernels only
  if (!coords || !row_ptr || !taps || !workspace) return IMMOCO_ERR_BAD_ARG;
  if (((uintptr_t)taps & 7) != 0 || ((uintptr_t)workspace & 255) != 0) return IMMOCO_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  const int D = grid->n_dims;
  const int64_t n_level = n_points << D;
  if (n_level == 0) {
    return cudaMemsetAsync(row_ptr, 0, ((size_t)grid->offset[grid->n_levels] + 1) * sizeof(uint32_t), s) == cudaSuccess
               ? 0 : IMMOCO_ERR_BAD_ARG;
  }
  const size_t key_bytes = align_up((size_t)n_level * 8, 256);
  unsigned long long* keys_a = (unsigned long long*)workspace;
  unsigned long long* keys_b = (unsigned long long*)((char*)workspace + key_bytes);
  void* temp = (char*)workspace + 2 * key_bytes;
  size_t temp_bytes = 0;
  if (cub::DeviceRadixSort::SortKeys(nullptr, temp_bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                     n_level, 0, 64) != cudaSuccess)
    return IMMOCO_ERR_BAD_ARG;
  if ((int64_t)(2 * key_bytes + temp_bytes) > workspace_bytes) return IMMOCO_ERR_BAD_ARG;
  const int n = (int)n_points;
  const unsigned blocks_pts = (unsigned)std::min<int64_t>(ceil_div64(n_points, kThreads), 148 * 16);
  const unsigned blocks_taps = (unsigned)std::min<int64_t>(ceil_div64(n_level, kThreads), 148 * 32);
  for (int l = 0; l < grid->n_levels; ++l) {
    if (D == 2) csr_emit_keys_kernel<2><<<blocks_pts, kThreads, 0, s>>>(*grid, coords, n, l, keys_a);
    else csr_emit_keys_kernel<3><<<blocks_pts, kThreads, 0, s>>>(*grid, coords, n, l, keys_a);
    IMMOCO_LAUNCH_CHECK();
    size_t tb = temp_bytes;
    const cudaError_t e = cub::DeviceRadixSort::SortKeys(temp, tb, (const unsigned long long*)keys_a, keys_b, n_level, 0,
                                                         32 + row_bits(grid->entries[l]), s);
    if (e != cudaSuccess) return (int)e;
    uint2* out = (uint2*)taps + (size_t)l * n_level;
    if (D == 2) csr_fill_taps_kernel<2><<<blocks_taps, kThreads, 0, s>>>(*grid, coords, n, l, keys_b, (uint32_t)n_level, out);
    else csr_fill_taps_kernel<3><<<blocks_taps, kThreads, 0, s>>>(*grid, coords, n, l, keys_b, (uint32_t)n_level, out);
    IMMOCO_LAUNCH_CHECK();
    const uint32_t entries = grid->entries[l];
    const int last = (l == grid->n_levels - 1) ? 1 : 0;
    csr_row_ptr_kernel<<<(entries + kThreads) / kThreads, kThreads, 0, s>>>(keys_b, (uint32_t)n_level, entries,
                                                                            (uint32_t)((int64_t)l * n_level), last,
                                                                            row_ptr + grid->offset[l]);
    IMMOCO_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int immoco_hashgrid_bwd_csr(const immoco_grid_desc* grid, const immoco_grid_csr* csr, const float* d_enc,
                                       float* grad_table, void* stream) {
  if (!grad_table) return IMMOCO_ERR_BAD_ARG;
  AdamConsts a = {};
  return run_gather(grid, csr, d_enc, grad_table, nullptr, nullptr, nullptr, a, false, stream);
}

extern "C" int immoco_hashgrid_bwd_csr_adam(const immoco_grid_desc* grid, const immoco_grid_csr* csr, const float* d_enc,
                                            float* table, float* exp_avg, float* exp_avg_sq, float* grad_table,
                                            double lr, double beta1, double beta2, double eps, int32_t step,
                                            void* stream) {
  if (!table || !exp_avg || !exp_avg_sq || step < 1) return IMMOCO_ERR_BAD_ARG;
  if ((((uintptr_t)table | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq | (uintptr_t)grad_table) & 7) != 0)
    return IMMOCO_ERR_BAD_ARG;
  const AdamScalars sc = adam_scalars(lr, beta1, beta2, eps, step);
  const AdamConsts a = {sc.omb1, sc.b2, sc.omb2, sc.step_size, sc.bc2_sqrt, sc.eps};
  return run_gather(grid, csr, d_enc, grad_table, table, exp_avg, exp_avg_sq, a, true, stream);
}
