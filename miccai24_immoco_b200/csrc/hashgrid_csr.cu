// Deterministic hash-grid backward: row-sorted tap list (CSR) built once per coordinate set, then a
// gather with a fixed summation order per table row -- optionally with the Adam update of that row fused in.
// Replaces tiny-cuda-nn's kernel_grid_backward (atomic scatter) + the table part of torch.optim.Adam.step
// on the IM-MoCo fit path (src/models/immoco.py:149-154,164-175): the fit's coordinates are registered
// buffers (immoco.py:72-80), i.e. WHICH point touches WHICH row with WHICH weight never changes.
//
//   build : per level  keys = (row << 32 | point * 2^D + corner)  -> radix sort (CUB, set-up only)
//           -> taps[k] = {point, weight}, row_ptr[row] = lower bound of the row in the sorted keys
//   gather: thread (narrow levels) or warp (coarse levels: hundreds of taps per row) per table row:
//           g[row] = sum_k w_k * d_enc[level][point_k] in tap order -> fixed order, no atomics;
//           rows without taps are skipped (their gradient / moments stay zero for ever).
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace {

constexpr int kThreads = 256;

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
csr_fill_taps_kernel(const __grid_constant__ immoco_grid_desc g, const float* __restrict__ coords, int level,
                     const unsigned long long* __restrict__ keys, uint32_t n_taps, uint2* __restrict__ taps) {
  const float scale = g.scale[level];
  for (uint32_t k = blockIdx.x * kThreads + threadIdx.x; k < n_taps; k += gridDim.x * kThreads) {
    const uint32_t tap = (uint32_t)(keys[k] & 0xffffffffull);
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
    taps[k] = make_uint2(point, __float_as_uint(w));
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

// taps per thread and trip of the narrow path: all index records first, then all gathers, then the
// ordered sum -- 4 independent L2 round trips in flight per thread (8 measured slower: 282 vs 269 us for the
// 3-D grid at C2; the gather sits at the measured L2 rate for unpaired 8-byte gathers, profiles/round2_l2_peaks.json)
constexpr int kUnroll = 4;

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
  const float2* __restrict__ go = d_enc + (size_t)level * n;
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
      const uint2 t0 = __ldg(taps + k);
      const bool two = k + 32 < e;
      const uint2 t1 = two ? __ldg(taps + k + 32) : make_uint2(0u, 0u);
      const float2 d0 = __ldg(go + t0.x);
      const float2 d1 = two ? __ldg(go + t1.x) : make_float2(0.f, 0.f);
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
    const uint32_t r = cta * kThreads + threadIdx.x;
    if (r >= entries) return;
    const uint32_t s = __ldg(row_ptr + off + r), e = __ldg(row_ptr + off + r + 1);
    if (s == e) return;
    const RowState st = load_row_state<ADAM>(off + r, table, m, v);
    float ax = 0.f, ay = 0.f;
    for (uint32_t k = s; k < e; k += kUnroll) {
      uint2 t[kUnroll];
      float2 d[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) t[u] = (k + u < e) ? __ldg(taps + k + u) : make_uint2(0u, 0u);
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) d[u] = (k + u < e) ? __ldg(go + t[u].x) : make_float2(0.f, 0.f);
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {                // weight 0 for the padding taps: adds +0
        ax = fmaf(__uint_as_float(t[u].y), d[u].x, ax);
        ay = fmaf(__uint_as_float(t[u].y), d[u].y, ay);
      }
    }
    finish_row<ADAM>(off + r, ax, ay, st, grad, table, m, v, a);
  }
}

int check(const immoco_grid_desc* g, int64_t n) {
  if (!g || n < 0) return IMMOCO_ERR_BAD_ARG;
  if (g->n_levels < 1 || g->n_levels > IMMOCO_MAX_LEVELS) return IMMOCO_ERR_BAD_ARG;
  if (g->n_dims != 2 && g->n_dims != 3) return IMMOCO_ERR_UNSUPPORTED;
  // tap ids (point * 2^D + corner) and tap offsets (level * n * 2^D + k) are 32-bit
  if ((n << g->n_dims) * (int64_t)g->n_levels >= ((int64_t)1 << 32)) return IMMOCO_ERR_UNSUPPORTED;
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
    if (D == 2) csr_fill_taps_kernel<2><<<blocks_taps, kThreads, 0, s>>>(*grid, coords, l, keys_b, (uint32_t)n_level, out);
    else csr_fill_taps_kernel<3><<<blocks_taps, kThreads, 0, s>>>(*grid, coords, l, keys_b, (uint32_t)n_level, out);
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
