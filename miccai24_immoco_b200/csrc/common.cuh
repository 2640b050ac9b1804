// Shared device helpers for the IM-MoCo sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "immoco_b200.h"

// SM count of the CURRENT device (B200: 148 = 2 dies x 74), queried once per device (fit.cu); grids are
// sized as multiples of it
int immoco_num_sms();
#define IMMOCO_NUM_SMS (immoco_num_sms())

#define IMMOCO_LAUNCH_CHECK()                      \
  do {                                             \
    cudaError_t e__ = cudaGetLastError();          \
    if (e__ != cudaSuccess) return (int)e__;       \
  } while (0)

// ---- programmatic dependent launch (PDL) -----------------------------------------------------------
// Every hot-path kernel is launched with cudaLaunchAttributeProgrammaticStreamSerialization and begins
// with pdl_wait() (griddepcontrol.wait: the preceding kernel of the stream has completed and its
// writes are visible), so the ~18 dependent launches of an iteration are resolved on the device
// instead of by the host-side stream scheduler: 826 -> 808 us per iteration on B200
// (profiles/round1_v5_pdl_ab.txt).  No kernel triggers its dependents early
// (griddepcontrol.launch_dependents at CTA start was measured SLOWER, 966 us: the early-resident CTAs
// of the next kernel take SM resources from the last wave of the running one).
// Work done BEFORE pdl_wait() may only read buffers written at least two kernels earlier (MLP weights).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

int immoco_pdl_enabled();   // fit.cu

// "first call on this device?" -- function attributes (dynamic shared-memory opt-in) are per device, so a
// launcher keeps one of these as a function-local static instead of a process-wide flag
struct DeviceOnce {
  bool done[64] = {};
  bool first() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
    if (done[dev]) return false;
    done[dev] = true;
    return true;
  }
};

template <typename... P, typename... A>
inline cudaError_t immoco_launch(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 A... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = immoco_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(args)...);
}

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------------
// hash-grid index maths (tiny-cuda-nn grid.h grid_index / coherent-prime hash, restated in
// oracle/immoco_oracle.py:grid_corner_index).  All arithmetic is uint32 and wraps.
// ---------------------------------------------------------------------------------------------
// physical row of hash index r under immoco_grid_desc::swizzle (0: identity): Gray code, then two bit
// positions exchanged -- a linear bijection on the index bits
__host__ __device__ __forceinline__ uint32_t grid_swizzle(uint32_t r, uint32_t swz) {
  if (swz == 0u) return r;
  r ^= r >> 1;
  const uint32_t a = swz & 0xffu, b = (swz >> 8) & 0xffu;
  const uint32_t x = ((r >> a) ^ (r >> b)) & 1u;
  return r ^ ((x << a) | (x << b));
}

// physical row of hash index r under a level's chunk tables (immoco_grid_desc::layout_lut): S is linear over the
// index bits, so it is the XOR of the images of three bit groups
__host__ __device__ __forceinline__ uint32_t grid_lut_row(const uint32_t* t, uint32_t r) {
  return t[r & 127u] ^ t[128u + ((r >> 7) & 63u)] ^ t[192u + ((r >> 13) & 63u)];
}
// the level's chunk tables when the descriptor stores the level under a general linear layout, else nullptr
__host__ __device__ __forceinline__ const uint32_t* grid_level_lut(const immoco_grid_desc& g, int level) {
  return (g.swizzle[level] == IMMOCO_LAYOUT_LUT && g.layout_lut) ? g.layout_lut + 256 * level : nullptr;
}
inline bool grid_has_lut(const immoco_grid_desc& g) {
  for (int l = 0; l < g.n_levels; ++l)
    if (g.swizzle[l] == IMMOCO_LAYOUT_LUT) return true;
  return false;
}

template <int D>
__host__ __device__ __forceinline__ uint32_t grid_index(const uint32_t (&q)[D], uint32_t hashed,
                                                        uint32_t entries, uint32_t res, uint32_t swz = 0u) {
  uint32_t idx;
  if (hashed) {
    idx = q[0];  // prime 1
    if (D > 1) idx ^= q[1] * 2654435761u;
    if (D > 2) idx ^= q[2] * 805459861u;
  } else {
    uint32_t stride = 1;
    idx = 0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      idx += q[d] * stride;
      stride *= res;
    }
  }
  // entries is a power of two for every level of the reference configuration; keep the general
  // modulo for others (round-up-to-8 dense levels of odd resolutions).
  if ((entries & (entries - 1)) != 0) return idx % entries;
  idx &= entries - 1;
  return (hashed && swz) ? grid_swizzle(idx, swz) : idx;
}

// pos = fmaf(scale, x, 0.5); cell = (uint32)(int)floor(pos); frac = pos - floor(pos)
__host__ __device__ __forceinline__ void grid_pos(float x, float scale, uint32_t& cell, float& frac) {
  float pos = fmaf(scale, x, 0.5f);
  float fl = floorf(pos);
  frac = pos - fl;
  cell = (uint32_t)(int)fl;
}

// ---------------------------------------------------------------------------------------------
// torch.optim.Adam (amsgrad=False, weight_decay=0, maximize=False), single-tensor formulation
// (src/models/immoco.py:149-154):  m = lerp(m, g, 1-b1);  v = b2*v + (1-b2)*g*g;
// p -= step_size * m / (sqrt(v)/bc2_sqrt + eps).  ONE definition, explicit roundings, so the stand-alone
// Adam kernels and the update fused into the hash-grid gather produce the same bits.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void adam_update(float& p, float& m, float& v, float g, float omb1, float b2, float omb2,
                                            float step_size, float bc2_sqrt, float eps) {
  m = __fmaf_rn(omb1, __fsub_rn(g, m), m);
  v = __fmaf_rn(__fmul_rn(omb2, g), g, __fmul_rn(b2, v));
  const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), bc2_sqrt), eps);
  p = __fmaf_rn(-step_size, __fdiv_rn(m, denom), p);
}

// bias corrections in double like torch (python floats), then rounded to fp32 scalars
struct AdamScalars {
  float omb1, b2, omb2, step_size, bc2_sqrt, eps;
};
AdamScalars adam_scalars(double lr, double beta1, double beta2, double eps, int step);   // fit.cu

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// a * conj(b)
__device__ __forceinline__ float2 cmul_conj(float2 a, float2 b) {
  return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == IMMOCO_ACT_RELU) return fmaxf(x, 0.0f);
  if (act == IMMOCO_ACT_TANH) return tanhf(x);
  return x;
}
// derivative expressed through the activation OUTPUT y
__device__ __forceinline__ float act_grad_from_out(float y, int act) {
  if (act == IMMOCO_ACT_RELU) return y > 0.0f ? 1.0f : 0.0f;
  if (act == IMMOCO_ACT_TANH) return 1.0f - y * y;
  return 1.0f;
}
