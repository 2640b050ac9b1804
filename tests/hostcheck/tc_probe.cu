// GPU bring-up probe for tcgen05 operand conventions (test-only; not part of the product library).
// One K=8 tf32 MMA, D[128 x 32] = A[128 x 8] . B[32 x 8]^T, with the operand variants the backward
// kernel relies on; dumps every D so the host can tell which conventions the hardware follows.
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../miccai24_immoco_b200/csrc/tc_common.cuh"

__device__ float a_val(int m, int k) { return (float)((m % 7) - 3) + 0.25f * (float)k; }
__device__ float b_val(int n, int k) { return (float)((n % 5) - 2) + 0.5f * (float)k; }

__global__ void __launch_bounds__(128) tc_probe_kernel(float* out, int variant_mask) {
  extern __shared__ __align__(128) float smem[];
  float* a_k = smem;                 // A K-major: LBO 2048 B, SBO 128
  float* b_k = a_k + 128 * 8;        // B K-major [32 x 8]: LBO 512, SBO 128
  float* b_mn = b_k + 32 * 8;        // B MN-major placement: SBO' 128 (mn groups), LBO' 1024 (k groups)
  float* misc = b_mn + 1024;
  uint64_t* bar = reinterpret_cast<uint64_t*>(misc);
  uint32_t* slot = reinterpret_cast<uint32_t*>(misc + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int k = 0; k < 8; ++k) a_k[(k >> 2) * 512 + (tid >> 3) * 32 + (tid & 7) * 4 + (k & 3)] = a_val(tid, k);
  for (int i = tid; i < 1024; i += 128) b_mn[i] = 0.f;
  __syncthreads();
  if (tid < 32) {
    for (int k = 0; k < 8; ++k) {
      b_k[(k >> 2) * 128 + (tid >> 3) * 32 + (tid & 7) * 4 + (k & 3)] = b_val(tid, k);
      b_mn[(tid >> 2) * 32 + (k >> 3) * 256 + (k & 7) * 4 + (tid & 3)] = b_val(tid, k);
    }
  }
  if (tid == 0) { tc::mbar_init(bar, 1); tc::mbar_fence_init(); }
  if (warp == 0) tc::tmem_alloc(slot, 256);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tm = *slot;
  const uint32_t trow = tm + ((uint32_t)(warp * 32) << 16);
  // A into TMEM columns [128, 136): thread = lane m
  {
    uint32_t v[32];
    for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(j < 8 ? a_val(tid, j) : 0.f);
    tc::tmem_st32(trow + 128, v);
    tc::tmem_st_wait();
  }
  tc::fence_before_sync();
  __syncthreads();
  const uint32_t sa = tc::smem_u32(a_k), sbk = tc::smem_u32(b_k), sbm = tc::smem_u32(b_mn);
  uint32_t phase = 0;
  for (int v = 0; v < 6; ++v) {
    if (tid == 0) {
      tc::fence_after_sync();
      const uint64_t da = tc::smem_desc(sa, 2048, 128);
      const uint64_t dbk = tc::smem_desc(sbk, 512, 128);
      const uint64_t dbm = tc::smem_desc(sbm, 1024, 128);    // lbo = k-group stride, sbo = mn-group stride
      const uint64_t dbm_sw = tc::smem_desc(sbm, 128, 1024); // swapped roles
      const uint32_t id_kk = tc::idesc_tf32(128, 32, 0, 0), id_kmn = tc::idesc_tf32(128, 32, 0, 1);
      if (v == 0) tc::mma_ss(tm, da, dbk, id_kk, 0);
      if (v == 1) tc::mma_ss(tm, da, dbm, id_kmn, 0);
      if (v == 2) tc::mma_ss(tm, da, dbm_sw, id_kmn, 0);
      if (v == 3) tc::mma_ts(tm, tm + 128, dbk, id_kk, 0);
      if (v == 4) tc::mma_ts(tm, tm + 128, dbm, id_kmn, 0);
      if (v == 5) tc::mma_ts(tm, tm + 128, dbm_sw, id_kmn, 0);
      tc::mma_commit(bar);
    }
    tc::mbar_wait(bar, phase);
    phase ^= 1;
    tc::fence_after_sync();
    uint32_t r[32];
    tc::tmem_ld32(trow, r);
    tc::tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(size_t)v * 4096 + tid * 32 + j] = __uint_as_float(r[j]);
    tc::fence_before_sync();
    __syncthreads();
  }
  if (warp == 0) tc::tmem_dealloc(tm, 256);
}

extern "C" int tc_probe(float* out_dev, void* stream) {
  tc_probe_kernel<<<1, 128, 16384, (cudaStream_t)stream>>>(out_dev, 0);
  return (int)cudaGetLastError();
}
