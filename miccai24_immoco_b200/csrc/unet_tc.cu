// kld-net 3x3 convolutions on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in
// TMEM), fp32 parity through the 3xTF32 split of tc_common.cuh.  Replaces the Conv2d(k=3, pad=1, no bias)
// layers of fastmri.models.Unet == src/models/unet.py:17-187 (used at src/test/test_immoco.py:17-20,50-58);
// csrc/unet.cu keeps the fp32 SIMT kernel for the 2-channel input layer and everything that is not a 3x3
// convolution.
//
// Implicit GEMM, no im2col: a CTA owns 16 x 8 output pixels of one image (= the 128 rows / TMEM lanes of the
// accumulator) and NT output channels (= its columns), and walks the input channels 8 at a time (one K = 8
// MMA step).  The 18 x 10 input window (halo included) of those channels is staged ONCE per step in shared
// memory as [channel quad][row][x][4 channels]: a pixel's 4 channels are one 16-byte row of a UMMA core
// matrix and 8 consecutive x are one core matrix, so the A operand of filter tap (ky, kx) is the SAME buffer
// read through a descriptor whose start address is advanced by ky rows + kx pixels
// (SBO = window row pitch, LBO = channel-quad pitch) -- nine shifted views instead of nine copies.
// The weights are pre-packed once per layer (immoco_unet_pack_conv3x3) as [channel quad][tap][cout][4],
// already split into tf32 hi / lo parts, so staging the B operand is a plain 16-byte copy.
// Three accumulators (hi*hi, lo*hi, hi*lo) keep consecutive MMAs independent; every kFlush steps they are
// folded into fp32 registers (TMEM accumulation truncates: long chains would bias the sum).
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int kTY = 16, kTX = 8;                    // output tile
constexpr int kInY = kTY + 2, kInX = kTX + 2;       // input window incl. halo
constexpr int kWin = kInY * kInX;                   // 180 pixels
constexpr int kThreads = 256;
constexpr int kCs = 8;                              // input channels per step
constexpr int kQuadBytes = kWin * 16;               // one 4-channel quad of the window: 2880 B
constexpr int kRowPitch = kInX * 16;                // 160 B
constexpr int kFlush = 16;                          // steps between accumulator flushes

template <int NT>
struct ConvSmem {
  static constexpr int a_bytes = 2 * kQuadBytes;                 // hi (or lo) part of one step: 8 channels
  static constexpr int b_bytes = 2 * 9 * NT * 16;                // hi (or lo): 2 quads x 9 taps x NT x 16 B
  static constexpr int stage_bytes = 2 * a_bytes + 2 * b_bytes;
  static constexpr int off_misc = 2 * stage_bytes;               // 2 mbarriers (16 B) + TMEM base (4 B)
  static constexpr int off_stats = off_misc + 32;                // [4 lane quadrants][NT][2] floats
  static constexpr int total = off_stats + 4 * NT * 2 * 4;
};

template <int NT>
__global__ void __launch_bounds__(kThreads)
conv3x3_tc_kernel(const float* __restrict__ in0, int c0, const float* __restrict__ in1, int c1,
                  const float4* __restrict__ w_hi, const float4* __restrict__ w_lo, float* __restrict__ out,
                  double* __restrict__ stats, int cout, int h, int w, int tiles_x) {
  using S = ConvSmem<NT>;
  constexpr int kHalf = NT / 2;                     // accumulator columns per thread
  constexpr uint32_t kCols = (3 * NT <= 128) ? 128u : 256u;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + S::off_misc);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::off_misc + 16);
  float* st = reinterpret_cast<float*>(smem + S::off_stats);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, cs = warp >> 2;        // TMEM lane quadrant, column half
  const int tile = blockIdx.x;
  const int ty0 = (tile / tiles_x) * kTY, tx0 = (tile % tiles_x) * kTX;
  const int co_base = blockIdx.y * NT, n = blockIdx.z;
  const int cin = c0 + c1, steps = cin / kCs;
  const size_t plane = (size_t)h * w;

  if (tid == 0) {
    tc::mbar_init(bar, 1);
    tc::mbar_init(bar + 1, 1);
    tc::mbar_fence_init();
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, kCols);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tm = *tmem_slot;
  const uint32_t trow = tm + ((uint32_t)(quad * 32) << 16);
  constexpr uint32_t idesc = tc::idesc_tf32(128, NT, 0, 0);

  float acc[kHalf];
#pragma unroll
  for (int j = 0; j < kHalf; ++j) acc[j] = 0.f;
  bool pending[2] = {false, false};
  uint32_t phase[2] = {0u, 0u};

  for (int s = 0; s < steps; ++s) {
    const int buf = s & 1;
    if (pending[buf]) {                 // the MMAs of step s - 2 still read this buffer
      tc::mbar_wait(bar + buf, phase[buf]);
      phase[buf] ^= 1u;
      pending[buf] = false;
    }
    unsigned char* base = smem + buf * S::stage_bytes;
    float* a_hi = reinterpret_cast<float*>(base);
    float* a_lo = reinterpret_cast<float*>(base + S::a_bytes);
    float4* b_hi = reinterpret_cast<float4*>(base + 2 * S::a_bytes);
    float4* b_lo = reinterpret_cast<float4*>(base + 2 * S::a_bytes + S::b_bytes);
    // ---- A: 8 channels x 18 x 10 window, zero outside the image; 4 adjacent lanes = the 4 channels of a pixel
    for (int idx = tid; idx < kCs * kWin; idx += kThreads) {
      const int cl = idx & 3, rest = idx >> 2;
      const int cq = rest / kWin, pix = rest - cq * kWin;
      const int iy = pix / kInX, ix = pix - iy * kInX;
      const int gy = ty0 + iy - 1, gx = tx0 + ix - 1, ch = s * kCs + cq * 4 + cl;
      float v = 0.f;
      if (gy >= 0 && gy < h && gx >= 0 && gx < w) {
        const float* src = ch < c0 ? in0 + ((size_t)n * c0 + ch) * plane : in1 + ((size_t)n * c1 + (ch - c0)) * plane;
        v = __ldg(src + (size_t)gy * w + gx);
      }
      const float hi = tc::tf32_hi(v);
      const int o = (cq * kWin + pix) * 4 + cl;
      a_hi[o] = hi;
      a_lo[o] = v - hi;
    }
    // ---- B: [quad][tap][NT] 16-byte rows of the packed weights
    for (int idx = tid; idx < 2 * 9 * NT; idx += kThreads) {
      const int nn = idx % NT, qt = idx / NT;       // qt = quad * 9 + tap
      const size_t g = ((size_t)(s * 2 * 9 + qt)) * cout + co_base + nn;
      b_hi[idx] = __ldg(w_hi + g);
      b_lo[idx] = __ldg(w_lo + g);
    }
    tc::fence_proxy_async();
    __syncthreads();                    // also orders the flush's TMEM reads before the MMAs that overwrite
    if (warp == 0 && tc::elect_one()) {
      tc::fence_after_sync();
      const uint32_t sa_hi = tc::smem_u32(a_hi), sa_lo = tc::smem_u32(a_lo);
      const uint32_t sb_hi = tc::smem_u32(b_hi), sb_lo = tc::smem_u32(b_lo);
      const bool fresh = (s % kFlush) == 0;
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const uint32_t a_off = (uint32_t)((tap / 3) * kRowPitch + (tap % 3) * 16);
        const uint32_t b_off = (uint32_t)(tap * NT * 16);
        const uint64_t da_hi = tc::smem_desc(sa_hi + a_off, kQuadBytes, kRowPitch);
        const uint64_t da_lo = tc::smem_desc(sa_lo + a_off, kQuadBytes, kRowPitch);
        const uint64_t db_hi = tc::smem_desc(sb_hi + b_off, 9 * NT * 16, 128);
        const uint64_t db_lo = tc::smem_desc(sb_lo + b_off, 9 * NT * 16, 128);
        const uint32_t accum = (fresh && tap == 0) ? 0u : 1u;
        tc::mma_ss(tm, da_hi, db_hi, idesc, accum);
        tc::mma_ss(tm + NT, da_lo, db_hi, idesc, accum);
        tc::mma_ss(tm + 2 * NT, da_hi, db_lo, idesc, accum);
      }
      tc::mma_commit(bar + buf);
    }
    pending[buf] = true;
    if ((s + 1) % kFlush == 0 || s == steps - 1) {
      // fold the three accumulators into registers (MMAs complete in order: this step done => all done)
      tc::mbar_wait(bar + buf, phase[buf]);
      phase[buf] ^= 1u;
      pending[buf] = false;
      tc::fence_after_sync();
#pragma unroll
      for (int j0 = 0; j0 < kHalf; j0 += 16) {
        uint32_t v0[16], v1[16], v2[16];
        tc::tmem_ld16(trow + cs * kHalf + j0, v0);
        tc::tmem_ld16(trow + NT + cs * kHalf + j0, v1);
        tc::tmem_ld16(trow + 2 * NT + cs * kHalf + j0, v2);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j)
          acc[j0 + j] += (__uint_as_float(v0[j]) + __uint_as_float(v1[j])) + __uint_as_float(v2[j]);
      }
      tc::fence_before_sync();
    }
  }

  // ---- raw output + instance statistics (per (image, channel) sum / sum of squares, fp64 atomics) ---------
  const int r = quad * 32 + lane;                   // TMEM lane = pixel of the tile
  const int gy = ty0 + (r >> 3), gx = tx0 + (r & 7);
  const bool live = gy < h && gx < w;
#pragma unroll
  for (int j = 0; j < kHalf; ++j) {
    const int col = cs * kHalf + j;
    const int co = co_base + col;
    const float v = live ? acc[j] : 0.f;
    if (live) out[((size_t)n * cout + co) * plane + (size_t)gy * w + gx] = v;
    const float sum = warp_sum(v), sq = warp_sum(v * v);
    if (lane == 0) {
      st[(quad * NT + col) * 2 + 0] = sum;
      st[(quad * NT + col) * 2 + 1] = sq;
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (tid < NT) {
    double sum = 0.0, sq = 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      sum += (double)st[(q * NT + tid) * 2 + 0];
      sq += (double)st[(q * NT + tid) * 2 + 1];
    }
    atomicAdd(stats + ((size_t)n * cout + co_base + tid) * 2 + 0, sum);
    atomicAdd(stats + ((size_t)n * cout + co_base + tid) * 2 + 1, sq);
  }
  if (warp == 0) tc::tmem_dealloc(tm, kCols);
}

// weight (cout, cin, 3, 3) -> packed [cin / 4][9][cout] float4 (4 consecutive input channels), tf32 hi / lo
__global__ void __launch_bounds__(256)
pack_conv3x3_kernel(const float* __restrict__ weight, float4* __restrict__ w_hi, float4* __restrict__ w_lo, int cout,
                    int cin) {
  const int total = (cin / 4) * 9 * cout;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int co = idx % cout, qt = idx / cout;
    const int q = qt / 9, tap = qt - q * 9;
    float v[4], hi[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      v[c] = __ldg(weight + ((size_t)co * cin + 4 * q + c) * 9 + tap);
      hi[c] = tc::tf32_hi(v[c]);
    }
    w_hi[idx] = make_float4(hi[0], hi[1], hi[2], hi[3]);
    w_lo[idx] = make_float4(v[0] - hi[0], v[1] - hi[1], v[2] - hi[2], v[3] - hi[3]);
  }
}

template <int NT>
int launch_conv(const float* in0, int c0, const float* in1, int c1, const float* w_hi, const float* w_lo, float* out,
                double* stats, int n, int cout, int h, int w, cudaStream_t s) {
  constexpr int smem = ConvSmem<NT>::total;
  static DeviceOnce once;
  if (once.first()) cudaFuncSetAttribute(conv3x3_tc_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int tiles_x = (w + kTX - 1) / kTX, tiles_y = (h + kTY - 1) / kTY;
  if (n > 65535 || cout / NT > 65535) return IMMOCO_ERR_UNSUPPORTED;
  conv3x3_tc_kernel<NT><<<dim3(tiles_x * tiles_y, cout / NT, n), kThreads, smem, s>>>(
      in0, c0, in1, c1, (const float4*)w_hi, (const float4*)w_lo, out, stats, cout, h, w, tiles_x);
  IMMOCO_LAUNCH_CHECK();
  return 0;
}

}  // namespace

extern "C" int immoco_unet_pack_conv3x3(const float* weight, float* w_hi, float* w_lo, int32_t cout, int32_t cin,
                                        void* stream) {
  if (!weight || !w_hi || !w_lo || cout < 1 || cin < 4 || (cin & 3) != 0) return IMMOCO_ERR_BAD_ARG;
  if ((((uintptr_t)w_hi | (uintptr_t)w_lo) & 15) != 0) return IMMOCO_ERR_BAD_ARG;
  const int total = (cin / 4) * 9 * cout;
  pack_conv3x3_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(weight, (float4*)w_hi, (float4*)w_lo, cout, cin);
  IMMOCO_LAUNCH_CHECK();
  return 0;
}

// in1 may be NULL (c1 = 0).  (c0 + c1) % 8 == 0 and cout % 32 == 0, otherwise IMMOCO_ERR_UNSUPPORTED (the
// caller keeps immoco_unet_conv3x3 for those).  stats: (n * cout * 2) doubles, ZEROED by the caller.
extern "C" int immoco_unet_conv3x3_tc(const float* in0, int32_t c0, const float* in1, int32_t c1, const float* w_hi,
                                      const float* w_lo, float* out, double* stats, int32_t n, int32_t cout, int32_t h,
                                      int32_t w, void* stream) {
  if (!in0 || !w_hi || !w_lo || !out || !stats || c0 < 1 || c1 < 0 || (c1 > 0 && !in1) || n < 0 || cout < 1 || h < 1 || w < 1)
    return IMMOCO_ERR_BAD_ARG;
  if (((c0 + c1) % kCs) != 0 || (cout % 32) != 0 || (c0 & 3) != 0) return IMMOCO_ERR_UNSUPPORTED;
  if (n == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (cout % 64 == 0) return launch_conv<64>(in0, c0, in1, c1, w_hi, w_lo, out, stats, n, cout, h, w, s);
  return launch_conv<32>(in0, c0, in1, c1, w_hi, w_lo, out, stats, n, cout, h, w, s);
}
