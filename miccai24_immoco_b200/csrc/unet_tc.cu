// kld-net 3x3 convolutions on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in
// TMEM), fp32 parity through the 3xTF32 split of tc_common.cuh.  Replaces the Conv2d(k=3, pad=1, no bias)
// layers of fastmri.models.Unet == src/models/unet.py:17-187 (used at src/test/test_immoco.py:17-20,50-58);
// csrc/unet.cu keeps the fp32 SIMT kernel for the 2-channel input layer and everything that is not a 3x3
// convolution.
//
// Implicit GEMM, no im2col.  A CTA owns MT horizontally adjacent tiles of 16 x 8 output pixels of one image
// (each tile = the 128 rows / TMEM lanes of an accumulator) and NT output channels (= its columns), and walks
// the input channels 8 at a time (one K = 8 MMA step per filter tap):
//   A operand: the 18 x (8 MT + 2) input window (halo included) of the step's 8 channels is staged ONCE in
//     shared memory as [channel quad][row][x][4 channels]: a pixel's 4 channels are one 16-byte row of a UMMA
//     core matrix and 8 consecutive x are one core matrix, so the operand of filter tap (ky, kx) of tile t is
//     the SAME buffer read through a descriptor whose start address is advanced by ky rows + (8 t + kx) pixels
//     (SBO = window row pitch, LBO = channel-quad pitch): nine shifted views instead of nine copies, and the
//     MT tiles share their halo columns.  Elements are split into tf32 hi / lo on the way (registers),
//     prefetched one step ahead.
//   B operand: the weights are packed once per layer (immoco_unet_pack_conv3x3) as
//     [cout tile][channel quad][tap][NT][4], already split into hi / lo, so a step's B operand is ONE
//     contiguous block: two bulk asynchronous copies (cp.async.bulk, the TMA engine) per step into a 3-deep
//     ring, completing on an mbarrier, issued two steps ahead by a dedicated issuer warp.
//   The MT tiles reuse the step's weights (the kernel is otherwise bound by re-reading them through L2).
//   Three accumulators per tile (hi*hi, lo*hi, hi*lo) keep consecutive MMAs independent; every kFlush steps
//   they are folded into fp32 registers (TMEM accumulation truncates: long chains would bias the sum,
//   DESIGN.md 4.1).
// Warps 0-7 stage A and run the epilogue (TMEM lane quadrant = warp & 3, column half = warp >> 2); warp 8
// issues the bulk copies and the MMAs.
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int kTY = 16, kTX = 8;                    // one output tile
constexpr int kInY = kTY + 2;
constexpr int kStagers = 256;
constexpr int kThreads = kStagers + 32;
constexpr int kCs = 8;                              // input channels per step
constexpr int kRing = 3;                            // shared-memory stages
constexpr int kFlush = 16;                          // steps between accumulator flushes

template <int NT, int MT>
struct ConvCfg {
  static constexpr int win_x = kTX * MT + 2;
  static constexpr int win = kInY * win_x;                       // window pixels
  static constexpr int quad_bytes = win * 16;                    // one 4-channel quad of the window
  static constexpr int row_pitch = win_x * 16;
  static constexpr int a_bytes = 2 * quad_bytes;                 // hi (or lo) part of one step: 8 channels
  static constexpr int b_bytes = 2 * 9 * NT * 16;                // hi (or lo): 2 quads x 9 taps x NT x 16 B
  static constexpr int stage_bytes = 2 * a_bytes + 2 * b_bytes;
  static constexpr int off_misc = kRing * stage_bytes;           // full[3], done[3] mbarriers + TMEM base
  static constexpr int off_stats = off_misc + 64;                // [4 lane quadrants][NT][2] floats
  static constexpr int total = off_stats + 4 * NT * 2 * 4;
  static constexpr int per_thread = (kCs * win + kStagers - 1) / kStagers;
  static constexpr int tmem_cols = 512;                          // MT * 3 * NT = 384 -> next power of two
  static_assert(MT * 3 * NT <= 512, "accumulators must fit TMEM");
};

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc::smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared bulk asynchronous copy (TMA engine, no tensor map), completion counted on `bar`
__device__ __forceinline__ void bulk_copy_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   tc::smem_u32(dst_smem)),
               "l"(src), "r"(bytes), "r"(tc::smem_u32(bar))
               : "memory");
}

template <int NT, int MT>
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_tc_kernel(const float* __restrict__ in0, int c0, const float* __restrict__ in1, int c1,
                  const float4* __restrict__ w_hi, const float4* __restrict__ w_lo, float* __restrict__ out,
                  double* __restrict__ stats, int cout, int h, int w, int tiles_x) {
  using S = ConvCfg<NT, MT>;
  constexpr int kHalf = NT / 2;                     // accumulator columns per stager thread and tile
  constexpr int kPer = S::per_thread;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::off_misc);        // B operand of a stage has landed
  uint64_t* done = full + kRing;                                          // the stage's MMAs have completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::off_misc + 48);
  float* st = reinterpret_cast<float*>(smem + S::off_stats);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool stager = tid < kStagers;
  const int quad = warp & 3, cs = (warp >> 2) & 1;  // TMEM lane quadrant, column half (stager warps)
  const int tile = blockIdx.x;
  const int ty0 = (tile / tiles_x) * kTY, tx0 = (tile % tiles_x) * (kTX * MT);
  const int co_tile = blockIdx.y, co_base = co_tile * NT, n = blockIdx.z;
  const int cin = c0 + c1, steps = cin / kCs;
  const size_t plane = (size_t)h * w;
  // this CTA's packed weights: [co_tile][cin / 4][9][NT] float4
  const float4* wh = w_hi + (size_t)co_tile * (cin / 4) * 9 * NT;
  const float4* wl = w_lo + (size_t)co_tile * (cin / 4) * 9 * NT;

  if (tid == 0) {
    for (int i = 0; i < kRing; ++i) {
      tc::mbar_init(full + i, 1);
      tc::mbar_init(done + i, 1);
    }
    tc::mbar_fence_init();
  }
  if (warp == 8) tc::tmem_alloc(tmem_slot, S::tmem_cols);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tm = *tmem_slot;

  if (!stager) {
    // =============================== issuer warp: bulk copies + MMAs ===================================
    constexpr uint32_t idesc = tc::idesc_tf32(128, NT, 0, 0);
    auto load_b = [&](int s) {
      unsigned char* base = smem + (s % kRing) * S::stage_bytes + 2 * S::a_bytes;
      mbar_expect_tx(full + (s % kRing), 2u * S::b_bytes);
      bulk_copy_g2s(base, wh + (size_t)s * 2 * 9 * NT, S::b_bytes, full + (s % kRing));
      bulk_copy_g2s(base + S::b_bytes, wl + (size_t)s * 2 * 9 * NT, S::b_bytes, full + (s % kRing));
    };
    if (tc::elect_one()) {              // (elect.sync inside the branch: ptxas keeps the MMA operands uniform)
      load_b(0);
      if (steps > 1) load_b(1);
    }
    __syncwarp();
    for (int s = 0; s < steps; ++s) {
      const int buf = s % kRing;
      __syncthreads();                  // A of step s is in shared memory (and flushed TMEM reads are done)
      if (tc::elect_one()) {
        tc::mbar_wait(full + buf, (uint32_t)((s / kRing) & 1));
        tc::fence_after_sync();
        unsigned char* base = smem + buf * S::stage_bytes;
        const uint32_t sa_hi = tc::smem_u32(base), sa_lo = sa_hi + S::a_bytes;
        const uint32_t sb_hi = sa_hi + 2 * S::a_bytes, sb_lo = sb_hi + S::b_bytes;
        const bool fresh = (s % kFlush) == 0;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const uint32_t b_off = (uint32_t)(tap * NT * 16);
          const uint64_t db_hi = tc::smem_desc(sb_hi + b_off, 9 * NT * 16, 128);
          const uint64_t db_lo = tc::smem_desc(sb_lo + b_off, 9 * NT * 16, 128);
          const uint32_t accum = (fresh && tap == 0) ? 0u : 1u;
#pragma unroll
          for (int t = 0; t < MT; ++t) {
            const uint32_t a_off = (uint32_t)((tap / 3) * S::row_pitch + (tap % 3 + kTX * t) * 16);
            const uint64_t da_hi = tc::smem_desc(sa_hi + a_off, S::quad_bytes, S::row_pitch);
            const uint64_t da_lo = tc::smem_desc(sa_lo + a_off, S::quad_bytes, S::row_pitch);
            const uint32_t d = tm + (uint32_t)(t * 3 * NT);
            tc::mma_ss(d, da_hi, db_hi, idesc, accum);
            tc::mma_ss(d + NT, da_lo, db_hi, idesc, accum);
            tc::mma_ss(d + 2 * NT, da_hi, db_lo, idesc, accum);
          }
        }
        tc::mma_commit(done + buf);
        if (s + 2 < steps) {            // refill the ring two steps ahead: that slot was read by step s - 1
          if (s >= 1) tc::mbar_wait(done + (s - 1) % kRing, (uint32_t)(((s - 1) / kRing) & 1));
          load_b(s + 2);
        }
      }
      __syncwarp();
    }
    __syncthreads();                    // epilogue done
    tc::tmem_dealloc(tm, S::tmem_cols);
    return;
  }

  // ================================== stager warps: A operand + epilogue ===================================
  // this thread's elements of a step: (channel within the step, window pixel) -> global / shared offsets
  int goff[kPer];         // pixel offset inside a channel plane, -1: outside the image (zero padding) / no element
  int soff[kPer];         // float offset inside the stage's A part | channel within the step << 24
#pragma unroll
  for (int k = 0; k < kPer; ++k) {
    const int idx = tid + k * kStagers;
    goff[k] = -1;
    soff[k] = -1;
    if (idx < kCs * S::win) {
      const int cl = idx & 3, rest = idx >> 2;
      const int cq = rest / S::win, pix = rest - cq * S::win;
      const int iy = pix / S::win_x, ix = pix - iy * S::win_x;
      const int gy = ty0 + iy - 1, gx = tx0 + ix - 1;
      if (gy >= 0 && gy < h && gx >= 0 && gx < w) goff[k] = gy * w + gx;
      soff[k] = ((cq * S::win + pix) * 4 + cl) | ((cq * 4 + cl) << 24);
    }
  }
  float pre[kPer];
  auto prefetch = [&](int s) {
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      pre[k] = 0.f;
      if (goff[k] >= 0) {
        const int ch = s * kCs + (soff[k] >> 24);
        const float* src = ch < c0 ? in0 + ((size_t)n * c0 + ch) * plane : in1 + ((size_t)n * c1 + (ch - c0)) * plane;
        pre[k] = __ldg(src + goff[k]);
      }
    }
  };
  const uint32_t trow = tm + ((uint32_t)(quad * 32) << 16);
  float acc[MT][kHalf];
#pragma unroll
  for (int t = 0; t < MT; ++t)
#pragma unroll
    for (int j = 0; j < kHalf; ++j) acc[t][j] = 0.f;

  prefetch(0);
  for (int s = 0; s < steps; ++s) {
    const int buf = s % kRing;
    // the slot was last read by the MMAs of step s - 3, which the issuer saw complete before the barrier of
    // step s - 1 (it waits for step s - 3 before refilling the ring in step s - 2)
    float* a_hi = reinterpret_cast<float*>(smem + buf * S::stage_bytes);
    float* a_lo = reinterpret_cast<float*>(smem + buf * S::stage_bytes + S::a_bytes);
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      if (soff[k] >= 0) {
        const float v = pre[k];
        const float hi = tc::tf32_hi(v);
        const int o = soff[k] & 0xFFFFFF;
        a_hi[o] = hi;
        a_lo[o] = v - hi;
      }
    }
    tc::fence_proxy_async();
    if (s + 1 < steps) prefetch(s + 1);          // in flight behind this step's MMAs
    tc::fence_before_sync();
    __syncthreads();
    if ((s + 1) % kFlush == 0 || s == steps - 1) {
      // fold the accumulators into registers (MMAs complete in order: this step done => all done)
      tc::mbar_wait(done + buf, (uint32_t)((s / kRing) & 1));
      tc::fence_after_sync();
#pragma unroll
      for (int t = 0; t < MT; ++t) {
#pragma unroll
        for (int j0 = 0; j0 < kHalf; j0 += 16) {
          uint32_t v0[16], v1[16], v2[16];
          const uint32_t col = (uint32_t)(t * 3 * NT + cs * kHalf + j0);
          tc::tmem_ld16(trow + col, v0);
          tc::tmem_ld16(trow + col + NT, v1);
          tc::tmem_ld16(trow + col + 2 * NT, v2);
          tc::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j)
            acc[t][j0 + j] += (__uint_as_float(v0[j]) + __uint_as_float(v1[j])) + __uint_as_float(v2[j]);
        }
      }
      tc::fence_before_sync();          // ordered before the next step's barrier, after which MMAs overwrite
    }
  }

  // ---- raw output + instance statistics (per (image, channel) sum / sum of squares, fp64 atomics) ---------
  const int r = quad * 32 + lane;                   // TMEM lane = pixel of a tile
  const int gy = ty0 + (r >> 3);
#pragma unroll
  for (int j = 0; j < kHalf; ++j) {
    const int col = cs * kHalf + j;
    const int co = co_base + col;
    float sum = 0.f, sq = 0.f;
#pragma unroll
    for (int t = 0; t < MT; ++t) {
      const int gx = tx0 + kTX * t + (r & 7);
      if (gy < h && gx < w) {
        const float v = acc[t][j];
        out[((size_t)n * cout + co) * plane + (size_t)gy * w + gx] = v;
        sum += v;
        sq = fmaf(v, v, sq);
      }
    }
    sum = warp_sum(sum);
    sq = warp_sum(sq);
    if (lane == 0) {
      st[(quad * NT + col) * 2 + 0] = sum;
      st[(quad * NT + col) * 2 + 1] = sq;
    }
  }
  // named barrier over the 256 stager threads only (the issuer warp is waiting at the final __syncthreads)
  asm volatile("bar.sync 1, 256;" ::: "memory");
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
  tc::fence_before_sync();
  __syncthreads();                      // releases the issuer warp, which frees the tensor memory
}

// weight (cout, cin, 3, 3) -> packed [cout / nt][cin / 4][9][nt] float4 (4 consecutive input channels), split into
// tf32 hi / lo parts
__global__ void __launch_bounds__(256)
pack_conv3x3_kernel(const float* __restrict__ weight, float4* __restrict__ w_hi, float4* __restrict__ w_lo, int cout,
                    int cin, int nt) {
  const int quads = cin / 4;
  const int total = quads * 9 * cout;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int nn = idx % nt;
    int rest = idx / nt;
    const int tap = rest % 9;
    rest /= 9;
    const int q = rest % quads, ct = rest / quads;
    const int co = ct * nt + nn;
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

// output channels per CTA: 64 when cout allows it, else 32 (both the packing and the kernel derive it from cout)
int nt_of(int cout) { return (cout % 64 == 0) ? 64 : 32; }

template <int NT, int MT>
int launch_conv(const float* in0, int c0, const float* in1, int c1, const float* w_hi, const float* w_lo, float* out,
                double* stats, int n, int cout, int h, int w, cudaStream_t s) {
  constexpr int smem = ConvCfg<NT, MT>::total;
  static DeviceOnce once;
  if (once.first()) cudaFuncSetAttribute(conv3x3_tc_kernel<NT, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int tiles_x = (w + kTX * MT - 1) / (kTX * MT), tiles_y = (h + kTY - 1) / kTY;
  if (n > 65535 || cout / NT > 65535) return IMMOCO_ERR_UNSUPPORTED;
  conv3x3_tc_kernel<NT, MT><<<dim3(tiles_x * tiles_y, cout / NT, n), kThreads, smem, s>>>(
      in0, c0, in1, c1, (const float4*)w_hi, (const float4*)w_lo, out, stats, cout, h, w, tiles_x);
  IMMOCO_LAUNCH_CHECK();
  return 0;
}

}  // namespace

extern "C" int immoco_unet_pack_conv3x3(const float* weight, float* w_hi, float* w_lo, int32_t cout, int32_t cin,
                                        void* stream) {
  if (!weight || !w_hi || !w_lo || cout < 1 || cin < 4 || (cin & 3) != 0 || (cout % 32) != 0) return IMMOCO_ERR_BAD_ARG;
  if ((((uintptr_t)w_hi | (uintptr_t)w_lo) & 15) != 0) return IMMOCO_ERR_BAD_ARG;
  const int total = (cin / 4) * 9 * cout;
  pack_conv3x3_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(weight, (float4*)w_hi, (float4*)w_lo, cout, cin,
                                                                          nt_of(cout));
  IMMOCO_LAUNCH_CHECK();
  return 0;
}

// in1 may be NULL (c1 = 0).  (c0 + c1) % 8 == 0, c0 % 4 == 0 and cout % 32 == 0, otherwise IMMOCO_ERR_UNSUPPORTED
// (the caller keeps immoco_unet_conv3x3 for those).  stats: (n * cout * 2) doubles, ZEROED by the caller.
extern "C" int immoco_unet_conv3x3_tc(const float* in0, int32_t c0, const float* in1, int32_t c1, const float* w_hi,
                                      const float* w_lo, float* out, double* stats, int32_t n, int32_t cout, int32_t h,
                                      int32_t w, void* stream) {
  if (!in0 || !w_hi || !w_lo || !out || !stats || c0 < 1 || c1 < 0 || (c1 > 0 && !in1) || n < 0 || cout < 1 || h < 1 || w < 1)
    return IMMOCO_ERR_BAD_ARG;
  if (((c0 + c1) % kCs) != 0 || (cout % 32) != 0 || (c0 & 3) != 0) return IMMOCO_ERR_UNSUPPORTED;
  if ((((uintptr_t)w_hi | (uintptr_t)w_lo) & 15) != 0) return IMMOCO_ERR_BAD_ARG;
  if (n == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (nt_of(cout) == 64) return launch_conv<64, 2>(in0, c0, in1, c1, w_hi, w_lo, out, stats, n, cout, h, w, s);
  return launch_conv<32, 4>(in0, c0, in1, c1, w_hi, w_lo, out, stats, n, cout, h, w, s);
}
