// kld-net inference kernels: the line-detection U-Net that produces the movement-group masks the
// IM-MoCo fit consumes (src/models/kld_net.py:4-11 -> fastmri Unet == src/models/unet.py:17-187 with
// InstanceNorm2d; caller src/test/test_immoco.py:50-58).  NCHW fp32, inference only.
//
//   conv3x3_kernel       3x3 / pad 1 / no bias over one or two (channel-concatenated) inputs; writes the
//                        RAW output and accumulates per-(image, channel) sum / sum-of-squares (fp64)
//   convt2x2_kernel      ConvTranspose2d(k=2, s=2, no bias): a 1x1 contraction to 4*Cout values per
//                        input pixel scattered to the 2x2 output block; same statistics
//   instnorm_lrelu_kernel  InstanceNorm2d(eps 1e-5, biased variance) + LeakyReLU(0.2) in place, optionally
//                        also writing the 2x2 average-pooled tensor the next level reads
//   conv1x1_bias_kernel  final Conv2d(chans -> out_chans, k=1) with bias
//
// fp32 SIMT with register tiles (4 pixels x 8/16 output channels per thread, operands staged in shared
// memory): the network is 37.7 GFLOP per 320x320 slice against ~860 ms of fitting, so parity with the
// fp32 reference was put before tensor cores here.
#include "common.cuh"

namespace {

constexpr int kTile = 16;          // output pixels per tile edge
constexpr int kCk = 8;             // input channels per shared-memory stage
constexpr int kThreads = 256;
constexpr int kInW = kTile + 2;    // input tile edge incl. halo
constexpr int kInStride = 20;      // padded row stride of the input tile (floats)

// CPT = output channels per thread; the CTA covers TCO = 4 * CPT output channels of a 16x16 tile.
template <int CPT>
__global__ void __launch_bounds__(kThreads)
conv3x3_kernel(const float* __restrict__ in0, int c0, const float* __restrict__ in1, int c1,
               const float* __restrict__ weight, float* __restrict__ out, double* __restrict__ stats,
               int cout, int h, int w, int co_groups) {
  constexpr int TCO = 4 * CPT;
  __shared__ __align__(16) float s_in[kCk][kInW][kInStride];
  __shared__ __align__(16) float s_w[kCk][9][TCO];
  const int cin = c0 + c1;
  const int n = blockIdx.z / co_groups, cg = blockIdx.z - n * co_groups;
  const int co_base = cg * TCO;
  const int ty0 = blockIdx.y * kTile, tx0 = blockIdx.x * kTile;
  const int tid = threadIdx.x;
  const int pg = tid & 63, og = tid >> 6;          // pixel group (16 rows x 4 groups of 4), cout group
  const int py = pg >> 2, px = (pg & 3) * 4;

  float acc[4][CPT];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int o = 0; o < CPT; ++o) acc[p][o] = 0.f;

  for (int cb = 0; cb < cin; cb += kCk) {
    // ---- stage inputs (zero padding at the image border and beyond the channel count) -----------
    for (int idx = tid; idx < kCk * kInW * kInW; idx += kThreads) {
      const int c = idx / (kInW * kInW), r = idx - c * (kInW * kInW);
      const int iy = r / kInW, ix = r - iy * kInW;
      const int gy = ty0 + iy - 1, gx = tx0 + ix - 1, ch = cb + c;
      float v = 0.f;
      if (ch < cin && gy >= 0 && gy < h && gx >= 0 && gx < w) {
        const float* src = ch < c0 ? in0 + ((size_t)n * c0 + ch) * h * w : in1 + ((size_t)n * c1 + (ch - c0)) * h * w;
        v = __ldg(src + (size_t)gy * w + gx);
      }
      s_in[c][iy][ix] = v;
    }
    // ---- stage weights: s_w[c][k][o] = W[co_base + o][cb + c][k] -----------------------------------
    for (int idx = tid; idx < kCk * 9 * TCO; idx += kThreads) {
      const int o = idx / (kCk * 9), r = idx - o * (kCk * 9);
      const int c = r / 9, k = r - c * 9;
      const int co = co_base + o, ch = cb + c;
      s_w[c][k][o] = (co < cout && ch < cin) ? __ldg(weight + ((size_t)co * cin + ch) * 9 + k) : 0.f;
    }
    __syncthreads();
#pragma unroll 2
    for (int c = 0; c < kCk; ++c) {
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        float v[6];
        const float4 a = *reinterpret_cast<const float4*>(&s_in[c][py + ky][px]);
        const float2 b = *reinterpret_cast<const float2*>(&s_in[c][py + ky][px + 4]);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          float wv[CPT];
#pragma unroll
          for (int o4 = 0; o4 < CPT; o4 += 4) {
            const float4 t = *reinterpret_cast<const float4*>(&s_w[c][ky * 3 + kx][og * CPT + o4]);
            wv[o4] = t.x; wv[o4 + 1] = t.y; wv[o4 + 2] = t.z; wv[o4 + 3] = t.w;
          }
#pragma unroll
          for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int o = 0; o < CPT; ++o) acc[p][o] = fmaf(v[p + kx], wv[o], acc[p][o]);
        }
      }
    }
    __syncthreads();
  }
  // ---- raw output + instance statistics --------------------------------------------------------------
  const int gy = ty0 + py;
#pragma unroll
  for (int o = 0; o < CPT; ++o) {
    const int co = co_base + og * CPT + o;
    float s = 0.f, q = 0.f;
    if (co < cout && gy < h) {
      float* dst = out + (((size_t)n * cout + co) * h + gy) * w + tx0 + px;
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        if (tx0 + px + p < w) {
          dst[p] = acc[p][o];
          s += acc[p][o];
          q = fmaf(acc[p][o], acc[p][o], q);
        }
      }
    }
    // the 64 threads of a cout group are two full warps: reduce in the warp, one fp64 atomic per warp
    s = warp_sum(s);
    q = warp_sum(q);
    if ((tid & 31) == 0 && co < cout) {
      atomicAdd(stats + ((size_t)n * cout + co) * 2 + 0, (double)s);
      atomicAdd(stats + ((size_t)n * cout + co) * 2 + 1, (double)q);
    }
  }
}

// ConvTranspose2d(k=2, s=2, no bias): weight (Cin, Cout, 2, 2); out (N, Cout, 2H, 2W) raw + statistics.
// Thread = one input pixel x 16 "virtual" outputs v = co*4 + a*2 + b (contiguous in the weight tensor).
__global__ void __launch_bounds__(kThreads)
convt2x2_kernel(const float* __restrict__ in, const float* __restrict__ weight, float* __restrict__ out,
                double* __restrict__ stats, int cin, int cout, int h, int w, int v_groups) {
  constexpr int VPT = 16, kCi = 32;
  __shared__ __align__(16) float s_w[kCi][VPT];
  const int n = blockIdx.z / v_groups, vg = blockIdx.z - n * v_groups;
  const int v0 = vg * VPT;
  const int pix = blockIdx.x * kThreads + threadIdx.x;
  const bool live = pix < h * w;
  float acc[VPT];
#pragma unroll
  for (int v = 0; v < VPT; ++v) acc[v] = 0.f;
  for (int cb = 0; cb < cin; cb += kCi) {
    for (int idx = threadIdx.x; idx < kCi * VPT; idx += kThreads) {
      const int c = idx / VPT, v = idx - c * VPT;
      s_w[c][v] = (cb + c < cin && v0 + v < cout * 4) ? __ldg(weight + (size_t)(cb + c) * cout * 4 + v0 + v) : 0.f;
    }
    __syncthreads();
    if (live) {
      for (int c = 0; c < kCi && cb + c < cin; ++c) {
        const float x = __ldg(in + ((size_t)n * cin + cb + c) * h * w + pix);
#pragma unroll
        for (int v4 = 0; v4 < VPT; v4 += 4) {
          const float4 t = *reinterpret_cast<const float4*>(&s_w[c][v4]);
          acc[v4] = fmaf(x, t.x, acc[v4]); acc[v4 + 1] = fmaf(x, t.y, acc[v4 + 1]);
          acc[v4 + 2] = fmaf(x, t.z, acc[v4 + 2]); acc[v4 + 3] = fmaf(x, t.w, acc[v4 + 3]);
        }
      }
    }
    __syncthreads();
  }
  const int i = live ? pix / w : 0, j = live ? pix - i * w : 0;
#pragma unroll
  for (int cq = 0; cq < VPT / 4; ++cq) {
    const int co = (v0 >> 2) + cq;
    float s = 0.f, q = 0.f;
    if (live && co < cout) {
      float* dst = out + (((size_t)n * cout + co) * (2 * h) + 2 * i) * (2 * w) + 2 * j;
      const float a = acc[4 * cq], b = acc[4 * cq + 1], c = acc[4 * cq + 2], d = acc[4 * cq + 3];
      *reinterpret_cast<float2*>(dst) = make_float2(a, b);
      *reinterpret_cast<float2*>(dst + 2 * w) = make_float2(c, d);
      s = (a + b) + (c + d);
      q = fmaf(a, a, fmaf(b, b, fmaf(c, c, d * d)));
    }
    s = warp_sum(s);
    q = warp_sum(q);
    if ((threadIdx.x & 31) == 0 && co < cout) {
      atomicAdd(stats + ((size_t)n * cout + co) * 2 + 0, (double)s);
      atomicAdd(stats + ((size_t)n * cout + co) * 2 + 1, (double)q);
    }
  }
}

// x <- LeakyReLU_0.2((x - mean) * rstd) per (image, channel) plane; pooled (may be null) <- 2x2 average.
// grid (chunks, N*C); each thread handles 2x2 blocks so the pooled value needs no second pass.
__global__ void __launch_bounds__(kThreads)
instnorm_lrelu_kernel(float* __restrict__ x, const double* __restrict__ stats, float* __restrict__ pooled,
                      int h, int w, float eps, float slope) {
  const int plane = blockIdx.y;
  const double cnt = (double)h * (double)w;
  const double mean_d = stats[2 * plane] / cnt;
  const double var_d = fmax(stats[2 * plane + 1] / cnt - mean_d * mean_d, 0.0);
  const float mean = (float)mean_d, rstd = (float)(1.0 / sqrt(var_d + (double)eps));
  float* xp = x + (size_t)plane * h * w;
  const int hb = (h + 1) / 2, wb = (w + 1) / 2;
  for (int idx = blockIdx.x * kThreads + threadIdx.x; idx < hb * wb; idx += gridDim.x * kThreads) {
    const int bi = idx / wb, bj = idx - bi * wb;
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int i = 2 * bi + (t >> 1), j = 2 * bj + (t & 1);
      if (i < h && j < w) {
        float v = (xp[(size_t)i * w + j] - mean) * rstd;
        v = v >= 0.f ? v : v * slope;
        xp[(size_t)i * w + j] = v;
        sum += v;
      }
    }
    if (pooled && 2 * bi + 1 < h && 2 * bj + 1 < w)
      pooled[((size_t)plane * (h / 2) + bi) * (w / 2) + bj] = sum * 0.25f;
  }
}

__global__ void __launch_bounds__(kThreads)
conv1x1_bias_kernel(const float* __restrict__ in, const float* __restrict__ weight, const float* __restrict__ bias,
                    float* __restrict__ out, int cin, int cout, int hw) {
  const int n = blockIdx.z, co = blockIdx.y;
  for (int p = blockIdx.x * kThreads + threadIdx.x; p < hw; p += gridDim.x * kThreads) {
    float acc = bias ? __ldg(bias + co) : 0.f;
    for (int c = 0; c < cin; ++c) acc = fmaf(__ldg(in + ((size_t)n * cin + c) * hw + p), __ldg(weight + (size_t)co * cin + c), acc);
    out[((size_t)n * cout + co) * hw + p] = acc;
  }
}

}  // namespace

// in1 may be NULL (c1 = 0).  stats: (n * cout * 2) doubles, ZEROED by the caller.
extern "C" int immoco_unet_conv3x3(const float* in0, int32_t c0, const float* in1, int32_t c1, const float* weight,
                                   float* out, double* stats, int32_t n, int32_t cout, int32_t h, int32_t w,
                                   void* stream) {
  if (!in0 || !weight || !out || !stats || c0 < 1 || c1 < 0 || (c1 > 0 && !in1) || n < 0 || cout < 1 || h < 1 || w < 1)
    return IMMOCO_ERR_BAD_ARG;
  if (n == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const int tx = (w + kTile - 1) / kTile, ty = (h + kTile - 1) / kTile;
  if (cout % 64 == 0 || cout > 32) {
    const int groups = (cout + 63) / 64;
    if ((int64_t)n * groups > 65535) return IMMOCO_ERR_UNSUPPORTED;
    conv3x3_kernel<16><<<dim3(tx, ty, n * groups), kThreads, 0, s>>>(in0, c0, in1, c1, weight, out, stats, cout, h, w, groups);
  } else {
    const int groups = (cout + 31) / 32;
    if ((int64_t)n * groups > 65535) return IMMOCO_ERR_UNSUPPORTED;
    conv3x3_kernel<8><<<dim3(tx, ty, n * groups), kThreads, 0, s>>>(in0, c0, in1, c1, weight, out, stats, cout, h, w, groups);
  }
  IMMOCO_LAUNCH_CHECK();
  return 0;
}

extern "C" int immoco_unet_convt2x2(const float* in, const float* weight, float* out, double* stats, int32_t n,
                                    int32_t cin, int32_t cout, int32_t h, int32_t w, void* stream) {
  if (!in || !weight || !out || !stats || n < 0 || cin < 1 || cout < 1 || h < 1 || w < 1) return IMMOCO_ERR_BAD_ARG;
  if (n == 0) return 0;
  const int groups = (cout * 4 + 15) / 16;
  if ((int64_t)n * groups > 65535) return IMMOCO_ERR_UNSUPPORTED;
  convt2x2_kernel<<<dim3((h * w + kThreads - 1) / kThreads, 1, n * groups), kThreads, 0, (cudaStream_t)stream>>>(
      in, weight, out, stats, cin, cout, h, w, groups);
  IMMOCO_LAUNCH_CHECK();
  return 0;
}

extern "C" int immoco_unet_instnorm_lrelu(float* x, const double* stats, float* pooled, int32_t planes, int32_t h,
                                          int32_t w, float eps, float slope, void* stream) {
  if (!x || !stats || planes < 0 || h < 1 || w < 1) return IMMOCO_ERR_BAD_ARG;
  if (planes == 0) return 0;
  if (planes > 65535) return IMMOCO_ERR_UNSUPPORTED;
  const int blocks = ((h + 1) / 2) * ((w + 1) / 2);
  int gx = (blocks + kThreads - 1) / kThreads;
  if (gx > 64) gx = 64;
  instnorm_lrelu_kernel<<<dim3(gx, planes), kThreads, 0, (cudaStream_t)stream>>>(x, stats, pooled, h, w, eps, slope);
  IMMOCO_LAUNCH_CHECK();
  return 0;
}

extern "C" int immoco_unet_conv1x1(const float* in, const float* weight, const float* bias, float* out, int32_t n,
                                   int32_t cin, int32_t cout, int32_t hw, void* stream) {
  if (!in || !weight || !out || n < 0 || cin < 1 || cout < 1 || hw < 1) return IMMOCO_ERR_BAD_ARG;
  if (n == 0) return 0;
  if (n > 65535 || cout > 65535) return IMMOCO_ERR_UNSUPPORTED;
  int gx = (hw + kThreads - 1) / kThreads;
  if (gx > 1024) gx = 1024;
  conv1x1_bias_kernel<<<dim3(gx, cout, n), kThreads, 0, (cudaStream_t)stream>>>(in, weight, bias, out, cin, cout, hw);
  IMMOCO_LAUNCH_CHECK();
  return 0;
}
