// Rigid bicubic resampling for the autofocusing baseline (src/models/autofocusing.py:8-91): M images,
// each warped by its own affine theta_m exactly as
//   F.grid_sample(img_m, F.affine_grid(theta_m, align_corners=True), mode="bicubic", align_corners=False)
// (zeros padding, cubic-convolution A = -0.75), and the backward pass reduced straight to d theta (the
// images are constants of the optimisation: autofocusing.py:29 builds them from the measured k-space).
#include "common.cuh"

namespace {

constexpr float kA = -0.75f;

__device__ __forceinline__ float cc1(float x) { return ((kA + 2.f) * x - (kA + 3.f)) * x * x + 1.f; }
__device__ __forceinline__ float cc2(float x) { return ((kA * x - 5.f * kA) * x + 8.f * kA) * x - 4.f * kA; }
__device__ __forceinline__ float dcc1(float x) { return (3.f * (kA + 2.f) * x - 2.f * (kA + 3.f)) * x; }
__device__ __forceinline__ float dcc2(float x) { return (3.f * kA * x - 10.f * kA) * x + 8.f * kA; }

__device__ __forceinline__ void cubic_w(float t, float (&w)[4]) {
  w[0] = cc2(t + 1.f); w[1] = cc1(t); w[2] = cc1(1.f - t); w[3] = cc2(2.f - t);
}
__device__ __forceinline__ void cubic_dw(float t, float (&d)[4]) {
  d[0] = dcc2(t + 1.f); d[1] = dcc1(t); d[2] = -dcc1(1.f - t); d[3] = -dcc2(2.f - t);
}

struct Sample { float x, y, ix, iy; };

__device__ __forceinline__ Sample sample_pos(const float* __restrict__ th, int i, int j, int h, int w) {
  Sample s;
  s.x = w > 1 ? (2.0f * j) / (float)(w - 1) - 1.0f : 0.0f;     // affine_grid base, align_corners=True
  s.y = h > 1 ? (2.0f * i) / (float)(h - 1) - 1.0f : 0.0f;
  const float gx = s.x * th[0] + s.y * th[1] + th[2];
  const float gy = s.x * th[3] + s.y * th[4] + th[5];
  s.ix = ((gx + 1.0f) * (float)w - 1.0f) * 0.5f;               // grid_sample, align_corners=False
  s.iy = ((gy + 1.0f) * (float)h - 1.0f) * 0.5f;
  return s;
}

// images (M, H, W) complex, theta (M, 6), out (M, H, W) complex
__global__ void __launch_bounds__(256)
bicubic_fwd_kernel(const float2* __restrict__ images, const float* __restrict__ theta, float2* __restrict__ out,
                   int h, int w) {
  const int m = blockIdx.y;
  const float* th = theta + 6 * m;
  const float2* img = images + (size_t)m * h * w;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < h * w; idx += gridDim.x * blockDim.x) {
    const int i = idx / w, j = idx - i * w;
    const Sample s = sample_pos(th, i, j, h, w);
    const float fx = floorf(s.ix), fy = floorf(s.iy);
    float wx[4], wy[4];
    cubic_w(s.ix - fx, wx);
    cubic_w(s.iy - fy, wy);
    const int x0 = (int)fx - 1, y0 = (int)fy - 1;
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int yy = y0 + a;
      if (yy < 0 || yy >= h) continue;
      float2 row = make_float2(0.f, 0.f);
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int xx = x0 + b;
        if (xx >= 0 && xx < w) {
          const float2 v = __ldg(img + (size_t)yy * w + xx);
          row.x = fmaf(wx[b], v.x, row.x);
          row.y = fmaf(wx[b], v.y, row.y);
        }
      }
      acc.x = fmaf(wy[a], row.x, acc.x);
      acc.y = fmaf(wy[a], row.y, acc.y);
    }
    out[(size_t)m * h * w + idx] = acc;
  }
}

// d_theta (M, 6) doubles (ZEROED by the caller) += sum over pixels of d_out . d sample / d theta
__global__ void __launch_bounds__(256)
bicubic_bwd_theta_kernel(const float2* __restrict__ images, const float* __restrict__ theta,
                         const float2* __restrict__ d_out, double* __restrict__ d_theta, int h, int w) {
  const int m = blockIdx.y;
  const float* th = theta + 6 * m;
  const float2* img = images + (size_t)m * h * w;
  float g[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < h * w; idx += gridDim.x * blockDim.x) {
    const int i = idx / w, j = idx - i * w;
    const Sample s = sample_pos(th, i, j, h, w);
    const float fx = floorf(s.ix), fy = floorf(s.iy);
    float wx[4], wy[4], dx[4], dy[4];
    cubic_w(s.ix - fx, wx);
    cubic_w(s.iy - fy, wy);
    cubic_dw(s.ix - fx, dx);
    cubic_dw(s.iy - fy, dy);
    const int x0 = (int)fx - 1, y0 = (int)fy - 1;
    const float2 go = __ldg(d_out + (size_t)m * h * w + idx);
    float gix = 0.f, giy = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int yy = y0 + a;
      if (yy < 0 || yy >= h) continue;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int xx = x0 + b;
        if (xx >= 0 && xx < w) {
          const float2 v = __ldg(img + (size_t)yy * w + xx);
          const float dot = v.x * go.x + v.y * go.y;
          gix = fmaf(dot, dx[b] * wy[a], gix);
          giy = fmaf(dot, wx[b] * dy[a], giy);
        }
      }
    }
    const float ggx = gix * (0.5f * (float)w), ggy = giy * (0.5f * (float)h);
    g[0] = fmaf(ggx, s.x, g[0]); g[1] = fmaf(ggx, s.y, g[1]); g[2] += ggx;
    g[3] = fmaf(ggy, s.x, g[3]); g[4] = fmaf(ggy, s.y, g[4]); g[5] += ggy;
  }
  __shared__ float red[6][8];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const float v = warp_sum(g[k]);
    if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    double acc = 0.0;
    for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) acc += (double)red[threadIdx.x][wv];
    atomicAdd(d_theta + 6 * m + threadIdx.x, acc);
  }
}

}  // namespace

extern "C" int immoco_rigid_bicubic_fwd(const float* images, const float* theta, float* out, int32_t n_mov,
                                        int32_t h, int32_t w, void* stream) {
  if (!images || !theta || !out || n_mov < 0 || h < 1 || w < 1) return IMMOCO_ERR_BAD_ARG;
  if (n_mov == 0) return 0;
  dim3 grid((unsigned)((h * w + 255) / 256), (unsigned)n_mov);
  bicubic_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float2*)images, theta, (float2*)out, h, w);
  IMMOCO_LAUNCH_CHECK();
  return 0;
}

extern "C" int immoco_rigid_bicubic_bwd_theta(const float* images, const float* theta, const float* d_out,
                                              double* d_theta, int32_t n_mov, int32_t h, int32_t w, void* stream) {
  if (!images || !theta || !d_out || !d_theta || n_mov < 0 || h < 1 || w < 1) return IMMOCO_ERR_BAD_ARG;
  if (n_mov == 0) return 0;
  int gx = (h * w + 255) / 256;
  if (gx > 148) gx = 148;
  bicubic_bwd_theta_kernel<<<dim3((unsigned)gx, (unsigned)n_mov), 256, 0, (cudaStream_t)stream>>>(
      (const float2*)images, theta, (const float2*)d_out, d_theta, h, w);
  IMMOCO_LAUNCH_CHECK();
  return 0;
}
