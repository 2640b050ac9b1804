// Rigid-motion k-space corruption on the device: the synthetic-input generator of every config
// (src/utils/motion_utils.py:121-202 motion_simulation2D).  Random draws stay on the host (they
// consume the torch RNG in the reference's order); the image work runs here:
//   moved[m] = grid_sample(image, affine_grid(theta_m, align_corners=True), bilinear, border,
//              align_corners=False)                                   (motion_utils.py:165-186)
//   k[:, w0_m:w1_m] = FFT(moved[m])[:, w0_m:w1_m]                      (motion_utils.py:188-196)
#include "common.cuh"

namespace {

// theta: (M, 6) row-major [t00 t01 t02; t10 t11 t12]; out (M, H, W) complex
__global__ void __launch_bounds__(256)
rigid_resample_kernel(const float2* __restrict__ image, const float* __restrict__ theta,
                      float2* __restrict__ out, int h, int w) {
  const int m = blockIdx.y;
  const float t00 = theta[6 * m + 0], t01 = theta[6 * m + 1], t02 = theta[6 * m + 2];
  const float t10 = theta[6 * m + 3], t11 = theta[6 * m + 4], t12 = theta[6 * m + 5];
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < h * w; idx += gridDim.x * blockDim.x) {
    const int i = idx / w, j = idx - i * w;
    // affine_grid base coordinates, align_corners=True: linspace(-1, 1, n)
    const float x = w > 1 ? (2.0f * j) / (float)(w - 1) - 1.0f : 0.0f;
    const float y = h > 1 ? (2.0f * i) / (float)(h - 1) - 1.0f : 0.0f;
    const float gx = x * t00 + y * t01 + t02;
    const float gy = x * t10 + y * t11 + t12;
    // grid_sample un-normalisation, align_corners=False; padding_mode="border": clip to the image
    float ix = ((gx + 1.0f) * (float)w - 1.0f) * 0.5f;
    float iy = ((gy + 1.0f) * (float)h - 1.0f) * 0.5f;
    ix = fminf(fmaxf(ix, 0.0f), (float)(w - 1));
    iy = fminf(fmaxf(iy, 0.0f), (float)(h - 1));
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    const int x0 = (int)fx0, y0 = (int)fy0;
    const float ax = ix - fx0, ay = iy - fy0;
    const float wgt[4] = {(1.f - ax) * (1.f - ay), ax * (1.f - ay), (1.f - ax) * ay, ax * ay};
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int xx = x0 + (t & 1), yy = y0 + (t >> 1);
      if (xx < w && yy < h) {
        const float2 v = __ldg(image + (size_t)yy * w + xx);
        acc.x = fmaf(wgt[t], v.x, acc.x);
        acc.y = fmaf(wgt[t], v.y, acc.y);
      }
    }
    out[(size_t)m * h * w + idx] = acc;
  }
}

// k[:, w0[m]:w1[m]] = k_moved[m][:, w0[m]:w1[m]]; mask[:, w0:w1] = 1 (int64 like the reference)
__global__ void __launch_bounds__(256)
replace_lines_kernel(float2* __restrict__ k, const float2* __restrict__ k_moved, long long* __restrict__ mask,
                     const int* __restrict__ w0, const int* __restrict__ w1, int h, int w) {
  const int m = blockIdx.y;
  const int a = max(0, w0[m]), b = min(w, w1[m]);
  const int nw = b - a;
  if (nw <= 0) return;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < h * nw; idx += gridDim.x * blockDim.x) {
    const int i = idx / nw, j = a + idx - i * nw;
    k[(size_t)i * w + j] = k_moved[((size_t)m * h + i) * w + j];
    if (mask) mask[(size_t)i * w + j] = 1;
  }
}

}  // namespace

extern "C" int immoco_rigid_resample(const float* image, const float* theta, float* out, int32_t n_mov,
                                     int32_t h, int32_t w, void* stream) {
  if (!image || !theta || !out || n_mov < 0 || h < 1 || w < 1) return IMMOCO_ERR_BAD_ARG;
  if (n_mov == 0) return 0;
  dim3 grid((unsigned)((h * w + 255) / 256), (unsigned)n_mov);
  rigid_resample_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float2*)image, theta, (float2*)out, h, w);
  IMMOCO_LAUNCH_CHECK();
  return 0;
}

// Movements are applied in order (a later window overwrites an earlier one, as the reference's loop does):
// one launch per movement keeps that order without atomics; M is small (<= 19 in the reference data).
extern "C" int immoco_replace_lines(float* k, const float* k_moved, int64_t* mask, const int32_t* w0,
                                    const int32_t* w1, int32_t n_mov, int32_t h, int32_t w, void* stream) {
  if (!k || !k_moved || !w0 || !w1 || n_mov < 0 || h < 1 || w < 1) return IMMOCO_ERR_BAD_ARG;
  for (int m = 0; m < n_mov; ++m) {
    dim3 grid((unsigned)((h * 9 + 255) / 256), 1);
    replace_lines_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        (float2*)k, (const float2*)k_moved + (size_t)m * h * w, (long long*)mask, w0 + m, w1 + m, h, w);
    IMMOCO_LAUNCH_CHECK();
  }
  return 0;
}
