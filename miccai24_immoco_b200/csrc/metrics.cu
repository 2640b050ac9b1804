// On-device image-quality metrics of the evaluation step that follows the fit
// (src/utils/evaluate.py:19-47 normalize / rmse / my_psnr, :57-80 calmetric2D with piq.ssim
// kernel_size=11, data_range=1; caller src/test/test_immoco.py:74-85 on central-half crops).
// One launch finds the per-image min / max, one launch does the rest: min-max normalisation on
// load, squared error, and SSIM with the separable 11-tap Gaussian window evaluated in shared memory.
#include "common.cuh"

namespace {

constexpr int kT = 32;            // outputs per tile edge
constexpr int kMaxTaps = 11;
constexpr int kThreads = 256;

struct View {                      // strided (batch, H, W) view; complex -> magnitude on load
  const float* p;
  int64_t img_stride, row_stride;  // in ELEMENTS (a complex element is 2 floats)
  int is_complex;
};

__device__ __forceinline__ float load_px(const View& v, int b, int i, int j) {
  const int64_t o = (int64_t)b * v.img_stride + (int64_t)i * v.row_stride + j;
  if (v.is_complex) {
    const float2 z = reinterpret_cast<const float2*>(v.p)[o];
    return sqrtf(z.x * z.x + z.y * z.y);
  }
  return v.p[o];
}
// value of the f x f average-pooled image at (i, j) (piq.ssim down-samples by f = max(1, round(min(H,W)/256)))
__device__ __forceinline__ float load_pooled(const View& v, int b, int i, int j, int f, float lo, float inv) {
  float s = 0.f;
  for (int di = 0; di < f; ++di)
    for (int dj = 0; dj < f; ++dj) s += (load_px(v, b, i * f + di, j * f + dj) - lo) * inv;
  return s / (float)(f * f);
}

// mm[b] = {min_pred, max_pred, min_gt, max_gt}
__global__ void __launch_bounds__(1024) minmax_kernel(View pred, View gt, float* __restrict__ mm, int h, int w) {
  const int b = blockIdx.x;
  float lo0 = INFINITY, hi0 = -INFINITY, lo1 = INFINITY, hi1 = -INFINITY;
  for (int idx = threadIdx.x; idx < h * w; idx += blockDim.x) {
    const int i = idx / w, j = idx - i * w;
    const float a = load_px(pred, b, i, j), c = load_px(gt, b, i, j);
    lo0 = fminf(lo0, a); hi0 = fmaxf(hi0, a);
    lo1 = fminf(lo1, c); hi1 = fmaxf(hi1, c);
  }
  __shared__ float red[4][32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo0 = fminf(lo0, __shfl_xor_sync(0xffffffffu, lo0, o));
    hi0 = fmaxf(hi0, __shfl_xor_sync(0xffffffffu, hi0, o));
    lo1 = fminf(lo1, __shfl_xor_sync(0xffffffffu, lo1, o));
    hi1 = fmaxf(hi1, __shfl_xor_sync(0xffffffffu, hi1, o));
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = lo0; red[1][warp] = hi0; red[2][warp] = lo1; red[3][warp] = hi1; }
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    lo0 = lane < nw ? red[0][lane] : INFINITY;
    hi0 = lane < nw ? red[1][lane] : -INFINITY;
    lo1 = lane < nw ? red[2][lane] : INFINITY;
    hi1 = lane < nw ? red[3][lane] : -INFINITY;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo0 = fminf(lo0, __shfl_xor_sync(0xffffffffu, lo0, o));
      hi0 = fmaxf(hi0, __shfl_xor_sync(0xffffffffu, hi0, o));
      lo1 = fminf(lo1, __shfl_xor_sync(0xffffffffu, lo1, o));
      hi1 = fmaxf(hi1, __shfl_xor_sync(0xffffffffu, hi1, o));
    }
    if (lane == 0) { mm[4 * b + 0] = lo0; mm[4 * b + 1] = hi0; mm[4 * b + 2] = lo1; mm[4 * b + 3] = hi1; }
  }
}

struct Taps { float g[kMaxTaps]; };

// acc[b] = {sum (x-y)^2 over the full-resolution image, sum of the SSIM map, map size, unused}
__global__ void __launch_bounds__(kThreads)
ssim_mse_kernel(View pred, View gt, const float* __restrict__ mm, double* __restrict__ acc, int h, int w,
                int f, int taps, Taps win) {
  const int b = blockIdx.z;
  const float lo_x = mm[4 * b + 0], inv_x = 1.0f / ((mm[4 * b + 1] - lo_x) + 1e-24f);
  const float lo_y = mm[4 * b + 2], inv_y = 1.0f / ((mm[4 * b + 3] - lo_y) + 1e-24f);
  const int hp = h / f, wp = w / f;                 // pooled size (avg_pool2d floors)
  const int oh = hp - taps + 1, ow = wp - taps + 1; // valid-convolution output size
  const int ti = blockIdx.y * kT, tj = blockIdx.x * kT;
  constexpr int kIn = kT + kMaxTaps - 1;
  __shared__ float sx[kIn][kIn + 1], sy[kIn][kIn + 1];
  __shared__ float hx[5][kIn][kT + 1];
  __shared__ double red[2][kThreads / 32];
  const int in_e = kT + taps - 1;

  // ---- squared error on the full-resolution normalised images: each pixel owned by one tile ------
  double se = 0.0;
  {
    const int i1 = (blockIdx.y == gridDim.y - 1) ? h : min(h, (ti + kT) * f);
    const int j1 = (blockIdx.x == gridDim.x - 1) ? w : min(w, (tj + kT) * f);
    const int i0 = ti * f, j0 = tj * f;
    const int nw = j1 - j0;
    for (int idx = threadIdx.x; idx < (i1 - i0) * nw; idx += kThreads) {
      const int i = i0 + idx / nw, j = j0 + idx % nw;
      const float x = (load_px(pred, b, i, j) - lo_x) * inv_x;
      const float y = (load_px(gt, b, i, j) - lo_y) * inv_y;
      const float d = x - y;
      se += (double)(d * d);
    }
  }
  // ---- SSIM tile ------------------------------------------------------------------------------------
  double ss = 0.0;
  if (oh > 0 && ow > 0 && ti < oh && tj < ow) {
    for (int idx = threadIdx.x; idx < in_e * in_e; idx += kThreads) {
      const int r = idx / in_e, c = idx - r * in_e;
      const int i = ti + r, j = tj + c;
      float x = 0.f, y = 0.f;
      if (i < hp && j < wp) {
        x = load_pooled(pred, b, i, j, f, lo_x, inv_x);
        y = load_pooled(gt, b, i, j, f, lo_y, inv_y);
      }
      sx[r][c] = x;
      sy[r][c] = y;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < in_e * kT; idx += kThreads) {      // horizontal pass
      const int r = idx / kT, c = idx - r * kT;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
      for (int t = 0; t < taps; ++t) {
        const float x = sx[r][c + t], y = sy[r][c + t], g = win.g[t];
        a0 = fmaf(g, x, a0); a1 = fmaf(g, y, a1);
        a2 = fmaf(g, x * x, a2); a3 = fmaf(g, y * y, a3); a4 = fmaf(g, x * y, a4);
      }
      hx[0][r][c] = a0; hx[1][r][c] = a1; hx[2][r][c] = a2; hx[3][r][c] = a3; hx[4][r][c] = a4;
    }
    __syncthreads();
    const float c1 = 0.01f * 0.01f, c2 = 0.03f * 0.03f;
    for (int idx = threadIdx.x; idx < kT * kT; idx += kThreads) {         // vertical pass + SSIM
      const int r = idx / kT, c = idx - r * kT;
      if (ti + r >= oh || tj + c >= ow) continue;
      float m[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
      for (int t = 0; t < taps; ++t) {
        const float g = win.g[t];
#pragma unroll
        for (int q = 0; q < 5; ++q) m[q] = fmaf(g, hx[q][r + t][c], m[q]);
      }
      const float mx = m[0], my = m[1];
      const float sxx = m[2] - mx * mx, syy = m[3] - my * my, sxy = m[4] - mx * my;
      const float cs = (2.f * sxy + c2) / (sxx + syy + c2);
      ss += (double)((2.f * mx * my + c1) / (mx * mx + my * my + c1) * cs);
    }
  }
  // ---- block reduction, one atomic per CTA and quantity --------------------------------------------
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    se += __shfl_xor_sync(0xffffffffu, se, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = se; red[1][warp] = ss; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, c = 0.0;
    for (int k = 0; k < kThreads / 32; ++k) { a += red[0][k]; c += red[1][k]; }
    atomicAdd(acc + 4 * b + 0, a);
    atomicAdd(acc + 4 * b + 1, c);
    if (blockIdx.x == 0 && blockIdx.y == 0) acc[4 * b + 2] = (oh > 0 && ow > 0) ? (double)oh * (double)ow : 0.0;
  }
}

// ---- HaarPSI (piq.haarpsi as called at src/utils/evaluate.py:76; restated in oracle/immoco_oracle.py:haarpsi01) ----
// pooled[b][which][i][j] = 2x2 average of 255 * min-max-normalised pixel (zero beyond the image: odd sizes are padded)
__global__ void __launch_bounds__(kThreads)
haar_pool_kernel(View pred, View gt, const float* __restrict__ mm, float* __restrict__ pooled, int h, int w, int hp,
                 int wp) {
  const int b = blockIdx.z, which = blockIdx.y;
  const View& v = which ? gt : pred;
  const float lo = mm[4 * b + 2 * which], inv = 255.0f / ((mm[4 * b + 2 * which + 1] - lo) + 1e-24f);
  float* out = pooled + ((size_t)(b * 2 + which) * hp) * wp;
  for (int idx = blockIdx.x * kThreads + threadIdx.x; idx < hp * wp; idx += gridDim.x * kThreads) {
    const int i = idx / wp, j = idx - i * wp;
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int y = 2 * i + (t >> 1), x = 2 * j + (t & 1);
      if (y < h && x < w) s += (load_px(v, b, y, x) - lo) * inv;
    }
    out[idx] = 0.25f * s;
  }
}

// Haar responses at kernel sizes 2 / 4 / 8 in two orientations ('same' zero padding: k/2 - 1 before, k/2 after),
// local similarity of the two finest scales, weight = larger magnitude at the coarsest; acc[b] += {sum sigmoid(alpha S) W, sum W}
__device__ __forceinline__ void haar_coeffs(const float* __restrict__ img, int hp, int wp, int i, int j, float (&co)[6]) {
#pragma unroll
  for (int sc = 0; sc < 3; ++sc) {
    const int k = 2 << sc, off = k / 2 - 1;
    float hsum = 0.f, vsum = 0.f;
    for (int a = 0; a < k; ++a) {
      const int y = i + a - off;
      if (y < 0 || y >= hp) continue;
      for (int c = 0; c < k; ++c) {
        const int x = j + c - off;
        if (x < 0 || x >= wp) continue;
        const float v = __ldg(img + (size_t)y * wp + x);
        hsum += (a < k / 2) ? v : -v;          // haar_filter: second half of the ROWS negative
        vsum += (c < k / 2) ? v : -v;          // its transpose: second half of the COLUMNS negative
      }
    }
    co[2 * sc] = hsum / (float)k;
    co[2 * sc + 1] = vsum / (float)k;
  }
}

__global__ void __launch_bounds__(kThreads)
haar_sim_kernel(const float* __restrict__ pooled, double* __restrict__ acc, int hp, int wp, float cst, float alpha) {
  const int b = blockIdx.z;
  const float* px = pooled + ((size_t)(b * 2) * hp) * wp;
  const float* py = px + (size_t)hp * wp;
  double num = 0.0, den = 0.0;
  for (int idx = blockIdx.x * kThreads + threadIdx.x; idx < hp * wp; idx += gridDim.x * kThreads) {
    const int i = idx / wp, j = idx - i * wp;
    float cx[6], cy[6];
    haar_coeffs(px, hp, wp, i, j, cx);
    haar_coeffs(py, hp, wp, i, j, cy);
#pragma unroll
    for (int o = 0; o < 2; ++o) {
      float sim = 0.f;
#pragma unroll
      for (int sc = 0; sc < 2; ++sc) {
        const float a = fabsf(cx[2 * sc + o]), bb = fabsf(cy[2 * sc + o]);
        sim += (2.f * a * bb + cst) / (a * a + bb * bb + cst);
      }
      sim *= 0.5f;
      const float wgt = fmaxf(fabsf(cx[4 + o]), fabsf(cy[4 + o]));
      num += (double)(wgt / (1.f + expf(-alpha * sim)));
      den += (double)wgt;
    }
  }
  __shared__ double red[2][kThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    num += __shfl_xor_sync(0xffffffffu, num, o);
    den += __shfl_xor_sync(0xffffffffu, den, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = num; red[1][threadIdx.x >> 5] = den; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, d = 0.0;
    for (int wv = 0; wv < kThreads / 32; ++wv) { a += red[0][wv]; d += red[1][wv]; }
    atomicAdd(acc + 2 * b, a);
    atomicAdd(acc + 2 * b + 1, d);
  }
}

}  // namespace

// minmax: the (batch, 4) floats immoco_metrics2d produced for the same views.  pooled: scratch of
// batch * 2 * hp * wp floats (hp, wp = the 2x2-pooled size of the zero-padded image).  acc: batch * 2 doubles, ZEROED.
extern "C" int immoco_haarpsi(const float* pred, int64_t pred_img_stride, int64_t pred_row_stride, int32_t pred_complex,
                              const float* gt, int64_t gt_img_stride, int64_t gt_row_stride, int32_t gt_complex,
                              int32_t batch, int32_t h, int32_t w, float c, float alpha, const float* minmax,
                              float* pooled, double* acc, void* stream) {
  if (!pred || !gt || !minmax || !pooled || !acc || batch < 0 || h < 1 || w < 1) return IMMOCO_ERR_BAD_ARG;
  if (batch == 0) return 0;
  if (batch > 65535) return IMMOCO_ERR_UNSUPPORTED;
  const int dp = (h % 2) > (w % 2) ? (h % 2) : (w % 2);
  const int hp = (h + dp) / 2, wp = (w + dp) / 2;
  if (h < 16 || w < 16) return IMMOCO_ERR_UNSUPPORTED;        // piq: the image must hold the 2^(scales+1) kernel
  cudaStream_t s = (cudaStream_t)stream;
  View vp{pred, pred_img_stride, pred_row_stride, pred_complex};
  View vg{gt, gt_img_stride, gt_row_stride, gt_complex};
  int gx = (hp * wp + kThreads - 1) / kThreads;
  if (gx > 1024) gx = 1024;
  haar_pool_kernel<<<dim3(gx, 2, batch), kThreads, 0, s>>>(vp, vg, minmax, pooled, h, w, hp, wp);
  IMMOCO_LAUNCH_CHECK();
  haar_sim_kernel<<<dim3(gx, 1, batch), kThreads, 0, s>>>(pooled, acc, hp, wp, c, alpha);
  IMMOCO_LAUNCH_CHECK();
  return 0;
}

// pred / gt: (batch, h, w) views with element strides (complex views are read as magnitudes).
// minmax: batch*4 floats scratch; acc: batch*4 doubles, ZEROED by the caller, receives
// {sum sq err, sum ssim map, ssim map size, -}.  kernel_size <= 11 (Gaussian sigma 1.5), pool >= 1.
extern "C" int immoco_metrics2d(const float* pred, int64_t pred_img_stride, int64_t pred_row_stride,
                                int32_t pred_complex, const float* gt, int64_t gt_img_stride,
                                int64_t gt_row_stride, int32_t gt_complex, int32_t batch, int32_t h, int32_t w,
                                int32_t kernel_size, int32_t pool, float* minmax, double* acc, void* stream) {
  if (!pred || !gt || !minmax || !acc || batch < 0 || h < 1 || w < 1 || pool < 1) return IMMOCO_ERR_BAD_ARG;
  if (kernel_size < 1 || kernel_size > kMaxTaps) return IMMOCO_ERR_UNSUPPORTED;
  if (batch == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  View vp{pred, pred_img_stride, pred_row_stride, pred_complex};
  View vg{gt, gt_img_stride, gt_row_stride, gt_complex};
  minmax_kernel<<<batch, 1024, 0, s>>>(vp, vg, minmax, h, w);
  IMMOCO_LAUNCH_CHECK();
  Taps win;
  double sum = 0.0, g[kMaxTaps];
  for (int t = 0; t < kernel_size; ++t) {
    const double c = t - (kernel_size - 1) / 2.0;
    g[t] = exp(-(c * c) / (2.0 * 1.5 * 1.5));
    sum += g[t];
  }
  for (int t = 0; t < kMaxTaps; ++t) win.g[t] = t < kernel_size ? (float)(g[t] / sum) : 0.f;
  const int hp = h / pool, wp = w / pool;
  const int oh = hp - kernel_size + 1, ow = wp - kernel_size + 1;
  // tiles cover the SSIM map and, through the last tile of each dimension, every full-resolution pixel
  const int gy = oh > 0 ? (oh + kT - 1) / kT : 1, gx = ow > 0 ? (ow + kT - 1) / kT : 1;
  dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)batch);
  ssim_mse_kernel<<<grid, kThreads, 0, s>>>(vp, vg, minmax, acc, h, w, pool, kernel_size, win);
  IMMOCO_LAUNCH_CHECK();
  return 0;
}
