// Motion forward model, its adjoint, the data-consistency loss and the gradient-entropy prior
// (src/models/immoco.py:91-111,170-172; src/utils/data_utils.py:29-34;
//  src/utils/losses.py:20-40).
//
// The movement-group masks are column indicators (src/utils/motion_utils.py:56-109), hence
//   K[:,l] = F_H( w0[l] * F_W(I)[:,l] + sum_m w_m[l] * F_W(I_m)[:,l] )
// so ONE full row pass (static image) + a pruned row DFT on the few lines of each group + ONE
// column pass replace the reference's (M+1) full 2-D FFTs, and the resampled images I_m never
// leave shared memory.  All transforms are centred and un-normalised.
#include "common.cuh"
#include "fft.cuh"
#include "fit_batch.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kRowsPerCta = 2;   // row pass: transforms per CTA
constexpr int kColsPerCta = 4;   // column pass: adjacent columns per CTA (64 B segments)

// a mod n in [0, n) without the integer-division routine (the pruned row DFTs take one per tap): the
// quotient is estimated in fp32 and corrected by at most one step either way; exact for |a| < 2^22
// (here |a| <= W^2 / 4)
__device__ __forceinline__ int wrap_mod_fast(int a, int n, float inv_n) {
  const int q = __float2int_rd((float)a * inv_n);
  int r = a - q * n;
  if (r < 0) r += n;
  if (r >= n) r -= n;
  return r;
}

// ------------------------------------------------------------------------------------------------
// Reproducible accumulation of the image cotangent (deterministic mode).  The adjoint row passes scatter
// into d_image from many CTAs; float atomics make the sum depend on arrival order.  Instead every
// contribution is converted to 64-bit fixed point and added with integer atomics (associative ->
// order-free), then converted back once (d_image_finalize_kernel).  The scale is a power of two chosen per
// iteration from mx = max |component of d_c| (column pass, integer atomicMax): a contribution is a (pruned)
// row DFT of at most W weighted entries, |c| <= 4 * W2 * mx with W2 = W rounded up to a power of two and 4 =
// slack for mask weights; 2^20 contributions may pile up on one pixel: scale = 2^(62 - 20 - 2 - log2 W2 - e)
// with mx < 2^e.  A typical contribution (~ sqrt(W) * mx) keeps ~2^-35 relative resolution, far below fp32.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int fx_exponent(uint32_t dmax_bits, int W) {
  const float mx = __uint_as_float(dmax_bits);
  int e = 0;
  if (mx > 0.0f && mx <= 3.0e38f) e = ilogbf(mx) + 1;
  const int lw = 32 - __clz(W - 1);
  return 40 - lw - e;
}
__device__ __forceinline__ double pow2_double(int p) {
  return __longlong_as_double((long long)(1023 + p) << 52);
}
__device__ __forceinline__ void fx_add(long long* __restrict__ fx, size_t idx, float c, double sc) {
  const long long q = __double2ll_rn((double)c * sc);
  if (q != 0) atomicAdd(reinterpret_cast<unsigned long long*>(fx) + idx, (unsigned long long)q);
}

// ------------------------------------------------------------------------------------------------
// Row pass: centred 1-D FFT along W of `rows` rows.
//   out[r][l] (+)= scale * out_w[l] * sum_j in_w[j] * in[r][j] * exp(-/+ 2 pi i (l-W/2)(j-W/2)/W)
// ------------------------------------------------------------------------------------------------
// ATOMIC: the result is ADDED with float atomics (entries whose output weight is zero are skipped) -- used
// by the fused launches below, where CTAs of the pruned motion rows add into the same buffer concurrently.
// MODE 0: plain store (optionally read-modify-write), 1: float atomics, 2: 64-bit fixed-point atomics into `fx`
template <bool INV, int MODE>
__device__ __forceinline__ void rows_body(const float2* __restrict__ in, float2* __restrict__ out, int rows, int W,
                                          const FftPlan& plan, const float2* __restrict__ tw_g,
                                          const float* __restrict__ in_w, const float* __restrict__ out_w,
                                          float scale, int accumulate, int cta, long long* __restrict__ fx = nullptr,
                                          double fx_sc = 0.0) {
  constexpr bool ATOMIC = MODE == 1;
  extern __shared__ __align__(16) float2 sm2[];
  float2* tw = sm2;
  float2* a = tw + W;
  float2* b = a + kRowsPerCta * W;
  const int r0 = cta * kRowsPerCta;
  const int nr = min(kRowsPerCta, rows - r0);
  const int half = W >> 1;
  for (int t = threadIdx.x; t < W; t += kThreads) tw[t] = __ldg(tw_g + t);
  for (int idx = threadIdx.x; idx < nr * W; idx += kThreads) {
    const int r = idx / W, j = idx - r * W;
    float2 v = __ldg(in + (size_t)(r0 + r) * W + j);
    if (in_w) { const float s = __ldg(in_w + j); v.x *= s; v.y *= s; }
    int jj = j + half; if (jj >= W) jj -= W;
    a[r * W + jj] = v;
  }
  __syncthreads();
  const float2* res = fft_smem<INV>(a, b, nr, W, plan, tw);
  for (int idx = threadIdx.x; idx < nr * W; idx += kThreads) {
    const int r = idx / W, l = idx - r * W;
    int ll = l + half; if (ll >= W) ll -= W;
    float2 v = res[r * W + ll];
    float s = scale;
    if (out_w) s *= __ldg(out_w + l);
    v.x *= s; v.y *= s;
    float2* o = out + (size_t)(r0 + r) * W + l;
    if (MODE == 2) {
      const size_t gi = 2 * ((size_t)(r0 + r) * W + l);
      fx_add(fx, gi, v.x, fx_sc);
      fx_add(fx, gi + 1, v.y, fx_sc);
    } else if (ATOMIC) {
      if (s != 0.0f) atomicAdd(o, v);
    } else {
      if (accumulate) { const float2 p = *o; v.x += p.x; v.y += p.y; }
      *o = v;
    }
  }
}

template <bool INV>
__global__ void __launch_bounds__(kThreads)
fft_rows_kernel(const float2* __restrict__ in, float2* __restrict__ out, int rows, int W,
                const __grid_constant__ FftPlan plan, const float2* __restrict__ tw_g,
                const float* __restrict__ in_w, const float* __restrict__ out_w, float scale,
                int accumulate) {
  pdl_wait();
  rows_body<INV, 0>(in, out, rows, W, plan, tw_g, in_w, out_w, scale, accumulate, blockIdx.x);
}

// ------------------------------------------------------------------------------------------------
// Column pass helpers: a CTA owns kColsPerCta adjacent columns of one (H,W) image.
// smem transform t = column, stride HP = H+1 (bank spread), input/output rolled by H/2.
// ------------------------------------------------------------------------------------------------
template <int CPC = kColsPerCta>
__device__ __forceinline__ void load_cols(float2* a, const float2* __restrict__ in, int H, int W,
                                          int l0, int nc, int HP) {
  const int half = H >> 1;
  for (int idx = threadIdx.x; idx < H * CPC; idx += kThreads) {
    const int i = idx / CPC, c = idx - i * CPC;
    if (c < nc) {
      int ii = i + half; if (ii >= H) ii -= H;
      a[c * HP + ii] = __ldg(in + (size_t)i * W + l0 + c);
    }
  }
}

template <bool INV>
__global__ void __launch_bounds__(kThreads)
fft_cols_kernel(const float2* __restrict__ in, float2* __restrict__ out, int H, int W,
                const __grid_constant__ FftPlan plan, const float2* __restrict__ tw_g, float scale) {
  pdl_wait();
  extern __shared__ __align__(16) float2 sm2[];
  const int HP = H + 1;
  float2* tw = sm2;
  float2* a = tw + H;
  float2* b = a + kColsPerCta * HP;
  const int img = blockIdx.y;
  const int l0 = blockIdx.x * kColsPerCta;
  const int nc = min(kColsPerCta, W - l0);
  const int half = H >> 1;
  in += (size_t)img * H * W;
  out += (size_t)img * H * W;
  for (int t = threadIdx.x; t < H; t += kThreads) tw[t] = __ldg(tw_g + t);
  load_cols(a, in, H, W, l0, nc, HP);
  __syncthreads();
  const float2* res = fft_smem<INV>(a, b, nc, HP, plan, tw);
  for (int idx = threadIdx.x; idx < H * kColsPerCta; idx += kThreads) {
    const int k = idx / kColsPerCta, c = idx - k * kColsPerCta;
    if (c < nc) {
      int kk = k + half; if (kk >= H) kk -= H;
      float2 v = res[c * HP + kk];
      v.x *= scale; v.y *= scale;
      out[(size_t)k * W + l0 + c] = v;
    }
  }
}

// Fused column pass of the fit loop:
//   K = F_H(C);  loss += sum |K - K_in|^2;  dC = F_H^H((K - K_in) / (H W))
// CPC adjacent columns per CTA: 4 (64-byte segments) when that still gives every SM a CTA, else 2
template <int CPC>
__global__ void __launch_bounds__(kThreads)
colpass_loss_kernel(const __grid_constant__ FitBatch batch, int H, int W, const __grid_constant__ FftPlan plan,
                    const float2* __restrict__ tw_g, int zero_input) {
  pdl_wait();
  const FitBatchInst& inst = batch.inst[blockIdx.z];        // blockIdx.z = instance (fit_batch.cuh)
  const float2* __restrict__ c_in = inst.c_tmp;
  const float2* __restrict__ k_in = inst.k_in;
  float2* __restrict__ k_out = inst.k_out;
  float2* __restrict__ d_c = inst.d_c;
  double* __restrict__ loss_acc = inst.loss_dc;
  double* __restrict__ loss_slots = inst.slots_dc;
  uint32_t* __restrict__ dmax_bits = inst.dmax;
  float2* __restrict__ zero_after_load = zero_input ? inst.c_tmp : nullptr;
  extern __shared__ __align__(16) float2 sm2[];
  __shared__ float red[kThreads / 32];
  __shared__ float redmax[kThreads / 32];
  const int HP = H + 1;
  float2* tw = sm2;
  float2* a = tw + H;
  float2* b = a + CPC * HP;
  const int l0 = blockIdx.x * CPC;
  const int nc = min(CPC, W - l0);
  const int half = H >> 1;
  for (int t = threadIdx.x; t < H; t += kThreads) tw[t] = __ldg(tw_g + t);
  load_cols<CPC>(a, c_in, H, W, l0, nc, HP);
  if (zero_after_load) {     // == c_in: the next iteration's fused row launch ADDS into a zeroed buffer
    for (int idx = threadIdx.x; idx < H * CPC; idx += kThreads) {
      const int i = idx / CPC, c = idx - i * CPC;
      if (c < nc) zero_after_load[(size_t)i * W + l0 + c] = make_float2(0.f, 0.f);
    }
  }
  // the measured k-space of this CTA's columns: requested now, consumed after the forward transform
  constexpr int kMaxPre = 4;                       // items per thread held in registers (H * CPC <= 1024)
  float2 kin_pre[kMaxPre];
#pragma unroll
  for (int u = 0; u < kMaxPre; ++u) {
    const int idx = threadIdx.x + u * kThreads;
    const int k = idx / CPC, c = idx - k * CPC;
    kin_pre[u] = (idx < H * CPC && c < nc) ? __ldg(k_in + (size_t)k * W + l0 + c) : make_float2(0.f, 0.f);
  }
  __syncthreads();
  float2* res = fft_smem<false>(a, b, nc, HP, plan, tw);
  float2* other = (res == a) ? b : a;
  const float inv_hw = 1.0f / ((float)H * (float)W);
  float part = 0.0f;
  int u_it = 0;
  for (int idx = threadIdx.x; idx < H * CPC; idx += kThreads, ++u_it) {
    const int k = idx / CPC, c = idx - k * CPC;
    if (c < nc) {
      int kk = k + half; if (kk >= H) kk -= H;
      const float2 v = res[c * HP + kk];
      const size_t g = (size_t)k * W + l0 + c;
      float2 t;
      if (u_it < kMaxPre) {
        t = kin_pre[0];
#pragma unroll
        for (int u = 1; u < kMaxPre; ++u) if (u == u_it) t = kin_pre[u];
      } else {
        t = __ldg(k_in + g);
      }
      k_out[g] = v;
      const float dx = v.x - t.x, dy = v.y - t.y;
      part = fmaf(dx, dx, part);
      part = fmaf(dy, dy, part);
      // cotangent goes back to the rolled slot it came from: input of the adjoint transform
      res[c * HP + kk] = make_float2(dx * inv_hw, dy * inv_hw);
    }
  }
  part = warp_sum(part);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int wv = 0; wv < kThreads / 32; ++wv) s += (double)red[wv];
    // fit loop: this CTA's slot, added in slot order at the end of the call (reproducible); else one atomic
    if (loss_slots) loss_slots[blockIdx.x] = s;
    else atomicAdd(loss_acc, s);
  }
  const float2* adj = fft_smem<true>(res, other, nc, HP, plan, tw);
  float mx = 0.0f;
  for (int idx = threadIdx.x; idx < H * CPC; idx += kThreads) {
    const int i = idx / CPC, c = idx - i * CPC;
    if (c < nc) {
      int ii = i + half; if (ii >= H) ii -= H;
      const float2 v = adj[c * HP + ii];
      d_c[(size_t)i * W + l0 + c] = v;
      mx = fmaxf(mx, fmaxf(fabsf(v.x), fabsf(v.y)));
    }
  }
  if (dmax_bits) {      // largest cotangent component: scale of the fixed-point image-cotangent accumulation
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) redmax[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
      float m = 0.0f;
      for (int wv = 0; wv < kThreads / 32; ++wv) m = fmaxf(m, redmax[wv]);
      if (!(m <= 3.0e38f)) m = 3.0e38f;                 // NaN / inf (diverged fit): saturate
      atomicMax(dmax_bits, __float_as_uint(m));         // non-negative floats order like their bit patterns
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Bilinear resampling geometry (ATen grid_sampler_2d, bilinear / zeros / align_corners=False,
// called at src/models/immoco.py:97-107).
// ------------------------------------------------------------------------------------------------
struct Taps {
  int x0, y0;
  float wx0, wx1, wy0, wy1;   // (x1-ix), (ix-x0), (y1-iy), (iy-y0)
};

__device__ __forceinline__ Taps make_taps(float gx, float gy, int H, int W) {
  const float ix = ((gx + 1.0f) * (float)W - 1.0f) / 2.0f;
  const float iy = ((gy + 1.0f) * (float)H - 1.0f) / 2.0f;
  const float fx0 = floorf(ix), fy0 = floorf(iy);
  Taps t;
  // clamp before the int cast: far-out samples (all taps out of range) stay out of range
  t.x0 = (int)fminf(fmaxf(fx0, -2.0f), (float)W + 1.0f);
  t.y0 = (int)fminf(fmaxf(fy0, -2.0f), (float)H + 1.0f);
  t.wx1 = ix - fx0; t.wx0 = (fx0 + 1.0f) - ix;
  t.wy1 = iy - fy0; t.wy0 = (fy0 + 1.0f) - iy;
  return t;
}

__device__ __forceinline__ float2 fetch(const float2* __restrict__ img, int y, int x, int H, int W) {
  if ((unsigned)x < (unsigned)W && (unsigned)y < (unsigned)H) return __ldg(img + (size_t)y * W + x);
  return make_float2(0.f, 0.f);
}

// One CTA per (row i, group m): resample row i of the image along group m's deformed grid into
// shared memory, then evaluate the row DFT only at the group's phase-encode lines.
__device__ __forceinline__ void motion_rows_fwd_body(const float2* __restrict__ image, const float2* __restrict__ disp,
                                                     const float2* __restrict__ ident, const immoco_lines& lines,
                                                     const float2* __restrict__ tw_g, float2* __restrict__ c_out,
                                                     int H, int W, int i, int m) {
  extern __shared__ __align__(16) float2 sm2[];
  float2* tw = sm2;
  float2* row = tw + W;
  const int l_beg = __ldg(lines.group_ofs + m), l_end = __ldg(lines.group_ofs + m + 1);
  if (l_beg == l_end) return;
  for (int t = threadIdx.x; t < W; t += kThreads) tw[t] = __ldg(tw_g + t);
  const size_t base = ((size_t)m * H + i) * W;
  for (int j = threadIdx.x; j < W; j += kThreads) {
    const float2 d = __ldg(disp + base + j);
    const float2 id = __ldg(ident + (size_t)i * W + j);
    const Taps t = make_taps(id.x + d.x, id.y + d.y, H, W);
    const float2 nw = fetch(image, t.y0, t.x0, H, W), ne = fetch(image, t.y0, t.x0 + 1, H, W);
    const float2 sw = fetch(image, t.y0 + 1, t.x0, H, W), se = fetch(image, t.y0 + 1, t.x0 + 1, H, W);
    const float wnw = t.wx0 * t.wy0, wne = t.wx1 * t.wy0, wsw = t.wx0 * t.wy1, wse = t.wx1 * t.wy1;
    row[j] = make_float2(nw.x * wnw + ne.x * wne + sw.x * wsw + se.x * wse,
                         nw.y * wnw + ne.y * wne + sw.y * wsw + se.y * wse);
  }
  __syncthreads();
  const int half = W >> 1;
  const float inv_w = 1.0f / (float)W;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int li = l_beg + warp; li < l_end; li += kThreads / 32) {
    const int l = __ldg(lines.line_idx + li);
    const int a = l - half;
    float ax = 0.f, ay = 0.f;
    for (int j = lane; j < W; j += 32) {
      const float2 w = tw[wrap_mod_fast(a * (j - half), W, inv_w)];
      const float2 v = row[j];
      ax += v.x * w.x - v.y * w.y;
      ay += v.x * w.y + v.y * w.x;
    }
    ax = warp_sum(ax);
    ay = warp_sum(ay);
    if (lane == 0) {
      const float s = __ldg(lines.line_w + li);
      atomicAdd(&c_out[(size_t)i * W + l].x, s * ax);
      atomicAdd(&c_out[(size_t)i * W + l].y, s * ay);
    }
  }
}

__global__ void __launch_bounds__(kThreads)
motion_rows_fwd_kernel(const float2* __restrict__ image, const float2* __restrict__ disp,
                       const float2* __restrict__ ident, const __grid_constant__ immoco_lines lines,
                       const float2* __restrict__ tw_g, float2* __restrict__ c_out, int H, int W) {
  pdl_wait();
  motion_rows_fwd_body(image, disp, ident, lines, tw_g, c_out, H, W, blockIdx.x, blockIdx.y);
}

// Fused row launch of the fit loop: blockIdx.y == 0 -> static row pass (2 rows per CTA, result ADDED into
// c_out, zero-weight lines skipped); blockIdx.y == 1 + m -> pruned rows of movement group m.  c_out must be
// zero on entry (colpass_loss_kernel re-zeroes it after loading).  One launch instead of two dependent ones.
__global__ void __launch_bounds__(kThreads)
rows_fwd_fused_kernel(const __grid_constant__ FitBatch batch, const float2* __restrict__ ident,
                      const __grid_constant__ FftPlan plan, const float2* __restrict__ tw_g, int H, int W) {
  pdl_wait();
  const FitBatchInst& inst = batch.inst[blockIdx.z];
  const float2* __restrict__ image = inst.image;
  const float2* __restrict__ disp = inst.disp;
  float2* __restrict__ c_out = inst.c_tmp;
  const immoco_lines& lines = inst.lines;
  if (blockIdx.y == 0) {
    if ((int)blockIdx.x * kRowsPerCta >= H) return;
    rows_body<false, 1>(image, c_out, H, W, plan, tw_g, nullptr, lines.static_w, 1.0f, 0, blockIdx.x);
  } else {
    if ((int)blockIdx.y - 1 >= lines.n_groups) return;
    motion_rows_fwd_body(image, disp, ident, lines, tw_g, c_out, H, W, blockIdx.x, (int)blockIdx.y - 1);
  }
}

// Adjoint of the kernel above: pruned inverse row DFT -> d(moved row) -> scatter into d_image
// (grid_sampler_2d_backward) and the cotangent of the PRE-tanh displacement.
template <bool DET>
__device__ __forceinline__ void motion_rows_bwd_body(const float2* __restrict__ d_c, const float2* __restrict__ image,
                                                     const float2* __restrict__ disp, const float2* __restrict__ ident,
                                                     const immoco_lines& lines, const float2* __restrict__ tw_g,
                                                     float2* __restrict__ d_image, float2* __restrict__ d_disp,
                                                     int pre_tanh, int H, int W, int i, int m,
                                                     long long* __restrict__ fx = nullptr, double fx_sc = 0.0) {
  // scatter of one tap's contribution: float2 atomic, or two fixed-point atomics (deterministic mode)
  auto scatter = [&](int y, int x, float w, float gr, float gi) {
    const size_t idx = (size_t)y * W + x;
    if (DET) {
      fx_add(fx, 2 * idx, w * gr, fx_sc);
      fx_add(fx, 2 * idx + 1, w * gi, fx_sc);
    } else {
      atomicAdd(d_image + idx, make_float2(w * gr, w * gi));
    }
  };
  extern __shared__ __align__(16) float2 sm2[];
  float2* tw = sm2;
  float2* gl = tw + W;
  int* la = reinterpret_cast<int*>(gl + lines.max_lines);
  const int l_beg = __ldg(lines.group_ofs + m), l_end = __ldg(lines.group_ofs + m + 1);
  const int nl = l_end - l_beg;
  const int half = W >> 1;
  for (int t = threadIdx.x; t < W; t += kThreads) tw[t] = __ldg(tw_g + t);
  for (int li = threadIdx.x; li < nl; li += kThreads) {
    const int l = __ldg(lines.line_idx + l_beg + li);
    const float s = __ldg(lines.line_w + l_beg + li);
    const float2 g = __ldg(d_c + (size_t)i * W + l);
    gl[li] = make_float2(s * g.x, s * g.y);
    la[li] = l - half;
  }
  __syncthreads();
  const size_t base = ((size_t)m * H + i) * W;
  const float mx = (float)W / 2.0f, my = (float)H / 2.0f;
  const float inv_w = 1.0f / (float)W;
  for (int j = threadIdx.x; j < W; j += kThreads) {
    const int bq = j - half;
    // requested before the pruned DFT so that the round trip hides behind it
    const float2 d = __ldg(disp + base + j);
    const float2 id = __ldg(ident + (size_t)i * W + j);
    float gr = 0.f, gi = 0.f;   // cotangent of the moved pixel (re, im)
    for (int li = 0; li < nl; ++li) {
      const float2 w = tw[wrap_mod_fast(la[li] * bq, W, inv_w)];
      const float2 g = gl[li];
      gr += g.x * w.x + g.y * w.y;     // g * conj(w)
      gi += g.y * w.x - g.x * w.y;
    }
    const Taps t = make_taps(id.x + d.x, id.y + d.y, H, W);
    const int x0 = t.x0, y0 = t.y0, x1 = t.x0 + 1, y1 = t.y0 + 1;
    const bool inx0 = (unsigned)x0 < (unsigned)W, inx1 = (unsigned)x1 < (unsigned)W;
    const bool iny0 = (unsigned)y0 < (unsigned)H, iny1 = (unsigned)y1 < (unsigned)H;
    float gix = 0.f, giy = 0.f;
    if (iny0 && inx0) {
      scatter(y0, x0, t.wx0 * t.wy0, gr, gi);
      const float2 v = __ldg(image + (size_t)y0 * W + x0);
      const float dot = v.x * gr + v.y * gi;
      gix -= dot * t.wy0; giy -= dot * t.wx0;
    }
    if (iny0 && inx1) {
      scatter(y0, x1, t.wx1 * t.wy0, gr, gi);
      const float2 v = __ldg(image + (size_t)y0 * W + x1);
      const float dot = v.x * gr + v.y * gi;
      gix += dot * t.wy0; giy -= dot * t.wx1;
    }
    if (iny1 && inx0) {
      scatter(y1, x0, t.wx0 * t.wy1, gr, gi);
      const float2 v = __ldg(image + (size_t)y1 * W + x0);
      const float dot = v.x * gr + v.y * gi;
      gix -= dot * t.wy1; giy += dot * t.wx0;
    }
    if (iny1 && inx1) {
      scatter(y1, x1, t.wx1 * t.wy1, gr, gi);
      const float2 v = __ldg(image + (size_t)y1 * W + x1);
      const float dot = v.x * gr + v.y * gi;
      gix += dot * t.wy1; giy += dot * t.wx1;
    }
    const float sx = pre_tanh ? (1.0f - d.x * d.x) : 1.0f;
    const float sy = pre_tanh ? (1.0f - d.y * d.y) : 1.0f;
    d_disp[base + j] = make_float2(sx * (mx * gix), sy * (my * giy));
  }
}

__global__ void __launch_bounds__(kThreads)
motion_rows_bwd_kernel(const float2* __restrict__ d_c, const float2* __restrict__ image,
                       const float2* __restrict__ disp, const float2* __restrict__ ident,
                       const __grid_constant__ immoco_lines lines, const float2* __restrict__ tw_g,
                       float2* __restrict__ d_image, float2* __restrict__ d_disp, int pre_tanh, int H,
                       int W) {
  pdl_wait();
  motion_rows_bwd_body<false>(d_c, image, disp, ident, lines, tw_g, d_image, d_disp, pre_tanh, H, W, blockIdx.x, blockIdx.y);
}

// Fused adjoint row launch of the fit loop: blockIdx.y == 0 -> adjoint static row pass, ADDED into d_image
// with float atomics; blockIdx.y == 1 + m -> adjoint of movement group m (scatter into d_image, d_disp).
// DET: every contribution goes to the 64-bit fixed-point plane `fx` (scale from dmax_bits, see fx_exponent);
// d_image itself (it holds the gradient-entropy term) is only touched by d_image_finalize_kernel afterwards.
template <bool DET>
__global__ void __launch_bounds__(kThreads)
rows_bwd_fused_kernel(const __grid_constant__ FitBatch batch, const float2* __restrict__ ident,
                      const __grid_constant__ FftPlan plan, const float2* __restrict__ tw_g, int H, int W) {
  pdl_wait();
  const FitBatchInst& inst = batch.inst[blockIdx.z];
  const float2* __restrict__ d_c = inst.d_c;
  const float2* __restrict__ image = inst.image;
  const float2* __restrict__ disp = inst.disp;
  float2* __restrict__ d_image = inst.d_image;
  float2* __restrict__ d_disp = inst.d_disp;
  long long* __restrict__ fx = inst.fx;
  const immoco_lines& lines = inst.lines;
  const double sc = DET ? pow2_double(fx_exponent(*inst.dmax, W)) : 0.0;
  if (blockIdx.y == 0) {
    if ((int)blockIdx.x * kRowsPerCta >= H) return;
    rows_body<true, DET ? 2 : 1>(d_c, d_image, H, W, plan, tw_g, lines.static_w, nullptr, 1.0f, 1, blockIdx.x, fx, sc);
  } else {
    if ((int)blockIdx.y - 1 >= lines.n_groups) return;
    motion_rows_bwd_body<DET>(d_c, image, disp, ident, lines, tw_g, d_image, d_disp, 1, H, W, blockIdx.x,
                              (int)blockIdx.y - 1, fx, sc);
  }
}

// d_image += fixed-point plane (converted back with the iteration's scale); the plane is re-zeroed for the
// next iteration.  One rounding per element whatever order the contributions arrived in.
__global__ void __launch_bounds__(kThreads)
d_image_finalize_kernel(const __grid_constant__ FitBatch batch, int n, int W) {
  pdl_wait();
  const FitBatchInst& inst = batch.inst[blockIdx.z];
  float* __restrict__ d_image = reinterpret_cast<float*>(inst.d_image);
  long long* __restrict__ fx = inst.fx;
  const double inv = pow2_double(-fx_exponent(*inst.dmax, W));
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
    const long long q = fx[i];
    if (q != 0) {
      d_image[i] += (float)((double)q * inv);
      fx[i] = 0;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Gradient entropy (src/utils/losses.py:20-40): GE = -sum G log(G + 1e-24),
// G[i,j] = |I[i,j]-I[i,j+1]| [j<W-1] + |I[i,j]-I[i+1,j]| [i<H-1].  Value + gather-form gradient.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float cabs2(float2 z) { return hypotf(z.x, z.y); }

__device__ __forceinline__ float ge_dloss_dg(float g) {
  const float ge = g + 1e-24f;
  return -(logf(ge) + g / ge);
}

__global__ void __launch_bounds__(kThreads)
grad_entropy_kernel(const __grid_constant__ FitBatch batch, float grad_scale, int accumulate, int H, int W) {
  pdl_wait();
  const FitBatchInst& inst = batch.inst[blockIdx.z];
  const float2* __restrict__ img = inst.image;
  float2* __restrict__ d_img = inst.d_image;
  double* __restrict__ loss_acc = inst.loss_ge;
  double* __restrict__ loss_slots = inst.slots_ge;
  __shared__ float red[kThreads / 32];
  const int P = H * W;
  float part = 0.0f;
  for (int idx = blockIdx.x * kThreads + threadIdx.x; idx < P; idx += gridDim.x * kThreads) {
    const int i = idx / W, j = idx - i * W;
    const float2 z = __ldg(img + idx);
    const bool has_r = j < W - 1, has_d = i < H - 1;
    float2 gsum = make_float2(0.f, 0.f);
    // own cell
    float2 dx = make_float2(0.f, 0.f), dy = make_float2(0.f, 0.f);
    if (has_r) { const float2 r = __ldg(img + idx + 1); dx = make_float2(z.x - r.x, z.y - r.y); }
    if (has_d) { const float2 dn = __ldg(img + idx + W); dy = make_float2(z.x - dn.x, z.y - dn.y); }
    const float ax = cabs2(dx), ay = cabs2(dy);
    const float g = ax + ay;
    part -= g * logf(g + 1e-24f);
    const float gg = ge_dloss_dg(g);
    if (ax > 0.f) { gsum.x += gg * dx.x / ax; gsum.y += gg * dx.y / ax; }
    if (ay > 0.f) { gsum.x += gg * dy.x / ay; gsum.y += gg * dy.y / ay; }
    // left neighbour's horizontal difference contains -z
    if (j > 0) {
      const float2 l = __ldg(img + idx - 1);
      const float2 ldx = make_float2(l.x - z.x, l.y - z.y);
      float2 ldy = make_float2(0.f, 0.f);
      if (has_d) { const float2 ld = __ldg(img + idx - 1 + W); ldy = make_float2(l.x - ld.x, l.y - ld.y); }
      const float lax = cabs2(ldx);
      const float lg = lax + cabs2(ldy);
      if (lax > 0.f) { const float s = ge_dloss_dg(lg) / lax; gsum.x -= s * ldx.x; gsum.y -= s * ldx.y; }
    }
    // upper neighbour's vertical difference contains -z
    if (i > 0) {
      const float2 u = __ldg(img + idx - W);
      const float2 udy = make_float2(u.x - z.x, u.y - z.y);
      float2 udx = make_float2(0.f, 0.f);
      if (has_r) { const float2 ur = __ldg(img + idx - W + 1); udx = make_float2(u.x - ur.x, u.y - ur.y); }
      const float uay = cabs2(udy);
      const float ug = cabs2(udx) + uay;
      if (uay > 0.f) { const float s = ge_dloss_dg(ug) / uay; gsum.x -= s * udy.x; gsum.y -= s * udy.y; }
    }
    float2 o = make_float2(grad_scale * gsum.x, grad_scale * gsum.y);
    if (accumulate) { const float2 p = d_img[idx]; o.x += p.x; o.y += p.y; }
    d_img[idx] = o;
  }
  part = warp_sum(part);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int wv = 0; wv < kThreads / 32; ++wv) s += (double)red[wv];
    if (loss_slots) loss_slots[blockIdx.x] = s;
    else atomicAdd(loss_acc, s);
  }
}

// ------------------------------------------------------------------------------------------------
// host-side launch helpers
// ------------------------------------------------------------------------------------------------
struct Plans {
  FftPlan h, w;
  bool ok;
};

Plans make_plans(int H, int W) {
  Plans p;
  p.ok = (H >= 2) && (W >= 2) && (H % 2 == 0) && (W % 2 == 0) && fft_make_plan(H, &p.h) && fft_make_plan(W, &p.w);
  return p;
}

template <typename K>
void allow_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

int launch_rows(const float* in, float* out, int rows, int W, const FftPlan& plan, const float* tw,
                const float* in_w, const float* out_w, float scale, int accumulate, bool inv,
                cudaStream_t s) {
  const size_t smem = (size_t)(W + 2 * kRowsPerCta * W) * sizeof(float2);
  if (smem > 200 * 1024) return IMMOCO_ERR_UNSUPPORTED;
  const int grid = (rows + kRowsPerCta - 1) / kRowsPerCta;
  if (inv) {
    allow_smem(fft_rows_kernel<true>, smem);
    immoco_launch(fft_rows_kernel<true>, dim3(grid), dim3(kThreads), smem, s, (const float2*)in, (float2*)out, rows, W, plan,
                                                      (const float2*)tw, in_w, out_w, scale, accumulate);
  } else {
    allow_smem(fft_rows_kernel<false>, smem);
    immoco_launch(fft_rows_kernel<false>, dim3(grid), dim3(kThreads), smem, s, (const float2*)in, (float2*)out, rows, W, plan,
                                                       (const float2*)tw, in_w, out_w, scale, accumulate);
  }
  IMMOCO_LAUNCH_CHECK();
  return 0;
}

size_t cols_smem(int H) { return (size_t)(H + 2 * kColsPerCta * (H + 1)) * sizeof(float2); }

int launch_cols(const float* in, float* out, int batch, int H, int W, const FftPlan& plan,
                const float* tw, float scale, bool inv, cudaStream_t s) {
  const size_t smem = cols_smem(H);
  if (smem > 200 * 1024) return IMMOCO_ERR_UNSUPPORTED;
  dim3 grid((W + kColsPerCta - 1) / kColsPerCta, batch);
  if (inv) {
    allow_smem(fft_cols_kernel<true>, smem);
    immoco_launch(fft_cols_kernel<true>, dim3(grid), dim3(kThreads), smem, s, (const float2*)in, (float2*)out, H, W, plan,
                                                      (const float2*)tw, scale);
  } else {
    allow_smem(fft_cols_kernel<false>, smem);
    immoco_launch(fft_cols_kernel<false>, dim3(grid), dim3(kThreads), smem, s, (const float2*)in, (float2*)out, H, W, plan,
                                                       (const float2*)tw, scale);
  }
  IMMOCO_LAUNCH_CHECK();
  return 0;
}

int launch_motion_fwd(const float* image, const float* disp, const float* ident,
                      const immoco_lines* lines, const float* tw_w, float* c_out, int H, int W,
                      cudaStream_t s) {
  if (lines->n_groups <= 0) return 0;
  const size_t smem = (size_t)2 * W * sizeof(float2);
  allow_smem(motion_rows_fwd_kernel, smem);
  dim3 grid(H, lines->n_groups);
  immoco_launch(motion_rows_fwd_kernel, dim3(grid), dim3(kThreads), smem, s, (const float2*)image, (const float2*)disp,
                                                     (const float2*)ident, *lines, (const float2*)tw_w,
                                                     (float2*)c_out, H, W);
  IMMOCO_LAUNCH_CHECK();
  return 0;
}

int launch_motion_bwd(const float* d_c, const float* image, const float* disp, const float* ident,
                      const immoco_lines* lines, const float* tw_w, float* d_image, float* d_disp,
                      int pre_tanh, int H, int W, cudaStream_t s) {
  if (lines->n_groups <= 0) return 0;
  const size_t smem = (size_t)W * sizeof(float2) + (size_t)lines->max_lines * (sizeof(float2) + sizeof(int));
  if (smem > 200 * 1024) return IMMOCO_ERR_UNSUPPORTED;
  allow_smem(motion_rows_bwd_kernel, smem);
  dim3 grid(H, lines->n_groups);
  immoco_launch(motion_rows_bwd_kernel, dim3(grid), dim3(kThreads), smem, s, (const float2*)d_c, (const float2*)image,
                                                     (const float2*)disp, (const float2*)ident, *lines,
                                                     (const float2*)tw_w, (float2*)d_image,
                                                     (float2*)d_disp, pre_tanh, H, W);
  IMMOCO_LAUNCH_CHECK();
  return 0;
}

}  // namespace

extern "C" int immoco_fft2c(const float* in, float* out, float* tmp, int32_t batch, int32_t h,
                            int32_t w, const float* tw_h, const float* tw_w, int32_t inverse,
                            float scale, void* stream) {
  if (batch < 0 || !in || !out || !tmp) return IMMOCO_ERR_BAD_ARG;
  const Plans p = make_plans(h, w);
  if (!p.ok) return IMMOCO_ERR_UNSUPPORTED;
  if (batch == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (int e = launch_rows(in, tmp, batch * h, w, p.w, tw_w, nullptr, nullptr, 1.0f, 0, inverse != 0, s)) return e;
  return launch_cols(tmp, out, batch, h, w, p.h, tw_h, scale, inverse != 0, s);
}

extern "C" int immoco_forward_model(const float* image, const float* disp, const float* ident,
                                    const immoco_lines* lines, const float* tw_h, const float* tw_w,
                                    float* c_tmp, float* k_out, int32_t h, int32_t w, void* stream) {
  if (!lines) return IMMOCO_ERR_BAD_ARG;
  const Plans p = make_plans(h, w);
  if (!p.ok) return IMMOCO_ERR_UNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
  if (int e = launch_rows(image, c_tmp, h, w, p.w, tw_w, nullptr, lines->static_w, 1.0f, 0, false, s)) return e;
  if (int e = launch_motion_fwd(image, disp, ident, lines, tw_w, c_tmp, h, w, s)) return e;
  return launch_cols(c_tmp, k_out, 1, h, w, p.h, tw_h, 1.0f, false, s);
}

extern "C" int immoco_forward_model_bwd(const float* d_k, const float* image, const float* disp,
                                        const float* ident, const immoco_lines* lines,
                                        const float* tw_h, const float* tw_w, float* c_tmp,
                                        float* d_image, float* d_disp, int32_t pre_tanh, int32_t h,
                                        int32_t w, void* stream) {
  if (!lines) return IMMOCO_ERR_BAD_ARG;
  const Plans p = make_plans(h, w);
  if (!p.ok) return IMMOCO_ERR_UNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
  if (int e = launch_cols(d_k, c_tmp, 1, h, w, p.h, tw_h, 1.0f, true, s)) return e;
  if (int e = launch_rows(c_tmp, d_image, h, w, p.w, tw_w, lines->static_w, nullptr, 1.0f, 1, true, s)) return e;
  return launch_motion_bwd(c_tmp, image, disp, ident, lines, tw_w, d_image, d_disp, pre_tanh, h, w, s);
}

// columns per CTA of the fused column pass: 4 (64-byte segments) when that still gives every SM a CTA, else 2
// (narrow images: 320 columns -> 80 CTAs of 4)
static int colpass_cpc(int w) { return ((w + kColsPerCta - 1) / kColsPerCta >= IMMOCO_NUM_SMS) ? kColsPerCta : 2; }
static int colpass_grid(int w) { const int cpc = colpass_cpc(w); return (w + cpc - 1) / cpc; }
static int grad_entropy_grid(int h, int w) { return (h * w + kThreads - 1) / kThreads; }

static int colpass_launch(const FitBatch& b, const float* tw_h, int32_t h, int32_t w, int zero_input, void* stream) {
  const Plans p = make_plans(h, w);
  if (!p.ok) return IMMOCO_ERR_UNSUPPORTED;
  if (b.n < 1 || b.n > kMaxFitBatch) return IMMOCO_ERR_BAD_ARG;
  const size_t smem = cols_smem(h);
  if (smem > 200 * 1024) return IMMOCO_ERR_UNSUPPORTED;
  const dim3 grid(colpass_grid(w), 1, b.n);
  if (colpass_cpc(w) == kColsPerCta) {
    allow_smem(colpass_loss_kernel<kColsPerCta>, smem);
    immoco_launch(colpass_loss_kernel<kColsPerCta>, grid, dim3(kThreads), smem, (cudaStream_t)stream, b, h, w, p.h,
                  (const float2*)tw_h, zero_input);
  } else {
    allow_smem(colpass_loss_kernel<2>, smem);
    immoco_launch(colpass_loss_kernel<2>, grid, dim3(kThreads), smem, (cudaStream_t)stream, b, h, w, p.h,
                  (const float2*)tw_h, zero_input);
  }
  IMMOCO_LAUNCH_CHECK();
  return 0;
}

static FitBatch one_colpass(const float* c, const float* k_in, float* k_out, float* d_c, double* loss_acc,
                            double* loss_slots, uint32_t* dmax_bits) {
  FitBatch b = {};
  b.n = 1;
  b.inst[0].c_tmp = (float2*)c;
  b.inst[0].k_in = (const float2*)k_in;
  b.inst[0].k_out = (float2*)k_out;
  b.inst[0].d_c = (float2*)d_c;
  b.inst[0].loss_dc = loss_acc;
  b.inst[0].slots_dc = loss_slots;
  b.inst[0].dmax = dmax_bits;
  return b;
}

extern "C" int immoco_colpass_loss(const float* c, const float* k_in, float* k_out, float* d_c,
                                   double* loss_acc, const float* tw_h, int32_t h, int32_t w,
                                   void* stream) {
  return colpass_launch(one_colpass(c, k_in, k_out, d_c, loss_acc, nullptr, nullptr), tw_h, h, w, 0, stream);
}
// fit.cu: column pass of a batch of instances; zero_input: the pass leaves its input buffer zeroed for the next
// iteration's fused row launch; the loss goes to per-CTA slots when an instance has slots_dc, the largest
// cotangent component to *dmax when it has one
int immoco_colpass_loss_batch(const FitBatch& b, const float* tw_h, int h, int w, int zero_input, void* stream) {
  return colpass_launch(b, tw_h, h, w, zero_input, stream);
}
int immoco_colpass_loss_slots(const float* c, const float* k_in, float* k_out, float* d_c, double* loss_acc,
                              const float* tw_h, int h, int w, double* loss_slots, uint32_t* dmax_bits, void* stream) {
  return colpass_launch(one_colpass(c, k_in, k_out, d_c, loss_acc, loss_slots, dmax_bits), tw_h, h, w, 0, stream);
}

int immoco_grad_entropy_batch(const FitBatch& b, float grad_scale, int accumulate, int h, int w, void* stream) {
  if (h < 1 || w < 1 || b.n < 1 || b.n > kMaxFitBatch) return IMMOCO_ERR_BAD_ARG;
  immoco_launch(grad_entropy_kernel, dim3(grad_entropy_grid(h, w), 1, b.n), dim3(kThreads), 0, (cudaStream_t)stream, b,
                grad_scale, accumulate, h, w);
  IMMOCO_LAUNCH_CHECK();
  return 0;
}
extern "C" int immoco_grad_entropy(const float* image, float grad_scale, double* loss_acc,
                                   float* d_image, int32_t accumulate, int32_t h, int32_t w,
                                   void* stream) {
  FitBatch b = {};
  b.n = 1;
  b.inst[0].image = (const float2*)image;
  b.inst[0].d_image = (float2*)d_image;
  b.inst[0].loss_ge = loss_acc;
  return immoco_grad_entropy_batch(b, grad_scale, accumulate, h, w, stream);
}

extern "C" int immoco_fit_loss_slots(int32_t h, int32_t w, int32_t out[2]) {
  if (h < 2 || w < 2 || !out) return IMMOCO_ERR_BAD_ARG;
  out[0] = colpass_grid(w);
  out[1] = grad_entropy_grid(h, w);
  return 0;
}

// d_image += the fixed-point cotangent plane of this iteration (deterministic mode), plane re-zeroed
int immoco_d_image_finalize_batch(const FitBatch& b, int h, int w, void* stream) {
  if (b.n < 1 || b.n > kMaxFitBatch) return IMMOCO_ERR_BAD_ARG;
  const int n = 2 * h * w;
  int grid = (n + kThreads - 1) / kThreads;
  if (grid > IMMOCO_NUM_SMS * 8) grid = IMMOCO_NUM_SMS * 8;
  immoco_launch(d_image_finalize_kernel, dim3(grid, 1, b.n), dim3(kThreads), 0, (cudaStream_t)stream, b, n, w);
  IMMOCO_LAUNCH_CHECK();
  return 0;
}

// used by fit.cu
int immoco_rows_static(const float* in, float* out, int h, int w, const float* tw_w,
                       const float* in_w, const float* out_w, int accumulate, bool inv, void* stream) {
  const Plans p = make_plans(h, w);
  if (!p.ok) return IMMOCO_ERR_UNSUPPORTED;
  return launch_rows(in, out, h, w, p.w, tw_w, in_w, out_w, 1.0f, accumulate, inv, (cudaStream_t)stream);
}
int immoco_motion_rows_fwd(const float* image, const float* disp, const float* ident,
                           const immoco_lines* lines, const float* tw_w, float* c_out, int h, int w,
                           void* stream) {
  return launch_motion_fwd(image, disp, ident, lines, tw_w, c_out, h, w, (cudaStream_t)stream);
}
int immoco_motion_rows_bwd(const float* d_c, const float* image, const float* disp, const float* ident,
                           const immoco_lines* lines, const float* tw_w, float* d_image, float* d_disp,
                           int h, int w, void* stream) {
  return launch_motion_bwd(d_c, image, disp, ident, lines, tw_w, d_image, d_disp, 1, h, w, (cudaStream_t)stream);
}

// Fused row launches of the fit loop (fit.cu) over a batch of instances of one shape.  Every instance's c_tmp
// must be zero on entry of the forward one.
static int max_lines_of(const FitBatch& b) {
  int m = 1;
  for (int i = 0; i < b.n; ++i)
    if (b.inst[i].lines.max_lines > m) m = b.inst[i].lines.max_lines;
  return m;
}
int immoco_rows_fwd_fused_batch(const FitBatch& b, const float* ident, const float* tw_w, int h, int w, void* stream) {
  const Plans p = make_plans(h, w);
  if (!p.ok) return IMMOCO_ERR_UNSUPPORTED;
  if (b.n < 1 || b.n > kMaxFitBatch) return IMMOCO_ERR_BAD_ARG;
  const size_t smem = (size_t)(w + 2 * kRowsPerCta * w) * sizeof(float2);     // >= the motion rows' 2 W
  if (smem > 200 * 1024) return IMMOCO_ERR_UNSUPPORTED;
  allow_smem(rows_fwd_fused_kernel, smem);
  dim3 grid(h, 1 + b.inst[0].lines.n_groups, b.n);
  immoco_launch(rows_fwd_fused_kernel, grid, dim3(kThreads), smem, (cudaStream_t)stream, b, (const float2*)ident, p.w,
                (const float2*)tw_w, h, w);
  IMMOCO_LAUNCH_CHECK();
  return 0;
}
// det: deterministic mode (fixed-point accumulation into inst.fx, d_image itself untouched until the finalize)
int immoco_rows_bwd_fused_batch(const FitBatch& b, const float* ident, const float* tw_w, int h, int w, bool det,
                                void* stream) {
  const Plans p = make_plans(h, w);
  if (!p.ok) return IMMOCO_ERR_UNSUPPORTED;
  if (b.n < 1 || b.n > kMaxFitBatch) return IMMOCO_ERR_BAD_ARG;
  for (int i = 0; i < b.n; ++i)
    if (det && (!b.inst[i].fx || !b.inst[i].dmax)) return IMMOCO_ERR_BAD_ARG;
  size_t smem = (size_t)(w + 2 * kRowsPerCta * w) * sizeof(float2);
  const size_t smem_m = (size_t)w * sizeof(float2) + (size_t)max_lines_of(b) * (sizeof(float2) + sizeof(int));
  if (smem_m > smem) smem = smem_m;
  if (smem > 200 * 1024) return IMMOCO_ERR_UNSUPPORTED;
  dim3 grid(h, 1 + b.inst[0].lines.n_groups, b.n);
  if (det) {
    allow_smem(rows_bwd_fused_kernel<true>, smem);
    immoco_launch(rows_bwd_fused_kernel<true>, grid, dim3(kThreads), smem, (cudaStream_t)stream, b, (const float2*)ident,
                  p.w, (const float2*)tw_w, h, w);
  } else {
    allow_smem(rows_bwd_fused_kernel<false>, smem);
    immoco_launch(rows_bwd_fused_kernel<false>, grid, dim3(kThreads), smem, (cudaStream_t)stream, b, (const float2*)ident,
                  p.w, (const float2*)tw_w, h, w);
  }
  IMMOCO_LAUNCH_CHECK();
  return 0;
}
