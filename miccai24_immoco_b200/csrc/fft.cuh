// Shared-memory mixed-radix Stockham FFT used by the centred 2-D transform of the IM-MoCo forward
// model (src/utils/data_utils.py:29-34).  Lengths on the path: 320 = 4.4.4.5, 640 = 4.4.4.2.5,
// 368 = 4.4.23.  Radix 4 and 2 have butterfly forms; any other prime factor is evaluated as a
// direct small DFT with a single combined twiddle lookup (exact table, one rounding).
//
// The per-butterfly routine is __host__ __device__ so tests/cpu can replay it without a GPU.
#pragma once
#include <cuda_runtime.h>

#define IMMOCO_FFT_MAX_STAGES 12

struct FftPlan {
  int n;
  int n_stages;
  int radix[IMMOCO_FFT_MAX_STAGES];
};

// Factorise n: 4s first, then one 2, then odd factors ascending. Returns false if a prime
// factor exceeds max_radix (direct-DFT cost is O(r^2) per butterfly).
inline bool fft_make_plan(int n, FftPlan* plan, int max_radix = 64) {
  plan->n = n;
  plan->n_stages = 0;
  if (n < 1) return false;
  int r = n;
  while (r % 4 == 0) { plan->radix[plan->n_stages++] = 4; r /= 4; if (plan->n_stages >= IMMOCO_FFT_MAX_STAGES - 2) return false; }
  if (r % 2 == 0) { plan->radix[plan->n_stages++] = 2; r /= 2; }
  for (int f = 3; r > 1; f += 2) {
    while (r % f == 0) {
      if (f > max_radix || plan->n_stages >= IMMOCO_FFT_MAX_STAGES) return false;
      plan->radix[plan->n_stages++] = f;
      r /= f;
    }
  }
  return true;
}

// tw[t] = exp(-2 pi i t / N). INV selects the conjugate.
template <bool INV>
__host__ __device__ __forceinline__ float2 fft_tw(const float2* tw, int t) {
  float2 w = tw[t];
  if (INV) w.y = -w.y;
  return w;
}

__host__ __device__ __forceinline__ float2 fft_cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// One radix-R butterfly `j` (0 <= j < N/R) of the Stockham stage with sub-transform size Ns.
// a / d for 0 <= a < 2^20, 0 < d <= 2^12 without the integer-division routine: (a + 0.5) / d is at least
// 0.5 / d away from an integer, far more than the rounding of the fp32 product (inv_d = 1.0f / d)
__host__ __device__ __forceinline__ int fft_div(int a, float inv_d) {
  return (int)(((float)a + 0.5f) * inv_d);
}

// Stage constants (uniform over the block): T = N / R butterflies per transform, twstep = N / (Ns R)
template <bool INV>
__host__ __device__ __forceinline__ void fft_butterfly(const float2* in, float2* out, int N, int R,
                                                       int Ns, int j, const float2* tw, int T, int twstep,
                                                       float inv_ns) {
  const int q = fft_div(j, inv_ns);
  const int k = j - q * Ns;
  const int j0 = q * Ns * R + k;
  if (R == 4) {
    float2 v0 = in[j];
    float2 v1 = fft_cmul(in[j + T], fft_tw<INV>(tw, k * twstep));
    float2 v2 = fft_cmul(in[j + 2 * T], fft_tw<INV>(tw, 2 * k * twstep));
    float2 v3 = fft_cmul(in[j + 3 * T], fft_tw<INV>(tw, 3 * k * twstep));
    const float2 t0 = make_float2(v0.x + v2.x, v0.y + v2.y);
    const float2 t1 = make_float2(v0.x - v2.x, v0.y - v2.y);
    const float2 t2 = make_float2(v1.x + v3.x, v1.y + v3.y);
    const float2 d = make_float2(v1.x - v3.x, v1.y - v3.y);
    // forward: (v1 - v3) * (-i) ; inverse: * (+i)
    const float2 t3 = INV ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);
    out[j0] = make_float2(t0.x + t2.x, t0.y + t2.y);
    out[j0 + Ns] = make_float2(t1.x + t3.x, t1.y + t3.y);
    out[j0 + 2 * Ns] = make_float2(t0.x - t2.x, t0.y - t2.y);
    out[j0 + 3 * Ns] = make_float2(t1.x - t3.x, t1.y - t3.y);
  } else if (R == 2) {
    const float2 v0 = in[j];
    const float2 v1 = fft_cmul(in[j + T], fft_tw<INV>(tw, k * twstep));
    out[j0] = make_float2(v0.x + v1.x, v0.y + v1.y);
    out[j0 + Ns] = make_float2(v0.x - v1.x, v0.y - v1.y);
  } else {
    // direct R-point DFT; twiddle of input q for output p: q*k*twstep + ((p*q) mod R) * (N/R)
    for (int p = 0; p < R; ++p) {
      float2 acc = make_float2(0.f, 0.f);
      int pq = 0;  // (p*q) mod R, incrementally
      for (int q = 0; q < R; ++q) {
        int t = q * k * twstep + pq * T;
        if (t >= N) t -= N;
        const float2 w = fft_tw<INV>(tw, t);
        const float2 v = in[j + q * T];
        acc.x += v.x * w.x - v.y * w.y;
        acc.y += v.x * w.y + v.y * w.x;
        pq += p;
        if (pq >= R) pq -= R;
      }
      out[j0 + p * Ns] = acc;
    }
  }
}

#ifdef __CUDACC__
// Cooperative FFT of `nfft` transforms held in shared memory (`a`, transform t at a + t*tstride),
// scratch `b` of the same shape. All threads of the block must call. Returns the buffer holding
// the result (natural order).
template <bool INV>
__device__ float2* fft_smem(float2* a, float2* b, int nfft, int tstride, const FftPlan& plan,
                            const float2* tw) {
  const int N = plan.n;
  int Ns = 1;
  for (int s = 0; s < plan.n_stages; ++s) {
    const int R = plan.radix[s];
    const int T = N / R;                       // uniform per stage: three divisions per stage, not per butterfly
    const int twstep = N / (Ns * R);
    const float inv_t = 1.0f / (float)T, inv_ns = 1.0f / (float)Ns;
    for (int idx = threadIdx.x; idx < nfft * T; idx += blockDim.x) {
      const int t = fft_div(idx, inv_t);
      const int j = idx - t * T;
      fft_butterfly<INV>(a + t * tstride, b + t * tstride, N, R, Ns, j, tw, T, twstep, inv_ns);
    }
    __syncthreads();
    float2* tmp = a; a = b; b = tmp;
    Ns *= R;
  }
  return a;
}
#endif
