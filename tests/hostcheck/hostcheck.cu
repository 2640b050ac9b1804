// Test-only host replay of the __host__ __device__ index / butterfly routines the kernels use
// (no GPU needed): lets the CPU test-suite pin the exact integer maths against the oracle.
#include <cmath>
#include <cstring>
#include <vector>

#include "../../miccai24_immoco_b200/csrc/common.cuh"
#include "../../miccai24_immoco_b200/csrc/fft.cuh"

extern "C" int hostcheck_fft(const float* in, float* out, int n, int nfft, int inverse) {
  FftPlan plan;
  if (!fft_make_plan(n, &plan)) return -1;
  std::vector<float2> tw(n), a(n), b(n);
  for (int t = 0; t < n; ++t) {
    const double ang = -2.0 * M_PI * (double)t / (double)n;
    tw[t] = make_float2((float)std::cos(ang), (float)std::sin(ang));
  }
  for (int f = 0; f < nfft; ++f) {
    std::memcpy(a.data(), in + (size_t)f * n * 2, sizeof(float2) * n);
    float2* pa = a.data();
    float2* pb = b.data();
    int Ns = 1;
    for (int s = 0; s < plan.n_stages; ++s) {
      const int R = plan.radix[s];
      const int T = n / R, twstep = n / (Ns * R);
      const float inv_ns = 1.0f / (float)Ns;
      for (int j = 0; j < n / R; ++j) {
        if (inverse) fft_butterfly<true>(pa, pb, n, R, Ns, j, tw.data(), T, twstep, inv_ns);
        else fft_butterfly<false>(pa, pb, n, R, Ns, j, tw.data(), T, twstep, inv_ns);
      }
      std::swap(pa, pb);
      Ns *= R;
    }
    std::memcpy(out + (size_t)f * n * 2, pa, sizeof(float2) * n);
  }
  return plan.n_stages;
}

// taps of one level for n points: idx[c][i] (uint32), w[c][i]
extern "C" int hostcheck_taps(const immoco_grid_desc* g, int level, const float* coords, int n,
                              uint32_t* idx, float* w) {
  const int D = g->n_dims;
  for (int i = 0; i < n; ++i) {
    uint32_t cell[3];
    float frac[3];
    for (int d = 0; d < D; ++d) grid_pos(coords[(size_t)i * D + d], g->scale[level], cell[d], frac[d]);
    for (int c = 0; c < (1 << D); ++c) {
      float ww = 1.0f;
      uint32_t q2[2], q3[3];
      for (int d = 0; d < D; ++d) {
        const int bit = (c >> d) & 1;
        const uint32_t q = cell[d] + (uint32_t)bit;
        if (D == 2) q2[d] = q; else q3[d] = q;
        ww = (d == 0) ? (bit ? frac[0] : 1.0f - frac[0]) : ww * (bit ? frac[d] : 1.0f - frac[d]);
      }
      const uint32_t id = (D == 2)
          ? grid_index<2>(q2, g->hashed[level], g->entries[level], g->resolution[level], g->swizzle[level])
          : grid_index<3>(q3, g->hashed[level], g->entries[level], g->resolution[level], g->swizzle[level]);
      idx[(size_t)c * n + i] = id;
      w[(size_t)c * n + i] = ww;
    }
  }
  return 0;
}
