// INR MLPs: out = W2[:2] . act(W1 . enc), one hidden layer, no biases.
// Replaces the network half of tcnn.NetworkWithInputEncoding (src/models/immoco.py:11-25,60-65).
//
// TEST-SIDE CHECKER (built into tests/checkers/_mlp_simt.so by __graft_entry__.build(); NOT part of
// libimmoco_b200.so): fp32 SIMT implementation, the exact-fp32 A/B anchor of the tcgen05 kernels.  A CTA owns tiles of 128 points; the hidden
// layer is processed in chunks of 64 neurons so the 64-wide motion MLP and the 256-wide image MLP
// share one code path.  Weights live in shared memory for the whole (persistent) CTA lifetime.
//
//   forward : acc[8pt][4n] register tiles over k=32, activation, layer-2 dot, 16-lane shuffle reduce
//   backward: recompute hidden chunk -> dh (smem) -> dE += dh.W1 (register tile 4pt x 4k)
//             gW1 += dh^T.E (8 accumulators / thread / chunk, kept across tiles), gW2 likewise;
//             weight gradients leave the CTA once, at the end (smem reduce + global atomics).
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kTile = 128;          // points per tile
constexpr int kIn = 32;             // encoded features (16 levels x 2)
constexpr int kSE = kTile + 4;      // row stride of Es[k][p]
constexpr int kSD = 64 + 4;         // row stride of dhs[p][n_in_chunk]

template <int WIDTH>
struct Layout {
  static constexpr int SW = WIDTH + 4;             // row stride of W1t[k][n]
  static constexpr int off_es = 0;
  static constexpr int off_w1t = off_es + kIn * kSE;
  static constexpr int off_w2 = off_w1t + kIn * SW;
  static constexpr int fwd_floats = off_w2 + 2 * WIDTH;
  static constexpr int off_dh = fwd_floats;
  static constexpr int off_do = off_dh + kTile * kSD;
  static constexpr int bwd_floats = off_do + kTile * 2;
};

template <int ACT>
__device__ __forceinline__ float act_f(float x) {
  if (ACT == IMMOCO_ACT_RELU) return fmaxf(x, 0.0f);
  if (ACT == IMMOCO_ACT_TANH) return tanhf(x);
  return x;
}
template <int ACT>
__device__ __forceinline__ float act_g(float y) {
  if (ACT == IMMOCO_ACT_RELU) return y > 0.0f ? 1.0f : 0.0f;
  if (ACT == IMMOCO_ACT_TANH) return 1.0f - y * y;
  return 1.0f;
}

template <int WIDTH>
__device__ __forceinline__ void load_weights(float* smem, const float* __restrict__ w1,
                                             const float* __restrict__ w2) {
  using L = Layout<WIDTH>;
  float* W1t = smem + L::off_w1t;
  float* W2s = smem + L::off_w2;
  for (int idx = threadIdx.x; idx < WIDTH * kIn; idx += kThreads) {
    const int nrn = idx >> 5, k = idx & 31;     // W1 is (WIDTH x 32) row-major
    W1t[k * L::SW + nrn] = __ldg(w1 + idx);
  }
  for (int idx = threadIdx.x; idx < 2 * WIDTH; idx += kThreads) W2s[idx] = __ldg(w2 + idx);
}

__device__ __forceinline__ void load_enc_tile(float* Es, const float2* __restrict__ enc, int n, int p0) {
  for (int idx = threadIdx.x; idx < 16 * kTile; idx += kThreads) {
    const int l = idx >> 7, p = idx & (kTile - 1);
    float2 v = make_float2(0.f, 0.f);
    if (p0 + p < n) v = __ldg(enc + (size_t)l * n + p0 + p);
    Es[(2 * l) * kSE + p] = v.x;
    Es[(2 * l + 1) * kSE + p] = v.y;
  }
}

// acc[i][j] = sum_k Es[k][8tp+i] * W1t[k][c*64+4tn+j]
template <int WIDTH>
__device__ __forceinline__ void hidden_chunk(const float* Es, const float* W1t, int c, int tp, int tn,
                                             float (&acc)[8][4]) {
  using L = Layout<WIDTH>;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
#pragma unroll 8
  for (int k = 0; k < kIn; ++k) {
    const float4 e0 = *reinterpret_cast<const float4*>(Es + k * kSE + 8 * tp);
    const float4 e1 = *reinterpret_cast<const float4*>(Es + k * kSE + 8 * tp + 4);
    const float4 wv = *reinterpret_cast<const float4*>(W1t + k * L::SW + c * 64 + 4 * tn);
    const float e[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
    const float w[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(e[i], w[j], acc[i][j]);
  }
}

template <int WIDTH, int ACT>
__global__ void __launch_bounds__(kThreads)
mlp_fwd_kernel(const float2* __restrict__ enc, const float* __restrict__ w1,
               const float* __restrict__ w2, float* __restrict__ out, int n, int out_tanh) {
  using L = Layout<WIDTH>;
  extern __shared__ __align__(16) float smem[];
  float* Es = smem + L::off_es;
  const float* W1t = smem + L::off_w1t;
  const float* W2s = smem + L::off_w2;
  load_weights<WIDTH>(smem, w1, w2);

  const int tn = threadIdx.x & 15, tp = threadIdx.x >> 4;
  const int n_tiles = (n + kTile - 1) / kTile;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int p0 = tile * kTile;
    __syncthreads();
    load_enc_tile(Es, enc, n, p0);
    __syncthreads();
    float o0[8], o1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o0[i] = o1[i] = 0.0f;
#pragma unroll 1
    for (int c = 0; c < WIDTH / 64; ++c) {
      float acc[8][4];
      hidden_chunk<WIDTH>(Es, W1t, c, tp, tn, acc);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float w20 = W2s[c * 64 + 4 * tn + j];
        const float w21 = W2s[WIDTH + c * 64 + 4 * tn + j];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float h = act_f<ACT>(acc[i][j]);
          o0[i] = fmaf(h, w20, o0[i]);
          o1[i] = fmaf(h, w21, o1[i]);
        }
      }
    }
    // reduce over the 16 lanes that share the same 8 points
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        o0[i] += __shfl_xor_sync(0xffffffffu, o0[i], o);
        o1[i] += __shfl_xor_sync(0xffffffffu, o1[i], o);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int p = p0 + 8 * tp + i;
      if (p < n) {
        if (tn == 2 * i) out[(size_t)p * 2] = out_tanh ? tanhf(o0[i]) : o0[i];
        if (tn == 2 * i + 1) out[(size_t)p * 2 + 1] = out_tanh ? tanhf(o1[i]) : o1[i];
      }
    }
  }
}

template <int WIDTH, int ACT>
__global__ void __launch_bounds__(kThreads, 1)
mlp_bwd_kernel(const float2* __restrict__ enc, const float* __restrict__ w1,
               const float* __restrict__ w2, const float* __restrict__ d_out,
               float2* __restrict__ d_enc, float* __restrict__ g_w1, float* __restrict__ g_w2, int n) {
  using L = Layout<WIDTH>;
  constexpr int C = WIDTH / 64;
  extern __shared__ __align__(16) float smem[];
  float* Es = smem + L::off_es;
  const float* W1t = smem + L::off_w1t;
  const float* W2s = smem + L::off_w2;
  float* dhs = smem + L::off_dh;
  float* dos = smem + L::off_do;
  load_weights<WIDTH>(smem, w1, w2);

  const int tid = threadIdx.x;
  const int tn = tid & 15, tp = tid >> 4;      // hidden tile: points 8tp+i, neurons c*64+4tn+j
  const int tk = tid & 7, tpp = tid >> 3;      // dE tile    : points 4tpp+i, features tk+8j
  const int nq = tid & 31, kq = tid >> 5;      // gW1 tile   : neurons c*64+2nq+{0,1}, features 4kq+kk

  float gw1[C][8];
  float gw2[C][2][4];
#pragma unroll
  for (int c = 0; c < C; ++c) {
#pragma unroll
    for (int x = 0; x < 8; ++x) gw1[c][x] = 0.0f;
#pragma unroll
    for (int j = 0; j < 4; ++j) gw2[c][0][j] = gw2[c][1][j] = 0.0f;
  }

  const int n_tiles = (n + kTile - 1) / kTile;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int p0 = tile * kTile;
    __syncthreads();
    load_enc_tile(Es, enc, n, p0);
    if (tid < kTile * 2) {
      const int p = p0 + (tid >> 1);
      dos[tid] = (p < n) ? __ldg(d_out + (size_t)p * 2 + (tid & 1)) : 0.0f;
    }
    __syncthreads();

    float dE[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dE[i][j] = 0.0f;

    float do0[8], do1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      do0[i] = dos[(8 * tp + i) * 2];
      do1[i] = dos[(8 * tp + i) * 2 + 1];
    }

#pragma unroll
    for (int c = 0; c < C; ++c) {
      {
        float acc[8][4];
        hidden_chunk<WIDTH>(Es, W1t, c, tp, tn, acc);
        float w20[4], w21[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          w20[j] = W2s[c * 64 + 4 * tn + j];
          w21[j] = W2s[WIDTH + c * 64 + 4 * tn + j];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float dh[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float h = act_f<ACT>(acc[i][j]);
            gw2[c][0][j] = fmaf(do0[i], h, gw2[c][0][j]);
            gw2[c][1][j] = fmaf(do1[i], h, gw2[c][1][j]);
            dh[j] = act_g<ACT>(h) * fmaf(w20[j], do0[i], w21[j] * do1[i]);
          }
          *reinterpret_cast<float4*>(dhs + (8 * tp + i) * kSD + 4 * tn) =
              make_float4(dh[0], dh[1], dh[2], dh[3]);
        }
      }
      __syncthreads();
      // dE[p][k] += sum_n dh[p][n] * W1[n][k]
#pragma unroll 4
      for (int n4 = 0; n4 < 64; n4 += 4) {
        float4 a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(dhs + (4 * tpp + i) * kSD + n4);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          b[j] = *reinterpret_cast<const float4*>(W1t + (tk + 8 * j) * L::SW + c * 64 + n4);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            dE[i][j] = fmaf(a[i].x, b[j].x, dE[i][j]);
            dE[i][j] = fmaf(a[i].y, b[j].y, dE[i][j]);
            dE[i][j] = fmaf(a[i].z, b[j].z, dE[i][j]);
            dE[i][j] = fmaf(a[i].w, b[j].w, dE[i][j]);
          }
      }
      // gW1[n][k] += sum_p dh[p][n] * E[p][k]
#pragma unroll 4
      for (int p4 = 0; p4 < kTile; p4 += 4) {
        float4 ev[4];
        float2 dv[4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) ev[kk] = *reinterpret_cast<const float4*>(Es + (4 * kq + kk) * kSE + p4);
#pragma unroll
        for (int pp = 0; pp < 4; ++pp) dv[pp] = *reinterpret_cast<const float2*>(dhs + (p4 + pp) * kSD + 2 * nq);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          gw1[c][kk] = fmaf(dv[0].x, ev[kk].x, gw1[c][kk]);
          gw1[c][kk] = fmaf(dv[1].x, ev[kk].y, gw1[c][kk]);
          gw1[c][kk] = fmaf(dv[2].x, ev[kk].z, gw1[c][kk]);
          gw1[c][kk] = fmaf(dv[3].x, ev[kk].w, gw1[c][kk]);
          gw1[c][4 + kk] = fmaf(dv[0].y, ev[kk].x, gw1[c][4 + kk]);
          gw1[c][4 + kk] = fmaf(dv[1].y, ev[kk].y, gw1[c][4 + kk]);
          gw1[c][4 + kk] = fmaf(dv[2].y, ev[kk].z, gw1[c][4 + kk]);
          gw1[c][4 + kk] = fmaf(dv[3].y, ev[kk].w, gw1[c][4 + kk]);
        }
      }
      __syncthreads();  // dhs (and, after the last chunk, Es) may be overwritten
    }
    // stage dE through Es for a coalesced plane write
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) Es[(tk + 8 * j) * kSE + 4 * tpp + i] = dE[i][j];
    __syncthreads();
    for (int idx = tid; idx < 16 * kTile; idx += kThreads) {
      const int l = idx >> 7, p = idx & (kTile - 1);
      if (p0 + p < n)
        d_enc[(size_t)l * n + p0 + p] = make_float2(Es[(2 * l) * kSE + p], Es[(2 * l + 1) * kSE + p]);
    }
  }

  // ---- weight gradients leave the CTA once -------------------------------------------------
  __syncthreads();
  float* red = dhs;  // [2][WIDTH] reduction buffer for gW2
  for (int idx = tid; idx < 2 * WIDTH; idx += kThreads) red[idx] = 0.0f;
  __syncthreads();
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      atomicAdd(red + c * 64 + 4 * tn + j, gw2[c][0][j]);
      atomicAdd(red + WIDTH + c * 64 + 4 * tn + j, gw2[c][1][j]);
    }
  __syncthreads();
  for (int idx = tid; idx < 2 * WIDTH; idx += kThreads) atomicAdd(g_w2 + idx, red[idx]);
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int nn = 0; nn < 2; ++nn)
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        atomicAdd(g_w1 + (size_t)(c * 64 + 2 * nq + nn) * kIn + 4 * kq + kk, gw1[c][nn * 4 + kk]);
}

template <typename K>
int resident_ctas(K kernel, int smem_bytes) {
  int per_sm = 1;
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, smem_bytes) != cudaSuccess || per_sm < 1)
    per_sm = 1;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return per_sm * sms;
}

template <int WIDTH, int ACT>
int launch_fwd(const float* enc, const float* w1, const float* w2, float* out, int n, int out_tanh,
               cudaStream_t s) {
  constexpr int smem = Layout<WIDTH>::fwd_floats * 4;
  static int ctas = resident_ctas(mlp_fwd_kernel<WIDTH, ACT>, smem);
  const int n_tiles = (n + kTile - 1) / kTile;
  const int grid = n_tiles < ctas ? n_tiles : ctas;
  mlp_fwd_kernel<WIDTH, ACT><<<grid, kThreads, smem, s>>>((const float2*)enc, w1, w2, out, n, out_tanh);
  IMMOCO_LAUNCH_CHECK();
  return 0;
}

template <int WIDTH, int ACT>
int launch_bwd(const float* enc, const float* w1, const float* w2, const float* d_out, float* d_enc,
               float* g_w1, float* g_w2, int n, cudaStream_t s) {
  constexpr int smem = Layout<WIDTH>::bwd_floats * 4;
  static int ctas = resident_ctas(mlp_bwd_kernel<WIDTH, ACT>, smem);
  const int n_tiles = (n + kTile - 1) / kTile;
  const int grid = n_tiles < ctas ? n_tiles : ctas;
  mlp_bwd_kernel<WIDTH, ACT><<<grid, kThreads, smem, s>>>((const float2*)enc, w1, w2, d_out,
                                                         (float2*)d_enc, g_w1, g_w2, n);
  IMMOCO_LAUNCH_CHECK();
  return 0;
}

}  // namespace

// ---- test-side entry points (same contracts as immoco_mlp_fwd / immoco_mlp_bwd of the product library) ----
extern "C" int immoco_simt_mlp_fwd(const float* enc, const float* w1, const float* w2, float* out,
                                   int64_t n_points, int32_t width, int32_t act, int32_t out_tanh,
                                   void* stream) {
  if (n_points < 0 || n_points > 0x3fffffff) return IMMOCO_ERR_BAD_ARG;
  if (n_points == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const int n = (int)n_points;
  if (width == 256 && act == IMMOCO_ACT_RELU) return launch_fwd<256, IMMOCO_ACT_RELU>(enc, w1, w2, out, n, out_tanh, s);
  if (width == 256 && act == IMMOCO_ACT_TANH) return launch_fwd<256, IMMOCO_ACT_TANH>(enc, w1, w2, out, n, out_tanh, s);
  if (width == 64 && act == IMMOCO_ACT_RELU) return launch_fwd<64, IMMOCO_ACT_RELU>(enc, w1, w2, out, n, out_tanh, s);
  if (width == 64 && act == IMMOCO_ACT_TANH) return launch_fwd<64, IMMOCO_ACT_TANH>(enc, w1, w2, out, n, out_tanh, s);
  return IMMOCO_ERR_UNSUPPORTED;
}

extern "C" int immoco_simt_mlp_bwd(const float* enc, const float* w1, const float* w2, const float* d_out,
                                   float* d_enc, float* g_w1, float* g_w2, int64_t n_points,
                                   int32_t width, int32_t act, void* stream) {
  if (n_points < 0 || n_points > 0x3fffffff) return IMMOCO_ERR_BAD_ARG;
  if (n_points == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const int n = (int)n_points;
  if (width == 256 && act == IMMOCO_ACT_RELU) return launch_bwd<256, IMMOCO_ACT_RELU>(enc, w1, w2, d_out, d_enc, g_w1, g_w2, n, s);
  if (width == 256 && act == IMMOCO_ACT_TANH) return launch_bwd<256, IMMOCO_ACT_TANH>(enc, w1, w2, d_out, d_enc, g_w1, g_w2, n, s);
  if (width == 64 && act == IMMOCO_ACT_RELU) return launch_bwd<64, IMMOCO_ACT_RELU>(enc, w1, w2, d_out, d_enc, g_w1, g_w2, n, s);
  if (width == 64 && act == IMMOCO_ACT_TANH) return launch_bwd<64, IMMOCO_ACT_TANH>(enc, w1, w2, d_out, d_enc, g_w1, g_w2, n, s);
  return IMMOCO_ERR_UNSUPPORTED;
}
