// Fused Adam and the per-iteration driver of the IM-MoCo fit
// (src/models/immoco.py:149-154 optimizer, :164-181 loop).
#include <math.h>

#include <mutex>

#include "common.cuh"
#include "fit_batch.cuh"

// launch helpers implemented in forward_model.cu
int immoco_rows_static(const float* in, float* out, int h, int w, const float* tw_w,
                       const float* in_w, const float* out_w, int accumulate, bool inv, void* stream);
int immoco_motion_rows_fwd(const float* image, const float* disp, const float* ident,
                           const immoco_lines* lines, const float* tw_w, float* c_out, int h, int w,
                           void* stream);
int immoco_motion_rows_bwd(const float* d_c, const float* image, const float* disp, const float* ident,
                           const immoco_lines* lines, const float* tw_w, float* d_image, float* d_disp,
                           int h, int w, void* stream);
int immoco_rows_fwd_fused_batch(const FitBatch& b, const float* ident, const float* tw_w, int h, int w, void* stream);
int immoco_rows_bwd_fused_batch(const FitBatch& b, const float* ident, const float* tw_w, int h, int w, bool det,
                                void* stream);
int immoco_colpass_loss_batch(const FitBatch& b, const float* tw_h, int h, int w, int zero_input, void* stream);
int immoco_colpass_loss_slots(const float* c, const float* k_in, float* k_out, float* d_c, double* loss_acc,
                              const float* tw_h, int h, int w, double* loss_slots, uint32_t* dmax_bits, void* stream);
int immoco_grad_entropy_batch(const FitBatch& b, float grad_scale, int accumulate, int h, int w, void* stream);
int immoco_d_image_finalize_batch(const FitBatch& b, int h, int w, void* stream);
// mlp_tc.cu
int immoco_mlp_bwd_tc_grid(int64_t n_points);
int immoco_mlp_fwd_tc_batch(const MlpFwdBatch& b, int64_t n_points, int32_t width, int32_t act, int32_t out_tanh,
                            void* stream);

namespace {

// torch.optim.Adam (amsgrad=False, weight_decay=0, maximize=False), single-tensor formulation:
//   m = lerp(m, g, 1-b1);  v = b2*v + (1-b2)*g*g;  p -= step_size * m / (sqrt(v)/bc2_sqrt + eps)
// One pass: reads p,g,m,v, writes p,m,v and zeroes g (28 B / parameter), 128-bit accesses.
#define IMMOCO_ADAM_LANE(c) adam_update(pi.c, mi.c, vi.c, gi.c, one_minus_b1, b2, one_minus_b2, step_size, bc2_sqrt, eps);

// U independent 128-bit items per thread per trip: all 4*U loads are issued before the first use so
// 64*U bytes per thread are in flight.  HINTS: gradients and moments are touched once per iteration
// -> streaming (evict-first) loads/stores; the parameters are written with the default policy so the
// tables can still be L2-resident when the next iteration's hash-grid gathers start.
template <int U, bool HINTS>
__global__ void __launch_bounds__(256)
adam_kernel(float4* __restrict__ p, float4* __restrict__ g, float4* __restrict__ m,
            float4* __restrict__ v, int64_t n4, float one_minus_b1, float b2, float one_minus_b2,
            float step_size, float bc2_sqrt, float eps, int zero_grad) {
  pdl_wait();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += stride * U) {
    float4 ga[U], ma[U], va[U], pa[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < n4) {
        if (HINTS) { ga[u] = __ldcs(g + i); ma[u] = __ldcs(m + i); va[u] = __ldcs(v + i); }
        else { ga[u] = g[i]; ma[u] = m[i]; va[u] = v[i]; }
        pa[u] = p[i];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < n4) {
        const float4 gi = ga[u];
        float4 mi = ma[u], vi = va[u], pi = pa[u];
        IMMOCO_ADAM_LANE(x) IMMOCO_ADAM_LANE(y) IMMOCO_ADAM_LANE(z) IMMOCO_ADAM_LANE(w)
        if (HINTS) { __stcs(m + i, mi); __stcs(v + i, vi); }
        else { m[i] = mi; v[i] = vi; }
        p[i] = pi;
        if (zero_grad) {
          if (HINTS) __stcs(g + i, make_float4(0.f, 0.f, 0.f, 0.f));
          else g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
  }
}

__global__ void adam_tail_kernel(float* p, float* g, float* m, float* v, int64_t begin, int64_t n,
                                 float one_minus_b1, float b2, float one_minus_b2, float step_size,
                                 float bc2_sqrt, float eps, int zero_grad) {
  pdl_wait();
  const int64_t i = begin + threadIdx.x;
  if (i < n) {
    const float gi = g[i];
    float pi = p[i], mi = m[i], vi = v[i];
    adam_update(pi, mi, vi, gi, one_minus_b1, b2, one_minus_b2, step_size, bc2_sqrt, eps);
    p[i] = pi; m[i] = mi; v[i] = vi;
    if (zero_grad) g[i] = 0.f;
  }
}

// Adam over an INR's MLP block with the gradient assembled from the backward kernel's per-CTA partial blocks
// (deterministic mode).  One WARP per 128-bit item: lane l adds blocks l, l + 32, ... in order, then a fixed
// butterfly combines the 32 lane sums -- a fixed summation tree whose depth is 5 loads + 5 shuffles instead of
// a 148-long dependent chain (a thread per item took 26 us for 3 K parameters; this takes a few).
__global__ void __launch_bounds__(256)
adam_partials_kernel(float4* __restrict__ p, const float4* __restrict__ part, int n_part, float4* __restrict__ m,
                     float4* __restrict__ v, int n4, float one_minus_b1, float b2, float one_minus_b2,
                     float step_size, float bc2_sqrt, float eps) {
  pdl_wait();
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= n4) return;
  float4 gi = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int c = lane; c < n_part; c += 32) {
    const float4 a = __ldg(part + (size_t)c * n4 + i);
    gi.x += a.x; gi.y += a.y; gi.z += a.z; gi.w += a.w;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    gi.x += __shfl_xor_sync(0xffffffffu, gi.x, o);
    gi.y += __shfl_xor_sync(0xffffffffu, gi.y, o);
    gi.z += __shfl_xor_sync(0xffffffffu, gi.z, o);
    gi.w += __shfl_xor_sync(0xffffffffu, gi.w, o);
  }
  if (lane != 0) return;
  float4 mi = m[i], vi = v[i], pi = p[i];
  IMMOCO_ADAM_LANE(x) IMMOCO_ADAM_LANE(y) IMMOCO_ADAM_LANE(z) IMMOCO_ADAM_LANE(w)
  m[i] = mi; v[i] = vi; p[i] = pi;
}

// loss[it] = sum of the iteration's per-CTA slots, added in slot order (immoco_fit::loss_slots)
__global__ void loss_reduce_kernel(const double* __restrict__ slots, int per_iter, int n_dc, double* __restrict__ loss,
                                   int it_begin, int it_end) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int it = it_begin + (t >> 1), which = t & 1;
  if (it >= it_end) return;
  const double* s = slots + (size_t)it * per_iter + (which ? n_dc : 0);
  const int n = which ? per_iter - n_dc : n_dc;
  double acc = 0.0;
  for (int k = 0; k < n; ++k) acc += s[k];
  loss[2 * (size_t)it + which] = acc;
}

}  // namespace
#undef IMMOCO_ADAM_LANE

// SM count of the current device, cached per device
int immoco_num_sms() {
  static int cache[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cache[dev] == 0) {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1) sms = 148;
    cache[dev] = sms;
  }
  return cache[dev];
}

AdamScalars adam_scalars(double lr, double beta1, double beta2, double eps, int step) {
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  AdamScalars a;
  a.step_size = (float)(lr / bc1);
  a.bc2_sqrt = (float)sqrt(bc2);
  a.omb1 = (float)(1.0 - beta1);
  a.omb2 = (float)(1.0 - beta2);
  a.b2 = (float)beta2;
  a.eps = (float)eps;
  return a;
}

// 1: the float-atomic path reduces the motion grid's hashed-level gradients inside the 64-wide MLP backward
// kernel (immoco_mlp_bwd_scatter); 0: separate scatter kernel over the feature planes.  Default set from the
// measurement in profiles/round2_fused_scatter.txt
static int g_fused_scatter = 0;
extern "C" int immoco_set_fused_scatter(int32_t on) { g_fused_scatter = on ? 1 : 0; return 0; }
extern "C" int immoco_get_fused_scatter(void) { return g_fused_scatter; }

static int g_deterministic = 0;
extern "C" int immoco_set_deterministic(int32_t on) { g_deterministic = on ? 1 : 0; return 0; }
extern "C" int immoco_get_deterministic(void) { return g_deterministic; }

static int g_pdl = 1;
int immoco_pdl_enabled() { return g_pdl; }
// 1 (default): hot-path kernels are launched with programmatic stream serialization (common.cuh)
extern "C" int immoco_set_pdl(int32_t on) { g_pdl = on ? 1 : 0; return 0; }

// default: one 128-bit item per thread, streaming (evict-first) gradients / moments -- measured inside the iteration
// (tools/adam_iter_ab.py): variant 0: 629.7, 3: 627.7, 1: 661, 4: 654, 2 / 5: 680+ us
static int g_adam_variant = 3, g_adam_ctas_per_sm = 32;
// tuning knobs (tools/adam_bench.py): variant = {U=1,2,4} x {plain, streaming hints}; CTAs per SM
extern "C" int immoco_set_adam_tuning(int32_t variant, int32_t ctas_per_sm) {
  if (variant < 0 || variant > 5 || ctas_per_sm < 1 || ctas_per_sm > 64) return IMMOCO_ERR_BAD_ARG;
  g_adam_variant = variant;
  g_adam_ctas_per_sm = ctas_per_sm;
  return 0;
}

extern "C" int immoco_adam_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq,
                                int64_t n, double lr, double beta1, double beta2, double eps,
                                int32_t step, int32_t zero_grad, void* stream) {
  if (n < 0 || step < 1) return IMMOCO_ERR_BAD_ARG;
  if (n == 0) return 0;
  if ((((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) != 0)
    return IMMOCO_ERR_BAD_ARG;
  const AdamScalars sc = adam_scalars(lr, beta1, beta2, eps, step);
  const float step_size = sc.step_size, bc2_sqrt = sc.bc2_sqrt, omb1 = sc.omb1, omb2 = sc.omb2;
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t n4 = n / 4;
  if (n4 > 0) {
    int64_t blocks = (n4 + 255) / 256;
    const int64_t cap = (int64_t)IMMOCO_NUM_SMS * g_adam_ctas_per_sm;
    if (blocks > cap) blocks = cap;
#define IMMOCO_ADAM_LAUNCH(U, H)                                                                          \
  immoco_launch(adam_kernel<U, H>, dim3((unsigned)blocks), dim3(256), 0, s, (float4*)params, (float4*)grads, (float4*)exp_avg,   \
                                                     (float4*)exp_avg_sq, n4, omb1, (float)beta2, omb2,   \
                                                     step_size, bc2_sqrt, (float)eps, zero_grad)
    switch (g_adam_variant) {
      case 0: IMMOCO_ADAM_LAUNCH(1, false); break;
      case 1: IMMOCO_ADAM_LAUNCH(2, false); break;
      case 2: IMMOCO_ADAM_LAUNCH(4, false); break;
      case 3: IMMOCO_ADAM_LAUNCH(1, true); break;
      case 4: IMMOCO_ADAM_LAUNCH(2, true); break;
      default: IMMOCO_ADAM_LAUNCH(4, true); break;
    }
#undef IMMOCO_ADAM_LAUNCH
    IMMOCO_LAUNCH_CHECK();
  }
  if (n4 * 4 < n) {
    immoco_launch(adam_tail_kernel, dim3(1), dim3(4), 0, s, params, grads, exp_avg, exp_avg_sq, n4 * 4, n, omb1, (float)beta2, omb2,
                                     step_size, bc2_sqrt, (float)eps, zero_grad);
    IMMOCO_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int immoco_adam_step_partials(float* params, const float* g_part, int32_t n_part, float* exp_avg,
                                         float* exp_avg_sq, int64_t n_mlp, double lr, double beta1, double beta2,
                                         double eps, int32_t step, void* stream) {
  if (n_mlp < 0 || (n_mlp & 3) != 0 || n_part < 0 || step < 1 || n_mlp > 0x7fffffff) return IMMOCO_ERR_BAD_ARG;
  if (n_mlp == 0) return 0;
  if ((((uintptr_t)params | (uintptr_t)g_part | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) != 0)
    return IMMOCO_ERR_BAD_ARG;
  const AdamScalars sc = adam_scalars(lr, beta1, beta2, eps, step);
  const int n4 = (int)(n_mlp / 4);
  immoco_launch(adam_partials_kernel, dim3((n4 + 7) / 8), dim3(256), 0, (cudaStream_t)stream, (float4*)params,
                (const float4*)g_part, (int)n_part, (float4*)exp_avg, (float4*)exp_avg_sq, n4, sc.omb1, sc.b2, sc.omb2,
                sc.step_size, sc.bc2_sqrt, sc.eps);
  IMMOCO_LAUNCH_CHECK();
  return 0;
}

extern "C" int immoco_abi_version(void) { return 3; }

extern "C" void immoco_struct_sizes(int32_t out[5]) {
  out[0] = (int32_t)sizeof(immoco_grid_desc);
  out[1] = (int32_t)sizeof(immoco_lines);
  out[2] = (int32_t)sizeof(immoco_fit);
  out[3] = (int32_t)sizeof(immoco_grid_csr);
  out[4] = (int32_t)sizeof(immoco_grid_taps);
}

// launches per iteration.  Float-atomic path: hashgrid fwd + mlp fwd (x2), gradient entropy, fused rows
// (static + motion groups), colpass, fused adjoint rows, mlp bwd + hashgrid bwd (dense levels + hashed levels:
// 2 launches) (x2), adam (x2) = 16.  Deterministic path: the hash-grid backward is ONE gather launch per INR,
// plus the fixed-point finalize and one partial-sum Adam per INR = 17; with the table update fused into the
// gather the two table-Adam launches go: 15.
static int g_fuse_rows = 1;
extern "C" int immoco_launches_per_iteration(int32_t m) { return m > 0 ? (g_fuse_rows ? 16 : 18) : 10; }
extern "C" int immoco_launches_per_iteration_mode(int32_t m, int32_t deterministic, int32_t fuse_adam) {
  if (!deterministic) return immoco_launches_per_iteration(m);
  const int per_inr = fuse_adam ? 5 : 6;     // hashgrid fwd, mlp fwd, mlp bwd, gather(+adam), adam partials(, adam table)
  return (m > 0 ? 2 : 1) * per_inr + 1 /*GE*/ + 3 /*rows fwd, colpass, rows bwd*/ + 1 /*finalize*/;
}

// 1 (default): the static row pass and the pruned motion rows of an iteration are ONE launch (forward) and
// ONE launch (adjoint); 0: the four separate launches (A/B check)
extern "C" int immoco_set_fused_rows(int32_t on) { g_fuse_rows = on ? 1 : 0; return 0; }

#define IMMOCO_TRY(expr)          \
  do {                            \
    int e__ = (expr);             \
    if (e__ != 0) return e__;     \
  } while (0)

// ---- optional per-kernel timing: a begin and an end CUDA event around every kernel of selected
//      iterations, recorded on the stream the kernel is launched on; nothing synchronises until
//      immoco_profile_read() / immoco_profile_timeline() ------------------------------------------
struct immoco_profile {
  int n_slots;                 // kernels per iteration
  int capacity;                // instrumented iterations the event pool holds
  int used;
  cudaEvent_t* ev;             // capacity * (2 * n_slots + 1) events: base, then (begin, end) per slot
};
static inline int prof_stride(const immoco_profile* p) { return 2 * p->n_slots + 1; }

extern "C" immoco_profile* immoco_profile_create(int32_t capacity) {
  if (capacity < 1) return nullptr;
  immoco_profile* p = new immoco_profile;
  p->n_slots = IMMOCO_PROFILE_SLOTS;
  p->capacity = capacity;
  p->used = 0;
  const int n = capacity * prof_stride(p);
  p->ev = new cudaEvent_t[n];
  for (int i = 0; i < n; ++i) {
    if (cudaEventCreate(&p->ev[i]) != cudaSuccess) { p->capacity = 0; break; }
  }
  return p;
}

extern "C" void immoco_profile_destroy(immoco_profile* p) {
  if (!p) return;
  for (int i = 0; i < p->capacity * prof_stride(p); ++i) cudaEventDestroy(p->ev[i]);
  delete[] p->ev;
  delete p;
}

// Sums per-slot milliseconds over the instrumented iterations into ms_sum[IMMOCO_PROFILE_SLOTS];
// returns the number of iterations summed (the caller must have synchronised the stream).
// Slots that launched nothing in an iteration (n_M = 0) were still bracketed by their two events.
extern "C" int immoco_profile_read(immoco_profile* p, float* ms_sum) {
  if (!p || !ms_sum) return IMMOCO_ERR_BAD_ARG;
  for (int k = 0; k < p->n_slots; ++k) ms_sum[k] = 0.f;
  for (int i = 0; i < p->used; ++i) {
    cudaEvent_t* e = p->ev + (size_t)i * prof_stride(p) + 1;
    for (int k = 0; k < p->n_slots; ++k) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, e[2 * k], e[2 * k + 1]) != cudaSuccess) return IMMOCO_ERR_BAD_ARG;
      ms_sum[k] += ms;
    }
  }
  const int n = p->used;
  p->used = 0;
  return n;
}

// begin / end of every slot of instrumented iteration `index`, in ms after that iteration's first
// event (a timeline across both streams); does not consume the samples.
extern "C" int immoco_profile_timeline(immoco_profile* p, int32_t index, float* begin_ms, float* end_ms) {
  if (!p || !begin_ms || !end_ms || index < 0 || index >= p->used) return IMMOCO_ERR_BAD_ARG;
  cudaEvent_t* base = p->ev + (size_t)index * prof_stride(p);
  for (int k = 0; k < p->n_slots; ++k) {
    if (cudaEventElapsedTime(&begin_ms[k], base[0], base[1 + 2 * k]) != cudaSuccess) return IMMOCO_ERR_BAD_ARG;
    if (cudaEventElapsedTime(&end_ms[k], base[0], base[2 + 2 * k]) != cudaSuccess) return IMMOCO_ERR_BAD_ARG;
  }
  return 0;
}

// ---- image-INR branch on a second stream --------------------------------------------------------
// The two INR branches of an iteration are independent and stress different units
// (hash-grid gathers / reductions: L2; MLPs: tensor pipe + SIMT epilogue; Adam: HBM).  The image
// branch runs on an auxiliary non-blocking stream.  Schedule of one iteration (=> same-stream order,
// -> event dependency):
//   aux : [Adam_i(it-1)] => HGi_f => MLPi_f => GE  ........................  MLPi_b => HGi_b => Adam_i
//   main: [Adam_m(it-1)] => HGm_f => MLPm_f =>(join_fwd)=> rows => motion_rows => colpass => rows_adj
//          => motion_rows_bwd => MLPm_b =>(mlp_done -> aux)=> HGm_b => Adam_m
// MLPi_b is released only after MLPm_b: both want every SM's tensor memory, and the motion chain is
// the critical path.  The whole image branch (bwd of iteration it, update, fwd of it+1) then hides
// behind HGm_b / Adam_m / HGm_f of the motion chain.
// One aux stream + its events per (device, caller stream) are created on first use and live for the
// process (the library's only hidden state), so concurrent fits on different streams do not share.
struct AuxStream {
  int dev = -1;
  cudaStream_t owner = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join_fwd = nullptr, mlp_done = nullptr, join_end = nullptr;
  // deferred gradient zeroing (see immoco_fit_run): a third stream and its events
  cudaStream_t zero_stream = nullptr;
  cudaEvent_t adam_i_done = nullptr, zero_start = nullptr, zero_done = nullptr;
  // deterministic mode: the adjoint row launch finished (the image branch converts the fixed-point plane)
  cudaEvent_t rows_done = nullptr;
};
static std::mutex g_aux_mutex;
static int g_aux_high_priority = 1;
static AuxStream g_aux_table[256];
static int g_aux_used = 0;
static void aux_destroy(AuxStream& a) {
  cudaEvent_t ev[8] = {a.fork, a.join_fwd, a.mlp_done, a.join_end, a.adam_i_done, a.zero_start, a.zero_done, a.rows_done};
  for (auto e : ev)
    if (e) cudaEventDestroy(e);
  if (a.stream) cudaStreamDestroy(a.stream);
  if (a.zero_stream) cudaStreamDestroy(a.zero_stream);
  a = AuxStream();
}
// frees the auxiliary stream sets of the CURRENT device (the caller guarantees no fit is in flight on them)
extern "C" int immoco_release_streams(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return IMMOCO_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lock(g_aux_mutex);
  int freed = 0, kept = 0;
  for (int i = 0; i < g_aux_used; ++i) {
    if (g_aux_table[i].dev == dev) {
      aux_destroy(g_aux_table[i]);
      ++freed;
    } else {
      if (kept != i) g_aux_table[kept] = g_aux_table[i];
      ++kept;
    }
  }
  for (int i = kept; i < g_aux_used; ++i) g_aux_table[i] = AuxStream();
  g_aux_used = kept;
  return freed;
}
static AuxStream* aux_for(cudaStream_t owner) {
  AuxStream* table = g_aux_table;
  int& used = g_aux_used;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lock(g_aux_mutex);
  for (int i = 0; i < used; ++i)
    if (table[i].dev == dev && table[i].owner == owner) return &table[i];
  if (used >= 256) return nullptr;
  AuxStream a;
  a.dev = dev;
  a.owner = owner;
  // the image branch is short and the motion chain waits on it at the forward join: give it the
  // higher scheduling priority so its CTAs are placed first whenever both streams have work queued
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  if (cudaStreamCreateWithPriority(&a.stream, cudaStreamNonBlocking, g_aux_high_priority ? prio_hi : prio_lo) !=
      cudaSuccess)
    return nullptr;
  if (cudaStreamCreateWithPriority(&a.zero_stream, cudaStreamNonBlocking, prio_lo) != cudaSuccess) return nullptr;
  cudaEvent_t* ev[8] = {&a.fork, &a.join_fwd, &a.mlp_done, &a.join_end, &a.adam_i_done, &a.zero_start, &a.zero_done,
                        &a.rows_done};
  for (auto e : ev)
    if (cudaEventCreateWithFlags(e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
  table[used] = a;
  return &table[used++];
}
static int g_overlap = 1, g_profile_overlap = 0;
// 1 (default): Adam leaves the gradients in place and ONE memset of the whole gradient vector runs on a third
// stream while the motion chain is in its SM-bound stretch (MLP forward, row / column passes), i.e. when the
// memory system is idle; 0: Adam zeroes the gradients itself (4 more bytes per parameter on the critical path)
static int g_deferred_zero = 1;
extern "C" int immoco_set_deferred_zero(int32_t on) { g_deferred_zero = on ? 1 : 0; return 0; }
extern "C" int immoco_set_branch_overlap(int32_t on) { g_overlap = on ? 1 : 0; return 0; }
// 1: instrumented iterations keep the two-stream schedule (timeline mode); 0 (default): they run
// serially on the caller's stream so per-kernel durations are contention-free.
extern "C" int immoco_set_profile_overlap(int32_t on) { g_profile_overlap = on ? 1 : 0; return 0; }

// K(slot, stream, launch-expression): brackets the launch(es) of one profile slot with its events
#define K(slot, strm, expr)                                                     \
  do {                                                                          \
    if (ev) cudaEventRecord(ev[1 + 2 * (slot)], (cudaStream_t)(strm));          \
    IMMOCO_TRY(expr);                                                           \
    if (ev) cudaEventRecord(ev[2 + 2 * (slot)], (cudaStream_t)(strm));          \
  } while (0)
static inline int nop() { return 0; }

// One batch of B fits of the same shape (h, w, m, network widths, grid descriptors), each with its own buffers,
// masks and parameters, advanced in lock step: the latency-bound kernels (row / column passes, gradient
// entropy, fixed-point finalize, MLP forward) are ONE launch over all instances (blockIdx.z / .y = instance),
// the kernels that fill the GPU by themselves (hash grid, MLP backward, Adam) are issued per instance.
static int fit_run_impl(const immoco_fit* const* fs, int B, int32_t it_begin, int32_t it_end,
                        const float* lambdas_host, void* stream, immoco_profile* prof, int32_t profile_every) {
  if (!fs || B < 1 || B > kMaxFitBatch || !lambdas_host || it_begin < 0 || it_end < it_begin) return IMMOCO_ERR_BAD_ARG;
  const immoco_fit* f0 = fs[0];
  if (!f0) return IMMOCO_ERR_BAD_ARG;
  const int H = f0->h, W = f0->w, M = f0->m;
  if (H < 2 || W < 2 || M < 0) return IMMOCO_ERR_BAD_ARG;
  const int wi = f0->width_image, wm = f0->width_motion;
  const bool det = f0->deterministic != 0;
  const bool fuse_adam = det && f0->fuse_adam != 0;
  const int64_t P = (int64_t)H * W, MP = P * M;
  for (int b = 0; b < B; ++b) {
    const immoco_fit* f = fs[b];
    if (!f || f->h != H || f->w != W || f->m != M || M != f->lines.n_groups) return IMMOCO_ERR_BAD_ARG;
    if (f->width_image != wi || f->width_motion != wm || f->act_image != f0->act_image || f->act_motion != f0->act_motion)
      return IMMOCO_ERR_BAD_ARG;
    if ((f->deterministic != 0) != det || (det && (f->fuse_adam != 0) != fuse_adam)) return IMMOCO_ERR_BAD_ARG;
    if (f->n_motion != f0->n_motion || f->n_image != f0->n_image || (f->n_motion & 3) != 0) return IMMOCO_ERR_BAD_ARG;
    if ((f->loss_slots == nullptr) != (f0->loss_slots == nullptr)) return IMMOCO_ERR_BAD_ARG;
    // tap-indexed image table: all fits of a batch share the shape, hence the tap list and the live-row count
    if ((f->taps_image.rows == nullptr) != (f0->taps_image.rows == nullptr)) return IMMOCO_ERR_BAD_ARG;
    if (grid_has_lut(f->grid_motion) != grid_has_lut(f0->grid_motion)) return IMMOCO_ERR_BAD_ARG;
    if (!det && f->taps_image.rows &&
        (f->taps_image.n_points != (int64_t)H * W || f->taps_image.n_active_rows != f0->taps_image.n_active_rows ||
         f->taps_image.n_active_rows < 0 || f->taps_image.n_active_rows > (int64_t)f->grid_image.offset[f->grid_image.n_levels]))
      return IMMOCO_ERR_BAD_ARG;
    if (det) {
      if (!f->loss_slots || !f->d_image_fx || !f->dc_max_bits || !f->mlp_part_image || !f->csr_image.row_ptr ||
          !f->csr_image.taps)
        return IMMOCO_ERR_BAD_ARG;
      if (M > 0 && (!f->mlp_part_motion || !f->csr_motion.row_ptr || !f->csr_motion.taps)) return IMMOCO_ERR_BAD_ARG;
      if (f->csr_image.n_points != P || (M > 0 && f->csr_motion.n_points != MP)) return IMMOCO_ERR_BAD_ARG;
    }
  }
  if (it_begin == it_end) return 0;
  // parameter views: [motion | image], each [W1 (width x 32) | W2 (16 x width) | table]
  const int64_t mlp_m = (int64_t)wm * 32 + 16 * (int64_t)wm;
  const int64_t mlp_i = (int64_t)wi * 32 + 16 * (int64_t)wi;
  const int64_t n_mot = f0->n_motion;
  // floats of the image INR that Adam and the gradient memset visit: everything, or (tap-indexed table) the MLP
  // block + the rows some pixel touches, rounded up to a 128-bit item
  const bool taps_i = !det && f0->taps_image.rows != nullptr;
  // motion table under a chunk-table layout (immoco_grid_desc::layout_lut): the grouped kernels, which rely on
  // coords_motion being (t_g, y_p, x_p) -- what make_grids((M, H, W)) produces (src/models/immoco.py:48-53)
  const bool grouped_m = M > 0 && grid_has_lut(f0->grid_motion);
  if (grouped_m && (det || g_fused_scatter)) return IMMOCO_ERR_UNSUPPORTED;
  int64_t n_img_live = f0->n_image;
  if (taps_i) {
    n_img_live = (mlp_i + 2 * f0->taps_image.n_active_rows + 3) / 4 * 4;
    if (n_img_live > f0->n_image) n_img_live = f0->n_image;
  }
  int slots[2] = {0, 0};
  if (f0->loss_slots) IMMOCO_TRY(immoco_fit_loss_slots(H, W, slots));
  const int per_iter = slots[0] + slots[1];
  const int n_part_i = det ? immoco_mlp_bwd_tc_grid(P) : 0;
  const int n_part_m = (det && M > 0) ? immoco_mlp_bwd_tc_grid(MP) : 0;

  cudaStream_t ms = (cudaStream_t)stream;
  // The reproducible path runs on ONE stream: its two big kernels (row-sorted gather + Adam of the motion and of the
  // image grid) are both bound by L2 random gathers and a 420 MB tap stream, and co-running them -- or either of them
  // beside the other branch -- costs more than it hides (C2: 1037 us per iteration on two streams, 732 on one;
  // profiles/round2_deterministic_mode_cost.txt).
  AuxStream* aux = (g_overlap && M > 0 && !det) ? aux_for(ms) : nullptr;
  bool forked = false;          // aux currently carries work that `ms` has not joined
  bool zero_pending = false;    // the previous iteration left its gradients for the deferred memset

  const bool fuse_rows = g_fuse_rows != 0 || det || B > 1;
  for (int b = 0; b < B; ++b) {
    const immoco_fit* f = fs[b];
    // the fused row launch ADDS into c_tmp; the column pass re-zeroes it for the next iteration
    if (fuse_rows && cudaMemsetAsync(f->c_tmp, 0, (size_t)P * 2 * sizeof(float), ms) != cudaSuccess) return IMMOCO_ERR_BAD_ARG;
    // the per-iteration maxima are integer atomicMax targets: clear the ones this call owns
    if (det && cudaMemsetAsync(f->dc_max_bits + it_begin, 0, (size_t)(it_end - it_begin) * sizeof(uint32_t), ms) != cudaSuccess)
      return IMMOCO_ERR_BAD_ARG;
  }
  auto zero_grads = [&](cudaStream_t st) {
    for (int b = 0; b < B; ++b) cudaMemsetAsync(fs[b]->grads, 0, (size_t)(n_mot + n_img_live) * sizeof(float), st);
  };

  for (int it = it_begin; it < it_end; ++it) {
    // ---- this iteration's view of the batch -----------------------------------------------------------
    FitBatch fb = {};
    MlpFwdBatch mi = {}, mm = {};
    fb.n = mi.n = mm.n = B;
    for (int b = 0; b < B; ++b) {
      const immoco_fit* f = fs[b];
      FitBatchInst& a = fb.inst[b];
      a.image = (const float2*)f->image;
      a.disp = (const float2*)f->disp;
      a.c_tmp = (float2*)f->c_tmp;
      a.k_in = (const float2*)f->k_in;
      a.k_out = (float2*)f->k_out;
      a.d_c = (float2*)f->d_c;
      a.d_image = (float2*)f->d_image;
      a.d_disp = (float2*)f->d_disp;
      a.fx = det ? (long long*)f->d_image_fx : nullptr;
      a.dmax = det ? f->dc_max_bits + it : nullptr;
      a.loss_dc = f->loss + 2 * (int64_t)it;
      a.loss_ge = a.loss_dc + 1;
      a.slots_dc = f->loss_slots ? f->loss_slots + (size_t)it * per_iter : nullptr;
      a.slots_ge = a.slots_dc ? a.slots_dc + slots[0] : nullptr;
      a.lines = f->lines;
      float* pm = f->params;
      float* pi = f->params + n_mot;
      mi.enc[b] = (const float2*)f->enc_image; mi.w1[b] = pi; mi.w2[b] = pi + (int64_t)wi * 32; mi.out[b] = (float2*)f->image;
      mm.enc[b] = (const float2*)f->enc_motion; mm.w1[b] = pm; mm.w2[b] = pm + (int64_t)wm * 32; mm.out[b] = (float2*)f->disp;
    }
    cudaEvent_t* ev = nullptr;
    if (prof && profile_every > 0 && (it % profile_every) == profile_every - 1 && prof->used < prof->capacity)
      ev = prof->ev + (size_t)(prof->used++) * prof_stride(prof);
    // instrumented iterations run serially on the caller's stream (contention-free durations) unless
    // timeline mode is on
    const bool two = aux && (!ev || g_profile_overlap);
    if (!two && forked) {       // fold the image branch back before a serial iteration
      cudaEventRecord(aux->join_end, aux->stream);
      cudaStreamWaitEvent(ms, aux->join_end, 0);
      forked = false;
    }
    if (two && !forked) {       // the image branch starts behind everything already on `ms`
      cudaEventRecord(aux->fork, ms);
      cudaStreamWaitEvent(aux->stream, aux->fork, 0);
      forked = true;
    }
    void* is = two ? (void*)aux->stream : stream;       // stream of the image-INR branch
    // deferred zeroing needs the auxiliary streams; serial (instrumented) iterations zero inside Adam.  The
    // deterministic path never zeroes: every gradient it consumes is WRITTEN in the same iteration.
    const bool defer_zero = two && g_deferred_zero != 0 && !det;
    if (zero_pending && !defer_zero) {      // the previous iteration deferred its zeroing, this one cannot
      if (aux) cudaStreamWaitEvent(ms, aux->adam_i_done, 0);
      zero_grads(ms);
      zero_pending = false;
    }
    if (ev) cudaEventRecord(ev[0], ms);
// per-instance launches of one profile slot
#define EACH(expr)                                    \
  [&]() -> int {                                      \
    for (int b = 0; b < B; ++b) {                     \
      const immoco_fit* f = fs[b];                    \
      [[maybe_unused]] float* pm = f->params;                  \
      [[maybe_unused]] float* pi = f->params + n_mot;          \
      [[maybe_unused]] float* gm = f->grads;                   \
      [[maybe_unused]] float* gi = f->grads + n_mot;           \
      [[maybe_unused]] float* mom1_m = f->exp_avg;             \
      [[maybe_unused]] float* mom2_m = f->exp_avg_sq;          \
      [[maybe_unused]] float* mom1_i = f->exp_avg + n_mot;     \
      [[maybe_unused]] float* mom2_i = f->exp_avg_sq + n_mot;  \
      int e__ = (expr);                               \
      if (e__ != 0) return e__;                       \
    }                                                 \
    return 0;                                         \
  }()
    // ---- forward -------------------------------------------------------------------------------
    K(0, is, EACH(taps_i ? immoco_hashgrid_fwd_taps(&f->grid_image, &f->taps_image, f->coords_image, pi + mlp_i, f->enc_image, P, is)
                         : immoco_hashgrid_fwd(&f->grid_image, f->coords_image, pi + mlp_i, f->enc_image, P, is)));
    K(1, is, immoco_mlp_fwd_tc_batch(mi, P, wi, f0->act_image, 0, is));
    // gradient entropy needs the image only; it initialises d_image (lambda folded in)
    K(7, is, immoco_grad_entropy_batch(fb, lambdas_host[it], 0, H, W, is));
    if (two) cudaEventRecord(aux->join_fwd, aux->stream);
    K(2, ms, M > 0 ? EACH(grouped_m ? immoco_hashgrid_fwd_grouped(&f->grid_motion, f->coords_motion, pm + mlp_m, f->enc_motion, P, M, stream)
                                    : immoco_hashgrid_fwd(&f->grid_motion, f->coords_motion, pm + mlp_m, f->enc_motion, MP, stream)) : nop());
    if (defer_zero && zero_pending) {
      // gradients of the previous iteration were consumed by both Adam launches (Adam_m precedes this point
      // on `ms`, Adam_i is awaited through its event): zero them now, beside the SM-bound kernels that follow
      cudaEventRecord(aux->zero_start, ms);
      cudaStreamWaitEvent(aux->zero_stream, aux->zero_start, 0);
      cudaStreamWaitEvent(aux->zero_stream, aux->adam_i_done, 0);
      zero_grads(aux->zero_stream);
      cudaEventRecord(aux->zero_done, aux->zero_stream);
    }
    K(3, ms, M > 0 ? immoco_mlp_fwd_tc_batch(mm, MP, wm, f0->act_motion, 1, stream) : nop());
    if (two) cudaStreamWaitEvent(ms, aux->join_fwd, 0);
    if (fuse_rows) {      // slot 4 (static row pass) is folded into slot 5, slot 8 into slot 9
      K(4, ms, nop());
      K(5, ms, immoco_rows_fwd_fused_batch(fb, f0->coords_image, f0->tw_w, H, W, stream));
      K(6, ms, immoco_colpass_loss_batch(fb, f0->tw_h, H, W, 1, stream));
      K(9, ms, immoco_rows_bwd_fused_batch(fb, f0->coords_image, f0->tw_w, H, W, det, stream));
      if (det) {
        // d_image += the fixed-point plane; on the image branch (it is the only consumer), beside MLPm_b
        if (two) { cudaEventRecord(aux->rows_done, ms); cudaStreamWaitEvent(aux->stream, aux->rows_done, 0); }
        K(8, is, immoco_d_image_finalize_batch(fb, H, W, is));
      } else {
        K(8, ms, nop());
      }
    } else {      // four separate launches (A/B check; single instance, float-atomic mode only)
      const immoco_fit* f = f0;
      const FitBatchInst& a = fb.inst[0];
      K(4, ms, immoco_rows_static(f->image, f->c_tmp, H, W, f->tw_w, nullptr, f->lines.static_w, 0, false, stream));
      K(5, ms, immoco_motion_rows_fwd(f->image, f->disp, f->coords_image, &f->lines, f->tw_w, f->c_tmp, H, W, stream));
      K(6, ms, immoco_colpass_loss_slots(f->c_tmp, f->k_in, f->k_out, f->d_c, a.loss_dc, f->tw_h, H, W, a.slots_dc, nullptr, stream));
      // ---- backward ----------------------------------------------------------------------------
      K(8, ms, immoco_rows_static(f->d_c, f->d_image, H, W, f->tw_w, f->lines.static_w, nullptr, 1, true, stream));
      K(9, ms, M > 0 ? immoco_motion_rows_bwd(f->d_c, f->image, f->disp, f->coords_image, &f->lines, f->tw_w,
                                              f->d_image, f->d_disp, H, W, stream) : nop());
    }
    if (defer_zero && zero_pending) {      // first accumulation into the gradients of this iteration
      cudaStreamWaitEvent(ms, aux->zero_done, 0);
      zero_pending = false;
    }
    if (det) {
      // ---- backward + update, reproducible: per-CTA weight-gradient blocks, row-sorted gathers -----------
      K(10, ms, M > 0 ? EACH(immoco_mlp_bwd_partials(f->enc_motion, pm, pm + (int64_t)wm * 32, f->d_disp, f->d_enc_motion,
                                                     f->mlp_part_motion, MP, wm, f->act_motion, stream)) : nop());
      if (two) { cudaEventRecord(aux->mlp_done, ms); cudaStreamWaitEvent(aux->stream, aux->mlp_done, 0); }
      if (fuse_adam) {
        K(11, ms, M > 0 ? EACH(immoco_hashgrid_bwd_csr_adam(&f->grid_motion, &f->csr_motion, f->d_enc_motion, pm + mlp_m,
                                                            mom1_m + mlp_m, mom2_m + mlp_m, nullptr, f->lr, f->beta1,
                                                            f->beta2, f->eps, it + 1, stream)) : nop());
      } else {
        K(11, ms, M > 0 ? EACH(immoco_hashgrid_bwd_csr(&f->grid_motion, &f->csr_motion, f->d_enc_motion, gm + mlp_m, stream)) : nop());
      }
      K(12, is, EACH(immoco_mlp_bwd_partials(f->enc_image, pi, pi + (int64_t)wi * 32, f->d_image, f->d_enc_image,
                                             f->mlp_part_image, P, wi, f->act_image, is)));
      if (fuse_adam) {
        K(13, is, EACH(immoco_hashgrid_bwd_csr_adam(&f->grid_image, &f->csr_image, f->d_enc_image, pi + mlp_i, mom1_i + mlp_i,
                                                    mom2_i + mlp_i, nullptr, f->lr, f->beta1, f->beta2, f->eps, it + 1, is)));
      } else {
        K(13, is, EACH(immoco_hashgrid_bwd_csr(&f->grid_image, &f->csr_image, f->d_enc_image, gi + mlp_i, is)));
      }
      if (ev) cudaEventRecord(ev[1 + 2 * 14], ms);
      if (M > 0) {
        IMMOCO_TRY(EACH(immoco_adam_step_partials(pm, f->mlp_part_motion, n_part_m, mom1_m, mom2_m, mlp_m, f->lr, f->beta1,
                                                  f->beta2, f->eps, it + 1, stream)));
        if (!fuse_adam)
          IMMOCO_TRY(EACH(immoco_adam_step(pm + mlp_m, gm + mlp_m, mom1_m + mlp_m, mom2_m + mlp_m, n_mot - mlp_m, f->lr,
                                           f->beta1, f->beta2, f->eps, it + 1, 0, stream)));
      }
      if (ev) { cudaEventRecord(ev[2 + 2 * 14], ms); cudaEventRecord(ev[1 + 2 * 15], (cudaStream_t)is); }
      IMMOCO_TRY(EACH(immoco_adam_step_partials(pi, f->mlp_part_image, n_part_i, mom1_i, mom2_i, mlp_i, f->lr, f->beta1,
                                                f->beta2, f->eps, it + 1, is)));
      if (!fuse_adam)
        IMMOCO_TRY(EACH(immoco_adam_step(pi + mlp_i, gi + mlp_i, mom1_i + mlp_i, mom2_i + mlp_i, f->n_image - mlp_i, f->lr,
                                         f->beta1, f->beta2, f->eps, it + 1, 0, is)));
      if (ev) cudaEventRecord(ev[2 + 2 * 15], (cudaStream_t)is);
      continue;
    }
    if (g_fused_scatter && M > 0 && wm == 64) {
      // the hashed levels' table gradients leave the MLP backward kernel directly (no feature-plane round trip);
      // the dense levels follow through d_enc
      K(10, ms, EACH(immoco_mlp_bwd_scatter(f->enc_motion, pm, pm + (int64_t)wm * 32, f->d_disp, f->d_enc_motion, gm,
                                            gm + (int64_t)wm * 32, &f->grid_motion, f->coords_motion, gm + mlp_m, MP, wm,
                                            f->act_motion, stream)));
      if (two) { cudaEventRecord(aux->mlp_done, ms); cudaStreamWaitEvent(aux->stream, aux->mlp_done, 0); }
      K(11, ms, EACH(immoco_hashgrid_bwd_dense_levels(&f->grid_motion, f->coords_motion, f->d_enc_motion, gm + mlp_m, MP,
                                                      stream)));
    } else {
      K(10, ms, M > 0 ? EACH(immoco_mlp_bwd(f->enc_motion, pm, pm + (int64_t)wm * 32, f->d_disp, f->d_enc_motion, gm,
                                            gm + (int64_t)wm * 32, MP, wm, f->act_motion, stream)) : nop());
      if (two) { cudaEventRecord(aux->mlp_done, ms); cudaStreamWaitEvent(aux->stream, aux->mlp_done, 0); }
      K(11, ms, M > 0 ? EACH(grouped_m ? immoco_hashgrid_bwd_grouped(&f->grid_motion, f->coords_motion, f->d_enc_motion, gm + mlp_m, P, M, stream)
                                       : immoco_hashgrid_bwd(&f->grid_motion, f->coords_motion, f->d_enc_motion, gm + mlp_m, MP, stream)) : nop());
    }
    K(12, is, EACH(immoco_mlp_bwd(f->enc_image, pi, pi + (int64_t)wi * 32, f->d_image, f->d_enc_image, gi,
                                  gi + (int64_t)wi * 32, P, wi, f->act_image, is)));
    K(13, is, EACH(taps_i ? immoco_hashgrid_bwd_taps(&f->grid_image, &f->taps_image, f->coords_image, f->d_enc_image, gi + mlp_i, P, is)
                          : immoco_hashgrid_bwd(&f->grid_image, f->coords_image, f->d_enc_image, gi + mlp_i, P, is)));
    // ---- update (zero_grad fused), one launch per INR so each follows its own branch ---------------
    // (the LAST iteration of the call zeroes inside Adam, so the gradients are clean when the call returns)
    const int adam_zeroes = (defer_zero && it + 1 < it_end) ? 0 : 1;
    K(14, ms, M > 0 ? EACH(immoco_adam_step(pm, gm, f->exp_avg, f->exp_avg_sq, n_mot, f->lr, f->beta1, f->beta2,
                                            f->eps, it + 1, adam_zeroes, stream)) : nop());
    K(15, is, EACH(immoco_adam_step(pi, gi, f->exp_avg + n_mot, f->exp_avg_sq + n_mot, n_img_live, f->lr,
                                    f->beta1, f->beta2, f->eps, it + 1, adam_zeroes, is)));
    if (!adam_zeroes) {
      cudaEventRecord(aux->adam_i_done, aux->stream);
      zero_pending = true;
    }
  }
#undef EACH
  if (forked) {
    cudaEventRecord(aux->join_end, aux->stream);
    cudaStreamWaitEvent(ms, aux->join_end, 0);
  }
  if (f0->loss_slots) {      // per-CTA loss slots of this call's iterations -> loss, added in slot order
    const int n = 2 * (it_end - it_begin);
    for (int b = 0; b < B; ++b) {
      loss_reduce_kernel<<<(n + 127) / 128, 128, 0, ms>>>(fs[b]->loss_slots, per_iter, slots[0], fs[b]->loss, it_begin, it_end);
      IMMOCO_LAUNCH_CHECK();
    }
  }
  return 0;
}

extern "C" int immoco_fit_run(const immoco_fit* f, int32_t it_begin, int32_t it_end,
                              const float* lambdas_host, void* stream, immoco_profile* prof,
                              int32_t profile_every) {
  return fit_run_impl(&f, 1, it_begin, it_end, lambdas_host, stream, prof, profile_every);
}

extern "C" int immoco_fit_run_batched(const immoco_fit* const* fits, int32_t n_fits, int32_t it_begin, int32_t it_end,
                                      const float* lambdas_host, void* stream, immoco_profile* prof,
                                      int32_t profile_every) {
  return fit_run_impl(fits, n_fits, it_begin, it_end, lambdas_host, stream, prof, profile_every);
}

extern "C" int immoco_max_fit_batch(void) { return kMaxFitBatch; }
