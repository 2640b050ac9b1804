// Fused Adam and the per-iteration driver of the IM-MoCo fit
// (src/models/immoco.py:149-154 optimizer, :164-181 loop).
#include <math.h>

#include "common.cuh"

// launch helpers implemented in forward_model.cu
int immoco_rows_static(const float* in, float* out, int h, int w, const float* tw_w,
                       const float* in_w, const float* out_w, int accumulate, bool inv, void* stream);
int immoco_motion_rows_fwd(const float* image, const float* disp, const float* ident,
                           const immoco_lines* lines, const float* tw_w, float* c_out, int h, int w,
                           void* stream);
int immoco_motion_rows_bwd(const float* d_c, const float* image, const float* disp, const float* ident,
                           const immoco_lines* lines, const float* tw_w, float* d_image, float* d_disp,
                           int h, int w, void* stream);

namespace {

// torch.optim.Adam (amsgrad=False, weight_decay=0, maximize=False), single-tensor formulation:
//   m = lerp(m, g, 1-b1);  v = b2*v + (1-b2)*g*g;  p -= step_size * m / (sqrt(v)/bc2_sqrt + eps)
// One pass: reads p,g,m,v, writes p,m,v and zeroes g (28 B / parameter), 128-bit accesses.
__global__ void __launch_bounds__(256)
adam_kernel(float4* __restrict__ p, float4* __restrict__ g, float4* __restrict__ m,
            float4* __restrict__ v, int64_t n4, float one_minus_b1, float b2, float one_minus_b2,
            float step_size, float bc2_sqrt, float eps, int zero_grad) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float4 gi = g[i];
    float4 mi = m[i], vi = v[i], pi = p[i];
#define IMMOCO_ADAM_LANE(c)                                              \
  mi.c = mi.c + one_minus_b1 * (gi.c - mi.c);                            \
  vi.c = fmaf(one_minus_b2 * gi.c, gi.c, b2 * vi.c);                     \
  pi.c = pi.c - step_size * (mi.c / (sqrtf(vi.c) / bc2_sqrt + eps));
    IMMOCO_ADAM_LANE(x) IMMOCO_ADAM_LANE(y) IMMOCO_ADAM_LANE(z) IMMOCO_ADAM_LANE(w)
#undef IMMOCO_ADAM_LANE
    m[i] = mi; v[i] = vi; p[i] = pi;
    if (zero_grad) g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

__global__ void adam_tail_kernel(float* p, float* g, float* m, float* v, int64_t begin, int64_t n,
                                 float one_minus_b1, float b2, float one_minus_b2, float step_size,
                                 float bc2_sqrt, float eps, int zero_grad) {
  const int64_t i = begin + threadIdx.x;
  if (i < n) {
    const float gi = g[i];
    const float mi = m[i] + one_minus_b1 * (gi - m[i]);
    const float vi = fmaf(one_minus_b2 * gi, gi, b2 * v[i]);
    p[i] = p[i] - step_size * (mi / (sqrtf(vi) / bc2_sqrt + eps));
    m[i] = mi; v[i] = vi;
    if (zero_grad) g[i] = 0.f;
  }
}

}  // namespace

extern "C" int immoco_adam_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq,
                                int64_t n, double lr, double beta1, double beta2, double eps,
                                int32_t step, int32_t zero_grad, void* stream) {
  if (n < 0 || step < 1) return IMMOCO_ERR_BAD_ARG;
  if (n == 0) return 0;
  if ((((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) != 0)
    return IMMOCO_ERR_BAD_ARG;
  // bias corrections in double like torch (python floats), then rounded to fp32 scalars
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  const float step_size = (float)(lr / bc1);
  const float bc2_sqrt = (float)sqrt(bc2);
  const float omb1 = (float)(1.0 - beta1);
  const float omb2 = (float)(1.0 - beta2);
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t n4 = n / 4;
  if (n4 > 0) {
    int64_t blocks = (n4 + 255) / 256;
    const int64_t cap = (int64_t)IMMOCO_NUM_SMS * 16;   // 16 resident 256-thread CTAs per SM
    if (blocks > cap) blocks = cap;
    adam_kernel<<<(unsigned)blocks, 256, 0, s>>>((float4*)params, (float4*)grads, (float4*)exp_avg,
                                                 (float4*)exp_avg_sq, n4, omb1, (float)beta2, omb2, step_size,
                                                 bc2_sqrt, (float)eps, zero_grad);
    IMMOCO_LAUNCH_CHECK();
  }
  if (n4 * 4 < n) {
    adam_tail_kernel<<<1, 4, 0, s>>>(params, grads, exp_avg, exp_avg_sq, n4 * 4, n, omb1, (float)beta2, omb2,
                                     step_size, bc2_sqrt, (float)eps, zero_grad);
    IMMOCO_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int immoco_abi_version(void) { return 1; }

extern "C" void immoco_struct_sizes(int32_t out[3]) {
  out[0] = (int32_t)sizeof(immoco_grid_desc);
  out[1] = (int32_t)sizeof(immoco_lines);
  out[2] = (int32_t)sizeof(immoco_fit);
}

// hashgrid fwd + mlp fwd (x2), rows, motion rows, colpass, GE, rows adj, motion rows bwd,
// mlp bwd + hashgrid bwd (x2), adam
extern "C" int immoco_launches_per_iteration(int32_t m) { return m > 0 ? 15 : 9; }

#define IMMOCO_TRY(expr)          \
  do {                            \
    int e__ = (expr);             \
    if (e__ != 0) return e__;     \
  } while (0)

// ---- optional per-kernel timing: CUDA events recorded on the launching stream around every
//      kernel of selected iterations; no synchronisation until immoco_profile_read() -----------
struct immoco_profile {
  int n_slots;                 // kernels per iteration
  int capacity;                // instrumented iterations the event pool holds
  int used;
  cudaEvent_t* ev;             // capacity * (n_slots + 1) events
};

extern "C" immoco_profile* immoco_profile_create(int32_t capacity) {
  if (capacity < 1) return nullptr;
  immoco_profile* p = new immoco_profile;
  p->n_slots = IMMOCO_PROFILE_SLOTS;
  p->capacity = capacity;
  p->used = 0;
  const int n = capacity * (p->n_slots + 1);
  p->ev = new cudaEvent_t[n];
  for (int i = 0; i < n; ++i) {
    if (cudaEventCreate(&p->ev[i]) != cudaSuccess) { p->capacity = 0; break; }
  }
  return p;
}

extern "C" void immoco_profile_destroy(immoco_profile* p) {
  if (!p) return;
  for (int i = 0; i < p->capacity * (p->n_slots + 1); ++i) cudaEventDestroy(p->ev[i]);
  delete[] p->ev;
  delete p;
}

// Sums per-slot milliseconds over the instrumented iterations into ms_sum[IMMOCO_PROFILE_SLOTS];
// returns the number of iterations summed (the caller must have synchronised the stream).
extern "C" int immoco_profile_read(immoco_profile* p, float* ms_sum) {
  if (!p || !ms_sum) return IMMOCO_ERR_BAD_ARG;
  for (int k = 0; k < p->n_slots; ++k) ms_sum[k] = 0.f;
  for (int i = 0; i < p->used; ++i) {
    cudaEvent_t* e = p->ev + (size_t)i * (p->n_slots + 1);
    for (int k = 0; k < p->n_slots; ++k) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, e[k], e[k + 1]) != cudaSuccess) return IMMOCO_ERR_BAD_ARG;
      ms_sum[k] += ms;
    }
  }
  const int n = p->used;
  p->used = 0;
  return n;
}

// ---- image-INR branch on a second stream --------------------------------------------------------
// The two INR branches of an iteration are independent and stress different units
// (hash-grid gathers / reductions: L2; MLPs: tensor pipe + SIMT epilogue), so the image branch runs on
// an auxiliary non-blocking stream, forked and joined with events.  One aux stream + 4 events per
// device are created on first use and live for the process (the library's only hidden state).
struct AuxStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork_fwd = nullptr, join_fwd = nullptr, fork_bwd = nullptr, join_bwd = nullptr;
  bool ok = false;
};
static AuxStream* aux_for_current_device() {
  static AuxStream table[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  AuxStream& a = table[dev];
  if (!a.ok) {
    if (cudaStreamCreateWithFlags(&a.stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    cudaEvent_t* ev[4] = {&a.fork_fwd, &a.join_fwd, &a.fork_bwd, &a.join_bwd};
    for (auto e : ev)
      if (cudaEventCreateWithFlags(e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    a.ok = true;
  }
  return &a;
}
static int g_overlap = 1;
extern "C" int immoco_set_branch_overlap(int32_t on) { g_overlap = on ? 1 : 0; return 0; }

#define IMMOCO_MARK()                                                  \
  do {                                                                 \
    if (ev) { cudaEventRecord(ev[slot], (cudaStream_t)stream); }       \
    ++slot;                                                            \
  } while (0)

extern "C" int immoco_fit_run(const immoco_fit* f, int32_t it_begin, int32_t it_end,
                              const float* lambdas_host, void* stream, immoco_profile* prof,
                              int32_t profile_every) {
  if (!f || !lambdas_host || it_begin < 0 || it_end < it_begin) return IMMOCO_ERR_BAD_ARG;
  const int H = f->h, W = f->w, M = f->m;
  if (H < 2 || W < 2 || M < 0 || M != f->lines.n_groups) return IMMOCO_ERR_BAD_ARG;
  const int64_t P = (int64_t)H * W, MP = P * M;
  const int wi = f->width_image, wm = f->width_motion;
  // parameter views: [motion | image], each [W1 (width x 32) | W2 (16 x width) | table]
  float* pm = f->params;
  float* pi = f->params + f->n_motion;
  float* gm = f->grads;
  float* gi = f->grads + f->n_motion;
  const int64_t mlp_m = (int64_t)wm * 32 + 16 * (int64_t)wm;
  const int64_t mlp_i = (int64_t)wi * 32 + 16 * (int64_t)wi;

  for (int it = it_begin; it < it_end; ++it) {
    double* loss = f->loss + 2 * (int64_t)it;
    cudaEvent_t* ev = nullptr;
    if (prof && profile_every > 0 && (it % profile_every) == profile_every - 1 && prof->used < prof->capacity)
      ev = prof->ev + (size_t)(prof->used++) * (prof->n_slots + 1);
    int slot = 0;
    // instrumented iterations run serially on the caller's stream so per-kernel event times mean something
    AuxStream* aux = (g_overlap && !ev && M > 0) ? aux_for_current_device() : nullptr;
    cudaStream_t ms = (cudaStream_t)stream;
    void* is = aux ? (void*)aux->stream : stream;       // stream of the image-INR branch
    IMMOCO_MARK();
    // ---- forward -------------------------------------------------------------------------------
    if (aux) { cudaEventRecord(aux->fork_fwd, ms); cudaStreamWaitEvent(aux->stream, aux->fork_fwd, 0); }
    IMMOCO_TRY(immoco_hashgrid_fwd(&f->grid_image, f->coords_image, pi + mlp_i, f->enc_image, P, is));
    IMMOCO_MARK();  // slot 0: hashgrid_fwd_image
    IMMOCO_TRY(immoco_mlp_fwd(f->enc_image, pi, pi + (int64_t)wi * 32, f->image, P, wi, f->act_image, 0, is));
    IMMOCO_MARK();  // 1: mlp_fwd_image
    if (aux) cudaEventRecord(aux->join_fwd, aux->stream);
    if (M > 0) {
      IMMOCO_TRY(immoco_hashgrid_fwd(&f->grid_motion, f->coords_motion, pm + mlp_m, f->enc_motion, MP, stream));
    }
    IMMOCO_MARK();  // 2: hashgrid_fwd_motion
    if (M > 0) {
      IMMOCO_TRY(immoco_mlp_fwd(f->enc_motion, pm, pm + (int64_t)wm * 32, f->disp, MP, wm, f->act_motion, 1, stream));
    }
    IMMOCO_MARK();  // 3: mlp_fwd_motion
    if (aux) cudaStreamWaitEvent(ms, aux->join_fwd, 0);
    IMMOCO_TRY(immoco_rows_static(f->image, f->c_tmp, H, W, f->tw_w, nullptr, f->lines.static_w, 0, false, stream));
    IMMOCO_MARK();  // 4: fft_rows
    IMMOCO_TRY(immoco_motion_rows_fwd(f->image, f->disp, f->coords_image, &f->lines, f->tw_w, f->c_tmp, H, W, stream));
    IMMOCO_MARK();  // 5: motion_rows_fwd
    IMMOCO_TRY(immoco_colpass_loss(f->c_tmp, f->k_in, f->k_out, f->d_c, loss, f->tw_h, H, W, stream));
    IMMOCO_MARK();  // 6: colpass_loss
    // ---- backward ------------------------------------------------------------------------------
    IMMOCO_TRY(immoco_grad_entropy(f->image, lambdas_host[it], loss + 1, f->d_image, 0, H, W, stream));
    IMMOCO_MARK();  // 7: grad_entropy
    IMMOCO_TRY(immoco_rows_static(f->d_c, f->d_image, H, W, f->tw_w, f->lines.static_w, nullptr, 1, true, stream));
    IMMOCO_MARK();  // 8: fft_rows_adj
    if (M > 0) {
      IMMOCO_TRY(immoco_motion_rows_bwd(f->d_c, f->image, f->disp, f->coords_image, &f->lines, f->tw_w,
                                        f->d_image, f->d_disp, H, W, stream));
    }
    IMMOCO_MARK();  // 9: motion_rows_bwd
    if (aux) { cudaEventRecord(aux->fork_bwd, ms); cudaStreamWaitEvent(aux->stream, aux->fork_bwd, 0); }
    if (M > 0) {
      IMMOCO_TRY(immoco_mlp_bwd(f->enc_motion, pm, pm + (int64_t)wm * 32, f->d_disp, f->d_enc_motion, gm,
                                gm + (int64_t)wm * 32, MP, wm, f->act_motion, stream));
    }
    IMMOCO_MARK();  // 10: mlp_bwd_motion
    if (M > 0) {
      IMMOCO_TRY(immoco_hashgrid_bwd(&f->grid_motion, f->coords_motion, f->d_enc_motion, gm + mlp_m, MP, stream));
    }
    IMMOCO_MARK();  // 11: hashgrid_bwd_motion
    IMMOCO_TRY(immoco_mlp_bwd(f->enc_image, pi, pi + (int64_t)wi * 32, f->d_image, f->d_enc_image, gi,
                              gi + (int64_t)wi * 32, P, wi, f->act_image, is));
    IMMOCO_MARK();  // 12: mlp_bwd_image
    IMMOCO_TRY(immoco_hashgrid_bwd(&f->grid_image, f->coords_image, f->d_enc_image, gi + mlp_i, P, is));
    IMMOCO_MARK();  // 13: hashgrid_bwd_image
    if (aux) { cudaEventRecord(aux->join_bwd, aux->stream); cudaStreamWaitEvent(ms, aux->join_bwd, 0); }
    // ---- update (zero_grad fused) ----------------------------------------------------------------
    IMMOCO_TRY(immoco_adam_step(f->params, f->grads, f->exp_avg, f->exp_avg_sq, f->n_motion + f->n_image,
                                f->lr, f->beta1, f->beta2, f->eps, it + 1, 1, stream));
    IMMOCO_MARK();  // 14: adam
  }
  return 0;
}
