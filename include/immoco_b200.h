/*
 * immoco_b200.h -- C ABI of the B200-native IM-MoCo hot path (libimmoco_b200.so).
 *
 * The reference (multimodallearning/MICCAI24_IMMoCo) is pure Python: it has no FFI layer of its
 * own.  The seam it uses for this path is the tiny-cuda-nn torch binding plus ATen ops; every
 * entry point below replaces one of those call sites (cited as file:line relative to the
 * reference root).  INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer owned by the caller
 *    (the torch caching allocator in the shipped host code); the library allocates no device memory
 *    (its only hidden state: one auxiliary stream set per caller stream, immoco_release_streams());
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*), never synchronises;
 *  - return value: 0 on success, otherwise the cudaError_t of the failed launch, or
 *    IMMOCO_ERR_* (negative) for rejected arguments;
 *  - complex tensors are interleaved (re, im) fp32 pairs, row-major (H, W);
 *  - "enc" feature planes are level-major: enc[level][point] = float2 (the 2 features of a level).
 */
#ifndef IMMOCO_B200_H_
#define IMMOCO_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IMMOCO_MAX_LEVELS 16
#define IMMOCO_ERR_BAD_ARG (-1)
#define IMMOCO_ERR_UNSUPPORTED (-2)

#define IMMOCO_ACT_NONE 0
#define IMMOCO_ACT_RELU 1
#define IMMOCO_ACT_TANH 2

/* Multiresolution hash-grid description (tiny-cuda-nn "Grid"/"Hash" encoding configured at
 * src/models/immoco.py:27-37; semantics restated in oracle/immoco_oracle.py:make_grid_levels). */
typedef struct immoco_grid_desc {
  int32_t n_dims;                           /* 2 (image INR) or 3 (motion INR)          */
  int32_t n_levels;                         /* <= IMMOCO_MAX_LEVELS                      */
  float scale[IMMOCO_MAX_LEVELS];           /* base * per_level_scale^l - 1              */
  uint32_t resolution[IMMOCO_MAX_LEVELS];   /* ceil(scale) + 1                           */
  uint32_t entries[IMMOCO_MAX_LEVELS];      /* rows of the level's table                 */
  uint32_t offset[IMMOCO_MAX_LEVELS + 1];   /* first row of each level (prefix sums)     */
  uint32_t hashed[IMMOCO_MAX_LEVELS];       /* 1: coherent-prime hash, 0: dense index    */
  /* Physical row layout of a hashed level (0 = the reference's own layout: row = index).  Otherwise the row
   * of hash index r is S(r): Gray code r ^ (r >> 1), then bit positions (swizzle & 0xff) and
   * ((swizzle >> 8) & 0xff) exchanged.  S is a linear bijection on the level's index bits, so it is only a
   * storage permutation (the fit engine permutes the table on the way in and out, `immoco.py:FitEngine`).
   * Why: the two dim-0 corners of a cell are indices r and r ^ (2^t - 1); the Gray code turns that into ONE
   * flipped bit (t - 1), and the exchange moves the one group whose t is large (coordinate +1: t = log2(res)
   * + 1) into the low bits, so both corners of every lane pair lie in one 128-byte line.
   * Used for power-of-two hashed levels only; ignored elsewhere. */
  uint32_t swizzle[IMMOCO_MAX_LEVELS];
  /* swizzle[l] == IMMOCO_LAYOUT_LUT: the row of hash index r is an ARBITRARY linear bijection S_l on the index
   * bits, given as chunk tables: S_l(r) = T[r & 127] ^ T[128 + ((r >> 7) & 63)] ^ T[192 + ((r >> 13) & 63)],
   * T = layout_lut + 256 * l (device memory, 256 words per level; entries <= 2^19).  Built by
   * miccai24_immoco_b200/encoding.py:GridSpec.linear_layout so that the 2 M rows ONE pixel corner needs over all M
   * movement groups and both dim-0 corners fall into one or two 128-byte lines; the grouped kernels
   * (immoco_hashgrid_fwd_grouped / _bwd_grouped) put those 2 M rows on adjacent lanes.  The generic entry points
   * honour the tables through their one-thread-per-point kernels (checkers); the row-sorted (csr) and the fused
   * MLP-scatter entry points reject such descriptors. */
  const uint32_t* layout_lut;
} immoco_grid_desc;
#define IMMOCO_LAYOUT_LUT 0xFFFFFFFFu

/* Column structure of the movement-group masks (src/utils/motion_utils.py:56-109 produces masks
 * that are constant along rows): K[:,l] = static_w[l]*F(I)[:,l] + sum_m w_ml * F(I_m)[:,l].   */
typedef struct immoco_lines {
  int32_t n_groups;            /* M                                                     */
  const int32_t* group_ofs;    /* device, M+1 prefix offsets into line_idx / line_w      */
  const int32_t* line_idx;     /* device, column index l of every (group, line) pair     */
  const float* line_w;         /* device, mask value of that column in that group        */
  const float* static_w;       /* device, W floats: 1 - sum_m masks[m, :, l]             */
  int32_t max_lines;           /* max lines of any one group (shared-memory sizing)      */
} immoco_lines;

/* ---- (1) hash-grid encoding: replaces the encoding half of tcnn.NetworkWithInputEncoding
 *          (src/models/immoco.py:60-65, called at :84-87 and :93) -------------------------- */
int immoco_hashgrid_fwd(const immoco_grid_desc* grid, const float* coords, const float* table,
                        float* enc, int64_t n_points, void* stream);
/* grad_table += scatter(d_enc)  (adjoint of the gather; K12 "kernel_grid_backward") */
int immoco_hashgrid_bwd(const immoco_grid_desc* grid, const float* coords, const float* d_enc,
                        float* grad_table, int64_t n_points, void* stream);

/* level-range variants of the two calls above: only levels [level_begin, level_end) */
int immoco_hashgrid_fwd_levels(const immoco_grid_desc* grid, const float* coords, const float* table,
                               float* enc, int64_t n_points, int32_t level_begin, int32_t level_end,
                               void* stream);
int immoco_hashgrid_bwd_levels(const immoco_grid_desc* grid, const float* coords, const float* d_enc,
                               float* grad_table, int64_t n_points, int32_t level_begin,
                               int32_t level_end, void* stream);
/* The same two passes for a GROUPED 3-D coordinate set: coords[(g * n_pixels + p) * 3 ..] = (t_g, y_p, x_p) -- the
 * first coordinate depends on the group only, the other two on the pixel only (make_grids((M, H, W)),
 * src/models/immoco.py:48-53,78-80).  The 2 M lanes of a "bundle" take the M groups x 2 dim-0 corners of one pixel, so
 * one gather / reduction instruction touches every 128-byte line that holds rows of that pixel corner once for ALL
 * groups (with an IMMOCO_LAYOUT_LUT layout: 2 lines instead of 4 at M = 4).  2 <= n_groups <= 16
 * (IMMOCO_ERR_UNSUPPORTED otherwise); a bundle has the next power of two >= 2 M lanes, the surplus ones idle.
 * Features bit-identical to immoco_hashgrid_fwd on the same table. */
int immoco_hashgrid_fwd_grouped(const immoco_grid_desc* grid, const float* coords, const float* table, float* enc,
                                int64_t n_pixels, int32_t n_groups, void* stream);
int immoco_hashgrid_bwd_grouped(const immoco_grid_desc* grid, const float* coords, const float* d_enc,
                                float* grad_table, int64_t n_pixels, int32_t n_groups, void* stream);
/* Kernel selection for A/B checks: 1 = lane-pair kernels (two adjacent lanes take the two dim-0
 * corners of one point; product path), 0 = one thread per (point, level). Same results to rounding. */
int immoco_set_hashgrid_impl(int32_t pair);
/* > 0: run the lane-pair kernels as persistent grids of this many 256-thread CTAs per SM (caps their
 * SM share when co-running); 0 (default): one CTA per (level, 128-point tile), hardware-balanced. */
int immoco_set_hashgrid_ctas_per_sm(int32_t ctas);
/* Same cap for the BACKWARD scatter kernels only (> 0 overrides the knob above for them): the scatter is
 * paced by the L2 atomic units, so a thin persistent grid loses little and leaves the SMs' thread slots
 * and registers to an SM-bound kernel of the other INR branch. */
int immoco_set_hashgrid_bwd_ctas_per_sm(int32_t ctas);

/* ---- (1b) deterministic hash-grid backward: the coordinates of an IM-MoCo fit are constant
 *          (src/models/immoco.py:72-80 registers them as buffers), so the scatter of
 *          kernel_grid_backward can be turned around ONCE per coordinate set into a row-sorted tap list
 *          (CSR): the taps of physical table row r are taps[row_ptr[r] .. row_ptr[r+1]), each tap = 8 bytes
 *          {uint32 point index, float interpolation weight}, ordered by (point, corner).  The backward
 *          pass is then a gather with a FIXED summation order per row: bit-reproducible, no atomics, and
 *          rows that no point touches are never written.  All arrays are caller-owned device memory. ---- */
typedef struct immoco_grid_csr {
  const uint32_t* row_ptr;   /* n_rows + 1 entries (n_rows = grid->offset[n_levels])                */
  const void* taps;          /* n_taps x {uint32 point, float weight}                               */
  int64_t n_taps;            /* n_points * 2^n_dims * n_levels                                      */
  int64_t n_points;
} immoco_grid_csr;
/* bytes of scratch immoco_hashgrid_csr_build needs (sort keys of one level, double-buffered, + radix-sort
 * temporaries); < 0: invalid arguments */
int64_t immoco_hashgrid_csr_workspace_bytes(const immoco_grid_desc* grid, int64_t n_points);
/* builds row_ptr (n_rows + 1 uint32) and taps (n_taps x 8 bytes) for `coords` ((n_points, n_dims) fp32) */
int immoco_hashgrid_csr_build(const immoco_grid_desc* grid, const float* coords, int64_t n_points,
                              uint32_t* row_ptr, void* taps, void* workspace, int64_t workspace_bytes,
                              void* stream);
/* grad_table[r] = sum_k taps[k].weight * d_enc[level(r)][taps[k].point] for every touched row r (WRITTEN, not
 * accumulated; untouched rows are left alone).  Same values as immoco_hashgrid_bwd up to summation order. */
int immoco_hashgrid_bwd_csr(const immoco_grid_desc* grid, const immoco_grid_csr* csr, const float* d_enc,
                            float* grad_table, void* stream);
/* the same gather with torch.optim.Adam fused in: the thread that owns row r applies the update to
 * params / exp_avg / exp_avg_sq of that row at once (rows nobody touches have g = m = v = 0 for ever and
 * never move under dense Adam either, src/models/immoco.py:149-154).  grad_table may be NULL. */
int immoco_hashgrid_bwd_csr_adam(const immoco_grid_desc* grid, const immoco_grid_csr* csr, const float* d_enc,
                                 float* table, float* exp_avg, float* exp_avg_sq, float* grad_table, double lr,
                                 double beta1, double beta2, double eps, int32_t step, void* stream);

/* ---- (1c) tap-indexed storage of the hashed levels.  With constant coordinates the rows a point touches can be
 *          looked up once (immoco_hashgrid_tap_rows: level-relative physical rows, [level][point][2^n_dims]
 *          uint32) and the hashed levels stored in ANY row order; the caller ranks the rows by first touch
 *          (miccai24_immoco_b200/immoco.py:GridTaps), which puts the rows nobody ever touches -- 47 % of the 2-D
 *          image grid's hashed rows at 320 x 320: g = m = v = 0 for ever under torch.optim.Adam
 *          (src/models/immoco.py:149-154) -- behind the n_active_rows live ones, where Adam and the gradient
 *          memset never go, and makes a pixel's first-touched rows consecutive.  `rows` then holds ABSOLUTE
 *          physical rows (relative to the table base) of the levels >= first_level, 16-byte aligned; levels
 *          below first_level keep their own index arithmetic.  2-D grids only. ---- */
typedef struct immoco_grid_taps {
  const uint32_t* rows;      /* (n_levels - first_level) x n_points x 4 absolute rows; NULL = not used      */
  int32_t first_level;       /* levels [first_level, n_levels) are tap-indexed                               */
  int32_t reserved;
  int64_t n_points;
  int64_t n_active_rows;     /* rows [0, n_active_rows) of the table hold everything any point touches       */
} immoco_grid_taps;
int immoco_hashgrid_tap_rows(const immoco_grid_desc* grid, const float* coords, int64_t n_points,
                             int32_t level_begin, int32_t level_end, uint32_t* rows, void* stream);
/* immoco_hashgrid_fwd / immoco_hashgrid_bwd over a table whose levels >= taps->first_level are stored in the
 * order `taps->rows` describes (forward: one launch, bit-identical features; backward: dense-level launch +
 * one indexed launch) */
int immoco_hashgrid_fwd_taps(const immoco_grid_desc* grid, const immoco_grid_taps* taps, const float* coords,
                             const float* table, float* enc, int64_t n_points, void* stream);
int immoco_hashgrid_bwd_taps(const immoco_grid_desc* grid, const immoco_grid_taps* taps, const float* coords,
                             const float* d_enc, float* grad_table, int64_t n_points, void* stream);

/* ---- (2) INR MLP: replaces the network half of tcnn.NetworkWithInputEncoding
 *          (configs at src/models/immoco.py:11-25).  One hidden layer of `width` (64 or 256),
 *          no biases, W1: width x 32, W2: 16 x width (rows >= 2 are padding).
 *          out[n][2]; out_tanh applies the outer .tanh() of immoco.py:93. ------------------- */
int immoco_mlp_fwd(const float* enc, const float* w1, const float* w2, float* out,
                   int64_t n_points, int32_t width, int32_t act, int32_t out_tanh, void* stream);
/* d_out is the cotangent of the PRE-tanh output. d_enc is written; g_w1/g_w2 are accumulated. */
int immoco_mlp_bwd(const float* enc, const float* w1, const float* w2, const float* d_out,
                   float* d_enc, float* g_w1, float* g_w2, int64_t n_points, int32_t width,
                   int32_t act, void* stream);
/* Deterministic weight gradients: every CTA of the tensor-core backward kernel writes ITS partial
 * [g_W1 | g_W2] (same layout as the MLP parameters, width*32 + 16*width floats) to
 * g_part[cta * n_mlp ..]; the CTA -> tile assignment is static, so the partials are reproducible and
 * the consumer (immoco_adam_step_partials) adds them in CTA order.  g_part must hold
 * immoco_mlp_bwd_partial_count(n_points) blocks and be ZERO outside the rows the kernel writes (the 14
 * padding rows of W2).  Tensor-core implementation only. */
int immoco_mlp_bwd_partials(const float* enc, const float* w1, const float* w2, const float* d_out,
                            float* d_enc, float* g_part, int64_t n_points, int32_t width, int32_t act,
                            void* stream);
int immoco_mlp_bwd_partial_count(int64_t n_points);
/* 64-wide network over a 3-D hash grid (the Motion INR, src/models/immoco.py:19-25,64-65): immoco_mlp_bwd with the
 * table gradients of the hashed levels reduced straight from the dE tile inside the kernel (lane pairs = the two
 * dim-0 corners of a point, same arithmetic as immoco_hashgrid_bwd) -- no feature-plane round trip for those
 * levels.  d_enc receives the DENSE levels only; immoco_hashgrid_bwd_dense_levels finishes them.  Together the two
 * calls equal immoco_mlp_bwd + immoco_hashgrid_bwd up to the order of the float atomics. */
int immoco_mlp_bwd_scatter(const float* enc, const float* w1, const float* w2, const float* d_out,
                           float* d_enc, float* g_w1, float* g_w2, const immoco_grid_desc* grid,
                           const float* coords, float* grad_table, int64_t n_points, int32_t width,
                           int32_t act, void* stream);
int immoco_hashgrid_bwd_dense_levels(const immoco_grid_desc* grid, const float* coords, const float* d_enc,
                                     float* grad_table, int64_t n_points, void* stream);
/* 1: immoco_fit_run uses the pair above for the motion branch of the float-atomic path; 0: separate kernels. */
int immoco_set_fused_scatter(int32_t on);
int immoco_get_fused_scatter(void);
/* (The fp32 SIMT A/B kernels of round 1 are no longer part of this library: they are built into the
 * test-side checker tests/checkers/_mlp_simt.so.  The product MLPs are tcgen05 kind::tf32, 3xTF32 split,
 * accumulators in TMEM.) */
/* element-wise helper for the autograd wrapper: d_pre = d_post * (1 - y^2) */
int immoco_tanh_bwd(const float* y, const float* d_post, float* d_pre, int64_t n, void* stream);

/* ---- (3) centred un-normalised 2-D FFT: replaces FFT/IFFT (src/utils/data_utils.py:29-34)
 *          inverse=0: forward (exp(-i..)); inverse=1: conjugate transform, scaled by `scale`
 *          (1/(H*W) gives IFFT, 1 gives the adjoint of FFT).  tw_h / tw_w: exp(-2 pi i t/N). -- */
int immoco_fft2c(const float* in, float* out, float* tmp, int32_t batch, int32_t h, int32_t w,
                 const float* tw_h, const float* tw_w, int32_t inverse, float scale, void* stream);

/* ---- (4) motion forward model: replaces IMMoCo.forward's grid_sample + FFT + mask-combine
 *          (src/models/immoco.py:91-111).  image (H,W) complex, disp (M,H,W,2) = tanh output,
 *          ident (H,W,2) identity grid.  k_out (H,W) complex.  c_tmp (H,W) complex scratch. -- */
int immoco_forward_model(const float* image, const float* disp, const float* ident,
                         const immoco_lines* lines, const float* tw_h, const float* tw_w,
                         float* c_tmp, float* k_out, int32_t h, int32_t w, void* stream);
/* adjoint: d_k (H,W) complex cotangent of k_out -> d_image (accumulated, +=) and d_disp (M,H,W,2,
 * written): cotangent of disp, or of the PRE-tanh motion-INR output when pre_tanh != 0
 * (multiplied by 1 - disp^2, fusing the backward of the .tanh() at immoco.py:93). */
int immoco_forward_model_bwd(const float* d_k, const float* image, const float* disp,
                             const float* ident, const immoco_lines* lines, const float* tw_h,
                             const float* tw_w, float* c_tmp, float* d_image, float* d_disp,
                             int32_t pre_tanh, int32_t h, int32_t w, void* stream);

/* ---- (5) fused column pass of the fit loop: K = colFFT(C); loss_acc[0] += sum|K-K_in|^2;
 *          d_c = colFFT^H((K-K_in)/(H*W))  (F.mse_loss, src/models/immoco.py:170-171) ------- */
int immoco_colpass_loss(const float* c, const float* k_in, float* k_out, float* d_c,
                        double* loss_acc, const float* tw_h, int32_t h, int32_t w, void* stream);

/* ---- (6) gradient-entropy prior (src/utils/losses.py:20-40): loss_acc[0] += GE(image);
 *          d_image = grad_scale * dGE/dimage (written when accumulate==0, else +=) --------- */
int immoco_grad_entropy(const float* image, float grad_scale, double* loss_acc, float* d_image,
                        int32_t accumulate, int32_t h, int32_t w, void* stream);

/* ---- (7) fused Adam (torch.optim.Adam defaults, src/models/immoco.py:149-154,166,175):
 *          m,v,p updated in one pass, grad zeroed (optimizer.zero_grad fused). ------------- */
int immoco_adam_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                     double lr, double beta1, double beta2, double eps, int32_t step,
                     int32_t zero_grad, void* stream);
/* Adam over the first n_mlp parameters with the gradient taken as sum_{c < n_part} g_part[c * n_mlp + i]
 * (added in CTA order: deterministic); params / moments point at the INR's MLP block. */
int immoco_adam_step_partials(float* params, const float* g_part, int32_t n_part, float* exp_avg,
                              float* exp_avg_sq, int64_t n_mlp, double lr, double beta1, double beta2,
                              double eps, int32_t step, void* stream);
/* tuning knobs of the Adam kernel (tools/adam_bench.py): variant 0..5 = {1,2,4 items per thread} x
 * {plain, streaming cache hints}; resident 256-thread CTAs per SM. Results are identical. */
int immoco_set_adam_tuning(int32_t variant, int32_t ctas_per_sm);

/* ---- (8) composite: iterations [it_begin, it_end) of the optimisation loop
 *          (src/models/immoco.py:164-181) on caller-provided buffers. ---------------------- */
typedef struct immoco_fit {
  int32_t h, w, m;                 /* image rows, phase-encode lines, movement groups        */
  immoco_grid_desc grid_image;     /* 2-D */
  immoco_grid_desc grid_motion;    /* 3-D */
  int32_t width_image, act_image;  /* 256, ReLU */
  int32_t width_motion, act_motion;/* 64, Tanh  */
  /* flat parameter vector [motion (n_motion) | image (n_image)], each [W1 | W2 | table] */
  int64_t n_motion, n_image;
  float* params; float* grads; float* exp_avg; float* exp_avg_sq;
  const float* coords_image;       /* (H*W, 2)  identity grid (x=col, y=row)                 */
  const float* coords_motion;      /* (M*H*W,3) (m,row,col)                                  */
  immoco_lines lines;
  const float* tw_h; const float* tw_w;
  const float* k_in;               /* (H,W) complex, normalised measured k-space             */
  float* enc_image;  float* d_enc_image;    /* 16 x P x 2                                    */
  float* enc_motion; float* d_enc_motion;   /* 16 x M*P x 2                                  */
  float* image; float* d_image;             /* P x 2                                         */
  float* disp;  float* d_disp;              /* M*P x 2                                       */
  float* c_tmp; float* d_c; float* k_out;   /* P x 2 each                                    */
  double* loss;                    /* 2 doubles per iteration: sum|dK|^2, GE (pre-zeroed)    */
  double lr, beta1, beta2, eps;
  /* ---- reproducible accumulation (all optional; NULL / 0 = the float-atomic path) ----------------------
   * loss_slots: per-iteration, per-CTA partial sums of the two loss terms, slots_per_iter =
   *   immoco_fit_loss_slots() doubles per iteration; immoco_fit_run adds them in slot order into `loss` at
   *   the end of the call (no floating-point atomics on the loss).
   * deterministic != 0 additionally requires every buffer below and makes two runs of the same call
   * bit-identical: hash-grid backward = row-sorted gather (csr_*), MLP weight gradients = per-CTA partials
   * added in CTA order (mlp_part_*), image cotangent accumulated in 64-bit fixed point (d_image_fx, scaled
   * per iteration by the largest column-pass cotangent, dc_max_bits[it]).  fuse_adam != 0: the table rows
   * are updated inside the gather kernel (no gradient round trip through HBM). */
  double* loss_slots;
  int32_t deterministic;
  int32_t fuse_adam;
  immoco_grid_csr csr_image;
  immoco_grid_csr csr_motion;
  float* mlp_part_image;           /* immoco_mlp_bwd_partial_count(P)   x n_mlp_image floats, zeroed */
  float* mlp_part_motion;          /* immoco_mlp_bwd_partial_count(M*P) x n_mlp_motion floats, zeroed */
  int64_t* d_image_fx;             /* 2*P int64, zeroed by the caller once                    */
  uint32_t* dc_max_bits;           /* one word per iteration, zeroed by the caller            */
  /* ---- float-atomic path only (ignored when deterministic): taps_image.rows != NULL = the image table's levels
   * >= first_level are stored tap-indexed (section 1c): the image branch runs immoco_hashgrid_fwd_taps /
   * _bwd_taps, and Adam + gradient zeroing stop after the MLP block + 2 * n_active_rows floats. */
  immoco_grid_taps taps_image;
} immoco_fit;
/* doubles per iteration of immoco_fit::loss_slots for an (h, w) slice: out[0] column-pass CTAs (data
 * consistency), out[1] gradient-entropy CTAs; slots_per_iter = out[0] + out[1] */
int immoco_fit_loss_slots(int32_t h, int32_t w, int32_t out[2]);
/* process-wide default consulted by the shipped host code when it builds a fit (the C ABI itself takes the
 * mode per call through immoco_fit::deterministic): 1 = bit-reproducible fits. */
int immoco_set_deterministic(int32_t on);
int immoco_get_deterministic(void);
/* destroys the auxiliary streams / events immoco_fit_run created for caller streams of the current device
 * (the library's only hidden state); safe to call when no fit is in flight.  Returns the number freed. */
int immoco_release_streams(void);

/* Optional per-kernel timing: events are recorded around every kernel of each iteration `it`
 * with it % profile_every == profile_every - 1 (prof may be NULL).  Slot order:
 * 0 hashgrid_fwd_image, 1 mlp_fwd_image, 2 hashgrid_fwd_motion, 3 mlp_fwd_motion, 4 fft_rows,
 * 5 motion_rows_fwd, 6 colpass_loss, 7 grad_entropy, 8 fft_rows_adj, 9 motion_rows_bwd,
 * 10 mlp_bwd_motion, 11 hashgrid_bwd_motion, 12 mlp_bwd_image, 13 hashgrid_bwd_image, 14 adam_motion,
 * 15 adam_image. */
#define IMMOCO_PROFILE_SLOTS 16
typedef struct immoco_profile immoco_profile;
immoco_profile* immoco_profile_create(int32_t capacity);
void immoco_profile_destroy(immoco_profile* prof);
int immoco_profile_read(immoco_profile* prof, float* ms_sum);
/* begin / end (ms after the iteration's first event) of every slot of instrumented iteration
 * `index`: a two-stream timeline when immoco_set_profile_overlap(1) is in force. */
int immoco_profile_timeline(immoco_profile* prof, int32_t index, float* begin_ms, float* end_ms);
int immoco_set_profile_overlap(int32_t on);

int immoco_fit_run(const immoco_fit* fit, int32_t it_begin, int32_t it_end,
                   const float* lambdas_host, void* stream, immoco_profile* prof,
                   int32_t profile_every);
/* The same loop over a BATCH of n_fits (<= immoco_max_fit_batch()) independent fits of ONE shape (h, w, m,
 * network widths, grid layouts, accumulation mode), advanced in lock step with the same lambda schedule: the
 * latency-bound kernels of an iteration (row / column passes, gradient entropy, MLP forward) are issued once
 * for all instances (grid z / y = instance), the kernels that fill the GPU alone (hash grid, MLP backward,
 * Adam) per instance.  Each instance gets exactly the result of its own immoco_fit_run (bit for bit in
 * deterministic mode).  This is "within a GPU, batch B instances" of src/test/test_immoco.py:45-72's loop. */
int immoco_fit_run_batched(const immoco_fit* const* fits, int32_t n_fits, int32_t it_begin, int32_t it_end,
                           const float* lambdas_host, void* stream, immoco_profile* prof,
                           int32_t profile_every);
int immoco_max_fit_batch(void);
/* 1 (default): the image-INR branch of every non-instrumented iteration runs on an internal
 * auxiliary stream, forked from / joined to `stream` with events; 0: everything on `stream`. */
int immoco_set_branch_overlap(int32_t on);
/* 1 (default): kernels are launched with programmatic dependent launch (griddepcontrol): a kernel's
 * launch and prologue overlap its predecessor's tail; 0: plain stream order. Results are identical. */
int immoco_set_pdl(int32_t on);

/* ---- (9) evaluation metrics on the device: replaces calmetric2D's normalize / my_psnr / piq.ssim /
 *          rmse (src/utils/evaluate.py:19-47,57-80; caller src/test/test_immoco.py:74-85).
 *          pred / gt: (batch, h, w) strided views (strides in elements; a complex view is read as its
 *          magnitude).  minmax: batch*4 floats scratch.  acc: batch*4 doubles ZEROED by the caller,
 *          receives {sum of squared error of the min-max-normalised images, sum of the SSIM map,
 *          SSIM map size, -}.  Gaussian window sigma 1.5, kernel_size <= 11, pool = piq's
 *          down-sampling factor max(1, round(min(h,w)/256)). -------------------------------------- */
int immoco_metrics2d(const float* pred, int64_t pred_img_stride, int64_t pred_row_stride,
                     int32_t pred_complex, const float* gt, int64_t gt_img_stride, int64_t gt_row_stride,
                     int32_t gt_complex, int32_t batch, int32_t h, int32_t w, int32_t kernel_size,
                     int32_t pool, float* minmax, double* acc, void* stream);

/* HaarPSI of the same min-max-normalised image pairs (piq.haarpsi(data_range=1, scales=3), evaluate.py:76):
 * minmax = what immoco_metrics2d wrote for these views; pooled = scratch of batch*2*hp*wp floats with
 * hp = (h + d) / 2, wp = (w + d) / 2, d = max(h % 2, w % 2); acc (batch*2 doubles, ZEROED by the caller) receives
 * {sum of sigmoid(alpha * similarity) * weight, sum of weight}; the score is (logit((acc0 + eps) / (acc1 + eps)) /
 * alpha)^2 with eps = FLT_EPSILON.  piq is not available here: parity unpinned (restated algorithm). */
int immoco_haarpsi(const float* pred, int64_t pred_img_stride, int64_t pred_row_stride, int32_t pred_complex,
                   const float* gt, int64_t gt_img_stride, int64_t gt_row_stride, int32_t gt_complex,
                   int32_t batch, int32_t h, int32_t w, float c, float alpha, const float* minmax, float* pooled,
                   double* acc, void* stream);

/* ---- (10) synthetic rigid motion: the image-domain half of motion_simulation2D
 *          (src/utils/motion_utils.py:121-202).  theta: (n_mov, 6) affine rows as handed to
 *          F.affine_grid(align_corners=True); out: (n_mov, h, w) complex, bilinear, border padding,
 *          sampled with align_corners=False (motion_utils.py:165-186). ------------------------------ */
int immoco_rigid_resample(const float* image, const float* theta, float* out, int32_t n_mov, int32_t h,
                          int32_t w, void* stream);
/* k[:, w0[m]:w1[m]] = k_moved[m][:, w0[m]:w1[m]] for m in order; mask (int64, may be NULL) gets 1 there
 * (motion_utils.py:188-198).  w0 / w1 are device arrays. */
int immoco_replace_lines(float* k, const float* k_moved, int64_t* mask, const int32_t* w0, const int32_t* w1,
                         int32_t n_mov, int32_t h, int32_t w, void* stream);

/* ---- (11) kld-net inference: the kernels of fastmri.models.Unet (== src/models/unet.py:17-187 with
 *          InstanceNorm2d), NCHW fp32, as used at src/test/test_immoco.py:17-20,50-58.
 *          conv3x3: pad 1, no bias, over the channel concat [in0 (c0) | in1 (c1, may be NULL/0)];
 *          writes the RAW output and accumulates per-(image, channel) {sum, sum of squares} into
 *          stats (n*cout*2 doubles, ZEROED by the caller).  convt2x2: ConvTranspose2d(k=2, s=2, no
 *          bias), weight (cin, cout, 2, 2), out (n, cout, 2h, 2w), same statistics.  instnorm_lrelu:
 *          in place over `planes` = n*c planes, biased variance, then LeakyReLU(slope); `pooled` (may be
 *          NULL) receives the 2x2 average pool (planes, h/2, w/2).  conv1x1: with bias (may be NULL). -- */
int immoco_unet_conv3x3(const float* in0, int32_t c0, const float* in1, int32_t c1, const float* weight,
                        float* out, double* stats, int32_t n, int32_t cout, int32_t h, int32_t w, void* stream);
/* The same convolution on the tensor cores (tcgen05 kind::tf32, 3xTF32 split, implicit GEMM over shifted views
 * of one staged input window): weights packed ONCE per layer by immoco_unet_pack_conv3x3 into w_hi / w_lo
 * ((cin / 4) * 9 * cout * 4 floats each: [channel quad][tap][cout][4], tf32 hi and lo parts).  Requires
 * (c0 + c1) % 8 == 0, c0 % 4 == 0, cout % 32 == 0 (IMMOCO_ERR_UNSUPPORTED otherwise: the 2-channel input
 * layer stays on immoco_unet_conv3x3).  Same outputs / statistics contract. */
int immoco_unet_pack_conv3x3(const float* weight, float* w_hi, float* w_lo, int32_t cout, int32_t cin, void* stream);
int immoco_unet_conv3x3_tc(const float* in0, int32_t c0, const float* in1, int32_t c1, const float* w_hi,
                           const float* w_lo, float* out, double* stats, int32_t n, int32_t cout, int32_t h,
                           int32_t w, void* stream);
int immoco_unet_convt2x2(const float* in, const float* weight, float* out, double* stats, int32_t n,
                         int32_t cin, int32_t cout, int32_t h, int32_t w, void* stream);
int immoco_unet_instnorm_lrelu(float* x, const double* stats, float* pooled, int32_t planes, int32_t h,
                               int32_t w, float eps, float slope, void* stream);
int immoco_unet_conv1x1(const float* in, const float* weight, const float* bias, float* out, int32_t n,
                        int32_t cin, int32_t cout, int32_t hw, void* stream);

/* ---- (12) autofocusing baseline (src/models/autofocusing.py:69-84): n_mov complex images, each warped
 *          by its own affine theta (n_mov, 6) as F.grid_sample(F.affine_grid(theta, align_corners=True),
 *          mode="bicubic", zeros padding, align_corners=False); the backward pass reduces the output
 *          cotangent straight to d theta (n_mov*6 doubles, ZEROED by the caller). ----------------------- */
int immoco_rigid_bicubic_fwd(const float* images, const float* theta, float* out, int32_t n_mov, int32_t h,
                             int32_t w, void* stream);
int immoco_rigid_bicubic_bwd_theta(const float* images, const float* theta, const float* d_out, double* d_theta,
                                   int32_t n_mov, int32_t h, int32_t w, void* stream);

/* library/ABI version and the number of kernel launches one fit iteration issues */
int immoco_abi_version(void);
/* sizeof(immoco_grid_desc), sizeof(immoco_lines), sizeof(immoco_fit), sizeof(immoco_grid_csr): lets a
 * foreign-language binding assert that its struct mirrors match this build. */
void immoco_struct_sizes(int32_t out[5]);
int immoco_launches_per_iteration(int32_t m);
/* the same for a fit with immoco_fit::deterministic / fuse_adam set */
int immoco_launches_per_iteration_mode(int32_t m, int32_t deterministic, int32_t fuse_adam);
/* 1 (default): immoco_fit_run issues the static row pass and the pruned motion rows of an iteration as ONE
 * launch forward and ONE launch for the adjoint (16 launches per iteration); 0: four separate launches
 * (18 per iteration; A/B check -- results agree to rounding). */
int immoco_set_fused_rows(int32_t on);
/* 1 (default): inside immoco_fit_run Adam leaves the gradients in place and one memset of the gradient vector
 * runs on a third stream during the motion chain's SM-bound stretch (the memory system is idle there);
 * 0: Adam zeroes the gradients itself.  Results are identical. */
int immoco_set_deferred_zero(int32_t on);

#ifdef __cplusplus
}
#endif
#endif /* IMMOCO_B200_H_ */
