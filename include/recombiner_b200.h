/*
 * recombiner_b200.h -- C ABI of the B200-native RECOMBINER hot path.
 *
 * One shared object (recombiner_b200/librecombiner_b200.so, built for sm_100a).
 * The reference (cambridge-mlg/RECOMBINER) has no FFI: its hot path is Python
 * calling ATen.  Each entry point below therefore names the reference *Python*
 * call site(s) it replaces (file:line relative to the reference checkout); the
 * Python host mirror in recombiner_b200/ binds them with ctypes.
 *
 * Conventions
 *  - every function returns 0 on success, <0 on error; rcb_last_error() gives a
 *    thread-local message.  Nothing here has a CPU fallback.
 *  - all pointers are BORROWED DEVICE pointers (caller allocates outputs and
 *    workspaces, the library never retains them past the call);
 *  - `stream` is a cudaStream_t (pass torch.cuda.current_stream().cuda_stream);
 *    calls are asynchronous w.r.t. the host;
 *  - float = f32, double = f64, int = int32; "rows" = datapoints or patches,
 *    "item" = (row, MC sample) pair with index row*S + s.
 */
#ifndef RECOMBINER_B200_H
#define RECOMBINER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* rcb_stream_t;

int rcb_version(void);
const char* rcb_last_error(void);

/* ------------------------------------------------------------------------- *
 * (a) fit path
 * ------------------------------------------------------------------------- */

/* Effective posterior -> reparameterised samples.
 * Replaces test_model.py:286-303 (mask-mix, column un-permute, group->param
 * gather, lpe draw) and utils.py:142-198 (hierarchical weight draw, one call per
 * level).  For element (row n, MC sample s, parameter p):
 *     q = g2p[p];  rr = row_map ? row_map[n] : n;  src row r = perm ? perm[rr*P + q] : rr
 *     mu~ = loc*(1-m) + sample*m,  sigma~ = softplus(log_scale)/6*(1-m) + 1e-15*m
 *     value = mu~ + sigma~ * eps
 * p <  n_w : written (accumulate=0) or added (accumulate=1) to hw[(n*S+s)*ld_hw + p]
 * p >= n_w : written to lpe[(n*S+s)*n_l + (p-n_w)]            (lpe may be NULL)
 * eps: explicit tensors (eps_w (rows,S,n_w), eps_l (S,rows,n_l)) when non-NULL,
 * else counter-based Philox4x32-10 keyed by (seed, step, tensor_id, element). */

/* Per-step scalars kept in device memory, so that a captured CUDA graph of one fit step can be replayed with a new
 * noise key and new Adam bias corrections: rcb_set_step_state writes them (arguments by value, stream-ordered),
 * and kernels given a non-NULL `dyn` read them instead of the by-value fields of their argument block. */
typedef struct {
  int64_t seed;            /* Philox key (the fit loop folds the epoch into its low word) */
  int step;
  float adam_step_size;    /* lr / (1 - b1^t) */
  float adam_bc2_sqrt;     /* sqrt(1 - b2^t) */
  float beta_scalar;       /* >= 0: replaces rcb_update_args.beta_scalar (the global beta of prior training,
                              prior_model.py:247); < 0: unused */
} rcb_step_state;
int rcb_set_step_state(rcb_step_state* dev, int64_t seed, int step, float adam_step_size, float adam_bc2_sqrt,
                       float beta_scalar, rcb_stream_t stream);

/* Adam step on one flat fp32 vector with torch.optim.Adam's default arithmetic (prior_model.py:224-227,253: the
 * optimiser of the shared mappings A_l / upsampler; the posteriors' Adam is fused into rcb_fit_update).
 * step_size = lr / (1 - b1^t), bc2_sqrt = sqrt(1 - b2^t); with dyn != NULL both are read from device memory so
 * that a captured step can be replayed.  All four buffers 16-byte aligned. */
int rcb_adam_flat(float* theta, const float* grad, float* m, float* v, int64_t n, float step_size, float bc2_sqrt,
                  float b1, float b2, float eps, const rcb_step_state* dyn, rcb_stream_t stream);

/* out[0] = scale * sum(sqerr[0..n)), out[1] = kl[0] (kl may be NULL): the two loss terms of one prior-training
 * step (prior_model.py:237-253) kept on the device in f64 instead of two .item() round trips per step. */
int rcb_step_stats(const float* sqerr, int n, double scale, const double* kl, double* out, rcb_stream_t stream);

typedef struct {
  const float* loc;        /* (src_rows, P) group order */
  const float* log_scale;  /* (src_rows, P) */
  const float* mask;       /* (src_rows, P) or NULL */
  const float* sample;     /* (src_rows, P) or NULL */
  const int* g2p;          /* (P) group_to_param, or NULL = identity */
  const int* perm;         /* (src_rows, P) per-column row permutation, or NULL */
  const int* row_map;      /* (rows) level-2/3 expansion: row -> src row, or NULL */
  const float* eps_w;      /* (rows, S, n_w) or NULL */
  const float* eps_l;      /* (S, rows, n_l) or NULL */
  float* hw;               /* (rows*S, ld_hw) */
  float* lpe;              /* (rows*S, n_l) or NULL; stitched (data*S, sp_total*lpe_c) when lpe_slot != NULL */
  const int* lpe_slot;     /* patch modalities: (rows_per_datum, n_l/lpe_c) slot of each latent position in the
                              datum's stitched grid (utils.py:71-90), or NULL */
  float* eps_w_store;      /* optional (rows, S, n_w): keep the generated noise for rcb_fit_update */
  float* eps_l_store;      /* optional (S, rows, n_l) */
  int64_t seed;
  int64_t row_offset;      /* global index of row 0 (multi-GPU shards) */
  int rows, S, P, n_w, n_l, ld_hw;
  int step, tensor_id, accumulate;
  int rows_per_datum, sp_total, lpe_c;   /* used with lpe_slot */
  const rcb_step_state* dyn;             /* optional: seed and step read from device memory */
  void* lpe_h;                           /* optional: the latent grid is written here as fp16 INSTEAD of lpe (rcb_gemm_tc_hh) */
  void* hw_h;                            /* optional: the weight samples are written here as fp16 INSTEAD of hw (row stride ld_hw
                                            fp16 elements; not with accumulate) for rcb_gemm_tc_h */
  const int* p2g;                        /* optional: inverse of g2p (parameter index of a stored column), lets the row
                                            kernel read the posterior coalesced */
  int fast_math;                         /* row-wise kernel only: softplus through the ex2 / lg2 approximations (as
                                            rcb_update_args.fast_math); 0 in the fp32 parity configuration */
} rcb_sample_args;
int rcb_fit_sample(const rcb_sample_args* a, rcb_stream_t stream);

/* C[M,N] = A[M,K] @ B[K,N] (+bias[n % bias_mod], LeakyReLU 0.01 if act=1).
 * Replaces the shared-operand `mm` of the linear reparameterisation
 * (test_model.py:348-349, prior_model.py:173-174) and its backward, and the
 * first upsampler stage on small latent grids (prior_model.py:48-50) once the
 * nearest-upsample + conv is folded into a dense map (rcb_fold_dense).
 * lda/ldb/ldc in elements, all multiples of 4; trans_a: A given as [K,M]. */
int rcb_gemm(const float* A, int lda, const float* B, int ldb, float* C, int ldc,
             int M, int N, int K, const float* bias, int bias_mod, int act,
             int trans_a, int accumulate, rcb_stream_t stream);

/* Tensor-core variant of rcb_gemm: C[M,N] = A[M,K] @ Bt[N,K]^T with TF32 tcgen05.mma
 * (fp32 accumulation in TMEM), TMA-fed, both operands K-major -- B is passed
 * TRANSPOSED ([N,K] row-major, leading dimension ldbt).  Same epilogue as rcb_gemm.
 * Tolerance vs fp32: ~1e-3 relative (10-bit mantissa operands), stated in the tests. */
int rcb_gemm_tc(const float* A, int lda, const float* Bt, int ldbt, float* C, int ldc,
                int M, int N, int K, const float* bias, int bias_mod, int act,
                int accumulate, rcb_stream_t stream);

/* rcb_gemm_tc with C written as fp16 (ldc in fp16 elements, N % 4 == 0, no accumulation into C). */
int rcb_gemm_tc_oh(const float* A, int lda, const float* Bt, int ldbt, void* C_h, int ldc,
                   int M, int N, int K, const float* bias, int bias_mod, int act, rcb_stream_t stream);

/* nb <= 4 products C_i[M,N_i] = A_i[M,K_i] @ Bt_i[N_i,K_i]^T in ONE launch (the per-layer reparameterisation
 * products h_w[seg_l] @ A_l, test_model.py:348-349, and their data gradients).  The arrays are host arrays of nb
 * entries; A_i share lda, C_i share ldc; in_half: A_i and Bt_i are fp16; C_i is multiplied by out_scale.  No bias /
 * activation / accumulation. */
int rcb_gemm_tc_batch(int nb, const void* const* A, int lda, const void* const* Bt, const int* ldbt,
                      float* const* C, int ldc, int M, const int* N, const int* K, int in_half, float out_scale,
                      rcb_stream_t stream);
/* fp16 A and fp16 Bt, fp32 C (K % 8 == 0, lda % 8 == 0, ldbt % 8 == 0): same epilogue as rcb_gemm_tc */
int rcb_gemm_tc_h(const void* A_h, int lda, const void* Bt_h, int ldbt, float* C, int ldc,
                  int M, int N, int K, const float* bias, int bias_mod, int act, int accumulate, rcb_stream_t stream);
/* fp16 A, fp16 Bt and fp16 C (K % 8 == 0; kind::f16 MMAs, twice the TF32 rate, fp32 accumulation) */
int rcb_gemm_tc_hh(const void* A_h, int lda, const void* Bt_h, int ldbt, void* C_h, int ldc,
                   int M, int N, int K, const float* bias, int bias_mod, int act, rcb_stream_t stream);

/* Geometry of one nearest-upsample + 'same' conv stage on a channel-last grid
 * (item, d, h, w, c).  2-D signals use d=1, fz=1, kz=1; 1-D also h=1, fy=1, ky=1.
 * prior_model.py:29-45. */
typedef struct {
  int d, h, w;     /* source grid */
  int fz, fy, fx;  /* nearest-upsample factors */
  int kz, ky, kx;  /* kernel extent (odd), padding (k-1)/2 */
  int ic, oc;
} rcb_upconv_geom;

/* Fold nearest-upsample into the conv taps (polyphase form): for phase (ry,rx)
 * and tap (ty,tx) in {0,1}^2,
 *   w_eff[rz][ry][rx][tz][ty][tx][ic][oc] = sum of w[oc][ic][kz][ky][kx] over the kernel taps
 *   that land on source offset (bz[rz]+tz, by[ry]+ty, bx[rx]+tx).
 * w: torch conv layout (oc, ic, kz, ky, kx).  w_eff_t is the [..][oc][ic] transpose
 * used by the data-gradient kernel. */
int rcb_fold_poly(const float* w, const rcb_upconv_geom* g, float* w_eff, float* w_eff_t,
                  rcb_stream_t stream);
/* Forward weights in K-major form for the tensor-core path:
 * w_eff_k[phase][oc][tap*ic + c] (same values as w_eff). */
int rcb_fold_poly_k(const float* w, const rcb_upconv_geom* g, float* w_eff_k, rcb_stream_t stream);
/* Dense fold for tiny grids (the 5-tap first stage, factor >= 4): m[(sy,sx,ic)][(oy,ox,oc)], and its transpose.
 * Only the structurally non-zero entries are written (an output pixel's taps reach at most 2 x 2 source pixels): m and
 * m_t must be ZERO before the first call; later calls with the same geometry overwrite the same entries. */
int rcb_fold_dense(const float* w, const rcb_upconv_geom* g, float* m, float* m_t,
                   rcb_stream_t stream);

/* out[item, y*fy+ry, x*fx+rx, :] = act(bias + sum_taps w_eff . src[item, y+.., x+.., :])
 * Replaces prior_model.py:52-57 (up2/conv2/act2, up3/conv3) with the upsample
 * never materialised. */
int rcb_upconv_fwd(const float* src, const float* w_eff, const float* bias, float* out,
                   const rcb_upconv_geom* g, int items, int act, rcb_stream_t stream);
/* d_src = (transpose of the above)(d_out) [* lrelu'(src_act) if src_act != NULL] */
int rcb_upconv_bwd(const float* d_out, const float* w_eff_t, const float* src_act, float* d_src,
                   const rcb_upconv_geom* g, int items, rcb_stream_t stream);

/* Tensor-core (tcgen05, TF32) variants of rcb_upconv_fwd / rcb_upconv_bwd: implicit GEMM whose
 * A operand is gathered by 5-D TMA boxes of the channel-last activation tensor (tap shift in
 * the forward, stride-f traversal in the data gradient; borders = TMA out-of-bounds zero fill).
 * fwd takes w_eff_k (rcb_fold_poly_k); bwd takes w_eff (rcb_fold_poly).  ic % 32 == 0. */
int rcb_upconv_fwd_tc(const float* src, const float* w_eff_k, const float* bias, float* out,
                      const rcb_upconv_geom* g, int items, int act, rcb_stream_t stream);
int rcb_upconv_bwd_tc(const float* d_out, const float* w_eff, const float* src_act, float* d_src,
                      const rcb_upconv_geom* g, int items, rcb_stream_t stream);
/* The x2 / 3-tap / 64 -> 16 channel stage (the last one of the 2-D upsamplers, prior_model.py:47-59) with fp16
 * source activations and fp16 weights (rcb_to_half of w_eff_k; other stages with ic % 64 == 0 run the general kernel
 * with kind::f16 MMAs): fp16 keeps the same 10 mantissa bits the TF32
 * MMAs read of an fp32 operand, at half the bytes.  fp32 accumulation, bias, activation and output. */
int rcb_upconv_fwd_tc_h(const void* src_h, const void* w_eff_k_h, const float* bias, float* out,
                        const rcb_upconv_geom* g, int items, int act, rcb_stream_t stream);
/* fp16 in and fp16 out (any stage with ic % 64 == 0): the persistent resident-weight kernels for the x2 / 3-tap stages
 * with 64 -> 64 channels (source grid 8 x 8, or >= 16 lines with h % 4 == 0) and 64 -> 16 channels (>= 16 lines), the
 * general kernel with kind::f16 MMAs otherwise */
int rcb_upconv_fwd_tc_hh(const void* src_h, const void* w_eff_k_h, const float* bias, void* out_h,
                         const rcb_upconv_geom* g, int items, int act, rcb_stream_t stream);
/* rcb_upconv_fwd_tc writing its activations as fp16 (for a following rcb_upconv_fwd_tc_h stage), and
 * rcb_upconv_bwd_tc reading the LeakyReLU mask from such fp16 activations (only the signs are used). */
int rcb_upconv_fwd_tc_oh(const float* src, const float* w_eff_k, const float* bias, void* out_h,
                         const rcb_upconv_geom* g, int items, int act, rcb_stream_t stream);
int rcb_upconv_bwd_tc_ah(const float* d_out, const float* w_eff, const void* src_act_h, float* d_src,
                         const rcb_upconv_geom* g, int items, rcb_stream_t stream);
/* Data gradient of the same x2 / 3-tap / 64 -> 16 channel stage with the 64 x 256 weight matrix resident in shared
 * memory (w_bwd_k = rcb_fold_poly_bwd_f2 of the rcb_fold_poly output).  act_kind: 0 no LeakyReLU mask, 1 src_act is
 * the producing stage's fp32 activations, 2 its fp16 activations. */
int rcb_fold_poly_bwd_f2(const float* w_eff, const rcb_upconv_geom* g, float* w_bwd_k, rcb_stream_t stream);
int rcb_upconv_bwd_f2(const float* d_out, const float* w_bwd_k, const void* src_act, int act_kind, float* d_src,
                      const rcb_upconv_geom* g, int items, rcb_stream_t stream);
/* The same kernel leaving d_src as fp16 multiplied by out_scale (clamped to the fp16 range): the operand form of
 * rcb_upconv_bwd_f2w below. */
int rcb_upconv_bwd_f2_oh(const float* d_out, const float* w_bwd_k, const void* src_act, int act_kind, void* d_src_h,
                         float out_scale, const rcb_upconv_geom* g, int items, rcb_stream_t stream);
/* fp16 in as well: d_out_h (any fixed unit, e.g. rcb_mlp_args.d_pe_h) and w_bwd_k_h = rcb_to_half of w_bwd_k; one
 * kind::f16 MMA per (a, b) product instead of two TF32 ones, half the gradient bytes. */
int rcb_upconv_bwd_f2_hh(const void* d_out_h, const void* w_bwd_k_h, const void* src_act, int act_kind, void* d_src_h,
                         float out_scale, const rcb_upconv_geom* g, int items, rcb_stream_t stream);
/* Data gradient of the x2 / 3-tap / 64 -> 64 channel stage (the middle one of the 2-D upsamplers) with fp16 operands
 * and the 16 weight blocks (128 KB) resident in shared memory: d_out_h is the fp16 gradient of the stage's output
 * (scaled, as rcb_upconv_bwd_f2_oh writes it), w_bwd_k_h = rcb_to_half of rcb_fold_poly_bwd_f2w(w_eff), src_act_h the
 * producing stage's fp16 activations (LeakyReLU mask; null = none); d_src is fp32 and multiplied by out_scale_inv.
 * Source grids of 8 x 8 (two items per tile) or >= 16 lines with h % 4 == 0; rcb_upconv_bwd_f2w_eligible says which. */
int rcb_upconv_bwd_f2w_eligible(const rcb_upconv_geom* g);
int rcb_fold_poly_bwd_f2w(const float* w_eff, const rcb_upconv_geom* g, float* w_bwd_k, rcb_stream_t stream);
int rcb_upconv_bwd_f2w(const void* d_out_h, const void* w_bwd_k_h, const void* src_act_h, float* d_src,
                       float out_scale_inv, const rcb_upconv_geom* g, int items, rcb_stream_t stream);
/* rcb_upconv_bwd_f2w leaving d_src as fp16 (saturating) multiplied by out_scale: the A operand of an fp16 GEMM for the
 * stage below (rcb_gemm_tc_batch with in_half, whose out_scale undoes the unit). */
int rcb_upconv_bwd_f2w_oh(const void* d_out_h, const void* w_bwd_k_h, const void* src_act_h, void* d_src_h,
                          float out_scale, const rcb_upconv_geom* g, int items, rcb_stream_t stream);
/* dst[i] = (fp16, round to nearest) src[i] */
int rcb_to_half(const float* src, void* dst, int64_t n, rcb_stream_t stream);

/* Weight gradients of the learned mappings (prior training only; the mappings are
 * frozen at compression time).  Autograd backward of prior_model.py:48-57,173-174.
 *  rcb_upconv_wgrad: d_w_eff[z][tap][ic][oc] = sum_{item,y,x} src_gather * d_out (polyphase
 *      form, split-K with f32 atomics; d_w_eff is zeroed by the call).
 *  rcb_unfold_poly / rcb_unfold_dense: adjoint of the folds -> conv layout (oc, ic, ky, kx).
 *  rcb_colsum: out[c % mod] = sum_r sum_{c' = c (mod mod)} X[r, c']   (bias gradients). */
int rcb_upconv_wgrad(const float* src, const float* d_out, float* d_w_eff, const rcb_upconv_geom* g,
                     int items, rcb_stream_t stream);
int rcb_unfold_poly(const float* d_w_eff, const rcb_upconv_geom* g, float* d_w, rcb_stream_t stream);
int rcb_unfold_dense(const float* d_m, const rcb_upconv_geom* g, float* d_w, rcb_stream_t stream);
int rcb_colsum(const float* x, int64_t rows, int cols, int mod, float* out, rcb_stream_t stream);

/* Fused per-item SIREN MLP (test_model.py:347-355 + loss :624-627 + their
 * autograd backward).  mode 0: forward, writes y_pred.  mode 1: forward +
 * squared-error + backward with dy = coef*(y_pred - y).  mode 2: forward +
 * backward with dy read from `dy`.  Activations never leave shared memory. */
typedef struct {
  const float* wt;      /* (items, ld_w) per-item weights, layer-major, bias first */
  const float* xt;      /* (x_rows, n_f, pix) Fourier features, transposed */
  const float* pe;      /* (items, pix, 16) */
  const float* y;       /* (rows, pix, out) targets (mode 1) */
  const float* dy;      /* (items, pix, out) (mode 2) */
  float* y_pred;        /* (items, pix, out) (mode 0) */
  float* d_pe;          /* (items, pix, 16) (modes 1,2) */
  float* d_wt;          /* (items, ld_w)    (modes 1,2) */
  float* sqerr;         /* (items) sum of squared error (mode 1) */
  const int64_t* pe_base; /* patch modalities: per item, pixel offset of the patch origin inside the stitched
                             pe / d_pe tensors (utils.py:104-116); NULL = item*pix */
  int64_t x_row_stride; /* 0 if all rows share one x */
  int64_t pitch_z, pitch_y; /* pixel pitches of the stitched grid (with pe_base) */
  int items, S, pix, n_f, out, ld_w, mode;
  int ph, pw;           /* patch extent along y and x (pixels); pix = pd*ph*pw (with pe_base) */
  float coef, w0;
  float d_wt_h_scale;   /* with d_wt_h: power-of-two factor applied before the fp16 rounding (undone by the consumer) */
  int ld_wh;            /* row stride of d_wt_h in fp16 elements */
  void* d_wt_h;         /* optional (rcb_mlp_tc, mode 1): the weight gradients are written here as fp16 (clamped to the
                           fp16 range) INSTEAD of d_wt, for an fp16-operand data-gradient GEMM (rcb_gemm_tc_batch) */
  int pe_half;          /* rcb_mlp_tc only: pe holds fp16 values (written by rcb_upconv_fwd_tc_hh) */
  /* rcb_mlp_tc only, optional: generate the Fourier inputs from the pixel index instead of reading xt.  x_tab is the
   * per-axis table rcb_fourier_table writes: for axis a and coordinate index i, 2*x_nfreq floats at
   * x_tab[x_off[a] + i * 2 * x_nfreq] = cos(pi c_i w_j), j < x_nfreq, then sin(pi c_i w_j)  (data/image.py:24-27 on the
   * pixel-centre grid of utils.py:265-284).  The pixel index is row-major over x_size[0 .. x_axes); n_f = 2 * x_axes *
   * x_nfreq.  The (n_f, pix) tensor X -- identical for every datapoint -- is then never read. */
  const float* x_tab;
  int x_axes, x_nfreq;
  int x_size[3], x_off[3];
  void* d_pe_h;         /* optional (rcb_mlp_tc): d pe is written here as fp16 INSTEAD of d_pe, in the units the kernel carries
                           its gradients in -- true d_pe / coef in mode 1, true d_pe * coef in mode 2 -- for the fp16 data
                           gradient of the upsampler (rcb_upconv_bwd_f2_hh, rcb_upconv_bwd_f2w undoes the unit) */
} rcb_mlp_args;
int rcb_mlp(const rcb_mlp_args* a, rcb_stream_t stream);
/* Same contract on tcgen05: two 128-pixel tiles of an item in flight per CTA (one 128-thread group
 * each, a thread per pixel row), two CTAs per SM.
 * Chain products (TF32 operands, fp32 accumulation in TMEM) read their A operand straight from
 * TMEM and are updated in place by the epilogue (tcgen05.ld -> sin / *cos -> tcgen05.st); the
 * weight/bias gradients accumulate in TMEM over the item's tiles from feature-major fp16 copies
 * in shared memory (X in [-1,1]; gradients carried in units of coef).  n_f = 16 only.
 * mode 1 needs coef > 0; in mode 2 `coef` (if > 0) is a scale applied to dy inside the kernel and
 * removed from the outputs (pick ~1/max|dy| so that fp16 keeps its 10-bit mantissa). */
int rcb_mlp_tc(const rcb_mlp_args* a, rcb_stream_t stream);

/* Per-axis Fourier feature table for rcb_mlp_args.x_tab, generated on the device from coordinate indices with the
 * reference's fp32 arithmetic: c_i = -1 + 2 * ((0.5 + i) / size) (utils.py:265-284), t = (c_i * w_j) * pi_f32,
 * tab[i * 2 * n_freq + j] = cos(t), tab[i * 2 * n_freq + n_freq + j] = sin(t).  freq: HOST array of n_freq <= 8
 * frequencies w_j = exp(linspace(0, ln 1024, n_freq)) as the loaders compute them (data/image.py:25). */
int rcb_fourier_table(float* tab, int size, const float* freq, int n_freq, rcb_stream_t stream);

/* Polyphase conv weight gradient on tcgen05 (2-D grids, k > 1): d_w_eff[phase][tap][ic][oc] as produced by
 * rcb_upconv_wgrad, from channel-major copies srcT [3 x-shifts][ic][items*h*w] (rcb_transpose_xshift) and doutT
 * [oc][items][fy*fx phases][h][w] (rcb_transpose_phases).  TF32 operands, fp32 accumulation; split-K partial sums
 * are added atomically. */
int rcb_upconv_wgrad_tc(const float* srcT, const float* doutT, float* d_w_eff, const rcb_upconv_geom* g,
                        int items, rcb_stream_t stream);

/* out[c][r] = in[r][c] (row strides ld_in, ld_out in elements).  Puts the (items, W) tensors of the
 * reparameterisation weight gradient dA_l = hw_l^T d_wt_l (prior_model.py:170-174 under autograd)
 * into the K-major form rcb_gemm_tc takes. */
int rcb_transpose(const float* in, int64_t ld_in, float* out, int64_t ld_out, int rows, int cols,
                  rcb_stream_t stream);
/* out[dx+1][c][r] = in[r+dx][c] inside the line of w pixels that row r belongs to, 0 outside (dx = -1, 0, 1). */
int rcb_transpose_xshift(const float* in, float* out, int64_t rows, int cols, int w, rcb_stream_t stream);
/* Channel-major copy of an upsampled channel-last tensor in (rows = items*(h*fy)*(w*fx), cols) with the rows
 * regrouped by phase: out[c][((item*fy + oy%fy)*fx + ox%fx)*h*w + (oy/fy)*w + ox/fx] = in[(item, oy, ox)][c]. */
int rcb_transpose_phases(const float* in, float* out, int64_t rows, int cols, int h, int w, int fy, int fx,
                         rcb_stream_t stream);

/* Gradient reduction over MC samples + beta-weighted closed-form KL gradient
 * (+ fused Adam).  Replaces the autograd backward of test_model.py:289-303,
 * calculate_kl :357-377 and Adam.step (:635).  Group-order element (r,q):
 *   g_mu  = (1-m) * sum_s d[n,s,p]        + beta*(mu-mu_p)/sig_p^2
 *   g_rho = ((1-m) * sum_s d[n,s,p]*eps   + beta*(sig/sig_p^2 - 1/sig)) * sigmoid(rho)/6
 * with p = p2g[q], n = inverse-permuted row.  adam=1 updates loc/log_scale in
 * place (torch.optim.Adam arithmetic); adam=0 writes g_loc/g_log_scale.
 * kl_out (double[1], optional) += sum beta*KL. */
typedef struct {
  float* loc; float* log_scale;          /* (src_rows, P) */
  const float* mask;                     /* or NULL */
  const float* p_loc; const float* p_log_scale;  /* (P) */
  const float* beta;                     /* (src_rows, G) or NULL -> beta_scalar */
  const int* group_idx;                  /* (P) */
  const int* p2g;                        /* (P) param_to_group or NULL */
  const int* perm_inv;                   /* (src_rows,P) or NULL */
  const int* row_children;               /* level-2/3: (src_rows, n_children) rows fed by src row; NULL = self */
  const float* d_hw;                     /* (rows*S, ld_hw) or NULL */
  const float* d_lpe;                    /* (rows*S, n_l) or NULL */
  const float* eps_w; const float* eps_l;
  const int* lpe_slot;                   /* as in rcb_sample_args, or NULL */
  float* g_loc; float* g_log_scale;      /* adam=0 outputs */
  float* m1_loc; float* v_loc; float* m1_ls; float* v_ls;  /* Adam state */
  double* kl_out;
  int64_t seed; int64_t row_offset;
  int src_rows, rows, n_children, S, P, n_w, n_l, ld_hw, G;
  int step, tensor_id, adam;
  int p_scale_direct;   /* 1: p_log_scale already holds sigma_p (prior training passes scales) */
  int rows_per_datum, sp_total, lpe_c;   /* used with lpe_slot */
  /* Adam: step_size = lr/(1-b1^t) and bc2_sqrt = sqrt(1-b2^t) are computed by the
   * host in f64 exactly as torch.optim.Adam does, then passed as f32. */
  float adam_step_size, adam_bc2_sqrt, b1, b2, adam_eps, beta_scalar, grad_scale;
  const rcb_step_state* dyn;             /* optional: seed, step and the Adam scalars read from device memory */
  /* optional (general layout): per-patch-row sample sums written by rcb_fit_reduce -- red_mu / red_sig (rows, n_w) for
   * the weights with THIS level's noise, red_mu_l / red_sig_l (rows, n_l) for the latent grid (level 1 only).  The data
   * gradient is then read from them instead of d_hw / d_lpe / eps_*. */
  const float* red_mu; const float* red_sig; const float* red_mu_l; const float* red_sig_l;
  int fast_math;        /* row-wise kernel only: softplus / sigmoid / log, the divisions and square roots of the KL gradient
                           and of Adam run on the special-function unit (ex2 / lg2 / rcp / rsqrt approximations, ~2 ulp;
                           log1p by a series below 0.1) instead of the correctly rounded library sequences.  0 in the fp32
                           parity configuration. */
} rcb_update_args;
int rcb_fit_update(const rcb_update_args* a, rcb_stream_t stream);

/* First half of the gradient reduction for the patch modalities (autograd backward of utils.py:142-198 with respect
 * to every level's mean and scale): per patch row and parameter, the sums over the S samples of the data gradient and
 * of gradient x noise for each of the n_levels levels (they share d_hw and differ in their noise).  eps_*: the noise
 * the sampling kernel kept (or the caller injected). */
typedef struct {
  const float* d_hw;            /* (rows*S, ld_hw) */
  const float* d_lpe;           /* stitched latent-grid gradient, or NULL if n_l == 0 */
  const float* eps_w[3];        /* (rows, S, n_w) per level */
  const float* eps_l;           /* (S, rows, n_l) */
  const int* lpe_slot;          /* stitching table of rcb_sample_args, or NULL */
  float* red_mu;                /* (rows, n_w) */
  float* red_sig[3];            /* (rows, n_w) per level */
  float* red_mu_l; float* red_sig_l;   /* (rows, n_l) */
  int rows, S, n_w, n_l, ld_hw, n_levels, rows_per_datum, sp_total, lpe_c;
} rcb_reduce_args;
int rcb_fit_reduce(const rcb_reduce_args* a, rcb_stream_t stream);

/* out[i] = softplus(in[i]) / 6 (threshold 20): the standard-deviation transform `st` of test_model.py:101 /
 * prior_model.py:88 in the arithmetic every kernel here uses, for the q_scale / p_scale inputs of rcb_rec_encode and
 * rcb_rec_decode (encoder and decoder then agree on every bit). */
int rcb_std_transform(const float* in, float* out, int64_t n, rcb_stream_t stream);

/* Per-(row, block) KL in nats, f64 accumulation (test_model.py:384-388). */
int rcb_group_kl(const float* loc, const float* log_scale, const float* p_loc,
                 const float* p_log_scale, const int* group_start, const int* group_end,
                 double* kl, int rows, int P, int G, rcb_stream_t stream);
/* Per-block beta annealing (test_model.py:404-413). coded: uint8 (rows,G). */
int rcb_anneal_beta(float* beta, const double* kl, const uint8_t* coded, int rows, int G,
                    double step, double upper, double lower, double bits, rcb_stream_t stream);
/* Largest-KL not-yet-coded block per row (test_model.py:809-817). */
int rcb_pick_block(const double* kl, const uint8_t* coded, int* block, int rows, int G,
                   rcb_stream_t stream);

/* ------------------------------------------------------------------------- *
 * (b) relative entropy coding
 * ------------------------------------------------------------------------- */

/* Standard-normal candidate table, dimension-major: table[(d)*n + k] =
 * f32(ndtri(f64(sobol_k,d))) clamped to +-100, for n candidates (n <= 2^30).
 * Replaces get_sobol_normal_sample (test_model.py:493-498): SobolEngine.draw
 * by random access from the engine's scrambled state (shift (D), words (D,30),
 * both int64 as torch stores them) + Cephes ndtri in f64. */
int rcb_rec_table(const int64_t* shift, const int64_t* words, float* table, int D, int n,
                  rcb_stream_t stream);

/* Batched A* / Gumbel-max coding of n_pairs (row, block) pairs
 * (sample_group + compress_group, test_model.py:501-533,586-595):
 *   log_w_k = sum_d [log q(z_kd) - log p(z_kd)] + g_k,  z_kd = mu_p + sig_p * s_kd  (f64)
 *   idx = first argmax_k.  If apply!=0: sample[row, start:end] = f32(z_idx),
 *   mask[...] = 1, beta[row, block] = 0, coded[row, block] = 1, idx_out[row*G+block] = idx.
 * tables: per-block pointer table (G entries) into dimension-major tables.
 * q_scale/p_scale are standard deviations (already softplus/6). */
typedef struct {
  const int* pair_row; const int* pair_block;   /* (n_pairs) */
  const float* q_loc; const float* q_scale;     /* (rows, P) group order */
  const float* p_loc; const float* p_scale;     /* (P) */
  const int* group_start; const int* group_end; /* (G) */
  const float* const* tables;                   /* (G) device array of device pointers */
  const double* gumbel;                         /* (n_cand) */
  int* idx_out;                                 /* apply: (rows,G) else (n_pairs) */
  float* z_out;                                 /* apply=0: (n_pairs, max_D) or NULL */
  double* logw_out;                             /* (n_pairs, n_cand) or NULL */
  float* sample; float* mask; float* beta; uint8_t* coded;   /* apply targets */
  int n_pairs, P, G, n_cand, max_D, apply;
  /* Scratch of the staged scoring kernel (candidate table streamed through shared memory by bulk async copies,
   * candidates split over several CTAs per run): >= n_pairs * 772 + 128 bytes.  Its first
   * (workspace_bytes - 64) / 772 * 4 bytes are per-pair arrival counters: ZERO before the first call; every launch
   * leaves them zero again, so one buffer (always passed with the same workspace_bytes) serves every later call on
   * the same stream.  NULL, a buffer that is too small, or n_cand % 4 != 0 select the unstaged kernel; both
   * produce bit-identical results. */
  void* workspace; int64_t workspace_bytes;
} rcb_rec_args;
int rcb_rec_encode(const rcb_rec_args* a, rcb_stream_t stream);

/* The (row, block) pairs of one round -- row i codes blocks[i] -- listed with equal blocks adjacent (rows_out,
 * blocks_out: n entries each), the order rcb_rec_encode wants for its table sharing.  Replaces the host-side sort of
 * round 1; the serial row loop it stands in for is test_model.py:807-818. */
int rcb_rec_order(const int* blocks, int* rows_out, int* blocks_out, int n, int G, rcb_stream_t stream);

/* Measurement aid: `ctas` x 256 threads x 16 independent chains of `iters` double-precision FMAs; *flop_out (host)
 * receives the FLOP count of the launch.  bench.py times it with CUDA events to get the FP64 peak the REC scoring
 * kernel's roofline is stated against. */
int rcb_ubench_dfma(double* scratch, int ctas, int iters, double* flop_out, rcb_stream_t stream);

/* Receiver side: regenerate z for (row, block, idx) -> sample[row, start:end]. */
int rcb_rec_decode(const int* pair_row, const int* pair_block, const int* idx,
                   const float* p_loc, const float* p_scale, const int* group_start,
                   const int* group_end, const float* const* tables, float* sample,
                   float* mask, int n_pairs, int P, int n_cand, rcb_stream_t stream);

/* ------------------------------------------------------------------------- *
 * (c) prior EM statistics (main_prior_training.py:157-172)
 * ------------------------------------------------------------------------- */
/* stats[0:P] = sum_n mu, [P:2P] = sum_n mu^2, [2P:3P] = sum_n sigma^2 (f64). */
int rcb_prior_suffstats(const float* loc, const float* log_scale, double* stats, int rows, int P,
                        rcb_stream_t stream);
/* mu_p = s0/N; sigma_p = sqrt(s2/N + (s1 - N mu_p^2)/(N-1)) from (all-reduced) stats. */
int rcb_prior_from_stats(const double* stats, float* p_loc, float* p_scale, int64_t n_total, int P,
                         rcb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* RECOMBINER_B200_H */
