// HBM-bound kernels of the fit path: reparameterised sampling, sample-reduction +
// KL gradient + Adam, per-block KL sums, beta annealing, block selection, and the
// EM sufficient statistics of prior training.
#include <cuda_fp16.h>
#include "common.cuh"

namespace rcb {

// element (row n, sample s, latent l) of the lpe / d_lpe tensors: plain (rows*S, n_l), or the
// stitched per-datum grid of the patch modalities
__device__ __forceinline__ int64_t lpe_index(const int* slot, int rows_per_datum, int sp_total, int C, int n_l, int S,
                                             int n, int s, int l) {
  if (!slot) return ((int64_t)n * S + s) * n_l + l;
  const int d = n / rows_per_datum, r = n - d * rows_per_datum;
  const int sp = l / C, c = l - sp * C;
  return (((int64_t)d * S + s) * sp_total + slot[(int64_t)r * (n_l / C) + sp]) * C + c;
}

// ----------------------------------------------------------------- sampling --
// General form (patch modalities: per-column row permutation, level expansion, stitched latent grid, accumulation of
// levels 2 / 3 into hw).  grid: (chunks of 1024 weight parameters + chunks of 1024 latent values, rows).  A thread owns
// the four parameters {c0 + t, c0 + 256 + t, c0 + 512 + t, c0 + 768 + t} of its chunk -- the four normals of ONE
// Philox block per sample (philox_normal4; a thread per parameter generated every block four times) -- and keeps their
// four posterior gathers in flight together.
__global__ void __launch_bounds__(256) sample_kernel(rcb_sample_args a) {
  if (a.dyn) { a.seed = a.dyn->seed; a.step = a.dyn->step; }
  const int n = blockIdx.y, t = threadIdx.x;
  const int w_chunks = (a.n_w + 1023) >> 10;
  const bool is_w = (int)blockIdx.x < w_chunks;
  const int c0 = (is_w ? blockIdx.x : blockIdx.x - w_chunks) << 10;
  const int lim = is_w ? a.n_w : a.n_l;
  if (!is_w && !(a.lpe || a.lpe_h)) return;
  const int r0 = a.row_map ? a.row_map[n] : n;
  float mu[4], sig[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = c0 + k * 256 + t;               // index inside the weight / latent part
    mu[k] = 0.f; sig[k] = 0.f;
    if (i < lim) {
      const int p = is_w ? i : a.n_w + i;
      const int q = a.g2p ? a.g2p[p] : p;
      const int r = a.perm ? a.perm[(int64_t)r0 * a.P + q] : r0;
      const int64_t e = (int64_t)r * a.P + q;
      const float m = a.mask ? a.mask[e] : 0.f;
      mu[k] = a.loc[e] * (1.f - m);
      if (a.sample) mu[k] += a.sample[e] * m;
      sig[k] = std_transform(a.log_scale[e]) * (1.f - m) + 1e-15f * m;
    }
  }
  const int64_t gn = a.row_offset + n;
  for (int s = 0; s < a.S; ++s) {
    const int64_t item = (int64_t)n * a.S + s;
    const float* ein = is_w ? (a.eps_w ? a.eps_w + item * a.n_w : nullptr)
                            : (a.eps_l ? a.eps_l + ((int64_t)s * a.rows + n) * a.n_l : nullptr);
    float* eout = is_w ? (a.eps_w_store ? a.eps_w_store + item * a.n_w : nullptr)
                       : (a.eps_l_store ? a.eps_l_store + ((int64_t)s * a.rows + n) * a.n_l : nullptr);
    float z[4];
    if (!ein) philox_normal4(a.seed, a.step, is_w ? a.tensor_id : a.tensor_id + 16, gn * a.S + s, (uint32_t)((c0 >> 2) + t), z);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = c0 + k * 256 + t;
      if (i >= lim) continue;
      const float eps = ein ? ein[i] : z[k];
      if (eout) eout[i] = eps;
      const float v = fmaf(sig[k], eps, mu[k]);
      if (is_w) {
        if (a.hw_h) {
          reinterpret_cast<__half*>(a.hw_h)[item * a.ld_hw + i] = __float2half_rn(v);
        } else {
          float* dst = a.hw + item * a.ld_hw + i;
          *dst = a.accumulate ? *dst + v : v;
        }
      } else {
        const int64_t li = lpe_index(a.lpe_slot, a.rows_per_datum, a.sp_total, a.lpe_c, a.n_l, a.S, n, s, i);
        if (a.lpe_h) reinterpret_cast<__half*>(a.lpe_h)[li] = __float2half_rn(v);
        else a.lpe[li] = v;
      }
    }
  }
}


// Row-per-CTA variant for the non-patch modalities (no row permutation / expansion / stitching):
// the row's posterior (group order) is staged once in shared memory, then every (sample,
// parameter) output is written coalesced, four Philox normals per block.
// dynamic smem: 4 * P floats.
__global__ void __launch_bounds__(256, 4) sample_rows_kernel(rcb_sample_args a) {
  if (a.dyn) { a.seed = a.dyn->seed; a.step = a.dyn->step; }
  extern __shared__ float sm[];
  float* s_mu = sm;                 // parameter order
  float* s_sig = sm + a.P;
  const int n = blockIdx.x;
  if (a.g2p && !a.p2g) {            // no inverse map given: gather the row through g2p
    for (int p = threadIdx.x; p < a.P; p += blockDim.x) {
      const int64_t e = (int64_t)n * a.P + a.g2p[p];
      const float m = a.mask ? a.mask[e] : 0.f;
      float mu = a.loc[e] * (1.f - m);
      if (a.sample) mu += a.sample[e] * m;
      s_mu[p] = mu;
      s_sig[p] = (a.fast_math ? std_transform_fast(a.log_scale[e]) : std_transform(a.log_scale[e])) * (1.f - m) + 1e-15f * m;
    }
  } else {                          // coalesced read in stored (group) order, scattered into parameter order
    for (int q = threadIdx.x; q < a.P; q += blockDim.x) {
      const int64_t e = (int64_t)n * a.P + q;
      const float m = a.mask ? a.mask[e] : 0.f;
      float mu = a.loc[e] * (1.f - m);
      if (a.sample) mu += a.sample[e] * m;
      const int p = a.p2g ? a.p2g[q] : q;
      s_mu[p] = mu;
      s_sig[p] = (a.fast_math ? std_transform_fast(a.log_scale[e]) : std_transform(a.log_scale[e])) * (1.f - m) + 1e-15f * m;
    }
  }
  __syncthreads();
  const int64_t gn = a.row_offset + n;
  const int t = threadIdx.x;                     // blockDim.x == 256: slot of the Philox chunk
  for (int s = 0; s < a.S; ++s) {
    const int64_t item = (int64_t)n * a.S + s;
    float* hw = a.hw_h ? nullptr : a.hw + item * a.ld_hw;
    __half* hw_h = a.hw_h ? reinterpret_cast<__half*>(a.hw_h) + item * a.ld_hw : nullptr;
    const float* ein = a.eps_w ? a.eps_w + item * a.n_w : nullptr;
    float* eout = a.eps_w_store ? a.eps_w_store + item * a.n_w : nullptr;
    for (int c0 = 0; c0 < a.n_w; c0 += 1024) {
      float z[4];
      if (!ein) philox_normal4(a.seed, a.step, a.tensor_id, gn * a.S + s, (uint32_t)((c0 >> 2) + t), z);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int p = c0 + k * 256 + t;
        if (p < a.n_w) {
          const float eps = ein ? ein[p] : z[k];
          if (eout) eout[p] = eps;
          const float val = fmaf(s_sig[p], eps, s_mu[p]);
          if (hw_h) hw_h[p] = __float2half_rn(val);
          else hw[p] = val;
        }
      }
    }
    if (a.lpe || a.lpe_h) {
      float* lpe = a.lpe_h ? nullptr : a.lpe + item * a.n_l;
      __half* lpe_h = a.lpe_h ? reinterpret_cast<__half*>(a.lpe_h) + item * a.n_l : nullptr;
      const float* lin = a.eps_l ? a.eps_l + ((int64_t)s * a.rows + n) * a.n_l : nullptr;
      float* lout = a.eps_l_store ? a.eps_l_store + ((int64_t)s * a.rows + n) * a.n_l : nullptr;
      for (int c0 = 0; c0 < a.n_l; c0 += 1024) {
        float z[4];
        if (!lin) philox_normal4(a.seed, a.step, a.tensor_id + 16, gn * a.S + s, (uint32_t)((c0 >> 2) + t), z);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int l = c0 + k * 256 + t;
          if (l < a.n_l) {
            const float eps = lin ? lin[l] : z[k];
            if (lout) lout[l] = eps;
            const float val = fmaf(s_sig[a.n_w + l], eps, s_mu[a.n_w + l]);
            if (lpe_h) lpe_h[l] = __float2half_rn(val);
            else lpe[l] = val;
          }
        }
      }
    }
  }
}

// Sample reduction of update_rows_kernel with the sample count as a compile-time constant (5: the fit loop, 1: prior
// training) and the noise kept by the sampling kernel: straight-line loads through per-sample base pointers, two
// parameters in flight per thread; sums over s in ascending order.
template <int SS>
__device__ __forceinline__ void reduce_row_samples(const rcb_update_args& a, int r, float* s_dmu, float* s_dsig) {
    // the fit loop's case (five samples, noise kept by the sampling kernel): straight-line loads through
    // per-sample base pointers, two parameters (twenty loads) in flight per thread; same summation order
    const float* dh = a.d_hw + (int64_t)r * SS * a.ld_hw;
    const float* ew = a.eps_w + (int64_t)r * SS * a.n_w;
    for (int p0 = threadIdx.x; p0 < a.n_w; p0 += 2 * blockDim.x) {
      const int p1 = p0 + blockDim.x;
      const bool two = p1 < a.n_w;
      float d0[SS], e0[SS], d1[SS], e1[SS];
#pragma unroll
      for (int k = 0; k < SS; ++k) {
        d0[k] = dh[(int64_t)k * a.ld_hw + p0];
        e0[k] = ew[(int64_t)k * a.n_w + p0];
        d1[k] = two ? dh[(int64_t)k * a.ld_hw + p1] : 0.f;
        e1[k] = two ? ew[(int64_t)k * a.n_w + p1] : 0.f;
      }
      float m0 = 0.f, s0 = 0.f, m1 = 0.f, s1 = 0.f;
#pragma unroll
      for (int k = 0; k < SS; ++k) {
        m0 += d0[k]; s0 = fmaf(d0[k], e0[k], s0);
        m1 += d1[k]; s1 = fmaf(d1[k], e1[k], s1);
      }
      s_dmu[p0] = m0; s_dsig[p0] = s0;
      if (two) { s_dmu[p1] = m1; s_dsig[p1] = s1; }
    }
    if (a.n_l > 0) {
      const float* dl = a.d_lpe + (int64_t)r * SS * a.n_l;
      const float* el = a.eps_l + (int64_t)r * a.n_l;
      const int64_t es = (int64_t)a.rows * a.n_l;            // eps_l is (S, rows, n_l)
      for (int l0 = threadIdx.x; l0 < a.n_l; l0 += 2 * blockDim.x) {
        const int l1 = l0 + blockDim.x;
        const bool two = l1 < a.n_l;
        float d0[SS], e0[SS], d1[SS], e1[SS];
#pragma unroll
        for (int k = 0; k < SS; ++k) {
          d0[k] = dl[(int64_t)k * a.n_l + l0];
          e0[k] = el[k * es + l0];
          d1[k] = two ? dl[(int64_t)k * a.n_l + l1] : 0.f;
          e1[k] = two ? el[k * es + l1] : 0.f;
        }
        float m0 = 0.f, s0 = 0.f, m1 = 0.f, s1 = 0.f;
#pragma unroll
        for (int k = 0; k < SS; ++k) {
          m0 += d0[k]; s0 = fmaf(d0[k], e0[k], s0);
          m1 += d1[k]; s1 = fmaf(d1[k], e1[k], s1);
        }
        s_dmu[a.n_w + l0] = m0; s_dsig[a.n_w + l0] = s0;
        if (two) { s_dmu[a.n_w + l1] = m1; s_dsig[a.n_w + l1] = s1; }
      }
    }
}

// Row-per-CTA variant of update_kernel for the same case: the per-sample gradients are reduced
// in parameter order (coalesced) into shared memory, then the KL gradient + Adam runs in group
// order (coalesced on the stored state).  Same arithmetic order as update_kernel.
// dynamic smem: 2 * P floats.
template <bool FAST>
__global__ void __launch_bounds__(256, 4) update_rows_kernel(rcb_update_args a) {
  if (a.dyn) {
    a.seed = a.dyn->seed; a.step = a.dyn->step; a.adam_step_size = a.dyn->adam_step_size; a.adam_bc2_sqrt = a.dyn->adam_bc2_sqrt;
    if (a.dyn->beta_scalar >= 0.f) a.beta_scalar = a.dyn->beta_scalar;      // prior training: the global beta of a captured step
  }
  extern __shared__ float sm[];
  float* s_dmu = sm;
  float* s_dsig = sm + a.P;
  const int r = blockIdx.x;
  const int64_t gn = a.row_offset + r;
  const bool stored_noise = a.eps_w && (a.n_l == 0 || a.eps_l);
  if (stored_noise && a.S == 5) {
    reduce_row_samples<5>(a, r, s_dmu, s_dsig);
  } else if (stored_noise && a.S == 1) {
    reduce_row_samples<1>(a, r, s_dmu, s_dsig);
  } else
  // loads of 2 parameters x up to 4 samples are issued together (memory-level parallelism);
  // the sums still run over s in ascending order
  for (int p0 = threadIdx.x; p0 < a.P; p0 += 2 * blockDim.x) {
    float d_mu[2] = {0.f, 0.f}, d_sig[2] = {0.f, 0.f};
    for (int s0 = 0; s0 < a.S; s0 += 4) {
      float d[2][4], eps[2][4];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int p = p0 + u * blockDim.x;
        const bool is_w = p < a.n_w;
        const int l = p - a.n_w;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int s = s0 + k;
          d[u][k] = 0.f; eps[u][k] = 0.f;
          if (p < a.P && s < a.S) {
            const int64_t item = (int64_t)r * a.S + s;
            if (is_w) {
              d[u][k] = a.d_hw[item * a.ld_hw + p];
              eps[u][k] = a.eps_w ? a.eps_w[item * a.n_w + p]
                                  : philox_normal(a.seed, a.step, a.tensor_id, gn * a.S + s, (uint32_t)p);
            } else {
              d[u][k] = a.d_lpe[item * a.n_l + l];
              eps[u][k] = a.eps_l ? a.eps_l[((int64_t)s * a.rows + r) * a.n_l + l]
                                  : philox_normal(a.seed, a.step, a.tensor_id + 16, gn * a.S + s, (uint32_t)l);
            }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (s0 + k < a.S) {
            d_mu[u] += d[u][k];
            d_sig[u] = fmaf(d[u][k], eps[u][k], d_sig[u]);
          }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int p = p0 + u * blockDim.x;
      if (p < a.P) { s_dmu[p] = d_mu[u]; s_dsig[p] = d_sig[u]; }
    }
  }
  __syncthreads();
  // Two stored parameters per trip: all of their loads are issued before any arithmetic or store (the stores of one
  // element would otherwise fence the loads of the next: the kernel sat at 8.6 long-scoreboard stalls per issue with
  // ~10 dependent loads in flight per thread).  Per element the arithmetic and its order are unchanged.
  float kl_term = 0.f;
  for (int q0 = threadIdx.x; q0 < a.P; q0 += 2 * blockDim.x) {
    float mu_[2], rho_[2], m_[2], beta_[2], mup_[2], rawp_[2], m1a_[2], va_[2], m1b_[2], vb_[2], sdm_[2], sds_[2];
    bool on_[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int q = q0 + u * blockDim.x;
      on_[u] = q < a.P;
      const int qq = on_[u] ? q : q0;
      const int64_t e = (int64_t)r * a.P + qq;
      const int p = a.p2g ? a.p2g[qq] : qq;
      mu_[u] = a.loc[e]; rho_[u] = a.log_scale[e];
      m_[u] = a.mask ? a.mask[e] : 0.f;
      beta_[u] = a.beta ? a.beta[(int64_t)r * a.G + a.group_idx[qq]] : a.beta_scalar;
      mup_[u] = a.p_loc[qq]; rawp_[u] = a.p_log_scale[qq];
      if (a.adam) { m1a_[u] = a.m1_loc[e]; va_[u] = a.v_loc[e]; m1b_[u] = a.m1_ls[e]; vb_[u] = a.v_ls[e]; }
      sdm_[u] = s_dmu[p]; sds_[u] = s_dsig[p];
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (!on_[u]) continue;
      const int q = q0 + u * blockDim.x;
      const int64_t e = (int64_t)r * a.P + q;
      const float mu = mu_[u], rho = rho_[u], m = m_[u];
      const float sig = FAST ? std_transform_fast(rho) : std_transform(rho);
      float d_mu = 0.f, d_sig = 0.f;
      if (m != 1.f) {
        d_mu = sdm_[u] * (a.grad_scale * (1.f - m));
        d_sig = sds_[u] * (a.grad_scale * (1.f - m));
      }
      const float beta = beta_[u];
      const float mu_p = mup_[u];
      const float sig_p = a.p_scale_direct ? rawp_[u] : (FAST ? std_transform_fast(rawp_[u]) : std_transform(rawp_[u]));
      const float dm = mu - mu_p;
      float inv_vp, vr, t1, inv_sig, lvr;
      if (FAST) {      // the same expressions through one reciprocal of sigma_p and one of sigma
        const float rp = __fdividef(1.f, sig_p);
        inv_vp = rp * rp;
        const float ratio = sig * rp;
        vr = ratio * ratio;
        t1 = (dm * rp) * (dm * rp);
        inv_sig = __fdividef(1.f, sig);
        lvr = __logf(vr);
      } else {
        inv_vp = 1.f / (sig_p * sig_p);
        const float ratio = sig / sig_p;
        vr = ratio * ratio;
        t1 = (dm / sig_p) * (dm / sig_p);
        inv_sig = 1.f / sig;
        lvr = logf(vr);
      }
      kl_term += beta * 0.5f * (vr + t1 - 1.f - lvr);
      const float g_mu = d_mu + beta * dm * inv_vp;
      const float g_sig = d_sig + beta * (sig * inv_vp - inv_sig);
      const float g_rho = g_sig * (FAST ? std_transform_grad_fast(rho) : std_transform_grad(rho));
      if (a.adam) {
        const float step_size = a.adam_step_size, bc2s = a.adam_bc2_sqrt;
        const float inv_bc2s = FAST ? __fdividef(1.f, bc2s) : 0.f;
        float m1 = m1a_[u], v = va_[u];
        m1 = m1 + (g_mu - m1) * (1.f - a.b1);
        v = v * a.b2 + (1.f - a.b2) * g_mu * g_mu;
        a.m1_loc[e] = m1; a.v_loc[e] = v;
        a.loc[e] = FAST ? mu - step_size * __fdividef(m1, fmaf(sqrt_fast(v), inv_bc2s, a.adam_eps))
                        : mu - step_size * (m1 / (sqrtf(v) / bc2s + a.adam_eps));
        m1 = m1b_[u]; v = vb_[u];
        m1 = m1 + (g_rho - m1) * (1.f - a.b1);
        v = v * a.b2 + (1.f - a.b2) * g_rho * g_rho;
        a.m1_ls[e] = m1; a.v_ls[e] = v;
        a.log_scale[e] = FAST ? rho - step_size * __fdividef(m1, fmaf(sqrt_fast(v), inv_bc2s, a.adam_eps))
                              : rho - step_size * (m1 / (sqrtf(v) / bc2s + a.adam_eps));
      } else {
        a.g_loc[e] = g_mu;
        a.g_log_scale[e] = g_rho;
      }
    }
  }
  if (a.kl_out) {
    __shared__ float red[8];
    const float s = warp_sum(kl_term);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < 8; ++w) t += (double)red[w];
      atomicAdd(a.kl_out, t);
    }
  }
}

// ------------------------------------------------ per-row sample reduction --
// First half of the gradient reduction for the general (patch) layout: per patch row n and parameter p (parameter
// order, coalesced), red_mu[n][p] = sum_s d[n,s,p] and, for every level l of the hierarchy, red_sig[l][n][p] =
// sum_s d[n,s,p] * eps_l[n,s,p] (the levels share the data gradient and differ in their noise, utils.py:142-191);
// the latent part likewise from the stitched d_lpe.  update_kernel then only sums these over the rows a stored
// parameter feeds (one for level 1, 4 .. 96 children for levels 2 / 3) instead of over (children x samples) scattered
// 4-byte loads.  grid: (ceil((n_w + n_l) / 256), rows).
__global__ void __launch_bounds__(256) reduce_samples_kernel(rcb_reduce_args a) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.y;
  if (p < a.n_w) {
    float m = 0.f, sg[3] = {0.f, 0.f, 0.f};
    for (int s = 0; s < a.S; ++s) {
      const int64_t item = (int64_t)n * a.S + s;
      const float d = a.d_hw[item * a.ld_hw + p];
      m += d;
#pragma unroll
      for (int l = 0; l < 3; ++l)
        if (l < a.n_levels) sg[l] = fmaf(d, a.eps_w[l][item * a.n_w + p], sg[l]);
    }
    a.red_mu[(int64_t)n * a.n_w + p] = m;
#pragma unroll
    for (int l = 0; l < 3; ++l)
      if (l < a.n_levels) a.red_sig[l][(int64_t)n * a.n_w + p] = sg[l];
  } else if (p < a.n_w + a.n_l) {
    const int l = p - a.n_w;
    float m = 0.f, sg = 0.f;
    for (int s = 0; s < a.S; ++s) {
      const float d = a.d_lpe[lpe_index(a.lpe_slot, a.rows_per_datum, a.sp_total, a.lpe_c, a.n_l, a.S, n, s, l)];
      m += d;
      sg = fmaf(d, a.eps_l[((int64_t)s * a.rows + n) * a.n_l + l], sg);
    }
    a.red_mu_l[(int64_t)n * a.n_l + l] = m;
    a.red_sig_l[(int64_t)n * a.n_l + l] = sg;
  }
}

// ------------------------------------------- gradient reduction + KL + Adam --
// grid: (ceil(P/256), src_rows).  One thread = one stored (row, group-order column).
__global__ void __launch_bounds__(256) update_kernel(rcb_update_args a) {
  if (a.dyn) {
    a.seed = a.dyn->seed; a.step = a.dyn->step; a.adam_step_size = a.dyn->adam_step_size; a.adam_bc2_sqrt = a.dyn->adam_bc2_sqrt;
    if (a.dyn->beta_scalar >= 0.f) a.beta_scalar = a.dyn->beta_scalar;      // prior training: the global beta of a captured step
  }
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  float kl_term = 0.f;
  if (q < a.P) {
    const int64_t e = (int64_t)r * a.P + q;
    const int p = a.p2g ? a.p2g[q] : q;
    const int rr = a.perm_inv ? a.perm_inv[e] : r;        // row of this level in parameter order
    const float mu = a.loc[e], rho = a.log_scale[e];
    const float m = a.mask ? a.mask[e] : 0.f;
    const float sig = std_transform(rho);
    // ---- reduce the per-sample gradients that this parameter produced
    float d_mu = 0.f, d_sig = 0.f;
    const bool is_w = p < a.n_w;
    const float* src = is_w ? a.d_hw : a.d_lpe;
    if (a.red_mu && m != 1.f) {
      // sample sums already reduced per patch row (rcb_fit_reduce): add the rows this parameter feeds
      const int nch = a.row_children ? a.n_children : 1;
      const float* rm = is_w ? a.red_mu : a.red_mu_l;
      const float* rs = is_w ? a.red_sig : a.red_sig_l;
      const int ld = is_w ? a.n_w : a.n_l, col = is_w ? p : p - a.n_w;
      for (int c0 = 0; c0 < nch; c0 += 8) {
        float dv[8], sv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int c = c0 + u;
          dv[u] = 0.f; sv[u] = 0.f;
          if (c < nch) {
            const int n = a.row_children ? a.row_children[(int64_t)rr * a.n_children + c] : rr;
            dv[u] = rm[(int64_t)n * ld + col];
            sv[u] = rs[(int64_t)n * ld + col];
          }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) { d_mu += dv[u]; d_sig += sv[u]; }
      }
      d_mu *= a.grad_scale * (1.f - m);
      d_sig *= a.grad_scale * (1.f - m);
    } else if (src && m != 1.f) {
      const int nch = a.row_children ? a.n_children : 1;
      // (child row, sample) pairs in the reference's order; eight gradient / noise loads are issued together
      // (a level-3 row of a patch modality sums over every patch of the datum: hundreds of terms per thread)
      const int total = nch * a.S;
      for (int t0 = 0; t0 < total; t0 += 8) {
        float dv[8], ev[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int t = t0 + u;
          dv[u] = 0.f; ev[u] = 0.f;
          if (t < total) {
            const int c = t / a.S, s = t - c * a.S;
            const int n = a.row_children ? a.row_children[(int64_t)rr * a.n_children + c] : rr;
            const int64_t gn = a.row_offset + n;
            if (is_w) {
              dv[u] = src[((int64_t)n * a.S + s) * a.ld_hw + p];
              ev[u] = a.eps_w ? a.eps_w[((int64_t)n * a.S + s) * a.n_w + p]
                              : philox_normal(a.seed, a.step, a.tensor_id, gn * a.S + s, (uint32_t)p);
            } else {
              const int l = p - a.n_w;
              dv[u] = src[lpe_index(a.lpe_slot, a.rows_per_datum, a.sp_total, a.lpe_c, a.n_l, a.S, n, s, l)];
              ev[u] = a.eps_l ? a.eps_l[((int64_t)s * a.rows + n) * a.n_l + l]
                              : philox_normal(a.seed, a.step, a.tensor_id + 16, gn * a.S + s, (uint32_t)l);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (t0 + u < total) {
            d_mu += dv[u];
            d_sig = fmaf(dv[u], ev[u], d_sig);
          }
        }
      }
      d_mu *= a.grad_scale * (1.f - m);
      d_sig *= a.grad_scale * (1.f - m);
    }
    // ---- closed-form KL(q||p) and its gradient, weighted by the block's beta
    const float beta = a.beta ? a.beta[(int64_t)r * a.G + a.group_idx[q]] : a.beta_scalar;
    const float mu_p = a.p_loc[q], sig_p = a.p_scale_direct ? a.p_log_scale[q] : std_transform(a.p_log_scale[q]);
    const float inv_vp = 1.f / (sig_p * sig_p);
    const float dm = mu - mu_p;
    const float ratio = sig / sig_p;
    const float vr = ratio * ratio;
    const float t1 = (dm / sig_p) * (dm / sig_p);
    kl_term = beta * 0.5f * (vr + t1 - 1.f - logf(vr));
    float g_mu = d_mu + beta * dm * inv_vp;
    float g_sig = d_sig + beta * (sig * inv_vp - 1.f / sig);
    float g_rho = g_sig * std_transform_grad(rho);
    if (a.adam) {
      const float step_size = a.adam_step_size, bc2s = a.adam_bc2_sqrt;
      float m1 = a.m1_loc[e], v = a.v_loc[e];
      m1 = m1 + (g_mu - m1) * (1.f - a.b1);
      v = v * a.b2 + (1.f - a.b2) * g_mu * g_mu;
      a.m1_loc[e] = m1; a.v_loc[e] = v;
      a.loc[e] = mu - step_size * (m1 / (sqrtf(v) / bc2s + a.adam_eps));
      m1 = a.m1_ls[e]; v = a.v_ls[e];
      m1 = m1 + (g_rho - m1) * (1.f - a.b1);
      v = v * a.b2 + (1.f - a.b2) * g_rho * g_rho;
      a.m1_ls[e] = m1; a.v_ls[e] = v;
      a.log_scale[e] = rho - step_size * (m1 / (sqrtf(v) / bc2s + a.adam_eps));
    } else {
      a.g_loc[e] = g_mu;
      a.g_log_scale[e] = g_rho;
    }
  }
  if (a.kl_out) {
    __shared__ float red[8];
    float s = warp_sum(kl_term);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < 8; ++w) t += (double)red[w];
      atomicAdd(a.kl_out, t);
    }
  }
}

// ----------------------------------------------------------- per-block KL ----
// grid: rows; one warp per block, lanes stride the block's contiguous columns.
__global__ void __launch_bounds__(256) group_kl_kernel(const float* __restrict__ loc, const float* __restrict__ log_scale,
                                                       const float* __restrict__ p_loc, const float* __restrict__ p_log_scale,
                                                       const int* __restrict__ gs, const int* __restrict__ ge,
                                                       double* __restrict__ kl, int P, int G) {
  const int r = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int g = warp; g < G; g += 8) {
    double acc = 0.0;
    for (int q = gs[g] + lane; q < ge[g]; q += 32) {
      const int64_t e = (int64_t)r * P + q;
      float sig = std_transform(log_scale[e]), sig_p = std_transform(p_log_scale[q]);
      float ratio = sig / sig_p, vr = ratio * ratio;
      float d = (loc[e] - p_loc[q]) / sig_p;
      acc += (double)(0.5f * (vr + d * d - 1.f - logf(vr)));
    }
    acc = warp_sum(acc);
    if (lane == 0) kl[(int64_t)r * G + g] = acc;
  }
}

__global__ void anneal_kernel(float* __restrict__ beta, const double* __restrict__ kl, const uint8_t* __restrict__ coded,
                              int64_t total, double step, double upper, double lower, double bits) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= total) return;
  if (coded && coded[i]) return;
  const double b = kl[i] / 0.6931471805599453;
  float v = beta[i];
  const float factor = (float)(1.0 + step);   // f64 1+step, then .float() (test_model.py:407,409)
  if (b > bits + upper) v = v * factor;
  if (b <= bits - lower) v = v / factor;
  beta[i] = fminf(fmaxf(v, 0.f), 10000.f);
}

// one warp per row: first index of the largest KL among not-yet-coded blocks
__global__ void pick_block_kernel(const double* __restrict__ kl, const uint8_t* __restrict__ coded, int* __restrict__ block,
                                  int rows, int G) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= rows) return;
  double best = -1e300;
  int bi = 0x7fffffff;
  for (int g = lane; g < G; g += 32) {
    double v = coded[(int64_t)r * G + g] ? -1e10 : kl[(int64_t)r * G + g] / 0.6931471805599453;
    if (v > best) { best = v; bi = g; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double ov = __shfl_xor_sync(0xffffffffu, best, o);
    int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  // a row whose KLs are all NaN (diverged fit) never wins a comparison: code block 0 instead of an out-of-range index
  if (lane == 0) block[r] = bi < G ? bi : 0;
}

__global__ void __launch_bounds__(256) to_half_kernel(const float* __restrict__ src, __half* __restrict__ dst, int64_t n) {
  const int64_t i0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
#pragma unroll
  for (int e = 0; e < 4; ++e)
    if (i0 + e < n) dst[i0 + e] = __float2half_rn(src[i0 + e]);
}

// out = softplus(in) / 6 with the same device arithmetic the fit kernels use for the posterior / prior scales, so that the
// encoder, the decoder and the fit agree on every bit of a standard deviation (test_model.py:101)
__global__ void __launch_bounds__(256) std_transform_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = std_transform(in[i]);
}

__global__ void set_step_state_kernel(rcb_step_state* dev, long long seed, int step, float ss, float bc, float beta) {
  dev->seed = seed; dev->step = step; dev->adam_step_size = ss; dev->adam_bc2_sqrt = bc; dev->beta_scalar = beta;
}

// Adam on one flat fp32 parameter vector (the shared mappings of prior training): the update torch.optim.Adam makes
// with default flags, theta -= step_size * m / (sqrt(v) / sqrt(1 - b2^t) + eps), step_size = lr / (1 - b1^t).
__global__ void __launch_bounds__(256) adam_flat_kernel(float* __restrict__ theta, const float* __restrict__ grad,
                                                        float* __restrict__ m, float* __restrict__ v, int64_t n, float step_size,
                                                        float bc2s, float b1, float b2, float eps, const rcb_step_state* dyn) {
  if (dyn) { step_size = dyn->adam_step_size; bc2s = dyn->adam_bc2_sqrt; }
  const int64_t i0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
  if (i0 >= n) return;
  if (i0 + 4 <= n) {
    const float4 g = *reinterpret_cast<const float4*>(grad + i0);
    float4 mm = *reinterpret_cast<const float4*>(m + i0), vv = *reinterpret_cast<const float4*>(v + i0);
    float4 t = *reinterpret_cast<const float4*>(theta + i0);
#define RCB_ADAM1(c)                                                     \
    mm.c = b1 * mm.c + (1.f - b1) * g.c;                                 \
    vv.c = b2 * vv.c + (1.f - b2) * g.c * g.c;                           \
    t.c = t.c - step_size * (mm.c / (sqrtf(vv.c) / bc2s + eps));
    RCB_ADAM1(x) RCB_ADAM1(y) RCB_ADAM1(z) RCB_ADAM1(w)
#undef RCB_ADAM1
    *reinterpret_cast<float4*>(m + i0) = mm;
    *reinterpret_cast<float4*>(v + i0) = vv;
    *reinterpret_cast<float4*>(theta + i0) = t;
  } else {
    for (int64_t i = i0; i < n; ++i) {
      const float g = grad[i];
      const float mm = b1 * m[i] + (1.f - b1) * g, vv = b2 * v[i] + (1.f - b2) * g * g;
      m[i] = mm; v[i] = vv;
      theta[i] = theta[i] - step_size * (mm / (sqrtf(vv) / bc2s + eps));
    }
  }
}

// out[0] = scale * sum(sqerr[0..n)), out[1] = kl[0]  (per-step loss terms of prior training, f64; one CTA)
__global__ void __launch_bounds__(256) step_stats_kernel(const float* __restrict__ sqerr, int n, double scale,
                                                         const double* __restrict__ kl, double* __restrict__ out) {
  __shared__ double red[8];
  double t = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) t += (double)sqerr[i];
  t = warp_sum(t);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += red[w];
    out[0] = scale * s;
    out[1] = kl ? kl[0] : 0.0;
  }
}

// ------------------------------------------------------------------ transpose --
// out[c][r] = in[r][c] through a padded 32 x 32 shared-memory tile (coalesced on both sides).
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ in, int64_t ld_in, float* __restrict__ out,
                                                        int64_t ld_out, int rows, int cols) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    tile[i][tx] = (r < rows && c < cols) ? in[(int64_t)r * ld_in + c] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (c < cols && r < rows) out[(int64_t)c * ld_out + r] = tile[tx][i];
  }
}

// Three channel-major copies of a channel-last tensor whose rows are (item, y, x) pixels, shifted by dx = -1, 0, +1
// along x with zeros outside the line: out[dx + 1][c][r] = (0 <= r % w + dx < w) ? in[r + dx][c] : 0.
// (TMA needs 16-byte aligned box starts, so a one-pixel shift along a tensor's innermost dimension cannot be a box
// coordinate; the tensor-core weight gradient picks the copy instead.)
__global__ void __launch_bounds__(256) transpose_xshift_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t rows,
                                                               int cols, int w) {
  __shared__ float tile[34][33];                    // rows r0 - 1 .. r0 + 32
  const int c0 = blockIdx.x * 32;
  const int64_t r0 = (int64_t)blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 34; i += 8) {
    const int64_t r = r0 - 1 + i;
    const int c = c0 + tx;
    tile[i][tx] = (r >= 0 && r < rows && c < cols) ? in[r * cols + c] : 0.f;
  }
  __syncthreads();
  const int64_t r = r0 + tx;
  if (r < rows) {
    const int x = (int)(r % w);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int dx = k - 1;
      const bool inside = x + dx >= 0 && x + dx < w;
#pragma unroll
      for (int i = ty; i < 32; i += 8) {
        const int c = c0 + i;
        if (c < cols) out[((int64_t)k * cols + c) * rows + r] = inside ? tile[tx + 1 + dx][i] : 0.f;
      }
    }
  }
}

// Same, with the rows (item, oy, ox) of an upsampled channel-last tensor regrouped by phase on the way out:
// out[c][((item * fy + oy % fy) * fx + ox % fx) * h * w + (oy / fy) * w + ox / fx] = in[(item, oy, ox)][c],
// i.e. every phase of the output grid becomes a dense (h, w) plane per channel (TMA cannot stride its innermost
// dimension; the tensor-core weight gradient reads one phase at a time).
__global__ void __launch_bounds__(256) transpose_phases_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t rows,
                                                               int cols, int h, int w, int fy, int fx) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32;
  const int64_t r0 = (int64_t)blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = ty; i < 32; i += 8) {
    const int64_t r = r0 + i;
    const int c = c0 + tx;
    tile[i][tx] = (r < rows && c < cols) ? in[r * cols + c] : 0.f;
  }
  __syncthreads();
  const int Wo = w * fx, Ho = h * fy;
  const int64_t r = r0 + tx;
  if (r < rows) {
    const int ox = (int)(r % Wo);
    const int64_t t = r / Wo;
    const int oy = (int)(t % Ho);
    const int64_t item = t / Ho;
    const int64_t ro = ((item * fy + oy % fy) * fx + ox % fx) * (int64_t)(h * w) + (int64_t)(oy / fy) * w + ox / fx;
#pragma unroll
    for (int i = ty; i < 32; i += 8) {
      const int c = c0 + i;
      if (c < cols) out[(int64_t)c * rows + ro] = tile[tx][i];
    }
  }
}

// Wide-tile forms of the two kernels above for cols <= 64 (the upsampler's channel counts): one CTA moves 128
// consecutive rows -- a contiguous 128 * cols block of the input, read as float4 -- through a padded shared-memory tile
// and writes 32 consecutive rows of one channel per warp instruction.  Same outputs; 4 x fewer, 4 x larger CTAs and no
// half-empty 32-column tiles for the 16-channel tensors.
// rows per CTA: a tile of ~32 KB whatever the channel count (128 rows of 64 channels, 512 rows of 16)
static inline int tw_rows(int cols) { return cols <= 16 ? 512 : (cols <= 32 ? 256 : 128); }

template <int MODE>        // 0: x-shifted copies (transpose_xshift), 1: phase planes (transpose_phases)
__global__ void __launch_bounds__(256) transpose_wide_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t rows,
                                                             int cols, int h, int w, int fy, int fx, int TW_ROWS) {
  extern __shared__ float tw_tile[];                  // [TW_ROWS + 2][cols + 1]; row 0 = the row before the block
  const int ldt = cols + 1;
  const int64_t r0 = (int64_t)blockIdx.x * TW_ROWS;
  const int nrow = (int)min((int64_t)TW_ROWS, rows - r0);
  // block rows r0 - 1 .. r0 + nrow (halo rows only matter for MODE 0)
  const int64_t first = MODE == 0 ? r0 - 1 : r0;
  const int nload = MODE == 0 ? nrow + 2 : nrow;
  const int64_t base = first * cols;
  const int total4 = nload * cols / 4;                // cols % 4 == 0
  for (int i = threadIdx.x; i < total4; i += 256) {
    const int64_t e = base + 4 * (int64_t)i;
    const int lr = (4 * i) / cols, lc = (4 * i) - lr * cols;
    const int64_t gr = first + lr;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gr >= 0 && gr < rows) v = *reinterpret_cast<const float4*>(in + e);
    float* t = tw_tile + (MODE == 0 ? lr : lr + 1) * ldt + lc;
    t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int rg = 0; rg < TW_ROWS / 32; ++rg) {
    const int lr = rg * 32 + lane;
    const int64_t r = r0 + lr;
    if (lr >= nrow) continue;
    if (MODE == 0) {
      const int x = (int)(r % w);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int dx = k - 1;
        const bool inside = x + dx >= 0 && x + dx < w;
        for (int c = warp; c < cols; c += 8)
          out[((int64_t)k * cols + c) * rows + r] = inside ? tw_tile[(lr + 1 + dx) * ldt + c] : 0.f;
      }
    } else {
      const int Wo = w * fx, Ho = h * fy;
      const int ox = (int)(r % Wo);
      const int64_t t = r / Wo;
      const int oy = (int)(t % Ho);
      const int64_t item = t / Ho;
      const int64_t ro = ((item * fy + oy % fy) * fx + ox % fx) * (int64_t)(h * w) + (int64_t)(oy / fy) * w + ox / fx;
      for (int c = warp; c < cols; c += 8) out[(int64_t)c * rows + ro] = tw_tile[(lr + 1) * ldt + c];
    }
  }
}

// ------------------------------------------------------ EM prior statistics --
// grid: (ceil(P/128), row chunks); f64 partial sums added with atomics.
__global__ void __launch_bounds__(128) suffstats_kernel(const float* __restrict__ loc, const float* __restrict__ log_scale,
                                                        double* __restrict__ stats, int rows, int P, int rows_per_block) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(rows, r0 + rows_per_block);
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  for (int r = r0; r < r1; ++r) {
    const int64_t e = (int64_t)r * P + p;
    float mu = loc[e], sg = std_transform(log_scale[e]);
    s0 += (double)mu;
    s1 += (double)mu * (double)mu;
    s2 += (double)(sg * sg);
  }
  atomicAdd(stats + p, s0);
  atomicAdd(stats + P + p, s1);
  atomicAdd(stats + 2 * (int64_t)P + p, s2);
}

__global__ void prior_from_stats_kernel(const double* __restrict__ stats, float* __restrict__ p_loc,
                                        float* __restrict__ p_scale, double n, int P) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const double mean = stats[p] / n;
  const double var_mu = (stats[P + p] - n * mean * mean) / (n - 1.0);   // unbiased, main_prior_training.py:158
  const double v = stats[2 * (int64_t)P + p] / n + var_mu;
  p_loc[p] = (float)mean;
  p_scale[p] = (float)sqrt(v > 0.0 ? v : 0.0);
}

}  // namespace rcb

using namespace rcb;

extern "C" int rcb_set_step_state(rcb_step_state* dev, int64_t seed, int step, float adam_step_size,
                                  float adam_bc2_sqrt, float beta_scalar, rcb_stream_t stream) {
  RCB_CHECK_ARG(dev != nullptr, "rcb_set_step_state: null pointer");
  set_step_state_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(dev, (long long)seed, step, adam_step_size, adam_bc2_sqrt,
                                                          beta_scalar);
  RCB_CHECK_LAUNCH("rcb_set_step_state");
  return 0;
}

extern "C" int rcb_std_transform(const float* in, float* out, int64_t n, rcb_stream_t stream) {
  RCB_CHECK_ARG(in && out && n > 0, "rcb_std_transform: bad arguments");
  const int64_t blocks = (n + 255) / 256;
  RCB_CHECK_ARG(blocks < (1ll << 31), "rcb_std_transform: tensor too large");
  std_transform_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(in, out, n);
  RCB_CHECK_LAUNCH("rcb_std_transform");
  return 0;
}

extern "C" int rcb_adam_flat(float* theta, const float* grad, float* m, float* v, int64_t n, float step_size,
                             float bc2_sqrt, float b1, float b2, float eps, const rcb_step_state* dyn, rcb_stream_t stream) {
  RCB_CHECK_ARG(theta && grad && m && v && n > 0, "rcb_adam_flat: bad arguments");
  RCB_CHECK_ARG((((uintptr_t)theta | (uintptr_t)grad | (uintptr_t)m | (uintptr_t)v) & 15) == 0, "rcb_adam_flat: buffers must be 16-byte aligned");
  const int64_t blocks = (n + 1023) / 1024;
  RCB_CHECK_ARG(blocks < (1ll << 31), "rcb_adam_flat: vector too long");
  adam_flat_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(theta, grad, m, v, n, step_size, bc2_sqrt, b1, b2, eps, dyn);
  RCB_CHECK_LAUNCH("rcb_adam_flat");
  return 0;
}

extern "C" int rcb_step_stats(const float* sqerr, int n, double scale, const double* kl, double* out, rcb_stream_t stream) {
  RCB_CHECK_ARG(sqerr && out && n > 0, "rcb_step_stats: bad arguments");
  step_stats_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(sqerr, n, scale, kl, out);
  RCB_CHECK_LAUNCH("rcb_step_stats");
  return 0;
}

extern "C" int rcb_fit_sample(const rcb_sample_args* a, rcb_stream_t stream) {
  RCB_CHECK_ARG(a != nullptr, "rcb_fit_sample: null args");
  RCB_CHECK_ARG(a->loc && a->log_scale && (a->hw || a->hw_h), "rcb_fit_sample: null tensor");
  RCB_CHECK_ARG(!(a->hw_h && a->accumulate), "rcb_fit_sample: the fp16 weight samples cannot be accumulated into");
  RCB_CHECK_ARG(a->rows > 0 && a->S > 0 && a->P > 0, "rcb_fit_sample: empty problem");
  RCB_CHECK_ARG(a->n_w + a->n_l == a->P || (a->n_l == 0 && a->n_w == a->P), "rcb_fit_sample: n_w+n_l != P");
  RCB_CHECK_ARG(a->ld_hw >= a->n_w, "rcb_fit_sample: ld_hw too small");
  RCB_CHECK_ARG(a->rows <= 65535, "rcb_fit_sample: at most 65535 rows per call");
  const size_t row_smem = 2 * sizeof(float) * (size_t)a->P;
  if (!a->perm && !a->row_map && !a->lpe_slot && !a->accumulate && row_smem <= 200 * 1024) {
    if (row_smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(sample_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)row_smem);
      if (e != cudaSuccess) { set_error("rcb_fit_sample: smem opt-in failed: %s", cudaGetErrorString(e)); return -1; }
    }
    sample_rows_kernel<<<a->rows, 256, row_smem, (cudaStream_t)stream>>>(*a);
    RCB_CHECK_LAUNCH("rcb_fit_sample");
    return 0;
  }
  dim3 grid(ceil_div(a->n_w, 1024) + ceil_div(a->n_l, 1024), a->rows);
  sample_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*a);
  RCB_CHECK_LAUNCH("rcb_fit_sample");
  return 0;
}

extern "C" int rcb_fit_reduce(const rcb_reduce_args* a, rcb_stream_t stream) {
  RCB_CHECK_ARG(a != nullptr, "rcb_fit_reduce: null args");
  RCB_CHECK_ARG(a->d_hw && a->red_mu && a->n_levels >= 1 && a->n_levels <= 3, "rcb_fit_reduce: bad arguments");
  for (int l = 0; l < a->n_levels; ++l)
    RCB_CHECK_ARG(a->eps_w[l] && a->red_sig[l], "rcb_fit_reduce: level %d needs its noise and its output", l);
  RCB_CHECK_ARG(a->n_l == 0 || (a->d_lpe && a->eps_l && a->red_mu_l && a->red_sig_l), "rcb_fit_reduce: latent part incomplete");
  RCB_CHECK_ARG(a->rows > 0 && a->rows <= 65535 && a->S > 0 && a->n_w > 0 && a->ld_hw >= a->n_w, "rcb_fit_reduce: bad shape");
  dim3 grid(ceil_div(a->n_w + a->n_l, 256), a->rows);
  reduce_samples_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*a);
  RCB_CHECK_LAUNCH("rcb_fit_reduce");
  return 0;
}

extern "C" int rcb_fit_update(const rcb_update_args* a, rcb_stream_t stream) {
  RCB_CHECK_ARG(a != nullptr, "rcb_fit_update: null args");
  RCB_CHECK_ARG(a->loc && a->log_scale && a->p_loc && a->p_log_scale, "rcb_fit_update: null tensor");
  RCB_CHECK_ARG(a->beta == nullptr || a->group_idx != nullptr, "rcb_fit_update: per-block beta needs group_idx");
  RCB_CHECK_ARG(a->src_rows > 0 && a->src_rows <= 65535 && a->P > 0, "rcb_fit_update: bad shape");
  if (a->adam) {
    RCB_CHECK_ARG(a->m1_loc && a->v_loc && a->m1_ls && a->v_ls && (a->dyn || a->adam_bc2_sqrt > 0.f), "rcb_fit_update: Adam state missing");
  } else {
    RCB_CHECK_ARG(a->g_loc && a->g_log_scale, "rcb_fit_update: gradient outputs missing");
  }
  const size_t row_smem = 2 * sizeof(float) * (size_t)a->P;
  RCB_CHECK_ARG(!a->red_mu || (a->red_sig && (a->n_l == 0 || (a->red_mu_l && a->red_sig_l))), "rcb_fit_update: reduced gradients incomplete");
  if (!a->perm_inv && !a->row_children && !a->lpe_slot && !a->red_mu && a->d_hw && (a->n_l == 0 || a->d_lpe) && a->n_w + a->n_l == a->P &&
      row_smem <= 200 * 1024) {
    if (row_smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(update_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)row_smem);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(update_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)row_smem);
      if (e != cudaSuccess) { set_error("rcb_fit_update: smem opt-in failed: %s", cudaGetErrorString(e)); return -1; }
    }
    if (a->fast_math) update_rows_kernel<true><<<a->src_rows, 256, row_smem, (cudaStream_t)stream>>>(*a);
    else update_rows_kernel<false><<<a->src_rows, 256, row_smem, (cudaStream_t)stream>>>(*a);
    RCB_CHECK_LAUNCH("rcb_fit_update");
    return 0;
  }
  dim3 grid(ceil_div(a->P, 256), a->src_rows);
  update_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*a);
  RCB_CHECK_LAUNCH("rcb_fit_update");
  return 0;
}

extern "C" int rcb_group_kl(const float* loc, const float* log_scale, const float* p_loc, const float* p_log_scale,
                            const int* group_start, const int* group_end, double* kl, int rows, int P, int G,
                            rcb_stream_t stream) {
  RCB_CHECK_ARG(loc && log_scale && p_loc && p_log_scale && group_start && group_end && kl, "rcb_group_kl: null tensor");
  if (rows <= 0 || G <= 0) return 0;
  group_kl_kernel<<<rows, 256, 0, (cudaStream_t)stream>>>(loc, log_scale, p_loc, p_log_scale, group_start, group_end, kl, P, G);
  RCB_CHECK_LAUNCH("rcb_group_kl");
  return 0;
}

extern "C" int rcb_anneal_beta(float* beta, const double* kl, const uint8_t* coded, int rows, int G, double step,
                               double upper, double lower, double bits, rcb_stream_t stream) {
  RCB_CHECK_ARG(beta && kl, "rcb_anneal_beta: null tensor");
  int64_t total = (int64_t)rows * G;
  if (total <= 0) return 0;
  anneal_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(beta, kl, coded, total, step, upper, lower, bits);
  RCB_CHECK_LAUNCH("rcb_anneal_beta");
  return 0;
}

extern "C" int rcb_pick_block(const double* kl, const uint8_t* coded, int* block, int rows, int G, rcb_stream_t stream) {
  RCB_CHECK_ARG(kl && coded && block, "rcb_pick_block: null tensor");
  if (rows <= 0) return 0;
  pick_block_kernel<<<ceil_div(rows, 8), 256, 0, (cudaStream_t)stream>>>(kl, coded, block, rows, G);
  RCB_CHECK_LAUNCH("rcb_pick_block");
  return 0;
}

extern "C" int rcb_to_half(const float* src, void* dst, int64_t n, rcb_stream_t stream) {
  RCB_CHECK_ARG(src && dst && n > 0, "rcb_to_half: bad arguments");
  const int64_t blocks = (n + 1023) / 1024;
  RCB_CHECK_ARG(blocks < (1ll << 31), "rcb_to_half: tensor too large");
  to_half_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, (__half*)dst, n);
  RCB_CHECK_LAUNCH("rcb_to_half");
  return 0;
}

extern "C" int rcb_transpose(const float* in, int64_t ld_in, float* out, int64_t ld_out, int rows, int cols,
                             rcb_stream_t stream) {
  RCB_CHECK_ARG(in && out, "rcb_transpose: null tensor");
  RCB_CHECK_ARG(rows > 0 && cols > 0 && ld_in >= cols && ld_out >= rows, "rcb_transpose: bad shape");
  dim3 grid(ceil_div(cols, 32), ceil_div(rows, 32));
  RCB_CHECK_ARG(grid.y <= 65535, "rcb_transpose: too many rows");
  transpose_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, ld_in, out, ld_out, rows, cols);
  RCB_CHECK_LAUNCH("rcb_transpose");
  return 0;
}

extern "C" int rcb_transpose_xshift(const float* in, float* out, int64_t rows, int cols, int w, rcb_stream_t stream) {
  RCB_CHECK_ARG(in && out, "rcb_transpose_xshift: null tensor");
  RCB_CHECK_ARG(rows > 0 && cols > 0 && w > 0 && rows % w == 0, "rcb_transpose_xshift: rows must be whole lines of w pixels");
  if (cols <= 64 && cols % 4 == 0 && (((uintptr_t)in) & 15) == 0) {
    const int tr = tw_rows(cols);
    const size_t smem = sizeof(float) * (tr + 2) * (size_t)(cols + 1);
    transpose_wide_kernel<0><<<(unsigned)ceil_div(rows, (int64_t)tr), 256, smem, (cudaStream_t)stream>>>(in, out, rows, cols, 1, w, 1, 1, tr);
    RCB_CHECK_LAUNCH("rcb_transpose_xshift");
    return 0;
  }
  dim3 grid(ceil_div(cols, 32), ceil_div(rows, 32));
  RCB_CHECK_ARG(grid.y <= 65535 * 32, "rcb_transpose_xshift: too many rows");
  transpose_xshift_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, out, rows, cols, w);
  RCB_CHECK_LAUNCH("rcb_transpose_xshift");
  return 0;
}

extern "C" int rcb_transpose_phases(const float* in, float* out, int64_t rows, int cols, int h, int w, int fy, int fx,
                                    rcb_stream_t stream) {
  RCB_CHECK_ARG(in && out, "rcb_transpose_phases: null tensor");
  RCB_CHECK_ARG(rows > 0 && cols > 0 && h > 0 && w > 0 && fy > 0 && fx > 0 && rows % ((int64_t)h * fy * w * fx) == 0,
                "rcb_transpose_phases: rows must be whole (h*fy, w*fx) grids");
  if (cols <= 64 && cols % 4 == 0 && (((uintptr_t)in) & 15) == 0) {
    const int tr = tw_rows(cols);
    const size_t smem = sizeof(float) * (tr + 2) * (size_t)(cols + 1);
    transpose_wide_kernel<1><<<(unsigned)ceil_div(rows, (int64_t)tr), 256, smem, (cudaStream_t)stream>>>(in, out, rows, cols, h, w, fy, fx, tr);
    RCB_CHECK_LAUNCH("rcb_transpose_phases");
    return 0;
  }
  dim3 grid(ceil_div(cols, 32), ceil_div(rows, 32));
  RCB_CHECK_ARG(grid.y <= 65535 * 32, "rcb_transpose_phases: too many rows");
  transpose_phases_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, out, rows, cols, h, w, fy, fx);
  RCB_CHECK_LAUNCH("rcb_transpose_phases");
  return 0;
}

extern "C" int rcb_prior_suffstats(const float* loc, const float* log_scale, double* stats, int rows, int P,
                                   rcb_stream_t stream) {
  RCB_CHECK_ARG(loc && log_scale && stats, "rcb_prior_suffstats: null tensor");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(stats, 0, sizeof(double) * 3 * (size_t)P, st);
  if (e != cudaSuccess) { set_error("rcb_prior_suffstats: memset failed: %s", cudaGetErrorString(e)); return -1; }
  if (rows <= 0) return 0;
  const int rpb = 64;
  dim3 grid(ceil_div(P, 128), ceil_div(rows, rpb));
  suffstats_kernel<<<grid, 128, 0, st>>>(loc, log_scale, stats, rows, P, rpb);
  RCB_CHECK_LAUNCH("rcb_prior_suffstats");
  return 0;
}

extern "C" int rcb_prior_from_stats(const double* stats, float* p_loc, float* p_scale, int64_t n_total, int P,
                                    rcb_stream_t stream) {
  RCB_CHECK_ARG(stats && p_loc && p_scale, "rcb_prior_from_stats: null tensor");
  RCB_CHECK_ARG(n_total >= 2, "rcb_prior_from_stats: need at least 2 rows for the unbiased variance");
  prior_from_stats_kernel<<<ceil_div(P, 256), 256, 0, (cudaStream_t)stream>>>(stats, p_loc, p_scale, (double)n_total, P);
  RCB_CHECK_LAUNCH("rcb_prior_from_stats");
  return 0;
}
