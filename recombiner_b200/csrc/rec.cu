// Relative entropy coding: candidate table generation (scrambled Sobol by random
// access + Cephes inverse normal CDF in f64), batched A*/Gumbel-max scoring with a
// block-level first-argmax, and the receiver-side regeneration.
//
// Reference: test_model.py:441-533 (get_sobol_normal_sample, sample_group),
// :586-595 (compress_group).  All scoring is f64, as in the reference.
#include <math_constants.h>

#include "tc_common.cuh"

namespace rcb {

// ---- Cephes ndtri (the algorithm behind scipy.stats.norm.ppf) ----------------
__device__ __forceinline__ double ndtri_p0(double x) {
  double r = -5.99633501014107895267E1;
  r = r * x + 9.80010754185999661536E1;
  r = r * x + -5.66762857469070293439E1;
  r = r * x + 1.39312609387279679503E1;
  r = r * x + -1.23916583867381258016E0;
  return r;
}
__device__ __forceinline__ double ndtri_q0(double x) {
  double r = x + 1.95448858338141759834E0;
  r = r * x + 4.67627912898881538453E0;
  r = r * x + 8.63602421390890590575E1;
  r = r * x + -2.25462687854119370527E2;
  r = r * x + 2.00260212380060660359E2;
  r = r * x + -8.20372256168333339912E1;
  r = r * x + 1.59056225126211695515E1;
  r = r * x + -1.18331621121330003142E0;
  return r;
}
__device__ __forceinline__ double ndtri_p1(double x) {
  double r = 4.05544892305962419923E0;
  r = r * x + 3.15251094599893866154E1;
  r = r * x + 5.71628192246421288162E1;
  r = r * x + 4.40805073893200834700E1;
  r = r * x + 1.46849561928858024014E1;
  r = r * x + 2.18663306850790267539E0;
  r = r * x + -1.40256079171354495875E-1;
  r = r * x + -3.50424626827848203418E-2;
  r = r * x + -8.57456785154685413611E-4;
  return r;
}
__device__ __forceinline__ double ndtri_q1(double x) {
  double r = x + 1.57799883256466749731E1;
  r = r * x + 4.53907635128879210584E1;
  r = r * x + 4.13172038254672030440E1;
  r = r * x + 1.50425385692907503408E1;
  r = r * x + 2.50464946208309415979E0;
  r = r * x + -1.42182922854787788574E-1;
  r = r * x + -3.80806407691578277194E-2;
  r = r * x + -9.33259480895457427372E-4;
  return r;
}
__device__ __forceinline__ double ndtri_p2(double x) {
  double r = 3.23774891776946035970E0;
  r = r * x + 6.91522889068984211695E0;
  r = r * x + 3.93881025292474443415E0;
  r = r * x + 1.33303460815807542389E0;
  r = r * x + 2.01485389549179081538E-1;
  r = r * x + 1.23716634817820021358E-2;
  r = r * x + 3.01581553508235416007E-4;
  r = r * x + 2.65806974686737550832E-6;
  r = r * x + 6.23974539184983293730E-9;
  return r;
}
__device__ __forceinline__ double ndtri_q2(double x) {
  double r = x + 6.02427039364742014255E0;
  r = r * x + 3.67983563856160859403E0;
  r = r * x + 1.37702099489081330271E0;
  r = r * x + 2.16236993594496635890E-1;
  r = r * x + 1.34204006088543189037E-2;
  r = r * x + 3.28014464682127739104E-4;
  r = r * x + 2.89247864745380683936E-6;
  r = r * x + 6.79019408009981274425E-9;
  return r;
}

__device__ double ndtri(double y0) {
  if (y0 <= 0.0) return -CUDART_INF;
  if (y0 >= 1.0) return CUDART_INF;
  bool negate = true;
  double y = y0;
  if (y > 1.0 - 0.13533528323661269189) { y = 1.0 - y; negate = false; }
  if (y > 0.13533528323661269189) {
    y = y - 0.5;
    double y2 = y * y;
    double x = y + y * (y2 * ndtri_p0(y2) / ndtri_q0(y2));
    return x * 2.50662827463100050242E0;
  }
  double x = sqrt(-2.0 * log(y));
  double x0 = x - log(x) / x;
  double z = 1.0 / x;
  double x1 = (x < 8.0) ? z * ndtri_p1(z) / ndtri_q1(z) : z * ndtri_p2(z) / ndtri_q2(z);
  x = x0 - x1;
  return negate ? -x : x;
}

// grid (ceil(n/256), D): table[d*n + k]
__global__ void __launch_bounds__(256) rec_table_kernel(const int64_t* __restrict__ shift, const int64_t* __restrict__ words,
                                                        float* __restrict__ table, int n, int nbits) {
  __shared__ int64_t w[30];
  const int d = blockIdx.y;
  if (threadIdx.x < 30) w[threadIdx.x] = words[(int64_t)d * 30 + threadIdx.x];
  __syncthreads();
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint32_t gray = (uint32_t)k ^ ((uint32_t)k >> 1);
  int64_t q = shift[d];
  for (int j = 0; j < nbits; ++j) q ^= w[j] & -(int64_t)((gray >> j) & 1u);
  const float u = __ll2float_rn(q) * 9.31322574615478515625e-10f;   // * 2^-30
  float s = (float)ndtri((double)u);
  s = fminf(fmaxf(s, -100.f), 100.f);
  table[(int64_t)d * n + k] = s;
}

// ---- scoring -----------------------------------------------------------------
constexpr int REC_THREADS = 256;
constexpr int REC_ILP = 4;
constexpr int REC_RPC = 4;        // pairs per CTA; consecutive pairs that share a block share the table loads

// Score R rows that code the SAME block against the block's candidate table: every table value is loaded
// once and used for R quadratic forms.  Per row the arithmetic (order of d, mixed f32/f64 roundings) is
// exactly that of a single-row pass, so indices and log-weights do not depend on the grouping.
// coef: [R][2*D] (A_d, B_d); c0[r]; best[r] / best_k[r]: per-thread running first-argmax.
template <int R>
__device__ __forceinline__ void rec_score_run(const rcb_rec_args& a, const float* __restrict__ tab, const double* coef,
                                              const double* c0, int D, int n, int pair0, double (&best)[REC_RPC],
                                              int (&best_k)[REC_RPC]) {
  const int tid = threadIdx.x;
  for (int k0 = tid; k0 < n; k0 += REC_THREADS * REC_ILP) {
    double acc[R][REC_ILP];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int u = 0; u < REC_ILP; ++u) acc[r][u] = 0.0;
    for (int d = 0; d < D; ++d) {
      const float* col = tab + (int64_t)d * n + k0;
      double sv[REC_ILP];
#pragma unroll
      for (int u = 0; u < REC_ILP; ++u) sv[u] = (k0 + u * REC_THREADS < n) ? (double)__ldg(col + u * REC_THREADS) : 0.0;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const double A = coef[(r * D + d) * 2], B = coef[(r * D + d) * 2 + 1];
#pragma unroll
        for (int u = 0; u < REC_ILP; ++u) acc[r][u] = fma(sv[u], fma(A, sv[u], B), acc[r][u]);
      }
    }
#pragma unroll
    for (int u = 0; u < REC_ILP; ++u) {
      const int k = k0 + u * REC_THREADS;
      if (k < n) {
        const double gk = a.gumbel[k];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const double lw = acc[r][u] + c0[r] + gk;
          if (a.logw_out) a.logw_out[(int64_t)(pair0 + r) * n + k] = lw;
          if (lw > best[r]) { best[r] = lw; best_k[r] = k; }     // k increases: keeps the first maximum
        }
      }
    }
  }
}

// One CTA per REC_RPC consecutive (row, block) pairs.  Runs of equal blocks inside the CTA are scored together;
// callers that sort their pairs by block (TestBNNmodel.compress_round does) get the table reuse.
__global__ void __launch_bounds__(REC_THREADS) rec_encode_kernel(rcb_rec_args a, int rpc) {
  extern __shared__ __align__(16) double coef[];   // [rpc][2*max_D] (A_d, B_d)
  __shared__ double red_v[REC_THREADS / 32];
  __shared__ int red_i[REC_THREADS / 32];
  __shared__ double c0_s[REC_RPC];
  __shared__ int best_s[REC_RPC];

  const int n = a.n_cand;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int p_begin = blockIdx.x * rpc, p_end = min(a.n_pairs, p_begin + rpc);

  for (int pair0 = p_begin; pair0 < p_end;) {
    const int blk = a.pair_block[pair0];
    if (blk < 0 || blk >= a.G) __trap();           // a diverged caller must fail loudly, not read out of range
    int R = 1;
    while (pair0 + R < p_end && a.pair_block[pair0 + R] == blk) ++R;
    const int start = a.group_start[blk], D = a.group_end[blk] - start;
    const float* __restrict__ tab = a.tables[blk];

    // per-dimension quadratic-form coefficients, reproducing the reference's mixed
    // precision: var = scale*scale and log(scale) are f32, everything else f64
    for (int r = 0; r < R; ++r) {
      const int row = a.pair_row[pair0 + r];
      double c_part = 0.0;
      for (int d = tid; d < D; d += REC_THREADS) {
        const float mq = a.q_loc[(int64_t)row * a.P + start + d], sq = a.q_scale[(int64_t)row * a.P + start + d];
        const float mp = a.p_loc[start + d], sp = a.p_scale[start + d];
        const double kp = 1.0 / (2.0 * (double)__fmul_rn(sp, sp));
        const double kq = 1.0 / (2.0 * (double)__fmul_rn(sq, sq));
        const double delta = (double)mp - (double)mq;
        const double spd = (double)sp;
        coef[(r * D + d) * 2] = spd * spd * (kp - kq);
        coef[(r * D + d) * 2 + 1] = -2.0 * spd * delta * kq;
        c_part += -delta * delta * kq + (double)logf(sp) - (double)logf(sq);
      }
      c_part = warp_sum(c_part);
      __syncthreads();
      if (lane == 0) red_v[warp] = c_part;
      __syncthreads();
      if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < REC_THREADS / 32; ++w) t += red_v[w];
        c0_s[r] = t;
      }
    }
    __syncthreads();
    double c0[REC_RPC], best[REC_RPC];
    int best_k[REC_RPC];
#pragma unroll
    for (int r = 0; r < REC_RPC; ++r) { c0[r] = r < R ? c0_s[r] : 0.0; best[r] = -CUDART_INF; best_k[r] = 0x7fffffff; }

    if (R == 1) rec_score_run<1>(a, tab, coef, c0, D, n, pair0, best, best_k);
    else if (R == 2) rec_score_run<2>(a, tab, coef, c0, D, n, pair0, best, best_k);
    else if (R == 3) rec_score_run<3>(a, tab, coef, c0, D, n, pair0, best, best_k);
    else rec_score_run<4>(a, tab, coef, c0, D, n, pair0, best, best_k);

    // block-level first-argmax, one row at a time
#pragma unroll
    for (int r = 0; r < REC_RPC; ++r) {
      if (r < R) {
        double bv = best[r];
        int bi = best_k[r];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          double ov = __shfl_xor_sync(0xffffffffu, bv, o);
          int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        __syncthreads();
        if (lane == 0) { red_v[warp] = bv; red_i[warp] = bi; }
        __syncthreads();
        if (tid == 0) {
          double v = red_v[0]; int i = red_i[0];
          for (int w = 1; w < REC_THREADS / 32; ++w)
            if (red_v[w] > v || (red_v[w] == v && red_i[w] < i)) { v = red_v[w]; i = red_i[w]; }
          best_s[r] = i;
          const int row = a.pair_row[pair0 + r];
          if (a.apply) {
            a.idx_out[(int64_t)row * a.G + blk] = i;
            if (a.beta) a.beta[(int64_t)row * a.G + blk] = 0.f;
            if (a.coded) a.coded[(int64_t)row * a.G + blk] = 1;
          } else {
            a.idx_out[pair0 + r] = i;
          }
        }
      }
    }
    __syncthreads();
    for (int r = 0; r < R; ++r) {
      const int kb = best_s[r];
      if (kb >= n) continue;   // n == 0
      const int row = a.pair_row[pair0 + r];
      for (int d = tid; d < D; d += REC_THREADS) {
        const double sd = (double)tab[(int64_t)d * n + kb];
        // z = mu_p + sigma_p * s in f64 with two roundings, then f32 on store (test_model.py:514,591)
        const float z = (float)__dadd_rn((double)a.p_loc[start + d], __dmul_rn((double)a.p_scale[start + d], sd));
        if (a.apply) {
          a.sample[(int64_t)row * a.P + start + d] = z;
          a.mask[(int64_t)row * a.P + start + d] = 1.f;
        } else if (a.z_out) {
          a.z_out[(int64_t)(pair0 + r) * a.max_D + d] = z;
        }
      }
    }
    __syncthreads();           // coef / best_s are rewritten by the next run
    pair0 += R;
  }
}

// ---- staged scoring (the path the coder normally runs) --------------------------
// Same arithmetic as rec_encode_kernel, restructured so that the FP64 pipe is not left waiting on table loads:
//  * the candidate table travels global -> shared memory as bulk async copies (cp.async.bulk + mbarrier, one
//    2 KB row of 512 candidates per dimension, 8 dimensions per stage, 4 stages in flight), issued by one thread
//    and decoupled from the threads that consume them;
//  * up to RS_RPC = 8 rows that code the same block share every table value (8 x 2 x 2 DFMA per LDS.64);
//  * the 2^16 candidates of a run are split over `splits` CTAs (grid.y) so that a round of a few hundred runs
//    still fills 148 SMs; each CTA leaves its (max, first index) in the workspace and the last one to arrive
//    (a counter per pair) reduces them in candidate order and commits the result.
// Per (row, candidate) the sum over d runs in ascending d with the same fma sequence, so indices, samples and
// log-weights are bit-identical to the unstaged kernel.
constexpr int RS_THREADS = 256, RS_ILP = 2, RS_CH = RS_THREADS * RS_ILP, RS_DS = 8, RS_NS = 4, RS_RPC = 8;
constexpr int RS_ROW_BYTES = RS_CH * 4, RS_STAGE_BYTES = RS_DS * RS_ROW_BYTES;

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int R>
__device__ __forceinline__ uint32_t rec_staged_run(const rcb_rec_args& a, const float* __restrict__ tab, const double* coef,
                                                const double* c0, int D, int n, int pair0, int chunk_begin, int chunk_end,
                                                uint32_t stages, uint64_t* full, uint32_t it_base,
                                                double* red_v, int* red_i, double* cta_v, int* cta_k) {
  const int tid = threadIdx.x;
  double best[R];
  int best_k[R];
#pragma unroll
  for (int r = 0; r < R; ++r) { best[r] = -CUDART_INF; best_k[r] = 0x7fffffff; }
  const int gpc = (D + RS_DS - 1) / RS_DS;                       // dimension groups per chunk
  const int total = (chunk_end - chunk_begin) * gpc;
  // producer state (thread 0): the next (chunk, group) to request
  int p_chunk = chunk_begin, p_g = 0, p_i = 0;
  auto issue = [&]() {
    const int d0 = p_g * RS_DS, nd = min(RS_DS, D - d0);
    const uint32_t bytes = (uint32_t)min(RS_CH, n - p_chunk * RS_CH) * 4u;
    const uint32_t s = (it_base + (uint32_t)p_i) % RS_NS;
    mbar_expect_tx(&full[s], (uint32_t)nd * bytes);
    const float* src = tab + (int64_t)d0 * n + (int64_t)p_chunk * RS_CH;
#pragma unroll 1
    for (int j = 0; j < nd; ++j) bulk_g2s(stages + s * RS_STAGE_BYTES + j * RS_ROW_BYTES, src + (int64_t)j * n, bytes, &full[s]);
    ++p_i;
    if (++p_g == gpc) { p_g = 0; ++p_chunk; }
  };
  if (tid == 0)
    for (int i = 0; i < RS_NS && i < total; ++i) issue();

  double acc[R][RS_ILP];
  int chunk = chunk_begin, g = 0;
  for (int i = 0; i < total; ++i) {
    const uint32_t gi = it_base + (uint32_t)i, s = gi % RS_NS;
    if (g == 0) {
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int u = 0; u < RS_ILP; ++u) acc[r][u] = 0.0;
    }
    const int k0 = chunk * RS_CH + tid * RS_ILP;
    const int d0 = g * RS_DS, nd = min(RS_DS, D - d0);
    mbar_wait(&full[s], (gi / RS_NS) & 1u);
    const uint32_t row0 = stages + s * RS_STAGE_BYTES + (uint32_t)tid * (RS_ILP * 4);
#pragma unroll
    for (int j = 0; j < RS_DS; ++j) {
      if (j < nd) {
        float v0, v1;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v0), "=f"(v1) : "r"(row0 + j * RS_ROW_BYTES));
        double sv[RS_ILP];
        sv[0] = (k0 < n) ? (double)v0 : 0.0;
        sv[1] = (k0 + 1 < n) ? (double)v1 : 0.0;
        const double2* cf = reinterpret_cast<const double2*>(coef) + (d0 + j);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const double2 ab = cf[r * D];
#pragma unroll
          for (int u = 0; u < RS_ILP; ++u) acc[r][u] = fma(sv[u], fma(ab.x, sv[u], ab.y), acc[r][u]);
        }
      }
    }
    __syncthreads();                                  // every thread is done with stage s
    if (tid == 0 && p_i < total) issue();             // refill it with the group RS_NS ahead
    if (g == gpc - 1) {
#pragma unroll
      for (int u = 0; u < RS_ILP; ++u) {
        const int k = k0 + u;
        if (k < n) {
          const double gk = a.gumbel[k];
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const double lw = acc[r][u] + c0[r] + gk;          // c0 lives in shared memory
            if (a.logw_out) a.logw_out[(int64_t)(pair0 + r) * n + k] = lw;
            if (lw > best[r]) { best[r] = lw; best_k[r] = k; }     // k increases per thread: keeps the first maximum
          }
        }
      }
      g = 0;
      ++chunk;
    } else {
      ++g;
    }
  }
  it_base += (uint32_t)total;
  // CTA-level first-argmax per row -> cta_v[r], cta_k[r]
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    double bv = best[r];
    int bi = best_k[r];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      double ov = __shfl_xor_sync(0xffffffffu, bv, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    __syncthreads();
    if (lane == 0) { red_v[warp] = bv; red_i[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
      double v = red_v[0]; int i = red_i[0];
      for (int w = 1; w < RS_THREADS / 32; ++w)
        if (red_v[w] > v || (red_v[w] == v && red_i[w] < i)) { v = red_v[w]; i = red_i[w]; }
      cta_v[r] = v; cta_k[r] = i;
    }
  }
  __syncthreads();
  return it_base;
}

// grid (ceil(n_pairs / rpc), splits).  ws: double part_v[n_pairs][splits]; int part_k[n_pairs][splits];
// unsigned arrived[n_pairs] (zero on entry, left zero on exit).
__global__ void __launch_bounds__(RS_THREADS, 2) rec_encode_staged_kernel(rcb_rec_args a, int rpc, int splits,
                                                                         double* part_v, int* part_k, unsigned* arrived) {
  extern __shared__ __align__(128) uint8_t rs_smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(rs_smem);                    // [RS_NS]
  const uint32_t stages = smem_u32(rs_smem + 128);
  double* coef = reinterpret_cast<double*>(rs_smem + 128 + RS_NS * RS_STAGE_BYTES);     // [rpc][D][2]
  __shared__ double red_v[RS_THREADS / 32];
  __shared__ int red_i[RS_THREADS / 32];
  __shared__ double c0_s[RS_RPC];
  __shared__ int best_s[RS_RPC];
  __shared__ double cta_v[RS_RPC];
  __shared__ int cta_k[RS_RPC];

  const int n = a.n_cand;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int p_begin = blockIdx.x * rpc, p_end = min(a.n_pairs, p_begin + rpc);
  const int split = blockIdx.y;
  const int n_chunks = (n + RS_CH - 1) / RS_CH, cps = (n_chunks + splits - 1) / splits;
  const int chunk_begin = min(n_chunks, split * cps), chunk_end = min(n_chunks, chunk_begin + cps);
  if (tid == 0) {
    for (int s = 0; s < RS_NS; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t it_base = 0;

  for (int pair0 = p_begin; pair0 < p_end;) {
    const int blk = a.pair_block[pair0];
    if (blk < 0 || blk >= a.G) __trap();           // a diverged caller must fail loudly, not read out of range
    int R = 1;
    while (pair0 + R < p_end && a.pair_block[pair0 + R] == blk) ++R;
    R = R >= 8 ? 8 : (R >= 4 ? 4 : (R >= 2 ? 2 : 1));      // scored in power-of-two sub-runs; the rest forms the next run
    const int start = a.group_start[blk], D = a.group_end[blk] - start;
    const float* __restrict__ tab = a.tables[blk];

    for (int r = 0; r < R; ++r) {                   // coefficients: identical to rec_encode_kernel
      const int row = a.pair_row[pair0 + r];
      double c_part = 0.0;
      for (int d = tid; d < D; d += RS_THREADS) {
        const float mq = a.q_loc[(int64_t)row * a.P + start + d], sq = a.q_scale[(int64_t)row * a.P + start + d];
        const float mp = a.p_loc[start + d], sp = a.p_scale[start + d];
        const double kp = 1.0 / (2.0 * (double)__fmul_rn(sp, sp));
        const double kq = 1.0 / (2.0 * (double)__fmul_rn(sq, sq));
        const double delta = (double)mp - (double)mq;
        const double spd = (double)sp;
        coef[(r * D + d) * 2] = spd * spd * (kp - kq);
        coef[(r * D + d) * 2 + 1] = -2.0 * spd * delta * kq;
        c_part += -delta * delta * kq + (double)logf(sp) - (double)logf(sq);
      }
      c_part = warp_sum(c_part);
      __syncthreads();
      if (lane == 0) red_v[warp] = c_part;
      __syncthreads();
      if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < RS_THREADS / 32; ++w) t += red_v[w];
        c0_s[r] = t;
      }
    }
    __syncthreads();
#define RCB_RS_RUN(RR) it_base = rec_staged_run<RR>(a, tab, coef, c0_s, D, n, pair0, chunk_begin, chunk_end, stages, full, it_base, \
                                          red_v, red_i, cta_v, cta_k)
    if (R == 8) RCB_RS_RUN(8);
    else if (R == 4) RCB_RS_RUN(4);
    else if (R == 2) RCB_RS_RUN(2);
    else RCB_RS_RUN(1);
#undef RCB_RS_RUN

    // cross-CTA reduction by the last CTA of each pair (candidate order: split ascending), then commit
    if (tid < R) {
      const int r = tid, pair = pair0 + r;
      double v = cta_v[r];
      int i = cta_k[r];
      int final_i = -1;
      if (splits > 1) {
        part_v[(int64_t)pair * splits + split] = v;
        part_k[(int64_t)pair * splits + split] = i;
        __threadfence();
        if (atomicAdd(&arrived[pair], 1u) == (unsigned)(splits - 1)) {
          __threadfence();
          v = -CUDART_INF; i = 0x7fffffff;
          for (int sp = 0; sp < splits; ++sp) {
            const double pv = __ldcg(&part_v[(int64_t)pair * splits + sp]);
            const int pk = __ldcg(&part_k[(int64_t)pair * splits + sp]);
            if (pv > v || (pv == v && pk < i)) { v = pv; i = pk; }
          }
          arrived[pair] = 0u;                  // ready for the next launch
          final_i = i;
        }
      } else {
        final_i = i;
      }
      best_s[r] = final_i;
      if (final_i >= 0) {
        const int row = a.pair_row[pair];
        if (a.apply) {
          a.idx_out[(int64_t)row * a.G + blk] = final_i;
          if (a.beta) a.beta[(int64_t)row * a.G + blk] = 0.f;
          if (a.coded) a.coded[(int64_t)row * a.G + blk] = 1;
        } else {
          a.idx_out[pair] = final_i;
        }
      }
    }
    __syncthreads();
    for (int r = 0; r < R; ++r) {
      const int kb = best_s[r];
      if (kb < 0 || kb >= n) continue;             // not the last CTA of this pair (or n == 0)
      const int row = a.pair_row[pair0 + r];
      for (int d = tid; d < D; d += RS_THREADS) {
        const double sd = (double)tab[(int64_t)d * n + kb];
        const float z = (float)__dadd_rn((double)a.p_loc[start + d], __dmul_rn((double)a.p_scale[start + d], sd));
        if (a.apply) {
          a.sample[(int64_t)row * a.P + start + d] = z;
          a.mask[(int64_t)row * a.P + start + d] = 1.f;
        } else if (a.z_out) {
          a.z_out[(int64_t)(pair0 + r) * a.max_D + d] = z;
        }
      }
    }
    __syncthreads();           // coef / best_s are rewritten by the next run
    pair0 += R;
  }
}

__global__ void rec_decode_kernel(const int* __restrict__ pair_row, const int* __restrict__ pair_block,
                                  const int* __restrict__ idx, const float* __restrict__ p_loc,
                                  const float* __restrict__ p_scale, const int* __restrict__ gs,
                                  const int* __restrict__ ge, const float* const* __restrict__ tables,
                                  float* __restrict__ sample, float* __restrict__ mask, int P, int n) {
  const int pair = blockIdx.x;
  const int row = pair_row[pair], blk = pair_block[pair];
  const int start = gs[blk], D = ge[blk] - start;
  const float* tab = tables[blk];
  const int k = idx[pair];
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const double s = (double)tab[(int64_t)d * n + k];
    const float z = (float)__dadd_rn((double)p_loc[start + d], __dmul_rn((double)p_scale[start + d], s));
    sample[(int64_t)row * P + start + d] = z;
    if (mask) mask[(int64_t)row * P + start + d] = 1.f;
  }
}

// Pairs of a round grouped by block (counting sort in one CTA): rows_out / blocks_out list the rows 0 .. n-1 so that
// equal blocks are adjacent -- the scoring kernel shares every table value between the rows of such a run.  The order
// inside a run is whatever the atomics give; each row's result is independent of the grouping (bit-identical).
__global__ void __launch_bounds__(1024) rec_order_kernel(const int* __restrict__ blocks, int* __restrict__ rows_out,
                                                         int* __restrict__ blocks_out, int n, int G) {
  extern __shared__ int ro_cnt[];                   // [G] counts, then running offsets
  for (int g = threadIdx.x; g < G; g += blockDim.x) ro_cnt[g] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int b = blocks[i];
    if (b < 0 || b >= G) __trap();
    atomicAdd(&ro_cnt[b], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int g = 0; g < G; ++g) { const int c = ro_cnt[g]; ro_cnt[g] = run; run += c; }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int b = blocks[i];
    const int pos = atomicAdd(&ro_cnt[b], 1);
    rows_out[pos] = i;
    blocks_out[pos] = b;
  }
}

// FP64 FMA peak of the device, for the roofline of the REC scoring kernel (bench.py): 16 independent DFMA chains per
// thread, nothing else in the loop.  2 * 16 * iters * threads FLOP per launch.
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double seed) {
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = seed + (double)(threadIdx.x + i);
  const double m = 1.0 + 1e-9 * seed, c = 1e-12;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], m, c);
  }
  double t = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) t += acc[i];
  if (t == 123.456) out[0] = t;                    // never true: keeps the chains alive
}

}  // namespace rcb

using namespace rcb;

extern "C" int rcb_rec_order(const int* blocks, int* rows_out, int* blocks_out, int n, int G, rcb_stream_t stream) {
  RCB_CHECK_ARG(blocks && rows_out && blocks_out && n > 0 && G > 0 && G <= 40000, "rcb_rec_order: bad arguments");
  const size_t smem = sizeof(int) * (size_t)G;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(rec_order_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("rcb_rec_order: smem opt-in failed: %s", cudaGetErrorString(e)); return -1; }
  }
  rec_order_kernel<<<1, 1024, smem, (cudaStream_t)stream>>>(blocks, rows_out, blocks_out, n, G);
  RCB_CHECK_LAUNCH("rcb_rec_order");
  return 0;
}

extern "C" int rcb_ubench_dfma(double* scratch, int ctas, int iters, double* flop_out, rcb_stream_t stream) {
  RCB_CHECK_ARG(scratch && ctas > 0 && iters > 0, "rcb_ubench_dfma: bad arguments");
  dfma_peak_kernel<<<ctas, 256, 0, (cudaStream_t)stream>>>(scratch, iters, 1.0);
  RCB_CHECK_LAUNCH("rcb_ubench_dfma");
  if (flop_out) *flop_out = 2.0 * 16.0 * (double)iters * 256.0 * (double)ctas;
  return 0;
}

extern "C" int rcb_rec_table(const int64_t* shift, const int64_t* words, float* table, int D, int n, rcb_stream_t stream) {
  RCB_CHECK_ARG(shift && words && table, "rcb_rec_table: null tensor");
  RCB_CHECK_ARG(D > 0 && D <= 65535 && n > 0 && n <= (1 << 30), "rcb_rec_table: bad shape D=%d n=%d", D, n);
  int nbits = 0;
  while (nbits < 30 && ((int64_t)1 << nbits) < n) ++nbits;     // gray(k) < 2^nbits for k < n
  dim3 grid(ceil_div(n, 256), D);
  rec_table_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(shift, words, table, n, nbits);
  RCB_CHECK_LAUNCH("rcb_rec_table");
  return 0;
}

extern "C" int rcb_rec_encode(const rcb_rec_args* a, rcb_stream_t stream) {
  RCB_CHECK_ARG(a != nullptr, "rcb_rec_encode: null args");
  RCB_CHECK_ARG(a->pair_row && a->pair_block && a->q_loc && a->q_scale && a->p_loc && a->p_scale &&
                a->group_start && a->group_end && a->tables && a->gumbel && a->idx_out, "rcb_rec_encode: null tensor");
  RCB_CHECK_ARG(!a->apply || (a->sample && a->mask), "rcb_rec_encode: apply needs sample and mask");
  RCB_CHECK_ARG(a->max_D > 0 && a->max_D <= 12000, "rcb_rec_encode: max_D %d out of range (1..12000)", a->max_D);
  RCB_CHECK_ARG(a->n_cand > 0, "rcb_rec_encode: no candidates");
  if (a->n_pairs <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  // staged kernel: needs 16-byte aligned table rows (n % 4 == 0), the caller's workspace and room for the coefficients
  if (a->n_cand % 4 == 0 && a->workspace) {
    int rpc = RS_RPC;
    const size_t fixed = 128 + (size_t)RS_NS * RS_STAGE_BYTES;
    while (rpc > 1 && fixed + sizeof(double) * 2 * (size_t)a->max_D * rpc > 100 * 1024) --rpc;
    const size_t smem = fixed + sizeof(double) * 2 * (size_t)a->max_D * rpc;
    if (smem <= 200 * 1024) {
      const int runs = ceil_div(a->n_pairs, rpc);
      const int n_chunks = ceil_div(a->n_cand, RS_CH);
      // enough CTAs for ~4 waves of 2 CTAs per SM, never more splits than chunks
      int splits = 1;
      while (splits < 64 && runs * splits < 148 * 8 && splits * 2 <= n_chunks) splits *= 2;
      // Workspace layout depends on its SIZE only, never on this call's n_pairs / splits: the arrival counters of
      // `cap` pairs come first (they must be zero on entry and are left zero), the partial results after them.  A
      // layout that moved with n_pairs would let one call's partial results land on the next call's counters.
      const size_t cap = a->workspace_bytes > 64 ? ((size_t)a->workspace_bytes - 64) / (4 + 64 * 12) : 0;
      if ((size_t)a->n_pairs <= cap) {
        uint8_t* w = reinterpret_cast<uint8_t*>(a->workspace);
        unsigned* arrived = reinterpret_cast<unsigned*>(w);
        double* part_v = reinterpret_cast<double*>(w + ((cap * 4 + 63) / 64) * 64);
        int* part_k = reinterpret_cast<int*>(part_v + (size_t)a->n_pairs * splits);
        cudaError_t e = cudaFuncSetAttribute(rec_encode_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("rcb_rec_encode: smem opt-in failed: %s", cudaGetErrorString(e)); return -1; }
        rec_encode_staged_kernel<<<dim3(runs, splits), RS_THREADS, smem, st>>>(*a, rpc, splits, part_v, part_k, arrived);
        RCB_CHECK_LAUNCH("rcb_rec_encode");
        return 0;
      }
    }
  }
  int rpc = REC_RPC;
  while (rpc > 1 && sizeof(double) * 2 * (size_t)a->max_D * rpc > 96 * 1024) --rpc;
  size_t smem = sizeof(double) * 2 * (size_t)a->max_D * rpc;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(rec_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("rcb_rec_encode: smem opt-in failed: %s", cudaGetErrorString(e)); return -1; }
  }
  rec_encode_kernel<<<ceil_div(a->n_pairs, rpc), REC_THREADS, smem, st>>>(*a, rpc);
  RCB_CHECK_LAUNCH("rcb_rec_encode");
  return 0;
}

extern "C" int rcb_rec_decode(const int* pair_row, const int* pair_block, const int* idx, const float* p_loc,
                              const float* p_scale, const int* group_start, const int* group_end,
                              const float* const* tables, float* sample, float* mask, int n_pairs, int P, int n_cand,
                              rcb_stream_t stream) {
  RCB_CHECK_ARG(pair_row && pair_block && idx && p_loc && p_scale && group_start && group_end && tables && sample,
                "rcb_rec_decode: null tensor");
  if (n_pairs <= 0) return 0;
  rec_decode_kernel<<<n_pairs, 128, 0, (cudaStream_t)stream>>>(pair_row, pair_block, idx, p_loc, p_scale, group_start,
                                                               group_end, tables, sample, mask, P, n_cand);
  RCB_CHECK_LAUNCH("rcb_rec_decode");
  return 0;
}
