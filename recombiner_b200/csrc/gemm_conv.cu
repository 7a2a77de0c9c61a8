// Plain GEMM, weight folding and polyphase up-convolution entry points.
#include "gemm_engine.cuh"

namespace rcb {

static int make_geom(const rcb_upconv_geom* g, PolyGeom* out) {
  RCB_CHECK_ARG(g != nullptr, "upconv: null geometry");
  RCB_CHECK_ARG(g->d > 0 && g->h > 0 && g->w > 0 && g->fz > 0 && g->fy > 0 && g->fx > 0, "upconv: bad grid/factors");
  RCB_CHECK_ARG((g->kz & 1) && (g->ky & 1) && (g->kx & 1), "upconv: kernel extents must be odd");
  RCB_CHECK_ARG(g->ic % 16 == 0 && g->oc % 16 == 0, "upconv: channels must be multiples of 16");
  int pz = (g->kz - 1) / 2, py = (g->ky - 1) / 2, px = (g->kx - 1) / 2;
  RCB_CHECK_ARG(pz < g->fz || pz == 0, "upconv: padding %d must be < factor %d", pz, g->fz);
  RCB_CHECK_ARG(py < g->fy || py == 0, "upconv: padding %d must be < factor %d", py, g->fy);
  RCB_CHECK_ARG(px < g->fx || px == 0, "upconv: padding %d must be < factor %d", px, g->fx);
  out->d = g->d; out->h = g->h; out->w = g->w;
  out->fz = g->fz; out->fy = g->fy; out->fx = g->fx;
  out->pz = pz; out->py = py; out->px = px;
  out->Tz = (pz == 0) ? 1 : 2;
  out->Ty = (py == 0) ? 1 : 2;
  out->Tx = (px == 0) ? 1 : 2;
  out->ic = g->ic; out->oc = g->oc;
  return 0;
}

// tap index in {0,1} of kernel position kk for phase r (folding is border-agnostic;
// borders are handled by the zero-padded gather)
__device__ __forceinline__ int tap_of(int r, int kk, int p, int f) {
  int num = r + kk - p;                       // offset on the upsampled grid
  int off = (num >= 0) ? num / f : -((-num + f - 1) / f);   // floor division
  int base = r < p ? -1 : 0;
  return off - base;
}

// w_eff[phase][tap][ic][oc] (and the [..][oc][ic] transpose)
__global__ void fold_poly_kernel(const float* __restrict__ w, PolyGeom g, int kz, int ky, int kx,
                                 float* __restrict__ w_eff, float* __restrict__ w_eff_t, float* __restrict__ w_eff_k) {
  int64_t total = (int64_t)g.phases() * g.taps() * g.ic * g.oc;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int o = e % g.oc; int64_t r = e / g.oc;
    int c = r % g.ic; r /= g.ic;
    int tap = r % g.taps(); int ph = r / g.taps();
    int tz, ty, tx, rz, ry, rx;
    g.split_tap(tap, tz, ty, tx);
    g.split_phase(ph, rz, ry, rx);
    float s = 0.f;
    for (int a0 = 0; a0 < kz; ++a0) {
      if (tap_of(rz, a0, g.pz, g.fz) != tz) continue;
      for (int a = 0; a < ky; ++a) {
        if (tap_of(ry, a, g.py, g.fy) != ty) continue;
        for (int b = 0; b < kx; ++b) {
          if (tap_of(rx, b, g.px, g.fx) != tx) continue;
          s += w[((((int64_t)o * g.ic + c) * kz + a0) * ky + a) * kx + b];
        }
      }
    }
    if (w_eff) w_eff[e] = s;
    if (w_eff_t) {
      int64_t seg = (int64_t)ph * g.taps() + tap;
      w_eff_t[(seg * g.oc + o) * g.ic + c] = s;
    }
    if (w_eff_k) w_eff_k[((int64_t)ph * g.oc + o) * ((int64_t)g.taps() * g.ic) + (int64_t)tap * g.ic + c] = s;
  }
}

// dense fold (2-D / 1-D grids only): m[(sy,sx,ic)][(oy,ox,oc)] = sum of the taps (a, b) of w[oc][ic][a][b] that carry
// source pixel (sy, sx) to output pixel (oy, ox) through the nearest-upsampling; m_t is its transpose.
// One thread per (output pixel, (oc, ic) pair): the taps of a pixel fall on at most 2 x 2 source pixels (KY - 1 <= fy,
// KX - 1 <= fx), whose sums are written; every other entry of m / m_t is structurally zero and is never touched (the
// caller zeroes the buffers once).  TRANSPOSED selects which of the two matrices is written and the thread order that
// makes its stores contiguous (oc fastest for m, ic for m_t).  (A thread per matrix ELEMENT re-deriving its taps with
// integer divisions cost 60 us per matrix for the 512 x 4096 cifar fold; a thread per pair walking all pixels, 170 us.)
template <int KY, int KX, bool TRANSPOSED>
__global__ void __launch_bounds__(128) fold_dense_kernel(const float* __restrict__ w, PolyGeom g, float* __restrict__ out) {
  const int H = g.h * g.fy, W = g.w * g.fx;
  const int64_t rows = (int64_t)g.h * g.w * g.ic, cols = (int64_t)H * W * g.oc;
  const int npairs = g.oc * g.ic;
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= (int64_t)npairs * H * W) return;
  const int pair = (int)(e % npairs), pixel = (int)(e / npairs);
  const int o = TRANSPOSED ? pair / g.ic : pair % g.oc, c = TRANSPOSED ? pair % g.ic : pair / g.oc;
  float wr[KY][KX];
#pragma unroll
  for (int a = 0; a < KY; ++a)
#pragma unroll
    for (int b = 0; b < KX; ++b) wr[a][b] = __ldg(w + (((int64_t)o * g.ic + c) * KY + a) * KX + b);
  {
    const int oy = pixel / W, ox = pixel - oy * W;
    const int sy0 = max(oy - g.py, 0) / g.fy;
    {
      const int sx0 = max(ox - g.px, 0) / g.fx;
      float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
      int ixs[KX];                                 // per tap column: 0 / 1 = which source column, -1 = outside
#pragma unroll
      for (int b = 0; b < KX; ++b) {
        const int ux = ox + b - g.px;
        ixs[b] = (ux < 0 || ux >= W) ? -1 : ux / g.fx - sx0;
      }
#pragma unroll
      for (int a = 0; a < KY; ++a) {
        const int uy = oy + a - g.py;
        if (uy < 0 || uy >= H) continue;
        const bool hi = uy / g.fy != sy0;
        float r0 = 0.f, r1 = 0.f;                  // this tap row's sums for the two source columns, in b order
#pragma unroll
        for (int b = 0; b < KX; ++b) {
          if (ixs[b] == 0) r0 += wr[a][b];
          else if (ixs[b] == 1) r1 += wr[a][b];
        }
        if (hi) { acc[1][0] += r0; acc[1][1] += r1; }
        else { acc[0][0] += r0; acc[0][1] += r1; }
      }
#pragma unroll
      for (int iy = 0; iy < 2; ++iy)
#pragma unroll
        for (int ix = 0; ix < 2; ++ix) {
          const int sy = sy0 + iy, sx = sx0 + ix;
          if (sy < g.h && sx < g.w) {
            const int64_t row = ((int64_t)sy * g.w + sx) * g.ic + c, col = ((int64_t)oy * W + ox) * g.oc + o;
            out[TRANSPOSED ? col * rows + row : row * cols + col] = acc[iy][ix];
          }
        }
    }
  }
}

}  // namespace rcb

using namespace rcb;

extern "C" int rcb_gemm(const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                        int M, int N, int K, const float* bias, int bias_mod, int act,
                        int trans_a, int accumulate, rcb_stream_t stream) {
  RCB_CHECK_ARG(A && B && C, "rcb_gemm: null operand");
  RCB_CHECK_ARG(lda % 4 == 0 && ldb % 4 == 0, "rcb_gemm: lda/ldb must be multiples of 4 (got %d, %d)", lda, ldb);
  RCB_CHECK_ARG(((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0), "rcb_gemm: A/B must be 16-byte aligned");
  RCB_CHECK_ARG(ldb >= ((N + 3) / 4) * 4, "rcb_gemm: ldb %d too small for N %d (padded to 4)", ldb, N);
  RCB_CHECK_ARG(!bias || bias_mod > 0, "rcb_gemm: bias_mod must be positive");
  cudaStream_t st = (cudaStream_t)stream;
  PlainC ep{C, ldc, M, bias, bias_mod, act, accumulate};
  if (trans_a) {
    TransA al{A, lda, M};
    return launch_engine("rcb_gemm(T)", al, B, ldb, 0, ep, M, N, K, 1, st);
  }
  PlainA al{A, lda, M};
  return launch_engine("rcb_gemm", al, B, ldb, 0, ep, M, N, K, 1, st);
}

extern "C" int rcb_fold_poly(const float* w, const rcb_upconv_geom* g, float* w_eff, float* w_eff_t,
                             rcb_stream_t stream) {
  PolyGeom pg;
  if (int rc = make_geom(g, &pg)) return rc;
  RCB_CHECK_ARG(w && w_eff, "rcb_fold_poly: null pointer");
  int64_t total = (int64_t)pg.phases() * pg.taps() * pg.ic * pg.oc;
  int blocks = (int)((total + 255) / 256); if (blocks > 8192) blocks = 8192;
  fold_poly_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w, pg, g->kz, g->ky, g->kx, w_eff, w_eff_t, nullptr);
  RCB_CHECK_LAUNCH("rcb_fold_poly");
  return 0;
}

extern "C" int rcb_fold_poly_k(const float* w, const rcb_upconv_geom* g, float* w_eff_k, rcb_stream_t stream) {
  PolyGeom pg;
  if (int rc = make_geom(g, &pg)) return rc;
  RCB_CHECK_ARG(w && w_eff_k, "rcb_fold_poly_k: null pointer");
  int64_t total = (int64_t)pg.phases() * pg.taps() * pg.ic * pg.oc;
  int blocks = (int)((total + 255) / 256); if (blocks > 8192) blocks = 8192;
  fold_poly_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w, pg, g->kz, g->ky, g->kx, nullptr, nullptr, w_eff_k);
  RCB_CHECK_LAUNCH("rcb_fold_poly_k");
  return 0;
}

extern "C" int rcb_fold_dense(const float* w, const rcb_upconv_geom* g, float* m, float* m_t,
                              rcb_stream_t stream) {
  PolyGeom pg;
  if (int rc = make_geom(g, &pg)) return rc;
  RCB_CHECK_ARG(w && m, "rcb_fold_dense: null pointer");
  RCB_CHECK_ARG(pg.d == 1 && pg.fz == 1 && g->kz == 1, "rcb_fold_dense: 1-D / 2-D grids only");
  RCB_CHECK_ARG((g->ky == 5 || g->ky == 1) && g->kx == 5 && g->ky - 1 <= pg.fy && g->kx - 1 <= pg.fx,
                "rcb_fold_dense: built for the 5-tap first stage (1 x 5 or 5 x 5) with factor >= 4");
  const int64_t threads = (int64_t)pg.oc * pg.ic * pg.h * pg.fy * pg.w * pg.fx;
  const int blocks = (int)((threads + 127) / 128);
  cudaStream_t st = (cudaStream_t)stream;
  if (m_t) {
    // the transposed matrix is the cheap one to fold (consecutive threads read neighbouring weights: 15 us against 51 us
    // for the 512 x 4096 cifar fold); m is then an exact transposed copy of it
    if (g->ky == 5) fold_dense_kernel<5, 5, true><<<blocks, 128, 0, st>>>(w, pg, m_t);
    else fold_dense_kernel<1, 5, true><<<blocks, 128, 0, st>>>(w, pg, m_t);
    RCB_CHECK_LAUNCH("rcb_fold_dense");
    const int64_t rows = (int64_t)pg.h * pg.w * pg.ic, cols = (int64_t)pg.h * pg.fy * pg.w * pg.fx * pg.oc;
    return rcb_transpose(m_t, rows, m, cols, (int)cols, (int)rows, stream);
  }
  if (g->ky == 5) fold_dense_kernel<5, 5, false><<<blocks, 128, 0, st>>>(w, pg, m);
  else fold_dense_kernel<1, 5, false><<<blocks, 128, 0, st>>>(w, pg, m);
  RCB_CHECK_LAUNCH("rcb_fold_dense");
  return 0;
}

extern "C" int rcb_upconv_fwd(const float* src, const float* w_eff, const float* bias, float* out,
                              const rcb_upconv_geom* g, int items, int act, rcb_stream_t stream) {
  PolyGeom pg;
  if (int rc = make_geom(g, &pg)) return rc;
  RCB_CHECK_ARG(src && w_eff && bias && out, "rcb_upconv_fwd: null pointer");
  int M = items * pg.vol(), K = pg.taps() * pg.ic, N = pg.oc, Z = pg.phases();
  RCB_CHECK_ARG((int64_t)items * pg.vol() < (1LL << 31), "rcb_upconv_fwd: too many rows");
  RCB_CHECK_ARG(Z <= 65535, "rcb_upconv_fwd: too many phases");
  ConvFwdA al{src, pg, M};
  ConvFwdC ep{out, pg, M, bias, act};
  return launch_engine("rcb_upconv_fwd", al, w_eff, N, (int64_t)K * N, ep, M, N, K, Z, (cudaStream_t)stream);
}

extern "C" int rcb_upconv_bwd(const float* d_out, const float* w_eff_t, const float* src_act, float* d_src,
                              const rcb_upconv_geom* g, int items, rcb_stream_t stream) {
  PolyGeom pg;
  if (int rc = make_geom(g, &pg)) return rc;
  RCB_CHECK_ARG(d_out && w_eff_t && d_src, "rcb_upconv_bwd: null pointer");
  int M = items * pg.vol(), K = pg.phases() * pg.taps() * pg.oc, N = pg.ic;
  RCB_CHECK_ARG((int64_t)items * pg.vol() < (1LL << 31), "rcb_upconv_bwd: too many rows");
  ConvBwdA al{d_out, pg, M};
  ConvBwdC ep{d_src, src_act, pg.ic, M};
  return launch_engine("rcb_upconv_bwd", al, w_eff_t, N, 0, ep, M, N, K, 1, (cudaStream_t)stream);
}
