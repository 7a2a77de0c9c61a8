// Weight gradients of the upsampler (prior training): polyphase conv wgrad as a
// split-K "TN" GEMM with gathered rows, the adjoints of the two weight folds, and a
// column-sum for bias gradients.  Reference: autograd backward of prior_model.py:47-59.
#include "gemm_engine.cuh"

namespace rcb {

constexpr int WG_BK = 16;
constexpr int WG_ROWS_PER_CHUNK = 4096;

// d_w_eff[z][tap][ic][oc] += sum_{m in chunk} src[m shifted by tap][ic] * d_out[m at phase z][oc]
// grid: (ic/64, k-chunks, fy*fx*Ty*Tx); block 256; tile 64(ic) x OC_T.
template <int OC_T>
__global__ void __launch_bounds__(256) upconv_wgrad_kernel(const float* __restrict__ src, const float* __restrict__ d_out,
                                                          float* __restrict__ d_w_eff, PolyGeom g, int M) {
  constexpr int BM = 64, BK = WG_BK;
  constexpr int TN = OC_T / 16;        // 16 x 16 thread grid: 4 ic x TN oc per thread
  __shared__ __align__(16) float As[BK][BM];
  __shared__ __align__(16) float Bs[BK][OC_T];
  const int tid = threadIdx.x;
  const int ic0 = blockIdx.x * BM;
  const int seg = blockIdx.z;
  int tz, ty, tx, rz, ry, rx;
  g.split_tap(seg % g.taps(), tz, ty, tx);
  g.split_phase(seg / g.taps(), rz, ry, rx);
  const int dz = g.base_z(rz) + tz, dy = g.base_y(ry) + ty, dx = g.base_x(rx) + tx;
  const int vol = g.vol(), Hout = g.h * g.fy, Wout = g.w * g.fx;
  const int m_begin = blockIdx.y * WG_ROWS_PER_CHUNK;
  const int m_end = min(M, m_begin + WG_ROWS_PER_CHUNK);

  const int cx = tid % 16, cy = tid / 16;      // cx -> oc group, cy -> ic group of 4
  float acc[4][TN];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int m0 = m_begin; m0 < m_end; m0 += BK) {
    // A: 16 rows x 64 ic = 256 float4 chunks -> one per thread
    {
      const int kk = tid / 16, c4 = tid % 16;
      const int m = m0 + kk;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < m_end) {
        int item = m / vol, rem = m - item * vol;
        int sz, sy, sx;
        g.split_row(rem, sz, sy, sx);
        int zz = sz + dz, yy = sy + dy, xx = sx + dx;
        if (zz >= 0 && zz < g.d && yy >= 0 && yy < g.h && xx >= 0 && xx < g.w)
          v = __ldg(reinterpret_cast<const float4*>(src + ((int64_t)item * vol + ((int64_t)zz * g.h + yy) * g.w + xx) * g.ic + ic0 + c4 * 4));
      }
      *reinterpret_cast<float4*>(&As[kk][c4 * 4]) = v;
    }
    // B: 16 rows x OC_T oc
    for (int c = tid; c < BK * OC_T / 4; c += 256) {
      const int kk = c / (OC_T / 4), c4 = c % (OC_T / 4);
      const int m = m0 + kk;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < m_end) {
        int item = m / vol, rem = m - item * vol;
        int sz, sy, sx;
        g.split_row(rem, sz, sy, sx);
        int oz = sz * g.fz + rz, oy = sy * g.fy + ry, ox = sx * g.fx + rx;
        v = __ldg(reinterpret_cast<const float4*>(d_out + ((((int64_t)item * g.d * g.fz + oz) * Hout + oy) * Wout + ox) * g.oc + c4 * 4));
      }
      *reinterpret_cast<float4*>(&Bs[kk][c4 * 4]) = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a4 = *reinterpret_cast<const float4*>(&As[kk][cy * 4]);
      float a[4] = {a4.x, a4.y, a4.z, a4.w};
      float b[TN];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[kk][cx * TN + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* out = d_w_eff + (int64_t)seg * g.ic * g.oc;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j)
      atomicAdd(out + (int64_t)(ic0 + cy * 4 + i) * g.oc + cx * TN + j, acc[i][j]);
}

__device__ __forceinline__ int tap_index(int r, int kk, int p, int f) {
  int num = r + kk - p;
  int off = (num >= 0) ? num / f : -((-num + f - 1) / f);
  return off - (r < p ? -1 : 0);
}

// d_w[o][c][a0][a][b] = sum over phases of d_w_eff[phase][tap(phase, a0, a, b)][c][o]
__global__ void unfold_poly_kernel(const float* __restrict__ d_w_eff, PolyGeom g, int kz, int ky, int kx, float* __restrict__ d_w) {
  int64_t total = (int64_t)g.oc * g.ic * kz * ky * kx;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int b = e % kx; int64_t r = e / kx;
    int a = r % ky; r /= ky;
    int a0 = r % kz; r /= kz;
    int c = r % g.ic; int o = r / g.ic;
    float s = 0.f;
    for (int rz = 0; rz < g.fz; ++rz) {
      int tz = tap_index(rz, a0, g.pz, g.fz);
      for (int ry = 0; ry < g.fy; ++ry) {
        int ty = tap_index(ry, a, g.py, g.fy);
        for (int rx = 0; rx < g.fx; ++rx) {
          int tx = tap_index(rx, b, g.px, g.fx);
          int64_t seg = (((int64_t)rz * g.fy + ry) * g.fx + rx) * g.taps() + ((tz * g.Ty + ty) * g.Tx + tx);
          s += d_w_eff[(seg * g.ic + c) * g.oc + o];
        }
      }
    }
    d_w[e] = s;
  }
}

// d_w[o][c][a][b] = sum over output pixels whose tap (a,b) lands inside the grid of
// d_m[(sy,sx,c)][(oy,ox,o)] with (sy,sx) the source pixel under that tap
// One thread per (tap row a, (oc, ic) pair), oc fastest: d_m's innermost index is oc, so a warp reads 128 contiguous
// bytes per (output pixel, source pixel).  The thread walks the output pixels whose tap row a is inside the grid and
// adds, for each of its KX taps, the entry of the source pixel the tap falls on (terms in (oy, ox) order).
template <int KY, int KX>
__global__ void __launch_bounds__(128) unfold_dense_kernel(const float* __restrict__ d_m, PolyGeom g, float* __restrict__ d_w) {
  const int H = g.h * g.fy, W = g.w * g.fx;
  const int64_t cols = (int64_t)H * W * g.oc;
  const int npairs = g.oc * g.ic;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= npairs * KY) return;
  const int pair = e % npairs, a = e / npairs;
  const int o = pair % g.oc, c = pair / g.oc;
  float acc[KX];
#pragma unroll
  for (int b = 0; b < KX; ++b) acc[b] = 0.f;
  for (int oy = 0; oy < H; ++oy) {
    const int uy = oy + a - g.py;
    if (uy < 0 || uy >= H) continue;
    const int sy = uy / g.fy;
    for (int ox = 0; ox < W; ++ox) {
      const int sx0 = max(ox - g.px, 0) / g.fx;
      const float* base = d_m + (((int64_t)sy * g.w) * g.ic + c) * cols + ((int64_t)oy * W + ox) * g.oc + o;
      const float v0 = base[(int64_t)sx0 * g.ic * cols];
      const float v1 = sx0 + 1 < g.w ? base[(int64_t)(sx0 + 1) * g.ic * cols] : 0.f;
#pragma unroll
      for (int b = 0; b < KX; ++b) {
        const int ux = ox + b - g.px;
        if (ux < 0 || ux >= W) continue;
        acc[b] += (ux / g.fx - sx0) ? v1 : v0;
      }
    }
  }
#pragma unroll
  for (int b = 0; b < KX; ++b) d_w[(((int64_t)o * g.ic + c) * KY + a) * KX + b] = acc[b];
}

// out[c % mod] += sum_r x[r][c]; grid (ceil(cols/256), row chunks)
__global__ void colsum_kernel(const float* __restrict__ x, int64_t rows, int cols, int mod, float* __restrict__ out,
                              int64_t rows_per_block) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  const int64_t r0 = blockIdx.y * rows_per_block;
  const int64_t r1 = min(rows, r0 + rows_per_block);
  float s = 0.f;
  for (int64_t r = r0; r < r1; ++r) s += x[r * cols + c];
  atomicAdd(out + (c % mod), s);
}

// narrow matrices (cols divides 256): walk the flat array, each thread keeps one column
__global__ void colsum_flat_kernel(const float* __restrict__ x, int64_t total, int cols, int mod, float* __restrict__ out,
                                   int64_t elems_per_block) {
  __shared__ float red[256];
  const int64_t e0 = blockIdx.x * elems_per_block;
  const int64_t e1 = min(total, e0 + elems_per_block);
  float s = 0.f;
  for (int64_t e = e0 + threadIdx.x; e < e1; e += 256) s += x[e];
  red[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x < cols) {
    float t = 0.f;
    for (int i = threadIdx.x; i < 256; i += cols) t += red[i];
    atomicAdd(out + (threadIdx.x % mod), t);
  }
}

static int geom_of(const rcb_upconv_geom* g, PolyGeom* pg) {
  RCB_CHECK_ARG(g != nullptr, "null geometry");
  pg->d = g->d; pg->h = g->h; pg->w = g->w; pg->fz = g->fz; pg->fy = g->fy; pg->fx = g->fx;
  pg->pz = (g->kz - 1) / 2; pg->py = (g->ky - 1) / 2; pg->px = (g->kx - 1) / 2;
  pg->Tz = pg->pz == 0 ? 1 : 2; pg->Ty = pg->py == 0 ? 1 : 2; pg->Tx = pg->px == 0 ? 1 : 2;
  pg->ic = g->ic; pg->oc = g->oc;
  return 0;
}

}  // namespace rcb

using namespace rcb;

extern "C" int rcb_upconv_wgrad(const float* src, const float* d_out, float* d_w_eff, const rcb_upconv_geom* g,
                                int items, rcb_stream_t stream) {
  PolyGeom pg;
  if (int rc = geom_of(g, &pg)) return rc;
  RCB_CHECK_ARG(src && d_out && d_w_eff, "rcb_upconv_wgrad: null pointer");
  RCB_CHECK_ARG(pg.ic % 64 == 0 && (pg.oc == 16 || pg.oc == 64), "rcb_upconv_wgrad: unsupported channels %d -> %d", pg.ic, pg.oc);
  cudaStream_t st = (cudaStream_t)stream;
  const int nseg = pg.phases() * pg.taps();
  RCB_CHECK_ARG(nseg <= 65535, "rcb_upconv_wgrad: too many (phase, tap) segments");
  cudaError_t e = cudaMemsetAsync(d_w_eff, 0, sizeof(float) * (size_t)nseg * pg.ic * pg.oc, st);
  if (e != cudaSuccess) { set_error("rcb_upconv_wgrad: memset failed: %s", cudaGetErrorString(e)); return -1; }
  const int M = items * pg.vol();
  if (M <= 0) return 0;
  dim3 grid(pg.ic / 64, ceil_div(M, WG_ROWS_PER_CHUNK), nseg);
  RCB_CHECK_ARG(grid.y <= 65535, "rcb_upconv_wgrad: too many rows");
  if (pg.oc == 64) upconv_wgrad_kernel<64><<<grid, 256, 0, st>>>(src, d_out, d_w_eff, pg, M);
  else upconv_wgrad_kernel<16><<<grid, 256, 0, st>>>(src, d_out, d_w_eff, pg, M);
  RCB_CHECK_LAUNCH("rcb_upconv_wgrad");
  return 0;
}

extern "C" int rcb_unfold_poly(const float* d_w_eff, const rcb_upconv_geom* g, float* d_w, rcb_stream_t stream) {
  PolyGeom pg;
  if (int rc = geom_of(g, &pg)) return rc;
  RCB_CHECK_ARG(d_w_eff && d_w, "rcb_unfold_poly: null pointer");
  int64_t total = (int64_t)pg.oc * pg.ic * g->kz * g->ky * g->kx;
  unfold_poly_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(d_w_eff, pg, g->kz, g->ky, g->kx, d_w);
  RCB_CHECK_LAUNCH("rcb_unfold_poly");
  return 0;
}

extern "C" int rcb_unfold_dense(const float* d_m, const rcb_upconv_geom* g, float* d_w, rcb_stream_t stream) {
  PolyGeom pg;
  if (int rc = geom_of(g, &pg)) return rc;
  RCB_CHECK_ARG(d_m && d_w, "rcb_unfold_dense: null pointer");
  RCB_CHECK_ARG(pg.d == 1 && pg.fz == 1 && g->kz == 1, "rcb_unfold_dense: 1-D / 2-D grids only");
  RCB_CHECK_ARG((g->ky == 5 || g->ky == 1) && g->kx == 5 && g->ky - 1 <= pg.fy && g->kx - 1 <= pg.fx,
                "rcb_unfold_dense: built for the 5-tap first stage (1 x 5 or 5 x 5) with factor >= 4");
  const int blocks = (pg.oc * pg.ic * g->ky + 127) / 128;
  if (g->ky == 5) unfold_dense_kernel<5, 5><<<blocks, 128, 0, (cudaStream_t)stream>>>(d_m, pg, d_w);
  else unfold_dense_kernel<1, 5><<<blocks, 128, 0, (cudaStream_t)stream>>>(d_m, pg, d_w);
  RCB_CHECK_LAUNCH("rcb_unfold_dense");
  return 0;
}

extern "C" int rcb_colsum(const float* x, int64_t rows, int cols, int mod, float* out, rcb_stream_t stream) {
  RCB_CHECK_ARG(x && out && cols > 0 && mod > 0, "rcb_colsum: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * (size_t)mod, st);
  if (e != cudaSuccess) { set_error("rcb_colsum: memset failed: %s", cudaGetErrorString(e)); return -1; }
  if (rows <= 0) return 0;
  if (256 % cols == 0) {
    const int64_t total = rows * cols, epb = 256 * 64;     // multiple of cols: a thread's column is fixed
    colsum_flat_kernel<<<ceil_div(total, epb), 256, 0, st>>>(x, total, cols, mod, out, epb);
    RCB_CHECK_LAUNCH("rcb_colsum");
    return 0;
  }
  const int64_t rpb = 256;
  dim3 grid(ceil_div(cols, 256), ceil_div(rows, rpb));
  RCB_CHECK_ARG(grid.y <= 65535, "rcb_colsum: too many rows");
  colsum_kernel<<<grid, 256, 0, st>>>(x, rows, cols, mod, out, rpb);
  RCB_CHECK_LAUNCH("rcb_colsum");
  return 0;
}
