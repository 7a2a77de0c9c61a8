// Fused per-item SIREN coordinate MLP: forward, squared-error and backward in one
// pass per 128-pixel tile, with all activations resident in shared memory.
// FP32 SIMT (exact-parity path).  One CTA per (row, MC sample) item.
//
// Reference semantics: test_model.py:347-355 (4 layers, x <- sin(w0 (xW+b)),
// last layer linear), loss test_model.py:624-627, weight layout
// test_model.py:269-280 (per layer: `out` biases, then W row-major (in,out)).
#include <type_traits>

#include "common.cuh"

namespace rcb {

constexpr int MLP_THREADS = 256;
constexpr int TILE = 128;   // pixels per tile
constexpr int LDP = 132;    // row stride of the [feature][pixel] tiles (conflict-free float4 rows)
constexpr int HID = 32;
constexpr int NPE = 16;     // positional-encoding channels appended to the Fourier features

template <int IN0, int OUT>
struct MlpLayout {
  static constexpr int F = IN0 - NPE;
  static constexpr int off0 = 0;
  static constexpr int off1 = HID * (IN0 + 1);
  static constexpr int off2 = off1 + HID * (HID + 1);
  static constexpr int off3 = off2 + HID * (HID + 1);
  static constexpr int n_w = off3 + OUT * (HID + 1);
  static constexpr int n_w_pad = (n_w + 3) / 4 * 4;
  static constexpr int X0_ROWS = (IN0 + 3) / 4 * 4;
  // shared-memory carve-up (floats)
  static constexpr int s_w = 0;
  static constexpr int s_wt1 = s_w + n_w_pad;
  static constexpr int s_wt2 = s_wt1 + HID * HID;
  static constexpr int s_wt0 = s_wt2 + HID * HID;          // [32][16]
  static constexpr int s_x0 = s_wt0 + HID * NPE;
  static constexpr int s_x1 = s_x0 + X0_ROWS * LDP;
  static constexpr int s_x2 = s_x1 + HID * LDP;
  static constexpr int s_x3 = s_x2 + HID * LDP;
  static constexpr int s_c0 = s_x3 + HID * LDP;
  static constexpr int s_c1 = s_c0 + HID * LDP;
  static constexpr int s_c2 = s_c1 + HID * LDP;
  static constexpr int s_dz = s_c2 + HID * LDP;
  static constexpr int s_dy = s_dz + HID * LDP;            // [4][LDP]
  static constexpr int s_red = s_dy + 4 * LDP;             // [8] block reduction
  static constexpr int total = s_red + 8;
};

// acc[p][j] = bias[j0+j] + sum_i X[i][px0+p] * W[i][j0+j]
template <int NIN>
__device__ __forceinline__ void tile_gemm(const float* __restrict__ X, const float* __restrict__ W, int ldw,
                                          int px0, int j0, float (&acc)[4][4]) {
#pragma unroll 8
  for (int i = 0; i < NIN; ++i) {
    float4 xa = *reinterpret_cast<const float4*>(X + i * LDP + px0);
    float4 wb = *reinterpret_cast<const float4*>(W + i * ldw + j0);
    float xs[4] = {xa.x, xa.y, xa.z, xa.w};
    float ws[4] = {wb.x, wb.y, wb.z, wb.w};
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[p][j] = fmaf(xs[p], ws[j], acc[p][j]);
  }
}

template <int IN0, int OUT, int MODE>
__global__ void __launch_bounds__(MLP_THREADS, 1) mlp_kernel(rcb_mlp_args a) {
  using L = MlpLayout<IN0, OUT>;
  extern __shared__ __align__(16) float smem[];
  float* Ws = smem + L::s_w;
  float* WT1 = smem + L::s_wt1;
  float* WT2 = smem + L::s_wt2;
  float* WT0 = smem + L::s_wt0;
  float* X0 = smem + L::s_x0;
  float* X1 = smem + L::s_x1;
  float* X2 = smem + L::s_x2;
  float* X3 = smem + L::s_x3;
  float* C0 = smem + L::s_c0;
  float* C1 = smem + L::s_c1;
  float* C2 = smem + L::s_c2;
  float* DZ = smem + L::s_dz;
  float* DY = smem + L::s_dy;
  float* RED = smem + L::s_red;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int item = blockIdx.x;
  const int row = item / a.S;
  const int pix = a.pix;
  const float w0 = a.w0;

  // ---- stage this item's weights (and the transposes the backward needs)
  const float* wt_g = a.wt + (int64_t)item * a.ld_w;
  for (int i = tid; i < L::n_w_pad; i += MLP_THREADS) Ws[i] = (i < L::n_w) ? wt_g[i] : 0.f;
  __syncthreads();
  if (MODE != 0) {
    for (int e = tid; e < HID * HID; e += MLP_THREADS) {
      int j = e / HID, i = e % HID;
      WT1[e] = Ws[L::off1 + HID + i * HID + j];
      WT2[e] = Ws[L::off2 + HID + i * HID + j];
    }
    for (int e = tid; e < HID * NPE; e += MLP_THREADS) {
      int j = e / NPE, c = e % NPE;
      WT0[e] = Ws[L::off0 + HID + (L::F + c) * HID + j];
    }
  }
  const float* W0 = Ws + L::off0 + HID;
  const float* W1 = Ws + L::off1 + HID;
  const float* W2 = Ws + L::off2 + HID;
  const float* W3 = Ws + L::off3 + OUT;
  const float* B0 = Ws + L::off0;
  const float* B1 = Ws + L::off1;
  const float* B2 = Ws + L::off2;
  const float* B3 = Ws + L::off3;

  const float* xt = a.xt + (int64_t)row * a.x_row_stride;
  // positional encodings: (items, pix, 16), or a window of the stitched per-datum grid
  const bool stitched = a.pe_base != nullptr;
  const int64_t pe_origin = stitched ? a.pe_base[item] : (int64_t)item * pix;
  const float* pe = a.pe + pe_origin * NPE;
  const int php = a.ph * a.pw;
  auto pe_off = [&](int gp) -> int64_t {      // pixel gp of this patch -> pixel offset from its origin
    if (!stitched) return gp;
    int z = gp / php, rem = gp - z * php;
    int yy = rem / a.pw, xx = rem - yy * a.pw;
    return (int64_t)z * a.pitch_z + (int64_t)yy * a.pitch_y + xx;
  };

  // gradient accumulators (persist over tiles)
  float gW[3][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  float gW0x[4] = {0.f, 0.f, 0.f, 0.f};   // rows 32.. of layer 0 when IN0 > 32
  float gW3 = 0.f, gB[3] = {0.f, 0.f, 0.f}, gB3 = 0.f, sq = 0.f;

  const int px0 = lane * 4;      // tile-GEMM mapping: lane -> 4 pixels, warp -> 4 features
  const int j0 = warp * 4;
  const int ip = tid / 16, jp = tid % 16;   // dW mapping: rows {ip, ip+16}, cols {jp, jp+16}

  for (int pix0 = 0; pix0 < pix; pix0 += TILE) {
    __syncthreads();   // previous tile fully consumed
    // ---- load X0 = [fourier | pe] transposed to [feature][pixel]
    for (int c = tid; c < L::F * (TILE / 4); c += MLP_THREADS) {
      int i = c / (TILE / 4), p4 = c % (TILE / 4);
      int p = pix0 + p4 * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p < pix) v = __ldg(reinterpret_cast<const float4*>(xt + (int64_t)i * pix + p));
      *reinterpret_cast<float4*>(X0 + i * LDP + p4 * 4) = v;
    }
    for (int c = tid; c < TILE * (NPE / 4); c += MLP_THREADS) {
      int p = c / (NPE / 4), c4 = c % (NPE / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (pix0 + p < pix) v = __ldg(reinterpret_cast<const float4*>(pe + pe_off(pix0 + p) * NPE + c4 * 4));
      X0[(L::F + c4 * 4 + 0) * LDP + p] = v.x;
      X0[(L::F + c4 * 4 + 1) * LDP + p] = v.y;
      X0[(L::F + c4 * 4 + 2) * LDP + p] = v.z;
      X0[(L::F + c4 * 4 + 3) * LDP + p] = v.w;
    }
    __syncthreads();

    // ---- forward, three sine layers
    auto sine_layer = [&](auto nin_tag, const float* Xin, const float* Wl, const float* Bl, float* Xout, float* Cout) {
      constexpr int NIN = decltype(nin_tag)::value;
      float acc[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float b = Bl[j0 + j];
#pragma unroll
        for (int p = 0; p < 4; ++p) acc[p][j] = b;
      }
      tile_gemm<NIN>(Xin, Wl, HID, px0, j0, acc);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float s[4], c[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          sincosf(w0 * acc[p][j], &s[p], &c[p]);
          c[p] *= w0;
        }
        *reinterpret_cast<float4*>(Xout + (j0 + j) * LDP + px0) = make_float4(s[0], s[1], s[2], s[3]);
        if (MODE != 0) *reinterpret_cast<float4*>(Cout + (j0 + j) * LDP + px0) = make_float4(c[0], c[1], c[2], c[3]);
      }
    };
    sine_layer(std::integral_constant<int, IN0>{}, X0, W0, B0, X1, C0);
    __syncthreads();
    sine_layer(std::integral_constant<int, HID>{}, X1, W1, B1, X2, C1);
    __syncthreads();
    sine_layer(std::integral_constant<int, HID>{}, X2, W2, B2, X3, C2);
    __syncthreads();

    // ---- last (linear) layer + loss / dy, one pixel per thread
    if (tid < TILE) {
      const int p = tid, gp = pix0 + p;
      const bool valid = gp < pix;
      float o[OUT];
#pragma unroll
      for (int k = 0; k < OUT; ++k) o[k] = B3[k];
#pragma unroll 8
      for (int i = 0; i < HID; ++i) {
        float x = X3[i * LDP + p];
#pragma unroll
        for (int k = 0; k < OUT; ++k) o[k] = fmaf(x, W3[i * OUT + k], o[k]);
      }
      if (MODE == 0) {
        if (valid) {
#pragma unroll
          for (int k = 0; k < OUT; ++k) a.y_pred[((int64_t)item * pix + gp) * OUT + k] = o[k];
        }
      } else if (MODE == 1) {
#pragma unroll
        for (int k = 0; k < OUT; ++k) {
          float r = valid ? o[k] - __ldg(a.y + ((int64_t)row * pix + gp) * OUT + k) : 0.f;
          sq = fmaf(r, r, sq);
          DY[k * LDP + p] = a.coef * r;
        }
      } else {
#pragma unroll
        for (int k = 0; k < OUT; ++k)
          DY[k * LDP + p] = valid ? __ldg(a.dy + ((int64_t)item * pix + gp) * OUT + k) : 0.f;
      }
    }
    if (MODE == 0) continue;
    __syncthreads();

    // ---- backward through the last layer
    if (tid < HID * OUT) {
      const int i = tid % HID, k = tid / HID;
      float s = 0.f;
#pragma unroll 8
      for (int p = 0; p < TILE; p += 4) {
        float4 x = *reinterpret_cast<const float4*>(X3 + i * LDP + p);
        float4 d = *reinterpret_cast<const float4*>(DY + k * LDP + p);
        s = fmaf(x.x, d.x, s); s = fmaf(x.y, d.y, s); s = fmaf(x.z, d.z, s); s = fmaf(x.w, d.w, s);
      }
      gW3 += s;
    } else if (tid >= 128 && tid < 128 + OUT) {
      const int k = tid - 128;
      float s = 0.f;
      for (int p = 0; p < TILE; p += 4) {
        float4 d = *reinterpret_cast<const float4*>(DY + k * LDP + p);
        s += (d.x + d.y) + (d.z + d.w);
      }
      gB3 += s;
    }
    {
      // dz2[i][p] = (sum_k dy[k][p] W3[i][k]) * C2[i][p]
      float dyv[OUT][4];
#pragma unroll
      for (int k = 0; k < OUT; ++k) {
        float4 d = *reinterpret_cast<const float4*>(DY + k * LDP + px0);
        dyv[k][0] = d.x; dyv[k][1] = d.y; dyv[k][2] = d.z; dyv[k][3] = d.w;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = j0 + j;
        float4 c = *reinterpret_cast<const float4*>(C2 + i * LDP + px0);
        float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < OUT; ++k) {
          float w = W3[i * OUT + k];
#pragma unroll
          for (int p = 0; p < 4; ++p) v[p] = fmaf(dyv[k][p], w, v[p]);
        }
        *reinterpret_cast<float4*>(DZ + i * LDP + px0) = make_float4(v[0] * c.x, v[1] * c.y, v[2] * c.z, v[3] * c.w);
      }
    }

    // ---- backward through the sine layers l = 2, 1, 0 (DZ holds dz_l)
    auto back_layer = [&](auto l_tag, const float* Xl, const float* WTl, const float* Cprev) {
      constexpr int LIDX = decltype(l_tag)::value;
      __syncthreads();
      // weight gradient: 2x2 tile per thread over the 128 pixels of this tile
      {
        float s00 = 0.f, s01 = 0.f, s10 = 0.f, s11 = 0.f;
#pragma unroll 4
        for (int p = 0; p < TILE; p += 4) {
          float4 xa = *reinterpret_cast<const float4*>(Xl + ip * LDP + p);
          float4 xb = *reinterpret_cast<const float4*>(Xl + (ip + 16) * LDP + p);
          float4 da = *reinterpret_cast<const float4*>(DZ + jp * LDP + p);
          float4 db = *reinterpret_cast<const float4*>(DZ + (jp + 16) * LDP + p);
          s00 = fmaf(xa.x, da.x, s00); s00 = fmaf(xa.y, da.y, s00); s00 = fmaf(xa.z, da.z, s00); s00 = fmaf(xa.w, da.w, s00);
          s01 = fmaf(xa.x, db.x, s01); s01 = fmaf(xa.y, db.y, s01); s01 = fmaf(xa.z, db.z, s01); s01 = fmaf(xa.w, db.w, s01);
          s10 = fmaf(xb.x, da.x, s10); s10 = fmaf(xb.y, da.y, s10); s10 = fmaf(xb.z, da.z, s10); s10 = fmaf(xb.w, da.w, s10);
          s11 = fmaf(xb.x, db.x, s11); s11 = fmaf(xb.y, db.y, s11); s11 = fmaf(xb.z, db.z, s11); s11 = fmaf(xb.w, db.w, s11);
        }
        gW[LIDX][0] += s00; gW[LIDX][1] += s01; gW[LIDX][2] += s10; gW[LIDX][3] += s11;
      }
      if (LIDX == 0 && IN0 > HID && tid < 16) {
        // extra input rows 32..IN0-1 of the first layer (IN0 = 34: two rows)
        float s[4] = {0.f, 0.f, 0.f, 0.f};
        for (int p = 0; p < TILE; p += 4) {
          float4 da = *reinterpret_cast<const float4*>(DZ + jp * LDP + p);
          float4 db = *reinterpret_cast<const float4*>(DZ + (jp + 16) * LDP + p);
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            if (HID + r < IN0) {
              float4 x = *reinterpret_cast<const float4*>(Xl + (HID + r) * LDP + p);
              s[2 * r] += x.x * da.x + x.y * da.y + x.z * da.z + x.w * da.w;
              s[2 * r + 1] += x.x * db.x + x.y * db.y + x.z * db.z + x.w * db.w;
            }
          }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) gW0x[r] += s[r];
      }
      if (tid < HID) {
        float s = 0.f;
#pragma unroll 8
        for (int p = 0; p < TILE; p += 4) {
          float4 d = *reinterpret_cast<const float4*>(DZ + tid * LDP + p);
          s += (d.x + d.y) + (d.z + d.w);
        }
        gB[LIDX] += s;
      }
      // input gradient
      if (LIDX > 0) {
        float acc[4][4];
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[p][j] = 0.f;
        tile_gemm<HID>(DZ, WTl, HID, px0, j0, acc);   // acc[p][i] = sum_j dz[j][p] W[i][j]
        __syncthreads();                               // everyone done reading DZ
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float4 c = *reinterpret_cast<const float4*>(Cprev + (j0 + j) * LDP + px0);
          *reinterpret_cast<float4*>(DZ + (j0 + j) * LDP + px0) =
              make_float4(acc[0][j] * c.x, acc[1][j] * c.y, acc[2][j] * c.z, acc[3][j] * c.w);
        }
      } else if (warp < NPE / 4) {
        // d pe = dx0 restricted to the positional-encoding inputs
        float acc[4][4];
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[p][j] = 0.f;
        tile_gemm<HID>(DZ, WTl, NPE, px0, j0, acc);
        float* dpe = a.d_pe + pe_origin * NPE;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          int gp = pix0 + px0 + p;
          if (gp < pix)
            *reinterpret_cast<float4*>(dpe + pe_off(gp) * NPE + j0) = make_float4(acc[p][0], acc[p][1], acc[p][2], acc[p][3]);
        }
      }
    };
    back_layer(std::integral_constant<int, 2>{}, X2, WT2, C1);
    back_layer(std::integral_constant<int, 1>{}, X1, WT1, C0);
    back_layer(std::integral_constant<int, 0>{}, X0, WT0, (const float*)nullptr);
  }

  if (MODE == 0) return;

  // ---- write the per-item weight gradients (same layout as the weights)
  float* g = a.d_wt + (int64_t)item * a.ld_w;
  const int offs[3] = {L::off0, L::off1, L::off2};
#pragma unroll
  for (int l = 0; l < 3; ++l) {
    float* gw = g + offs[l] + HID;
    gw[ip * HID + jp] = gW[l][0];
    gw[ip * HID + jp + 16] = gW[l][1];
    gw[(ip + 16) * HID + jp] = gW[l][2];
    gw[(ip + 16) * HID + jp + 16] = gW[l][3];
    if (tid < HID) g[offs[l] + tid] = gB[l];
  }
  if (IN0 > HID && tid < 16) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      if (HID + r < IN0) {
        g[L::off0 + HID + (HID + r) * HID + jp] = gW0x[2 * r];
        g[L::off0 + HID + (HID + r) * HID + jp + 16] = gW0x[2 * r + 1];
      }
    }
  }
  if (tid < HID * OUT) g[L::off3 + OUT + (tid % HID) * OUT + tid / HID] = gW3;
  if (tid >= 128 && tid < 128 + OUT) g[L::off3 + (tid - 128)] = gB3;

  if (MODE == 1) {
    sq = warp_sum(sq);
    if (lane == 0) RED[warp] = sq;
    __syncthreads();
    if (tid == 0) {
      float s = 0.f;
      for (int w = 0; w < MLP_THREADS / 32; ++w) s += RED[w];
      a.sqerr[item] = s;
    }
  }
}

template <int IN0, int OUT>
static int launch_mlp(const rcb_mlp_args* a, cudaStream_t st) {
  using L = MlpLayout<IN0, OUT>;
  size_t smem = sizeof(float) * L::total;
  RCB_CHECK_ARG(a->ld_w >= L::n_w, "rcb_mlp: ld_w %d < weight count %d", a->ld_w, L::n_w);
  RCB_CHECK_ARG(a->n_f == L::F, "rcb_mlp: n_f %d does not match in_dim %d", a->n_f, IN0);
#define RCB_MLP_LAUNCH(MODE)                                                                         \
  do {                                                                                               \
    cudaError_t e = cudaFuncSetAttribute(mlp_kernel<IN0, OUT, MODE>,                                 \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
    if (e != cudaSuccess) { set_error("rcb_mlp: smem opt-in failed: %s", cudaGetErrorString(e)); return -1; } \
    mlp_kernel<IN0, OUT, MODE><<<a->items, MLP_THREADS, smem, st>>>(*a);                             \
  } while (0)
  if (a->mode == 0) RCB_MLP_LAUNCH(0);
  else if (a->mode == 1) RCB_MLP_LAUNCH(1);
  else RCB_MLP_LAUNCH(2);
#undef RCB_MLP_LAUNCH
  RCB_CHECK_LAUNCH("rcb_mlp");
  return 0;
}

}  // namespace rcb

using namespace rcb;

extern "C" int rcb_mlp(const rcb_mlp_args* a, rcb_stream_t stream) {
  RCB_CHECK_ARG(a != nullptr, "rcb_mlp: null args");
  RCB_CHECK_ARG(a->items > 0 && a->S > 0 && a->pix > 0, "rcb_mlp: empty problem");
  RCB_CHECK_ARG(a->pix % 4 == 0, "rcb_mlp: pixel count %d must be a multiple of 4", a->pix);
  RCB_CHECK_ARG(a->mode >= 0 && a->mode <= 2, "rcb_mlp: bad mode %d", a->mode);
  RCB_CHECK_ARG(a->wt && a->xt && a->pe, "rcb_mlp: null input");
  RCB_CHECK_ARG(a->mode != 0 || a->y_pred, "rcb_mlp: mode 0 needs y_pred");
  RCB_CHECK_ARG(a->mode != 1 || (a->y && a->sqerr), "rcb_mlp: mode 1 needs y and sqerr");
  RCB_CHECK_ARG(a->mode != 2 || a->dy, "rcb_mlp: mode 2 needs dy");
  RCB_CHECK_ARG(a->mode == 0 || (a->d_pe && a->d_wt), "rcb_mlp: backward needs d_pe and d_wt");
  RCB_CHECK_ARG(!a->pe_base || (a->ph > 0 && a->pw > 0 && a->pix % (a->ph * a->pw) == 0),
                "rcb_mlp: stitched addressing needs the patch extent (ph, pw) dividing pix");
  cudaStream_t st = (cudaStream_t)stream;
  const int in0 = a->n_f + NPE;
  if (in0 == 32 && a->out == 3) return launch_mlp<32, 3>(a, st);
  if (in0 == 32 && a->out == 1) return launch_mlp<32, 1>(a, st);
  if (in0 == 34 && a->out == 3) return launch_mlp<34, 3>(a, st);
  set_error("rcb_mlp: unsupported INR shape in=%d out=%d (supported: 32->3, 32->1, 34->3 with 3x32 hidden)", in0, a->out);
  return -2;
}
