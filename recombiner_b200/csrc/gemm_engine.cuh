// FP32 SIMT tile engine: C = gather(A) @ B with pluggable A-row addressing and
// epilogue.  One engine serves the plain shared-operand GEMMs of the linear
// reparameterisation and the polyphase (upsample-folded) convolutions as implicit
// GEMMs, forward and data-gradient.  This is the exact-fp32 parity path; the
// tcgen05 path replaces the core for the tensor-core precision modes.
#pragma once
#include "common.cuh"

namespace rcb {

constexpr int GEMM_BK = 16;
constexpr int GEMM_THREADS = 256;

// ---------------------------------------------------------------- A loaders --
struct PlainA {
  static constexpr bool kTrans = false;
  const float* A;
  int lda;
  int M;
  struct Row { const float* p; };
  __device__ __forceinline__ Row row(int m, int) const {
    Row r; r.p = (m < M) ? A + (int64_t)m * lda : nullptr; return r;
  }
  __device__ __forceinline__ const float* chunk(const Row& r, int k, int) const {
    return r.p ? r.p + k : nullptr;
  }
};

struct TransA {  // A given as [K][M]
  static constexpr bool kTrans = true;
  const float* A;
  int lda;
  int M;
};

// polyphase geometry shared by the conv loaders / epilogues (3-D; 2-D uses d=fz=Tz=1)
struct PolyGeom {
  int d, h, w, fz, fy, fx, pz, py, px, Tz, Ty, Tx, ic, oc;
  __device__ __forceinline__ int base_z(int rz) const { return rz < pz ? -1 : 0; }
  __device__ __forceinline__ int base_y(int ry) const { return ry < py ? -1 : 0; }
  __device__ __forceinline__ int base_x(int rx) const { return rx < px ? -1 : 0; }
  __host__ __device__ __forceinline__ int vol() const { return d * h * w; }
  __host__ __device__ __forceinline__ int phases() const { return fz * fy * fx; }
  __host__ __device__ __forceinline__ int taps() const { return Tz * Ty * Tx; }
  __device__ __forceinline__ void split_row(int rem, int& z, int& y, int& x) const {
    z = rem / (h * w); rem -= z * h * w;
    y = rem / w; x = rem - y * w;
  }
  __device__ __forceinline__ void split_phase(int p, int& rz, int& ry, int& rx) const {
    rz = p / (fy * fx); p -= rz * fy * fx;
    ry = p / fx; rx = p - ry * fx;
  }
  __device__ __forceinline__ void split_tap(int t, int& tz, int& ty, int& tx) const {
    tz = t / (Ty * Tx); t -= tz * Ty * Tx;
    ty = t / Tx; tx = t - ty * Tx;
  }
};

// forward: row m = (item, sz, sy, sx); k = tap*ic + c; z = phase
struct ConvFwdA {
  static constexpr bool kTrans = false;
  const float* src;
  PolyGeom g;
  int M;
  struct Row { const float* base; int sz, sy, sx; };
  __device__ __forceinline__ Row row(int m, int z) const {
    Row r;
    if (m >= M) { r.base = nullptr; r.sz = r.sy = r.sx = 0; return r; }
    int vol = g.vol();
    int item = m / vol, rem = m - item * vol;
    int sz, sy, sx, rz, ry, rx;
    g.split_row(rem, sz, sy, sx);
    g.split_phase(z, rz, ry, rx);
    r.sz = sz + g.base_z(rz);
    r.sy = sy + g.base_y(ry);
    r.sx = sx + g.base_x(rx);
    r.base = src + (int64_t)item * vol * g.ic;
    return r;
  }
  __device__ __forceinline__ const float* chunk(const Row& r, int k, int) const {
    if (!r.base) return nullptr;
    int tap = k / g.ic, c = k - tap * g.ic;
    int tz, ty, tx;
    g.split_tap(tap, tz, ty, tx);
    int zz = r.sz + tz, yy = r.sy + ty, xx = r.sx + tx;
    if (zz < 0 || zz >= g.d || yy < 0 || yy >= g.h || xx < 0 || xx >= g.w) return nullptr;
    return r.base + (((int64_t)zz * g.h + yy) * g.w + xx) * g.ic + c;
  }
};

// data gradient: row m = (item, z, y, x) source pixel; k = (phase*taps + tap)*oc + o
struct ConvBwdA {
  static constexpr bool kTrans = false;
  const float* d_out;
  PolyGeom g;
  int M;
  struct Row { const float* base; int z, y, x; };
  __device__ __forceinline__ Row row(int m, int) const {
    Row r;
    if (m >= M) { r.base = nullptr; r.z = r.y = r.x = 0; return r; }
    int vol = g.vol();
    int item = m / vol, rem = m - item * vol;
    g.split_row(rem, r.z, r.y, r.x);
    r.base = d_out + (int64_t)item * vol * g.phases() * g.oc;
    return r;
  }
  __device__ __forceinline__ const float* chunk(const Row& r, int k, int) const {
    if (!r.base) return nullptr;
    int seg = k / g.oc, o = k - seg * g.oc;
    int tap = seg % g.taps(), ph = seg / g.taps();
    int tz, ty, tx, rz, ry, rx;
    g.split_tap(tap, tz, ty, tx);
    g.split_phase(ph, rz, ry, rx);
    int sz = r.z - g.base_z(rz) - tz, sy = r.y - g.base_y(ry) - ty, sx = r.x - g.base_x(rx) - tx;
    if (sz < 0 || sz >= g.d || sy < 0 || sy >= g.h || sx < 0 || sx >= g.w) return nullptr;
    int oz = sz * g.fz + rz, oy = sy * g.fy + ry, ox = sx * g.fx + rx;
    return r.base + (((int64_t)oz * (g.h * g.fy) + oy) * (g.w * g.fx) + ox) * g.oc + o;
  }
};

// ---------------------------------------------------------------- epilogues --
__device__ __forceinline__ float lrelu(float v) { return v > 0.f ? v : 0.01f * v; }

struct PlainC {
  float* C;
  int ldc, M;
  const float* bias;
  int bias_mod, act, accumulate;
  __device__ __forceinline__ float* row_ptr(int m, int) const { return m < M ? C + (int64_t)m * ldc : nullptr; }
  __device__ __forceinline__ float xform(float v, float* p, int n) const {
    if (accumulate) v += p[n];
    if (bias) v += bias[n % bias_mod];
    if (act) v = lrelu(v);
    return v;
  }
};

struct ConvFwdC {
  float* out;
  PolyGeom g;
  int M;
  const float* bias;
  int act;
  __device__ __forceinline__ float* row_ptr(int m, int z) const {
    if (m >= M) return nullptr;
    int vol = g.vol();
    int item = m / vol, rem = m - item * vol;
    int sz, sy, sx, rz, ry, rx;
    g.split_row(rem, sz, sy, sx);
    g.split_phase(z, rz, ry, rx);
    int oz = sz * g.fz + rz, oy = sy * g.fy + ry, ox = sx * g.fx + rx;
    return out + ((((int64_t)item * g.d * g.fz + oz) * (g.h * g.fy) + oy) * (g.w * g.fx) + ox) * g.oc;
  }
  __device__ __forceinline__ float xform(float v, float*, int n) const {
    v += bias[n];
    return act ? lrelu(v) : v;
  }
};

struct ConvBwdC {
  float* d_src;
  const float* src_act;  // post-activation of the producing stage, or NULL
  int ic, M;
  __device__ __forceinline__ float* row_ptr(int m, int) const { return m < M ? d_src + (int64_t)m * ic : nullptr; }
  __device__ __forceinline__ float xform(float v, float* p, int n) const {
    if (src_act) {
      float a = src_act[(p - d_src) + n];
      v *= (a > 0.f ? 1.f : 0.01f);
    }
    return v;
  }
};

// ------------------------------------------------------------------- engine --
template <class AL, class EP, int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(GEMM_THREADS)
gemm_engine(AL al, const float* __restrict__ Bmat, int ldb, int64_t b_zstride, EP ep, int M, int N, int K) {
  constexpr int BK = GEMM_BK;
  constexpr int NT = GEMM_THREADS;
  static_assert((BM / TM) * (BN / TN) == NT, "tile/thread mismatch");
  static_assert(TM % 4 == 0, "TM must be a multiple of 4");
  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];

  const int tid = threadIdx.x;
  const int z = blockIdx.z;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const float* __restrict__ Bz = Bmat + (int64_t)z * b_zstride;

  // ---- A tile: BM x BK, as 4-float chunks along k (or along m when transposed)
  constexpr int A_CHUNKS = BM * (BK / 4) / NT;
  constexpr int B_CHUNKS = (BK * BN / 4 + NT - 1) / NT;
  float4 a_reg[A_CHUNKS];
  float4 b_reg[B_CHUNKS];

  const int a_row = tid % BM;  // row owned by this thread for every chunk (non-transposed)

  auto load_a = [&](int k0, auto&& rowstate) {
#pragma unroll
    for (int i = 0; i < A_CHUNKS; ++i) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if constexpr (!AL::kTrans) {
        int kc = (tid + i * NT) / BM;
        int k = k0 + kc * 4;
        if (k < K) {
          const float* p = al.chunk(rowstate, k, z);
          if (p) {
            if (k + 4 <= K) {
              v = __ldg(reinterpret_cast<const float4*>(p));
            } else {
              v.x = __ldg(p);
              if (k + 1 < K) v.y = __ldg(p + 1);
              if (k + 2 < K) v.z = __ldg(p + 2);
            }
          }
        }
      } else {
        int c = tid + i * NT;
        int kk = c / (BM / 4), m4 = c % (BM / 4);
        int k = k0 + kk, m = m0 + m4 * 4;
        if (k < K && m < al.M) {
          const float* p = al.A + (int64_t)k * al.lda + m;
          if (m + 4 <= al.M) {
            v = __ldg(reinterpret_cast<const float4*>(p));
          } else {
            v.x = __ldg(p);
            if (m + 1 < al.M) v.y = __ldg(p + 1);
            if (m + 2 < al.M) v.z = __ldg(p + 2);
          }
        }
      }
      a_reg[i] = v;
    }
  };
  auto store_a = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_CHUNKS; ++i) {
      if constexpr (!AL::kTrans) {
        int kc = (tid + i * NT) / BM;
        As[buf][kc * 4 + 0][a_row] = a_reg[i].x;
        As[buf][kc * 4 + 1][a_row] = a_reg[i].y;
        As[buf][kc * 4 + 2][a_row] = a_reg[i].z;
        As[buf][kc * 4 + 3][a_row] = a_reg[i].w;
      } else {
        int c = tid + i * NT;
        int kk = c / (BM / 4), m4 = c % (BM / 4);
        *reinterpret_cast<float4*>(&As[buf][kk][m4 * 4]) = a_reg[i];
      }
    }
  };
  auto load_b = [&](int k0) {
#pragma unroll
    for (int i = 0; i < B_CHUNKS; ++i) {
      int c = tid + i * NT;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < BK * BN / 4) {
        int kk = c / (BN / 4), n4 = c % (BN / 4);
        int k = k0 + kk, n = n0 + n4 * 4;
        if (k < K && n < N) v = __ldg(reinterpret_cast<const float4*>(Bz + (int64_t)k * ldb + n));
      }
      b_reg[i] = v;
    }
  };
  auto store_b = [&](int buf) {
#pragma unroll
    for (int i = 0; i < B_CHUNKS; ++i) {
      int c = tid + i * NT;
      if (c < BK * BN / 4) {
        int kk = c / (BN / 4), n4 = c % (BN / 4);
        *reinterpret_cast<float4*>(&Bs[buf][kk][n4 * 4]) = b_reg[i];
      }
    }
  };

  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  auto compute = [&](int buf) {
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        float4 t = *reinterpret_cast<const float4*>(&As[buf][kk][ty * TM + i]);
        a[i] = t.x; a[i + 1] = t.y; a[i + 2] = t.z; a[i + 3] = t.w;
      }
      if constexpr (TN % 4 == 0) {
#pragma unroll
        for (int j = 0; j < TN; j += 4) {
          float4 t = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * TN + j]);
          b[j] = t.x; b[j + 1] = t.y; b[j + 2] = t.z; b[j + 3] = t.w;
        }
      } else {
        static_assert(TN == 2, "TN must be 2 or a multiple of 4");
        float2 t = *reinterpret_cast<const float2*>(&Bs[buf][kk][tx * TN]);
        b[0] = t.x; b[1] = t.y;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  };

  const int nk = (K + BK - 1) / BK;
  if constexpr (!AL::kTrans) {
    auto rs = al.row(m0 + a_row, z);
    load_a(0, rs);
    load_b(0);
    store_a(0);
    store_b(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
      int cur = kt & 1;
      if (kt + 1 < nk) { load_a((kt + 1) * BK, rs); load_b((kt + 1) * BK); }
      compute(cur);
      if (kt + 1 < nk) { store_a(cur ^ 1); store_b(cur ^ 1); }
      __syncthreads();
    }
  } else {
    int dummy = 0;
    load_a(0, dummy);
    load_b(0);
    store_a(0);
    store_b(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
      int cur = kt & 1;
      if (kt + 1 < nk) { load_a((kt + 1) * BK, dummy); load_b((kt + 1) * BK); }
      compute(cur);
      if (kt + 1 < nk) { store_a(cur ^ 1); store_b(cur ^ 1); }
      __syncthreads();
    }
  }

  // ---- epilogue
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int m = m0 + ty * TM + i;
    float* p = ep.row_ptr(m, z);
    if (!p) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = n0 + tx * TN + j;
      if (n < N) p[n] = ep.xform(acc[i][j], p, n);
    }
  }
}

template <class AL, class EP>
static int launch_engine(const char* name, AL al, const float* B, int ldb, int64_t b_zstride, EP ep,
                         int M, int N, int K, int Z, cudaStream_t st) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if (N <= 16) {
    dim3 grid(ceil_div(M, 256), ceil_div(N, 16), Z);
    gemm_engine<AL, EP, 256, 16, 8, 2><<<grid, GEMM_THREADS, 0, st>>>(al, B, ldb, b_zstride, ep, M, N, K);
  } else if (N <= 64 || (int64_t)ceil_div(M, 128) * ceil_div(N, 128) < 296) {
    dim3 grid(ceil_div(M, 128), ceil_div(N, 64), Z);
    gemm_engine<AL, EP, 128, 64, 8, 4><<<grid, GEMM_THREADS, 0, st>>>(al, B, ldb, b_zstride, ep, M, N, K);
  } else {
    dim3 grid(ceil_div(M, 128), ceil_div(N, 128), Z);
    gemm_engine<AL, EP, 128, 128, 8, 8><<<grid, GEMM_THREADS, 0, st>>>(al, B, ldb, b_zstride, ep, M, N, K);
  }
  RCB_CHECK_LAUNCH(name);
  return 0;
}

}  // namespace rcb
