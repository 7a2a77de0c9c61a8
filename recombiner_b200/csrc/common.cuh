// Shared device/host helpers for the RECOMBINER B200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/recombiner_b200.h"

namespace rcb {

void set_error(const char* fmt, ...);

#define RCB_CHECK_ARG(cond, ...)          \
  do {                                    \
    if (!(cond)) {                        \
      rcb::set_error(__VA_ARGS__);        \
      return -2;                          \
    }                                     \
  } while (0)

#define RCB_CHECK_LAUNCH(name)                                            \
  do {                                                                    \
    cudaError_t e__ = cudaGetLastError();                                 \
    if (e__ != cudaSuccess) {                                             \
      rcb::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
      return -1;                                                          \
    }                                                                     \
  } while (0)

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// sigma = softplus(x)/6 with torch's threshold-20 linearisation
// (reference test_model.py:101, prior_model.py:88).
__device__ __forceinline__ float std_transform(float x) {
  float sp = x > 20.f ? x : log1pf(expf(x));
  return sp / 6.f;
}
// d sigma / d raw = sigmoid(x)/6 (1/6 above the threshold)
__device__ __forceinline__ float std_transform_grad(float x) {
  float sg = x > 20.f ? 1.f : 1.f / (1.f + expf(-x));
  return sg / 6.f;
}

// The same two functions on the special-function unit (ex2 / lg2 / rcp approximations, ~2 ulp each).  log1p(t) keeps its
// relative accuracy for small t = e^x through the alternating series below 0.1 (truncation < 2e-6 relative); above it
// lg2(1 + t) is accurate to ~3e-6 relative.  Used by the update kernel when rcb_update_args.fast_math is set.
__device__ __forceinline__ float std_transform_fast(float x) {
  const float t = __expf(x);
  const float series = t * (1.f - t * (0.5f - t * (0.33333334f - t * (0.25f - t * 0.2f))));
  const float sp = x > 20.f ? x : (t < 0.1f ? series : __logf(1.f + t));
  return sp * 0.16666667f;
}
__device__ __forceinline__ float std_transform_grad_fast(float x) {
  const float sg = x > 20.f ? 1.f : __fdividef(1.f, 1.f + __expf(-x));
  return sg * 0.16666667f;
}
__device__ __forceinline__ float sqrt_fast(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// ---- counter-based Philox4x32-10 ------------------------------------------
struct Philox {
  static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  static constexpr uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  __host__ __device__ static inline void round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#ifdef __CUDA_ARCH__
    uint32_t hi0 = __umulhi(M0, c[0]), hi1 = __umulhi(M1, c[2]);
#else
    uint32_t hi0 = (uint32_t)(((uint64_t)M0 * c[0]) >> 32), hi1 = (uint32_t)(((uint64_t)M1 * c[2]) >> 32);
#endif
    uint32_t lo0 = M0 * c[0], lo1 = M1 * c[2];
    uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  }
  __host__ __device__ static inline void block(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      round(c, k0, k1);
      k0 += W0; k1 += W1;
    }
  }
};

// Standard normals keyed by (seed, step, tensor_id, item, index), item = global row * S + sample.
// One Philox block yields two Box-Muller pairs = the normals of indices {c + t, c + 256 + t,
// c + 512 + t, c + 768 + t} of a 1024-index chunk c (block number "quad" = chunk * 256 + t), so a
// 256-thread CTA that owns a chunk reads and writes it fully coalesced.  philox_normal4 and
// philox_normal return the same values: the sampling kernel and the gradient kernel agree
// whether or not the noise is kept in memory.
__device__ __forceinline__ void philox_normal4(int64_t seed, int step, int tensor_id, int64_t item, uint32_t quad,
                                               float (&z)[4]) {
  uint32_t c[4] = {quad, (uint32_t)item, (uint32_t)tensor_id ^ ((uint32_t)((uint64_t)item >> 32) << 8), (uint32_t)step};
  Philox::block(c, (uint32_t)seed, (uint32_t)((uint64_t)seed >> 32));
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    // u1 in (0,1], u2 in [0,1)
    const float u1 = ((float)(c[2 * h] >> 8) + 1.0f) * (1.0f / 16777216.0f);
    const float u2 = (float)(c[2 * h + 1] >> 8) * (1.0f / 16777216.0f);
    // __logf has ~2^-22 absolute error near 1: a slightly positive result must not become sqrt(negative) = NaN
    const float r = sqrt_fast(fmaxf(-2.0f * __logf(u1), 0.0f));     // MUFU: the radius does not need a correctly rounded root
    float sn, cs;
    __sincosf(6.2831853071795865f * u2, &sn, &cs);
    z[2 * h] = r * cs;
    z[2 * h + 1] = r * sn;
  }
}
__device__ __forceinline__ float philox_normal(int64_t seed, int step, int tensor_id, int64_t item, uint32_t idx) {
  float z[4];
  philox_normal4(seed, step, tensor_id, item, ((idx >> 10) << 8) | (idx & 255u), z);
  const uint32_t k = (idx >> 8) & 3u;
  return k == 0 ? z[0] : (k == 1 ? z[1] : (k == 2 ? z[2] : z[3]));
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace rcb
