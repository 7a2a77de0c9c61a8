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

// One standard normal per (seed, step, tensor_id, element): Philox block on the
// element counter, Box-Muller on the first two words.  The same function is
// called by the sampling kernel and by the gradient kernel, so eps is never stored.
__device__ __forceinline__ float philox_normal(int64_t seed, int step, int tensor_id, uint64_t elem) {
  uint32_t c[4] = {(uint32_t)elem, (uint32_t)(elem >> 32), (uint32_t)tensor_id, (uint32_t)step};
  Philox::block(c, (uint32_t)seed, (uint32_t)((uint64_t)seed >> 32));
  // u1 in (0,1], u2 in [0,1)
  float u1 = ((float)(c[0] >> 8) + 1.0f) * (1.0f / 16777216.0f);
  float u2 = (float)(c[1] >> 8) * (1.0f / 16777216.0f);
  float r = sqrtf(-2.0f * logf(u1));
  float sn, cs;
  sincospif(2.0f * u2, &sn, &cs);
  return r * cs;
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace rcb
