// kind::f16 MMA with MN-major A and B (pixel-major rows): D[64][32] = A^T B, A = [X (32 features) | ones block] via LBO.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include "../../recombiner_b200/csrc/tc_common.cuh"
using namespace rcb;
namespace rcb { void set_error(const char*, ...) {} }

__device__ __forceinline__ void umma_f16(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint64_t desc_mn64(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;      // SWIZZLE_64B
  return d;
}
// byte offset of half (k, mn < 32) in a [K][64 B] pixel-major buffer with the 64-byte swizzle
__device__ __host__ inline int off64(int k, int mn) { return k * 64 + ((((mn >> 3) ^ ((k >> 1) & 3))) << 4) + (mn & 7) * 2; }

__global__ void kern(int swap, float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  uint8_t* X = smem;                 // 128 x 64 B
  uint8_t* Z = smem + 8192;          // 128 x 64 B
  uint8_t* O = smem + 16384;         // ones block 128 x 64 B
  for (int e = threadIdx.x; e < 128 * 32; e += blockDim.x) {
    const int k = e / 32, m = e % 32;
    *(__half*)(X + off64(k, m)) = __float2half((float)((k * 7 + m * 3) % 5 - 2));
    *(__half*)(Z + off64(k, m)) = __float2half((float)((k * 3 + m) % 7 - 3));
    *(__half*)(O + off64(k, m)) = __float2half(m == 0 ? 1.f : 0.f);
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) tmem_alloc(&slot, 32);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = slot;
  if (threadIdx.x < 32) {
    if (elect_one()) {
      const uint32_t lbo = 16384, sbo = 512;
      // instruction descriptor: D = f32, A = B = f16, both MN-major, N = 32, M = 64
      const uint32_t id = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);
      for (int kk = 0; kk < 8; ++kk) {
        const uint64_t da = swap ? desc_mn64(smem_u32(X) + kk * 1024, sbo, lbo) : desc_mn64(smem_u32(X) + kk * 1024, lbo, sbo);
        const uint64_t db = swap ? desc_mn64(smem_u32(Z) + kk * 1024, sbo, lbo) : desc_mn64(smem_u32(Z) + kk * 1024, lbo, sbo);
        umma_f16(tb, da, db, id, kk ? 1u : 0u);
      }
      umma_commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t v[16];
  for (int c = 0; c < 2; ++c) {
    tmem_ld16(tb + ((uint32_t)(warp * 32) << 16) + c * 16, v);
    for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * 32 + c * 16 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tb, 32);
}

int main(int argc, char** argv) {
  float* d; cudaMalloc(&d, 128 * 32 * 4);
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int swap = 0; swap < 2; ++swap) {
    cudaMemset(d, 0, 128 * 32 * 4);
    kern<<<1, 128, 32 * 1024>>>(swap, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("swap=%d: %s\n", swap, cudaGetErrorString(e)); return 1; }
    static float h[128 * 32];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    // expected: row i (TMEM lane (i % 16) + 32 * (i / 16)): i < 32: sum_k X[k][i] Z[k][n]; i == 32: sum_k Z[k][n]; else 0
    int bad = 0; double maxerr = 0;
    for (int i = 0; i < 64; ++i) {
      const int lane = (i % 16) + 32 * (i / 16);
      for (int n = 0; n < 32; ++n) {
        double ref = 0;
        for (int k = 0; k < 128; ++k) {
          const double z = (double)((k * 3 + n) % 7 - 3);
          const double x = i < 32 ? (double)((k * 7 + i * 3) % 5 - 2) : (i == 32 ? 1.0 : 0.0);
          ref += x * z;
        }
        const double err = fabs(h[lane * 32 + n] - ref);
        if (err > maxerr) maxerr = err;
        if (err > 1e-3 && bad < 6) { printf("  swap=%d row %d col %d: got %g want %g\n", swap, i, n, h[lane * 32 + n], ref); }
        bad += err > 1e-3;
      }
    }
    printf("swap=%d: %d mismatches, max err %g\n", swap, bad, maxerr);
  }
  return 0;
}
