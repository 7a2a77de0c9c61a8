// Microbenchmark: issue + completion time of small tcgen05 MMAs, dependent vs independent accumulators.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../recombiner_b200/csrc/tc_common.cuh"
using namespace rcb;
namespace rcb { void set_error(const char*, ...) {} }

__device__ __forceinline__ void umma_tf32_ts(uint32_t d, uint32_t a, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint32_t idesc_f16(int n, int m) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }
__device__ __forceinline__ uint32_t idesc_t32(int n, int m) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }

// mode: 0 tf32 SS, 1 tf32 TS, 2 f16 SS ; nd = number of distinct accumulators rotated; n = MMA N; m = MMA M
__global__ void bench(int mode, int nd, int n, int m, int count, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = slot;
  if (threadIdx.x < 32) {
   if (elect_one()) {
    const uint32_t sa = smem_u32(smem), sb = sa + 32 * 1024;
    const uint64_t da = smem_desc_sw128(sa), db = smem_desc_sw128(sb);
    const uint32_t id = mode == 2 ? idesc_f16(n, m) : idesc_t32(n, m);
    uint32_t ph = 0;
    for (int rep = 0; rep < 3; ++rep) {
      long long t0 = clock64();
      if (mode == 0) {
#pragma unroll 8
        for (int i = 0; i < count; ++i) umma_tf32(tb + 256 + (uint32_t)((i & (nd - 1)) * n), da + (uint64_t)((i & 3) * 2), db + (uint64_t)((i & 3) * 2), id, i >= nd ? 1u : 0u);
      } else if (mode == 1) {
#pragma unroll 8
        for (int i = 0; i < count; ++i) umma_tf32_ts(tb + 256 + (uint32_t)((i & (nd - 1)) * n), tb + (uint32_t)((i & 3) * 8), db + (uint64_t)((i & 3) * 2), id, i >= nd ? 1u : 0u);
      } else {
#pragma unroll 8
        for (int i = 0; i < count; ++i) umma_f16(tb + 256 + (uint32_t)((i & (nd - 1)) * n), da + (uint64_t)((i & 3) * 2), db + (uint64_t)((i & 3) * 2), id, i >= nd ? 1u : 0u);
      }
      long long t1 = clock64();
      umma_commit(&bar);
      long long t2 = clock64();
      mbar_wait(&bar, ph); ph ^= 1;
      long long t3 = clock64();
      if (rep == 2) { out[0] = t1 - t0; out[1] = t2 - t1; out[2] = t3 - t2; }
    }
   }
   __syncwarp();
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tb, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const char* names[3] = {"tf32 SS", "tf32 TS", "f16  SS"};
  for (int mode = 0; mode < 3; ++mode)
    for (int m : {128, 64})
      for (int n : {16, 32, 64, 128})
        for (int nd : {1, 2, 4})
          for (int count : {16, 64}) {
            if (nd * n > 256) continue;
            if (m == 64 && mode == 1 && false) continue;
            bench<<<1, 128, 64 * 1024>>>(mode, nd, n, m, count, d);
            long long h[3];
            cudaError_t e = cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) { printf("ERR %s\n", cudaGetErrorString(e)); return 1; }
            printf("%s M=%3d N=%3d accs=%d count=%2d : issue %6lld  commit %4lld  drain %6lld  total/MMA %.1f\n", names[mode], m, n, nd, count,
                   h[0], h[1], h[2], (double)(h[0] + h[1] + h[2]) / count);
          }
  return 0;
}
