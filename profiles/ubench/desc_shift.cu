// Does a K-major SW128 smem descriptor with a row-shifted start address (+ base_offset) and SBO = 2048 read
// the rows (y*16 + x + shift) of a 16-rows-per-line tile?  B = identity, so D[m][n] = A[row(m)][n].
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../recombiner_b200/csrc/tc_common.cuh"
using namespace rcb;
namespace rcb { void set_error(const char*, ...) {} }

__device__ __forceinline__ uint64_t desc_sw128_ex(uint32_t saddr, uint32_t sbo_bytes, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void k(int shift, int use_base_off, float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int ROWS = 20 * 16;                       // 20 lines of 16 rows
  float* A = (float*)smem;                        // ROWS x 32 floats, swizzled
  float* B = (float*)(smem + 48 * 1024);          // 32 x 32 identity, swizzled
  for (int e = threadIdx.x; e < ROWS * 32; e += blockDim.x) {
    const int r = e / 32, c = e % 32;
    const float v = (c & 1) ? (float)c : (float)r;
    *(float*)((uint8_t*)A + r * 128 + ((((c >> 2) ^ (r & 7)) << 4) | ((c & 3) << 2))) = v;
  }
  for (int e = threadIdx.x; e < 32 * 32; e += blockDim.x) {
    const int r = e / 32, c = e % 32;
    *(float*)((uint8_t*)B + r * 128 + ((((c >> 2) ^ (r & 7)) << 4) | ((c & 3) << 2))) = (r == c) ? 1.f : 0.f;
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) tmem_alloc(&slot, 32);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = slot;
  if (threadIdx.x < 32) {
    if (elect_one()) {
      const uint32_t sa = smem_u32(A) + (uint32_t)shift * 128u;
      const uint64_t da = desc_sw128_ex(sa, 2048, use_base_off ? ((sa >> 7) & 7) : 0);
      const uint64_t db = smem_desc_sw128(smem_u32(B));
      const uint32_t id = idesc_tf32(32);
      for (int kk = 0; kk < 4; ++kk) umma_tf32(tb, da + kk * 2, db + kk * 2, id, kk ? 1u : 0u);
      umma_commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t v[16];
  tmem_ld16(tb + ((uint32_t)(warp * 32) << 16), v);
  for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * 16 + j] = __uint_as_float(v[j]);
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tb, 32);
}

int main() {
  float* d; cudaMalloc(&d, 128 * 16 * 4);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  static float h[128 * 16];
  for (int ubo = 0; ubo < 2; ++ubo)
    for (int shift : {0, 1, 2, 7, 8, 16, 17, 18, 33, 34}) {
      k<<<1, 128, 64 * 1024>>>(shift, ubo, d);
      cudaError_t e = cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) { printf("ERR shift %d: %s\n", shift, cudaGetErrorString(e)); return 1; }
      int bad_row = 0, bad_col = 0;
      for (int m = 0; m < 128; ++m) {
        const int expect = (m / 8) * 16 + (m % 8) + shift;
        if ((int)h[m * 16 + 0] != expect || (int)h[m * 16 + 2] != expect) ++bad_row;
        if ((int)h[m * 16 + 1] != 1 || (int)h[m * 16 + 3] != 3 || (int)h[m * 16 + 15] != 15) ++bad_col;
      }
      printf("base_off=%d shift=%2d : wrong rows %3d  wrong cols %3d   m=0..9 ->", ubo, shift, bad_row, bad_col);
      for (int m = 0; m < 10; ++m) printf(" %d/%d", (int)h[m * 16], (int)h[m * 16 + 1]);
      printf("\n");
    }
  return 0;
}
