// kind::f16 MMA with the A operand in TMEM: which packing of the K elements does the hardware expect?
// A[128][K=32] written by tcgen05.st, 16 columns of packed half2 per row; variant 0: column c = (k = 2c, 2c+1).
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include "../../recombiner_b200/csrc/tc_common.cuh"
using namespace rcb;
namespace rcb { void set_error(const char*, ...) {} }

__device__ __forceinline__ void umma_f16_ts(uint32_t d, uint32_t a, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__host__ __device__ inline float aval(int m, int k) { return (float)((m * 3 + k * 5) % 7 - 3); }
__host__ __device__ inline float bval(int n, int k) { return (float)((n * 2 + k * 3) % 5 - 2); }

__global__ void kern(int kstep_cols, float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  // B: K-major [N = 32 rows][K = 32 halves = 64 B], 64-byte swizzle (8-row groups 512 B apart)
  for (int e = threadIdx.x; e < 32 * 32; e += blockDim.x) {
    const int n = e / 32, k = e % 32;
    const int off = n * 64 + ((((k >> 3) ^ ((n >> 1) & 3))) << 4) + (k & 7) * 2;
    *(__half*)(smem + off) = __float2half(bval(n, k));
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) tmem_alloc(&slot, 64);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, m = warp * 32 + lane;
  {
    uint32_t v[16];
    for (int c = 0; c < 16; ++c) {
      const __half2 h = __floats2half2_rn(aval(m, 2 * c), aval(m, 2 * c + 1));
      v[c] = *reinterpret_cast<const uint32_t*>(&h);
    }
    tmem_st16(tb + ((uint32_t)(warp * 32) << 16), v);       // A in columns [0, 16)
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (threadIdx.x < 32) {
    if (elect_one()) {
      const uint32_t id = (1u << 4) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint64_t db = smem_desc_sw64(smem_u32(smem));
      for (int kk = 0; kk < 2; ++kk)      // K = 16 per instruction: 32 B further in B, kstep_cols columns further in A
        umma_f16_ts(tb + 32, tb + (uint32_t)(kk * kstep_cols), db + (uint64_t)(kk * 2), id, kk ? 1u : 0u);
      umma_commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  uint32_t v[16];
  for (int c = 0; c < 2; ++c) {
    tmem_ld16(tb + ((uint32_t)(warp * 32) << 16) + 32 + c * 16, v);
    for (int j = 0; j < 16; ++j) out[m * 32 + c * 16 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tb, 64);
}

int main() {
  float* d; cudaMalloc(&d, 128 * 32 * 4);
  for (int cols = 8; cols <= 16; cols += 8) {
    cudaMemset(d, 0, 128 * 32 * 4);
    kern<<<1, 128, 16 * 1024>>>(cols, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("cols=%d: %s\n", cols, cudaGetErrorString(e)); return 1; }
    static float h[128 * 32];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 32; ++n) {
        double ref = 0;
        for (int k = 0; k < 32; ++k) ref += (double)aval(m, k) * bval(n, k);
        if (fabs(h[m * 32 + n] - ref) > 1e-3) { if (bad < 4) printf("  cols=%d (%d,%d): got %g want %g\n", cols, m, n, h[m * 32 + n], ref); ++bad; }
      }
    printf("k-step of %d columns: %d mismatches\n", cols, bad);
  }
  return 0;
}
