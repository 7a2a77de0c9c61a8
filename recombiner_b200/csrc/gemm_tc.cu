// tcgen05 / TMEM / TMA GEMM for the shared-operand contractions of the fit path
// (linear reparameterisation fwd/bwd, dense-folded first upsampler stage):
//     C[M,N] = A[M,K] @ Bt[N,K]^T  (+ bias[n % bias_mod], LeakyReLU)
// fp32 in HBM, TF32 tensor-core math (kind::tf32), fp32 accumulation in TMEM.
//
// One CTA per 128 x BN output tile, warp-specialised:
//   warp 0   : TMA producer  (cp.async.bulk.tensor.2d, 128B swizzle, 3-stage mbarrier ring)
//   warp 1   : TMEM allocator + single-thread tcgen05.mma issuer
//   warps 2-5: epilogue (tcgen05.ld 32x32b -> staged through the idle pipeline stages -> bias/activation ->
//              coalesced global stores); two CTAs per SM
// Both operands are K-major (A row-major [M,K]; B passed transposed as [N,K]); ragged M/N/K
// edges are handled by TMA out-of-bounds zero fill plus guarded stores.
#include <cuda_fp16.h>
#include "tc_common.cuh"

namespace rcb {

constexpr int TC_BK = 32;            // 32 tf32 = 128 bytes = one swizzle row

constexpr int GEMM_STAGES = 3;       // 3 x 32 KB: two CTAs per SM, one's epilogue under the other's main loop

template <int BN>
struct TcSmem {
  static constexpr int A_BYTES = TC_BM * TC_BK * 4;
  static constexpr int B_BYTES = BN * TC_BK * 4;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = GEMM_STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + 128 + 1024;   // barriers + slack for 1024-B alignment
  static constexpr int EPI_LD = BN + 4;                // floats per staged accumulator row (conflict-free float4 rows)
  static_assert(4 * 32 * EPI_LD * 4 <= BAR_OFF, "epilogue staging must fit the pipeline stages");
};

template <int BN>
__device__ __forceinline__ void gemm_tc_body(const CUtensorMap* tmA, const CUtensorMap* tmB,
                                             float* __restrict__ C, int ldc, int M, int N, int K,
                                             const float* __restrict__ bias, int bias_mod, int act, int accumulate,
                                             int out_half, int in_half, float out_scale = 1.0f) {
  using S = TcSmem<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + S::BAR_OFF);
  uint64_t* empty = full + GEMM_STAGES;
  uint64_t* tmem_full = empty + GEMM_STAGES;
  uint32_t* tmem_slot = (uint32_t*)(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * TC_BM, n0 = blockIdx.y * BN;
  const int BKE = in_half ? 2 * TC_BK : TC_BK;          // elements per 128-byte operand row
  const int nkb = (K + BKE - 1) / BKE;

  if (threadIdx.x == 0) {
    for (int s = 0; s < GEMM_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    // whole warp converged, one elected lane issues (see elect_one)
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % GEMM_STAGES;
      const uint32_t ph = (kb / GEMM_STAGES) & 1;
      mbar_wait(&empty[s], ph ^ 1);
      if (elect_one()) {
        uint8_t* a_dst = smem + s * S::STAGE_BYTES;
        uint8_t* b_dst = a_dst + S::A_BYTES;
        mbar_expect_tx(&full[s], S::STAGE_BYTES);
        tma_load_2d(tmA, &full[s], a_dst, kb * BKE, m0);
        tma_load_2d(tmB, &full[s], b_dst, kb * BKE, n0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===== MMA issuer (whole warp converged, one elected lane issues) =====
    // instruction descriptor: D=f32, A=B=tf32, both K-major, N=BN, M=128
    const uint32_t idesc = (1u << 4) | (in_half ? 0u : ((2u << 7) | (2u << 10))) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % GEMM_STAGES;
      const uint32_t ph = (kb / GEMM_STAGES) & 1;
      mbar_wait(&full[s], ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        const uint32_t a_addr = smem_u32(smem + s * S::STAGE_BYTES);
        const uint32_t b_addr = a_addr + S::A_BYTES;
        const uint64_t da = smem_desc_sw128(a_addr), db = smem_desc_sw128(b_addr);
#pragma unroll
        for (int k = 0; k < TC_BK / 8; ++k) {
          // advance 8 tf32 = 32 bytes inside the swizzle row: +2 in the (>>4) address field
          if (in_half) {
            asm volatile(
                "{\n\t"
                ".reg .pred p;\n\t"
                "setp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
                "}" ::"r"(tmem_base), "l"(da + (uint64_t)(k * 2)), "l"(db + (uint64_t)(k * 2)), "r"(idesc), "r"((kb | k) ? 1u : 0u) : "memory");
          } else {
            umma_tf32(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
          }
        }
        umma_commit(&empty[s]);          // frees the smem stage once these MMAs retire
        if (kb == nkb - 1) umma_commit(tmem_full);   // accumulator complete
      }
      __syncwarp();
    }
  } else {
    // ===== epilogue: TMEM -> registers -> shared (row per lane) -> global (row per instruction) =====
    // All MMAs and therefore all TMA fills are done when tmem_full fires: the pipeline stages are free.
    mbar_wait(tmem_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    float* wbuf = reinterpret_cast<float*>(smem) + (size_t)q * 32 * S::EPI_LD;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t v[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
            "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
            "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
            "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(wbuf + lane * S::EPI_LD + c0 + j) =
            make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
    }
    __syncwarp();
    // lanes walk a row: BN / 4 float4 chunks per row, 32 / (BN / 4) rows per instruction
    constexpr int LPR = BN / 4;
    constexpr int RPI = 32 / LPR;
    const int sub = lane / LPR, n = n0 + (lane % LPR) * 4;
    // a lane keeps its four columns for every row it writes: bias fetched once
    float bv[4] = {0.f, 0.f, 0.f, 0.f};
    if (bias) {
#pragma unroll
      for (int t = 0; t < 4; ++t) if (n + t < N) bv[t] = bias[(n + t) % bias_mod];
    }
    const float slope = act ? 0.01f : 1.0f;
    // four rows per trip: the shared-memory loads (and the C loads of an accumulating call) are issued together
#pragma unroll 1
    for (int r0 = 0; r0 < 32; r0 += 4 * RPI) {
      float4 f4[4], c4[4];
      int rows[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int rr = r0 + u * RPI + sub;
        rows[u] = (rr < 32 && n < N) ? m0 + q * 32 + rr : M;
        f4[u] = *reinterpret_cast<const float4*>(wbuf + (rr & 31) * S::EPI_LD + (lane % LPR) * 4);
        c4[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (accumulate && rows[u] < M) {
          const float* crow = C + (int64_t)rows[u] * ldc;
          if (n + 4 <= N) c4[u] = *reinterpret_cast<const float4*>(crow + n);
          else { c4[u].x = crow[n]; if (n + 1 < N) c4[u].y = crow[n + 1]; if (n + 2 < N) c4[u].z = crow[n + 2]; }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (rows[u] >= M) continue;
        float o[4] = {f4[u].x * out_scale + c4[u].x + bv[0], f4[u].y * out_scale + c4[u].y + bv[1],
                      f4[u].z * out_scale + c4[u].z + bv[2], f4[u].w * out_scale + c4[u].w + bv[3]};
#pragma unroll
        for (int t = 0; t < 4; ++t) o[t] = o[t] > 0.f ? o[t] : slope * o[t];
        float* crow = C + (int64_t)rows[u] * ldc;
        if (out_half) {                      // C is fp16 (N % 4 == 0, no accumulate: checked by the launcher)
          const __half2 lo = __floats2half2_rn(o[0], o[1]), hi = __floats2half2_rn(o[2], o[3]);
          *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(C) + (int64_t)rows[u] * ldc + n) =
              make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
        } else if (n + 4 <= N) {
          *reinterpret_cast<float4*>(crow + n) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
          for (int t = 0; t < 4 && n + t < N; ++t) crow[n + t] = o[t];
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN) : "memory");
  }
}

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 2)
gemm_tf32_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    float* __restrict__ C, int ldc, int M, int N, int K,
                    const float* __restrict__ bias, int bias_mod, int act, int accumulate, int out_half, int in_half) {
  gemm_tc_body<BN>(&tmA, &tmB, C, ldc, M, N, K, bias, bias_mod, act, accumulate, out_half, in_half);
}

// Several GEMMs that share M and the row strides in one launch (blockIdx.z picks the problem): the four per-layer
// reparameterisation products are 360 CTAs each -- 1.2 waves of the 296 CTA slots -- but 3.7 waves together.
constexpr int GEMM_MAX_BATCH = 4;
struct GemmBatch {
  CUtensorMap a[GEMM_MAX_BATCH], b[GEMM_MAX_BATCH];
  float* C[GEMM_MAX_BATCH];
  int N[GEMM_MAX_BATCH], K[GEMM_MAX_BATCH];
};
template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 2)
gemm_tc_batch_kernel(const __grid_constant__ GemmBatch g, int ldc, int M, int in_half, float out_scale) {
  const int z = blockIdx.z;
  if ((int)blockIdx.y * BN >= g.N[z]) return;
  gemm_tc_body<BN>(&g.a[z], &g.b[z], g.C[z], ldc, M, g.N[z], g.K[z], nullptr, 1, 0, 0, 0, in_half, out_scale);
}

// ---- host side: tensor maps through the driver entry point (no libcuda link) -------------
EncodeTiledFn tc_get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D fp32 tensor [rows, cols] with row stride ld (elements), box = [box_rows, 32 cols], 128B swizzle
static int make_map_2d(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                       int esize = 4) {
  EncodeTiledFn enc = tc_get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return -1; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * (cuuint64_t)esize};
  cuuint32_t box[2] = {(cuuint32_t)(128 / esize), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return -1; }
  return 0;
}

template <int BN>
static int launch_tc(const float* A, int lda, const float* Bt, int ldbt, float* C, int ldc, int M, int N, int K,
                     const float* bias, int bias_mod, int act, int accumulate, cudaStream_t st, int out_half = 0, int in_half = 0) {
  CUtensorMap tmA, tmB;
  const int es = in_half ? 2 : 4;
  if (int rc = make_map_2d(&tmA, A, M, K, lda, TC_BM, es)) return rc;
  if (int rc = make_map_2d(&tmB, Bt, N, K, ldbt, BN, es)) return rc;
  const int smem = TcSmem<BN>::TOTAL;
  static uint64_t configured = 0;       // one bit per device: the opt-in is per device (and per template instance)
  int dev = 0;
  cudaGetDevice(&dev);
  if (!((configured >> (dev & 63)) & 1u)) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tf32_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { set_error("rcb_gemm_tc: smem opt-in failed: %s", cudaGetErrorString(e)); return -1; }
    configured |= 1ull << (dev & 63);
  }
  dim3 grid(ceil_div(M, TC_BM), ceil_div(N, BN));
  gemm_tf32_tc_kernel<BN><<<grid, TC_THREADS, smem, st>>>(tmA, tmB, C, ldc, M, N, K, bias, bias_mod, act, accumulate, out_half, in_half);
  RCB_CHECK_LAUNCH("rcb_gemm_tc");
  return 0;
}

}  // namespace rcb

using namespace rcb;

extern "C" int rcb_gemm_tc(const float* A, int lda, const float* Bt, int ldbt, float* C, int ldc, int M, int N, int K,
                           const float* bias, int bias_mod, int act, int accumulate, rcb_stream_t stream) {
  RCB_CHECK_ARG(A && Bt && C, "rcb_gemm_tc: null operand");
  RCB_CHECK_ARG(M > 0 && N > 0 && K > 0, "rcb_gemm_tc: empty problem");
  RCB_CHECK_ARG(lda % 4 == 0 && ldbt % 4 == 0 && ldc % 4 == 0, "rcb_gemm_tc: leading dimensions must be multiples of 4");
  RCB_CHECK_ARG(((uintptr_t)A % 16 == 0) && ((uintptr_t)Bt % 16 == 0) && ((uintptr_t)C % 16 == 0),
                "rcb_gemm_tc: operands must be 16-byte aligned");
  RCB_CHECK_ARG(!bias || bias_mod > 0, "rcb_gemm_tc: bias_mod must be positive");
  cudaStream_t st = (cudaStream_t)stream;
  // 128-wide tiles unless they would leave most SMs idle (skinny problems of prior training: M = 1024 rows, N = 512):
  // 64-wide tiles double the CTA count for the same K loop
  const int64_t ctas128 = (int64_t)ceil_div(M, TC_BM) * ceil_div(N, 128);
  if (N > 64 && ctas128 >= 96) return launch_tc<128>(A, lda, Bt, ldbt, C, ldc, M, N, K, bias, bias_mod, act, accumulate, st);
  return launch_tc<64>(A, lda, Bt, ldbt, C, ldc, M, N, K, bias, bias_mod, act, accumulate, st);
}

// Same GEMM, C written as fp16 (for a following fp16-operand stage).  N % 4 == 0, no accumulation into C.
extern "C" int rcb_gemm_tc_oh(const float* A, int lda, const float* Bt, int ldbt, void* C_h, int ldc, int M, int N, int K,
                              const float* bias, int bias_mod, int act, rcb_stream_t stream) {
  RCB_CHECK_ARG(A && Bt && C_h, "rcb_gemm_tc_oh: null operand");
  RCB_CHECK_ARG(M > 0 && N > 0 && K > 0 && N % 4 == 0, "rcb_gemm_tc_oh: N must be a positive multiple of 4");
  RCB_CHECK_ARG(lda % 4 == 0 && ldbt % 4 == 0 && ldc % 4 == 0, "rcb_gemm_tc_oh: leading dimensions must be multiples of 4");
  RCB_CHECK_ARG(((uintptr_t)A % 16 == 0) && ((uintptr_t)Bt % 16 == 0) && ((uintptr_t)C_h % 8 == 0),
                "rcb_gemm_tc_oh: operands must be 16-byte aligned (C: 8)");
  RCB_CHECK_ARG(!bias || bias_mod > 0, "rcb_gemm_tc_oh: bias_mod must be positive");
  cudaStream_t st = (cudaStream_t)stream;
  float* C = reinterpret_cast<float*>(C_h);
  if (N > 64) return launch_tc<128>(A, lda, Bt, ldbt, C, ldc, M, N, K, bias, bias_mod, act, 0, st, 1);
  return launch_tc<64>(A, lda, Bt, ldbt, C, ldc, M, N, K, bias, bias_mod, act, 0, st, 1);
}

// fp16 A and Bt (K-major, 64 elements per 128-byte row, kind::f16), fp16 C.
extern "C" int rcb_gemm_tc_hh(const void* A_h, int lda, const void* Bt_h, int ldbt, void* C_h, int ldc, int M, int N, int K,
                              const float* bias, int bias_mod, int act, rcb_stream_t stream) {
  RCB_CHECK_ARG(A_h && Bt_h && C_h, "rcb_gemm_tc_hh: null operand");
  RCB_CHECK_ARG(M > 0 && N > 0 && K > 0 && N % 4 == 0 && K % 8 == 0, "rcb_gemm_tc_hh: N %% 4 == 0 and K %% 8 == 0 required");
  RCB_CHECK_ARG(lda % 8 == 0 && ldbt % 8 == 0 && ldc % 4 == 0, "rcb_gemm_tc_hh: leading dimensions must keep rows 16-byte aligned");
  RCB_CHECK_ARG(((uintptr_t)A_h % 16 == 0) && ((uintptr_t)Bt_h % 16 == 0) && ((uintptr_t)C_h % 8 == 0),
                "rcb_gemm_tc_hh: operands must be 16-byte aligned (C: 8)");
  RCB_CHECK_ARG(!bias || bias_mod > 0, "rcb_gemm_tc_hh: bias_mod must be positive");
  cudaStream_t st = (cudaStream_t)stream;
  const float* A = reinterpret_cast<const float*>(A_h);
  const float* Bt = reinterpret_cast<const float*>(Bt_h);
  float* C = reinterpret_cast<float*>(C_h);
  if (N > 64) return launch_tc<128>(A, lda, Bt, ldbt, C, ldc, M, N, K, bias, bias_mod, act, 0, st, 1, 1);
  return launch_tc<64>(A, lda, Bt, ldbt, C, ldc, M, N, K, bias, bias_mod, act, 0, st, 1, 1);
}

// fp16 A and Bt, fp32 C with the full epilogue of rcb_gemm_tc.
extern "C" int rcb_gemm_tc_h(const void* A_h, int lda, const void* Bt_h, int ldbt, float* C, int ldc, int M, int N, int K,
                             const float* bias, int bias_mod, int act, int accumulate, rcb_stream_t stream) {
  RCB_CHECK_ARG(A_h && Bt_h && C, "rcb_gemm_tc_h: null operand");
  RCB_CHECK_ARG(M > 0 && N > 0 && K > 0 && K % 8 == 0, "rcb_gemm_tc_h: K %% 8 == 0 required");
  RCB_CHECK_ARG(lda % 8 == 0 && ldbt % 8 == 0 && ldc % 4 == 0, "rcb_gemm_tc_h: leading dimensions must keep rows 16-byte aligned");
  RCB_CHECK_ARG(((uintptr_t)A_h % 16 == 0) && ((uintptr_t)Bt_h % 16 == 0) && ((uintptr_t)C % 16 == 0),
                "rcb_gemm_tc_h: operands must be 16-byte aligned");
  RCB_CHECK_ARG(!bias || bias_mod > 0, "rcb_gemm_tc_h: bias_mod must be positive");
  cudaStream_t st = (cudaStream_t)stream;
  const float* A = reinterpret_cast<const float*>(A_h);
  const float* Bt = reinterpret_cast<const float*>(Bt_h);
  if (N > 64) return launch_tc<128>(A, lda, Bt, ldbt, C, ldc, M, N, K, bias, bias_mod, act, accumulate, st, 0, 1);
  return launch_tc<64>(A, lda, Bt, ldbt, C, ldc, M, N, K, bias, bias_mod, act, accumulate, st, 0, 1);
}

// nb <= 4 products C_i[M, N_i] = A_i[M, K_i] @ Bt_i[N_i, K_i]^T in one launch; A_i share the row stride lda, C_i share ldc;
// in_half: A_i and Bt_i are fp16 (K_i % 8 == 0), else fp32 / TF32.  No bias, activation or accumulation.
extern "C" int rcb_gemm_tc_batch(int nb, const void* const* A, int lda, const void* const* Bt, const int* ldbt,
                                 float* const* C, int ldc, int M, const int* N, const int* K, int in_half,
                                 float out_scale, rcb_stream_t stream) {
  RCB_CHECK_ARG(nb >= 1 && nb <= GEMM_MAX_BATCH && A && Bt && ldbt && C && N && K, "rcb_gemm_tc_batch: 1..4 problems");
  const int al = in_half ? 8 : 4;
  RCB_CHECK_ARG(M > 0 && lda % al == 0 && ldc % 4 == 0, "rcb_gemm_tc_batch: bad M or leading dimensions");
  GemmBatch g;
  int n_max = 0;
  for (int i = 0; i < GEMM_MAX_BATCH; ++i) {
    const int j = i < nb ? i : 0;             // unused slots repeat problem 0 (never launched: grid.z = nb)
    RCB_CHECK_ARG(A[j] && Bt[j] && C[j] && N[j] > 0 && K[j] > 0 && ldbt[j] % al == 0, "rcb_gemm_tc_batch: bad problem");
    RCB_CHECK_ARG(!in_half || K[j] % 8 == 0, "rcb_gemm_tc_batch: fp16 operands need K %% 8 == 0");
    RCB_CHECK_ARG(((uintptr_t)A[j] % 16 == 0) && ((uintptr_t)Bt[j] % 16 == 0) && ((uintptr_t)C[j] % 16 == 0),
                  "rcb_gemm_tc_batch: operands must be 16-byte aligned");
    if (int rc = make_map_2d(&g.a[i], reinterpret_cast<const float*>(A[j]), M, K[j], lda, TC_BM, in_half ? 2 : 4)) return rc;
    if (int rc = make_map_2d(&g.b[i], reinterpret_cast<const float*>(Bt[j]), N[j], K[j], ldbt[j], 128, in_half ? 2 : 4)) return rc;
    g.C[i] = C[j]; g.N[i] = N[j]; g.K[i] = K[j];
    if (i < nb && N[j] > n_max) n_max = N[j];
  }
  const int smem = TcSmem<128>::TOTAL;
  static uint64_t configured = 0;       // one bit per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (!((configured >> (dev & 63)) & 1u)) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_batch_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { set_error("rcb_gemm_tc_batch: smem opt-in failed: %s", cudaGetErrorString(e)); return -1; }
    configured |= 1ull << (dev & 63);
  }
  dim3 grid(ceil_div(M, TC_BM), ceil_div(n_max, 128), nb);
  gemm_tc_batch_kernel<128><<<grid, TC_THREADS, smem, (cudaStream_t)stream>>>(g, ldc, M, in_half, out_scale);
  RCB_CHECK_LAUNCH("rcb_gemm_tc_batch");
  return 0;
}
