// Tensor-core fused per-item SIREN MLP (forward + squared error + backward): TWO 128-pixel tiles of the item in flight
// per CTA (software-pipelined by the same 256 threads: while tile A's MMAs run, the threads do
// tile B's epilogue) and two CTAs per SM, i.e. four tiles in flight per SM.
//
// What makes four tiles fit:
//  * the chain  Z_l = X_l W_l,  y = X_3 W_3,  dX_l = dZ_l W_l^T  (tcgen05 kind::tf32, A operand in
//    TMEM) ping-pongs between two 32-column TMEM regions per tile and updates them in place
//    (tcgen05.ld -> sin / *cos -> tcgen05.st into the same lane and columns), so a tile owns 64
//    TMEM columns instead of 128;
//  * the activations the weight gradients need later are kept as feature-major fp16 copies in
//    shared memory (X in [-1,1]: fp16 keeps the same 10-bit mantissa TF32 does), written by the
//    forward epilogue while the values are in registers; dW_l = X_l^T dZ_l is a kind::f16 MMA
//    (K = 16 pixels per instruction, half the instructions and bytes of TF32).  The gradients
//    dZ_l^T are stored as fp16 too, in units of 1/coef (mode 1: the chain carries residuals,
//    O(1)) or of the caller's scale (mode 2), which keeps them in fp16's normal range; the
//    accumulators are fp32 and are unscaled when they leave TMEM;
//  * the derivative factors cos(.) are held packed as half2.
// As in the second design the A tiles carry a row of ones (accumulator row 32 = bias gradient),
// w0 is folded into the staged weights, and TF32 rounding is +0x1000 on the fp32 pattern.
// Reference semantics: test_model.py:347-355, 624-627; weight layout :269-280.
#include <cuda_fp16.h>

#include "tc_common.cuh"

namespace rcb {
namespace v3 {

constexpr int MT_EPI = 256;
constexpr int MT_THREADS = 256;        // two threads per pixel row, 16 features each; the warps take turns issuing the MMAs

// TMEM columns: tile slot s owns [64 s, 64 s + 64) = regions R0, R1; weight-gradient accumulators after
constexpr uint32_t TM_DW = 128, TM_COLS = 256;

struct Sm {
  static constexpr int XT_KB = 40 * 128;            // one 64-pixel K block (fp16): 32 feature rows, ones row, 7 zero rows
  static constexpr int XT_BYTES = 2 * XT_KB;
  // per tile slot
  static constexpr int XT1 = 0, XT2 = XT_BYTES, XT3 = 2 * XT_BYTES;   // X_l^T [2 K blocks][40][64 px]; XT3 later holds X_0^T
  static constexpr int DZT = 3 * XT_BYTES;          // dZ_l^T [2 K blocks][32][64 px]
  static constexpr int DZ3 = DZT + 2 * 4096;        // dy^T   [2 K blocks][16][64 px] (rows >= OUT stay zero)
  static constexpr int SLOT = DZ3 + 2 * 2048;
  // per CTA
  static constexpr int WF = 2 * SLOT;               // 3 forward B tiles  [j][i] (tf32)
  static constexpr int WB = WF + 3 * 4096;          // 3 backward B tiles [i][j] (layer 0: the 16 pe inputs)
  static constexpr int W3 = WB + 3 * 4096;          // output B tile [16 (OUT used)][32]
  static constexpr int PLAIN = W3 + 2048;           // floats: [0,96) w0*b_l, [96,100) b3, [104,108) sq partials, [128,256) W3[j][4]
  static constexpr int BAR = PLAIN + 1024;
  static constexpr int TOTAL = BAR + 64;
};
static_assert(Sm::SLOT % 1024 == 0 && Sm::WF % 1024 == 0 && Sm::W3 % 1024 == 0, "swizzled tiles need 1024-B alignment");
static_assert(Sm::SLOT + Sm::XT3 + Sm::XT_KB + 128 * 128 <= Sm::BAR, "M=128 reads past the last A tile must stay in the allocation");

// byte offset of element (row, col) in a tile of 128-byte rows with the 128-byte swizzle; 4- and 2-byte elements
__device__ __forceinline__ uint32_t swz4(int row, int col) {
  return (uint32_t)(row * 128 + ((((col >> 2) ^ (row & 7)) << 4) | ((col & 3) << 2)));
}
__device__ __forceinline__ uint32_t swz2(int row, int col) {
  return (uint32_t)(row * 128 + ((((col >> 3) ^ (row & 7)) << 4) | ((col & 7) << 1)));
}
__device__ __forceinline__ uint32_t rnd_tf32(float x) { return __float_as_uint(x) + 0x1000u; }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void sts16(uint32_t addr, __half v) {
  asm volatile("st.shared.b16 [%0], %1;" ::"r"(addr), "h"(__half_as_ushort(v)) : "memory");
}
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem: lane = row, one fp32 column per K element] * B[smem descriptor], TF32
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], fp16 operands, fp32 accumulation
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// instruction descriptor: D = f32, A = B = f16, both K-major, M = 128
__device__ __forceinline__ uint32_t idesc_f16(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  const __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_h2(uint32_t v) { return __half22float2(*reinterpret_cast<const __half2*>(&v)); }

#ifdef RCB_MLP_PROFILE
__device__ long long rcb_prof_buf[4 * 1024];
#define PROF(id)                                                                         \
  do {                                                                                   \
    if (prof_slot >= 0 && prof_n < 511) {                                                \
      rcb_prof_buf[prof_slot * 1024 + 2 * prof_n] = (id);                               \
      rcb_prof_buf[prof_slot * 1024 + 2 * prof_n + 1] = clock64();                      \
      ++prof_n;                                                                          \
    }                                                                                    \
  } while (0)
#else
#define PROF(id) do {} while (0)
#endif

template <int OUT, int MODE>
__global__ void __launch_bounds__(MT_THREADS, 2) mlp_tc_kernel(rcb_mlp_args a) {
#ifdef RCB_MLP_PROFILE
  int prof_n = 0;
  const int prof_slot = (blockIdx.x == 3000 && (threadIdx.x == 0 || threadIdx.x == 64)) ? (threadIdx.x == 0 ? 0 : 1)
                        : ((blockIdx.x == 3001 && (threadIdx.x == 0 || threadIdx.x == 64)) ? (threadIdx.x == 0 ? 2 : 3) : -1);
#endif
  constexpr int F = 16, HID = 32, NPE = 16;
  constexpr int off0 = 0, off1 = HID * (32 + 1), off2 = off1 + HID * (HID + 1), off3 = off2 + HID * (HID + 1);
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  float* plain = (float*)(smem + Sm::PLAIN);
  uint64_t* bar_ready = (uint64_t*)(smem + Sm::BAR);       // [2] epilogue -> MMA (256 arrivals), one per tile slot
  uint64_t* bar_mma = bar_ready + 2;                       // [2] MMA -> epilogue (commit of the stage's MMAs)
  uint32_t* tmem_slot = (uint32_t*)(bar_ready + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x;
  const int row_item = item / a.S;
  const int pix = a.pix;
  const int ntiles = (pix + 127) / 128;
  const float w0 = a.w0;
  const float* wt_g = a.wt + (int64_t)item * a.ld_w;
  // gradients travel in units of `unscale`: mode 1 carries residuals (dy / coef), mode 2 dy * coef
  const float gscale = MODE == 2 ? (a.coef > 0.f ? a.coef : 1.f) : 1.f;
  const float unscale = MODE == 1 ? a.coef : 1.f / gscale;

  if ((sbase & 1023u) != 0u) __trap();                     // the swizzled tiles assume a 1024-B aligned window
  if (threadIdx.x == 0) {
    mbar_init(&bar_ready[0], MT_EPI);
    mbar_init(&bar_ready[1], MT_EPI);
    mbar_init(&bar_mma[0], 1);
    mbar_init(&bar_mma[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(tmem_slot, TM_COLS);
  // ---- stage the item's weights (w0 folded in), the constant rows of the A tiles, zero padding of dy^T
  {
    const int t = threadIdx.x;
    for (int e = t; e < 3 * HID * HID; e += MT_EPI) {
      const int l = e / (HID * HID), r = (e / HID) % HID, c = e % HID;          // W_l[i = r][j = c]
      const int off = l == 0 ? off0 : (l == 1 ? off1 : off2);
      const uint32_t w = rnd_tf32(w0 * wt_g[off + HID + r * HID + c]);
      sts32(sbase + Sm::WF + l * 4096 + swz4(c, r), w);                         // forward B: rows j, K = i
      if (l > 0) sts32(sbase + Sm::WB + l * 4096 + swz4(r, c), w);              // backward B: rows i, K = j
      else if (r >= F) sts32(sbase + Sm::WB + swz4(r - F, c), w);               // layer 0: pe inputs only
    }
    for (int e = t; e < 16 * HID; e += MT_EPI) {
      const int k = e / HID, j = e % HID;                                       // W_3[j][k] -> rows k, K = j
      sts32(sbase + Sm::W3 + swz4(k, j), k < OUT ? rnd_tf32(wt_g[off3 + OUT + j * OUT + k]) : 0u);
    }
    for (int e = t; e < 3 * HID; e += MT_EPI) plain[e] = w0 * wt_g[(e / HID == 0 ? off0 : (e / HID == 1 ? off1 : off2)) + e % HID];
    if (t < 4) plain[96 + t] = t < OUT ? wt_g[off3 + t] : 0.f;
    for (int e = t; e < HID * 4; e += MT_EPI) plain[128 + e] = (e % 4 < OUT) ? wt_g[off3 + OUT + (e / 4) * OUT + e % 4] : 0.f;
    if (MODE != 0) {
      // rows 32..39 of every K block of every A tile: ones row + zeros (32-bit words = fp16 pairs)
      for (int e = t; e < 2 * 3 * 2 * 8 * 32; e += MT_EPI) {
        const int s = e / 1536, b = (e / 512) % 3, kb = (e / 256) % 2, rr = 32 + (e / 32) % 8, c = e % 32;
        sts32(sbase + s * Sm::SLOT + b * Sm::XT_BYTES + kb * Sm::XT_KB + rr * 128 + c * 4, rr == 32 ? 0x3c003c00u : 0u);
      }
      for (int e = t; e < 2 * 1024; e += MT_EPI) sts32(sbase + (e / 1024) * Sm::SLOT + Sm::DZ3 + (e % 1024) * 4, 0u);
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ---- MMA issue.  The stage's MMAs are issued by ONE warp, right after it has published its own
  // share: the whole warp waits for the 256 arrivals (converged), one elected lane issues and
  // commits.  Issue blocks while the tensor pipe is busy, so the warps take turns (stage, slot) ->
  // warp: no single warp carries all of it.  MMAs of consecutive stages of a slot are ordered by
  // completion (commit -> wait -> publish); the slots only share the weight-gradient accumulators,
  // where the in-order pipe makes every D += A B atomic.
  uint32_t ph_ready[2] = {0u, 0u};
  // 128 x n x 32 chain product (TF32), A in TMEM columns [a_col, a_col + 32)
  auto chain = [&](uint32_t d_col, uint32_t a_col, int b_off, uint32_t idesc) {
    const uint64_t db = smem_desc_sw128(sbase + b_off);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      umma_tf32_ts(tmem_base + d_col, tmem_base + a_col + (uint32_t)(k * 8), db + (uint64_t)(k * 2), idesc, k ? 1u : 0u);
  };
  // d[feature i (+ ones row)][j] += sum over the tile's 128 pixels of X^T[i][px] dZ^T[j][px]   (fp16 operands)
  auto wgrad = [&](uint32_t d_col, int xt_off, int dz_off, int dz_kb, uint32_t idesc, bool first) {
#pragma unroll
    for (int kb = 0; kb < 2; ++kb) {
      const uint64_t da = smem_desc_sw128(sbase + xt_off + kb * Sm::XT_KB);
      const uint64_t db = smem_desc_sw128(sbase + dz_off + kb * dz_kb);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_f16(tmem_base + d_col, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (first && kb == 0 && k == 0) ? 0u : 1u);
    }
  };
  // stages of one tile: 0-2 sine layers, 3 output layer, 4-6 backward.  R0 = 64 s, R1 = 64 s + 32.
  auto issue = [&](int s, int stage, bool first) {
    const bool mine = warp == ((2 * stage + s) & 7);
    if (mine) {
      mbar_wait(&bar_ready[s], ph_ready[s]);
      tc_fence_after();
    }
    if (mine && elect_one()) {
      const uint32_t t32 = idesc_tf32(32), t16 = idesc_tf32(16), h32 = idesc_f16(32), h16 = idesc_f16(16);
      const uint32_t R0 = (uint32_t)(64 * s), R1 = R0 + 32;
      const int so = s * Sm::SLOT;
      switch (stage) {
        case 0: chain(R1, R0, Sm::WF, t32); break;                              // Z0 = X0 W0
        case 1: chain(R0, R1, Sm::WF + 4096, t32); break;                       // Z1 = X1 W1
        case 2: chain(R1, R0, Sm::WF + 8192, t32); break;                       // Z2 = X2 W2
        case 3: chain(R0, R1, Sm::W3, t16); break;                              // y  = X3 W3
        case 4:
          chain(R0, R1, Sm::WB + 8192, t32);                                    // dX2 = dZ2 W2^T
          wgrad(TM_DW + 96, so + Sm::XT3, so + Sm::DZ3, 2048, h16, first);      // dW3 = X3^T dy
          wgrad(TM_DW + 64, so + Sm::XT2, so + Sm::DZT, 4096, h32, first);      // dW2 = X2^T dZ2
          break;
        case 5:
          chain(R1, R0, Sm::WB + 4096, t32);                                    // dX1 = dZ1 W1^T
          wgrad(TM_DW + 32, so + Sm::XT1, so + Sm::DZT, 4096, h32, first);      // dW1 = X1^T dZ1
          break;
        default:
          chain(R0, R1, Sm::WB, t16);                                           // d pe = dZ0 W0[pe rows]^T
          wgrad(TM_DW, so + Sm::XT3, so + Sm::DZT, 4096, h32, first);           // dW0 = X0^T dZ0
          break;
      }
      umma_commit(&bar_mma[s]);
    }
    ph_ready[s] ^= 1;
    __syncwarp();
  };

  const int q = warp & 3;                 // TMEM lane quarter
  const int hh = warp >> 2;               // which 16 of the 32 features
  const int r = q * 32 + lane;            // pixel row inside the tile
  const int j0 = hh * 16;
  const uint32_t tm = tmem_base + ((uint32_t)(q * 32) << 16);
  const float* xt = a.xt + (int64_t)row_item * a.x_row_stride;
  const bool stitched = a.pe_base != nullptr;
  const int64_t pe_origin = stitched ? a.pe_base[item] : (int64_t)item * pix;
  const int php = a.ph * a.pw;
  auto pe_off = [&](int gp) -> int64_t {
    if (!stitched) return gp;
    int z = gp / php, rem = gp - z * php;
    int yy = rem / a.pw, xx = rem - yy * a.pw;
    return (int64_t)z * a.pitch_z + (int64_t)yy * a.pitch_y + xx;
  };
  // feature-major fp16 element of this thread's pixel: K block r / 64, column r % 64
  const uint32_t t_kb = (uint32_t)(q >> 1);
  const int t_col = (q & 1) * 32 + lane;
  auto store_t = [&](uint32_t tile, int kb_bytes, const float (&v)[16]) {     // tile = address of K block 0
    const uint32_t base = tile + t_kb * (uint32_t)kb_bytes;
#pragma unroll
    for (int j = 0; j < 16; ++j) sts16(base + swz2(j0 + j, t_col), __float2half_rn(v[j]));
  };
  // this thread's 16 input features of pixel gp: Fourier features (half 0) or positional encodings (half 1)
  auto load_x0 = [&](int gp, uint32_t (&v)[16]) {
    const bool ok = gp < pix;
    if (hh == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = ok ? __float_as_uint(__ldg(xt + (int64_t)i * pix + gp)) : 0u;
    } else {
      const uint4* p = reinterpret_cast<const uint4*>(a.pe + (pe_origin + (ok ? pe_off(gp) : 0)) * NPE);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint4 t4 = ok ? __ldg(p + c) : make_uint4(0u, 0u, 0u, 0u);
        v[c * 4] = t4.x; v[c * 4 + 1] = t4.y; v[c * 4 + 2] = t4.z; v[c * 4 + 3] = t4.w;
      }
    }
  };
  auto publish = [&](int s, bool smem_written) {   // hand this thread's share of the stage's operands to the MMAs
    tmem_st_wait();
    PROF(310 + s);
    if (smem_written) fence_async_smem();
    PROF(320 + s);
    tc_fence_before();
    mbar_arrive(&bar_ready[s]);
    PROF(300 + s);
  };
  uint32_t ph_mma[2] = {0u, 0u};
  int prof_stage[2] = {0, 0};
  auto wait_mma = [&](int s) {
    PROF(100 + prof_stage[s] * 10 + s);
    mbar_wait(&bar_mma[s], ph_mma[s]);
    ph_mma[s] ^= 1;
    tc_fence_after();
    PROF(200 + prof_stage[s] * 10 + s);
    prof_stage[s] = (prof_stage[s] + 1) % 7;
  };
  float sq = 0.f;
  uint32_t xin[2][16];
  load_x0(r, xin[0]);
  if (ntiles > 1) load_x0(128 + r, xin[1]);

  for (int t0 = 0; t0 < ntiles; t0 += 2) {
    const bool act1 = t0 + 1 < ntiles;
    const bool first = t0 == 0;
    uint32_t cs[2][3][8];                         // cos(.) of the three sine layers, packed half2
    float dy[2][OUT];
    // ---- X0 -> TMEM (R0)
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      if (s == 1 && !act1) continue;
#pragma unroll
      for (int i = 0; i < 16; ++i) xin[s][i] += 0x1000u;                 // TF32 round-to-nearest of the inputs
      tmem_st16(tm + (uint32_t)(64 * s) + j0, xin[s]);
      publish(s, false);
      issue(s, 0, first && s == 0);
    }
    // ---- three sine layers: X_{l+1} = sin(acc + b') back into the accumulator's columns, fp16 copy of X_{l+1}^T
#pragma unroll
    for (int l = 0; l < 3; ++l) {
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        if (s == 1 && !act1) continue;
        const uint32_t reg = tm + (uint32_t)(64 * s + ((l & 1) ? 0 : 32)) + j0;   // Z0 -> R1, Z1 -> R0, Z2 -> R1
        wait_mma(s);
        uint32_t acc[16];
        tmem_ld16_issue(reg, acc);
        tmem_ld_wait();
        PROF(400 + s);
        float x[16];
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          const float z0 = __uint_as_float(acc[j]) + plain[l * 32 + j0 + j];
          const float z1 = __uint_as_float(acc[j + 1]) + plain[l * 32 + j0 + j + 1];
          x[j] = __sinf(z0); x[j + 1] = __sinf(z1);
          cs[s][l][j >> 1] = pack_h2(__cosf(z0), __cosf(z1));
          acc[j] = rnd_tf32(x[j]); acc[j + 1] = rnd_tf32(x[j + 1]);
        }
        PROF(410 + s);
        tmem_st16(reg, acc);
        PROF(420 + s);
        if (MODE != 0) store_t(sbase + s * Sm::SLOT + (l == 0 ? Sm::XT1 : (l == 1 ? Sm::XT2 : Sm::XT3)), Sm::XT_KB, x);
        PROF(430 + s);
        publish(s, MODE != 0);
        issue(s, l + 1, first && s == 0);
      }
    }
    // ---- output layer (on the tensor core), loss and dy; dZ2 = (dy W3^T) * cos in place of X3; dy^T, dZ2^T
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      if (s == 1 && !act1) continue;
      const int gp = (t0 + s) * 128 + r;
      const bool valid = gp < pix;
      wait_mma(s);
      {
        uint32_t yv[16];
        tmem_ld16_issue(tm + (uint32_t)(64 * s), yv);       // both halves of a row read the same columns
        tmem_ld_wait();
        if (MODE == 0) {
          if (hh == 0 && valid)
#pragma unroll
            for (int k = 0; k < OUT; ++k) a.y_pred[((int64_t)item * pix + gp) * OUT + k] = __uint_as_float(yv[k]) + plain[96 + k];
        } else {
#pragma unroll
          for (int k = 0; k < OUT; ++k) {
            if (MODE == 1) {
              const float rr = valid ? __uint_as_float(yv[k]) + plain[96 + k] - __ldg(a.y + ((int64_t)row_item * pix + gp) * OUT + k) : 0.f;
              if (hh == 0) sq = fmaf(rr, rr, sq);
              dy[s][k] = rr;
            } else {
              dy[s][k] = valid ? gscale * __ldg(a.dy + ((int64_t)item * pix + gp) * OUT + k) : 0.f;
            }
          }
        }
      }
      if (MODE == 0) continue;
      const int so = s * Sm::SLOT;
      if (hh == 0) {
#pragma unroll
        for (int k = 0; k < OUT; ++k) sts16(sbase + so + Sm::DZ3 + t_kb * 2048 + swz2(k, t_col), __float2half_rn(dy[s][k]));
      }
      float dz[16];
      uint32_t dzr[16];
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        const float2 c2 = unpack_h2(cs[s][2][j >> 1]);
        const float4 wa = *(const float4*)(plain + 128 + (j0 + j) * 4);
        const float4 wb = *(const float4*)(plain + 128 + (j0 + j + 1) * 4);
        float va = dy[s][0] * wa.x, vb = dy[s][0] * wb.x;
        if (OUT > 1) { va = fmaf(dy[s][1], wa.y, va); vb = fmaf(dy[s][1], wb.y, vb); }
        if (OUT > 2) { va = fmaf(dy[s][2], wa.z, va); vb = fmaf(dy[s][2], wb.z, vb); }
        dz[j] = va * c2.x; dz[j + 1] = vb * c2.y;
        dzr[j] = rnd_tf32(dz[j]); dzr[j + 1] = rnd_tf32(dz[j + 1]);
      }
      tmem_st16(tm + (uint32_t)(64 * s + 32) + j0, dzr);    // R1: A operand of dX2
      store_t(sbase + so + Sm::DZT, 4096, dz);
      publish(s, true);
      issue(s, 4, first && s == 0);
    }
    if (MODE != 0) {
      // ---- dZ1 (from R0, in place), dZ0 (from R1, in place) = data gradient * cos; X0^T reloaded for dW0
#pragma unroll
      for (int l = 1; l >= 0; --l) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          if (s == 1 && !act1) continue;
          const int so = s * Sm::SLOT;
          const uint32_t reg = tm + (uint32_t)(64 * s + (l == 1 ? 0 : 32)) + j0;
          uint32_t x0[16];
          if (l == 0) load_x0((t0 + s) * 128 + r, x0);        // in flight while the MMAs finish
          wait_mma(s);
          uint32_t acc[16];
          tmem_ld16_issue(reg, acc);
          tmem_ld_wait();
          float dz[16];
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            const float2 c2 = unpack_h2(cs[s][l][j >> 1]);
            dz[j] = __uint_as_float(acc[j]) * c2.x; dz[j + 1] = __uint_as_float(acc[j + 1]) * c2.y;
            acc[j] = rnd_tf32(dz[j]); acc[j + 1] = rnd_tf32(dz[j + 1]);
          }
          tmem_st16(reg, acc);
          store_t(sbase + so + Sm::DZT, 4096, dz);
          if (l == 0) {
            float xf[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) xf[j] = __uint_as_float(x0[j]);
            store_t(sbase + so + Sm::XT3, Sm::XT_KB, xf);
          }
          publish(s, true);
          issue(s, l == 1 ? 5 : 6, first && s == 0);
        }
      }
    }
    // ---- the next pair's inputs travel while the last MMAs of this pair run
    if (t0 + 2 < ntiles) load_x0((t0 + 2) * 128 + r, xin[0]);
    if (t0 + 3 < ntiles) load_x0((t0 + 3) * 128 + r, xin[1]);
    if (MODE == 0) continue;
    // ---- d pe (16 columns of R0: 8 per half), back in true units
    float dpe[2][8];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      if (s == 1 && !act1) continue;
      wait_mma(s);
      uint32_t acc[16];
      tmem_ld16_issue(tm + (uint32_t)(64 * s), acc);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 8; ++c) dpe[s][c] = unscale * __uint_as_float(hh ? acc[8 + c] : acc[c]);
    }
    // both halves of a pixel row read the same 16 columns, and the next pair's X0 goes into them:
    // nobody may run ahead into the next pair before every thread has its d pe
    tc_fence_before();
    asm volatile("bar.sync 1, 256;" ::: "memory");
    tc_fence_after();
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      if (s == 1 && !act1) continue;
      const int gp = (t0 + s) * 128 + r;
      if (gp < pix) {
        float* dst = a.d_pe + (pe_origin + pe_off(gp)) * NPE + hh * 8;
        *(float4*)(dst) = make_float4(dpe[s][0], dpe[s][1], dpe[s][2], dpe[s][3]);
        *(float4*)(dst + 4) = make_float4(dpe[s][4], dpe[s][5], dpe[s][6], dpe[s][7]);
      }
    }
  }

  if (MODE != 0) {
    // ---- weight gradients: TMEM lane i < 32 = input feature i, lane 32 = the ones row (bias)
    float* g = a.d_wt + (int64_t)item * a.ld_w;
    if (q == 0 || q == 1) {
      uint32_t acc[16];
      const float sc = w0 * unscale;
#pragma unroll
      for (int l = 0; l < 3; ++l) {
        tmem_ld16_issue(tm + TM_DW + (uint32_t)(l * 32 + j0), acc);
        tmem_ld_wait();
        const int off = l == 0 ? off0 : (l == 1 ? off1 : off2);
        float* dst = q == 0 ? g + off + HID + lane * HID + j0 : g + off + j0;
        if (q == 0 || lane == 0) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
            *(float4*)(dst + c * 4) = make_float4(sc * __uint_as_float(acc[c * 4]), sc * __uint_as_float(acc[c * 4 + 1]),
                                                  sc * __uint_as_float(acc[c * 4 + 2]), sc * __uint_as_float(acc[c * 4 + 3]));
        }
      }
      if (hh == 0) {
        tmem_ld16_issue(tm + TM_DW + 96, acc);
        tmem_ld_wait();
        if (q == 0) {
#pragma unroll
          for (int k = 0; k < OUT; ++k) g[off3 + OUT + lane * OUT + k] = unscale * __uint_as_float(acc[k]);
        } else if (lane == 0) {
#pragma unroll
          for (int k = 0; k < OUT; ++k) g[off3 + k] = unscale * __uint_as_float(acc[k]);
        }
      }
    }
    if (MODE == 1) {
      sq = warp_sum(sq);
      if (hh == 0 && lane == 0) plain[104 + q] = sq;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (MODE == 1 && threadIdx.x == 0) a.sqerr[item] = (plain[104] + plain[105]) + (plain[106] + plain[107]);
  if (warp == 0) tmem_dealloc(tmem_base, TM_COLS);
}

template <int OUT>
static int launch(const rcb_mlp_args* a, cudaStream_t st) {
#define RCB_MT_LAUNCH(MODE)                                                                                   \
  do {                                                                                                        \
    cudaError_t e = cudaFuncSetAttribute(mlp_tc_kernel<OUT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         Sm::TOTAL);                                                          \
    if (e != cudaSuccess) { set_error("rcb_mlp_tc: smem opt-in failed: %s", cudaGetErrorString(e)); return -1; } \
    mlp_tc_kernel<OUT, MODE><<<a->items, MT_THREADS, Sm::TOTAL, st>>>(*a);                                   \
  } while (0)
  if (a->mode == 0) RCB_MT_LAUNCH(0);
  else if (a->mode == 1) RCB_MT_LAUNCH(1);
  else RCB_MT_LAUNCH(2);
#undef RCB_MT_LAUNCH
  RCB_CHECK_LAUNCH("rcb_mlp_tc");
  return 0;
}

}  // namespace v3
}  // namespace rcb

using namespace rcb;

#ifdef RCB_MLP_PROFILE
extern "C" int rcb_mlp_prof_read(long long* host) {
  return (int)cudaMemcpyFromSymbol(host, rcb::v3::rcb_prof_buf, sizeof(long long) * 4 * 1024);
}
#endif

extern "C" int rcb_mlp_tc(const rcb_mlp_args* a, rcb_stream_t stream) {
  RCB_CHECK_ARG(a != nullptr, "rcb_mlp_tc: null args");
  RCB_CHECK_ARG(a->items > 0 && a->S > 0 && a->pix > 0, "rcb_mlp_tc: empty problem");
  RCB_CHECK_ARG(a->mode >= 0 && a->mode <= 2, "rcb_mlp_tc: bad mode %d", a->mode);
  RCB_CHECK_ARG(a->wt && a->xt && a->pe, "rcb_mlp_tc: null input");
  RCB_CHECK_ARG(a->n_f == 16, "rcb_mlp_tc: built for 32 input features (16 Fourier + 16 pe); use rcb_mlp for other shapes");
  RCB_CHECK_ARG(a->mode != 0 || a->y_pred, "rcb_mlp_tc: mode 0 needs y_pred");
  RCB_CHECK_ARG(a->mode != 1 || (a->y && a->sqerr && a->coef > 0.f), "rcb_mlp_tc: mode 1 needs y, sqerr and coef > 0");
  RCB_CHECK_ARG(a->mode != 2 || a->dy, "rcb_mlp_tc: mode 2 needs dy");
  RCB_CHECK_ARG(a->mode == 0 || (a->d_pe && a->d_wt), "rcb_mlp_tc: backward needs d_pe and d_wt");
  RCB_CHECK_ARG(a->ld_w % 4 == 0, "rcb_mlp_tc: ld_w must be a multiple of 4");
  RCB_CHECK_ARG(!a->pe_base || (a->ph > 0 && a->pw > 0), "rcb_mlp_tc: stitched addressing needs the patch extent");
  cudaStream_t st = (cudaStream_t)stream;
  if (a->out == 3) return v3::launch<3>(a, st);
  if (a->out == 1) return v3::launch<1>(a, st);
  set_error("rcb_mlp_tc: unsupported output width %d", a->out);
  return -2;
}
