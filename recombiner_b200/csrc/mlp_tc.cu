// Tensor-core fused per-item SIREN MLP (forward + squared error + backward): TWO 128-pixel tiles of
// the item in flight per CTA -- warps 0-3 own the even tiles, warps 4-7 the odd ones, one thread per
// pixel row with all 32 features, so one group's epilogue overlaps the other group's MMA round
// trip -- and two CTAs per SM, i.e. four tiles in flight per SM.
//
// What makes four tiles fit:
//  * the chain  Z_l = X_l W_l,  y = X_3 W_3,  dX_l = dZ_l W_l^T  (tcgen05 kind::f16, the A operand as
//    packed fp16 pairs in TMEM) ping-pongs between two 32-column TMEM regions per tile and updates
//    them in place (tcgen05.ld -> sin / *cos -> tcgen05.st into the same lane and columns), so a
//    tile owns 64 TMEM columns instead of 128;
//  * the activations the weight gradients need later are kept as feature-major fp16 copies in
//    shared memory (X in [-1,1]: fp16 keeps the same 10-bit mantissa TF32 does), written by the
//    forward epilogue while the values are in registers; dW_l = X_l^T dZ_l is a kind::f16 MMA
//    (K = 16 pixels per instruction, half the instructions and bytes of TF32).  The gradients
//    dZ_l^T are stored as fp16 too, in units of 1/coef (mode 1: the chain carries residuals,
//    O(1)) or of the caller's scale (mode 2), which keeps them in fp16's normal range; the
//    accumulators are fp32 and are unscaled when they leave TMEM;
//  * the derivative factors cos(.) are held packed as half2.
// A tile costs SIX MMA round trips: the first product of the group's NEXT tile (Z0 = X0 W0) is issued in the same batch
// as the last one of the current tile (d pe = dZ0 W0^T), whose result is read back at the top of the next tile; the inputs
// X0 are kept as an fp16 copy in shared memory for dW0 (no second load), and the constant (1, 1) block that picks up the
// biases has TMEM columns of its own (written once).
// The chain products of the last two backward stages are committed before their weight-gradient MMAs (early commits,
// see EARLY5 / EARLY6 in the kernel).  The A tiles of the weight gradients carry a row of ones (accumulator row 32 =
// bias gradient) and w0 is folded into the staged weights.
// Reference semantics: test_model.py:347-355, 624-627; weight layout :269-280.
#include <cuda_fp16.h>

#include "tc_common.cuh"

namespace rcb {
namespace mlp {

constexpr int MT_THREADS = 256;        // 2 groups x 128 threads: one thread per pixel row of the group's tile
constexpr int MT_GROUP = 128;

// TMEM columns: tile slot s owns [64 s, 64 s + 64) = regions R0, R1; weight-gradient accumulators after (112 columns);
// [240 + 8 s, 248 + 8 s): the slot's bias block of the chain's A operand (K elements (1, 1, 0 ..); 34 inputs: (x32, x33, 1, 1, 0 ..))
constexpr uint32_t TM_DW = 128, TM_ONES = 240, TM_COLS = 256;

struct Sm {
  // fp16 operands of the weight-gradient MMAs, pixel-major (MN-major for the MMA): one 64-byte row of 32 features per
  // pixel, 64-byte swizzle, 8-pixel groups 512 B apart
  static constexpr int XP = 128 * 64;
  // per tile slot
  // X_0 (the tile's inputs, for dW0) comes first: the second atom of its 34-input operand (over X_2) must lie above it
  static constexpr int XT0 = 0, XT1 = XP, XT2 = 2 * XP, XT3 = 3 * XP;   // X_0 .. X_3
  static constexpr int DZT = 4 * XP;                // dZ_l
  static constexpr int DZ3 = 5 * XP;                // dy^T   [2 K blocks][16][64 px], K-major (rows >= OUT stay zero)
  static constexpr int SLOT = DZ3 + 2 * 2048;
  // per CTA
  // second MN atom of every A operand: feature 32 = 1 (bias row), 33..63 = 0.  Every pixel row is the same, so ONE K = 16
  // step (16 pixel rows, 1 KB) serves all eight steps: the descriptor's leading-byte offset shrinks by 1 KB per step
  static constexpr int ONES = 2 * SLOT;
  static constexpr int WT = 32 * 64;                // one fp16 weight tile: 32 rows of 32 K elements (64 B), 64-byte swizzle
  static constexpr int WF = ONES + 1024;            // 3 forward B tiles  [j][i]
  static constexpr int WB = WF + 3 * WT;            // 3 backward B tiles [i][j] (layer 0: the 16 pe inputs)
  static constexpr int W3 = WB + 3 * WT;            // output B tile [16 (OUT used)][32]
  static constexpr int WF0B = W3 + WT;              // forward B tile of layer 0 for input features 32..47 (34-input INRs: video)
  // the biases of the sine layers ride the chain MMAs: a third K = 16 block per layer whose rows hold (hi, lo) = the fp16
  // split of w0 * b_l[j] in K elements 0, 1, against two constant 1.0 in the A operand (exact to 2^-22 relative)
  static constexpr int WBI = WF0B + WT;             // 3 bias tiles [j][16 K], as 64-byte rows like the others (34 inputs: layer 0 uses WF0B)
  static constexpr int PLAIN = WBI + 3 * WT;        // floats: [96,100) b3, [104,112) sq partials, [128,256) W3[j][4]
  static constexpr int BAR = PLAIN + 1024;
  static constexpr int TOTAL = BAR + 64;
};
static_assert(2 * (Sm::TOTAL + 1024) <= 228 * 1024, "two CTAs per SM");
static_assert(Sm::SLOT % 1024 == 0 && Sm::WF % 1024 == 0 && Sm::W3 % 1024 == 0, "swizzled tiles need 1024-B alignment");

// byte offset of element (row, col) in a tile of 128-byte rows with the 128-byte swizzle; 4- and 2-byte elements
__device__ __forceinline__ uint32_t swz4(int row, int col) {
  return (uint32_t)(row * 128 + ((((col >> 2) ^ (row & 7)) << 4) | ((col & 3) << 2)));
}
__device__ __forceinline__ uint32_t swz2(int row, int col) {
  return (uint32_t)(row * 128 + ((((col >> 3) ^ (row & 7)) << 4) | ((col & 7) << 1)));
}
// byte offset of fp16 element (row, col < 32) in a tile of 64-byte rows with the 64-byte swizzle
__device__ __forceinline__ uint32_t swz64(int row, int col) {
  return (uint32_t)(row * 64 + ((((col >> 3) ^ ((row >> 1) & 3)) << 4) | ((col & 7) << 1)));
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
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
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
// D[tmem] (+)= A[tmem: lane = row, two fp16 K elements per 32-bit column] * B[smem descriptor], fp16 operands
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// D = f32, A = B = f16, both K-major, M = 128 (the chain products)
__device__ __forceinline__ uint32_t idesc_f16_m128(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
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
// instruction descriptor: D = f32, A = B = f16, both K-major, M = 64 (the weight-gradient tiles have 33 useful
// rows: half the shared-memory reads of M = 128).  Accumulator row i lives in TMEM lane (i % 16) + 32 * (i / 16).
__device__ __forceinline__ uint32_t idesc_f16(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);
}
// MN-major operand (pixel rows of 64 B, 64-byte swizzle): leading-byte offset = distance to the next 32-feature atom,
// stride-byte offset = 512 (eight pixel rows); one K = 16 step is 1024 B further
__device__ __forceinline__ uint64_t smem_desc_pm(uint32_t saddr, uint32_t lbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;      // SWIZZLE_64B
  return d;
}
constexpr uint32_t IDESC_A_MN = 1u << 15, IDESC_B_MN = 1u << 16;
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  const __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}
// same, saturating at +-65504 instead of producing inf (one F2FP.SATFINITE; NaN stays NaN)
__device__ __forceinline__ uint32_t pack_h2_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float2 unpack_h2(uint32_t v) { return __half22float2(*reinterpret_cast<const __half2*>(&v)); }
// fp16 pair (lo * c.x, hi * c.y) for a packed fp16 pair c.  Two packed forms were measured and are SLOWER than the plain
// unpack + two FMULs: one HMUL2 on the packed words (144 fewer instructions per pixel row and tile; 0.589 against
// 0.565 ms) and one FMUL2 (mul.f32x2) per pair (48 fewer; 0.535 against 0.517 ms).
__device__ __forceinline__ uint32_t mul_h2(float lo, float hi, uint32_t c) {
  const float2 c2 = unpack_h2(c);
  return pack_h2(lo * c2.x, hi * c2.y);
}

// Predicated read-only loads whose destination registers are zeroed BEFORE the load is issued.  The form
// `v = ok ? __ldg(p) : 0` compiles to the load followed by a predicated zeroing of the same registers, and a write to the
// destination of a load in flight waits for the load even when its predicate is off: the thread sat at the load site
// for the whole memory latency (2.6 % of all stall samples of the kernel) instead of at the first use.
__device__ __forceinline__ void ldg16_if(bool ok, const void* p, uint32_t& x, uint32_t& y, uint32_t& z, uint32_t& w) {
  x = y = z = w = 0u;
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "setp.ne.b32 q, %4, 0;\n\t"
      "@q ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%5];\n\t"
      "}" : "+r"(x), "+r"(y), "+r"(z), "+r"(w) : "r"((uint32_t)ok), "l"(p));
}
__device__ __forceinline__ void ldg4_if(bool ok, const void* p, uint32_t& x) {
  x = 0u;
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "setp.ne.b32 q, %1, 0;\n\t"
      "@q ld.global.nc.u32 %0, [%2];\n\t"
      "}" : "+r"(x) : "r"((uint32_t)ok), "l"(p));
}

__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
        "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
        "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}

#ifdef RCB_MLP_PROFILE
__device__ long long rcb_prof_buf[4 * 1024];
#define PROF(id)                                                                         \
  do {                                                                                   \
    if (prof_slot >= 0 && prof_n < 511) {                                                \
      rcb_prof_buf[prof_slot * 1024 + 2 * prof_n] = (id);                                \
      rcb_prof_buf[prof_slot * 1024 + 2 * prof_n + 1] = clock64();                       \
      ++prof_n;                                                                          \
    }                                                                                    \
  } while (0)
#else
#define PROF(id) do {} while (0)
#endif

// F = Fourier features per pixel: 16 (32 INR inputs; cifar, kodak, audio, protein) or 18 (34 inputs; video).  With 34
// inputs the first layer runs three K = 16 steps (features 32, 33 live in a third, zero-padded K block) and the second
// 32-feature atom of dW0's A operand is a per-tile block [x32, x33, 1, 0, ...] instead of the shared bias-only block.
// D = signal dimensionality (1, 2, 3): only used to generate the Fourier inputs from the pixel index (a.x_tab).
template <int OUT, int MODE, int F, int D>
__global__ void __launch_bounds__(MT_THREADS, 2) mlp_tc_kernel(rcb_mlp_args a) {
#ifdef RCB_MLP_PROFILE
  int prof_n = 0;
  const int prof_slot = (blockIdx.x == 3000 && (threadIdx.x == 32 || threadIdx.x == 160)) ? (threadIdx.x == 32 ? 0 : 1) : -1;
#endif
  constexpr int HID = 32, NPE = 16, IN = F + NPE;
  // Early commits.  Stage 5: the chain product dX1 is committed before the stage's weight-gradient MMAs are even issued,
  // so its epilogue neither waits for them nor for the other group's turn.  That epilogue then must not write into dZ^T,
  // which dW1 is still reading: dZ0^T goes over X3^T, whose last reader (dW3, stage 4) is complete by then.  Stage 6
  // likewise commits after d pe and the next tile's first product; dW0 follows uncommitted (the group's next commit covers
  // it) and the next tile's X0^T, which would overwrite its operand, is stored one stage later.  (Stage 4 has no free
  // buffer for dZ1^T; with a barrier of their own for the weight-gradient MMAs and three more waits per tile the same idea
  // measured slower.)
  constexpr bool EARLY5 = true, EARLY6 = true;
  // dW3 (eight N = 16 MMAs) leaves the batch of stage 4, the longest wait of a tile, for the uncommitted tail of stage 5;
  // its operand X3^T must then survive the last epilogue, so dZ0^T goes over X2^T instead (32-input form only: the
  // 34-input form keeps its second dW0 atom there)
  constexpr bool DW3_LATE = IN <= 32;
  constexpr int DZ0 = DW3_LATE ? Sm::XT2 : (EARLY5 ? Sm::XT3 : Sm::DZT);
  constexpr bool WIDE = IN > 32;
  static_assert(F == 16 || F == 18, "16 or 18 Fourier features");
  constexpr int off0 = 0, off1 = HID * (IN + 1), off2 = off1 + HID * (HID + 1), off3 = off2 + HID * (HID + 1);
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  float* plain = (float*)(smem + Sm::PLAIN);
  uint64_t* bar_ready = (uint64_t*)(smem + Sm::BAR);       // [2] epilogue -> MMA (128 arrivals), one per group
  uint64_t* bar_mma = bar_ready + 2;                       // [2] MMA -> epilogue (commit of the stage's MMAs)
  uint32_t* tmem_slot = (uint32_t*)(bar_ready + 4);
  volatile int* wg_turn = (volatile int*)(tmem_slot + 1);    // [3] next tile allowed to issue its weight-gradient MMAs, per backward stage

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

  const int g = warp >> 2;                // group = tile slot: tiles g, g + 2, g + 4, ...
  const int q = warp & 3;                 // TMEM lane quarter
  const int r = q * 32 + lane;            // pixel row inside the tile
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
  // per-item bases, so that a pixel's pe / d pe address is one 32-bit multiply-add away (unstitched grids)
  const char* pe_item = reinterpret_cast<const char*>(a.pe) + pe_origin * (NPE * (a.pe_half ? 2 : 4));
  char* dpe_item = a.d_pe_h ? reinterpret_cast<char*>(a.d_pe_h) + pe_origin * (NPE * 2)
                            : reinterpret_cast<char*>(a.d_pe) + pe_origin * (NPE * 4);
  // coordinate decode of the generated inputs: shifts when the inner extents are powers of two (every shipped shape)
  int x_sh[3] = {0, 0, 0};
  bool x_pow2 = true;
#pragma unroll
  for (int ax = 1; ax < D; ++ax) {
    const int sz = a.x_size[ax];
    x_sh[ax] = 31 - __clz(sz);
    x_pow2 = x_pow2 && (sz & (sz - 1)) == 0;
  }
  // the 32 input features of pixel gp: 16 Fourier features, 16 positional encodings (raw fp32 patterns)
  constexpr int NFQ = F / (2 * D);          // frequencies per axis
  static_assert(2 * D * NFQ == F, "Fourier feature count does not match the dimensionality");
  auto load_x0 = [&](int gp, uint32_t (&v)[IN]) {
    const bool ok = gp < pix;
    if (a.x_tab) {
      // generated inputs: per-axis rows [cos(pi c w_0..), sin(pi c w_0..)] looked up by the coordinate indices of the
      // pixel; feature order [cos axis 0 .., cos axis 1 .., sin axis 0 .., sin axis 1 ..] (data/image.py:25-27)
      int rem = ok ? gp : 0;
#pragma unroll
      for (int ax = D - 1; ax >= 0; --ax) {
        const int sz = a.x_size[ax];
        int i;
        if (ax == 0) { i = rem; }
        else if (x_pow2) { i = rem & (sz - 1); rem >>= x_sh[ax]; }
        else { i = rem % sz; rem /= sz; }
        const float* row = a.x_tab + a.x_off[ax] + i * (2 * NFQ);
        if (NFQ % 4 == 0) {
#pragma unroll
          for (int c = 0; c < NFQ / 4; ++c) {
            const int ic = ax * NFQ + 4 * c, is = D * NFQ + ax * NFQ + 4 * c;
            ldg16_if(ok, reinterpret_cast<const float4*>(row) + c, v[ic], v[ic + 1], v[ic + 2], v[ic + 3]);
            ldg16_if(ok, reinterpret_cast<const float4*>(row + NFQ) + c, v[is], v[is + 1], v[is + 2], v[is + 3]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < NFQ; ++j) {
            ldg4_if(ok, row + j, v[ax * NFQ + j]);
            ldg4_if(ok, row + NFQ + j, v[D * NFQ + ax * NFQ + j]);
          }
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < F; ++i) ldg4_if(ok, xt + (int64_t)i * pix + (ok ? gp : 0), v[i]);
    }
    if (F == 16 && a.pe_half) {      // fp16 positional encodings: v[16..23] are already the packed pairs the chain operand needs
      const uint4* p = reinterpret_cast<const uint4*>(stitched ? pe_item + (ok ? pe_off(gp) : 0) * (NPE * 2)
                                                               : pe_item + (uint32_t)(ok ? gp : 0) * (uint32_t)(NPE * 2));
#pragma unroll
      for (int c = 0; c < 2; ++c) ldg16_if(ok, p + c, v[16 + c * 4], v[17 + c * 4], v[18 + c * 4], v[19 + c * 4]);
#pragma unroll
      for (int c = 24; c < 32; ++c) v[c] = 0u;
      return;
    }
    const uint4* p = reinterpret_cast<const uint4*>(stitched ? pe_item + (ok ? pe_off(gp) : 0) * (NPE * 4)
                                                             : pe_item + (uint32_t)(ok ? gp : 0) * (uint32_t)(NPE * 4));
#pragma unroll
    for (int c = 0; c < 4; ++c) ldg16_if(ok, p + c, v[F + c * 4], v[F + 1 + c * 4], v[F + 2 + c * 4], v[F + 3 + c * 4]);
  };
  uint32_t xin[IN];
  if (g < ntiles) load_x0(g * 128 + r, xin);    // first tile's inputs travel under the weight staging
#ifndef RCB_MLP_NO_L2_PREFETCH
  {
    // The weight samples were written a whole upsampler pass ago and have left L2 by now: every CTA starts on ~13 KB of
    // DRAM misses.  Ask L2 for the rows of the CTA that will run about one CTA lifetime from now (two per SM are resident).
    int sms;
    asm("mov.u32 %0, %%nsmid;" : "=r"(sms));
    const int ahead = item + 2 * sms;
    const int lines = (a.ld_w * 4 + 127) / 128;
    if (ahead < a.items && (int)threadIdx.x < lines)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(a.wt + (int64_t)ahead * a.ld_w) + threadIdx.x * 128));
  }
#endif

  PROF(1);
  if ((sbase & 1023u) != 0u) __trap();                     // the swizzled tiles assume a 1024-B aligned window
  if (threadIdx.x == 0) {
    mbar_init(&bar_ready[0], MT_GROUP);
    mbar_init(&bar_ready[1], MT_GROUP);
    mbar_init(&bar_mma[0], 1);
    mbar_init(&bar_mma[1], 1);
    wg_turn[0] = wg_turn[1] = wg_turn[2] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(tmem_slot, TM_COLS);
  // ---- stage the item's weights (w0 folded in), the constant rows of the A tiles, zero padding of dy^T
  {
    const int t = threadIdx.x;
    // all global loads first (they are independent), then the swizzled stores: the st.shared asm
    // statements are ordering points for the compiler, and a load-store-load-store chain would pay
    // the full memory latency twelve times per thread
    float wv[12], w3v[2], b3v = 0.f, w3p = 0.f;
#pragma unroll
    for (int i = 0; i < 12; ++i) {
      const int e = t + i * MT_THREADS;                                         // W_l[i = r][j = c]
      const int l = e / (HID * HID), r = (e / HID) % HID, c = e % HID;
      wv[i] = wt_g[(l == 0 ? off0 : (l == 1 ? off1 : off2)) + HID + r * HID + c];
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int e = t + i * MT_THREADS, k = e / HID, j = e % HID;               // W_3[j][k] -> rows k, K = j
      w3v[i] = k < OUT ? wt_g[off3 + OUT + j * OUT + k] : 0.f;
    }
    if (t < OUT) b3v = wt_g[off3 + t];
    if (t < HID * 4 && t % 4 < OUT) w3p = wt_g[off3 + OUT + (t / 4) * OUT + t % 4];
#pragma unroll
    for (int i = 0; i < 12; ++i) {
      const int e = t + i * MT_THREADS;
      const int l = e / (HID * HID), r = (e / HID) % HID, c = e % HID;
      const __half w = __float2half_rn(w0 * wv[i]);
      sts16(sbase + Sm::WF + l * Sm::WT + swz64(c, r), w);                      // forward B: rows j, K = i
      if (l > 0) sts16(sbase + Sm::WB + l * Sm::WT + swz64(r, c), w);           // backward B: rows i, K = j
      else if (r >= F) sts16(sbase + Sm::WB + swz64(r - F, c), w);              // layer 0: pe inputs only
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int e = t + i * MT_THREADS, k = e / HID, j = e % HID;
      sts16(sbase + Sm::W3 + swz64(k, j), __float2half_rn(k < OUT ? w3v[i] : 0.f));
    }
    if (WIDE) {
      for (int e = t; e < Sm::WT / 4; e += MT_THREADS) sts32(sbase + Sm::WF0B + e * 4, 0u);      // K 2..15 of the third block stay zero
    }
    // bias tiles: every 4-byte word written exactly once (row j = 64 bytes, the (hi, lo) pair is word 0 of the swizzled
    // chunk 0); loads first, stores after, like the weights above
    constexpr int BW = 3 * Sm::WT / 4 / MT_THREADS;
    uint32_t bw[BW];
#pragma unroll
    for (int i = 0; i < BW; ++i) {
      const int e = t + i * MT_THREADS;
      const int l = e / (Sm::WT / 4), w = e % (Sm::WT / 4), j = w >> 4, ww = w & 15;
      bw[i] = 0u;
      // 34 inputs: the bias block of the A operand is (x32, x33, 1, 1, 0 ..) for all three layers, so the pair sits in K
      // elements 2, 3 (word 1) and K elements 0, 1 stay zero
      if (ww == ((((j >> 1) & 3) << 2) | (WIDE ? 1 : 0)) && !(WIDE && l == 0)) {
        const float b = w0 * wt_g[(l == 0 ? off0 : (l == 1 ? off1 : off2)) + j];
        const __half hi = __float2half_rn(b), lo = __float2half_rn(b - __half2float(hi));
        bw[i] = (uint32_t)__half_as_ushort(hi) | ((uint32_t)__half_as_ushort(lo) << 16);
      }
    }
#pragma unroll
    for (int i = 0; i < BW; ++i) sts32(sbase + Sm::WBI + (t + i * MT_THREADS) * 4, bw[i]);
    if (t < 4) plain[96 + t] = b3v;
    if (t < HID * 4) plain[128 + t] = w3p;
    if (MODE != 0) {
      // the shared second atom of the A operands: per pixel row, feature 32 = 1.0 (chunk 0 of the swizzled row), rest 0
      if (t < 16 * 16) {
        const int k = t >> 4, w = t & 15;
        sts32(sbase + Sm::ONES + k * 64 + w * 4, w == (((k >> 1) & 3) << 2) ? 0x00003c00u : 0u);
      }
      for (int e = t; e < 2 * 1024; e += MT_THREADS) sts32(sbase + (e / 1024) * Sm::SLOT + Sm::DZ3 + (e % 1024) * 4, 0u);
    }
  }
  PROF(2);
  if (WIDE) {
    __syncthreads();                                  // the zero fill of the third K block is complete
    const int t = threadIdx.x;
    if (t < (IN - 32) * HID) {
      const int r = 32 + t / HID, c = t % HID;
      const __half w = __float2half_rn(w0 * wt_g[off0 + HID + r * HID + c]);
      sts16(sbase + Sm::WF0B + swz64(c, r - 32), w);                            // forward B: rows j, K = i - 32
      sts16(sbase + Sm::WB + swz64(r - F, c), w);                               // backward B (layer 0): pe input r - F
    }
    if (t < HID) {                                                              // layer 0's bias: K elements 2, 3 of the same block
      const float b = w0 * wt_g[off0 + t];
      const __half hi = __float2half_rn(b), lo = __float2half_rn(b - __half2float(hi));
      sts32(sbase + Sm::WF0B + swz64(t, 2), (uint32_t)__half_as_ushort(hi) | ((uint32_t)__half_as_ushort(lo) << 16));
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  PROF(3);
  const uint32_t tmem_base = *tmem_slot;

  const uint32_t tm = tmem_base + ((uint32_t)(q * 32) << 16);
  const uint32_t R0 = (uint32_t)(64 * g), R1 = R0 + 32, RB = TM_ONES + (uint32_t)(8 * g);
  const int so = g * Sm::SLOT;

  if (MODE != 0) {
    // the weight-gradient accumulators are shared by both groups: start them at zero, always accumulate
    if (g == 0) {
      uint32_t z[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) z[i] = 0u;
      tmem_st32(tm + TM_DW, z);
      tmem_st32(tm + TM_DW + 32, z);
      tmem_st32(tm + TM_DW + 64, z);
      tmem_st32(tm + TM_DW + 80, z);      // columns [208, 240): overlaps the previous store, covers DW3
      tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }

  // ---- MMA issue.  The stage's MMAs are issued by ONE warp of the group, right after it has published
  // its own share: the whole warp waits for the group's 128 arrivals (converged), one elected lane
  // issues and commits.  Issue blocks while the tensor pipe is busy, so the group's warps take turns.
  // Consecutive stages of a tile are ordered by completion (commit -> wait -> publish); the groups
  // only share the weight-gradient accumulators, where the in-order pipe makes every D += A B atomic.
  uint32_t ph_ready = 0u;
  // chain product over K = 32: two K = 16 steps; the A operand sits in the first 16 columns of its region as packed fp16
  auto chain = [&](uint32_t d_col, uint32_t a_col, int b_off, uint32_t idesc) {
    const uint64_t db = smem_desc_sw64(sbase + b_off);
#pragma unroll
    for (int k = 0; k < 2; ++k)
      umma_f16_ts(tmem_base + d_col, tmem_base + a_col + (uint32_t)(k * 8), db + (uint64_t)(k * 2), idesc, k ? 1u : 0u);
  };
  // third K = 16 block of a sine layer's product: the group's bias block (RB) against the layer's bias tile
  auto bias = [&](uint32_t d_col, int b_off, uint32_t idesc) {
    umma_f16_ts(tmem_base + d_col, tmem_base + RB, smem_desc_sw64(sbase + b_off), idesc, 1u);
  };
  // dW = X^T dZ over the tile's 128 pixels: both operands pixel-major (MN-major), eight K = 16 steps of 1024 B.  The
  // second 32-feature atom of A is the shared 1 KB ones block (its distance shrinks with every step) or, for the 34-input
  // dW0, a per-tile block that advances with the first
  auto wgrad = [&](uint32_t d_col, int x_off, int dz_off, uint32_t idesc, int atom2_off = Sm::ONES, bool atom2_fixed = true) {
    const uint32_t xa = sbase + x_off;
    const uint64_t da = smem_desc_pm(xa, sbase + atom2_off - xa), db = smem_desc_pm(sbase + dz_off, 16);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint64_t step = (uint64_t)(k * 64);
      umma_f16(tmem_base + d_col, atom2_fixed ? da + step - (step << 16) : da + step, db + step, idesc, 1u);
    }
  };
  // dW3 = X3^T dy: A pixel-major, B = dy^T K-major [2 K blocks][16][64 px]
  auto wgrad_dy = [&](uint32_t d_col, int x_off, int dy_off, uint32_t idesc) {
    const uint32_t xa = sbase + x_off;
    const uint64_t da = smem_desc_pm(xa, sbase + Sm::ONES - xa);
#pragma unroll
    for (int kb = 0; kb < 2; ++kb) {
      const uint64_t db = smem_desc_sw128(sbase + dy_off + kb * 2048);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t step = (uint64_t)((kb * 4 + k) * 64);
        umma_f16(tmem_base + d_col, da + step - (step << 16), db + (uint64_t)(k * 2), idesc, 1u);
      }
    }
  };
  // stages of one tile: 0-2 sine layers, 3 output layer, 4-6 backward.  Stage 6 also carries stage 0 of the group's NEXT
  // tile (`next`): its inputs are already in R0[0, 16) and the bias block, and nothing it touches is read by stage 6
  // except R1, which the in-order pipe has consumed (A of d pe) before the next tile's Z0 overwrites it.
  auto issue = [&](int stage, int tile = 0, bool next = false) {
    const bool mine = q == (stage & 3);
    if (mine) {
      mbar_wait(&bar_ready[g], ph_ready);
      tc_fence_after();
    }
    if (mine && elect_one()) {
      const uint32_t t32 = idesc_f16_m128(32), t16 = idesc_f16_m128(16);
      const uint32_t h32 = idesc_f16(32) | IDESC_A_MN | IDESC_B_MN, h16 = idesc_f16(16) | IDESC_A_MN;
      switch (stage) {
        case 0:                                                                 // Z0 = X0 W0 + b0
          chain(R1, R0, Sm::WF, t32);
          bias(R1, WIDE ? Sm::WF0B : Sm::WBI, t32);
          break;
        case 1:                                                                 // Z1 = X1 W1 + b1
          chain(R0, R1, Sm::WF + Sm::WT, t32);
          bias(R0, Sm::WBI + Sm::WT, t32);
          break;
        case 2:                                                                 // Z2 = X2 W2 + b2
          chain(R1, R0, Sm::WF + 2 * Sm::WT, t32);
          bias(R1, Sm::WBI + 2 * Sm::WT, t32);
          break;
        case 3: chain(R0, R1, Sm::W3, t16); break;                              // y  = X3 W3
        // The two groups add into the same weight-gradient accumulators.  The tensor pipe runs MMAs in the order
        // they are issued, so the tiles take turns (tile t after tile t - 1, per stage): the fp32 summation order,
        // and with it every bit of the gradients, does not depend on how the groups happen to interleave.
        case 4:
          chain(R0, R1, Sm::WB + 2 * Sm::WT, t32);                                    // dX2 = dZ2 W2^T
#ifndef RCB_NO_TURN
          while (wg_turn[0] != tile) {}
#endif
          if (!DW3_LATE) wgrad_dy(TM_DW + 96, so + Sm::XT3, so + Sm::DZ3, h16);   // dW3 = X3^T dy
          wgrad(TM_DW + 64, so + Sm::XT2, so + Sm::DZT, h32);                   // dW2 = X2^T dZ2
          wg_turn[0] = tile + 1;
          break;
        case 5:
          chain(R1, R0, Sm::WB + Sm::WT, t32);                                    // dX1 = dZ1 W1^T
          if (EARLY5) umma_commit(&bar_mma[g]);
#ifndef RCB_NO_TURN
          while (wg_turn[1] != tile) {}
#endif
          wgrad(TM_DW + 32, so + Sm::XT1, so + Sm::DZT, h32);                   // dW1 = X1^T dZ1
          if (DW3_LATE) wgrad_dy(TM_DW + 96, so + Sm::XT3, so + Sm::DZ3, h16);    // dW3 = X3^T dy
          wg_turn[1] = tile + 1;
          break;
        default:
          chain(R0 + 16u, R1, Sm::WB, t16);                                     // d pe = dZ0 W0[pe rows]^T  -> R0[16, 32)
          if (EARLY6 && next) {
            chain(R1, R0, Sm::WF, t32);
            bias(R1, WIDE ? Sm::WF0B : Sm::WBI, t32);
            umma_commit(&bar_mma[g]);
          }
#ifndef RCB_NO_TURN
          while (wg_turn[2] != tile) {}
#endif
          if (WIDE) wgrad(TM_DW, so + Sm::XT0, so + DZ0, h32, so + Sm::XT2, false);  // dW0 = X0^T dZ0, 34 inputs + bias row
          else wgrad(TM_DW, so + Sm::XT0, so + DZ0, h32);                       // dW0 = X0^T dZ0
          wg_turn[2] = tile + 1;
          if (next && !EARLY6) {                                                // the next tile's Z0 = X0 W0 + b0
            chain(R1, R0, Sm::WF, t32);
            bias(R1, WIDE ? Sm::WF0B : Sm::WBI, t32);
          }
          break;
      }
      if (!(EARLY5 && stage == 5) && !(EARLY6 && stage == 6 && next)) umma_commit(&bar_mma[g]);
    }
    ph_ready ^= 1;
    __syncwarp();
  };

  // feature-major fp16 element of this thread's pixel: K block r / 64, column r % 64
  const uint32_t t_kb = (uint32_t)(q >> 1);
  const int t_col = (q & 1) * 32 + lane;
  // features f0 .. f0+15 of this thread's pixel row of a pixel-major buffer <- v[0..15] (fp16, round to nearest):
  // two 16-byte chunks of the 64-byte row, chunk index XORed with bits 1-2 of the row (64-byte swizzle)
  const uint32_t pm_row = (uint32_t)r * 64u, pm_x = ((uint32_t)r >> 1) & 3u;
  // hp[0..7] = the 16 features f0 .. f0+15 as packed fp16 pairs (the same words that go to TMEM as the chain operand)
  auto store_p16 = [&](uint32_t buf, int f0, const uint32_t* hp) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const uint32_t addr = buf + pm_row + (((uint32_t)((f0 >> 3) + c) ^ pm_x) << 4);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(hp[4 * c]), "r"(hp[4 * c + 1]),
                   "r"(hp[4 * c + 2]), "r"(hp[4 * c + 3]) : "memory");
    }
  };
  // hand this thread's share of the stage's operands to the MMAs.  `fence_smem`: the stage's MMAs read shared memory this
  // thread wrote (through the async proxy); the proxy fence covers every earlier store of the thread, so the forward
  // stages, whose fp16 copies are only read from stage 4 on, publish without it
  auto publish = [&](bool fence_smem) {
    PROF(300);
    tmem_st_wait();
    if (fence_smem) fence_async_smem();
    tc_fence_before();
    mbar_arrive(&bar_ready[g]);
    PROF(301);
  };
  uint32_t ph_mma = 0u;
  auto wait_mma = [&]() {
    PROF(100);
    mbar_wait(&bar_mma[g], ph_mma);
    ph_mma ^= 1;
    tc_fence_after();
    PROF(200);
  };
  PROF(4);
  float sq = 0.f;
  constexpr uint32_t ONES2 = 0x3c003c00u;           // (1.0, 1.0) as packed fp16
  uint32_t x0p[16];                                 // the tile's 32 first inputs as packed fp16 pairs
  uint32_t wide_x = 0u;                             // 34 inputs: (x32, x33) of the tile whose inputs are in TMEM
  uint32_t wide_n = 0u;                             // ... and of the tile whose inputs are in x0p
  // the inputs in `xin` -> x0p (as soon as the loads have landed: 16 live registers instead of 32) ...
  auto pack_x0 = [&]() {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      x0p[i] = (F == 16 && a.pe_half && i >= 8) ? xin[8 + i] : pack_h2(__uint_as_float(xin[2 * i]), __uint_as_float(xin[2 * i + 1]));
    if (WIDE) wide_n = pack_h2(__uint_as_float(xin[32]), __uint_as_float(xin[IN - 1]));
  };
  // ... -> the chain operand R0[0, 16) and (34 inputs) the per-tile bias block: inputs 32, 33, then the two ones
  auto put_x0 = [&]() {
    tmem_st16(tm + R0, x0p);
    if (WIDE) {
      wide_x = wide_n;
      const uint32_t b8[8] = {wide_x, ONES2, 0u, 0u, 0u, 0u, 0u, 0u};
      tmem_st8(tm + RB, b8);
    }
  };
  // d pe of the tile whose pixel is gp: 16 columns R0[16, 32), back in true units
  auto read_dpe = [&](int gp) {
    uint32_t acc[16];
    tmem_ld16_issue(tm + R0 + 16u, acc);
    tmem_ld_wait();
    if (gp >= pix) return;
    if (a.d_pe_h != nullptr) {       // fp16, still in the chain's units
      uint32_t h[8];
#pragma unroll
      for (int c = 0; c < 8; ++c)
        h[c] = pack_h2_sat(__uint_as_float(acc[2 * c]), __uint_as_float(acc[2 * c + 1]));
      uint4* dst = reinterpret_cast<uint4*>(stitched ? dpe_item + pe_off(gp) * (NPE * 2) : dpe_item + (uint32_t)gp * (uint32_t)(NPE * 2));
      dst[0] = make_uint4(h[0], h[1], h[2], h[3]);
      dst[1] = make_uint4(h[4], h[5], h[6], h[7]);
    } else {
      float4* dst = reinterpret_cast<float4*>(stitched ? dpe_item + pe_off(gp) * (NPE * 4) : dpe_item + (uint32_t)gp * (uint32_t)(NPE * 4));
#pragma unroll
      for (int c = 0; c < 4; ++c)
        dst[c] = make_float4(unscale * __uint_as_float(acc[4 * c]), unscale * __uint_as_float(acc[4 * c + 1]),
                             unscale * __uint_as_float(acc[4 * c + 2]), unscale * __uint_as_float(acc[4 * c + 3]));
    }
  };

  if (g < ntiles) {
    // the group's bias block (written once unless it carries inputs) and its first tile's inputs
    if (!WIDE) {
      const uint32_t ones8[8] = {ONES2, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
      tmem_st8(tm + RB, ones8);
    }
    pack_x0();
    put_x0();
    publish(false);
    issue(0);
  }
  for (int tile = g; tile < ntiles; tile += 2) {
    const int gp = tile * 128 + r;
    const bool valid = gp < pix;
    const bool has_next = tile + 2 < ntiles;
    uint32_t cs[3][16];                           // cos(.) of the three sine layers, packed half2
    float dy[OUT];                                // targets / output gradient first, then dy in the chain's units
    // ---- Z0 of this tile is in flight (issued alone or with the previous tile's last stage, whose d pe comes back here)
    wait_mma();
    if (MODE != 0) {
      if (tile > g) read_dpe(gp - 256);
      // X0^T for dW0: its buffer was an operand of the previous tile's last stage until the wait above
      if (!EARLY6) {
        store_p16(sbase + so + Sm::XT0, 0, x0p);
        store_p16(sbase + so + Sm::XT0, 16, x0p + 8);
      }
    }
    // ---- three sine layers: X_{l+1} = sin(acc + b') back into the accumulator's columns, fp16 copy of X_{l+1}^T
#pragma unroll
    for (int l = 0; l < 3; ++l) {
      const uint32_t reg = tm + ((l & 1) ? R0 : R1);                     // Z0 -> R1, Z1 -> R0, Z2 -> R1
      if (l > 0) wait_mma();
      if (EARLY6 && l == 1 && MODE != 0) {        // the previous tile's dW0 has retired with this stage's commit
        store_p16(sbase + so + Sm::XT0, 0, x0p);
        store_p16(sbase + so + Sm::XT0, 16, x0p + 8);
      }
      uint32_t acc[32];
      tmem_ld32_issue(reg, acc);
      tmem_ld_wait();
      uint32_t xp[16];                                                   // X_{l+1} as packed fp16 pairs
#pragma unroll
      for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          const float z0 = __uint_as_float(acc[16 * h + j]), z1 = __uint_as_float(acc[16 * h + j + 1]);     // bias included
          xp[(16 * h + j) >> 1] = pack_h2(__sinf(z0), __sinf(z1));
          cs[l][(16 * h + j) >> 1] = pack_h2(__cosf(z0), __cosf(z1));
        }
        if (MODE != 0) store_p16(sbase + so + (l == 0 ? Sm::XT1 : (l == 1 ? Sm::XT2 : Sm::XT3)), 16 * h, xp + 8 * h);
      }
      tmem_st16(reg, xp);
      if (l == 2 && MODE != 0) {
        // the pixel's targets (mode 1) or output gradient (mode 2): in flight under the output layer's round trip
        const float* src = MODE == 1 ? a.y + ((int64_t)row_item * pix + (valid ? gp : 0)) * OUT
                                     : a.dy + ((int64_t)item * pix + (valid ? gp : 0)) * OUT;
#pragma unroll
        for (int k = 0; k < OUT; ++k) {
          uint32_t t;
          ldg4_if(valid, src + k, t);
          dy[k] = __uint_as_float(t);
        }
      }
      publish(false);
      issue(l + 1);
    }
    // ---- output layer (on the tensor core), loss and dy; dZ2 = (dy W3^T) * cos in place of X3; dy^T, dZ2^T
    wait_mma();
    {
      uint32_t yv[16];
      tmem_ld16_issue(tm + R0, yv);
      tmem_ld_wait();
      if (MODE == 0) {
        if (valid)
#pragma unroll
          for (int k = 0; k < OUT; ++k) a.y_pred[((int64_t)item * pix + gp) * OUT + k] = __uint_as_float(yv[k]) + plain[96 + k];
      } else {
#pragma unroll
        for (int k = 0; k < OUT; ++k) {
          if (MODE == 1) {
            const float rr = valid ? __uint_as_float(yv[k]) + plain[96 + k] - dy[k] : 0.f;
            sq = fmaf(rr, rr, sq);
            dy[k] = rr;
          } else {
            dy[k] *= gscale;
          }
        }
      }
    }
    if (MODE == 0) {
      if (has_next) {
        load_x0(gp + 256, xin);
        pack_x0();
        put_x0();
        publish(false);
        issue(0);
      }
      continue;
    }
#pragma unroll
    for (int k = 0; k < OUT; ++k) sts16(sbase + so + Sm::DZ3 + t_kb * 2048 + swz2(k, t_col), __float2half_rn(dy[k]));
    {
      uint32_t dzp[16];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          const float4 wa = *(const float4*)(plain + 128 + (16 * h + j) * 4);
          const float4 wb = *(const float4*)(plain + 128 + (16 * h + j + 1) * 4);
          float va = dy[0] * wa.x, vb = dy[0] * wb.x;
          if (OUT > 1) { va = fmaf(dy[1], wa.y, va); vb = fmaf(dy[1], wb.y, vb); }
          if (OUT > 2) { va = fmaf(dy[2], wa.z, va); vb = fmaf(dy[2], wb.z, vb); }
          dzp[(16 * h + j) >> 1] = mul_h2(va, vb, cs[2][(16 * h + j) >> 1]);
        }
        store_p16(sbase + so + Sm::DZT, 16 * h, dzp + 8 * h);
      }
      tmem_st16(tm + R1, dzp);                                           // A operand of dX2
    }
    publish(true);                                                       // covers X_0 .. X_3^T and dy^T as well
    issue(4, tile);
    // ---- dZ1 (from R0, in place), dZ0 (from R1, in place) = data gradient * cos; the next tile's inputs join the last stage
#pragma unroll
    for (int l = 1; l >= 0; --l) {
      const uint32_t reg = tm + (l == 1 ? R0 : R1);
      // the next tile's inputs: requested a whole backward stage before they are packed (one stage later they still cost
      // 4 % of the warp time in front of the first pack; yet another stage earlier the 32 extra live registers cost more
      // than the latency).  Unconditional (a pixel index past the end loads nothing) so that neither xin nor x0p is live
      // across the rest of the loop body
      if (l == 1) load_x0(has_next ? gp + 256 : pix, xin);
      wait_mma();
      // two passes of 16 columns: 16 live accumulator registers while the next tile's 32 raw inputs are still in flight
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t acc[16], dzp[8];
        tmem_ld16_issue(reg + (uint32_t)(16 * h), acc);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; j += 2)
          dzp[j >> 1] = mul_h2(__uint_as_float(acc[j]), __uint_as_float(acc[j + 1]), cs[l][(16 * h + j) >> 1]);
        store_p16(sbase + so + (l == 0 ? DZ0 : Sm::DZT), 16 * h, dzp);
        tmem_st8(reg + (uint32_t)(8 * h), dzp);
      }
      if (l == 0) {
        pack_x0();
        if (WIDE) {
          // second atom of dW0's A operand: [x32, x33, 1, 0, ...] -- accumulator rows 32, 33 = the last two inputs, row 34 = bias
          // (over the dead X_2^T)
          const uint32_t row = sbase + so + Sm::XT2 + pm_row;
#pragma unroll
          for (int c = 0; c < 4; ++c)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + (((uint32_t)c ^ pm_x) << 4)),
                         "r"(c == 0 ? wide_x : 0u), "r"(c == 0 ? 0x00003c00u : 0u), "r"(0u), "r"(0u) : "memory");
        }
        // R0[0, 16) (dZ1, consumed by stage 5) and the bias block (last read by stage 2) are free: the next tile's inputs
        if (has_next) put_x0();
      }
      publish(true);
      issue(l == 1 ? 5 : 6, tile, has_next);
    }
  }
  if (MODE != 0 && g < ntiles) {                   // d pe of the group's last tile
    wait_mma();
    read_dpe((ntiles - 1 - ((ntiles - 1 - g) & 1)) * 128 + r);
  }

  PROF(5);
  if (MODE != 0) {
    // both groups' MMAs have retired once every thread is here
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    PROF(6);
    // ---- weight gradients (M = 64 accumulators): input feature i < 16 in TMEM lane i, 16 <= i < 32 in lane
    //      32 + (i - 16), the ones row (bias) in lane 64; group 0's warps take columns 0-15, group 1's 16-31
    float* gw = a.d_wt + (int64_t)item * a.ld_w;
    __half* gwh = a.d_wt_h ? reinterpret_cast<__half*>(a.d_wt_h) + (int64_t)item * a.ld_wh : nullptr;
    auto to_h = [](float v) { return fminf(fmaxf(v, -65504.f), 65504.f); };      // clamp: an overflow must not turn into inf
    const int j0 = g * 16;
    if (q <= 2) {
      // all four accumulator reads first, one wait: their latencies overlap instead of adding up
      uint32_t acc_l[3][16], acc3[16];
#pragma unroll
      for (int l = 0; l < 3; ++l) tmem_ld16_issue(tm + TM_DW + (uint32_t)(l * 32 + j0), acc_l[l]);
      if (g == 0) tmem_ld16_issue(tm + TM_DW + 96, acc3);
      tmem_ld_wait();
      const float sc = w0 * unscale * (gwh ? a.d_wt_h_scale : 1.f);
      const float sc3 = unscale * (gwh ? a.d_wt_h_scale : 1.f);
      const bool wrow_n = q < 2 && lane < 16, brow_n = q == 2 && lane == 0;
      const int irow_n = q * 16 + lane;
#pragma unroll
      for (int l = 0; l < 3; ++l) {
        const uint32_t (&acc)[16] = acc_l[l];
        const int off = l == 0 ? off0 : (l == 1 ? off1 : off2);
        // 34 inputs: layer 0 has two more weight rows (accumulator rows 32, 33 = lanes 64, 65) and its bias in row 34
        const bool wide0 = WIDE && l == 0;
        const bool wrow = wrow_n || (wide0 && q == 2 && lane < IN - 32);
        const bool brow = wide0 ? (q == 2 && lane == IN - 32) : brow_n;
        const int irow = (wide0 && q == 2) ? 32 + lane : irow_n;
        const int didx = wrow ? off + HID + irow * HID + j0 : off + j0;
        float* dst = gw + didx;
        if ((wrow || brow) && gwh) {
          uint32_t hv[8];
#pragma unroll
          for (int c = 0; c < 8; ++c)
            hv[c] = pack_h2_sat(sc * __uint_as_float(acc[2 * c]), sc * __uint_as_float(acc[2 * c + 1]));
          *reinterpret_cast<uint4*>(gwh + didx) = make_uint4(hv[0], hv[1], hv[2], hv[3]);
          *reinterpret_cast<uint4*>(gwh + didx + 8) = make_uint4(hv[4], hv[5], hv[6], hv[7]);
        } else if (wrow || brow) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
            *(float4*)(dst + c * 4) = make_float4(sc * __uint_as_float(acc[c * 4]), sc * __uint_as_float(acc[c * 4 + 1]),
                                                  sc * __uint_as_float(acc[c * 4 + 2]), sc * __uint_as_float(acc[c * 4 + 3]));
        }
      }
      if (g == 0) {
        const uint32_t (&acc)[16] = acc3;
        const bool wrow = wrow_n, brow = brow_n;
        const int irow = irow_n;
        if (wrow) {
#pragma unroll
          for (int k = 0; k < OUT; ++k) {
            if (gwh) gwh[off3 + OUT + irow * OUT + k] = __float2half_rn(to_h(sc3 * __uint_as_float(acc[k])));
            else gw[off3 + OUT + irow * OUT + k] = unscale * __uint_as_float(acc[k]);
          }
        } else if (brow) {
#pragma unroll
          for (int k = 0; k < OUT; ++k) {
            if (gwh) gwh[off3 + k] = __float2half_rn(to_h(sc3 * __uint_as_float(acc[k])));
            else gw[off3 + k] = unscale * __uint_as_float(acc[k]);
          }
        }
      }
    }
    if (MODE == 1) {
      sq = warp_sum(sq);
      if (lane == 0) plain[104 + warp] = sq;
    }
  }
  PROF(7);
  tc_fence_before();
  __syncthreads();
  PROF(8);
  if (MODE == 1 && threadIdx.x == 0)
    a.sqerr[item] = ((plain[104] + plain[105]) + (plain[106] + plain[107])) + ((plain[108] + plain[109]) + (plain[110] + plain[111]));
  if (warp == 0) tmem_dealloc(tmem_base, TM_COLS);
}

template <int OUT, int F, int D>
static int launch(const rcb_mlp_args* a, cudaStream_t st) {
#define RCB_MT_LAUNCH(MODE)                                                                                   \
  do {                                                                                                        \
    cudaError_t e = cudaFuncSetAttribute(mlp_tc_kernel<OUT, MODE, F, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         Sm::TOTAL);                                                          \
    if (e != cudaSuccess) { set_error("rcb_mlp_tc: smem opt-in failed: %s", cudaGetErrorString(e)); return -1; } \
    mlp_tc_kernel<OUT, MODE, F, D><<<a->items, MT_THREADS, Sm::TOTAL, st>>>(*a);                             \
  } while (0)
  if (a->mode == 0) RCB_MT_LAUNCH(0);
  else if (a->mode == 1) RCB_MT_LAUNCH(1);
  else RCB_MT_LAUNCH(2);
#undef RCB_MT_LAUNCH
  RCB_CHECK_LAUNCH("rcb_mlp_tc");
  return 0;
}

}  // namespace mlp
}  // namespace rcb

using namespace rcb;

#ifdef RCB_MLP_PROFILE
extern "C" int rcb_mlp_prof_read(long long* host) {
  return (int)cudaMemcpyFromSymbol(host, rcb::mlp::rcb_prof_buf, sizeof(long long) * 4 * 1024);
}
#endif

namespace rcb {
// one thread per (coordinate index, frequency); fp32 arithmetic in the reference's order, accurate sincosf
__global__ void fourier_table_kernel(float* __restrict__ tab, int size, int nf, float w0, float w1, float w2, float w3, float w4,
                                     float w5, float w6, float w7) {
  const float w[8] = {w0, w1, w2, w3, w4, w5, w6, w7};
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= size * nf) return;
  const int i = e / nf, j = e - i * nf;
  const float c = __fadd_rn(-1.0f, __fmul_rn(2.0f, __fdiv_rn(0.5f + (float)i, (float)size)));   // utils.py:265-284
  const float t = __fmul_rn(__fmul_rn(c, w[j]), 3.14159274101257324f);                         // (c * w) * float32(pi)
  float sn, cs;
  sincosf(t, &sn, &cs);
  tab[i * 2 * nf + j] = cs;
  tab[i * 2 * nf + nf + j] = sn;
}
}  // namespace rcb

extern "C" int rcb_fourier_table(float* tab, int size, const float* freq, int n_freq, rcb_stream_t stream) {
  RCB_CHECK_ARG(tab && freq && size > 0 && n_freq > 0 && n_freq <= 8, "rcb_fourier_table: bad arguments");
  float w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int j = 0; j < n_freq; ++j) w[j] = freq[j];
  const int total = size * n_freq;
  rcb::fourier_table_kernel<<<(total + 127) / 128, 128, 0, (cudaStream_t)stream>>>(tab, size, n_freq, w[0], w[1], w[2], w[3], w[4],
                                                                                  w[5], w[6], w[7]);
  RCB_CHECK_LAUNCH("rcb_fourier_table");
  return 0;
}

extern "C" int rcb_mlp_tc(const rcb_mlp_args* a, rcb_stream_t stream) {
  RCB_CHECK_ARG(a != nullptr, "rcb_mlp_tc: null args");
  RCB_CHECK_ARG(a->items > 0 && a->S > 0 && a->pix > 0, "rcb_mlp_tc: empty problem");
  RCB_CHECK_ARG(a->mode >= 0 && a->mode <= 2, "rcb_mlp_tc: bad mode %d", a->mode);
  RCB_CHECK_ARG(a->wt && (a->xt || a->x_tab) && a->pe, "rcb_mlp_tc: null input");
  RCB_CHECK_ARG(a->n_f == 16 || a->n_f == 18, "rcb_mlp_tc: built for 16 or 18 Fourier features (+ 16 pe); use rcb_mlp for other shapes");
  RCB_CHECK_ARG(!a->pe_half || a->n_f == 16, "rcb_mlp_tc: fp16 positional encodings need the 32-input form");
  RCB_CHECK_ARG(a->mode != 0 || a->y_pred, "rcb_mlp_tc: mode 0 needs y_pred");
  RCB_CHECK_ARG(a->mode != 1 || (a->y && a->sqerr && a->coef > 0.f), "rcb_mlp_tc: mode 1 needs y, sqerr and coef > 0");
  RCB_CHECK_ARG(a->mode != 2 || a->dy, "rcb_mlp_tc: mode 2 needs dy");
  RCB_CHECK_ARG(a->mode == 0 || ((a->d_pe || a->d_pe_h) && a->d_wt), "rcb_mlp_tc: backward needs d_pe (or d_pe_h) and d_wt");
  RCB_CHECK_ARG(a->ld_w % 4 == 0, "rcb_mlp_tc: ld_w must be a multiple of 4");
  RCB_CHECK_ARG(!a->pe_base || (a->ph > 0 && a->pw > 0), "rcb_mlp_tc: stitched addressing needs the patch extent");
  cudaStream_t st = (cudaStream_t)stream;
  // D only matters with generated inputs: the 16-feature forms exist for 1-D (8 frequencies) and 2-D (4 per axis)
  // signals, the 18-feature one for 3-D (3 per axis); without x_tab any D serves
  int D = a->n_f == 18 ? 3 : 2;
  if (a->x_tab) {
    RCB_CHECK_ARG(a->x_axes >= 1 && a->x_axes <= 3 && 2 * a->x_axes * a->x_nfreq == a->n_f, "rcb_mlp_tc: x_tab does not describe %d features", a->n_f);
    RCB_CHECK_ARG((a->n_f == 18) == (a->x_axes == 3) && (((uintptr_t)a->x_tab) & 15) == 0, "rcb_mlp_tc: unsupported generated-input shape");
    int64_t prod = 1;
    for (int i = 0; i < a->x_axes; ++i) {
      RCB_CHECK_ARG(a->x_size[i] > 0 && a->x_off[i] >= 0 && a->x_off[i] % 4 == 0, "rcb_mlp_tc: bad x_tab axis %d", i);
      prod *= a->x_size[i];
    }
    RCB_CHECK_ARG(prod == a->pix, "rcb_mlp_tc: x_size does not multiply to pix");
    D = a->x_axes;
  }
  if (a->out == 3 && a->n_f == 18) return mlp::launch<3, 18, 3>(a, st);
  if (a->out == 3 && a->n_f == 16) return D == 1 ? mlp::launch<3, 16, 1>(a, st) : mlp::launch<3, 16, 2>(a, st);
  if (a->out == 1 && a->n_f == 16) return D == 1 ? mlp::launch<1, 16, 1>(a, st) : mlp::launch<1, 16, 2>(a, st);
  set_error("rcb_mlp_tc: unsupported shape (out %d, %d Fourier features)", a->out, a->n_f);
  return -2;
}
