// Tensor-core fused per-item SIREN MLP (forward + squared error + backward), second design:
// activations never leave the SM.  One CTA per (row, MC sample) item, TWO CTAs per SM so that
// one item's epilogue overlaps the other's MMA round trip.
//
// Every contraction of a 128-pixel tile is a tcgen05 TF32 MMA with fp32 accumulation in TMEM:
//   chain   Z_l = X_l W_l, y = X_3 W_3, dX_l = dZ_l W_l^T       A operand read straight from TMEM
//                                                               (the epilogue writes sin(.) / dZ back
//                                                               with tcgen05.st, no shared-memory hop)
//   wgrad   dW_l = X_l^T dZ_l                                   A/B = feature-major [feature][pixel]
//                                                               copies in shared memory (K = pixels)
// Shared-memory A tiles of the weight-gradient MMAs carry a row of ones after the 32 feature
// rows, so accumulator row 32 is the bias gradient; rows 33..127 of the M=128 MMA read whatever
// follows in shared memory and land in TMEM lanes nobody reads.
// w0 is folded into the staged weights and biases (W' = w0 W, b' = w0 b): the epilogue is
// sin(acc + b'), the stored derivative factor is cos(.), and the chain carries dZ/w0; the
// weight/bias gradients are scaled by w0 once per item when they leave TMEM.
// TF32 rounding (round-to-nearest-away) is an integer add of 0x1000 on the fp32 pattern: the
// tensor core ignores the 13 low mantissa bits.
//
// TMEM (256 columns per CTA): S0 S1 S2 = X0 X1 X2 (later dZ0 dZ1 dZ2), ACC = chain accumulator
// (also holds X3 for the output MMA), DW0..DW3 = weight-gradient accumulators kept over the
// item's tiles, Y = output accumulator.
// Reference semantics: test_model.py:347-355, 624-627; weight layout :269-280.
#include "tc_common.cuh"

namespace rcb {

constexpr int MT_THREADS = 256;        // two threads per pixel row, 16 features each; thread 0 also issues the MMAs
constexpr int MT_EPI = 256;

constexpr uint32_t TM_S0 = 0, TM_S1 = 32, TM_S2 = 64, TM_ACC = 96, TM_DW = 128, TM_Y = 240, TM_COLS = 256;

struct MtSmem {
  static constexpr int XT_KB = 40 * 128;            // one 32-pixel K block: 32 feature rows, ones row, 7 zero rows
  static constexpr int XT_BYTES = 4 * XT_KB;
  static constexpr int XT0 = 0, XT1 = XT_BYTES;     // X_l^T  [4 K blocks][40][32 px]
  static constexpr int DZT = 2 * XT_BYTES;          // dZ_l^T [4 K blocks][32][32 px]
  static constexpr int DZ3 = DZT + 4 * 4096;        // dy^T   [4 K blocks][16][32 px] (rows >= OUT stay zero)
  static constexpr int WF = DZ3 + 4 * 2048;         // 3 forward B tiles  [j][i]
  static constexpr int WB = WF + 3 * 4096;          // 3 backward B tiles [i][j] (layer 0: the 16 pe inputs)
  static constexpr int W3 = WB + 3 * 4096;          // output B tile [16 (OUT used)][32]
  static constexpr int PLAIN = W3 + 2048;           // floats: [0,96) w0*b_l, [96,100) b3, [128,256) W3[j][4], [256,264) scratch
  static constexpr int BAR = PLAIN + 2048;
  static constexpr int TOTAL = BAR + 64 + 1024;     // + slack for the 1024-B alignment of the base
};
static_assert(MtSmem::WF % 1024 == 0 && MtSmem::WB % 1024 == 0 && MtSmem::W3 % 1024 == 0, "swizzled tiles need 1024-B alignment");
static_assert(MtSmem::XT1 + 3 * MtSmem::XT_KB + 128 * 128 <= MtSmem::BAR, "M=128 reads past the last A tile must stay in the allocation");

// byte offset of element (row, col) in a tile of 128-byte rows with the 128-byte swizzle
__device__ __forceinline__ uint32_t swz(int row, int col) {
  return (uint32_t)(row * 128 + ((((col >> 2) ^ (row & 7)) << 4) | ((col & 3) << 2)));
}
__device__ __forceinline__ uint32_t rnd_tf32(float x) { return __float_as_uint(x) + 0x1000u; }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
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
// D[tmem] (+)= A[tmem: lane = row, one fp32 column per K element] * B[smem descriptor]
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}

#ifdef RCB_MLP_PROFILE
__device__ long long rcb_prof_buf[4 * 512];
#define PROF(id)                                                                         \
  do {                                                                                   \
    if (prof_slot >= 0 && prof_n < 255) {                                                \
      rcb_prof_buf[prof_slot * 512 + 2 * prof_n] = (id);                                 \
      rcb_prof_buf[prof_slot * 512 + 2 * prof_n + 1] = clock64();                        \
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
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  float* plain = (float*)(smem + MtSmem::PLAIN);
  uint64_t* bar_ready = (uint64_t*)(smem + MtSmem::BAR);   // epilogue -> MMA (256 arrivals)
  uint64_t* bar_mma = bar_ready + 1;                       // MMA -> epilogue (commit of the stage's MMAs)
  uint32_t* tmem_slot = (uint32_t*)(bar_ready + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x;
  const int row_item = item / a.S;
  const int pix = a.pix;
  const int ntiles = (pix + 127) / 128;
  const float w0 = a.w0;
  const float* wt_g = a.wt + (int64_t)item * a.ld_w;

  if (threadIdx.x == 0) {
    mbar_init(bar_ready, MT_EPI);
    mbar_init(bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(tmem_slot, TM_COLS);
  // ---- stage the item's weights (w0 folded in), the constant rows of the A tiles and the zero
  //      padding of dy^T
  {
    const int t = threadIdx.x;
    for (int e = t; e < 3 * HID * HID; e += MT_EPI) {
      const int l = e / (HID * HID), r = (e / HID) % HID, c = e % HID;          // W_l[i = r][j = c]
      const int off = l == 0 ? off0 : (l == 1 ? off1 : off2);
      const uint32_t w = rnd_tf32(w0 * wt_g[off + HID + r * HID + c]);
      sts32(sbase + MtSmem::WF + l * 4096 + swz(c, r), w);                      // forward B: rows j, K = i
      if (l > 0) sts32(sbase + MtSmem::WB + l * 4096 + swz(r, c), w);           // backward B: rows i, K = j
      else if (r >= F) sts32(sbase + MtSmem::WB + swz(r - F, c), w);            // layer 0: pe inputs only
    }
    for (int e = t; e < 16 * HID; e += MT_EPI) {
      const int k = e / HID, j = e % HID;                                       // W_3[j][k] -> rows k, K = j
      sts32(sbase + MtSmem::W3 + swz(k, j), k < OUT ? rnd_tf32(wt_g[off3 + OUT + j * OUT + k]) : 0u);
    }
    for (int e = t; e < 3 * HID; e += MT_EPI) plain[e] = w0 * wt_g[(e / HID == 0 ? off0 : (e / HID == 1 ? off1 : off2)) + e % HID];
    if (t < 4) plain[96 + t] = t < OUT ? wt_g[off3 + t] : 0.f;
    for (int e = t; e < HID * 4; e += MT_EPI) plain[128 + e] = (e % 4 < OUT) ? wt_g[off3 + OUT + (e / 4) * OUT + e % 4] : 0.f;
    if (MODE != 0) {
      for (int e = t; e < 2 * 4 * 8 * 32; e += MT_EPI) {                        // rows 32..39 of every K block
        const int b = e / 1024, kb = (e / 256) % 4, rr = 32 + (e / 32) % 8, c = e % 32;
        sts32(sbase + (b ? MtSmem::XT1 : MtSmem::XT0) + kb * MtSmem::XT_KB + swz(rr, c), rr == 32 ? 0x3f800000u : 0u);
      }
      for (int e = t; e < 4 * 16 * 32; e += MT_EPI) sts32(sbase + MtSmem::DZ3 + e * 4, 0u);
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ---- MMA issue (thread 0 only, right after it has published its own part of the stage)
  uint32_t ph_ready = 0;
  // 128 x n x 32 chain product, A in TMEM columns [a_col, a_col + 32)
  auto chain = [&](uint32_t d_col, uint32_t a_col, int b_off, uint32_t idesc) {
    const uint64_t db = smem_desc_sw128(sbase + b_off);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      umma_tf32_ts(tmem_base + d_col, tmem_base + a_col + (uint32_t)(k * 8), db + (uint64_t)(k * 2), idesc, k ? 1u : 0u);
  };
  // d[feature i (+ ones row)][j] += sum over the tile's 128 pixels of X^T[i][px] dZ^T[j][px]
  auto wgrad = [&](uint32_t d_col, int xt_off, int dz_off, int dz_kb, uint32_t idesc, bool first) {
#pragma unroll
    for (int kb = 0; kb < 4; ++kb) {
      const uint64_t da = smem_desc_sw128(sbase + xt_off + kb * MtSmem::XT_KB);
      const uint64_t db = smem_desc_sw128(sbase + dz_off + kb * dz_kb);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_tf32(tmem_base + d_col, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (first && kb == 0 && k == 0) ? 0u : 1u);
    }
  };
  // stages of one tile: 0-2 sine layers, 3 output layer, 4-6 backward
  auto issue = [&](int stage, bool first) {
    if (threadIdx.x == 0) {
      const uint32_t id32 = idesc_tf32(32), id16 = idesc_tf32(16);
      mbar_wait(bar_ready, ph_ready);
      tc_fence_after();
      PROF(100 + stage);
      switch (stage) {
        case 0: chain(TM_ACC, TM_S0, MtSmem::WF, id32); break;
        case 1: chain(TM_ACC, TM_S1, MtSmem::WF + 4096, id32); break;
        case 2: chain(TM_ACC, TM_S2, MtSmem::WF + 8192, id32); break;
        case 3: chain(TM_Y, TM_ACC, MtSmem::W3, id16); break;
        case 4:
          wgrad(TM_DW + 96, MtSmem::XT0, MtSmem::DZ3, 2048, id16, first);  // dW3 = X3^T dy
          wgrad(TM_DW + 64, MtSmem::XT1, MtSmem::DZT, 4096, id32, first);  // dW2 = X2^T dZ2
          chain(TM_ACC, TM_S2, MtSmem::WB + 8192, id32);                   // dX2 = dZ2 W2^T
          break;
        case 5:
          wgrad(TM_DW + 32, MtSmem::XT0, MtSmem::DZT, 4096, id32, first);  // dW1 = X1^T dZ1
          chain(TM_ACC, TM_S1, MtSmem::WB + 4096, id32);                   // dX1 = dZ1 W1^T
          break;
        default:
          wgrad(TM_DW, MtSmem::XT1, MtSmem::DZT, 4096, id32, first);       // dW0 = X0^T dZ0
          chain(TM_ACC, TM_S0, MtSmem::WB, id16);                          // d pe = dZ0 W0[pe rows]^T
          break;
      }
      umma_commit(bar_mma);
      PROF(110 + stage);
    }
    ph_ready ^= 1;
    __syncwarp();
  };

  const int q = warp & 3;                 // TMEM lane quarter = 32-pixel K block of the weight-gradient tiles
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
  // transposed (feature-major) element of this thread's pixel: feature f of K block q
  const uint32_t xt0_addr = sbase + MtSmem::XT0 + q * MtSmem::XT_KB;
  const uint32_t xt1_addr = sbase + MtSmem::XT1 + q * MtSmem::XT_KB;
  const uint32_t dzt_addr = sbase + MtSmem::DZT + q * 4096;
  const uint32_t dz3_addr = sbase + MtSmem::DZ3 + q * 2048;
  auto store_t = [&](uint32_t base, const uint32_t (&v)[16]) {
#pragma unroll
    for (int j = 0; j < 16; ++j) sts32(base + swz(j0 + j, lane), v[j]);
  };
  // this thread's 16 input features of pixel gp: Fourier features (half 0) or positional encodings (half 1);
  // raw fp32 patterns (TF32 rounding happens when they are consumed, so the loads stay in flight)
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
  auto publish = [&](bool smem_written) {        // hand this thread's part of the stage's operands to the MMAs
    tmem_st_wait();
    if (smem_written) fence_async_smem();
    tc_fence_before();
    mbar_arrive(bar_ready);
  };
  uint32_t ph_mma = 0;
  auto wait_mma = [&]() {
    mbar_wait(bar_mma, ph_mma);
    ph_mma ^= 1;
    tc_fence_after();
  };
  float sq = 0.f;
  uint32_t xin[16];
  load_x0(r, xin);

  for (int tile = 0; tile < ntiles; ++tile) {
    const int gp = tile * 128 + r;
    const bool valid = gp < pix;
    const bool first = tile == 0;
    // ---- X0 -> TMEM
    PROF(0);
#pragma unroll
    for (int i = 0; i < 16; ++i) xin[i] += 0x1000u;                // TF32 round-to-nearest of the inputs
    tmem_st16(tm + TM_S0 + j0, xin);
    publish(false);
    PROF(1);
    issue(0, first);
    // ---- three sine layers: X_{l+1} = sin(acc + b'), cs_l = cos(acc + b')
    float cs[3][16];
#pragma unroll
    for (int l = 0; l < 3; ++l) {
      wait_mma();
      PROF(10 + l);
      uint32_t acc[16];
      tmem_ld16_issue(tm + TM_ACC + j0, acc);
      tmem_ld_wait();
      PROF(20 + l);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float z = __uint_as_float(acc[j]) + plain[l * 32 + j0 + j];
        cs[l][j] = __cosf(z);
        acc[j] = rnd_tf32(__sinf(z));
      }
      tmem_st16(tm + (l == 0 ? TM_S1 : (l == 1 ? TM_S2 : TM_ACC)) + j0, acc);
      PROF(30 + l);
      publish(false);
      PROF(40 + l);
      issue(l + 1, first);
    }
    // ---- output layer (on the tensor core), loss and dy; both halves of a row read the same columns
    wait_mma();
    PROF(50);
    float dy[OUT];
    {
      uint32_t yv[16];
      tmem_ld16_issue(tm + TM_Y, yv);
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
            dy[k] = a.coef * rr;
          } else {
            dy[k] = valid ? __ldg(a.dy + ((int64_t)item * pix + gp) * OUT + k) : 0.f;
          }
        }
      }
    }
    if (MODE == 0) {
      if (tile + 1 < ntiles) load_x0(gp + 128, xin);
      continue;
    }
    // ---- X3^T, dy^T; dZ2 = (dy W3^T) * cos -> TMEM (A of the next data-gradient MMA) and dZ2^T; X2^T
    {
      uint32_t xv[16];
      tmem_ld16_issue(tm + TM_ACC + j0, xv);
      tmem_ld_wait();
      store_t(xt0_addr, xv);
      if (hh == 0) {
#pragma unroll
        for (int k = 0; k < OUT; ++k) sts32(dz3_addr + swz(k, lane), rnd_tf32(dy[k]));
      }
      tmem_ld16_issue(tm + TM_S2 + j0, xv);
      tmem_ld_wait();
      store_t(xt1_addr, xv);
      uint32_t dz[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float4 w4 = *(const float4*)(plain + 128 + (j0 + j) * 4);
        float v = dy[0] * w4.x;
        if (OUT > 1) v = fmaf(dy[1], w4.y, v);
        if (OUT > 2) v = fmaf(dy[2], w4.z, v);
        dz[j] = rnd_tf32(v * cs[2][j]);
      }
      store_t(dzt_addr, dz);
      tmem_st16(tm + TM_S2 + j0, dz);
      PROF(51);
      publish(true);
      PROF(52);
      issue(4, first);
    }
    // ---- dZ1, dZ0: data gradient from the tensor core times cos; X1^T, X0^T
#pragma unroll
    for (int l = 1; l >= 0; --l) {
      wait_mma();
      PROF(60 + l);
      uint32_t acc[16], xv[16];
      tmem_ld16_issue(tm + TM_ACC + j0, acc);
      tmem_ld16_issue(tm + (l == 1 ? TM_S1 : TM_S0) + j0, xv);
      tmem_ld_wait();
      store_t(l == 1 ? xt0_addr : xt1_addr, xv);
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = rnd_tf32(__uint_as_float(acc[j]) * cs[l][j]);
      store_t(dzt_addr, acc);
      tmem_st16(tm + (l == 1 ? TM_S1 : TM_S0) + j0, acc);
      PROF(62 + l);
      publish(true);
      PROF(64 + l);
      issue(l == 1 ? 5 : 6, first);
    }
    // ---- next tile's inputs travel while the last MMAs of this tile run
    if (tile + 1 < ntiles) load_x0(gp + 128, xin);
    // ---- d pe (16 columns: 8 per half); the chain carries dZ / w0
    PROF(70);
    wait_mma();
    PROF(71);
    {
      uint32_t acc[16];
      tmem_ld16_issue(tm + TM_ACC, acc);
      tmem_ld_wait();
      if (valid) {
        float* dst = a.d_pe + (pe_origin + pe_off(gp)) * NPE + hh * 8;
        const int b = hh * 8;
        *(float4*)(dst) = make_float4(__uint_as_float(acc[b]), __uint_as_float(acc[b + 1]), __uint_as_float(acc[b + 2]), __uint_as_float(acc[b + 3]));
        *(float4*)(dst + 4) = make_float4(__uint_as_float(acc[b + 4]), __uint_as_float(acc[b + 5]), __uint_as_float(acc[b + 6]), __uint_as_float(acc[b + 7]));
      }
    }
  }

  if (MODE != 0) {
    // ---- weight gradients: TMEM lane i < 32 = input feature i, lane 32 = the ones row (bias)
    float* g = a.d_wt + (int64_t)item * a.ld_w;
    if (q == 0 || q == 1) {
      uint32_t acc[16];
#pragma unroll
      for (int l = 0; l < 3; ++l) {
        tmem_ld16_issue(tm + TM_DW + (uint32_t)(l * 32 + j0), acc);
        tmem_ld_wait();
        const int off = l == 0 ? off0 : (l == 1 ? off1 : off2);
        float* dst = q == 0 ? g + off + HID + lane * HID + j0 : g + off + j0;
        if (q == 0 || lane == 0) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
            *(float4*)(dst + c * 4) = make_float4(w0 * __uint_as_float(acc[c * 4]), w0 * __uint_as_float(acc[c * 4 + 1]),
                                                  w0 * __uint_as_float(acc[c * 4 + 2]), w0 * __uint_as_float(acc[c * 4 + 3]));
        }
      }
      if (hh == 0) {
        tmem_ld16_issue(tm + TM_DW + 96, acc);
        tmem_ld_wait();
        if (q == 0) {
#pragma unroll
          for (int k = 0; k < OUT; ++k) g[off3 + OUT + lane * OUT + k] = __uint_as_float(acc[k]);
        } else if (lane == 0) {
#pragma unroll
          for (int k = 0; k < OUT; ++k) g[off3 + k] = __uint_as_float(acc[k]);
        }
      }
    }
    if (MODE == 1) {
      sq = warp_sum(sq);
      if (hh == 0 && lane == 0) plain[256 + q] = sq;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (MODE == 1 && threadIdx.x == 0) a.sqerr[item] = (plain[256] + plain[257]) + (plain[258] + plain[259]);
  if (warp == 0) tmem_dealloc(tmem_base, TM_COLS);
}

template <int OUT>
static int launch_mlp_tc(const rcb_mlp_args* a, cudaStream_t st) {
#define RCB_MT_LAUNCH(MODE)                                                                                   \
  do {                                                                                                        \
    cudaError_t e = cudaFuncSetAttribute(mlp_tc_kernel<OUT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         MtSmem::TOTAL);                                                      \
    if (e != cudaSuccess) { set_error("rcb_mlp_tc: smem opt-in failed: %s", cudaGetErrorString(e)); return -1; } \
    mlp_tc_kernel<OUT, MODE><<<a->items, MT_THREADS, MtSmem::TOTAL, st>>>(*a);                                \
  } while (0)
  if (a->mode == 0) RCB_MT_LAUNCH(0);
  else if (a->mode == 1) RCB_MT_LAUNCH(1);
  else RCB_MT_LAUNCH(2);
#undef RCB_MT_LAUNCH
  RCB_CHECK_LAUNCH("rcb_mlp_tc");
  return 0;
}

}  // namespace rcb

using namespace rcb;

#ifdef RCB_MLP_PROFILE
extern "C" int rcb_mlp_prof_read(long long* host) {
  return (int)cudaMemcpyFromSymbol(host, rcb_prof_buf, sizeof(long long) * 4 * 512);
}
#endif

extern "C" int rcb_mlp_tc(const rcb_mlp_args* a, rcb_stream_t stream) {
  RCB_CHECK_ARG(a != nullptr, "rcb_mlp_tc: null args");
  RCB_CHECK_ARG(a->items > 0 && a->S > 0 && a->pix > 0, "rcb_mlp_tc: empty problem");
  RCB_CHECK_ARG(a->mode >= 0 && a->mode <= 2, "rcb_mlp_tc: bad mode %d", a->mode);
  RCB_CHECK_ARG(a->wt && a->xt && a->pe, "rcb_mlp_tc: null input");
  RCB_CHECK_ARG(a->n_f == 16, "rcb_mlp_tc: the tensor-core MLP is built for 32 input features (16 Fourier + 16 pe); "
                              "use rcb_mlp for other shapes");
  RCB_CHECK_ARG(a->mode != 0 || a->y_pred, "rcb_mlp_tc: mode 0 needs y_pred");
  RCB_CHECK_ARG(a->mode != 1 || (a->y && a->sqerr), "rcb_mlp_tc: mode 1 needs y and sqerr");
  RCB_CHECK_ARG(a->mode != 2 || a->dy, "rcb_mlp_tc: mode 2 needs dy");
  RCB_CHECK_ARG(a->mode == 0 || (a->d_pe && a->d_wt), "rcb_mlp_tc: backward needs d_pe and d_wt");
  RCB_CHECK_ARG(a->ld_w % 4 == 0, "rcb_mlp_tc: ld_w must be a multiple of 4");
  RCB_CHECK_ARG(!a->pe_base || (a->ph > 0 && a->pw > 0), "rcb_mlp_tc: stitched addressing needs the patch extent");
  cudaStream_t st = (cudaStream_t)stream;
  if (a->out == 3) return launch_mlp_tc<3>(a, st);
  if (a->out == 1) return launch_mlp_tc<1>(a, st);
  set_error("rcb_mlp_tc: unsupported output width %d", a->out);
  return -2;
}
