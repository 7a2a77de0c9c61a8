// Tensor-core version of the fused per-item SIREN MLP (forward + squared error + backward):
// every contraction of the 128-pixel tile runs on tcgen05 (TF32 operands, fp32 TMEM
// accumulators); CUDA cores only do the epilogues (bias, sin/cos, loss, masks).
//
// One CTA per (row, MC sample) item.  warp 0 = MMA issuer; warps 1-8 = 256 epilogue threads,
// two per pixel row (each owns 16 of the 32 features; TMEM lane quarter = warp % 4).
// Per 128-pixel tile, chained through TMEM and shared memory without touching HBM:
//   X0 -> [X0 W0] -> sin -> X1 -> [X1 W1] -> sin -> X2 -> [X2 W2] -> sin -> X3 -> y (CUDA cores, 32->out)
//   dy -> dZ2 -> [dZ2 W2^T] -> dZ1 -> [dZ1 W1^T] -> dZ0 -> [dZ0 W0pe^T] -> d pe
//   [X0|X1|X2|X3]^T [dZ0|dZ1|dZ2|dy]  -> all four weight gradients in ONE 128x128 accumulator that
//   stays in TMEM over the item's tiles.  (TF32 MN-major operands would need the 32B-base
//   swizzle, incompatible with the K-major tiles of the chain, so the epilogues also write a
//   feature-major copy [feature][pixel] of every activation / gradient tile for this GEMM.)
// Chain tiles are [128 px][32 features] fp32, 128-byte rows, 128B-swizzled exactly as a TMA
// box would write them (written here from registers; fence.proxy.async before the MMA).
// Reference semantics: test_model.py:347-355, 624-627; weight layout :269-280.
#include <type_traits>

#include "tc_common.cuh"

namespace rcb {

constexpr int MT_THREADS = 288;
constexpr int MT_EPI = 256;
constexpr int TILE_BYTES = 128 * 128;
constexpr int WTILE_BYTES = 32 * 128;

// byte offset of element (row, col) in a [rows][32 fp32] tile with 128-byte swizzle
__device__ __forceinline__ uint32_t swz(int row, int col) {
  return (uint32_t)(row * 128 + ((((col >> 2) ^ (row & 7)) << 4) | ((col & 3) << 2)));
}
__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// sin/cos with two-term Cody-Waite reduction to [-pi, pi] and the MUFU approximations
// (abs error ~4e-7 there); arguments are w0*z with |w0*z| of at most a few hundred
__device__ __forceinline__ void fast_sincos(float x, float* s, float* c) {
  const float k = rintf(x * 0.15915494309189535f);
  float r = fmaf(-k, 6.2831855f, x);
  r = fmaf(-k, -1.7484555e-7f, r);
  *s = __sinf(r);
  *c = __cosf(r);
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

struct MtSmem {
  // all tiles 1024-B aligned
  static constexpr int XA = 0;                              // px-major activation tile (A of the layer GEMMs)
  static constexpr int DZA = TILE_BYTES;                    // px-major gradient tile (A of the data-gradient GEMMs)
  static constexpr int TA = 2 * TILE_BYTES;                 // feature-major [4 px-blocks][128 = 4 layers x 32][32 px]
  static constexpr int TB = 6 * TILE_BYTES;                 // same for [dZ0 | dZ1 | dZ2 | dy]
  static constexpr int WF = 10 * TILE_BYTES;                // 3 forward weight tiles  [j][i]
  static constexpr int WB = WF + 3 * WTILE_BYTES;           // 3 backward weight tiles [i][j] (layer 0: 16 pe rows)
  static constexpr int PLAIN = WB + 3 * WTILE_BYTES;        // biases (3*32 + 4), W3 (32*4), dy exchange (128*4)
  static constexpr int BAR = PLAIN + 1024 + 2048;
  static constexpr int TOTAL = BAR + 128 + 1024;
};

template <int OUT, int MODE>
__global__ void __launch_bounds__(MT_THREADS, 1) mlp_tc_kernel(rcb_mlp_args a) {
  constexpr int F = 16, HID = 32, NPE = 16;
  constexpr int off0 = 0, off1 = HID * (32 + 1), off2 = off1 + HID * (HID + 1), off3 = off2 + HID * (HID + 1);
  constexpr int n_w = off3 + OUT * (HID + 1);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  float* plain = (float*)(smem + MtSmem::PLAIN);           // [0,96): b0,b1,b2; [96,100): b3; [128,256): W3 [i][4]
  uint64_t* bar_ready = (uint64_t*)(smem + MtSmem::BAR);   // epilogue -> MMA (256 arrivals)
  uint64_t* bar_mma = bar_ready + 1;                       // MMA -> epilogue (chain GEMM done)
  uint64_t* bar_dw = bar_ready + 2;                        // MMA -> epilogue (weight-gradient GEMM done)
  uint32_t* tmem_slot = (uint32_t*)(bar_ready + 3);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x;
  const int row_item = item / a.S;
  const int pix = a.pix;
  const int ntiles = (pix + 127) / 128;
  const float* wt_g = a.wt + (int64_t)item * a.ld_w;

  if (threadIdx.x == 0) {
    mbar_init(bar_ready, MT_EPI);
    mbar_init(bar_mma, 1);
    mbar_init(bar_dw, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  // ---- stage the item's weights: swizzled tf32 tiles for the MMAs, plain copies for the epilogues
  if (warp > 0) {
    const int t = threadIdx.x - 32;
    for (int e = t; e < 3 * HID * HID; e += MT_EPI) {
      const int l = e / (HID * HID), r = (e / HID) % HID, c = e % HID;          // W_l[i=r][j=c]
      const int off = l == 0 ? off0 : (l == 1 ? off1 : off2);
      const float w = to_tf32(wt_g[off + HID + r * HID + c]);
      *(float*)(smem + MtSmem::WF + l * WTILE_BYTES + swz(c, r)) = w;           // forward B: rows j, K = i
      if (l > 0) *(float*)(smem + MtSmem::WB + l * WTILE_BYTES + swz(r, c)) = w; // backward B: rows i, K = j
      else if (r >= F) *(float*)(smem + MtSmem::WB + swz(r - F, c)) = w;        // layer 0: pe inputs only
    }
    for (int e = t; e < 3 * HID; e += MT_EPI) plain[e] = wt_g[(e / HID == 0 ? off0 : (e / HID == 1 ? off1 : off2)) + e % HID];
    if (t < OUT) plain[96 + t] = wt_g[off3 + t];
    for (int e = t; e < HID * 4; e += MT_EPI) plain[128 + e] = (e % 4 < OUT) ? wt_g[off3 + OUT + (e / 4) * OUT + e % 4] : 0.f;
    // rows 96..127 of the feature-major gradient tiles hold dy (OUT rows) and zeros
    for (int e = t; e < 4 * 32 * 32; e += MT_EPI) {
      const int kb = e / 1024, rr = 96 + (e / 32) % 32, c = e % 32;
      *(float*)(smem + MtSmem::TB + kb * TILE_BYTES + swz(rr, c)) = 0.f;
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_chain = tmem_base;            // 32 columns: chain accumulator
  const uint32_t tm_dw = tmem_base + 32;          // 128 columns: [X0..X3]^T [dZ0..dZ2, dy]

  if (warp == 0) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      const uint32_t id32 = idesc_tf32(32), id16 = idesc_tf32(16);
      const uint32_t id_dw = idesc_tf32(128);
      const uint32_t sX = smem_u32(smem + MtSmem::XA), sDZ = smem_u32(smem + MtSmem::DZA);
      const uint32_t sTA = smem_u32(smem + MtSmem::TA), sTB = smem_u32(smem + MtSmem::TB);
      const uint32_t sWF = smem_u32(smem + MtSmem::WF), sWB = smem_u32(smem + MtSmem::WB);
      uint32_t ph = 0;
      auto gemm = [&](uint32_t a_addr, uint32_t b_addr, uint32_t idesc) {
        mbar_wait(bar_ready, ph); ph ^= 1;
        tc_fence_after();
        const uint64_t da = smem_desc_sw128(a_addr), db = smem_desc_sw128(b_addr);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_tf32(tm_chain, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, k ? 1u : 0u);
      };
      for (int tile = 0; tile < ntiles; ++tile) {
        for (int l = 0; l < 3; ++l) {
          gemm(sX, sWF + l * WTILE_BYTES, id32);
          umma_commit(bar_mma);
        }
        if (MODE != 0) {
          gemm(sDZ, sWB + 2 * WTILE_BYTES, id32);                    // dX2 = dZ2 W2^T
          umma_commit(bar_mma);
          gemm(sDZ, sWB + 1 * WTILE_BYTES, id32);                    // dX1 = dZ1 W1^T
          umma_commit(bar_mma);
          gemm(sDZ, sWB, id16);                                      // d pe = dZ0 W0[pe rows]^T
          umma_commit(bar_mma);
          // weight gradients of all layers: [4 x 32 features] x [128 gradient columns], K = 128 pixels
#pragma unroll
          for (int kb = 0; kb < 4; ++kb) {
            const uint64_t da = smem_desc_sw128(sTA + kb * TILE_BYTES), db = smem_desc_sw128(sTB + kb * TILE_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_tf32(tm_dw, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), id_dw, (tile | kb | k) ? 1u : 0u);
          }
          umma_commit(bar_dw);
        }
      }
    }
  } else {
    // ================================ epilogue threads ================================
    const int q = warp & 3;                 // TMEM lane quarter
    const int hh = (warp - 1) >> 2;         // which 16 of the 32 features
    const int r = q * 32 + lane;            // pixel row inside the tile
    const int j0 = hh * 16;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const float w0 = a.w0;
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
    uint32_t ph_mma = 0, ph_dw = 0;
    float gb[3][16];                        // bias-gradient partial sums of this row
#pragma unroll
    for (int l = 0; l < 3; ++l)
#pragma unroll
      for (int j = 0; j < 16; ++j) gb[l][j] = 0.f;
    float gb3[OUT];
#pragma unroll
    for (int k = 0; k < OUT; ++k) gb3[k] = 0.f;
    float sq = 0.f;

    float* dy_x = plain + 256;              // [128 rows][4]: dy exchange between the two halves of a row
    // this thread's 16 values of pixel row r: px-major chain tile (optional) + feature-major copy
    auto store16 = [&](uint8_t* chain_tile, uint8_t* t_tiles, int feat_base, const float (&v)[16]) {
      float w[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) w[j] = to_tf32(v[j]);
      if (chain_tile) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          *(float4*)(chain_tile + swz(r, j0 + c * 4)) = make_float4(w[c * 4], w[c * 4 + 1], w[c * 4 + 2], w[c * 4 + 3]);
      }
      if (MODE != 0) {
        uint8_t* blk = t_tiles + (r >> 5) * TILE_BYTES;
#pragma unroll
        for (int j = 0; j < 16; ++j) *(float*)(blk + swz(feat_base + j0 + j, r & 31)) = w[j];
      }
    };

    for (int tile = 0; tile < ntiles; ++tile) {
      const int pix0 = tile * 128;
      const int gp = pix0 + r;
      const bool valid = gp < pix;
      if (MODE != 0 && tile > 0) { mbar_wait(bar_dw, ph_dw); ph_dw ^= 1; }   // previous tile's dW GEMM read everything
      // ---- X0 = [fourier | pe]: half 0 loads the Fourier features, half 1 the positional encodings
      {
        float v[16];
        if (hh == 0) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = valid ? __ldg(xt + (int64_t)i * pix + gp) : 0.f;
        } else {
          const float4* p = reinterpret_cast<const float4*>(a.pe + (pe_origin + (valid ? pe_off(gp) : 0)) * NPE);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float4 t = valid ? __ldg(p + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            v[c * 4] = t.x; v[c * 4 + 1] = t.y; v[c * 4 + 2] = t.z; v[c * 4 + 3] = t.w;
          }
        }
        store16(smem + MtSmem::XA, smem + MtSmem::TA, 0, v);
      }
      fence_async_smem();
      mbar_arrive(bar_ready);

      // ---- three sine layers
      float cs[3][16];
#pragma unroll
      for (int l = 0; l < 3; ++l) {
        mbar_wait(bar_mma, ph_mma); ph_mma ^= 1;
        tc_fence_after();
        uint32_t acc[16];
        tmem_ld16(tm_chain + lane_addr + (uint32_t)j0, acc);
        float x[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float s, c;
          fast_sincos(w0 * (__uint_as_float(acc[j]) + plain[l * 32 + j0 + j]), &s, &c);
          x[j] = s;
          cs[l][j] = w0 * c;
        }
        store16(smem + MtSmem::XA, smem + MtSmem::TA, (l + 1) * 32, x);
        tc_fence_before();
        if (l < 2) { fence_async_smem(); mbar_arrive(bar_ready); }
      }
      epi_sync();                            // X3 rows complete (both halves) for the generic-proxy reads below
      // ---- last (linear) layer, loss and dy: one thread per pixel row
      if (hh == 0) {
        float o[OUT];
#pragma unroll
        for (int k = 0; k < OUT; ++k) o[k] = plain[96 + k];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float4 xv = *(const float4*)(smem + MtSmem::XA + swz(r, c * 4));
          const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
          for (int t = 0; t < 4; ++t)
#pragma unroll
            for (int k = 0; k < OUT; ++k) o[k] = fmaf(xs[t], plain[128 + (c * 4 + t) * 4 + k], o[k]);
        }
        if (MODE == 0) {
          if (valid)
#pragma unroll
            for (int k = 0; k < OUT; ++k) a.y_pred[((int64_t)item * pix + gp) * OUT + k] = o[k];
        } else {
          float dyv[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int k = 0; k < OUT; ++k) {
            if (MODE == 1) {
              float rr = valid ? o[k] - __ldg(a.y + ((int64_t)row_item * pix + gp) * OUT + k) : 0.f;
              sq = fmaf(rr, rr, sq);
              dyv[k] = a.coef * rr;
            } else {
              dyv[k] = valid ? __ldg(a.dy + ((int64_t)item * pix + gp) * OUT + k) : 0.f;
            }
            gb3[k] += dyv[k];
          }
          *(float4*)(dy_x + r * 4) = make_float4(dyv[0], dyv[1], dyv[2], dyv[3]);
#pragma unroll
          for (int k = 0; k < OUT; ++k)
            *(float*)(smem + MtSmem::TB + (r >> 5) * TILE_BYTES + swz(96 + k, r & 31)) = to_tf32(dyv[k]);
        }
      }
      epi_sync();                            // dy visible to the partner half; X3 reads done before XA is reused
      if (MODE == 0) continue;
      // ---- dZ2 = (dy W3^T) * 30 cos
      {
        float4 d4 = *(const float4*)(dy_x + r * 4);
        const float dyv[4] = {d4.x, d4.y, d4.z, d4.w};
        float dz[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float v = 0.f;
#pragma unroll
          for (int k = 0; k < OUT; ++k) v = fmaf(dyv[k], plain[128 + (j0 + j) * 4 + k], v);
          dz[j] = v * cs[2][j];
          gb[2][j] += dz[j];
        }
        store16(smem + MtSmem::DZA, smem + MtSmem::TB, 2 * 32, dz);
      }
      fence_async_smem();
      mbar_arrive(bar_ready);
      // ---- dZ1, dZ0 through the tensor core
#pragma unroll
      for (int l = 1; l >= 0; --l) {
        mbar_wait(bar_mma, ph_mma); ph_mma ^= 1;
        tc_fence_after();
        uint32_t acc[16];
        tmem_ld16(tm_chain + lane_addr + (uint32_t)j0, acc);
        float dz[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          dz[j] = __uint_as_float(acc[j]) * cs[l][j];
          gb[l][j] += dz[j];
        }
        store16(smem + MtSmem::DZA, smem + MtSmem::TB, l * 32, dz);
        tc_fence_before();
        fence_async_smem();
        mbar_arrive(bar_ready);
      }
      // ---- d pe (16 columns: 8 per half)
      {
        mbar_wait(bar_mma, ph_mma); ph_mma ^= 1;
        tc_fence_after();
        uint32_t acc[16];
        tmem_ld16(tm_chain + lane_addr, acc);        // all 16; this half keeps [hh*8, hh*8+8)
        if (valid) {
          float* dst = a.d_pe + (pe_origin + pe_off(gp)) * NPE + hh * 8;
          const int b = hh * 8;
          *(float4*)(dst) = make_float4(__uint_as_float(acc[b]), __uint_as_float(acc[b + 1]), __uint_as_float(acc[b + 2]), __uint_as_float(acc[b + 3]));
          *(float4*)(dst + 4) = make_float4(__uint_as_float(acc[b + 4]), __uint_as_float(acc[b + 5]), __uint_as_float(acc[b + 6]), __uint_as_float(acc[b + 7]));
        }
        tc_fence_before();
      }
    }

    if (MODE != 0) {
      // ---- weight gradients: row f of the 128x128 accumulator belongs to layer q = f / 32
      mbar_wait(bar_dw, ph_dw);
      tc_fence_after();
      float* g = a.d_wt + (int64_t)item * a.ld_w;
      {
        uint32_t acc[16];
        tmem_ld16(tm_dw + lane_addr + (uint32_t)(q * 32 + j0), acc);
        const int i = lane;                              // input feature of layer q
        if (q < 3) {
          const int off = q == 0 ? off0 : (q == 1 ? off1 : off2);
#pragma unroll
          for (int c = 0; c < 4; ++c)
            *(float4*)(g + off + HID + i * HID + j0 + c * 4) =
                make_float4(__uint_as_float(acc[c * 4]), __uint_as_float(acc[c * 4 + 1]), __uint_as_float(acc[c * 4 + 2]), __uint_as_float(acc[c * 4 + 3]));
        } else if (hh == 0) {
#pragma unroll
          for (int k = 0; k < OUT; ++k) g[off3 + OUT + i * OUT + k] = __uint_as_float(acc[k]);
        }
      }
      // ---- bias gradients: reduce the per-row partial sums over the 128 rows (4 warps per half)
      float* scratch = (float*)(smem + MtSmem::XA);      // tiles are dead now: [8 warps][64]
      epi_sync();
#pragma unroll
      for (int l = 0; l < 3; ++l)
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float v = warp_sum(gb[l][j]);
          if (lane == 0) scratch[(warp - 1) * 64 + l * 16 + j] = v;
        }
#pragma unroll
      for (int k = 0; k < OUT; ++k) {
        float v = warp_sum(gb3[k]);
        if (lane == 0) scratch[(warp - 1) * 64 + 48 + k] = v;
      }
      sq = warp_sum(sq);
      if (lane == 0) scratch[(warp - 1) * 64 + 60] = sq;
      epi_sync();
      const int t = threadIdx.x - 32;
      if (t < 96) {                                      // (layer, feature): feature half = (j / 16)
        const int l = t / 32, j = t % 32, h2 = j / 16;
        float s = 0.f;
        for (int w = 0; w < 4; ++w) s += scratch[(h2 * 4 + w) * 64 + l * 16 + (j % 16)];
        g[(l == 0 ? off0 : (l == 1 ? off1 : off2)) + j] = s;
      } else if (t < 96 + OUT) {
        float s = 0.f;
        for (int w = 0; w < 4; ++w) s += scratch[w * 64 + 48 + (t - 96)];
        g[off3 + (t - 96)] = s;
      } else if (t == 128 && MODE == 1) {
        float s = 0.f;
        for (int w = 0; w < 4; ++w) s += scratch[w * 64 + 60];
        a.sqerr[item] = s;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
  (void)n_w;
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
