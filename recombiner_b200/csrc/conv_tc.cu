// Tensor-core polyphase up-convolution: TMA-gathered implicit GEMM on tcgen05 (TF32, fp32
// accumulation in TMEM).  No im2col and no materialised upsampling: each k-block of the
// A operand is ONE 5-D TMA box of the channel-last activation tensor, shifted by the tap
// offset (forward) or strided by the upsampling factor (data gradient); image borders are
// the TMA unit's out-of-bounds zero fill.
//
//   forward : out[item, s*f + r, :] = act(bias + sum_{tap,ic} src[item, s + b_r + tap, ic] * w_eff[r][tap][ic][:])
//             one CTA = 128 source pixels x a group of phases r, one TMEM accumulator per phase
//   backward: d_src[item, s, :] = lrelu'(.) * sum_{r,tap,oc} d_out[item, (s - b_r - tap)*f + r, oc] * w_eff[r][tap][:][oc]
//             one CTA = 128 source pixels, K = phases*taps*oc
// Warp roles as in gemm_tc.cu.  Reference semantics: prior_model.py:47-59.
#include <cstdlib>
#include <cstring>
#include <cuda_fp16.h>
#include "gemm_engine.cuh"
#include "tc_common.cuh"

namespace rcb {

__device__ __forceinline__ uint2 pack_h4(const float4& o) {
  const __half2 lo = __floats2half2_rn(o.x, o.y), hi = __floats2half2_rn(o.z, o.w);
  return make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
}

// same, saturating at +-65504 instead of producing inf (F2FP.SATFINITE)
__device__ __forceinline__ uint2 pack_h4_sat(const float4& o) {
  uint2 r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r.x) : "f"(o.y), "f"(o.x));
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r.y) : "f"(o.w), "f"(o.z));
  return r;
}

__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// instruction descriptor: D = f32, A = B = f16, both K-major, M = 128
__device__ __forceinline__ uint32_t idesc_f16_m128(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}

struct ConvTile {
  int tx, ty, tz, ni;            // box extent in source pixels / items (tx*ty*tz*ni <= 128)
  int ntx, nty, ntz;             // tiles per axis
};

struct ConvTcArgs {
  PolyGeom g;
  ConvTile t;
  int items;
  int phases_per_cta;            // forward: TMEM accumulators per CTA
  int kblocks;                   // k-blocks per (phase[,tap]) unit
  int bk;                        // 32 (128B swizzle) or 16 (64B swizzle)
  int act;
  int out_half;                  // forward: out is fp16 (the next stage reads it as an fp16 MMA operand)
  int in_half;                   // forward: src and the weights are fp16 (64 channels per 128-byte row, kind::f16)
  int act_half;                  // backward: src_act is fp16
  int b_off, stage_bytes, bar_off;   // shared-memory layout (bytes): B tile offset in a stage, stage size, barriers
  int nstages, epi_off;              // forward: ring depth and the epilogue staging area (4 warps x 32 x (oc + 4) floats)
  const float* bias;             // forward
  const float* src_act;          // backward (post-activation of the producing stage) or NULL
  float* out;
};

constexpr int CT_MAX_PHASES = 16;

// stage = [A tile: 128 rows x row_bytes][B tile: n rows x row_bytes], both 1024-B aligned; sized per
// problem so that several CTAs fit on an SM (their prologues/epilogues overlap the others' main loops)
static void smem_layout(ConvTcArgs* a, int row_bytes, int b_rows, int* total, int nstages = TC_STAGES, int epi_bytes = 0) {
  int a_bytes = (TC_BM * row_bytes + 1023) / 1024 * 1024;
  int b_bytes = (b_rows * row_bytes + 1023) / 1024 * 1024;
  a->b_off = a_bytes;
  a->stage_bytes = a_bytes + b_bytes;
  a->nstages = nstages;
  a->epi_off = nstages * a->stage_bytes;
  a->bar_off = a->epi_off + (epi_bytes + 1023) / 1024 * 1024;
  *total = a->bar_off + 512 + 1024;
}

__device__ __forceinline__ void tile_origin(const ConvTcArgs& a, int tile, int& item0, int& z0, int& y0, int& x0) {
  int t = tile;
  x0 = (t % a.t.ntx) * a.t.tx; t /= a.t.ntx;
  y0 = (t % a.t.nty) * a.t.ty; t /= a.t.nty;
  z0 = (t % a.t.ntz) * a.t.tz; t /= a.t.ntz;
  item0 = t * a.t.ni;
}
// tile row -> (item, z, y, x) of the source pixel; false if outside the tensor
__device__ __forceinline__ bool tile_row(const ConvTcArgs& a, int r, int item0, int z0, int y0, int x0,
                                         int& item, int& z, int& y, int& x) {
  int t = r;
  x = x0 + t % a.t.tx; t /= a.t.tx;
  y = y0 + t % a.t.ty; t /= a.t.ty;
  z = z0 + t % a.t.tz; t /= a.t.tz;
  item = item0 + t;
  return t < a.t.ni && item < a.items && z < a.g.d && y < a.g.h && x < a.g.w;
}

// ------------------------------------------------------------------------------ forward --
__global__ void __launch_bounds__(TC_THREADS)
upconv_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, ConvTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + a.bar_off);
  uint64_t* empty = full + TC_STAGES;
  uint64_t* acc_full = empty + TC_STAGES;                 // one per phase handled by this CTA
  uint32_t* tmem_slot = (uint32_t*)(acc_full + CT_MAX_PHASES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const PolyGeom& g = a.g;
  const int OC = g.oc;
  int item0, z0, y0, x0;
  tile_origin(a, blockIdx.x, item0, z0, y0, x0);
  const int ph0 = blockIdx.y * a.phases_per_cta;
  const int nph = min(a.phases_per_cta, g.phases() - ph0);
  const int kb_per_phase = g.taps() * a.kblocks;          // kblocks = ic / KC
  const int KC = a.in_half ? 64 : 32;                     // channels per 128-byte operand row
  const uint32_t a_bytes = (uint32_t)(a.t.tx * a.t.ty * a.t.tz * a.t.ni) * 128u;
  const uint32_t b_bytes = (uint32_t)OC * 128u;
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < a.phases_per_cta * OC) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < a.nstages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }   // nstages <= TC_STAGES
    for (int p = 0; p < CT_MAX_PHASES; ++p) mbar_init(&acc_full[p], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // warps 0 and 1 stay converged; one elected lane issues the TMA loads / MMAs (see elect_one)
  if (warp == 0) {
    int s = 0;
    uint32_t par = 1;
    for (int p = 0; p < nph; ++p) {
      int rz, ry, rx;
      g.split_phase(ph0 + p, rz, ry, rx);
      const int bz = g.base_z(rz), by = g.base_y(ry), bx = g.base_x(rx);
      int kb = 0;
      for (int tz = 0; tz < g.Tz; ++tz)
        for (int ty = 0; ty < g.Ty; ++ty)
          for (int tx = 0; tx < g.Tx; ++tx)
            for (int cb = 0; cb < a.kblocks; ++cb, ++kb) {
              mbar_wait(&empty[s], par);
              if (elect_one()) {
                uint8_t* a_dst = smem + s * a.stage_bytes;
                mbar_expect_tx(&full[s], a_bytes + b_bytes);
                tma_load_5d(&tmA, &full[s], a_dst, cb * KC, x0 + bx + tx, y0 + by + ty, z0 + bz + tz, item0);
                tma_load_2d(&tmB, &full[s], a_dst + a.b_off, kb * KC, (ph0 + p) * OC);
              }
              __syncwarp();
              if (++s == a.nstages) { s = 0; par ^= 1; }
            }
    }
  } else if (warp == 1) {
    const uint32_t idesc = a.in_half ? idesc_f16_m128(OC) : idesc_tf32(OC);
    int s = 0;
    uint32_t par = 0;
    for (int p = 0; p < nph; ++p) {
      for (int kb = 0; kb < kb_per_phase; ++kb) {
        mbar_wait(&full[s], par);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(smem + s * a.stage_bytes);
          const uint64_t da = smem_desc_sw128(a_addr), db = smem_desc_sw128(a_addr + a.b_off);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (a.in_half) umma_f16_ss(tmem_base + (uint32_t)(p * OC), da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
            else umma_tf32(tmem_base + (uint32_t)(p * OC), da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
          }
          umma_commit(&empty[s]);
          if (kb == kb_per_phase - 1) umma_commit(&acc_full[p]);
        }
        __syncwarp();
        if (++s == a.nstages) { s = 0; par ^= 1; }
      }
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    int item, z, y, x;
    const bool valid = tile_row(a, r, item0, z0, y0, x0, item, z, y, x);
    const int Ho = g.h * g.fy, Wo = g.w * g.fx, Do = g.d * g.fz;
    // element offset of this row's phase-(0,0,0) output pixel, or -1
    const int64_t base0 = valid ? ((((int64_t)item * Do + (int64_t)z * g.fz) * Ho + (int64_t)y * g.fy) * Wo + (int64_t)x * g.fx) * OC : -1;
    // Each warp stages its 32 x OC accumulator rows of a phase in its own shared-memory area (row per lane out of
    // TMEM), then writes them back with consecutive lanes on consecutive 16-B chunks of a row: whole 128-byte
    // lines per store instruction instead of 32 scattered ones.
    const int ldw = OC + 4;
    float* wbuf = reinterpret_cast<float*>(smem + a.epi_off) + (size_t)q * 32 * ldw;
    const int lanes_per_row = OC / 4, rows_per_iter = 32 / lanes_per_row;
    const int sub = lane / lanes_per_row, ch = (lane % lanes_per_row) * 4;
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + ch));
    for (int p = 0; p < nph; ++p) {
      int rz, ry, rx;
      g.split_phase(ph0 + p, rz, ry, rx);
      const int64_t ph_off = (((int64_t)rz * Ho + ry) * Wo + rx) * OC;
      mbar_wait(&acc_full[p], 0);
      tc_fence_after();
      for (int c0 = 0; c0 < OC; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(p * OC + c0), v);
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          *reinterpret_cast<float4*>(wbuf + lane * ldw + c0 + j) =
              make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
      }
      __syncwarp();
      // four walk steps per trip: the shuffles and shared-memory loads of all four are issued before the stores
      for (int r0 = 0; r0 < 32; r0 += 4 * rows_per_iter) {
        float4 f[4];
        int64_t bb[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int rr = r0 + u * rows_per_iter + sub;
          bb[u] = __shfl_sync(0xffffffffu, base0, rr & 31);
          if (sub >= rows_per_iter || rr >= 32) bb[u] = -1;
          f[u] = *reinterpret_cast<const float4*>(wbuf + (rr & 31) * ldw + ch);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (bb[u] >= 0) {
            float4 o = f[u];
            o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;
            if (a.act) {
              o.x = o.x > 0.f ? o.x : 0.01f * o.x; o.y = o.y > 0.f ? o.y : 0.01f * o.y;
              o.z = o.z > 0.f ? o.z : 0.01f * o.z; o.w = o.w > 0.f ? o.w : 0.01f * o.w;
            }
            if (a.out_half) *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(a.out) + bb[u] + ph_off + ch) = pack_h4(o);
            else *reinterpret_cast<float4*>(a.out + bb[u] + ph_off + ch) = o;
          }
        }
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

// ------------------------------------------------------------------- forward, halo tile --
// 2-D grids: one CTA = 8 x 16 source pixels of one item.  Per 32-channel k-block the (8+2) x (16+2)
// neighbourhood is fetched ONCE (one 5-D TMA box of 18 lines x 10 pixels, image borders zero-filled) and
// every (phase, tap) MMA reads its shifted 128 rows out of that tile through the shared-memory
// descriptor alone: start address + (line, pixel) shift, 8-row groups 1280 B (= one line) apart.  The
// 128-byte swizzle is a function of the absolute shared-memory address on both the TMA and the MMA
// side, so a row-shifted start needs no base offset (checked on hardware).  A traffic drops from
// (phases x taps) tiles per k-block to 2.25 tiles.  The weight tiles of all taps of one
// (k-block, phase) travel as one ring stage (one barrier wait per 4 * taps MMAs), and the
// single-thread loops carry no integer division.
struct ConvHaloArgs {
  PolyGeom g;
  int items, tiles_x, tiles_y;
  int phases_per_cta, kblocks;         // kblocks = ic / 32
  int act, out_half;
  int b_bytes, nbs;                    // weight ring: nbs stages of taps x (oc x 128 B)
  int a_off, bar_off;                  // shared-memory layout
  const float* bias;
  float* out;
};
constexpr int HALO_LINES = 18, HALO_PITCH = 10;                 // (16 + 2) lines of (8 + 2) pixel rows
constexpr int HALO_BYTES = HALO_LINES * HALO_PITCH * 128;       // 23040 per k-block
constexpr int HALO_BUF = (HALO_BYTES + 1023) / 1024 * 1024;     // double-buffered, 1024-B aligned
constexpr int HALO_NBS_MAX = 4;

#ifdef RCB_CONV_PROFILE
__device__ long long rcb_conv_prof[256];
#define CPROF(i) do { if (blockIdx.x == 5000 && blockIdx.y == 0 && lane == 0) rcb_conv_prof[i] = clock64(); } while (0)
#else
#define CPROF(i) do {} while (0)
#endif

__global__ void __launch_bounds__(TC_THREADS)
upconv_fwd_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, ConvHaloArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* a_full = (uint64_t*)(smem + a.bar_off);       // [2] halo tile landed
  uint64_t* a_empty = a_full + 2;                          // [2] halo tile consumed
  uint64_t* b_full = a_empty + 2;                          // [nbs]
  uint64_t* b_empty = b_full + HALO_NBS_MAX;               // [nbs]
  uint64_t* acc_full = b_empty + HALO_NBS_MAX;             // [1]
  uint32_t* tmem_slot = (uint32_t*)(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 1) CPROF(0);
  const PolyGeom& g = a.g;
  const int OC = g.oc;
  int t = blockIdx.x;
  const int x0 = (t % a.tiles_x) * 8; t /= a.tiles_x;
  const int y0 = (t % a.tiles_y) * 16; t /= a.tiles_y;
  const int item = t;
  const int ph0 = blockIdx.y * a.phases_per_cta;
  const int nph = min(a.phases_per_cta, g.phases() - ph0);
  const int stage_bytes = 4 * a.b_bytes;                   // Ty * Tx = 4 taps
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < a.phases_per_cta * OC) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < a.nbs; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer: halo tile per k-block, then the weight tiles of its phases (all taps per stage)
    int s = 0;
    uint32_t par = 1;                               // parity to wait on b_empty[s]: first lap passes
    auto load_halo = [&](int kb) {
      const int ab = kb & 1;
      mbar_wait(&a_empty[ab], ((kb >> 1) & 1) ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&a_full[ab], HALO_BYTES);
        tma_load_5d(&tmA, &a_full[ab], smem + a.a_off + ab * HALO_BUF, kb * 32, x0 - 1, y0 - 1, 0, item);
      }
      __syncwarp();
    };
    load_halo(0);
    for (int kb = 0; kb < a.kblocks; ++kb) {
      if (kb + 1 < a.kblocks) load_halo(kb + 1);    // the next tile travels under this k-block's MMAs
      for (int p = 0; p < nph; ++p) {
        mbar_wait(&b_empty[s], par);
        if (elect_one()) {
          mbar_expect_tx(&b_full[s], (uint32_t)stage_bytes);
#pragma unroll
          for (int tap = 0; tap < 4; ++tap)
            tma_load_2d(&tmB, &b_full[s], smem + s * stage_bytes + tap * a.b_bytes, tap * g.ic + kb * 32, (ph0 + p) * OC);
        }
        __syncwarp();
        if (++s == a.nbs) { s = 0; par ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer
    CPROF(1);
    const uint32_t idesc = idesc_tf32(OC);
    // descriptor template: K-major, 128-byte swizzle, 8-row groups one line (2048 B) apart
    const uint64_t da_hi = ((uint64_t)1 << 16) | ((uint64_t)((HALO_PITCH * 128) >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    int s = 0;
    uint32_t par = 0;
    for (int kb = 0; kb < a.kblocks; ++kb) {
      const int ab = kb & 1;
      const uint32_t a_addr = smem_u32(smem + a.a_off + ab * HALO_BUF);
      mbar_wait(&a_full[ab], (kb >> 1) & 1);
      tc_fence_after();
      CPROF(2 + kb * 40);
      int ry = ph0 / g.fx, rx = ph0 - ry * g.fx;    // 2-D: phase = ry * fx + rx
      for (int p = 0; p < nph; ++p) {
        const int by = g.base_y(ry), bx = g.base_x(rx);
        mbar_wait(&b_full[s], par);
        tc_fence_after();
        CPROF(3 + kb * 40 + p);
        if (elect_one()) {
          const uint32_t b_addr = smem_u32(smem + s * stage_bytes);
#pragma unroll
          for (int tap = 0; tap < 4; ++tap) {
            const uint32_t shift_rows = (uint32_t)((1 + by + (tap >> 1)) * HALO_PITCH + (1 + bx + (tap & 1)));
            const uint64_t da = da_hi | (uint64_t)(((a_addr + shift_rows * 128u) & 0x3FFFF) >> 4);
            const uint64_t db = smem_desc_sw128(b_addr + tap * a.b_bytes);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_tf32(tmem_base + (uint32_t)(p * OC), da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | tap | k) ? 1u : 0u);
          }
          umma_commit(&b_empty[s]);
          if (p == nph - 1) {
            umma_commit(&a_empty[ab]);
            if (kb == a.kblocks - 1) umma_commit(acc_full);
          }
        }
        __syncwarp();
        if (++s == a.nbs) { s = 0; par ^= 1; }
        if (++rx == g.fx) { rx = 0; ++ry; }
      }
    }
  } else {
    // ===== epilogue: row m = (line m / 8, pixel m % 8) of the tile
    const int q = warp & 3;
    const int Wo = g.w * g.fx;
    const int64_t row_pitch = (int64_t)Wo * OC;
    if (warp == 2) CPROF(100);
    mbar_wait(acc_full, 0);
    tc_fence_after();
    if (warp == 2) CPROF(101);
    const int W = g.fx * OC;                               // floats one source pixel contributes to an output line
    if (a.phases_per_cta % g.fx == 0 && W <= 128 && 128 % W == 0 && 4 * 32 * (W + 4) * 4 <= 2 * HALO_BUF) {
      // All MMAs are done, so the halo buffers are free: each warp stages its 32 rows x (fx phases x OC) there and
      // writes them back with consecutive lanes on consecutive 16-B chunks.  The fx phases of one output line are
      // adjacent in memory (pixel 2x, 2x+1, ...), so a store instruction covers whole 128-byte lines.
      const int ldw = W + 4;
      float* wbuf = reinterpret_cast<float*>(smem + a.a_off) + (size_t)q * 32 * ldw;
      const int lanes_per_row = W / 4, rows_per_iter = 32 / lanes_per_row;
      const int sub = lane / lanes_per_row, ch = (lane % lanes_per_row) * 4;
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + ch % OC));
      int ry = ph0 / g.fx;
      for (int p = 0; p < nph; p += g.fx, ++ry) {
        for (int c0 = 0; c0 < W; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(p * OC + c0), v);
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(wbuf + lane * ldw + c0 + j) =
                make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
        }
        __syncwarp();
        for (int r0 = 0; r0 < 32; r0 += rows_per_iter) {
          const int rr = r0 + sub, m = q * 32 + rr;
          const int y = y0 + (m >> 3), x = x0 + (m & 7);
          if (y < g.h && x < g.w) {
            float4 f = *reinterpret_cast<const float4*>(wbuf + rr * ldw + ch);
            f.x += b4.x; f.y += b4.y; f.z += b4.z; f.w += b4.w;
            if (a.act) {
              f.x = f.x > 0.f ? f.x : 0.01f * f.x; f.y = f.y > 0.f ? f.y : 0.01f * f.y;
              f.z = f.z > 0.f ? f.z : 0.01f * f.z; f.w = f.w > 0.f ? f.w : 0.01f * f.w;
            }
            const int64_t off = (((int64_t)item * g.h * g.fy + (int64_t)y * g.fy + ry) * Wo + (int64_t)x * g.fx) * OC + ch;
            if (a.out_half) *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(a.out) + off) = pack_h4(f);
            else *reinterpret_cast<float4*>(a.out + off) = f;
          }
        }
        __syncwarp();
      }
    } else {
      const int m = q * 32 + lane;
      const int y = y0 + (m >> 3), x = x0 + (m & 7);
      const bool valid = y < g.h && x < g.w;
      const int64_t obase = (((int64_t)item * g.h * g.fy + (int64_t)y * g.fy) * Wo + (int64_t)x * g.fx) * OC;
      int ry = ph0 / g.fx, rx = ph0 - ry * g.fx;
      for (int p = 0; p < nph; ++p) {
        const int64_t orow = obase + ry * row_pitch + rx * OC;
        for (int c0 = 0; c0 < OC; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(p * OC + c0), v);
          if (valid) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + c0 + j));
              float o[4] = {__uint_as_float(v[j]) + b4.x, __uint_as_float(v[j + 1]) + b4.y,
                            __uint_as_float(v[j + 2]) + b4.z, __uint_as_float(v[j + 3]) + b4.w};
              if (a.act) {
#pragma unroll
                for (int tt = 0; tt < 4; ++tt) o[tt] = o[tt] > 0.f ? o[tt] : 0.01f * o[tt];
              }
              const float4 o4 = make_float4(o[0], o[1], o[2], o[3]);
              if (a.out_half) *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(a.out) + orow + c0 + j) = pack_h4(o4);
              else *reinterpret_cast<float4*>(a.out + orow + c0 + j) = o4;
            }
          }
        }
        if (++rx == g.fx) { rx = 0; ++ry; }
      }
    }
    if (warp == 2) CPROF(102);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) CPROF(103);
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

// ------------------------------------------------- forward, factor 2, resident weights --
// The last upsampler stage (x2, 3-tap) is where the forward pass moves most of its bytes.  A tile's 16
// (phase, tap) products read only 9 distinct shifts of the halo tile, and products that share a shift differ only
// in the weight rows.  With the four phase accumulators laid side by side in TMEM in the order
// (0,0) (0,1) (1,1) (1,0), phases that share a shift are neighbours, so one MMA of N = members * oc serves them
// all: 10 MMAs per k-step instead of 16 (the tensor core's shared-memory operand read, ~128 B/clk, is what
// bounds these thin-N MMAs).  The weights (16 oc-row blocks per k-block, 64 KB for 64 -> 16 channels) are loaded
// once per CTA; the CTAs are persistent (one per SM), walk the tiles round-robin, and keep two accumulator sets
// so that the epilogue of a tile runs under the MMAs of the next.
constexpr int F2_GROUPS = 10, F2_STAGES = 4;
struct ConvF2Args {
  PolyGeom g;
  int items, tiles_x, tiles_y, n_tiles;
  int kblocks, act, out_half;
  int w_bytes_kb;                       // bytes of resident weights per k-block: 16 blocks of oc rows x 128 B
  int a_off, epi_off, bar_off;          // shared-memory layout
  int grp_shift[F2_GROUPS];             // halo row shift of the group's A operand
  int grp_col[F2_GROUPS];               // first accumulator column (within one accumulator set)
  int grp_n[F2_GROUPS];                 // MMA N
  int grp_row[F2_GROUPS];               // first weight row of the group within a k-block
  int mem_phase[16], mem_tap[16];       // weight block b (oc rows) = w_eff_k[mem_phase[b]][:, mem_tap[b]]
  int phase_col[4];                     // accumulator column of phase ry * 2 + rx
  const float* bias;
  float* out;
};

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
               ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cta(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// HALF: the source activations and the weights are fp16 (the 10-bit mantissa TF32 keeps of an fp32 operand anyway):
// 64 channels per 128-byte row, kind::f16 MMAs, half the shared-memory operand bytes per MMA.
template <bool HALF>
__global__ void __launch_bounds__(TC_THREADS, 1)
upconv_fwd_f2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmO, const __grid_constant__ ConvF2Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* a_full = (uint64_t*)(smem + a.bar_off);        // [F2_STAGES]
  uint64_t* a_empty = a_full + F2_STAGES;                   // [F2_STAGES]
  uint64_t* acc_full = a_empty + F2_STAGES;                 // [2]
  uint64_t* acc_empty = acc_full + 2;                       // [2], one arrival per epilogue warp
  uint64_t* w_full = acc_empty + 2;                         // [1]
  uint32_t* tmem_slot = (uint32_t*)(w_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const PolyGeom& g = a.g;
  const int OC = g.oc;
  const int acc_cols = 4 * OC;
  constexpr int KC = HALF ? 64 : 32;                        // channels per 128-byte operand row
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < 2 * acc_cols) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < F2_STAGES; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
    mbar_init(w_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto tile_origin = [&](int t, int& item, int& y0, int& x0) {
    x0 = (t % a.tiles_x) * 8; t /= a.tiles_x;
    y0 = (t % a.tiles_y) * 16; t /= a.tiles_y;
    item = t;
  };

  if (warp == 0) {
    // ===== TMA producer: the resident weights, then one halo tile per (tile, k-block)
    if (elect_one()) {
      mbar_expect_tx(w_full, (uint32_t)(a.kblocks * a.w_bytes_kb));
      for (int kb = 0; kb < a.kblocks; ++kb)
        for (int b = 0; b < 16; ++b)
          tma_load_2d(&tmB, w_full, smem + kb * a.w_bytes_kb + b * OC * 128, a.mem_tap[b] * g.ic + kb * KC, a.mem_phase[b] * OC);
    }
    __syncwarp();
    int s = 0;
    uint32_t par = 1;
    for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
      int item, y0, x0;
      tile_origin(t, item, y0, x0);
      for (int kb = 0; kb < a.kblocks; ++kb) {
        mbar_wait(&a_empty[s], par);
        if (elect_one()) {
          mbar_expect_tx(&a_full[s], HALO_BYTES);
          tma_load_5d(&tmA, &a_full[s], smem + a.a_off + s * HALO_BUF, kb * KC, x0 - 1, y0 - 1, 0, item);
        }
        __syncwarp();
        if (++s == F2_STAGES) { s = 0; par ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer
    const uint64_t da_hi = ((uint64_t)1 << 16) | ((uint64_t)((HALO_PITCH * 128) >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    uint32_t idesc[F2_GROUPS];
#pragma unroll
    for (int j = 0; j < F2_GROUPS; ++j) idesc[j] = HALF ? idesc_f16_m128(a.grp_n[j]) : idesc_tf32(a.grp_n[j]);
    mbar_wait(w_full, 0);
    int s = 0;
    uint32_t par = 0;
    int it = 0;
    for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x, ++it) {
      const int buf = it & 1;
      mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t acc = tmem_base + (uint32_t)(buf * acc_cols);
      for (int kb = 0; kb < a.kblocks; ++kb) {
        mbar_wait(&a_full[s], par);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(smem + a.a_off + s * HALO_BUF);
          const uint32_t w_addr = smem_u32(smem + kb * a.w_bytes_kb);
#pragma unroll
          for (int j = 0; j < F2_GROUPS; ++j) {
            const uint64_t da = da_hi | (uint64_t)(((a_addr + (uint32_t)a.grp_shift[j] * 128u) & 0x3FFFF) >> 4);
            const uint64_t db = smem_desc_sw128(w_addr + (uint32_t)a.grp_row[j] * 128u);
#pragma unroll
            for (int k = 0; k < 4; ++k) { // group 0 covers every column, so it alone starts the accumulation
              if (HALF) umma_f16_ss(acc + (uint32_t)a.grp_col[j], da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc[j], (kb | j | k) ? 1u : 0u);
              else umma_tf32(acc + (uint32_t)a.grp_col[j], da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc[j], (kb | j | k) ? 1u : 0u);
            }
          }
          umma_commit(&a_empty[s]);
          if (kb == a.kblocks - 1) umma_commit(&acc_full[buf]);
        }
        __syncwarp();
        if (++s == F2_STAGES) { s = 0; par ^= 1; }
      }
    }
  } else {
    // ===== epilogue: row m = (line m / 8, pixel m % 8) of the tile.  The two phases of an output line are
    // adjacent in memory (2 * oc = 32 floats = 128 B per source pixel), so each warp stages its 32 rows as
    // 128-byte swizzled rows -- bias and LeakyReLU applied in registers -- and one TMA store per output line
    // parity writes them out (the store clips ragged tiles); two staging buffers per warp.
    const int q = warp & 3;
    uint8_t* stage = smem + a.epi_off + q * 2 * 4096;
    float bias[16];
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + j));
      bias[j] = b4.x; bias[j + 1] = b4.y; bias[j + 2] = b4.z; bias[j + 3] = b4.w;
    }
    const float slope = a.act ? 0.01f : 1.0f;
    int it = 0;
    for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x, ++it) {
      const int buf = it & 1;
      int item, y0, x0;
      tile_origin(t, item, y0, x0);
      mbar_wait(&acc_full[buf], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * acc_cols);
#pragma unroll
      for (int ry = 0; ry < 2; ++ry) {
        uint32_t v[2][16];
        tmem_ld16_nowait(acc + (uint32_t)a.phase_col[ry * 2], v[0]);
        tmem_ld16_nowait(acc + (uint32_t)a.phase_col[ry * 2 + 1], v[1]);
        if (lane == 0) bulk_wait_read<1>();                // this buffer's previous store has left shared memory
        __syncwarp();
        tmem_wait_ld();
        if (ry == 1) {                                     // accumulators drained: the next tile but one may start
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cta(&acc_empty[buf]);
        }
        if (a.out_half) {                                  // fp16 output: 64-byte rows, 64-byte swizzle (16-byte chunk ^ bits 1-2 of the row:
          uint8_t* row = stage + ry * 4096 + lane * 64;    // unswizzled, 32 lanes x 16 B at a 64-byte stride were 16 wavefronts per store)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float f = __uint_as_float(v[c >> 1][(c & 1) * 8 + e]) + bias[(c & 1) * 8 + e];
              o[e] = f > 0.f ? f : slope * f;
            }
            const uint2 lo = pack_h4(make_float4(o[0], o[1], o[2], o[3])), hi = pack_h4(make_float4(o[4], o[5], o[6], o[7]));
            *reinterpret_cast<uint4*>(row + ((c ^ ((lane >> 1) & 3)) * 16)) = make_uint4(lo.x, lo.y, hi.x, hi.y);
          }
        } else {
        uint8_t* row = stage + ry * 4096 + lane * 128;
#pragma unroll
        for (int rx = 0; rx < 2; ++rx)
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float f = __uint_as_float(v[rx][j + e]) + bias[j + e];
              o[e] = f > 0.f ? f : slope * f;
            }
            const int chunk = (rx * 4 + (j >> 2)) ^ (lane & 7);
            *reinterpret_cast<float4*>(row + chunk * 16) = make_float4(o[0], o[1], o[2], o[3]);
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          tma_store_5d(&tmO, stage + ry * 4096, 0, x0, ry, y0 + q * 4, item);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    if (lane == 0) bulk_wait_read<0>();
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

#ifdef RCB_CONV_PROFILE
extern "C" int rcb_conv_prof_read(long long* host) { return (int)cudaMemcpyFromSymbol(host, rcb::rcb_conv_prof, sizeof(long long) * 256); }
#endif

// ----------------------------------------------------------------------------- backward --
template <int BK>   // 32: 128-byte rows / swizzle; 16: 64-byte rows / swizzle (oc = 16)
__global__ void __launch_bounds__(TC_THREADS)
upconv_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, ConvTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + a.bar_off);
  uint64_t* empty = full + TC_STAGES;
  uint64_t* acc_full = empty + TC_STAGES;
  uint32_t* tmem_slot = (uint32_t*)(acc_full + CT_MAX_PHASES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 1) CPROF(0);
  const PolyGeom& g = a.g;
  const int IC = g.ic;                                     // N of this GEMM
  int item0, z0, y0, x0;
  tile_origin(a, blockIdx.x, item0, z0, y0, x0);
  const int nseg = g.phases() * g.taps();
  const int nkb = nseg * a.kblocks;                        // kblocks = oc / BK
  constexpr uint32_t ROW_BYTES = BK * 4;
  const uint32_t a_bytes = (uint32_t)(a.t.tx * a.t.ty * a.t.tz * a.t.ni) * ROW_BYTES;
  const uint32_t b_bytes = (uint32_t)IC * ROW_BYTES;
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < IC) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(&acc_full[0], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // nested (phase, tap, k-block) counters: the single-thread loop carries no integer division
    int it = 0, seg = 0;
    for (int rz = 0; rz < g.fz; ++rz)
      for (int ry = 0; ry < g.fy; ++ry)
        for (int rx = 0; rx < g.fx; ++rx)
          for (int tz = 0; tz < g.Tz; ++tz)
            for (int ty = 0; ty < g.Ty; ++ty)
              for (int tx = 0; tx < g.Tx; ++tx, ++seg) {
                // output pixel feeding source pixel s through (phase, tap): (s - b - t) * f + r, stride f per source step
                const int cx = (x0 - g.base_x(rx) - tx) * g.fx + rx;
                const int cy = (y0 - g.base_y(ry) - ty) * g.fy + ry;
                const int cz = (z0 - g.base_z(rz) - tz) * g.fz + rz;
                for (int kb = 0; kb < a.kblocks; ++kb, ++it) {
                  const int s = it % TC_STAGES;
                  mbar_wait(&empty[s], ((it / TC_STAGES) & 1) ^ 1);
                  if (elect_one()) {
                    uint8_t* a_dst = smem + s * a.stage_bytes;
                    mbar_expect_tx(&full[s], a_bytes + b_bytes);
                    tma_load_5d(&tmA, &full[s], a_dst, kb * BK, cx, cy, cz, item0);
                    tma_load_2d(&tmB, &full[s], a_dst + a.b_off, kb * BK, seg * IC);
                  }
                  __syncwarp();
                }
              }
  } else if (warp == 1) {
    CPROF(1);
    const uint32_t idesc = idesc_tf32(IC);
    for (int it = 0; it < nkb; ++it) {
      const int s = it % TC_STAGES;
      mbar_wait(&full[s], (it / TC_STAGES) & 1);
      tc_fence_after();
      if (it < 90) CPROF(2 + it);
      if (elect_one()) {
        const uint32_t a_addr = smem_u32(smem + s * a.stage_bytes);
        const uint64_t da = BK == 32 ? smem_desc_sw128(a_addr) : smem_desc_sw64(a_addr);
        const uint64_t db = BK == 32 ? smem_desc_sw128(a_addr + a.b_off) : smem_desc_sw64(a_addr + a.b_off);
#pragma unroll
        for (int k = 0; k < BK / 8; ++k)
          umma_tf32(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (it | k) ? 1u : 0u);
        umma_commit(&empty[s]);
        if (it == nkb - 1) umma_commit(&acc_full[0]);
      }
      __syncwarp();
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    int item, z, y, x;
    const bool valid = tile_row(a, r, item0, z0, y0, x0, item, z, y, x);
    // Walk order of the coalesced write-back: consecutive lanes on consecutive 16-B chunks of a row.
    const int64_t m = valid ? (((int64_t)item * g.d + z) * g.h + y) * g.w + x : -1;
    const int lanes_per_row = IC / 4;                        // 16-B chunks per row: 4, 16 or 32
    const int rows_per_iter = 32 / lanes_per_row;
    const int sub = lane / lanes_per_row, ch = (lane % lanes_per_row) * 4;
    // The LeakyReLU mask of the producing stage does not depend on the MMAs: fetch it while they run and
    // keep one bit per element (4 bits per walk step).
    uint64_t lrelu_bits[2] = {0ull, 0ull};
    if (a.src_act) {
      const int nsteps = 32 / rows_per_iter;
      for (int s0 = 0; s0 < nsteps; s0 += 8) {            // eight loads in flight per lane
        float4 s4[8];
        uint2 h2[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int rr = (s0 + u) * rows_per_iter + sub;
          const int64_t mm = __shfl_sync(0xffffffffu, m, rr & 31);
          const bool ok = s0 + u < nsteps && sub < rows_per_iter && mm >= 0;
          if (a.act_half)
            h2[u] = ok ? __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(a.src_act) + mm * IC + ch)) : make_uint2(0u, 0u);
          else
            s4[u] = ok ? __ldg(reinterpret_cast<const float4*>(a.src_act + mm * IC + ch)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int step = s0 + u;
          uint64_t b;
          if (a.act_half)       // an fp16 is positive exactly when its bits are a positive int16
            b = (uint64_t)(((short)(h2[u].x & 0xffffu) > 0 ? 1u : 0u) | ((int)h2[u].x >= 0x10000 ? 2u : 0u) |
                           ((short)(h2[u].y & 0xffffu) > 0 ? 4u : 0u) | ((int)h2[u].y >= 0x10000 ? 8u : 0u));
          else
            b = (uint64_t)((s4[u].x > 0.f ? 1u : 0u) | (s4[u].y > 0.f ? 2u : 0u) | (s4[u].z > 0.f ? 4u : 0u) | (s4[u].w > 0.f ? 8u : 0u));
          if (step < 16) lrelu_bits[0] |= b << ((step & 15) * 4);
          else lrelu_bits[1] |= b << ((step & 15) * 4);
        }
      }
    }
    if (warp == 2) CPROF(100);
    mbar_wait(&acc_full[0], 0);
    tc_fence_after();
    if (warp == 2) CPROF(101);
    // All MMAs (and therefore all TMA fills) are done: the pipeline stages are free.  Each warp
    // stages its 32 x IC accumulator rows there (row-per-lane out of TMEM, padded rows), then
    // walks them again so that the gradient stores are coalesced.
    const int ldw = IC + 4;                                 // floats per staged row (bank-conflict-free float4 rows)
    float* wbuf = reinterpret_cast<float*>(smem) + (size_t)q * 32 * ldw;
    for (int c0 = 0; c0 < IC; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
      for (int j = 0; j < 16; j += 4)
        *reinterpret_cast<float4*>(wbuf + lane * ldw + c0 + j) =
            make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
    }
    __syncwarp();
    for (int s0 = 0; s0 * rows_per_iter < 32; s0 += 4) {
      float4 f[4];
      int64_t mm[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int rr = (s0 + u) * rows_per_iter + sub;
        mm[u] = __shfl_sync(0xffffffffu, m, rr & 31);
        if (sub >= rows_per_iter || rr >= 32) mm[u] = -1;
        f[u] = *reinterpret_cast<const float4*>(wbuf + (rr & 31) * ldw + ch);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (mm[u] >= 0) {
          float4 o = f[u];
          if (a.src_act) {
            const int step = s0 + u;
            const uint32_t b = (uint32_t)((step < 16 ? lrelu_bits[0] : lrelu_bits[1]) >> ((step & 15) * 4));
            o.x *= (b & 1u) ? 1.f : 0.01f; o.y *= (b & 2u) ? 1.f : 0.01f;
            o.z *= (b & 4u) ? 1.f : 0.01f; o.w *= (b & 8u) ? 1.f : 0.01f;
          }
          *reinterpret_cast<float4*>(a.out + mm[u] * IC + ch) = o;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

// ------------------------------------- forward, factor 2, 64 -> 64 channels, fp16 --
// The middle upsampler stage (x2, 3-tap, 64 -> 64) in the same persistent form, with fp16 activations in and out and
// fp16 weights: one 128-byte row per pixel is the whole K range, and the 16 (phase, tap) weight blocks (128 KB) stay
// in shared memory.  Accumulators: 4 phases x 64 = 256 TMEM columns, two sets.  A tile is 16 eight-pixel row groups,
// one halo row pitch (10 pixels) apart: 16 lines of one item (source grids with >= 16 lines), or 8 lines x 2 items
// interleaved line by line (8 x 8 grids; the TMA box puts the item dimension between x and y), so that the same
// shifted-descriptor trick serves both.  Output rows (pixel, column parity) are staged as swizzled 128-byte rows and
// leave through one TMA store per warp and line parity.
constexpr int B2_THREADS = 320;                                 // warp 0: TMA, warp 1: MMA, warps 2-9: epilogue
constexpr int F2W_STAGES = 2, F2W_STAGE_BYTES = 26 * 1024, F2W_W_BLOCK = 64 * 128;
struct ConvF2WArgs {
  PolyGeom g;
  int items, tiles_x, tiles_y, n_tiles;
  int ipt;                              // items per tile: 1 (8 px x 16 lines) or 2 (8 x 8 grids)
  int act, halo_bytes;
  int a_off, epi_off, bias_off, bar_off;
  int grp_shift[F2_GROUPS], grp_col[F2_GROUPS], grp_n[F2_GROUPS], grp_row[F2_GROUPS];
  int mem_phase[16], mem_tap[16];
  int phase_col[4];
  const float* bias;
};

__global__ void __launch_bounds__(B2_THREADS, 1)
upconv_fwd_f2w_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmO, const __grid_constant__ ConvF2WArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* a_full = (uint64_t*)(smem + a.bar_off);        // [F2W_STAGES]
  uint64_t* a_empty = a_full + F2W_STAGES;
  uint64_t* acc_full = a_empty + F2W_STAGES;                // [2]
  uint64_t* acc_empty = acc_full + 2;                       // [2], one arrival per epilogue warp
  uint64_t* w_full = acc_empty + 2;
  uint32_t* tmem_slot = (uint32_t*)(w_full + 1);
  float* bias_s = reinterpret_cast<float*>(smem + a.bias_off);
  constexpr int OC = 64, ACC_COLS = 4 * OC;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const PolyGeom& g = a.g;

  if (threadIdx.x == 0) {
    for (int i = 0; i < F2W_STAGES; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8); }
    mbar_init(w_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < OC) bias_s[threadIdx.x] = a.bias[threadIdx.x];
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // tile -> first item, first line, first pixel
  auto tile_origin = [&](int t, int& item, int& y0, int& x0) {
    if (a.ipt == 2) { item = 2 * t; y0 = 0; x0 = 0; return; }
    x0 = (t % a.tiles_x) * 8; t /= a.tiles_x;
    y0 = (t % a.tiles_y) * 16; t /= a.tiles_y;
    item = t;
  };

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(w_full, 16u * F2W_W_BLOCK);
      for (int b = 0; b < 16; ++b)
        tma_load_2d(&tmB, w_full, smem + b * F2W_W_BLOCK, a.mem_tap[b] * g.ic, a.mem_phase[b] * OC);
    }
    __syncwarp();
    int s = 0;
    uint32_t par = 1;
    for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
      int item, y0, x0;
      tile_origin(t, item, y0, x0);
      mbar_wait(&a_empty[s], par);
      if (elect_one()) {
        mbar_expect_tx(&a_full[s], (uint32_t)a.halo_bytes);
        if (a.ipt == 2) tma_load_4d(&tmA, &a_full[s], smem + a.a_off + s * F2W_STAGE_BYTES, 0, -1, item, -1);
        else tma_load_4d(&tmA, &a_full[s], smem + a.a_off + s * F2W_STAGE_BYTES, 0, x0 - 1, y0 - 1, item);
      }
      __syncwarp();
      if (++s == F2W_STAGES) { s = 0; par ^= 1; }
    }
  } else if (warp == 1) {
    const uint64_t da_hi = ((uint64_t)1 << 16) | ((uint64_t)((HALO_PITCH * 128) >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    uint32_t idesc[F2_GROUPS];
#pragma unroll
    for (int j = 0; j < F2_GROUPS; ++j) idesc[j] = idesc_f16_m128(a.grp_n[j]);
    const uint32_t w_addr = smem_u32(smem);
    mbar_wait(w_full, 0);
    int s = 0;
    uint32_t par = 0;
    int it = 0;
    for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x, ++it) {
      const int buf = it & 1;
      mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t acc = tmem_base + (uint32_t)(buf * ACC_COLS);
      mbar_wait(&a_full[s], par);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_addr = smem_u32(smem + a.a_off + s * F2W_STAGE_BYTES);
#pragma unroll
        for (int j = 0; j < F2_GROUPS; ++j) {
          const uint64_t da = da_hi | (uint64_t)(((a_addr + (uint32_t)a.grp_shift[j] * 128u) & 0x3FFFF) >> 4);
          const uint64_t db = smem_desc_sw128(w_addr + (uint32_t)a.grp_row[j] * 128u);
#pragma unroll
          for (int k = 0; k < 4; ++k)     // group 0 covers every column, so it alone starts the accumulation
            umma_f16_ss(acc + (uint32_t)a.grp_col[j], da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc[j], (j | k) ? 1u : 0u);
        }
        umma_commit(&a_empty[s]);
        umma_commit(&acc_full[buf]);
      }
      __syncwarp();
      if (++s == F2W_STAGES) { s = 0; par ^= 1; }
    }
  } else {
    // EIGHT epilogue warps: warp (q, hh) converts the 32 pixel rows of TMEM lane quarter q and the 32 channels of half hh,
    // one (line parity, column parity) phase at a time: 32 rows x 32 channels as fp16 = 32 swizzled 64-byte rows = one TMA
    // store (two when the tile holds two items); two 2 KB buffers per warp alternate, and the next phase's accumulators
    // are already on their way out of TMEM while this one is converted.  (Four warps with 256 columns per lane each were
    // the bottleneck of this kernel: the MMAs of a tile finished long before its conversion.)
    const int q = warp & 3, hh = (warp - 2) >> 2, ew = warp - 2;
    uint8_t* stage = smem + a.epi_off + ew * 4096;
    const int gl = lane >> 3, px = lane & 7;
    // staged row: [line][px] (one item) or [item][line][px] (two items: group = line * 2 + item)
    const int row = (a.ipt == 2 ? ((gl & 1) * 2 + (gl >> 1)) : gl) * 8 + px;
    const float slope = a.act ? 0.01f : 1.0f;
    const float* bias_h = bias_s + hh * 32;
    int it = 0;
    for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x, ++it) {
      const int buf = it & 1;
      int item, y0, x0;
      tile_origin(t, item, y0, x0);
      mbar_wait(&acc_full[buf], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * ACC_COLS + hh * 32);
      uint32_t v[2][2][16];
#pragma unroll
      for (int c = 0; c < 2; ++c) tmem_ld16_nowait(acc + (uint32_t)(a.phase_col[0] + c * 16), v[0][c]);
      tmem_wait_ld();
#pragma unroll
      for (int ph = 0; ph < 4; ++ph) {                     // ph = ry * 2 + rx
        const int ry = ph >> 1, rx = ph & 1;
        if (ph < 3) {
#pragma unroll
          for (int c = 0; c < 2; ++c) tmem_ld16_nowait(acc + (uint32_t)(a.phase_col[ph + 1] + c * 16), v[(ph + 1) & 1][c]);
        }
        if (lane == 0) bulk_wait_read<1>();                // the store that last read this buffer has left it
        __syncwarp();
        uint8_t* dst = stage + (ph & 1) * 2048 + row * 64;
#pragma unroll
        for (int c = 0; c < 4; ++c) {                      // chunk c: channels hh * 32 + 8c .. + 7 as fp16
          const float4 b0 = *reinterpret_cast<const float4*>(bias_h + c * 8), b1 = *reinterpret_cast<const float4*>(bias_h + c * 8 + 4);
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
          float o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float f = __uint_as_float(v[ph & 1][c >> 1][(c & 1) * 8 + e]) + bb[e];
            o[e] = f > 0.f ? f : slope * f;
          }
          const uint2 lo = pack_h4(make_float4(o[0], o[1], o[2], o[3])), hi = pack_h4(make_float4(o[4], o[5], o[6], o[7]));
          *reinterpret_cast<uint4*>(dst + ((c ^ ((row >> 1) & 3)) * 16)) = make_uint4(lo.x, lo.y, hi.x, hi.y);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          const uint8_t* sb = stage + (ph & 1) * 2048;
          if (a.ipt == 2) {
            tma_store_5d(&tmO, sb, hh * 32, rx, 0, ry, item * g.h + 2 * q);
            if (item + 1 < a.items) tma_store_5d(&tmO, sb + 1024, hh * 32, rx, 0, ry, (item + 1) * g.h + 2 * q);
          } else if (y0 + 4 * q < g.h) {                    // h % 4 == 0: a warp's four lines are all inside or all outside
            tma_store_5d(&tmO, sb, hh * 32, rx, x0, ry, item * g.h + y0 + 4 * q);
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (ph < 3) {
          tmem_wait_ld();
        } else {                                           // accumulators drained: the next tile but one may start
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cta(&acc_empty[buf]);
        }
      }
    }
    if (lane == 0) bulk_wait_read<0>();
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// --------------------------------------- data gradient, factor 2, resident weights --
// The adjoint of the stage above.  A source pixel s receives from the 4 x 4 output pixels 2s - 1 + (a, b):
//   d_src[s][ic] = mask(s, ic) * sum_{a, b, oc} d_out[2s - 1 + (a, b)][oc] * W[a][b][oc][ic]
// Seen as (item, y, line parity ry, x, (column parity rx, oc)) the output gradient has one dense 128-byte row per
// source pixel and line parity, so the (8 + 2) x (16 + 2) neighbourhood of a tile travels as TWO halo boxes (ry = 0, 1)
// and every (a, b) product reads a row-shifted window of one of them: a <-> (ry, dy) = (1,-1) (0,0) (1,0) (0,+1), and
// likewise b <-> (rx, dx), where rx picks the 64-byte half of the row.  The 64 x 256 weight matrix (64 KB) stays in
// shared memory; CTAs are persistent, accumulators double-buffered, the result leaves through TMA stores.
constexpr int B2_STAGES = 4;
constexpr int B2_MSTAGES = 2;                                   // mask tiles in flight (fp16 variant)
constexpr int HALO_BUF_H = 12 * 1024;                           // an fp16 halo box (18 x 10 rows of 64 bytes)
struct ConvB2Args {
  PolyGeom g;
  int items, tiles_x, tiles_y, n_tiles;
  int act_kind;                          // LeakyReLU mask of the producing stage: 0 none, 1 fp32, 2 fp16 activations
  int a_off, m_off, epi_off, bar_off;
  int out_half;                          // d_src leaves as fp16, multiplied by out_scale (tmO then maps fp16 rows)
  float out_scale;
  const void* src_act;
};

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

// w_bwd_k[ic][a * 64 + b * 16 + oc] = w_eff[phase (ry_a, rx_b)][tap (ty_a, tx_b)][ic][oc]
__global__ void fold_bwd_f2_kernel(const float* __restrict__ w_eff, float* __restrict__ w_bwd_k, int ic, int oc) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int K = 16 * oc;
  if (e >= ic * K) return;
  const int c = e / K, k = e - c * K;
  const int a = k / (4 * oc), b = (k / oc) & 3, o = k % oc;
  const int ry = (a == 0 || a == 2) ? 1 : 0, ty = (a < 2) ? 1 : 0;
  const int rx = (b == 0 || b == 2) ? 1 : 0, tx = (b < 2) ? 1 : 0;
  w_bwd_k[e] = w_eff[((size_t)((ry * 2 + rx) * 4 + ty * 2 + tx) * ic + c) * oc + o];
}

// HIN: d_out is fp16 (64-byte rows, 64-byte swizzle, one kind::f16 K step per product) and so are the weights
template <bool HIN>
__global__ void __launch_bounds__(B2_THREADS, 1)
upconv_bwd_f2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmM,
                     const __grid_constant__ ConvB2Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* a_full = (uint64_t*)(smem + a.bar_off);        // [B2_STAGES]
  uint64_t* a_empty = a_full + B2_STAGES;
  uint64_t* acc_full = a_empty + B2_STAGES;                 // [2]
  uint64_t* acc_empty = acc_full + 2;                       // [2], one arrival per epilogue warp
  uint64_t* w_full = acc_empty + 2;
  uint64_t* m_full = w_full + 1;                            // [B2_MSTAGES] mask tiles (HIN with fp16 activations)
  uint64_t* m_empty = m_full + B2_MSTAGES;                  // one arrival per epilogue warp
  uint32_t* tmem_slot = (uint32_t*)(m_empty + B2_MSTAGES);
  const bool tma_mask = HIN && a.act_kind == 2;
  constexpr int IC = 64, W_BLOCK = IC * 128;                // one 128-byte-wide K block of the weights
  constexpr int W_BLOCKS = HIN ? 4 : 8;                     // K = 256: 64 halves or 32 floats per block row
  constexpr int ROW_BYTES = HIN ? 64 : 128;                 // one source pixel and line parity of d_out
  constexpr int A_BUF = HIN ? HALO_BUF_H : HALO_BUF;
  constexpr uint32_t TMEM_COLS = 128;                       // two accumulator sets of 64 columns

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const PolyGeom& g = a.g;

  if (threadIdx.x == 0) {
    for (int i = 0; i < B2_STAGES; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8); }
    mbar_init(w_full, 1);
    for (int i = 0; i < B2_MSTAGES; ++i) { mbar_init(&m_full[i], 1); mbar_init(&m_empty[i], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto tile_origin = [&](int t, int& item, int& y0, int& x0) {
    x0 = (t % a.tiles_x) * 8; t /= a.tiles_x;
    y0 = (t % a.tiles_y) * 16; t /= a.tiles_y;
    item = t;
  };

  if (warp == 0) {
    // ===== TMA producer: the weights once, then the two halo boxes (line parity 0, 1) of every tile
    if (elect_one()) {
      mbar_expect_tx(w_full, (uint32_t)(W_BLOCKS * W_BLOCK));
      for (int kb = 0; kb < W_BLOCKS; ++kb) tma_load_2d(&tmB, w_full, smem + kb * W_BLOCK, kb * (HIN ? 64 : 32), 0);
    }
    __syncwarp();
    int s = 0, it = 0;
    uint32_t par = 1;
    for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x, ++it) {
      int item, y0, x0;
      tile_origin(t, item, y0, x0);
      if (tma_mask) {
        // the tile's 128 rows of the producing stage's fp16 activations (only their signs are read): they ride the same
        // pipeline as the operands, so no epilogue warp ever waits on a global load
        const int ms = it % B2_MSTAGES;
        mbar_wait(&m_empty[ms], ((it / B2_MSTAGES) & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&m_full[ms], 128u * 128u);
          tma_load_4d(&tmM, &m_full[ms], smem + a.m_off + ms * 16384, 0, x0, y0, item);
        }
        __syncwarp();
      }
      for (int ry = 0; ry < 2; ++ry) {
        mbar_wait(&a_empty[s], par);
        if (elect_one()) {
          mbar_expect_tx(&a_full[s], (uint32_t)(HALO_LINES * HALO_PITCH * ROW_BYTES));
          tma_load_5d(&tmA, &a_full[s], smem + a.a_off + s * A_BUF, 0, x0 - 1, ry, y0 - 1, item);
        }
        __syncwarp();
        if (++s == B2_STAGES) { s = 0; par ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer
    const uint64_t da_hi = ((uint64_t)1 << 16) | ((uint64_t)((HALO_PITCH * ROW_BYTES) >> 4) << 32) | ((uint64_t)1 << 46) |
                           ((uint64_t)(HIN ? 4 : 2) << 61);
    const uint32_t idesc = HIN ? idesc_f16_m128(IC) : idesc_tf32(IC);
    const uint32_t w_addr = smem_u32(smem);
    mbar_wait(w_full, 0);
    int s = 0;
    uint32_t par = 0;
    int it = 0;
    for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x, ++it) {
      const int buf = it & 1;
      mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t acc = tmem_base + (uint32_t)(buf * IC);
#pragma unroll
      for (int ry = 0; ry < 2; ++ry) {
        mbar_wait(&a_full[s], par);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(smem + a.a_off + s * A_BUF);
#pragma unroll
          for (int ai = 0; ai < 2; ++ai) {
            // line parity 0 serves a = 1 (dy = 0) and a = 3 (dy = +1); parity 1 serves a = 2 (dy = 0) and a = 0 (dy = -1)
            const int aa = ry == 0 ? (ai == 0 ? 1 : 3) : (ai == 0 ? 2 : 0);
            const int dy = ry == 0 ? (ai == 0 ? 0 : 1) : (ai == 0 ? 0 : -1);
#pragma unroll
            for (int b = 0; b < 4; ++b) {
              const int rx = (b == 0 || b == 2) ? 1 : 0;
              const int dx = b == 0 ? -1 : (b == 3 ? 1 : 0);
              const uint32_t shift_rows = (uint32_t)((1 + dy) * HALO_PITCH + (1 + dx));
              if (HIN) {      // 16 oc = 32 bytes = one K step; the weight block of `aa` holds all four b
                const uint64_t da = (da_hi | (uint64_t)(((a_addr + shift_rows * 64u) & 0x3FFFF) >> 4)) + (uint64_t)(rx * 2);
                const uint64_t db = smem_desc_sw128(w_addr + (uint32_t)(aa * W_BLOCK)) + (uint64_t)(b * 2);
                umma_f16_ss(acc, da, db, idesc, (ry | ai | b) ? 1u : 0u);
              } else {
                const uint64_t da = (da_hi | (uint64_t)(((a_addr + shift_rows * 128u) & 0x3FFFF) >> 4)) + (uint64_t)(rx * 4);
                const uint64_t db = smem_desc_sw128(w_addr + (uint32_t)((aa * 2 + (b >> 1)) * W_BLOCK)) + (uint64_t)((b & 1) * 4);
#pragma unroll
                for (int k = 0; k < 2; ++k)
                  umma_tf32(acc, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (ry | ai | b | k) ? 1u : 0u);
              }
            }
          }
          umma_commit(&a_empty[s]);
          if (ry == 1) umma_commit(&acc_full[buf]);
        }
        __syncwarp();
        if (++s == B2_STAGES) { s = 0; par ^= 1; }
      }
    }
  } else {
    // ===== epilogue, EIGHT warps: accumulator row m = (line m / 8, pixel m % 8); warp (q, hh) owns the 32 rows of TMEM
    // lane quarter q and the 32 channels of half hh.  (With four warps -- one per scheduler -- converting a whole
    // 64-channel row each, the ~700 dependent instructions per tile of this chain, not HBM, bounded the kernel.)
    const int q = warp & 3, hh = (warp - 2) >> 2, ew = warp - 2;
    uint8_t* stage = smem + a.epi_off + ew * 2 * 4096;
    int it = 0;
    for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x, ++it) {
      const int buf = it & 1;
      int item, y0, x0;
      tile_origin(t, item, y0, x0);
      const int m = q * 32 + lane;
      // LeakyReLU mask: the 32 sign bits of this lane's half row
      uint32_t bits = ~0u;
      if (tma_mask) {
        const int ms = it % B2_MSTAGES;
        mbar_wait(&m_full[ms], (it / B2_MSTAGES) & 1);
        const uint8_t* row = smem + a.m_off + ms * 16384 + m * 128;
        bits = 0u;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint4 hv = *reinterpret_cast<const uint4*>(row + (((hh * 4 + c) ^ (m & 7)) * 16));
          const uint32_t w4[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
          for (int e = 0; e < 4; ++e)                       // an fp16 is positive exactly when its bits are a positive int16
            bits |= (((short)(w4[e] & 0xffffu) > 0 ? 1u : 0u) | ((int)w4[e] >= 0x10000 ? 2u : 0u)) << (c * 8 + e * 2);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive_cta(&m_empty[ms]);
      } else if (a.act_kind == 2) {                         // fp16 activations straight from global memory
        const int y = y0 + (m >> 3), x = x0 + (m & 7);
        bits = 0u;
        if (y < g.h && x < g.w) {
          const uint4* p = reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(a.src_act) +
                                                          (((int64_t)item * g.h + y) * g.w + x) * IC + hh * 32);
          uint4 hv[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) hv[c] = __ldg(p + c);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint32_t w4[4] = {hv[c].x, hv[c].y, hv[c].z, hv[c].w};
#pragma unroll
            for (int e = 0; e < 4; ++e)
              bits |= (((short)(w4[e] & 0xffffu) > 0 ? 1u : 0u) | ((int)w4[e] >= 0x10000 ? 2u : 0u)) << (c * 8 + e * 2);
          }
        }
      } else if (a.act_kind == 1) {                         // fp32 activations: each lane walks its own half row
        bits = 0u;
        const int y = y0 + (m >> 3), x = x0 + (m & 7);
        if (y < g.h && x < g.w) {
          const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.src_act) +
                                                            (((int64_t)item * g.h + y) * g.w + x) * IC + hh * 32);
          float4 f[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) f[c] = __ldg(p + c);
#pragma unroll
          for (int c = 0; c < 8; ++c)
            bits |= ((f[c].x > 0.f ? 1u : 0u) | (f[c].y > 0.f ? 2u : 0u) | (f[c].z > 0.f ? 4u : 0u) | (f[c].w > 0.f ? 8u : 0u)) << (c * 4);
        }
      }
      mbar_wait(&acc_full[buf], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * IC + hh * 32);
      uint32_t v[2][16];
      tmem_ld16_nowait(acc, v[0]);
      tmem_ld16_nowait(acc + 16u, v[1]);
      if (lane == 0) bulk_wait_read<1>();                  // the store that last read this buffer has left shared memory
      __syncwarp();
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cta(&acc_empty[buf]);
      uint8_t* sbuf = stage + buf * 4096;
      if (a.out_half) {                                     // 32 channels = one 64-byte fp16 row, 64-byte swizzle
        uint8_t* row = sbuf + lane * 64;
        const float s1 = a.out_scale, s0 = 0.01f * a.out_scale;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e)
            o[e] = __uint_as_float(v[c >> 1][(c & 1) * 8 + e]) * (((bits >> (c * 8 + e)) & 1u) ? s1 : s0);
          const uint2 lo = pack_h4_sat(make_float4(o[0], o[1], o[2], o[3])), hi = pack_h4_sat(make_float4(o[4], o[5], o[6], o[7]));
          *reinterpret_cast<uint4*>(row + ((c ^ ((lane >> 1) & 3)) * 16)) = make_uint4(lo.x, lo.y, hi.x, hi.y);
        }
      } else {
        uint8_t* row = sbuf + lane * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            o[e] = __uint_as_float(v[c >> 2][(c & 3) * 4 + e]) * (((bits >> (c * 4 + e)) & 1u) ? 1.f : 0.01f);
          *reinterpret_cast<float4*>(row + ((c ^ (lane & 7)) * 16)) = make_float4(o[0], o[1], o[2], o[3]);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        tma_store_4d(&tmO, sbuf, hh * 32, x0, y0 + q * 4, item);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    if (lane == 0) bulk_wait_read<0>();
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------- data gradient, factor 2, 64 -> 64 channels, fp16 operands, resident weights --
// Same adjoint as above for the middle stage of the 2-D upsamplers.  With 64 output channels a (line parity, column
// parity) phase of the output gradient has one dense 128-byte fp16 row per source pixel, so the neighbourhood of a tile
// travels as FOUR halo boxes (TMA element strides of 2 along x and y pick a phase), each feeding the 2 x 2 products
// (a, b) whose parities it matches through row-shifted descriptors.  The 16 weight blocks W[a][b] (64 ic rows x 64 oc,
// 128 KB as fp16) stay resident.  A tile is 16 lines x 8 pixels of one item, or two whole 8 x 8 items interleaved line
// by line (the box puts the item dimension between x and y, as in the forward kernel above).  The incoming gradient
// is fp16 in units of 1 / out_scale_inv (written by the kernel above with out_half); the result is fp32, true units.
constexpr int B2W_STAGES = 3;
struct ConvB2WArgs {
  PolyGeom g;
  int items, tiles_x, tiles_y, n_tiles;
  int ipt;                               // items per tile: 1 (8 px x 16 lines) or 2 (8 x 8 grids)
  int halo_bytes;
  int a_off, epi_off, bar_off;
  float out_scale_inv;
  int out_half;                          // d_src is fp16 (saturating), else fp32
  const void* src_act;                   // fp16 activations of the producing stage (LeakyReLU mask) or null
  void* d_src;
};

__global__ void __launch_bounds__(B2_THREADS, 1)
upconv_bwd_f2w_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmO, const __grid_constant__ ConvB2WArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* a_full = (uint64_t*)(smem + a.bar_off);        // [B2W_STAGES]
  uint64_t* a_empty = a_full + B2W_STAGES;
  uint64_t* acc_full = a_empty + B2W_STAGES;                // [2]
  uint64_t* acc_empty = acc_full + 2;                       // [2], one arrival per epilogue warp
  uint64_t* w_full = acc_empty + 2;
  uint32_t* tmem_slot = (uint32_t*)(w_full + 1);
  constexpr int IC = 64;
  constexpr uint32_t TMEM_COLS = 128;                       // two accumulator sets of 64 columns

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const PolyGeom& g = a.g;

  if (threadIdx.x == 0) {
    for (int i = 0; i < B2W_STAGES; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8); }
    mbar_init(w_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto tile_origin = [&](int t, int& item, int& y0, int& x0) {
    if (a.ipt == 2) { item = 2 * t; y0 = 0; x0 = 0; return; }
    x0 = (t % a.tiles_x) * 8; t /= a.tiles_x;
    y0 = (t % a.tiles_y) * 16; t /= a.tiles_y;
    item = t;
  };

  if (warp == 0) {
    // ===== TMA producer: the weights once, then the four phase boxes (line parity, column parity) of every tile
    if (elect_one()) {
      mbar_expect_tx(w_full, 16u * F2W_W_BLOCK);
      for (int b = 0; b < 16; ++b) tma_load_2d(&tmB, w_full, smem + b * F2W_W_BLOCK, b * 64, 0);
    }
    __syncwarp();
    int s = 0;
    uint32_t par = 1;
    for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
      int item, y0, x0;
      tile_origin(t, item, y0, x0);
      for (int ph = 0; ph < 4; ++ph) {
        const int ry = ph >> 1, rx = ph & 1;
        mbar_wait(&a_empty[s], par);
        if (elect_one()) {
          mbar_expect_tx(&a_full[s], (uint32_t)a.halo_bytes);
          uint8_t* dst = smem + a.a_off + s * F2W_STAGE_BYTES;
          if (a.ipt == 2) tma_load_4d(&tmA, &a_full[s], dst, 0, rx - 2, item, ry - 2);
          else tma_load_4d(&tmA, &a_full[s], dst, 0, 2 * (x0 - 1) + rx, 2 * (y0 - 1) + ry, item);
        }
        __syncwarp();
        if (++s == B2W_STAGES) { s = 0; par ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer
    const uint64_t da_hi = ((uint64_t)1 << 16) | ((uint64_t)((HALO_PITCH * 128) >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    const uint32_t idesc = idesc_f16_m128(IC);
    const uint32_t w_addr = smem_u32(smem);
    const int dy_rows = HALO_PITCH * a.ipt;
    mbar_wait(w_full, 0);
    int s = 0;
    uint32_t par = 0;
    int it = 0;
    for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x, ++it) {
      const int buf = it & 1;
      mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t acc = tmem_base + (uint32_t)(buf * IC);
#pragma unroll
      for (int ph = 0; ph < 4; ++ph) {
        const int ry = ph >> 1, rx = ph & 1;
        mbar_wait(&a_full[s], par);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(smem + a.a_off + s * F2W_STAGE_BYTES);
#pragma unroll
          for (int ai = 0; ai < 2; ++ai) {
            // parity 0 serves index 1 (shift 0) and 3 (shift +1); parity 1 serves 2 (shift 0) and 0 (shift -1)
            const int aa = ry == 0 ? (ai == 0 ? 1 : 3) : (ai == 0 ? 2 : 0);
            const int dy = ry == 0 ? (ai == 0 ? 0 : 1) : (ai == 0 ? 0 : -1);
#pragma unroll
            for (int bi = 0; bi < 2; ++bi) {
              const int bb = rx == 0 ? (bi == 0 ? 1 : 3) : (bi == 0 ? 2 : 0);
              const int dx = rx == 0 ? (bi == 0 ? 0 : 1) : (bi == 0 ? 0 : -1);
              const uint32_t shift_rows = (uint32_t)((1 + dy) * dy_rows + (1 + dx));
              const uint64_t da = da_hi | (uint64_t)(((a_addr + shift_rows * 128u) & 0x3FFFF) >> 4);
              const uint64_t db = smem_desc_sw128(w_addr + (uint32_t)((aa * 4 + bb) * F2W_W_BLOCK));
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16_ss(acc, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (ph | ai | bi | k) ? 1u : 0u);
            }
          }
          umma_commit(&a_empty[s]);
          if (ph == 3) umma_commit(&acc_full[buf]);
        }
        __syncwarp();
        if (++s == B2W_STAGES) { s = 0; par ^= 1; }
      }
    }
  } else {
    // ===== epilogue, eight warps: accumulator row m = (line group m / 8, pixel m % 8); warp (q, hh) owns the 32 rows of
    // TMEM lane quarter q -- four lines of one item, or two lines of each of the tile's two items (group = line * 2 +
    // item) -- and the 32 channels of half hh.  Each lane holds 128 contiguous bytes of the fp32 result: they go straight
    // from registers to global memory (eight 16-byte stores), which leaves the shared memory to a third operand stage.
    const int q = warp & 3, hh = (warp - 2) >> 2, ew = warp - 2;
    uint8_t* scratch = smem + a.epi_off + ew * 128;
    uint4 h[4];                                             // mask loads run one tile ahead
    auto load_mask = [&](int t) {
      int item, y0, x0;
      tile_origin(t, item, y0, x0);
#pragma unroll
      for (int i = 0; i < 4; ++i) {                         // load i: line group i of the warp's four, 8 pixels x 4 chunks of 8 halves
        const int yy = a.ipt == 2 ? 2 * q + (i >> 1) : y0 + q * 4 + i;
        const int ii = a.ipt == 2 ? item + (i & 1) : item;
        const int xx = x0 + (lane >> 2);
        h[i] = (t < a.n_tiles && yy < g.h && xx < g.w && ii < a.items)
                   ? __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(a.src_act) +
                                                          (((int64_t)ii * g.h + yy) * g.w + xx) * IC) + hh * 4 + (lane & 3))
                   : make_uint4(0u, 0u, 0u, 0u);
      }
    };
    if (a.src_act != nullptr) load_mask(blockIdx.x);
    int it = 0;
    for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x, ++it) {
      const int buf = it & 1;
      int item, y0, x0;
      tile_origin(t, item, y0, x0);
      uint32_t bits = ~0u;
      if (a.src_act != nullptr) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t w4[4] = {h[i].x, h[i].y, h[i].z, h[i].w};
          uint32_t b8 = 0u;
#pragma unroll
          for (int e = 0; e < 4; ++e)                       // an fp16 is positive exactly when its bits are a positive int16
            b8 |= (((short)(w4[e] & 0xffffu) > 0 ? 1u : 0u) | ((int)w4[e] >= 0x10000 ? 2u : 0u)) << (e * 2);
          scratch[i * 32 + lane] = (uint8_t)b8;             // = [row i * 8 + lane / 4][chunk lane % 4]
        }
        __syncwarp();
        bits = *reinterpret_cast<const uint32_t*>(scratch + lane * 4);
        __syncwarp();
      }
      mbar_wait(&acc_full[buf], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * IC + hh * 32);
      uint32_t v[2][16];
      tmem_ld16_nowait(acc, v[0]);
      tmem_ld16_nowait(acc + 16u, v[1]);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cta(&acc_empty[buf]);
      if (a.src_act != nullptr) load_mask(t + (int)gridDim.x);
      // this lane's pixel
      const int gi = lane >> 3, px = lane & 7;
      const int yy = a.ipt == 2 ? 2 * q + (gi >> 1) : y0 + q * 4 + gi;
      const int ii = a.ipt == 2 ? item + (gi & 1) : item;
      const int xx = x0 + px;
      if (yy < g.h && xx < g.w && ii < a.items) {
        const int64_t off = (((int64_t)ii * g.h + yy) * g.w + xx) * IC + hh * 32;
        const float s1 = a.out_scale_inv, s0 = 0.01f * a.out_scale_inv;
        if (a.out_half) {
          uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(a.d_src) + off);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e)
              o[e] = __uint_as_float(v[c >> 1][(c & 1) * 8 + e]) * (((bits >> (c * 8 + e)) & 1u) ? s1 : s0);
            const uint2 lo = pack_h4_sat(make_float4(o[0], o[1], o[2], o[3])), hi = pack_h4_sat(make_float4(o[4], o[5], o[6], o[7]));
            dst[c] = make_uint4(lo.x, lo.y, hi.x, hi.y);
          }
        } else {
          float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(a.d_src) + off);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)
              o[e] = __uint_as_float(v[c >> 2][(c & 3) * 4 + e]) * (((bits >> (c * 4 + e)) & 1u) ? s1 : s0);
            dst[c] = make_float4(o[0], o[1], o[2], o[3]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---------------------------------------------------------------------------------- host --
static int make_geom_tc(const rcb_upconv_geom* g, PolyGeom* out) {
  RCB_CHECK_ARG(g != nullptr, "upconv_tc: null geometry");
  int pz = (g->kz - 1) / 2, py = (g->ky - 1) / 2, px = (g->kx - 1) / 2;
  out->d = g->d; out->h = g->h; out->w = g->w; out->fz = g->fz; out->fy = g->fy; out->fx = g->fx;
  out->pz = pz; out->py = py; out->px = px;
  out->Tz = pz == 0 ? 1 : 2; out->Ty = py == 0 ? 1 : 2; out->Tx = px == 0 ? 1 : 2;
  out->ic = g->ic; out->oc = g->oc;
  return 0;
}

// 128 source pixels per tile, widest along x; `limit` caps the box extent per axis
// (strided data-gradient boxes must stay <= 256 elements: extent * factor <= 256)
static ConvTile choose_tile(const PolyGeom& g, int items, int lim_x, int lim_y, int lim_z) {
  ConvTile t;
  auto mn = [](int a, int b) { return a < b ? a : b; };
  t.tx = mn(mn(g.w, 128), lim_x);
  t.ty = mn(mn(g.h, 128 / t.tx), lim_y);
  t.tz = mn(mn(g.d, 128 / (t.tx * t.ty)), lim_z);
  t.ni = mn(items, 128 / (t.tx * t.ty * t.tz));
  if (t.ni < 1) t.ni = 1;
  t.ntx = ceil_div(g.w, t.tx); t.nty = ceil_div(g.h, t.ty); t.ntz = ceil_div(g.d, t.tz);
  return t;
}

// 5-D map over a channel-last activation tensor (items, D, H, W, C), optional traversal strides
static int make_map_5d(CUtensorMap* map, const void* base, int items, int D, int H, int W, int C, const ConvTile& t,
                       int box_c, int sz, int sy, int sx, CUtensorMapSwizzle swz, int esize = 4) {
  EncodeTiledFn enc = tc_get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return -1; }
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)items};
  const cuuint64_t es = (cuuint64_t)esize;
  cuuint64_t strides[4] = {(cuuint64_t)C * es, (cuuint64_t)W * C * es, (cuuint64_t)H * W * C * es, (cuuint64_t)D * H * W * C * es};
  cuuint32_t box[5] = {(cuuint32_t)box_c, (cuuint32_t)(t.tx * sx), (cuuint32_t)(t.ty * sy), (cuuint32_t)(t.tz * sz), (cuuint32_t)t.ni};
  cuuint32_t estr[5] = {1, (cuuint32_t)sx, (cuuint32_t)sy, (cuuint32_t)sz, 1};
  // a strided box of extent n*s covers exactly n elements only if it does not run past the last one
  for (int i = 1; i < 4; ++i) box[i] = box[i] - (estr[i] - 1);
  CUresult r = enc(map, esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(5d) failed with CUresult %d", (int)r); return -1; }
  return 0;
}

static int make_map_b(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int box_rows, int box_cols,
                      CUtensorMapSwizzle swz, int esize = 4) {
  EncodeTiledFn enc = tc_get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return -1; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * (cuuint64_t)esize};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(weights) failed with CUresult %d", (int)r); return -1; }
  return 0;
}

template <class K>
static int opt_in_smem(K kernel, const char* name, int bytes = 200 * 1024) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) { set_error("%s: smem opt-in failed: %s", name, cudaGetErrorString(e)); return -1; }
  return 0;
}

// x2 stage with 16 output channels: persistent CTAs, resident weights, phases that share a shift in one MMA
static bool f2_eligible(const PolyGeom& g, bool half) {
  return g.d == 1 && g.Tz == 1 && g.Ty == 2 && g.Tx == 2 && g.h >= 16 && g.fy == 2 && g.fx == 2 && g.py == 1 && g.px == 1 &&
         g.oc == 16 && g.ic <= 64 && g.ic % (half ? 64 : 32) == 0;
}
static int launch_f2(const void* src, const void* w_eff_k, const float* bias, float* out, const PolyGeom& g, int items,
                     int act, rcb_stream_t stream, bool half, bool out_half = false) {
  ConvF2Args f;
  f.g = g; f.items = items;
  f.tiles_x = ceil_div(g.w, 8); f.tiles_y = ceil_div(g.h, 16);
  f.n_tiles = f.tiles_x * f.tiles_y * items;
  const int KC = half ? 64 : 32, es = half ? 2 : 4;
  f.kblocks = g.ic / KC; f.act = act; f.out_half = out_half ? 1 : 0; f.bias = bias; f.out = out;
  f.w_bytes_kb = 16 * g.oc * 128;
  // accumulator order (ry, rx) = (0,0) (0,1) (1,1) (1,0); a phase's tap for the shift (dy, dx) is
  // (dy - base_y(ry), dx - base_x(rx)) with base(0) = -1, base(1) = 0
  const int order[4] = {0, 1, 3, 2};
  for (int i = 0; i < 4; ++i) f.phase_col[order[i]] = i * g.oc;
  // groups: {dy, dx, first slot, members}
  const int groups[F2_GROUPS][4] = {{0, 0, 0, 4}, {-1, 0, 0, 2}, {0, 1, 1, 2}, {1, 0, 2, 2}, {0, -1, 0, 1},
                                    {0, -1, 3, 1}, {-1, -1, 0, 1}, {-1, 1, 1, 1}, {1, 1, 2, 1}, {1, -1, 3, 1}};
  int nb = 0;
  for (int j = 0; j < F2_GROUPS; ++j) {
    const int dy = groups[j][0], dx = groups[j][1], slot = groups[j][2], mem = groups[j][3];
    f.grp_shift[j] = (1 + dy) * HALO_PITCH + (1 + dx);
    f.grp_col[j] = slot * g.oc;
    f.grp_n[j] = mem * g.oc;
    f.grp_row[j] = nb * g.oc;
    for (int i = 0; i < mem; ++i, ++nb) {
      const int ph = order[slot + i], ry = ph >> 1, rx = ph & 1;
      const int ty = dy - (ry == 0 ? -1 : 0), tx = dx - (rx == 0 ? -1 : 0);
      RCB_CHECK_ARG(nb < 16 && ty >= 0 && ty < 2 && tx >= 0 && tx < 2, "rcb_upconv_fwd_tc: bad shift table");
      f.mem_phase[nb] = ph; f.mem_tap[nb] = ty * 2 + tx;
    }
  }
  RCB_CHECK_ARG(nb == 16, "rcb_upconv_fwd_tc: bad shift table");
  f.a_off = f.kblocks * f.w_bytes_kb;
  f.epi_off = f.a_off + F2_STAGES * HALO_BUF;
  f.bar_off = f.epi_off + 4 * 2 * 4096;            // per epilogue warp: two buffers of 32 rows x 128 B
  const int smem_total = f.bar_off + 512 + 1024;
  ConvTile box;
  box.tx = HALO_PITCH; box.ty = HALO_LINES; box.tz = 1; box.ni = 1; box.ntx = box.nty = box.ntz = 1;
  CUtensorMap tmA, tmB;
  if (int rc = make_map_5d(&tmA, src, items, g.d, g.h, g.w, g.ic, box, KC, 1, 1, 1, CU_TENSOR_MAP_SWIZZLE_128B, es)) return rc;
  if (int rc = make_map_b(&tmB, w_eff_k, (int64_t)g.phases() * g.oc, (int64_t)g.taps() * g.ic, g.oc, KC,
                          CU_TENSOR_MAP_SWIZZLE_128B, es)) return rc;
  CUtensorMap tmO;                            // out as (item, y, line parity, x, 2 * oc): one 128-byte row per source pixel
  {
    EncodeTiledFn enc = tc_get_encode();
    const cuuint64_t oes = out_half ? 2 : 4;
    const cuuint64_t W2 = 2 * (cuuint64_t)g.oc, line = (cuuint64_t)g.w * W2 * oes;
    cuuint64_t dims[5] = {W2, (cuuint64_t)g.w, 2, (cuuint64_t)g.h, (cuuint64_t)items};
    cuuint64_t strides[4] = {W2 * oes, line, 2 * line, (cuuint64_t)g.h * 2 * line};
    cuuint32_t obox[5] = {(cuuint32_t)W2, 8, 1, 4, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&tmO, out_half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)out, dims,
                     strides, obox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     out_half ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(out) failed with CUresult %d", (int)r); return -1; }
  }
  if (int rc = half ? opt_in_smem(upconv_fwd_f2_kernel<true>, "rcb_upconv_fwd_tc") : opt_in_smem(upconv_fwd_f2_kernel<false>, "rcb_upconv_fwd_tc")) return rc;
  static int n_sm = 0;
  if (n_sm == 0) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); }
  const int grid = f.n_tiles < n_sm ? f.n_tiles : n_sm;
  if (half) upconv_fwd_f2_kernel<true><<<grid, TC_THREADS, smem_total, (cudaStream_t)stream>>>(tmA, tmB, tmO, f);
  else upconv_fwd_f2_kernel<false><<<grid, TC_THREADS, smem_total, (cudaStream_t)stream>>>(tmA, tmB, tmO, f);
  RCB_CHECK_LAUNCH("rcb_upconv_fwd_tc");
  return 0;
}

static bool f2w_eligible(const PolyGeom& g) {
  return g.d == 1 && g.Tz == 1 && g.Ty == 2 && g.Tx == 2 && g.fy == 2 && g.fx == 2 && g.py == 1 && g.px == 1 &&
         g.oc == 64 && g.ic == 64 && ((g.h == 8 && g.w == 8) || (g.h >= 16 && g.h % 4 == 0));
}
static int encode_map_t(CUtensorMap* map, CUtensorMapDataType dt, const void* base, int rank, const cuuint64_t* dims,
                        const cuuint64_t* strides, const cuuint32_t* box, const char* what) {
  EncodeTiledFn enc = tc_get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return -1; }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(map, dt, (cuuint32_t)rank, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)r); return -1; }
  return 0;
}
// fp16 in, fp16 out, fp16 K-major weights
static int launch_f2w(const void* src_h, const void* w_eff_k_h, const float* bias, void* out_h, const PolyGeom& g, int items,
                      int act, rcb_stream_t stream) {
  ConvF2WArgs f;
  f.g = g; f.items = items;
  f.ipt = (g.h == 8 && g.w == 8) ? 2 : 1;
  f.tiles_x = ceil_div(g.w, 8); f.tiles_y = ceil_div(g.h, 16);
  f.n_tiles = f.ipt == 2 ? ceil_div(items, 2) : f.tiles_x * f.tiles_y * items;
  f.act = act; f.bias = bias;
  f.halo_bytes = (f.ipt == 2 ? 10 * 2 * HALO_PITCH : HALO_LINES * HALO_PITCH) * 128;
  const int dy_rows = HALO_PITCH * f.ipt;
  const int order[4] = {0, 1, 3, 2};
  for (int i = 0; i < 4; ++i) f.phase_col[order[i]] = i * 64;
  const int groups[F2_GROUPS][4] = {{0, 0, 0, 4}, {-1, 0, 0, 2}, {0, 1, 1, 2}, {1, 0, 2, 2}, {0, -1, 0, 1},
                                    {0, -1, 3, 1}, {-1, -1, 0, 1}, {-1, 1, 1, 1}, {1, 1, 2, 1}, {1, -1, 3, 1}};
  int nb = 0;
  for (int j = 0; j < F2_GROUPS; ++j) {
    const int dy = groups[j][0], dx = groups[j][1], slot = groups[j][2], mem = groups[j][3];
    f.grp_shift[j] = (1 + dy) * dy_rows + (1 + dx);
    f.grp_col[j] = slot * 64;
    f.grp_n[j] = mem * 64;
    f.grp_row[j] = nb * 64;
    for (int i = 0; i < mem; ++i, ++nb) {
      const int ph = order[slot + i], ry = ph >> 1, rx = ph & 1;
      const int ty = dy - (ry == 0 ? -1 : 0), tx = dx - (rx == 0 ? -1 : 0);
      RCB_CHECK_ARG(nb < 16 && ty >= 0 && ty < 2 && tx >= 0 && tx < 2, "rcb_upconv_fwd_tc_hh: bad shift table");
      f.mem_phase[nb] = ph; f.mem_tap[nb] = ty * 2 + tx;
    }
  }
  RCB_CHECK_ARG(nb == 16, "rcb_upconv_fwd_tc_hh: bad shift table");
  f.a_off = 16 * F2W_W_BLOCK;
  f.epi_off = f.a_off + F2W_STAGES * F2W_STAGE_BYTES;
  f.bias_off = f.epi_off + 4 * 8192;
  f.bar_off = f.bias_off + 1024;
  const int smem_total = f.bar_off + 512 + 1024;
  CUtensorMap tmA, tmB, tmO;
  {
    const cuuint64_t px = 128, line = (cuuint64_t)g.w * px, img = (cuuint64_t)g.h * line;
    if (f.ipt == 2) {     // (channels, x, item, y): the two items of a tile alternate line by line in shared memory
      cuuint64_t dims[4] = {64, (cuuint64_t)g.w, (cuuint64_t)items, (cuuint64_t)g.h};
      cuuint64_t strides[3] = {px, img, line};
      cuuint32_t box[4] = {64, HALO_PITCH, 2, 10};
      if (int rc = encode_map_t(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, src_h, 4, dims, strides, box, "src")) return rc;
    } else {
      cuuint64_t dims[4] = {64, (cuuint64_t)g.w, (cuuint64_t)g.h, (cuuint64_t)items};
      cuuint64_t strides[3] = {px, line, img};
      cuuint32_t box[4] = {64, HALO_PITCH, HALO_LINES, 1};
      if (int rc = encode_map_t(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, src_h, 4, dims, strides, box, "src")) return rc;
    }
  }
  if (int rc = make_map_b(&tmB, w_eff_k_h, (int64_t)g.phases() * g.oc, (int64_t)g.taps() * g.ic, 64, 64,
                          CU_TENSOR_MAP_SWIZZLE_128B, 2)) return rc;
  {   // out as ((item, y), line parity, px, column parity, 64 channels): a phase of 8 pixels x n lines is one box
    const cuuint64_t px = 128, line = 2 * (cuuint64_t)g.w * px;
    cuuint64_t dims[5] = {64, 2, (cuuint64_t)g.w, 2, (cuuint64_t)g.h * (cuuint64_t)items};
    cuuint64_t strides[4] = {px, 2 * px, line, 2 * line};
    cuuint32_t box[5] = {32, 1, 8, 1, (cuuint32_t)(f.ipt == 2 ? 2 : 4)};     // a warp's 32 rows x 32 channels of one phase
    EncodeTiledFn enc = tc_get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return -1; }
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&tmO, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, (void*)out_h, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(out, fp16) failed with CUresult %d", (int)r); return -1; }
  }
  cudaError_t e = cudaFuncSetAttribute(upconv_fwd_f2w_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_total);
  if (e != cudaSuccess) { set_error("rcb_upconv_fwd_tc_hh: smem opt-in failed: %s", cudaGetErrorString(e)); return -1; }
  static int n_sm = 0;
  if (n_sm == 0) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); }
  const int grid = f.n_tiles < n_sm ? f.n_tiles : n_sm;
  upconv_fwd_f2w_kernel<<<grid, B2_THREADS, smem_total, (cudaStream_t)stream>>>(tmA, tmB, tmO, f);
  RCB_CHECK_LAUNCH("rcb_upconv_fwd_tc_hh");
  return 0;
}

static bool b2_eligible(const PolyGeom& g) {
  return g.d == 1 && g.Tz == 1 && g.Ty == 2 && g.Tx == 2 && g.h >= 16 && g.fy == 2 && g.fx == 2 && g.py == 1 && g.px == 1 &&
         g.oc == 16 && g.ic == 64;
}
static int encode_map(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                      const cuuint32_t* box, const char* what) {
  EncodeTiledFn enc = tc_get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return -1; }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, (void*)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)r); return -1; }
  return 0;
}
static int launch_b2(const void* d_out, const void* w_bwd_k, const void* src_act, int act_kind, void* d_src,
                     const PolyGeom& g, int items, rcb_stream_t stream, int out_half = 0, float out_scale = 1.f, int in_half = 0) {
  ConvB2Args f;
  f.g = g; f.items = items;
  f.out_half = out_half; f.out_scale = out_scale;
  f.tiles_x = ceil_div(g.w, 8); f.tiles_y = ceil_div(g.h, 16);
  f.n_tiles = f.tiles_x * f.tiles_y * items;
  f.act_kind = src_act ? act_kind : 0; f.src_act = src_act;
  f.a_off = (in_half ? 4 : 8) * 64 * 128;
  f.m_off = f.a_off + B2_STAGES * (in_half ? HALO_BUF_H : HALO_BUF);
  f.epi_off = f.m_off + (in_half ? B2_MSTAGES * 16384 : 0);
  f.bar_off = f.epi_off + 8 * 2 * 4096;               // staging buffers of the eight epilogue warps
  const int smem_total = f.bar_off + 512 + 1024;
  CUtensorMap tmA, tmB, tmO, tmM;
  memset(&tmM, 0, sizeof(tmM));
  if (in_half && f.act_kind == 2) {   // the producing stage's fp16 activations, one 128-row tile per box
    const cuuint64_t px = 128, line = (cuuint64_t)g.w * px;
    cuuint64_t dims[4] = {64, (cuuint64_t)g.w, (cuuint64_t)g.h, (cuuint64_t)items};
    cuuint64_t strides[3] = {px, line, (cuuint64_t)g.h * line};
    cuuint32_t box[4] = {64, 8, 16, 1};
    if (int rc = encode_map_t(&tmM, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, src_act, 4, dims, strides, box, "src_act")) return rc;
  }
  {   // d_out as (item, y, line parity, x, (column parity, oc)): one 128-byte (fp16: 64-byte) row per source pixel and line parity
    const cuuint64_t row = in_half ? 64 : 128, line = (cuuint64_t)g.w * row;
    cuuint64_t dims[5] = {32, (cuuint64_t)g.w, 2, (cuuint64_t)g.h, (cuuint64_t)items};
    cuuint64_t strides[4] = {row, line, 2 * line, (cuuint64_t)g.h * 2 * line};
    cuuint32_t box[5] = {32, HALO_PITCH, 1, HALO_LINES, 1};
    if (in_half) {
      EncodeTiledFn enc = tc_get_encode();
      if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return -1; }
      cuuint32_t estr[5] = {1, 1, 1, 1, 1};
      CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, (void*)d_out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(d_out, fp16) failed with CUresult %d", (int)r); return -1; }
    } else if (int rc = encode_map(&tmA, d_out, 5, dims, strides, box, "d_out")) return rc;
  }
  if (in_half) {
    if (int rc = make_map_b(&tmB, w_bwd_k, 64, 256, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B, 2)) return rc;
  } else if (int rc = make_map_b(&tmB, w_bwd_k, 64, 256, 64, 32, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  {
    const cuuint64_t px = out_half ? 128 : 256;
    cuuint64_t dims[4] = {64, (cuuint64_t)g.w, (cuuint64_t)g.h, (cuuint64_t)items};
    cuuint64_t strides[3] = {px, (cuuint64_t)g.w * px, (cuuint64_t)g.h * g.w * px};
    cuuint32_t box[4] = {32, 8, 4, 1};                  // a warp's 32 rows x 32 channels
    if (out_half) {
      EncodeTiledFn enc = tc_get_encode();
      if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return -1; }
      cuuint32_t estr[4] = {1, 1, 1, 1};
      CUresult r = enc(&tmO, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, (void*)d_src, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(d_src, fp16) failed with CUresult %d", (int)r); return -1; }
    } else if (int rc = encode_map(&tmO, d_src, 4, dims, strides, box, "d_src")) return rc;
  }
  static int n_sm = 0;
  if (n_sm == 0) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); }
  const int grid = f.n_tiles < n_sm ? f.n_tiles : n_sm;
  if (in_half) {
    if (int rc = opt_in_smem(upconv_bwd_f2_kernel<true>, "rcb_upconv_bwd_f2", smem_total)) return rc;
    upconv_bwd_f2_kernel<true><<<grid, B2_THREADS, smem_total, (cudaStream_t)stream>>>(tmA, tmB, tmO, tmM, f);
  } else {
    if (int rc = opt_in_smem(upconv_bwd_f2_kernel<false>, "rcb_upconv_bwd_f2", smem_total)) return rc;
    upconv_bwd_f2_kernel<false><<<grid, B2_THREADS, smem_total, (cudaStream_t)stream>>>(tmA, tmB, tmO, tmM, f);
  }
  RCB_CHECK_LAUNCH("rcb_upconv_bwd_f2");
  return 0;
}

static bool b2w_eligible(const PolyGeom& g) {
  return g.d == 1 && g.Tz == 1 && g.Ty == 2 && g.Tx == 2 && g.fy == 2 && g.fx == 2 && g.py == 1 && g.px == 1 &&
         g.oc == 64 && g.ic == 64 && ((g.h == 8 && g.w == 8) || (g.h >= 16 && g.h % 4 == 0));
}
// fp16 d_out (scaled), fp16 K-major weights [ic][(a, b, oc)], fp16 mask activations, fp32 d_src
static int launch_b2w(const void* d_out_h, const void* w_bwd_k_h, const void* src_act_h, void* d_src, const PolyGeom& g,
                      int items, float out_scale_inv, rcb_stream_t stream, int out_half = 0) {
  ConvB2WArgs f;
  f.g = g; f.items = items; f.out_half = out_half;
  f.ipt = (g.h == 8 && g.w == 8) ? 2 : 1;
  f.tiles_x = ceil_div(g.w, 8); f.tiles_y = ceil_div(g.h, 16);
  f.n_tiles = f.ipt == 2 ? ceil_div(items, 2) : f.tiles_x * f.tiles_y * items;
  f.halo_bytes = (f.ipt == 2 ? 10 * 2 * HALO_PITCH : HALO_LINES * HALO_PITCH) * 128;
  f.out_scale_inv = out_scale_inv; f.src_act = src_act_h; f.d_src = d_src;
  f.a_off = 16 * F2W_W_BLOCK;
  f.epi_off = f.a_off + B2W_STAGES * F2W_STAGE_BYTES;
  f.bar_off = f.epi_off + 8 * 128;                    // mask scratch of the eight epilogue warps
  const int smem_total = f.bar_off + 512 + 1024;
  EncodeTiledFn enc = tc_get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return -1; }
  CUtensorMap tmA, tmB, tmO;
  {   // d_out (item, 2h, 2w, 64 oc fp16), read with element strides 2 along x and y: one (line, column) parity per box
    const cuuint64_t px = 128, line = 2 * (cuuint64_t)g.w * px, img = 2 * (cuuint64_t)g.h * line;
    const cuuint32_t bx = 2 * HALO_PITCH - 1;
    CUresult r;
    if (f.ipt == 2) {
      cuuint64_t dims[4] = {64, 2 * (cuuint64_t)g.w, (cuuint64_t)items, 2 * (cuuint64_t)g.h};
      cuuint64_t strides[3] = {px, img, line};
      cuuint32_t box[4] = {64, bx, 2, 2 * 10 - 1};
      cuuint32_t estr[4] = {1, 2, 1, 2};
      r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, (void*)d_out_h, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
      cuuint64_t dims[4] = {64, 2 * (cuuint64_t)g.w, 2 * (cuuint64_t)g.h, (cuuint64_t)items};
      cuuint64_t strides[3] = {px, line, img};
      cuuint32_t box[4] = {64, bx, 2 * HALO_LINES - 1, 1};
      cuuint32_t estr[4] = {1, 2, 2, 1};
      r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, (void*)d_out_h, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(d_out, strided) failed with CUresult %d", (int)r); return -1; }
  }
  if (int rc = make_map_b(&tmB, w_bwd_k_h, 64, 1024, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B, 2)) return rc;
  memset(&tmO, 0, sizeof(tmO));            // the result leaves through plain stores (kernel parameter kept for ABI symmetry)
  cudaError_t e = cudaFuncSetAttribute(upconv_bwd_f2w_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_total);
  if (e != cudaSuccess) { set_error("rcb_upconv_bwd_f2w: smem opt-in failed: %s", cudaGetErrorString(e)); return -1; }
  static int n_sm = 0;
  if (n_sm == 0) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); }
  const int grid = f.n_tiles < n_sm ? f.n_tiles : n_sm;
  upconv_bwd_f2w_kernel<<<grid, B2_THREADS, smem_total, (cudaStream_t)stream>>>(tmA, tmB, tmO, f);
  RCB_CHECK_LAUNCH("rcb_upconv_bwd_f2w");
  return 0;
}

}  // namespace rcb

using namespace rcb;

// w_eff_k: forward weights in K-major form [phase][oc][tap*ic] (rcb_fold_poly_k).
static int upconv_fwd_tc_impl(const float* src, const float* w_eff_k, const float* bias, float* out,
                              const rcb_upconv_geom* geo, int items, int act, rcb_stream_t stream, int out_half,
                              int in_half = 0) {
  PolyGeom g;
  if (int rc = make_geom_tc(geo, &g)) return rc;
  RCB_CHECK_ARG(src && w_eff_k && bias && out, "rcb_upconv_fwd_tc: null pointer");
  RCB_CHECK_ARG(g.ic % (in_half ? 64 : 32) == 0 && g.oc % 16 == 0 && g.oc <= 128, "rcb_upconv_fwd_tc: unsupported channels %d -> %d", g.ic, g.oc);
  if (items <= 0) return 0;
  if (in_half && f2_eligible(g, true) && !out_half) return launch_f2(src, w_eff_k, bias, out, g, items, act, stream, true);
  if (in_half && out_half && f2w_eligible(g) && !getenv("RCB_NO_F2W")) return launch_f2w(src, w_eff_k, bias, out, g, items, act, stream);
  if (in_half && out_half && f2_eligible(g, true)) return launch_f2(src, w_eff_k, bias, out, g, items, act, stream, true, true);
  if (f2_eligible(g, false) && !out_half && !getenv("RCB_NO_F2")) return launch_f2(src, w_eff_k, bias, out, g, items, act, stream, false);
  if (!in_half && g.d == 1 && g.Tz == 1 && g.Ty == 2 && g.Tx == 2 && g.h >= 16) {
    // 2-D grid with full 8 x 16 tiles: halo-tile kernel
    ConvHaloArgs h;
    h.g = g; h.items = items;
    h.tiles_x = ceil_div(g.w, 8); h.tiles_y = ceil_div(g.h, 16);
    int ppc = 256 / g.oc;                       // <= 256 TMEM columns: two CTAs per SM
    if (ppc < 1) ppc = 1;
    if (ppc > g.phases()) ppc = g.phases();
    h.phases_per_cta = ppc;
    h.kblocks = g.ic / 32;
    h.act = act; h.out_half = out_half; h.bias = bias; h.out = out;
    h.b_bytes = (g.oc * 128 + 1023) / 1024 * 1024;
    h.nbs = 24576 / (4 * h.b_bytes);            // about 24 KB of weight ring, at least double-buffered
    if (h.nbs > HALO_NBS_MAX) h.nbs = HALO_NBS_MAX;
    if (h.nbs < 2) h.nbs = 2;
    h.a_off = h.nbs * 4 * h.b_bytes;
    h.bar_off = h.a_off + 2 * HALO_BUF;
    const int smem_total = h.bar_off + 512 + 1024;
    ConvTile box;                               // 16 x 18 pixels of one item
    box.tx = HALO_PITCH; box.ty = HALO_LINES; box.tz = 1; box.ni = 1; box.ntx = box.nty = box.ntz = 1;
    CUtensorMap tmA, tmB;
    if (int rc = make_map_5d(&tmA, src, items, g.d, g.h, g.w, g.ic, box, 32, 1, 1, 1, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    if (int rc = make_map_b(&tmB, w_eff_k, (int64_t)g.phases() * g.oc, (int64_t)g.taps() * g.ic, g.oc, 32,
                            CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    if (int rc = opt_in_smem(upconv_fwd_halo_kernel, "rcb_upconv_fwd_tc")) return rc;
    dim3 grid(h.tiles_x * h.tiles_y * items, ceil_div(g.phases(), ppc));
    upconv_fwd_halo_kernel<<<grid, TC_THREADS, smem_total, (cudaStream_t)stream>>>(tmA, tmB, h);
    RCB_CHECK_LAUNCH("rcb_upconv_fwd_tc");
    return 0;
  }
  ConvTcArgs a;
  a.g = g;
  a.t = choose_tile(g, items, 128, 128, 128);
  a.items = items;
  int ppc = 512 / g.oc;
  if (ppc > CT_MAX_PHASES) ppc = CT_MAX_PHASES;
  if (ppc > g.phases()) ppc = g.phases();
  a.phases_per_cta = ppc;
  const int KC = in_half ? 64 : 32, es = in_half ? 2 : 4;
  a.kblocks = g.ic / KC;
  a.bk = 32;
  a.act = act; a.out_half = out_half; a.in_half = in_half; a.act_half = 0; a.bias = bias; a.src_act = nullptr; a.out = out;
  int smem_total;
  smem_layout(&a, 128, g.oc, &smem_total, 3, 4 * 32 * (g.oc + 4) * 4);
  CUtensorMap tmA, tmB;
  if (int rc = make_map_5d(&tmA, src, items, g.d, g.h, g.w, g.ic, a.t, KC, 1, 1, 1, CU_TENSOR_MAP_SWIZZLE_128B, es)) return rc;
  if (int rc = make_map_b(&tmB, w_eff_k, (int64_t)g.phases() * g.oc, (int64_t)g.taps() * g.ic, g.oc, KC,
                          CU_TENSOR_MAP_SWIZZLE_128B, es)) return rc;
  if (int rc = opt_in_smem(upconv_fwd_tc_kernel, "rcb_upconv_fwd_tc")) return rc;
  dim3 grid(a.t.ntx * a.t.nty * a.t.ntz * ceil_div(items, a.t.ni), ceil_div(g.phases(), ppc));
  upconv_fwd_tc_kernel<<<grid, TC_THREADS, smem_total, (cudaStream_t)stream>>>(tmA, tmB, a);
  RCB_CHECK_LAUNCH("rcb_upconv_fwd_tc");
  return 0;
}

extern "C" int rcb_upconv_fwd_tc(const float* src, const float* w_eff_k, const float* bias, float* out,
                                 const rcb_upconv_geom* geo, int items, int act, rcb_stream_t stream) {
  return upconv_fwd_tc_impl(src, w_eff_k, bias, out, geo, items, act, stream, 0);
}
// same, writing the activations as fp16 for a following rcb_upconv_fwd_tc_h stage
extern "C" int rcb_upconv_fwd_tc_oh(const float* src, const float* w_eff_k, const float* bias, void* out_h,
                                    const rcb_upconv_geom* geo, int items, int act, rcb_stream_t stream) {
  return upconv_fwd_tc_impl(src, w_eff_k, bias, reinterpret_cast<float*>(out_h), geo, items, act, stream, 1);
}

// fp16 source activations and fp16 K-major weights (rcb_to_half of the rcb_fold_poly_k output), ic % 64 == 0.  The
// x2 / 16-channel stage runs the resident-weight kernel, everything else the general one with kind::f16 MMAs.
extern "C" int rcb_upconv_fwd_tc_h(const void* src_h, const void* w_eff_k_h, const float* bias, float* out,
                                   const rcb_upconv_geom* geo, int items, int act, rcb_stream_t stream) {
  return upconv_fwd_tc_impl(reinterpret_cast<const float*>(src_h), reinterpret_cast<const float*>(w_eff_k_h), bias, out, geo,
                            items, act, stream, 0, 1);
}
// fp16 in, fp16 out
extern "C" int rcb_upconv_fwd_tc_hh(const void* src_h, const void* w_eff_k_h, const float* bias, void* out_h,
                                    const rcb_upconv_geom* geo, int items, int act, rcb_stream_t stream) {
  return upconv_fwd_tc_impl(reinterpret_cast<const float*>(src_h), reinterpret_cast<const float*>(w_eff_k_h), bias,
                            reinterpret_cast<float*>(out_h), geo, items, act, stream, 1, 1);
}

// Data gradient of the x2 / 3-tap / 64 -> 16 channel stage with resident weights.  w_bwd_k: rcb_fold_poly_bwd_f2 of w_eff.
// act_kind: 0 no mask, 1 src_act holds fp32 activations, 2 fp16 activations (rcb_upconv_fwd_tc_oh).
extern "C" int rcb_fold_poly_bwd_f2(const float* w_eff, const rcb_upconv_geom* geo, float* w_bwd_k, rcb_stream_t stream) {
  PolyGeom g;
  if (int rc = make_geom_tc(geo, &g)) return rc;
  RCB_CHECK_ARG(w_eff && w_bwd_k, "rcb_fold_poly_bwd_f2: null pointer");
  RCB_CHECK_ARG(b2_eligible(g), "rcb_fold_poly_bwd_f2: only 2-D x2 stages with 64 -> 16 channels and h >= 16");
  const int n = g.ic * 16 * g.oc;
  fold_bwd_f2_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(w_eff, w_bwd_k, g.ic, g.oc);
  RCB_CHECK_LAUNCH("rcb_fold_poly_bwd_f2");
  return 0;
}
extern "C" int rcb_upconv_bwd_f2(const float* d_out, const float* w_bwd_k, const void* src_act, int act_kind, float* d_src,
                                 const rcb_upconv_geom* geo, int items, rcb_stream_t stream) {
  PolyGeom g;
  if (int rc = make_geom_tc(geo, &g)) return rc;
  RCB_CHECK_ARG(d_out && w_bwd_k && d_src, "rcb_upconv_bwd_f2: null pointer");
  RCB_CHECK_ARG(act_kind >= 0 && act_kind <= 2, "rcb_upconv_bwd_f2: act_kind must be 0, 1 or 2");
  RCB_CHECK_ARG(b2_eligible(g), "rcb_upconv_bwd_f2: only 2-D x2 stages with 64 -> 16 channels and h >= 16");
  if (items <= 0) return 0;
  return launch_b2(d_out, w_bwd_k, src_act, act_kind, d_src, g, items, stream);
}

// same, leaving d_src as fp16 multiplied by out_scale (a power of two that brings the gradient to O(1))
extern "C" int rcb_upconv_bwd_f2_oh(const float* d_out, const float* w_bwd_k, const void* src_act, int act_kind, void* d_src_h,
                                    float out_scale, const rcb_upconv_geom* geo, int items, rcb_stream_t stream) {
  PolyGeom g;
  if (int rc = make_geom_tc(geo, &g)) return rc;
  RCB_CHECK_ARG(d_out && w_bwd_k && d_src_h, "rcb_upconv_bwd_f2_oh: null pointer");
  RCB_CHECK_ARG(act_kind >= 0 && act_kind <= 2, "rcb_upconv_bwd_f2_oh: act_kind must be 0, 1 or 2");
  RCB_CHECK_ARG(out_scale > 0.f, "rcb_upconv_bwd_f2_oh: out_scale must be positive");
  RCB_CHECK_ARG(b2_eligible(g), "rcb_upconv_bwd_f2_oh: only 2-D x2 stages with 64 -> 16 channels and h >= 16");
  if (items <= 0) return 0;
  return launch_b2(d_out, w_bwd_k, src_act, act_kind, d_src_h, g, items, stream, 1, out_scale);
}
// fp16 in (d_out_h in any fixed unit, w_bwd_k_h = rcb_to_half of the folded weights) and fp16 out, same unit x out_scale
extern "C" int rcb_upconv_bwd_f2_hh(const void* d_out_h, const void* w_bwd_k_h, const void* src_act, int act_kind, void* d_src_h,
                                    float out_scale, const rcb_upconv_geom* geo, int items, rcb_stream_t stream) {
  PolyGeom g;
  if (int rc = make_geom_tc(geo, &g)) return rc;
  RCB_CHECK_ARG(d_out_h && w_bwd_k_h && d_src_h, "rcb_upconv_bwd_f2_hh: null pointer");
  RCB_CHECK_ARG(act_kind >= 0 && act_kind <= 2, "rcb_upconv_bwd_f2_hh: act_kind must be 0, 1 or 2");
  RCB_CHECK_ARG(out_scale > 0.f, "rcb_upconv_bwd_f2_hh: out_scale must be positive");
  RCB_CHECK_ARG(b2_eligible(g), "rcb_upconv_bwd_f2_hh: only 2-D x2 stages with 64 -> 16 channels and h >= 16");
  if (items <= 0) return 0;
  return launch_b2(d_out_h, w_bwd_k_h, src_act, act_kind, d_src_h, g, items, stream, 1, out_scale, 1);
}
extern "C" int rcb_upconv_bwd_f2w_eligible(const rcb_upconv_geom* geo) {
  PolyGeom g;
  if (make_geom_tc(geo, &g)) return 0;
  return b2w_eligible(g) ? 1 : 0;
}
extern "C" int rcb_fold_poly_bwd_f2w(const float* w_eff, const rcb_upconv_geom* geo, float* w_bwd_k, rcb_stream_t stream) {
  PolyGeom g;
  if (int rc = make_geom_tc(geo, &g)) return rc;
  RCB_CHECK_ARG(w_eff && w_bwd_k, "rcb_fold_poly_bwd_f2w: null pointer");
  RCB_CHECK_ARG(b2w_eligible(g), "rcb_fold_poly_bwd_f2w: only 2-D x2 stages with 64 -> 64 channels (8 x 8 grids or >= 16 lines)");
  const int n = g.ic * 16 * g.oc;
  fold_bwd_f2_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(w_eff, w_bwd_k, g.ic, g.oc);
  RCB_CHECK_LAUNCH("rcb_fold_poly_bwd_f2w");
  return 0;
}
extern "C" int rcb_upconv_bwd_f2w(const void* d_out_h, const void* w_bwd_k_h, const void* src_act_h, float* d_src,
                                  float out_scale_inv, const rcb_upconv_geom* geo, int items, rcb_stream_t stream) {
  PolyGeom g;
  if (int rc = make_geom_tc(geo, &g)) return rc;
  RCB_CHECK_ARG(d_out_h && w_bwd_k_h && d_src, "rcb_upconv_bwd_f2w: null pointer");
  RCB_CHECK_ARG(out_scale_inv > 0.f, "rcb_upconv_bwd_f2w: out_scale_inv must be positive");
  RCB_CHECK_ARG(b2w_eligible(g), "rcb_upconv_bwd_f2w: only 2-D x2 stages with 64 -> 64 channels (8 x 8 grids or >= 16 lines)");
  if (items <= 0) return 0;
  return launch_b2w(d_out_h, w_bwd_k_h, src_act_h, d_src, g, items, out_scale_inv, stream);
}
// same with d_src left as fp16 (saturating), multiplied by out_scale: the operand of an fp16 GEMM for the stage below
extern "C" int rcb_upconv_bwd_f2w_oh(const void* d_out_h, const void* w_bwd_k_h, const void* src_act_h, void* d_src_h,
                                     float out_scale, const rcb_upconv_geom* geo, int items, rcb_stream_t stream) {
  PolyGeom g;
  if (int rc = make_geom_tc(geo, &g)) return rc;
  RCB_CHECK_ARG(d_out_h && w_bwd_k_h && d_src_h, "rcb_upconv_bwd_f2w_oh: null pointer");
  RCB_CHECK_ARG(out_scale > 0.f, "rcb_upconv_bwd_f2w_oh: out_scale must be positive");
  RCB_CHECK_ARG(b2w_eligible(g), "rcb_upconv_bwd_f2w_oh: only 2-D x2 stages with 64 -> 64 channels (8 x 8 grids or >= 16 lines)");
  if (items <= 0) return 0;
  return launch_b2w(d_out_h, w_bwd_k_h, src_act_h, d_src_h, g, items, out_scale, stream, 1);
}

// w_eff: [phase][tap][ic][oc] as produced by rcb_fold_poly (already K-major for this GEMM).
static int upconv_bwd_tc_impl(const float* d_out, const float* w_eff, const float* src_act, float* d_src,
                              const rcb_upconv_geom* geo, int items, rcb_stream_t stream, int act_half) {
  PolyGeom g;
  if (int rc = make_geom_tc(geo, &g)) return rc;
  RCB_CHECK_ARG(d_out && w_eff && d_src, "rcb_upconv_bwd_tc: null pointer");
  RCB_CHECK_ARG(g.ic % 16 == 0 && g.ic <= 128 && (g.oc == 16 || g.oc % 32 == 0),
                "rcb_upconv_bwd_tc: unsupported channels %d -> %d", g.ic, g.oc);
  if (items <= 0) return 0;
  ConvTcArgs a;
  a.g = g;
  a.t = choose_tile(g, items, 256 / g.fx, 256 / g.fy, 256 / g.fz);
  a.items = items;
  a.phases_per_cta = 1;
  const int bk = (g.oc == 16) ? 16 : 32;
  a.bk = bk;
  a.kblocks = g.oc / bk;
  a.act = 0; a.out_half = 0; a.in_half = 0; a.act_half = act_half; a.bias = nullptr; a.src_act = src_act; a.out = d_src;
  int smem_total;
  smem_layout(&a, bk * 4, g.ic, &smem_total);
  RCB_CHECK_ARG(4 * 32 * (g.ic + 4) * 4 <= a.bar_off, "rcb_upconv_bwd_tc: epilogue staging does not fit the pipeline stages");
  const CUtensorMapSwizzle swz = bk == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUtensorMap tmA, tmB;
  if (int rc = make_map_5d(&tmA, d_out, items, g.d * g.fz, g.h * g.fy, g.w * g.fx, g.oc, a.t, bk, g.fz, g.fy, g.fx, swz)) return rc;
  if (int rc = make_map_b(&tmB, w_eff, (int64_t)g.phases() * g.taps() * g.ic, g.oc, g.ic, bk, swz)) return rc;
  dim3 grid(a.t.ntx * a.t.nty * a.t.ntz * ceil_div(items, a.t.ni));
  if (bk == 32) {
    if (int rc = opt_in_smem(upconv_bwd_tc_kernel<32>, "rcb_upconv_bwd_tc")) return rc;
    upconv_bwd_tc_kernel<32><<<grid, TC_THREADS, smem_total, (cudaStream_t)stream>>>(tmA, tmB, a);
  } else {
    if (int rc = opt_in_smem(upconv_bwd_tc_kernel<16>, "rcb_upconv_bwd_tc")) return rc;
    upconv_bwd_tc_kernel<16><<<grid, TC_THREADS, smem_total, (cudaStream_t)stream>>>(tmA, tmB, a);
  }
  RCB_CHECK_LAUNCH("rcb_upconv_bwd_tc");
  return 0;
}
extern "C" int rcb_upconv_bwd_tc(const float* d_out, const float* w_eff, const float* src_act, float* d_src,
                                 const rcb_upconv_geom* geo, int items, rcb_stream_t stream) {
  return upconv_bwd_tc_impl(d_out, w_eff, src_act, d_src, geo, items, stream, 0);
}
// same, with the producing stage's activations stored as fp16 (rcb_upconv_fwd_tc_oh): only their signs are read
extern "C" int rcb_upconv_bwd_tc_ah(const float* d_out, const float* w_eff, const void* src_act_h, float* d_src,
                                    const rcb_upconv_geom* geo, int items, rcb_stream_t stream) {
  RCB_CHECK_ARG(src_act_h != nullptr, "rcb_upconv_bwd_tc_ah: null activations");
  return upconv_bwd_tc_impl(d_out, w_eff, reinterpret_cast<const float*>(src_act_h), d_src, geo, items, stream, 1);
}

// ===================================================================== weight gradient ====
// Polyphase conv weight gradient of the upsampler (prior training) on tcgen05:
//   d_w_eff[phase][tap][ic][oc] = sum over (item, source pixel s) of src[s + shift(phase, tap)][ic] * d_out[s*f + r][oc]
// One TMEM accumulator (ic x oc) per (phase, tap) segment; K = pixels.  Both operands are read from
// channel-major copies (srcT [ic][items*h*w], doutT [oc][items*Ho*Wo], made by rcb_transpose), so a K block of
// 32 consecutive source pixels (of the flattened (y, x) grid of an item) is a 128-byte K-major row per channel: A tiles
// are 3-D TMA boxes, shifted by whole lines through the pixel coordinate (outside the item = zero fill) and along x by
// picking one of three pre-shifted copies (a TMA box start must be 16-byte aligned); doutT is regrouped by phase (rcb_transpose_phases), so B tiles are dense boxes of one phase plane.  The distinct shifts of a CTA's segments are
// loaded once per K block (9 tiles for 16 (phase, tap) pairs at factor 2).  Split-K over CTAs, partial sums
// are added atomically (like the SIMT kernel they replace).
namespace rcb {

constexpr int WG_MAX_SEG = 16, WG_MAX_SHIFT = 9, WG_MAX_GROUPS = 16;

// one group of phases handled by a CTA (blockIdx.y): its first phase and its set of distinct source shifts
struct WgGroup {
  int phase0, nshift;
  int shift_dy[WG_MAX_SHIFT], shift_dx[WG_MAX_SHIFT];
  int seg_shift[WG_MAX_SEG];          // per CTA-local segment: index of its shift
};

struct WgTcArgs {
  int h, w, fy, fx, ic, oc, items;
  int kb_per_item;                    // K block = 32 consecutive source pixels of the flattened (y, x) grid of one item
  int kb_total, kb_per_cta;
  int phases_per_cta;                 // segments per CTA = 4 * phases_per_cta
  int a_bytes, b_bytes, stage_bytes, bar_off;     // stage sized for the group with the most shifts
  float* d_w_eff;
  WgGroup groups[WG_MAX_GROUPS];      // all phase groups run in ONE launch (grid.y): they share the SMs instead of
                                      // queueing behind one another on the stream
};

__global__ void __launch_bounds__(TC_THREADS)
upconv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, WgTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + a.bar_off);          // [2]
  uint64_t* empty = full + 2;                               // [2]
  uint64_t* acc_full = empty + 2;
  uint32_t* tmem_slot = (uint32_t*)(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nseg = 4 * a.phases_per_cta;
  const WgGroup& gr = a.groups[blockIdx.y];
  const int ph0 = gr.phase0;
  const int kb0 = blockIdx.x * a.kb_per_cta;
  const int nkb = min(a.kb_per_cta, a.kb_total - kb0);
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < nseg * a.oc) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (nkb <= 0) {                       // uniform per CTA
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
    return;
  }

  if (warp == 0) {
    // ===== TMA producer: K block -> (item, first pixel) by counters
    int item = kb0 / a.kb_per_item, kbi = kb0 - item * a.kb_per_item;
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb & 1;
      mbar_wait(&empty[s], ((kb >> 1) & 1) ^ 1);
      if (elect_one()) {
        uint8_t* st = smem + s * a.stage_bytes;
        mbar_expect_tx(&full[s], (uint32_t)(gr.nshift * a.a_bytes + a.phases_per_cta * a.b_bytes));
        const int p0 = kbi * 32;
        // a line shift is +-w pixels of the flattened grid; pixels before the first / after the last line of the item
        // are outside the tensor dimension and come back as zeros
        for (int i = 0; i < gr.nshift; ++i)
          tma_load_3d(&tmA, &full[s], st + i * a.a_bytes, p0 + gr.shift_dy[i] * a.w, item, (gr.shift_dx[i] + 1) * a.ic);
        for (int p = 0; p < a.phases_per_cta; ++p)
          tma_load_3d(&tmB, &full[s], st + gr.nshift * a.a_bytes + p * a.b_bytes, p0, item * (a.fy * a.fx) + ph0 + p, 0);
      }
      __syncwarp();
      if (++kbi == a.kb_per_item) { kbi = 0; ++item; }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: per K block, 4 k-steps for every (phase, tap) segment
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a.oc >> 3) << 17) | ((uint32_t)(a.ic >> 4) << 24);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb & 1;
      mbar_wait(&full[s], (kb >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t st = smem_u32(smem + s * a.stage_bytes);
        for (int sg = 0; sg < nseg; ++sg) {
          const uint64_t da = smem_desc_sw128(st + gr.seg_shift[sg] * a.a_bytes);
          const uint64_t db = smem_desc_sw128(st + gr.nshift * a.a_bytes + (sg >> 2) * a.b_bytes);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_tf32(tmem_base + (uint32_t)(sg * a.oc), da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
        }
        umma_commit(&empty[s]);
        if (kb == nkb - 1) umma_commit(acc_full);
      }
      __syncwarp();
    }
  } else {
    // ===== epilogue: accumulator row = input channel.  M = 128: TMEM lane = row; M = 64: row i in lane (i % 16) + 32 (i / 16)
    const int q = warp & 3;
    const int row = a.ic == 128 ? q * 32 + lane : q * 16 + lane;
    const bool has_row = a.ic == 128 || lane < 16;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    for (int sg = 0; sg < nseg; ++sg) {
      const int seg_global = (ph0 + (sg >> 2)) * 4 + (sg & 3);
      float* dst = a.d_w_eff + ((int64_t)seg_global * a.ic + row) * a.oc;
      for (int c0 = 0; c0 < a.oc; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(sg * a.oc + c0), v);
        if (has_row) {
#pragma unroll
          for (int j = 0; j < 16; ++j) atomicAdd(dst + c0 + j, __uint_as_float(v[j]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

// 3-D map over a channel-major tensor [channels][planes][pixels] (fp32), box = 32 consecutive pixels of one plane x box_ch channels
static int make_map_cm(CUtensorMap* map, const float* base, int channels, int planes, int pixels, int box_ch) {
  EncodeTiledFn enc = tc_get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return -1; }
  cuuint64_t dims[3] = {(cuuint64_t)pixels, (cuuint64_t)planes, (cuuint64_t)channels};
  cuuint64_t strides[2] = {(cuuint64_t)pixels * 4, (cuuint64_t)planes * pixels * 4};
  cuuint32_t box[3] = {32, 1, (cuuint32_t)box_ch};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(channel-major) failed with CUresult %d", (int)r); return -1; }
  return 0;
}

}  // namespace rcb

// srcT: [3 x-shifts][ic][items*h*w] (rcb_transpose_xshift), doutT: [oc][items][fy*fx phases][h][w] (rcb_transpose_phases).
// d_w_eff: [phases][taps][ic][oc], overwritten.
extern "C" int rcb_upconv_wgrad_tc(const float* srcT, const float* doutT, float* d_w_eff, const rcb_upconv_geom* geo,
                                   int items, rcb_stream_t stream) {
  PolyGeom g;
  if (int rc = make_geom_tc(geo, &g)) return rc;
  RCB_CHECK_ARG(srcT && doutT && d_w_eff, "rcb_upconv_wgrad_tc: null pointer");
  RCB_CHECK_ARG(g.d == 1 && g.Tz == 1 && g.Ty == 2 && g.Tx == 2, "rcb_upconv_wgrad_tc: 2-D grids with k > 1 only");
  RCB_CHECK_ARG((g.ic == 64 || g.ic == 128) && g.oc % 16 == 0 && g.oc <= 128, "rcb_upconv_wgrad_tc: unsupported channels %d -> %d", g.ic, g.oc);
  WgTcArgs a;
  a.h = g.h; a.w = g.w; a.fy = g.fy; a.fx = g.fx; a.ic = g.ic; a.oc = g.oc; a.items = items;
  RCB_CHECK_ARG((g.h * g.w) % 32 == 0 && g.w % 4 == 0, "rcb_upconv_wgrad_tc: %d x %d grid does not tile into 32-pixel K blocks", g.h, g.w);
  a.kb_per_item = g.h * g.w / 32;
  cudaStream_t st = (cudaStream_t)stream;
  const int nphase = g.fy * g.fx;
  cudaError_t e = cudaMemsetAsync(d_w_eff, 0, sizeof(float) * (size_t)nphase * 4 * g.ic * g.oc, st);
  if (e != cudaSuccess) { set_error("rcb_upconv_wgrad_tc: memset failed: %s", cudaGetErrorString(e)); return -1; }
  if (items <= 0) return 0;
  a.kb_total = items * a.kb_per_item;
  int ppc = 256 / (4 * g.oc);                    // <= 256 TMEM columns per CTA
  if (ppc < 1) ppc = 1;
  if (ppc > nphase) ppc = nphase;
  while (nphase % ppc) --ppc;
  a.phases_per_cta = ppc;
  a.a_bytes = g.ic * 128;
  a.b_bytes = (g.oc * 128 + 1023) / 1024 * 1024;
  a.d_w_eff = d_w_eff;
  CUtensorMap tmA, tmB;
  if (int rc = make_map_cm(&tmA, srcT, 3 * g.ic, items, g.h * g.w, g.ic)) return rc;
  if (int rc = make_map_cm(&tmB, doutT, g.oc, items * nphase, g.h * g.w, g.oc)) return rc;
  if (int rc = opt_in_smem(upconv_wgrad_tc_kernel, "rcb_upconv_wgrad_tc")) return rc;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int ngroups = nphase / ppc;
  RCB_CHECK_ARG(ngroups <= WG_MAX_GROUPS, "rcb_upconv_wgrad_tc: too many phase groups (%d)", ngroups);
  int max_shift = 0;
  for (int grp = 0; grp < ngroups; ++grp) {     // each phase group has its own set of distinct shifts
    WgGroup& b = a.groups[grp];
    b.phase0 = grp * ppc;
    b.nshift = 0;
    for (int lp = 0; lp < ppc; ++lp)
      for (int t = 0; t < 4; ++t) {
        const int ph = b.phase0 + lp, ry = ph / g.fx, rx = ph % g.fx;
        const int dy = (ry < g.py ? -1 : 0) + (t >> 1), dx = (rx < g.px ? -1 : 0) + (t & 1);
        int idx = -1;
        for (int i = 0; i < b.nshift; ++i) if (b.shift_dy[i] == dy && b.shift_dx[i] == dx) idx = i;
        if (idx < 0) {
          RCB_CHECK_ARG(b.nshift < WG_MAX_SHIFT, "rcb_upconv_wgrad_tc: too many shifts");
          idx = b.nshift; b.shift_dy[idx] = dy; b.shift_dx[idx] = dx; ++b.nshift;
        }
        b.seg_shift[lp * 4 + t] = idx;
      }
    if (b.nshift > max_shift) max_shift = b.nshift;
  }
  a.stage_bytes = max_shift * a.a_bytes + ppc * a.b_bytes;
  a.bar_off = 2 * a.stage_bytes;
  const int smem_total = a.bar_off + 256 + 1024;
  RCB_CHECK_ARG(smem_total <= 200 * 1024, "rcb_upconv_wgrad_tc: stage does not fit shared memory");
  int ksplit = (sms + ngroups - 1) / ngroups;    // the phase groups run concurrently (grid.y): they share the SMs
  if (ksplit > a.kb_total) ksplit = a.kb_total;
  a.kb_per_cta = (a.kb_total + ksplit - 1) / ksplit;
  const int nct = (a.kb_total + a.kb_per_cta - 1) / a.kb_per_cta;
  upconv_wgrad_tc_kernel<<<dim3(nct, ngroups), TC_THREADS, smem_total, st>>>(tmA, tmB, a);
  RCB_CHECK_LAUNCH("rcb_upconv_wgrad_tc");
  return 0;
}
