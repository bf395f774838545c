// Training kernels of the DiffNet residual stack on tcgen05: the forward that keeps what the
// backward needs, and the backward itself — dgrad, wgrad and the fused gate backward — of ResidualBlock.forward
// (nnsvs/diffsinger/denoiser.py:54-66) inside the training step of nnsvs/bin/train_acoustic_multitrack.py:358-380.
//
// Two GEMM kernels carry all of it:
//
//  svsk_seggemm_bf16   frame-major:  D[b][t][n] = sum_s sum_k  X_s[b][t + shift_s][k] * Wp[n][koff_s + k]
//      M = 128 frames on the TMEM lanes, N <= 256 output columns per CTA, K streamed through a 4-stage TMA ring; every
//      segment s is its own NTC bf16 tensor read at a row offset (rows outside [0, T) of a track are zero-filled by the
//      TMA unit = the conv's zero padding, forward and transposed).  The epilogue mode selects the layer:
//        GATE_FWD   y = D + b1 (packed rows: 128 gate rows then the 128 filter rows of the same channels)
//                   -> ypre (kept for the backward) and z = sigmoid(y_gate) * tanh(y_filter)
//        RES_SKIP   o = D + bout: x' = (x + o[:C]) / sqrt 2 -> x' and xd' = x' + dp_next (the next conv's input);
//                   skip32 (+)= o[C:]
//        GATE_BWD   dz = D (= do . Wout) -> dy_gate = dz tanh(yf) s (1 - s), dy_filter = dz s (1 - tanh^2 yf), s = sigmoid(yg)
//        ADD_SCALE  (D + add) * alpha [* (mask > 0)]   dgrad: u_l = (conv^T dy_l + u_{l+1}) / sqrt 2
//        PLAIN      act(D + bias) * alpha [* (mask > 0)] (+= into an fp32 buffer): head, tail and their backward
//
//  svsk_wgrad_bf16     channel-major: dW[n][koff_s + k] = sum_b sum_t  P[b][n][t] * Q_s[b][k][t + shift_s]
//      both operands are [B][channels][T] (time contiguous = K-major), M = 128 rows of P, N = 128 rows of Q_s, K = all
//      frames of all tracks; one CTA owns one 128 x 128 tile of dW for the whole contraction (no split-K, no atomics:
//      bit-reproducible gradients).
//
// plus the layout change [B][T][N] -> [B][N][T] the wgrad operands need and the weight packing of all layers in one
// launch.  Bias gradients and the step-embedding gradient are column sums of dy over time: the caller appends indicator
// rows (all frames / first d / last d of each track) to a wgrad operand, so they come out of the same GEMM.
#include <cuda_bf16.h>

#include "sm100_ptx.cuh"
#include "svsk_common.cuh"
#include "tma_util.cuh"

namespace svsk {

constexpr int kSgStages = 4;
constexpr int kSgABytes = 128 * 128;  // 128 frames x 64 bf16
constexpr int kSgMaxSeg = 4;

struct SegGemmArgs {
  int nseg;
  int shift[kSgMaxSeg];
  int kb[kSgMaxSeg];  // 64-wide k-blocks per segment
  int B, T, tiles_per_track, Nblk, Nrows, mode, C, init, act, accumulate;
  int ld_in0, ld_mask, ld_out0, ld_out1, ld_outf;
  float alpha;
  const float* bias;
  const __nv_bfloat16* in0;
  const __nv_bfloat16* mask;
  const float* dp_next;
  __nv_bfloat16* out0;
  __nv_bfloat16* out1;
  float* outf;
};

struct __align__(8) SegGemmBarriers {
  uint64_t full[kSgStages];
  uint64_t empty[kSgStages];
  uint64_t d_full;
  uint32_t tmem_base;
};

__device__ __forceinline__ void store16_bf16(__nv_bfloat16* p, const float (&v)[16]) {
  uint32_t pk[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) pk[i] = ptx::pack_bf16(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  *reinterpret_cast<uint4*>(p + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
}

__device__ __forceinline__ void load16_bf16(const __nv_bfloat16* p, float (&v)[16]) {
  const uint4 a = *reinterpret_cast<const uint4*>(p);
  const uint4 b = *reinterpret_cast<const uint4*>(p + 8);
  const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
  }
}

__global__ void __launch_bounds__(192, 1)
seggemm_bf16_kernel(const __grid_constant__ CUtensorMap tm0, const __grid_constant__ CUtensorMap tm1,
                    const __grid_constant__ CUtensorMap tm2, const __grid_constant__ CUtensorMap tm3,
                    const __grid_constant__ CUtensorMap tm_w, const SegGemmArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stage_bytes = kSgABytes + a.Nblk * 128;
  SegGemmBarriers* bars = reinterpret_cast<SegGemmBarriers*>(smem + kSgStages * stage_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / a.tiles_per_track, t0 = (blockIdx.x % a.tiles_per_track) * 128;
  const int n0 = blockIdx.y * a.Nblk;
  const int N = min(a.Nblk, a.Nrows - n0);
  int iters = 0;
  for (int s = 0; s < a.nseg; ++s) iters += a.kb[s];
  const uint32_t tmem_cols = a.Nblk <= 32 ? 32 : (a.Nblk <= 64 ? 64 : (a.Nblk <= 128 ? 128 : 256));

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm0);
    ptx::prefetch_tmap(&tm_w);
    for (int i = 0; i < kSgStages; ++i) {
      ptx::mbar_init(&bars->full[i], 1);
      ptx::mbar_init(&bars->empty[i], 1);
    }
    ptx::mbar_init(&bars->d_full, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(&bars->tmem_base, tmem_cols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      const CUtensorMap* maps[kSgMaxSeg] = {&tm0, &tm1, &tm2, &tm3};
      int st = 0, kcol = 0;
      uint32_t ph = 0;
      for (int s = 0; s < a.nseg; ++s) {
        for (int kb = 0; kb < a.kb[s]; ++kb, ++kcol) {
          ptx::mbar_wait(&bars->empty[st], ph ^ 1);
          uint8_t* As = smem + st * stage_bytes;
          ptx::mbar_arrive_expect_tx(&bars->full[st], stage_bytes);
          ptx::tma_load_3d(As, maps[s], &bars->full[st], kb * 64, t0 + a.shift[s], b);
          ptx::tma_load_2d(As + kSgABytes, &tm_w, &bars->full[st], kcol * 64, n0);
          if (++st == kSgStages) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = ptx::umma_idesc_bf16_f32(128, (uint32_t)((N + 15) / 16 * 16));
      int st = 0;
      uint32_t ph = 0;
      for (int it = 0; it < iters; ++it) {
        ptx::mbar_wait(&bars->full[st], ph);
        ptx::tc_fence_after();
        const uint32_t a0 = ptx::smem_u32(smem + st * stage_bytes);
        const uint32_t b0 = a0 + kSgABytes;
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4)
          ptx::umma_bf16(tmem, ptx::umma_desc_k_sw128(a0 + k4 * 32), ptx::umma_desc_k_sw128(b0 + k4 * 32), idesc,
                         (it | k4) != 0);
        ptx::umma_commit(&bars->empty[st]);
        if (++st == kSgStages) { st = 0; ph ^= 1; }
      }
      ptx::umma_commit(&bars->d_full);
    }
  } else {
    const int q = warp & 3;
    const int t = t0 + q * 32 + lane;
    ptx::mbar_wait(&bars->d_full, 0);
    ptx::tc_fence_after();
    const bool ok = t < a.T;
    const size_t row = (size_t)b * a.T + (ok ? t : 0);
    const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16);
    if (a.mode == SVSK_SEG_GATE_FWD) {
      // columns [0,128): gate rows, [128,256): filter rows of channels blockIdx.y * 128 ..
      __nv_bfloat16* yp = a.out0 + row * a.ld_out0 + n0;
      __nv_bfloat16* zp = a.out1 + row * a.ld_out1 + blockIdx.y * 128;
      for (int c0 = 0; c0 < 128; c0 += 16) {
        uint32_t rg[16], rf[16];
        ptx::tmem_ld16(tbase + c0, rg);
        ptx::tmem_ld16(tbase + 128 + c0, rf);
        ptx::tmem_ld_wait();
        float g[16], f[16], z[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          g[i] = __uint_as_float(rg[i]) + a.bias[n0 + c0 + i];
          f[i] = __uint_as_float(rf[i]) + a.bias[n0 + 128 + c0 + i];
          z[i] = ptx::sigmoid_approx(g[i]) * ptx::tanh_approx(f[i]);
        }
        if (ok) {
          store16_bf16(yp + c0, g);
          store16_bf16(yp + 128 + c0, f);
          store16_bf16(zp + c0, z);
        }
      }
    } else if (a.mode == SVSK_SEG_GATE_BWD) {
      // D = dz of channels blockIdx.y * 128 ..; ypre holds their gate / filter pre-activations at packed columns
      const int p0 = blockIdx.y * 256;
      const __nv_bfloat16* yp = a.in0 + row * a.ld_in0 + p0;
      __nv_bfloat16* dyp = a.out0 + row * a.ld_out0 + p0;
      for (int c0 = 0; c0 < 128; c0 += 16) {
        uint32_t r[16];
        ptx::tmem_ld16(tbase + c0, r);
        ptx::tmem_ld_wait();
        float g[16], f[16];
        load16_bf16(yp + c0, g);
        load16_bf16(yp + 128 + c0, f);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float dz = __uint_as_float(r[i]);
          const float s = ptx::sigmoid_approx(g[i]), th = ptx::tanh_approx(f[i]);
          g[i] = dz * th * s * (1.f - s);
          f[i] = dz * s * (1.f - th * th);
        }
        if (ok) {
          store16_bf16(dyp + c0, g);
          store16_bf16(dyp + 128 + c0, f);
        }
      }
    } else if (a.mode == SVSK_SEG_RES_SKIP) {
      // blockIdx.y == 0: residual rows [0, C); == 1: skip rows [C, 2C)
      const float rs2 = 0.70710678118654752440f;
      for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t r[16];
        ptx::tmem_ld16(tbase + c0, r);
        ptx::tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]) + a.bias[n0 + c0 + i];
        if (!ok) continue;
        if (blockIdx.y == 0) {
          float x[16];
          load16_bf16(a.in0 + row * a.ld_in0 + c0, x);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = (x[i] + v[i]) * rs2;
          if (a.out0) store16_bf16(a.out0 + row * a.ld_out0 + c0, v);
          if (a.out1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += a.dp_next ? a.dp_next[(size_t)b * a.C + c0 + i] : 0.f;
            store16_bf16(a.out1 + row * a.ld_out1 + c0, v);
          }
        } else {
          float* sp = a.outf + row * a.ld_outf + c0;
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            float4 o = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            if (!a.init) {
              const float4 p = *reinterpret_cast<const float4*>(sp + i);
              o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
            }
            *reinterpret_cast<float4*>(sp + i) = o;
          }
        }
      }
    } else {  // SVSK_SEG_ADD_SCALE / SVSK_SEG_PLAIN
      for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t r[16];
        ptx::tmem_ld16(tbase + c0, r);
        ptx::tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]) + (a.bias ? a.bias[n0 + c0 + i] : 0.f);
        if (!ok) continue;
        if (a.in0) {
          float x[16];
          load16_bf16(a.in0 + row * a.ld_in0 + n0 + c0, x);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += x[i];
        }
        if (a.act == SVSK_ACT_RELU) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] *= a.alpha;
        if (a.mask) {
          float m[16];
          load16_bf16(a.mask + row * a.ld_mask + n0 + c0, m);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = m[i] > 0.f ? v[i] : 0.f;
        }
        if (a.out0) store16_bf16(a.out0 + row * a.ld_out0 + n0 + c0, v);
        if (a.out1) {  // the same rows plus a per-track vector (the head's x0 + dp_0)
          float w[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) w[i] = v[i] + (a.dp_next ? a.dp_next[(size_t)b * a.Nrows + n0 + c0 + i] : 0.f);
          store16_bf16(a.out1 + row * a.ld_out1 + n0 + c0, w);
        }
        if (a.outf) {
          float* op = a.outf + row * a.ld_outf + n0 + c0;
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            float4 o = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            if (a.accumulate) {
              const float4 p = *reinterpret_cast<const float4*>(op + i);
              o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
            }
            *reinterpret_cast<float4*>(op + i) = o;
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem, tmem_cols);
}

// ------------------------------------------------------------------------------------------------ wgrad
constexpr int kWgStages = 6;
constexpr int kWgTile = 128 * 128;  // 128 rows x 64 frames (bf16)
constexpr int kWgMaxSeg = 5;

struct WgradArgs {
  int nseg;
  int shift[kWgMaxSeg];
  int ktiles[kWgMaxSeg];  // 128-row tiles of Q_s
  int krows[kWgMaxSeg];   // rows of Q_s
  int koff[kWgMaxSeg];    // first dW column of segment s
  int B, T, Prows, ldw, accumulate;
  long long split_stride;  // floats between the partial results of consecutive track groups (gridDim.z of them)
  float* dW;
};

struct __align__(8) WgradBarriers {
  uint64_t full[kWgStages];
  uint64_t empty[kWgStages];
  uint64_t d_full;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(192, 1)
wgrad_bf16_kernel(const __grid_constant__ CUtensorMap tm_p, const __grid_constant__ CUtensorMap tq0,
                  const __grid_constant__ CUtensorMap tq1, const __grid_constant__ CUtensorMap tq2,
                  const __grid_constant__ CUtensorMap tq3, const __grid_constant__ CUtensorMap tq4, const WgradArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  WgradBarriers* bars = reinterpret_cast<WgradBarriers*>(smem + kWgStages * 2 * kWgTile);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * 128;
  int seg = 0, kt = blockIdx.y;
  while (seg < a.nseg - 1 && kt >= a.ktiles[seg]) { kt -= a.ktiles[seg]; ++seg; }
  const int k0 = kt * 128;
  const int tblocks = (a.T + 63) / 64;
  // gridDim.z track groups, each with its own partial dW (summed by the caller: deterministic, no atomics)
  const int b_begin = (int)(((long long)a.B * blockIdx.z) / gridDim.z), b_end = (int)(((long long)a.B * (blockIdx.z + 1)) / gridDim.z);
  const int iters = (b_end - b_begin) * tblocks;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_p);
    for (int i = 0; i < kWgStages; ++i) {
      ptx::mbar_init(&bars->full[i], 1);
      ptx::mbar_init(&bars->empty[i], 1);
    }
    ptx::mbar_init(&bars->d_full, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(&bars->tmem_base, 128);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      const CUtensorMap* maps[kWgMaxSeg] = {&tq0, &tq1, &tq2, &tq3, &tq4};
      const CUtensorMap* tq = maps[seg];
      const int shift = a.shift[seg];
      int st = 0;
      uint32_t ph = 0;
      for (int b = b_begin; b < b_end; ++b) {
        for (int tb = 0; tb < tblocks; ++tb) {
          ptx::mbar_wait(&bars->empty[st], ph ^ 1);
          uint8_t* Ps = smem + st * 2 * kWgTile;
          ptx::mbar_arrive_expect_tx(&bars->full[st], 2 * kWgTile);
          ptx::tma_load_3d(Ps, &tm_p, &bars->full[st], tb * 64, n0, b);
          ptx::tma_load_3d(Ps + kWgTile, tq, &bars->full[st], tb * 64 + shift, k0, b);
          if (++st == kWgStages) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = ptx::umma_idesc_bf16_f32(128, 128);
      int st = 0;
      uint32_t ph = 0;
      for (int it = 0; it < iters; ++it) {
        ptx::mbar_wait(&bars->full[st], ph);
        ptx::tc_fence_after();
        const uint32_t a0 = ptx::smem_u32(smem + st * 2 * kWgTile);
        const uint32_t b0 = a0 + kWgTile;
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4)
          ptx::umma_bf16(tmem, ptx::umma_desc_k_sw128(a0 + k4 * 32), ptx::umma_desc_k_sw128(b0 + k4 * 32), idesc,
                         (it | k4) != 0);
        ptx::umma_commit(&bars->empty[st]);
        if (++st == kWgStages) { st = 0; ph ^= 1; }
      }
      if (iters > 0) ptx::umma_commit(&bars->d_full);
    }
  } else {
    const int q = warp & 3;
    const int n = n0 + q * 32 + lane;
    if (iters > 0) {
      ptx::mbar_wait(&bars->d_full, 0);
      ptx::tc_fence_after();
    }
    const bool ok = n < a.Prows;
    const int ncols = min(128, a.krows[seg] - k0);
    float* out = a.dW + (size_t)blockIdx.z * a.split_stride + (size_t)(ok ? n : 0) * a.ldw + a.koff[seg] + k0;
    for (int c0 = 0; c0 < 128; c0 += 16) {
      uint32_t r[16];
      if (iters > 0) {
        ptx::tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + c0, r);
        ptx::tmem_ld_wait();
      } else {   // an empty track group (more groups than tracks) contributes zeros
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = 0u;
      }
      if (!ok || c0 >= ncols) continue;
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        float4 o = make_float4(__uint_as_float(r[i]), __uint_as_float(r[i + 1]), __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
        if (a.accumulate) {
          const float4 p = *reinterpret_cast<const float4*>(out + c0 + i);
          o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
        }
        *reinterpret_cast<float4*>(out + c0 + i) = o;
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem, 128);
}

// ------------------------------------------------------------------------------------------------ helpers
// y[b][row0 + j * N + n][t] = x[b][t + shift_j][n] (0 where t + shift_j falls outside [0, T)), j < nshift, n < N, t < Tp:
// x [B][T][ldx] bf16, y [B][out_rows][Tp] bf16.  The shifts live here and not in the wgrad kernel's TMA coordinates because
// a tiled TMA load traps unless its start coordinate along the innermost dimension is 16-byte aligned (measured on B200:
// shifts of +-8 bf16 frames work, +-4 and +-1 raise an illegal-instruction fault).  64 x 64 tiles, 16-byte global accesses.
__global__ void __launch_bounds__(256)
ntc_to_nct_bf16_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int T, int N, int ldx, int Tp,
                       int out_rows, int row0, int shift0, int shift1, int shift2) {
  __shared__ uint32_t tile[64][33];                 // [frame][channel pair]; 33 words per row: conflict-free column reads
  const int ntiles_n = (N + 63) / 64;
  const int j = blockIdx.y / ntiles_n;
  const int n0 = (blockIdx.y % ntiles_n) * 64;
  const int shift = j == 0 ? shift0 : (j == 1 ? shift1 : shift2);
  const int b = blockIdx.z, t0 = blockIdx.x * 64;
  {
    const int cseg = threadIdx.x & 7;                // 8 channels = 16 bytes
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int r = (threadIdx.x >> 3) + rr * 32;
      const int t = t0 + r + shift, n = n0 + cseg * 8;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (t >= 0 && t < T && n < N) v = *reinterpret_cast<const uint4*>(x + ((size_t)b * T + t) * ldx + n);
      tile[r][cseg * 4 + 0] = v.x; tile[r][cseg * 4 + 1] = v.y; tile[r][cseg * 4 + 2] = v.z; tile[r][cseg * 4 + 3] = v.w;
    }
  }
  __syncthreads();
  {
    const int tseg = threadIdx.x & 7;                // 8 frames = 16 bytes
    const int t = t0 + tseg * 8;
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int nl = (threadIdx.x >> 3) + rr * 32;   // channel within the tile
      const int n = n0 + nl;
      if (n >= N || t >= Tp) continue;
      const int w = nl >> 1, hi = nl & 1;
      uint32_t e[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t word = tile[tseg * 8 + k][w];
        e[k] = hi ? (word >> 16) : (word & 0xFFFFu);
      }
      const uint4 o = make_uint4(e[0] | (e[1] << 16), e[2] | (e[3] << 16), e[4] | (e[5] << 16), e[6] | (e[7] << 16));
      *reinterpret_cast<uint4*>(y + ((size_t)b * out_rows + row0 + (size_t)j * N + n) * Tp + t) = o;
    }
  }
}

static inline __host__ __device__ int packed_row_of(int r, int C) {
  if (r < C) return 256 * (r / 128) + (r % 128);
  const int c = r - C;
  return 256 * (c / 128) + 128 + (c % 128);
}

// All layers in one launch.  Inputs are the fp32 parameters stacked over the layers:
//   wd [L][2C][C][3] dilated conv, wc [L][2C][H] conditioner, wo [L][2C][C] output projection.
// Outputs (bf16): w1p [L][2C][3C+H] (packed rows; K = tap -d | tap 0 | tap +d | cond), woutp [L][2C][C] (reference rows),
//   woutT [L][C][2C] (dz = do . Wout), w1T [L][C][3*2C] (dgrad: K = packed dy columns of tap 0 | 1 | 2),
//   wcT [L][H][2C] (dcond = dy . Wcond, K = packed dy columns).
__global__ void diffnet_train_pack_kernel(const float* __restrict__ wd, const float* __restrict__ wc, const float* __restrict__ wo,
                                          __nv_bfloat16* __restrict__ w1p, __nv_bfloat16* __restrict__ woutp,
                                          __nv_bfloat16* __restrict__ woutT, __nv_bfloat16* __restrict__ w1T,
                                          __nv_bfloat16* __restrict__ wcT, int C, int H) {
  const int l = blockIdx.y, r = blockIdx.x;  // r: reference output row in [0, 2C)
  const int K1 = 3 * C + H, C2 = 2 * C;
  const int pr = packed_row_of(r, C);
  const float* wdl = wd + (size_t)l * C2 * C * 3 + (size_t)r * C * 3;
  const float* wcl = wc + (size_t)l * C2 * H + (size_t)r * H;
  const float* wol = wo + (size_t)l * C2 * C + (size_t)r * C;
  for (int k = threadIdx.x; k < 3 * C; k += blockDim.x) {
    const int j = k / C, ci = k - j * C;
    const __nv_bfloat16 v = __float2bfloat16_rn(wdl[ci * 3 + j]);
    w1p[((size_t)l * C2 + pr) * K1 + k] = v;
    w1T[((size_t)l * C + ci) * (3 * C2) + j * C2 + pr] = v;
  }
  for (int k = threadIdx.x; k < H; k += blockDim.x) {
    const __nv_bfloat16 v = __float2bfloat16_rn(wcl[k]);
    w1p[((size_t)l * C2 + pr) * K1 + 3 * C + k] = v;
    wcT[((size_t)l * H + k) * C2 + pr] = v;
  }
  for (int k = threadIdx.x; k < C; k += blockDim.x) {
    const __nv_bfloat16 v = __float2bfloat16_rn(wol[k]);
    woutp[((size_t)l * C2 + r) * C + k] = v;
    woutT[((size_t)l * C + k) * C2 + r] = v;
  }
}

static int set_max_smem(const void* fn, const char* what) {
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  if (e != cudaSuccess) return fail((int)e, "%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
  return 0;
}

}  // namespace svsk

using namespace svsk;

extern "C" int svsk_seggemm_bf16(const svsk_seggemm_params* pp, void* stream) {
  SVSK_REQUIRE(pp != nullptr, SVSK_E_ARG, "seggemm_bf16: null params");
  const svsk_seggemm_params& p = *pp;
  SVSK_REQUIRE(p.nseg >= 1 && p.nseg <= kSgMaxSeg && p.wp && p.B > 0 && p.T > 0 && p.Nrows >= 16 && p.Nrows % 16 == 0, SVSK_E_ARG,
               "seggemm_bf16: nseg=%d B=%d T=%d Nrows=%d", p.nseg, p.B, p.T, p.Nrows);
  int Ktot = 0;
  for (int s = 0; s < p.nseg; ++s) {
    SVSK_REQUIRE(p.x[s] && p.kx[s] > 0 && p.kx[s] % 64 == 0 && p.ldx[s] >= p.kx[s] && p.ldx[s] % 8 == 0, SVSK_E_ARG,
                 "seggemm_bf16: segment %d: K=%d ld=%d (K %% 64, ld %% 8)", s, p.kx[s], p.ldx[s]);
    SVSK_REQUIRE((reinterpret_cast<uintptr_t>(p.x[s]) & 15) == 0, SVSK_E_ALIGN, "seggemm_bf16: segment %d not 16-byte aligned", s);
    Ktot += p.kx[s];
  }
  int Nblk = 256;
  switch (p.mode) {
    case SVSK_SEG_GATE_FWD:
      SVSK_REQUIRE(p.Nrows % 256 == 0 && p.bias && p.out0 && p.out1 && p.ld_out0 >= p.Nrows && p.ld_out1 >= p.Nrows / 2, SVSK_E_ARG,
                   "seggemm_bf16: GATE_FWD needs 2C %% 256 == 0, bias, ypre and z");
      break;
    case SVSK_SEG_GATE_BWD:
      Nblk = 128;
      SVSK_REQUIRE(p.Nrows % 128 == 0 && p.in0 && p.out0 && p.ld_in0 >= 2 * p.Nrows && p.ld_out0 >= 2 * p.Nrows, SVSK_E_ARG,
                   "seggemm_bf16: GATE_BWD needs C %% 128 == 0, ypre and dy");
      break;
    case SVSK_SEG_RES_SKIP:
      Nblk = p.Nrows / 2;
      SVSK_REQUIRE((Nblk == 128 || Nblk == 256) && p.bias && p.in0 && p.outf && p.C == Nblk && p.ld_in0 >= Nblk && p.ld_outf >= Nblk &&
                       p.ld_outf % 4 == 0, SVSK_E_ARG, "seggemm_bf16: RES_SKIP needs C in {128, 256}, bias, x and skip32");
      break;
    case SVSK_SEG_ADD_SCALE:
    case SVSK_SEG_PLAIN:
      Nblk = p.Nrows < 256 ? p.Nrows : 256;
      SVSK_REQUIRE(p.out0 || p.outf, SVSK_E_ARG, "seggemm_bf16: no output");
      SVSK_REQUIRE(!p.outf || p.ld_outf % 4 == 0, SVSK_E_ALIGN, "seggemm_bf16: fp32 output pitch must be a multiple of 4");
      break;
    default:
      return fail(SVSK_E_ARG, "seggemm_bf16: unknown mode %d", p.mode);
  }
  SVSK_REQUIRE((!p.out0 || p.ld_out0 % 8 == 0) && (!p.out1 || p.ld_out1 % 8 == 0) && (!p.in0 || p.ld_in0 % 8 == 0) &&
                   (!p.mask || p.ld_mask % 8 == 0), SVSK_E_ALIGN, "seggemm_bf16: bf16 pitches must be multiples of 8");
  int rc = require_sm100();
  if (rc) return rc;

  CUtensorMap tm[kSgMaxSeg], tm_w;
  for (int s = 0; s < kSgMaxSeg; ++s) {
    const int u = s < p.nseg ? s : 0;
    uint64_t dims[3] = {(uint64_t)p.kx[u], (uint64_t)p.T, (uint64_t)p.B};
    uint64_t str[2] = {(uint64_t)p.ldx[u] * 2, (uint64_t)p.T * p.ldx[u] * 2};
    uint32_t box[3] = {64, 128, 1};
    if ((rc = make_tmap_bf16(&tm[s], p.x[u], 3, dims, str, box))) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)Ktot, (uint64_t)p.Nrows};
    uint64_t str[1] = {(uint64_t)Ktot * 2};
    uint32_t box[2] = {64, (uint32_t)Nblk};
    if ((rc = make_tmap_bf16(&tm_w, p.wp, 2, dims, str, box))) return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    if ((rc = set_max_smem((const void*)seggemm_bf16_kernel, "seggemm_bf16"))) return rc;
    attr_set = true;
  }
  SegGemmArgs a{};
  a.nseg = p.nseg;
  for (int s = 0; s < p.nseg; ++s) { a.shift[s] = p.shift[s]; a.kb[s] = p.kx[s] / 64; }
  a.B = p.B; a.T = p.T; a.tiles_per_track = (p.T + 127) / 128; a.Nblk = Nblk; a.Nrows = p.Nrows; a.mode = p.mode; a.C = p.C;
  a.init = p.init; a.act = p.act; a.accumulate = p.accumulate;
  a.ld_in0 = p.ld_in0; a.ld_mask = p.ld_mask; a.ld_out0 = p.ld_out0; a.ld_out1 = p.ld_out1; a.ld_outf = p.ld_outf;
  a.alpha = p.alpha;
  a.bias = p.bias; a.in0 = (const __nv_bfloat16*)p.in0; a.mask = (const __nv_bfloat16*)p.mask; a.dp_next = p.dp_next;
  a.out0 = (__nv_bfloat16*)p.out0; a.out1 = (__nv_bfloat16*)p.out1; a.outf = p.outf;
  const int smem_bytes = kSgStages * (kSgABytes + Nblk * 128) + (int)sizeof(SegGemmBarriers) + 1024;
  dim3 grid((unsigned)(p.B * a.tiles_per_track), (unsigned)((p.Nrows + Nblk - 1) / Nblk));
  seggemm_bf16_kernel<<<grid, 192, smem_bytes, as_stream(stream)>>>(tm[0], tm[1], tm[2], tm[3], tm_w, a);
  return check_launch("seggemm_bf16");
}

extern "C" int svsk_wgrad_bf16(const svsk_wgrad_params* pp, void* stream) {
  SVSK_REQUIRE(pp != nullptr, SVSK_E_ARG, "wgrad_bf16: null params");
  const svsk_wgrad_params& p = *pp;
  SVSK_REQUIRE(p.p && p.dW && p.nseg >= 1 && p.nseg <= kWgMaxSeg && p.B > 0 && p.T > 0 && p.Prows > 0 && p.Tp >= p.T && p.Tp % 8 == 0,
               SVSK_E_ARG, "wgrad_bf16: nseg=%d B=%d T=%d Tp=%d (Tp %% 8)", p.nseg, p.B, p.T, p.Tp);
  SVSK_REQUIRE(p.ldw % 4 == 0 && (reinterpret_cast<uintptr_t>(p.dW) & 15) == 0, SVSK_E_ALIGN, "wgrad_bf16: dW pitch %% 4, 16-byte aligned");
  int rc = require_sm100();
  if (rc) return rc;
  WgradArgs a{};
  a.nseg = p.nseg;
  int tiles = 0, koff = 0;
  CUtensorMap tq[kWgMaxSeg], tm_p;
  auto make = [&](CUtensorMap* m, const void* base, int rows) {
    uint64_t dims[3] = {(uint64_t)p.T, (uint64_t)rows, (uint64_t)p.B};
    uint64_t str[2] = {(uint64_t)p.Tp * 2, (uint64_t)rows * p.Tp * 2};
    uint32_t box[3] = {64, 128, 1};
    return make_tmap_bf16(m, base, 3, dims, str, box);
  };
  if ((rc = make(&tm_p, p.p, p.Prows))) return rc;
  for (int s = 0; s < kWgMaxSeg; ++s) {
    const int u = s < p.nseg ? s : 0;
    SVSK_REQUIRE(p.q[u] && p.qrows[u] > 0 && p.qrows[u] % 16 == 0, SVSK_E_ARG, "wgrad_bf16: segment %d: rows=%d (%% 16)", u, p.qrows[u]);
    SVSK_REQUIRE(p.shift[u] % 8 == 0, SVSK_E_ALIGN,
                 "wgrad_bf16: segment %d: shift %d is not a multiple of 8 frames (TMA needs a 16-byte aligned start along time; "
                 "let svsk_ntc_to_nct_bf16 write the shifted copy)", u, p.shift[u]);
    if ((rc = make(&tq[s], p.q[u], p.qrows[u]))) return rc;
    if (s < p.nseg) {
      a.shift[s] = p.shift[s];
      a.krows[s] = p.qrows[s];
      a.ktiles[s] = (p.qrows[s] + 127) / 128;
      a.koff[s] = koff;
      koff += p.qrows[s];
      tiles += a.ktiles[s];
    }
  }
  SVSK_REQUIRE(p.ldw >= koff, SVSK_E_ARG, "wgrad_bf16: dW pitch %d < %d columns", p.ldw, koff);
  SVSK_REQUIRE(p.splits >= 1 && p.splits <= 64 && (p.splits == 1 || p.split_stride >= (long long)p.Prows * p.ldw), SVSK_E_ARG,
               "wgrad_bf16: splits=%d needs split_stride >= Prows * ldw", p.splits);
  a.B = p.B; a.T = p.T; a.Prows = p.Prows; a.ldw = p.ldw; a.accumulate = p.accumulate; a.dW = p.dW;
  a.split_stride = p.split_stride;
  static bool attr_set = false;
  if (!attr_set) {
    if ((rc = set_max_smem((const void*)wgrad_bf16_kernel, "wgrad_bf16"))) return rc;
    attr_set = true;
  }
  const int smem_bytes = kWgStages * 2 * kWgTile + (int)sizeof(WgradBarriers) + 1024;
  dim3 grid((unsigned)((p.Prows + 127) / 128), (unsigned)tiles, (unsigned)p.splits);
  wgrad_bf16_kernel<<<grid, 192, smem_bytes, as_stream(stream)>>>(tm_p, tq[0], tq[1], tq[2], tq[3], tq[4], a);
  return check_launch("wgrad_bf16");
}

extern "C" int svsk_ntc_to_nct_bf16(const void* x, void* y, int B, int T, int N, int ldx, int Tp, int out_rows, int row0,
                                    int nshift, const int* shifts, void* stream) {
  SVSK_REQUIRE(x && y && B > 0 && T > 0 && N > 0 && N % 8 == 0 && ldx >= N && ldx % 8 == 0 && Tp >= T && Tp % 8 == 0 && B <= 65535,
               SVSK_E_ARG, "ntc_to_nct_bf16: bad args (N, ldx, Tp multiples of 8)");
  SVSK_REQUIRE(nshift >= 1 && nshift <= 3 && shifts && row0 >= 0 && out_rows >= row0 + nshift * N, SVSK_E_ARG,
               "ntc_to_nct_bf16: 1..3 shifted copies of N rows must fit rows %d.. of %d", row0, out_rows);
  SVSK_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0, SVSK_E_ALIGN,
               "ntc_to_nct_bf16: 16-byte aligned tensors");
  dim3 grid((unsigned)((Tp + 63) / 64), (unsigned)(nshift * ((N + 63) / 64)), (unsigned)B);
  ntc_to_nct_bf16_kernel<<<grid, 256, 0, as_stream(stream)>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, T, N, ldx, Tp, out_rows,
                                                              row0, shifts[0], nshift > 1 ? shifts[1] : 0, nshift > 2 ? shifts[2] : 0);
  return check_launch("ntc_to_nct_bf16");
}

extern "C" int svsk_diffnet_train_pack(const float* wd, const float* wc, const float* wo, void* w1p, void* woutp, void* woutT,
                                       void* w1T, void* wcT, int L, int C, int H, void* stream) {
  SVSK_REQUIRE(wd && wc && wo && w1p && woutp && woutT && w1T && wcT, SVSK_E_ARG, "diffnet_train_pack: null");
  SVSK_REQUIRE(L > 0 && L <= 65535 && (C == 128 || C == 256) && H > 0 && H % 64 == 0, SVSK_E_ARG,
               "diffnet_train_pack: need C in {128,256}, H %% 64 == 0 (L=%d C=%d H=%d)", L, C, H);
  diffnet_train_pack_kernel<<<dim3(2 * C, L), 256, 0, as_stream(stream)>>>(
      wd, wc, wo, (__nv_bfloat16*)w1p, (__nv_bfloat16*)woutp, (__nv_bfloat16*)woutT, (__nv_bfloat16*)w1T, (__nv_bfloat16*)wcT, C, H);
  return check_launch("diffnet_train_pack");
}
