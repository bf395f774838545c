// Feed-forward and k=7 convolution layers of the FFConvLSTM encoder on tcgen05 (nnsvs/model.py:837-859,914-915,922;
// SURVEY.md §8(f) row 1), plus the small kernels around them.
//
// svsk_tapgemm_bf16:  Y[b][t][co] = act( bias[co] + sum_j sum_ci W_j[co][ci] * X[b][t + j][ci] ),  j < ksize
// over frame-major bf16 activations.  The convolution reads a buffer the caller has already padded in time (reflection
// rows written by svsk_reflect_pad_rows_bf16), so every tap is a plain row-shifted TMA box and the kernel is one GEMM
// loop over ksize x ceil(Cin/64) k-blocks; ksize = 1 is nn.Linear.  BatchNorm (eval) is folded into W and bias when the
// weights are packed.  Grid = (128-frame tiles) x (256-wide blocks of output channels); M = 128 frames on the TMEM lanes.
#include <cuda_bf16.h>

#include "sm100_ptx.cuh"
#include "svsk_common.cuh"
#include "tma_util.cuh"

namespace svsk {

constexpr int kTgStages = 4;
constexpr int kTgABytes = 128 * 128;  // 128 frames x 64 bf16

struct TapGemmArgs {
  const float* bias;
  __nv_bfloat16* y_b;
  float* y_f;
  int B, T, Cout, ksize, KB, tiles_per_track;
  int Tp_x, Tp_y, y_row0, ldy_b, ldy_f, act, bw;
};

struct __align__(8) TapGemmBarriers {
  uint64_t full[kTgStages];
  uint64_t empty[kTgStages];
  uint64_t d_full;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(192, 1)
tapgemm_bf16_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w, const TapGemmArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stage_bytes = kTgABytes + a.bw * 128;
  TapGemmBarriers* bars = reinterpret_cast<TapGemmBarriers*>(smem + kTgStages * stage_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / a.tiles_per_track, t0 = (blockIdx.x % a.tiles_per_track) * 128;
  const int co0 = blockIdx.y * 256;
  const int N = min(256, a.Cout - co0);
  const int iters = a.ksize * a.KB;
  const uint32_t tmem_cols = N <= 32 ? 32 : (N <= 64 ? 64 : (N <= 128 ? 128 : 256));

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_x);
    ptx::prefetch_tmap(&tm_w);
    for (int i = 0; i < kTgStages; ++i) {
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
      int s = 0;
      uint32_t ph = 0;
      const int row0 = b * a.Tp_x + t0;
      for (int j = 0; j < a.ksize; ++j) {
        for (int kb = 0; kb < a.KB; ++kb) {
          ptx::mbar_wait(&bars->empty[s], ph ^ 1);
          uint8_t* As = smem + s * stage_bytes;
          ptx::mbar_arrive_expect_tx(&bars->full[s], stage_bytes);
          ptx::tma_load_2d(As, &tm_x, &bars->full[s], kb * 64, row0 + j);
          ptx::tma_load_2d(As + kTgABytes, &tm_w, &bars->full[s], kb * 64, j * a.Cout + co0);
          if (++s == kTgStages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = ptx::umma_idesc_bf16_f32(128, N);
      int s = 0;
      uint32_t ph = 0;
      for (int it = 0; it < iters; ++it) {
        ptx::mbar_wait(&bars->full[s], ph);
        ptx::tc_fence_after();
        const uint32_t a0 = ptx::smem_u32(smem + s * stage_bytes);
        const uint32_t b0 = a0 + kTgABytes;
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4)
          ptx::umma_bf16(tmem, ptx::umma_desc_k_sw128(a0 + k4 * 32), ptx::umma_desc_k_sw128(b0 + k4 * 32), idesc,
                         (it | k4) != 0);
        ptx::umma_commit(&bars->empty[s]);
        if (++s == kTgStages) { s = 0; ph ^= 1; }
      }
      ptx::umma_commit(&bars->d_full);
    }
  } else {
    const int q = warp & 3;
    const int t = t0 + q * 32 + lane;
    ptx::mbar_wait(&bars->d_full, 0);
    ptx::tc_fence_after();
    const bool ok = t < a.T;
    float* yf = a.y_f ? a.y_f + ((size_t)b * a.T + t) * a.ldy_f + co0 : nullptr;
    __nv_bfloat16* yb = a.y_b ? a.y_b + ((size_t)b * a.Tp_y + a.y_row0 + t) * a.ldy_b + co0 : nullptr;
    const bool vec_f = yf && (a.ldy_f % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.y_f) & 15) == 0);
    const bool vec_b = yb && (a.ldy_b % 8 == 0) && ((reinterpret_cast<uintptr_t>(a.y_b) & 15) == 0);
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t r[16];
      ptx::tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + c0, r);
      ptx::tmem_ld_wait();
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        v[i] = __uint_as_float(r[i]) + (a.bias ? a.bias[co0 + c0 + i] : 0.f);
        if (a.act == SVSK_ACT_RELU) v[i] = fmaxf(v[i], 0.f);
      }
      if (!ok) continue;
      if (yf) {
        if (vec_f) {
#pragma unroll
          for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(yf + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) yf[c0 + i] = v[i];
        }
      }
      if (yb) {
        if (vec_b) {
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) pk[i] = ptx::pack_bf16(v[2 * i], v[2 * i + 1]);
          *reinterpret_cast<uint4*>(yb + c0) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(yb + c0 + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) yb[c0 + i] = __float2bfloat16_rn(v[i]);
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem, tmem_cols);
}

// wp[j][co][k] = bf16( w[co][k][j] * scale[co] ), zero for Cin <= k < Kp
__global__ void tapgemm_pack_kernel(const float* __restrict__ w, const float* __restrict__ scale, __nv_bfloat16* __restrict__ wp,
                                    int Cout, int Cin, int ksize, int Kp) {
  const long long n = (long long)ksize * Cout * Kp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % Kp);
    const int co = (int)((i / Kp) % Cout);
    const int j = (int)(i / ((long long)Kp * Cout));
    float v = 0.f;
    if (k < Cin) v = w[((size_t)co * Cin + k) * ksize + j] * (scale ? scale[co] : 1.f);
    wp[i] = __float2bfloat16_rn(v);
  }
}

// buf [B][Tp][C]: data in rows pad .. pad+T-1; rows pad-1-i <- row pad+1+i and rows pad+T+i <- row pad+T-2-i (i < pad).
__global__ void reflect_pad_rows_kernel(__nv_bfloat16* buf, int B, int Tp, int C8, int T, int pad) {
  const long long n = (long long)B * 2 * pad * C8;
  uint4* base = reinterpret_cast<uint4*>(buf);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8);
    const int r = (int)((i / C8) % (2 * pad));
    const int b = (int)(i / ((long long)C8 * 2 * pad));
    int dst, src;
    if (r < pad) { dst = pad - 1 - r; src = pad + 1 + r; }
    else { const int k = r - pad; dst = pad + T + k; src = pad + T - 2 - k; }
    base[((size_t)b * Tp + dst) * C8 + c] = base[((size_t)b * Tp + src) * C8 + c];
  }
}

// One thread per input row: y = x with the one-hot block [start, start+V) replaced by the exact one-hot of its argmax
// (first maximum, like torch.argmax; an all-zero block therefore selects index 0 — model.py:908).
__global__ void encoder_front_kernel(const float* __restrict__ x, float* __restrict__ yf, __nv_bfloat16* __restrict__ yb,
                                     long long rows, int in_dim, int start, int V, int ldy_f, int ldy_b) {
  const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float* xr = x + r * in_dim;
  int best = 0;
  if (V > 0) {
    float bv = xr[start];
    for (int i = 1; i < V; ++i) {
      const float v = xr[start + i];
      if (v > bv) { bv = v; best = i; }
    }
  }
  for (int i = 0; i < in_dim; ++i) {
    float v = xr[i];
    if (i >= start && i < start + V) v = (i - start == best) ? 1.f : 0.f;
    if (yf) yf[r * ldy_f + i] = v;
    if (yb) yb[r * ldy_b + i] = __float2bfloat16_rn(v);
  }
  if (yf) for (int i = in_dim; i < ldy_f; ++i) yf[r * ldy_f + i] = 0.f;
  if (yb) for (int i = in_dim; i < ldy_b; ++i) yb[r * ldy_b + i] = __float2bfloat16_rn(0.f);
}

}  // namespace svsk

using namespace svsk;

extern "C" int svsk_tapgemm_pack_bf16(const float* w, const float* scale, void* wp, int Cout, int Cin, int ksize, void* stream) {
  SVSK_REQUIRE(w && wp && Cout > 0 && Cin > 0 && ksize >= 1 && ksize <= 15, SVSK_E_ARG, "tapgemm_pack_bf16: bad args");
  const int Kp = (Cin + 63) / 64 * 64;
  const long long n = (long long)ksize * Cout * Kp;
  const unsigned grid = (unsigned)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
  tapgemm_pack_kernel<<<grid, 256, 0, as_stream(stream)>>>(w, scale, reinterpret_cast<__nv_bfloat16*>(wp), Cout, Cin, ksize, Kp);
  return check_launch("tapgemm_pack_bf16");
}

extern "C" int svsk_reflect_pad_rows_bf16(void* buf, int B, int Tp, int C, int T, int pad, void* stream) {
  SVSK_REQUIRE(buf && B > 0 && C > 0 && C % 8 == 0 && pad >= 1, SVSK_E_ARG, "reflect_pad_rows_bf16: bad args (C %% 8)");
  SVSK_REQUIRE(T > pad && Tp >= T + 2 * pad, SVSK_E_ARG, "reflect_pad_rows_bf16: needs T > pad and Tp >= T + 2 pad (T=%d Tp=%d pad=%d)",
               T, Tp, pad);
  SVSK_REQUIRE((reinterpret_cast<uintptr_t>(buf) & 15) == 0, SVSK_E_ALIGN, "reflect_pad_rows_bf16: buffer must be 16-byte aligned");
  const long long n = (long long)B * 2 * pad * (C / 8);
  reflect_pad_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(reinterpret_cast<__nv_bfloat16*>(buf), B, Tp,
                                                                                     C / 8, T, pad);
  return check_launch("reflect_pad_rows_bf16");
}

extern "C" int svsk_encoder_front(const float* x, float* y_f32, void* y_bf16, long long rows, int in_dim, int onehot_start,
                                  int onehot_len, int ldy_f, int ldy_b, void* stream) {
  SVSK_REQUIRE(x && (y_f32 || y_bf16) && rows > 0 && in_dim > 0, SVSK_E_ARG, "encoder_front: bad args");
  SVSK_REQUIRE(onehot_len >= 0 && onehot_start >= 0 && onehot_start + onehot_len <= in_dim, SVSK_E_ARG,
               "encoder_front: one-hot block [%d, %d) outside the %d input columns", onehot_start, onehot_start + onehot_len, in_dim);
  SVSK_REQUIRE((!y_f32 || ldy_f >= in_dim) && (!y_bf16 || ldy_b >= in_dim), SVSK_E_ARG, "encoder_front: output pitch");
  encoder_front_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, as_stream(stream)>>>(
      x, y_f32, reinterpret_cast<__nv_bfloat16*>(y_bf16), rows, in_dim, onehot_start, onehot_len, ldy_f, ldy_b);
  return check_launch("encoder_front");
}

extern "C" int svsk_tapgemm_bf16(const svsk_tapgemm_bf16_params* pp, void* stream) {
  SVSK_REQUIRE(pp != nullptr, SVSK_E_ARG, "tapgemm_bf16: null params");
  const svsk_tapgemm_bf16_params& p = *pp;
  SVSK_REQUIRE(p.x && p.wp && (p.y_bf16 || p.y_f32), SVSK_E_ARG, "tapgemm_bf16: null tensor");
  SVSK_REQUIRE(p.B > 0 && p.T > 0 && p.Cin > 0 && p.ksize >= 1 && p.ksize <= 15, SVSK_E_ARG, "tapgemm_bf16: B=%d T=%d Cin=%d k=%d",
               p.B, p.T, p.Cin, p.ksize);
  SVSK_REQUIRE(p.Cout >= 16 && p.Cout % 16 == 0, SVSK_E_ARG, "tapgemm_bf16: Cout=%d must be a multiple of 16", p.Cout);
  SVSK_REQUIRE(p.ldx >= p.Cin && p.ldx % 8 == 0, SVSK_E_ALIGN, "tapgemm_bf16: ldx=%d (>= Cin, %% 8)", p.ldx);
  SVSK_REQUIRE(p.Tp_x >= p.T + p.ksize - 1, SVSK_E_ARG, "tapgemm_bf16: input holds %d rows per track, needs T + k - 1 = %d", p.Tp_x,
               p.T + p.ksize - 1);
  SVSK_REQUIRE(!p.y_bf16 || (p.ldy_b >= p.Cout && p.y_row0 >= 0 && p.Tp_y >= p.y_row0 + p.T), SVSK_E_ARG, "tapgemm_bf16: bf16 output geometry");
  SVSK_REQUIRE(!p.y_f32 || p.ldy_f >= p.Cout, SVSK_E_ARG, "tapgemm_bf16: ldy_f");
  SVSK_REQUIRE((long long)p.B * p.Tp_x < (1ll << 31), SVSK_E_ARG, "tapgemm_bf16: too many rows");
  int rc = require_sm100();
  if (rc) return rc;

  const int Kp = (p.Cin + 63) / 64 * 64;
  const int bw = p.Cout < 256 ? p.Cout : 256;
  CUtensorMap tm_x, tm_w;
  {
    uint64_t dims[2] = {(uint64_t)p.Cin, (uint64_t)p.B * p.Tp_x};
    uint64_t str[1] = {(uint64_t)p.ldx * 2};
    uint32_t box[2] = {64, 128};
    if ((rc = make_tmap_bf16(&tm_x, p.x, 2, dims, str, box))) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)Kp, (uint64_t)p.ksize * p.Cout};
    uint64_t str[1] = {(uint64_t)Kp * 2};
    uint32_t box[2] = {64, (uint32_t)bw};
    if ((rc = make_tmap_bf16(&tm_w, p.wp, 2, dims, str, box))) return rc;
  }
  const int stage_bytes = kTgABytes + bw * 128;
  const int smem_bytes = kTgStages * stage_bytes + (int)sizeof(TapGemmBarriers) + 1024;
  int dev = 0;
  cudaGetDevice(&dev);
  static bool attr_set[64] = {false};
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(tapgemm_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) return fail((int)e, "tapgemm_bf16: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  TapGemmArgs a;
  a.bias = p.bias;
  a.y_b = reinterpret_cast<__nv_bfloat16*>(p.y_bf16);
  a.y_f = p.y_f32;
  a.B = p.B; a.T = p.T; a.Cout = p.Cout; a.ksize = p.ksize; a.KB = Kp / 64;
  a.tiles_per_track = (p.T + 127) / 128;
  a.Tp_x = p.Tp_x; a.Tp_y = p.Tp_y; a.y_row0 = p.y_row0; a.ldy_b = p.ldy_b; a.ldy_f = p.ldy_f; a.act = p.act; a.bw = bw;
  dim3 grid((unsigned)(p.B * a.tiles_per_track), (unsigned)((p.Cout + 255) / 256));
  tapgemm_bf16_kernel<<<grid, 192, smem_bytes, as_stream(stream)>>>(tm_x, tm_w, a);
  return check_launch("tapgemm_bf16");
}

// ---------------------------------------------------------------------------------------------------------------------
// BatchNorm1d in TRAINING mode (model.py:839-852 under module.train(): the diffusion recipe's encoders have dropout = 0, so
// this is all that differs from the eval forward): statistics of a [B][C][T] fp32 tensor over all B * T positions of a
// channel, padded frames included, as torch.nn.BatchNorm1d takes them.  One CTA per channel, fp64 sums; the running
// buffers move by `momentum` towards (mean, unbiased variance).
namespace svsk {

__global__ void __launch_bounds__(256) bn_batch_stats_kernel(const float* __restrict__ x, int B, int C, int T,
                                                             float* __restrict__ mean, float* __restrict__ var,
                                                             float* __restrict__ run_mean, float* __restrict__ run_var,
                                                             float momentum) {
  const int c = blockIdx.x;
  double s = 0.0, ss = 0.0;
  for (int b = 0; b < B; ++b) {
    const float* row = x + ((size_t)b * C + c) * T;
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
      const double v = row[t];
      s += v;
      ss += v * v;
    }
  }
  __shared__ double sh[2][8];
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s; sh[1][threadIdx.x >> 5] = ss; }
  __syncthreads();
  if (threadIdx.x == 0) {
    s = ss = 0.0;
    for (int w = 0; w < 8; ++w) { s += sh[0][w]; ss += sh[1][w]; }
    const double n = (double)B * T;
    const double m = s / n;
    double v = ss / n - m * m;
    if (v < 0.0) v = 0.0;
    mean[c] = (float)m;
    var[c] = (float)v;
    if (run_mean) run_mean[c] = (float)((1.0 - momentum) * run_mean[c] + momentum * m);
    if (run_var) run_var[c] = (float)((1.0 - momentum) * run_var[c] + momentum * v * (n > 1.0 ? n / (n - 1.0) : 1.0));
  }
}

__global__ void bn_apply_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ mean,
                                const float* __restrict__ var, const float* __restrict__ gamma, const float* __restrict__ beta,
                                float eps, int relu, int C, int T, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = (int)((i / T) % C);
  float v = (x[i] - mean[c]) * (1.0f / sqrtf(var[c] + eps));
  v = v * (gamma ? gamma[c] : 1.0f) + (beta ? beta[c] : 0.0f);
  y[i] = relu ? fmaxf(v, 0.f) : v;
}

}  // namespace svsk

extern "C" int svsk_bn_batch_stats_f32(const float* x, int B, int C, int T, float* mean, float* var, float* running_mean,
                                       float* running_var, float momentum, void* stream) {
  SVSK_REQUIRE(x && mean && var, SVSK_E_ARG, "bn_batch_stats_f32: null tensor");
  SVSK_REQUIRE(B > 0 && C > 0 && T > 0 && C <= 65535 * 32, SVSK_E_ARG, "bn_batch_stats_f32: bad B/C/T");
  SVSK_REQUIRE(momentum >= 0.f && momentum <= 1.f, SVSK_E_ARG, "bn_batch_stats_f32: momentum %f", (double)momentum);
  int rc = require_sm100();
  if (rc) return rc;
  bn_batch_stats_kernel<<<(unsigned)C, 256, 0, as_stream(stream)>>>(x, B, C, T, mean, var, running_mean, running_var, momentum);
  return check_launch("bn_batch_stats_f32");
}

extern "C" int svsk_bn_apply_f32(const float* x, float* y, const float* mean, const float* var, const float* gamma,
                                 const float* beta, float eps, int relu, int B, int C, int T, void* stream) {
  SVSK_REQUIRE(x && y && mean && var, SVSK_E_ARG, "bn_apply_f32: null tensor");
  SVSK_REQUIRE(B > 0 && C > 0 && T > 0 && eps > 0.f, SVSK_E_ARG, "bn_apply_f32: bad B/C/T/eps");
  int rc = require_sm100();
  if (rc) return rc;
  const size_t n = (size_t)B * C * T;
  bn_apply_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(x, y, mean, var, gamma, beta, eps, relu, C, T, n);
  return check_launch("bn_apply_f32");
}
