// uSFGAN front-end kernels that write the channel-last bf16 tensors of the tensor-core stacks DIRECTLY (SURVEY §8(f)
// row 2), instead of fp32 NCT intermediates followed by a layout/precision conversion:
//   * svsk_upsample_fused: the whole UpsampleNetwork (nnsvs/usfgan/layers/upsample.py:61-128: per scale s a nearest-
//     neighbour stretch by s and a (1, 2s+1) single-channel smoothing conv with zero padding s, shared by all aux
//     channels) as ONE pass from frame rate to sample rate.  All stages are linear and channel-independent, so each
//     output sample is a weighted sum of a handful of input frames; a thread derives the weights of its sample by
//     walking the stages top-down (exact at the sequence ends: the zero padding of every stage is applied where the
//     reference applies it), then applies them to all channels out of a shared-memory copy of the frames.
//     Staged path at config 3: 6.1 ms (five fp32 tensors written, the last one 1.4 GB, + NCT->NTC); fused: one 0.7 GB write.
//   * svsk_expand1_bf16: a 1 -> C pointwise Conv1d (generator.py conv_first_sine / conv_first_noise) straight to NTC bf16.
//   * svsk_usfgan_source: the sine-based source signal and the pitch-dependent dilation factors from frame-level F0
//     (nnsvs/usfgan/utils/features.py:56-75, 145-164), one pass over the samples after a per-track scan over the frames.
#include <cuda_bf16.h>

#include "sm100_ptx.cuh"
#include "svsk_common.cuh"
#include "usfgan_fr.cuh"

namespace svsk {

constexpr int kUpMaxStages = 6;
constexpr int kUpWin = 8;      // composite window: at most this many entries per stage level
constexpr int kUpFrames = 48;  // frames staged per block of 128 samples
constexpr int kUpMaxA = 128;

struct UpsampleArgs {
  const float* c;      // [B][A][F]
  const float* taps;   // stage k: 2*s_k+1 taps, concatenated
  __nv_bfloat16* out_b;  // [B][T][Ap] or null
  float* out_f;          // [B][T][Ap] or null
  int B, A, Ap, F, T, nst, hop;
  int scale[kUpMaxStages], tap_off[kUpMaxStages];
};

__global__ void __launch_bounds__(128) upsample_fused_kernel(const UpsampleArgs a) {
  __shared__ float wbuf[2][kUpWin][128];
  __shared__ float cs[kUpMaxA][kUpFrames + 1];
  __shared__ float taps_s[128];
  const int b = blockIdx.y, t0 = blockIdx.x * 128, tid = threadIdx.x;
  const int t = t0 + tid;
  // frames this block can touch: every stage reaches at most s_k samples of its own rate = one frame, plus the floors
  const int margin = a.nst + 1;
  const int f_lo = max(0, t0 / a.hop - margin), f_hi = min(a.F - 1, (t0 + 127) / a.hop + margin);
  const int nf = f_hi - f_lo + 1;
  for (int i = tid; i < a.A * nf; i += 128) {
    const int ch = i / nf, fr = i - ch * nf;
    cs[ch][fr] = a.c[((size_t)b * a.A + ch) * a.F + f_lo + fr];
  }
  int ntaps = 0;
  for (int k = 0; k < a.nst; ++k) ntaps += 2 * a.scale[k] + 1;
  for (int i = tid; i < ntaps; i += 128) taps_s[i] = a.taps[i];
  __syncthreads();
  if (t >= a.T) return;

  // composite weights, top-down: level nst is the output (one entry, weight 1), level 0 the frames
  int lo = t, n = 1, cur = 0;
  wbuf[0][0][tid] = 1.f;
  int len_k = a.T;  // length of the level-k signal
  for (int k = a.nst - 1; k >= 0; --k) {
    const int s = a.scale[k];
    const float* w = taps_s + a.tap_off[k];
    const int len_prev = len_k / s;
    // conv input index u = (entry index) + j - s must lie in [0, len_k); it reads stretched sample u = prev[u / s]
    const int u_min = max(lo - s, 0), u_max = min(lo + n - 1 + s, len_k - 1);
    const int plo = u_min / s, pn = u_max / s - plo + 1;  // pn <= kUpWin is checked at launch
    const int nxt = cur ^ 1;
    for (int i = 0; i < pn; ++i) wbuf[nxt][i][tid] = 0.f;
    for (int i = 0; i < n; ++i) {
      const float wv = wbuf[cur][i][tid];
      const int base = lo + i - s;
      for (int j = 0; j <= 2 * s; ++j) {
        const int u = base + j;
        if (u >= 0 && u < len_k) wbuf[nxt][u / s - plo][tid] += wv * w[j];
      }
    }
    lo = plo; n = pn; cur = nxt; len_k = len_prev;
  }
  // apply to all channels; 8 channels = one 16-byte bf16 store
  float wv[kUpWin];
#pragma unroll
  for (int i = 0; i < kUpWin; ++i) wv[i] = i < n ? wbuf[cur][i][tid] : 0.f;
  const int fo = lo - f_lo;
  const size_t orow = ((size_t)b * a.T + t) * a.Ap;
  for (int c0 = 0; c0 < a.Ap; c0 += 8) {
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float s = 0.f;
      if (c0 + e < a.A) {
#pragma unroll
        for (int i = 0; i < kUpWin; ++i)
          if (i < n) s = fmaf(wv[i], cs[c0 + e][fo + i], s);
      }
      acc[e] = s;
    }
    if (a.out_b) {
      *reinterpret_cast<uint4*>(a.out_b + orow + c0) = make_uint4(ptx::pack_bf16(acc[0], acc[1]), ptx::pack_bf16(acc[2], acc[3]),
                                                                  ptx::pack_bf16(acc[4], acc[5]), ptx::pack_bf16(acc[6], acc[7]));
    }
    if (a.out_f) {
      *reinterpret_cast<float4*>(a.out_f + orow + c0) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      *reinterpret_cast<float4*>(a.out_f + orow + c0 + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
  }
}

__global__ void __launch_bounds__(256) expand1_bf16_kernel(const float* __restrict__ x, long long x_batch_stride,
                                                           const float* __restrict__ w, const float* __restrict__ bias,
                                                           __nv_bfloat16* __restrict__ out, int T, int C) {
  __shared__ float ws[256], bs[256];
  for (int i = threadIdx.x; i < C; i += 256) { ws[i] = w[i]; bs[i] = bias ? bias[i] : 0.f; }
  __syncthreads();
  const int b = blockIdx.y;
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t >= T) return;
  const float xv = x[(size_t)b * x_batch_stride + t];
  __nv_bfloat16* o = out + ((size_t)b * T + t) * C;
  for (int c0 = 0; c0 < C; c0 += 8) {
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = fmaf(ws[c0 + e], xv, bs[c0 + e]);
    *reinterpret_cast<uint4*>(o + c0) = make_uint4(ptx::pack_bf16(v[0], v[1]), ptx::pack_bf16(v[2], v[3]),
                                                   ptx::pack_bf16(v[4], v[5]), ptx::pack_bf16(v[6], v[7]));
  }
}

// ------------------------------------------------------------------------------------------------ source signal
// SignalGenerator.sinusoid (features.py:145-164):
//   rad = (hold(f0) / fs) mod 1 ; sine = vuv * sin(2 pi cumsum(rad)) * sine_amp + randn * (vuv * na + (1 - vuv) * na / 3)
// The reference's cumsum is torch's: on a CPU host it accumulates the fp32 values in fp64 and rounds every prefix to fp32
// (verified bit for bit; a sequential fp32 sum drifts by 1.8 cycles over 720 000 samples).  F0 is constant over the hop
// samples of a frame, so prefix(f * hop + k) = P[f] + (k + 1) * rad_f with P[f] = hop * sum_{f' < f} rad_f' — both exact
// to the last bit or so of an fp64 — which makes the scan a per-track prefix over F frames (kernel 1, one block per track)
// and an independent evaluation per sample (kernel 2).  Everything else follows the reference's fp32 expression order.
__global__ void __launch_bounds__(1024) source_frame_scan_kernel(const double* __restrict__ f0, double* __restrict__ prefix,
                                                                 int F, int hop, float fs) {
  __shared__ double warp_tot[32];
  __shared__ double carry_s;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0) carry_s = 0.0;
  __syncthreads();
  for (int base = 0; base < F; base += 1024) {
    const int f = base + tid;
    double v = 0.0;
    if (f < F) {
      const float x = __fdiv_rn((float)f0[(size_t)b * F + f], fs);
      v = (double)hop * (double)(x - floorf(x));
    }
    double incl = v;                                   // inclusive scan inside the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double n = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += n;
    }
    if (lane == 31) warp_tot[w] = incl;
    __syncthreads();
    if (w == 0) {
      double t = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double n = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += n;
      }
      warp_tot[lane] = t;                              // inclusive totals of the warps
    }
    __syncthreads();
    const double before = carry_s + (w > 0 ? warp_tot[w - 1] : 0.0);
    if (f < F) prefix[(size_t)b * F + f] = before + incl - v;   // exclusive
    __syncthreads();
    if (tid == 1023) carry_s = before + incl;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) source_samples_kernel(const double* __restrict__ f0, const double* __restrict__ prefix,
                                                             const float* __restrict__ noise, float* __restrict__ sine_out,
                                                             long long sine_bstride, float* __restrict__ d_out, int F, int hop,
                                                             float fs, int dense_factor, float sine_amp, float noise_amp) {
  const long long T = (long long)F * hop;
  const int b = blockIdx.y;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const int f = (int)(t / hop), k = (int)(t - (long long)f * hop);
  const double f064 = f0[(size_t)b * F + f];
  const float f032 = (float)f064;                      // torch.FloatTensor(f0)
  if (sine_out) {
    const float vuv = f032 > 0.f ? 1.f : 0.f;
    const float x = __fdiv_rn(f032, fs);
    const float rad = x - floorf(x);                   // remainder(x, 1), exact
    const float c32 = (float)(prefix[(size_t)b * F + f] + (double)(k + 1) * (double)rad);
    const float phase = __fmul_rn(__fmul_rn(c32, 2.0f), 3.14159274101257324f);   // cumsum * 2 * np.pi, fp32 scalars
    float v = __fmul_rn(__fmul_rn(vuv, sinf(phase)), sine_amp);
    if (noise) {
      const float amp = __fadd_rn(__fmul_rn(vuv, noise_amp), __fdiv_rn(__fmul_rn(1.0f - vuv, noise_amp), 3.0f));
      v = __fadd_rn(v, __fmul_rn(amp, noise[(size_t)b * T + t]));
    }
    sine_out[(size_t)b * sine_bstride + t] = v;
  }
  if (d_out) {
    // dilated_factor (features.py:56-75), float64 like the reference's numpy expression: fs / f0 / dense_factor,
    // unvoiced frames first set to fs / dense_factor
    const double fsd = (double)fs;
    const double f0d = f064 == 0.0 ? fsd / (double)dense_factor : f064;
    d_out[(size_t)b * T + t] = (float)(fsd / f0d / (double)dense_factor);
  }
}

// Frame-rate aux projection, operand A of the block kernel (usfgan_fr.cuh): the upsampler's impulse responses, one
// thread per sample, re-ordered into the 16-frame window of the sample's tile.
__global__ void usfgan_aux_weights_kernel(const float* __restrict__ imp, __nv_bfloat16* __restrict__ u, int T, int Tpad,
                                          int hop, int reach) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= Tpad) return;
  uint32_t pk[8];
  const int fb = usfgan_frame_base(t & ~127, reach, hop);
#pragma unroll
  for (int k = 0; k < 16; k += 2) {
    float v0 = 0.f, v1 = 0.f;
    if (t < T) {
      v0 = imp[(size_t)((fb + k) & 15) * T + t];
      v1 = imp[(size_t)((fb + k + 1) & 15) * T + t];
    }
    pk[k >> 1] = ptx::pack_bf16(v0, v1);
  }
  uint4* dst = reinterpret_cast<uint4*>(u + (size_t)t * 16);
  dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
}

// The aux upsampler as U . c at the blocks' frame window (usfgan_fr.cuh): out[b][t][ch] = sum_k imp[(fb + k) mod 16][t] *
// cin[b][ch][fb + k], fb = usfgan_frame_base(tile of t).  Exactly what upsample.py:61-128 computes (the stages are linear and
// channel-wise; imp holds their response, boundaries included, to unit impulses), in one pass at the write rate of the
// sample-rate tensor: the periodicity estimator is the only consumer of sample-rate aux features left once the blocks
// take their projection at frame rate, and the staged-arithmetic kernel above needs 1.1 ms for 6 x 720 000 samples.
__global__ void __launch_bounds__(256)
upsample_frames_kernel(const float* __restrict__ imp, const float* __restrict__ cin, __nv_bfloat16* __restrict__ out, int A,
                       int Ap, int Tf, int T, int hop, int reach) {
  extern __shared__ __align__(16) float up_s[];
  float* c_s = up_s;                                                       // [16 frames][Ap]
  float* u_s = up_s + 16 * Ap;                                             // [128 samples][17] (padded rows)
  uint4* o_s = reinterpret_cast<uint4*>(u_s + 128 * 17 + 3);               // [128 samples][Ap / 8] output staging
  o_s = reinterpret_cast<uint4*>(reinterpret_cast<uintptr_t>(o_s) & ~uintptr_t(15));
  const int b = blockIdx.y, t0 = blockIdx.x * 128;
  const int fb = usfgan_frame_base(t0, reach, hop);
  for (int i = threadIdx.x; i < 16 * Ap; i += 256) {
    const int ch = i >> 4, k = i & 15, f = fb + k;
    c_s[k * Ap + ch] = (ch < A && f >= 0 && f < Tf) ? cin[((size_t)b * A + ch) * Tf + f] : 0.f;
  }
  for (int i = threadIdx.x; i < 128 * 16; i += 256) {
    const int row = i & 127, k = i >> 7, t = t0 + row;
    u_s[row * 17 + k] = t < T ? imp[(size_t)((fb + k) & 15) * T + t] : 0.f;
  }
  __syncthreads();
  // thread = (sample, half of the channel groups): its 16 weights stay in registers, the frame window is read as warp-wide
  // broadcasts (all lanes of a warp walk the same channel groups), 8 FMAs per 2 shared-memory loads
  const int G = Ap >> 3, row = threadIdx.x & 127, half = threadIdx.x >> 7;
  const int g0 = half ? (G + 1) / 2 : 0, g1 = half ? G : (G + 1) / 2;
  float u[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) u[k] = u_s[row * 17 + k];
  for (int cg = g0; cg < g1; ++cg) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float4 c0 = *reinterpret_cast<const float4*>(c_s + k * Ap + cg * 8);
      const float4 c1 = *reinterpret_cast<const float4*>(c_s + k * Ap + cg * 8 + 4);
      acc[0] = fmaf(u[k], c0.x, acc[0]); acc[1] = fmaf(u[k], c0.y, acc[1]); acc[2] = fmaf(u[k], c0.z, acc[2]);
      acc[3] = fmaf(u[k], c0.w, acc[3]); acc[4] = fmaf(u[k], c1.x, acc[4]); acc[5] = fmaf(u[k], c1.y, acc[5]);
      acc[6] = fmaf(u[k], c1.z, acc[6]); acc[7] = fmaf(u[k], c1.w, acc[7]);
    }
    o_s[row * G + cg] = make_uint4(ptx::pack_bf16(acc[0], acc[1]), ptx::pack_bf16(acc[2], acc[3]),
                                   ptx::pack_bf16(acc[4], acc[5]), ptx::pack_bf16(acc[6], acc[7]));
  }
  __syncthreads();
  // the tile's rows are contiguous in the output: coalesced 16-byte stores
  const int rows = min(128, T - t0);
  uint4* dst = reinterpret_cast<uint4*>(out + ((size_t)b * T + t0) * Ap);
  for (int i = threadIdx.x; i < rows * G; i += 256) dst[i] = o_s[i];
}

}  // namespace svsk

using namespace svsk;

extern "C" int svsk_upsample_fused(const float* c, const float* taps, const int32_t* scales, int n_stages, int B, int A, int F,
                                   void* out_bf16, float* out_f32, int Ap, void* stream) {
  SVSK_REQUIRE(c && taps && scales && (out_bf16 || out_f32), SVSK_E_ARG, "upsample_fused: null");
  SVSK_REQUIRE(n_stages >= 1 && n_stages <= kUpMaxStages, SVSK_E_ARG, "upsample_fused: %d stages (1..%d)", n_stages, kUpMaxStages);
  SVSK_REQUIRE(B > 0 && B <= 65535 && A >= 1 && A <= kUpMaxA && F >= 1 && Ap >= A && Ap % 8 == 0, SVSK_E_ARG,
               "upsample_fused: bad shape B=%d A=%d F=%d Ap=%d", B, A, F, Ap);
  UpsampleArgs a;
  a.c = c; a.taps = taps; a.out_b = (__nv_bfloat16*)out_bf16; a.out_f = out_f32;
  a.B = B; a.A = A; a.Ap = Ap; a.F = F; a.nst = n_stages;
  long long hop = 1;
  int off = 0, ntaps = 0;
  for (int k = 0; k < n_stages; ++k) {
    SVSK_REQUIRE(scales[k] >= 2 && scales[k] <= 16, SVSK_E_ARG, "upsample_fused: scale %d (2..16)", scales[k]);
    a.scale[k] = scales[k];
    a.tap_off[k] = off;
    off += 2 * scales[k] + 1;
    ntaps += 2 * scales[k] + 1;
    hop *= scales[k];
    // a window of n entries reads (n - 1 + 2 s) / s + 2 entries of the level below
    SVSK_REQUIRE((kUpWin - 1 + 2 * scales[k]) / scales[k] + 2 <= kUpWin, SVSK_E_ARG,
                 "upsample_fused: scale %d needs a wider composite window than %d (scales >= 2 work)", scales[k], kUpWin);
  }
  for (int k = n_stages; k < kUpMaxStages; ++k) { a.scale[k] = 1; a.tap_off[k] = 0; }
  SVSK_REQUIRE(ntaps <= 128, SVSK_E_ARG, "upsample_fused: %d taps (at most 128)", ntaps);
  SVSK_REQUIRE(hop * F < (1ll << 31), SVSK_E_ARG, "upsample_fused: output too long");
  a.hop = (int)hop;
  a.T = (int)(hop * F);
  SVSK_REQUIRE(127 / a.hop + 2 + 2 * (n_stages + 1) <= kUpFrames, SVSK_E_ARG,
               "upsample_fused: hop %d too small for the %d-frame staging buffer", a.hop, kUpFrames);
  SVSK_REQUIRE(out_bf16 == nullptr || ((uintptr_t)out_bf16 % 16) == 0, SVSK_E_ALIGN, "upsample_fused: out_bf16 alignment");
  SVSK_REQUIRE(out_f32 == nullptr || ((uintptr_t)out_f32 % 16) == 0, SVSK_E_ALIGN, "upsample_fused: out_f32 alignment");
  upsample_fused_kernel<<<dim3(ceil_div(a.T, 128), B), 128, 0, as_stream(stream)>>>(a);
  return check_launch("upsample_fused");
}

extern "C" int svsk_expand1_bf16(const float* x, long long x_batch_stride, const float* w, const float* bias, void* out,
                                 int B, int T, int C, void* stream) {
  SVSK_REQUIRE(x && w && out, SVSK_E_ARG, "expand1_bf16: null");
  SVSK_REQUIRE(B > 0 && B <= 65535 && T > 0 && C >= 8 && C <= 256 && C % 8 == 0, SVSK_E_ARG, "expand1_bf16: bad shape");
  SVSK_REQUIRE(((uintptr_t)out % 16) == 0, SVSK_E_ALIGN, "expand1_bf16: out alignment");
  expand1_bf16_kernel<<<dim3(ceil_div(T, 256), B), 256, 0, as_stream(stream)>>>(x, x_batch_stride, w, bias, (__nv_bfloat16*)out,
                                                                                T, C);
  return check_launch("expand1_bf16");
}

extern "C" int svsk_usfgan_source(const double* f0, const float* noise, float* sine_out, long long sine_batch_stride, float* d_out,
                                  double* scratch, int B, int F, int hop, int sample_rate, int dense_factor, float sine_amp,
                                  float noise_amp, void* stream) {
  SVSK_REQUIRE(f0 && scratch && (sine_out || d_out), SVSK_E_ARG, "usfgan_source: null");
  SVSK_REQUIRE(B > 0 && B <= 65535 && F > 0 && hop > 0 && sample_rate > 0 && dense_factor > 0, SVSK_E_ARG,
               "usfgan_source: B=%d F=%d hop=%d fs=%d dense_factor=%d", B, F, hop, sample_rate, dense_factor);
  SVSK_REQUIRE(!sine_out || sine_batch_stride >= (long long)F * hop, SVSK_E_ARG, "usfgan_source: sine batch stride");
  if (sine_out) {
    source_frame_scan_kernel<<<B, 1024, 0, as_stream(stream)>>>(f0, scratch, F, hop, (float)sample_rate);
    int rc = check_launch("usfgan_source (frame scan)");
    if (rc) return rc;
  }
  const long long T = (long long)F * hop;
  dim3 grid((unsigned)((T + 255) / 256), (unsigned)B);
  source_samples_kernel<<<grid, 256, 0, as_stream(stream)>>>(f0, scratch, noise_amp > 0.f ? noise : nullptr, sine_out,
                                                             sine_batch_stride, d_out, F, hop, (float)sample_rate, dense_factor,
                                                             sine_amp, noise_amp);
  return check_launch("usfgan_source");
}

extern "C" int svsk_usfgan_aux_weights(const float* imp, void* u, int T, int hop, int reach, void* stream) {
  SVSK_REQUIRE(imp && u, SVSK_E_ARG, "usfgan_aux_weights: null tensor");
  SVSK_REQUIRE(T > 0 && hop >= 1 && reach >= 0, SVSK_E_ARG, "usfgan_aux_weights: T=%d hop=%d reach=%d", T, hop, reach);
  SVSK_REQUIRE((reinterpret_cast<uintptr_t>(u) & 15) == 0, SVSK_E_ALIGN, "usfgan_aux_weights: u must be 16-byte aligned");
  // 16 impulse channels tell frames apart only if no sample hears two frames that are 16 apart
  SVSK_REQUIRE(2ll * reach < 15ll * hop, SVSK_E_ARG, "usfgan_aux_weights: reach %d too long for hop %d", reach, hop);
  const int Tpad = (T + 127) / 128 * 128;
  usfgan_aux_weights_kernel<<<(Tpad + 255) / 256, 256, 0, as_stream(stream)>>>(imp, (__nv_bfloat16*)u, T, Tpad, hop, reach);
  return check_launch("usfgan_aux_weights");
}

extern "C" int svsk_usfgan_frame_base(int t0, int reach, int hop) { return usfgan_frame_base(t0, reach, hop); }

extern "C" int svsk_upsample_frames_bf16(const float* imp, const float* cin, void* out, int B, int A, int Ap, int Tf, int T, int hop,
                                         int reach, void* stream) {
  SVSK_REQUIRE(imp && cin && out, SVSK_E_ARG, "upsample_frames_bf16: null tensor");
  SVSK_REQUIRE(B > 0 && B <= 65535 && A >= 1 && Ap >= A && Ap % 8 == 0 && Ap <= 112 && Tf > 0 && T > 0, SVSK_E_ARG,
               "upsample_frames_bf16: bad shape (Ap <= 112) B=%d A=%d Ap=%d Tf=%d T=%d", B, A, Ap, Tf, T);
  SVSK_REQUIRE(hop >= 1 && reach >= 0 && (127 + 2 * (long long)reach) / hop <= 7 && 2ll * reach < 15ll * hop, SVSK_E_ARG,
               "upsample_frames_bf16: a 128-sample tile must reach at most 8 frames (hop=%d reach=%d)", hop, reach);
  SVSK_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, SVSK_E_ALIGN, "upsample_frames_bf16: out must be 16-byte aligned");
  const int smem = (16 * Ap + 128 * 17 + 8) * (int)sizeof(float) + 128 * Ap * 2;
  upsample_frames_kernel<<<dim3((T + 127) / 128, B), 256, smem, as_stream(stream)>>>(imp, cin, (__nv_bfloat16*)out, A, Ap, Tf, T,
                                                                                   hop, reach);
  return check_launch("upsample_frames_bf16");
}
