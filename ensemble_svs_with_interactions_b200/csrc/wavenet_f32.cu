// Fused WaveNet residual block in fp32: replaces ResSkipBlock._forward (nnsvs/wavenet/modules.py:88-122)
// with ONE launch per layer —
//   y = causal_conv_k_dil(x) + conv1x1c(c) + b          (taps at t - (k-1-j) d, zeros before the start)
//   z = tanh(y[:G/2]) * sigmoid(y[G/2:])
//   skips (+)= conv1x1_skip(z) + b_skip ;  x_out = conv1x1_out(z) + b_out + x      (no sqrt(1/2): modules.py:119-121)
// instead of five (conv, conditioning conv, gate, skip conv, out conv).  The acoustic WaveNet is tiny (71 168 MAC per
// frame and layer, B x T = 400 frames in the reference's test): it is launch-latency bound, so the kernel stays on the
// CUDA cores in exact fp32 (this is the reference's own arithmetic; tolerance 1e-6, not a bf16 path) and spends its
// effort on having a frame tile's whole block in one CTA: inputs (all taps + conditioning) staged once in shared
// memory, the gate pre-activations and z never leave it.  Weights are pre-transposed ([K][out], folded weight norm) so
// that a warp reads 32 consecutive output channels of one k per load.
#include "svsk_common.cuh"

namespace svsk {

constexpr int kWnTT = 16;  // frames per CTA
constexpr int kWnKC = 32;  // weight rows staged per chunk

// acc[i] += sum_k w[k][col] * in[k][tg * 8 + i] for one (column, half tile) item per thread, K rows of weights streamed
// through shared memory in chunks of kWnKC rows: all threads fetch the NEXT chunk into registers (coalesced float4, many
// loads in flight) before they compute on the current one, so the L2 latency of a chunk hides behind 32 k-steps of FMAs.
// ncols is a multiple of 4; w_s holds kWnKC * ncols floats.  Must be called by all 256 threads.
__device__ __forceinline__ void wn_gemm_tile(const float* __restrict__ w, int K, int ncols, const float* in_s, float* w_s,
                                             int n_items, float (&acc)[2][8]) {
  const int vec_per_chunk = kWnKC * ncols / 4;           // float4s per chunk
  constexpr int kMaxV = 8;                               // per thread: chunk of 32 x 256 floats at most
  float4 nxt[kMaxV];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int v = 0; v < kMaxV; ++v) {
      const int i = threadIdx.x + v * 256;
      nxt[v] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < vec_per_chunk) {
        const int r = (i * 4) / ncols;
        if (k0 + r < K) nxt[v] = *reinterpret_cast<const float4*>(w + (size_t)k0 * ncols + (size_t)i * 4);
      }
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < K; k0 += kWnKC) {
    __syncthreads();                                      // the previous chunk has been consumed
#pragma unroll
    for (int v = 0; v < kMaxV; ++v) {
      const int i = threadIdx.x + v * 256;
      if (i < vec_per_chunk) *reinterpret_cast<float4*>(w_s + (size_t)i * 4) = nxt[v];
    }
    __syncthreads();
    if (k0 + kWnKC < K) fetch(k0 + kWnKC);
    const int kn = min(kWnKC, K - k0);
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int item = threadIdx.x + it * 256;
      if (item >= n_items) break;
      const int col = item % ncols, tg = item / ncols;
      const float* xs = in_s + (size_t)k0 * kWnTT + tg * 8;
#pragma unroll 8
      for (int k = 0; k < kn; ++k) {
        const float wv = w_s[k * ncols + col];
        const float4 a = *reinterpret_cast<const float4*>(xs + k * kWnTT);
        const float4 c4 = *reinterpret_cast<const float4*>(xs + k * kWnTT + 4);
        acc[it][0] = fmaf(wv, a.x, acc[it][0]); acc[it][1] = fmaf(wv, a.y, acc[it][1]);
        acc[it][2] = fmaf(wv, a.z, acc[it][2]); acc[it][3] = fmaf(wv, a.w, acc[it][3]);
        acc[it][4] = fmaf(wv, c4.x, acc[it][4]); acc[it][5] = fmaf(wv, c4.y, acc[it][5]);
        acc[it][6] = fmaf(wv, c4.z, acc[it][6]); acc[it][7] = fmaf(wv, c4.w, acc[it][7]);
      }
    }
  }
}

__global__ void __launch_bounds__(256) wavenet_block_f32_kernel(const svsk_wavenet_block_params p) {
  extern __shared__ float smem[];
  const int K1 = p.ksize * p.R + p.Cc, Gh = p.G / 2, O2 = p.S + p.R;
  float* in_s = smem;                                        // [K1][kWnTT]
  float* y_s = in_s + K1 * kWnTT;                            // [G][kWnTT]
  float* z_s = y_s + p.G * kWnTT;                            // [G/2][kWnTT]
  float* w_s = z_s + Gh * kWnTT;                             // [kWnKC][max(G, O2)]
  const int b = blockIdx.y, t0 = blockIdx.x * kWnTT;
  const float* xb = p.x + (size_t)b * p.R * p.T;
  const float* cb = p.c + (size_t)b * p.Cc * p.T;
  // ---- stage the inputs: row j * R + ci = x[ci][t - (k-1-j) d], rows k R .. = c
  for (int i = threadIdx.x; i < K1 * kWnTT; i += 256) {
    const int r = i / kWnTT, tt = i % kWnTT;
    const int t = t0 + tt;
    float v = 0.f;
    if (t < p.T) {
      if (r < p.ksize * p.R) {
        const int j = r / p.R, ci = r - j * p.R;
        const int u = t - (p.ksize - 1 - j) * p.dilation;
        if (u >= 0) v = xb[(size_t)ci * p.T + u];
      } else {
        v = cb[(size_t)(r - p.ksize * p.R) * p.T + t];
      }
    }
    in_s[i] = v;
  }
  // ---- gate pre-activations: item = (output channel, half tile of 8 frames), up to two items per thread
  {
    const int n_items = p.G * (kWnTT / 8);
    float acc[2][8];
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int item = threadIdx.x + it * 256;
      const float bias = (p.b1 && item < n_items) ? p.b1[item % p.G] : 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[it][i] = bias;
    }
    wn_gemm_tile(p.w1t, K1, p.G, in_s, w_s, n_items, acc);   // (its first __syncthreads also publishes in_s)
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int item = threadIdx.x + it * 256;
      if (item >= n_items) break;
      const int co = item % p.G, tg = item / p.G;
#pragma unroll
      for (int i = 0; i < 8; ++i) y_s[co * kWnTT + tg * 8 + i] = acc[it][i];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Gh * kWnTT; i += 256) {
    const float a = y_s[i], g = y_s[Gh * kWnTT + i];
    z_s[i] = tanhf(a) * (1.f / (1.f + expf(-g)));
  }
  // ---- skip (columns [0, S)) and out (columns [S, S + R)) projections of z
  {
    const int n_items = O2 * (kWnTT / 8);
    float acc[2][8];
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int item = threadIdx.x + it * 256;
      const float bias = item < n_items ? p.b2[item % O2] : 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[it][i] = bias;
    }
    wn_gemm_tile(p.w2t, Gh, O2, z_s, w_s, n_items, acc);     // (its first __syncthreads also publishes z_s)
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int item = threadIdx.x + it * 256;
      if (item >= n_items) break;
      const int o = item % O2, tg = item / O2;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int t = t0 + tg * 8 + i;
        if (t >= p.T) break;
        if (o < p.S) {
          float* sp = p.skips + ((size_t)b * p.S + o) * p.T + t;
          *sp = p.first ? acc[it][i] : *sp + acc[it][i];
        } else {
          const size_t off = ((size_t)b * p.R + (o - p.S)) * p.T + t;
          p.x_out[off] = acc[it][i] + p.x[off];
        }
      }
    }
  }
}

// w1t[(j * R + ci) * G + co] = conv.weight[co][ci][j];  w1t[(k R + cc) * G + co] = conv1x1c.weight[co][cc]
// w2t[k * (S + R) + o] = o < S ? skip.weight[o][k] : out.weight[o - S][k]
__global__ void wavenet_pack_kernel(const float* __restrict__ wconv, const float* __restrict__ wc, const float* __restrict__ wskip,
                                    const float* __restrict__ wout, float* __restrict__ w1t, float* __restrict__ w2t, int R, int G,
                                    int S, int Cc, int ksize) {
  const int K1 = ksize * R + Cc, Gh = G / 2, O2 = S + R;
  const int n1 = K1 * G, n2 = Gh * O2;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n1 + n2; i += gridDim.x * blockDim.x) {
    if (i < n1) {
      const int r = i / G, co = i % G;
      if (r < ksize * R) { const int j = r / R, ci = r - j * R; w1t[i] = wconv[((size_t)co * R + ci) * ksize + j]; }
      else w1t[i] = wc[(size_t)co * Cc + (r - ksize * R)];
    } else {
      const int q = i - n1, k = q / O2, o = q % O2;
      w2t[q] = o < S ? wskip[(size_t)o * Gh + k] : wout[(size_t)(o - S) * Gh + k];
    }
  }
}

}  // namespace svsk

using namespace svsk;

extern "C" int svsk_wavenet_pack_f32(const float* wconv, const float* wc, const float* wskip, const float* wout, float* w1t,
                                     float* w2t, int R, int G, int S, int Cc, int ksize, void* stream) {
  SVSK_REQUIRE(wconv && wc && wskip && wout && w1t && w2t, SVSK_E_ARG, "wavenet_pack_f32: null");
  SVSK_REQUIRE(R > 0 && G > 0 && G % 2 == 0 && S > 0 && Cc > 0 && ksize >= 1, SVSK_E_ARG, "wavenet_pack_f32: bad sizes");
  const int n = (ksize * R + Cc) * G + (G / 2) * (S + R);
  wavenet_pack_kernel<<<(n + 255) / 256, 256, 0, as_stream(stream)>>>(wconv, wc, wskip, wout, w1t, w2t, R, G, S, Cc, ksize);
  return check_launch("wavenet_pack_f32");
}

extern "C" int svsk_wavenet_block_f32(const svsk_wavenet_block_params* pp, void* stream) {
  SVSK_REQUIRE(pp != nullptr, SVSK_E_ARG, "wavenet_block_f32: null params");
  const svsk_wavenet_block_params& p = *pp;
  SVSK_REQUIRE(p.x && p.c && p.w1t && p.w2t && p.b2 && p.x_out && p.skips, SVSK_E_ARG, "wavenet_block_f32: null tensor");
  SVSK_REQUIRE(p.B > 0 && p.B <= 65535 && p.T > 0 && p.R > 0 && p.G > 0 && p.G % 2 == 0 && p.S > 0 && p.Cc > 0 && p.ksize >= 1 &&
                   p.dilation >= 1, SVSK_E_ARG, "wavenet_block_f32: bad sizes");
  SVSK_REQUIRE(p.x != p.x_out, SVSK_E_ARG, "wavenet_block_f32: x and x_out must differ (later frames read earlier taps)");
  SVSK_REQUIRE(p.G % 4 == 0 && (p.S + p.R) % 4 == 0 && p.G <= 256 && p.S + p.R <= 256, SVSK_E_ARG,
               "wavenet_block_f32: gate channels G=%d and S + R=%d must be multiples of 4, at most 256", p.G, p.S + p.R);
  const int wcols = p.G > p.S + p.R ? p.G : p.S + p.R;
  const size_t smem = (((size_t)(p.ksize * p.R + p.Cc) + p.G + p.G / 2) * kWnTT + (size_t)kWnKC * wcols) * sizeof(float);
  SVSK_REQUIRE(smem <= 200 * 1024, SVSK_E_ARG, "wavenet_block_f32: %zu bytes of shared memory (k R + cin + 1.5 G too large)", smem);
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(wavenet_block_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail((int)e, "wavenet_block_f32: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    smem_set = smem;
  }
  dim3 grid((unsigned)ceil_div(p.T, kWnTT), (unsigned)p.B);
  wavenet_block_f32_kernel<<<grid, 256, smem, as_stream(stream)>>>(p);
  return check_launch("wavenet_block_f32");
}
