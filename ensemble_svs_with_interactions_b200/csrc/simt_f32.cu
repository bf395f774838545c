// fp32 exact path: CUDA-core kernels in the reference's NCT layout.
// These are the "fp32" precision mode of the drop-in modules and the small glue (step-embedding MLP,
// DDPM update, pitch-dependent tap indices, aux upsampling) of the bf16 tensor-core mode.
#include "svsk_common.cuh"
#include <cmath>

namespace svsk {

// ----------------------------------------------------------------------------------------------
// conv1d as a tiled sum of taps
// ----------------------------------------------------------------------------------------------
constexpr int kTT = 64;    // time tile
constexpr int kCOT = 64;   // out-channel tile
constexpr int kCIT = 8;    // in-channel chunk
constexpr int kMaxK = 8;   // max kernel size

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case SVSK_ACT_RELU: return fmaxf(v, 0.f);
    case SVSK_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    case SVSK_ACT_MISH: {
      // x * tanh(softplus(x)); softplus with torch's threshold of 20 (F.softplus default)
      float sp = v > 20.f ? v : log1pf(expf(v));
      return v * tanhf(sp);
    }
    default: return v;
  }
}

// source index for output t and tap j; returns -1 for "contributes zero"
__device__ __forceinline__ int tap_source(int t, int j, const svsk_conv1d_f32_params& p, int T_in, int b) {
  if (p.pad_mode == SVSK_PAD_INDEXED) {
    if (j == 1) return t;
    const int32_t* idx = (j == 0) ? p.idx_past : p.idx_future;
    return idx[(size_t)b * p.T + t];
  }
  if (p.pad_mode == SVSK_PAD_VALID) return t + j * p.dilation;
  int u = t + (j - p.tap_origin) * p.dilation;
  if (p.pad_mode == SVSK_PAD_ZEROS) return (u < 0 || u >= T_in) ? -1 : u;
  if (p.pad_mode == SVSK_PAD_REFLECT) {
    if (u < 0) u = -u;
    if (u >= T_in) u = 2 * (T_in - 1) - u;
    return (u < 0 || u >= T_in) ? -1 : u;  // only when dilation >= T (rejected on the host)
  }
  // replicate
  return u < 0 ? 0 : (u >= T_in ? T_in - 1 : u);
}

__global__ void __launch_bounds__(256) conv1d_f32_kernel(const svsk_conv1d_f32_params p, int T_in) {
  __shared__ float Ws[kCIT * kMaxK][kCOT];
  __shared__ float Xs[kCIT * kMaxK][kTT];
  __shared__ int Src[kMaxK][kTT];

  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int t0 = blockIdx.x * kTT, co0 = blockIdx.y * kCOT, b = blockIdx.z;
  const int ks = p.ksize;

  for (int i = threadIdx.x; i < ks * kTT; i += 256) {
    int j = i / kTT, tt = i % kTT;
    int t = t0 + tt;
    Src[j][tt] = (t < p.T) ? tap_source(t, j, p, T_in, b) : -1;
  }
  __syncthreads();

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const float* xb = p.x + (size_t)b * p.Cin * T_in;
  for (int ci0 = 0; ci0 < p.Cin; ci0 += kCIT) {
    const int rows = min(kCIT, p.Cin - ci0) * ks;
    // weights: Ws[ci*ks + j][co]
    for (int i = threadIdx.x; i < kCIT * ks * kCOT; i += 256) {
      int r = i / kCOT, co = i % kCOT;
      float v = 0.f;
      if (r < rows && co0 + co < p.Cout) {
        int ci = ci0 + r / ks, j = r % ks;
        v = p.w[((size_t)(co0 + co) * p.Cin + ci) * ks + j];
      }
      Ws[r][co] = v;
    }
    // inputs: Xs[ci*ks + j][t]
    for (int i = threadIdx.x; i < kCIT * ks * kTT; i += 256) {
      int r = i / kTT, tt = i % kTT;
      float v = 0.f;
      if (r < rows) {
        int ci = ci0 + r / ks, j = r % ks;
        int u = Src[j][tt];
        if (u >= 0) {
          v = xb[(size_t)ci * T_in + u];
          if (p.in_relu) v = fmaxf(v, 0.f);
          if (p.in_bias) v += p.in_bias[(size_t)b * p.Cin + ci];
        }
      }
      Xs[r][tt] = v;
    }
    __syncthreads();
    for (int r = 0; r < rows; ++r) {
      float a[4], x[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = Ws[r][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) x[j] = Xs[r][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], x[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int co = co0 + ty + 16 * i;
    if (co >= p.Cout) continue;
    float bias = p.bias ? p.bias[co] : 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int t = t0 + tx + 16 * j;
      if (t >= p.T) continue;
      size_t o = ((size_t)b * p.Cout + co) * p.T + t;
      float v = acc[i][j] + bias;
      if (p.residual) v += p.residual[o];
      v *= p.out_scale;
      v = apply_act(v, p.act);
      if (p.accumulate) v += p.y[o];
      p.y[o] = v;
    }
  }
}

// ----------------------------------------------------------------------------------------------
// elementwise kernels
// ----------------------------------------------------------------------------------------------
__global__ void gated_act_kernel(const float* __restrict__ y, float* __restrict__ z, int H, int T, int order,
                                 size_t total) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  size_t b = i / ((size_t)H * T);
  size_t r = i - b * (size_t)H * T;
  float a = y[b * 2 * H * T + r];
  float g = y[b * 2 * H * T + (size_t)H * T + r];
  float sa = 1.f / (1.f + expf(-(order == SVSK_GATE_SIGMOID_TANH ? a : g)));
  float th = tanhf(order == SVSK_GATE_SIGMOID_TANH ? g : a);
  z[i] = sa * th;
}

__global__ void diffnet_residual_skip_kernel(const float* __restrict__ o, float* __restrict__ x,
                                             float* __restrict__ skip, int C, int T, int init_skip, size_t total) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  size_t b = i / ((size_t)C * T);
  size_t r = i - b * (size_t)C * T;
  const float* ob = o + b * 2 * C * T;
  // (x + residual) / sqrt(2.0): the reference divides by the double constant (denoiser.py:66)
  x[i] = (x[i] + ob[r]) / 1.4142135623730951f;
  float s = ob[(size_t)C * T + r];
  skip[i] = init_skip ? s : skip[i] + s;
}

__global__ void scale_act_kernel(const float* __restrict__ x, float* __restrict__ y, size_t n, float alpha, int act) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = alpha * apply_act(x[i], act);
}

__global__ void sinusoidal_embedding_kernel(const int64_t* __restrict__ t, float* __restrict__ out, int B, int dim,
                                            float neg_scale) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int half = dim / 2;
  if (i >= B * half) return;
  int b = i / half, k = i % half;
  float f = expf((float)k * neg_scale);
  float arg = (float)t[b] * f;
  out[(size_t)b * dim + k] = sinf(arg);
  out[(size_t)b * dim + half + k] = cosf(arg);
}

struct DdpmTables {
  const float *sra, *srm1, *c1, *c2, *plv;
};

__global__ void ddpm_update_kernel(const float* __restrict__ x, const float* __restrict__ eps,
                                   const float* __restrict__ z, float* __restrict__ out,
                                   const int64_t* __restrict__ t, DdpmTables tab, size_t per_batch, int clip) {
  int b = blockIdx.y;
  int64_t tb = t[b];
  float a = tab.sra[tb], bb = tab.srm1[tb], c1 = tab.c1[tb], c2 = tab.c2[tb];
  float sigma = tb == 0 ? 0.f : expf(0.5f * tab.plv[tb]);
  size_t base = (size_t)b * per_batch;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_batch; i += (size_t)gridDim.x * blockDim.x) {
    float xv = x[base + i];
    float x0 = a * xv - bb * eps[base + i];
    if (clip) x0 = fminf(fmaxf(x0, -1.f), 1.f);
    float mean = c1 * x0 + c2 * xv;
    out[base + i] = mean + sigma * z[base + i];
  }
}

__global__ void q_sample_kernel(const float* __restrict__ x0, const float* __restrict__ noise, float* __restrict__ out,
                                const int64_t* __restrict__ t, const float* __restrict__ sac,
                                const float* __restrict__ somac, size_t per_batch) {
  int b = blockIdx.y;
  float a = sac[t[b]], s = somac[t[b]];
  size_t base = (size_t)b * per_batch;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_batch; i += (size_t)gridDim.x * blockDim.x)
    out[base + i] = a * x0[base + i] + s * noise[base + i];
}

__global__ void plms_transfer_kernel(const float* __restrict__ x, const float* __restrict__ nz, float* __restrict__ out,
                                     const int64_t* __restrict__ t, int interval, const float* __restrict__ ac,
                                     size_t per_batch) {
  int b = blockIdx.y;
  int64_t tb = t[b];
  int64_t tp = tb - interval < 0 ? 0 : tb - interval;
  float a_t = ac[tb], a_prev = ac[tp];
  float a_t_sq = sqrtf(a_t), a_prev_sq = sqrtf(a_prev);
  float cx = 1.f / (a_t_sq * (a_t_sq + a_prev_sq));
  float cn = 1.f / (a_t_sq * (sqrtf((1.f - a_prev) * a_t) + sqrtf((1.f - a_t) * a_prev)));
  float d = a_prev - a_t;
  size_t base = (size_t)b * per_batch;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_batch; i += (size_t)gridDim.x * blockDim.x) {
    float xv = x[base + i];
    out[base + i] = xv + d * (cx * xv - cn * nz[base + i]);
  }
}

struct LinComb {
  const float* in[4];
  float coef[4];
  int n_in;
};
__global__ void lincomb_kernel(LinComb lc, float* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = 0.f;
  for (int k = 0; k < lc.n_in; ++k) v += lc.coef[k] * lc.in[k][i];
  out[i] = v;
}

__global__ void pd_index_kernel(const float* __restrict__ d, int32_t* __restrict__ ip, int32_t* __restrict__ ifu, int T,
                                int dilation, size_t total) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int t = (int)(i % T);
  // same fp32 operation order as the reference (index.py:27-47): product, then sum with the float index, then
  // round-half-even; intrinsics keep the compiler from contracting into an FMA (different rounding).
  float prod = __fmul_rn(d[i], (float)dilation);
  float sp = __fadd_rn(-prod, (float)(t - T));
  float sf = __fadd_rn(prod, (float)t);
  long long p = (long long)rintf(sp) + T;
  long long f = (long long)rintf(sf);
  ip[i] = p < 0 ? -1 : (p >= T ? -1 : (int32_t)p);
  ifu[i] = f >= T ? -1 : (f < 0 ? -1 : (int32_t)f);
}

__global__ void upsample_smooth_kernel(const float* __restrict__ in, const float* __restrict__ taps,
                                       float* __restrict__ out, int Tin, int scale, size_t total) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int Tout = Tin * scale;
  size_t r = i / Tout;
  int t = (int)(i - r * Tout);
  const float* row = in + r * Tin;
  float acc = 0.f;
  for (int j = 0; j <= 2 * scale; ++j) {
    int u = t + j - scale;
    if (u >= 0 && u < Tout) acc = fmaf(taps[j], row[u / scale], acc);
  }
  out[i] = acc;
}

__global__ void periodic_mix_kernel(const float* __restrict__ a, const float* __restrict__ h,
                                    const float* __restrict__ n, float* __restrict__ s, float* __restrict__ h2,
                                    float* __restrict__ n2, size_t cnt) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cnt) return;
  float av = a[i];
  float hv = av * h[i];
  float nv = (1.0f - av) * n[i];
  if (h2) h2[i] = hv;
  if (n2) n2[i] = nv;
  s[i] = hv + nv;
}

// y[b][co] = act(bias[co] + sum_ci w[co][ci] x[b][ci]): one warp per output channel keeps its weight row in
// registers and walks over the (small) batch.  Used for the step-embedding MLP and the per-layer tap-bias tables.
__global__ void __launch_bounds__(256) linear_f32_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ bias, float* __restrict__ y, int Bt,
                                                         int Cin, int Cout, int act) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int co = blockIdx.x * 8 + warp;
  if (co >= Cout) return;
  const float* wr = w + (size_t)co * Cin;
  const float bv = bias ? bias[co] : 0.f;
  for (int b0 = blockIdx.y; b0 < Bt; b0 += gridDim.y) {
    const float* xr = x + (size_t)b0 * Cin;
    float acc = 0.f;
    for (int ci = lane; ci < Cin; ci += 32) acc = fmaf(wr[ci], xr[ci], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) y[(size_t)b0 * Cout + co] = apply_act(acc + bv, act);
  }
}

inline unsigned blocks_for(size_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

}  // namespace svsk

using namespace svsk;

extern "C" int svsk_conv1d_f32(const svsk_conv1d_f32_params* pp, void* stream) {
  SVSK_REQUIRE(pp != nullptr, SVSK_E_ARG, "conv1d_f32: null params");
  svsk_conv1d_f32_params p = *pp;
  SVSK_REQUIRE(p.x && p.w && p.y, SVSK_E_ARG, "conv1d_f32: null tensor");
  SVSK_REQUIRE(p.B > 0 && p.Cin > 0 && p.Cout > 0 && p.T > 0, SVSK_E_ARG, "conv1d_f32: empty shape B=%d Cin=%d Cout=%d T=%d",
               p.B, p.Cin, p.Cout, p.T);
  SVSK_REQUIRE(p.ksize >= 1 && p.ksize <= kMaxK, SVSK_E_ARG, "conv1d_f32: ksize %d not in [1,%d]", p.ksize, kMaxK);
  SVSK_REQUIRE(p.dilation >= 1, SVSK_E_ARG, "conv1d_f32: dilation %d", p.dilation);
  SVSK_REQUIRE(p.pad_mode >= SVSK_PAD_ZEROS && p.pad_mode <= SVSK_PAD_INDEXED, SVSK_E_ARG, "conv1d_f32: pad_mode %d",
               p.pad_mode);
  SVSK_REQUIRE(p.B <= 65535, SVSK_E_ARG, "conv1d_f32: B=%d > 65535", p.B);
  int T_in = p.T;
  if (p.pad_mode == SVSK_PAD_VALID) T_in = p.T + (p.ksize - 1) * p.dilation;
  if (p.pad_mode == SVSK_PAD_INDEXED)
    SVSK_REQUIRE(p.ksize == 3 && p.idx_past && p.idx_future, SVSK_E_ARG, "conv1d_f32: INDEXED needs ksize 3 + index arrays");
  if (p.pad_mode == SVSK_PAD_REFLECT) {
    int reach = max(p.tap_origin, p.ksize - 1 - p.tap_origin) * p.dilation;
    SVSK_REQUIRE(reach < p.T, SVSK_E_ARG, "conv1d_f32: reflect padding %d needs T > %d (T=%d)", reach, reach, p.T);
  }
  dim3 grid(ceil_div(p.T, kTT), ceil_div(p.Cout, kCOT), p.B);
  conv1d_f32_kernel<<<grid, 256, 0, as_stream(stream)>>>(p, T_in);
  return check_launch("conv1d_f32");
}

extern "C" int svsk_linear_f32(const float* x, const float* w, const float* bias, float* y, int Bt, int Cin, int Cout,
                               int act, void* stream) {
  SVSK_REQUIRE(x && w && y && Bt > 0 && Cin > 0 && Cout > 0, SVSK_E_ARG, "linear_f32: bad args");
  dim3 grid(ceil_div(Cout, 8), Bt < 64 ? Bt : 64);
  linear_f32_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, w, bias, y, Bt, Cin, Cout, act);
  return check_launch("linear_f32");
}

extern "C" int svsk_gated_act_f32(const float* y, float* z, int B, int H, int T, int order, void* stream) {
  SVSK_REQUIRE(y && z && B > 0 && H > 0 && T > 0, SVSK_E_ARG, "gated_act_f32: bad args");
  SVSK_REQUIRE(order == 0 || order == 1, SVSK_E_ARG, "gated_act_f32: order %d", order);
  size_t total = (size_t)B * H * T;
  gated_act_kernel<<<blocks_for(total, 256), 256, 0, as_stream(stream)>>>(y, z, H, T, order, total);
  return check_launch("gated_act_f32");
}

extern "C" int svsk_diffnet_residual_skip_f32(const float* o, float* x, float* skip, int B, int C, int T,
                                              int init_skip, void* stream) {
  SVSK_REQUIRE(o && x && skip && B > 0 && C > 0 && T > 0, SVSK_E_ARG, "diffnet_residual_skip_f32: bad args");
  size_t total = (size_t)B * C * T;
  diffnet_residual_skip_kernel<<<blocks_for(total, 256), 256, 0, as_stream(stream)>>>(o, x, skip, C, T, init_skip, total);
  return check_launch("diffnet_residual_skip_f32");
}

extern "C" int svsk_scale_act_f32(const float* x, float* y, size_t n, float alpha, int act, void* stream) {
  SVSK_REQUIRE(x && y, SVSK_E_ARG, "scale_act_f32: null");
  if (n == 0) return 0;
  scale_act_kernel<<<blocks_for(n, 256), 256, 0, as_stream(stream)>>>(x, y, n, alpha, act);
  return check_launch("scale_act_f32");
}

extern "C" int svsk_sinusoidal_embedding_f32(const int64_t* t, float* out, int B, int dim, void* stream) {
  SVSK_REQUIRE(t && out && B > 0 && dim >= 4 && dim % 2 == 0, SVSK_E_ARG, "sinusoidal_embedding_f32: bad args");
  // -ln(10000)/(half-1) evaluated in double then rounded once, as the reference's python float does (denoiser.py:22-23)
  float neg_scale = (float)(-log(10000.0) / (double)(dim / 2 - 1));
  sinusoidal_embedding_kernel<<<blocks_for((size_t)B * dim / 2, 128), 128, 0, as_stream(stream)>>>(t, out, B, dim,
                                                                                                  neg_scale);
  return check_launch("sinusoidal_embedding_f32");
}

static unsigned per_batch_blocks(size_t per_batch) {
  size_t b = (per_batch + 255) / 256;
  return (unsigned)(b > 1184 ? 1184 : (b == 0 ? 1 : b));  // 8 x 148 SMs, grid-stride beyond
}

extern "C" int svsk_ddpm_update_f32(const float* x, const float* eps, const float* z, float* out, const int64_t* t,
                                    const float* sra, const float* srm1, const float* c1, const float* c2,
                                    const float* plv, int B, size_t per_batch, int clip_denoised, void* stream) {
  SVSK_REQUIRE(x && eps && z && out && t && sra && srm1 && c1 && c2 && plv, SVSK_E_ARG, "ddpm_update_f32: null");
  SVSK_REQUIRE(B > 0 && B <= 65535 && per_batch > 0, SVSK_E_ARG, "ddpm_update_f32: bad shape");
  DdpmTables tab{sra, srm1, c1, c2, plv};
  ddpm_update_kernel<<<dim3(per_batch_blocks(per_batch), B), 256, 0, as_stream(stream)>>>(x, eps, z, out, t, tab,
                                                                                         per_batch, clip_denoised);
  return check_launch("ddpm_update_f32");
}

extern "C" int svsk_q_sample_f32(const float* x0, const float* noise, float* out, const int64_t* t, const float* sac,
                                 const float* somac, int B, size_t per_batch, void* stream) {
  SVSK_REQUIRE(x0 && noise && out && t && sac && somac, SVSK_E_ARG, "q_sample_f32: null");
  SVSK_REQUIRE(B > 0 && B <= 65535 && per_batch > 0, SVSK_E_ARG, "q_sample_f32: bad shape");
  q_sample_kernel<<<dim3(per_batch_blocks(per_batch), B), 256, 0, as_stream(stream)>>>(x0, noise, out, t, sac, somac,
                                                                                      per_batch);
  return check_launch("q_sample_f32");
}

extern "C" int svsk_plms_transfer_f32(const float* x, const float* noise_t, float* out, const int64_t* t, int interval,
                                      const float* alphas_cumprod, int B, size_t per_batch, void* stream) {
  SVSK_REQUIRE(x && noise_t && out && t && alphas_cumprod, SVSK_E_ARG, "plms_transfer_f32: null");
  SVSK_REQUIRE(B > 0 && B <= 65535 && per_batch > 0 && interval >= 0, SVSK_E_ARG, "plms_transfer_f32: bad shape");
  plms_transfer_kernel<<<dim3(per_batch_blocks(per_batch), B), 256, 0, as_stream(stream)>>>(x, noise_t, out, t, interval,
                                                                                           alphas_cumprod, per_batch);
  return check_launch("plms_transfer_f32");
}

extern "C" int svsk_lincomb_f32(const float* const* in, const float* coef, int n_in, float* out, size_t n, void* stream) {
  SVSK_REQUIRE(in && coef && out && n_in >= 1 && n_in <= 4, SVSK_E_ARG, "lincomb_f32: bad args");
  LinComb lc{};
  lc.n_in = n_in;
  for (int i = 0; i < n_in; ++i) {
    SVSK_REQUIRE(in[i] != nullptr, SVSK_E_ARG, "lincomb_f32: null input %d", i);
    lc.in[i] = in[i];
    lc.coef[i] = coef[i];
  }
  if (n == 0) return 0;
  lincomb_kernel<<<blocks_for(n, 256), 256, 0, as_stream(stream)>>>(lc, out, n);
  return check_launch("lincomb_f32");
}

extern "C" int svsk_pd_index(const float* d, int32_t* idx_past, int32_t* idx_future, int B, int T, int dilation,
                             void* stream) {
  SVSK_REQUIRE(d && idx_past && idx_future && B > 0 && T > 0 && dilation >= 1, SVSK_E_ARG, "pd_index: bad args");
  size_t total = (size_t)B * T;
  pd_index_kernel<<<blocks_for(total, 256), 256, 0, as_stream(stream)>>>(d, idx_past, idx_future, T, dilation, total);
  return check_launch("pd_index");
}

extern "C" int svsk_upsample_smooth_f32(const float* in, const float* taps, float* out, int R, int Tin, int scale,
                                        void* stream) {
  SVSK_REQUIRE(in && taps && out && R > 0 && Tin > 0 && scale >= 1, SVSK_E_ARG, "upsample_smooth_f32: bad args");
  size_t total = (size_t)R * Tin * scale;
  upsample_smooth_kernel<<<blocks_for(total, 256), 256, 0, as_stream(stream)>>>(in, taps, out, Tin, scale, total);
  return check_launch("upsample_smooth_f32");
}

extern "C" int svsk_periodic_mix_f32(const float* a, const float* h, const float* n, float* s, float* h2, float* n2,
                                     size_t cnt, void* stream) {
  SVSK_REQUIRE(a && h && n && s, SVSK_E_ARG, "periodic_mix_f32: null");
  if (cnt == 0) return 0;
  periodic_mix_kernel<<<blocks_for(cnt, 256), 256, 0, as_stream(stream)>>>(a, h, n, s, h2, n2, cnt);
  return check_launch("periodic_mix_f32");
}
