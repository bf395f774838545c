// Acoustic post-processing between the diffusion models and the vocoder on the device:
//   * svsk_filtfilt_f32 — nnsvs.dsp.lowpass_filter (dsp.py:10-33; called per feature dimension at gen.py:1500-1513):
//     scipy.signal.filtfilt with its defaults, i.e. odd extension by `pad` samples, a direct-form-II-transposed IIR pass
//     forward started from zi * ext[0], the same pass backward started from zi * y[last], extension dropped;
//   * svsk_variance_scaling_f32 — nnsvs.postfilters.variance_scaling (postfilters.py:9-46; gen.py:1394-1418).
// Both walk feature trajectories [track][frame][dim] with one thread per (track, dim): neighbouring threads read
// neighbouring dims of the same frame (coalesced), the recurrences run in fp64 like the reference's numpy / scipy code.
// The work is tiny (hundreds of trajectories) and purely latency-bound; what it buys is that the features never leave
// the GPU between GaussianDiffusion.inference and the vocoder.
#include "svsk_common.cuh"

namespace svsk {

constexpr int kMaxOrder = 8;

struct FiltArgs {
  const float* x;
  float* y;
  double* scratch;
  const int32_t* lengths;
  double b[kMaxOrder + 1], a[kMaxOrder + 1], zi[kMaxOrder];
  int B, T, D, pad, min_len;
};

template <int N>
__device__ __forceinline__ double df2t_step(const double (&b)[N + 1], const double (&a)[N + 1], double (&z)[N], double x) {
  // y_t -> z0_t -> y_{t+1} is the serial chain: everything that does not need y is computed off it, leaving one add and
  // one fma per sample (same arithmetic as scipy's lfilter up to the rounding of b0 x + z0)
  const double y = z[0] + b[0] * x;
#pragma unroll
  for (int i = 0; i < N - 1; ++i) z[i] = fma(-a[i + 1], y, fma(b[i + 1], x, z[i + 1]));
  z[N - 1] = fma(-a[N], y, b[N] * x);
  return y;
}

template <int N>
__global__ void filtfilt_kernel(const FiltArgs p) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x, bi = blockIdx.y;
  if (d >= p.D) return;
  const int len = p.lengths ? min(max(p.lengths[bi], 0), p.T) : p.T;
  const float* x = p.x + (size_t)bi * p.T * p.D + d;
  float* y = p.y + (size_t)bi * p.T * p.D + d;
  for (int t = len; t < p.T; ++t) y[(size_t)t * p.D] = x[(size_t)t * p.D];   // padding frames pass through
  if (len <= p.min_len) {                                                       // dsp.py:26-28: too short, returned as is
    for (int t = 0; t < len; ++t) y[(size_t)t * p.D] = x[(size_t)t * p.D];
    return;
  }
  double b[N + 1], a[N + 1], z[N];
#pragma unroll
  for (int i = 0; i <= N; ++i) { b[i] = p.b[i]; a[i] = p.a[i]; }
  const int pad = p.pad, E = len + 2 * pad;
  double* s = p.scratch + (size_t)bi * (p.T + 2 * pad) * p.D + d;
  const double x0 = x[0], xl = x[(size_t)(len - 1) * p.D];
  // odd extension of the first `len` frames, branch-free so that a chunk's loads issue back to back:
  // ext(e) = base + sign * x[idx]
  auto ext = [&](int e) -> double {
    const bool head = e < pad, tail = e >= pad + len;
    const int idx = head ? pad - e : (tail ? 2 * len - 2 + pad - e : e - pad);
    const double v = (double)x[(size_t)idx * p.D];
    return head ? 2.0 * x0 - v : (tail ? 2.0 * xl - v : v);
  };
  const double e0 = ext(0);
#pragma unroll
  for (int i = 0; i < N; ++i) z[i] = p.zi[i] * e0;
  // the recurrence is serial but its inputs are not: fetch kChunk samples at once (independent loads in flight), then run
  // the dependent chain over registers — one thread per trajectory cannot hide a load per step otherwise
  constexpr int kChunk = 32;
  double v[kChunk];
  double last = 0.0;
  for (int e0 = 0; e0 < E; e0 += kChunk) {
#pragma unroll
    for (int i = 0; i < kChunk; ++i) v[i] = ext(min(e0 + i, E - 1));
#pragma unroll
    for (int i = 0; i < kChunk; ++i) {
      if (e0 + i < E) {
        last = df2t_step<N>(b, a, z, v[i]);
        s[(size_t)(e0 + i) * p.D] = last;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < N; ++i) z[i] = p.zi[i] * last;
  for (int e0 = E - 1; e0 >= 0; e0 -= kChunk) {
#pragma unroll
    for (int i = 0; i < kChunk; ++i) v[i] = s[(size_t)max(e0 - i, 0) * p.D];
#pragma unroll
    for (int i = 0; i < kChunk; ++i) {
      const int e = e0 - i;
      if (e >= 0) {
        const double r = df2t_step<N>(b, a, z, v[i]);
        if (e >= pad && e < pad + len) y[(size_t)(e - pad) * p.D] = (float)r;
      }
    }
  }
}

template <int N>
static void launch_filtfilt(const FiltArgs& a, cudaStream_t st) {
  dim3 grid((unsigned)((a.D + 63) / 64), (unsigned)a.B);
  filtfilt_kernel<N><<<grid, 64, 0, st>>>(a);
}

constexpr int kVsSlices = 8;  // threads along time per (track, dim)

__global__ void __launch_bounds__(64 * kVsSlices)
variance_scaling_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ gv,
                        const uint8_t* __restrict__ mask, const int32_t* __restrict__ lengths, int offset, int B, int T, int D) {
  __shared__ double red[kVsSlices][64];
  __shared__ int cnt[kVsSlices][64];
  const int dx = threadIdx.x, sl = threadIdx.y;
  const int d = blockIdx.x * 64 + dx, bi = blockIdx.y;
  const bool live = d < D;
  const int len = lengths ? min(max(lengths[bi], 0), T) : T;
  const float* xs = x + (size_t)bi * T * D + (live ? d : 0);
  float* ys = y + (size_t)bi * T * D + (live ? d : 0);
  const uint8_t* m = mask ? mask + (size_t)bi * T : nullptr;
  double sum = 0.0;
  int n = 0;
  if (live)
    for (int t = sl; t < len; t += kVsSlices)
      if (!m || m[t]) { sum += xs[(size_t)t * D]; ++n; }
  red[sl][dx] = sum;
  cnt[sl][dx] = n;
  __syncthreads();
  sum = 0.0;
  n = 0;
#pragma unroll
  for (int i = 0; i < kVsSlices; ++i) { sum += red[i][dx]; n += cnt[i][dx]; }   // same order on every slice: identical mu
  const double mu = n ? sum / n : 0.0;
  __syncthreads();
  double ss = 0.0;
  if (live)
    for (int t = sl; t < len; t += kVsSlices)
      if (!m || m[t]) { const double v = xs[(size_t)t * D] - mu; ss = fma(v, v, ss); }
  red[sl][dx] = ss;
  __syncthreads();
  ss = 0.0;
#pragma unroll
  for (int i = 0; i < kVsSlices; ++i) ss += red[i][dx];
  if (!live) return;
  const bool scale = n > 0 && d >= offset;                  // postfilters.py:24-25: no note frames -> unchanged
  const double g = scale ? sqrt((double)gv[d] / (ss / n)) : 1.0;
  for (int t = sl; t < T; t += kVsSlices) {
    const float v = xs[(size_t)t * D];
    ys[(size_t)t * D] = (scale && t < len && (!m || m[t])) ? (float)fma(g, (double)v - mu, mu) : v;
  }
}

// y[r][d] = mode 0: x * a[d] + b[d]   |   mode 1: (x - b[d]) / a[d]      (feature scalers, nnsvs/util.py:288-292,335-339)
__global__ void scale_features_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ a,
                                      const float* __restrict__ b, int mode, long long n, int D) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(i % D);
    const float v = x[i];
    y[i] = mode == 0 ? v * a[d] + b[d] : (v - b[d]) / a[d];
  }
}

// MDN head (nnsvs/mdn.py:45-74,165-212), one thread per (row, output dim): raw [rows][ld] holds the three Linears' outputs
// side by side ([log_pi | log_sigma | mu], G*D columns each, component-major).  Writes log_softmax over the G components
// of log_pi and copies of the other two as [rows][G][D]; optionally the (sigma, mu) of the most probable component.
__global__ void mdn_head_kernel(const float* __restrict__ raw, float* __restrict__ log_pi, float* __restrict__ log_sigma,
                                float* __restrict__ mu, float* __restrict__ best_sigma, float* __restrict__ best_mu, long long rows,
                                int G, int D, int ld) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= rows * D) return;
  const long long r = i / D;
  const int d = (int)(i % D);
  const float* p = raw + r * ld + d;
  float mx = -INFINITY;
  int best = 0;
  for (int g = 0; g < G; ++g) {
    const float v = p[(size_t)g * D];
    if (v > mx) { mx = v; best = g; }      // first maximum, like torch.max
  }
  float sum = 0.f;
  for (int g = 0; g < G; ++g) sum += expf(p[(size_t)g * D] - mx);
  const float lse = mx + logf(sum);
  const size_t GD = (size_t)G * D;
  for (int g = 0; g < G; ++g) {
    const size_t o = (size_t)r * GD + (size_t)g * D + d;
    if (log_pi) log_pi[o] = p[(size_t)g * D] - lse;
    if (log_sigma) log_sigma[o] = p[GD + (size_t)g * D];
    if (mu) mu[o] = p[2 * GD + (size_t)g * D];
  }
  if (best_sigma) best_sigma[r * D + d] = expf(p[GD + (size_t)best * D]);
  if (best_mu) best_mu[r * D + d] = p[2 * GD + (size_t)best * D];
}

}  // namespace svsk

using namespace svsk;

extern "C" int svsk_filtfilt_f32(const float* x, float* y, double* scratch, const int32_t* lengths, const double* b, const double* a,
                                 const double* zi, int order, int pad, int min_len, int B, int T, int D, void* stream) {
  SVSK_REQUIRE(x && y && scratch && b && a && zi, SVSK_E_ARG, "filtfilt_f32: null argument");
  SVSK_REQUIRE(order >= 1 && order <= kMaxOrder, SVSK_E_ARG, "filtfilt_f32: order %d (1..%d)", order, kMaxOrder);
  SVSK_REQUIRE(B > 0 && T > 0 && D > 0 && pad >= 1 && min_len >= pad, SVSK_E_ARG,
               "filtfilt_f32: B=%d T=%d D=%d pad=%d min_len=%d (sequences must be longer than the extension)", B, T, D, pad, min_len);
  SVSK_REQUIRE(a[0] == 1.0, SVSK_E_ARG, "filtfilt_f32: a[0] must be 1 (normalised coefficients)");
  FiltArgs p = {};
  p.x = x; p.y = y; p.scratch = scratch; p.lengths = lengths;
  for (int i = 0; i <= order; ++i) { p.b[i] = b[i]; p.a[i] = a[i]; }
  for (int i = 0; i < order; ++i) p.zi[i] = zi[i];
  p.B = B; p.T = T; p.D = D; p.pad = pad; p.min_len = min_len;
  cudaStream_t st = as_stream(stream);
  switch (order) {
    case 1: launch_filtfilt<1>(p, st); break;
    case 2: launch_filtfilt<2>(p, st); break;
    case 3: launch_filtfilt<3>(p, st); break;
    case 4: launch_filtfilt<4>(p, st); break;
    case 5: launch_filtfilt<5>(p, st); break;
    case 6: launch_filtfilt<6>(p, st); break;
    case 7: launch_filtfilt<7>(p, st); break;
    default: launch_filtfilt<8>(p, st); break;
  }
  return check_launch("filtfilt_f32");
}

extern "C" int svsk_variance_scaling_f32(const float* x, float* y, const float* gv, const uint8_t* note_mask, const int32_t* lengths,
                                         int offset, int B, int T, int D, void* stream) {
  SVSK_REQUIRE(x && y && gv && B > 0 && T > 0 && D > 0 && offset >= 0, SVSK_E_ARG, "variance_scaling_f32: bad args");
  dim3 grid((unsigned)((D + 63) / 64), (unsigned)B);
  variance_scaling_kernel<<<grid, dim3(64, kVsSlices), 0, as_stream(stream)>>>(x, y, gv, note_mask, lengths, offset, B, T, D);
  return check_launch("variance_scaling_f32");
}

extern "C" int svsk_scale_features_f32(const float* x, float* y, const float* a, const float* b, int mode, long long rows, int D,
                                       void* stream) {
  SVSK_REQUIRE(x && y && a && b && rows > 0 && D > 0 && (mode == 0 || mode == 1), SVSK_E_ARG, "scale_features_f32: bad args");
  const long long n = rows * D;
  const unsigned grid = (unsigned)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
  scale_features_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, y, a, b, mode, n, D);
  return check_launch("scale_features_f32");
}

extern "C" int svsk_mdn_head_f32(const float* raw, float* log_pi, float* log_sigma, float* mu, float* best_sigma, float* best_mu,
                                 long long rows, int G, int D, int ld, void* stream) {
  SVSK_REQUIRE(raw && rows > 0 && G >= 1 && D >= 1 && ld >= 3 * G * D, SVSK_E_ARG, "mdn_head_f32: bad args (ld >= 3 G D)");
  SVSK_REQUIRE(log_pi || log_sigma || mu || best_sigma || best_mu, SVSK_E_ARG, "mdn_head_f32: no output requested");
  const long long n = rows * D;
  mdn_head_kernel<<<(unsigned)((n + 127) / 128), 128, 0, as_stream(stream)>>>(raw, log_pi, log_sigma, mu, best_sigma, best_mu, rows, G, D, ld);
  return check_launch("mdn_head_f32");
}
