// Library-level entry points: version, last-error string, device check, layout conversion kernels.
#include "svsk_common.cuh"

#include <cuda_bf16.h>
#include <mutex>

namespace svsk {

char* last_error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}

int require_sm100() {
  static std::mutex mu;
  static int cached[64];
  static bool have[64] = {false};
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail((int)e, "cudaGetDevice: %s", cudaGetErrorString(e));
  if (dev >= 0 && dev < 64) {
    std::lock_guard<std::mutex> lk(mu);
    if (have[dev]) return cached[dev] == 10 ? 0 : fail(SVSK_E_ARCH, "device %d is sm_%d0, libsvsk needs sm_100", dev, cached[dev]);
  }
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) return fail((int)e, "cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
  if (dev >= 0 && dev < 64) {
    std::lock_guard<std::mutex> lk(mu);
    cached[dev] = major;
    have[dev] = true;
  }
  if (major != 10) return fail(SVSK_E_ARCH, "device %d is sm_%d0, libsvsk needs sm_100", dev, major);
  return 0;
}

// [B][C][T] fp32 -> [B][T][Cp] (bf16 and/or fp32).  32x32 smem transpose tiles.
__global__ void nct_to_ntc_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ ob, float* __restrict__ of,
                                  int C, int T, int Cp) {
  __shared__ float tile[32][33];
  int b = blockIdx.z;
  int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    int c = c0 + i, t = t0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && t < T) ? x[((size_t)b * C + c) * T + t] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    int t = t0 + i, c = c0 + threadIdx.x;
    if (t < T && c < Cp) {
      float v = tile[threadIdx.x][i];
      size_t o = ((size_t)b * T + t) * Cp + c;
      if (ob) ob[o] = __float2bfloat16_rn(v);
      if (of) of[o] = v;
    }
  }
}

__global__ void ntc_to_nct_kernel(const float* __restrict__ x, float* __restrict__ y, int C, int T, int Cp, float alpha) {
  __shared__ float tile[32][33];
  int b = blockIdx.z;
  int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    int t = t0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (t < T && c < C) ? x[((size_t)b * T + t) * Cp + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    int c = c0 + i, t = t0 + threadIdx.x;
    if (c < C && t < T) y[((size_t)b * C + c) * T + t] = alpha * tile[threadIdx.x][i];
  }
}

__global__ void ntc_bf16_to_nct_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, int C, int T, int Cp) {
  __shared__ float tile[32][33];
  int b = blockIdx.z;
  int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    int t = t0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (t < T && c < C) ? __bfloat162float(x[((size_t)b * T + t) * Cp + c]) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    int c = c0 + i, t = t0 + threadIdx.x;
    if (c < C && t < T) y[((size_t)b * C + c) * T + t] = tile[threadIdx.x][i];
  }
}

__global__ void cast_scale_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, size_t n, float alpha,
                                       int relu) {
  size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    float4 v = *reinterpret_cast<const float4*>(x + i);
    float a = v.x * alpha, b = v.y * alpha, c = v.z * alpha, d = v.w * alpha;
    if (relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); c = fmaxf(c, 0.f); d = fmaxf(d, 0.f); }
    __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&lo);
    pk.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(y + i) = pk;
  } else {
    for (; i < n; ++i) {
      float a = x[i] * alpha;
      if (relu) a = fmaxf(a, 0.f);
      y[i] = __float2bfloat16_rn(a);
    }
  }
}

}  // namespace svsk

using namespace svsk;

extern "C" const char* svsk_last_error(void) { return last_error_buffer(); }
extern "C" int svsk_version(void) { return SVSK_VERSION; }

extern "C" int svsk_device_check(int device) {
  int major = 0;
  cudaError_t e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
  if (e != cudaSuccess) return fail((int)e, "svsk_device_check: %s", cudaGetErrorString(e));
  if (major != 10) return fail(SVSK_E_ARCH, "device %d is sm_%d0, libsvsk needs sm_100", device, major);
  return 0;
}

extern "C" int svsk_nct_to_ntc(const float* x, void* out_bf16, float* out_f32, int B, int C, int T, int Cp, void* stream) {
  SVSK_REQUIRE(x && (out_bf16 || out_f32), SVSK_E_ARG, "nct_to_ntc: null");
  SVSK_REQUIRE(B > 0 && B <= 65535 && C > 0 && T > 0 && Cp >= C, SVSK_E_ARG, "nct_to_ntc: bad shape");
  dim3 grid(ceil_div(T, 32), ceil_div(Cp, 32), B);
  nct_to_ntc_kernel<<<grid, dim3(32, 8), 0, as_stream(stream)>>>(x, (__nv_bfloat16*)out_bf16, out_f32, C, T, Cp);
  return check_launch("nct_to_ntc");
}

extern "C" int svsk_ntc_to_nct_f32(const float* x, float* y, int B, int C, int T, int Cp, float alpha, void* stream) {
  SVSK_REQUIRE(x && y, SVSK_E_ARG, "ntc_to_nct_f32: null");
  SVSK_REQUIRE(B > 0 && B <= 65535 && C > 0 && T > 0 && Cp >= C, SVSK_E_ARG, "ntc_to_nct_f32: bad shape");
  dim3 grid(ceil_div(T, 32), ceil_div(C, 32), B);
  ntc_to_nct_kernel<<<grid, dim3(32, 8), 0, as_stream(stream)>>>(x, y, C, T, Cp, alpha);
  return check_launch("ntc_to_nct_f32");
}

extern "C" int svsk_ntc_bf16_to_nct_f32(const void* x, float* y, int B, int C, int T, int Cp, void* stream) {
  SVSK_REQUIRE(x && y, SVSK_E_ARG, "ntc_bf16_to_nct_f32: null");
  SVSK_REQUIRE(B > 0 && B <= 65535 && C > 0 && T > 0 && Cp >= C, SVSK_E_ARG, "ntc_bf16_to_nct_f32: bad shape");
  dim3 grid(ceil_div(T, 32), ceil_div(C, 32), B);
  ntc_bf16_to_nct_kernel<<<grid, dim3(32, 8), 0, as_stream(stream)>>>((const __nv_bfloat16*)x, y, C, T, Cp);
  return check_launch("ntc_bf16_to_nct_f32");
}

extern "C" int svsk_cast_scale_bf16(const float* x, void* y, size_t n, float alpha, int relu, void* stream) {
  SVSK_REQUIRE(x && y, SVSK_E_ARG, "cast_scale_bf16: null");
  SVSK_REQUIRE(((uintptr_t)x % 16) == 0 && ((uintptr_t)y % 8) == 0, SVSK_E_ALIGN, "cast_scale_bf16: alignment");
  if (n == 0) return 0;
  size_t threads = (n + 3) / 4;
  cast_scale_bf16_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, as_stream(stream)>>>(x, (__nv_bfloat16*)y, n, alpha, relu);
  return check_launch("cast_scale_bf16");
}
