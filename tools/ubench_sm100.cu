// Micro-benchmark entry point (profiling aid, not on the product path): cycles per tcgen05.mma for the operand
// shapes the block kernels use, operands resident in shared memory (no TMA traffic), one issuing thread.
#include <cuda_bf16.h>

#include "sm100_ptx.cuh"  // -I ensemble_svs_with_interactions_b200/csrc (csrc/build.py:build_ubench)
#include "svsk_common.cuh"

namespace svsk {

template <int kCtaGroup>
__global__ void __launch_bounds__(128, 1) ubench_umma_kernel(int N, int iters, int advance, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 192 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_mbar_init(); }
  if (warp == 1) {
    if (kCtaGroup == 2) { ptx::tmem_alloc2(&tmem_base, 512); ptx::tmem_relinquish2(); }
    else { ptx::tmem_alloc(&tmem_base, 512); ptx::tmem_relinquish(); }
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  if (kCtaGroup == 2) ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_base;
  const uint32_t rank = kCtaGroup == 2 ? ptx::cluster_ctarank() : 0;
  if (warp == 0 && lane == 0 && rank == 0) {
    const uint32_t idesc = ptx::umma_idesc_bf16_f32(kCtaGroup == 2 ? 256 : 128, N);
    const uint32_t a0 = ptx::smem_u32(smem);             // A tiles: 16 KB each (128 rows x 64)
    const uint32_t b0 = a0 + 4 * 16384;                  // B tiles: up to 32 KB each (256 rows x 64)
    if (advance >= 32) {
      // issue-loop mode: groups of 4 MMAs with optional per-group extras, never waiting for the MMAs themselves:
      // bit0 = two try_waits on a long-completed barrier, bit1 = tcgen05.fence::after_thread_sync, bit2 = commit to a
      // barrier nobody waits on, bit3 = test_wait instead of try_wait.  Pipe-bound = 4 x nominal per group.
      const int v = advance - 32;
      __shared__ uint64_t done_bar, sink_bar;
      ptx::mbar_init(&done_bar, 1);
      ptx::mbar_init(&sink_bar, 1);
      ptx::fence_mbar_init();
      ptx::mbar_arrive(&done_bar);  // phase 0 of done_bar is complete from here on
      const long long t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        if (v & 1) {
          if (v & 8) { ptx::mbar_wait_spin(&done_bar, 0); ptx::mbar_wait_spin(&done_bar, 0); }
          else { ptx::mbar_wait(&done_bar, 0); ptx::mbar_wait(&done_bar, 0); }
        }
        if (v & 2) ptx::tc_fence_after();
        for (int g = 0; g < 4; ++g) {
          if (kCtaGroup == 2) ptx::umma2_bf16(tmem, ptx::umma_desc_k_sw128(a0 + g * 32), ptx::umma_desc_k_sw128(b0 + g * 32), idesc, 1);
          else ptx::umma_bf16(tmem, ptx::umma_desc_k_sw128(a0 + g * 32), ptx::umma_desc_k_sw128(b0 + g * 32), idesc, 1);
        }
        if (v & 4) { if (kCtaGroup == 2) ptx::umma_commit2_mc(&sink_bar, 1); else ptx::umma_commit(&sink_bar); }
      }
      const long long t1 = clock64();
      if (kCtaGroup == 2) ptx::umma_commit2_mc(&bar, 1); else ptx::umma_commit(&bar);
      ptx::mbar_wait(&bar, 0);
      const long long t2 = clock64();
      out[blockIdx.x * 2] = t1 - t0;
      out[blockIdx.x * 2 + 1] = t2 - t0;
    } else if (advance >= 16) {
      // serial mode: groups of (advance - 16 + 1) MMAs, each followed by commit + wait: (group time) - (MMA time) = latency
      // from the last MMA's completion to the issuing thread seeing the barrier flip
      const int group = advance - 16 + 1;
      const long long t0 = clock64();
      uint32_t ph = 0;
      for (int i = 0; i < iters; ++i) {
        for (int g = 0; g < group; ++g) {
          if (kCtaGroup == 2) ptx::umma2_bf16(tmem, ptx::umma_desc_k_sw128(a0), ptx::umma_desc_k_sw128(b0), idesc, 1);
          else ptx::umma_bf16(tmem, ptx::umma_desc_k_sw128(a0), ptx::umma_desc_k_sw128(b0), idesc, 1);
        }
        if (kCtaGroup == 2) ptx::umma_commit2_mc(&bar, 1); else ptx::umma_commit(&bar);
        ptx::mbar_wait_spin(&bar, ph);
        ph ^= 1;
      }
      const long long t1 = clock64();
      out[blockIdx.x * 2] = t1 - t0;
      out[blockIdx.x * 2 + 1] = t1 - t0;
    } else {
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const int sel = advance ? (i & 3) : 0;
      const uint32_t aa = a0 + sel * 16384 + (i & 3) * 32 * (advance ? 1 : 0);
      const uint32_t bb = b0 + sel * 32768;
      if (kCtaGroup == 2) ptx::umma2_bf16(tmem, ptx::umma_desc_k_sw128(aa), ptx::umma_desc_k_sw128(bb), idesc, i != 0);
      else ptx::umma_bf16(tmem, ptx::umma_desc_k_sw128(aa), ptx::umma_desc_k_sw128(bb), idesc, i != 0);
    }
    const long long t1 = clock64();
    if (kCtaGroup == 2) ptx::umma_commit2_mc(&bar, 1); else ptx::umma_commit(&bar);
    ptx::mbar_wait(&bar, 0);
    const long long t2 = clock64();
    out[blockIdx.x * 2] = t1 - t0;      // issue time
    out[blockIdx.x * 2 + 1] = t2 - t0;  // until the last MMA has completed
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (kCtaGroup == 2) ptx::cluster_sync_all();
  if (warp == 1) { if (kCtaGroup == 2) ptx::tmem_dealloc2(tmem, 512); else ptx::tmem_dealloc(tmem, 512); }
}

// Correctness probe: does a K-major SWIZZLE_128B operand descriptor work when its start address is offset by a whole
// number of 128-byte rows (not a multiple of the 1024-byte swizzle atom)?  A = rows r0 .. r0+127 of a window tile,
// B = 64x64 identity, so D[m][n] must equal win[r0+m][n].  mode 0: base-offset field 0; mode 1: (addr >> 7) & 7.
__global__ void __launch_bounds__(128, 1) ubench_rowshift_kernel(const __nv_bfloat16* win, int rows, int r0, int mode, float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* wsm = smem;               // window: rows x 128 bytes
  uint8_t* bsm = smem + 256 * 128;   // identity: 64 rows x 128 bytes
  for (int i = threadIdx.x; i < rows * 8; i += 128) {
    const int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(wsm + ptx::sw128_offset((uint32_t)r, (uint32_t)c)) =
        *reinterpret_cast<const uint4*>(win + (size_t)r * 64 + c * 8);
  }
  for (int i = threadIdx.x; i < 64 * 64; i += 128) {
    const int n = i >> 6, k = i & 63;
    reinterpret_cast<__nv_bfloat16*>(bsm + ptx::sw128_offset((uint32_t)n, (uint32_t)(k >> 3)))[k & 7] =
        __float2bfloat16_rn(n == k ? 1.f : 0.f);
  }
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_mbar_init(); }
  if (warp == 1) { ptx::tmem_alloc(&tmem_base, 64); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_base;
  if (threadIdx.x == 0) {
    const uint32_t idesc = ptx::umma_idesc_bf16_f32(128, 64);
    const uint32_t a_addr = ptx::smem_u32(wsm) + (uint32_t)r0 * 128u, b_addr = ptx::smem_u32(bsm);
    for (int k = 0; k < 4; ++k) {
      uint64_t da = ptx::umma_desc_k_sw128(a_addr + k * 32);
      if (mode == 1) da |= (uint64_t)((a_addr >> 7) & 7) << 49;
      ptx::umma_bf16(tmem, da, ptx::umma_desc_k_sw128(b_addr + k * 32), idesc, k != 0);
    }
    ptx::umma_commit(&bar);
  }
  ptx::mbar_wait(&bar, 0);
  ptx::tc_fence_after();
  for (int c0 = 0; c0 < 64; c0 += 16) {
    uint32_t r[16];
    ptx::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
    ptx::tmem_ld_wait();
    for (int e = 0; e < 16; ++e) out[(size_t)(warp * 32 + lane) * 64 + c0 + e] = __uint_as_float(r[e]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem, 64);
}


// Throughput of the special-function unit and of candidate gate formulations z = tanh(a) * sigmoid(b) (profiling aid for
// the uSFGAN / DiffNet epilogues).  Every thread keeps 8 independent chains; out[block] = cycles for `iters` rounds.
__device__ __forceinline__ float ub_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ub_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// 2^y for y in [-126, 126] on the FMA pipe: round-to-nearest split + degree-4 polynomial on [-0.5, 0.5] + exponent splice
__device__ __forceinline__ float ub_exp2_fma(float y) {
  const float magic = 12582912.f;  // 1.5 * 2^23
  const float yr = y + magic;
  const float n = yr - magic;
  const float f = y - n;
  float p = fmaf(f, 0.0096181291f, 0.0555041087f);
  p = fmaf(p, f, 0.2402265070f);
  p = fmaf(p, f, 0.6931471806f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(yr) << 23));
}
__global__ void ubench_sfu_kernel(int mode, int iters, unsigned long long* out, float* sink) {
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = 0.001f * (threadIdx.x + 1) + 0.1f * i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float v = x[i];
      if (mode == 0) v = ptx::tanh_approx(v);
      else if (mode == 1) v = ub_ex2(v) - 1.0f;
      else if (mode == 2) v = ub_rcp(v) + 0.5f;
      else if (mode == 3) v = fmaf(v, 0.999f, 0.001f);
      else if (mode == 4) {  // gate, as the kernels have it: 2 MUFU.TANH
        v = ptx::tanh_approx(v + 0.1f) * fmaf(ptx::tanh_approx(0.5f * (v - 0.2f)), 0.5f, 0.5f) + 0.3f;
      } else if (mode == 5) {  // gate with 2 EX2 + 1 RCP
        const float a = fminf(fmaxf(v + 0.1f, -10.f), 10.f), b = fminf(fmaxf(v - 0.2f, -30.f), 30.f);
        const float e1 = ub_ex2(a * -2.885390082f), e2 = ub_ex2(b * -1.442695041f);
        v = (1.f - e1) * ub_rcp((1.f + e1) * (1.f + e2)) + 0.3f;
      } else if (mode == 6) {  // gate with both exponentials on the FMA pipe + 1 RCP
        const float a = fminf(fmaxf(v + 0.1f, -10.f), 10.f), b = fminf(fmaxf(v - 0.2f, -30.f), 30.f);
        const float e1 = ub_exp2_fma(a * -2.885390082f), e2 = ub_exp2_fma(b * -1.442695041f);
        v = (1.f - e1) * ub_rcp((1.f + e1) * (1.f + e2)) + 0.3f;
      } else if (mode == 7) {  // one exponential on the SFU, one on the FMA pipe
        const float a = fminf(fmaxf(v + 0.1f, -10.f), 10.f), b = fminf(fmaxf(v - 0.2f, -30.f), 30.f);
        const float e1 = ub_ex2(a * -2.885390082f), e2 = ub_exp2_fma(b * -1.442695041f);
        v = (1.f - e1) * ub_rcp((1.f + e1) * (1.f + e2)) + 0.3f;
      }
      x[i] = v;
    }
  }
  const long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc += x[i];
  if (acc == 12345.678f) sink[0] = acc;
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
}


// TMEM read throughput: every warp reads its own lane quarter, `cols` fp32 columns per round, with tcgen05.ld
// 32x32b.x16 / .x32 / .x64 and one wait::ld per round (mode 0) or per instruction (mode 1).
__device__ __forceinline__ void ub_tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, "
      "%25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]),
        "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]),
        "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__global__ void ubench_tmem_ld_kernel(int width, int per_inst_wait, int iters, unsigned long long* out, float* sink) {
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { ptx::tmem_alloc(&tmem_base, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    // one round = 128 columns of this warp's 32 lanes = 16 KB
    if (width == 16) {
#pragma unroll
      for (int c = 0; c < 128; c += 16) {
        uint32_t r[16];
        ptx::tmem_ld16(tmem + ((it * 128 + c) & 511), r);
        if (per_inst_wait) ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) acc ^= r[i];
      }
    } else {
#pragma unroll
      for (int c = 0; c < 128; c += 32) {
        uint32_t r[32];
        ub_tmem_ld32(tmem + ((it * 128 + c) & 511), r);
        if (per_inst_wait) ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc ^= r[i];
      }
    }
    ptx::tmem_ld_wait();
  }
  const long long t1 = clock64();
  if (acc == 0x12345678u) sink[0] = 1.f;
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
  if (warp == 0) ptx::tmem_dealloc(tmem_base, 512);
}

}  // namespace svsk

using namespace svsk;

// out: [grid][2] cycles (issue, complete).  cta_group 1 or 2; grid = number of CTAs (even for cta_group 2).
extern "C" SVSK_API int svsk_ubench_umma(int cta_group, int N, int iters, int advance, int grid, unsigned long long* out,
                                         void* stream) {
  SVSK_REQUIRE(out && (cta_group == 1 || cta_group == 2) && N >= 16 && N <= 256 && N % 16 == 0 && iters > 0 && grid > 0,
               SVSK_E_ARG, "ubench_umma: bad args");
  int rc = require_sm100();
  if (rc) return rc;
  const int smem_bytes = 193 * 1024;
  if (cta_group == 1) {
    cudaFuncSetAttribute(ubench_umma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    ubench_umma_kernel<1><<<grid, 128, smem_bytes, as_stream(stream)>>>(N, iters, advance, out);
  } else {
    SVSK_REQUIRE(grid % 2 == 0, SVSK_E_ARG, "ubench_umma: grid must be even for cta_group 2");
    cudaFuncSetAttribute(ubench_umma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = as_stream(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, ubench_umma_kernel<2>, N, iters, advance, out);
    if (e != cudaSuccess) return fail((int)e, "ubench_umma: %s", cudaGetErrorString(e));
  }
  return check_launch("ubench_umma");
}

extern "C" SVSK_API int svsk_ubench_rowshift(const void* win, int rows, int r0, int mode, float* out, void* stream) {
  SVSK_REQUIRE(win && out && rows >= 128 && rows <= 256 && r0 >= 0 && r0 + 128 <= rows, SVSK_E_ARG, "ubench_rowshift: bad args");
  int rc = require_sm100();
  if (rc) return rc;
  const int smem_bytes = 256 * 128 + 64 * 128 + 1024;
  ubench_rowshift_kernel<<<1, 128, smem_bytes, as_stream(stream)>>>((const __nv_bfloat16*)win, rows, r0, mode, out);
  return check_launch("ubench_rowshift");
}

// out [grid] cycles; every thread does iters x 8 operations of `mode` (see ubench_sfu_kernel).
extern "C" SVSK_API int svsk_ubench_sfu(int mode, int warps, int iters, int grid, unsigned long long* out, float* sink, void* stream) {
  SVSK_REQUIRE(out && sink && warps >= 1 && warps <= 32 && iters > 0 && grid > 0, SVSK_E_ARG, "ubench_sfu: bad args");
  ubench_sfu_kernel<<<grid, warps * 32, 0, as_stream(stream)>>>(mode, iters, out, sink);
  return check_launch("ubench_sfu");
}

// out [grid] cycles for `iters` rounds of 128 columns x 32 lanes per warp (16 KB per warp and round).
extern "C" SVSK_API int svsk_ubench_tmem_ld(int width, int per_inst_wait, int warps, int iters, int grid, unsigned long long* out,
                                            float* sink, void* stream) {
  SVSK_REQUIRE(out && sink && (width == 16 || width == 32) && warps >= 1 && warps <= 32 && iters > 0 && grid > 0, SVSK_E_ARG,
               "ubench_tmem_ld: bad args");
  int rc = require_sm100();
  if (rc) return rc;
  ubench_tmem_ld_kernel<<<grid, warps * 32, 0, as_stream(stream)>>>(width, per_inst_wait, iters, out, sink);
  return check_launch("ubench_tmem_ld");
}
