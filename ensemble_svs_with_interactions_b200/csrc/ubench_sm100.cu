// Micro-benchmark entry point (profiling aid, not on the product path): cycles per tcgen05.mma for the operand
// shapes the block kernels use, operands resident in shared memory (no TMA traffic), one issuing thread.
#include <cuda_bf16.h>

#include "sm100_ptx.cuh"
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
  ptx::tc_fence_before();
  __syncthreads();
  if (kCtaGroup == 2) ptx::cluster_sync_all();
  if (warp == 1) { if (kCtaGroup == 2) ptx::tmem_dealloc2(tmem, 512); else ptx::tmem_dealloc(tmem, 512); }
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
