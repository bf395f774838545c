// Time-major bf16 GEMM with fused bias/activation epilogue on tcgen05: Y[n][co] = act(A[n][:] . W[co][:] + b[co]).
// Replaces the 1x1 projections around the stacks (denoiser.py:110-112,121-123; generator.py:461-466,492-493).
// M = 128 time rows per CTA (TMEM lanes), N = Cout (<= 256 columns), K streamed in 64-channel TMA boxes.
#include <cuda_bf16.h>

#include "sm100_ptx.cuh"
#include "svsk_common.cuh"
#include "tma_util.cuh"

namespace svsk {

constexpr int kLinStages = 4;
constexpr int kLinABytes = 128 * 128;  // 128 rows x 64 bf16

struct LinearArgs {
  const float* bias;
  __nv_bfloat16* y_b;
  float* y_f;
  long long N;
  int K, Cout, ldy_b, ldy_f, act, tmem_cols;
  // transposed bf16 output (svsk_usfgan_aux_frames): row n = (track n / t_rows, frame n % t_rows), column co ->
  // y_t[track * t_bstride + co * t_ld + frame]
  __nv_bfloat16* y_t;
  long long t_bstride;
  int t_rows, t_ld;
  int nstages;
  // gridDim.y > 1 with row-major outputs (svsk_diffnet_cond_project_bf16): chunk blockIdx.y of Cout weight rows is
  // written to y_b + blockIdx.y * y_chunk_stride (elements)
  long long y_chunk_stride;
};

struct __align__(8) LinearBarriers {
  uint64_t full[kLinStages];
  uint64_t empty[kLinStages];
  uint64_t d_full;
  uint32_t tmem_base;
};

__device__ __forceinline__ float lin_act(float v, int act) {
  if (act == SVSK_ACT_RELU) return fmaxf(v, 0.f);
  if (act == SVSK_ACT_SIGMOID) return ptx::sigmoid_approx(v);
  return v;
}

__global__ void __launch_bounds__(192, 1)
linear_bf16_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w,
                   const LinearArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int wbytes = a.Cout * 128;
  const int stage_bytes = kLinABytes + wbytes;
  const int nstages = a.nstages;  // <= kLinStages; fewer when K is short, so that several CTAs share an SM
  LinearBarriers* bars = reinterpret_cast<LinearBarriers*>(smem + nstages * stage_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long n0 = (long long)blockIdx.x * 128;
  const int KB = (a.K + 63) / 64;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_a);
    ptx::prefetch_tmap(&tm_w);
    for (int i = 0; i < kLinStages; ++i) {
      ptx::mbar_init(&bars->full[i], 1);
      ptx::mbar_init(&bars->empty[i], 1);
    }
    ptx::mbar_init(&bars->d_full, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(&bars->tmem_base, a.tmem_cols);
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
      for (int kb = 0; kb < KB; ++kb) {
        ptx::mbar_wait(&bars->empty[s], ph ^ 1);
        uint8_t* As = smem + s * stage_bytes;
        ptx::mbar_arrive_expect_tx(&bars->full[s], stage_bytes);
        ptx::tma_load_2d(As, &tm_a, &bars->full[s], kb * 64, (int)n0);
        ptx::tma_load_2d(As + kLinABytes, &tm_w, &bars->full[s], kb * 64, (int)blockIdx.y * a.Cout);  // y: chunk of weight rows
        if (++s == nstages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = ptx::umma_idesc_bf16_f32(128, a.Cout);
      int s = 0;
      uint32_t ph = 0;
      for (int kb = 0; kb < KB; ++kb) {
        ptx::mbar_wait(&bars->full[s], ph);
        ptx::tc_fence_after();
        const uint32_t a0 = ptx::smem_u32(smem + s * stage_bytes);
        const uint32_t b0 = a0 + kLinABytes;
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4)
          ptx::umma_bf16(tmem, ptx::umma_desc_k_sw128(a0 + k4 * 32), ptx::umma_desc_k_sw128(b0 + k4 * 32), idesc,
                         (kb | k4) != 0);
        ptx::umma_commit(&bars->empty[s]);
        if (++s == nstages) { s = 0; ph ^= 1; }
      }
      ptx::umma_commit(&bars->d_full);
    }
  } else {
    const int q = warp & 3;
    const long long n = n0 + q * 32 + lane;
    ptx::mbar_wait(&bars->d_full, 0);
    ptx::tc_fence_after();
    const bool vec_f = a.y_f && (a.ldy_f % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.y_f) & 15) == 0);
    const bool vec_b = a.y_b && (a.ldy_b % 8 == 0) && ((reinterpret_cast<uintptr_t>(a.y_b) & 15) == 0);
    __nv_bfloat16* yt = nullptr;
    if (a.y_t && n < a.N) {
      const long long trk = n / a.t_rows;
      yt = a.y_t + trk * a.t_bstride + (n - trk * a.t_rows) + (size_t)blockIdx.y * a.Cout * a.t_ld;
    }
    __nv_bfloat16* const y_b = a.y_b ? a.y_b + (size_t)blockIdx.y * a.y_chunk_stride : nullptr;
    for (int c0 = 0; c0 < a.Cout; c0 += 16) {
      uint32_t r[16];
      ptx::tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + c0, r);
      ptx::tmem_ld_wait();
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = lin_act(__uint_as_float(r[i]) + (a.bias ? a.bias[c0 + i] : 0.f), a.act);
      if (yt) {  // a warp writes 32 consecutive frames of one output row per store: 64-byte segments
#pragma unroll
        for (int i = 0; i < 16; ++i) yt[(size_t)(c0 + i) * a.t_ld] = __float2bfloat16_rn(v[i]);
      }
      if (n < a.N) {
        if (a.y_f) {
          float* dst = a.y_f + n * a.ldy_f + c0;
          if (vec_f) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) dst[i] = v[i];
          }
        }
        if (y_b) {
          __nv_bfloat16* dst = y_b + n * a.ldy_b + c0;
          if (vec_b) {
            uint32_t pk[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
              pk[i] = *reinterpret_cast<uint32_t*>(&h);
            }
            *reinterpret_cast<uint4*>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4*>(dst + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) dst[i] = __float2bfloat16_rn(v[i]);
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem, a.tmem_cols);
}

}  // namespace svsk

using namespace svsk;

extern "C" int svsk_linear_bf16(const svsk_linear_bf16_params* pp, void* stream) {
  SVSK_REQUIRE(pp != nullptr, SVSK_E_ARG, "linear_bf16: null params");
  const svsk_linear_bf16_params& p = *pp;
  SVSK_REQUIRE(p.a && p.w && (p.y_bf16 || p.y_f32), SVSK_E_ARG, "linear_bf16: null tensor");
  SVSK_REQUIRE(p.N > 0 && p.N < (1ll << 31), SVSK_E_ARG, "linear_bf16: N=%lld", (long long)p.N);
  SVSK_REQUIRE(p.K > 0 && p.K % 8 == 0, SVSK_E_ARG, "linear_bf16: K=%d must be a multiple of 8", p.K);
  SVSK_REQUIRE(p.Cout >= 16 && p.Cout <= 256 && p.Cout % 16 == 0, SVSK_E_ARG,
               "linear_bf16: Cout=%d must be a multiple of 16 in [16,256]", p.Cout);
  SVSK_REQUIRE(p.lda >= p.K && p.lda % 8 == 0, SVSK_E_ALIGN, "linear_bf16: lda=%d", p.lda);
  SVSK_REQUIRE(!p.y_bf16 || p.ldy_b >= p.Cout, SVSK_E_ARG, "linear_bf16: ldy_b");
  SVSK_REQUIRE(!p.y_f32 || p.ldy_f >= p.Cout, SVSK_E_ARG, "linear_bf16: ldy_f");
  int rc = require_sm100();
  if (rc) return rc;

  CUtensorMap tm_a, tm_w;
  {
    uint64_t dims[2] = {(uint64_t)p.K, (uint64_t)p.N};
    uint64_t str[1] = {(uint64_t)p.lda * 2};
    uint32_t box[2] = {64, 128};
    if ((rc = make_tmap_bf16(&tm_a, p.a, 2, dims, str, box))) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)p.K, (uint64_t)p.Cout};
    uint64_t str[1] = {(uint64_t)p.K * 2};
    uint32_t box[2] = {64, (uint32_t)p.Cout};
    if ((rc = make_tmap_bf16(&tm_w, p.w, 2, dims, str, box))) return rc;
  }
  const int stage_bytes = kLinABytes + p.Cout * 128;
  const int smem_bytes = kLinStages * stage_bytes + (int)sizeof(LinearBarriers) + 1024;
  int dev = 0;
  cudaGetDevice(&dev);
  static bool attr_set[64] = {false};
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(linear_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) return fail((int)e, "linear_bf16: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  LinearArgs a;
  a.bias = p.bias;
  a.y_b = (__nv_bfloat16*)p.y_bf16;
  a.y_f = p.y_f32;
  a.N = p.N;
  a.K = p.K;
  a.Cout = p.Cout;
  a.ldy_b = p.ldy_b;
  a.ldy_f = p.ldy_f;
  a.act = p.act;
  a.tmem_cols = p.Cout <= 32 ? 32 : (p.Cout <= 64 ? 64 : (p.Cout <= 128 ? 128 : 256));
  a.nstages = kLinStages;
  a.y_t = nullptr;
  a.t_bstride = 0;
  a.t_rows = 1;
  a.t_ld = 0;
  a.y_chunk_stride = 0;
  unsigned grid = (unsigned)((p.N + 127) / 128);
  linear_bf16_kernel<<<grid, 192, smem_bytes, as_stream(stream)>>>(tm_a, tm_w, a);
  return check_launch("linear_bf16");
}

extern "C" int svsk_usfgan_aux_frames(const void* cin, const void* w, void* q, int B, int Tf, int Ap, int R, int q_ld, int q_fpad,
                                      void* stream) {
  SVSK_REQUIRE(cin && w && q, SVSK_E_ARG, "usfgan_aux_frames: null tensor");
  SVSK_REQUIRE(B > 0 && Tf > 0 && (long long)B * Tf < (1ll << 31), SVSK_E_ARG, "usfgan_aux_frames: B=%d Tf=%d", B, Tf);
  SVSK_REQUIRE(Ap > 0 && Ap % 8 == 0, SVSK_E_ALIGN, "usfgan_aux_frames: Ap=%d must be a multiple of 8", Ap);
  SVSK_REQUIRE(R > 0 && R % 16 == 0, SVSK_E_ARG, "usfgan_aux_frames: R=%d must be a multiple of 16", R);
  SVSK_REQUIRE(q_fpad >= 0 && q_ld >= q_fpad + Tf, SVSK_E_ARG, "usfgan_aux_frames: q_ld=%d < q_fpad + Tf = %d", q_ld, q_fpad + Tf);
  int rc = require_sm100();
  if (rc) return rc;
  int dev = 0;
  cudaGetDevice(&dev);
  static bool attr_set[64] = {false};
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(linear_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) return fail((int)e, "usfgan_aux_frames: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  const long long N = (long long)B * Tf;
  CUtensorMap tm_a;
  {
    uint64_t dims[2] = {(uint64_t)Ap, (uint64_t)N};
    uint64_t str[1] = {(uint64_t)Ap * 2};
    uint32_t box[2] = {64, 128};
    if ((rc = make_tmap_bf16(&tm_a, cin, 2, dims, str, box))) return rc;
  }
  // rows = frames of all tracks; `chunk` output rows of w per CTA (gridDim.y chunks: ONE launch for all blocks' projections
  // when R is a multiple of 128 — 55 chunks at the recipe's depth; 28 launches of 256 rows took 0.47 ms at config 3, mostly
  // launch boundaries), stored transposed: frame index innermost
  const int chunk = R % 256 == 0 ? 256 : (R % 128 == 0 ? 128 : 0);
  for (int r0 = 0; r0 < R; r0 += 256) {
    const int cout = chunk ? chunk : (R - r0 < 256 ? R - r0 : 256);
    const int rows = chunk ? R : cout;   // weight rows the map covers
    CUtensorMap tm_w;
    uint64_t dims[2] = {(uint64_t)Ap, (uint64_t)rows};
    uint64_t str[1] = {(uint64_t)Ap * 2};
    uint32_t box[2] = {64, (uint32_t)cout};
    if ((rc = make_tmap_bf16(&tm_w, static_cast<const __nv_bfloat16*>(w) + (size_t)r0 * Ap, 2, dims, str, box))) return rc;
    LinearArgs a;
    a.bias = nullptr;
    a.y_b = nullptr;
    a.y_f = nullptr;
    a.N = N;
    a.K = Ap;
    a.Cout = cout;
    a.ldy_b = a.ldy_f = 0;
    a.act = SVSK_ACT_NONE;
    a.tmem_cols = cout <= 32 ? 32 : (cout <= 64 ? 64 : (cout <= 128 ? 128 : 256));
    a.y_t = static_cast<__nv_bfloat16*>(q) + (size_t)r0 * q_ld + q_fpad;
    a.t_bstride = (long long)R * q_ld;
    a.t_rows = Tf;
    a.t_ld = q_ld;
    a.y_chunk_stride = 0;
    // K = Ap is one or two k-blocks: as many stages as k-blocks, so that three or four CTAs fit an SM — the kernel is a
    // latency chain per CTA (load, a few MMAs, 2-byte transposed stores), 0.45 -> 0.2 ms at config 3
    a.nstages = (Ap + 63) / 64 < kLinStages ? (Ap + 63) / 64 : kLinStages;
    const int smem_bytes = a.nstages * (kLinABytes + cout * 128) + (int)sizeof(LinearBarriers) + 1024;
    const unsigned gy = chunk ? (unsigned)(R / chunk) : 1u;
    SVSK_REQUIRE(gy <= 65535u, SVSK_E_ARG, "usfgan_aux_frames: R=%d too large", R);
    linear_bf16_kernel<<<dim3((unsigned)((N + 127) / 128), gy), 192, smem_bytes, as_stream(stream)>>>(tm_a, tm_w, a);
    if ((rc = check_launch("usfgan_aux_frames"))) return rc;
    if (chunk) break;
  }
  return 0;
}

// conditioner_projection(cond) of all layers of a DiffNet (denoiser.py:59) in ONE launch, for a sampling run that reuses
// it across its K denoiser calls: p[blk][n][0..255] = cond[n][:] . wcp[blk*256 + r][:], blk = layer * (2C/256) + output block
// (gridDim.y chunks of 256 weight rows; two ring stages so that two CTAs share an SM).  svsk_diffnet_pcond_pack_bf16 then
// lays the result out for svsk_diffnet_stack_bf16.
extern "C" int svsk_diffnet_cond_project_bf16(const void* cond, const void* wcp, void* p, long long N, int H, int nblk,
                                              void* stream) {
  SVSK_REQUIRE(cond && wcp && p, SVSK_E_ARG, "diffnet_cond_project_bf16: null tensor");
  SVSK_REQUIRE(N > 0 && N < (1ll << 31), SVSK_E_ARG, "diffnet_cond_project_bf16: N=%lld", N);
  SVSK_REQUIRE(H > 0 && H % 64 == 0 && H <= 512, SVSK_E_ARG, "diffnet_cond_project_bf16: H=%d (need a multiple of 64, at most 512)", H);
  SVSK_REQUIRE(nblk >= 1 && nblk <= 65535, SVSK_E_ARG, "diffnet_cond_project_bf16: nblk=%d", nblk);
  SVSK_REQUIRE(((uintptr_t)p % 16) == 0, SVSK_E_ALIGN, "diffnet_cond_project_bf16: p must be 16-byte aligned");
  int rc = require_sm100();
  if (rc) return rc;
  int dev = 0;
  cudaGetDevice(&dev);
  static bool attr_set[64] = {false};
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(linear_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) return fail((int)e, "diffnet_cond_project_bf16: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  CUtensorMap tm_a, tm_w;
  {
    uint64_t dims[2] = {(uint64_t)H, (uint64_t)N};
    uint64_t str[1] = {(uint64_t)H * 2};
    uint32_t box[2] = {64, 128};
    if ((rc = make_tmap_bf16(&tm_a, cond, 2, dims, str, box))) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)H, (uint64_t)nblk * 256};
    uint64_t str[1] = {(uint64_t)H * 2};
    uint32_t box[2] = {64, 256};
    if ((rc = make_tmap_bf16(&tm_w, wcp, 2, dims, str, box))) return rc;
  }
  LinearArgs a;
  a.bias = nullptr;
  a.y_b = static_cast<__nv_bfloat16*>(p);
  a.y_f = nullptr;
  a.N = N;
  a.K = H;
  a.Cout = 256;
  a.ldy_b = 256;
  a.ldy_f = 0;
  a.act = SVSK_ACT_NONE;
  a.tmem_cols = 256;
  a.y_t = nullptr;
  a.t_bstride = 0;
  a.t_rows = 1;
  a.t_ld = 0;
  a.y_chunk_stride = N * 256;
  a.nstages = 2;
  const int smem_bytes = a.nstages * (kLinABytes + 256 * 128) + (int)sizeof(LinearBarriers) + 1024;
  linear_bf16_kernel<<<dim3((unsigned)((N + 127) / 128), (unsigned)nblk), 192, smem_bytes, as_stream(stream)>>>(tm_a, tm_w, a);
  return check_launch("diffnet_cond_project_bf16");
}
