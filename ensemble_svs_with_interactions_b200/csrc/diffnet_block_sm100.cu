// Fused DiffNet residual block on sm_100a: TMA -> tcgen05.mma (TMEM accumulators) -> gated epilogue -> second
// tcgen05 GEMM from shared memory -> residual / skip epilogue.  Replaces ResidualBlock.forward
// (nnsvs/diffsinger/denoiser.py:54-66) with ONE launch per layer.
//
// Orientation: output channels are the MMA M dimension (128-row blocks), time is N (a tile of NT frames, NT a
// multiple of 16 chosen on the host so that B*ceil(T/NT) tiles fill the 148 SMs in as few waves as possible),
// and the contraction K runs over [x(t-d) ; x(t) ; x(t+d) ; cond(t)] = 3C + H channels.  Activations are NTC
// bf16, so every tap is the same [T][C] tensor loaded at a row offset of -d / 0 / +d by TMA; rows outside [0,T)
// are zero-filled by the TMA unit, which IS the convolution's zero padding.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM alloc), warps 2-5 = epilogue
// (TMEM lane quarter = warp_idx % 4).  Weight rows are packed so that each pair of 128-row M blocks holds the
// gate rows and the matching filter rows of 128 channels: the epilogue thread that owns TMEM lane r reads gate and
// filter of the same channel from two column ranges of its own lane.  Pair p's gating overlaps pair p+1's MMAs;
// the residual epilogue of the second GEMM overlaps the skip half's MMAs.
#include <cuda_bf16.h>
#include <cstdlib>

#include "sm100_ptx.cuh"
#include "svsk_common.cuh"
#include "tma_util.cuh"

namespace svsk {

constexpr int kStageABytes = 256 * 128;  // 256 weight rows x 64 bf16 (two 128-row M blocks)
constexpr int kMaxStages = 6;
constexpr int kSmemLimit = 232448;       // 227 KB opt-in dynamic shared memory per CTA
constexpr int kTmemCols = 512;

struct DiffnetBlockArgs {
  float* x32;
  float* skip32;
  __nv_bfloat16* xb_out;
  const float* stepbias;
  const float* bout;
  int B, T, C, H, dilation, sb_stride, init_skip, write_x, NT, nstages;
  unsigned long long* dbg;  // optional [num_ctas][16] clock64 timeline (SVSK_DIFFNET_TIMELINE)
  int dbg_flags;            // ablation switches for profiling only
};

struct __align__(8) DiffnetBarriers {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t d1_full[2];
  uint64_t d2_full[2];
  uint64_t g_ready;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(192, 1)
diffnet_block_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_cond,
                     const __grid_constant__ CUtensorMap tm_w1, const __grid_constant__ CUtensorMap tm_wout,
                     const DiffnetBlockArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int NT = a.NT, C = a.C, H = a.H;
  const int stageB = NT * 128;
  const int stage_bytes = kStageABytes + stageB;
  const int CB = C / 64;            // 64-channel K blocks per tap
  const int KB1 = 3 * CB + H / 64;  // K blocks of the first GEMM
  const int KB2 = CB;               // K blocks of the second GEMM
  const int pairs = (2 * C) / 256;  // 256-row (gate block + filter block) pairs
  uint8_t* g_smem = smem + a.nstages * stage_bytes;  // G: KB2 blocks of [NT rows][64 ch] bf16, 128B-swizzled
  DiffnetBarriers* bars = reinterpret_cast<DiffnetBarriers*>(g_smem + KB2 * stageB);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t0 = blockIdx.x * NT, b = blockIdx.y;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_x);
    ptx::prefetch_tmap(&tm_cond);
    ptx::prefetch_tmap(&tm_w1);
    ptx::prefetch_tmap(&tm_wout);
    for (int i = 0; i < a.nstages; ++i) {
      ptx::mbar_init(&bars->full[i], 1);
      ptx::mbar_init(&bars->empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&bars->d1_full[i], 1);
      ptx::mbar_init(&bars->d2_full[i], 1);
    }
    ptx::mbar_init(&bars->g_ready, 128);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(&bars->tmem_base, kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  unsigned long long* dbg = a.dbg ? a.dbg + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 16 : nullptr;
#define SVSK_STAMP(i) do { if (dbg) dbg[i] = clock64(); } while (0)
  if (threadIdx.x == 0) SVSK_STAMP(0);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int p = 0; p < pairs; ++p) {
        for (int kb = 0; kb < KB1; ++kb) {
          ptx::mbar_wait(&bars->empty[s], ph ^ 1);
          uint8_t* As = smem + s * stage_bytes;
          uint8_t* Bs = As + kStageABytes;
          ptx::mbar_arrive_expect_tx(&bars->full[s], kStageABytes + stageB);
          ptx::tma_load_2d(As, &tm_w1, &bars->full[s], kb * 64, p * 256);
          if (kb < 3 * CB) {
            int j = kb / CB, cb = kb - j * CB;
            ptx::tma_load_3d(Bs, &tm_x, &bars->full[s], cb * 64, t0 + (j - 1) * a.dilation, b);
          } else {
            ptx::tma_load_3d(Bs, &tm_cond, &bars->full[s], (kb - 3 * CB) * 64, t0, b);
          }
          if (++s == a.nstages) { s = 0; ph ^= 1; }
        }
      }
      for (int p = 0; p < pairs; ++p) {
        for (int kb = 0; kb < KB2; ++kb) {
          ptx::mbar_wait(&bars->empty[s], ph ^ 1);
          uint8_t* As = smem + s * stage_bytes;
          ptx::mbar_arrive_expect_tx(&bars->full[s], kStageABytes);
          ptx::tma_load_2d(As, &tm_wout, &bars->full[s], kb * 64, p * 256);
          if (++s == a.nstages) { s = 0; ph ^= 1; }
        }
      }
      SVSK_STAMP(1);  // all loads issued
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    if (lane == 0) {
      const uint32_t idesc = ptx::umma_idesc_bf16_f32(128, NT);
      int s = 0;
      uint32_t ph = 0;
      for (int p = 0; p < pairs; ++p) {
        for (int kb = 0; kb < KB1; ++kb) {
          ptx::mbar_wait(&bars->full[s], ph);
          ptx::tc_fence_after();
          const uint32_t a0 = ptx::smem_u32(smem + s * stage_bytes);
          const uint32_t b0 = a0 + kStageABytes;
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              ptx::umma_bf16(tmem + p * 2 * NT + half * NT, ptx::umma_desc_k_sw128(a0 + half * 16384 + k4 * 32),
                             ptx::umma_desc_k_sw128(b0 + k4 * 32), idesc, (kb | k4) != 0);
            }
          }
          ptx::umma_commit(&bars->empty[s]);
          if (++s == a.nstages) { s = 0; ph ^= 1; }
        }
        ptx::umma_commit(&bars->d1_full[p]);
        SVSK_STAMP(2 + p);  // GEMM1 pair p issued
      }
      // second GEMM: A = Wout rows (two M blocks per stage), B = gated activations resident in smem
      ptx::mbar_wait(&bars->g_ready, 0);
      ptx::tc_fence_after();
      SVSK_STAMP(4);  // G ready seen by the MMA thread
      const uint32_t g0 = ptx::smem_u32(g_smem);
      for (int p = 0; p < pairs; ++p) {
        for (int kb = 0; kb < KB2; ++kb) {
          ptx::mbar_wait(&bars->full[s], ph);
          ptx::tc_fence_after();
          const uint32_t a0 = ptx::smem_u32(smem + s * stage_bytes);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              ptx::umma_bf16(tmem + (2 * p + half) * NT, ptx::umma_desc_k_sw128(a0 + half * 16384 + k4 * 32),
                             ptx::umma_desc_k_sw128(g0 + kb * stageB + k4 * 32), idesc, (kb | k4) != 0);
            }
          }
          ptx::umma_commit(&bars->empty[s]);
          if (++s == a.nstages) { s = 0; ph ^= 1; }
        }
        ptx::umma_commit(&bars->d2_full[p]);
        SVSK_STAMP(5 + p);  // GEMM2 pair p issued
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps (TMEM lane quarter q)
    const int q = warp & 3;
    const uint32_t tlane = (uint32_t)(q * 32) << 16;
    const int d = a.dilation, T = a.T;
    const float* sb = a.stepbias + (size_t)b * a.sb_stride;
    const int twoC = 2 * C;

    // ---- epilogue 1: bias + step embedding (masked at the sequence ends) + sigmoid*tanh -> G (bf16, smem)
    for (int p = 0; p < pairs; ++p) {
      const int prg = p * 256 + q * 32 + lane;  // packed gate row; filter row = prg + 128
      const float g_l = sb[prg], g_c = sb[twoC + prg], g_r = sb[2 * twoC + prg];
      const float f_l = sb[prg + 128], f_c = sb[twoC + prg + 128], f_r = sb[2 * twoC + prg + 128];
      const int kc = p * 128 + q * 32 + lane;  // gated channel = K index of the second GEMM
      uint8_t* gdst = g_smem + (kc >> 6) * stageB + ((kc & 7) << 1);
      const uint32_t chunk16 = (uint32_t)((kc & 63) >> 3);

      ptx::mbar_wait(&bars->d1_full[p], 0);
      ptx::tc_fence_after();
      if (warp == 2 && lane == 0) SVSK_STAMP(7 + 2 * p);   // D1 pair p complete
      for (int c0 = 0; c0 < NT; c0 += 16) {
        if (a.dbg_flags & 2) break;
        uint32_t rg[16], rf[16];
        ptx::tmem_ld16(tmem + tlane + p * 2 * NT + c0, rg);
        ptx::tmem_ld16(tmem + tlane + p * 2 * NT + NT + c0, rf);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int t = t0 + c0 + i;
          const bool has_l = (t - d) >= 0, has_r = (t + d) < T;
          float gv = __uint_as_float(rg[i]) + g_c + (has_l ? g_l : 0.f) + (has_r ? g_r : 0.f);
          float fv = __uint_as_float(rf[i]) + f_c + (has_l ? f_l : 0.f) + (has_r ? f_r : 0.f);
          float z = ptx::sigmoid_approx(gv) * ptx::tanh_approx(fv);
          *reinterpret_cast<__nv_bfloat16*>(gdst + ptx::sw128_offset((uint32_t)(c0 + i), chunk16)) =
              __float2bfloat16_rn(z);
        }
      }
      if (warp == 2 && lane == 0) SVSK_STAMP(8 + 2 * p);  // gating of pair p done
    }
    if (warp == 2 && lane == 0) SVSK_STAMP(11);  // gating done
    ptx::tc_fence_before();
    ptx::fence_proxy_async_smem();  // generic-proxy writes of G -> visible to the tensor core's async proxy
    ptx::mbar_arrive(&bars->g_ready);

    // ---- epilogue 2: residual rows -> x32 (in place) + bf16 copy ; skip rows -> skip32 (accumulate)
    for (int p = 0; p < pairs; ++p) {
      ptx::mbar_wait(&bars->d2_full[p], 0);
      ptx::tc_fence_after();
      if (warp == 2 && lane == 0) SVSK_STAMP(12 + p);  // D2 pair p complete
      for (int half = 0; half < 2; ++half) {
        if (a.dbg_flags & 1) break;
        const int mb = 2 * p + half;
        const int orow = mb * 128 + q * 32 + lane;
        const bool is_res = (mb * 128) < C;  // warp-uniform
        if (is_res && !a.write_x) continue;
        const int ch = is_res ? orow : orow - C;
        const float bo = a.bout[orow];
        for (int c0 = 0; c0 < NT; c0 += 16) {
          uint32_t r[16];
          ptx::tmem_ld16(tmem + tlane + mb * NT + c0, r);
          ptx::tmem_ld_wait();
          if (is_res) {
            float xv[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int t = t0 + c0 + i;
              xv[i] = (t < T) ? a.x32[((size_t)b * T + t) * C + ch] : 0.f;
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int t = t0 + c0 + i;
              if (t < T) {
                const size_t idx = ((size_t)b * T + t) * C + ch;
                const float v = (xv[i] + __uint_as_float(r[i]) + bo) * 0.70710678118654752f;
                a.x32[idx] = v;
                a.xb_out[idx] = __float2bfloat16_rn(v);
              }
            }
          } else if (a.init_skip) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int t = t0 + c0 + i;
              if (t < T) a.skip32[((size_t)b * T + t) * C + ch] = __uint_as_float(r[i]) + bo;
            }
          } else {
            float sv[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int t = t0 + c0 + i;
              sv[i] = (t < T) ? a.skip32[((size_t)b * T + t) * C + ch] : 0.f;
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int t = t0 + c0 + i;
              if (t < T) a.skip32[((size_t)b * T + t) * C + ch] = sv[i] + __uint_as_float(r[i]) + bo;
            }
          }
        }
      }
    }
  }

  if (warp == 2 && lane == 0) SVSK_STAMP(14);  // epilogue done
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem, kTmemCols);
  if (threadIdx.x == 0) SVSK_STAMP(15);
#undef SVSK_STAMP
}

// reference row r in [0,2C) -> packed row (gate rows of channel block q at 256q.., filter rows at 256q+128..)
static inline int packed_row(int r, int C) {
  if (r < C) return 256 * (r / 128) + (r % 128);
  int c = r - C;
  return 256 * (c / 128) + 128 + (c % 128);
}

__global__ void diffnet_pack_kernel(const float* __restrict__ dw, const float* __restrict__ cw,
                                    const float* __restrict__ ow, __nv_bfloat16* __restrict__ w1p,
                                    __nv_bfloat16* __restrict__ woutp, int C, int H) {
  const int K1 = 3 * C + H;
  const int r = blockIdx.x;  // reference row
  int pr;
  if (r < C) pr = 256 * (r / 128) + (r % 128);
  else { int c = r - C; pr = 256 * (c / 128) + 128 + (c % 128); }
  for (int k = threadIdx.x; k < K1; k += blockDim.x) {
    float v;
    if (k < 3 * C) { int j = k / C, ci = k - j * C; v = dw[((size_t)r * C + ci) * 3 + j]; }
    else v = cw[(size_t)r * H + (k - 3 * C)];
    w1p[(size_t)pr * K1 + k] = __float2bfloat16_rn(v);
  }
  for (int k = threadIdx.x; k < C; k += blockDim.x) woutp[(size_t)r * C + k] = __float2bfloat16_rn(ow[(size_t)r * C + k]);
}

static int choose_time_tile(int B, int T, int C, int num_sms) {
  const int max_nt = (C == 256) ? 128 : 128;
  int best = 0;
  long best_cost = 0;
  for (int nt = max_nt; nt >= 32; nt -= 16) {
    long tiles = (long)B * ((T + nt - 1) / nt);
    long waves = (tiles + num_sms - 1) / num_sms;
    long cost = waves * (nt + 24);  // +24: fixed per-tile overhead (pipeline fill, epilogue tail) in column units
    if (best == 0 || cost < best_cost) { best = nt; best_cost = cost; }
  }
  return best;
}

}  // namespace svsk

using namespace svsk;

extern "C" int svsk_diffnet_packed_row(int reference_row, int C) {
  if (C <= 0 || C % 128 != 0 || reference_row < 0 || reference_row >= 2 * C) return -1;
  return packed_row(reference_row, C);
}

extern "C" int svsk_diffnet_pack_block(const float* dilated_w, const float* cond_w, const float* out_w, void* w1p,
                                       void* woutp, int C, int H, void* stream) {
  SVSK_REQUIRE(dilated_w && cond_w && out_w && w1p && woutp, SVSK_E_ARG, "diffnet_pack_block: null");
  SVSK_REQUIRE(C > 0 && C % 128 == 0 && C <= 256 && H > 0 && H % 64 == 0, SVSK_E_ARG,
               "diffnet_pack_block: need C in {128,256}, H %% 64 == 0 (C=%d H=%d)", C, H);
  diffnet_pack_kernel<<<2 * C, 256, 0, as_stream(stream)>>>(dilated_w, cond_w, out_w, (__nv_bfloat16*)w1p,
                                                            (__nv_bfloat16*)woutp, C, H);
  return check_launch("diffnet_pack_block");
}

extern "C" int svsk_diffnet_block_bf16(const svsk_diffnet_block_params* pp, void* stream) {
  SVSK_REQUIRE(pp != nullptr, SVSK_E_ARG, "diffnet_block_bf16: null params");
  const svsk_diffnet_block_params& p = *pp;
  SVSK_REQUIRE(p.xb_in && p.xb_out && p.x32 && p.skip32 && p.cond && p.w1p && p.woutp && p.stepbias && p.bout,
               SVSK_E_ARG, "diffnet_block_bf16: null tensor");
  SVSK_REQUIRE(p.xb_in != p.xb_out, SVSK_E_ARG, "diffnet_block_bf16: xb_in and xb_out must differ (halo reads)");
  SVSK_REQUIRE(p.C == 128 || p.C == 256, SVSK_E_ARG, "diffnet_block_bf16: C=%d (need 128 or 256)", p.C);
  SVSK_REQUIRE(p.H > 0 && p.H % 64 == 0, SVSK_E_ARG, "diffnet_block_bf16: H=%d (need a multiple of 64)", p.H);
  SVSK_REQUIRE(p.B > 0 && p.B <= 65535 && p.T > 0 && p.dilation >= 1, SVSK_E_ARG, "diffnet_block_bf16: bad B/T/dilation");
  SVSK_REQUIRE(p.stepbias_batch_stride == 0 || p.stepbias_batch_stride >= 6 * p.C, SVSK_E_ARG,
               "diffnet_block_bf16: stepbias stride %d", p.stepbias_batch_stride);
  int rc = require_sm100();
  if (rc) return rc;

  int dev = 0, num_sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);

  int NT = p.time_tile ? p.time_tile : choose_time_tile(p.B, p.T, p.C, num_sms);
  SVSK_REQUIRE(NT >= 32 && NT <= 128 && NT % 16 == 0, SVSK_E_ARG, "diffnet_block_bf16: time_tile %d", NT);
  SVSK_REQUIRE((2 * p.C / 128) * NT <= kTmemCols, SVSK_E_ARG, "diffnet_block_bf16: tile does not fit TMEM");

  const int stage_bytes = kStageABytes + NT * 128;
  const int g_bytes = (p.C / 64) * NT * 128;
  const int fixed = g_bytes + (int)sizeof(DiffnetBarriers) + 1024 /*alignment slack*/;
  int nstages = (kSmemLimit - fixed) / stage_bytes;
  if (nstages > kMaxStages) nstages = kMaxStages;
  SVSK_REQUIRE(nstages >= 2, SVSK_E_ARG, "diffnet_block_bf16: not enough shared memory for 2 stages");
  const int smem_bytes = nstages * stage_bytes + fixed;

  CUtensorMap tm_x, tm_cond, tm_w1, tm_wout;
  {
    uint64_t dims[3] = {(uint64_t)p.C, (uint64_t)p.T, (uint64_t)p.B};
    uint64_t str[2] = {(uint64_t)p.C * 2, (uint64_t)p.T * p.C * 2};
    uint32_t box[3] = {64, (uint32_t)NT, 1};
    if ((rc = make_tmap_bf16(&tm_x, p.xb_in, 3, dims, str, box))) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)p.H, (uint64_t)p.T, (uint64_t)p.B};
    uint64_t str[2] = {(uint64_t)p.H * 2, (uint64_t)p.T * p.H * 2};
    uint32_t box[3] = {64, (uint32_t)NT, 1};
    if ((rc = make_tmap_bf16(&tm_cond, p.cond, 3, dims, str, box))) return rc;
  }
  {
    const uint64_t K1 = 3 * (uint64_t)p.C + p.H;
    uint64_t dims[2] = {K1, (uint64_t)2 * p.C};
    uint64_t str[1] = {K1 * 2};
    uint32_t box[2] = {64, 256};
    if ((rc = make_tmap_bf16(&tm_w1, p.w1p, 2, dims, str, box))) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)p.C, (uint64_t)2 * p.C};
    uint64_t str[1] = {(uint64_t)p.C * 2};
    uint32_t box[2] = {64, 256};
    if ((rc = make_tmap_bf16(&tm_wout, p.woutp, 2, dims, str, box))) return rc;
  }

  static bool attr_set[64] = {false};
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(diffnet_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
    if (e != cudaSuccess) return fail((int)e, "diffnet_block_bf16: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  DiffnetBlockArgs a;
  a.x32 = p.x32;
  a.skip32 = p.skip32;
  a.xb_out = (__nv_bfloat16*)p.xb_out;
  a.stepbias = p.stepbias;
  a.bout = p.bout;
  a.B = p.B; a.T = p.T; a.C = p.C; a.H = p.H;
  a.dilation = p.dilation;
  a.sb_stride = p.stepbias_batch_stride;
  a.init_skip = p.init_skip;
  a.write_x = p.write_x;
  a.NT = NT;
  a.nstages = nstages;
  a.dbg = nullptr;
  a.dbg_flags = 0;
  if (const char* e = getenv("SVSK_DIFFNET_TIMELINE")) a.dbg = reinterpret_cast<unsigned long long*>(strtoull(e, nullptr, 0));
  if (const char* e = getenv("SVSK_DIFFNET_ABLATE")) a.dbg_flags = atoi(e);
  dim3 grid(ceil_div(p.T, NT), p.B);
  diffnet_block_kernel<<<grid, 192, smem_bytes, as_stream(stream)>>>(tm_x, tm_cond, tm_w1, tm_wout, a);
  return check_launch("diffnet_block_bf16");
}
