// Fused uSFGAN / QPPWG residual block, CTA-pair version (tcgen05 cta_group::2) — same contract as usfgan_block_sm100.cu
// (FixedBlock / AdaptiveBlock forward, nnsvs/usfgan/layers/residual_block.py:123-157,198-234 + pd_indexing index.py:12-54).
//
// Why a pair: the single-CTA kernel is bound by its ONE MMA-issuing thread — per 128-sample tile it spends ~2.4k cycles
// in barrier waits / fences / commits around only ~1.9k cycles of tensor-pipe work (profiles/r01_usfgan_block_role_
// accounting.log).  With cta_group::2 one issuing thread drives 256 samples per instruction (M = 256: 128 TMEM lanes in each
// CTA), each CTA stages only its half of the weights (N split: CTA0 the tanh rows, CTA1 the sigmoid rows; 44 KB instead
// of 88 KB resident), and the freed shared memory deepens the activation ring to 8 slots.
//
// Warps (15): 0 TMA producer | 1 MMA issuer (leader CTA) + TMEM | 2 forwarder ("my slot s landed" -> leader's ready[s])
//             3-6 gather producers (cp.async rows for adaptive taps / reflected boundary tiles) | 7-14 epilogue.
#include <cuda_bf16.h>
#include <cstdlib>

#include "sm100_ptx.cuh"
#include "svsk_common.cuh"
#include "tma_util.cuh"

namespace svsk {

constexpr int kVTile = 128 * 128;   // activation slot: 128 rows x 64 bf16
constexpr int kVWTile = 64 * 128;   // this CTA's half of one W1 k-block: 64 rows x 64 bf16
constexpr int kVMaxStages = 8;
constexpr int kVThreads = 480;

struct Usfgan2Args {
  const __nv_bfloat16* xb_in;
  const float* bias1;
  const float* bout;
  const int32_t* idx_past;
  const int32_t* idx_future;
  int B, T, A, dilation, adaptive, nstages, akb, last_ksteps, tiles_per_row, total_tiles;  // tiles = 256-row pair tiles
  float out_scale;
  int out_relu;
  unsigned long long* dbg;
  int dbg_flags;
};

struct __align__(8) Usfgan2Barriers {
  uint64_t full[kVMaxStages];   // this CTA's slot landed: 1 producer arrival (+ TMA bytes) + 128 gather-thread arrivals
  uint64_t ready[kVMaxStages];  // leader only: both CTAs' slot s landed (2 forwarder arrivals)
  uint64_t empty[kVMaxStages];  // one multicast tcgen05.commit
  uint64_t d1_full[2], g_full[2], d2_full[2];
  uint64_t w_full;
  uint32_t tmem_base;
};

__device__ __forceinline__ void v2_cp_async_16(void* dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(ptx::smem_u32(dst)), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void v2_cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(ptx::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void v2_mbar_arrive_n(uint64_t* bar, uint32_t n) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(ptx::smem_u32(bar)), "r"(n) : "memory");
}
__device__ __forceinline__ bool v2_tile_needs_gather(int t0, int T, int d, int adaptive) {
  if (adaptive) return true;
  const int last = min(t0 + 127, T - 1);
  return (t0 - d < 0) || (last + d >= T);
}

template <bool kProf>
__global__ void __launch_bounds__(kVThreads, 1)
usfgan_block2_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_aux,
                     const __grid_constant__ CUtensorMap tm_w1, const __grid_constant__ CUtensorMap tm_wout,
                     const __grid_constant__ CUtensorMap tm_xout, const Usfgan2Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int KB = 3 + a.akb;
  uint8_t* w1_s = smem;                       // KB half-tiles [64 rows][64]
  uint8_t* wout_s = w1_s + KB * kVWTile;      // [32 rows][64] = 4 KB
  uint8_t* ring = wout_s + 4096;
  uint8_t* gbuf = ring + a.nstages * kVTile;  // 3 x 16 KB rotating: G then the output tile of tile n % 3
  float* bias_s = reinterpret_cast<float*>(gbuf + 3 * kVTile);
  Usfgan2Barriers* bars = reinterpret_cast<Usfgan2Barriers*>(bias_s + 192);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = a.T;
  const uint32_t rank = ptx::cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_x);
    ptx::prefetch_tmap(&tm_aux);
    ptx::prefetch_tmap(&tm_w1);
    ptx::prefetch_tmap(&tm_wout);
    ptx::prefetch_tmap(&tm_xout);
    for (int i = 0; i < a.nstages; ++i) {
      ptx::mbar_init(&bars->full[i], 129);
      ptx::mbar_init(&bars->ready[i], 2);
      ptx::mbar_init(&bars->empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&bars->d1_full[i], 1);
      ptx::mbar_init(&bars->g_full[i], 16);  // 8 epilogue warps x 2 CTAs
      ptx::mbar_init(&bars->d2_full[i], 1);
    }
    ptx::mbar_init(&bars->w_full, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc2(&bars->tmem_base, 512);
    ptx::tmem_relinquish2();
  }
  for (int i = threadIdx.x; i < 192; i += kVThreads) bias_s[i] = i < 128 ? a.bias1[i] : a.bout[i - 128];
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs, own rows / own weight half)
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(&bars->w_full, KB * kVWTile + 4096);
      for (int kb = 0; kb < KB; ++kb) ptx::tma_load_2d(w1_s + kb * kVWTile, &tm_w1, &bars->w_full, kb * 64, (int)rank * 64);
      ptx::tma_load_2d(wout_s, &tm_wout, &bars->w_full, 0, (int)rank * 32);
      int s = 0;
      uint32_t ph = 0;
      for (int tile = pair; tile < a.total_tiles; tile += npairs) {
        const int b = tile / a.tiles_per_row, t0 = (tile - b * a.tiles_per_row) * 256 + (int)rank * 128;
        const bool gather = v2_tile_needs_gather(t0, T, a.dilation, a.adaptive);
        for (int kb = 0; kb < KB; ++kb) {
          ptx::mbar_wait(&bars->empty[s], ph ^ 1);
          uint8_t* slot = ring + s * kVTile;
          if (kb == 1) {
            ptx::mbar_arrive_expect_tx(&bars->full[s], kVTile);
            ptx::tma_load_3d(slot, &tm_x, &bars->full[s], 0, t0, b);
          } else if (kb >= 3) {
            ptx::mbar_arrive_expect_tx(&bars->full[s], kVTile);
            ptx::tma_load_3d(slot, &tm_aux, &bars->full[s], (kb - 3) * 64, t0, b);
          } else if (!gather) {
            ptx::mbar_arrive_expect_tx(&bars->full[s], kVTile);
            ptx::tma_load_3d(slot, &tm_x, &bars->full[s], 0, t0 + (kb - 1) * a.dilation, b);
          } else {
            ptx::mbar_arrive(&bars->full[s]);  // the gather warps bring the data
          }
          if (++s == a.nstages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer: one thread of the leader CTA
    if (rank == 0 && lane == 0) {
      const uint32_t idesc1 = ptx::umma_idesc_bf16_f32(256, 128);
      const uint32_t idesc2 = ptx::umma_idesc_bf16_f32(256, 64);
      ptx::mbar_wait(&bars->w_full, 0);  // the peer's weights: covered by its forwarder (it waits w_full first)
      ptx::tc_fence_after();
      const uint32_t w1a = ptx::smem_u32(w1_s), woa = ptx::smem_u32(wout_s), ga = ptx::smem_u32(gbuf);
      int s = 0;
      uint32_t ph = 0;
      int n_issued = 0;
      for (int tile = pair;; tile += npairs) {
        const bool have = tile < a.total_tiles;
        if (have) {
          const int p = n_issued & 1;
          for (int kb = 0; kb < KB; ++kb) {
            ptx::mbar_wait(&bars->ready[s], ph);
            ptx::fence_proxy_async_smem();
            ptx::tc_fence_after();
            const uint32_t a0 = ptx::smem_u32(ring + s * kVTile);
            const int ks = (kb == KB - 1) ? a.last_ksteps : 4;
            for (int k4 = 0; k4 < ks; ++k4)
              ptx::umma2_bf16(tmem + p * 128, ptx::umma_desc_k_sw128(a0 + k4 * 32),
                              ptx::umma_desc_k_sw128(w1a + kb * kVWTile + k4 * 32), idesc1, (kb | k4) != 0);
            ptx::umma_commit2_mc(&bars->empty[s], 3);
            if (++s == a.nstages) { s = 0; ph ^= 1; }
          }
          ptx::umma_commit2_mc(&bars->d1_full[p], 3);
        }
        if (n_issued > 0) {
          const int m = n_issued - 1, p = m & 1;
          ptx::mbar_wait(&bars->g_full[p], (m >> 1) & 1);
          ptx::tc_fence_after();
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            ptx::umma2_bf16(tmem + 256 + p * 64, ptx::umma_desc_k_sw128(ga + (m % 3) * kVTile + k4 * 32),
                            ptx::umma_desc_k_sw128(woa + k4 * 32), idesc2, k4 != 0);
          ptx::umma_commit2_mc(&bars->d2_full[p], 3);
        }
        if (!have) break;
        ++n_issued;
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ forwarder: my slot landed -> leader's ready[s]
    if (lane == 0) {
      ptx::mbar_wait(&bars->w_full, 0);  // ready[0] of the first tile also vouches for this CTA's resident weights
      int s = 0;
      uint32_t ph = 0;
      for (int tile = pair; tile < a.total_tiles; tile += npairs) {
        for (int kb = 0; kb < KB; ++kb) {
          ptx::mbar_wait(&bars->full[s], ph);
          ptx::fence_proxy_async_smem();  // cp.async rows (generic proxy) before the tensor core's async-proxy reads
          if (rank == 0) ptx::mbar_arrive(&bars->ready[s]);
          else ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&bars->ready[s]), 0));
          if (++s == a.nstages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp < 7) {
    // ------------------------------------------------------------------ gather producers (thread = row of the tile)
    const int r = threadIdx.x - 96;
    int s = 0;
    uint32_t ph = 0;
    for (int tile = pair; tile < a.total_tiles; tile += npairs) {
      const int b = tile / a.tiles_per_row, t0 = (tile - b * a.tiles_per_row) * 256 + (int)rank * 128;
      const bool gather = v2_tile_needs_gather(t0, T, a.dilation, a.adaptive);
      const int t = t0 + r;
      for (int kb = 0; kb < KB; ++kb) {
        ptx::mbar_wait_warp(&bars->empty[s], ph ^ 1);
        if (gather && (kb == 0 || kb == 2)) {
          int src = -1;
          if (t < T) {
            if (a.adaptive) {
              src = (kb == 0 ? a.idx_past : a.idx_future)[(size_t)b * T + t];
            } else {
              src = t + (kb - 1) * a.dilation;
              if (src < 0) src = -src;
              if (src >= T) src = 2 * (T - 1) - src;
            }
          }
          const bool ok = src >= 0 && src < T;
          const uint8_t* g = reinterpret_cast<const uint8_t*>(a.xb_in + ((size_t)b * T + (ok ? src : 0)) * 64);
          uint8_t* slot = ring + s * kVTile;
#pragma unroll
          for (int c = 0; c < 8; ++c)
            v2_cp_async_16(slot + ptx::sw128_offset((uint32_t)r, (uint32_t)c), g + c * 16, ok ? 16u : 0u);
          v2_cp_async_arrive_noinc(&bars->full[s]);
        } else if (lane == 0) {
          v2_mbar_arrive_n(&bars->full[s], 32);  // this warp's 32 arrivals for a slot the TMA producer fills
        }
        if (++s == a.nstages) { s = 0; ph ^= 1; }
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  } else {
    // ------------------------------------------------------------------ epilogue (thread = one sample, half the columns)
    const int q = warp & 3;
    const int sub = (warp - 7) >> 2;
    const int row = q * 32 + lane;
    const uint32_t tlane = (uint32_t)(q * 32) << 16;
    const bool elected = (warp == 7 && lane == 0);
    // Software-pipelined: iteration `it` gates tile it (so GEMM2(it) can be queued) and THEN finishes tile it-1, whose
    // GEMM2 ran behind GEMM1(it) in the in-order tensor pipe while this warp group was gating.
    const int my_tiles = a.total_tiles > pair ? (a.total_tiles - 1 - pair) / npairs + 1 : 0;
    long long acc_d1 = 0, acc_gate = 0, acc_d2 = 0, acc_e2 = 0, acc_sync = 0;
    uint4 xr_prev[2][2], xr_cur[2][2];
#pragma unroll
    for (int i = 0; i < 2; ++i) xr_prev[i][0] = xr_prev[i][1] = xr_cur[i][0] = xr_cur[i][1] = make_uint4(0, 0, 0, 0);
    for (int it = 0; it <= my_tiles; ++it) {
      if (it < my_tiles) {
        const int tile = pair + it * npairs;
        const int b = tile / a.tiles_per_row, t0 = (tile - b * a.tiles_per_row) * 256 + (int)rank * 128;
        const int t = t0 + row, p = it & 1;
        // residual row of tile `it`, prefetched now and consumed one iteration later
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int c0 = 16 * (2 * i + sub);
          if (t < T) {
            const uint4* src = reinterpret_cast<const uint4*>(a.xb_in + ((size_t)b * T + t) * 64 + c0);
            xr_cur[i][0] = __ldg(src);
            xr_cur[i][1] = __ldg(src + 1);
          } else {
            xr_cur[i][0] = xr_cur[i][1] = make_uint4(0, 0, 0, 0);
          }
        }
        uint8_t* gb = gbuf + (it % 3) * kVTile;
        long long c_0 = (kProf ? clock64() : 0ll);
        if (it >= 3) {  // buffer it%3 was the output tile of tile it-3: its TMA store must have finished reading
          if (elected) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          ptx::named_bar_sync(2, 256);
        }
        acc_sync += (kProf ? clock64() : 0ll) - c_0;
        c_0 = (kProf ? clock64() : 0ll);
        ptx::mbar_wait_warp(&bars->d1_full[p], (it >> 1) & 1);
        ptx::tc_fence_after();
        acc_d1 += (kProf ? clock64() : 0ll) - c_0;
        c_0 = (kProf ? clock64() : 0ll);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          if ((kProf ? a.dbg_flags : 0) & 1) break;
          const int c0 = 16 * (2 * i + sub);
          uint32_t ra[16], rb[16];
          ptx::tmem_ld16(tmem + tlane + p * 128 + c0, ra);
          ptx::tmem_ld16(tmem + tlane + p * 128 + 64 + c0, rb);
          ptx::tmem_ld_wait();
          uint32_t o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float z0 = ptx::tanh_approx(__uint_as_float(ra[2 * e]) + bias_s[c0 + 2 * e]) *
                             ptx::sigmoid_approx(__uint_as_float(rb[2 * e]) + bias_s[64 + c0 + 2 * e]);
            const float z1 = ptx::tanh_approx(__uint_as_float(ra[2 * e + 1]) + bias_s[c0 + 2 * e + 1]) *
                             ptx::sigmoid_approx(__uint_as_float(rb[2 * e + 1]) + bias_s[64 + c0 + 2 * e + 1]);
            o[e] = ptx::pack_bf16(z0, z1);
          }
          ptx::st_shared_v4(gb + ptx::sw128_offset((uint32_t)row, (uint32_t)(c0 >> 3)), o[0], o[1], o[2], o[3]);
          ptx::st_shared_v4(gb + ptx::sw128_offset((uint32_t)row, (uint32_t)(c0 >> 3) + 1), o[4], o[5], o[6], o[7]);
        }
        ptx::tc_fence_before();
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {  // one arrival per warp (count 32 folded into the barrier's expected count of 16 warps)
          if (rank == 0) ptx::mbar_arrive(&bars->g_full[p]);
          else ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&bars->g_full[p]), 0));
        }
        acc_gate += (kProf ? clock64() : 0ll) - c_0;
      }
      if (it >= 1) {
        const int m = it - 1, p = m & 1;
        const int tile = pair + m * npairs;
        const int b = tile / a.tiles_per_row, t0 = (tile - b * a.tiles_per_row) * 256 + (int)rank * 128;
        uint8_t* gb = gbuf + (m % 3) * kVTile;
        long long c_0 = (kProf ? clock64() : 0ll);
        ptx::mbar_wait_warp(&bars->d2_full[p], (m >> 1) & 1);
        ptx::tc_fence_after();
        acc_d2 += (kProf ? clock64() : 0ll) - c_0;
        c_0 = (kProf ? clock64() : 0ll);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          if ((kProf ? a.dbg_flags : 0) & 1) break;
          const int c0 = 16 * (2 * i + sub);
          uint32_t rd[16];
          ptx::tmem_ld16(tmem + tlane + 256 + p * 64 + c0, rd);
          ptx::tmem_ld_wait();
          const uint32_t xw[8] = {xr_prev[i][0].x, xr_prev[i][0].y, xr_prev[i][0].z, xr_prev[i][0].w,
                                  xr_prev[i][1].x, xr_prev[i][1].y, xr_prev[i][1].z, xr_prev[i][1].w};
          uint32_t o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float lo = (__uint_as_float(rd[2 * e]) + bias_s[128 + c0 + 2 * e] + ptx::bf16_lo(xw[e])) * a.out_scale;
            float hi = (__uint_as_float(rd[2 * e + 1]) + bias_s[128 + c0 + 2 * e + 1] + ptx::bf16_hi(xw[e])) * a.out_scale;
            if (a.out_relu) { lo = fmaxf(lo, 0.f); hi = fmaxf(hi, 0.f); }
            o[e] = ptx::pack_bf16(lo, hi);
          }
          ptx::st_shared_v4(gb + ptx::sw128_offset((uint32_t)row, (uint32_t)(c0 >> 3)), o[0], o[1], o[2], o[3]);
          ptx::st_shared_v4(gb + ptx::sw128_offset((uint32_t)row, (uint32_t)(c0 >> 3) + 1), o[4], o[5], o[6], o[7]);
        }
        ptx::tc_fence_before();
        ptx::fence_proxy_async_smem();
        acc_e2 += (kProf ? clock64() : 0ll) - c_0;
        c_0 = (kProf ? clock64() : 0ll);
        ptx::named_bar_sync(3, 256);
        if (elected) {
          ptx::tma_store_3d(&tm_xout, gb, 0, t0, b);
          ptx::bulk_commit_group();
        }
        acc_sync += (kProf ? clock64() : 0ll) - c_0;
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) { xr_prev[i][0] = xr_cur[i][0]; xr_prev[i][1] = xr_cur[i][1]; }
    }
    if (elected) ptx::bulk_wait_read_all();
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();  // the peer's smem / TMEM are used by the leader's MMAs until here
  if (warp == 1) ptx::tmem_dealloc2(tmem, 512);
}

}  // namespace svsk

using namespace svsk;

extern "C" int svsk_usfgan_block2_bf16(const svsk_usfgan_block_params* pp, void* stream) {
  SVSK_REQUIRE(pp != nullptr, SVSK_E_ARG, "usfgan_block2_bf16: null params");
  const svsk_usfgan_block_params& p = *pp;
  SVSK_REQUIRE(p.xb_in && p.xb_out && p.aux && p.w1p && p.woutp && p.bias1 && p.bout, SVSK_E_ARG,
               "usfgan_block2_bf16: null tensor");
  SVSK_REQUIRE(p.xb_in != p.xb_out, SVSK_E_ARG, "usfgan_block2_bf16: xb_in and xb_out must differ (taps read neighbours)");
  SVSK_REQUIRE(p.B > 0 && p.T > 0 && p.A >= 1 && p.A <= 320 && p.A % 8 == 0, SVSK_E_ARG,
               "usfgan_block2_bf16: bad shape B=%d T=%d A=%d (aux row pitch must be a multiple of 16 bytes)", p.B, p.T, p.A);
  if (p.adaptive) SVSK_REQUIRE(p.idx_past && p.idx_future, SVSK_E_ARG, "usfgan_block2_bf16: adaptive needs tap indices");
  else SVSK_REQUIRE(p.dilation >= 1 && p.dilation < p.T, SVSK_E_ARG,
                    "usfgan_block2_bf16: reflect padding needs T > dilation (T=%d, dilation=%d)", p.T, p.dilation);
  SVSK_REQUIRE((long long)p.B * ((p.T + 255) / 256) < (1ll << 30), SVSK_E_ARG, "usfgan_block2_bf16: too many tiles");
  int rc = require_sm100();
  if (rc) return rc;

  const int akb = (p.A + 63) / 64, KB = 3 + akb;
  const int K1p = 3 * 64 + akb * 64;
  const int fixed = KB * kVWTile + 4096 + 3 * kVTile + 192 * 4 + (int)sizeof(Usfgan2Barriers) + 1024;
  int nstages = (232448 - fixed) / kVTile;
  if (nstages > kVMaxStages) nstages = kVMaxStages;
  SVSK_REQUIRE(nstages >= 3, SVSK_E_ARG, "usfgan_block2_bf16: not enough shared memory (aux too wide)");
  const int smem_bytes = fixed + nstages * kVTile;

  CUtensorMap tm_x, tm_aux, tm_w1, tm_wout, tm_xout;
  {
    uint64_t dims[3] = {64, (uint64_t)p.T, (uint64_t)p.B};
    uint64_t str[2] = {128, (uint64_t)p.T * 128};
    uint32_t box[3] = {64, 128, 1};
    if ((rc = make_tmap_bf16(&tm_x, p.xb_in, 3, dims, str, box))) return rc;
    if ((rc = make_tmap_bf16(&tm_xout, p.xb_out, 3, dims, str, box))) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)p.A, (uint64_t)p.T, (uint64_t)p.B};
    uint64_t str[2] = {(uint64_t)p.A * 2, (uint64_t)p.T * p.A * 2};
    uint32_t box[3] = {64, 128, 1};
    if ((rc = make_tmap_bf16(&tm_aux, p.aux, 3, dims, str, box))) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)K1p, 128};
    uint64_t str[1] = {(uint64_t)K1p * 2};
    uint32_t box[2] = {64, 64};
    if ((rc = make_tmap_bf16(&tm_w1, p.w1p, 2, dims, str, box))) return rc;
  }
  {
    uint64_t dims[2] = {64, 64};
    uint64_t str[1] = {128};
    uint32_t box[2] = {64, 32};
    if ((rc = make_tmap_bf16(&tm_wout, p.woutp, 2, dims, str, box))) return rc;
  }
  int dev = 0, num_sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  static bool attr_set[64] = {false};
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(usfgan_block2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) return fail((int)e, "usfgan_block2_bf16: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  Usfgan2Args a;
  a.xb_in = (const __nv_bfloat16*)p.xb_in;
  a.bias1 = p.bias1;
  a.bout = p.bout;
  a.idx_past = p.idx_past;
  a.idx_future = p.idx_future;
  a.B = p.B; a.T = p.T; a.A = p.A;
  a.dilation = p.adaptive ? 0 : p.dilation;
  a.adaptive = p.adaptive;
  a.nstages = nstages;
  a.akb = akb;
  a.last_ksteps = (p.A - (akb - 1) * 64 + 15) / 16;
  a.tiles_per_row = (p.T + 255) / 256;
  a.total_tiles = p.B * a.tiles_per_row;
  a.out_scale = p.out_scale;
  a.out_relu = p.out_relu;
  a.dbg = nullptr;
  a.dbg_flags = 0;
  int pairs = num_sms / 2;
  if (a.total_tiles < pairs) pairs = a.total_tiles;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(kVThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = as_stream(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, usfgan_block2_kernel<false>, tm_x, tm_aux, tm_w1, tm_wout, tm_xout, a);
  if (e != cudaSuccess) return fail((int)e, "usfgan_block2_bf16: launch: %s", cudaGetErrorString(e));
  return check_launch("usfgan_block2_bf16");
}
