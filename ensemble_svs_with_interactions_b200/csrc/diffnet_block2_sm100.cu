// Fused DiffNet residual block, CTA-pair version (tcgen05 cta_group::2): replaces ResidualBlock.forward
// (nnsvs/diffsinger/denoiser.py:54-66) with one launch per layer.
//
// Why a CTA pair: the single-CTA kernel (diffnet_block_sm100.cu) is bound by SHARED-MEMORY bandwidth, not by the
// tensor pipe — per MMA cycle it writes 117 B (TMA) and reads 149 B (operands) against 128 B/cycle (measured:
// profiles/r01_diffnet_block_timeline.md).  With cta_group::2 each SM stages only half of every weight tile and an
// MMA of N = 256 re-uses each activation row for 256 output channels: ~109 B/cycle.
//
// Orientation: TIME is the MMA M dimension (256 frames per pair, 128 TMEM lanes per CTA), output channels are N.
//   GEMM1 (per 128-channel block j):  D1_j[256 t][gate 128 | filter 128] = X[256 t][K] . W1p[256 rows of block j][K]^T
//           K = [x(t-d) ; x(t) ; x(t+d) ; cond(t)] = 3C + H, the taps being the same NTC tensor at row offsets -d/0/+d
//           (TMA zero-fills rows outside [0,T) = the conv's zero padding).  B rows: CTA0 stages the gate rows, CTA1
//           the filter rows of the block; every thread (= one frame) then sees gate and filter of a channel in its
//           own TMEM lane.
//   gating: + bias + step-embedding taps (masked where a tap falls outside the sequence) ; sigmoid*tanh ; bf16 ->
//           swizzled smem tile G (the A operand of GEMM2) with 16-byte stores.
//   GEMM2:  D2[256 t][512] = G[256 t][C] . Woutp^T, residual half then skip half.
//   epilogue 2: no global loads/stores from registers at all (the register RMW version spent 46k of 87k cycles
//           waiting on them): the old x tile is TMA-loaded into ring slots that GEMM2 leaves free, x' = (x+r)/sqrt2 is
//           written over it (bf16) and TMA-stored; the skip half is staged as fp32 in 32-column slabs and added to the
//           global skip sum by TMA reduce-add (cp.reduce.async.bulk .add.f32, performed at L2), or stored on layer 0.
//           TMA clips rows >= T.  The residual stream is carried in bf16 (x32 is not used by this kernel).
// Warps: 0 = TMA producer (both CTAs), 1 = MMA issuer (leader CTA) + TMEM alloc, 2..9 = epilogue (two warps per TMEM
// lane quarter, alternating 16-column chunks).
#include <cuda_bf16.h>
#include <cstdlib>

#include "sm100_ptx.cuh"
#include "svsk_common.cuh"
#include "tma_util.cuh"

namespace svsk {

constexpr int k2TileBytes = 128 * 128;     // 128 rows x 64 bf16
constexpr int k2StageBytes = 2 * k2TileBytes;  // A (activations) + B (this CTA's half of a 256-row weight block)
constexpr int k2MaxStages = 6;
constexpr int k2SmemLimit = 232448;
constexpr int k2TmemCols = 512;
constexpr int k2Threads = 320;

struct Diffnet2Args {
  float* x32;
  float* skip32;
  __nv_bfloat16* xb_out;
  const float* stepbias;
  const float* bout;
  int B, T, C, H, dilation, sb_stride, init_skip, write_x, nstages;
  unsigned long long* dbg;
  int dbg_flags;  // profiling ablations: 4 = skip MMAs, 8 = skip TMA loads after the first ring fill, 1 = skip epilogue-2 global I/O
};

struct __align__(8) Diffnet2Barriers {
  uint64_t full[k2MaxStages];       // this CTA's stage landed (own TMA bytes)
  uint64_t peer_full[k2MaxStages];  // leader only: the peer CTA's stage landed (forwarded by the peer's warp 1)
  uint64_t empty[k2MaxStages];
  uint64_t d1_full[2];
  uint64_t d2_full[2];
  uint64_t g_ready;
  uint64_t xold_full;  // old x tile (C/64 TMA loads) landed
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(k2Threads, 1)
diffnet_block2_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_cond,
                      const __grid_constant__ CUtensorMap tm_w1, const __grid_constant__ CUtensorMap tm_wout,
                      const __grid_constant__ CUtensorMap tm_xout, const __grid_constant__ CUtensorMap tm_skip,
                      const Diffnet2Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int C = a.C, H = a.H, T = a.T;
  const int CB = C / 64;
  const int KB1 = 3 * CB + H / 64;
  const int KB2 = CB;
  const int NB = (2 * C) / 256;  // 256-column output blocks of either GEMM
  const int twoC = 2 * C;
  uint8_t* g_smem = smem + a.nstages * k2StageBytes;
  float* sb_full = reinterpret_cast<float*>(g_smem + KB2 * k2TileBytes);
  float* sb_l = sb_full + twoC;
  float* sb_r = sb_l + twoC;
  float* bo_s = sb_r + twoC;
  Diffnet2Barriers* bars = reinterpret_cast<Diffnet2Barriers*>(bo_s + twoC);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const int b = blockIdx.y;
  const int t_cta0 = (blockIdx.x >> 1) * 256 + (int)rank * 128;  // first frame of this CTA's 128 TMEM lanes
  unsigned long long* dbg = a.dbg ? a.dbg + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 32 : nullptr;
#define SVSK_STAMP(i) do { if (dbg) dbg[i] = clock64(); } while (0)
  if (threadIdx.x == 0) SVSK_STAMP(0);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_x);
    ptx::prefetch_tmap(&tm_cond);
    ptx::prefetch_tmap(&tm_w1);
    ptx::prefetch_tmap(&tm_wout);
    ptx::prefetch_tmap(&tm_xout);
    ptx::prefetch_tmap(&tm_skip);
    for (int i = 0; i < a.nstages; ++i) {
      ptx::mbar_init(&bars->full[i], 1);       // own producer's arrive.expect_tx
      ptx::mbar_init(&bars->peer_full[i], 1);  // one remote arrive per use
      ptx::mbar_init(&bars->empty[i], 1);      // one multicast tcgen05.commit
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&bars->d1_full[i], 1);
      ptx::mbar_init(&bars->d2_full[i], 1);
    }
    ptx::mbar_init(&bars->g_ready, 2 * 256);  // every epilogue thread of both CTAs
    ptx::mbar_init(&bars->xold_full, CB);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc2(&bars->tmem_base, k2TmemCols);
    ptx::tmem_relinquish2();
  }
  // Programmatic dependent launch: everything above overlaps the previous kernel's tail; nothing produced by an
  // earlier kernel is read before this point.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (warp >= 2) {
    // per-column biases -> smem: sb_full = centre + left + right tap terms (what an interior frame gets)
    const float* sb = a.stepbias + (size_t)b * a.sb_stride;
    for (int i = threadIdx.x - 64; i < twoC; i += 256) {
      const float l = sb[i], c = sb[twoC + i], r = sb[2 * twoC + i];
      sb_full[i] = c + l + r;
      sb_l[i] = l;
      sb_r[i] = r;
      bo_s[i] = a.bout[i];
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      long long acc_pe = 0;
      for (int j = 0; j < NB; ++j) {
        for (int kb = 0; kb < KB1; ++kb) {
          const long long c_0 = dbg ? clock64() : 0ll;
          ptx::mbar_wait(&bars->empty[s], ph ^ 1);
          if (dbg) acc_pe += clock64() - c_0;
          uint8_t* As = smem + s * k2StageBytes;
          uint8_t* Bs = As + k2TileBytes;
          if ((a.dbg_flags & 8) && (j * KB1 + kb) >= a.nstages) {
            ptx::mbar_arrive(&bars->full[s]);
            if (++s == a.nstages) { s = 0; ph ^= 1; }
            continue;
          }
          ptx::mbar_arrive_expect_tx(&bars->full[s], k2StageBytes);
          if (kb < 3 * CB) {
            const int jt = kb / CB, cb = kb - jt * CB;
            ptx::tma_load_3d(As, &tm_x, &bars->full[s], cb * 64, t_cta0 + (jt - 1) * a.dilation, b);
          } else {
            ptx::tma_load_3d(As, &tm_cond, &bars->full[s], (kb - 3 * CB) * 64, t_cta0, b);
          }
          ptx::tma_load_2d(Bs, &tm_w1, &bars->full[s], kb * 64, j * 256 + (int)rank * 128);
          if (++s == a.nstages) { s = 0; ph ^= 1; }
        }
      }
      for (int j = 0; j < NB; ++j) {
        for (int kb = 0; kb < KB2; ++kb) {
          ptx::mbar_wait(&bars->empty[s], ph ^ 1);
          uint8_t* Bs = smem + s * k2StageBytes + k2TileBytes;
          if (a.dbg_flags & 8) {
            ptx::mbar_arrive(&bars->full[s]);
            if (++s == a.nstages) { s = 0; ph ^= 1; }
            continue;
          }
          ptx::mbar_arrive_expect_tx(&bars->full[s], k2TileBytes);
          ptx::tma_load_2d(Bs, &tm_wout, &bars->full[s], kb * 64, j * 256 + (int)rank * 128);
          const int n2 = j * KB2 + kb;  // GEMM2 only uses the B half of a slot: its A half takes x channel block n2
          if (a.write_x && n2 < CB) {
            ptx::mbar_arrive_expect_tx(&bars->xold_full, k2TileBytes);
            ptx::tma_load_3d(smem + s * k2StageBytes, &tm_x, &bars->xold_full, n2 * 64, t_cta0, b);
          }
          if (++s == a.nstages) { s = 0; ph ^= 1; }
        }
      }
      SVSK_STAMP(1);
      if (dbg) dbg[16] = acc_pe;  // producer: cycles waiting for free slots during GEMM1
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer: one thread of the leader CTA
    if (rank == 0 && lane == 0) {
      const uint32_t idesc = ptx::umma_idesc_bf16_f32(256, 256);
      const uint32_t ring_lo = ptx::umma_desc_lo(ptx::smem_u32(smem)), g_lo = ptx::umma_desc_lo(ptx::smem_u32(g_smem));
      int s = 0;
      uint32_t ph = 0;
      long long acc_wf = 0, acc_wp = 0, acc_is = 0;
      for (int j = 0; j < NB; ++j) {
        for (int kb = 0; kb < KB1; ++kb) {
          const long long c_0 = dbg ? clock64() : 0ll;
          ptx::mbar_wait(&bars->full[s], ph);
          const long long c_1 = dbg ? clock64() : 0ll;
          ptx::mbar_wait(&bars->peer_full[s], ph);
          const long long c_2 = dbg ? clock64() : 0ll;
          acc_wf += c_1 - c_0;
          acc_wp += c_2 - c_1;
          ptx::tc_fence_after();
          const uint32_t a_lo = ring_lo + s * (k2StageBytes >> 4), b_lo = a_lo + (k2TileBytes >> 4);
          if (!(a.dbg_flags & 4)) {
            ptx::umma2_bf16_lo(tmem + j * 256, a_lo, b_lo, idesc, kb != 0);
            ptx::umma2_bf16_lo(tmem + j * 256, a_lo + 2, b_lo + 2, idesc, 1);
            ptx::umma2_bf16_lo(tmem + j * 256, a_lo + 4, b_lo + 4, idesc, 1);
            ptx::umma2_bf16_lo(tmem + j * 256, a_lo + 6, b_lo + 6, idesc, 1);
          }
          ptx::umma_commit2_mc(&bars->empty[s], 3);
          if (dbg) acc_is += clock64() - c_2;
          if (++s == a.nstages) { s = 0; ph ^= 1; }
        }
        ptx::umma_commit2_mc(&bars->d1_full[j], 3);
        SVSK_STAMP(2 + j);
      }
      if (dbg) { dbg[17] = acc_wf; dbg[18] = acc_wp; dbg[19] = acc_is; }  // GEMM1: wait own stage / peer stage / issue+commit
      ptx::mbar_wait(&bars->g_ready, 0);
      ptx::tc_fence_after();
      SVSK_STAMP(4);
      for (int j = 0; j < NB; ++j) {
        for (int kb = 0; kb < KB2; ++kb) {
          ptx::mbar_wait(&bars->full[s], ph);
          ptx::mbar_wait(&bars->peer_full[s], ph);
          ptx::tc_fence_after();
          const uint32_t b_lo = ring_lo + s * (k2StageBytes >> 4) + (k2TileBytes >> 4);
          const uint32_t gk_lo = g_lo + kb * (k2TileBytes >> 4);
          if (!(a.dbg_flags & 4)) {
            ptx::umma2_bf16_lo(tmem + j * 256, gk_lo, b_lo, idesc, kb != 0);
            ptx::umma2_bf16_lo(tmem + j * 256, gk_lo + 2, b_lo + 2, idesc, 1);
            ptx::umma2_bf16_lo(tmem + j * 256, gk_lo + 4, b_lo + 4, idesc, 1);
            ptx::umma2_bf16_lo(tmem + j * 256, gk_lo + 6, b_lo + 6, idesc, 1);
          }
          ptx::umma_commit2_mc(&bars->empty[s], 3);
          if (++s == a.nstages) { s = 0; ph ^= 1; }
        }
        ptx::umma_commit2_mc(&bars->d2_full[j], 3);
        SVSK_STAMP(5 + j);
      }
    } else if (rank == 1 && lane == 0) {
      // peer CTA: forward "my stage s has landed" to the leader's peer_full[s].  (Having the peer's TMA count bytes
      // directly on the leader's barrier — the cta_group::2 TMA form — measured 2.5x slower: 19 B/cycle/SM.)
      int s = 0;
      uint32_t ph = 0;
      const int total = NB * (KB1 + KB2);
      for (int it = 0; it < total; ++it) {
        ptx::mbar_wait(&bars->full[s], ph);
        ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&bars->peer_full[s]), 0));
        if (++s == a.nstages) { s = 0; ph ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps (thread = one frame)
    const int q = warp & 3;           // TMEM lane quarter this warp may read
    const int sub = (warp - 2) >> 2;  // the two warps of a quarter alternate 16-column chunks
    const int row = q * 32 + lane;
    const int t = t_cta0 + row;
    const uint32_t tlane = (uint32_t)(q * 32) << 16;
    const bool has_l = (t - a.dilation) >= 0, has_r = (t + a.dilation) < T;
    const bool stamp = (warp == 2 && lane == 0);

    // ---- epilogue 1: gating -> G
    for (int j = 0; j < NB; ++j) {
      ptx::mbar_wait(&bars->d1_full[j], 0);
      ptx::tc_fence_after();
      if (stamp) SVSK_STAMP(7 + 2 * j);
      uint32_t rgb[2][16], rfb[2][16];
      ptx::tmem_ld16(tmem + tlane + j * 256 + 16 * sub, rgb[0]);
      ptx::tmem_ld16(tmem + tlane + j * 256 + 128 + 16 * sub, rfb[0]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c0 = 16 * (2 * i + sub);
        ptx::tmem_ld_wait();
        if (i + 1 < 4) {  // next chunk's TMEM loads fly while this chunk is gated
          ptx::tmem_ld16(tmem + tlane + j * 256 + c0 + 32, rgb[(i + 1) & 1]);
          ptx::tmem_ld16(tmem + tlane + j * 256 + 128 + c0 + 32, rfb[(i + 1) & 1]);
        }
        const uint32_t* rg = rgb[i & 1];
        const uint32_t* rf = rfb[i & 1];
        const int pg = j * 256 + c0, pf = pg + 128;
        float z[16];
#pragma unroll
        for (int e = 0; e < 16; e += 4) {
          const float4 bg = *reinterpret_cast<const float4*>(sb_full + pg + e);
          const float4 bf = *reinterpret_cast<const float4*>(sb_full + pf + e);
          float gv[4] = {__uint_as_float(rg[e]) + bg.x, __uint_as_float(rg[e + 1]) + bg.y,
                         __uint_as_float(rg[e + 2]) + bg.z, __uint_as_float(rg[e + 3]) + bg.w};
          float fv[4] = {__uint_as_float(rf[e]) + bf.x, __uint_as_float(rf[e + 1]) + bf.y,
                         __uint_as_float(rf[e + 2]) + bf.z, __uint_as_float(rf[e + 3]) + bf.w};
          if (!has_l) {
#pragma unroll
            for (int u = 0; u < 4; ++u) { gv[u] -= sb_l[pg + e + u]; fv[u] -= sb_l[pf + e + u]; }
          }
          if (!has_r) {
#pragma unroll
            for (int u = 0; u < 4; ++u) { gv[u] -= sb_r[pg + e + u]; fv[u] -= sb_r[pf + e + u]; }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) z[e + u] = ptx::sigmoid_approx(gv[u]) * ptx::tanh_approx(fv[u]);
        }
        const int kc0 = j * 128 + c0;  // first gated channel of the chunk = K index of GEMM2
        uint8_t* gk = g_smem + (kc0 >> 6) * k2TileBytes;
        const uint32_t ch16 = (uint32_t)((kc0 & 63) >> 3);
        ptx::st_shared_v4(gk + ptx::sw128_offset((uint32_t)row, ch16), ptx::pack_bf16(z[0], z[1]),
                          ptx::pack_bf16(z[2], z[3]), ptx::pack_bf16(z[4], z[5]), ptx::pack_bf16(z[6], z[7]));
        ptx::st_shared_v4(gk + ptx::sw128_offset((uint32_t)row, ch16 + 1), ptx::pack_bf16(z[8], z[9]),
                          ptx::pack_bf16(z[10], z[11]), ptx::pack_bf16(z[12], z[13]), ptx::pack_bf16(z[14], z[15]));
      }
      if (stamp) SVSK_STAMP(8 + 2 * j);
    }
    ptx::tc_fence_before();
    ptx::fence_proxy_async_smem();  // G (generic-proxy stores) -> visible to the tensor cores' async proxy
    if (rank == 0) ptx::mbar_arrive(&bars->g_ready);
    else ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&bars->g_ready), 0));
    if (stamp) SVSK_STAMP(11);

    // ---- epilogue 2: residual -> in-place over the old x tile -> TMA store ; skip -> fp32 slabs -> TMA reduce-add
    const bool elected = (warp == 2 && lane == 0);
    const int slot0 = (NB * KB1) % a.nstages;          // ring slot of GEMM2's first k-block
    const float s2 = 0.70710678118654752f;
    int skip_slab = 0;                                  // running index of 32-column skip slabs of this CTA
    for (int j = 0; j < NB; ++j) {
      ptx::mbar_wait(&bars->d2_full[j], 0);
      ptx::tc_fence_after();
      if (stamp) SVSK_STAMP(12 + j);
      const int res_cols = min(max(C - j * 256, 0), 256);   // residual columns in this 256-column block
      // residual part
      if (res_cols > 0 && a.write_x) {
        ptx::mbar_wait(&bars->xold_full, 0);
#pragma unroll 1
        for (int i = 0; i < res_cols / 32; ++i) {
          const int c0 = 16 * (2 * i + sub);
          const int oc0 = j * 256 + c0;  // output channel = residual channel
          uint32_t r[16];
          ptx::tmem_ld16(tmem + tlane + j * 256 + c0, r);
          ptx::tmem_ld_wait();
          uint8_t* xt = smem + ((slot0 + (oc0 >> 6)) % a.nstages) * k2StageBytes;  // x tile of channel block oc0/64
          const uint32_t ch16 = (uint32_t)((oc0 & 63) >> 3);
          uint8_t* p0 = xt + ptx::sw128_offset((uint32_t)row, ch16);
          uint8_t* p1 = xt + ptx::sw128_offset((uint32_t)row, ch16 + 1);
          const uint4 xa = ptx::ld_shared_v4(p0), xb = ptx::ld_shared_v4(p1);
          const uint32_t xo[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
          uint32_t o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float lo = (ptx::bf16_lo(xo[e]) + __uint_as_float(r[2 * e]) + bo_s[oc0 + 2 * e]) * s2;
            const float hi = (ptx::bf16_hi(xo[e]) + __uint_as_float(r[2 * e + 1]) + bo_s[oc0 + 2 * e + 1]) * s2;
            o[e] = ptx::pack_bf16(lo, hi);
          }
          ptx::st_shared_v4(p0, o[0], o[1], o[2], o[3]);
          ptx::st_shared_v4(p1, o[4], o[5], o[6], o[7]);
        }
        ptx::fence_proxy_async_smem();
        ptx::named_bar_sync(1, 256);
        if (elected) {
          for (int cb = j * 4; cb < j * 4 + res_cols / 64; ++cb)
            ptx::tma_store_3d(&tm_xout, smem + ((slot0 + cb) % a.nstages) * k2StageBytes, cb * 64, t_cta0, b);
          ptx::bulk_commit_group();
        }
      }
      // skip part: columns [res_cols, 256) of this block, 32 at a time (one 128-byte fp32 row per frame)
      if (res_cols < 256) {
        if (j != NB - 1) __trap();  // skip columns only live in the last block: all MMAs are done, ring B halves + G free
#pragma unroll 1
        for (int i = res_cols / 32; i < 8; i += 2, skip_slab += 2) {
          uint32_t r0[16], r1[16];
          ptx::tmem_ld16(tmem + tlane + j * 256 + 16 * (2 * i + sub), r0);
          ptx::tmem_ld16(tmem + tlane + j * 256 + 16 * (2 * i + 2 + sub), r1);
          ptx::tmem_ld_wait();
          uint8_t* slab[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {  // slab buffers: B halves of the ring slots first, then 16 KB pieces of G
            const int n = skip_slab + u;
            slab[u] = (n < a.nstages) ? smem + n * k2StageBytes + k2TileBytes : g_smem + (n - a.nstages) * k2TileBytes;
          }
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const uint32_t* r = u ? r1 : r0;
            const int oc0 = j * 256 + 16 * (2 * (i + u) + sub);
#pragma unroll
            for (int e = 0; e < 16; e += 4) {
              const float4 bo = *reinterpret_cast<const float4*>(bo_s + oc0 + e);
              ptx::st_shared_v4f(slab[u] + ptx::sw128_offset((uint32_t)row, (uint32_t)(sub * 4 + (e >> 2))),
                                 __uint_as_float(r[e]) + bo.x, __uint_as_float(r[e + 1]) + bo.y,
                                 __uint_as_float(r[e + 2]) + bo.z, __uint_as_float(r[e + 3]) + bo.w);
            }
          }
          ptx::fence_proxy_async_smem();
          ptx::named_bar_sync(1, 256);
          if (elected) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int ch0 = j * 256 + 32 * (i + u) - C;  // first skip channel of the slab
              if (a.init_skip) ptx::tma_store_3d(&tm_skip, slab[u], ch0, t_cta0, b);
              else ptx::tma_reduce_add_3d(&tm_skip, slab[u], ch0, t_cta0, b);
            }
            ptx::bulk_commit_group();
          }
        }
      }
    }
    if (elected) ptx::bulk_wait_read_all();  // smem may be released; global visibility comes with grid completion
    if (stamp) SVSK_STAMP(14);
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();  // the peer's smem / TMEM are in use by the leader's MMAs until here
  if (warp == 1) ptx::tmem_dealloc2(tmem, k2TmemCols);
  if (threadIdx.x == 0) SVSK_STAMP(15);
#undef SVSK_STAMP
}

}  // namespace svsk

using namespace svsk;

extern "C" int svsk_diffnet_block2_bf16(const svsk_diffnet_block_params* pp, void* stream) {
  SVSK_REQUIRE(pp != nullptr, SVSK_E_ARG, "diffnet_block2_bf16: null params");
  const svsk_diffnet_block_params& p = *pp;
  SVSK_REQUIRE(p.xb_in && p.xb_out && p.skip32 && p.cond && p.w1p && p.woutp && p.stepbias && p.bout, SVSK_E_ARG,
               "diffnet_block2_bf16: null tensor");
  SVSK_REQUIRE(p.xb_in != p.xb_out, SVSK_E_ARG, "diffnet_block2_bf16: xb_in and xb_out must differ (halo reads)");
  SVSK_REQUIRE(p.C == 128 || p.C == 256, SVSK_E_ARG, "diffnet_block2_bf16: C=%d (need 128 or 256)", p.C);
  SVSK_REQUIRE(p.H > 0 && p.H % 64 == 0, SVSK_E_ARG, "diffnet_block2_bf16: H=%d (need a multiple of 64)", p.H);
  SVSK_REQUIRE(p.B > 0 && p.B <= 65535 && p.T > 0 && p.dilation >= 1, SVSK_E_ARG, "diffnet_block2_bf16: bad B/T/dilation");
  SVSK_REQUIRE(p.stepbias_batch_stride == 0 || p.stepbias_batch_stride >= 6 * p.C, SVSK_E_ARG,
               "diffnet_block2_bf16: stepbias stride %d", p.stepbias_batch_stride);
  SVSK_REQUIRE(((uintptr_t)p.skip32 % 16) == 0 && ((uintptr_t)p.xb_out % 16) == 0, SVSK_E_ALIGN,
               "diffnet_block2_bf16: skip32 / xb_out must be 16-byte aligned");
  int rc = require_sm100();
  if (rc) return rc;

  const int g_bytes = (p.C / 64) * k2TileBytes;
  const int fixed = g_bytes + 4 * 2 * p.C * (int)sizeof(float) + (int)sizeof(Diffnet2Barriers) + 1024;
  int nstages = (k2SmemLimit - fixed) / k2StageBytes;
  if (nstages > k2MaxStages) nstages = k2MaxStages;
  SVSK_REQUIRE(nstages >= 2, SVSK_E_ARG, "diffnet_block2_bf16: not enough shared memory");
  const int smem_bytes = nstages * k2StageBytes + fixed;

  CUtensorMap tm_x, tm_cond, tm_w1, tm_wout, tm_xout, tm_skip;
  {
    uint64_t dims[3] = {(uint64_t)p.C, (uint64_t)p.T, (uint64_t)p.B};
    uint64_t str[2] = {(uint64_t)p.C * 2, (uint64_t)p.T * p.C * 2};
    uint32_t box[3] = {64, 128, 1};
    if ((rc = make_tmap_bf16(&tm_x, p.xb_in, 3, dims, str, box))) return rc;
    if ((rc = make_tmap_bf16(&tm_xout, p.xb_out, 3, dims, str, box))) return rc;
    uint64_t str4[2] = {(uint64_t)p.C * 4, (uint64_t)p.T * p.C * 4};
    uint32_t box4[3] = {32, 128, 1};
    if ((rc = make_tmap_f32(&tm_skip, p.skip32, 3, dims, str4, box4))) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)p.H, (uint64_t)p.T, (uint64_t)p.B};
    uint64_t str[2] = {(uint64_t)p.H * 2, (uint64_t)p.T * p.H * 2};
    uint32_t box[3] = {64, 128, 1};
    if ((rc = make_tmap_bf16(&tm_cond, p.cond, 3, dims, str, box))) return rc;
  }
  {
    const uint64_t K1 = 3 * (uint64_t)p.C + p.H;
    uint64_t dims[2] = {K1, (uint64_t)2 * p.C};
    uint64_t str[1] = {K1 * 2};
    uint32_t box[2] = {64, 128};
    if ((rc = make_tmap_bf16(&tm_w1, p.w1p, 2, dims, str, box))) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)p.C, (uint64_t)2 * p.C};
    uint64_t str[1] = {(uint64_t)p.C * 2};
    uint32_t box[2] = {64, 128};
    if ((rc = make_tmap_bf16(&tm_wout, p.woutp, 2, dims, str, box))) return rc;
  }

  int dev = 0;
  cudaGetDevice(&dev);
  static bool attr_set[64] = {false};
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(diffnet_block2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, k2SmemLimit);
    if (e != cudaSuccess) return fail((int)e, "diffnet_block2_bf16: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  Diffnet2Args a;
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
  a.nstages = nstages;
  a.dbg = nullptr;
  a.dbg_flags = 0;
  if (const char* e = getenv("SVSK_DIFFNET_TIMELINE")) a.dbg = reinterpret_cast<unsigned long long*>(strtoull(e, nullptr, 0));
  if (const char* e = getenv("SVSK_DIFFNET_ABLATE")) a.dbg_flags = atoi(e);

  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * ceil_div(p.T, 256), p.B);
  cfg.blockDim = dim3(k2Threads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = as_stream(stream);
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = getenv("SVSK_NO_PDL") ? 1 : 2;
  cudaError_t e = cudaLaunchKernelEx(&cfg, diffnet_block2_kernel, tm_x, tm_cond, tm_w1, tm_wout, tm_xout, tm_skip, a);
  if (e != cudaSuccess) return fail((int)e, "diffnet_block2_bf16: launch: %s", cudaGetErrorString(e));
  return check_launch("diffnet_block2_bf16");
}
