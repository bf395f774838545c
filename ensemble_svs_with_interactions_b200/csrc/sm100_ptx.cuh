// Thin inline-PTX wrappers for the sm_100a features libsvsk uses: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 (alloc / mma / commit / ld / fences), UMMA shared-memory and instruction descriptors.
// Bit layouts follow the PTX ISA "tcgen05" chapter (matrix descriptor / instruction descriptor tables).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdint>

namespace svsk {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Non-blocking probe (mbarrier.test_wait): lets the MMA issuer look at the NEXT stage's barrier before it issues the current
// stage's MMAs, so the probe's latency overlaps the issue instead of sitting on the serial chain.
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must end as a launch failure the host can see, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
#pragma unroll 1  // nvcc otherwise unrolls this spin loop ~50x at every call site (tens of KB of SASS, I-cache misses)
  for (uint32_t spin = 0; spin < (1u << 22); ++spin) {
    if (mbar_try_wait(bar, parity)) return;
  }
  __trap();
}

// Pure polling wait (mbarrier.test_wait never suspends the thread): for the single-thread roles on a pipeline's critical
// handshake path, where the wake-up latency of a suspended try_wait would add to every hop.
__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
#pragma unroll 1
  for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
    if (mbar_test(bar, parity)) return;
  }
  __trap();
}

// Whole-warp wait with ONE polling lane: 32 lanes spinning on try_wait hammer the shared-memory port that the tensor
// core needs for its operands (measured on the uSFGAN block kernel).  __syncwarp orders memory for the other lanes.
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity) {
  if ((threadIdx.x & 31) == 0) mbar_wait(bar, parity);
  __syncwarp();
}

// ---------------------------------------------------------------- proxies / fences
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// ---------------------------------------------------------------- TMEM allocation (one full warp calls these)
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ---------------------------------------------------------------- UMMA descriptors
// Shared-memory matrix descriptor, K-major operand stored as rows of 128 B (64 bf16) with the 128-byte swizzle
// (what a TMA box of {64, rows} with CU_TENSOR_MAP_SWIZZLE_128B writes).  8-row groups are 1024 B apart (SBO).
//   [0,14) start address >> 4 | [16,30) LBO >> 4 (unused for swizzled K-major; 1) | [32,46) SBO >> 4 = 64
//   [46,48) version = 1 (sm_100) | [61,64) layout type = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  uint64_t d = (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// The same descriptor as two 32-bit halves: only the low half depends on the address, so the MMA-issuing thread (whose
// scalar instructions sit on the tensor pipe's critical path, tcgen05.mma issue being blocking) computes
// lo = umma_desc_lo(base) once per operand tile and adds 2 per 32-byte K step; hi is the constant below.
constexpr uint32_t kUmmaDescHiSw128 = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO=64, version=1, SWIZZLE_128B
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t smem_addr) { return ((smem_addr & 0x3FFFF) >> 4) | (1u << 16); }

// Instruction descriptor for kind::f16, bf16 x bf16 -> fp32, both operands K-major.
//   [4,6) D format = 1 (F32) | [7,10) A format = 1 (BF16) | [10,13) B format = 1 (BF16)
//   [15] A major = 0 (K) | [16] B major = 0 (K) | [17,23) N >> 3 | [24,29) M >> 4
__device__ __forceinline__ uint32_t umma_idesc_bf16_f32(uint32_t M, uint32_t N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// D[tmem] (+)= A[smem] . B[smem]^T ; issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kUmmaDescHiSw128)
      : "memory");
}
__device__ __forceinline__ void umma2_bf16_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kUmmaDescHiSw128)
      : "memory");
}
// KSTEPS K-steps of one 64-wide k-block (the descriptor low words advance by 2 per 32 bytes) PLUS a non-blocking probe
// of another mbarrier whose result is consumed only after the MMAs have been issued.  Measured (tools/ubench_umma.py,
// "issue loop"): an mbarrier try_wait/test_wait whose result is needed right away costs the issuing thread ~160 cycles
// even when the phase completed long ago, and a loop of {2 waits, 4 MMAs} runs at 632 cycles per group whatever the MMA
// shape — slower than the tensor pipe needs for the four MMAs (256 / 512 cycles at N = 128 / 256).  Probing the NEXT
// ring entry's barrier here hides that latency behind the MMA issue.
#define SVSK_UMMA_X4_PROBE(NAME, CG)                                                                                    \
  __device__ __forceinline__ bool NAME(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,                  \
                                       uint32_t accumulate_first, int ksteps, uint64_t* probe_bar, uint32_t probe_parity) { \
    uint32_t ok;                                                                                                        \
    asm volatile(                                                                                                       \
        "{\n\t.reg .pred p, q, k2, k3, k4;\n\t.reg .b64 da, db;\n\t.reg .b32 al, bl;\n\t"                              \
        "mbarrier.test_wait.parity.shared::cta.b64 q, [%8], %9;\n\t"                                                    \
        "setp.ne.b32 p, %5, 0;\n\t"                                                                                     \
        "setp.gt.s32 k2, %7, 1;\n\t"                                                                                    \
        "setp.gt.s32 k3, %7, 2;\n\t"                                                                                    \
        "setp.gt.s32 k4, %7, 3;\n\t"                                                                                    \
        "mov.b64 da, {%2, %6};\n\t"                                                                                     \
        "mov.b64 db, {%3, %6};\n\t"                                                                                     \
        "tcgen05.mma.cta_group::" #CG ".kind::f16 [%1], da, db, %4, p;\n\t"                                             \
        "setp.eq.b32 p, %4, %4;\n\t"                                                                                    \
        "add.u32 al, %2, 2;\n\t"                                                                                        \
        "add.u32 bl, %3, 2;\n\t"                                                                                        \
        "mov.b64 da, {al, %6};\n\t"                                                                                     \
        "mov.b64 db, {bl, %6};\n\t"                                                                                     \
        "@k2 tcgen05.mma.cta_group::" #CG ".kind::f16 [%1], da, db, %4, p;\n\t"                                         \
        "add.u32 al, %2, 4;\n\t"                                                                                        \
        "add.u32 bl, %3, 4;\n\t"                                                                                        \
        "mov.b64 da, {al, %6};\n\t"                                                                                     \
        "mov.b64 db, {bl, %6};\n\t"                                                                                     \
        "@k3 tcgen05.mma.cta_group::" #CG ".kind::f16 [%1], da, db, %4, p;\n\t"                                         \
        "add.u32 al, %2, 6;\n\t"                                                                                        \
        "add.u32 bl, %3, 6;\n\t"                                                                                        \
        "mov.b64 da, {al, %6};\n\t"                                                                                     \
        "mov.b64 db, {bl, %6};\n\t"                                                                                     \
        "@k4 tcgen05.mma.cta_group::" #CG ".kind::f16 [%1], da, db, %4, p;\n\t"                                         \
        "selp.u32 %0, 1, 0, q;\n\t}\n"                                                                                 \
        : "=r"(ok)                                                                                                      \
        : "r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate_first), "r"(kUmmaDescHiSw128), "r"(ksteps),     \
          "r"(smem_u32(probe_bar)), "r"(probe_parity)                                                                   \
        : "memory");                                                                                                    \
    return ok != 0;                                                                                                     \
  }
SVSK_UMMA_X4_PROBE(umma_bf16_x4_probe, 1)
SVSK_UMMA_X4_PROBE(umma2_bf16_x4_probe, 2)
#undef SVSK_UMMA_X4_PROBE

// Arrive on an mbarrier once every previously issued tcgen05.mma of this thread has completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ---------------------------------------------------------------- TMEM -> registers
// 32 lanes x 16 consecutive fp32 columns: thread i of the warp gets lane (base_lane + i), columns c..c+15.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- math
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_approx(float x) { return fmaf(tanh_approx(0.5f * x), 0.5f, 0.5f); }
// Packed fp32 pairs (FADD2 / FMUL2 / FFMA2 on sm_100: one issue slot for two lanes' worth of IEEE arithmetic — the gating
// loops are issue-bound, not MUFU-bound; results are bit-identical to the scalar forms).
__device__ __forceinline__ uint64_t f2_pack(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
// z = sigmoid(g) * tanh(f) for two columns at once; same arithmetic as sigmoid_approx(g) * tanh_approx(f) per lane
__device__ __forceinline__ uint64_t f2_gate(uint64_t g, uint64_t f) {
  const uint64_t half = f2_pack(0.5f, 0.5f);
  float g0, g1, f0, f1;
  f2_unpack(f2_mul(g, half), g0, g1);
  f2_unpack(f, f0, f1);
  const uint64_t sg = f2_fma(f2_pack(tanh_approx(g0), tanh_approx(g1)), half, half);
  return f2_mul(sg, f2_pack(tanh_approx(f0), tanh_approx(f1)));
}


// ---------------------------------------------------------------- clusters / CTA pairs (cta_group::2)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_smem_addr` in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
// Remote arrive with the default (.release.cta) semantics, as CUTLASS's ClusterBarrier::arrive(cta_id) does.  The
// .release.cluster / .acquire.cluster forms compile to MEMBAR.ALL.GPU / CCTL.IVALL per use (measured: ~600 extra
// cycles per pipeline stage); the data these barriers order is moved by the async proxy (TMA, tcgen05), which is
// ordered by the mbarrier itself and by fence.proxy.async, not by generic-proxy release/acquire.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
#pragma unroll 1
  for (uint32_t spin = 0; spin < (1u << 22); ++spin) {
    if (mbar_try_wait_cluster(bar, parity)) return;
  }
  __trap();
}
// TMA loads of a CTA pair: data lands in THIS CTA's smem, completion bytes are counted on `mbar_cluster_addr`
// (a shared::cluster address, normally the leader CTA's barrier).
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* m, uint32_t mbar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(mbar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(void* dst, const CUtensorMap* m, uint32_t mbar_cluster_addr, int c0, int c1,
                                                int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(mbar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// One MMA over the CTA pair: M = 256 (128 rows of A and of D per CTA), B's N rows split half/half between the CTAs.
// Issued by ONE thread of the leader CTA; descriptors are the leader's, the peer uses the same smem offsets.
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all earlier MMAs of this thread are done) on the barrier at this smem offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit2_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
// 16-byte store into another CTA's shared memory (address from mapa)
__device__ __forceinline__ void st_cluster_v4(uint32_t cluster_addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(cluster_addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// Before signalling another CTA that generic-proxy stores into ITS shared memory are done and may be read by its tensor
// cores (async proxy): make them visible cluster-wide, then cross the proxy.
__device__ __forceinline__ void fence_proxy_async_cluster_release() {
  asm volatile("fence.acq_rel.cluster;\n\tfence.proxy.async.shared::cluster;" ::: "memory");
}
__device__ __forceinline__ void st_shared_v4(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(p)), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// ---------------------------------------------------------------- TMA stores / reductions (smem -> global), bulk groups
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// TMA load multicast to the CTAs of `cta_mask` (bit = rank in the cluster): the tile lands at the same CTA-relative
// shared-memory offset in each of them and each one's mbarrier (same offset) gets the complete_tx.
__device__ __forceinline__ void tma_load_3d_mc(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                               uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5}], "
      "[%2], %6;" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "h"(cta_mask)
      : "memory");
}
// global[tile] += smem[tile] (element type and shape come from the tensor map; fp32 here), performed at L2
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(const void* p) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_u32(p)));
  return v;
}
__device__ __forceinline__ uint2 ld_shared_v2(const void* p) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_shared_v4f(const void* p) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
  return v;
}
__device__ __forceinline__ void st_shared_v4f(void* p, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(p)), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

// byte offset of element (row, 16-byte chunk) inside a 128B-swizzled K-major tile whose base is 1024-aligned
__device__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t chunk16) {
  return row * 128u + ((chunk16 ^ (row & 7u)) << 4);
}

}  // namespace ptx
}  // namespace svsk
