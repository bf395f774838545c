// Bidirectional LSTM recurrence, one thread-block cluster per (track, direction) — the sequential part of the
// FFConvLSTM encoder in front of the denoiser (nnsvs/model.py:861-868,917-919).
//
// The input half of the gates (W_ih x_t + b_ih + b_hh, all frames, both directions) is a plain GEMM done beforehand; this
// kernel runs   a_t = pre_t + W_hh h_{t-1};  c_t = s(f) c_{t-1} + s(i) tanh(g);  h_t = s(o) tanh(c_t)   (gate rows in
// torch order i, f, g, o) over the first `length` frames of each track, from the last of them for the reverse direction;
// frames beyond `length` are written as zeros (pack_padded_sequence / pad_packed_sequence semantics).
//
// Latency is everything here (T dependent steps), so nothing is re-read:
//  * the 4H x H recurrent matrix of one direction is split over the NC <= 8 CTAs of a cluster (H/NC units each, all four
//    gates of a unit on the same CTA) and lives in REGISTERS: thread (row, segment) keeps its <= 128 fp32 weights;
//  * h_{t-1} sits in shared memory of every CTA (two buffers, by step parity); a step is: broadcast-read h, 128 FMAs,
//    shuffle-reduce the segments, gate activation on the row's own thread, one __syncthreads, cell update on warp 0,
//    and the new h values go to all CTAs of the cluster as 16-byte st.async messages that complete a transaction
//    barrier in the receiver — no cluster-wide barrier per step;
//  * the gate pre-activations of step t+2 are fetched while step t computes.
#include <cuda_bf16.h>

#include "sm100_ptx.cuh"
#include "svsk_common.cuh"

namespace svsk {

constexpr int kLstmMaxThreads = 256;

struct LstmArgs {
  const float* pre;
  const float* w_hh;
  const int32_t* lengths;
  float* h_f32;
  __nv_bfloat16* h_bf16;
  long long pre_sb, pre_st, pre_sr, hf_sb, hf_st, hf_sr, hb_sb, hb_st;
  int B, T, H, ndir;
  int NC, U, S, seg_len;  // CTAs per cluster, units per CTA, segments per row, real elements per segment
};

__device__ __forceinline__ void st_async_v4(uint32_t cluster_addr, float a, float b, float c, float d, uint32_t cluster_mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(cluster_addr),
               "f"(a), "f"(b), "f"(c), "f"(d), "r"(cluster_mbar)
               : "memory");
}

__device__ __forceinline__ float sigmoid_exact(float x) { return 1.f / (1.f + expf(-x)); }

template <int SEGT>
__global__ void __launch_bounds__(kLstmMaxThreads, 1) lstm_recurrence_kernel(const LstmArgs a) {
  constexpr int kPitch = SEGT + 4;  // the second segment starts 4 banks off the first: both broadcasts in one wavefront
  __shared__ __align__(16) float hbuf[2][2 * kPitch];
  __shared__ __align__(16) float gates[4 * 32];
  __shared__ __align__(16) float stage[32];
  __shared__ __align__(8) uint64_t hbar[2];

  const int tid = threadIdx.x, lane = tid & 31;
  const int c = (int)ptx::cluster_ctarank();
  const int cluster_id = blockIdx.x / a.NC;
  const int b = cluster_id / a.ndir, d = cluster_id % a.ndir;
  const int U = a.U, S = a.S, H = a.H;
  const int seg = tid % S, lrow = tid / S;          // lrow = gate * U + unit
  const bool active = lrow < 4 * U;                  // the block is padded to whole warps
  const int gate = active ? lrow / U : 0, u = active ? lrow % U : 0;
  const int unit = c * U + u;                        // hidden unit of this row
  const int len = a.lengths ? min(max(a.lengths[b], 0), a.T) : a.T;

  // recurrent weights -> registers (zeros beyond the real segment)
  float w[SEGT];
  {
    const float* src = a.w_hh + ((size_t)d * 4 * H + (size_t)gate * H + unit) * H + seg * a.seg_len;
#pragma unroll
    for (int k = 0; k < SEGT; ++k) w[k] = (active && k < a.seg_len) ? src[k] : 0.f;
  }
  for (int i = tid; i < 2 * 2 * kPitch; i += blockDim.x) (&hbuf[0][0])[i] = 0.f;
  if (tid == 0) {
    ptx::mbar_init(&hbar[0], 1);
    ptx::mbar_init(&hbar[1], 1);
    ptx::fence_mbar_init();
  }
  __syncthreads();
  ptx::cluster_sync_all();  // every CTA's buffers and barriers exist before anybody sends

  const float* pre = a.pre + (size_t)b * a.pre_sb + ((size_t)d * 4 * H + (size_t)gate * H + unit) * a.pre_sr;
  const bool gate_thread = active && seg == 0;
  auto frame_of = [&](int step) { return d == 0 ? step : len - 1 - step; };
  float p0 = 0.f, p1 = 0.f;  // pre-activations of steps t, t+1
  if (gate_thread) {
    if (len > 0) p0 = pre[(size_t)frame_of(0) * a.pre_st];
    if (len > 1) p1 = pre[(size_t)frame_of(1) * a.pre_st];
  }
  float cell = 0.f;          // warp 0, lane = unit
  const uint32_t hbuf_addr = ptx::smem_u32(&hbuf[0][0]);
  const uint32_t hbar_addr = ptx::smem_u32(&hbar[0]);
  const int quads = U >> 2, msgs = a.NC * quads;

  for (int t = 0; t < len; ++t) {
    const int cur = t & 1, nxt = cur ^ 1;
    if (tid == 0 && t + 1 < len) ptx::mbar_arrive_expect_tx(&hbar[nxt], (uint32_t)H * 4u);
    float p2 = 0.f;
    if (gate_thread && t + 2 < len) p2 = pre[(size_t)frame_of(t + 2) * a.pre_st];
    if (t > 0) ptx::mbar_wait(&hbar[cur], ((t - 1) >> 1) & 1);   // h_{t-1} complete in hbuf[cur]

    const float* hs = &hbuf[cur][seg * kPitch];
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
#pragma unroll
    for (int k = 0; k < SEGT; k += 4) {
      const float4 hv = *reinterpret_cast<const float4*>(hs + k);
      acc0 = fmaf(w[k], hv.x, acc0);
      acc1 = fmaf(w[k + 1], hv.y, acc1);
      acc2 = fmaf(w[k + 2], hv.z, acc2);
      acc3 = fmaf(w[k + 3], hv.w, acc3);
    }
    float dot = (acc0 + acc1) + (acc2 + acc3);
    if (S == 2) dot += __shfl_xor_sync(0xffffffffu, dot, 1);
    if (gate_thread) {
      const float v = dot + p0;
      gates[gate * 32 + u] = (gate == 2) ? tanhf(v) : sigmoid_exact(v);
    }
    p0 = p1;
    p1 = p2;
    __syncthreads();

    if (tid < 32) {
      float h = 0.f;
      if (lane < U) {
        const float gi = gates[lane], gf = gates[32 + lane], gg = gates[64 + lane], go = gates[96 + lane];
        cell = fmaf(gf, cell, gi * gg);
        h = go * tanhf(cell);
        stage[lane] = h;
      }
      __syncwarp();
      if (t + 1 < len) {
        for (int m = lane; m < msgs; m += 32) {
          const int dst = m / quads, q = m - dst * quads;
          const float4 hv = *reinterpret_cast<const float4*>(&stage[4 * q]);
          const int ug = c * U + 4 * q;                              // first of the four units
          const int pos = (ug / a.seg_len) * kPitch + (ug % a.seg_len);
          st_async_v4(ptx::mapa(hbuf_addr + (uint32_t)(nxt * 2 * kPitch + pos) * 4u, (uint32_t)dst), hv.x, hv.y, hv.z, hv.w,
                      ptx::mapa(hbar_addr + (uint32_t)nxt * 8u, (uint32_t)dst));
        }
      }
      if (lane < U) {
        const int frame = frame_of(t);
        const int col = d * H + c * U + lane;
        if (a.h_f32) a.h_f32[(size_t)b * a.hf_sb + (size_t)frame * a.hf_st + (size_t)col * a.hf_sr] = h;
        if (a.h_bf16) a.h_bf16[(size_t)b * a.hb_sb + (size_t)frame * a.hb_st + col] = __float2bfloat16_rn(h);
      }
      __syncwarp();  // stage is rewritten next step
    }
  }

  // frames past the end of the packed sequence read as zeros
  for (int i = tid; i < (a.T - len) * U; i += blockDim.x) {
    const int frame = len + i / U, col = d * H + c * U + i % U;
    if (a.h_f32) a.h_f32[(size_t)b * a.hf_sb + (size_t)frame * a.hf_st + (size_t)col * a.hf_sr] = 0.f;
    if (a.h_bf16) a.h_bf16[(size_t)b * a.hb_sb + (size_t)frame * a.hb_st + col] = __float2bfloat16_rn(0.f);
  }
  ptx::cluster_sync_all();  // nobody leaves while a peer may still address its shared memory
}

template <int SEGT>
static int launch_lstm(const LstmArgs& a, int threads, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(a.B * a.ndir * a.NC));
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)a.NC;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, lstm_recurrence_kernel<SEGT>, a);
  if (e != cudaSuccess) return fail((int)e, "lstm_f32: launch: %s", cudaGetErrorString(e));
  return 0;
}

// H -> (CTAs per cluster, units per CTA, segments per row, segment length); false when H has no layout here.
static bool lstm_layout(int H, int* NC, int* U, int* S, int* seg_len) {
  if (H < 4 || H > 256 || H % 4) return false;
  for (int nc = 1; nc <= 8; nc *= 2) {
    if (H % (4 * nc)) continue;
    if (H / nc > 32) continue;
    *NC = nc;
    *U = H / nc;
    *S = H > 128 ? 2 : 1;
    if (H % (4 * *S)) return false;
    *seg_len = H / *S;
    return true;
  }
  return false;
}

}  // namespace svsk

using namespace svsk;

extern "C" int svsk_lstm_supported(int H) {
  int NC, U, S, seg;
  return lstm_layout(H, &NC, &U, &S, &seg) ? 1 : 0;
}

extern "C" int svsk_lstm_f32(const svsk_lstm_params* pp, void* stream) {
  SVSK_REQUIRE(pp != nullptr, SVSK_E_ARG, "lstm_f32: null params");
  const svsk_lstm_params& p = *pp;
  SVSK_REQUIRE(p.pre && p.w_hh && (p.h_f32 || p.h_bf16), SVSK_E_ARG, "lstm_f32: null tensor");
  SVSK_REQUIRE(p.B > 0 && p.T > 0 && (p.ndir == 1 || p.ndir == 2), SVSK_E_ARG, "lstm_f32: B=%d T=%d ndir=%d", p.B, p.T, p.ndir);
  LstmArgs a = {};
  SVSK_REQUIRE(lstm_layout(p.H, &a.NC, &a.U, &a.S, &a.seg_len), SVSK_E_ARG,
               "lstm_f32: hidden size %d has no cluster layout (need H <= 256, H %% 4 == 0 and H / 2^k <= 32 units per CTA)", p.H);
  if (int rc = require_sm100()) return rc;
  a.pre = p.pre; a.w_hh = p.w_hh; a.lengths = p.lengths; a.h_f32 = p.h_f32; a.h_bf16 = reinterpret_cast<__nv_bfloat16*>(p.h_bf16);
  a.pre_sb = p.pre_stride_b; a.pre_st = p.pre_stride_t; a.pre_sr = p.pre_stride_r;
  a.hf_sb = p.hf_stride_b; a.hf_st = p.hf_stride_t; a.hf_sr = p.hf_stride_c;
  a.hb_sb = p.hb_stride_b; a.hb_st = p.hb_stride_t;
  a.B = p.B; a.T = p.T; a.H = p.H; a.ndir = p.ndir;
  const int threads = (4 * a.U * a.S + 31) & ~31;
  SVSK_REQUIRE(threads <= kLstmMaxThreads, SVSK_E_ARG, "lstm_f32: %d threads", threads);
  cudaStream_t st = as_stream(stream);
  int rc;
  if (a.seg_len <= 8) rc = launch_lstm<8>(a, threads, st);
  else if (a.seg_len <= 16) rc = launch_lstm<16>(a, threads, st);
  else if (a.seg_len <= 32) rc = launch_lstm<32>(a, threads, st);
  else if (a.seg_len <= 64) rc = launch_lstm<64>(a, threads, st);
  else rc = launch_lstm<128>(a, threads, st);
  return rc;
}
