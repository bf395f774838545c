// Frame-rate aux projection of the uSFGAN blocks (shared by the block kernel and the kernel that builds the weights).
//
// The reference upsamples the aux features to sample rate (nnsvs/usfgan/layers/upsample.py:61-128: per stage a nearest
// stretch and a (1, 2s+1) smoothing filter with zero padding) and every block projects them with its bias-free
// conv1x1_aux (residual_block.py:74,100,139-142).  Both steps are linear and the upsampler treats every channel alike, so
//     conv1x1_aux(upsample(c))[t] = sum_f U[t, f] * Q[f],     Q = conv1x1_aux(c)  at FRAME rate,
// where U[t, f] is the upsampler's response at sample t to a unit impulse at frame f.  U[t, .] is non-zero only for
// frames within `reach` samples of t (reach = sum_i s_i * hop / (s_1 ... s_i)), so a 128-sample tile touches a handful
// of frames: the block kernel adds the projection with ONE K = 16 MMA per tile, A = U[tile] (128 x 16), B = Q for the 16
// frames starting at usfgan_frame_base(t0) — instead of streaming 80 sample-rate channels through 5 K-steps.
#pragma once

namespace svsk {

// First of the 16 frames a tile starting at sample t0 multiplies with: the first frame that can reach the tile, rounded
// down to a multiple of 8 (16-byte aligned rows of Q for cp.async).  May be negative (down to -q_fpad).
__host__ __device__ inline int usfgan_frame_base(int t0, int reach, int hop) {
  const int a = t0 - reach;
  const int f = a >= 0 ? a / hop : -((-a + hop - 1) / hop);
  return f & ~7;
}

// Last frame a tile can reach; the host checks usfgan_frame_last - usfgan_frame_base <= 15 for every tile.
__host__ __device__ inline int usfgan_frame_last(int t0, int reach, int hop) { return (t0 + 127 + reach) / hop; }

}  // namespace svsk
