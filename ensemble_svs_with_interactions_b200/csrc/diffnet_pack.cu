// Weight packing for the DiffNet tensor-core kernels (diffnet_block3 / diffnet_stack / training): the 2C output rows of
// the dilated conv (nnsvs/diffsinger/denoiser.py:33-35,45-52) are permuted so that each 256-row block holds the gate
// rows and then the filter rows of the same 128 channels — the epilogue thread that owns TMEM lane r reads the gate and
// the filter of one channel from two column ranges of its own lane.  K order: [tap -d | tap 0 | tap +d | conditioner].
#include <cuda_bf16.h>

#include "svsk_common.cuh"

namespace svsk {

// reference row r in [0,2C) -> packed row (gate rows of channel block q at 256q.., filter rows at 256q+128..)
static inline __host__ __device__ int packed_row(int r, int C) {
  if (r < C) return 256 * (r / 128) + (r % 128);
  const int c = r - C;
  return 256 * (c / 128) + 128 + (c % 128);
}

__global__ void diffnet_pack_kernel(const float* __restrict__ dw, const float* __restrict__ cw,
                                    const float* __restrict__ ow, __nv_bfloat16* __restrict__ w1p,
                                    __nv_bfloat16* __restrict__ woutp, int C, int H) {
  const int K1 = 3 * C + H;
  const int r = blockIdx.x;  // reference row
  const int pr = packed_row(r, C);
  for (int k = threadIdx.x; k < K1; k += blockDim.x) {
    float v;
    if (k < 3 * C) { int j = k / C, ci = k - j * C; v = dw[((size_t)r * C + ci) * 3 + j]; }
    else v = cw[(size_t)r * H + (k - 3 * C)];
    w1p[(size_t)pr * K1 + k] = __float2bfloat16_rn(v);
  }
  for (int k = threadIdx.x; k < C; k += blockDim.x) woutp[(size_t)r * C + k] = __float2bfloat16_rn(ow[(size_t)r * C + k]);
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
