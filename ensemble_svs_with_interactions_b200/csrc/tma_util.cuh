// Host-side construction of TMA tensor maps (bf16, 128-byte swizzle, zero OOB fill) with a small cache.
// cuTensorMapEncodeTiled is resolved through cudaGetDriverEntryPoint so libsvsk.so does not link libcuda
// (it has to load on GPU-less build machines for the symbol-export tests).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace svsk {

// dims/strides innermost first; rank 2 or 3; box_inner is always 64 bf16 (= 128 B swizzle span).
// strides_bytes[i] is the byte stride of dimension i+1.  Returns 0 or an SVSK_E* / cudaError code.
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box);
// same for fp32 tensors (box_inner 32 floats = 128 B swizzle span)
int make_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box);

}  // namespace svsk
