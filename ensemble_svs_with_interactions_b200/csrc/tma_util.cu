#include "tma_util.cuh"

#include <cstring>
#include <mutex>
#include <vector>

#include "svsk_common.cuh"

namespace svsk {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn resolve_encode() {
  static std::once_flag once;
  static EncodeTiledFn fn = nullptr;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct Key {
  const void* base;
  int rank;
  int dtype;
  uint64_t dims[3];
  uint64_t strides[2];
  uint32_t box[3];
  bool operator==(const Key& o) const { return std::memcmp(this, &o, sizeof(Key)) == 0; }
};
struct Entry {
  Key k;
  CUtensorMap m;
};

static int make_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, CUtensorMapDataType dtype) {
  SVSK_REQUIRE(rank == 2 || rank == 3, SVSK_E_ARG, "tmap: rank %d", rank);
  SVSK_REQUIRE(((uintptr_t)base % 16) == 0, SVSK_E_ALIGN, "tmap: base pointer not 16-byte aligned");
  for (int i = 0; i < rank - 1; ++i)
    SVSK_REQUIRE(strides_bytes[i] % 16 == 0, SVSK_E_ALIGN, "tmap: stride %d = %llu bytes is not a multiple of 16", i,
                 (unsigned long long)strides_bytes[i]);
  for (int i = 0; i < rank; ++i) SVSK_REQUIRE(box[i] >= 1 && box[i] <= 256, SVSK_E_ARG, "tmap: box[%d]=%u", i, box[i]);

  Key k;
  std::memset(&k, 0, sizeof(k));
  k.base = base;
  k.rank = rank;
  k.dtype = (int)dtype;
  for (int i = 0; i < rank; ++i) {
    k.dims[i] = dims[i];
    k.box[i] = box[i];
  }
  for (int i = 0; i < rank - 1; ++i) k.strides[i] = strides_bytes[i];

  static std::mutex mu;
  static std::vector<Entry> cache;
  {
    std::lock_guard<std::mutex> lk(mu);
    for (const Entry& e : cache)
      if (e.k == k) {
        *out = e.m;
        return 0;
      }
  }
  EncodeTiledFn enc = resolve_encode();
  SVSK_REQUIRE(enc != nullptr, SVSK_E_DRIVER, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t gdims[3] = {1, 1, 1};
  cuuint64_t gstr[2] = {0, 0};
  cuuint32_t gbox[3] = {1, 1, 1}, estr[3] = {1, 1, 1};
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
  }
  for (int i = 0; i < rank - 1; ++i) gstr[i] = strides_bytes[i];
  CUtensorMap m;
  CUresult r = enc(&m, dtype, (cuuint32_t)rank, const_cast<void*>(base), gdims, gstr, gbox,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SVSK_REQUIRE(r == CUDA_SUCCESS, SVSK_E_DRIVER, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  {
    std::lock_guard<std::mutex> lk(mu);
    if (cache.size() >= 4096) cache.clear();
    cache.push_back(Entry{k, m});
  }
  *out = m;
  return 0;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box) {
  return make_tmap(out, base, rank, dims, strides_bytes, box, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
}
int make_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box) {
  return make_tmap(out, base, rank, dims, strides_bytes, box, CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
}

}  // namespace svsk
