// Host-side TMA tensor-map encoding (cuTensorMapEncodeTiled resolved at run time so the library links and loads on a
// GPU-less build box without libcuda.so), plus a per-thread memo of encoded maps.
#include <cstring>
#include <map>
#include <vector>

#include "tcgen05.cuh"

namespace rmem {

namespace {
#ifdef RMEM_OPERAND_BF16
constexpr CUtensorMapDataType kTmaType = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
#else
constexpr CUtensorMapDataType kTmaType = CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
#endif

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn resolve_encode() {
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
      set_error("cuTensorMapEncodeTiled entry point unavailable: %s", cudaGetErrorString(e));
      return nullptr;
    }
    encode = reinterpret_cast<EncodeFn>(fn);
  }
  return encode;
}

struct MapKey {
  uint64_t v[16];
  bool operator<(const MapKey& o) const { return memcmp(v, o.v, sizeof(v)) < 0; }
};
}  // namespace

int tma_encode(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
               const uint32_t* box, const uint32_t* estr, int swizzle128) {
  RMEM_REQUIRE(rank >= 2 && rank <= 5, "tma_encode: rank %d", rank);
  EncodeFn encode = resolve_encode();
  if (!encode) return RMEM_ERR_CUDA;
  cuuint64_t d[5], s[4];
  cuuint32_t b[5], e[5];
  for (int i = 0; i < rank; ++i) {
    d[i] = dims[i];
    b[i] = box[i];
    e[i] = estr ? estr[i] : 1;
    if (i < rank - 1) s[i] = strides_bytes[i];
  }
  CUresult r = encode(map, kTmaType, (cuuint32_t)rank, const_cast<void*>(base), d, s, b, e,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: CUresult %d (rank=%d dims=%llu,%llu,%llu stride0=%llu box=%u,%u,%u)", (int)r,
              rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)strides_bytes[0], box[0], box[1],
              rank > 2 ? box[2] : 0);
    return RMEM_ERR_CUDA;
  }
  return RMEM_OK;
}

int tma_encode_cached(const CUtensorMap** out, const void* base, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* estr, int swizzle128) {
  // std::map nodes are stable, so the returned pointer stays valid for the thread's lifetime.
  static thread_local std::map<MapKey, CUtensorMap> cache;
  MapKey k;
  memset(&k, 0, sizeof(k));
  k.v[0] = reinterpret_cast<uint64_t>(base);
  k.v[1] = (uint64_t)rank | ((uint64_t)(swizzle128 ? 1 : 0) << 8);
  for (int i = 0; i < rank; ++i) {
    k.v[2 + i] = dims[i];
    k.v[7 + i] = ((uint64_t)box[i] << 32) | (estr ? estr[i] : 1);
    if (i < rank - 1) k.v[12 + i] = strides_bytes[i];
  }
  auto it = cache.find(k);
  if (it == cache.end()) {
    if (cache.size() > 8192) cache.clear();
    CUtensorMap m;
    RMEM_TRY(tma_encode(&m, base, rank, dims, strides_bytes, box, estr, swizzle128));
    it = cache.emplace(k, m).first;
  }
  *out = &it->second;
  return RMEM_OK;
}

}  // namespace rmem
