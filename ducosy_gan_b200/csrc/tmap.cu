// see tmap.cuh
#include <cstring>
#include <mutex>
#include <unordered_map>

#include "tmap.cuh"

namespace ducosy {
namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static std::once_flag once;
  static EncodeTiledFn fn = nullptr;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct Key {
  uint64_t v[16];   // base, dtype|rank, gdim[5], gstr[4], box[5]
  bool operator==(const Key& o) const { return std::memcmp(v, o.v, sizeof(v)) == 0; }
};
struct KeyHash {
  size_t operator()(const Key& k) const {
    uint64_t h = 1469598103934665603ull;
    for (uint64_t x : k.v) {
      h ^= x;
      h *= 1099511628211ull;
      h ^= h >> 29;
    }
    return size_t(h);
  }
};

constexpr size_t kMaxEntries = 8192;   // a generator forward + backward uses a few hundred distinct descriptors
std::mutex g_mu;
std::unordered_map<Key, CUtensorMap, KeyHash> g_maps;

}  // namespace

int encode_tiled_cached(CUtensorMap* out, CUtensorMapDataType dt, int rank, const void* base, const cuuint64_t* gdim,
                        const cuuint64_t* gstr, const cuuint32_t* box, const char* who) {
  DUCOSY_CHECK(rank >= 2 && rank <= 5, DUCOSY_ERR_ARG, "%s: tensor map rank %d", who, rank);
  Key k{};
  k.v[0] = reinterpret_cast<uint64_t>(base);
  k.v[1] = (uint64_t(dt) << 8) | uint64_t(rank);
  for (int i = 0; i < rank; ++i) k.v[2 + i] = gdim[i];
  for (int i = 0; i + 1 < rank; ++i) k.v[7 + i] = gstr[i];
  for (int i = 0; i < rank; ++i) k.v[11 + i] = box[i];
  {
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_maps.find(k);
    if (it != g_maps.end()) {
      *out = it->second;
      return 0;
    }
  }
  EncodeTiledFn encode = encode_fn();
  DUCOSY_CHECK(encode != nullptr, DUCOSY_ERR_CUDA, "%s: cuTensorMapEncodeTiled is not available (no CUDA driver?)", who);
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUresult r = encode(out, dt, cuuint32_t(rank), const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DUCOSY_CHECK(r == CUDA_SUCCESS, DUCOSY_ERR_CUDA, "%s: cuTensorMapEncodeTiled failed with %d", who, int(r));
  std::lock_guard<std::mutex> lock(g_mu);
  if (g_maps.size() >= kMaxEntries) g_maps.clear();
  g_maps.emplace(k, *out);
  return 0;
}

}  // namespace ducosy
