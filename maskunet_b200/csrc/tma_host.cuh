// Host-side TMA descriptor construction without linking libcuda: the driver entry point is fetched
// through the runtime (cudaGetDriverEntryPoint), so the library loads on machines with no driver and
// only fails -- loudly -- when a tcgen05 entry point is actually called.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <array>
#include <cstring>
#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace mu {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) {
    set_error("cuTensorMapEncodeTiled unavailable (cudaGetDriverEntryPoint: %s)", cudaGetErrorString(e));
    return nullptr;
  }
  fn = reinterpret_cast<PFN_encodeTiled>(p);
  return fn;
}

// ---- descriptor cache: a training step re-uses the same (pointer, shape) pairs every iteration --
// parameters, the caching allocator's recycled activation blocks -- so an encoded CUtensorMap is kept per key instead
// of re-encoding on every launch.  A tensor map is a pure function of its key (base pointer, dims, box, kind): a hit
// can never be stale, whatever the buffer holds now.  Mutex-guarded, process-wide (the device is part of the key only
// through the pointer: unified addressing makes device pointers unique per process), bounded: it is cleared when full.
struct TmapKey {
  std::array<uint64_t, 8> v;
  bool operator==(const TmapKey& o) const { return v == o.v; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = 0xcbf29ce484222325ull;
    for (uint64_t x : k.v) {
      h ^= x;
      h *= 0x100000001b3ull;
      h ^= h >> 29;
    }
    return (size_t)h;
  }
};
struct TmapCache {
  std::mutex mu;
  std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> map;
  uint64_t hits = 0, misses = 0;
};
TmapCache& tmap_cache();          // one instance for the library (capi.cu)
constexpr size_t kTmapCacheMax = 4096;

// look up / fill: returns true on a hit (map filled); on a miss the caller encodes and calls tmap_cache_put
inline bool tmap_cache_get(const TmapKey& key, CUtensorMap* out) {
  TmapCache& c = tmap_cache();
  std::lock_guard<std::mutex> lock(c.mu);
  auto it = c.map.find(key);
  if (it == c.map.end()) {
    ++c.misses;
    return false;
  }
  ++c.hits;
  std::memcpy(out, &it->second, sizeof(CUtensorMap));
  return true;
}
inline void tmap_cache_put(const TmapKey& key, const CUtensorMap* m) {
  TmapCache& c = tmap_cache();
  std::lock_guard<std::mutex> lock(c.mu);
  if (c.map.size() >= kTmapCacheMax) c.map.clear();
  std::memcpy(&c.map[key], m, sizeof(CUtensorMap));
}
inline TmapKey tmap_key(int kind, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t d3, uint64_t b0,
                        uint64_t b1) {
  return TmapKey{{(uint64_t)kind, (uint64_t)reinterpret_cast<uintptr_t>(base), d0, d1, d2, d3, b0, b1}};
}

// bf16 tensor [batch][rows][cols] (cols contiguous); box = [1][box_rows][64 cols], 128-byte swizzle.
// Out-of-bounds rows are zero-filled on load and dropped on store.
inline int make_tmap_bf16_3d(CUtensorMap* map, const void* base, int cols, int rows, int batch, int box_rows) {
  const TmapKey key = tmap_key(1, base, cols, rows, batch, 0, box_rows, 0);
  if (tmap_cache_get(key, map)) return 0;
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return MU_ERR_DRIVER;
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)cols * 2, (cuuint64_t)cols * 2 * (cuuint64_t)rows};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: CUresult %d (cols %d rows %d batch %d box_rows %d)", (int)r, cols, rows,
              batch, box_rows);
    return MU_ERR_DRIVER;
  }
  tmap_cache_put(key, map);
  return 0;
}

// bf16 tensor [batch][rows][16 cols]; box = [1][box_rows][8 cols] (16-byte rows), NO swizzle: a box lands as consecutive
// 8-row x 16-byte core matrices, the canonical no-swizzle K-major UMMA operand (SBO = 128 bytes).
inline int make_tmap_bf16_3d_w16(CUtensorMap* map, const void* base, int rows, int batch, int box_rows) {
  const TmapKey key = tmap_key(5, base, 16, rows, batch, 0, box_rows, 0);
  if (tmap_cache_get(key, map)) return 0;
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return MU_ERR_DRIVER;
  cuuint64_t dims[3] = {16, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t strides[2] = {32, (cuuint64_t)32 * (cuuint64_t)rows};
  cuuint32_t box[3] = {8, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(w16) failed: CUresult %d (rows %d batch %d box_rows %d)", (int)r, rows, batch, box_rows);
    return MU_ERR_DRIVER;
  }
  tmap_cache_put(key, map);
  return 0;
}

// fp32 tensor [batch][rows][cols]; box = [1][box_rows][32 cols] (128-byte rows), 128-byte swizzle.
// Used as the destination of TMA reduce-add stores (rows past `rows` are dropped).
inline int make_tmap_f32_3d(CUtensorMap* map, const void* base, int cols, int rows, int batch, int box_rows) {
  const TmapKey key = tmap_key(2, base, cols, rows, batch, 0, box_rows, 0);
  if (tmap_cache_get(key, map)) return 0;
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return MU_ERR_DRIVER;
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)cols * 4, (cuuint64_t)cols * 4 * (cuuint64_t)rows};
  cuuint32_t box[3] = {32, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(f32) failed: CUresult %d (cols %d rows %d batch %d box_rows %d)", (int)r, cols,
              rows, batch, box_rows);
    return MU_ERR_DRIVER;
  }
  tmap_cache_put(key, map);
  return 0;
}

// bf16 channels-last activation [B][H][W][C] as a 4-D tensor (C, W, H, B); box = [1][box_h][box_w][64 channels],
// 128-byte swizzle.  Coordinates may be negative / past the edge: those elements are zero-filled on load (the
// zero padding of a 3x3 convolution) and dropped on store.
inline int make_tmap_bf16_nhwc(CUtensorMap* map, const void* base, int C, int W, int H, int B, int box_w, int box_h) {
  const TmapKey key = tmap_key(3, base, C, W, H, B, box_w, box_h);
  if (tmap_cache_get(key, map)) return 0;
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return MU_ERR_DRIVER;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * W, (cuuint64_t)C * 2 * W * H};
  cuuint32_t box[4] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(nhwc) failed: CUresult %d (C %d W %d H %d B %d box %dx%d)", (int)r, C, W, H, B,
              box_w, box_h);
    return MU_ERR_DRIVER;
  }
  tmap_cache_put(key, map);
  return 0;
}

}  // namespace mu
