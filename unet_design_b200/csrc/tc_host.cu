// Host-side CUtensorMap encoding for the TMA-fed kernels.  cuTensorMapEncodeTiled is a DRIVER entry
// point: it is resolved at run time through cudaGetDriverEntryPoint so that the library links against
// the CUDA runtime only and still loads (and exports every symbol) on a machine without a driver.
// Encoded maps are cached by (address, shape, strides, box): the caching allocator hands the same
// buffers back every step, so steady-state training encodes nothing.  The cache is the library's only
// global state; it is immutable per key and guarded by a mutex (threading).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>

#include "tc_common.cuh"

namespace ub {
namespace tc {
namespace {

using EncodeTiledFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                   const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn resolve_encoder() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

struct Key {
    uint64_t v[18];
    bool operator==(const Key &o) const { return std::memcmp(v, o.v, sizeof(v)) == 0; }
};
struct KeyHash {
    size_t operator()(const Key &k) const {
        uint64_t h = 1469598103934665603ull;
        for (uint64_t x : k.v) { h ^= x; h *= 1099511628211ull; }
        return (size_t)h;
    }
};

std::mutex g_mu;
std::unordered_map<Key, CUtensorMap, KeyHash> g_cache;

}  // namespace

namespace {
int encode_tensor_map(CUtensorMap *out, const void *base, int rank, const int64_t *dims, const int64_t *strides_elems,
                      const int *box, const int *elem_strides, int esize);
}

int encode_bf16_tensor_map(CUtensorMap *out, const void *base, int rank, const int64_t *dims,
                           const int64_t *strides_elems, const int *box, const int *elem_strides) {
    return encode_tensor_map(out, base, rank, dims, strides_elems, box, elem_strides, 2);
}

int encode_f32_tensor_map(CUtensorMap *out, const void *base, int rank, const int64_t *dims,
                          const int64_t *strides_elems, const int *box) {
    return encode_tensor_map(out, base, rank, dims, strides_elems, box, nullptr, 4);
}

namespace {
int encode_tensor_map(CUtensorMap *out, const void *base, int rank, const int64_t *dims, const int64_t *strides_elems,
                      const int *box, const int *elem_strides, int esize) {
    if (!out || !base || rank < 2 || rank > 4) return UB200_E_BADARG;
    const CUtensorMapDataType dtype = esize == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    Key key{};
    key.v[0] = reinterpret_cast<uint64_t>(base);
    key.v[1] = (uint64_t)rank | ((uint64_t)esize << 8);
    for (int i = 0; i < rank; ++i) {
        key.v[2 + i] = (uint64_t)dims[i];
        key.v[6 + i] = (uint64_t)box[i];
        if (i + 1 < rank) key.v[10 + i] = (uint64_t)strides_elems[i];
        key.v[14 + i] = (uint64_t)(elem_strides ? elem_strides[i] : 1);
    }
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_cache.find(key);
        if (it != g_cache.end()) { *out = it->second; return UB200_OK; }
    }
    EncodeTiledFn enc = resolve_encoder();
    if (!enc) return UB200_E_NODEVICE;
    cuuint64_t gdim[4], gstr[3];
    cuuint32_t gbox[4], estr[4];
    for (int i = 0; i < rank; ++i) {
        const int es = elem_strides ? elem_strides[i] : 1;
        if (dims[i] <= 0 || box[i] <= 0 || es < 1 || es > 8 || box[i] * es > 256) return UB200_E_UNSUPPORTED;
        gdim[i] = (cuuint64_t)dims[i];
        gbox[i] = (cuuint32_t)(box[i] * es);      // bounding box; ceil(gbox / es) = box elements are loaded
        estr[i] = (cuuint32_t)es;
        if (i + 1 < rank) {
            if (strides_elems[i] <= 0 || (strides_elems[i] * esize) % 16 != 0) return UB200_E_UNSUPPORTED;
            gstr[i] = (cuuint64_t)strides_elems[i] * esize;
        }
    }
    const int row_bytes = box[0] * esize;
    CUtensorMapSwizzle swz = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                           : row_bytes == 64  ? CU_TENSOR_MAP_SWIZZLE_64B
                           : row_bytes == 32  ? CU_TENSOR_MAP_SWIZZLE_32B
                                              : CU_TENSOR_MAP_SWIZZLE_NONE;
    if (swz == CU_TENSOR_MAP_SWIZZLE_NONE) return UB200_E_UNSUPPORTED;
    CUtensorMap m;
    CUresult r = enc(&m, dtype, (cuuint32_t)rank, const_cast<void *>(base), gdim, gstr, gbox,
                     estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_ERROR_INVALID_CONTEXT || r == CUDA_ERROR_NOT_INITIALIZED) {
        // A thread that has made no runtime call yet (autograd runs backward on its own worker threads) has no current
        // driver context, and this is a DRIVER entry point: let the runtime bind the primary context, then retry once.
        (void)cudaFree(nullptr);
        r = enc(&m, dtype, (cuuint32_t)rank, const_cast<void *>(base), gdim, gstr, gbox, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) {
        if (getenv("UB200_DEBUG")) {
            fprintf(stderr, "cuTensorMapEncodeTiled failed (%d): base %p rank %d", (int)r, base, rank);
            for (int i = 0; i < rank; ++i)
                fprintf(stderr, " | dim %llu box %u es %u stride %llu", (unsigned long long)gdim[i], gbox[i], estr[i],
                        (unsigned long long)(i + 1 < rank ? gstr[i] : 0));
            fprintf(stderr, "\n");
        }
        return UB200_E_UNSUPPORTED;
    }
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (g_cache.size() > 65536) g_cache.clear();   // bounded; a changing allocator pattern just re-encodes
        g_cache.emplace(key, m);
    }
    *out = m;
    return UB200_OK;
}
}  // namespace

}  // namespace tc
}  // namespace ub
