// api.cu — library-level entry points of include/tspn_b200.h: version, errors, device
// check, host-side batch layout, tensor-map encoding.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace tspn {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_arch() {
    static thread_local int cached_dev = -1;
    static thread_local int cached_res = TSPN_ECUDA;
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        set_error("cudaGetDevice failed: %s", cudaGetErrorString(e));
        return TSPN_ECUDA;
    }
    if (dev == cached_dev) {
        if (cached_res == TSPN_EARCH) set_error("device %d is not compute capability 10.x (sm_100a only, no fallback)", dev);
        return cached_res;
    }
    int major = 0, minor = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    cached_dev = dev;
    if (major != 10) {
        set_error("device %d is compute capability %d.%d; libtspn_b200 is sm_100a only (no fallback)", dev, major, minor);
        cached_res = TSPN_EARCH;
    } else {
        cached_res = TSPN_OK;
    }
    return cached_res;
}

int num_sms() {
    static thread_local int cached_dev = -1;
    static thread_local int sms = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && dev != cached_dev) {
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cached_dev = dev;
    }
    return sms;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode_tensor_map(CUtensorMap* map, CUtensorMapDataType dtype, uint32_t rank, const void* base,
                      const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                      CUtensorMapSwizzle swizzle) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
        if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
            set_error("cuTensorMapEncodeTiled not available from the driver");
            return TSPN_ECUDA;
        }
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    cuuint64_t gdim[5];
    cuuint64_t gstr[4];
    cuuint32_t bdim[5];
    cuuint32_t estr[5];
    for (uint32_t i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bdim[i] = box[i];
        estr[i] = 1;
        if (i + 1 < rank) gstr[i] = strides_bytes[i];
    }
    CUresult r = fn(map, dtype, rank, const_cast<void*>(base), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %u, dims %llu x %llu, box %u x %u)", (int)r,
                  rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), box[0],
                  rank > 1 ? box[1] : 0);
        return TSPN_ECUDA;
    }
    return TSPN_OK;
}

}  // namespace tspn

extern "C" {

int tspn_version(void) { return TSPN_ABI_VERSION; }

int tspn_last_error(char* buf, int len) {
    int n = (int)strlen(tspn::g_err);
    if (buf && len > 0) {
        int c = n < len - 1 ? n : len - 1;
        memcpy(buf, tspn::g_err, c);
        buf[c] = 0;
    }
    return n;
}

int tspn_check_device(void) { return tspn::check_arch(); }

int tspn_geo_chunk(int64_t max_t) { return max_t <= 512 ? 512 : (max_t <= 1024 ? 1024 : 2048); }

int tspn_build_video_table(int num_videos, const int32_t* n_tracklets, const int32_t* n_frames, int table_rows,
                           int geo_chunk, int64_t* table_host, int64_t* totals) {
    TSPN_REQUIRE(num_videos >= 0 && (num_videos == 0 || (n_tracklets && n_frames)) && table_host && totals,
                 TSPN_EBADARG, "tspn_build_video_table: null argument");
    if (table_rows == 0) table_rows = num_videos;
    TSPN_REQUIRE(table_rows >= num_videos, TSPN_EBADARG, "tspn_build_video_table: table_rows=%d < num_videos=%d",
                 table_rows, num_videos);
    TSPN_REQUIRE(geo_chunk == 0 || geo_chunk == 512 || geo_chunk == 1024 || geo_chunk == 2048, TSPN_EBADARG,
                 "tspn_build_video_table: geo_chunk=%d (0, 512, 1024 or 2048)", geo_chunk);
    int64_t trk = 0, pairs = 0, geo = 0, items = 0, boxes = 0, scores = 0, max_n = 0, max_t = 0, max_chunks = 1;
    for (int v = 0; v < num_videos; ++v)
        if (n_frames[v] > max_t) max_t = n_frames[v];
    const int64_t chunk = geo_chunk ? geo_chunk : tspn_geo_chunk(max_t);
    // rows [num_videos, table_rows): empty videos; row table_rows: the sentinel (offset columns = totals)
    for (int v = 0; v <= table_rows; ++v) {
        const bool real = v < num_videos;
        const int64_t n = real ? n_tracklets[v] : 0, t = real ? n_frames[v] : (v < table_rows ? 1 : 0);
        if (real)
            TSPN_REQUIRE(n >= 0 && t >= 1, TSPN_ESHAPE, "video %d: need N >= 0 and T >= 1 (got N=%lld T=%lld)", v,
                         (long long)n, (long long)t);
        const int64_t tp = (t + 3) / 4 * 4, tb = (t + 7) / 8 * 8;
        int64_t* r = table_host + (int64_t)v * TSPN_VT_COLS;
        for (int c = 0; c < TSPN_VT_COLS; ++c) r[c] = 0;
        r[TSPN_VT_N] = n;
        r[TSPN_VT_T] = t;
        r[TSPN_VT_TP] = tp;
        r[TSPN_VT_TB] = tb;
        r[TSPN_VT_TRK_OFF] = trk;
        r[TSPN_VT_PAIR_OFF] = pairs;
        r[TSPN_VT_GEO_OFF] = geo;
        r[TSPN_VT_ITEM_OFF] = items;
        r[TSPN_VT_BOX_OFF] = boxes;
        r[TSPN_VT_SCORE_OFF] = scores;
        const int64_t p = n * (n - 1 > 0 ? n - 1 : 0);
        const int64_t nchunks = (t + chunk - 1) / chunk;
        trk += n;
        pairs += p;
        geo += p * TSPN_GEO_CHANNELS * tp;
        items += n >= 2 ? n * ((n - 1 + TSPN_GEO_OBJ_GROUP - 1) / TSPN_GEO_OBJ_GROUP) * nchunks : 0;
        boxes += n * tb;
        scores += n * n;
        if (n > max_n) max_n = n;
        if (n >= 2 && nchunks > max_chunks) max_chunks = nchunks;
    }
    totals[TSPN_TOT_TRACKLETS] = trk;
    totals[TSPN_TOT_PAIRS] = pairs;
    totals[TSPN_TOT_GEO_FLOATS] = geo;
    totals[TSPN_TOT_ITEMS] = items;
    totals[TSPN_TOT_BOXES] = boxes;
    totals[TSPN_TOT_SCORES] = scores;
    totals[TSPN_TOT_MAX_N] = max_n;
    totals[TSPN_TOT_MAX_T] = max_t;
    totals[TSPN_TOT_GEO_CHUNK] = chunk;
    totals[TSPN_TOT_MAX_CHUNKS] = max_chunks;
    return TSPN_OK;
}

}  // extern "C"
