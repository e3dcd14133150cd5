// Host-side runtime of libvtc: error state, device checks, TMA descriptor encoding.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <mutex>

#include "common.cuh"
#include "tma_host.h"

namespace vtc {

static thread_local char g_last_error[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    set_last_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    return VTC_ERR_CUDA;
}

int device_sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

int SmemOptIn::ensure(const void* func, size_t bytes, bool max_carveout) {
    int dev = -1;
    VTC_CUDA(cudaGetDevice(&dev));
    const bool slot = dev >= 0 && dev < 64;
    if (slot && cur[dev].load(std::memory_order_acquire) >= bytes) return VTC_OK;
    VTC_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
    if (max_carveout) VTC_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    if (slot) cur[dev].store(bytes, std::memory_order_release);      // racing threads set the same value twice: harmless
    return VTC_OK;
}

int check_arch() {
    static int cached[64] = {0};   // 0 unknown, 1 ok, -1 bad
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice (no CUDA device: libvtc has no CPU path)", __FILE__, __LINE__);
    if (dev >= 0 && dev < 64 && cached[dev] == 1) return VTC_OK;
    int major = 0, minor = 0;
    VTC_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    VTC_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    if (major != 10) {
        set_last_error("device %d is sm_%d%d; libvtc is built for sm_100a only (no fallback path)", dev, major, minor);
        return VTC_ERR_ARCH;
    }
    if (dev >= 0 && dev < 64) cached[dev] = 1;
    return VTC_OK;
}

// ---- TMA descriptors ----------------------------------------------------------------------------
// cuTensorMapEncodeTiled is fetched through the runtime so that libvtc.so has no link-time dependency on
// libcuda (the library must load, and export its symbols, on a machine without a driver).
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    });
    return fn;
}

static int make_tmap(CUtensorMap* map, CUtensorMapDataType dtype, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box) {
    PFN_encodeTiled enc = get_encode();
    VTC_REQUIRE(enc != nullptr, VTC_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    VTC_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, VTC_ERR_ARG, "TMA base pointer must be 16-byte aligned");
    cuuint64_t gdim[5];
    cuuint64_t gstr[5];
    cuuint32_t bdim[5];
    cuuint32_t estr[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bdim[i] = box[i];
        estr[i] = 1;
        if (i > 0) {
            gstr[i - 1] = strides_bytes[i - 1];
            VTC_REQUIRE(gstr[i - 1] % 16 == 0, VTC_ERR_SHAPE, "TMA stride %llu not a multiple of 16 bytes",
                        (unsigned long long)gstr[i - 1]);
        }
    }
    CUresult r = enc(map, dtype, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu box %u,%u)", (int)r, rank,
                       (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
        return VTC_ERR_CUDA;
    }
    return VTC_OK;
}

int make_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box) {
    return make_tmap(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rank, dims, strides_bytes, box);
}
int make_tmap_f32(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box) {
    return make_tmap(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, rank, dims, strides_bytes, box);
}

}  // namespace vtc

extern "C" {
int vtc_version(void) { return VTC_VERSION; }
const char* vtc_last_error(void) { return vtc::g_last_error; }
int vtc_check_device(void) { return vtc::check_arch(); }
uint64_t vtc_launch_count(void) { return vtc::g_launches.load(std::memory_order_relaxed); }
}
