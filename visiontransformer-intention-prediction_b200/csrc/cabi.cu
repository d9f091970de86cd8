// cabi.cu -- library-level entry points of the C ABI (include/bevgen.h).
#include "bevgen_common.cuh"

#include <atomic>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

namespace bevgen {
std::atomic<long long> g_launch_count{0};

namespace {
struct TimedLaunch { const char* name; cudaEvent_t e0, e1; };
std::atomic<bool> g_timing{false};
std::mutex g_timing_mu;                    // guards the event lists below (launchers may be called from several host threads)
std::vector<TimedLaunch> g_pending;
std::vector<cudaEvent_t> g_event_pool;
cudaEvent_t g_open = nullptr;
const char* g_open_name = nullptr;

cudaEvent_t get_event() {
    if (!g_event_pool.empty()) { cudaEvent_t e = g_event_pool.back(); g_event_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
}  // namespace

// cudaFuncAttributeMaxDynamicSharedMemorySize belongs to the (kernel, device) pair, not to the process: a second GPU
// driven from the same process needs its own call.  Remembered per pair under a mutex; raising it again is a no-op.
int ensure_dyn_smem(const void* kernel, int bytes) {
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, int> done;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    std::lock_guard<std::mutex> lk(mu);
    int& have = done[std::make_pair(kernel, dev)];
    if (bytes <= have) return 0;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return (int)e;
    have = bytes;
    return 0;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (no link-time dependency on libcuda).
int make_grid_tensor_map(CUtensorMap* map, float* base, uint64_t planes, uint32_t box_cols, uint32_t box_rows) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static std::mutex mu;
    static EncodeFn fn = nullptr;
    {
        std::lock_guard<std::mutex> lk(mu);
        if (!fn) {
            void* p = nullptr;
            cudaDriverEntryPointQueryResult q;
            const cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
            if (e != cudaSuccess) return (int)e;
            if (q != cudaDriverEntryPointSuccess || !p) return BEVGEN_E_UNSUPPORTED;
            fn = reinterpret_cast<EncodeFn>(p);
        }
    }
    const cuuint64_t gdim[3] = {(cuuint64_t)kW, (cuuint64_t)kH, (cuuint64_t)planes};
    const cuuint64_t gstride[2] = {(cuuint64_t)kW * sizeof(float), (cuuint64_t)kPlane * sizeof(float)};   // bytes, dims 1 and 2
    const cuuint32_t box[3] = {box_cols, box_rows, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : BEVGEN_E_INVALID_ARG;
}

// Timing is a single-caller developer / bench facility: begin/end pairs of one launch are bracketed by the same lock
// holder only in the sense that the lists stay consistent; concurrent launchers may interleave their records.
void timing_begin(const char* name, cudaStream_t st) {
    if (!g_timing) return;
    std::lock_guard<std::mutex> lk(g_timing_mu);
    g_open = get_event();
    g_open_name = name;
    cudaEventRecord(g_open, st);
}
void timing_end(cudaStream_t st) {
    if (!g_timing) return;
    std::lock_guard<std::mutex> lk(g_timing_mu);
    if (!g_open) return;
    cudaEvent_t e1 = get_event();
    cudaEventRecord(e1, st);
    g_pending.push_back({g_open_name, g_open, e1});
    g_open = nullptr;
}
}  // namespace bevgen

extern "C" void bevgen_timing_enable(int on) { bevgen::g_timing = on != 0; }

// Synchronises the recorded events, writes "name count total_ms\n" lines into buf, clears the record.
extern "C" int bevgen_timing_collect(char* buf, size_t cap) {
    using namespace bevgen;
    std::lock_guard<std::mutex> lk(g_timing_mu);
    std::map<std::string, std::pair<long long, double>> agg;
    for (auto& t : g_pending) {
        float ms = 0.f;
        cudaEventSynchronize(t.e1);
        if (cudaEventElapsedTime(&ms, t.e0, t.e1) == cudaSuccess) { auto& a = agg[t.name]; a.first += 1; a.second += ms; }
        g_event_pool.push_back(t.e0);
        g_event_pool.push_back(t.e1);
    }
    g_pending.clear();
    std::string out;
    char line[160];
    for (auto& kv : agg) {
        snprintf(line, sizeof line, "%s %lld %.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
        out += line;
    }
    if (!buf || cap == 0) return (int)out.size();
    const size_t n = out.size() < cap - 1 ? out.size() : cap - 1;
    memcpy(buf, out.data(), n);
    buf[n] = 0;
    return (int)n;
}

// Sweep store upload: the four columns of one sweep, host -> device, asynchronous on `stream` when the sources are pinned.
extern "C" int bevgen_store_put_cols(void* dst_x, void* dst_y, void* dst_z, void* dst_i, const void* src_x, const void* src_y,
                                     const void* src_z, const void* src_i, int64_t n, bevgen_stream_t stream) {
    if (n < 0) return BEVGEN_E_INVALID_ARG;
    if (n == 0) return 0;
    if (!dst_x || !dst_y || !dst_z || !dst_i || !src_x || !src_y || !src_z || !src_i) return BEVGEN_E_INVALID_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    BEVGEN_CUDA(cudaMemcpyAsync(dst_x, src_x, (size_t)n * 2, cudaMemcpyHostToDevice, st));
    BEVGEN_CUDA(cudaMemcpyAsync(dst_y, src_y, (size_t)n * 2, cudaMemcpyHostToDevice, st));
    BEVGEN_CUDA(cudaMemcpyAsync(dst_z, src_z, (size_t)n * 2, cudaMemcpyHostToDevice, st));
    BEVGEN_CUDA(cudaMemcpyAsync(dst_i, src_i, (size_t)n, cudaMemcpyHostToDevice, st));
    return 0;
}

extern "C" int bevgen_abi_version(void) { return BEVGEN_ABI_VERSION; }

extern "C" void bevgen_constants(double* out) {
    using namespace bevgen;
    out[0] = kH; out[1] = kW; out[2] = kVoxel; out[3] = kOffX; out[4] = kOffY;
    out[5] = kZMin; out[6] = kZMax; out[7] = kZ; out[8] = kMapC; out[9] = kZSpan;
}

extern "C" const char* bevgen_error_string(int code) {
    if (code == 0) return "ok";
    if (code == BEVGEN_E_INVALID_ARG) return "bevgen: invalid argument";
    if (code == BEVGEN_E_WORKSPACE) return "bevgen: workspace missing or too small";
    if (code == BEVGEN_E_UNSUPPORTED) return "bevgen: unsupported size";
    if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
    return "bevgen: unknown error";
}

extern "C" int64_t bevgen_launch_count(void) { return bevgen::g_launch_count.load(); }
