// bevgen_common.cuh -- shared constants, exact-arithmetic device helpers, PTX wrappers (sm_100a).
#pragma once

#include <cuda.h>              // CUtensorMap (types only: the encoder is looked up at run time, libcuda is not linked)
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/bevgen.h"

namespace bevgen {

// ---- grid constants: reference constants.py:28-47 ------------------------------------------
constexpr int kH = BEVGEN_GRID_H;          // 400 rows   (ego x: +60 m at row 0 ... -20 m)
constexpr int kW = BEVGEN_GRID_W;          // 720 cols   (ego y: -72 m ... +72 m)
constexpr int kZ = BEVGEN_ZBINS;           // 29 height bins per sweep
constexpr int kMapC = BEVGEN_MAP_CHANNELS; // 9
constexpr double kVoxel = 0.2;
constexpr double kOffX = 360.0;            // BEV_PIXEL_OFFSET_X
constexpr double kOffY = 300.0;            // BEV_PIXEL_OFFSET_Y
constexpr double kZMin = -2.0;
constexpr double kZMax = 3.8;
constexpr double kZSpan = kZMax - kZMin;   // folded in IEEE double exactly like Python's 3.8-(-2.0)
constexpr int kPlane = kH * kW;            // 288 000
constexpr int kMaxRects = BEVGEN_MAX_DROP_RECTS;

// map bit planes (map_raster.cu): 768-bit rows, one plane per channel
constexpr int kMapWordsPerRow = 24;                    // 768 bits >= 720 columns
constexpr int kMapPlaneWords = kH * kMapWordsPerRow;   // 9 600 u32 per channel

extern std::atomic<long long> g_launch_count;   // host-side counter (cabi.cu)

// opt a kernel into `bytes` of dynamic shared memory on the CURRENT device (tracked per (kernel, device); cabi.cu)
int ensure_dyn_smem(const void* kernel, int bytes);
#define BEVGEN_DYN_SMEM(kernel, bytes)                                                   \
    do {                                                                                 \
        const int rc__ = ::bevgen::ensure_dyn_smem((const void*)(kernel), (int)(bytes)); \
        if (rc__) return rc__;                                                           \
    } while (0)

#define BEVGEN_CHECK_LAUNCH()                                  \
    do {                                                       \
        ++::bevgen::g_launch_count;                            \
        cudaError_t e__ = cudaGetLastError();                  \
        if (e__ != cudaSuccess) return (int)e__;               \
    } while (0)

// augmentation of map BIT PLANES straight into float32 grids (augment.cu; used by bevgen_map_raster_augmented):
// frames b with do_rotate / do_scale in params_host: planes [B,9,400,24] u32 -> dst [B,9,400,720]
int launch_augment_bitplanes(const uint32_t* planes, float* dst, int B, const bevgen_aug_params* params,
                             const bevgen_aug_params* params_host, cudaStream_t st);

// optional per-kernel timing (bench.py): CUDA events recorded on the launching stream around each launch
void timing_begin(const char* name, cudaStream_t st);
void timing_end(cudaStream_t st);

#define BEVGEN_LAUNCH(name, st, ...)                           \
    do {                                                       \
        ::bevgen::timing_begin(name, st);                      \
        __VA_ARGS__;                                           \
        ::bevgen::timing_end(st);                              \
        BEVGEN_CHECK_LAUNCH();                                 \
    } while (0)

#define BEVGEN_CUDA(call)                                      \
    do {                                                       \
        cudaError_t e__ = (call);                              \
        if (e__ != cudaSuccess) return (int)e__;               \
    } while (0)

// ---- exact arithmetic ------------------------------------------------------------------------
// All binning math uses the round-to-nearest intrinsics so that neither -fmad nor any other
// contraction can change a bit; the only fused operations are the explicit __fma_rn chains that
// restate OpenBLAS's dgemm accumulation.

struct Pose34 { double m[12]; };   // rows 0..2 of the 4x4 rel_tf (row 3 is 0 0 0 1)

__device__ __forceinline__ void xform_point(const double* __restrict__ T, double x, double y, double z,
                                            double& ox, double& oy, double& oz) {
    // utils.py:32: (T @ [x y z 1]^T), per row: fma chain k = 0..3 starting from +0.0
    ox = __fma_rn(T[3], 1.0, __fma_rn(T[2], z, __fma_rn(T[1], y, __fma_rn(T[0], x, 0.0))));
    oy = __fma_rn(T[7], 1.0, __fma_rn(T[6], z, __fma_rn(T[5], y, __fma_rn(T[4], x, 0.0))));
    oz = __fma_rn(T[11], 1.0, __fma_rn(T[10], z, __fma_rn(T[9], y, __fma_rn(T[8], x, 0.0))));
}

// utils.py:80-95 evaluated literally (three IEEE float64 divisions).  Returns the voxel packed as
// (zb << 20) | (py << 10) | px, or -1 when the point is masked out (utils.py:83-87).  Kept out of line (it is the
// rare path) and free of reference parameters so that nothing is forced into local memory.
static __device__ __noinline__ int lidar_voxel_exact(double x, double y, double z) {
    const double fx = floor(__dadd_rn(kOffX, __ddiv_rn(y, kVoxel)));   // utils.py:80
    const double fy = floor(__dsub_rn(kOffY, __ddiv_rn(x, kVoxel)));   // utils.py:81
    // NaN / +-inf / out-of-range fail these compares, as the reference's INT_MIN cast does
    if (!(fx >= 0.0 && fx < (double)kW && fy >= 0.0 && fy < (double)kH && z >= kZMin && z < kZMax)) return -1;
    const double fz = floor(__dmul_rn(__ddiv_rn(__dsub_rn(z, kZMin), kZSpan), (double)kZ));   // utils.py:94
    int b = (int)fz;
    b = b < 0 ? 0 : (b > kZ - 1 ? kZ - 1 : b);                                                // utils.py:95
    return (b << 20) | ((int)fy << 10) | (int)fx;
}

// Same result, bit for bit, without the divisions on the common path.
//   reference:  r = fl(360 + fl(y / 0.2d)),  0.2d = 0.2(1 + 5.6e-17)
//   fast:       t = fl(fma(y, 5, 360))
// For -1 <= value <= 721:  |r - t| < (|5y| + |360 + 5y|) * 2^-51 < 1e-12, so floor(r) == floor(t) whenever
// t is further than kEps = 1e-6 from an integer; the (rare) near-integer points take the literal path.
// Same argument for 300 - x/0.2 and for (z + 2)/5.8d*29 vs (z + 2)*5  (29/5.8d = 5(1 + 3e-17)).
__device__ __forceinline__ bool lidar_voxel(double x, double y, double z, int& py, int& px, int& zb) {
    if (!(z >= kZMin && z < kZMax)) return false;                       // utils.py:86-87 (exact compares; rejects NaN z)
    const double tx = __fma_rn(y, 5.0, kOffX);
    const double ty = __fma_rn(x, -5.0, kOffY);
    const double tz = __dmul_rn(__dsub_rn(z, kZMin), 5.0);
    // floor via saturating conversions; NaN converts to 0 but then fails the "safely inside a cell" test below
    const int ix = __double2int_rd(tx), iy = __double2int_rd(ty), iz = __double2int_rd(tz);
    constexpr double kEps = 1e-6;
    const double dx = __dsub_rn(tx, (double)ix), dy = __dsub_rn(ty, (double)iy), dz = __dsub_rn(tz, (double)iz);
    // written positively so that NaN / saturated values fall through to the literal path
    const bool safe = dx >= kEps && dx <= 1.0 - kEps && dy >= kEps && dy <= 1.0 - kEps && dz >= kEps && dz <= 1.0 - kEps;
    if (safe) {                                                          // the common path: no packing round trip
        if ((unsigned)ix >= (unsigned)kW || (unsigned)iy >= (unsigned)kH) return false;
        px = ix; py = iy; zb = min(iz, kZ - 1);
        return true;
    }
    // near a cell boundary (or non-finite): quick reject far outside the grid, else evaluate literally
    if (!(tx > -0.5 && tx < (double)kW + 0.5 && ty > -0.5 && ty < (double)kH + 0.5)) return false;
    const int packed = lidar_voxel_exact(x, y, z);
    if (packed < 0) return false;
    px = packed & 1023; py = (packed >> 10) & 1023; zb = packed >> 20;
    return true;
}

// utils.py:52-60 for one map vertex; T = 3x3 row-major.  Returns rounded (half-even) pixel coords as
// doubles-turned-int64 semantics: values outside int32 are clamped far outside the grid.
__device__ __forceinline__ void map_vertex_px(const double* __restrict__ T, double wx, double wy,
                                              long long& px, long long& py) {
    const double ex = __fma_rn(T[2], 1.0, __fma_rn(T[1], wy, __fma_rn(T[0], wx, 0.0)));
    const double ey = __fma_rn(T[5], 1.0, __fma_rn(T[4], wy, __fma_rn(T[3], wx, 0.0)));
    const double rx = rint(__dadd_rn(kOffX, __ddiv_rn(ey, kVoxel)));   // np.round = half-to-even
    const double ry = rint(__dsub_rn(kOffY, __ddiv_rn(ex, kVoxel)));
    // astype(int): out-of-range / NaN -> INT64_MIN (x86 cvttsd2si)
    px = (rx >= -9223372036854775808.0 && rx < 9223372036854775808.0) ? (long long)rx : (long long)0x8000000000000000LL;
    py = (ry >= -9223372036854775808.0 && ry < 9223372036854775808.0) ? (long long)ry : (long long)0x8000000000000000LL;
}

// per-frame dropout rectangles, (y0, x0, ph, pw) as drawn by utils.py:490-494
__device__ __forceinline__ bool in_rects(const int* __restrict__ r, int n, int y, int x) {
    bool hit = false;
#pragma unroll
    for (int k = 0; k < kMaxRects; ++k) {
        if (k < n) {
            const int y0 = r[4 * k], x0 = r[4 * k + 1];
            hit |= (y >= y0) & (y < y0 + r[4 * k + 2]) & (x >= x0) & (x < x0 + r[4 * k + 3]);
        }
    }
    return hit;
}

// ---- PTX wrappers: bulk async copy (TMA, non-tensor form) -------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared::cta -> global, completion tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bulk_store_s2g(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {   // <= N groups may still be READING their smem source
    asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory");
}
// order this thread's generic-proxy smem writes before later async-proxy (TMA) reads
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }


// tiled (tensor-map) form: one instruction stores a whole [1 x rows x cols] box of a 3-D tensor from a dense smem tile
__device__ __forceinline__ void tensor_store_3d(const CUtensorMap* map, int x, int y, int z, const void* ssrc) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
                 :: "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(z), "r"(smem_u32(ssrc)) : "memory");
}
__device__ __forceinline__ void tensor_store_3d_hint(const CUtensorMap* map, int x, int y, int z, const void* ssrc, uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%1, %2, %3}], [%4], %5;"
                 :: "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(z), "r"(smem_u32(ssrc)), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// host: tensor map over a float32 grid viewed as [planes, 400, 720] with a [1, box_rows, box_cols] box (cabi.cu)
int make_grid_tensor_map(CUtensorMap* map, float* base, uint64_t planes, uint32_t box_cols, uint32_t box_rows);

// ---- PTX wrappers: per-thread async copy global -> shared (LDGSTS), 8 bytes ---------------------
__device__ __forceinline__ void cp_async_8(void* sdst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(smem_u32(sdst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }   // <= N groups pending

}  // namespace bevgen
