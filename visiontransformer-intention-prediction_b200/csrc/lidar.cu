// lidar.cu -- multi-sweep LiDAR aggregation -> height-binned max-intensity BEV grid (sm_100a).
//
// Replaces, for a whole batch of frames in two or three launches:
//   transform_points            reference utils.py:27-33   (per sweep, dataset.py:339-340)
//   create_intentnet_lidar_bev  reference utils.py:62-106
//
// Three selectable algorithms, identical bits (DESIGN.md section 3.1 has the measurements and the variants that were
// tried and removed):
//   TILED (default)   k_bin_fixed       one pass: points -> fixed-capacity per-tile record bins (+ per-sweep overflow lists)
//                     k_tile_store      one short-lived CTA per 28.8 KB tile: shared-memory tile, warp-aggregated
//                                       (__match_any_sync) atomicMax, ONE TMA bulk store -- each output byte written once
//                     k_overflow_red    overflow records as global RED.MAX
//   ATOMIC            k_zero_fill_tma   one short-lived CTA per 28.8 KB chunk: TMA bulk store of a zero tile
//                     k_scatter_atomic  one thread per point: FP64 pose FMA chain, exact binning, RED.MAX into the grid
//                                       (faster on dense / skewed scenes, where the tile bins overflow)
//   MEMSET_ATOMIC     cudaMemsetAsync + k_scatter_atomic (library fill; kept only as a comparison point)
// and, on top of the tiled path's record bins, generation FUSED with the resampling augmentations (lidar_aug_tile.cuh:
// k_rebin_aug + k_tile_aug), entry point bevgen_lidar_bev_augmented.
#include "aug_geom.cuh"

#include <cuda_fp16.h>

#include <cstdlib>
#include <vector>



namespace bevgen {

namespace {

constexpr int kBinThreads = 256;
constexpr int kBinItems = 4;
constexpr int kBinChunk = kBinThreads * kBinItems;     // points per CTA
constexpr int kRowBytes = kW * 4;                      // 2 880 (multiple of 16: bulk-copy legal)

struct LidarIn {
    // MODE 3: AV2 sweep columns as stored on disk (dataset.py:311 reads x, y, z float16 + intensity uint8): four arrays,
    // sweeps concatenated inside each, 7 B / point, each 16-byte aligned and readable up to the next multiple of 8 rows
    const __half* col_x = nullptr; const __half* col_y = nullptr; const __half* col_z = nullptr; const uint8_t* col_i = nullptr;
    const uint2* pts_h4;       // MODE 2: 4 x fp16 (x, y, z, intensity) = 8 B / point (AV2's on-disk precision)
    const float4* pts4;        // MODE 0
    const double* rel_tf;      // MODE 0: [B*S,16]
    const double* pts3;        // MODE 1
    const float* intensity;    // MODE 1
    const int32_t* sweep_off;  // [B*S+1] CSR offsets: sweep k owns rows [sweep_off[k], sweep_off[k+1]) -- and, always, the
                               // slice [sweep_off[k] - sweep_off[0], ...) of the overflow list
    const int32_t* sweep_start = nullptr;   // [B*S] or null.  Sweep STORE form: sweep k owns rows [sweep_start[k], sweep_start[k] +
                               // sweep_off[k+1] - sweep_off[k]) of a resident point store, so frames may share sweeps (sliding
                               // windows) and sweeps need not be contiguous; sweep_off is then only the prefix sum of lengths
    const uint8_t* flip;       // [B] or null
    const int32_t* rects;      // [B,5,4] or null
    const int32_t* nrects;     // [B] or null
    int S;
    // alternative source of the per-frame fusions (bevgen_lidar_bev_augmented): flip always; dropout rectangles only for
    // frames that are not resampled afterwards (the fused tile kernel applies them to its output instead)
    const bevgen_aug_params* aug = nullptr;   // [B] or null
    int n_points = 0x7FFFFFFF;                // rows in the point array: sweep ranges are clamped to [0, n_points)
};

// Row range of sweep `bs`, made safe against malformed CSR offsets (they live in device memory, the host cannot check them
// without a synchronisation): clamped to the point array, empty when it is not ordered after the first offset.
__device__ __forceinline__ void sweep_range(const LidarIn& in, int bs, int& beg, int& end) {
    const int lb0 = in.sweep_off[0], lb = in.sweep_off[bs];
    int len = in.sweep_off[bs + 1] - lb;
    if (lb0 < 0 || lb < lb0 || len < 0) len = 0;       // (also keeps the sweep's overflow-list slice inside the list)
    beg = in.sweep_start ? in.sweep_start[bs] : lb;
    if (beg < 0 || beg > in.n_points) { beg = 0; len = 0; }
    end = beg + (len < in.n_points - beg ? len : in.n_points - beg);
}

inline LidarIn make_in(const uint2* h4, const float4* p4, const double* rel_tf, const double* p3, const float* inten,
                       const int32_t* off, const uint8_t* flip, const int32_t* rects, const int32_t* nrects, int S,
                       const bevgen_aug_params* aug = nullptr) {
    LidarIn in;
    in.pts_h4 = h4; in.pts4 = p4; in.rel_tf = rel_tf; in.pts3 = p3; in.intensity = inten; in.sweep_off = off;
    in.flip = flip; in.rects = rects; in.nrects = nrects; in.S = S; in.aug = aug;
    return in;
}

// Raw point as loaded from HBM.  Loads are issued for ALL of a thread's items before the first use, so
// kBinItems independent 16 B (MODE 0) / 28 B (MODE 1) requests are in flight per thread.
template <int MODE> struct RawPoint;
template <> struct RawPoint<0> { float4 p; };
template <> struct RawPoint<1> { double x, y, z; float v; };
template <> struct RawPoint<2> { uint2 h; };

template <int MODE>
__device__ __forceinline__ RawPoint<MODE> load_point(const LidarIn& in, long long i, bool in_range);
template <>
__device__ __forceinline__ RawPoint<0> load_point<0>(const LidarIn& in, long long i, bool in_range) {
    RawPoint<0> r;
    r.p = in_range ? __ldcs(in.pts4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);   // streaming 16 B load, fully coalesced
    return r;
}
template <>
__device__ __forceinline__ RawPoint<1> load_point<1>(const LidarIn& in, long long i, bool in_range) {
    RawPoint<1> r;
    r.x = r.y = r.z = 0.0; r.v = 0.f;
    if (in_range) {
        r.x = __ldcs(in.pts3 + 3 * i); r.y = __ldcs(in.pts3 + 3 * i + 1); r.z = __ldcs(in.pts3 + 3 * i + 2);
        r.v = __ldcs(in.intensity + i);
    }
    return r;
}

template <>
__device__ __forceinline__ RawPoint<2> load_point<2>(const LidarIn& in, long long i, bool in_range) {
    RawPoint<2> r;
    r.h = in_range ? __ldcs(in.pts_h4 + i) : make_uint2(0u, 0u);               // streaming 8 B load, fully coalesced
    return r;
}

// MODE 3: one thread takes 8 CONSECUTIVE rows of the four columns: three 16 B loads + one 8 B load per 8 points (the
// row-interleaved formats issue one load per point), a warp reads 512 contiguous bytes of each fp16 column.  `i8` is a
// multiple of 8, so every vector is aligned; rows outside the sweep's [beg, end) are masked by the caller.
struct ColGroup { uint4 x, y, z; uint2 v; };
__device__ __forceinline__ ColGroup load_cols(const LidarIn& in, uint32_t i8, bool in_range) {
    ColGroup g;
    g.x = g.y = g.z = make_uint4(0u, 0u, 0u, 0u); g.v = make_uint2(0u, 0u);
    if (in_range) {
        g.x = __ldcs(reinterpret_cast<const uint4*>(in.col_x + i8));
        g.y = __ldcs(reinterpret_cast<const uint4*>(in.col_y + i8));
        g.z = __ldcs(reinterpret_cast<const uint4*>(in.col_z + i8));
        g.v = __ldcs(reinterpret_cast<const uint2*>(in.col_i + i8));
    }
    return g;
}
__device__ __forceinline__ float half_at(const uint4& q, int u) {          // u compile-time after unrolling
    const uint32_t w = u < 2 ? q.x : (u < 4 ? q.y : (u < 6 ? q.z : q.w));
    return __half2float(__ushort_as_half((unsigned short)((u & 1) ? (w >> 16) : (w & 0xFFFFu))));
}
__device__ __forceinline__ RawPoint<0> col_point(const ColGroup& g, int u) {
    RawPoint<0> r;          // fp16 -> fp32 and uint8 -> fp32 are exact: same values the reference widens to float64 / float32
    const uint32_t w = u < 4 ? g.v.x : g.v.y;
    r.p = make_float4(half_at(g.x, u), half_at(g.y, u), half_at(g.z, u), (float)((w >> (8 * (u & 3))) & 0xFFu));
    return r;
}

// One point -> voxel (row py, column px after the optional flip, height bin zb) + intensity bits, or invalid.
// MODE 0 / 2 transform first.
struct Voxel { int py, px, zb; uint32_t bits; };

__device__ __forceinline__ bool finish_voxel(double x, double y, double z, float v, bool flip, const int* sRect, int nrect,
                                             Voxel& o) {
    // max with the +0.0 initial grid: only strictly positive intensities can change a voxel
    // (documented domain: finite or +inf, >= 0; then float max == unsigned max on the bit pattern)
    if (!(v > 0.0f)) return false;
    if (!lidar_voxel(x, y, z, o.py, o.px, o.zb)) return false;
    if (flip) o.px = kW - 1 - o.px;                                // utils.py:399 fused
    if (nrect && in_rects(sRect, nrect, o.py, o.px)) return false;  // utils.py:496 fused
    o.bits = __float_as_uint(v);
    return true;
}

template <int MODE>
__device__ __forceinline__ bool point_voxel(const RawPoint<MODE>& raw, const double* sT, bool flip, const int* sRect, int nrect,
                                            Voxel& o);
template <>
__device__ __forceinline__ bool point_voxel<0>(const RawPoint<0>& raw, const double* sT, bool flip, const int* sRect, int nrect,
                                               Voxel& o) {
    if (!(raw.p.w > 0.0f)) return false;                          // skip the FP64 work for no-op points
    double x, y, z;
    xform_point(sT, (double)raw.p.x, (double)raw.p.y, (double)raw.p.z, x, y, z);
    return finish_voxel(x, y, z, raw.p.w, flip, sRect, nrect, o);
}
template <>
__device__ __forceinline__ bool point_voxel<2>(const RawPoint<2>& raw, const double* sT, bool flip, const int* sRect, int nrect,
                                               Voxel& o) {
    // fp16 -> fp32 is exact, so this equals the reference reading float16 columns and widening to float64
    const __half2 xy = *reinterpret_cast<const __half2*>(&raw.h.x), zi = *reinterpret_cast<const __half2*>(&raw.h.y);
    RawPoint<0> f;
    f.p = make_float4(__low2float(xy), __high2float(xy), __low2float(zi), __high2float(zi));
    return point_voxel<0>(f, sT, flip, sRect, nrect, o);
}
template <>
__device__ __forceinline__ bool point_voxel<1>(const RawPoint<1>& raw, const double*, bool flip, const int* sRect, int nrect,
                                               Voxel& o) {
    return finish_voxel(raw.x, raw.y, raw.z, raw.v, flip, sRect, nrect, o);
}

__device__ __forceinline__ void load_frame_consts(const LidarIn& in, int bs, int b, int mode, double* sT, int* sRect,
                                                  int& nrect, bool& flip) {
    if (mode != 1 && threadIdx.x < 12) sT[threadIdx.x] = in.rel_tf[(size_t)bs * 16 + threadIdx.x];
    if (in.aug) {
        const bevgen_aug_params* P = in.aug + b;
        nrect = (P->do_rotate | P->do_scale) ? 0 : P->n_rects;
        nrect = nrect > kMaxRects ? kMaxRects : (nrect < 0 ? 0 : nrect);
        if ((int)threadIdx.x < 4 * nrect) sRect[threadIdx.x] = P->rects[threadIdx.x];
        flip = P->do_flip != 0;
        return;
    }
    nrect = in.nrects ? in.nrects[b] : 0;
    nrect = nrect > kMaxRects ? kMaxRects : nrect;
    if ((int)threadIdx.x < 4 * nrect) sRect[threadIdx.x] = in.rects[(size_t)b * kMaxRects * 4 + threadIdx.x];
    flip = in.flip ? (in.flip[b] != 0) : false;
}

// ---- fill geometry: one chunk = 10 grid rows of one channel = 28 800 contiguous bytes -------------------------
constexpr int kFillRows = 10;
constexpr int kFillBytes = kFillRows * kRowBytes;                    // 28 800 B
constexpr int kFillPerSweep = kZ * kH / kFillRows;                   // 1160 chunks per (frame, sweep)

// ---- "binned": single binning pass into fixed-capacity tile bins + one short-lived raster CTA per tile ---------
// Tile = (frame, sweep, z-bin, band of 10 rows) = 7 200 floats = 28 800 CONTIGUOUS bytes of the output, i.e. exactly
// one chunk of the fast TMA fill kernel.  1160 tiles per sweep.
//   k_bin_fixed      per point: transform, bin, rank inside the CTA (shared atomics); per touched tile ONE global
//                    atomicAdd claims the CTA's slots; slot < kBinCap -> bins[tile][slot] = record, else -> overflow
//                    list (dense scenes; one list-append atomic per CTA).  One pass over the points, no histogram
//                    scan, no intermediate copy.
//   k_tile_store     one 128-thread CTA per tile: zero a 28.8 KB shared-memory tile, apply the bin's records with
//                    warp-aggregated (__match_any_sync + __reduce_max_sync) shared atomicMax, ONE TMA bulk store.
//                    7 CTAs are resident per SM.
//   k_overflow_red   the (rare) overflow records as global RED.MAX after the tiles have landed.
// Every output byte is written exactly once by the TMA; no global read-modify-write except for overflow records.
constexpr int kBinCap = 128;                                         // records per tile bin (avg 40 at 100k pts/sweep)
constexpr int kTilesPerSweep = kFillPerSweep;                        // 1160
constexpr int kTileElems = kFillRows * kW;                           // 7200 floats per tile
struct OvfRec { unsigned long long idx; uint32_t bits; uint32_t pad; };
// Row-band bin records: 8 bytes (cell inside the tile, intensity float bits) -- or 4 bytes where every intensity is a float
// whose low 16 bits are zero, which holds for AV2's uint8 intensities (MODE 3: an integer 0..255 has at most 8 significant
// bits): cell (13 bits) | upper half of the float (sign 0, exponent, 7 mantissa bits) << 13.  Same bins array, 4-byte stride;
// half the record bytes written by k_bin_fixed and, what matters, read by k_tile_store in the middle of its write stream.
template <int MODE> constexpr bool kRec4 = MODE == 3;
__device__ __forceinline__ uint32_t band_rec_pack(uint32_t cell, uint32_t bits) { return cell | ((bits >> 16) << 13); }
__device__ __forceinline__ uint2 band_rec_unpack(uint32_t w) { return make_uint2(w & 8191u, (w >> 13) << 16); }
__device__ __forceinline__ uint2 band_rec_load(const uint2* __restrict__ bins, size_t idx, int rec4) {
    return rec4 ? band_rec_unpack(__ldg(reinterpret_cast<const uint32_t*>(bins) + idx)) : __ldg(bins + idx);
}

// CTA-aggregated: a CTA takes AggCfg<ITEMS>::kChunk consecutive points of one sweep, ranks its records per tile with SHARED
// atomics, then claims each touched tile's slots with ONE global atomic (instead of one per record) and writes
// the records out.  Shared memory (8 192-point CTAs): records (64 KB) + ranks (16 KB) + per-tile count/base (9 KB).
constexpr int kAggThreads = 512;
// Points per thread: 16 (8 192 points per CTA, 91 KB of shared memory) for large launches -- fewest global slot claims
// per record -- and 8 (4 096 points, 50 KB) for small ones, where twice as many half-size CTAs fill the GPU (a single frame
// is 130 big CTAs on 296 slots; B200, us per frame big -> small at batch 1 / 2 / 4: 24.1 -> 19.7, 14.9 -> 14.4, 11.8 -> 11.2;
// launch_bin() picks).
constexpr int kAggItemsBig = 16, kAggItemsSmall = 8;
template <int ITEMS> struct AggCfg {
    static constexpr int kChunk = kAggThreads * ITEMS;
    static constexpr int kSmem = kChunk * 8 + kChunk * 2 + 2 * kTilesPerSweep * 4;   // records, ranks, per-tile count / base
};

template <int MODE, int ITEMS>
__global__ void __launch_bounds__(kAggThreads, 2) k_bin_fixed(LidarIn in, uint32_t* __restrict__ cnt, uint2* __restrict__ bins,
                                                           uint32_t* __restrict__ ovf_cnt, OvfRec* __restrict__ ovf, int bs0, int l2_hints) {
    extern __shared__ __align__(16) unsigned char agg_smem[];
    uint2* const recs = reinterpret_cast<uint2*>(agg_smem);                                   // (bin << 13 | cell, bits)
    constexpr int kAggItems = ITEMS, kAggChunk = AggCfg<ITEMS>::kChunk;
    uint16_t* const rank = reinterpret_cast<uint16_t*>(agg_smem + kAggChunk * 8);
    uint32_t* const hist = reinterpret_cast<uint32_t*>(agg_smem + kAggChunk * 10);
    uint32_t* const base = hist + kTilesPerSweep;
    const int bs = bs0 + blockIdx.y, b = bs / in.S, tid = threadIdx.x;
    int beg, end;
    sweep_range(in, bs, beg, end);
    const uint32_t org = MODE == 3 ? ((uint32_t)beg & ~7u) : (uint32_t)beg;   // column vectors are 8-row aligned
    if ((long long)org + (long long)blockIdx.x * kAggChunk >= end) return;
    __shared__ double sT[12];
    __shared__ int sRect[kMaxRects * 4];
    __shared__ uint32_t s_spill_n, s_spill_base;
    int nrect; bool flip;
    load_frame_consts(in, bs, b, MODE, sT, sRect, nrect, flip);
    // The grid's x extent comes from the caller's max_sweep_points (a sizing hint): a sweep longer than that is still
    // binned completely, its CTAs walk the remaining chunks (normally this loop runs once).
#pragma unroll 1
    for (uint32_t chunk0 = org + blockIdx.x * (uint32_t)kAggChunk; chunk0 < (uint32_t)end; chunk0 += gridDim.x * (uint32_t)kAggChunk) {
        for (int i = tid; i < kTilesPerSweep; i += kAggThreads) hist[i] = 0;
        __syncthreads();
        // phase 1: records + per-tile rank inside the CTA (shared atomics)
        auto emit = [&](int k, bool ok, const Voxel& v) {
            uint32_t word = 0xFFFFFFFFu;
            if (ok) {
                const uint32_t band = (uint32_t)v.py / kFillRows;
                const uint32_t bin = (uint32_t)v.zb * (kH / kFillRows) + band;
                const uint32_t tcell = ((uint32_t)v.py - band * kFillRows) * kW + (uint32_t)v.px;
                word = (bin << 13) | tcell;
                rank[k] = (uint16_t)atomicAdd(&hist[bin], 1u);
            }
            recs[k] = make_uint2(word, v.bits);
        };
        if constexpr (MODE == 3) {
            // 8 consecutive rows per thread and group: 56 B per thread in flight (4 vector loads), 2 x 512 threads per SM
#pragma unroll 1
            for (int g = 0; g < kAggItems / 8; ++g) {
                const uint32_t i8 = chunk0 + (uint32_t)((g * kAggThreads + tid) * 8);
                const ColGroup grp = load_cols(in, i8, i8 < (uint32_t)end);
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint32_t i = i8 + (uint32_t)u;
                    Voxel v;
                    v.bits = 0;
                    const bool ok = i >= (uint32_t)beg && i < (uint32_t)end && point_voxel<0>(col_point(grp, u), sT, flip, sRect, nrect, v);
                    emit((g * 8 + u) * kAggThreads + tid, ok, v);       // slot index: conflict-free, any bijection works
                }
            }
        } else {
#pragma unroll 1
            for (int g = 0; g < kAggItems; g += 4) {
                RawPoint<MODE> raw[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t i = chunk0 + (uint32_t)((g + u) * kAggThreads + tid);      // < 2^32: n_points < 2^31, chunk < 2^13
                    raw[u] = load_point<MODE>(in, (long long)i, i < (uint32_t)end);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int k = (g + u) * kAggThreads + tid;
                    const uint32_t i = chunk0 + (uint32_t)k;
                    Voxel v;
                    v.bits = 0;
                    const bool ok = (i < (uint32_t)end) && point_voxel<MODE>(raw[u], sT, flip, sRect, nrect, v);
                    emit(k, ok, v);
                }
            }
        }
        __syncthreads();
        // phase 2: one global claim per touched tile (all of a thread's claims are in flight together)
        {
            constexpr int kPer = (kTilesPerSweep + kAggThreads - 1) / kAggThreads;       // 3
            uint32_t c[kPer], b0[kPer];
#pragma unroll
            for (int j = 0; j < kPer; ++j) {
                const int bin = tid + j * kAggThreads;
                c[j] = bin < kTilesPerSweep ? hist[bin] : 0u;
                b0[j] = 0u;
                if (c[j]) b0[j] = atomicAdd(cnt + (size_t)bs * kTilesPerSweep + bin, c[j]);
            }
#pragma unroll
            for (int j = 0; j < kPer; ++j)
                if (c[j]) base[tid + j * kAggThreads] = b0[j];
        }
        __syncthreads();
        // phase 3: records to their slots; spills are ranked inside the CTA first (one shared atomic per warp) ...
        if (tid == 0) s_spill_n = 0;
        __syncthreads();
        const int lane = tid & 31;
        const uint32_t lt = (1u << lane) - 1u;
        uint64_t hint_pol = 0;
        if (l2_hints) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(hint_pol));   // small batches: see l2_hints_for()
#pragma unroll 2
        for (int it = 0; it < kAggItems; ++it) {
            const int k = it * kAggThreads + tid;
            const uint2 r = recs[k];
            const bool ok = r.x != 0xFFFFFFFFu;
            const uint32_t bin = r.x >> 13, tcell = r.x & 8191u;
            const uint32_t tile = (uint32_t)bs * kTilesPerSweep + bin;
            const uint32_t slot = ok ? base[bin] + rank[k] : 0u;
            const bool spill = ok && slot >= (uint32_t)kBinCap;
            if (ok && !spill) {
                if (kRec4<MODE>) {                                       // AV2 intensities: 4-byte records (band_rec_pack)
                    uint32_t* dst = reinterpret_cast<uint32_t*>(bins) + (size_t)tile * kBinCap + slot;
                    const uint32_t w = band_rec_pack(tcell, r.y);
                    if (hint_pol) asm volatile("st.global.L2::cache_hint.b32 [%0], %1, %2;" :: "l"(dst), "r"(w), "l"(hint_pol) : "memory");
                    else *dst = w;
                } else {
                    uint2* dst = bins + (size_t)tile * kBinCap + slot;
                    if (hint_pol) asm volatile("st.global.L2::cache_hint.v2.b32 [%0], {%1, %2}, %3;" :: "l"(dst), "r"(tcell), "r"(r.y), "l"(hint_pol) : "memory");
                    else *dst = make_uint2(tcell, r.y);
                }
            }
            const uint32_t sp = __ballot_sync(0xFFFFFFFFu, spill);
            if (sp) {
                uint32_t lbase = 0;
                if (lane == __ffs(sp) - 1) lbase = atomicAdd(&s_spill_n, (uint32_t)__popc(sp));
                lbase = __shfl_sync(0xFFFFFFFFu, lbase, __ffs(sp) - 1);
                // a spilled record keeps its CTA-local overflow index (< 8192) in `rank`, flagged by bit 15 for phase 4
                if (spill) rank[k] = (uint16_t)(0x8000u | (lbase + __popc(sp & lt)));
            }
        }
        __syncthreads();
        // ... phase 4: ONE global list-append atomic per CTA, then the spilled records go out
        if (s_spill_n == 0) continue;                                    // uniform
        // per-sweep lists: sweep bs owns ovf[sweep_off[bs] - sweep_off[0] ...) (a point spills at most once), count in ovf_cnt[bs]
        if (tid == 0) s_spill_base = atomicAdd(ovf_cnt + bs, s_spill_n);
        __syncthreads();
        const uint32_t obase = (uint32_t)(in.sweep_off[bs] - in.sweep_off[0]) + s_spill_base;
#pragma unroll 1
        for (int it = 0; it < kAggItems; ++it) {
            const int k = it * kAggThreads + tid;
            const uint2 r = recs[k];
            const uint32_t rk = rank[k];
            if (r.x == 0xFFFFFFFFu || !(rk & 0x8000u)) continue;           // (ranks of valid records are < 8192: bit 15 = spilled)
            OvfRec o;
            o.idx = (unsigned long long)((uint32_t)bs * kTilesPerSweep + (r.x >> 13)) * kTileElems + (r.x & 8191u);
            o.bits = r.y; o.pad = 0;
            ovf[obase + (rk & 0x7FFFu)] = o;
        }
    }   // chunk loop
}

__global__ void __launch_bounds__(128) k_tile_store(const uint32_t* __restrict__ cnt, const uint2* __restrict__ bins,
                                                    float* __restrict__ out, uint32_t t0, int l2_hints, int rec4) {
    extern __shared__ __align__(128) unsigned char ztile[];
    uint32_t* const tile = reinterpret_cast<uint32_t*>(ztile);
    const int tid = threadIdx.x, lane = tid & 31;
    const size_t t = (size_t)t0 + blockIdx.x;
    uint32_t n = __ldcs(cnt + t);
    for (int i = tid; i < kFillBytes / 16; i += 128) reinterpret_cast<uint4*>(ztile)[i] = make_uint4(0, 0, 0, 0);
    n = n < (uint32_t)kBinCap ? n : (uint32_t)kBinCap;
    // only the records that exist are read (measured: speculative reads of empty slots cost more HBM read traffic in
    // the write stream than the dependent load costs in latency -- 7 resident CTAs per SM hide it)
    const bool have = (uint32_t)tid < n;
    uint2 rec = make_uint2(0xFFFFFFFFu, 0u);
    if (have) rec = rec4 ? band_rec_unpack(__ldcs(reinterpret_cast<const uint32_t*>(bins) + t * kBinCap + tid))
                         : __ldcs(bins + t * kBinCap + tid);
    __syncthreads();
    if (n) {                                                    // uniform
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, rec.x);      // dedupe equal cells inside the warp
        const uint32_t m = __reduce_max_sync(peers, rec.y);
        if (have && lane == __ffs(peers) - 1) atomicMax(&tile[rec.x], m);
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (tid != 0) return;
    if (l2_hints) {                                            // output lines leave L2 first, the record bins stay (l2_hints_for)
        uint64_t pol;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                     :: "l"(reinterpret_cast<unsigned char*>(out) + t * kFillBytes), "r"(smem_u32(ztile)), "r"((uint32_t)kFillBytes), "l"(pol) : "memory");
    } else {
        bulk_store_s2g(reinterpret_cast<unsigned char*>(out) + t * kFillBytes, ztile, kFillBytes);
    }
    bulk_commit();
    bulk_wait_read<0>();
}

// grid (kOvfCtas, B*S): sweep blockIdx.y's overflow list.  With `aug`, sweeps of frames that are resampled are skipped
// (k_tile_aug consumes their overflow records itself).
constexpr int kOvfCtas = 4;
__global__ void __launch_bounds__(256) k_overflow_red(const uint32_t* __restrict__ ovf_cnt, const OvfRec* __restrict__ ovf,
                                                      const int32_t* __restrict__ sweep_off, uint32_t* __restrict__ out,
                                                      const bevgen_aug_params* __restrict__ aug, int S) {
    const int bs = blockIdx.y;
    const uint32_t n = ovf_cnt[bs];
    if (n == 0) return;
    if (aug) { const bevgen_aug_params* P = aug + bs / S; if (P->do_rotate | P->do_scale) return; }
    const OvfRec* list = ovf + (sweep_off[bs] - sweep_off[0]);
    for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < n; i += kOvfCtas * 256) {
        const OvfRec r = list[i];
        atomicMax(out + r.idx, r.bits);
    }
}

#include "lidar_aug_tile.cuh"

// ---- zero fill by TMA, short-lived CTAs ---------------------------------------------------------------
// Each CTA zero-initialises one shared-memory tile (kFillBytes) and bulk-stores it ONCE
// (cp.async.bulk.global.shared::cta -> UBLKCP), then exits.  Measured on B200 (tools/fillbench.cu, 2.67 GB):
//   this form 7.5 TB/s  >  cudaMemsetAsync 7.3 TB/s  >  persistent TMA or st.global.v4 grid-stride loops 6.3 TB/s.
__global__ void __launch_bounds__(128) k_zero_fill_tma(unsigned char* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char ztile[];
    for (int i = threadIdx.x; i < kFillBytes / 16; i += 128) reinterpret_cast<uint4*>(ztile)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x != 0) return;
    bulk_store_s2g(out + (size_t)blockIdx.x * kFillBytes, ztile, kFillBytes);
    bulk_commit();
    bulk_wait_read<0>();             // the tile must stay valid until the TMA engine has read it
}

// ---- atomic scatter into the zero-filled grid (no binning) ----------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(kBinThreads) k_scatter_atomic(LidarIn in, uint32_t* __restrict__ out, int bs0) {
    const int bs = bs0 + blockIdx.y, b = bs / in.S;
    int beg, end;
    sweep_range(in, bs, beg, end);
    if ((long long)(MODE == 3 ? (beg & ~7) : beg) + (long long)blockIdx.x * kBinChunk >= end) return;
    __shared__ double sT[12];
    __shared__ int sRect[kMaxRects * 4];
    int nrect; bool flip;
    load_frame_consts(in, bs, b, MODE, sT, sRect, nrect, flip);
    __syncthreads();
    // grid x extent = the caller's max_sweep_points hint; longer sweeps are still covered (loop normally runs once)
    const uint32_t org = MODE == 3 ? ((uint32_t)beg & ~7u) : (uint32_t)beg;
#pragma unroll 1
    for (uint32_t chunk0 = org + blockIdx.x * (uint32_t)kBinChunk; chunk0 < (uint32_t)end; chunk0 += gridDim.x * (uint32_t)kBinChunk) {
        uint32_t* addr[kBinItems];
        uint32_t val[kBinItems];
        if constexpr (MODE == 3) {
            static_assert(kBinItems == 4, "two threads share one 8-row column group");
            // thread t takes rows [4t, 4t+4) of the chunk: the 16 B / 8 B vectors are fetched by the pair of threads that share
            // them (the second fetch hits L1), which keeps the chunk geometry of the row formats
            const uint32_t i8 = chunk0 + (uint32_t)(threadIdx.x >> 1) * 8u;
            const ColGroup grp = load_cols(in, i8, i8 < (uint32_t)end);
            const int h = (threadIdx.x & 1) * 4;
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const uint32_t i = i8 + (uint32_t)(h + it);
                Voxel v;
                addr[it] = nullptr; val[it] = 0;
                const RawPoint<0> rp = h ? col_point(grp, 4 + it) : col_point(grp, it);
                if (i >= (uint32_t)beg && i < (uint32_t)end && point_voxel<0>(rp, sT, flip, sRect, nrect, v)) {
                    addr[it] = out + ((size_t)(bs * kZ + v.zb) * kH + v.py) * kW + v.px;
                    val[it] = v.bits;
                }
            }
        } else {
            RawPoint<MODE> raw[kBinItems];
#pragma unroll
            for (int it = 0; it < kBinItems; ++it) {
                const uint32_t i = chunk0 + (uint32_t)(it * kBinThreads + threadIdx.x);
                raw[it] = load_point<MODE>(in, (long long)i, i < (uint32_t)end);
            }
#pragma unroll
            for (int it = 0; it < kBinItems; ++it) {
                const uint32_t i = chunk0 + (uint32_t)(it * kBinThreads + threadIdx.x);
                Voxel v;
                addr[it] = nullptr; val[it] = 0;
                if (i < (uint32_t)end && point_voxel<MODE>(raw[it], sT, flip, sRect, nrect, v)) {
                    addr[it] = out + ((size_t)(bs * kZ + v.zb) * kH + v.py) * kW + v.px;
                    val[it] = v.bits;
                }
            }
        }
#pragma unroll
        for (int it = 0; it < kBinItems; ++it)
            if (addr[it]) atomicMax(addr[it], val[it]);
    }   // chunk loop
}

template <typename T>
__global__ void __launch_bounds__(256) k_transform_points(const T* __restrict__ pts, long long n, long long stride,
                                                          Pose34 P, double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    double ox, oy, oz;
    xform_point(P.m, (double)pts[i * stride], (double)pts[i * stride + 1], (double)pts[i * stride + 2], ox, oy, oz);
    out[3 * i] = ox; out[3 * i + 1] = oy; out[3 * i + 2] = oz;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// k_bin_fixed launch: big CTAs unless they would not even fill two waves of the GPU's 2 x 148 CTA slots
template <int MODE>
int launch_bin(const LidarIn& in, int n_sweeps, int max_sweep_points, uint32_t* cnt, uint2* bins, uint32_t* ovf_cnt, OvfRec* ovf,
               int hints, cudaStream_t st) {
    const int pad = MODE == 3 ? 7 : 0;           // column vectors are 8-row aligned: a sweep's first chunk may start up to 7 rows early
    const long long big_ctas = (long long)((max_sweep_points + pad + AggCfg<kAggItemsBig>::kChunk - 1) / AggCfg<kAggItemsBig>::kChunk) * n_sweeps;
    static const int forced = [] { const char* e = getenv("BEVGEN_BIN_ITEMS"); return e ? atoi(e) : 0; }();   // measurements only
    const bool small = forced ? forced == kAggItemsSmall : big_ctas < 4 * 148;
    if (small) {
        using C = AggCfg<kAggItemsSmall>;
        BEVGEN_DYN_SMEM((k_bin_fixed<MODE, kAggItemsSmall>), C::kSmem);
        const dim3 grid((max_sweep_points + pad + C::kChunk - 1) / C::kChunk, n_sweeps);
        BEVGEN_LAUNCH("k_bin_fixed", st, (k_bin_fixed<MODE, kAggItemsSmall><<<grid, kAggThreads, C::kSmem, st>>>(in, cnt, bins, ovf_cnt, ovf, 0, hints)));
    } else {
        using C = AggCfg<kAggItemsBig>;
        BEVGEN_DYN_SMEM((k_bin_fixed<MODE, kAggItemsBig>), C::kSmem);
        const dim3 grid((max_sweep_points + pad + C::kChunk - 1) / C::kChunk, n_sweeps);
        BEVGEN_LAUNCH("k_bin_fixed", st, (k_bin_fixed<MODE, kAggItemsBig><<<grid, kAggThreads, C::kSmem, st>>>(in, cnt, bins, ovf_cnt, ovf, 0, hints)));
    }
    return 0;
}

// L2 residency hints for the tiled path: the record bins are written evict_last and the output tiles evict_first, so that
// k_tile_store finds its records in L2 instead of reading HBM in the middle of its own write stream.  Pays only while the
// bins of the whole launch fit in L2 next to the stream -- measured on B200, k_tile_store us/frame without -> with hints:
// 2 frames 54.0 -> 48.2, 4: 51.9 -> 46.3, 8: 50.6 -> 46.7, 16: 49.8 -> 48.8, 32: 49.3 -> 49.9 (profiles/r2_exp_l2_hints.txt).
// With 4-byte records (AV2 intensities) the bins are half as large and the hints pay up to 32 frames: 16: 48.5 -> 46.5,
// 32: 47.8 -> 47.5, 40: 47.7 -> 47.7, 48: 47.7 -> 48.0 (profiles/r2_exp_rec4.txt).
// BEVGEN_L2_HINTS=0/1 forces them off / on for measurements.
int l2_hints_for(int sweeps_in_launch, bool rec4) {
    static const int forced = [] { const char* e = getenv("BEVGEN_L2_HINTS"); return e ? atoi(e) : -1; }();
    if (forced >= 0) return forced != 0;
    return sweeps_in_launch <= (rec4 ? 320 : 160);
}

// workspace of the tiled algorithm: [cnt u32 per tile | ovf_cnt] [bins: kBinCap records per tile] [overflow list]
struct TiledWorkspace {
    size_t cnt, bins, ovf, cnt2, bins2, bytes;
};

// n_resampled > 0 adds the per-(channel, output tile) bins of the fused generation+augmentation path (lidar_aug_tile.cuh)
TiledWorkspace tiled_layout(long long total_points, int B, int S, int n_resampled = 0) {
    TiledWorkspace w;
    const size_t ntiles = (size_t)B * S * kTilesPerSweep;
    size_t off = 0;
    w.cnt = off;  off += align_up((ntiles + (size_t)B * S + 64) * 4, 256);      // per-tile counts, then per-sweep overflow counts
    w.bins = off; off += align_up(ntiles * kBinCap * sizeof(uint2), 256);
    w.ovf = off;  off += align_up((size_t)total_points * sizeof(OvfRec), 256);
    const size_t nbins2 = (size_t)n_resampled * S * kZ * kATTiles;
    w.cnt2 = off;  off += align_up(nbins2 * 4, 256);
    w.bins2 = off; off += align_up(nbins2 * kATCap * sizeof(uint2), 256);
    w.bytes = off;
    return w;
}

// total_points = rows the sweeps of this launch add up to (sizes the overflow lists); store_rows = rows of the point array
// the ranges are clamped to (-1: the same -- CSR form, where the sweeps tile the array)
template <int MODE>
int launch_lidar(LidarIn in, int B, int S, long long total_points, int max_sweep_points, float* out,
                 void* workspace, size_t workspace_bytes, int algo, cudaStream_t st, long long store_rows = -1) {
    if (store_rows < 0) store_rows = total_points;
    if (B <= 0 || S <= 0 || total_points < 0 || total_points > 0x7FFFFFFFll || store_rows > 0x7FFFFFFFll || max_sweep_points < 0 ||
        !out || !in.sweep_off)
        return BEVGEN_E_INVALID_ARG;
    in.n_points = (int)store_rows;
    if ((in.rects == nullptr) != (in.nrects == nullptr)) return BEVGEN_E_INVALID_ARG;
    if ((reinterpret_cast<uintptr_t>(out) & 15u) != 0) return BEVGEN_E_INVALID_ARG;      // TMA bulk stores need 16 B alignment
    const size_t ntiles = (size_t)B * S * kTilesPerSweep;
    if (ntiles > 0x7FFFFFFFull) return BEVGEN_E_UNSUPPORTED;
    const int pad = MODE == 3 ? 7 : 0;         // column vectors are 8-row aligned: a sweep's first chunk may start up to 7 rows early
    const int chunks = max_sweep_points > 0 ? (max_sweep_points + pad + kBinChunk - 1) / kBinChunk : 0;
    const dim3 bin_grid(chunks > 0 ? chunks : 1, B * S);
    if (bin_grid.y > 65535u) return BEVGEN_E_UNSUPPORTED;

    if (algo == BEVGEN_ALGO_ATOMIC || algo == BEVGEN_ALGO_MEMSET_ATOMIC) {
        // zero fill (our TMA fill kernel, or cudaMemsetAsync as the library comparison point) + one atomic scatter launch
        if (algo == BEVGEN_ALGO_MEMSET_ATOMIC) {
            timing_begin("cudaMemsetAsync", st);
            BEVGEN_CUDA(cudaMemsetAsync(out, 0, (size_t)B * S * kZ * kPlane * sizeof(float), st));
            timing_end(st);
        } else {
            BEVGEN_LAUNCH("k_zero_fill_tma", st,
                          k_zero_fill_tma<<<(unsigned)ntiles, 128, kFillBytes, st>>>(reinterpret_cast<unsigned char*>(out)));
        }
        if (chunks > 0) {
            BEVGEN_LAUNCH("k_scatter_atomic", st,
                          k_scatter_atomic<MODE><<<bin_grid, kBinThreads, 0, st>>>(in, reinterpret_cast<uint32_t*>(out), 0));
        }
        return 0;
    }
    if (algo == BEVGEN_ALGO_TILED) {
        // one binning pass into fixed-capacity tile bins, one raster CTA per tile (smem tile + ONE TMA store), overflow REDs
        const TiledWorkspace L = tiled_layout(total_points, B, S);
        if (!workspace || workspace_bytes < L.bytes) return BEVGEN_E_WORKSPACE;
        const int hints = l2_hints_for(B * S, kRec4<MODE>);
        char* wp = static_cast<char*>(workspace);
        uint32_t* cnt = reinterpret_cast<uint32_t*>(wp + L.cnt);
        uint32_t* ovf_cnt = cnt + ntiles;
        uint2* bins = reinterpret_cast<uint2*>(wp + L.bins);
        OvfRec* ovf = reinterpret_cast<OvfRec*>(wp + L.ovf);
        BEVGEN_CUDA(cudaMemsetAsync(cnt, 0, (ntiles + (size_t)B * S + 64) * 4, st));
        // (Measured: splitting the batch into frame groups, with or without overlapping the binning of group g+1 with
        //  the stores of group g on a second stream, is slower than these three whole-batch launches.)
        if (chunks > 0) {
            const int rc = launch_bin<MODE>(in, B * S, max_sweep_points, cnt, bins, ovf_cnt, ovf, hints, st);
            if (rc) return rc;
        }
        BEVGEN_LAUNCH("k_tile_store", st, k_tile_store<<<(unsigned)ntiles, 128, kFillBytes, st>>>(cnt, bins, out, 0u, hints, (int)kRec4<MODE>));
        if (chunks > 0)
            BEVGEN_LAUNCH("k_overflow_red", st, k_overflow_red<<<dim3(kOvfCtas, B * S), 256, 0, st>>>(
                              ovf_cnt, ovf, in.sweep_off, reinterpret_cast<uint32_t*>(out), nullptr, S));
        return 0;
    }
    return BEVGEN_E_INVALID_ARG;
}

// Generation + augmentation of a whole batch, every output byte written once (bevgen_lidar_bev_augmented):
//   k_bin_fixed over all frames (flip fused for every frame; dropout rectangles fused for the frames that are not resampled)
//   k_tile_store + k_overflow_red for the runs of frames that neither rotate nor scale
//   k_tile_aug<ROT,SCL> per class for the others
template <int MODE>
int launch_lidar_augmented(LidarIn in, int B, int S, long long total_points, int max_sweep_points,
                           const bevgen_aug_params* params_host, float* out, void* workspace, size_t workspace_bytes,
                           cudaStream_t st, long long store_rows = -1) {
    if (store_rows < 0) store_rows = total_points;
    if (B <= 0 || S <= 0 || total_points < 0 || total_points > 0x7FFFFFFFll || store_rows > 0x7FFFFFFFll || max_sweep_points < 0 ||
        !out || !in.sweep_off || !in.aug || !params_host)
        return BEVGEN_E_INVALID_ARG;
    in.n_points = (int)store_rows;
    if ((reinterpret_cast<uintptr_t>(out) & 15u) != 0) return BEVGEN_E_INVALID_ARG;
    const size_t ntiles = (size_t)B * S * kTilesPerSweep;
    if (ntiles > 0x7FFFFFFFull || (unsigned)(B * S) > 65535u || B > 32767) return BEVGEN_E_UNSUPPORTED;
    int n_resampled = 0;
    for (int b = 0; b < B; ++b) {
        if (!aug_tile_supported(params_host[b])) return BEVGEN_E_UNSUPPORTED;
        n_resampled += (params_host[b].do_rotate != 0 || params_host[b].do_scale != 0) ? 1 : 0;
    }
    const TiledWorkspace L = tiled_layout(total_points, B, S, n_resampled);
    if (!workspace || workspace_bytes < L.bytes) return BEVGEN_E_WORKSPACE;
    const int hints = l2_hints_for(B * S, kRec4<MODE>);
    char* wp = static_cast<char*>(workspace);
    uint32_t* cnt = reinterpret_cast<uint32_t*>(wp + L.cnt);
    uint32_t* ovf_cnt = cnt + ntiles;
    uint2* bins = reinterpret_cast<uint2*>(wp + L.bins);
    OvfRec* ovf = reinterpret_cast<OvfRec*>(wp + L.ovf);
    uint32_t* cnt2 = reinterpret_cast<uint32_t*>(wp + L.cnt2);
    uint2* bins2 = reinterpret_cast<uint2*>(wp + L.bins2);
    BEVGEN_CUDA(cudaMemsetAsync(cnt, 0, (ntiles + (size_t)B * S + 64) * 4, st));
    BEVGEN_DYN_SMEM((k_tile_aug<true, false>), (AugTileCfg<true, false>::kSmem));
    BEVGEN_DYN_SMEM((k_tile_aug<false, true>), (AugTileCfg<false, true>::kSmem));
    BEVGEN_DYN_SMEM((k_tile_aug<true, true>), (AugTileCfg<true, true>::kSmem));
    const bool have_points = max_sweep_points > 0;
    if (have_points) {
        const int rc = launch_bin<MODE>(in, B * S, max_sweep_points, cnt, bins, ovf_cnt, ovf, hints, st);
        if (rc) return rc;
    }
    auto cls_of = [&](int b) { return (params_host[b].do_rotate != 0 ? 1 : 0) | (params_host[b].do_scale != 0 ? 2 : 0); };
    // runs of frames that are not resampled: plain tile stores
    bool any_plain = false;
    for (int b = 0; b < B;) {
        if (cls_of(b) != 0) { ++b; continue; }
        int e = b;
        while (e < B && cls_of(e) == 0) ++e;
        const unsigned t0 = (unsigned)((size_t)b * S * kTilesPerSweep), nt = (unsigned)((size_t)(e - b) * S * kTilesPerSweep);
        BEVGEN_LAUNCH("k_tile_store", st, k_tile_store<<<nt, 128, kFillBytes, st>>>(cnt, bins, out, t0, hints, (int)kRec4<MODE>));
        any_plain = true;
        b = e;
    }
    if (any_plain && have_points)
        BEVGEN_LAUNCH("k_overflow_red", st, k_overflow_red<<<dim3(kOvfCtas, B * S), 256, 0, st>>>(
                          ovf_cnt, ovf, in.sweep_off, reinterpret_cast<uint32_t*>(out), in.aug, S));
    CUtensorMap out_map;                                     // [B*S*29, 400, 720] view of `out`, 40 x 72 boxes (k_tile_aug's store)
    if (n_resampled > 0) {
        const int rc = make_grid_tensor_map(&out_map, out, (uint64_t)B * S * kZ, kATW, kATH);
        if (rc) return rc;
    }
    // slots of the per-(channel, output tile) bins, in class order; ONE re-binning launch for all resampled frames
    std::vector<short> slot_of(B, (short)-1);
    {
        int next_slot = 0;
        for (int cls = 1; cls < 4; ++cls)
            for (int b = 0; b < B; ++b)
                if (cls_of(b) == cls) slot_of[b] = (short)next_slot++;
        AugTileFrames fl;
        fl.n = 0;
        auto flush_rebin = [&]() -> int {
            if (fl.n == 0) return 0;
            BEVGEN_LAUNCH("k_rebin_aug", st, k_rebin_aug<<<dim3(kZ, S, fl.n), kRebinThreads, 0, st>>>(
                              cnt, bins, ovf_cnt, ovf, in.sweep_off, cnt2, bins2, in.aug, fl, S, (int)kRec4<MODE>));
            fl.n = 0;
            return 0;
        };
        for (int b = 0; b < B; ++b) {
            if (cls_of(b) == 0) continue;
            fl.src[fl.n] = fl.dst[fl.n] = (short)b;
            fl.slot[fl.n] = slot_of[b];
            if (++fl.n == 60) { const int rc = flush_rebin(); if (rc) return rc; }
        }
        const int rc = flush_rebin();
        if (rc) return rc;
    }
    for (int cls = 1; cls < 4; ++cls) {
        AugTileFrames fl;
        fl.n = 0;
        auto flush = [&]() -> int {
            if (fl.n == 0) return 0;
            const dim3 grid(kATTiles, S, fl.n);
            switch (cls) {
                case 1:
                    BEVGEN_LAUNCH("k_tile_aug", st, k_tile_aug<true, false><<<grid, kATThreads, AugTileCfg<true, false>::kSmem, st>>>(
                                      cnt, bins, ovf_cnt, ovf, in.sweep_off, cnt2, bins2, out_map, in.aug, fl, S, hints, (int)kRec4<MODE>));
                    break;
                case 2:
                    BEVGEN_LAUNCH("k_tile_aug", st, k_tile_aug<false, true><<<grid, kATThreads, AugTileCfg<false, true>::kSmem, st>>>(
                                      cnt, bins, ovf_cnt, ovf, in.sweep_off, cnt2, bins2, out_map, in.aug, fl, S, hints, (int)kRec4<MODE>));
                    break;
                default:
                    BEVGEN_LAUNCH("k_tile_aug", st, k_tile_aug<true, true><<<grid, kATThreads, AugTileCfg<true, true>::kSmem, st>>>(
                                      cnt, bins, ovf_cnt, ovf, in.sweep_off, cnt2, bins2, out_map, in.aug, fl, S, hints, (int)kRec4<MODE>));
                    break;
            }
            fl.n = 0;
            return 0;
        };
        for (int b = 0; b < B; ++b) {
            if (cls_of(b) != cls) continue;
            fl.src[fl.n] = fl.dst[fl.n] = (short)b;
            fl.slot[fl.n] = slot_of[b];
            if (++fl.n == 60) { const int rc = flush(); if (rc) return rc; }
        }
        const int rc = flush();
        if (rc) return rc;
    }
    return 0;
}

}  // namespace

}  // namespace bevgen

using namespace bevgen;

#ifdef BEVGEN_AT_PROFILE
extern "C" int bevgen_debug_at_profile(unsigned long long* out16, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out16, g_at_prof, sizeof(unsigned long long) * 16);
    if (reset) { unsigned long long z[16] = {}; cudaMemcpyToSymbol(g_at_prof, z, sizeof z); }
    return 0;
}
#endif

namespace bevgen { namespace {
__global__ void __launch_bounds__(256) k_overflow_total(const uint32_t* __restrict__ ovf_cnt, int n, uint32_t* __restrict__ out) {
    __shared__ uint32_t s;
    if (threadIdx.x == 0) s = 0;
    __syncthreads();
    uint32_t v = 0;
    for (int i = threadIdx.x; i < n; i += 256) v += ovf_cnt[i];
    atomicAdd(&s, v);
    __syncthreads();
    if (threadIdx.x == 0) *out = s;
}
} }

// Number of records the last TILED launch on `workspace` (same total_points, B, S) sent to the overflow lists, written to
// the device word out_dev on `stream`: what a caller needs to decide whether a scene is dense enough for ALGO_ATOMIC.
extern "C" int bevgen_lidar_overflow_total(const void* workspace, size_t workspace_bytes, int64_t total_points, int B, int S,
                                           uint32_t* out_dev, bevgen_stream_t stream) {
    if (!workspace || !out_dev || total_points < 0 || B <= 0 || S <= 0) return BEVGEN_E_INVALID_ARG;
    const TiledWorkspace L = tiled_layout(total_points, B, S);
    if (workspace_bytes < L.bytes) return BEVGEN_E_WORKSPACE;
    const uint32_t* cnt = reinterpret_cast<const uint32_t*>(static_cast<const char*>(workspace) + L.cnt);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    BEVGEN_LAUNCH("k_overflow_total", st, (k_overflow_total<<<1, 256, 0, st>>>(cnt + (size_t)B * S * kTilesPerSweep, B * S, out_dev)));
    return 0;
}

extern "C" size_t bevgen_lidar_aug_workspace_bytes(int64_t total_points, int B, int S, int n_resampled) {
    if (total_points < 0 || B <= 0 || S <= 0 || n_resampled < 0 || n_resampled > B) return 0;
    return tiled_layout(total_points, B, S, n_resampled).bytes;
}

extern "C" int bevgen_aug_fused_supported(const bevgen_aug_params* params_host, int B) {
    if (!params_host || B < 0) return 0;
    for (int b = 0; b < B; ++b)
        if (!aug_tile_supported(params_host[b])) return 0;
    return 1;
}

extern "C" int bevgen_lidar_bev_augmented(const void* pts, int pts_f16, const int32_t* sweep_offsets, const double* rel_tf,
                                          int B, int S, int64_t total_points, int max_sweep_points,
                                          const bevgen_aug_params* params, const bevgen_aug_params* params_host, float* out,
                                          void* workspace, size_t workspace_bytes, bevgen_stream_t stream) {
    if (total_points > 0 && (!pts || !rel_tf)) return BEVGEN_E_INVALID_ARG;
    if ((reinterpret_cast<uintptr_t>(pts) & (pts_f16 ? 7u : 15u)) != 0) return BEVGEN_E_INVALID_ARG;
    LidarIn in = make_in(pts_f16 ? static_cast<const uint2*>(pts) : nullptr, pts_f16 ? nullptr : static_cast<const float4*>(pts),
                         rel_tf, nullptr, nullptr, sweep_offsets, nullptr, nullptr, nullptr, S, params);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return pts_f16 ? launch_lidar_augmented<2>(in, B, S, total_points, max_sweep_points, params_host, out, workspace, workspace_bytes, st)
                   : launch_lidar_augmented<0>(in, B, S, total_points, max_sweep_points, params_host, out, workspace, workspace_bytes, st);
}

extern "C" size_t bevgen_lidar_workspace_bytes(int64_t total_points, int B, int S) {
    if (total_points < 0 || B <= 0 || S <= 0) return 0;
    return tiled_layout(total_points, B, S).bytes;
}

extern "C" int bevgen_lidar_bev_fused(const float* pts4, const int32_t* sweep_offsets, const double* rel_tf, int B, int S,
                                      int64_t total_points, int max_sweep_points, const uint8_t* flip,
                                      const int32_t* drop_rects, const int32_t* drop_counts, float* out, void* workspace,
                                      size_t workspace_bytes, int algo, bevgen_stream_t stream) {
    if (total_points > 0 && (!pts4 || !rel_tf)) return BEVGEN_E_INVALID_ARG;
    if ((reinterpret_cast<uintptr_t>(pts4) & 15u) != 0) return BEVGEN_E_INVALID_ARG;
    LidarIn in = make_in(nullptr, reinterpret_cast<const float4*>(pts4), rel_tf, nullptr, nullptr, sweep_offsets, flip, drop_rects, drop_counts, S);
    return launch_lidar<0>(in, B, S, total_points, max_sweep_points, out, workspace, workspace_bytes, algo,
                           static_cast<cudaStream_t>(stream));
}

extern "C" int bevgen_lidar_bev_fused_f16(const void* pts_h4, const int32_t* sweep_offsets, const double* rel_tf, int B, int S,
                                          int64_t total_points, int max_sweep_points, const uint8_t* flip,
                                          const int32_t* drop_rects, const int32_t* drop_counts, float* out, void* workspace,
                                          size_t workspace_bytes, int algo, bevgen_stream_t stream) {
    if (total_points > 0 && (!pts_h4 || !rel_tf)) return BEVGEN_E_INVALID_ARG;
    if ((reinterpret_cast<uintptr_t>(pts_h4) & 7u) != 0) return BEVGEN_E_INVALID_ARG;
    LidarIn in = make_in(static_cast<const uint2*>(pts_h4), nullptr, rel_tf, nullptr, nullptr, sweep_offsets, flip, drop_rects, drop_counts, S);
    return launch_lidar<2>(in, B, S, total_points, max_sweep_points, out, workspace, workspace_bytes, algo,
                           static_cast<cudaStream_t>(stream));
}

namespace {
// AV2 column arrays: 16-byte aligned (vector loads), and -- because a thread reads whole 8-row vectors -- allocated up to
// the next multiple of 8 rows past total_points (the Python layer pads; the padding rows are never used as points).
bool cols_ok(const void* x, const void* y, const void* z, const void* i, int64_t total_points) {
    if (total_points == 0) return true;
    if (!x || !y || !z || !i) return false;
    return ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(z) |
             reinterpret_cast<uintptr_t>(i)) & 15u) == 0;
}
LidarIn make_cols(const void* x, const void* y, const void* z, const uint8_t* i, const double* rel_tf, const int32_t* off,
                  const uint8_t* flip, const int32_t* rects, const int32_t* nrects, int S, const bevgen_aug_params* aug) {
    LidarIn in = make_in(nullptr, nullptr, rel_tf, nullptr, nullptr, off, flip, rects, nrects, S, aug);
    in.col_x = static_cast<const __half*>(x); in.col_y = static_cast<const __half*>(y); in.col_z = static_cast<const __half*>(z);
    in.col_i = i;
    return in;
}
}  // namespace

extern "C" int bevgen_lidar_bev_fused_cols(const void* x_f16, const void* y_f16, const void* z_f16, const uint8_t* intensity_u8,
                                           const int32_t* sweep_offsets, const double* rel_tf, int B, int S, int64_t total_points,
                                           int max_sweep_points, const uint8_t* flip, const int32_t* drop_rects,
                                           const int32_t* drop_counts, float* out, void* workspace, size_t workspace_bytes,
                                           int algo, bevgen_stream_t stream) {
    if (total_points > 0 && !rel_tf) return BEVGEN_E_INVALID_ARG;
    if (!cols_ok(x_f16, y_f16, z_f16, intensity_u8, total_points)) return BEVGEN_E_INVALID_ARG;
    return launch_lidar<3>(make_cols(x_f16, y_f16, z_f16, intensity_u8, rel_tf, sweep_offsets, flip, drop_rects, drop_counts, S, nullptr),
                           B, S, total_points, max_sweep_points, out, workspace, workspace_bytes, algo, static_cast<cudaStream_t>(stream));
}

extern "C" int bevgen_lidar_bev_augmented_cols(const void* x_f16, const void* y_f16, const void* z_f16, const uint8_t* intensity_u8,
                                               const int32_t* sweep_offsets, const double* rel_tf, int B, int S,
                                               int64_t total_points, int max_sweep_points, const bevgen_aug_params* params,
                                               const bevgen_aug_params* params_host, float* out, void* workspace,
                                               size_t workspace_bytes, bevgen_stream_t stream) {
    if (total_points > 0 && !rel_tf) return BEVGEN_E_INVALID_ARG;
    if (!cols_ok(x_f16, y_f16, z_f16, intensity_u8, total_points)) return BEVGEN_E_INVALID_ARG;
    return launch_lidar_augmented<3>(make_cols(x_f16, y_f16, z_f16, intensity_u8, rel_tf, sweep_offsets, nullptr, nullptr, nullptr, S, params),
                                     B, S, total_points, max_sweep_points, params_host, out, workspace, workspace_bytes,
                                     static_cast<cudaStream_t>(stream));
}

// ---- sweep STORE form: the column arrays are a resident store of many sweeps; a frame names its sweeps by start row -------
extern "C" int bevgen_lidar_bev_store_cols(const void* x_f16, const void* y_f16, const void* z_f16, const uint8_t* intensity_u8,
                                           int64_t store_rows, const int32_t* sweep_start, const int32_t* sweep_offsets,
                                           const double* rel_tf, int B, int S, int64_t total_points, int max_sweep_points,
                                           const uint8_t* flip, const int32_t* drop_rects, const int32_t* drop_counts, float* out,
                                           void* workspace, size_t workspace_bytes, int algo, bevgen_stream_t stream) {
    if (store_rows < 0 || !sweep_start || (total_points > 0 && !rel_tf)) return BEVGEN_E_INVALID_ARG;
    if (!cols_ok(x_f16, y_f16, z_f16, intensity_u8, store_rows)) return BEVGEN_E_INVALID_ARG;
    LidarIn in = make_cols(x_f16, y_f16, z_f16, intensity_u8, rel_tf, sweep_offsets, flip, drop_rects, drop_counts, S, nullptr);
    in.sweep_start = sweep_start;
    return launch_lidar<3>(in, B, S, total_points, max_sweep_points, out, workspace, workspace_bytes, algo,
                           static_cast<cudaStream_t>(stream), store_rows);
}

extern "C" int bevgen_lidar_bev_store_augmented_cols(const void* x_f16, const void* y_f16, const void* z_f16,
                                                     const uint8_t* intensity_u8, int64_t store_rows, const int32_t* sweep_start,
                                                     const int32_t* sweep_offsets, const double* rel_tf, int B, int S,
                                                     int64_t total_points, int max_sweep_points, const bevgen_aug_params* params,
                                                     const bevgen_aug_params* params_host, float* out, void* workspace,
                                                     size_t workspace_bytes, bevgen_stream_t stream) {
    if (store_rows < 0 || !sweep_start || (total_points > 0 && !rel_tf)) return BEVGEN_E_INVALID_ARG;
    if (!cols_ok(x_f16, y_f16, z_f16, intensity_u8, store_rows)) return BEVGEN_E_INVALID_ARG;
    LidarIn in = make_cols(x_f16, y_f16, z_f16, intensity_u8, rel_tf, sweep_offsets, nullptr, nullptr, nullptr, S, params);
    in.sweep_start = sweep_start;
    return launch_lidar_augmented<3>(in, B, S, total_points, max_sweep_points, params_host, out, workspace, workspace_bytes,
                                     static_cast<cudaStream_t>(stream), store_rows);
}

extern "C" int bevgen_lidar_bev_ego_f64(const double* pts3, const float* intensity, const int32_t* sweep_offsets, int B, int S,
                                        int64_t total_points, int max_sweep_points, const uint8_t* flip,
                                        const int32_t* drop_rects, const int32_t* drop_counts, float* out, void* workspace,
                                        size_t workspace_bytes, int algo, bevgen_stream_t stream) {
    if (total_points > 0 && (!pts3 || !intensity)) return BEVGEN_E_INVALID_ARG;
    LidarIn in = make_in(nullptr, nullptr, nullptr, pts3, intensity, sweep_offsets, flip, drop_rects, drop_counts, S);
    return launch_lidar<1>(in, B, S, total_points, max_sweep_points, out, workspace, workspace_bytes, algo,
                           static_cast<cudaStream_t>(stream));
}

static int launch_transform(const void* pts, bool f64, int64_t n, int64_t stride, const double* T_host, double* out,
                            cudaStream_t st) {
    if (n < 0 || stride < 3 || !T_host) return BEVGEN_E_INVALID_ARG;
    if (n == 0) return 0;
    if (!pts || !out) return BEVGEN_E_INVALID_ARG;
    Pose34 P;
    for (int i = 0; i < 12; ++i) P.m[i] = T_host[i];
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (f64) BEVGEN_LAUNCH("k_transform_points", st, k_transform_points<double><<<grid, 256, 0, st>>>(static_cast<const double*>(pts), n, stride, P, out));
    else     BEVGEN_LAUNCH("k_transform_points", st, k_transform_points<float><<<grid, 256, 0, st>>>(static_cast<const float*>(pts), n, stride, P, out));
    return 0;
}

extern "C" int bevgen_transform_points_f32(const float* pts, int64_t n, int64_t stride, const double* T_host, double* out,
                                           bevgen_stream_t stream) {
    return launch_transform(pts, false, n, stride, T_host, out, static_cast<cudaStream_t>(stream));
}
extern "C" int bevgen_transform_points_f64(const double* pts, int64_t n, int64_t stride, const double* T_host, double* out,
                                           bevgen_stream_t stream) {
    return launch_transform(pts, true, n, stride, T_host, out, static_cast<cudaStream_t>(stream));
}
