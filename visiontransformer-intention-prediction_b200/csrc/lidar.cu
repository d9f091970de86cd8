// lidar.cu -- multi-sweep LiDAR aggregation -> height-binned max-intensity BEV grid (sm_100a).
//
// Replaces, for a whole batch of frames in four launches:
//   transform_points            reference utils.py:27-33   (per sweep, dataset.py:339-340)
//   create_intentnet_lidar_bev  reference utils.py:62-106
//
// Data flow (HBM):   points [N,4] f32 --K_bin_count--> tmp records [N] u64 + row histogram
//                    --K_bin_scan--> row-bin starts --K_bin_scatter--> records sorted by (frame,sweep,row)
//                    --K_raster_store--> out [B, S*29, 400, 720] f32 (every byte written exactly once)
//
// A "tile" is one (frame, sweep, row): 29 height channels x 720 columns = 83 520 B.  The raster kernel
// keeps two such tiles in shared memory per SM, applies the tile's records with warp-aggregated
// (__match_any_sync) shared-memory atomicMax, and flushes the tile with 29 TMA bulk stores
// (cp.async.bulk.global.shared::cta, 2 880 B each = one channel row).  Zero fill is the untouched part
// of the tile, so no memset pass and no read-modify-write of the 334 MB/frame output ever happens.
#include "bevgen_common.cuh"

#include <cstdlib>
#include <vector>

namespace bevgen {

namespace {

constexpr uint32_t kInvalid = 0xFFFFFFFFu;
constexpr int kBinThreads = 256;
constexpr int kBinItems = 4;
constexpr int kBinChunk = kBinThreads * kBinItems;     // points per CTA in the two binning passes
constexpr int kRasterThreads = 256;
constexpr int kTileBytes = kRowTileElems * 4;          // 83 520
constexpr int kRowBytes = kW * 4;                      // 2 880 (multiple of 16: bulk-copy legal)

struct LidarIn {
    const float4* pts4;        // MODE 0
    const double* rel_tf;      // MODE 0: [B*S,16]
    const double* pts3;        // MODE 1
    const float* intensity;    // MODE 1
    const int32_t* sweep_off;  // [B*S+1]
    const uint8_t* flip;       // [B] or null
    const int32_t* rects;      // [B,5,4] or null
    const int32_t* nrects;     // [B] or null
    int S;
};

// Raw point as loaded from HBM.  Loads are issued for ALL of a thread's items before the first use, so
// kBinItems independent 16 B (MODE 0) / 28 B (MODE 1) requests are in flight per thread.
template <int MODE> struct RawPoint;
template <> struct RawPoint<0> { float4 p; };
template <> struct RawPoint<1> { double x, y, z; float v; };

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

// One point -> (row, cell-in-row-tile, intensity bits) or invalid.  MODE 0 transforms first.
template <int MODE>
__device__ __forceinline__ bool point_record(const RawPoint<MODE>& raw, const double* sT, bool flip,
                                             const int* sRect, int nrect, int& row, uint32_t& cell, uint32_t& bits);
template <>
__device__ __forceinline__ bool point_record<0>(const RawPoint<0>& raw, const double* sT, bool flip,
                                                const int* sRect, int nrect, int& row, uint32_t& cell, uint32_t& bits) {
    // max with the +0.0 initial grid: only strictly positive intensities can change a voxel
    // (documented domain: finite or +inf, >= 0; then float max == unsigned max on the bit pattern)
    if (!(raw.p.w > 0.0f)) return false;
    double x, y, z;
    xform_point(sT, (double)raw.p.x, (double)raw.p.y, (double)raw.p.z, x, y, z);
    int py, px, zb;
    if (!lidar_voxel(x, y, z, py, px, zb)) return false;
    if (flip) px = kW - 1 - px;                           // utils.py:399 fused
    if (nrect && in_rects(sRect, nrect, py, px)) return false;   // utils.py:496 fused
    row = py;
    cell = (uint32_t)(zb * kW + px);
    bits = __float_as_uint(raw.p.w);
    return true;
}
template <>
__device__ __forceinline__ bool point_record<1>(const RawPoint<1>& raw, const double*, bool flip,
                                                const int* sRect, int nrect, int& row, uint32_t& cell, uint32_t& bits) {
    if (!(raw.v > 0.0f)) return false;
    int py, px, zb;
    if (!lidar_voxel(raw.x, raw.y, raw.z, py, px, zb)) return false;
    if (flip) px = kW - 1 - px;
    if (nrect && in_rects(sRect, nrect, py, px)) return false;
    row = py;
    cell = (uint32_t)(zb * kW + px);
    bits = __float_as_uint(raw.v);
    return true;
}

__device__ __forceinline__ void load_frame_consts(const LidarIn& in, int bs, int b, int mode, double* sT, int* sRect,
                                                  int& nrect, bool& flip) {
    if (mode == 0 && threadIdx.x < 12) sT[threadIdx.x] = in.rel_tf[(size_t)bs * 16 + threadIdx.x];
    nrect = in.nrects ? in.nrects[b] : 0;
    nrect = nrect > kMaxRects ? kMaxRects : nrect;
    if ((int)threadIdx.x < 4 * nrect) sRect[threadIdx.x] = in.rects[(size_t)b * kMaxRects * 4 + threadIdx.x];
    flip = in.flip ? (in.flip[b] != 0) : false;
}

// ---- pass 1: per-point record + per-(frame,sweep,row) histogram --------------------------------
template <int MODE>
__global__ void __launch_bounds__(kBinThreads) k_bin_count(LidarIn in, uint2* __restrict__ tmp, uint32_t* __restrict__ counts) {
    const int bs = blockIdx.y, b = bs / in.S;
    const int beg = in.sweep_off[bs], end = in.sweep_off[bs + 1];
    const long long chunk0 = (long long)beg + (long long)blockIdx.x * kBinChunk;
    if (chunk0 >= end) return;
    __shared__ double sT[12];
    __shared__ int sRect[kMaxRects * 4];
    __shared__ uint32_t hist[kH];
    int nrect; bool flip;
    load_frame_consts(in, bs, b, MODE, sT, sRect, nrect, flip);
    for (int r = threadIdx.x; r < kH; r += kBinThreads) hist[r] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    RawPoint<MODE> raw[kBinItems];
#pragma unroll
    for (int it = 0; it < kBinItems; ++it) {
        const long long i = chunk0 + it * kBinThreads + threadIdx.x;
        raw[it] = load_point<MODE>(in, i, i < end);
    }
#pragma unroll
    for (int it = 0; it < kBinItems; ++it) {
        const long long i = chunk0 + it * kBinThreads + threadIdx.x;
        int row = 0; uint32_t cell = 0, bits = 0;
        const bool ok = (i < end) && point_record<MODE>(raw[it], sT, flip, sRect, nrect, row, cell, bits);
        // warp-aggregated histogram: one shared atomic per distinct row per warp
        const uint32_t key = ok ? (uint32_t)row : 0xFFFFu;
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, key);
        if (ok && lane == __ffs(peers) - 1) atomicAdd(&hist[row], (uint32_t)__popc(peers));
        if (i < end) tmp[i] = make_uint2(ok ? (((uint32_t)row << 15) | cell) : kInvalid, bits);
    }
    __syncthreads();
    for (int r = threadIdx.x; r < kH; r += kBinThreads)
        if (hist[r]) atomicAdd(&counts[(size_t)bs * kH + r], hist[r]);
}

// ---- pass 1b: exclusive scan of each sweep's 400 row counts -> absolute starts in `sorted` ------
__global__ void __launch_bounds__(512) k_bin_scan(const int32_t* __restrict__ sweep_off, const uint32_t* __restrict__ counts,
                                                  uint32_t* __restrict__ starts) {
    const int bs = blockIdx.x, t = threadIdx.x, lane = t & 31, w = t >> 5;
    __shared__ uint32_t wsum[16];
    const uint32_t c = t < kH ? counts[(size_t)bs * kH + t] : 0u;
    uint32_t v = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, v, d); if (lane >= d) v += n; }
    if (lane == 31) wsum[w] = v;
    __syncthreads();
    if (w == 0) {
        uint32_t s = lane < 16 ? wsum[lane] : 0u;
#pragma unroll
        for (int d = 1; d < 16; d <<= 1) { const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, s, d); if (lane >= d) s += n; }
        if (lane < 16) wsum[lane] = s;
    }
    __syncthreads();
    if (t < kH) starts[(size_t)bs * kH + t] = (uint32_t)sweep_off[bs] + (w ? wsum[w - 1] : 0u) + v - c;
}

// ---- pass 2: scatter records into their row bins ------------------------------------------------
__global__ void __launch_bounds__(kBinThreads) k_bin_scatter(const int32_t* __restrict__ sweep_off, int S,
                                                             const uint2* __restrict__ tmp, const uint32_t* __restrict__ starts,
                                                             uint32_t* __restrict__ cursor, uint2* __restrict__ sorted) {
    const int bs = blockIdx.y;
    const int beg = sweep_off[bs], end = sweep_off[bs + 1];
    const long long chunk0 = (long long)beg + (long long)blockIdx.x * kBinChunk;
    if (chunk0 >= end) return;
    __shared__ uint32_t hist[kH];
    __shared__ uint32_t base[kH];
    for (int r = threadIdx.x; r < kH; r += kBinThreads) hist[r] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    uint2 rec[kBinItems];
    uint32_t lrank[kBinItems];
#pragma unroll
    for (int it = 0; it < kBinItems; ++it) {
        const long long i = chunk0 + it * kBinThreads + threadIdx.x;
        rec[it] = (i < end) ? tmp[i] : make_uint2(kInvalid, 0u);
        const bool ok = rec[it].x != kInvalid;
        const uint32_t key = ok ? (rec[it].x >> 15) : 0xFFFFu;
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, key);
        lrank[it] = 0;
        if (ok) {   // claim-then-broadcast: the group leader reserves popc(peers) slots of the CTA-local bin
            const int leader = __ffs(peers) - 1;
            uint32_t b0 = 0;
            if (lane == leader) b0 = atomicAdd(&hist[key], (uint32_t)__popc(peers));
            b0 = __shfl_sync(peers, b0, leader);
            lrank[it] = b0 + (uint32_t)__popc(peers & lt);
        }
    }
    __syncthreads();
    for (int r = threadIdx.x; r < kH; r += kBinThreads)
        if (hist[r]) base[r] = starts[(size_t)bs * kH + r] + atomicAdd(&cursor[(size_t)bs * kH + r], hist[r]);
    __syncthreads();
#pragma unroll
    for (int it = 0; it < kBinItems; ++it)
        if (rec[it].x != kInvalid)
            sorted[base[rec[it].x >> 15] + lrank[it]] = make_uint2(rec[it].x & 0x7FFFu, rec[it].y);
}

// ---- pass 3: shared-memory tile raster + TMA bulk store (persistent, double buffered) -----------
__global__ void __launch_bounds__(kRasterThreads, 1) k_raster_store(const uint2* __restrict__ sorted,
                                                                    const uint32_t* __restrict__ starts,
                                                                    const uint32_t* __restrict__ counts,
                                                                    float* __restrict__ out, int num_tiles) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint32_t* const buf0 = reinterpret_cast<uint32_t*>(smem_raw);                   // two contiguous tiles
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < 2 * kRowTileElems; i += kRasterThreads) buf0[i] = 0u;
    uint32_t prev_start0 = 0u, prev_cnt0 = 0u, prev_start1 = 0u, prev_cnt1 = 0u;   // last occupant of each buffer
    __syncthreads();

    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        const int cur = it & 1;
        uint32_t* const tile = buf0 + cur * kRowTileElems;
        const uint32_t start = starts[t], cnt = counts[t];
        // the store issued from this buffer two tiles ago must have finished READING shared memory
        if (tid < kZ) bulk_wait_read<1>();
        __syncthreads();
        // restore the all-zero tile: clear exactly the cells the previous occupant touched
        const uint32_t pstart = cur ? prev_start1 : prev_start0, pcnt = cur ? prev_cnt1 : prev_cnt0;
        for (uint32_t i = tid; i < pcnt; i += kRasterThreads) tile[sorted[pstart + i].x] = 0u;
        __syncthreads();
        // apply this tile's records: warp-aggregated max (dedupe equal cells inside the warp first)
        for (uint32_t i0 = 0; i0 < cnt; i0 += kRasterThreads) {
            const uint32_t i = i0 + tid;
            const bool have = i < cnt;
            uint2 r = make_uint2(0xFFFFFFFFu, 0u);
            if (have) r = sorted[start + i];
            const uint32_t peers = __match_any_sync(0xFFFFFFFFu, r.x);
            const uint32_t m = __reduce_max_sync(peers, r.y);
            if (have && lane == __ffs(peers) - 1) atomicMax(&tile[r.x], m);
        }
        fence_proxy_async_smem();     // generic-proxy writes -> visible to the async proxy (TMA)
        __syncthreads();
        if (tid < kZ) {
            // tile t = (frame*S + sweep)*400 + row ; channel = (frame*S + sweep)*29 + zbin
            const int bs = t / kH, row = t - bs * kH;
            float* dst = out + ((size_t)(bs * kZ + tid) * kH + row) * kW;
            bulk_store_s2g(dst, tile + tid * kW, kRowBytes);
            bulk_commit();
        }
        if (cur) { prev_start1 = start; prev_cnt1 = cnt; } else { prev_start0 = start; prev_cnt0 = cnt; }
    }
    if (tid < kZ) bulk_wait_all<0>();
}

// ---- fused persistent kernel: TMA zero fill + L2-resident atomic scatter --------------------------
// One cooperative launch for the whole batch; warp-specialised CTAs:
//   producer warps (kProducers, lane 0 only): walk the fill chunks g = 0 .. B*S*1160-1; each chunk is ONE TMA
//       bulk store of an immutable all-zero shared-memory tile (kFillBytes) to out + g*kFillBytes -- no LSU
//       traffic, no registers.  When the store has completed (bulk wait_group) the producer publishes it:
//       proxy fence + __threadfence + atomicAdd(done[sweep of g]).
//   consumer warps: walk the scatter chunks (kScatChunk points each, sweep-major).  A sweep region (29 channels
//       = 33.4 MB = 1160 fill chunks) may only be scattered into after ALL its fill chunks have landed:
//       ld.acquire spin on done[js] == 1160.  Then transform, bin and RED.MAX into out, and publish
//       scat[js] += 1 per finished chunk.
//   flow control: a producer may fill sweep j only once sweep j - kWindow has been completely scattered
//       (scat[j - kWindow] == chunks per sweep).  This keeps the fill front at most kWindow sweep regions
//       (66.8 MB) ahead of the scatter front, so the atomics hit lines that are still dirty in the 126 MB L2
//       (measured: 61 G RED/s on L2-resident lines vs 22 G RED/s when every RED has to fetch its sector from
//       HBM, tools/redbench.cu).
// Dependencies are acyclic (fill j <- scatter j-2 <- fill j-2 <- ...), producers publish everything they hold
// before they block, all CTAs are co-resident (cooperative launch): no deadlock.  Every spin is bounded.
constexpr int kFillRows = 10;
constexpr int kFillBytes = kFillRows * kRowBytes;                    // 28 800 B
constexpr int kFillPerSweep = kZ * kH / kFillRows;                   // 1160 chunks per (frame, sweep)
constexpr int kProducers = 4;                                        // producer warps per CTA
constexpr int kConsumerThreads = 128;
constexpr int kFusedThreads = kProducers * 32 + kConsumerThreads;    // 256
constexpr int kScatItems = 4;                                        // points per consumer thread per chunk
constexpr int kScatChunk = kConsumerThreads * kScatItems;            // 512 points
constexpr int kWindow = 2;
constexpr unsigned kSpinLimit = 20000000u;

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" :: "n"(kConsumerThreads) : "memory"); }

// returns false on timeout (never hang the GPU)
__device__ __forceinline__ bool spin_until(const uint32_t* ctr, uint32_t target, uint32_t* err_flag) {
    unsigned spins = 0;
    while (ld_acquire_u32(ctr) < target) {
        __nanosleep(64);
        if (++spins > kSpinLimit) { atomicExch(err_flag, 1u); return false; }
    }
    return true;
}

template <int MODE, bool DEDUPE>
__global__ void __launch_bounds__(kFusedThreads) k_fill_scatter(LidarIn in, uint32_t* __restrict__ out,
                                                                uint32_t* __restrict__ done, uint32_t* __restrict__ scat,
                                                                uint32_t* __restrict__ err_flag,
                                                                int num_sweeps_total, int chunks_per_sweep) {
    extern __shared__ __align__(128) unsigned char zero_tile[];
    __shared__ double sT[12];
    __shared__ int sRect[kMaxRects * 4];
    __shared__ int s_ok;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < kFillBytes / 16; i += kFusedThreads) reinterpret_cast<uint4*>(zero_tile)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) s_ok = 1;          // sticky: only ever cleared (on a spin timeout)
    fence_proxy_async_smem();
    __syncthreads();

    if (warp < kProducers) {
        // ===================== producer: TMA zero fill =====================
        if (lane != 0) return;
        const long long total_fill = (long long)num_sweeps_total * kFillPerSweep;
        const long long stride = (long long)gridDim.x * kProducers;
        long long prev = -1;
        int cleared = kWindow - 1;           // sweeps <= cleared may be filled without asking
        bool ok = true;
        auto publish_prev = [&](bool all) {
            if (prev < 0) return;
            if (all) bulk_wait_all<0>(); else bulk_wait_all<1>();
            asm volatile("fence.proxy.async;" ::: "memory");
            __threadfence();
            atomicAdd(&done[prev / kFillPerSweep], 1u);
            prev = -1;
        };
        for (long long g = (long long)blockIdx.x * kProducers + warp; g < total_fill && ok; g += stride) {
            const int j = (int)(g / kFillPerSweep);
            if (j > cleared) {               // flow control: sweep j - kWindow must be fully scattered
                publish_prev(true);          // never block while holding an unpublished chunk
                ok = spin_until(&scat[j - kWindow], (uint32_t)chunks_per_sweep, err_flag);
                cleared = j;
            }
            bulk_store_s2g(reinterpret_cast<unsigned char*>(out) + g * kFillBytes, zero_tile, kFillBytes);
            bulk_commit();
            const long long mine = g;
            publish_prev(false);             // one store stays in flight while the previous one is published
            prev = mine;
        }
        publish_prev(true);
        return;
    }

    // ===================== consumers: transform + bin + RED.MAX =====================
    const int ctid = tid - kProducers * 32;
    const long long total_scat = (long long)num_sweeps_total * chunks_per_sweep;
    int cur_bs = -1;                 // sweep whose constants are in shared memory
    int ready_bs = -1;               // sweep known to be completely filled
    for (long long hs = blockIdx.x; hs < total_scat; hs += gridDim.x) {
        const int bs = (int)(hs / chunks_per_sweep);
        const int c = (int)(hs - (long long)bs * chunks_per_sweep);
        const int beg = in.sweep_off[bs], end = in.sweep_off[bs + 1];
        const long long chunk0 = (long long)beg + (long long)c * kScatChunk;
        if (chunk0 < end) {
            // issue the point loads first: their latency overlaps the wait for the fill
            RawPoint<MODE> raw[kScatItems];
#pragma unroll
            for (int it = 0; it < kScatItems; ++it) {
                const long long i = chunk0 + it * kConsumerThreads + ctid;
                raw[it] = load_point<MODE>(in, i, i < end);
            }
            if (bs != ready_bs) {
                if (ctid == 0 && !spin_until(&done[bs], (uint32_t)kFillPerSweep, err_flag)) s_ok = 0;
                consumer_sync();                                 // also orders the previous chunk's sT readers
                ready_bs = bs;
            }
            if (bs != cur_bs) {
                if (MODE == 0 && ctid < 12) sT[ctid] = in.rel_tf[(size_t)bs * 16 + ctid];
                const int b0 = bs / in.S;
                int nr = in.nrects ? in.nrects[b0] : 0;
                nr = nr > kMaxRects ? kMaxRects : nr;
                if (ctid < 4 * nr) sRect[ctid] = in.rects[(size_t)b0 * kMaxRects * 4 + ctid];
                cur_bs = bs;
                consumer_sync();
            }
            const int b = bs / in.S;
            int nrect = in.nrects ? in.nrects[b] : 0;
            nrect = nrect > kMaxRects ? kMaxRects : nrect;
            const bool flip = in.flip ? (in.flip[b] != 0) : false;
            if (s_ok) {
                uint32_t* region = out + (size_t)bs * kZ * kPlane;
#pragma unroll
                for (int it = 0; it < kScatItems; ++it) {
                    const long long i = chunk0 + it * kConsumerThreads + ctid;
                    int row = 0; uint32_t cell = 0, bits = 0;
                    const bool ok = (i < end) && point_record<MODE>(raw[it], sT, flip, sRect, nrect, row, cell, bits);
                    const uint32_t zb = cell / kW, px = cell - zb * kW;
                    const uint32_t vox = (zb * kH + (uint32_t)row) * kW + px;
                    if (DEDUPE) {      // warp-aggregated: one RED per distinct voxel per warp
                        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, ok ? vox : 0xFFFFFFFFu);
                        const uint32_t m = __reduce_max_sync(peers, bits);
                        if (ok && lane == __ffs(peers) - 1) atomicMax(region + vox, m);
                    } else if (ok) {
                        atomicMax(region + vox, bits);
                    }
                }
            }
            consumer_sync();                                     // every thread's REDs are issued
        }
        if (ctid == 0) atomicAdd(&scat[bs], 1u);                 // publish: this chunk of sweep bs is scattered
    }
}

// ---- pass 3'': one short-lived CTA per tile, shared-memory-privatised tile + TMA bulk store ---------------------
// The CTA zero-fills a [29 x 720] float tile in shared memory, applies the tile's records with warp-aggregated
// (__match_any_sync + __reduce_max_sync) shared-memory atomicMax, and flushes the tile with 29 TMA bulk stores
// (one 2 880 B channel row each).  No global atomics at all; every output byte is written exactly once.
// Two CTAs are resident per SM (2 x 83.5 KB of shared memory), so one tile is being built while the other drains.
constexpr int kTileSmemThreads = 256;
__global__ void __launch_bounds__(kTileSmemThreads, 2) k_tile_smem_store(const uint2* __restrict__ sorted,
                                                                         const uint32_t* __restrict__ starts,
                                                                         const uint32_t* __restrict__ counts,
                                                                         float* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint32_t* const tile = reinterpret_cast<uint32_t*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31;
    const int t = blockIdx.x;
    const int bs = t / kH, row = t - bs * kH;
    const uint32_t start = starts[t], cnt = counts[t];
    uint2 first = make_uint2(0xFFFFFFFFu, 0u);
    if ((uint32_t)tid < cnt) first = sorted[start + tid];     // in flight while the tile is being zeroed
    for (int i = tid; i < kRowTileElems / 4; i += kTileSmemThreads) reinterpret_cast<uint4*>(tile)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    for (uint32_t i0 = 0; i0 < cnt; i0 += kTileSmemThreads) {
        const uint32_t i = i0 + tid;
        uint2 r = make_uint2(0xFFFFFFFFu, 0u);
        if (i0 == 0) r = first; else if (i < cnt) r = sorted[start + i];
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, r.x);
        const uint32_t m = __reduce_max_sync(peers, r.y);
        if (r.x != 0xFFFFFFFFu && lane == __ffs(peers) - 1) atomicMax(&tile[r.x], m);
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (tid < kZ) {
        float* dst = out + ((size_t)(bs * kZ + tid) * kH + row) * kW;
        bulk_store_s2g(dst, tile + tid * kW, kRowBytes);
        bulk_commit();
        bulk_wait_read<0>();
    }
}

// ---- pass 3': one short-lived CTA per tile: TMA zero fill of the tile, then its records as L2-resident REDs ---
// Tile = (frame, sweep, row) = 29 channel rows of 2 880 B.  The CTA bulk-stores ONE shared zero row 29 times
// (cp.async.bulk.global.shared::cta, one store per channel row), waits for the stores to complete, and then
// applies the tile's own records (from the row-binned list) with RED.MAX.  The lines were written by this very
// CTA microseconds earlier, so the atomics hit the L2 and never read HBM (no read-modify-write traffic).
// No shared-memory tile -> ~16 CTAs per SM in flight, which is what the short-lived TMA fill needs to reach
// ~7.5 TB/s (tools/fillbench.cu).
constexpr int kTileThreads = 128;
template <bool DEDUPE>
__global__ void __launch_bounds__(kTileThreads) k_tile_fill_red(const uint2* __restrict__ sorted, const uint32_t* __restrict__ starts,
                                                                const uint32_t* __restrict__ counts, uint32_t* __restrict__ out) {
    __shared__ __align__(128) uint4 zrow[kRowBytes / 16];
    const int tid = threadIdx.x, lane = tid & 31;
    const int t = blockIdx.x;                                 // tile id = (frame*S + sweep)*400 + row
    const int bs = t / kH, row = t - bs * kH;
    for (int i = tid; i < kRowBytes / 16; i += kTileThreads) zrow[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
    const uint32_t start = starts[t], cnt = counts[t];
    uint2 first = make_uint2(0xFFFFFFFFu, 0u);
    if ((uint32_t)tid < cnt) first = sorted[start + tid];     // record load in flight during the fill
    __syncthreads();
    uint32_t* tile0 = out + ((size_t)bs * kZ * kH + row) * kW;   // channel stride = kPlane
    if (tid < kZ) {
        bulk_store_s2g(tile0 + (size_t)tid * kPlane, zrow, kRowBytes);
        bulk_commit();
        bulk_wait_all<0>();                                   // zero rows have landed
        asm volatile("fence.proxy.async;" ::: "memory");
    }
    if (cnt == 0) return;                                     // uniform
    __syncthreads();
    for (uint32_t i0 = 0; i0 < cnt; i0 += kTileThreads) {
        const uint32_t i = i0 + tid;
        uint2 r = make_uint2(0xFFFFFFFFu, 0u);
        if (i0 == 0) r = first; else if (i < cnt) r = sorted[start + i];
        const bool have = r.x != 0xFFFFFFFFu;
        const uint32_t zb = r.x / kW, px = r.x - zb * kW;
        if (DEDUPE) {                                         // warp-aggregated: one RED per distinct voxel per warp
            const uint32_t peers = __match_any_sync(0xFFFFFFFFu, r.x);
            const uint32_t m = __reduce_max_sync(peers, r.y);
            if (have && lane == __ffs(peers) - 1) atomicMax(tile0 + (size_t)zb * kPlane + px, m);
        } else if (have) {
            atomicMax(tile0 + (size_t)zb * kPlane + px, r.y);
        }
    }
}

// ---- "binned": single binning pass into fixed-capacity tile bins + one short-lived raster CTA per tile ---------
// Tile = (frame, sweep, z-bin, band of 10 rows) = 7 200 floats = 28 800 CONTIGUOUS bytes of the output, i.e. exactly
// one chunk of the fast TMA fill kernel.  1160 tiles per sweep.
//   k_bin_fixed      per point: transform, bin, slot = atomicAdd(cnt[tile]); slot < kBinCap -> bins[tile][slot] = record,
//                    else -> overflow list (dense scenes).  A full bin is detected with a plain load first, so a hot
//                    voxel costs at most kBinCap counter atomics.  One pass, no histogram scan, no intermediate copy.
//   k_tile_store     one 128-thread CTA per tile: zero a 28.8 KB shared-memory tile, apply the bin's records with
//                    warp-aggregated (__match_any_sync + __reduce_max_sync) shared atomicMax, ONE TMA bulk store.
//                    The record loads do not depend on the count load (fixed bin address), 7 CTAs are resident per SM.
//   k_overflow_red   the (rare) overflow records as global RED.MAX after the tiles have landed.
// Every output byte is written exactly once by the TMA; no global read-modify-write except for overflow records.
constexpr int kBinCap = 128;                                         // records per tile bin (avg 40 at 100k pts/sweep)
constexpr int kTilesPerSweep = kFillPerSweep;                        // 1160
constexpr int kTileElems = kFillRows * kW;                           // 7200 floats per tile
struct OvfRec { unsigned long long idx; uint32_t bits; uint32_t pad; };

template <int MODE>
__global__ void __launch_bounds__(kBinThreads) k_bin_fixed(LidarIn in, uint32_t* __restrict__ cnt, uint2* __restrict__ bins,
                                                           uint32_t* __restrict__ ovf_cnt, OvfRec* __restrict__ ovf) {
    const int bs = blockIdx.y, b = bs / in.S;
    const int beg = in.sweep_off[bs], end = in.sweep_off[bs + 1];
    const long long chunk0 = (long long)beg + (long long)blockIdx.x * kBinChunk;
    if (chunk0 >= end) return;
    __shared__ double sT[12];
    __shared__ int sRect[kMaxRects * 4];
    int nrect; bool flip;
    load_frame_consts(in, bs, b, MODE, sT, sRect, nrect, flip);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    RawPoint<MODE> raw[kBinItems];
#pragma unroll
    for (int it = 0; it < kBinItems; ++it) {
        const long long i = chunk0 + it * kBinThreads + threadIdx.x;
        raw[it] = load_point<MODE>(in, i, i < end);
    }
#pragma unroll
    for (int it = 0; it < kBinItems; ++it) {
        const long long i = chunk0 + it * kBinThreads + threadIdx.x;
        int row = 0; uint32_t cell = 0, bits = 0;
        const bool ok = (i < end) && point_record<MODE>(raw[it], sT, flip, sRect, nrect, row, cell, bits);
        const uint32_t zb = cell / kW, px = cell - zb * kW;
        const uint32_t band = (uint32_t)row / kFillRows;
        const uint32_t tile = (uint32_t)bs * kTilesPerSweep + zb * (kH / kFillRows) + band;
        const uint32_t tcell = ((uint32_t)row - band * kFillRows) * kW + px;
        // warp-aggregated slot claim: one counter atomic per distinct tile per warp
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, ok ? tile : 0xFFFFFFFFu);
        uint32_t slot = 0xFFFFFFFFu;
        if (ok) {
            const int leader = __ffs(peers) - 1;
            uint32_t base = 0;
            if (lane == leader) {
                base = __ldcg(cnt + tile);                                   // cheap test: bin already full?
                if (base < (uint32_t)kBinCap) base = atomicAdd(cnt + tile, (uint32_t)__popc(peers));
            }
            base = __shfl_sync(peers, base, leader);
            slot = base + (uint32_t)__popc(peers & lt);
        }
        const bool spill = ok && slot >= (uint32_t)kBinCap;
        if (ok && !spill) bins[(size_t)tile * kBinCap + slot] = make_uint2(tcell, bits);
        // overflow: one list-append atomic per warp
        const uint32_t sp = __ballot_sync(0xFFFFFFFFu, spill);
        if (sp) {
            uint32_t obase = 0;
            if (lane == __ffs(sp) - 1) obase = atomicAdd(ovf_cnt, (uint32_t)__popc(sp));
            obase = __shfl_sync(0xFFFFFFFFu, obase, __ffs(sp) - 1);
            if (spill) {
                OvfRec r;
                r.idx = (unsigned long long)tile * kTileElems + tcell;
                r.bits = bits; r.pad = 0;
                ovf[obase + __popc(sp & lt)] = r;
            }
        }
    }
}

__global__ void __launch_bounds__(128) k_tile_store(const uint32_t* __restrict__ cnt, const uint2* __restrict__ bins,
                                                    float* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char ztile[];
    uint32_t* const tile = reinterpret_cast<uint32_t*>(ztile);
    const int tid = threadIdx.x, lane = tid & 31;
    const size_t t = blockIdx.x;
    const uint2 rec = __ldcs(bins + t * kBinCap + tid);          // speculative: independent of the count load
    uint32_t n = __ldcs(cnt + t);
    for (int i = tid; i < kFillBytes / 16; i += 128) reinterpret_cast<uint4*>(ztile)[i] = make_uint4(0, 0, 0, 0);
    n = n < (uint32_t)kBinCap ? n : (uint32_t)kBinCap;
    __syncthreads();
    if (n) {                                                    // uniform
        const bool have = (uint32_t)tid < n;
        const uint32_t key = have ? rec.x : 0xFFFFFFFFu;
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, key);
        const uint32_t m = __reduce_max_sync(peers, rec.y);
        if (have && lane == __ffs(peers) - 1) atomicMax(&tile[rec.x], m);
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (tid != 0) return;
    bulk_store_s2g(reinterpret_cast<unsigned char*>(out) + t * kFillBytes, ztile, kFillBytes);
    bulk_commit();
    bulk_wait_read<0>();
}

__global__ void __launch_bounds__(256) k_overflow_red(const uint32_t* __restrict__ ovf_cnt, const OvfRec* __restrict__ ovf,
                                                      uint32_t* __restrict__ out) {
    const uint32_t n = *ovf_cnt;
    for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        const OvfRec r = ovf[i];
        atomicMax(out + r.idx, r.bits);
    }
}

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

// ---- ticketed single-launch fill + scatter --------------------------------------------------------------
// One ordinary launch for the whole batch.  Every CTA draws a ticket (atomicAdd) that fixes its role:
//   per sweep (frame-major):  kTicketFill fill CTAs (kTicketStores TMA bulk stores of a zero tile each, then
//                             done[sweep] += 1 once the store has completed), followed by
//                             chunks_per_sweep scatter CTAs (kBinChunk points each: transform, bin, RED.MAX).
// A scatter CTA spins until done[sweep] == kTicketFill.  Tickets are handed out in start order, so every CTA a
// scatter CTA waits for has already started and never waits itself: deadlock-free without a cooperative
// launch.  The hardware's bounded CTA residency is the flow control: the fill front cannot run more than about
// one sweep region ahead, so the atomics hit lines that are still dirty in L2 (61 G RED/s instead of 22).
constexpr int kTicketRows = 5;
constexpr int kTicketBytes = kTicketRows * kRowBytes;                // 14 400 B zero tile
constexpr int kTicketStores = 4;                                     // bulk stores per fill CTA (57.6 KB contiguous)
constexpr int kTicketFill = kZ * kH / (kTicketRows * kTicketStores); // 580 fill CTAs per (frame, sweep)
constexpr int kTicketThreads = 128;
constexpr int kTicketItems = 4;
constexpr int kTicketChunk = kTicketThreads * kTicketItems;          // 512 points per scatter CTA

template <int MODE>
__global__ void __launch_bounds__(kTicketThreads, 8) k_fill_scatter_ticket(LidarIn in, uint32_t* __restrict__ out,
                                                                        uint32_t* __restrict__ ticket,
                                                                        uint32_t* __restrict__ done,
                                                                        uint32_t* __restrict__ err_flag, int chunks_per_sweep) {
    extern __shared__ __align__(128) unsigned char ztile[];
    __shared__ uint32_t s_ticket;
    __shared__ int s_ok;
    __shared__ double sT[12];
    __shared__ int sRect[kMaxRects * 4];
    const int tid = threadIdx.x;
    if (tid == 0) { s_ticket = atomicAdd(ticket, 1u); s_ok = 1; }
    for (int i = tid; i < kTicketBytes / 16; i += kTicketThreads) reinterpret_cast<uint4*>(ztile)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
    __syncthreads();
    const uint32_t per_sweep = (uint32_t)(kTicketFill + chunks_per_sweep);
    const uint32_t bs = s_ticket / per_sweep, k = s_ticket - bs * per_sweep;

    if (k < (uint32_t)kTicketFill) {
        // ---------------- fill CTA ----------------
        if (tid != 0) return;
        unsigned char* dst = reinterpret_cast<unsigned char*>(out) + ((size_t)bs * kTicketFill + k) * (size_t)(kTicketBytes * kTicketStores);
#pragma unroll
        for (int q = 0; q < kTicketStores; ++q) bulk_store_s2g(dst + (size_t)q * kTicketBytes, ztile, kTicketBytes);
        bulk_commit();
        bulk_wait_all<0>();                                  // writes performed
        asm volatile("fence.proxy.async;" ::: "memory");
        __threadfence();
        atomicAdd(&done[bs], 1u);
        return;
    }
    // ---------------- scatter CTA ----------------
    const int c = (int)(k - kTicketFill);
    const int b = (int)bs / in.S;
    const int beg = in.sweep_off[bs], end = in.sweep_off[bs + 1];
    const long long chunk0 = (long long)beg + (long long)c * kTicketChunk;
    if (chunk0 >= end) return;
    int nrect; bool flip;
    load_frame_consts(in, (int)bs, b, MODE, sT, sRect, nrect, flip);
    RawPoint<MODE> raw[kTicketItems];
#pragma unroll
    for (int it = 0; it < kTicketItems; ++it) {     // point loads in flight while we wait for the fill
        const long long i = chunk0 + it * kTicketThreads + tid;
        raw[it] = load_point<MODE>(in, i, i < end);
    }
    if (tid == 0) {
        unsigned spins = 0;
        while (ld_acquire_u32(&done[bs]) < (uint32_t)kTicketFill) {
            __nanosleep(64);
            if (++spins > kSpinLimit) { s_ok = 0; atomicExch(err_flag, 1u); break; }   // never hang the GPU
        }
    }
    __syncthreads();
    if (!s_ok) return;
    uint32_t* region = out + (size_t)bs * kZ * kPlane;
#pragma unroll
    for (int it = 0; it < kTicketItems; ++it) {
        const long long i = chunk0 + it * kTicketThreads + tid;
        int row; uint32_t cell, bits;
        if (i < end && point_record<MODE>(raw[it], sT, flip, sRect, nrect, row, cell, bits)) {
            const uint32_t zb = cell / kW, px = cell - zb * kW;
            atomicMax(region + ((size_t)zb * kH + row) * kW + px, bits);
        }
    }
}

// ---- atomic scatter into the zero-filled grid (no binning) ----------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(kBinThreads) k_scatter_atomic(LidarIn in, uint32_t* __restrict__ out, int bs0) {
    const int bs = bs0 + blockIdx.y, b = bs / in.S;
    const int beg = in.sweep_off[bs], end = in.sweep_off[bs + 1];
    const long long chunk0 = (long long)beg + (long long)blockIdx.x * kBinChunk;
    if (chunk0 >= end) return;
    __shared__ double sT[12];
    __shared__ int sRect[kMaxRects * 4];
    int nrect; bool flip;
    load_frame_consts(in, bs, b, MODE, sT, sRect, nrect, flip);
    __syncthreads();
    RawPoint<MODE> raw[kBinItems];
#pragma unroll
    for (int it = 0; it < kBinItems; ++it) {
        const long long i = chunk0 + it * kBinThreads + threadIdx.x;
        raw[it] = load_point<MODE>(in, i, i < end);
    }
    uint32_t* addr[kBinItems];
    uint32_t val[kBinItems];
#pragma unroll
    for (int it = 0; it < kBinItems; ++it) {
        const long long i = chunk0 + it * kBinThreads + threadIdx.x;
        int row; uint32_t cell, bits;
        addr[it] = nullptr; val[it] = 0;
        if (i < end && point_record<MODE>(raw[it], sT, flip, sRect, nrect, row, cell, bits)) {
            const uint32_t zb = cell / kW, px = cell - zb * kW;
            addr[it] = out + ((size_t)(bs * kZ + zb) * kH + row) * kW + px;
            val[it] = bits;
        }
    }
#pragma unroll
    for (int it = 0; it < kBinItems; ++it)
        if (addr[it]) atomicMax(addr[it], val[it]);
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

// workspace layout: [counts | cursor | starts] (u32 per row bin)  [tmp | sorted] (u64 per point)
struct WorkspaceLayout {
    size_t counts, cursor, starts, tmp, sorted, bytes;
};

WorkspaceLayout layout(long long total_points, int B, int S) {
    WorkspaceLayout w;
    const size_t bins = (size_t)B * S * kH;
    size_t off = 0;
    w.counts = off; off += align_up(bins * 4, 256);
    w.cursor = off; off += align_up(bins * 4, 256);
    w.starts = off; off += align_up(bins * 4, 256);
    w.tmp = off;    off += align_up((size_t)total_points * 8, 256);
    w.sorted = off; off += align_up((size_t)total_points * 8, 256);
    w.bytes = off;
    return w;
}

int g_num_sms = 0;

template <int MODE>
int launch_lidar(const LidarIn& in, int B, int S, long long total_points, int max_sweep_points, float* out,
                 void* workspace, size_t workspace_bytes, int algo, cudaStream_t st) {
    if (B <= 0 || S <= 0 || total_points < 0 || max_sweep_points < 0 || !out || !in.sweep_off) return BEVGEN_E_INVALID_ARG;
    if ((in.rects == nullptr) != (in.nrects == nullptr)) return BEVGEN_E_INVALID_ARG;
    if ((size_t)B * S * kH > 0x7FFFFFFFull) return BEVGEN_E_UNSUPPORTED;
    const int chunks = max_sweep_points > 0 ? (max_sweep_points + kBinChunk - 1) / kBinChunk : 0;
    const dim3 bin_grid(chunks > 0 ? chunks : 1, B * S);
    if (bin_grid.y > 65535u) return BEVGEN_E_UNSUPPORTED;
    if (algo == 1 || algo == 4) {
        // zero fill (1: our TMA fill kernel, 4: cudaMemsetAsync for comparison) + one atomic scatter launch
        if (algo == 4) {
            timing_begin("cudaMemsetAsync", st);
            BEVGEN_CUDA(cudaMemsetAsync(out, 0, (size_t)B * S * kZ * kPlane * sizeof(float), st));
            timing_end(st);
        } else {
            if ((reinterpret_cast<uintptr_t>(out) & 15u) != 0) return BEVGEN_E_INVALID_ARG;
            const long long nchunks = (long long)B * S * kFillPerSweep;
            if (nchunks > 0x7FFFFFFFll) return BEVGEN_E_UNSUPPORTED;
            BEVGEN_LAUNCH("k_zero_fill_tma", st,
                          k_zero_fill_tma<<<(unsigned)nchunks, 128, kFillBytes, st>>>(reinterpret_cast<unsigned char*>(out)));
        }
        if (chunks > 0) {
            BEVGEN_LAUNCH("k_scatter_atomic", st,
                          k_scatter_atomic<MODE><<<bin_grid, kBinThreads, 0, st>>>(in, reinterpret_cast<uint32_t*>(out), 0));
        }
        return 0;
    }
    if (algo == 9) {   // binned: one binning pass + per-tile raster CTAs + overflow
        const size_t ntiles = (size_t)B * S * kTilesPerSweep;
        if (ntiles > 0x7FFFFFFFull) return BEVGEN_E_UNSUPPORTED;
        size_t off = 0;
        const size_t o_cnt = off;  off += align_up((ntiles + 64) * 4, 256);
        const size_t o_bins = off; off += align_up(ntiles * kBinCap * sizeof(uint2), 256);
        const size_t o_ovf = off;  off += align_up((size_t)total_points * sizeof(OvfRec), 256);
        if (!workspace || workspace_bytes < off) return BEVGEN_E_WORKSPACE;
        if ((reinterpret_cast<uintptr_t>(out) & 15u) != 0) return BEVGEN_E_INVALID_ARG;
        char* wp = static_cast<char*>(workspace);
        uint32_t* cnt = reinterpret_cast<uint32_t*>(wp + o_cnt);
        uint32_t* ovf_cnt = cnt + ntiles;
        uint2* bins = reinterpret_cast<uint2*>(wp + o_bins);
        OvfRec* ovf = reinterpret_cast<OvfRec*>(wp + o_ovf);
        BEVGEN_CUDA(cudaMemsetAsync(cnt, 0, (ntiles + 64) * 4, st));
        if (chunks > 0) BEVGEN_LAUNCH("k_bin_fixed", st, k_bin_fixed<MODE><<<bin_grid, kBinThreads, 0, st>>>(in, cnt, bins, ovf_cnt, ovf));
        BEVGEN_LAUNCH("k_tile_store", st, k_tile_store<<<(unsigned)ntiles, 128, kFillBytes, st>>>(cnt, bins, out));
        if (chunks > 0) BEVGEN_LAUNCH("k_overflow_red", st, k_overflow_red<<<296, 256, 0, st>>>(ovf_cnt, ovf, reinterpret_cast<uint32_t*>(out)));
        return 0;
    }
    if (algo == 5) {   // ticketed single-launch fill + scatter
        const size_t nsw_sz = (size_t)B * S;
        const size_t need = (nsw_sz + 64) * sizeof(uint32_t);
        if (!workspace || workspace_bytes < need) return BEVGEN_E_WORKSPACE;
        if ((reinterpret_cast<uintptr_t>(out) & 15u) != 0) return BEVGEN_E_INVALID_ARG;
        uint32_t* done = static_cast<uint32_t*>(workspace);
        uint32_t* ticket = done + nsw_sz;
        uint32_t* err_flag = ticket + 1;
        BEVGEN_CUDA(cudaMemsetAsync(done, 0, need, st));
        const int cps = max_sweep_points > 0 ? (max_sweep_points + kTicketChunk - 1) / kTicketChunk : 0;
        const long long total = (long long)nsw_sz * (kTicketFill + cps);
        if (total > 0x7FFFFFFFll) return BEVGEN_E_UNSUPPORTED;
        BEVGEN_LAUNCH("k_fill_scatter_ticket", st,
                      k_fill_scatter_ticket<MODE><<<(unsigned)total, kTicketThreads, kTicketBytes, st>>>(
                          in, reinterpret_cast<uint32_t*>(out), ticket, done, err_flag, cps));
        return 0;
    }
    if (algo == 2 || algo == 3) {   // fused fill + scatter (3 = with warp-level dedupe)
        const size_t nsw_sz = (size_t)B * S;
        const size_t need = (2 * nsw_sz + 64) * sizeof(uint32_t);
        if (!workspace || workspace_bytes < need) return BEVGEN_E_WORKSPACE;
        if ((reinterpret_cast<uintptr_t>(out) & 15u) != 0) return BEVGEN_E_INVALID_ARG;
        uint32_t* done = static_cast<uint32_t*>(workspace);
        uint32_t* scat = done + nsw_sz;
        uint32_t* err_flag = scat + nsw_sz;
        BEVGEN_CUDA(cudaMemsetAsync(done, 0, need, st));
        void (*kern)(LidarIn, uint32_t*, uint32_t*, uint32_t*, uint32_t*, int, int) =
            algo == 3 ? k_fill_scatter<MODE, true> : k_fill_scatter<MODE, false>;
        static int grid_cached[4] = {0, 0, 0, 0};
        int& grid = grid_cached[(algo - 2) * 2 + MODE];
        if (grid == 0) {
            int dev = 0, sms = 0, occ = 0;
            BEVGEN_CUDA(cudaGetDevice(&dev));
            BEVGEN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
            BEVGEN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kFusedThreads, kFillBytes));
            if (occ < 1) return BEVGEN_E_UNSUPPORTED;
            if (occ > 4) occ = 4;
            grid = sms * occ;
        }
        uint32_t* outu = reinterpret_cast<uint32_t*>(out);
        int nsw = B * S, cps = max_sweep_points > 0 ? (max_sweep_points + kScatChunk - 1) / kScatChunk : 1;
        LidarIn in_copy = in;
        void* args[] = {&in_copy, &outu, &done, &scat, &err_flag, &nsw, &cps};
        timing_begin("k_fill_scatter", st);
        BEVGEN_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kern), dim3(grid), dim3(kFusedThreads), args,
                                                (size_t)kFillBytes, st));
        timing_end(st);
        ++g_launch_count;
        return 0;
    }
    if (algo != 0 && algo != 6 && algo != 7 && algo != 8) return BEVGEN_E_INVALID_ARG;
    if (!workspace || workspace_bytes < bevgen_lidar_workspace_bytes(total_points, B, S)) return BEVGEN_E_WORKSPACE;
    if ((reinterpret_cast<uintptr_t>(out) & 15u) != 0) return BEVGEN_E_INVALID_ARG;   // bulk store needs 16 B alignment
    const WorkspaceLayout L = layout(total_points, B, S);
    char* wp = static_cast<char*>(workspace);
    struct { uint32_t *counts, *cursor, *starts; uint2 *tmp, *sorted; } w = {
        reinterpret_cast<uint32_t*>(wp + L.counts), reinterpret_cast<uint32_t*>(wp + L.cursor),
        reinterpret_cast<uint32_t*>(wp + L.starts), reinterpret_cast<uint2*>(wp + L.tmp),
        reinterpret_cast<uint2*>(wp + L.sorted)};
    const size_t bins = (size_t)B * S * kH;
    // counts and cursor are adjacent: one clear
    BEVGEN_CUDA(cudaMemsetAsync(w.counts, 0, L.starts - L.counts, st));
    if (chunks > 0) {
        BEVGEN_LAUNCH("k_bin_count", st, k_bin_count<MODE><<<bin_grid, kBinThreads, 0, st>>>(in, w.tmp, w.counts));
    }
    BEVGEN_LAUNCH("k_bin_scan", st, k_bin_scan<<<B * S, 512, 0, st>>>(in.sweep_off, w.counts, w.starts));
    if (chunks > 0) {
        BEVGEN_LAUNCH("k_bin_scatter", st,
                      k_bin_scatter<<<bin_grid, kBinThreads, 0, st>>>(in.sweep_off, S, w.tmp, w.starts, w.cursor, w.sorted));
    }
    if (algo == 8) {
        if (bins > 0x7FFFFFFFull) return BEVGEN_E_UNSUPPORTED;
        static bool attr8 = false;
        if (!attr8) {
            BEVGEN_CUDA(cudaFuncSetAttribute(k_tile_smem_store, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileBytes));
            attr8 = true;
        }
        BEVGEN_LAUNCH("k_tile_smem_store", st, k_tile_smem_store<<<(unsigned)bins, kTileSmemThreads, kTileBytes, st>>>(
                          w.sorted, w.starts, w.counts, out));
        return 0;
    }
    if (algo == 6 || algo == 7) {
        if (bins > 0x7FFFFFFFull) return BEVGEN_E_UNSUPPORTED;
        if (algo == 7) BEVGEN_LAUNCH("k_tile_fill_red", st, k_tile_fill_red<true><<<(unsigned)bins, kTileThreads, 0, st>>>(
                                         w.sorted, w.starts, w.counts, reinterpret_cast<uint32_t*>(out)));
        else BEVGEN_LAUNCH("k_tile_fill_red", st, k_tile_fill_red<false><<<(unsigned)bins, kTileThreads, 0, st>>>(
                               w.sorted, w.starts, w.counts, reinterpret_cast<uint32_t*>(out)));
        return 0;
    }
    if (g_num_sms == 0) {
        int dev = 0;
        BEVGEN_CUDA(cudaGetDevice(&dev));
        BEVGEN_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
        BEVGEN_CUDA(cudaFuncSetAttribute(k_raster_store, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kTileBytes));
    }
    const int tiles = (int)bins;
    const int grid = tiles < g_num_sms ? tiles : g_num_sms;   // persistent: one CTA per SM
    BEVGEN_LAUNCH("k_raster_store", st,
                  k_raster_store<<<grid, kRasterThreads, 2 * kTileBytes, st>>>(w.sorted, w.starts, w.counts, out, tiles));
    return 0;
}

}  // namespace

}  // namespace bevgen

using namespace bevgen;

extern "C" size_t bevgen_lidar_workspace_bytes(int64_t total_points, int B, int S) {
    if (total_points < 0 || B <= 0 || S <= 0) return 0;
    // max over the algorithms: row-binned CSR layout vs fixed-capacity tile bins + overflow list
    const size_t ntiles = (size_t)B * S * kTilesPerSweep;
    const size_t binned = align_up((ntiles + 64) * 4, 256) + align_up(ntiles * kBinCap * sizeof(uint2), 256) +
                          align_up((size_t)total_points * sizeof(OvfRec), 256);
    const size_t csr = layout(total_points, B, S).bytes;
    return binned > csr ? binned : csr;
}

extern "C" int bevgen_lidar_bev_fused(const float* pts4, const int32_t* sweep_offsets, const double* rel_tf, int B, int S,
                                      int64_t total_points, int max_sweep_points, const uint8_t* flip,
                                      const int32_t* drop_rects, const int32_t* drop_counts, float* out, void* workspace,
                                      size_t workspace_bytes, int algo, bevgen_stream_t stream) {
    if (total_points > 0 && (!pts4 || !rel_tf)) return BEVGEN_E_INVALID_ARG;
    if ((reinterpret_cast<uintptr_t>(pts4) & 15u) != 0) return BEVGEN_E_INVALID_ARG;
    LidarIn in{reinterpret_cast<const float4*>(pts4), rel_tf, nullptr, nullptr, sweep_offsets, flip, drop_rects, drop_counts, S};
    return launch_lidar<0>(in, B, S, total_points, max_sweep_points, out, workspace, workspace_bytes, algo,
                           static_cast<cudaStream_t>(stream));
}

extern "C" int bevgen_lidar_bev_ego_f64(const double* pts3, const float* intensity, const int32_t* sweep_offsets, int B, int S,
                                        int64_t total_points, int max_sweep_points, const uint8_t* flip,
                                        const int32_t* drop_rects, const int32_t* drop_counts, float* out, void* workspace,
                                        size_t workspace_bytes, int algo, bevgen_stream_t stream) {
    if (total_points > 0 && (!pts3 || !intensity)) return BEVGEN_E_INVALID_ARG;
    LidarIn in{nullptr, nullptr, pts3, intensity, sweep_offsets, flip, drop_rects, drop_counts, S};
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
