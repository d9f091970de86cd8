// map_raster.cu -- HD-map polygon / lane-boundary rasterization into the 9 BEV map channels (sm_100a).
//
// Replaces reference utils.py:108-182 (rasterize_map_ego_centric) including the pieces it reaches through
// OpenCV: world_to_bev_pixel (utils.py:47-60), cv2.fillPoly (utils.py:155,157,159,180) and cv2.polylines
// (utils.py:162,164,169,173).  All drawing is "write 1" (union), so primitives are order-free:
//
//   K_map_draw    one warp per primitive (lane segment or crosswalk) per frame:
//                   vertex transform (FP64 FMA chain + IEEE divide + rint, utils.py:52-60), in-grid filter with
//                   order-preserving warp compaction (utils.py:140-144), 8-connected Bresenham lines in closed
//                   form (one pixel per lane), even-odd scanline fill in OpenCV's 16.16 fixed point
//                   (crossings rank-sorted inside the warp), all OR-ed into per-frame bit planes
//                   [9][400][24 x u32] with atomicOr (L2 resident: 345 600 B per frame)
//   K_map_expand  bit planes -> float32 [B,9,400,720] with coalesced float4 stores; horizontal flip and
//                   dropout rectangles (utils.py:399, :496) are fused into this store.
#include "bevgen_common.cuh"

namespace bevgen {
namespace {

constexpr int kWordsPerRow = kMapWordsPerRow;          // 768 bits >= 720 columns
constexpr int kPlaneWords = kMapPlaneWords;            // 9 600 u32 per channel
constexpr int kFrameWords = kMapC * kPlaneWords;       // 86 400 u32 = 345 600 B per frame
constexpr int kDrawWarps = 4;
constexpr int kDrawThreads = kDrawWarps * 32;

struct MapIn {
    const double* verts;       // [V,2]
    const int32_t* pl_off;     // [P+1]
    const int32_t* lanes;      // [L,4]
    const int32_t* cws;        // [CW]
    const int32_t* frame_map;  // [B,4] lane_begin, lane_count, cw_begin, cw_count
    const double* ego_T;       // [B,9]
};

__device__ __forceinline__ void set_px(uint32_t* __restrict__ planes, uint32_t chmask, int x, int y) {
    const uint32_t bit = 1u << (x & 31);
    const int w = y * kWordsPerRow + (x >> 5);
    while (chmask) {
        const int c = __ffs(chmask) - 1;
        chmask &= chmask - 1;
        atomicOr(planes + c * kPlaneWords + w, bit);
    }
}

// cv2 Line(): LineIterator(connectivity 8, left_to_right).  Pixel k of the segment in closed form
// (SURVEY.md A.3): minor-axis offset c_k = floor((2*b*k + a - 1) / (2*a)).
__device__ __forceinline__ void draw_segment(uint32_t* planes, uint32_t chmask, short2 p0, short2 p1, int lane) {
    int x1 = p0.x, y1 = p0.y, dx = p1.x - p0.x, dy = p1.y - p0.y;
    if (dx < 0) { x1 = p1.x; y1 = p1.y; dx = -dx; dy = -dy; }       // iterate left -> right
    int ystep = 1;
    if (dy < 0) { dy = -dy; ystep = -1; }
    const bool steep = dy > dx;
    const int a = steep ? dy : dx, b = steep ? dx : dy;
    for (int k = lane; k <= a; k += 32) {
        const int c = a ? (2 * b * k + a - 1) / (2 * a) : 0;
        const int x = steep ? x1 + c : x1 + k;
        const int y = steep ? y1 + ystep * k : y1 + ystep * c;
        set_px(planes, chmask, x, y);
    }
}

// Transform + round + in-grid filter of one polyline, order-preserving compaction into dst (shared).
__device__ __forceinline__ int compact_polyline(const MapIn& in, const double* T, int pl, short2* dst, int lane) {
    if (pl < 0) return 0;
    const int beg = in.pl_off[pl], end = in.pl_off[pl + 1];
    int n = 0;
    for (int i0 = beg; i0 < end; i0 += 32) {
        const int i = i0 + lane;
        bool ok = false;
        long long px = 0, py = 0;
        if (i < end) {
            map_vertex_px(T, in.verts[2 * (size_t)i], in.verts[2 * (size_t)i + 1], px, py);
            ok = px >= 0 && px < kW && py >= 0 && py < kH;                    // utils.py:140-143
        }
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, ok);
        if (ok) dst[n + __popc(m & ((1u << lane) - 1u))] = make_short2((short)px, (short)py);
        n += __popc(m);
    }
    __syncwarp();
    return n;
}

// polygon vertex p of "left + flipud(right)" (utils.py:153)
__device__ __forceinline__ short2 poly_vertex(const short2* L, int nl, const short2* R, int nr, int p) {
    return p < nl ? L[p] : R[nr - 1 - (p - nl)];
}

// one interior row of fillPoly, the whole warp cooperating (any number of crossings); see fill_polygon
__device__ void fill_row_cooperative(uint32_t* planes, uint32_t chmask, const short2* L, int nl, const short2* R, int nr,
                                     int* xs, int* sorted, int lane, int y) {
    const int n = nl + nr;
    int k = 0;
    for (int e0 = 0; e0 < n; e0 += 32) {
        const int e = e0 + lane;
        bool act = false;
        int xc = 0;
        if (e < n) {
            const short2 p0 = poly_vertex(L, nl, R, nr, e == 0 ? n - 1 : e - 1), p1 = poly_vertex(L, nl, R, nr, e);
            if (p0.y != p1.y) {
                const short2 top = p0.y < p1.y ? p0 : p1, bot = p0.y < p1.y ? p1 : p0;
                if (top.y <= y && y < bot.y) {
                    const int dxfix = (((int)p1.x - (int)p0.x) * 65536) / ((int)p1.y - (int)p0.y);
                    xc = (int)((long long)top.x * 65536 + (long long)(y - top.y) * dxfix);
                    act = true;
                }
            }
        }
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, act);
        if (act) xs[k + __popc(m & ((1u << lane) - 1u))] = xc;
        k += __popc(m);
    }
    __syncwarp();
    if (k >= 2) {
        // rank sort; ties broken by index -> a permutation
        for (int i = lane; i < k; i += 32) {
            const int xi = xs[i];
            int r = 0;
            for (int j = 0; j < k; ++j) { const int xj = xs[j]; r += (xj < xi) || (xj == xi && j < i); }
            sorted[r] = xi;
        }
        __syncwarp();
        for (int pidx = lane; 2 * pidx + 1 < k; pidx += 32) {
            int xa = (sorted[2 * pidx] + 65535) >> 16, xb = sorted[2 * pidx + 1] >> 16;   // ceil left, floor right
            if (xa < kW && xb >= 0) {
                xa = max(xa, 0); xb = min(xb, kW - 1);
                for (int w = xa >> 5; w <= (xb >> 5); ++w) {
                    if (max(xa, w << 5) > min(xb, (w << 5) + 31)) continue;
                    const int lo = max(xa, w << 5) & 31, hi = min(xb, (w << 5) + 31) & 31;
                    const uint32_t bits = (0xFFFFFFFFu >> (31 - hi)) & (0xFFFFFFFFu << lo);
                    uint32_t cm = chmask;
                    while (cm) {
                        const int c = __ffs(cm) - 1;
                        cm &= cm - 1;
                        atomicOr(planes + c * kPlaneWords + y * kWordsPerRow + w, bits);
                    }
                }
            }
        }
    }
    __syncwarp();
}

// cv2.fillPoly of one contour: closed outline + even-odd scanline interior.
__device__ void fill_polygon(uint32_t* planes, uint32_t chmask, const short2* L, int nl, const short2* R, int nr,
                             int* xs, int* sorted, int lane) {
    const int n = nl + nr;
    int ymin = kH, ymax = -1;
    for (int e = lane; e < n; e += 32) {
        const short2 v = poly_vertex(L, nl, R, nr, e);
        ymin = min(ymin, (int)v.y); ymax = max(ymax, (int)v.y);
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        ymin = min(ymin, __shfl_xor_sync(0xFFFFFFFFu, ymin, d));
        ymax = max(ymax, __shfl_xor_sync(0xFFFFFFFFu, ymax, d));
    }
    // outline: edges v[e-1] -> v[e], closed
    for (int e = 0; e < n; ++e)
        draw_segment(planes, chmask, poly_vertex(L, nl, R, nr, e == 0 ? n - 1 : e - 1), poly_vertex(L, nl, R, nr, e), lane);
    // interior: rows [ymin, ymax): an edge is active on [y_top, y_bottom).  One row per lane: every lane walks all edges
    // (vertex reads are shared-memory broadcasts) and keeps its row's crossings sorted in registers; a chunk of 32 rows
    // in which some row has more than kMaxCross crossings falls back to the warp-cooperative form below.
    constexpr int kMaxCross = 8;
    for (int y0 = ymin; y0 < ymax; y0 += 32) {
        const int y = y0 + lane;
        const bool row_on = y < ymax;
        int xr[kMaxCross];
#pragma unroll
        for (int j = 0; j < kMaxCross; ++j) xr[j] = 0x7FFFFFFF;
        int k = 0;
        short2 p0 = poly_vertex(L, nl, R, nr, n - 1);
        for (int e = 0; e < n; ++e) {
            const short2 p1 = poly_vertex(L, nl, R, nr, e);
            if (p0.y != p1.y) {                                               // uniform
                const short2 top = p0.y < p1.y ? p0 : p1, bot = p0.y < p1.y ? p1 : p0;
                if (row_on && top.y <= y && y < bot.y) {
                    // dx = trunc(((x1 - x0) << 16) / (y1 - y0)) with (P0,P1) in contour order (C division)
                    const int dxfix = (((int)p1.x - (int)p0.x) * 65536) / ((int)p1.y - (int)p0.y);
                    int v = (int)((long long)top.x * 65536 + (long long)(y - top.y) * dxfix);
                    ++k;
                    // insert into the sorted register array (unrolled; a value pushed past the end is dropped, k remembers it)
#pragma unroll
                    for (int j = 0; j < kMaxCross; ++j) {
                        const int lo = min(xr[j], v);
                        v = max(xr[j], v);
                        xr[j] = lo;
                    }
                }
            }
            p0 = p1;
        }
        if (__any_sync(0xFFFFFFFFu, k > kMaxCross)) {
            for (int yy = y0; yy < min(y0 + 32, ymax); ++yy) fill_row_cooperative(planes, chmask, L, nl, R, nr, xs, sorted, lane, yy);
            continue;
        }
#pragma unroll
        for (int pidx = 0; pidx < kMaxCross / 2; ++pidx) {
            if (2 * pidx + 1 >= k) break;
            int xa = (xr[2 * pidx] + 65535) >> 16, xb = xr[2 * pidx + 1] >> 16;   // ceil left, floor right
            if (xa < kW && xb >= 0) {
                xa = max(xa, 0); xb = min(xb, kW - 1);
                for (int w = xa >> 5; w <= (xb >> 5); ++w) {
                    if (max(xa, w << 5) > min(xb, (w << 5) + 31)) continue;
                    const int lo = max(xa, w << 5) & 31, hi = min(xb, (w << 5) + 31) & 31;
                    const uint32_t bits = (0xFFFFFFFFu >> (31 - hi)) & (0xFFFFFFFFu << lo);
                    uint32_t cm = chmask;
                    while (cm) {
                        const int c = __ffs(cm) - 1;
                        cm &= cm - 1;
                        atomicOr(planes + c * kPlaneWords + y * kWordsPerRow + w, bits);
                    }
                }
            }
        }
    }
}

__global__ void __launch_bounds__(kDrawThreads) k_map_draw(MapIn in, uint32_t* __restrict__ bitplanes, int maxv) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    const int prim = blockIdx.x * kDrawWarps + warp;
    const int lane_begin = in.frame_map[4 * b], lane_count = in.frame_map[4 * b + 1];
    const int cw_begin = in.frame_map[4 * b + 2], cw_count = in.frame_map[4 * b + 3];
    if (prim >= lane_count + cw_count) return;
    // per-warp scratch: L[maxv], R[maxv] (short2), xs[2*maxv], sorted[2*maxv] (int)
    unsigned char* base = smem + (size_t)warp * ((size_t)maxv * 2 * sizeof(short2) + (size_t)maxv * 4 * sizeof(int));
    short2* L = reinterpret_cast<short2*>(base);
    short2* R = L + maxv;
    int* xs = reinterpret_cast<int*>(R + maxv);
    int* sorted = xs + 2 * maxv;
    const double* T = in.ego_T + (size_t)b * 9;
    uint32_t* planes = bitplanes + (size_t)b * kFrameWords;

    if (prim < lane_count) {
        const int32_t* ln = in.lanes + 4 * (size_t)(lane_begin + prim);
        const int flags = ln[2];
        const int nl = compact_polyline(in, T, ln[0], L, lane);
        const int nr = compact_polyline(in, T, ln[1], R, lane);
        if (nl > 1 && nr > 1) {                                              // utils.py:152-154 (n >= 4 >= 3)
            const uint32_t cm = 1u | ((flags & 1) ? (1u << 4) : 0u) | ((flags & 2) ? (1u << 5) : 0u);   // utils.py:155-159
            fill_polygon(planes, cm, L, nl, R, nr, xs, sorted, lane);
        }
        const int lm = (flags >> 4) & 15, rm = (flags >> 8) & 15;
        if (nl > 1) {                                                        // utils.py:161-162, :167-169
            const uint32_t cm = (1u << 1) | (lm ? (1u << lm) : 0u);
            for (int i = 0; i + 1 < nl; ++i) draw_segment(planes, cm, L[i], L[i + 1], lane);
        }
        if (nr > 1) {                                                        // utils.py:163-164, :171-173
            const uint32_t cm = (1u << 2) | (rm ? (1u << rm) : 0u);
            for (int i = 0; i + 1 < nr; ++i) draw_segment(planes, cm, R[i], R[i + 1], lane);
        }
    } else {
        const int pl = in.cws[cw_begin + (prim - lane_count)];
        const int n = compact_polyline(in, T, pl, L, lane);
        if (n >= 3) fill_polygon(planes, 1u << 3, L, n, R, 0, xs, sorted, lane);   // utils.py:179-180
    }
}

// bit planes -> float32, 4 pixels per thread (one float4 store); flip + dropout fused
// One short-lived 128-thread CTA per (frame, channel, band of 10 rows) = 28 800 contiguous output bytes: the band's 240
// bit-plane words -> a float32 tile in shared memory -> ONE TMA bulk store (the shape that reaches 7.5 TB/s for the
// LiDAR tiles, lidar.cu); flip (utils.py:399-400) and dropout rectangles (utils.py:496-497) fused.
constexpr int kExpRows = 10;
constexpr int kExpBytes = kExpRows * kW * 4;            // 28 800
constexpr int kExpBands = kH / kExpRows;                // 40
// With `aug` (bevgen_map_raster_augmented) flip / rectangles come from the frame's parameters, and frames that rotate or
// scale are left to the bit-plane gather (augment.cu).
__global__ void __launch_bounds__(128) k_map_expand(const uint32_t* __restrict__ bitplanes, float* __restrict__ out,
                                                    const uint8_t* __restrict__ flip, const int32_t* __restrict__ rects,
                                                    const int32_t* __restrict__ nrects, const bevgen_aug_params* __restrict__ aug) {
    extern __shared__ __align__(128) unsigned char mtile[];
    __shared__ uint32_t words[kExpRows * kWordsPerRow];
    __shared__ int sRect[kMaxRects * 4];
    const int tid = threadIdx.x;
    const unsigned t = blockIdx.x;                      // (b*9 + c)*40 + band
    const unsigned bc = t / kExpBands, band = t - bc * kExpBands, b = bc / kMapC;
    const int y0 = (int)band * kExpRows;
    const uint32_t* src = bitplanes + (size_t)bc * kPlaneWords + y0 * kWordsPerRow;
    for (int i = tid; i < kExpRows * kWordsPerRow; i += 128) words[i] = __ldg(src + i);
    bool fl = flip ? (flip[b] != 0) : false;
    int nr = nrects ? nrects[b] : 0;
    const int32_t* rsrc = rects ? rects + (size_t)b * kMaxRects * 4 : nullptr;
    if (aug) {
        const bevgen_aug_params* P = aug + b;
        if (P->do_rotate | P->do_scale) return;         // whole CTA
        fl = P->do_flip != 0;
        nr = P->n_rects < 0 ? 0 : P->n_rects;
        rsrc = P->rects;
    }
    nr = nr > kMaxRects ? kMaxRects : nr;
    if (tid < 4 * nr) sRect[tid] = rsrc[tid];
    __syncthreads();
    constexpr int kRow4 = kW / 4;                       // 180 float4 per row
    for (int i = tid; i < kExpRows * kRow4; i += 128) {
        const int yl = i / kRow4, x4 = i - yl * kRow4;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int x = 4 * x4 + j;
            const int sx = fl ? kW - 1 - x : x;         // utils.py:399-400
            bool on = (words[yl * kWordsPerRow + (sx >> 5)] >> (sx & 31)) & 1u;
            if (on && nr && in_rects(sRect, nr, y0 + yl, x)) on = false;    // utils.py:496-497
            v[j] = on ? 1.0f : 0.0f;
        }
        reinterpret_cast<float4*>(mtile)[i] = make_float4(v[0], v[1], v[2], v[3]);
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (tid != 0) return;
    bulk_store_s2g(reinterpret_cast<unsigned char*>(out) + (size_t)t * kExpBytes, mtile, kExpBytes);
    bulk_commit();
    bulk_wait_read<0>();
}

struct Mat3 { double m[9]; };
__global__ void __launch_bounds__(256) k_world_to_bev_pixel_v(const double* __restrict__ xy, long long m, Mat3 T,
                                                              long long* __restrict__ out) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= m) return;
    long long px, py;
    map_vertex_px(T.m, xy[2 * i], xy[2 * i + 1], px, py);
    out[2 * i] = px; out[2 * i + 1] = py;
}

}  // namespace
}  // namespace bevgen

using namespace bevgen;

extern "C" size_t bevgen_map_workspace_bytes(int B) {
    return B > 0 ? (size_t)B * kFrameWords * sizeof(uint32_t) : 0;
}

namespace {
// clear the bit planes and draw every frame's primitives into them
int draw_maps(const double* verts, const int32_t* pl_off, const int32_t* lanes, const int32_t* cws, const int32_t* frame_map,
              const double* ego_T, int B, int max_prims, int max_polyline_verts, uint32_t* planes, cudaStream_t st) {
    BEVGEN_CUDA(cudaMemsetAsync(planes, 0, bevgen_map_workspace_bytes(B), st));
    if (max_prims > 0) {
        if (!verts || !pl_off) return BEVGEN_E_INVALID_ARG;
        MapIn in{verts, pl_off, lanes, cws, frame_map, ego_T};
        if (max_polyline_verts < 1 || max_polyline_verts > BEVGEN_MAX_POLYLINE_VERTS) return BEVGEN_E_UNSUPPORTED;
        const int maxv = (max_polyline_verts + 31) / 32 * 32;
        const size_t smem = (size_t)kDrawWarps * ((size_t)maxv * 2 * sizeof(short2) + (size_t)maxv * 4 * sizeof(int));
        if (smem > 48 * 1024) BEVGEN_DYN_SMEM(k_map_draw, smem);
        const dim3 grid((max_prims + kDrawWarps - 1) / kDrawWarps, B);
        BEVGEN_LAUNCH("k_map_draw", st, k_map_draw<<<grid, kDrawThreads, smem, st>>>(in, planes, maxv));
    }
    return 0;
}
}  // namespace

extern "C" int bevgen_map_raster(const double* verts, const int32_t* pl_off, const int32_t* lanes, const int32_t* cws,
                                 const int32_t* frame_map, const double* ego_T, int B, int max_prims, int max_polyline_verts,
                                 const uint8_t* flip,
                                 const int32_t* drop_rects, const int32_t* drop_counts, float* out, void* workspace,
                                 size_t workspace_bytes, bevgen_stream_t stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (B <= 0 || max_prims < 0 || !out || !frame_map || !ego_T) return BEVGEN_E_INVALID_ARG;
    if ((drop_rects == nullptr) != (drop_counts == nullptr)) return BEVGEN_E_INVALID_ARG;
    if (!workspace || workspace_bytes < bevgen_map_workspace_bytes(B)) return BEVGEN_E_WORKSPACE;
    if (B > 65535) return BEVGEN_E_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(out) & 15u) != 0) return BEVGEN_E_INVALID_ARG;      // TMA bulk stores need 16 B alignment
    uint32_t* planes = static_cast<uint32_t*>(workspace);
    const int rc = draw_maps(verts, pl_off, lanes, cws, frame_map, ego_T, B, max_prims, max_polyline_verts, planes, st);
    if (rc) return rc;
    BEVGEN_LAUNCH("k_map_expand", st,
                  k_map_expand<<<(unsigned)(B * kMapC * kExpBands), 128, kExpBytes, st>>>(planes, out, flip, drop_rects, drop_counts, nullptr));
    return 0;
}

extern "C" int bevgen_map_raster_augmented(const double* verts, const int32_t* pl_off, const int32_t* lanes, const int32_t* cws,
                                           const int32_t* frame_map, const double* ego_T, int B, int max_prims,
                                           int max_polyline_verts, const bevgen_aug_params* params,
                                           const bevgen_aug_params* params_host, float* out, void* workspace,
                                           size_t workspace_bytes, bevgen_stream_t stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (B <= 0 || max_prims < 0 || !out || !frame_map || !ego_T || !params || !params_host) return BEVGEN_E_INVALID_ARG;
    if (!workspace || workspace_bytes < bevgen_map_workspace_bytes(B)) return BEVGEN_E_WORKSPACE;
    if (B > 32767) return BEVGEN_E_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(out) & 15u) != 0) return BEVGEN_E_INVALID_ARG;
    uint32_t* planes = static_cast<uint32_t*>(workspace);
    int rc = draw_maps(verts, pl_off, lanes, cws, frame_map, ego_T, B, max_prims, max_polyline_verts, planes, st);
    if (rc) return rc;
    bool any_plain = false, any_resampled = false;
    for (int b = 0; b < B; ++b) {
        const bool r = params_host[b].do_rotate != 0 || params_host[b].do_scale != 0;
        any_resampled |= r; any_plain |= !r;
    }
    if (any_plain)      // frames that neither rotate nor scale: flip / dropout fused into the expand; the others' CTAs exit at once
        BEVGEN_LAUNCH("k_map_expand", st,
                      k_map_expand<<<(unsigned)(B * kMapC * kExpBands), 128, kExpBytes, st>>>(planes, out, nullptr, nullptr, nullptr, params));
    if (any_resampled) rc = launch_augment_bitplanes(planes, out, B, params, params_host, st);
    return rc;
}

extern "C" int bevgen_world_to_bev_pixel(const double* xy, int64_t m, const double* T_host, int64_t* out,
                                         bevgen_stream_t stream) {
    if (m < 0 || !T_host) return BEVGEN_E_INVALID_ARG;
    if (m == 0) return 0;
    if (!xy || !out) return BEVGEN_E_INVALID_ARG;
    Mat3 T;
    for (int i = 0; i < 9; ++i) T.m[i] = T_host[i];
    k_world_to_bev_pixel_v<<<(unsigned)((m + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        xy, m, T, reinterpret_cast<long long*>(out));
    BEVGEN_CHECK_LAUNCH();
    return 0;
}
