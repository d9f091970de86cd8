/*
 * bevgen_oracle.c -- CPU restatement ("oracle") of the reference's BEV input generation.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library, and only as the checker (or the timed CPU baseline), never as the thing
 * shipped.  The product path is the sm_100a library under
 * visiontransformer-intention-prediction_b200/csrc and fails loudly without it.
 *
 * Plain scalar C, one thread, no third-party code.  Every function cites the reference
 * file:line (relative to the reference repo root) whose arithmetic it restates.  The
 * OpenCV / OpenBLAS behaviour that the reference reaches through cv2 / numpy is restated
 * from the published algorithms (opencv-python 4.x drawing.cpp / imgwarp.cpp / resize.cpp,
 * OpenBLAS dgemm k-ordered FMA accumulation) and pinned by tests/test_oracle_vs_reference.py
 * against the real reference functions executed in the build container, and by the
 * committed fixtures under tests/golden/ everywhere else.
 *
 * Build: see oracle/Makefile  (gcc -O2 -ffp-contract=off: no implicit FMA contraction;
 * the explicit fma() calls below are the only fused operations).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- grid constants: reference constants.py:28-47 ------------------------------------ */
#define GRID_H 400            /* GRID_HEIGHT_PX   constants.py:28 */
#define GRID_W 720            /* GRID_WIDTH_PX    constants.py:28 */
#define VOXEL 0.2             /* VOXEL_SIZE_M     constants.py:29 */
#define OFF_X 360.0           /* BEV_PIXEL_OFFSET_X = 720/2.0      constants.py:38 */
#define OFF_Y 300.0           /* BEV_PIXEL_OFFSET_Y = 400*3.0/4.0  constants.py:39 */
#define Z_MIN (-2.0)          /* constants.py:42 */
#define Z_MAX 3.8             /* constants.py:42 */
#define ZBINS 29              /* LIDAR_HEIGHT_CHANNELS constants.py:43 */
#define MAP_CH 9              /* MAP_CHANNELS constants.py:47 */

#define ORC_API __attribute__((visibility("default")))

ORC_API int orc_abi_version(void) { return 1; }

ORC_API void orc_constants(double *out /*10*/) {
    out[0] = GRID_H; out[1] = GRID_W; out[2] = VOXEL; out[3] = OFF_X; out[4] = OFF_Y;
    out[5] = Z_MIN;  out[6] = Z_MAX;  out[7] = ZBINS; out[8] = MAP_CH; out[9] = Z_MAX - Z_MIN;
}

/* numpy's float64 -> int64 astype on x86-64 (cvttsd2si): out-of-range / NaN -> INT64_MIN */
static inline int64_t f64_to_i64(double v) {
    if (!(v >= -9223372036854775808.0 && v < 9223372036854775808.0)) return INT64_MIN;
    return (int64_t)v;
}

/* =======================================================================================
 * Pose transform: utils.py:27-33.  numpy promotes the points to float64 (hstack with
 * float64 ones, utils.py:31) and multiplies by the float64 4x4 through OpenBLAS dgemm,
 * which accumulates each output element as an FMA chain in k order starting from 0.
 * ===================================================================================== */
static inline void xform_row(const double *T, double x, double y, double z, double *o) {
    for (int r = 0; r < 3; ++r) {
        double acc = fma(T[r * 4 + 0], x, 0.0);
        acc = fma(T[r * 4 + 1], y, acc);
        acc = fma(T[r * 4 + 2], z, acc);
        acc = fma(T[r * 4 + 3], 1.0, acc);
        o[r] = acc;
    }
}

ORC_API void orc_transform_points_f64(const double *pts, int64_t n, int64_t stride,
                                      const double *T /*16*/, double *out /*n*3*/) {
    for (int64_t i = 0; i < n; ++i)
        xform_row(T, pts[i * stride], pts[i * stride + 1], pts[i * stride + 2], out + i * 3);
}

ORC_API void orc_transform_points_f32(const float *pts, int64_t n, int64_t stride,
                                      const double *T, double *out) {
    for (int64_t i = 0; i < n; ++i)
        xform_row(T, (double)pts[i * stride], (double)pts[i * stride + 1],
                  (double)pts[i * stride + 2], out + i * 3);
}

/* =======================================================================================
 * LiDAR binning + max scatter: utils.py:62-106.
 *   pixel_x = floor(360.0 + y/0.2)  (utils.py:80)     pixel_y = floor(300.0 - x/0.2) (utils.py:81)
 *   valid   = in-grid and Z_MIN <= z < Z_MAX          (utils.py:83-87)
 *   z_bin   = clip(floor((z - Z_MIN)/(Z_MAX - Z_MIN)*29), 0, 28)   (utils.py:94-95)
 *   bev[sweep*29 + z_bin, pixel_y, pixel_x] = max(., intensity)    (utils.py:97-105)
 * Returns 1 and the flat voxel index when the point lands in the grid, else 0.
 * ===================================================================================== */
static inline int lidar_voxel(double x, double y, double z, int *py_out, int *px_out, int *zb_out) {
    int64_t px = f64_to_i64(floor(OFF_X + y / VOXEL));
    int64_t py = f64_to_i64(floor(OFF_Y - x / VOXEL));
    if (!(px >= 0 && px < GRID_W && py >= 0 && py < GRID_H && z >= Z_MIN && z < Z_MAX)) return 0;
    int64_t zb = f64_to_i64(floor((z - Z_MIN) / (Z_MAX - Z_MIN) * (double)ZBINS));
    if (zb < 0) zb = 0;
    if (zb > ZBINS - 1) zb = ZBINS - 1;
    *py_out = (int)py; *px_out = (int)px; *zb_out = (int)zb;
    return 1;
}

/* np.maximum.at on a +0.0-initialised float32 grid, restricted to the documented intensity
 * domain (finite or +inf, >= +0.0): plain float max. */
static inline void vox_max(float *cell, float v) { if (v > *cell) *cell = v; }

/* One sweep of already ego-frame float64 points (what dataset.py:340-347 hands over). */
ORC_API void orc_lidar_sweep_f64(const double *pts /*n*3*/, const float *intensity, int64_t n,
                                 int sweep_idx, float *bev /*[S*29,400,720]*/) {
    for (int64_t i = 0; i < n; ++i) {
        int py, px, zb;
        if (!lidar_voxel(pts[i * 3], pts[i * 3 + 1], pts[i * 3 + 2], &py, &px, &zb)) continue;
        size_t idx = ((size_t)(sweep_idx * ZBINS + zb) * GRID_H + (size_t)py) * GRID_W + (size_t)px;
        vox_max(bev + idx, intensity[i]);
    }
}

/* Fused frame: raw sweep-frame float32 (x,y,z,intensity) rows + per-sweep rel_tf
 * (dataset.py:339-340 followed by utils.py:62-106).  bev must be zero-initialised
 * (utils.py:69-70) by the caller; sweeps with an empty range leave their 29 channels zero
 * (utils.py:75-76). */
ORC_API void orc_lidar_frame_fused(const float *pts4, const int32_t *sweep_offsets /*S+1*/,
                                   const double *rel_tf /*S*16*/, int S, float *bev) {
    for (int s = 0; s < S; ++s) {
        const double *T = rel_tf + (size_t)s * 16;
        for (int64_t i = sweep_offsets[s]; i < sweep_offsets[s + 1]; ++i) {
            const float *p = pts4 + i * 4;
            double e[3];
            xform_row(T, (double)p[0], (double)p[1], (double)p[2], e);
            int py, px, zb;
            if (!lidar_voxel(e[0], e[1], e[2], &py, &px, &zb)) continue;
            size_t idx = ((size_t)(s * ZBINS + zb) * GRID_H + (size_t)py) * GRID_W + (size_t)px;
            vox_max(bev + idx, p[3]);
        }
    }
}

/* =======================================================================================
 * Map vertex pipeline: utils.py:47-60 (world_to_bev_pixel) with the 3x3 from utils.py:35-45.
 *   ego = T . [x y 1]^T (dgemm FMA chain, k order)       utils.py:52-54
 *   px  = 360.0 + ego_y/0.2 ; py = 300.0 - ego_x/0.2      utils.py:56-57
 *   np.round (half-to-even) then astype(int)              utils.py:60
 * ===================================================================================== */
static inline void map_vertex_px(const double *T /*9*/, double wx, double wy, int64_t *px, int64_t *py) {
    double ex = fma(T[2], 1.0, fma(T[1], wy, fma(T[0], wx, 0.0)));
    double ey = fma(T[5], 1.0, fma(T[4], wy, fma(T[3], wx, 0.0)));
    *px = f64_to_i64(nearbyint(OFF_X + ey / VOXEL));
    *py = f64_to_i64(nearbyint(OFF_Y - ex / VOXEL));
}

ORC_API void orc_world_to_bev_pixel(const double *xy /*m*2*/, int64_t m, const double *T /*9*/,
                                    int64_t *out /*m*2: px,py*/) {
    for (int64_t i = 0; i < m; ++i) map_vertex_px(T, xy[i * 2], xy[i * 2 + 1], out + i * 2, out + i * 2 + 1);
}

/* =======================================================================================
 * cv2.polylines / fillPoly outline: 8-connected Bresenham, thickness 1, LINE_8, shift 0
 * (OpenCV drawing.cpp Line() -> LineIterator(connectivity 8, left_to_right=true)).
 * Reached from utils.py:155-180.  All vertices are inside the image (utils.py:140-144
 * drops the others) so no clipping branch exists here.
 * ===================================================================================== */
static void line8(uint8_t *img, int x1, int y1, int x2, int y2) {
    int dx = x2 - x1, dy = y2 - y1;
    if (dx < 0) { x1 = x2; y1 = y2; dx = -dx; dy = -dy; }   /* iterate left -> right */
    int ystep = 1;
    if (dy < 0) { dy = -dy; ystep = -1; }
    int steep = dy > dx;
    int a = steep ? dy : dx, b = steep ? dx : dy;
    int err = a - 2 * b, plus = 2 * a, minus = -2 * b;
    int x = x1, y = y1;
    for (int k = 0; k <= a; ++k) {
        img[(size_t)y * GRID_W + x] = 1;
        int m = err < 0;
        err += minus + (m ? plus : 0);
        if (steep) { y += ystep; if (m) x += 1; }
        else       { x += 1;     if (m) y += ystep; }
    }
}

ORC_API void orc_polyline(uint8_t *img /*[400,720]*/, const int32_t *v /*n*2: x,y*/, int n) {
    /* cv2.polylines(isClosed=False): consecutive segments, both end points drawn.
     * n == 1 never reaches cv2 in the reference (len > 1 guards, utils.py:161-173). */
    for (int i = 0; i + 1 < n; ++i) line8(img, v[2 * i], v[2 * i + 1], v[2 * i + 2], v[2 * i + 3]);
}

/* cv2.fillPoly, single contour (OpenCV drawing.cpp CollectPolyEdges + FillEdgeCollection):
 * closed outline via line8, then even-odd scanline spans in 16.16 fixed point:
 *   edge x at its top row = top.x << 16 ; dx = trunc(((x1 - x0) << 16) / (y1 - y0))
 *   edge active on rows [y_top, y_bottom)
 *   per row: sort crossings, pair them, fill [ceil(xa), floor(xb)] clipped to the image. */
typedef struct { int y0, y1; int64_t x, dx; } orc_edge;

static int cmp_i64(const void *a, const void *b) {
    int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
    return (x > y) - (x < y);
}

ORC_API void orc_fill_poly(uint8_t *img, const int32_t *v /*n*2*/, int n) {
    if (n <= 0) return;
    orc_edge *edges = (orc_edge *)malloc(sizeof(orc_edge) * (size_t)n);
    int64_t *xs = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    int ne = 0, ymin = INT32_MAX, ymax = INT32_MIN;
    int px0 = v[2 * (n - 1)], py0 = v[2 * (n - 1) + 1];
    for (int i = 0; i < n; ++i) {
        int px1 = v[2 * i], py1 = v[2 * i + 1];
        line8(img, px0, py0, px1, py1);
        if (py0 != py1) {
            orc_edge e;
            /* C integer division truncates toward zero, as cv2's int64 division does */
            e.dx = (((int64_t)px1 - (int64_t)px0) * 65536) / ((int64_t)py1 - (int64_t)py0);
            if (py0 < py1) { e.y0 = py0; e.y1 = py1; e.x = (int64_t)px0 * 65536; }
            else           { e.y0 = py1; e.y1 = py0; e.x = (int64_t)px1 * 65536; }
            if (e.y0 < ymin) ymin = e.y0;
            if (e.y1 > ymax) ymax = e.y1;
            edges[ne++] = e;
        }
        px0 = px1; py0 = py1;
    }
    if (ne >= 2) {
        if (ymax > GRID_H) ymax = GRID_H;
        for (int y = ymin; y < ymax; ++y) {
            int k = 0;
            for (int e = 0; e < ne; ++e)
                if (edges[e].y0 <= y && y < edges[e].y1)
                    xs[k++] = edges[e].x + (int64_t)(y - edges[e].y0) * edges[e].dx;
            qsort(xs, (size_t)k, sizeof(int64_t), cmp_i64);
            for (int j = 0; j + 1 < k; j += 2) {
                int64_t xa = (xs[j] + 65535) >> 16, xb = xs[j + 1] >> 16;
                if (xa < GRID_W && xb >= 0) {
                    if (xa < 0) xa = 0;
                    if (xb >= GRID_W) xb = GRID_W - 1;
                    for (int64_t x = xa; x <= xb; ++x) img[(size_t)y * GRID_W + (size_t)x] = 1;
                }
            }
        }
    }
    free(edges); free(xs);
}

/* =======================================================================================
 * Map rasterization: utils.py:108-182 on a pre-parsed ("flat") map.
 *   verts      world xy float64 [V,2] (only entries that were dicts with 'x','y': utils.py:134)
 *   pl_off     polyline CSR offsets [P+1] into verts
 *   lanes      [L,4] int32: left polyline id, right polyline id, flags, reserved
 *              flags bit0 is_intersection (utils.py:156), bit1 lane_type=="BUS" (utils.py:158),
 *              bits 4-7 left mark channel (0 none | 6 | 7 | 8), bits 8-11 right mark channel
 *              (utils.py:166-173); polyline id -1 = boundary key absent / empty
 *   cws        [CW] polyline ids of crosswalk 'polygon' lists (utils.py:175-180)
 *   out        uint8 [9,400,720], must be zero on entry (utils.py:146)
 * ===================================================================================== */
static int compact_px(const double *verts, const int32_t *pl_off, int pl, const double *T, int32_t *dst) {
    if (pl < 0) return 0;
    int n = 0;
    for (int i = pl_off[pl]; i < pl_off[pl + 1]; ++i) {
        int64_t px, py;
        map_vertex_px(T, verts[2 * (size_t)i], verts[2 * (size_t)i + 1], &px, &py);
        /* utils.py:140-144: vertices outside the grid are dropped, order preserved */
        if (px >= 0 && px < GRID_W && py >= 0 && py < GRID_H) { dst[2 * n] = (int32_t)px; dst[2 * n + 1] = (int32_t)py; ++n; }
    }
    return n;
}

ORC_API void orc_map_raster(const double *verts, const int32_t *pl_off, int P,
                            const int32_t *lanes, int L, const int32_t *cws, int CW,
                            const double *T /*9*/, uint8_t *out) {
    int maxv = 1;
    for (int p = 0; p < P; ++p) { int c = pl_off[p + 1] - pl_off[p]; if (c > maxv) maxv = c; }
    int32_t *lp = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)maxv);
    int32_t *rp = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)maxv);
    int32_t *poly = (int32_t *)malloc(sizeof(int32_t) * 4 * (size_t)maxv);
    const size_t plane = (size_t)GRID_H * GRID_W;
    for (int l = 0; l < L; ++l) {
        const int32_t *ln = lanes + 4 * (size_t)l;
        int flags = ln[2];
        int nl = compact_px(verts, pl_off, ln[0], T, lp);
        int nr = compact_px(verts, pl_off, ln[1], T, rp);
        if (nl > 1 && nr > 1) {                                   /* utils.py:152 */
            int n = 0;                                            /* left + flipud(right), utils.py:153 */
            for (int i = 0; i < nl; ++i) { poly[2 * n] = lp[2 * i]; poly[2 * n + 1] = lp[2 * i + 1]; ++n; }
            for (int i = nr - 1; i >= 0; --i) { poly[2 * n] = rp[2 * i]; poly[2 * n + 1] = rp[2 * i + 1]; ++n; }
            if (n >= 3) {                                         /* utils.py:154 */
                orc_fill_poly(out + 0 * plane, poly, n);          /* utils.py:155 */
                if (flags & 1) orc_fill_poly(out + 4 * plane, poly, n);   /* utils.py:156-157 */
                if (flags & 2) orc_fill_poly(out + 5 * plane, poly, n);   /* utils.py:158-159 */
            }
        }
        if (nl > 1) orc_polyline(out + 1 * plane, lp, nl);        /* utils.py:161-162 */
        if (nr > 1) orc_polyline(out + 2 * plane, rp, nr);        /* utils.py:163-164 */
        int lm = (flags >> 4) & 15, rm = (flags >> 8) & 15;
        if (lm && nl > 1) orc_polyline(out + (size_t)lm * plane, lp, nl);   /* utils.py:167-169 */
        if (rm && nr > 1) orc_polyline(out + (size_t)rm * plane, rp, nr);   /* utils.py:171-173 */
    }
    for (int c = 0; c < CW; ++c) {
        int n = compact_px(verts, pl_off, cws[c], T, lp);
        if (n >= 3) orc_fill_poly(out + 3 * plane, lp, n);        /* utils.py:179-180 */
    }
    free(lp); free(rp); free(poly);
}

ORC_API void orc_u8_to_f32(const uint8_t *src, float *dst, int64_t n) {   /* utils.py:182 */
    for (int64_t i = 0; i < n; ++i) dst[i] = (float)src[i];
}

/* =======================================================================================
 * Augmentation: utils.py:394-498 (image part only; the GT-box arithmetic stays on the host).
 * ===================================================================================== */

/* utils.py:399-400: np.flip(axis=2) -- reverse W */
ORC_API void orc_flip_w(const float *src, float *dst, int C) {
    for (size_t r = 0; r < (size_t)C * GRID_H; ++r)
        for (int x = 0; x < GRID_W; ++x) dst[r * GRID_W + x] = src[r * GRID_W + (GRID_W - 1 - x)];
}

/* cv2.getRotationMatrix2D((360.0, 200.0), angle_deg, 1.0)  -- utils.py:425-428 */
ORC_API void orc_rotation_matrix(double angle_deg, double *M /*6*/) {
    double cx = GRID_W / 2.0, cy = GRID_H / 2.0;
    double a = angle_deg * (3.1415926535897932384626433832795 / 180.0);
    double alpha = cos(a) * 1.0, beta = sin(a) * 1.0;
    M[0] = alpha; M[1] = beta;  M[2] = (1 - alpha) * cx - beta * cy;
    M[3] = -beta; M[4] = alpha; M[5] = beta * cx + (1 - alpha) * cy;
}

/* cv2.warpAffine's inversion of the forward matrix (imgwarp.cpp, no WARP_INVERSE_MAP) */
ORC_API void orc_invert_affine(const double *Min, double *M /*6*/) {
    memcpy(M, Min, 6 * sizeof(double));
    double D = M[0] * M[4] - M[1] * M[3];
    D = D != 0 ? 1. / D : 0;
    double A11 = M[4] * D, A22 = M[0] * D;
    M[0] = A11; M[1] *= -D; M[3] *= -D; M[4] = A22;
    double b1 = -M[0] * M[2] - M[1] * M[5];
    double b2 = -M[3] * M[2] - M[4] * M[5];
    M[2] = b1; M[5] = b2;
}

static inline int sat_round_i32(double v) {          /* cv::saturate_cast<int>(double) = cvRound */
    double r = nearbyint(v);
    if (r >= 2147483647.0) return INT32_MAX;
    if (r <= -2147483648.0) return INT32_MIN;
    return (int)r;
}
static inline int sat_i16(int v) { return v < -32768 ? -32768 : (v > 32767 ? 32767 : v); }

/* cv2.warpAffine(ch, M, (720,400), INTER_LINEAR, BORDER_CONSTANT, 0) per channel -- utils.py:430-439.
 * Minv is the already inverted 2x3 (dst -> src).  Coordinates: 10-bit fixed point, rounded
 * to a 1/32-pixel grid; bilinear weights and the 4-tap sum in float32, left to right. */
ORC_API void orc_warp_affine(const float *src, float *dst, int C, const double *Minv /*6*/) {
    static int adelta[GRID_W], bdelta[GRID_W];
    for (int x = 0; x < GRID_W; ++x) {
        adelta[x] = sat_round_i32(Minv[0] * x * 1024.0);
        bdelta[x] = sat_round_i32(Minv[3] * x * 1024.0);
    }
    const size_t plane = (size_t)GRID_H * GRID_W;
    for (int y = 0; y < GRID_H; ++y) {
        int X0 = sat_round_i32((Minv[1] * y + Minv[2]) * 1024.0) + 16;
        int Y0 = sat_round_i32((Minv[4] * y + Minv[5]) * 1024.0) + 16;
        for (int x = 0; x < GRID_W; ++x) {
            int X = (X0 + adelta[x]) >> 5, Y = (Y0 + bdelta[x]) >> 5;
            int sx = sat_i16(X >> 5), sy = sat_i16(Y >> 5);
            float fx = (float)(X & 31) * (1.f / 32), fy = (float)(Y & 31) * (1.f / 32);
            float tx0 = 1.f - fx, tx1 = fx, ty0 = 1.f - fy, ty1 = fy;
            float w0 = ty0 * tx0, w1 = ty0 * tx1, w2 = ty1 * tx0, w3 = ty1 * tx1;
            int in00 = (sx >= 0 && sx < GRID_W && sy >= 0 && sy < GRID_H);
            int in01 = (sx + 1 >= 0 && sx + 1 < GRID_W && sy >= 0 && sy < GRID_H);
            int in10 = (sx >= 0 && sx < GRID_W && sy + 1 >= 0 && sy + 1 < GRID_H);
            int in11 = (sx + 1 >= 0 && sx + 1 < GRID_W && sy + 1 >= 0 && sy + 1 < GRID_H);
            for (int c = 0; c < C; ++c) {
                const float *S = src + (size_t)c * plane;
                float v0 = in00 ? S[(size_t)sy * GRID_W + sx] : 0.f;
                float v1 = in01 ? S[(size_t)sy * GRID_W + sx + 1] : 0.f;
                float v2 = in10 ? S[(size_t)(sy + 1) * GRID_W + sx] : 0.f;
                float v3 = in11 ? S[(size_t)(sy + 1) * GRID_W + sx + 1] : 0.f;
                float r = v0 * w0;
                r = r + v1 * w1;
                r = r + v2 * w2;
                r = r + v3 * w3;
                dst[(size_t)c * plane + (size_t)y * GRID_W + x] = r;
            }
        }
    }
}

/* cv2.resize(ch, (new_w,new_h), INTER_LINEAR) + centre crop / zero pad -- utils.py:455-476.
 * Separable bilinear (horizontal then vertical) in float32, no FMA; coordinates
 * (float)((i+0.5)*scale-0.5) with scale = 1/(dst/src) in float64 (resize.cpp, non-IPP path). */
typedef struct { int s0, s1; float w0, w1; } orc_tap;

static void resize_taps(int dsize, int ssize, orc_tap *t) {
    double inv_scale = (double)dsize / (double)ssize;
    double scale = 1. / inv_scale;
    for (int d = 0; d < dsize; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= (float)s;
        if (s < 0) { f = 0; s = 0; }
        if (s >= ssize - 1) { f = 0; s = ssize - 1; }
        t[d].s0 = s;
        t[d].s1 = s + 1 < ssize ? s + 1 : ssize - 1;
        t[d].w0 = 1.f - f;
        t[d].w1 = f;
    }
}

ORC_API void orc_scale_dims(double s, int *new_h, int *new_w) {      /* utils.py:457: int() truncation */
    *new_h = (int)(GRID_H * s);
    *new_w = (int)(GRID_W * s);
}

ORC_API void orc_scale(const float *src, float *dst, int C, double s) {
    int nh, nw;
    orc_scale_dims(s, &nh, &nw);
    orc_tap *tx = (orc_tap *)malloc(sizeof(orc_tap) * (size_t)nw);
    orc_tap *ty = (orc_tap *)malloc(sizeof(orc_tap) * (size_t)nh);
    resize_taps(nw, GRID_W, tx);
    resize_taps(nh, GRID_H, ty);
    const size_t plane = (size_t)GRID_H * GRID_W;
    int crop = s > 1.0;                                            /* utils.py:464 */
    int h0 = crop ? (nh - GRID_H) / 2 : (GRID_H - nh) / 2;         /* utils.py:465 / :469 */
    int w0 = crop ? (nw - GRID_W) / 2 : (GRID_W - nw) / 2;         /* utils.py:466 / :470 */
    for (int c = 0; c < C; ++c) {
        const float *S = src + (size_t)c * plane;
        float *D = dst + (size_t)c * plane;
        for (int y = 0; y < GRID_H; ++y) {
            for (int x = 0; x < GRID_W; ++x) {
                int ry = crop ? y + h0 : y - h0, rx = crop ? x + w0 : x - w0;
                float r = 0.f;
                if (ry >= 0 && ry < nh && rx >= 0 && rx < nw) {
                    const orc_tap *a = tx + rx, *b = ty + ry;
                    float top = S[(size_t)b->s0 * GRID_W + a->s0] * a->w0 + S[(size_t)b->s0 * GRID_W + a->s1] * a->w1;
                    float bot = S[(size_t)b->s1 * GRID_W + a->s0] * a->w0 + S[(size_t)b->s1 * GRID_W + a->s1] * a->w1;
                    r = top * b->w0 + bot * b->w1;
                }
                D[(size_t)y * GRID_W + x] = r;
            }
        }
    }
    free(tx); free(ty);
}

/* utils.py:496-497: zero rectangles [y0:y0+ph, x0:x0+pw] in every channel (numpy slicing clips) */
ORC_API void orc_dropout(float *grid, int C, const int32_t *rects /*n*4: y0,x0,ph,pw*/, int n) {
    const size_t plane = (size_t)GRID_H * GRID_W;
    for (int r = 0; r < n; ++r) {
        int y0 = rects[4 * r], x0 = rects[4 * r + 1], y1 = y0 + rects[4 * r + 2], x1 = x0 + rects[4 * r + 3];
        if (y1 > GRID_H) y1 = GRID_H;
        if (x1 > GRID_W) x1 = GRID_W;
        for (int c = 0; c < C; ++c)
            for (int y = y0; y < y1; ++y)
                for (int x = x0; x < x1; ++x) grid[(size_t)c * plane + (size_t)y * GRID_W + x] = 0.f;
    }
}
