/*
 * oracle/step_oracle.c -- CPU restatement of the t-pose optimisation step.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under t-pose_b200/ (the product) may
 * include, link or call this file.  It is used by tests/, by
 * __graft_entry__.smoke() and by the cpu_baseline / --impl reference legs of
 * bench.py, always as the checker or as the timed CPU baseline.
 *
 * What it restates (paths relative to /root/reference):
 *   software/triangulate/shader/triangle.vs:45-98   vertex stage (variant
 *        selection, dp cooling, RATIO divide, per-mode resets)
 *   software/triangulate/shader/triangle.fs:25-44   mode 0 / mode 1 fragments
 *   software/triangulate/shader/gradient.cs:19-36   gradient scatter
 *   software/triangulate/shader/shift.cs:16-47      vertex update
 *   software/warp/shader/triangle.vs:63-65          warp dp cooling
 *   software/warp/shader/triangle.fs:38-53          warp mode 0 / mode 1
 *   software/warp/shader/shift.cs:45                warp step size
 * It does the work the way the GL pipeline does it: 13*NT instanced
 * triangles, every covered pixel visited once per instance and per pass,
 * float32 per-fragment arithmetic, int32 accumulators that wrap.  No moment
 * shortcut, no prefix tables -- that is what makes it an independent check of
 * the CUDA path, which uses both.
 *
 * PARITY STATUS: "parity unpinned" for the raster stage.  The reference's own
 * implementation is GLSL on the OpenGL fixed-function rasteriser and cannot
 * run here (no GL/EGL/TinyEngine); the reference holds no golden vectors for
 * this path.  The conventions the GL driver would
 * decide are therefore fixed HERE, once, and the CUDA path must agree with
 * this file bit for bit:
 *   - raster grid W x H, image row 0 = top of the window, pixel centres at
 *     (i+0.5, r+0.5);
 *   - window transform in float32, one rounding per operation, no FMA:
 *        xn = (p.x + Dx) / RATIO ; yn = p.y + Dy
 *        xw = xn*(0.5f*W) + 0.5f*W ; yw = yn*(0.5f*H) + 0.5f*H   (y up)
 *   - snap to 1/256 pixel: X = lrintf(clamp(xw, +-2^15) * 256), same for y,
 *     then flip to y-down: Y = 256*H - Yup;
 *   - coverage: exact int64 edge functions, top-left fill rule in the y-down
 *     frame, both windings rasterise (glDisable(GL_CULL_FACE),
 *     software/triangulate/main.cpp:56), zero-area triangles cover nothing,
 *     fragments only inside [0,W) x [0,H);
 *   - texel fetch at the native raster: 255.0f * (b / 255.0f), which is b.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define TPO_TRIANGULATE 0
#define TPO_WARP 1

typedef struct { int32_t x[3], y[3]; } tpo_tri_fx;

/* Test hook: the reference's own test programs ship shader copies with other cooling laws
 * (tests/opt_min_energy_edgeflip/shader/triangle.vs:59-76).  which < 0: follow the flavour. */
static int g_dp_law = -1;
void tpo_set_dp_law(int which) { g_dp_law = which; }

/* triangle.vs:60-62 (triangulate), warp/shader/triangle.vs:63-65 (warp) */
float tpo_dp(int NT, int flavour)
{
    float dp = 0.05f;
    if (g_dp_law == 2) return dp;            /* no cooling at all: the oldest traces of the reference's test program */
    if (g_dp_law >= 0) flavour = g_dp_law;
    if (flavour == TPO_WARP) dp = dp / (1.0f + 9.0f * (float)NT / 1000.0f);
    else                     dp = dp / (1.0f + 4.0f * (float)NT / 3000.0f);
    return dp;
}

static int32_t snap256(float w)
{
    const float lim = 32768.0f;
    w = fminf(fmaxf(w, -lim), lim);
    return (int32_t)lrintf(w * 256.0f);
}

/* triangle.vs:47-84: instance g -> three snapped window vertices (y down) */
static void variant_vertices(const float *pts, const int32_t *tri, int NT, int g,
                             float ratio, int W, int H, float dp, tpo_tri_fx *out)
{
    const int k = g / NT, t = g % NT;
    const float hw = 0.5f * (float)W, hh = 0.5f * (float)H;
    for (int j = 0; j < 3; j++) {
        const int v = tri[4 * t + j];
        float dx = 0.0f, dy = 0.0f;
        if (k > 0 && (k - 1) / 4 == j) {
            switch ((k - 1) % 4) {
            case 0: dx = dp; break;
            case 1: dx = -dp; break;
            case 2: dy = dp; break;
            default: dy = -dp; break;
            }
        }
        float tx = pts[2 * v] + dx;
        float ty = pts[2 * v + 1] + dy;
        tx = tx / ratio;
        float xw = tx * hw; xw = xw + hw;
        float yw = ty * hh; yw = yw + hh;
        out->x[j] = snap256(xw);
        out->y[j] = 256 * H - snap256(yw);
    }
}

void tpo_variant_vertices(const float *pts, const int32_t *tri, int NT, float ratio,
                          int W, int H, int flavour, int32_t *out6 /* 13*NT*6 */)
{
    const float dp = tpo_dp(NT, flavour);
    for (int g = 0; g < 13 * NT; g++) {
        tpo_tri_fx f;
        variant_vertices(pts, tri, NT, g, ratio, W, H, dp, &f);
        for (int j = 0; j < 3; j++) { out6[6 * g + 2 * j] = f.x[j]; out6[6 * g + 2 * j + 1] = f.y[j]; }
    }
}

/* Coverage iterator: calls the per-fragment body for every covered pixel. */
typedef struct {
    int64_t A[3], B[3], C[3]; /* E_i(px,py) = A*px + B*py + C  (in 1/256 px units) */
    int bias[3];              /* 0 if the edge owns its boundary (top/left), 1 otherwise */
    int x0, x1, y0, y1;       /* clipped pixel bbox, inclusive */
    int empty;
} tpo_setup;

static void tri_setup(const tpo_tri_fx *f, int W, int H, tpo_setup *s)
{
    int64_t X[3] = { f->x[0], f->x[1], f->x[2] }, Y[3] = { f->y[0], f->y[1], f->y[2] };
    int64_t area2 = (X[1] - X[0]) * (Y[2] - Y[0]) - (X[2] - X[0]) * (Y[1] - Y[0]);
    s->empty = 0;
    if (area2 == 0) { s->empty = 1; return; }
    if (area2 < 0) { int64_t q = X[1]; X[1] = X[2]; X[2] = q; q = Y[1]; Y[1] = Y[2]; Y[2] = q; }
    for (int i = 0; i < 3; i++) {
        const int a = i, b = (i + 1) % 3;
        const int64_t dx = X[b] - X[a], dy = Y[b] - Y[a];
        /* E(p) = dx*(py-ya) - dy*(px-xa) ; interior is E > 0 */
        s->A[i] = -dy; s->B[i] = dx; s->C[i] = -dx * Y[a] + dy * X[a];
        const int owns = (dy < 0) || (dy == 0 && dx > 0); /* left edge or top edge */
        s->bias[i] = owns ? 0 : 1;
    }
    int64_t xmin = X[0], xmax = X[0], ymin = Y[0], ymax = Y[0];
    for (int i = 1; i < 3; i++) {
        if (X[i] < xmin) xmin = X[i]; if (X[i] > xmax) xmax = X[i];
        if (Y[i] < ymin) ymin = Y[i]; if (Y[i] > ymax) ymax = Y[i];
    }
    /* pixel i has centre 256*i+128 */
    int64_t px0 = (xmin - 128) >> 8, px1 = (xmax - 128 + 255) >> 8;
    int64_t py0 = (ymin - 128) >> 8, py1 = (ymax - 128 + 255) >> 8;
    if (px0 < 0) px0 = 0; if (py0 < 0) py0 = 0;
    if (px1 > W - 1) px1 = W - 1; if (py1 > H - 1) py1 = H - 1;
    if (px0 > px1 || py0 > py1) { s->empty = 1; return; }
    s->x0 = (int)px0; s->x1 = (int)px1; s->y0 = (int)py0; s->y1 = (int)py1;
}

#define TPO_COVERED_BEGIN(s, i, r)                                                  \
    if (!(s).empty)                                                                 \
        for (int r = (s).y0; r <= (s).y1; r++) {                                    \
            const int64_t cy_ = 256 * (int64_t)r + 128;                             \
            const int64_t cx_ = 256 * (int64_t)(s).x0 + 128;                        \
            int64_t e0_ = (s).A[0] * cx_ + (s).B[0] * cy_ + (s).C[0];               \
            int64_t e1_ = (s).A[1] * cx_ + (s).B[1] * cy_ + (s).C[1];               \
            int64_t e2_ = (s).A[2] * cx_ + (s).B[2] * cy_ + (s).C[2];               \
            for (int i = (s).x0; i <= (s).x1; i++) {                                \
                if (e0_ >= (s).bias[0] && e1_ >= (s).bias[1] && e2_ >= (s).bias[2]) {
#define TPO_COVERED_END(s)                                                          \
                }                                                                   \
                e0_ += 256 * (s).A[0]; e1_ += 256 * (s).A[1]; e2_ += 256 * (s).A[2]; \
            }                                                                       \
        }

/*
 * Mode 0 (triangle.fs:27-35; warp triangle.fs:40-44).  The vertex stage zeroes
 * cn[g] (and ca[g] in the triangulate flavour) first (triangle.vs:88-91).
 * ca is ivec4 per variant: 13*NT*4 int32, .w untouched by fragments.
 */
void tpo_mode0(const uint8_t *rgba, int W, int H, const float *pts, const int32_t *tri,
               int NT, float ratio, int flavour, int32_t *cn, int32_t *ca)
{
    const float dp = tpo_dp(NT, flavour);
#pragma omp parallel for schedule(dynamic, 16)
    for (int g = 0; g < 13 * NT; g++) {
        tpo_tri_fx f; tpo_setup s;
        variant_vertices(pts, tri, NT, g, ratio, W, H, dp, &f);
        tri_setup(&f, W, H, &s);
        uint32_t n = 0, cr = 0, cg = 0, cb = 0;
        if (flavour == TPO_WARP) {
            TPO_COVERED_BEGIN(s, i, r) n += 1u; TPO_COVERED_END(s)
            cn[g] = (int32_t)n;
        } else {
            TPO_COVERED_BEGIN(s, i, r)
                const uint8_t *px = rgba + 4 * ((size_t)r * W + i);
                const float fr = (float)px[0] / 255.0f, fg = (float)px[1] / 255.0f, fb = (float)px[2] / 255.0f;
                n += 1u;
                cr += (uint32_t)(int32_t)(255.0f * fr);
                cg += (uint32_t)(int32_t)(255.0f * fg);
                cb += (uint32_t)(int32_t)(255.0f * fb);
            TPO_COVERED_END(s)
            cn[g] = (int32_t)n;
            ca[4 * g + 0] = (int32_t)cr; ca[4 * g + 1] = (int32_t)cg; ca[4 * g + 2] = (int32_t)cb; ca[4 * g + 3] = 0;
        }
    }
}

/*
 * Mode 1 (triangle.fs:37-43; warp triangle.fs:46-53).  ten[g] is zeroed by the
 * vertex stage, gr[] is zeroed for every vertex a triangle references
 * (triangle.vs:93-98).  The mean colour is ca/cn (integer division) in the
 * triangulate flavour and the uploaded ca itself in the warp flavour.
 */
void tpo_mode1(const uint8_t *rgba, int W, int H, const float *pts, const int32_t *tri,
               int NT, float ratio, int flavour, const int32_t *cn, const int32_t *ca,
               int32_t *ten, int32_t *gr)
{
    const float dp = tpo_dp(NT, flavour);
    for (int t = 0; t < NT; t++)
        for (int j = 0; j < 3; j++) { gr[2 * tri[4 * t + j]] = 0; gr[2 * tri[4 * t + j] + 1] = 0; }
#pragma omp parallel for schedule(dynamic, 16)
    for (int g = 0; g < 13 * NT; g++) {
        tpo_tri_fx f; tpo_setup s;
        variant_vertices(pts, tri, NT, g, ratio, W, H, dp, &f);
        tri_setup(&f, W, H, &s);
        float m0, m1, m2; int have = 1;
        if (flavour == TPO_WARP) {
            m0 = (float)ca[4 * g]; m1 = (float)ca[4 * g + 1]; m2 = (float)ca[4 * g + 2];
        } else if (cn[g] > 0) {
            m0 = (float)(ca[4 * g] / cn[g]); m1 = (float)(ca[4 * g + 1] / cn[g]); m2 = (float)(ca[4 * g + 2] / cn[g]);
        } else { have = 0; m0 = m1 = m2 = 0.0f; }
        uint32_t e = 0;
        TPO_COVERED_BEGIN(s, i, r)
            float d0 = 0.0f, d1 = 0.0f, d2 = 0.0f;
            if (have) {
                const uint8_t *px = rgba + 4 * ((size_t)r * W + i);
                d0 = 255.0f * ((float)px[0] / 255.0f) - m0;
                d1 = 255.0f * ((float)px[1] / 255.0f) - m1;
                d2 = 255.0f * ((float)px[2] / 255.0f) - m2;
            }
            float dd = d0 * d0; dd = dd + d1 * d1; dd = dd + d2 * d2;
            e += (uint32_t)(int32_t)(0.5f * dd);
        TPO_COVERED_END(s)
        ten[g] = (int32_t)e;
    }
}

/* gradient.cs:27-34 (identical arithmetic in the warp copy, :27-49) */
void tpo_gradient(const int32_t *tri, int NT, const int32_t *ten, int32_t *gr)
{
    for (int t = 0; t < NT; t++)
        for (int j = 0; j < 3; j++) {
            const int v = tri[4 * t + j];
            const uint32_t gx = (uint32_t)ten[(1 + 4 * j) * NT + t] - (uint32_t)ten[(2 + 4 * j) * NT + t];
            const uint32_t gy = (uint32_t)ten[(3 + 4 * j) * NT + t] - (uint32_t)ten[(4 + 4 * j) * NT + t];
            gr[2 * v] = (int32_t)((uint32_t)gr[2 * v] + gx);
            gr[2 * v + 1] = (int32_t)((uint32_t)gr[2 * v + 1] + gy);
        }
}

/* shift.cs:16-47; step 5e-5 (triangulate) / 3e-5 (warp, warp/shader/shift.cs:45) */
void tpo_shift(float *pts, int NP, const int32_t *gr, float ratio, int flavour)
{
    const float step = (flavour == TPO_WARP) ? 0.00003f : 0.00005f;
    for (int v = 4; v < NP; v++) {
        float gx = (float)gr[2 * v], gy = (float)gr[2 * v + 1];
        float x = pts[2 * v], y = pts[2 * v + 1];
        if (x <= -ratio) { x = -ratio; gx = 0.0f; }
        else if (x >= ratio) { x = ratio; gx = 0.0f; }
        if (y <= -1.0f) { y = -1.0f; gy = 0.0f; }
        else if (y >= 1.0f) { y = 1.0f; gy = 0.0f; }
        float sx = step * gx; sx = sx / 256.0f; sx = sx / 256.0f;
        float sy = step * gy; sy = sy / 256.0f; sy = sy / 256.0f;
        pts[2 * v] = x - sx;
        pts[2 * v + 1] = y - sy;
    }
}

/*
 * Per-pixel triangle-ID raster of the NT base triangles (instances 0..NT-1 of
 * the draw): the id a pixel ends up with is the LAST instance that covers it
 * (draw order, no depth rejection between coplanar instances), -1 if none.
 */
void tpo_raster_ids(const float *pts, const int32_t *tri, int NT, float ratio,
                    int W, int H, int32_t *ids)
{
    for (size_t q = 0; q < (size_t)W * H; q++) ids[q] = -1;
    for (int t = 0; t < NT; t++) {
        tpo_tri_fx f; tpo_setup s;
        variant_vertices(pts, tri, NT, t, ratio, W, H, 0.0f, &f);
        tri_setup(&f, W, H, &s);
        TPO_COVERED_BEGIN(s, i, r) ids[(size_t)r * W + i] = t; TPO_COVERED_END(s)
    }
}

/*
 * One frame of software/triangulate (main.cpp:196-197 then the pipeline's
 * computecolors, :180-184): energy at the current positions with the colours
 * accumulated at those same positions, gradient, shift, then colours at the
 * new positions.  cn/ca must hold mode-0 results for `pts` on entry.
 * On return: ten = energies at the OLD positions, pts = new positions,
 * cn_next/ca_next = mode 0 at the new positions (cn/ca are left untouched so
 * the caller can see what the reference reads back at main.cpp:201-204).
 */
void tpo_frame_triangulate(const uint8_t *rgba, int W, int H, float *pts, int NP,
                           const int32_t *tri, int NT, float ratio,
                           const int32_t *cn, const int32_t *ca,
                           int32_t *ten, int32_t *gr, int32_t *cn_next, int32_t *ca_next)
{
    tpo_mode1(rgba, W, H, pts, tri, NT, ratio, TPO_TRIANGULATE, cn, ca, ten, gr);
    tpo_gradient(tri, NT, ten, gr);
    tpo_shift(pts, NP, gr, ratio, TPO_TRIANGULATE);
    tpo_mode0(rgba, W, H, pts, tri, NT, ratio, TPO_TRIANGULATE, cn_next, ca_next);
}

/*
 * One frame of software/warp (main.cpp:220-222): count, energy against the
 * OTHER view's image with the uploaded colours, gradient, shift.
 * ca holds the 13x replicated colours (triangulation.hpp:633-641).
 */
void tpo_frame_warp(const uint8_t *rgba_other, int W, int H, float *pts, int NP,
                    const int32_t *tri, int NT, float ratio, const int32_t *ca,
                    int32_t *cn, int32_t *ten, int32_t *gr)
{
    tpo_mode0(rgba_other, W, H, pts, tri, NT, ratio, TPO_WARP, cn, (int32_t *)0);
    tpo_mode1(rgba_other, W, H, pts, tri, NT, ratio, TPO_WARP, cn, ca, ten, gr);
    tpo_gradient(tri, NT, ten, gr);
    tpo_shift(pts, NP, gr, ratio, TPO_WARP);
}

/*
 * The 1/1.5 window (Tiny::window(w/1.5, h/1.5), software/triangulate/main.cpp:53): the raster has
 * rw x rh pixels and every fragment fetches the source image at its pixel centre through a
 * bilinear filter.  The GL driver owns that filter's weight precision, so it is fixed HERE:
 * float32, one rounding per operation, clamp-to-edge, and the filtered value is truncated to a
 * byte the way mode 0 truncates int(255*c) (triangle.fs:31-33).  STATED DIVERGENCE: the
 * reference's mode 1 uses the un-truncated filtered value (triangle.fs:40-42); here both modes see
 * the truncated texel, so that the energy stays an exact integer function of the raster image.
 * The CUDA path (k_resample, csrc/table.cuh) evaluates the same expression tree.
 */
void tpo_resample(const uint8_t *src, int w, int h, uint8_t *dst, int rw, int rh)
{
#pragma omp parallel for schedule(static)
    for (int j = 0; j < rh; j++)
        for (int i = 0; i < rw; i++) {
            float fx = (float)i + 0.5f; fx = fx * (float)w; fx = fx / (float)rw; fx = fx - 0.5f;
            float fy = (float)j + 0.5f; fy = fy * (float)h; fy = fy / (float)rh; fy = fy - 0.5f;
            const float flx = floorf(fx), fly = floorf(fy);
            const float ax = fx - flx, ay = fy - fly;
            const float bx = 1.0f - ax, by = 1.0f - ay;
            int x0 = (int)flx, y0 = (int)fly;
            int x1 = x0 + 1 < w - 1 ? x0 + 1 : w - 1, y1 = y0 + 1 < h - 1 ? y0 + 1 : h - 1;
            if (x0 < 0) x0 = 0; if (y0 < 0) y0 = 0;
            for (int c = 0; c < 3; c++) {
                const float p00 = (float)src[4 * ((size_t)y0 * w + x0) + c], p10 = (float)src[4 * ((size_t)y0 * w + x1) + c];
                const float p01 = (float)src[4 * ((size_t)y1 * w + x0) + c], p11 = (float)src[4 * ((size_t)y1 * w + x1) + c];
                float top = p00 * bx; { float u = p10 * ax; top = top + u; }
                float bot = p01 * bx; { float u = p11 * ax; bot = bot + u; }
                float v = top * by; { float u = bot * ay; v = v + u; }
                int q = (int)v; if (q < 0) q = 0; if (q > 255) q = 255;
                dst[4 * ((size_t)j * rw + i) + c] = (uint8_t)q;
            }
            dst[4 * ((size_t)j * rw + i) + 3] = 255;
        }
}

/* geterr's running sum (source/triangulation.hpp:653-674): float32, index order */
float tpo_sum_energy(const int32_t *ten, int NT)
{
    float s = 0.0f;
    for (int i = 0; i < NT; i++) { float e = 0.0f; e += (float)ten[i]; s += e; }
    return s;
}
