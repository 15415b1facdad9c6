// raster.cuh -- per-pixel triangle-ID raster of the NT base instances (what mode 2 of
// software/triangulate/shader/triangle.fs:46-50 shows, and what the fragment stage iterates
// over in modes 0/1).  One warp per triangle: rows in sequence, lanes across the span, so the
// stores of a row are coalesced.  Overlaps resolve to the highest instance index (draw order).
#pragma once
#include "common.cuh"

namespace tpose_dev {

__global__ void __launch_bounds__(256)
k_raster_ids(const float2* __restrict__ pts, const int4* __restrict__ tri, int NT,
             float ratio, int W, int H, int* __restrict__ ids)
{
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (t >= NT) return;
    const int4 tv = tri[t];
    Xform xf{ratio, 0.5f * (float)W, 0.5f * (float)H, 256 * H};
    int2 f[3];
    f[0] = to_fixed(pts[tv.x], 0.0f, 0.0f, xf);
    f[1] = to_fixed(pts[tv.y], 0.0f, 0.0f, xf);
    f[2] = to_fixed(pts[tv.z], 0.0f, 0.0f, xf);
    // every lane walks all three edges row by row (stride 1); a row inside the triangle is
    // crossed by exactly two of them under the half-open row rule
    EdgeWalk w[3];
    int rlo = H, rhi = 0;
#pragma unroll
    for (int e = 0; e < 3; e++) {
        w[e].init(f[e], f[(e + 1) % 3], H, 0, 1);
        if (!w[e].done()) { rlo = min(rlo, w[e].r); rhi = max(rhi, w[e].r_end); }
    }
    for (int r = rlo; r < rhi; r++) {
        int cmin = W, cmax = 0;
#pragma unroll
        for (int e = 0; e < 3; e++) {
            if (!w[e].done() && w[e].r == r) {
                const int c = w[e].cross(W);
                cmin = min(cmin, c); cmax = max(cmax, c);
                w[e].next(1);
            }
        }
        int* row = ids + (size_t)r * W;
        for (int x = cmin + lane; x < cmax; x += 32) atomicMax(row + x, t);
    }
}

// ---- headless display (software/view of the reference) ---------------------------------------------
// view/shader/triangle.vs:44-46: every corner is drawn at mix(p, op, s); glm::mix(x, y, a) = x*(1-a) + y*a
__global__ void k_mix_points(const float2* __restrict__ p, const float2* __restrict__ op, int np, float s, float2* __restrict__ out)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= np) return;
    const float2 a = p[v], b = op[v];
    const float w = __fsub_rn(1.0f, s);
    out[v] = make_float2(__fadd_rn(__fmul_rn(a.x, w), __fmul_rn(b.x, s)), __fadd_rn(__fmul_rn(a.y, w), __fmul_rn(b.y, s)));
}

// view/shader/triangle.fs:22-24 (mode 2): fragColor = vec3(ca[index].rgb) / 255 on a black clear
// colour; an RGBA8 target stores that as the clamped integer.  ids < 0: nothing drawn there.
__global__ void k_flat_colors(const int* __restrict__ ids, const int4* __restrict__ ca, int npix, uchar4* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    const int t = ids[i];
    uchar4 c = make_uchar4(0, 0, 0, 255);
    if (t >= 0) {
        const int4 v = ca[t];
        c = make_uchar4((unsigned char)min(max(v.x, 0), 255), (unsigned char)min(max(v.y, 0), 255), (unsigned char)min(max(v.z, 0), 255), 255);
    }
    out[i] = c;
}

// The viewer's wireframe (view/main.cpp:58-69,112-121: GL_LINE_STRIP over the corners of every triangle, white,
// view/shader/linestrip.fs).  GL leaves the exact fragments of a line to the driver (diamond-exit rule "or
// equivalent"); the convention here, shared with the oracle: end points snapped like the triangles' (1/256 px), a
// line is x-major if |dx| >= |dy|; an x-major line produces one fragment in every pixel column whose centre
// x lies in [min x, max x) -- in the row that holds the line's exact height at that centre -- and a y-major
// line the same with the axes swapped.  Zero-length lines draw nothing.  One thread per (triangle, side).
__global__ void k_wire(const float2* __restrict__ pts, const int4* __restrict__ tri, int NT, float ratio, int W, int H,
                       uchar4* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3 * NT) return;
    const int4 tv = tri[i / 3];
    const int k = i % 3;
    const int va = k == 0 ? tv.x : (k == 1 ? tv.y : tv.z), vb = k == 0 ? tv.y : (k == 1 ? tv.z : tv.x);
    Xform xf{ratio, 0.5f * (float)W, 0.5f * (float)H, 256 * H};
    int2 a = to_fixed(pts[va], 0.0f, 0.0f, xf), b = to_fixed(pts[vb], 0.0f, 0.0f, xf);
    const long long dx = (long long)b.x - a.x, dy = (long long)b.y - a.y;
    if (dx == 0 && dy == 0) return;
    const bool xmajor = llabs(dx) >= llabs(dy);
    // u = major axis, v = minor axis
    long long ua = xmajor ? a.x : a.y, va2 = xmajor ? a.y : a.x, ub = xmajor ? b.x : b.y, vb2 = xmajor ? b.y : b.x;
    if (ua > ub) { long long t = ua; ua = ub; ub = t; t = va2; va2 = vb2; vb2 = t; }
    const long long du = ub - ua, dv = vb2 - va2;
    const int ulim = xmajor ? W : H, vlim = xmajor ? H : W;
    int c0 = (int)max((ua - 128 + 255) >> 8, 0ll), c1 = (int)min((ub - 128 + 255) >> 8, (long long)ulim);   // centres in [ua, ub)
    for (int c = c0; c < c1; c++) {
        const long long m = 256ll * c + 128 - ua;                     // 0 <= m < du
        const long long v = va2 + floordiv64(dv * m, du);             // exact height (1/256 px) at the centre, rounded down
        const long long r = v >> 8;
        if (r < 0 || r >= vlim) continue;
        const int x = xmajor ? c : (int)r, y = xmajor ? (int)r : c;
        out[(size_t)y * W + x] = make_uchar4(255, 255, 255, 255);
    }
}

}  // namespace tpose_dev