// carry.cuh -- triangulation::warp / reversewarp on the device (source/triangulation.hpp:465-520 with
// barycentric / intriangle / cartesian of source/utility.hpp:28-83): every non-boundary point is
// located in the FIRST triangle (index order) of the source embedding that strictly contains it and
// re-expressed with the same barycentrics in the target embedding.  The reference does this on the
// host by brute force, NP x NT 3x3 inverses per level; here one thread owns a point and scans the
// triangles in index order, so "first match wins" holds without any inter-thread ordering.
//
// Arithmetic is GLM's, operation for operation (cofactor determinant / inverse, column-major
// mat3 * vec3, products summed left to right), in float32 without FMA, so results equal the host
// path of include/tpose/triangulation bit for bit (tests/test_schedule_gpu.py).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tpose_dev {

constexpr int kCarryThreads = 128;
constexpr int kCarryTile = 256;      // triangles staged per pass

struct CarryTri { float2 a, b, c; int4 idx; };

__global__ void __launch_bounds__(kCarryThreads)
k_carry(const int4* __restrict__ tri, int nt, const float2* __restrict__ from, const float2* __restrict__ to,
        float2* __restrict__ pts, int npts, float ratio)
{
    __shared__ CarryTri s_tri[kCarryTile];
    const int i = blockIdx.x * kCarryThreads + threadIdx.x;
    float2 p = make_float2(0.0f, 0.0f);
    bool open = false;                       // still looking for a triangle
    if (i < npts) {
        p = pts[i];
        // triangulation::boundary(vec2) (triangulation.hpp:132-140): such points are never moved
        open = !(p.x <= -ratio || p.y <= -1.0f || p.x >= ratio || p.y >= 1.0f);
    }
    for (int t0 = 0; t0 < nt; t0 += kCarryTile) {
        __syncthreads();
        for (int k = threadIdx.x; k < kCarryTile && t0 + k < nt; k += kCarryThreads) {
            const int4 t = tri[t0 + k];
            s_tri[k] = CarryTri{from[t.x], from[t.y], from[t.z], t};
        }
        __syncthreads();
        if (!__syncthreads_or(open)) break;
        if (!open) continue;
        const int lim = min(kCarryTile, nt - t0);
        for (int k = 0; k < lim; k++) {
            const float2 a = s_tri[k].a, b = s_tri[k].b, c = s_tri[k].c;
            // intriangle (utility.hpp:43-53): zero-length sides are rejected first
            if ((a.x - b.x) * (a.x - b.x) + (a.y - b.y) * (a.y - b.y) == 0.0f) continue;
            if ((b.x - c.x) * (b.x - c.x) + (b.y - c.y) * (b.y - c.y) == 0.0f) continue;
            if ((c.x - a.x) * (c.x - a.x) + (c.y - a.y) * (c.y - a.y) == 0.0f) continue;
            // R = mat3(1, a.x, a.y,  1, b.x, b.y,  1, c.x, c.y): columns (1,ax,ay), (1,bx,by), (1,cx,cy)
            // m[col][row]:  m00 = 1 m01 = ax m02 = ay | m10 = 1 m11 = bx m12 = by | m20 = 1 m21 = cx m22 = cy
            const float m01 = a.x, m02 = a.y, m11 = b.x, m12 = b.y, m21 = c.x, m22 = c.y;
            const float c0 = m11 * m22 - m21 * m12;
            const float c1 = m01 * m22 - m21 * m02;
            const float c2 = m01 * m12 - m11 * m02;
            const float det = 1.0f * c0 - 1.0f * c1 + 1.0f * c2;
            if ((double)fabsf(det) < 1E-8) continue;             // barycentric -> (1,1,1): not inside
            const float inv = 1.0f / det;
            const float r00 = +(m11 * m22 - m21 * m12) * inv;
            const float r10 = -(1.0f * m22 - 1.0f * m12) * inv;
            const float r20 = +(1.0f * m21 - 1.0f * m11) * inv;
            const float r01 = -(m01 * m22 - m21 * m02) * inv;
            const float r11 = +(1.0f * m22 - 1.0f * m02) * inv;
            const float r21 = -(1.0f * m21 - 1.0f * m01) * inv;
            const float r02 = +(m01 * m12 - m11 * m02) * inv;
            const float r12 = -(1.0f * m12 - 1.0f * m02) * inv;
            const float r22 = +(1.0f * m11 - 1.0f * m01) * inv;
            // inverse(R) * vec3(1, p.x, p.y)
            const float sx = r00 * 1.0f + r10 * p.x + r20 * p.y;
            const float sy = r01 * 1.0f + r11 * p.x + r21 * p.y;
            const float sz = r02 * 1.0f + r12 * p.x + r22 * p.y;
            if (sx <= 0.0f || sx >= 1.0f || sy <= 0.0f || sy >= 1.0f || sz <= 0.0f || sz >= 1.0f) continue;
            // cartesian (utility.hpp:79-83): s.x * v[t.x] + s.y * v[t.y] + s.z * v[t.z]
            const int4 t = s_tri[k].idx;
            const float2 ta = to[t.x], tb = to[t.y], tc = to[t.z];
            p.x = (sx * ta.x + sy * tb.x) + sz * tc.x;
            p.y = (sx * ta.y + sy * tb.y) + sz * tc.y;
            pts[i] = p;
            open = false;
            break;
        }
    }
}

}  // namespace tpose_dev
