// carry.cuh -- triangulation::warp / reversewarp on the device (source/triangulation.hpp:465-520 with
// barycentric / intriangle / cartesian of source/utility.hpp:28-83): every non-boundary point is
// located in the FIRST triangle (index order) of the source embedding that strictly contains it and
// re-expressed with the same barycentrics in the target embedding.  The reference does this on the
// host by brute force, NP x NT 3x3 inverses per level; here eight lanes own a point and scan the
// triangles of a tile interleaved, the inverse of a triangle is formed once per tile, and "first match
// wins" is the smallest index the eight lanes found.
//
// Arithmetic is GLM's, operation for operation (cofactor determinant / inverse, column-major
// mat3 * vec3, products summed left to right), in float32 without FMA, so results equal the host
// path of include/tpose/triangulation bit for bit (tests/test_schedule_gpu.py).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tpose_dev {

constexpr int kCarryThreads = 128;
constexpr int kCarryLanes = 8;       // lanes per point: lane l looks at triangles l, l + 8, ... of a tile
constexpr int kCarryTile = 256;      // triangles staged per pass

// inverse(R) of one source triangle (R = mat3(1, a.x, a.y,  1, b.x, b.y,  1, c.x, c.y), m[col][row]) as barycentric()
// forms it, computed ONCE per tile by the staging threads instead of once per (point, triangle): the same float
// operations in the same order, so the same bits.  ok = 0: intriangle rejects the triangle before looking at p
// (a zero-length side, utility.hpp:43-53; |det| < 1E-8, barycentric -> (1,1,1)).
struct CarryTri { float r00, r10, r20, r01, r11, r21, r02, r12, r22; int ok; int4 idx; };

__device__ __forceinline__ CarryTri carry_prepare(float2 a, float2 b, float2 c, int4 idx)
{
    CarryTri o;
    o.idx = idx; o.ok = 0;
    o.r00 = o.r10 = o.r20 = o.r01 = o.r11 = o.r21 = o.r02 = o.r12 = o.r22 = 0.0f;
    if ((a.x - b.x) * (a.x - b.x) + (a.y - b.y) * (a.y - b.y) == 0.0f) return o;
    if ((b.x - c.x) * (b.x - c.x) + (b.y - c.y) * (b.y - c.y) == 0.0f) return o;
    if ((c.x - a.x) * (c.x - a.x) + (c.y - a.y) * (c.y - a.y) == 0.0f) return o;
    // m00 = 1 m01 = ax m02 = ay | m10 = 1 m11 = bx m12 = by | m20 = 1 m21 = cx m22 = cy
    const float m01 = a.x, m02 = a.y, m11 = b.x, m12 = b.y, m21 = c.x, m22 = c.y;
    const float c0 = m11 * m22 - m21 * m12;
    const float c1 = m01 * m22 - m21 * m02;
    const float c2 = m01 * m12 - m11 * m02;
    const float det = 1.0f * c0 - 1.0f * c1 + 1.0f * c2;
    if ((double)fabsf(det) < 1E-8) return o;
    const float inv = 1.0f / det;
    o.r00 = +(m11 * m22 - m21 * m12) * inv;
    o.r10 = -(1.0f * m22 - 1.0f * m12) * inv;
    o.r20 = +(1.0f * m21 - 1.0f * m11) * inv;
    o.r01 = -(m01 * m22 - m21 * m02) * inv;
    o.r11 = +(1.0f * m22 - 1.0f * m02) * inv;
    o.r21 = -(1.0f * m21 - 1.0f * m01) * inv;
    o.r02 = +(m01 * m12 - m11 * m02) * inv;
    o.r12 = -(1.0f * m12 - 1.0f * m02) * inv;
    o.r22 = +(1.0f * m11 - 1.0f * m01) * inv;
    o.ok = 1;
    return o;
}

// kCarryLanes lanes per point.  "The first triangle in index order that contains the point" = tiles in order, and
// inside a tile the smallest index any lane found (a lane stops at its own first hit: its later ones are larger).
__global__ void __launch_bounds__(kCarryThreads)
k_carry(const int4* __restrict__ tri, int nt, const float2* __restrict__ from, const float2* __restrict__ to,
        float2* __restrict__ pts, int npts, float ratio)
{
    __shared__ CarryTri s_tri[kCarryTile];
    const int lane = threadIdx.x % kCarryLanes;
    const int i = (blockIdx.x * kCarryThreads + threadIdx.x) / kCarryLanes;
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
            s_tri[k] = carry_prepare(from[t.x], from[t.y], from[t.z], t);
        }
        __syncthreads();
        if (!__syncthreads_or(open)) break;
        const int lim = min(kCarryTile, nt - t0);
        int hit = kCarryTile;                // this lane's first containing triangle of the tile
        float sx = 0.0f, sy = 0.0f, sz = 0.0f;
        if (open)
            for (int k = lane; k < lim; k += kCarryLanes) {
                const CarryTri& q = s_tri[k];
                if (!q.ok) continue;
                // inverse(R) * vec3(1, p.x, p.y), then intriangle's strict test
                sx = q.r00 * 1.0f + q.r10 * p.x + q.r20 * p.y;
                sy = q.r01 * 1.0f + q.r11 * p.x + q.r21 * p.y;
                sz = q.r02 * 1.0f + q.r12 * p.x + q.r22 * p.y;
                if (sx <= 0.0f || sx >= 1.0f || sy <= 0.0f || sy >= 1.0f || sz <= 0.0f || sz >= 1.0f) continue;
                hit = k;
                break;
            }
        // the point's lanes sit in one warp (8 divides 32): smallest index among them
        int first = hit;
#pragma unroll
        for (int o = kCarryLanes / 2; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
        if (open && first < kCarryTile) {
            if (hit == first) {
                // cartesian (utility.hpp:79-83): s.x * v[t.x] + s.y * v[t.y] + s.z * v[t.z]
                const int4 t = s_tri[first].idx;
                const float2 ta = to[t.x], tb = to[t.y], tc = to[t.z];
                p.x = (sx * ta.x + sy * tb.x) + sz * tc.x;
                p.y = (sx * ta.y + sy * tb.y) + sz * tc.y;
                pts[i] = p;
            }
            open = false;
        }
    }
}

}  // namespace tpose_dev
