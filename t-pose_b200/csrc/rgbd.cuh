// rgbd.cuh -- the plane-fit meshing core of software/rgbd (main.cpp:244-349 of the reference): every
// triangle of a finished triangulation gathers the depth points whose colour-image pixel falls strictly
// inside it, fits a plane through them (mean position, mean normal), un-projects its three corners onto
// that plane, and every vertex ends up at the average of the corners its triangles produced.
//
// The reference does the gather on the host by brute force -- for each triangle a loop over ALL points
// (bounding box, then tpose::intriangle), O(NT * Npoints) -- and leans on an unreleased library (piolib) for
// three geometry helpers.  They are restated here as a plain pinhole model, consistent with each other:
//     project(p)            = (fx * p.x / p.z + ppx, fy * p.y / p.z + ppy)
//     unproject(u, mean, n) = s * d, d = ((u.x - ppx) / fx, (u.y - ppy) / fy, 1), s = dot(n, mean) / dot(n, d)
// (the ray through pixel u intersected with the plane through `mean` with normal n; w = 1).
//
//   k_rgbd_fit     one block per triangle: all threads test a chunk of points at a time (bounding box in pixel
//                  space, then barycentric / intriangle with GLM's arithmetic -- carry.cuh); the hits of a
//                  chunk are summed by one thread IN INDEX ORDER, because the reference iterates a
//                  std::set<size_t> and float addition is not associative.  Then the plane and the three
//                  un-projected corners of the triangle.
//   k_rgbd_gather  one thread per vertex: sums the corners its triangles produced in triangle order
//                  (main.cpp:322-327) and divides by their number (:338-339; 0/0 = NaN for a vertex none of
//                  whose triangles holds a point, as coded).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tpose_dev {

constexpr int kRgbdThreads = 256;

struct RgbdIntr { float fx, fy, ppx, ppy; };

// T (mesh -> pixel of the colour image) and TB = inverse(T) as the caller built them (rgbd/main.cpp:94-99), column-major
// like glm::mat3: m[3 * col + row]
struct RgbdMap { float T[9], TB[9]; };
// glm mat3 * vec3(p, 1), .xy: (m[0][r] x + m[1][r] y) + m[2][r] * 1
__device__ __forceinline__ float2 rgbd_apply(const float* m, float2 p)
{
    return make_float2(__fadd_rn(__fadd_rn(__fmul_rn(m[0], p.x), __fmul_rn(m[3], p.y)), __fmul_rn(m[6], 1.0f)),
                       __fadd_rn(__fadd_rn(__fmul_rn(m[1], p.x), __fmul_rn(m[4], p.y)), __fmul_rn(m[7], 1.0f)));
}
__device__ __forceinline__ float2 rgbd_T(float2 p, const RgbdMap& m) { return rgbd_apply(m.T, p); }

// tpose::intriangle (utility.hpp:43-53) with barycentric (:28-39): strict interior, GLM's inverse
__device__ __forceinline__ bool rgbd_intriangle(float2 p, float2 a, float2 b, float2 c)
{
    if ((a.x - b.x) * (a.x - b.x) + (a.y - b.y) * (a.y - b.y) == 0.0f) return false;
    if ((b.x - c.x) * (b.x - c.x) + (b.y - c.y) * (b.y - c.y) == 0.0f) return false;
    if ((c.x - a.x) * (c.x - a.x) + (c.y - a.y) * (c.y - a.y) == 0.0f) return false;
    const float m01 = a.x, m02 = a.y, m11 = b.x, m12 = b.y, m21 = c.x, m22 = c.y;
    const float c0 = __fsub_rn(__fmul_rn(m11, m22), __fmul_rn(m21, m12));
    const float c1 = __fsub_rn(__fmul_rn(m01, m22), __fmul_rn(m21, m02));
    const float c2 = __fsub_rn(__fmul_rn(m01, m12), __fmul_rn(m11, m02));
    const float det = __fadd_rn(__fsub_rn(c0, c1), c2);
    if ((double)fabsf(det) < 1E-8) return false;
    const float inv = __fdiv_rn(1.0f, det);
    const float r00 = __fmul_rn(c0, inv), r10 = __fmul_rn(-__fsub_rn(m22, m12), inv), r20 = __fmul_rn(__fsub_rn(m21, m11), inv);
    const float r01 = __fmul_rn(-c1, inv), r11 = __fmul_rn(__fsub_rn(m22, m02), inv), r21 = __fmul_rn(-__fsub_rn(m21, m01), inv);
    const float r02 = __fmul_rn(c2, inv), r12 = __fmul_rn(-__fsub_rn(m12, m02), inv), r22 = __fmul_rn(__fsub_rn(m11, m01), inv);
    const float sx = __fadd_rn(__fadd_rn(r00, __fmul_rn(r10, p.x)), __fmul_rn(r20, p.y));
    const float sy = __fadd_rn(__fadd_rn(r01, __fmul_rn(r11, p.x)), __fmul_rn(r21, p.y));
    const float sz = __fadd_rn(__fadd_rn(r02, __fmul_rn(r12, p.x)), __fmul_rn(r22, p.y));
    return !(sx <= 0.0f || sx >= 1.0f || sy <= 0.0f || sy >= 1.0f || sz <= 0.0f || sz >= 1.0f);
}

__device__ __forceinline__ float4 rgbd_unproject(float2 u, const RgbdIntr& k, float3 mean, float3 n)
{
    const float dx = __fdiv_rn(__fsub_rn(u.x, k.ppx), k.fx), dy = __fdiv_rn(__fsub_rn(u.y, k.ppy), k.fy);
    const float num = __fadd_rn(__fadd_rn(__fmul_rn(n.x, mean.x), __fmul_rn(n.y, mean.y)), __fmul_rn(n.z, mean.z));
    const float den = __fadd_rn(__fadd_rn(__fmul_rn(n.x, dx), __fmul_rn(n.y, dy)), n.z);
    const float s = __fdiv_rn(num, den);
    return make_float4(__fmul_rn(s, dx), __fmul_rn(s, dy), s, 1.0f);
}

// corners: 3 float4 per triangle; has[t] = 1 if the triangle holds at least one point
__global__ void __launch_bounds__(kRgbdThreads)
k_rgbd_fit(const int4* __restrict__ tri, int nt, const float2* __restrict__ pts, RgbdMap map,
           const float2* __restrict__ pixels, const float4* __restrict__ points3, const float4* __restrict__ normals, int npts,
           RgbdIntr intr, float4* __restrict__ corners, int* __restrict__ has, int* __restrict__ count)
{
    __shared__ unsigned char s_hit[kRgbdThreads];
    const int t = blockIdx.x;
    if (t >= nt) return;
    const int4 tv = tri[t];
    const float2 a = pts[tv.x], b = pts[tv.y], c = pts[tv.z];
    // bounding box in mesh space, mapped to pixels (main.cpp:260-270); TB * pixel brings a pixel back (:281)
    const float2 lo = rgbd_T(make_float2(fminf(fminf(a.x, b.x), c.x), fminf(fminf(a.y, b.y), c.y)), map);
    const float2 hi = rgbd_T(make_float2(fmaxf(fmaxf(a.x, b.x), c.x), fmaxf(fmaxf(a.y, b.y), c.y)), map);
    float3 mean = make_float3(0.0f, 0.0f, 0.0f), nrm = make_float3(0.0f, 0.0f, 0.0f);
    int n = 0;
    for (int c0 = 0; c0 < npts; c0 += kRgbdThreads) {
        const int i = c0 + threadIdx.x;
        bool hit = false;
        if (i < npts) {
            const float2 px = pixels[i];
            if (!(px.x < lo.x) && !(px.x > hi.x) && !(px.y < lo.y) && !(px.y > hi.y)) {
                hit = rgbd_intriangle(rgbd_apply(map.TB, px), a, b, c);
            }
        }
        s_hit[threadIdx.x] = hit ? 1 : 0;
        if (__syncthreads_or(hit)) {
            if (threadIdx.x == 0) {
                const int lim = min(kRgbdThreads, npts - c0);
                for (int k = 0; k < lim; k++)
                    if (s_hit[k]) {
                        const float4 p = points3[c0 + k], q = normals[c0 + k];
                        mean.x = __fadd_rn(mean.x, p.x); mean.y = __fadd_rn(mean.y, p.y); mean.z = __fadd_rn(mean.z, p.z);
                        nrm.x = __fadd_rn(nrm.x, q.x); nrm.y = __fadd_rn(nrm.y, q.y); nrm.z = __fadd_rn(nrm.z, q.z);
                        n++;
                    }
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        has[t] = n > 0;
        if (count) count[t] = n;
        if (n > 0) {
            const float fn = (float)n;
            mean = make_float3(__fdiv_rn(mean.x, fn), __fdiv_rn(mean.y, fn), __fdiv_rn(mean.z, fn));
            nrm = make_float3(__fdiv_rn(nrm.x, fn), __fdiv_rn(nrm.y, fn), __fdiv_rn(nrm.z, fn));
            // glm::normalize: v * inversesqrt(dot(v, v)) with inversesqrt = 1 / sqrt
            const float d = __fadd_rn(__fadd_rn(__fmul_rn(nrm.x, nrm.x), __fmul_rn(nrm.y, nrm.y)), __fmul_rn(nrm.z, nrm.z));
            const float is = __fdiv_rn(1.0f, __fsqrt_rn(d));
            nrm = make_float3(__fmul_rn(nrm.x, is), __fmul_rn(nrm.y, is), __fmul_rn(nrm.z, is));
            corners[3 * t + 0] = rgbd_unproject(rgbd_T(a, map), intr, mean, nrm);
            corners[3 * t + 1] = rgbd_unproject(rgbd_T(b, map), intr, mean, nrm);
            corners[3 * t + 2] = rgbd_unproject(rgbd_T(c, map), intr, mean, nrm);
        }
    }
}

__global__ void k_rgbd_gather(const int4* __restrict__ tri, int nt, int np, const float4* __restrict__ corners,
                              const int* __restrict__ has, float4* __restrict__ verts, float* __restrict__ ring)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= np) return;
    float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    float n = 0.0f;
    for (int t = 0; t < nt; t++) {
        if (!has[t]) continue;
        const int4 tv = tri[t];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const int vj = j == 0 ? tv.x : (j == 1 ? tv.y : tv.z);
            if (vj != v) continue;
            const float4 p = corners[3 * t + j];
            acc.x = __fadd_rn(acc.x, p.x); acc.y = __fadd_rn(acc.y, p.y); acc.z = __fadd_rn(acc.z, p.z); acc.w = __fadd_rn(acc.w, p.w);
            n = __fadd_rn(n, 1.0f);
        }
    }
    verts[v] = make_float4(__fdiv_rn(acc.x, n), __fdiv_rn(acc.y, n), __fdiv_rn(acc.z, n), __fdiv_rn(acc.w, n));
    if (ring) ring[v] = n;
}

}  // namespace tpose_dev
