// eval.cuh -- the fused evaluation kernel: for every triangle, the six pixel moments of all 13
// finite-difference variants, from 27 edge integrals over the row-prefix moment table; then, as
// the flags ask, cn / ca / ten in the reference SSBO layout, the gradient scatter and the vertex
// update.  One launch replaces the two 13*NT-instance draws plus gradient.cs and shift.cs
// (software/triangulate/main.cpp:121-155; shaders triangle.vs/.fs, gradient.cs, shift.cs).
//
// A launch can also carry the host-side checks of the PREVIOUS frame (EV_CHECK_PREV), so that a
// run of frames needs no host round trip and no serial check kernel between frames:
//   * geterr (source/triangulation.hpp:653-674) over the previous frame's base energies, by one
//     thread of block 0's service warp, overlapped with the edge walks of this launch;
//   * a conservative test of the per-frame topology edits of software/triangulate/main.cpp:316-346
//     (prune / >0.8*pi flip / short-edge collapse) on the host's ("watch") topology at the current
//     points, by the service warps.
// If either fires, the last block does NOT apply this launch's shift and raises the stop flag:
// points, and the other buffer set, are then exactly what the reference holds after the frame
// whose checks fired.  The evaluation done by this launch is speculative and discarded.
#pragma once
#include "common.cuh"

namespace tpose_dev {

enum EvalFlags : unsigned {
    EV_WRITE_CN    = 1u,   // mode 0: cn[g] = fragment count                       (triangle.fs:30)
    EV_WRITE_CA    = 2u,   // mode 0, triangulate flavour: ca[g].rgb               (triangle.fs:31-33)
    EV_WRITE_TEN   = 4u,   // mode 1: ten[g]                                       (triangle.fs:37-43)
    EV_MEAN_STORED = 8u,   // mean colour comes from the ca/cn buffers, not from this pass
    EV_MEAN_DIVIDE = 16u,  // mean = ca/cn (triangulate, triangle.fs:40); else mean = ca (warp :49-50)
    EV_GRAD_SHIFT  = 32u,  // gradient.cs + shift.cs fused behind the energies
    EV_CHECK_PREV  = 64u,  // previous frame's geterr + topology watch ride on this launch
    EV_FINAL       = 128u, // last launch of a run: checks only, never shifts
    EV_HONOUR_STOP = 256u  // launch is a no-op once the job's stop flag is up
};

// trigger bits (mirrors include/tpose_cuda.h)
constexpr int kTrigConverged = 1, kTrigMaxSteps = 2, kTrigTopology = 4;

// run_state words
enum { RS_STOP = 0, RS_STEPS = 1, RS_TRIG = 2, RS_SPARE = 3 };

// Everything one image-step needs; lives in device memory so that a batch of independent jobs can
// share one launch (blockIdx.y = job).
struct EvalJob {
    const uint4* table;        // row-prefix moments, pitch entries per row
    int pitch, W, H;
    float2* pts;
    const int4* tri;
    int NT, NP;
    float ratio, dp, step;
    int* cn[2]; int4* ca[2]; int* ten[2];   // two buffer sets, alternating per launch
    int2* gr;                  // pgradbuf as the reference sees it (written by the vertex update)
    int2* gr_acc;              // accumulation scratch, zero between launches
    const int4* wtri;          // watch topology (the host's triangles / halfedges), may be null
    const int* whe;
    float* err_state;          // [0] toterr [1] |relerr| [2] newerr [3] converged flag of the last check
    int* run_state;            // RS_*
    unsigned* trig;            // [0] topology trigger accumulator
    unsigned* ticket;          // last-block election counter
};

template <int G>
__device__ __forceinline__ void group_reduce(Moments& m)
{
#pragma unroll
    for (int o = G >> 1; o > 0; o >>= 1) {
        m.n   += __shfl_xor_sync(0xffffffffu, m.n, o);
        m.odd += __shfl_xor_sync(0xffffffffu, m.odd, o);
        m.r   += __shfl_xor_sync(0xffffffffu, m.r, o);
        m.g   += __shfl_xor_sync(0xffffffffu, m.g, o);
        m.b   += __shfl_xor_sync(0xffffffffu, m.b, o);
        m.s2  += __shfl_xor_sync(0xffffffffu, m.s2, o);
    }
}

// shift.cs:16-47 for one vertex; gx, gy are the integer gradient sums
__device__ __forceinline__ float2 shift_vertex(float2 p, int gxi, int gyi, float ratio, float step)
{
    float gx = (float)gxi, gy = (float)gyi;
    if (p.x <= -ratio) { p.x = -ratio; gx = 0.0f; }
    else if (p.x >= ratio) { p.x = ratio; gx = 0.0f; }
    if (p.y <= -1.0f) { p.y = -1.0f; gy = 0.0f; }
    else if (p.y >= 1.0f) { p.y = 1.0f; gy = 0.0f; }
    const float sx = __fdiv_rn(__fdiv_rn(__fmul_rn(step, gx), 256.0f), 256.0f);
    const float sy = __fdiv_rn(__fdiv_rn(__fmul_rn(step, gy), 256.0f), 256.0f);
    p.x = __fsub_rn(p.x, sx);
    p.y = __fsub_rn(p.y, sy);
    return p;
}

// ---- per-frame topology watch (conservative superset of triangulate/main.cpp:316-346) --------

__device__ __forceinline__ bool on_boundary(float2 p, float ratio)   // triangulation.hpp:132-140
{
    return p.x <= -ratio || p.y <= -1.0f || p.x >= ratio || p.y >= 1.0f;
}
__device__ __forceinline__ float2 vsub(float2 a, float2 b) { return make_float2(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y)); }
__device__ __forceinline__ float vdot(float2 a, float2 b) { return __fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)); }
__device__ __forceinline__ float vlen(float2 a) { return __fsqrt_rn(vdot(a, a)); }
// argument of the acos in triangulation::angle (triangulation.hpp:105-116); 1 if a side is degenerate
__device__ __forceinline__ float angle_cos(float2 paa, float2 pab, float2 pac)
{
    const float2 u = vsub(paa, pac), v = vsub(pab, pac);
    const float lu = vlen(u), lv = vlen(v);
    if (lu == 0.0f || lv == 0.0f) return 1.0f;
    return __fdiv_rn(__fdiv_rn(vdot(u, v), lu), lv);
}
// the ccw predicate of triangulation::flip (triangulation.hpp:257-259)
__device__ __forceinline__ bool ccw(float2 A, float2 B, float2 C)
{
    return __fmul_rn(__fsub_rn(C.y, A.y), __fsub_rn(B.x, A.x)) > __fmul_rn(__fsub_rn(B.y, A.y), __fsub_rn(C.x, A.x));
}
__device__ __forceinline__ int corner(const int4& t, int k) { return k == 0 ? t.x : (k == 1 ? t.y : t.z); }

__device__ bool watch_triangle(const EvalJob& J, int t)
{
    const int4 tv = J.wtri[t];
    const float2 p[3] = {J.pts[tv.x], J.pts[tv.y], J.pts[tv.z]};
    const int he[3] = {J.whe[3 * t], J.whe[3 * t + 1], J.whe[3 * t + 2]};
    // cos of the angle opposite halfedge k (at corner k+2)
    float cs[3];
#pragma unroll
    for (int k = 0; k < 3; k++) cs[k] = angle_cos(p[k], p[(k + 1) % 3], p[(k + 2) % 3]);
    // (c) collapse (triangulation.hpp:330-340): shortest side <= 0.01 (tiny safety margin)
    const float l0 = vlen(vsub(p[1], p[0])), l1 = vlen(vsub(p[2], p[1])), l2 = vlen(vsub(p[0], p[2]));
    if (fminf(l0, fminf(l1, l2)) <= 0.010001f) return true;
    // (a) prune (triangulation.hpp:220-241): all corners on the boundary, an open side, every
    //     angle outside (0, pi): acos argument not strictly inside (-1, 1)
    if (on_boundary(p[0], J.ratio) && on_boundary(p[1], J.ratio) && on_boundary(p[2], J.ratio) &&
        (he[0] < 0 || he[1] < 0 || he[2] < 0)) {
        bool flat = true;
#pragma unroll
        for (int k = 0; k < 3; k++) if (fabsf(cs[k]) < 0.99999f) flat = false;   // NaN compares false -> flat
        if (flat) return true;
    }
    // (b) angle > 0.8*pi (main.cpp:325-334) with a flippable edge: twin exists and the quad is convex
#pragma unroll
    for (int k = 0; k < 3; k++) {
        if (!(cs[k] < -0.8089f)) continue;                      // cos(0.8 pi) = -0.80902
        const int hb = he[k];
        if (hb < 0) continue;
        const int tb = hb / 3, kb = hb - 3 * tb;
        if (tb >= J.NT) return true;                            // inconsistent watch data: let the host look
        const int4 tvb = J.wtri[tb];
        const float2 A = p[k], C = p[(k + 2) % 3];
        const float2 B = J.pts[corner(tvb, kb)], D = J.pts[corner(tvb, (kb + 2) % 3)];
        if (ccw(A, C, D) == ccw(B, C, D) || ccw(A, B, C) == ccw(A, B, D)) continue;
        return true;
    }
    return false;
}

// geterr (source/triangulation.hpp:653-674): float32 running sum of the base energies in index
// order -- order matters, so it is one thread.
__device__ void geterr_serial(const int* __restrict__ ten_base, int nt, float* state, double rel_tol)
{
    float newerr = 0.0f;
    int i = 0;
    for (; i + 4 <= nt; i += 4) {
        const int4 v = *reinterpret_cast<const int4*>(ten_base + i);
        newerr = __fadd_rn(newerr, (float)v.x);
        newerr = __fadd_rn(newerr, (float)v.y);
        newerr = __fadd_rn(newerr, (float)v.z);
        newerr = __fadd_rn(newerr, (float)v.w);
    }
    for (; i < nt; i++) newerr = __fadd_rn(newerr, (float)ten_base[i]);
    const float tot = state[0];
    const float rel = __fdiv_rn(__fsub_rn(tot, newerr), tot);
    state[0] = newerr;
    state[1] = fabsf(rel);
    state[2] = newerr;
    // the reference compares the float against a double literal (`geterr(&tr) < 1E-4`)
    state[3] = ((double)fabsf(rel) < rel_tol) ? 1.0f : 0.0f;
}

// G lanes cooperate on one edge variant; TPB triangles per block; one extra service warp.
template <int G, int TPB>
struct EvalCfg {
    static constexpr int kItems = TPB * kEdgeVariants;
    static constexpr int kWork = ((kItems * G + 31) / 32) * 32;   // threads that walk edges
    static constexpr int kThreads = kWork + 32;                   // + service warp
};

template <int G, int TPB>
__global__ void __launch_bounds__(EvalCfg<G, TPB>::kThreads)
k_eval(const EvalJob* __restrict__ jobs, int set, unsigned flags, double rel_tol)
{
    using Cfg = EvalCfg<G, TPB>;
    constexpr int kItems = Cfg::kItems;
    __shared__ EvalJob J;
    __shared__ int2 s_fx[TPB][3][5];            // snapped vertices: [triangle][corner][0 base,1..4 moved]
    __shared__ long long s_edge[kItems][4];     // signed edge integrals: r, g, b, s2
    __shared__ int s_edge_n[kItems][2];         //                        n, odd
    __shared__ int s_ten[TPB][kVariants];
    __shared__ int4 s_tri[TPB];
    __shared__ bool s_last;

    const int tid = threadIdx.x;
    if (tid < (int)(sizeof(EvalJob) / sizeof(int)))
        reinterpret_cast<int*>(&J)[tid] = reinterpret_cast<const int*>(jobs + blockIdx.y)[tid];
    __syncthreads();
    const int nblocks = (J.NT + TPB - 1) / TPB;
    if ((int)blockIdx.x >= nblocks) return;
    if ((flags & EV_HONOUR_STOP) && __ldcg(&J.run_state[RS_STOP])) return;
    const int t0 = blockIdx.x * TPB;
    int* const cn = J.cn[set]; int4* const ca = J.ca[set]; int* const ten = J.ten[set];

    // phase 0: snapped positions of the 15 (corner, move) combinations of each triangle
    if (tid < TPB * 15) {
        const int tb = tid / 15, j = (tid % 15) / 5, s5 = tid % 5;
        const int t = t0 + tb;
        if (t < J.NT) {
            const int4 tv = J.tri[t];
            if (j == 0 && s5 == 0) s_tri[tb] = tv;
            const int v = (j == 0) ? tv.x : (j == 1) ? tv.y : tv.z;
            const float2 pv = J.pts[v];
            float dx = 0.0f, dy = 0.0f;
            if (s5 > 0) fd_offset(s5 - 1, J.dp, dx, dy);
            Xform xf{J.ratio, 0.5f * (float)J.W, 0.5f * (float)J.H, 256 * J.H};
            s_fx[tb][j][s5] = to_fixed(pv, dx, dy, xf);
        }
    }
    __syncthreads();

    if (tid < Cfg::kWork) {
        // phase 1: edge integrals.  item = (triangle, edge e, move q): q=0 base, 1..4 corner e moved,
        // 5..8 corner e+1 moved.
        const int it = tid / G, lane = tid % G;
        Moments m; m.zero();
        int sign = 0;
        if (it < kItems) {
            const int tb = it / kEdgeVariants, ev = it % kEdgeVariants;
            const int e = ev / 9, q = ev % 9;
            if (t0 + tb < J.NT) {
                const int ja = e, jb = (e + 1) % 3;
                const int2 a = s_fx[tb][ja][(q >= 1 && q <= 4) ? q : 0];
                const int2 b = s_fx[tb][jb][(q >= 5) ? q - 4 : 0];
                EdgeWalk w;
                w.init(a, b, J.H, lane, G);
                sign = w.sign;
                const uint4* __restrict__ table = J.table;
                const int pitch = J.pitch, W = J.W;
#pragma unroll 4
                while (!w.done()) {
                    const int c = w.cross(W);
                    const uint4 en = __ldg(table + (size_t)w.r * pitch + c);
                    m.add_entry(en, c);
                    w.next(G);
                }
            }
        }
        group_reduce<G>(m);
        if (it < kItems && lane == 0) {
            s_edge[it][0] = sign * m.r;  s_edge[it][1] = sign * m.g;
            s_edge[it][2] = sign * m.b;  s_edge[it][3] = sign * m.s2;
            s_edge_n[it][0] = sign * m.n; s_edge_n[it][1] = sign * m.odd;
        }
    } else if (flags & EV_CHECK_PREV) {
        // service warp: the previous frame's host-side checks
        const int sl = tid - Cfg::kWork;
        if (blockIdx.x == 0 && sl == 31) geterr_serial(J.ten[set ^ 1], J.NT, J.err_state, rel_tol);
        if (J.wtri && sl < TPB && t0 + sl < J.NT && watch_triangle(J, t0 + sl)) atomicOr(J.trig, (unsigned)kTrigTopology);
    }
    __syncthreads();

    // phase 2: one thread per (triangle, variant k): three edge integrals -> moments -> outputs
    if (tid < TPB * kVariants) {
        const int tb = tid / kVariants, k = tid % kVariants;
        const int t = t0 + tb;
        if (t < J.NT) {
            int i0, i1, i2;
            const int base = tb * kEdgeVariants;
            if (k == 0) { i0 = base; i1 = base + 9; i2 = base + 18; }
            else {
                const int j = (k - 1) >> 2, s = (k - 1) & 3;
                i0 = base + 9 * j + 1 + s;                  // edge j   : its first corner moved
                i1 = base + 9 * ((j + 2) % 3) + 5 + s;      // edge j+2 : its second corner moved
                i2 = base + 9 * ((j + 1) % 3);              // opposite edge, unchanged
            }
            long long sr = s_edge[i0][0] + s_edge[i1][0] + s_edge[i2][0];
            long long sg = s_edge[i0][1] + s_edge[i1][1] + s_edge[i2][1];
            long long sb = s_edge[i0][2] + s_edge[i1][2] + s_edge[i2][2];
            long long s2 = s_edge[i0][3] + s_edge[i1][3] + s_edge[i2][3];
            int n   = s_edge_n[i0][0] + s_edge_n[i1][0] + s_edge_n[i2][0];
            int odd = s_edge_n[i0][1] + s_edge_n[i1][1] + s_edge_n[i2][1];
            if (n < 0) { n = -n; odd = -odd; sr = -sr; sg = -sg; sb = -sb; s2 = -s2; }
            const int g = k * J.NT + t;
            if (flags & EV_WRITE_CN) cn[g] = n;
            if (flags & EV_WRITE_CA) ca[g] = make_int4((int)sr, (int)sg, (int)sb, 0);
            if (flags & EV_WRITE_TEN) {
                int m0 = 0, m1 = 0, m2 = 0; bool have = true;
                if (flags & EV_MEAN_STORED) {
                    const int4 c = ca[g];
                    if (flags & EV_MEAN_DIVIDE) {
                        const int cnt = cn[g];
                        if (cnt > 0) { m0 = c.x / cnt; m1 = c.y / cnt; m2 = c.z / cnt; } else have = false;
                    } else { m0 = c.x; m1 = c.y; m2 = c.z; }
                } else {
                    if (n > 0) { m0 = (int)(sr / n); m1 = (int)(sg / n); m2 = (int)(sb / n); } else have = false;
                }
                const long long e = have ? energy_closed_form(n, odd, sr, sg, sb, s2, m0, m1, m2) : 0ll;
                const int e32 = (int)(unsigned)(unsigned long long)e;   // int32 accumulator wraps
                ten[g] = e32;
                s_ten[tb][k] = e32;
            }
        }
    }
    if (!(flags & EV_GRAD_SHIFT)) return;
    __syncthreads();

    // phase 3: gradient.cs:27-34 -- central differences scattered to the three corners
    if (tid < TPB * 3) {
        const int tb = tid / 3, j = tid % 3;
        const int t = t0 + tb;
        if (t < J.NT) {
            const int4 tv = s_tri[tb];
            const int v = (j == 0) ? tv.x : (j == 1) ? tv.y : tv.z;
            const unsigned gx = (unsigned)s_ten[tb][1 + 4 * j] - (unsigned)s_ten[tb][2 + 4 * j];
            const unsigned gy = (unsigned)s_ten[tb][3 + 4 * j] - (unsigned)s_ten[tb][4 + 4 * j];
            atomicAdd((unsigned*)&J.gr_acc[v].x, gx);
            atomicAdd((unsigned*)&J.gr_acc[v].y, gy);
        }
    }
    // phase 4: the last block of this job to get here owns the commit: either the vertex update
    // (shift.cs) or, if the previous frame's checks fired, the stop.
    __syncthreads();
    if (tid == 0) {
        __threadfence();      // cumulative: covers this block's writes and atomics (ordered by the barrier)
        s_last = (atomicAdd(J.ticket, 1u) == (unsigned)nblocks - 1u);
        if (s_last) __threadfence();
    }
    __syncthreads();
    if (!s_last) return;
    bool stop = false;
    if (flags & EV_CHECK_PREV) {
        const bool conv = __ldcg(&J.err_state[3]) != 0.0f;
        const unsigned topo = __ldcg(J.trig);
        stop = conv || topo != 0u || (flags & EV_FINAL);
        if (tid == 0) {
            *J.trig = 0u;
            if (stop) {
                J.run_state[RS_STOP] = 1;
                J.run_state[RS_TRIG] = (conv ? kTrigConverged : 0) | (topo ? kTrigTopology : 0) |
                                       ((!conv && !topo) ? kTrigMaxSteps : 0);
            }
        }
    }
    for (int v = tid; v < J.NP; v += blockDim.x) {
        const int2 g = __ldcg(&J.gr_acc[v]);
        J.gr_acc[v] = make_int2(0, 0);
        if (stop) continue;
        J.gr[v] = g;
        if (v >= 4) J.pts[v] = shift_vertex(J.pts[v], g.x, g.y, J.ratio, J.step);
    }
    if (tid == 0) {
        *J.ticket = 0u;
        if (!stop) J.run_state[RS_STEPS] += 1;
    }
}

// un-fused gradient.cs / shift.cs, for the literal doshift() entry point
__global__ void k_zero_gradient(int2* gr, int np)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < np) gr[v] = make_int2(0, 0);
}

__global__ void k_gradient(const int4* __restrict__ tri, int nt, const int* __restrict__ ten, int2* gr)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nt) return;
    const int4 tv = tri[t];
    const int vs[3] = {tv.x, tv.y, tv.z};
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const unsigned gx = (unsigned)ten[(1 + 4 * j) * nt + t] - (unsigned)ten[(2 + 4 * j) * nt + t];
        const unsigned gy = (unsigned)ten[(3 + 4 * j) * nt + t] - (unsigned)ten[(4 + 4 * j) * nt + t];
        atomicAdd((unsigned*)&gr[vs[j]].x, gx);
        atomicAdd((unsigned*)&gr[vs[j]].y, gy);
    }
}

__global__ void k_shift(float2* pts, int np, const int2* __restrict__ gr, float ratio, float step)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < 4 || v >= np) return;
    const int2 g = gr[v];
    pts[v] = shift_vertex(pts[v], g.x, g.y, ratio, step);
}

}  // namespace tpose_dev
