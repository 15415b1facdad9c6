// eval.cuh -- the fused evaluation kernel: for every triangle, the six pixel moments of all 13
// finite-difference variants, from 27 edge integrals over the row-prefix moment table; then, as
// the flags ask, cn / ca / ten in the reference SSBO layout, the gradient scatter and the vertex
// update.  One launch replaces the two 13*NT-instance draws plus gradient.cs and shift.cs
// (software/triangulate/main.cpp:121-155; shaders triangle.vs/.fs, gradient.cs, shift.cs).
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
    EV_HISTORY     = 64u   // copy the base-row energies to the history slot (for the err kernel)
};

struct EvalParams {
    const uint4* __restrict__ table;  // row-prefix moments, pitch entries per row
    int pitch, W, H;
    float2* pts;
    const int4* __restrict__ tri;
    int NT, NP;
    float ratio, dp, step;
    int* cn; int4* ca; int* ten;
    int2* gr;                  // pgradbuf as the reference sees it (written by the vertex update)
    int2* gr_acc;              // accumulation scratch, zero between launches
    const int* stop;           // device flag: non-zero => this launch is a no-op (tpose_cuda_run)
    int* hist;                 // NT ints, base energies of this evaluation (EV_HISTORY)
    unsigned* ticket;          // last-block election counter
    unsigned flags;
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

// G lanes cooperate on one edge variant; TPB triangles per block.
template <int G, int TPB>
__global__ void __launch_bounds__(((TPB * kEdgeVariants * G + 31) / 32) * 32)
k_eval(const EvalParams p)
{
    constexpr int kItems = TPB * kEdgeVariants;
    __shared__ int2 s_fx[TPB][3][5];            // snapped vertices: [triangle][corner][0 base,1..4 moved]
    __shared__ long long s_edge[kItems][4];     // signed edge integrals: r, g, b, s2
    __shared__ int s_edge_n[kItems][2];         //                        n, odd
    __shared__ int s_ten[TPB][kVariants];
    __shared__ int4 s_tri[TPB];
    __shared__ bool s_last;

    const int tid = threadIdx.x;
    const int t0 = blockIdx.x * TPB;
    if (p.stop && *p.stop) return;

    // phase 0: snapped positions of the 15 (corner, move) combinations of each triangle
    if (tid < TPB * 15) {
        const int tb = tid / 15, j = (tid % 15) / 5, s5 = tid % 5;
        const int t = t0 + tb;
        if (t < p.NT) {
            const int4 tv = p.tri[t];
            if (j == 0 && s5 == 0) s_tri[tb] = tv;
            const int v = (j == 0) ? tv.x : (j == 1) ? tv.y : tv.z;
            const float2 pv = p.pts[v];
            float dx = 0.0f, dy = 0.0f;
            if (s5 > 0) fd_offset(s5 - 1, p.dp, dx, dy);
            Xform xf{p.ratio, 0.5f * (float)p.W, 0.5f * (float)p.H, 256 * p.H};
            s_fx[tb][j][s5] = to_fixed(pv, dx, dy, xf);
        }
    }
    __syncthreads();

    // phase 1: edge integrals.  item = (triangle, edge e, move q): q=0 base, 1..4 corner e moved,
    // 5..8 corner e+1 moved.
    {
        const int it = tid / G, lane = tid % G;
        Moments m; m.zero();
        int sign = 0;
        if (it < kItems) {
            const int tb = it / kEdgeVariants, ev = it % kEdgeVariants;
            const int e = ev / 9, q = ev % 9;
            if (t0 + tb < p.NT) {
                const int ja = e, jb = (e + 1) % 3;
                const int2 a = s_fx[tb][ja][(q >= 1 && q <= 4) ? q : 0];
                const int2 b = s_fx[tb][jb][(q >= 5) ? q - 4 : 0];
                EdgeWalk w;
                w.init(a, b, p.H, lane, G);
                sign = w.sign;
#pragma unroll 4
                while (!w.done()) {
                    const int c = w.cross(p.W);
                    const uint4 en = __ldg(p.table + (size_t)w.r * p.pitch + c);
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
    }
    __syncthreads();

    // phase 2: one thread per (triangle, variant k): three edge integrals -> moments -> outputs
    if (tid < TPB * kVariants) {
        const int tb = tid / kVariants, k = tid % kVariants;
        const int t = t0 + tb;
        if (t < p.NT) {
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
            const int g = k * p.NT + t;
            if (p.flags & EV_WRITE_CN) p.cn[g] = n;
            if (p.flags & EV_WRITE_CA) p.ca[g] = make_int4((int)sr, (int)sg, (int)sb, 0);
            if (p.flags & EV_WRITE_TEN) {
                int m0 = 0, m1 = 0, m2 = 0; bool have = true;
                if (p.flags & EV_MEAN_STORED) {
                    const int4 c = p.ca[g];
                    if (p.flags & EV_MEAN_DIVIDE) {
                        const int cnt = p.cn[g];
                        if (cnt > 0) { m0 = c.x / cnt; m1 = c.y / cnt; m2 = c.z / cnt; } else have = false;
                    } else { m0 = c.x; m1 = c.y; m2 = c.z; }
                } else {
                    if (n > 0) { m0 = (int)(sr / n); m1 = (int)(sg / n); m2 = (int)(sb / n); } else have = false;
                }
                const long long e = have ? energy_closed_form(n, odd, sr, sg, sb, s2, m0, m1, m2) : 0ll;
                const int e32 = (int)(unsigned)(unsigned long long)e;   // int32 accumulator wraps
                p.ten[g] = e32;
                s_ten[tb][k] = e32;
                if ((p.flags & EV_HISTORY) && k == 0) p.hist[t] = e32;
            }
        }
    }
    if (!(p.flags & EV_GRAD_SHIFT)) return;
    __syncthreads();

    // phase 3: gradient.cs:27-34 -- central differences scattered to the three corners
    if (tid < TPB * 3) {
        const int tb = tid / 3, j = tid % 3;
        const int t = t0 + tb;
        if (t < p.NT) {
            const int4 tv = s_tri[tb];
            const int v = (j == 0) ? tv.x : (j == 1) ? tv.y : tv.z;
            const unsigned gx = (unsigned)s_ten[tb][1 + 4 * j] - (unsigned)s_ten[tb][2 + 4 * j];
            const unsigned gy = (unsigned)s_ten[tb][3 + 4 * j] - (unsigned)s_ten[tb][4 + 4 * j];
            atomicAdd((unsigned*)&p.gr_acc[v].x, gx);
            atomicAdd((unsigned*)&p.gr_acc[v].y, gy);
        }
    }
    // phase 4: the last block to get here owns the vertex update (shift.cs) and re-zeroes the scratch
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(p.ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int v = tid; v < p.NP; v += blockDim.x) {
        const int2 g = __ldcg(&p.gr_acc[v]);
        p.gr_acc[v] = make_int2(0, 0);
        p.gr[v] = g;
        if (v >= 4) p.pts[v] = shift_vertex(p.pts[v], g.x, g.y, p.ratio, p.step);
    }
    if (tid == 0) *p.ticket = 0u;
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

// geterr (source/triangulation.hpp:653-674): float32 running sum of the base energies in index
// order -- order matters, so it is one thread -- then relerr against the persistent toterr.
// state: [0] toterr, [1] |relerr| of this call, [2] newerr
__global__ void k_geterr(const int* __restrict__ ten_base, int nt, float* state, float rel_tol,
                         int* stop_flag, int* steps_done)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (stop_flag && *stop_flag) return;
    if (steps_done) *steps_done += 1;
    float newerr = 0.0f;
    for (int i = 0; i < nt; i++) newerr = __fadd_rn(newerr, (float)ten_base[i]);
    const float tot = state[0];
    const float rel = __fdiv_rn(__fsub_rn(tot, newerr), tot);
    state[0] = newerr;
    state[1] = fabsf(rel);
    state[2] = newerr;
    if (stop_flag && fabsf(rel) < rel_tol) *stop_flag = 1;
}

}  // namespace tpose_dev
