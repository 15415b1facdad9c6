// eval.cuh -- the evaluation kernels of one optimisation step: for every triangle, the six pixel
// moments of all 13 finite-difference variants from edge integrals over the row-prefix moment
// table; then, as the flags ask, cn / ca / ten in the reference SSBO layout, the gradient scatter
// and the vertex update.  Two launches replace the two 13*NT-instance draws plus gradient.cs and
// shift.cs (software/triangulate/main.cpp:121-155; shaders triangle.vs/.fs, gradient.cs, shift.cs):
//
//   k_edges  one work item per (unique mesh edge, move q): q = 0 the edge as it is, q = 1..4 its
//            first end point moved by (+dp,0),(-dp,0),(0,+dp),(0,-dp), q = 5..8 its second one.
//            An interior edge is shared by two triangles and its nine integrals serve both (with
//            opposite sign), so the table is walked once per edge, not once per triangle side.
//            Results go to a 32-byte record per item (48 where the colour sums need 64 bits).
//   k_corners one thread per (triangle, variant k): three records -> moments -> cn/ca/ten.  The
//            moved variants are grouped by the vertex they move (one warp per vertex), so
//            gradient.cs becomes a warp reduction and shift.cs follows in the same warp.
//
// A step can also carry the host-side checks of the PREVIOUS frame (EV_CHECK_PREV), so that a run
// of frames needs no host round trip and no serial check between frames:
//   * geterr (source/triangulation.hpp:653-674) over the previous frame's base energies, by k_edges' service
//     block -- the in-order float32 sum, as a plain chain (geterr_warp) or block-parallel (geterr_block) --
//     overlapped with the edge walks;
//   * a conservative test of the per-frame topology edits of software/triangulate/main.cpp:316-346
//     (prune / >0.8*pi flip / short-edge collapse) on the host's ("watch") topology at the current
//     points, by the service block as well.
// If either fires, k_corners does NOT apply this step's shift and raises the stop
// flag: points, and the other buffer set, are then exactly what the reference holds after the
// frame whose checks fired.  The evaluation done by this step is speculative and discarded.
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
    EV_CHECK_PREV  = 64u,  // previous frame's geterr + topology watch ride on this step
    EV_FINAL       = 128u, // last step of a run: checks only, never shifts
    EV_HONOUR_STOP = 256u, // step is a no-op once the job's stop flag is up
    EV_HISTORY     = 512u  // record every frame (points, base energies, toterr, watch mask); flips alone do not stop the run
};

// trigger bits (mirrors include/tpose_cuda.h)
constexpr int kTrigConverged = 1, kTrigMaxSteps = 2, kTrigTopology = 4;

// run_state words
enum { RS_STOP = 0, RS_STEPS = 1, RS_TRIG = 2, RS_SPARE = 3 };

constexpr int kEdgeMoves = 9;        // per unique edge: as it is + 4 moves of each end point

// signed integral of one (edge, move) item, in the direction first -> second end point
// (six 64-bit words: the tile path adds partial records into it with 64-bit bulk reductions)
struct __align__(16) EdgeRec {
    long long r, g, b, s2;
    long long n, odd;
};

// The same record where 255*W*H < 2^31 (every image up to 3840x2160): no sum of the path leaves 32 bits except
// the squared one, so k_corners adds half the words.  Used by k_edges<G, unsigned> -> k_corners<EdgeRecNarrow>.
struct __align__(16) EdgeRecNarrow {
    int r, g, b, n;
    long long s2;
    int odd, pad;
};

// Everything one image-step needs; lives in device memory so that a batch of independent jobs can
// share one launch (blockIdx.y = job).
struct EvalJob {
    const uint4* table;        // row-prefix moments, pitch entries per row
    int pitch, W, H;
    float2* pts;
    const int4* tri;
    int NT, NP, NE;
    int flip;                  // 1: this job's buffer sets are swapped relative to the launch's `set` argument
    float ratio, dp, step;
    int* cn[2]; int4* ca[2]; int* ten[2];   // two buffer sets, alternating per step
    int2* gr;                  // pgradbuf as the reference sees it (written by the vertex update)
    const int2* edges;         // NE unique edges (first, second end point), derived from tri at upload
    const int4* tri_edges;     // per triangle side j: edge id, bit 31 set if the side runs second -> first
    const int* corner_off;     // NP+1 offsets into `corners`: the triangle corners that sit on each vertex
    const int* corners;        // 3*NT entries, triangle << 2 | corner, grouped by vertex
    EdgeRec* rec;              // NE*9 records
    // tile path (tiles.cuh)
    const void* tmap;          // CUtensorMap of the raw image (device memory), box 64 x TH
    const uint4* coarse;       // prefix at every tile column's left border, column-major [tx][row]
    int ntx, nty;              // tiles per row / column
    int2* esnap;               // NE x 10 snapped end points of the nine moves
    unsigned* ebox;            // NE packed tile ranges (tx0 | tx1 << 8 | ty0 << 16 | ty1 << 24)
    unsigned* gbox;            // the same, reduced over groups of 32 edges
    const int4* wtri;          // watch topology (the host's triangles / halfedges), may be null
    const int* whe;
    unsigned* wmask;           // one bit per watched triangle: set where a per-frame edit may apply
    // per-frame history of a run (EV_HISTORY), frame f of the run at [f * NP], [f * NT], [f], [f * words]
    float2* hist_pts;          // the points frame f left
    int* hist_ten;             // the base energies frame f evaluated (what the reference reads back as terr)
    float* hist_err;           // toterr after frame f's geterr
    unsigned* hist_mask;       // the watch mask at the points frame f left
    int2* hist_gr;             // pgradbuf as frame f left it (stays on the device; see tpose_cuda_history_restore_gradient)
    float* err_state;          // [0] toterr [1] |relerr| [2] newerr [3] converged flag of the last check
    int* run_state;            // RS_*
    int4* geterr_scratch;      // kGeterrScratchBytes (geterr_block)
    unsigned* trig;            // [0] topology trigger accumulator
};

// What a single-job launch can be told as kernel arguments because it never changes for a context (a captured graph
// freezes its arguments; everything that does change lives in the job record): the run's stop flag, and the arrays
// whose first element a thread needs is known from its block and thread index alone.  Those loads are then issued
// BEFORE the job record is in, so the record's round trip and the first link of the dependent chain (edge ->
// end points -> table; corner range -> corners -> ...) overlap.  All null for a batch.
struct Early {
    const int* stop;           // RS_STOP of a run, or null
    const int2* edges; int edges_cap;
    const int* corner_off; const float2* pts; int np_cap;
};

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

// 0: nothing can apply; 1: only the 0.8*pi flip can (it never changes what is uploaded, so it needs
// no stop by itself); 2: prune or collapse can -- or the watch data are inconsistent
__device__ int watch_triangle(const EvalJob& J, int t)
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
    if (fminf(l0, fminf(l1, l2)) <= 0.010001f) return 2;
    // (a) prune (triangulation.hpp:220-241): all corners on the boundary, an open side, every
    //     angle outside (0, pi): acos argument not strictly inside (-1, 1)
    if (on_boundary(p[0], J.ratio) && on_boundary(p[1], J.ratio) && on_boundary(p[2], J.ratio) &&
        (he[0] < 0 || he[1] < 0 || he[2] < 0)) {
        bool flat = true;
#pragma unroll
        for (int k = 0; k < 3; k++) if (fabsf(cs[k]) < 0.99999f) flat = false;   // NaN compares false -> flat
        if (flat) return 2;
    }
    // (b) angle > 0.8*pi (main.cpp:325-334) with a flippable edge: twin exists and the quad is convex
#pragma unroll
    for (int k = 0; k < 3; k++) {
        if (!(cs[k] < -0.8089f)) continue;                      // cos(0.8 pi) = -0.80902
        const int hb = he[k];
        if (hb < 0) continue;
        const int tb = hb / 3, kb = hb - 3 * tb;
        if (tb >= J.NT) return 2;                               // inconsistent watch data: let the host look
        const int4 tvb = J.wtri[tb];
        const float2 A = p[k], C = p[(k + 2) % 3];
        const float2 B = J.pts[corner(tvb, kb)], D = J.pts[corner(tvb, (kb + 2) % 3)];
        if (ccw(A, C, D) == ccw(B, C, D) || ccw(A, B, C) == ccw(A, B, D)) continue;
        return 1;
    }
    return 0;
}

// geterr (source/triangulation.hpp:653-674): float32 running sum of the base energies in index order.  Order
// matters (float addition is not associative), so taken literally it is one dependent chain of NT additions: 4 ns
// each on this part however the addends are fed (one thread chasing global loads: 28 us at 3 000 triangles; one warp
// fed by shuffles: 12 us, 41 us at 10 000) -- and k_edges cannot end before its service block does.  What is serial
// in it is much less than that (geterr_block):
//
// With non-negative addends the running sum s never decreases.  While s < 2^24 every addition is exact.  Inside a
// binade [2^e, 2^(e+1)), e >= 24, s is a multiple of U = 2^(e-23) and adding a = q U + r (0 <= r < U) gives
// s + U (q + c): c = 0 if r < U/2, 1 if r > U/2, and on an exact tie (r = U/2, round to even) c = parity of (s/U + q).
// So as long as the sum stays inside one binade, a run of additions is a function of ONE bit of s -- the parity of
// s/U, which is the last mantissa bit of the float -- that returns an increment in units of U: (inc_0, inc_1).  Two
// such runs compose: inc_p = A.inc_p + B.inc_[(p + A.inc_p) & 1].
//   1. every thread takes a contiguous chunk, sums it exactly (int64) and a block scan gives the exact prefix P in
//      front of the chunk; the binade of P is the thread's GUESS for the binade the float sum is in there;
//   2. the thread folds its chunk into (inc_0, inc_1) under that guess and leaves it in the job's global scratch;
//   3. warp 0 walks with the real float sum (the same in every lane): lane g holds group g -- eight chunk summaries and
//      their fold -- and hands them round with shuffles.  A summary is used only if the sum is in the guessed binade
//      before it and still in it after it (monotone: then it was there all the way, and the increment is exact);
//      otherwise the group's eight chunk summaries are tried the same way, and a chunk whose summary fails (it
//      straddles a power of two, holds a negative energy, ...) is added the plain way, one FADD per element.
// Nothing is assumed about how well the guess fits: a wrong guess costs time, never a bit.  About one group in four
// and seven chunks fall back on a 1.5e9 total (seven binades above 2^24).  Model and randomized check against the
// serial float32 loop: tests/geterr_model.py + tests/test_geterr_cpu.py; on the device: tests/test_step_gpu.py.
constexpr int kGeterrThreads = 256, kGeterrGroup = 8;         // 32 groups: one per lane of the warp that walks
// scratch of geterr_block in GLOBAL memory: one int4 (inc_0, inc_1, guess, -) per thread
constexpr int kGeterrScratchBytes = kGeterrThreads * 16;
// Below this many energies a run takes the plain chain.  Kernel alone the routines cross near 1 500 elements (chain
// 8.1 us + 4.4 ns per element, block routine ~12 us + 0.7 ns); inside k_edges, where the service block shares its SM
// with walkers, a frame at 3 000 triangles costs the same either way (21.4 / 21.9 us) and the block routine wins from
// there on (10 000: 26 against 55 us).
constexpr int kGeterrBlockMin = 3072;

// the verdict of geterr from the new sum
__device__ __forceinline__ void geterr_finish(float newerr, float* state, double rel_tol)
{
    const float tot = state[0];
    const float rel = __fdiv_rn(__fsub_rn(tot, newerr), tot);
    state[0] = newerr;
    state[1] = fabsf(rel);
    state[2] = newerr;
    // the reference compares the float against a double literal (`geterr(&tr) < 1E-4`)
    state[3] = ((double)fabsf(rel) < rel_tol) ? 1.0f : 0.0f;
}

// The literal chain by one warp: every lane loads its share of the next 256 energies (coalesced, converted to float,
// issued one round ahead of their use), and all lanes add the same 256 values in index order, fetching each from the
// lane that holds it with a shuffle.  Elements past nt are +0.0f: s + 0 = s exactly (s is never -0).
// No shared memory on purpose (geterr_block neither): a few KB of it in k_edges make every SM re-split its L1 / shared
// memory between k_edges and k_corners -- measured, +1.7 us on the cold k_edges and 1.5 % on the warm step.
__device__ void geterr_warp(const int* __restrict__ ten_base, int nt, float* state, double rel_tol, int lane)
{
    constexpr int K = 8;                       // registers per lane and round: 32 * K energies per round
    float cur[K], nxt[K];
    auto load = [&](int base, float (&r)[K]) {
#pragma unroll
        for (int j = 0; j < K; j++) {
            const int i = base + 32 * j + lane;
            r[j] = i < nt ? (float)__ldcg(ten_base + i) : 0.0f;
        }
    };
    float newerr = 0.0f;
    load(0, cur);
    for (int base = 0; base < nt; base += 32 * K) {
        if (base + 32 * K < nt) load(base + 32 * K, nxt);
#pragma unroll
        for (int j = 0; j < K; j++)
#pragma unroll
            for (int l = 0; l < 32; l++) newerr = __fadd_rn(newerr, __shfl_sync(0xffffffffu, cur[j], l));
#pragma unroll
        for (int j = 0; j < K; j++) cur[j] = nxt[j];
    }
    if (lane == 0) geterr_finish(newerr, state, rel_tol);
}

// a summary as the walk wants it: the two increments as floats (exact: < 2^24 units of a power of two), the binade's
// lower bound (0 below 2^24, where every sum is exact) and the bound the sum must stay under; lim = 0: no summary
__device__ __forceinline__ float4 geterr_pack(unsigned long long inc0, unsigned long long inc1, int e)
{
    if (e < 0 || inc0 >= (1ull << 24) || inc1 >= (1ull << 24)) return make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    const int sh = max(e - 23, 0);
    const float unit = __uint_as_float((unsigned)(127 + sh) << 23);
    return make_float4(__fmul_rn((float)(unsigned)inc0, unit), __fmul_rn((float)(unsigned)inc1, unit),
                       sh > 0 ? __uint_as_float((unsigned)(127 + e) << 23) : 0.0f, __uint_as_float((unsigned)(127 + max(e, 23) + 1) << 23));
}
// s + increment if s is in the summary's binade before and after
__device__ __forceinline__ bool geterr_try(float& s, const float4 S)
{
    const float o0 = __fadd_rn(s, S.x), o1 = __fadd_rn(s, S.y);       // (both, so that the parity test is off the chain)
    const float out = (__float_as_uint(s) & 1u) ? o1 : o0;
    if (!(s >= S.z) || !(out < S.w)) return false;
    s = out;
    return true;
}

__device__ __forceinline__ float4 shfl4(const float4 v, int src)
{
    return make_float4(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src), __shfl_sync(0xffffffffu, v.z, src), __shfl_sync(0xffffffffu, v.w, src));
}

// `scratch`: kGeterrScratchBytes of global memory of this job; nthreads = kGeterrThreads
__device__ void geterr_block(const int* __restrict__ ten_base, int nt, float* state, double rel_tol, int tid, int nthreads, int4* scratch)
{
    __shared__ unsigned long long s_wsum[kGeterrThreads / 32];
    const int T = nthreads;
    const int C = 4 * ((nt + 4 * T - 1) / (4 * T));            // chunk length, a multiple of 4: 16-byte loads
    const int lane = tid & 31, wid = tid >> 5;
    const int i0 = min(tid * C, nt), i1 = min(i0 + C, nt);
    // four energies as the exact integer values of their floats; past the end: 0
    auto load4 = [&](int i, unsigned long long (&a)[4], bool& neg) {
        int4 v = make_int4(0, 0, 0, 0);
        if (i + 4 <= nt) v = __ldg(reinterpret_cast<const int4*>(ten_base + i));
        else { if (i < nt) v.x = __ldg(ten_base + i); if (i + 1 < nt) v.y = __ldg(ten_base + i + 1); if (i + 2 < nt) v.z = __ldg(ten_base + i + 2); }
        neg |= (v.x | v.y | v.z | v.w) < 0;
        a[0] = (unsigned long long)(long long)(float)max(v.x, 0); a[1] = (unsigned long long)(long long)(float)max(v.y, 0);
        a[2] = (unsigned long long)(long long)(float)max(v.z, 0); a[3] = (unsigned long long)(long long)(float)max(v.w, 0);
    };
    // 1. exact chunk sum of the float values of the energies, and the exact prefix in front of the chunk
    unsigned long long sum = 0ull; bool neg = false;
    for (int i = i0; i < i1; i += 4) {
        unsigned long long a[4];
        load4(i, a, neg);
        sum += (a[0] + a[1]) + (a[2] + a[3]);
    }
    unsigned long long incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    if (lane == 31) s_wsum[wid] = incl;
    __syncthreads();
    unsigned long long P0 = incl - sum;
    for (int w = 0; w < wid; w++) P0 += s_wsum[w];
    // 2. the chunk as a function of the parity of s / U, under the binade of P0 (e: the guess; -1: no summary; -2: empty chunk)
    {
        int e = -2;
        unsigned long long t0 = 0ull, t1 = 0ull;              // increments in units of U for entry parity 0 / 1
        if (i0 < i1) {
            e = P0 ? 63 - __clzll((long long)P0) : 0;
            const int sh = max(e - 23, 0);
            const unsigned long long mask = (1ull << sh) - 1ull, half = (1ull << sh) >> 1;
            bool dummy = false;
            for (int i = i0; i < i1; i += 4) {
                unsigned long long a[4];
                load4(i, a, dummy);
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const unsigned long long q = a[k] >> sh, r = a[k] & mask;
                    const unsigned long long tie = (sh > 0 && r == half) ? 1ull : 0ull, up = (sh > 0 && r > half) ? 1ull : 0ull;
                    t0 += q + up + (tie & (q + t0));           // parity before the add = (entry parity + increments so far) & 1
                    t1 += q + up + (tie & (q + t1 + 1ull));
                }
            }
            if (neg || t0 >= (1ull << 24) || t1 >= (1ull << 24)) e = -1;
        }
        scratch[tid] = make_int4((int)(unsigned)t0, (int)(unsigned)t1, e, 0);
    }
    __syncthreads();                                           // (orders the block's global writes before the reads below)
    if (wid != 0) return;
    // 3. warp 0 walks: lane g holds group g -- its eight chunk summaries and their fold -- and every lane carries the same sum
    float4 S[kGeterrGroup], G;
    {
        int ge = -2; unsigned long long g0 = 0ull, g1 = 0ull;
#pragma unroll
        for (int k = 0; k < kGeterrGroup; k++) {
            const int4 c = scratch[lane * kGeterrGroup + k];
            const unsigned a0 = (unsigned)c.x, a1 = (unsigned)c.y; const int ek = c.z;
            S[k] = ek == -2 ? make_float4(0.0f, 0.0f, 0.0f, 1e38f) : geterr_pack(a0, a1, ek);        // (empty: + 0, always fine)
            if (ek == -2) continue;
            if (ek < 0 || ge == -1 || (ge >= 0 && ek != ge)) { ge = -1; continue; }
            ge = ek;
            g0 += (g0 & 1ull) ? a1 : a0;                       // parity in front of the chunk = (entry parity + increments so far) & 1
            g1 += ((g1 + 1ull) & 1ull) ? a1 : a0;
        }
        G = ge == -2 ? make_float4(0.0f, 0.0f, 0.0f, 1e38f) : geterr_pack(g0, g1, ge);
    }
    float newerr = 0.0f;
    for (int g = 0; g < 32; g++) {
        if (geterr_try(newerr, shfl4(G, g))) continue;
#pragma unroll
        for (int k = 0; k < kGeterrGroup; k++) {
            if (geterr_try(newerr, shfl4(S[k], g))) continue;
            const int t = g * kGeterrGroup + k;
            const int j0 = min(t * C, nt), j1 = min(j0 + C, nt);
            for (int i = j0; i < j1; i++) newerr = __fadd_rn(newerr, (float)__ldg(ten_base + i));
        }
    }
    if (lane == 0) geterr_finish(newerr, state, rel_tol);
}

// geterr over an array of energies on its own (tpose_cuda_geterr): state = err_state
// (`by_block`: 1 the block routine, 0 the warp chain, -1 whichever a run would take for this n)
__global__ void __launch_bounds__(kGeterrThreads) k_geterr(const int* __restrict__ energies, int n, float* state, double rel_tol, int by_block, int4* scratch)
{
    if (by_block < 0) by_block = n >= kGeterrBlockMin;
    if (by_block) geterr_block(energies, n, state, rel_tol, threadIdx.x, kGeterrThreads, scratch);
    else if (threadIdx.x < 32) geterr_warp(energies, n, state, rel_tol, threadIdx.x);
}

// ---- the service block (shared by k_edges and k_setup) ---------------------------------------------
// Pulls the static index arrays k_corners is about to chase through into L2, then runs the previous
// frame's host-side checks: geterr (geterr_block) and the topology watch.
__device__ void service_block(EvalJob& J, int set, unsigned flags, double rel_tol, int frame, int tid, int nthreads, bool single_job)
{
    if (!(flags & EV_FINAL) && single_job) {
        auto prefetch = [&](const void* base, size_t bytes) {
            const char* p = reinterpret_cast<const char*>(base);
            for (size_t o = (size_t)tid * 128; o < bytes; o += (size_t)nthreads * 128)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p + o));
        };
        prefetch(J.corner_off, sizeof(int) * ((size_t)J.NP + 1));
        prefetch(J.corners, sizeof(int) * 3 * (size_t)J.NT);
        prefetch(J.tri_edges, sizeof(int4) * (size_t)J.NT);
    }
    if (!(flags & EV_CHECK_PREV)) return;
    if (tid == 0) *J.trig = 0u;
    const int words = (J.NT + 31) / 32;
    const bool hist = (flags & EV_HISTORY) && frame > 0;          // this evaluation checks frame - 1
    unsigned* const mask = hist ? J.hist_mask + (size_t)(frame - 1) * words : J.wmask;
    if (J.wtri || hist) for (int w = tid; w < words; w += nthreads) mask[w] = 0u;   // (no watch topology: an empty mask)
    __syncthreads();
    // geterr.  Small meshes: the last warp runs the plain chain while the others watch.  Larger ones: the parallel part
    // by everybody, then thread 0 walks the summaries while the others start on the watch.
    const bool by_block = J.NT >= kGeterrBlockMin;
    if (by_block) geterr_block(J.ten[set ^ 1], J.NT, J.err_state, rel_tol, tid, nthreads, J.geterr_scratch);
    else if (tid >= nthreads - 32) geterr_warp(J.ten[set ^ 1], J.NT, J.err_state, rel_tol, tid & 31);
    if (hist && tid == (by_block ? 0 : nthreads - 32)) J.hist_err[frame - 1] = J.err_state[0];
    if (J.wtri && (by_block || tid < nthreads - 32)) {
        const int watchers = by_block ? nthreads : nthreads - 32;
        bool hit = false;
        for (int t = tid; t < J.NT; t += watchers) {
            const int code = watch_triangle(J, t);
            if (code) atomicOr(&mask[t >> 5], 1u << (t & 31));
            hit |= (flags & EV_HISTORY) ? code == 2 : code != 0;
        }
        if (hit) atomicOr(J.trig, (unsigned)kTrigTopology);
    }
}

// ---- k_edges ----------------------------------------------------------------------------------

constexpr int kEdgeThreads = 256;

// Development aid (-DTPOSE_TIMELINE): first / last globaltimer stamp over all threads at a few
// points of the two kernels, read back with tpose_cuda_debug_timeline.
#ifdef TPOSE_TIMELINE
constexpr int kTlWarps = 8192;
__device__ unsigned long long g_timeline[2][6][kTlWarps];
__device__ __forceinline__ void tl_mark(int kernel, int phase)
{
    __syncwarp();
    if ((threadIdx.x & 31) == 0) {
        unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
        if (w < kTlWarps) g_timeline[kernel][phase][w] = t;
    }
}
#define TL(k, p) tl_mark(k, p)
#else
#define TL(k, p)
#endif

template <typename T> __device__ __forceinline__ T shfl_xor_any(T v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }

// Walks the rows of one edge (this lane's share: rows r, r+G, ...) and sums the table entries at
// the crossings.  ACC is the accumulator type of the three colour sums: 32 bits are enough when
// 255*W*H < 2^31 (every image up to 3840x2160), else 64.
// Rows are taken BATCH at a time: all loads of a batch are issued before the first one is
// consumed, so a lane pays one memory latency per batch, not per row; whole batches first, then
// one tail batch whose slots past the end of the edge read a zero entry.
// k_edges exists in two shapes (capi.cu picks per launch):
//   wide   batches of 8, 61-64 registers, 4 blocks per SM -- the default;
//   dense  batches of 4, 48 registers, 5 blocks per SM: 740 resident blocks instead of 592.  For a single image whose
//          items need more than 592 blocks but not more than 740 (1080p / 3 000 triangles: 641) the wide shape
//          sends 49 blocks through the whole dependent chain a second time while the machine idles; the dense one
//          takes everything in one pass (cold k_edges 16.4 -> 14.4 us, cold step +11 %, measured).
constexpr int kWalkBatchWide = 8, kWalkBatchDense = 4;
constexpr int kEdgeBlocksWide = 4, kEdgeBlocksDense = 5;

// A table entry through the read-only path with the L2 fetch size pinned to 64 bytes (SASS: LDG.E.LTC64B.128.CONSTANT).
// A plain load that misses L2 makes it fetch the whole 128-byte line from DRAM (measured on B200 with a strided
// probe: 128 B of dram__bytes_read per 16-byte load at any stride >= 128, every load flavour; 64 B with this
// qualifier or with a bulk copy; cudaLimitMaxL2FetchGranularity changes nothing).  The walk is a gather -- the next
// row's crossing is in another line -- so the second half of the line is bytes nobody asked for.
__device__ __forceinline__ uint4 ldg_entry(const uint4* p)
{
    uint4 v;
    asm("ld.global.nc.L2::64B.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// one batch of BATCH rows.  TAIL: the edge may end inside the batch -- a slot past its end reads entry 0
// (column 0 of row 0: an all-zero prefix), an unconditional load of zeros instead of a predicated load into
// zeroed registers; full batches carry no per-slot test at all.
template <int G, int BATCH, typename ACC, typename REM, bool TAIL>
__device__ __forceinline__ void walk_batch(const int rr, const int r_end, int& q, const int dq, REM& rem, const REM den, const REM dr,
                                           const uint4* __restrict__ table, int& row, const int rstep, const int W,
                                           unsigned& n, unsigned& odd, ACC& sr, ACC& sg, ACC& sb, unsigned long long& s2)
{
    uint4 e[BATCH];
    int c[BATCH];
#pragma unroll
    for (int u = 0; u < BATCH; u++) {
        const bool on = !TAIL || rr + u * G < r_end;
        const int cc = min(max(q, 0), W);
        c[u] = on ? cc : 0;
        e[u] = ldg_entry(table + (on ? row + cc : 0));      // entry index: (W+1)*H <= 4096*8192 fits 32 bits
        q += dq; rem += dr;
        if (rem >= den) { rem -= den; ++q; }
        row += rstep;
    }
#pragma unroll
    for (int u = 0; u < BATCH; u++) {
        n += (unsigned)c[u]; odd += e[u].x >> 20; sr += (ACC)(e[u].x & 0xFFFFFu);
        sg += (ACC)e[u].y; sb += (ACC)e[u].z; s2 += e[u].w;
    }
}

template <int G, int BATCH, typename ACC, typename REM>
__device__ __forceinline__ void walk_rows(int rr, const int r_end, int q, const int dq, REM rem, const REM den, const REM dr,
                                          const uint4* __restrict__ table, int row, const int rstep, const int W,
                                          unsigned& n, unsigned& odd, ACC& sr, ACC& sg, ACC& sb, unsigned long long& s2)
{
    for (; rr + (BATCH - 1) * G < r_end; rr += BATCH * G)
        walk_batch<G, BATCH, ACC, REM, false>(rr, r_end, q, dq, rem, den, dr, table, row, rstep, W, n, odd, sr, sg, sb, s2);
    if (rr < r_end)
        walk_batch<G, BATCH, ACC, REM, true>(rr, r_end, q, dq, rem, den, dr, table, row, rstep, W, n, odd, sr, sg, sb, s2);
}

template <int G, int BATCH, typename ACC>
__device__ __forceinline__ void walk_edge(EdgeWalk& w, const uint4* __restrict__ table, int pitch, int W,
                                          unsigned& n, unsigned& odd, ACC& sr, ACC& sg, ACC& sb, unsigned long long& s2)
{
    if (w.den < (1ll << 30))                   // every on-screen edge: remainders fit 32 bits
        walk_rows<G, BATCH, ACC, int>(w.r, w.r_end, w.q, w.dq, (int)w.rem, (int)w.den, (int)w.dr, table, w.r * pitch, G * pitch, W,
                               n, odd, sr, sg, sb, s2);
    else
        walk_rows<G, BATCH, ACC, long long>(w.r, w.r_end, w.q, w.dq, w.rem, w.den, w.dr, table, w.r * pitch, G * pitch, W,
                                     n, odd, sr, sg, sb, s2);
}

// G lanes per (edge, move) item.  One block of each job's row is the service block.
template <int G, typename ACC, bool DENSE>
__global__ void __launch_bounds__(kEdgeThreads, DENSE ? kEdgeBlocksDense : kEdgeBlocksWide)
k_edges(const EvalJob* __restrict__ jobs, int set, unsigned flags, double rel_tol, int frame, const Early early)
{
    __shared__ EvalJob J;
    const int tid = threadIdx.x;
    TL(0, 0);
    // a single-job run hands its stop flag over as an argument: the launches behind the stop of a flat chain
    // then cost one load, not a job record and a dependent load
    if (early.stop && __ldcg(early.stop)) return;
    // ... and its first edge can be on its way before the job record is (single job: block 0 is the service block)
    int2 ev_early = make_int2(0, 0);
    const int e_early = (((int)blockIdx.x - 1) * (kEdgeThreads / G) + tid / G) / kEdgeMoves;
    const bool have_early = early.edges && blockIdx.x > 0 && e_early < early.edges_cap;
    if (have_early) ev_early = __ldg(early.edges + e_early);
    if (tid < (int)(sizeof(EvalJob) / sizeof(int)))
        reinterpret_cast<int*>(&J)[tid] = reinterpret_cast<const int*>(jobs + blockIdx.y)[tid];
    __syncthreads();
    TL(0, 1);
    if ((flags & EV_HONOUR_STOP) && __ldcg(&J.run_state[RS_STOP])) return;
    set ^= J.flip;

    // One block of each job's row is the service block.  For a single job it is block 0: the in-order float sum of
    // geterr is the longest dependent chain of the launch, so it is dispatched first and runs under the walkers (as
    // the last block it started when the first walker finished, and the kernel waited for it: 25.8 -> 23.2 us per
    // frame of a run at 1080p / 3000 triangles).  In a batch the other jobs' walkers cover it anyway and first place
    // costs 10-15 % (config 5, measured), so there it stays the last block of its row.
    const int service = gridDim.y == 1 ? 0 : (int)gridDim.x - 1;

    // the last evaluation of a run only carries the checks of the last frame: nothing to integrate
    if ((flags & EV_FINAL) && (int)blockIdx.x != service) return;

    if ((int)blockIdx.x == service) {
        service_block(J, set, flags, rel_tol, frame, tid, kEdgeThreads, gridDim.y == 1);
        return;
    }
    const int worker = (int)blockIdx.x - ((int)blockIdx.x > service), nworkers = (int)gridDim.x - 1;

    constexpr int kItemsPerBlock = kEdgeThreads / G;
    const int lane = tid % G;
    const int nitems = J.NE * kEdgeMoves;
    const Xform xf{J.ratio, 0.5f * (float)J.W, 0.5f * (float)J.H, 256 * J.H};
    // the grid does not depend on the mesh size (so a captured graph survives topology changes):
    // blocks stride over the items
    for (int base = worker * kItemsPerBlock; base < nitems; base += nworkers * kItemsPerBlock) {
        const int item = base + tid / G;
        unsigned n = 0, odd = 0; ACC sr = 0, sg = 0, sb = 0; unsigned long long s2 = 0;
        int sign = 0;
        if (item < nitems) {
            const int e = item / kEdgeMoves, q = item - e * kEdgeMoves;
            const int2 ev = (have_early && e == e_early) ? ev_early : __ldg(J.edges + e);
            const float2 pa = J.pts[ev.x], pb = J.pts[ev.y];
#ifdef TPOSE_TIMELINE
            if (pa.x + pb.x != 12345.0f) TL(0, 2);
#endif
            float dxa = 0.0f, dya = 0.0f, dxb = 0.0f, dyb = 0.0f;
            if (q >= 1 && q <= 4) fd_offset(q - 1, J.dp, dxa, dya);
            if (q >= 5) fd_offset(q - 5, J.dp, dxb, dyb);
            const int2 a = to_fixed(pa, dxa, dya, xf), b = to_fixed(pb, dxb, dyb, xf);
            EdgeWalk w;
            w.init(a, b, J.H, lane, G);
            sign = w.sign;
#ifdef TPOSE_TIMELINE
            if (w.q != 0x7fffffff) TL(0, 3);
#endif
            walk_edge<G, DENSE ? kWalkBatchDense : kWalkBatchWide, ACC>(w, J.table, J.pitch, J.W, n, odd, sr, sg, sb, s2);
#ifdef TPOSE_TIMELINE
            if (n + s2 != 0x7fffffffffull) TL(0, 4);
#endif
        }
#pragma unroll
        for (int o = G >> 1; o > 0; o >>= 1) {
            n += shfl_xor_any(n, o); odd += shfl_xor_any(odd, o);
            sr += shfl_xor_any(sr, o); sg += shfl_xor_any(sg, o); sb += shfl_xor_any(sb, o); s2 += shfl_xor_any(s2, o);
        }
        if (item < nitems && lane == 0) {
            if (sizeof(ACC) == 4) {          // narrow sums: 32-byte records
                EdgeRecNarrow rec;
                rec.r = sign * (int)sr; rec.g = sign * (int)sg; rec.b = sign * (int)sb; rec.n = sign * (int)n;
                rec.s2 = sign * (long long)s2; rec.odd = sign * (int)odd; rec.pad = 0;
                reinterpret_cast<EdgeRecNarrow*>(J.rec)[item] = rec;
            } else {
                EdgeRec rec;
                rec.r = sign * (long long)sr; rec.g = sign * (long long)sg; rec.b = sign * (long long)sb;
                rec.s2 = sign * (long long)s2;
                rec.n = sign * (long long)n; rec.odd = sign * (long long)odd;
                J.rec[item] = rec;
            }
        }
    }
    TL(0, 5);
}

// ---- k_corners ----------------------------------------------------------------------------------
// Variant k = 1+4j+s of triangle t moves corner j, and only the vertex at that corner reads its
// energy (gradient.cs:27-34).  So the 12 moved variants of every triangle are evaluated grouped by
// VERTEX: one warp per vertex walks the vertex's corner list (built at upload), four lanes per
// corner (s = 0..3), and the central differences are summed inside the warp -- no atomics, no
// fences, no inter-block ordering.  The base variants (k = 0) go to the remaining blocks.

constexpr int kCornerThreads = 256;
constexpr int kCornerWarps = kCornerThreads / 32;
constexpr int kCornerBlocksPerSM = 5;    // 48 registers x 256 threads (the wide-record variant holds 4)

// the three sums of a variant: 64-bit for the wide records, wrapping 32-bit for the narrow ones
template <typename REC> struct SumOf { typedef long long type; };
template <> struct SumOf<EdgeRecNarrow> { typedef int type; };

template <typename REC, typename S>
__device__ __forceinline__ void add_rec(const REC* __restrict__ recs, int packed, int qlocal,
                                        S& sr, S& sg, S& sb, long long& s2, int& n, int& odd);

template <>
__device__ __forceinline__ void add_rec<EdgeRec, long long>(const EdgeRec* __restrict__ recs, int packed, int qlocal,
                                                            long long& sr, long long& sg, long long& sb, long long& s2, int& n, int& odd)
{
    // side runs second -> first: its first corner is the edge's second end point, and the sign flips
    const bool rev = packed < 0;
    const int e = packed & 0x7fffffff;
    const int q = (rev && qlocal) ? (qlocal <= 4 ? qlocal + 4 : qlocal - 4) : qlocal;
    const longlong2* p = reinterpret_cast<const longlong2*>(recs + (size_t)e * kEdgeMoves + q);
    const longlong2 v0 = __ldcg(p), v1 = __ldcg(p + 1), v2 = __ldcg(p + 2);
    if (rev) { sr -= v0.x; sg -= v0.y; sb -= v1.x; s2 -= v1.y; n -= (int)v2.x; odd -= (int)v2.y; }
    else     { sr += v0.x; sg += v0.y; sb += v1.x; s2 += v1.y; n += (int)v2.x; odd += (int)v2.y; }
}

template <>
__device__ __forceinline__ void add_rec<EdgeRecNarrow, int>(const EdgeRecNarrow* __restrict__ recs, int packed, int qlocal,
                                                            int& sr, int& sg, int& sb, long long& s2, int& n, int& odd)
{
    const bool rev = packed < 0;
    const int e = packed & 0x7fffffff;
    const int q = (rev && qlocal) ? (qlocal <= 4 ? qlocal + 4 : qlocal - 4) : qlocal;
    const int4* p = reinterpret_cast<const int4*>(recs + (size_t)e * kEdgeMoves + q);
    const int4 v0 = __ldcg(p), v1 = __ldcg(p + 1);
    const long long vs2 = (long long)(((unsigned long long)(unsigned)v1.y << 32) | (unsigned)v1.x);
    if (rev) { sr -= v0.x; sg -= v0.y; sb -= v0.z; n -= v0.w; s2 -= vs2; odd -= v1.z; }
    else     { sr += v0.x; sg += v0.y; sb += v0.z; n += v0.w; s2 += vs2; odd += v1.z; }
}

// floor(a / n) for three non-negative 32-bit numerators and one divisor 1 <= n < 2^31 with a single hardware
// division: M = ceil(2^32 / n), q' = mulhi(a, M) is the quotient or one more (a * (M n - 2^32) < 2^32 n), fixed by
// one multiply-compare.  The reference divides ivec3 by int (triangle.fs:40).
__device__ __forceinline__ void div3_nonneg(unsigned a0, unsigned a1, unsigned a2, unsigned n, int& q0, int& q1, int& q2)
{
    if (n == 1u) { q0 = (int)a0; q1 = (int)a1; q2 = (int)a2; return; }
    const unsigned M = 0xFFFFFFFFu / n + 1u;                 // ceil(2^32 / n) for n >= 2 (2^32 itself is never a multiple's successor here)
    unsigned t0 = __umulhi(a0, M), t1 = __umulhi(a1, M), t2 = __umulhi(a2, M);
    if ((unsigned long long)t0 * n > a0) --t0;
    if ((unsigned long long)t1 * n > a1) --t1;
    if ((unsigned long long)t2 * n > a2) --t2;
    q0 = (int)t0; q1 = (int)t1; q2 = (int)t2;
}

// variant k of triangle t: three edge records -> moments -> cn / ca / ten as the flags ask.
// Returns the (wrapped int32) energy.
template <typename REC>
__device__ __forceinline__ int eval_variant(const EvalJob& J, unsigned flags, int* __restrict__ cn, int4* __restrict__ ca,
                                            int* __restrict__ ten, int t, int k)
{
    typedef typename SumOf<REC>::type S;
    const REC* recs = reinterpret_cast<const REC*>(J.rec);
    const int4 te = __ldg(J.tri_edges + t);
    // side j has its first corner moved, side j+2 its second one; the third side is as it is
    const int j = (k - 1) >> 2, s = (k - 1) & 3, j2 = (j >= 1) ? j - 1 : 2;
    const int q0 = (k == 0) ? 0 : (j == 0 ? 1 + s : (j2 == 0 ? 5 + s : 0));
    const int q1 = (k == 0) ? 0 : (j == 1 ? 1 + s : (j2 == 1 ? 5 + s : 0));
    const int q2 = (k == 0) ? 0 : (j == 2 ? 1 + s : (j2 == 2 ? 5 + s : 0));
    S sr = 0, sg = 0, sb = 0; long long s2 = 0; int n = 0, odd = 0;
    add_rec<REC, S>(recs, te.x, q0, sr, sg, sb, s2, n, odd);
    add_rec<REC, S>(recs, te.y, q1, sr, sg, sb, s2, n, odd);
    add_rec<REC, S>(recs, te.z, q2, sr, sg, sb, s2, n, odd);
    if (n < 0) { n = -n; odd = -odd; sr = -sr; sg = -sg; sb = -sb; s2 = -s2; }
    const int g = k * J.NT + t;
    if (flags & EV_WRITE_CN) cn[g] = n;
    if (flags & EV_WRITE_CA) ca[g] = make_int4((int)sr, (int)sg, (int)sb, 0);
    int e32 = 0;
    if (flags & EV_WRITE_TEN) {
        int m0 = 0, m1 = 0, m2 = 0; bool have = true;
        if (flags & EV_MEAN_STORED) {
            const int4 c = ca[g];
            if (flags & EV_MEAN_DIVIDE) {
                const int cnt = cn[g];
                if (cnt > 0) { m0 = c.x / cnt; m1 = c.y / cnt; m2 = c.z / cnt; } else have = false;
            } else { m0 = c.x; m1 = c.y; m2 = c.z; }
        } else {
            // |sums| <= 255*n < 2^31 here, so the quotient is a 32-bit one unless an accumulator
            // would have wrapped -- then take the general path
            if (n > 0) {
                if (sizeof(S) == 4 || (sr == (int)sr && sg == (int)sg && sb == (int)sb)) {
                    if ((int)sr >= 0 && (int)sg >= 0 && (int)sb >= 0) div3_nonneg((unsigned)(int)sr, (unsigned)(int)sg, (unsigned)(int)sb, (unsigned)n, m0, m1, m2);
                    else { m0 = (int)sr / n; m1 = (int)sg / n; m2 = (int)sb / n; }
                } else { m0 = (int)((long long)sr / n); m1 = (int)((long long)sg / n); m2 = (int)((long long)sb / n); }
            } else have = false;
        }
        const long long e = have ? energy_closed_form(n, odd, (long long)sr, (long long)sg, (long long)sb, s2, m0, m1, m2) : 0ll;
        e32 = (int)(unsigned)(unsigned long long)e;   // int32 accumulator wraps
        ten[g] = e32;
    }
    return e32;
}

template <typename REC>
__global__ void __launch_bounds__(kCornerThreads)
k_corners(const EvalJob* __restrict__ jobs, int set, unsigned flags, int frame, int base_blocks, const Early early)
{
    __shared__ EvalJob J;
    const int tid = threadIdx.x;
    TL(1, 0);
    if (early.stop && __ldcg(early.stop)) return;     // (see k_edges; a block that sees the stop THIS launch raised skips work that is discarded anyway)
    // the corner range and the position of this warp's first vertex, requested before the job record arrives
    const int v_early = (int)blockIdx.x * kCornerWarps + (tid >> 5);
    const bool have_early = early.corner_off && v_early < early.np_cap;
    int c0_early = 0, c1_early = 0; float2 p_early = make_float2(0.0f, 0.0f);
    if (have_early) {
        c0_early = __ldg(early.corner_off + v_early); c1_early = __ldg(early.corner_off + v_early + 1);
        if ((tid & 31) == 0) p_early = early.pts[v_early];
    }
    if (tid < (int)(sizeof(EvalJob) / sizeof(int)))
        reinterpret_cast<int*>(&J)[tid] = reinterpret_cast<const int*>(jobs + blockIdx.y)[tid];
    __syncthreads();
    TL(1, 1);
    // RS_STOP is raised by this kernel's first block only when this very launch decided to stop;
    // a block that sees it then skips work that is discarded anyway
    if ((flags & EV_HONOUR_STOP) && __ldcg(&J.run_state[RS_STOP])) return;
    set ^= J.flip;
    int* const cn = J.cn[set]; int4* const ca = J.ca[set]; int* const ten = J.ten[set];

    const int nvb = (int)gridDim.x - base_blocks;           // blocks that take vertices; the last base_blocks take the base variants
    if ((flags & EV_FINAL) && blockIdx.x != 0) return;      // check-only evaluation: block 0 records the verdict
    if ((int)blockIdx.x >= nvb) {
        // base variants, one thread each
        const bool hist = (flags & EV_HISTORY) && frame >= 0 && !(flags & EV_FINAL);
        for (int t = ((int)blockIdx.x - nvb) * kCornerThreads + tid; t < J.NT; t += base_blocks * kCornerThreads) {
            const int e = eval_variant<REC>(J, flags, cn, ca, ten, t, 0);
            if (hist) J.hist_ten[(size_t)frame * J.NT + t] = e;       // harmless if this frame turns out to be discarded
        }
        TL(1, 2);
        return;
    }

    // the previous frame's verdict (written by k_edges' service block of this step)
    bool stop = false, conv = false; unsigned topo = 0u;
    if (flags & EV_CHECK_PREV) {
        conv = __ldcg(&J.err_state[3]) != 0.0f;
        topo = __ldcg(J.trig);
        stop = conv || topo != 0u || (flags & EV_FINAL);
    }
    if (blockIdx.x == 0 && tid == 0 && (flags & EV_GRAD_SHIFT)) {
        if (stop) {
            J.run_state[RS_STOP] = 1;
            J.run_state[RS_TRIG] = (conv ? kTrigConverged : 0) | (topo ? kTrigTopology : 0) |
                                   ((!conv && !topo) ? kTrigMaxSteps : 0);
        } else {
            J.run_state[RS_STEPS] += 1;
        }
    }

    if (flags & EV_FINAL) return;
    const int lane = tid & 31, c_in = lane >> 2, s = lane & 3;
    for (int v = (int)blockIdx.x * kCornerWarps + (tid >> 5); v < J.NP; v += nvb * kCornerWarps) {
        const bool pre = have_early && v == v_early;
        const int c0 = pre ? c0_early : __ldg(J.corner_off + v), c1 = pre ? c1_early : __ldg(J.corner_off + v + 1);
        float2 p = make_float2(0.0f, 0.0f);
        if (lane == 0) p = pre ? p_early : J.pts[v];
        unsigned gx = 0u, gy = 0u;
        for (int cb = c0; cb < c1; cb += 8) {
            const int c = cb + c_in;
            if (c < c1) {
                const int packed = __ldg(J.corners + c);          // triangle << 2 | corner
                const int t = packed >> 2, j = packed & 3;
                const unsigned e = (unsigned)eval_variant<REC>(J, flags, cn, ca, ten, t, 1 + 4 * j + s);
                // gradient.cs:27-34: x from the +-dp moves in x (s = 0, 1), y from s = 2, 3
                const unsigned signed_e = (s & 1) ? 0u - e : e;
                if (s < 2) gx += signed_e; else gy += signed_e;
            }
        }
        if (!(flags & EV_GRAD_SHIFT)) continue;
#ifdef TPOSE_TIMELINE
        if (gx != 0x12345u) TL(1, 3);
#endif
        gx = __reduce_add_sync(0xffffffffu, gx);
        gy = __reduce_add_sync(0xffffffffu, gy);
        if (lane == 0 && !stop) {
            // A vertex no triangle references is never zeroed by the vertex stage (triangle.vs:93-98 resets
            // gr only for the three vertices of each instance) and gradient.cs adds nothing to it: its
            // entry stays what an earlier frame left there, and shift.cs:16-47 goes on applying it.
            if (c0 == c1) { const int2 stale = J.gr[v]; gx = (unsigned)stale.x; gy = (unsigned)stale.y; }
            else J.gr[v] = make_int2((int)gx, (int)gy);
            if (v >= 4) { p = shift_vertex(p, (int)gx, (int)gy, J.ratio, J.step); J.pts[v] = p; }
            if ((flags & EV_HISTORY) && frame >= 0) {
                J.hist_pts[(size_t)frame * J.NP + v] = p;
                J.hist_gr[(size_t)frame * J.NP + v] = make_int2((int)gx, (int)gy);
            }
        }
    }
    TL(1, 4);
}

// ---- k_eval_list -----------------------------------------------------------------------------------
// A handful of instances evaluated on their own: entry i = (v0, v1, v2, k) -- three vertex indices into the
// context's current points and a variant number k (0 = as it is, 1 + 4j + s = corner j moved, triangle.vs:66-78)
// -- gives cn, ca and the mode-1 energy of that one instance with its own mean colour (triangle.fs:27-43,
// triangulate flavour).  One warp per entry: the three sides are walked by all 32 lanes (rows interleaved)
// over the row-prefix table, exactly like k_edges walks an edge.  This is what the flip set of a converged
// frame needs (software/triangulate/main.cpp:287-306 re-rasterises all 13*NT instances twice to read the
// base energies of the flipped pairs, then of all triangles): the instances are independent, so evaluating
// just those gives the same numbers.
constexpr int kListThreads = 256;

template <typename ACC>
__global__ void __launch_bounds__(kListThreads)
k_eval_list(const uint4* __restrict__ table, int pitch, int W, int H, const float2* __restrict__ pts, float ratio, float dp,
            const int4* __restrict__ entries, int n, int* __restrict__ ten, int* __restrict__ cn, int4* __restrict__ ca)
{
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (kListThreads / 32) + (threadIdx.x >> 5);
    if (i >= n) return;
    const int4 en = __ldg(entries + i);
    const Xform xf{ratio, 0.5f * (float)W, 0.5f * (float)H, 256 * H};
    const int vs[3] = {en.x, en.y, en.z};
    int2 q[3];
#pragma unroll
    for (int j = 0; j < 3; j++) {
        float dx = 0.0f, dy = 0.0f;
        if (en.w > 0 && ((en.w - 1) >> 2) == j) fd_offset((en.w - 1) & 3, dp, dx, dy);
        q[j] = to_fixed(pts[vs[j]], dx, dy, xf);
    }
    long long sr = 0, sg = 0, sb = 0, s2 = 0; int cnt = 0, odd = 0;
#pragma unroll
    for (int j = 0; j < 3; j++) {
        EdgeWalk w;
        w.init(q[j], q[(j + 1) % 3], H, lane, 32);
        unsigned un = 0, uo = 0; ACC ur = 0, ug = 0, ub = 0; unsigned long long us2 = 0;
        walk_edge<32, kWalkBatchWide, ACC>(w, table, pitch, W, un, uo, ur, ug, ub, us2);
        const int sgn = w.sign;
        cnt += sgn * (int)un; odd += sgn * (int)uo;
        sr += sgn * (long long)ur; sg += sgn * (long long)ug; sb += sgn * (long long)ub; s2 += sgn * (long long)us2;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o); odd += __shfl_xor_sync(0xffffffffu, odd, o);
        sr += __shfl_xor_sync(0xffffffffu, sr, o); sg += __shfl_xor_sync(0xffffffffu, sg, o);
        sb += __shfl_xor_sync(0xffffffffu, sb, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (lane != 0) return;
    if (cnt < 0) { cnt = -cnt; odd = -odd; sr = -sr; sg = -sg; sb = -sb; s2 = -s2; }
    cn[i] = cnt;
    ca[i] = make_int4((int)sr, (int)sg, (int)sb, 0);
    long long e = 0;
    if (cnt > 0) {
        int m0, m1, m2;                   // the mean colour exactly as eval_variant forms it
        if (sr == (int)sr && sg == (int)sg && sb == (int)sb) { m0 = (int)sr / cnt; m1 = (int)sg / cnt; m2 = (int)sb / cnt; }
        else { m0 = (int)(sr / cnt); m1 = (int)(sg / cnt); m2 = (int)(sb / cnt); }
        e = energy_closed_form(cnt, odd, sr, sg, sb, s2, m0, m1, m2);
    }
    ten[i] = (int)(unsigned)(unsigned long long)e;
}

// un-fused gradient.cs / shift.cs, for the literal doshift() entry point
// triangle.vs:93-98: only vertices some triangle references are reset
__global__ void k_zero_gradient(int2* gr, int np, const int* __restrict__ corner_off)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < np && corner_off[v + 1] > corner_off[v]) gr[v] = make_int2(0, 0);
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
