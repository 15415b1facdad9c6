// mview.cuh -- the multiview half of the hot path (source/multiview.hpp of the reference), batched:
//   k_f8point     one CTA per image pair: Hartley normalisation (multiview.hpp:62-87), the 9x9 normal
//                 matrix A^T A of the N x 9 design matrix (:150-162) accumulated with FP64 tensor-core
//                 MMAs (mma.sync.m8n8k4.f64 -> DMMA in SASS; the one genuinely dense contraction of
//                 the path), its null vector by Jacobi, the rank-2 projection (:173-176) and the
//                 de-normalisation (:178-179);
//   k_triangulate one thread per match: optimal correction (:415-527), then HDLT (:370-380) with the
//                 pose of E = K^T F K under check = 3 (:541-621).
// A^T A in FP64 instead of a float32 SVD of A: float32 accumulation of A^T A misses the reference by
// 1e-3, float64 by 1e-12 (hard part 5).  tcgen05 has no FP64 kind and a minimum
// M of 64; the warp-level DMMA is the right size for a 9-wide problem.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/tpose/detail/smallmath.h"

namespace tpose_dev {

constexpr int kF8Threads = 256;

struct Norm { float cx, cy, scale; };

constexpr int kNormChunk = 2048;          // values staged per pass of the in-order float32 sums

// multiview.hpp:62-87 -- centroid, mean distance, scale sqrt(2)/dist, LITERALLY: the reference adds
// float32 values in index order (`c += p`, `dist += length(p)`), and float addition is not
// associative, so the sums are taken by one thread in that order; the addends (p, then
// length(p - c) with glm's sqrt(dot)) are produced by the whole block, a chunk at a time.
// s_val: 2 * kNormChunk floats.  All threads get the result.
__device__ Norm normalise_stats(const float2* __restrict__ p, int n, float* s_val, float* s_out)
{
    float cx = 0.0f, cy = 0.0f;
    for (int c0 = 0; c0 < n; c0 += kNormChunk) {
        const int m = min(kNormChunk, n - c0);
        for (int i = threadIdx.x; i < m; i += kF8Threads) { s_val[2 * i] = p[c0 + i].x; s_val[2 * i + 1] = p[c0 + i].y; }
        __syncthreads();
        if (threadIdx.x == 0) for (int i = 0; i < m; i++) { cx = __fadd_rn(cx, s_val[2 * i]); cy = __fadd_rn(cy, s_val[2 * i + 1]); }
        __syncthreads();
    }
    if (threadIdx.x == 0) { s_out[0] = __fdiv_rn(cx, (float)n); s_out[1] = __fdiv_rn(cy, (float)n); }
    __syncthreads();
    Norm r;
    r.cx = s_out[0]; r.cy = s_out[1];
    float dist = 0.0f;
    for (int c0 = 0; c0 < n; c0 += kNormChunk) {
        const int m = min(kNormChunk, n - c0);
        for (int i = threadIdx.x; i < m; i += kF8Threads) {
            const float dx = __fsub_rn(p[c0 + i].x, r.cx), dy = __fsub_rn(p[c0 + i].y, r.cy);
            s_val[i] = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));       // glm::length
        }
        __syncthreads();
        if (threadIdx.x == 0) for (int i = 0; i < m; i++) dist = __fadd_rn(dist, s_val[i]);
        __syncthreads();
    }
    if (threadIdx.x == 0) s_out[2] = (float)(sqrt(2.0) / (double)__fdiv_rn(dist, (float)n));   // `float scale = sqrt(2) / dist`
    __syncthreads();
    r.scale = s_out[2];
    __syncthreads();
    return r;
}

// tpose::sm::jacobi_eigen<9> (cyclic Jacobi, smallmath.h) with the rotations in the same order and
// every element updated by the same expression, so the result is the serial one bit for bit -- but
// the nine rows / columns of a rotation are updated by nine lanes at once (and V by nine more).
// Called by one full warp; A (destroyed) and V are 81 doubles in shared memory.
__device__ void jacobi_eigen9_warp(double* A, double* V)
{
    constexpr int N = 9;
    const int lane = threadIdx.x & 31;
    for (int i = lane; i < N * N; i += 32) V[i] = (i / N == i % N) ? 1.0 : 0.0;
    __syncwarp();
    for (int sweep = 0; sweep < 30; sweep++) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < N; i++) {                       // (every lane: same values, same order as the serial code)
            diag += A[i * N + i] * A[i * N + i];
            for (int j = i + 1; j < N; j++) off += A[i * N + j] * A[i * N + j];
        }
        if (off <= 1e-36 * diag || off == 0.0) break;
        for (int p = 0; p < N - 1; p++)
            for (int q = p + 1; q < N; q++) {
                const double apq = A[p * N + q];
                if (apq == 0.0) continue;                   // (uniform: every lane read the same element)
                const double theta = (A[q * N + q] - A[p * N + p]) / (2.0 * apq);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                __syncwarp();
                if (lane < N) {                             // columns p, q
                    const int k = lane;
                    const double akp = A[k * N + p], akq = A[k * N + q];
                    A[k * N + p] = c * akp - s * akq;
                    A[k * N + q] = s * akp + c * akq;
                } else if (lane < 2 * N) {
                    const int k = lane - N;
                    const double vkp = V[k * N + p], vkq = V[k * N + q];
                    V[k * N + p] = c * vkp - s * vkq;
                    V[k * N + q] = s * vkp + c * vkq;
                }
                __syncwarp();
                if (lane < N) {                             // rows p, q
                    const int k = lane;
                    const double apk = A[p * N + k], aqk = A[q * N + k];
                    A[p * N + k] = c * apk - s * aqk;
                    A[q * N + k] = s * apk + c * aqk;
                }
                __syncwarp();
            }
    }
    __syncwarp();
}

__device__ __forceinline__ float2 normalise_point(float2 p, const Norm& nm)
{
    return make_float2(__fmul_rn(__fsub_rn(p.x, nm.cx), nm.scale), __fmul_rn(__fsub_rn(p.y, nm.cy), nm.scale));
}

__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// The shared solve of F_8Point (multiview.hpp:150-176) and of one F_Sampson iteration (:197-238):
// the 9x9 normal matrix of the (optionally re-weighted) N x 9 design matrix on the FP64 tensor
// cores, its null vector by Jacobi, the rank-2 projection.  s_Fw != nullptr: rows are scaled by
// the Sampson weight computed from the float32 matrix s_Fw (and by wuser[i] if given).  Thread 0
// returns with F (row-major, rank 2, float32-rounded like the reference's Matrix3f) -- in the
// NORMALISED frame.  All threads must call; ends with a barrier.
__device__ void solve_normalised(const float2* __restrict__ pa, const float2* __restrict__ pb, int n, const Norm& na,
                                 const Norm& nb, const float* s_Fw, const float* __restrict__ wuser, double* s_G,
                                 double* s_V, double* F)
{
    for (int i = threadIdx.x; i < 81; i += kF8Threads) s_G[i] = 0.0;
    __syncthreads();
    float Fw[9], Fwt[9];
    if (s_Fw)
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) { Fw[r * 3 + c] = s_Fw[r * 3 + c]; Fwt[r * 3 + c] = s_Fw[c * 3 + r]; }

    // 8x8 block of A^T A on the FP64 tensor cores: per MMA one warp contracts 4 matches.
    // A-operand (8x4, row = lane/4, k = lane%4) and B-operand (4x8, k = lane%4, col = lane/4) are the
    // same number: element `lane/4` of the design row of match `4*chunk + lane%4`.
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int col = lane >> 2, kk = lane & 3;
    double d0 = 0.0, d1 = 0.0, col8 = 0.0, g88 = 0.0;
    const int nchunks = (n + 3) >> 2;
    for (int ch = warp; ch < nchunks; ch += kF8Threads / 32) {
        const int i = 4 * ch + kk;
        double e = 0.0, e8 = 0.0;
        if (i < n) {
            const float2 a = normalise_point(pa[i], na), b = normalise_point(pb[i], nb);
            float w = 1.0f;
            if (s_Fw) {     // W = 1 / (L.x^2 + L.y^2 + R.x^2 + R.y^2), L = eline(F^T, pB), R = eline(F, pA)  (:199-203)
                const float lx = __fadd_rn(__fadd_rn(__fmul_rn(Fwt[0], b.x), __fmul_rn(Fwt[1], b.y)), Fwt[2]);
                const float ly = __fadd_rn(__fadd_rn(__fmul_rn(Fwt[3], b.x), __fmul_rn(Fwt[4], b.y)), Fwt[5]);
                const float lz = __fadd_rn(__fadd_rn(__fmul_rn(Fwt[6], b.x), __fmul_rn(Fwt[7], b.y)), Fwt[8]);
                const float rx = __fadd_rn(__fadd_rn(__fmul_rn(Fw[0], a.x), __fmul_rn(Fw[1], a.y)), Fw[2]);
                const float ry = __fadd_rn(__fadd_rn(__fmul_rn(Fw[3], a.x), __fmul_rn(Fw[4], a.y)), Fw[5]);
                const float rz = __fadd_rn(__fadd_rn(__fmul_rn(Fw[6], a.x), __fmul_rn(Fw[7], a.y)), Fw[8]);
                const float Lx = __fdiv_rn(lx, lz), Ly = __fdiv_rn(ly, lz), Rx = __fdiv_rn(rx, rz), Ry = __fdiv_rn(ry, rz);
                float den = __fmul_rn(Lx, Lx);
                den = __fadd_rn(den, __fmul_rn(Ly, Ly));
                den = __fadd_rn(den, __fmul_rn(Rx, Rx));
                den = __fadd_rn(den, __fmul_rn(Ry, Ry));
                w = __fdiv_rn(1.0f, den);
                if (wuser) w = __fmul_rn(wuser[i], w);
            }
            // row elements as the reference forms them: ((w * first) * second), float32
            float el;
            switch (col) {
            case 0: el = s_Fw ? __fmul_rn(__fmul_rn(w, a.x), b.x) : __fmul_rn(a.x, b.x); break;
            case 1: el = s_Fw ? __fmul_rn(__fmul_rn(w, a.y), b.x) : __fmul_rn(a.y, b.x); break;
            case 2: el = s_Fw ? __fmul_rn(w, b.x) : b.x; break;
            case 3: el = s_Fw ? __fmul_rn(__fmul_rn(w, a.x), b.y) : __fmul_rn(a.x, b.y); break;
            case 4: el = s_Fw ? __fmul_rn(__fmul_rn(w, a.y), b.y) : __fmul_rn(a.y, b.y); break;
            case 5: el = s_Fw ? __fmul_rn(w, b.y) : b.y; break;
            case 6: el = s_Fw ? __fmul_rn(w, a.x) : a.x; break;
            default: el = s_Fw ? __fmul_rn(w, a.y) : a.y; break;
            }
            e = (double)el;
            e8 = (double)w;                                   // ninth column: w * 1
        }
        col8 += e * e8;
        if (col == 0) g88 += e8 * e8;
        dmma_m8n8k4(d0, d1, e, e);
    }
    // accumulator fragment: row = lane/4, columns 2*(lane%4) and +1
    atomicAdd(&s_G[col * 9 + 2 * kk], d0);
    atomicAdd(&s_G[col * 9 + 2 * kk + 1], d1);
    // ninth row / column
    col8 += __shfl_xor_sync(0xffffffffu, col8, 1);
    col8 += __shfl_xor_sync(0xffffffffu, col8, 2);
    g88 += __shfl_xor_sync(0xffffffffu, g88, 1);
    g88 += __shfl_xor_sync(0xffffffffu, g88, 2);
    if (kk == 0) { atomicAdd(&s_G[col * 9 + 8], col8); atomicAdd(&s_G[8 * 9 + col], col8); }
    if (lane == 0) atomicAdd(&s_G[80], g88);
    __syncthreads();

    if (threadIdx.x < 32) jacobi_eigen9_warp(s_G, s_V);
    __syncthreads();
    if (threadIdx.x == 0) {
        const int m = tpose::sm::argmin_diag<9>(s_G);
        for (int i = 0; i < 9; i++) F[i] = (double)(float)s_V[i * 9 + m];      // VectorXf f -> Matrix3f F
        tpose::sm::rank2_project(F);
        for (int i = 0; i < 9; i++) F[i] = (double)(float)F[i];                // held as Matrix3f
    }
    __syncthreads();
}

// F = HB^T F HA, then / F(2,2); H = [[s,0,-cx s],[0,s,-cy s],[0,0,1]] with float32 entries (multiview.hpp:80-84,178-179)
__device__ void denormalise(const double* F, const Norm& na, const Norm& nb, double* R)
{
    const double HA[9] = {na.scale, 0, (double)__fmul_rn(-na.cx, na.scale), 0, na.scale, (double)__fmul_rn(-na.cy, na.scale), 0, 0, 1};
    const double HB[9] = {nb.scale, 0, (double)__fmul_rn(-nb.cx, nb.scale), 0, nb.scale, (double)__fmul_rn(-nb.cy, nb.scale), 0, 0, 1};
    double T[9];
    tpose::sm::mat3_tmul(HB, F, T);
    tpose::sm::mat3_mul(T, HA, R);
    const double s = R[8];
    for (int i = 0; i < 9; i++) R[i] /= s;
}

// offsets[nb+1]: match range of each pair inside A / B.  F9: nb x 9 floats, row-major.
// iters == 0: F_8Point (multiview.hpp:130-183).  iters > 0: F_Sampson (:187-299), i.e. the 8-point
// result, then `iters` solves re-weighted by the Sampson denominators of the current F; wuser
// (may be null) are the per-match weights of the three-argument overload.  The first iteration
// uses the DE-normalised 8-point F with the normalised points, like the reference (:189-205).
__global__ void __launch_bounds__(kF8Threads)
k_f8point(const float2* __restrict__ A, const float2* __restrict__ B, const int* __restrict__ offsets,
          const float* __restrict__ wuser, int iters, float* __restrict__ F9)
{
    __shared__ float s_val[2 * kNormChunk];
    __shared__ float s_out[4];
    __shared__ double s_G[81];                 // A^T A
    __shared__ double s_V[81];
    __shared__ float s_F[9];
    const int pair = blockIdx.x;
    const int o0 = offsets[pair], n = offsets[pair + 1] - o0;
    float* out = F9 + 9 * pair;
    if (n <= 0) {                              // multiview.hpp:139-142: empty set -> identity
        if (threadIdx.x < 9) out[threadIdx.x] = (threadIdx.x % 4 == 0) ? 1.0f : 0.0f;
        return;
    }
    const float2* pa = A + o0;
    const float2* pb = B + o0;
    const Norm na = normalise_stats(pa, n, s_val, s_out);
    const Norm nb = normalise_stats(pb, n, s_val, s_out);
    double F[9], R[9];
    solve_normalised(pa, pb, n, na, nb, nullptr, nullptr, s_G, s_V, F);
    if (threadIdx.x == 0) {
        denormalise(F, na, nb, R);
        for (int i = 0; i < 9; i++) s_F[i] = (float)R[i];
    }
    __syncthreads();
    for (int k = 0; k < iters; k++) {
        solve_normalised(pa, pb, n, na, nb, s_F, wuser ? wuser + o0 : nullptr, s_G, s_V, F);
        if (threadIdx.x == 0) for (int i = 0; i < 9; i++) s_F[i] = (float)F[i];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (iters > 0) { denormalise(F, na, nb, R); for (int i = 0; i < 9; i++) out[i] = (float)R[i]; }
        else for (int i = 0; i < 9; i++) out[i] = s_F[i];
    }
}

// One thread per match of one pair (blockIdx.y = pair).  F9 / K9: per pair, row-major float.
// X4: homogeneous 3-D points divided by w (multiview.hpp:617-619); Ac / Bc (may be null) receive the
// corrected matches.  do_correct = 0 skips the optimal correction (plain HDLT of the raw matches).
__global__ void __launch_bounds__(128)
k_triangulate(const float* __restrict__ F9, const float* __restrict__ K9, const float2* __restrict__ A,
              const float2* __restrict__ B, const int* __restrict__ offsets, int do_correct,
              float4* __restrict__ X4, float2* __restrict__ Ac, float2* __restrict__ Bc)
{
    __shared__ double s_F[9], s_KPA[12], s_KPB[12];
    const int pair = blockIdx.y;
    const int o0 = offsets[pair], n = offsets[pair + 1] - o0;
    if ((int)(blockIdx.x * blockDim.x) >= n) return;
    if (threadIdx.x == 0) {
        double F[9], K[9], T[9], E[9], R1[9], R2[9], t[3];
        for (int i = 0; i < 9; i++) { F[i] = F9[9 * pair + i]; K[i] = K9[9 * pair + i]; s_F[i] = F[i]; }
        tpose::sm::mat3_tmul(K, F, T);                     // E = K^T F K (multiview.hpp:541)
        tpose::sm::mat3_mul(T, K, E);
        tpose::sm::get_pose(E, R1, R2, t);                 // check = 3: PB = [R2 | -t] (:584-588)
        for (int r = 0; r < 3; r++) {
            // K*PA = [K | 0]; K*PB = [K R2 | -K t]; both held as float32 like the reference's MatrixXf
            for (int c = 0; c < 3; c++) {
                s_KPA[r * 4 + c] = K[r * 3 + c];
                s_KPB[r * 4 + c] = (double)(float)(K[r * 3] * R2[c] + K[r * 3 + 1] * R2[3 + c] + K[r * 3 + 2] * R2[6 + c]);
            }
            s_KPA[r * 4 + 3] = 0.0;
            s_KPB[r * 4 + 3] = (double)(float)(-(K[r * 3] * t[0] + K[r * 3 + 1] * t[1] + K[r * 3 + 2] * t[2]));
        }
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double a[2] = {A[o0 + i].x, A[o0 + i].y}, b[2] = {B[o0 + i].x, B[o0 + i].y};
    if (do_correct) {
        tpose::sm::correct_match(s_F, a, b);
        a[0] = (double)(float)a[0]; a[1] = (double)(float)a[1];       // stored back into vec2 (:523-524)
        b[0] = (double)(float)b[0]; b[1] = (double)(float)b[1];
    }
    if (Ac) Ac[o0 + i] = make_float2((float)a[0], (float)a[1]);
    if (Bc) Bc[o0 + i] = make_float2((float)b[0], (float)b[1]);
    double X[4];
    tpose::sm::hdlt(s_KPA, s_KPB, a, b, X);
    X4[o0 + i] = make_float4((float)(X[0] / X[3]), (float)(X[1] / X[3]), (float)(X[2] / X[3]), 1.0f);
}

// ---- match selection by two-way warp inconsistency (tests/compute_fundamental_mat/model.h:1-82) -----
// The reference keeps this step as commented-out code: from the two warped stacks of a view pair
// (tr = the level as warp left it, trW = the same level of the other direction's .warp stack) it
// takes, for every vertex i of A and then of B,
//     d_i = length(trW.originpoints[i] - tr.points[i])                            model.h:22,52
// and keeps the vertex as a match (weight 1/d_i) unless d_i > max_dist (:27,:60; 2e-7 as coded),
// either end lies on the domain boundary (:30-32) or outside the central half of the domain
// (:34-41).  A's matches run originpoints -> points, B's points -> originpoints (:24-25,:54-55),
// both through T, which -- a row-major literal in a column-major glm::mat3 -- maps
// (x, y) -> ((0.5f/RATIO) x, (-0.5f/RATIO) y) (model.h:7-11).
// One block; the kept matches are compacted in index order (A's first), so the output is the
// vector the reference's push_backs would build.  out_n[0] = number of matches.
constexpr int kSelectThreads = 512;

__device__ __forceinline__ bool domain_boundary(float2 p, float ratio)           // triangulation.hpp:132-140
{
    return p.x <= -ratio || p.y <= -1.0f || p.x >= ratio || p.y >= 1.0f;
}

__global__ void __launch_bounds__(kSelectThreads)
k_select_matches(const float2* __restrict__ originA, const float2* __restrict__ pointsA, const float2* __restrict__ worigA, int npa,
                 const float2* __restrict__ pointsB, const float2* __restrict__ originB, const float2* __restrict__ worigB, int npb,
                 float ratio, double max_dist, float2* __restrict__ X, float2* __restrict__ Y, float* __restrict__ Wt,
                 float* __restrict__ dist_out, int* __restrict__ out_n)
{
    __shared__ int s_warp[kSelectThreads / 32];
    __shared__ int s_base;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_base = 0;
    __syncthreads();
    const float cx = __fdiv_rn(0.5f, ratio), cy = __fdiv_rn(-0.5f, ratio);
    const int total = npa + npb;
    for (int c0 = 0; c0 < total; c0 += kSelectThreads) {
        const int i = c0 + tid;
        bool keep = false;
        float2 pX = make_float2(0.0f, 0.0f), pY = pX;
        float d = 0.0f;
        if (i < total) {
            const bool isA = i < npa;
            const int k = isA ? i : i - npa;
            const float2 wo = isA ? worigA[k] : worigB[k];
            const float2 pt = isA ? pointsA[k] : pointsB[k];
            const float2 og = isA ? originA[k] : originB[k];
            const float dx = __fsub_rn(wo.x, pt.x), dy = __fsub_rn(wo.y, pt.y);
            d = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
            if (dist_out) dist_out[i] = d;
            pX = isA ? og : pt;                    // A: originpoints -> points; B: points -> originpoints
            pY = isA ? pt : og;
            keep = !((double)d > max_dist) && !domain_boundary(pX, ratio) && !domain_boundary(pY, ratio);
            const double hx = (double)ratio / 2.0, hy = 1.0 / 2.0;
            if ((double)pX.x < -hx || (double)pX.x > hx || (double)pX.y < -hy || (double)pX.y > hy) keep = false;
            if ((double)pY.x < -hx || (double)pY.x > hx || (double)pY.y < -hy || (double)pY.y > hy) keep = false;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_warp[wid] = __popc(bal);
        __syncthreads();
        int before = s_base;
        for (int w2 = 0; w2 < wid; w2++) before += s_warp[w2];
        if (keep) {
            const int at = before + __popc(bal & ((1u << lane) - 1u));
            // T * vec3(p, 1): (c x + 0 y) + 0 in float32, like glm's mat3 * vec3
            X[at] = make_float2(__fadd_rn(__fadd_rn(__fmul_rn(cx, pX.x), __fmul_rn(0.0f, pX.y)), 0.0f),
                                __fadd_rn(__fadd_rn(__fmul_rn(0.0f, pX.x), __fmul_rn(cy, pX.y)), 0.0f));
            Y[at] = make_float2(__fadd_rn(__fadd_rn(__fmul_rn(cx, pY.x), __fmul_rn(0.0f, pY.y)), 0.0f),
                                __fadd_rn(__fadd_rn(__fmul_rn(0.0f, pY.x), __fmul_rn(cy, pY.y)), 0.0f));
            Wt[at] = __fdiv_rn(1.0f, d);
        }
        __syncthreads();
        if (tid == 0) { int t = s_base; for (int w2 = 0; w2 < kSelectThreads / 32; w2++) t += s_warp[w2]; s_base = t; }
        __syncthreads();
    }
    if (tid == 0) out_n[0] = s_base;
}

// HDLT alone with caller-supplied 3x4 projections (multiview.hpp:370-380); X is NOT divided by w,
// its sign is normalised so that the largest-magnitude component is positive.
__global__ void __launch_bounds__(128)
k_hdlt(const float* __restrict__ PA12, const float* __restrict__ PB12, const float2* __restrict__ A,
       const float2* __restrict__ B, int n, float4* __restrict__ X4)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double PA[12], PB[12];
    for (int k = 0; k < 12; k++) { PA[k] = PA12[k]; PB[k] = PB12[k]; }
    const double a[2] = {A[i].x, A[i].y}, b[2] = {B[i].x, B[i].y};
    double X[4];
    tpose::sm::hdlt(PA, PB, a, b, X);
    int big = 0;
    for (int k = 1; k < 4; k++) if (fabs(X[k]) > fabs(X[big])) big = k;
    const double s = X[big] < 0.0 ? -1.0 : 1.0;
    X4[i] = make_float4((float)(s * X[0]), (float)(s * X[1]), (float)(s * X[2]), (float)(s * X[3]));
}

}  // namespace tpose_dev
