// tpose/detail/smallmath.h -- small dense linear algebra in double, written once for the host
// headers (tpose/utility, tpose/multiview) and the CUDA kernels (csrc/mview.cuh).
//
// These routines stand where the reference calls Eigen (JacobiSVD / EigenSolver, source/
// multiview.hpp and source/utility.hpp:104-118).  Eigen's float32 two-sided Jacobi cannot be
// reproduced bit for bit without Eigen, so everything here is float64 and the conventions Eigen
// leaves open are fixed explicitly (see svd3_proper and realroots_sorted).
#ifndef TPOSE_DETAIL_SMALLMATH_H
#define TPOSE_DETAIL_SMALLMATH_H

#include <math.h>

#if defined(__CUDACC__)
#define TPOSE_HD __host__ __device__
#else
#define TPOSE_HD
#endif

namespace tpose {
namespace sm {

// ---- symmetric eigen-decomposition by cyclic Jacobi rotations ------------------------------------
// A (N x N, row-major, symmetric) is destroyed: its diagonal ends up holding the eigenvalues;
// V's COLUMNS are the eigenvectors.  Converges quadratically; 30 sweeps is far more than needed.
template <int N>
TPOSE_HD inline void jacobi_eigen(double* A, double* V)
{
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++) V[i * N + j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 30; sweep++) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < N; i++) {
            diag += A[i * N + i] * A[i * N + i];
            for (int j = i + 1; j < N; j++) off += A[i * N + j] * A[i * N + j];
        }
        if (off <= 1e-36 * diag || off == 0.0) break;
        for (int p = 0; p < N - 1; p++)
            for (int q = p + 1; q < N; q++) {
                const double apq = A[p * N + q];
                if (apq == 0.0) continue;
                const double theta = (A[q * N + q] - A[p * N + p]) / (2.0 * apq);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < N; k++) {           // columns p, q
                    const double akp = A[k * N + p], akq = A[k * N + q];
                    A[k * N + p] = c * akp - s * akq;
                    A[k * N + q] = s * akp + c * akq;
                }
                for (int k = 0; k < N; k++) {           // rows p, q
                    const double apk = A[p * N + k], aqk = A[q * N + k];
                    A[p * N + k] = c * apk - s * aqk;
                    A[q * N + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < N; k++) {
                    const double vkp = V[k * N + p], vkq = V[k * N + q];
                    V[k * N + p] = c * vkp - s * vkq;
                    V[k * N + q] = s * vkp + c * vkq;
                }
            }
    }
}

// index of the smallest diagonal entry after jacobi_eigen
template <int N>
TPOSE_HD inline int argmin_diag(const double* A)
{
    int m = 0;
    for (int i = 1; i < N; i++) if (A[i * N + i] < A[m * N + m]) m = i;
    return m;
}

TPOSE_HD inline void cross3(const double* a, const double* b, double* c)
{
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

// ---- 3x3 SVD, M = U diag(S) V^T, row-major ----------------------------------------------------------
// Gauge (Eigen leaves it open; GetPose and the root selection of the optimal correction depend on it):
//   * singular values descending;
//   * third columns are cross products of the first two, so det U = det V = +1;
//   * of the two remaining joint sign choices, the one where the largest-magnitude component of
//     V's third column is positive.
// Meant for rank >= 2 matrices (fundamental / essential matrices): u_i = M v_i / s_i for i < 2.
TPOSE_HD inline void svd3_proper(const double* M, double* U, double* S, double* V)
{
    double G[9], E[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double g = 0.0;
            for (int k = 0; k < 3; k++) g += M[k * 3 + i] * M[k * 3 + j];
            G[i * 3 + j] = g;
        }
    jacobi_eigen<3>(G, E);
    int ord[3] = {0, 1, 2};                              // sort eigenvalues descending
    for (int a = 0; a < 2; a++)
        for (int b = a + 1; b < 3; b++)
            if (G[ord[b] * 3 + ord[b]] > G[ord[a] * 3 + ord[a]]) { const int t = ord[a]; ord[a] = ord[b]; ord[b] = t; }
    double v[3][3], u[3][3];
    for (int c = 0; c < 2; c++) {
        for (int r = 0; r < 3; r++) v[c][r] = E[r * 3 + ord[c]];
        const double lam = G[ord[c] * 3 + ord[c]];
        S[c] = sqrt(lam > 0.0 ? lam : 0.0);
        double n2 = 0.0;
        for (int r = 0; r < 3; r++) {
            u[c][r] = M[r * 3 + 0] * v[c][0] + M[r * 3 + 1] * v[c][1] + M[r * 3 + 2] * v[c][2];
            n2 += u[c][r] * u[c][r];
        }
        const double inv = n2 > 0.0 ? 1.0 / sqrt(n2) : 0.0;
        for (int r = 0; r < 3; r++) u[c][r] *= inv;
    }
    {   // re-orthogonalise u1 against u0 (they are orthogonal up to rounding already)
        double d = u[0][0] * u[1][0] + u[0][1] * u[1][1] + u[0][2] * u[1][2], n2 = 0.0;
        for (int r = 0; r < 3; r++) { u[1][r] -= d * u[0][r]; n2 += u[1][r] * u[1][r]; }
        const double inv = n2 > 0.0 ? 1.0 / sqrt(n2) : 0.0;
        for (int r = 0; r < 3; r++) u[1][r] *= inv;
    }
    const double lam2 = G[ord[2] * 3 + ord[2]];
    S[2] = sqrt(lam2 > 0.0 ? lam2 : 0.0);
    cross3(v[0], v[1], v[2]);
    cross3(u[0], u[1], u[2]);
    int big = 0;
    for (int r = 1; r < 3; r++) if (fabs(v[2][r]) > fabs(v[2][big])) big = r;
    if (v[2][big] < 0.0)
        for (int r = 0; r < 3; r++) { v[0][r] = -v[0][r]; u[0][r] = -u[0][r]; v[2][r] = -v[2][r]; u[2][r] = -u[2][r]; }
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < 3; r++) { U[r * 3 + c] = u[c][r]; V[r * 3 + c] = v[c][r]; }
}

// rank-2 projection F - s3 u3 v3^T of a 3x3 matrix (multiview.hpp:173-176), in place
TPOSE_HD inline void rank2_project(double* F)
{
    double G[9], E[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double g = 0.0;
            for (int k = 0; k < 3; k++) g += F[k * 3 + i] * F[k * 3 + j];
            G[i * 3 + j] = g;
        }
    jacobi_eigen<3>(G, E);
    const int m = argmin_diag<3>(G);
    const double v[3] = {E[0 * 3 + m], E[1 * 3 + m], E[2 * 3 + m]};
    for (int r = 0; r < 3; r++) {
        const double fv = F[r * 3] * v[0] + F[r * 3 + 1] * v[1] + F[r * 3 + 2] * v[2];
        for (int c = 0; c < 3; c++) F[r * 3 + c] -= fv * v[c];
    }
}

TPOSE_HD inline void mat3_mul(const double* A, const double* B, double* C)
{
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}
TPOSE_HD inline void mat3_tmul(const double* A, const double* B, double* C)   // A^T B
{
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) C[i * 3 + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
}
TPOSE_HD inline void mat3_mult(const double* A, const double* B, double* C)   // A B^T
{
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) C[i * 3 + j] = A[i * 3] * B[j * 3] + A[i * 3 + 1] * B[j * 3 + 1] + A[i * 3 + 2] * B[j * 3 + 2];
}

// ---- eigenvalues of a small real matrix, for the companion matrix of utility.hpp:104-118 ----------
// Elimination to Hessenberg form, then shifted QR with deflation.  a: n x n row-major, destroyed.
// Real eigenvalues get wi == 0 exactly, which is what realroots tests for (utility.hpp:128).
// Returns false if an eigenvalue failed to converge in 60 iterations.
TPOSE_HD inline bool hqr_eigenvalues(double* a, int n, double* wr, double* wi)
{
#define TPOSE_A(i, j) a[(i) * n + (j)]
    for (int m = 1; m < n - 1; m++) {
        int piv = m; double big = 0.0;
        for (int i = m; i < n; i++) if (fabs(TPOSE_A(i, m - 1)) > fabs(big)) { big = TPOSE_A(i, m - 1); piv = i; }
        if (piv != m) {
            for (int j = m - 1; j < n; j++) { const double t = TPOSE_A(piv, j); TPOSE_A(piv, j) = TPOSE_A(m, j); TPOSE_A(m, j) = t; }
            for (int j = 0; j < n; j++) { const double t = TPOSE_A(j, piv); TPOSE_A(j, piv) = TPOSE_A(j, m); TPOSE_A(j, m) = t; }
        }
        if (big == 0.0) continue;
        for (int i = m + 1; i < n; i++) {
            double f = TPOSE_A(i, m - 1);
            if (f == 0.0) continue;
            f /= big; TPOSE_A(i, m - 1) = f;
            for (int j = m; j < n; j++) TPOSE_A(i, j) -= f * TPOSE_A(m, j);
            for (int j = 0; j < n; j++) TPOSE_A(j, m) += f * TPOSE_A(j, i);
        }
    }
    for (int i = 2; i < n; i++) for (int j = 0; j < i - 1; j++) TPOSE_A(i, j) = 0.0;
    double norm = 0.0;
    for (int i = 0; i < n; i++) for (int j = (i > 0 ? i - 1 : 0); j < n; j++) norm += fabs(TPOSE_A(i, j));
    int hi = n - 1; double shift = 0.0;
    while (hi >= 0) {
        int iter = 0;
        for (;;) {
            int lo;
            for (lo = hi; lo >= 1; lo--) {                       // negligible sub-diagonal element?
                double s = fabs(TPOSE_A(lo - 1, lo - 1)) + fabs(TPOSE_A(lo, lo));
                if (s == 0.0) s = norm;
                if (fabs(TPOSE_A(lo, lo - 1)) + s == s) { TPOSE_A(lo, lo - 1) = 0.0; break; }
            }
            double x = TPOSE_A(hi, hi);
            if (lo == hi) { wr[hi] = x + shift; wi[hi] = 0.0; hi--; break; }
            double y = TPOSE_A(hi - 1, hi - 1), w = TPOSE_A(hi, hi - 1) * TPOSE_A(hi - 1, hi);
            if (lo == hi - 1) {                                  // 2x2 block: real pair or conjugates
                const double p = 0.5 * (y - x), q = p * p + w, z = sqrt(fabs(q));
                x += shift;
                if (q >= 0.0) {
                    const double zz = p + (p >= 0.0 ? z : -z);
                    wr[hi - 1] = x + zz; wr[hi] = (zz != 0.0) ? x - w / zz : x + zz;
                    wi[hi - 1] = wi[hi] = 0.0;
                } else {
                    wr[hi - 1] = wr[hi] = x + p;
                    wi[hi - 1] = z; wi[hi] = -z;
                }
                hi -= 2;
                break;
            }
            if (iter == 60) return false;
            if (iter == 10 || iter == 20) {                      // exceptional shift
                shift += x;
                for (int i = 0; i <= hi; i++) TPOSE_A(i, i) -= x;
                const double s = fabs(TPOSE_A(hi, hi - 1)) + fabs(TPOSE_A(hi - 1, hi - 2));
                y = x = 0.75 * s; w = -0.4375 * s * s;
            }
            ++iter;
            int m; double p = 0, q = 0, r = 0;
            for (m = hi - 2; m >= lo; m--) {
                const double z = TPOSE_A(m, m), rr = x - z, ss = y - z;
                p = (rr * ss - w) / TPOSE_A(m + 1, m) + TPOSE_A(m, m + 1);
                q = TPOSE_A(m + 1, m + 1) - z - rr - ss;
                r = TPOSE_A(m + 2, m + 1);
                const double s = fabs(p) + fabs(q) + fabs(r);
                p /= s; q /= s; r /= s;
                if (m == lo) break;
                const double u = fabs(TPOSE_A(m, m - 1)) * (fabs(q) + fabs(r));
                const double v = fabs(p) * (fabs(TPOSE_A(m - 1, m - 1)) + fabs(z) + fabs(TPOSE_A(m + 1, m + 1)));
                if (u + v == v) break;
            }
            for (int i = m + 2; i <= hi; i++) { TPOSE_A(i, i - 2) = 0.0; if (i != m + 2) TPOSE_A(i, i - 3) = 0.0; }
            for (int k = m; k <= hi - 1; k++) {                  // Francis double-shift sweep
                if (k != m) {
                    p = TPOSE_A(k, k - 1); q = TPOSE_A(k + 1, k - 1); r = (k != hi - 1) ? TPOSE_A(k + 2, k - 1) : 0.0;
                    x = fabs(p) + fabs(q) + fabs(r);
                    if (x != 0.0) { p /= x; q /= x; r /= x; }
                }
                double s = sqrt(p * p + q * q + r * r);
                if (p < 0) s = -s;
                if (s == 0.0) continue;
                if (k == m) { if (lo != m) TPOSE_A(k, k - 1) = -TPOSE_A(k, k - 1); }
                else TPOSE_A(k, k - 1) = -s * x;
                p += s; x = p / s; y = q / s; const double z = r / s; q /= p; r /= p;
                for (int j = k; j <= hi; j++) {
                    p = TPOSE_A(k, j) + q * TPOSE_A(k + 1, j);
                    if (k != hi - 1) { p += r * TPOSE_A(k + 2, j); TPOSE_A(k + 2, j) -= p * z; }
                    TPOSE_A(k + 1, j) -= p * y; TPOSE_A(k, j) -= p * x;
                }
                const int last = hi < k + 3 ? hi : k + 3;
                for (int i = lo; i <= last; i++) {
                    p = x * TPOSE_A(i, k) + y * TPOSE_A(i, k + 1);
                    if (k != hi - 1) { p += z * TPOSE_A(i, k + 2); TPOSE_A(i, k + 2) -= p * r; }
                    TPOSE_A(i, k + 1) -= p * q; TPOSE_A(i, k) -= p;
                }
            }
        }
    }
#undef TPOSE_A
    return true;
}

TPOSE_HD inline double horner(double x, const double* a, int n)   // a[0] + a[1] x + ... + a[n-1] x^(n-1)
{
    double acc = a[n - 1];
    for (int i = n - 2; i >= 0; i--) acc = acc * x + a[i];
    return acc;
}

// Real roots of a0 + ... + a6 t^6 the way utility.hpp:122-138 finds them (companion-matrix
// eigenvalues with an exactly zero imaginary part, 25 Newton steps each), in ASCENDING order of the
// unpolished eigenvalue (the reference keeps Eigen's eigenvalue order, which is not reproducible).
TPOSE_HD inline int realroots_sorted(const double* a, double* out)
{
    double C[36], wr[6], wi[6];
    for (int i = 0; i < 36; i++) C[i] = 0.0;
    for (int k = 0; k < 6; k++) C[k * 6 + 5] = -a[k] / a[6];
    for (int k = 1; k < 6; k++) C[k * 6 + (k - 1)] = 1.0;
    if (!hqr_eigenvalues(C, 6, wr, wi)) return 0;
    int n = 0;
    for (int r = 0; r < 6; r++)
        if (wi[r] == 0.0) out[n++] = wr[r];
    for (int i = 1; i < n; i++) {                                // insertion sort
        const double v = out[i];
        int j = i - 1;
        while (j >= 0 && out[j] > v) { out[j + 1] = out[j]; j--; }
        out[j + 1] = v;
    }
    const double d[6] = {a[1] * 1.0, a[2] * 2.0, a[3] * 3.0, a[4] * 4.0, a[5] * 5.0, a[6] * 6.0};
    for (int i = 0; i < n; i++)
        for (int it = 0; it < 25; it++) out[i] -= horner(out[i], a, 7) / horner(out[i], d, 6);
    return n;
}

// ---- multiview.hpp:415-527: Hartley-Sturm optimal correction of one match ----------------------------
// F: row-major 3x3 (the float32 values of the reference's Matrix3f, held in doubles).  A, B are
// updated in place.  The selection loop keeps the reference's bug (cost evaluated at the loop
// index, multiview.hpp:496).  Returns the number of real roots found (0: match left untouched;
// the reference would read R[0] of an empty vector).
TPOSE_HD inline int correct_match(const double* F, double* A, double* B)
{
    const double TAi[9] = {1, 0, A[0], 0, 1, A[1], 0, 0, 1};
    const double TBi[9] = {1, 0, B[0], 0, 1, B[1], 0, 0, 1};
    double T[9], Fs[9];
    mat3_tmul(TBi, F, T);
    mat3_mul(T, TAi, Fs);
    for (int i = 0; i < 9; i++) Fs[i] = (double)(float)Fs[i];    // the reference stores this in a MatrixXf
    double U[9], S[3], V[9];
    svd3_proper(Fs, U, S, V);
    double eA[3] = {V[2], V[5], V[8]}, eB[3] = {U[2], U[5], U[8]};
    const double na = sqrt(eA[0] * eA[0] + eA[1] * eA[1]), nb = sqrt(eB[0] * eB[0] + eB[1] * eB[1]);
    for (int i = 0; i < 3; i++) { eA[i] /= na; eB[i] /= nb; }
    const double RA[9] = {eA[0], eA[1], 0, -eA[1], eA[0], 0, 0, 0, 1};
    const double RB[9] = {eB[0], eB[1], 0, -eB[1], eB[0], 0, 0, 0, 1};
    double Fr[9];
    mat3_mul(RB, Fs, T);
    mat3_mult(T, RA, Fr);
    const double m = eA[2], n = eB[2];
    const double a = Fr[4], b = Fr[5], c = Fr[7], d = Fr[8];

    double co[7];
    co[0] = b*b*c*d - a*b*d*d;
    co[1] = b*b*b*b + (b*b*c*c - a*a*d*d) + 2.0*b*b*d*d*n*n + d*d*d*d*n*n*n*n;
    co[2] = (a*b*c*c - a*a*c*d) + 4.0*a*b*b*b + 2.0*(b*b*c*d - a*b*d*d)*m*m + 4.0*(a*b*d*d + b*b*c*d)*n*n + 4.0*c*d*d*d*n*n*n*n;
    co[3] = 6.0*a*a*b*b + 2.0*(b*b*c*c - a*a*d*d)*m*m + 2.0*a*a*d*d*n*n + 8.0*a*b*c*d*n*n + 2.0*b*b*c*c*n*n + 6.0*c*c*d*d*n*n*n*n;
    co[4] = (b*b*c*d - a*b*d*d)*m*m*m*m + 4.0*a*a*a*b + 2.0*(a*b*c*c - a*a*c*d)*m*m + 4.0*(a*a*c*d + a*b*c*c)*n*n + 4.0*c*c*c*d*n*n*n*n;
    co[5] = a*a*a*a + (b*b*c*c - a*a*d*d)*m*m*m*m + 2.0*a*a*c*c*n*n + c*c*c*c*n*n*n*n;
    co[6] = (a*b*c*c - a*a*c*d)*m*m*m*m;

    double R[6];
    const int nr = realroots_sorted(co, R);
    if (nr == 0) return 0;
    auto cost = [&](double t) {
        const double ct = c * t + d, at = a * t + b;
        return t * t / (1.0 + m * m * t * t) + ct * ct / (at * at + n * n * ct * ct);
    };
    double t = R[0], minerr = cost(R[0]);
    for (int r = 1; r < nr; r++) {
        const double cur = cost((double)r);                      // sic
        if (cur < minerr) { t = R[r]; minerr = cur; }
    }
    const double LA[3] = {t * m, 1.0, -t};
    const double LB[3] = {-n * (c * t + d), a * t + b, c * t + d};
    const double XA[3] = {-LA[0] * LA[2], -LA[1] * LA[2], LA[0] * LA[0] + LA[1] * LA[1]};
    const double XB[3] = {-LB[0] * LB[2], -LB[1] * LB[2], LB[0] * LB[0] + LB[1] * LB[1]};
    double ya[3], yb[3], za[3], zb[3];
    for (int i = 0; i < 3; i++) {                                // RA^T XA, RB^T XB
        ya[i] = RA[0 * 3 + i] * XA[0] + RA[1 * 3 + i] * XA[1] + RA[2 * 3 + i] * XA[2];
        yb[i] = RB[0 * 3 + i] * XB[0] + RB[1 * 3 + i] * XB[1] + RB[2 * 3 + i] * XB[2];
    }
    for (int i = 0; i < 3; i++) {
        za[i] = TAi[i * 3] * ya[0] + TAi[i * 3 + 1] * ya[1] + TAi[i * 3 + 2] * ya[2];
        zb[i] = TBi[i * 3] * yb[0] + TBi[i * 3 + 1] * yb[1] + TBi[i * 3 + 2] * yb[2];
    }
    A[0] = za[0] / za[2]; A[1] = za[1] / za[2];
    B[0] = zb[0] / zb[2]; B[1] = zb[1] / zb[2];
    return nr;
}

// ---- multiview.hpp:391-411: essential matrix -> (R1, R2, t) under the svd3_proper gauge ---------------
TPOSE_HD inline void get_pose(const double* E, double* R1, double* R2, double* t)
{
    const double W[9] = {0, -1, 0, 1, 0, 0, 0, 0, 1};
    double U[9], S[3], V[9], T[9];
    svd3_proper(E, U, S, V);
    t[0] = U[2]; t[1] = U[5]; t[2] = U[8];
    mat3_mul(U, W, T);
    mat3_mult(T, V, R1);
    const double Wt[9] = {0, 1, 0, -1, 0, 0, 0, 0, 1};
    mat3_mul(U, Wt, T);
    mat3_mult(T, V, R2);
}

// ---- multiview.hpp:370-380: homogeneous DLT, X = last right singular vector of the 4x4 system ---------
// PA, PB: 3x4 row-major.  The sign of X is arbitrary (callers divide by X[3]).
TPOSE_HD inline void hdlt(const double* PA, const double* PB, const double* xa, const double* xb, double* X)
{
    double H[16];
    for (int j = 0; j < 4; j++) {
        // float32 rows, like the reference's MatrixXf (one rounding per operation)
        const float a2 = (float)PA[8 + j], b2 = (float)PB[8 + j];
        const float h0 = (float)xa[0] * a2, h1 = (float)xa[1] * a2, h2 = (float)xb[0] * b2, h3 = (float)xb[1] * b2;
        H[0 * 4 + j] = (double)(h0 - (float)PA[j]);
        H[1 * 4 + j] = (double)(h1 - (float)PA[4 + j]);
        H[2 * 4 + j] = (double)(h2 - (float)PB[j]);
        H[3 * 4 + j] = (double)(h3 - (float)PB[4 + j]);
    }
    double G[16], E[16];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            double g = 0.0;
            for (int k = 0; k < 4; k++) g += H[k * 4 + i] * H[k * 4 + j];
            G[i * 4 + j] = g;
        }
    jacobi_eigen<4>(G, E);
    const int m = argmin_diag<4>(G);
    for (int r = 0; r < 4; r++) X[r] = E[r * 4 + m];
}

}  // namespace sm
}  // namespace tpose

#endif
