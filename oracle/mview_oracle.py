"""numpy restatement of the multiview half of the hot path (source/multiview.hpp of the reference).

TEST INFRASTRUCTURE ONLY -- imported by tests/ and by tests/golden/make_golden.py, never by the
product.

The reference computes with Eigen (JacobiSVD<MatrixXf>, EigenSolver<MatrixXd>) in float32, which is
not available here, so this restatement keeps the reference's float32 *data* (normalised points and
the rows of the design matrix are rounded to float32 exactly where the reference stores floats) and
does the decompositions in float64 LAPACK.  PARITY STATUS: "parity unpinned" against Eigen itself;
pinned against cv2.findFundamentalMat(FM_8POINT) on the reference's own match list
(tests/sfm_match_test/data.txt, 4e-5 element-wise) -- see tests/golden/make_golden.py.

Conventions this file fixes (Eigen leaves them to its implementation; the CUDA path follows these):
  * singular vectors: U and V of a 3x3 SVD are made proper rotations (det = +1, third columns =
    cross products of the first two) and, of the two joint sign choices that leaves, the one where
    the largest-magnitude component of V.col(2) is positive.  GetPose (multiview.hpp:398-409 takes
    t = U.col(2); check = 3 picks [R2 | -t]) and the buggy root selection depend on it;
  * real roots of the sextic are taken in ascending order before the selection loop of
    multiview.hpp:492-501, which is restated literally INCLUDING its bug (cost evaluated at the loop
    index r, not at the root R[r]).
"""
import numpy as np

f32 = np.float32


def _seqsum32(v):
    """float32 running sum in index order (what `for(p) c += p` does)."""
    return np.cumsum(np.asarray(v, f32), dtype=f32)[-1]


def camera():
    """multiview.hpp:35-52 (note: 1/fx on the diagonal, as coded)."""
    # `float px = 488.421 / 960.0f;` -- a double division rounded to float once
    px = f32(488.421 / 960.0); py = f32(268.8 / 960.0)
    fx = f32(673.101 / 960.0); fy = f32(673.328 / 960.0)
    return np.array([[f32(1.0) / fx, 0, px], [0, f32(1.0) / fy, py], [0, 0, 1]], f32)


def normalize(points):
    """multiview.hpp:62-87.  Returns (normalised float32 points, H float32)."""
    p = np.asarray(points, f32).copy()
    n = f32(len(p))
    c = np.array([_seqsum32(p[:, 0]) / n, _seqsum32(p[:, 1]) / n], f32)
    p = (p - c).astype(f32)
    ln = np.sqrt((p[:, 0] * p[:, 0] + p[:, 1] * p[:, 1]).astype(f32)).astype(f32)
    dist = f32(_seqsum32(ln) / n)
    scale = f32(np.sqrt(2.0) / np.float64(dist))
    p = (p * scale).astype(f32)
    H = np.array([[scale, 0, f32(-c[0] * scale)], [0, scale, f32(-c[1] * scale)], [0, 0, 1]], f32)
    return p, H


def design_matrix(pa, pb, w=None):
    """multiview.hpp:150-162 (and :208-220 with weights): float32 rows."""
    xa, ya, xb, yb = pa[:, 0], pa[:, 1], pb[:, 0], pb[:, 1]
    one = np.ones_like(xa)
    A = np.stack([xa * xb, ya * xb, xb, xa * yb, ya * yb, yb, xa, ya, one], 1).astype(f32)
    if w is not None:
        A = (np.asarray(w, f32)[:, None] * A).astype(f32)
    return A


def _rank2(F):
    U, S, Vt = np.linalg.svd(F)
    S[2] = 0.0
    return U @ np.diag(S) @ Vt


def f_8point(pa, pb):
    """multiview.hpp:130-183.  float64 result (compare after casting to float32)."""
    pa = np.asarray(pa, f32); pb = np.asarray(pb, f32)
    if len(pa) != len(pb) or len(pa) == 0:
        return np.eye(3)
    na, HA = normalize(pa)
    nb, HB = normalize(pb)
    A = design_matrix(na, nb).astype(np.float64)
    _, _, Vt = np.linalg.svd(A, full_matrices=True)
    F = Vt[8].reshape(3, 3)
    F = _rank2(F)
    F = HB.astype(np.float64).T @ F @ HA.astype(np.float64)
    return F / F[2, 2]


def svd3_proper(M):
    """3x3 SVD with det(U) = det(V) = +1 (sign of the third columns chosen)."""
    U, S, Vt = np.linalg.svd(np.asarray(M, np.float64))
    V = Vt.T.copy()
    U = U.copy()
    U[:, 2] = np.cross(U[:, 0], U[:, 1])
    V[:, 2] = np.cross(V[:, 0], V[:, 1])
    if V[np.argmax(np.abs(V[:, 2])), 2] < 0:
        for Q in (U, V):
            Q[:, 0] = -Q[:, 0]
            Q[:, 2] = -Q[:, 2]
    return U, S, V


def hdlt(PA, PB, xa, xb):
    """multiview.hpp:370-380: last right singular vector of the 4x4 system (float32 rows)."""
    PA = np.asarray(PA, f32); PB = np.asarray(PB, f32)
    H = np.stack([f32(xa[0]) * PA[2] - PA[0], f32(xa[1]) * PA[2] - PA[1],
                  f32(xb[0]) * PB[2] - PB[0], f32(xb[1]) * PB[2] - PB[1]]).astype(f32)
    _, _, Vt = np.linalg.svd(H.astype(np.float64))
    return Vt[3]


def get_pose(E):
    """multiview.hpp:391-411 with the det=+1 convention above.  Returns R1, R2, t."""
    W = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1]], np.float64)
    U, _, V = svd3_proper(E)
    return U @ W @ V.T, U @ W.T @ V.T, U[:, 2].copy()


def sextic_coefficients(a, b, c, d, m, n):
    """multiview.hpp:478-484 (doubles)."""
    a0 = b*b*c*d - a*b*d*d
    a1 = b*b*b*b + (b*b*c*c - a*a*d*d) + 2.0*b*b*d*d*n*n + d*d*d*d*n*n*n*n
    a2 = (a*b*c*c - a*a*c*d) + 4.0*a*b*b*b + 2.0*(b*b*c*d - a*b*d*d)*m*m + 4.0*(a*b*d*d + b*b*c*d)*n*n + 4.0*c*d*d*d*n*n*n*n
    a3 = 6.0*a*a*b*b + 2.0*(b*b*c*c - a*a*d*d)*m*m + 2.0*a*a*d*d*n*n + 8.0*a*b*c*d*n*n + 2.0*b*b*c*c*n*n + 6.0*c*c*d*d*n*n*n*n
    a4 = (b*b*c*d - a*b*d*d)*m*m*m*m + 4.0*a*a*a*b + 2.0*(a*b*c*c - a*a*c*d)*m*m + 4.0*(a*a*c*d + a*b*c*c)*n*n + 4.0*c*c*c*d*n*n*n*n
    a5 = a*a*a*a + (b*b*c*c - a*a*d*d)*m*m*m*m + 2.0*a*a*c*c*n*n + c*c*c*c*n*n*n*n
    a6 = (a*b*c*c - a*a*c*d)*m*m*m*m
    return np.array([a0, a1, a2, a3, a4, a5, a6], np.float64)


def horner(x, a):
    r = a[-1]
    for k in range(len(a) - 2, -1, -1):
        r = r * x + a[k]
    return r


def realroots(a):
    """utility.hpp:104-138: companion-matrix eigenvalues with exactly zero imaginary part, polished
    by 25 Newton steps; returned in ascending order (our convention)."""
    K = len(a) - 1
    Cm = np.zeros((K, K))
    Cm[:, K - 1] = -a[:K] / a[K]
    for k in range(1, K):
        Cm[k, k - 1] = 1.0
    ev = np.linalg.eigvals(Cm)
    rr = sorted(float(z.real) for z in ev if z.imag == 0)
    d = np.array([a[1] * 1.0, a[2] * 2.0, a[3] * 3.0, a[4] * 4.0, a[5] * 5.0, a[6] * 6.0])
    out = []
    for r in rr:
        for _ in range(25):
            r -= horner(r, a) / horner(r, d)
        out.append(r)
    return out


def correct_match(F, A, B):
    """multiview.hpp:415-527: Hartley-Sturm optimal correction of one match; float32 matrices where
    the reference holds Matrix3f, doubles inside the polynomial.  Returns corrected (A, B)."""
    F = np.asarray(F, f32).astype(np.float64)
    TAi = np.array([[1, 0, A[0]], [0, 1, A[1]], [0, 0, 1]], np.float64)     # TA.inverse()
    TBi = np.array([[1, 0, B[0]], [0, 1, B[1]], [0, 0, 1]], np.float64)
    F = (TBi.T @ F @ TAi).astype(f32).astype(np.float64)                     # TB^-T F TA^-1
    U, _, V = svd3_proper(F)
    eA = V[:, 2] / np.sqrt(V[0, 2] ** 2 + V[1, 2] ** 2)
    eB = U[:, 2] / np.sqrt(U[0, 2] ** 2 + U[1, 2] ** 2)
    RA = np.array([[eA[0], eA[1], 0], [-eA[1], eA[0], 0], [0, 0, 1]])
    RB = np.array([[eB[0], eB[1], 0], [-eB[1], eB[0], 0], [0, 0, 1]])
    F = RB @ F @ RA.T
    m, n = eA[2], eB[2]
    a, b, c, d = F[1, 1], F[1, 2], F[2, 1], F[2, 2]

    def S(t):
        return t * t / (1.0 + m * m * t * t) + (c * t + d) ** 2 / ((a * t + b) ** 2 + n * n * (c * t + d) ** 2)

    R = realroots(sextic_coefficients(a, b, c, d, m, n))
    t = R[0]
    minerr = S(R[0])
    for r in range(1, len(R)):
        cur = S(float(r))                    # sic: multiview.hpp:496 evaluates at the index
        if cur < minerr:
            t, minerr = R[r], cur
    LA = np.array([t * m, 1.0, -t])
    LB = np.array([-n * (c * t + d), a * t + b, c * t + d])
    XA = np.array([-LA[0] * LA[2], -LA[1] * LA[2], LA[0] ** 2 + LA[1] ** 2])
    XB = np.array([-LB[0] * LB[2], -LB[1] * LB[2], LB[0] ** 2 + LB[1] ** 2])
    XA = TAi @ RA.T @ XA
    XB = TBi @ RB.T @ XB
    return XA[:2] / XA[2], XB[:2] / XB[2]


def triangulate(F, K, A, B, correct=True):
    """multiview.hpp:529-627: corrected matches -> E = K^T F K -> pose (check = 3: [R2 | -t]) ->
    HDLT with K*PA, K*PB -> X / X(3).  Returns (N,4) float64."""
    F = np.asarray(F, f32).astype(np.float64)
    K = np.asarray(K, f32).astype(np.float64)
    A = np.asarray(A, f32).astype(np.float64).copy()
    B = np.asarray(B, f32).astype(np.float64).copy()
    if correct:
        for i in range(len(A)):
            A[i], B[i] = correct_match(F, A[i], B[i])
        A = A.astype(f32).astype(np.float64); B = B.astype(f32).astype(np.float64)   # stored back into vec2
    E = K.T @ F @ K
    R1, R2, t = get_pose(E)
    PA = np.hstack([np.eye(3), np.zeros((3, 1))])
    PB = np.hstack([R2, -t[:, None]])
    KPA, KPB = (K @ PA).astype(f32), (K @ PB).astype(f32)
    out = np.zeros((len(A), 4))
    for i in range(len(A)):
        X = hdlt(KPA, KPB, A[i], B[i])
        out[i] = X / X[3]
    return out


def eline(F, p):
    """multiview.hpp:108-113: l = F (p,1) in float32 (glm mat3*vec3: products summed left to right),
    divided by its z.  utility.hpp:146-153 (fromEigen3) is the identity on the mathematical matrix."""
    F = np.asarray(F, f32)
    x, y = f32(p[0]), f32(p[1])
    l = np.array([f32(f32(f32(F[i, 0] * x) + f32(F[i, 1] * y)) + F[i, 2]) for i in range(3)], f32)
    return (l / l[2]).astype(f32)


def f_sampson(pa, pb, w=None, iters=100):
    """multiview.hpp:187-242 (and :244-299 with per-match weights w): 100 re-weighted 8-point solves.
    Kept literally, including the first iteration's mix of the DE-normalised initial F with the
    normalised points (:189-205).  float32 where the reference stores floats, float64 LAPACK for
    the decompositions.  Returns float64 (compare after casting to float32)."""
    pa = np.asarray(pa, f32); pb = np.asarray(pb, f32)
    F = f_8point(pa, pb).astype(f32)
    if len(pa) != len(pb) or len(pa) == 0:
        return F.astype(np.float64)
    na, HA = normalize(pa)
    nb, HB = normalize(pb)
    n = len(na)
    for _ in range(iters):
        Ft = F.T.copy()
        # vectorised eline for all matches (same float32 operation order as eline above)
        def lines(M, P):
            out = np.empty((n, 3), f32)
            for i in range(3):
                out[:, i] = ((M[i, 0] * P[:, 0]).astype(f32) + (M[i, 1] * P[:, 1]).astype(f32)).astype(f32) + M[i, 2]
            return (out / out[:, 2:3]).astype(f32)
        L = lines(Ft, nb)
        R = lines(F, na)
        den = (L[:, 0] * L[:, 0]).astype(f32)
        den = (den + (L[:, 1] * L[:, 1]).astype(f32)).astype(f32)
        den = (den + (R[:, 0] * R[:, 0]).astype(f32)).astype(f32)
        den = (den + (R[:, 1] * R[:, 1]).astype(f32)).astype(f32)
        W = (f32(1.0) / den).astype(f32)
        if w is not None:
            W = (np.asarray(w, f32) * W).astype(f32)
        xa, ya, xb, yb = na[:, 0], na[:, 1], nb[:, 0], nb[:, 1]
        cols = [((W * xa).astype(f32) * xb).astype(f32), ((W * ya).astype(f32) * xb).astype(f32), (W * xb).astype(f32),
                ((W * xa).astype(f32) * yb).astype(f32), ((W * ya).astype(f32) * yb).astype(f32), (W * yb).astype(f32),
                (W * xa).astype(f32), (W * ya).astype(f32), W]
        A = np.stack(cols, 1).astype(np.float64)
        _, _, Vt = np.linalg.svd(A, full_matrices=True)
        F = _rank2(Vt[8].astype(f32).astype(np.float64).reshape(3, 3)).astype(f32)
    Fd = HB.astype(np.float64).T @ F.astype(np.float64) @ HA.astype(np.float64)
    return Fd / Fd[2, 2]


def select_matches(originA, pointsA, warp_originA, pointsB, originB, warp_originB, ratio, max_dist=0.0000002):
    """tests/compute_fundamental_mat/model.h:1-82 (commented out in the reference), loop for loop:
    returns (matchX, matchY, weights, dist2D) -- dist2D for all vertices of A then of B."""
    R = f32(ratio)
    cx, cy = f32(0.5) / R, f32(-0.5) / R                   # T, a row-major literal in a column-major glm::mat3 (model.h:7-11)

    def T(p):
        return np.array([(cx * p[0] + f32(0.0) * p[1]) + f32(0.0), (f32(0.0) * p[0] + cy * p[1]) + f32(0.0)], f32)

    def boundary(p):                                       # triangulation.hpp:132-140
        return p[0] <= -R or p[1] <= f32(-1) or p[0] >= R or p[1] >= f32(1)

    def central(p):
        return not (float(p[0]) < -float(R) / 2.0 or float(p[0]) > float(R) / 2.0 or float(p[1]) < -0.5 or float(p[1]) > 0.5)

    X, Y, W, D = [], [], [], []
    for (first, second, worig, a_side) in ((originA, pointsA, warp_originA, True), (pointsB, originB, warp_originB, False)):
        first = np.asarray(first, f32).reshape(-1, 2); second = np.asarray(second, f32).reshape(-1, 2)
        worig = np.asarray(worig, f32).reshape(-1, 2)
        pts = second if a_side else first                  # tr.points of this view
        for i in range(len(first)):
            dv = (worig[i] - pts[i]).astype(f32)
            d = f32(np.sqrt(f32(f32(dv[0] * dv[0]) + f32(dv[1] * dv[1]))))
            D.append(d)
            pX, pY = first[i], second[i]
            if float(d) > max_dist:
                continue
            if boundary(pX) or boundary(pY):
                continue
            if not central(pX) or not central(pY):
                continue
            X.append(T(pX)); Y.append(T(pY))
            with np.errstate(divide="ignore"):
                W.append(f32(1.0) / d)
    return (np.array(X, f32).reshape(-1, 2), np.array(Y, f32).reshape(-1, 2), np.array(W, f32), np.array(D, f32))
