"""Regenerates tests/golden/sfm_matches.npz from the reference's own match list
(/root/reference/tests/sfm_match_test/data.txt, 174 hand matches of imageA/imageB, read by
io::readmatches and divided by 1200 -- tests/sfm_match_test/main.cpp:26-30).

Run in the build container (needs /root/reference); the .npz travels to the GPU box.
Contents: A, B  float32 (174,2)   the matches as the reference feeds them to multiview.hpp
          F     float64 (3,3)     oracle/mview_oracle.f_8point(A, B)
          F_cv  float64 (3,3)     cv2.findFundamentalMat(FM_8POINT) on the same matches (the pin)
          K     float32 (3,3)     mview::Camera()
          X, Ac, Bc               oracle triangulate(F,K,A,B) and the corrected matches
          X_nc                    the same without the optimal correction
          F_sampson float64 (3,3) oracle/mview_oracle.f_sampson(A, B) (100 iterations)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import mview_oracle as M  # noqa: E402


def main():
    import cv2
    d = np.loadtxt("/root/reference/tests/sfm_match_test/data.txt").astype(np.float32)
    A = (d[:, :2] / np.float32(1200)).astype(np.float32)
    B = (d[:, 2:] / np.float32(1200)).astype(np.float32)
    F = M.f_8point(A, B)
    Fcv, _ = cv2.findFundamentalMat(A.astype(np.float64), B.astype(np.float64), cv2.FM_8POINT)
    K = M.camera()
    Ff = F.astype(np.float32)
    Ac, Bc = A.astype(np.float64).copy(), B.astype(np.float64).copy()
    for i in range(len(A)):
        Ac[i], Bc[i] = M.correct_match(Ff, Ac[i], Bc[i])
    X = M.triangulate(Ff, K, A, B)
    Xnc = M.triangulate(Ff, K, A, B, correct=False)
    Fs = M.f_sampson(A, B)
    np.savez(os.path.join(HERE, "sfm_matches.npz"), A=A, B=B, F=F, F_cv=Fcv, K=K, X=X, X_nc=Xnc, Ac=Ac, Bc=Bc, F_sampson=Fs)
    print("max |F - F_cv| =", np.abs(F - Fcv).max())


if __name__ == "__main__":
    main()
