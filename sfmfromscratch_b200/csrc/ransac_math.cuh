// ransac_math.cuh -- float64 building blocks of the two-view RANSAC kernels (ransac.cu).
// Host+device so that tests/native/ransac_host_check.cu can run the very same arithmetic on the
// CPU against the oracle where no GPU exists; the library itself only calls them from kernels.
#pragma once
#include <cmath>
#include <cuda_runtime.h>

#define SFM_HD __host__ __device__

// Separately rounded operations (numpy evaluates these expressions without contraction).
SFM_HD __forceinline__ double rmul(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    volatile double r = a * b; return r;
#endif
}
SFM_HD __forceinline__ double radd(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    volatile double r = a + b; return r;
#endif
}
SFM_HD __forceinline__ double rsub(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dsub_rn(a, b);
#else
    volatile double r = a - b; return r;
#endif
}
SFM_HD __forceinline__ double rsqrt_d(double a) {
#ifdef __CUDA_ARCH__
    return rsqrt(a);
#else
    return 1.0 / sqrt(a);
#endif
}

// One-sided (Hestenes) Jacobi SVD of an NxN matrix held column-wise as G[r][c]:
// on return the columns of G are sigma_i * u_i and the columns of V the right
// singular vectors (G_in = G_out * V^T).  High relative accuracy, no sorting.
template <int N>
SFM_HD __forceinline__ void jacobi_svd(double (&G)[N][N], double (&V)[N][N]) {
#pragma unroll
    for (int r = 0; r < N; ++r)
#pragma unroll
        for (int c = 0; c < N; ++c) V[r][c] = (r == c) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 40; ++sweep) {
        bool rotated = false;
#pragma unroll
        for (int p = 0; p < N - 1; ++p) {
#pragma unroll
            for (int q = p + 1; q < N; ++q) {
                double a = 0.0, b = 0.0, g = 0.0;
#pragma unroll
                for (int r = 0; r < N; ++r) {
                    a = fma(G[r][p], G[r][p], a);
                    b = fma(G[r][q], G[r][q], b);
                    g = fma(G[r][p], G[r][q], g);
                }
                if (g != 0.0 && fabs(g) > 1e-16 * sqrt(a * b)) {
                    rotated = true;
                    const double zeta = (b - a) / (2.0 * g);
                    const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(fma(zeta, zeta, 1.0)));
                    const double c = rsqrt_d(fma(t, t, 1.0));
                    const double s = c * t;
#pragma unroll
                    for (int r = 0; r < N; ++r) {
                        const double gp = G[r][p], gq = G[r][q];
                        G[r][p] = fma(c, gp, -s * gq);
                        G[r][q] = fma(s, gp, c * gq);
                        const double vp = V[r][p], vq = V[r][q];
                        V[r][p] = fma(c, vp, -s * vq);
                        V[r][q] = fma(s, vp, c * vq);
                    }
                }
            }
        }
        if (!rotated) break;
    }
}

SFM_HD __forceinline__ void mat3_mul(const double* A, const double* B, double* C) {   // C = A B, row-major
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            C[r * 3 + c] = fma(A[r * 3 + 2], B[6 + c], fma(A[r * 3 + 1], B[3 + c], A[r * 3] * B[c]));
}

SFM_HD __forceinline__ double det3(const double* A) {
    return A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) + A[2] * (A[3] * A[7] - A[4] * A[6]);
}

// SFM.py:163-178 for the 8 sampled points: returns the normalised x, y and the
// similarity (s, tx, ty) with T = [[s,0,tx],[0,s,ty],[0,0,1]].
SFM_HD __forceinline__ void normalize8(const double* x, const double* y, double* xn, double* yn,
                                           double& s, double& tx, double& ty) {
    double sx = x[0], sy = y[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) { sx = radd(sx, x[i]); sy = radd(sy, y[i]); }
    const double cx = sx * 0.125, cy = sy * 0.125;
    double d[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const double dx = rsub(x[i], cx), dy = rsub(y[i], cy);
        d[i] = sqrt(radd(rmul(dx, dx), rmul(dy, dy)));
    }
    const double md = radd(radd(radd(d[0], d[1]), radd(d[2], d[3])), radd(radd(d[4], d[5]), radd(d[6], d[7]))) * 0.125;
    s = 1.4142135623730951 / md;
    tx = rmul(-s, cx);
    ty = rmul(-s, cy);
#pragma unroll
    for (int i = 0; i < 8; ++i) { xn[i] = radd(rmul(x[i], s), tx); yn[i] = radd(rmul(y[i], s), ty); }
}

// SFM.py:189-236.
// `degenerate` (may be null) is set when the 8x9 design matrix is numerically rank-deficient (repeated or
// co-motion-free correspondences): its null space then has more than one dimension and the vector LAPACK
// returns for it -- hence the reference's F for this sample -- is decided by rounding noise.
SFM_HD void fundamental_8pt(const double* x1, const double* y1, const double* x2, const double* y2, double* F,
                            bool* degenerate = nullptr) {
    double a1[8], b1[8], a2[8], b2[8], s1, t1x, t1y, s2, t2x, t2y;
    normalize8(x1, y1, a1, b1, s1, t1x, t1y);
    normalize8(x2, y2, a2, b2, s2, t2x, t2y);
    // B = A^T, 9 x 8: column j is the j-th correspondence's row of the design matrix
    double B[9][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        B[0][j] = rmul(a1[j], a2[j]); B[1][j] = rmul(b1[j], a2[j]); B[2][j] = a2[j];
        B[3][j] = rmul(a1[j], b2[j]); B[4][j] = rmul(b1[j], b2[j]); B[5][j] = b2[j];
        B[6][j] = a1[j];              B[7][j] = b1[j];              B[8][j] = 1.0;
    }
    // Householder QR of B; the last column of Q spans the null space of A (np.linalg.svd(A)[2][-1]
    // up to sign when A has rank 8).
    double tau[8];
    double dmin = INFINITY, dmax = 0.0;                     // |diagonal of R|: its spread measures the rank
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        double nrm2 = 0.0;
#pragma unroll
        for (int r = j; r < 9; ++r) nrm2 = fma(B[r][j], B[r][j], nrm2);
        const double nrm = sqrt(nrm2);
        dmin = fmin(dmin, nrm); dmax = fmax(dmax, nrm);
        if (nrm == 0.0) { tau[j] = 0.0; continue; }
        const double alpha = B[j][j];
        const double beta = -copysign(nrm, alpha);
        const double inv = 1.0 / (alpha - beta);
        tau[j] = (beta - alpha) / beta;
#pragma unroll
        for (int r = j + 1; r < 9; ++r) B[r][j] *= inv;
#pragma unroll
        for (int c = j + 1; c < 8; ++c) {
            double w = B[j][c];
#pragma unroll
            for (int r = j + 1; r < 9; ++r) w = fma(B[r][j], B[r][c], w);
            w *= tau[j];
            B[j][c] -= w;
#pragma unroll
            for (int r = j + 1; r < 9; ++r) B[r][c] = fma(-w, B[r][j], B[r][c]);
        }
    }
    if (degenerate) *degenerate = !(dmin > 1e-10 * dmax);
    double z[9] = {0, 0, 0, 0, 0, 0, 0, 0, 1.0};
#pragma unroll
    for (int j = 7; j >= 0; --j) {
        double w = z[j];
#pragma unroll
        for (int r = j + 1; r < 9; ++r) w = fma(B[r][j], z[r], w);
        w *= tau[j];
        z[j] -= w;
#pragma unroll
        for (int r = j + 1; r < 9; ++r) z[r] = fma(-w, B[r][j], z[r]);
    }
    // rank-2 projection: drop the smallest singular triplet
    double G[3][3], V[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) G[r][c] = z[r * 3 + c];
    jacobi_svd<3>(G, V);
    double n2[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) n2[c] = G[0][c] * G[0][c] + G[1][c] * G[1][c] + G[2][c] * G[2][c];
    const int k = (n2[0] <= n2[1] && n2[0] <= n2[2]) ? 0 : ((n2[1] <= n2[2]) ? 1 : 2);
    double F2[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            double v = 0.0;
#pragma unroll
            for (int i = 0; i < 3; ++i) v = (i == k) ? v : fma(G[r][i], V[c][i], v);
            F2[r * 3 + c] = v;
        }
    // T2^T F2 T1
    double M[9];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        M[c] = s2 * F2[c];
        M[3 + c] = s2 * F2[3 + c];
        M[6 + c] = fma(t2x, F2[c], fma(t2y, F2[3 + c], F2[6 + c]));
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        F[r * 3] = M[r * 3] * s1;
        F[r * 3 + 1] = M[r * 3 + 1] * s1;
        F[r * 3 + 2] = fma(M[r * 3], t1x, fma(M[r * 3 + 1], t1y, M[r * 3 + 2]));
    }
}

// SFM.py:56-80: the four (R, T) readings of E = K2^T F K1.  The set is the reference's; the
// ORDER is canonical (Ra = M + N, Rb = M - N with M = u1 v0^T - u0 v1^T, N = u2 v2^T, each
// negated if its determinant is negative; T = u2), whereas the reference's order follows the
// signs LAPACK happens to give the singular vectors -- the host maps one onto the other.
SFM_HD void pose_candidates(const double* F, const double* K1, const double* K2, double* cand) {
    double K2t[9], tmp[9], E[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) K2t[r * 3 + c] = K2[c * 3 + r];
    mat3_mul(K2t, F, tmp);
    mat3_mul(tmp, K1, E);
    double G[3][3], V[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) G[r][c] = E[r * 3 + c];
    jacobi_svd<3>(G, V);
    double n2[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) n2[c] = G[0][c] * G[0][c] + G[1][c] * G[1][c] + G[2][c] * G[2][c];
    int i0 = 0, i1 = 1, i2 = 2;
    if (n2[i0] < n2[i1]) { int t = i0; i0 = i1; i1 = t; }
    if (n2[i1] < n2[i2]) { int t = i1; i1 = i2; i2 = t; }
    if (n2[i0] < n2[i1]) { int t = i0; i0 = i1; i1 = t; }
    double u0[3], u1[3], v0[3], v1[3];
    const double r0 = 1.0 / sqrt(n2[i0]), r1 = 1.0 / sqrt(n2[i1]);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        // dynamic column pick without dynamic indexing
        const double g0 = (i0 == 0) ? G[r][0] : (i0 == 1 ? G[r][1] : G[r][2]);
        const double g1 = (i1 == 0) ? G[r][0] : (i1 == 1 ? G[r][1] : G[r][2]);
        u0[r] = g0 * r0; u1[r] = g1 * r1;
        v0[r] = (i0 == 0) ? V[r][0] : (i0 == 1 ? V[r][1] : V[r][2]);
        v1[r] = (i1 == 0) ? V[r][0] : (i1 == 1 ? V[r][1] : V[r][2]);
    }
    const double u2[3] = {u0[1] * u1[2] - u0[2] * u1[1], u0[2] * u1[0] - u0[0] * u1[2], u0[0] * u1[1] - u0[1] * u1[0]};
    const double v2[3] = {v0[1] * v1[2] - v0[2] * v1[1], v0[2] * v1[0] - v0[0] * v1[2], v0[0] * v1[1] - v0[1] * v1[0]};
    double Ra[9], Rb[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double m = u1[r] * v0[c] - u0[r] * v1[c], nn = u2[r] * v2[c];
            Ra[r * 3 + c] = m + nn;
            Rb[r * 3 + c] = m - nn;
        }
    const double sa = det3(Ra) < 0 ? -1.0 : 1.0, sb = det3(Rb) < 0 ? -1.0 : 1.0;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        cand[i] = cand[12 + i] = sa * Ra[i];
        cand[24 + i] = cand[36 + i] = sb * Rb[i];
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        cand[9 + i] = cand[33 + i] = u2[i];
        cand[21 + i] = cand[45 + i] = -u2[i];
    }
}

// SFM.py:143-153 / :82-95: the distance of p2 from the epipolar line F p1.
SFM_HD __forceinline__ bool is_inlier(const double* F, double2 a, double2 b, double thr) {
    const double l0 = fma(F[2], 1.0, fma(F[1], a.y, rmul(F[0], a.x)));
    const double l1 = fma(F[5], 1.0, fma(F[4], a.y, rmul(F[3], a.x)));
    const double l2 = fma(F[8], 1.0, fma(F[7], a.y, rmul(F[6], a.x)));
    const double num = fabs(radd(radd(rmul(l0, b.x), rmul(l1, b.y)), l2));
    const double den = sqrt(radd(rmul(l0, l0), rmul(l1, l1)));
    return (num / den) < thr;
}


// SFM.py:308-309: P = K @ [R | T], row-major 3x4.
SFM_HD __forceinline__ void projection3x4(const double* K, const double* R, const double* T, double* P) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int q = 0; q < 3; ++q)
            P[r * 4 + q] = fma(K[r * 3 + 2], R[6 + q], fma(K[r * 3 + 1], R[3 + q], K[r * 3] * R[q]));
        P[r * 4 + 3] = fma(K[r * 3 + 2], T[2], fma(K[r * 3 + 1], T[1], K[r * 3] * T[0]));
    }
}

// SFM.py:104-124 for one correspondence: DLT triangulation (:239-253, smallest right singular
// vector of the 4x4 system) and the depth test in both cameras.  NaN depths pass, as they do in
// the reference's `<` comparisons.
SFM_HD __forceinline__ bool point_in_front(const double* P1, const double* P2, const double* Rb, const double* Tb,
                                           const double* Rc, const double* Tc, double2 a, double2 b) {
    double G[4][4], V[4][4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        G[0][q] = rsub(rmul(a.x, P1[8 + q]), P1[q]);
        G[1][q] = rsub(rmul(a.y, P1[8 + q]), P1[4 + q]);
        G[2][q] = rsub(rmul(b.x, P2[8 + q]), P2[q]);
        G[3][q] = rsub(rmul(b.y, P2[8 + q]), P2[4 + q]);
    }
    jacobi_svd<4>(G, V);
    double n2[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) n2[q] = G[0][q] * G[0][q] + G[1][q] * G[1][q] + G[2][q] * G[2][q] + G[3][q] * G[3][q];
    int k = 0;
#pragma unroll
    for (int q = 1; q < 4; ++q) if (n2[q] < n2[k]) k = q;
    double X[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) X[r] = (k == 0) ? V[r][0] : (k == 1 ? V[r][1] : (k == 2 ? V[r][2] : V[r][3]));
    const double X0 = X[0] / X[3], X1 = X[1] / X[3], X2 = X[2] / X[3];
    const double zb = fma(Rb[8], X2, fma(Rb[7], X1, Rb[6] * X0)) + Tb[2];
    const double zc = fma(Rc[8], X2, fma(Rc[7], X1, Rc[6] * X0)) + Tc[2];
    return !((zb < 1e-6) || (zc < 1e-6));
}
