// Per-observation math of the stereo local-map builder: reprojection residuals and
// Jacobians of one landmark seen by two consecutive stereo frames, and its blocks of the normal
// equations.  Plain C++ (host + device): the CUDA kernel (builder.cu) and the CPU unit test
// (tests/helpers/builder_math_host.cpp, compiled with g++) include the same file.
//
// There is no reference code for this step (LinearSFM starts from finished localmap_*.txt files);
// the conventions are fixed by the reference's geometry (LinearSFMImp.cpp:132-143: a point maps as
// X_cam = R (X - t), R = Rx(gamma) Ry(beta) Rz(alpha), passive) and by SURVEY Appendix A:
//   map frame = first stereo frame; state = pose of the second frame (t, alpha, beta, gamma) + the
//   landmarks in the map frame; camera looks along +x, y left, z up; right camera at y = -baseline;
//   measurement (uL, vL, uR) = (cx - f y/x, cy - f z/x, cx - f (y+b)/x), isotropic noise sigma.
#pragma once
#include "geom.cuh"

#ifdef __CUDACC__
#define LSFM_HD __host__ __device__ inline
#else
#define LSFM_HD inline
#endif

namespace bld {

struct Cam { double f, b, cx, cy, w; };        // w = 1 / sigma^2

// pose-dependent part, once per map and iteration
struct PoseLin {
    double t[3];
    double R[9], dA[9], dB[9], dG[9];           // R(alpha,beta,gamma) and its partials, row-major
};

LSFM_HD void pose_lin(const double *pose, PoseLin &P)
{
    P.t[0] = pose[0]; P.t[1] = pose[1]; P.t[2] = pose[2];
    geom::rot_derivs(pose[3], pose[4], pose[5], P.R, P.dA, P.dB, P.dG);
}

LSFM_HD void project(const Cam &c, const double *X, double *z)
{
    const double ix = 1.0 / X[0];
    z[0] = c.cx - c.f * X[1] * ix;
    z[1] = c.cy - c.f * X[2] * ix;
    z[2] = c.cx - c.f * (X[1] + c.b) * ix;
}

// d(uL, vL, uR) / d(x, y, z), row-major 3x3
LSFM_HD void project_jac(const Cam &c, const double *X, double *J)
{
    const double ix = 1.0 / X[0], f = c.f;
    J[0] = f * X[1] * ix * ix;          J[1] = -f * ix; J[2] = 0.0;
    J[3] = f * X[2] * ix * ix;          J[4] = 0.0;     J[5] = -f * ix;
    J[6] = f * (X[1] + c.b) * ix * ix;  J[7] = -f * ix; J[8] = 0.0;
}

// landmark from its stereo observation in the map frame (initial guess)
LSFM_HD void triangulate(const Cam &c, const double *z, double *X)
{
    // disparity floored at that of a landmark 500 baselines away: pixel noise can push
    // the disparity of a distant landmark to zero or below
    const double dmin = c.f / 500.0;
    const double disp = (z[0] - z[2]) > dmin ? (z[0] - z[2]) : dmin;
    const double x = c.f * c.b / disp;
    X[0] = x;
    X[1] = (c.cx - z[0]) * x / c.f;
    X[2] = (c.cy - z[1]) * x / c.f;
}

// Blocks of one landmark at the linearisation point (P, X):
//   V  (3x3)  = w (J0^T J0 + JX1^T JX1)         gF (3) = w (J0^T r0 + JX1^T r1)
//   W  (6x3)  = w JP1^T JX1                      gP (6) = w JP1^T r1
//   Ub (6x6)  = w JP1^T JP1
// with r = z - h(x), J0 = dh/dX in the map frame, JX1 = dh/dXc R, JP1 = [-JX1 | dh/dXc dR (X - t)].
// Returns the squared weighted residual w (|r0|^2 + |r1|^2).
LSFM_HD double feature_blocks(const Cam &c, const PoseLin &P, const double *X, const double *z0,
                              const double *z1, double *V, double *W, double *Ub, double *gF, double *gP)
{
    double h0[3], h1[3], J0[9], Jp1[9], JX1[9], JP1[18];
    project(c, X, h0);
    project_jac(c, X, J0);
    const double d[3] = {X[0] - P.t[0], X[1] - P.t[1], X[2] - P.t[2]};
    double Xc[3], a[3], b[3], g[3];
    geom::mat3_vec(P.R, d, Xc);
    geom::mat3_vec(P.dA, d, a);
    geom::mat3_vec(P.dB, d, b);
    geom::mat3_vec(P.dG, d, g);
    project(c, Xc, h1);
    project_jac(c, Xc, Jp1);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0.0;
            for (int k = 0; k < 3; k++) s += Jp1[3 * i + k] * P.R[3 * k + j];
            JX1[3 * i + j] = s;
        }
    for (int i = 0; i < 3; i++) {
        JP1[6 * i + 0] = -JX1[3 * i + 0]; JP1[6 * i + 1] = -JX1[3 * i + 1]; JP1[6 * i + 2] = -JX1[3 * i + 2];
        JP1[6 * i + 3] = Jp1[3 * i] * a[0] + Jp1[3 * i + 1] * a[1] + Jp1[3 * i + 2] * a[2];
        JP1[6 * i + 4] = Jp1[3 * i] * b[0] + Jp1[3 * i + 1] * b[1] + Jp1[3 * i + 2] * b[2];
        JP1[6 * i + 5] = Jp1[3 * i] * g[0] + Jp1[3 * i + 1] * g[1] + Jp1[3 * i + 2] * g[2];
    }
    const double r0[3] = {z0[0] - h0[0], z0[1] - h0[1], z0[2] - h0[2]};
    const double r1[3] = {z1[0] - h1[0], z1[1] - h1[1], z1[2] - h1[2]};
    const double w = c.w;
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) {
            double s = 0.0;
            for (int k = 0; k < 3; k++) s += J0[3 * k + i] * J0[3 * k + j] + JX1[3 * k + i] * JX1[3 * k + j];
            V[3 * i + j] = w * s;
        }
        double s = 0.0;
        for (int k = 0; k < 3; k++) s += J0[3 * k + i] * r0[k] + JX1[3 * k + i] * r1[k];
        gF[i] = w * s;
    }
    for (int i = 0; i < 6; i++) {
        for (int j = 0; j < 3; j++) {
            double s = 0.0;
            for (int k = 0; k < 3; k++) s += JP1[6 * k + i] * JX1[3 * k + j];
            W[3 * i + j] = w * s;
        }
        for (int j = 0; j < 6; j++) {
            double s = 0.0;
            for (int k = 0; k < 3; k++) s += JP1[6 * k + i] * JP1[6 * k + j];
            Ub[6 * i + j] = w * s;
        }
        double s = 0.0;
        for (int k = 0; k < 3; k++) s += JP1[6 * k + i] * r1[k];
        gP[i] = w * s;
    }
    return w * (r0[0] * r0[0] + r0[1] * r0[1] + r0[2] * r0[2] + r1[0] * r1[0] + r1[1] * r1[1] + r1[2] * r1[2]);
}

// symmetric 3x3 inverse (cofactors); returns the determinant
LSFM_HD double inv3_sym(const double *a, double *o)
{
    const double c00 = a[4] * a[8] - a[5] * a[7];
    const double c10 = a[5] * a[6] - a[3] * a[8];
    const double c20 = a[3] * a[7] - a[4] * a[6];
    const double det = a[0] * c00 + a[1] * c10 + a[2] * c20;
    const double id = 1.0 / det;
    o[0] = c00 * id;
    o[1] = o[3] = (a[2] * a[7] - a[1] * a[8]) * id;
    o[2] = o[6] = (a[1] * a[5] - a[2] * a[4]) * id;
    o[4] = (a[0] * a[8] - a[2] * a[6]) * id;
    o[5] = o[7] = (a[2] * a[3] - a[0] * a[5]) * id;
    o[8] = (a[0] * a[4] - a[1] * a[3]) * id;
    return det;
}

// solve the SPD 6x6 system S x = e in place (Cholesky); false if a pivot is not positive
LSFM_HD bool solve6_spd(double *S, double *e)
{
    for (int j = 0; j < 6; j++) {
        double d = S[6 * j + j];
        for (int q = 0; q < j; q++) d -= S[6 * j + q] * S[6 * j + q];
        if (!(d > 0.0)) return false;
        d = sqrt(d);
        S[6 * j + j] = d;
        for (int i = j + 1; i < 6; i++) {
            double v = S[6 * i + j];
            for (int q = 0; q < j; q++) v -= S[6 * i + q] * S[6 * j + q];
            S[6 * i + j] = v / d;
        }
    }
    for (int i = 0; i < 6; i++) {
        double v = e[i];
        for (int q = 0; q < i; q++) v -= S[6 * i + q] * e[q];
        e[i] = v / S[6 * i + i];
    }
    for (int i = 5; i >= 0; i--) {
        double v = e[i];
        for (int q = i + 1; q < 6; q++) v -= S[6 * q + i] * e[q];
        e[i] = v / S[6 * i + i];
    }
    return true;
}

} // namespace bld
