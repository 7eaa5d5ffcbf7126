// Rotation helpers of the frame transform, FP64, host+device.
// Conventions restated from the reference (formulas only, independent code):
//   R(alpha,beta,gamma) = Rx(gamma) Ry(beta) Rz(alpha), passive, row-major   LinearSFMImp.cpp:132-143
//   YPR extraction with the exact-zero gimbal test and PI = 3.1415926         LinearSFMImp.cpp:145-177, Imp.h:57
//   d(angles)/d(parameter) through the atan formulas                          LinearSFMImp.cpp:282-334
#pragma once
#include <math.h>

#ifndef __CUDACC__
#define __host__
#define __device__
#endif

#define LSFM_PI 3.1415926   // the reference's literal (LinearSFMImp.h:57) -- needed for parity

namespace geom {

struct SinCos { double sa, ca, sb, cb, sg, cg; };

__host__ __device__ inline SinCos sincos3(double a, double b, double g)
{
    SinCos s;
    s.sa = sin(a); s.ca = cos(a); s.sb = sin(b); s.cb = cos(b); s.sg = sin(g); s.cg = cos(g);
    return s;
}

__host__ __device__ inline void rot_from(const SinCos &s, double *R)
{
    R[0] = s.cb * s.ca;                       R[1] = s.cb * s.sa;                       R[2] = -s.sb;
    R[3] = s.sg * s.sb * s.ca - s.cg * s.sa;  R[4] = s.sg * s.sb * s.sa + s.cg * s.ca;  R[5] = s.sg * s.cb;
    R[6] = s.cg * s.sb * s.ca + s.sg * s.sa;  R[7] = s.cg * s.sb * s.sa - s.sg * s.ca;  R[8] = s.cg * s.cb;
}

__host__ __device__ inline void rot_ypr(double a, double b, double g, double *R)
{
    rot_from(sincos3(a, b, g), R);
}

// R and its partial derivatives wrt alpha, beta, gamma (what lmj_Rderivation returns, 179-280)
__host__ __device__ inline void rot_derivs(double a, double b, double g, double *R, double *dA,
                                           double *dB, double *dG)
{
    SinCos s = sincos3(a, b, g);
    rot_from(s, R);
    dA[0] = -s.cb * s.sa;                      dA[1] = s.cb * s.ca;                       dA[2] = 0.0;
    dA[3] = -s.sg * s.sb * s.sa - s.cg * s.ca; dA[4] = s.sg * s.sb * s.ca - s.cg * s.sa;  dA[5] = 0.0;
    dA[6] = -s.cg * s.sb * s.sa + s.sg * s.ca; dA[7] = s.cg * s.sb * s.ca + s.sg * s.sa;  dA[8] = 0.0;

    dB[0] = -s.sb * s.ca;        dB[1] = -s.sb * s.sa;        dB[2] = -s.cb;
    dB[3] = s.sg * s.cb * s.ca;  dB[4] = s.sg * s.cb * s.sa;  dB[5] = -s.sg * s.sb;
    dB[6] = s.cg * s.cb * s.ca;  dB[7] = s.cg * s.cb * s.sa;  dB[8] = -s.cg * s.sb;

    dG[0] = 0.0;                               dG[1] = 0.0;                               dG[2] = 0.0;
    dG[3] = s.cg * s.sb * s.ca + s.sg * s.sa;  dG[4] = s.cg * s.sb * s.sa - s.sg * s.ca;  dG[5] = s.cg * s.cb;
    dG[6] = -s.sg * s.sb * s.ca + s.cg * s.sa; dG[7] = -s.sg * s.sb * s.sa - s.cg * s.ca; dG[8] = -s.sg * s.cb;
}

// angles of R (LinearSFMImp.cpp:162-177)
__host__ __device__ inline void ypr_of(const double *R, double &a, double &b, double &g)
{
    b = atan2(-R[2], sqrt(R[0] * R[0] + R[1] * R[1]));
    double cb = cos(b);
    if (cb == 0) { a = 0; b = LSFM_PI / 2; g = atan2(R[1], R[4]); }
    else { a = atan2(R[1] / cb, R[0] / cb); g = atan2(R[5] / cb, R[8] / cb); }
}

// angles of R^T without forming the transpose (LinearSFMImp.cpp:145-160)
__host__ __device__ inline void ypr_of_transpose(const double *R, double &a, double &b, double &g)
{
    b = atan2(-R[6], sqrt(R[0] * R[0] + R[3] * R[3]));
    double cb = cos(b);
    if (cb == 0) { a = 0; b = LSFM_PI / 2; g = atan2(R[3], R[4]); }
    else { a = atan2(R[3] / cb, R[0] / cb); g = atan2(R[7] / cb, R[8] / cb); }
}

// derivative of (alpha,beta,gamma)(Ri) along the matrix direction dRi (lmj_dRi, 282-307)
// `tr` selects the transposed indexing of lmj_dRiTT (309-334).
__host__ __device__ inline void dangles(const double *dRi, const double *Ri, bool tr, double *out)
{
    const int i1 = tr ? 3 : 1, i2 = tr ? 6 : 2, i5 = tr ? 7 : 5;
    double r0 = Ri[0], r1 = Ri[i1], r2 = Ri[i2], r5 = Ri[i5], r8 = Ri[8];
    double d0 = dRi[0], d1 = dRi[i1], d2 = dRi[i2], d5 = dRi[i5], d8 = dRi[8];
    double F1 = r1 / r0, F3 = r5 / r8;
    double F5 = r0 * r0 + r1 * r1;
    double F4 = sqrt(F5);
    double F2 = -r2 / F4;
    double dF1 = (d1 * r0 - r1 * d0) / (r0 * r0);
    double dF3 = (d5 * r8 - r5 * d8) / (r8 * r8);
    double dF5 = 2 * r0 * d0 + 2 * r1 * d1;
    double dF4 = dF5 / (2 * sqrt(F5));
    double dF2 = (-d2 * F4 + r2 * dF4) / F5;
    out[0] = dF1 / (1 + F1 * F1);
    out[1] = dF2 / (1 + F2 * F2);
    out[2] = dF3 / (1 + F3 * F3);
}

// C = A * B^T  (3x3, row-major)  (lmj_TimesRRT, 336-347)
__host__ __device__ inline void mul_abt(const double *A, const double *B, double *C)
{
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            C[3 * i + j] = A[3 * i] * B[3 * j] + A[3 * i + 1] * B[3 * j + 1] + A[3 * i + 2] * B[3 * j + 2];
}

__host__ __device__ inline void mat3_vec(const double *A, const double *x, double *y)
{
#pragma unroll
    for (int i = 0; i < 3; i++) y[i] = A[3 * i] * x[0] + A[3 * i + 1] * x[1] + A[3 * i + 2] * x[2];
}

} // namespace geom
