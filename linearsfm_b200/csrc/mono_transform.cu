// Batched frame transform of MONOCULAR maps: row a15 of SURVEY 8(a).
//
// Reference: CLinearSFMImp::lmj_Transform_PF3DMono, LinearSFMImp.cpp:3173-6509
//   forward similarity transform of the state           3216-3306
//   constants of the inverse map at the new state       3311-3365
//   Jacobians J1 (diagonal), J2 (column of the old-Ref slot), J3 (column of the old-ScaP slot,
//   translation only) + gauge clean-up                  3371-3710
//   U congruence, 9 product families, routing           3713-4986
//   W/V congruence, routing                             4989-6503
//
// J = D + C2 e_a^T + C3 e_b^T  (a = slot of the old Ref pose, b = slot of the old ScaP pose), so
// I' = J^T I J has nine families per block; blocks of the pair (x,a) live in slot x of the first m
// leading U slots, blocks of (x,b) in slot m+x (the pair (a,b) therefore exists twice, which is why
// the mono solver accumulates U into S with +=, LinearSFMImp.cpp:6866-6876).  The arithmetic here is
// written from those formulas, not translated; the routing rules are the reference's.
// First (correctness) version: one thread per pose / U block / feature, FP64 atomics for shared
// targets; the stereo path's pose-major / pipelined kernels are the model for the fast version.
#include "ops.h"
#include "geom.cuh"
#include "small_mat.cuh"
#include <cub/cub.cuh>
#include <climits>

namespace {

struct MonoConst {
    double R[9], t[3], Scale;          // forward: x' = R (x - t) / Scale
    int slotRef, slotSca;              // slots of the NEW Ref / ScaP poses (gauge clean-up)
    int a, b;                          // slots of the OLD Ref / ScaP poses (column targets)
    int newFix, oldFix, signOut;
    double t3[3], Rq[9], QA[9], QB[9], QG[9];
    double s, s2, sigT[3], sigTT[3], sigA, sigB, sigG;
    int clearJ2Fix, clearJ3;
};

struct PoseJ3 { double J1[36], J2[36], J3[36]; };

__global__ void k_mono_find(const DMap *__restrict__ in, const int *__restrict__ posePre, int K,
                            int totPose, const int *__restrict__ newRef, const int *__restrict__ newSca,
                            int *__restrict__ slots /* [4][K]: newRef, newSca, oldRef, oldSca */)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= totPose) return;
    int k = seg_find(posePre, K, g);
    int p = g - posePre[k];
    int no = in[k].poseNo[p];
    if (no == -newRef[k]) atomicMin(&slots[0 * K + k], p);
    if (no == -newSca[k]) atomicMin(&slots[1 * K + k], p);
    if (no == -in[k].Ref) atomicMin(&slots[2 * K + k], p);
    if (no == -in[k].ScaP) atomicMin(&slots[3 * K + k], p);
}

__global__ void k_mono_const_fwd(const DMap *__restrict__ in, int K, const int *__restrict__ slots,
                                 const int *__restrict__ newFix, MonoConst *__restrict__ mc)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const DMap &M = in[k];
    MonoConst c;
    c.slotRef = slots[0 * K + k]; c.slotSca = slots[1 * K + k];
    c.a = slots[2 * K + k]; c.b = slots[3 * K + k];
    c.newFix = newFix[k]; c.oldFix = M.Fix;
    const double *x = M.poseVal + 6 * (size_t)c.slotRef;
    c.t[0] = x[0]; c.t[1] = x[1]; c.t[2] = x[2];
    geom::rot_ypr(x[3], x[4], x[5], c.R);
    const double *y = M.poseVal + 6 * (size_t)c.slotSca;
    double d[3] = {y[0] - c.t[0], y[1] - c.t[1], y[2] - c.t[2]}, ts[3];
    geom::mat3_vec(c.R, d, ts);
    c.Scale = fabs(ts[c.newFix]);
    c.signOut = ts[c.newFix] >= 0 ? 1 : -1;
    c.clearJ2Fix = (c.a == c.slotSca);           // pos3 == pos2   (3703)
    c.clearJ3 = (c.b == c.slotRef);              // pos4 == pos1   (3709)
    mc[k] = c;
}

__global__ void k_mono_state(const DMap *__restrict__ in, DMap *__restrict__ out,
                             const int *__restrict__ posePre, const int *__restrict__ featPre, int K,
                             int totPose, int totFeat, const MonoConst *__restrict__ mc)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= totPose + totFeat) return;
    if (g < totPose) {
        int k = seg_find(posePre, K, g);
        int p = g - posePre[k];
        const MonoConst &c = mc[k];
        const double *x = in[k].poseVal + 6 * (size_t)p;
        double *y = out[k].poseVal + 6 * (size_t)p;
        out[k].poseNo[p] = in[k].poseNo[p];                               // no relabelling (3262)
        double d[3] = {x[0] - c.t[0], x[1] - c.t[1], x[2] - c.t[2]}, tn[3];
        geom::mat3_vec(c.R, d, tn);
        double R2[9], R3[9], a[3];
        geom::rot_ypr(x[3], x[4], x[5], R2);
        geom::mul_abt(R2, c.R, R3);
        geom::ypr_of(R3, a[0], a[1], a[2]);
        y[0] = tn[0] / c.Scale; y[1] = tn[1] / c.Scale; y[2] = tn[2] / c.Scale;
        y[3] = a[0]; y[4] = a[1]; y[5] = a[2];
        if (p == c.slotRef) { for (int q = 0; q < 6; q++) y[q] = 0.0; }   // 3282-3290
        if (p == c.slotSca) y[c.newFix] = (double)c.signOut;              // 3291-3294
    } else {
        int gf = g - totPose;
        int k = seg_find(featPre, K, gf);
        int f = gf - featPre[k];
        const MonoConst &c = mc[k];
        const double *x = in[k].featVal + 3 * (size_t)f;
        double *y = out[k].featVal + 3 * (size_t)f;
        out[k].featNo[f] = in[k].featNo[f];
        double d[3] = {x[0] - c.t[0], x[1] - c.t[1], x[2] - c.t[2]}, tn[3];
        geom::mat3_vec(c.R, d, tn);
        y[0] = tn[0] / c.Scale; y[1] = tn[1] / c.Scale; y[2] = tn[2] / c.Scale;
    }
}

// constants of the inverse map, evaluated at the NEW values of slots a (old Ref) and b (old ScaP)
__global__ void k_mono_const_inv(const DMap *__restrict__ out, int K, MonoConst *__restrict__ mc)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    MonoConst c = mc[k];
    const double *x = out[k].poseVal + 6 * (size_t)c.a;
    c.t3[0] = x[0]; c.t3[1] = x[1]; c.t3[2] = x[2];
    geom::rot_derivs(x[3], x[4], x[5], c.Rq, c.QA, c.QB, c.QG);
    const double *y = out[k].poseVal + 6 * (size_t)c.b;
    double d[3] = {y[0] - c.t3[0], y[1] - c.t3[1], y[2] - c.t3[2]}, ts[3], v[3];
    geom::mat3_vec(c.Rq, d, ts);
    c.s = fabs(ts[c.oldFix]);
    c.s2 = c.s * c.s;
    double sg = ts[c.oldFix] >= 0 ? 1.0 : -1.0;
    for (int q = 0; q < 3; q++) { c.sigT[q] = -c.Rq[3 * c.oldFix + q] * sg; c.sigTT[q] = c.Rq[3 * c.oldFix + q] * sg; }
    geom::mat3_vec(c.QA, d, v); c.sigA = v[c.oldFix] * sg;
    geom::mat3_vec(c.QB, d, v); c.sigB = v[c.oldFix] * sg;
    geom::mat3_vec(c.QG, d, v); c.sigG = v[c.oldFix] * sg;
    mc[k] = c;
}

// translation parts of the Jacobians for a point with new value x (pose position or feature)
__device__ __forceinline__ void mono_point_jac(const MonoConst &c, const double *x, double *D /*3x3*/,
                                               double *C2 /*3x6*/, double *C3 /*3x3*/)
{
    double d[3] = {x[0] - c.t3[0], x[1] - c.t3[1], x[2] - c.t3[2]}, e[3], va[3], vb[3], vg[3];
    geom::mat3_vec(c.Rq, d, e);
    geom::mat3_vec(c.QA, d, va);
    geom::mat3_vec(c.QB, d, vb);
    geom::mat3_vec(c.QG, d, vg);
#pragma unroll
    for (int r = 0; r < 3; r++) {
#pragma unroll
        for (int q = 0; q < 3; q++) {
            D[3 * r + q] = c.Rq[3 * r + q] / c.s;
            C2[6 * r + q] = (-c.Rq[3 * r + q] * c.s - e[r] * c.sigT[q]) / c.s2;
            C3[3 * r + q] = (-e[r] * c.sigTT[q]) / c.s2;
        }
        C2[6 * r + 3] = (va[r] * c.s - e[r] * c.sigA) / c.s2;
        C2[6 * r + 4] = (vb[r] * c.s - e[r] * c.sigB) / c.s2;
        C2[6 * r + 5] = (vg[r] * c.s - e[r] * c.sigG) / c.s2;
    }
}

__global__ void k_mono_posejac(const DMap *__restrict__ out, const int *__restrict__ posePre, int K,
                               int totPose, const MonoConst *__restrict__ mc, PoseJ3 *__restrict__ pj)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= totPose) return;
    int k = seg_find(posePre, K, g);
    int p = g - posePre[k];
    const MonoConst &c = mc[k];
    const double *x = out[k].poseVal + 6 * (size_t)p;
    PoseJ3 J;
#pragma unroll
    for (int i = 0; i < 36; i++) { J.J1[i] = 0.0; J.J2[i] = 0.0; J.J3[i] = 0.0; }
    double D[9], C2[18], C3[9];
    mono_point_jac(c, x, D, C2, C3);
    double R2[9], dA2[9], dB2[9], dG2[9], Ri[9], tmp[9], v[3];
    geom::rot_derivs(x[3], x[4], x[5], R2, dA2, dB2, dG2);
    geom::mul_abt(R2, c.Rq, Ri);
    double A1[9], A2[9];               // angle blocks: columns = d(angles)/d(own angle X), d/d(slot-a angle X)
    geom::mul_abt(dA2, c.Rq, tmp); geom::dangles(tmp, Ri, false, v); A1[0] = v[0]; A1[3] = v[1]; A1[6] = v[2];
    geom::mul_abt(dB2, c.Rq, tmp); geom::dangles(tmp, Ri, false, v); A1[1] = v[0]; A1[4] = v[1]; A1[7] = v[2];
    geom::mul_abt(dG2, c.Rq, tmp); geom::dangles(tmp, Ri, false, v); A1[2] = v[0]; A1[5] = v[1]; A1[8] = v[2];
    geom::mul_abt(R2, c.QA, tmp); geom::dangles(tmp, Ri, false, v); A2[0] = v[0]; A2[3] = v[1]; A2[6] = v[2];
    geom::mul_abt(R2, c.QB, tmp); geom::dangles(tmp, Ri, false, v); A2[1] = v[0]; A2[4] = v[1]; A2[7] = v[2];
    geom::mul_abt(R2, c.QG, tmp); geom::dangles(tmp, Ri, false, v); A2[2] = v[0]; A2[5] = v[1]; A2[8] = v[2];
    // J1 = own derivative (3476-3493)
    for (int r = 0; r < 3; r++)
        for (int q = 0; q < 3; q++) { J.J1[6 * r + q] = D[3 * r + q]; J.J1[6 * (r + 3) + 3 + q] = A1[3 * r + q]; }
    // column of slot a: into J1 for the slot itself, J2 otherwise (3495-3556)
    double *Ja = (p == c.a) ? J.J1 : J.J2;
    for (int r = 0; r < 3; r++)
        for (int q = 0; q < 3; q++) {
            Ja[6 * r + q] += C2[6 * r + q];
            Ja[6 * r + 3 + q] += C2[6 * r + 3 + q];
            Ja[6 * (r + 3) + 3 + q] += A2[3 * r + q];
        }
    // column of slot b (translation only): into J1 for the slot itself, J3 otherwise (3558-3581)
    double *Jb = (p == c.b) ? J.J1 : J.J3;
    for (int r = 0; r < 3; r++)
        for (int q = 0; q < 3; q++) Jb[6 * r + q] += C3[3 * r + q];
    // gauge clean-up (3691-3710)
    if (p == c.slotRef) for (int i = 0; i < 36; i++) J.J1[i] = 0.0;
    if (p == c.slotSca) for (int r = 0; r < 6; r++) J.J1[6 * r + c.newFix] = 0.0;
    if (c.clearJ2Fix) for (int r = 0; r < 6; r++) J.J2[6 * r + c.newFix] = 0.0;
    if (c.clearJ3) for (int i = 0; i < 36; i++) J.J3[i] = 0.0;
    pj[g] = J;
}

__global__ void k_mono_uinit(DMap *__restrict__ out, const int *__restrict__ posePre, int K, int totPose,
                             const MonoConst *__restrict__ mc)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= totPose) return;
    int k = seg_find(posePre, K, g);
    int i = g - posePre[k];
    int m = out[k].m, a = mc[k].a, b = mc[k].b;
    out[k].Ui[i] = i <= a ? i : a;
    out[k].Uj[i] = i <= a ? a : i;
    out[k].Ui[m + i] = i <= a ? i : b;       // the reference tests i<=posID here too (3753-3764)
    out[k].Uj[m + i] = i <= a ? b : i;
    double *u = out[k].U + 36 * (size_t)i, *u2 = out[k].U + 36 * (size_t)(m + i);
    for (int q = 0; q < 36; q++) { u[q] = 0.0; u2[q] = 0.0; }
}

__global__ void k_mono_count_u(const DMap *__restrict__ in, const int *__restrict__ uPre, int K, int totU,
                               const MonoConst *__restrict__ mc, int *__restrict__ flag)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g > totU) return;
    if (g == totU) { flag[g] = 0; return; }
    int k = seg_find(uPre, K, g);
    int bi = g - uPre[k];
    int i = in[k].Ui[bi], j = in[k].Uj[bi], a = mc[k].a, b = mc[k].b;
    flag[g] = (i != a && j != a && i != b && j != b) ? 1 : 0;
}

__global__ void k_mono_count_w(const DMap *__restrict__ in, const int *__restrict__ featPre, int K,
                               int totFeat, const MonoConst *__restrict__ mc, int *__restrict__ cnt)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g > totFeat) return;
    if (g == totFeat) { cnt[g] = 0; return; }
    int k = seg_find(featPre, K, g);
    int f = g - featPre[k];
    const DMap &M = in[k];
    int a = mc[k].a, b = mc[k].b, c = 2;
    for (int j = M.wPtr[f]; j < M.wPtr[f + 1]; j++) c += (M.photo[j] != a && M.photo[j] != b);
    cnt[g] = c;
}

__global__ void k_mono_sizes(const int *__restrict__ uPre, const int *__restrict__ featPre, int K,
                             const int *__restrict__ uScan, const int *__restrict__ fScan,
                             int *__restrict__ nSurv, int *__restrict__ nWnew)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    nSurv[k] = uScan[uPre[k + 1]] - uScan[uPre[k]];
    nWnew[k] = fScan[featPre[k + 1]] - fScan[featPre[k]];
}

__device__ __forceinline__ void add6(double *dst, const double *P, bool asis, bool transposed)
{
    for (int r = 0; r < 6; r++)
        for (int q = 0; q < 6; q++) {
            double v = 0.0;
            if (asis) v += P[6 * r + q];
            if (transposed) v += P[6 * q + r];
            if (asis || transposed) atomicAdd(dst + 6 * r + q, v);
        }
}

// X^T I Y for dense 6x6
__device__ __forceinline__ void xtiy(const double *X, const double *I, const double *Y, double *P)
{
    double T[36];
    sm::mtm<6, 6, 6>(X, I, T);
    sm::mm<6, 6, 6>(T, Y, P);
}

// one thread per old U block: nine families (3767-4984)
__global__ void __launch_bounds__(64)
k_mono_ucong(const DMap *__restrict__ in, DMap *__restrict__ out, const int *__restrict__ uPre,
             const int *__restrict__ posePre, int K, int totU, const MonoConst *__restrict__ mc,
             const PoseJ3 *__restrict__ pj, const int *__restrict__ uScan)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= totU) return;
    int k = seg_find(uPre, K, g);
    int bi = g - uPre[k];
    const DMap &M = in[k];
    const int m = M.m, a = mc[k].a, b = mc[k].b;
    const int i = M.Ui[bi], j = M.Uj[bi];
    const bool off = (i != j);
    double I[36], P[36];
    sm::load<36>(M.U + 36 * (size_t)bi, I);
    const PoseJ3 &Ji = pj[posePre[k] + i];
    const PoseJ3 &Jj = pj[posePre[k] + j];
    double *Un = out[k].U;
    // C2_i^T I C2_j -> (a,a): slot a
    xtiy(Ji.J2, I, Jj.J2, P); add6(Un + 36 * (size_t)a, P, true, off);
    // C2_i^T I D_j -> (a,j): slot j
    xtiy(Ji.J2, I, Jj.J1, P); add6(Un + 36 * (size_t)j, P, j >= a, (j <= a) && off);
    // C2_i^T I C3_j -> (a,b): slot b
    xtiy(Ji.J2, I, Jj.J3, P); add6(Un + 36 * (size_t)b, P, b > a, (b < a) && off);
    // D_i^T I D_j -> (i,j)
    xtiy(Ji.J1, I, Jj.J1, P);
    if (i == a) add6(Un + 36 * (size_t)j, P, true, false);
    else if (j == a) add6(Un + 36 * (size_t)i, P, true, false);
    else if (i == b) add6(Un + 36 * (size_t)(m + j), P, true, false);
    else if (j == b) add6(Un + 36 * (size_t)(m + i), P, true, false);
    else {
        int slot = 2 * m + (uScan[g] - uScan[uPre[k]]);
        sm::store<36>(Un + 36 * (size_t)slot, P);
        out[k].Ui[slot] = i;
        out[k].Uj[slot] = j;
    }
    // D_i^T I C2_j -> (i,a): slot i
    xtiy(Ji.J1, I, Jj.J2, P); add6(Un + 36 * (size_t)i, P, i <= a, (i >= a) && off);
    // D_i^T I C3_j -> (i,b): slot m+i
    xtiy(Ji.J1, I, Jj.J3, P); add6(Un + 36 * (size_t)(m + i), P, i <= b, (i >= b) && off);
    // C3_i^T I C3_j -> (b,b): slot m+b
    xtiy(Ji.J3, I, Jj.J3, P); add6(Un + 36 * (size_t)(m + b), P, true, off);
    // C3_i^T I D_j -> (b,j): slot m+j
    xtiy(Ji.J3, I, Jj.J1, P); add6(Un + 36 * (size_t)(m + j), P, j >= b, (j <= b) && off);
    // C3_i^T I C2_j -> (b,a): slot b
    xtiy(Ji.J3, I, Jj.J2, P); add6(Un + 36 * (size_t)b, P, b < a, (b > a) && off);
}

// one thread per feature (5017-6501)
__global__ void __launch_bounds__(64)
k_mono_wv(const DMap *__restrict__ in, DMap *__restrict__ out, const int *__restrict__ featPre,
          const int *__restrict__ posePre, int K, int totFeat, const MonoConst *__restrict__ mc,
          const PoseJ3 *__restrict__ pj, const int *__restrict__ fScan)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= totFeat) return;
    int k = seg_find(featPre, K, g);
    int f = g - featPre[k];
    const DMap &M = in[k];
    const DMap &O = out[k];
    const MonoConst &c = mc[k];
    const int m = M.m, a = c.a, b = c.b;
    double D[9], C2[18], C3f[9], C3[18];
    mono_point_jac(c, O.featVal + 3 * (size_t)f, D, C2, C3f);
    for (int r = 0; r < 3; r++)
        for (int q = 0; q < 6; q++) C3[6 * r + q] = (q < 3) ? C3f[3 * r + q] : 0.0;
    if (c.clearJ2Fix) for (int r = 0; r < 3; r++) C2[6 * r + c.newFix] = 0.0;
    if (c.clearJ3) for (int i = 0; i < 18; i++) C3[i] = 0.0;
    double V[9], VD[9], VC2[18], VC3[18], Vn[9];
    sm::load<9>(M.V + 9 * (size_t)f, V);
    sm::mm<3, 3, 3>(V, D, VD);
    sm::mm<3, 3, 6>(V, C2, VC2);
    sm::mm<3, 3, 6>(V, C3, VC3);
    sm::mtm<3, 3, 3>(D, VD, Vn);
    sm::store<9>(O.V + 9 * (size_t)f, Vn);
    double Wa[18], Wb[18];                      // new blocks (a,f) and (b,f)
    sm::mtm<6, 3, 3>(C2, VD, Wa);               // C2_f^T V D_f
    sm::mtm<6, 3, 3>(C3, VD, Wb);               // C3_f^T V D_f
    double Uaa[36], Ubb[36], Uab[36];           // (a,a), (b,b) before symmetrising the W part; (a,b) oriented
    sm::mtm<6, 3, 6>(C2, VC2, Uaa);
    sm::mtm<6, 3, 6>(C3, VC3, Ubb);
    {
        double X[36];
        sm::mtm<6, 3, 6>(C2, VC3, X);           // C2^T V C3 = block (a,b)
        for (int r = 0; r < 6; r++)
            for (int q = 0; q < 6; q++) Uab[6 * r + q] = (a < b) ? X[6 * r + q] : (a > b ? X[6 * q + r] : 0.0);
    }
    int o0 = fScan[g] - fScan[featPre[k]];
    O.wPtr[f] = o0;
    O.photo[o0] = a; O.feature[o0] = f;
    O.photo[o0 + 1] = b; O.feature[o0 + 1] = f;
    int onext = o0 + 2;
    double *Un = O.U;
    for (int j = M.wPtr[f]; j < M.wPtr[f + 1]; j++) {
        int p = M.photo[j];
        const PoseJ3 &J = pj[posePre[k] + p];
        double W[18], WD[18], WC2[36], WC3[36], T[36];
        sm::load<18>(M.W + 18 * (size_t)j, W);
        sm::mm<6, 3, 3>(W, D, WD);
        sm::mm<6, 3, 6>(W, C2, WC2);
        sm::mm<6, 3, 6>(W, C3, WC3);
        double a1[18];
        sm::mtm<6, 6, 3>(J.J1, WD, a1);                        // D_p^T W D_f
        if (p == a) { for (int i = 0; i < 18; i++) Wa[i] += a1[i]; }
        else if (p == b) { for (int i = 0; i < 18; i++) Wb[i] += a1[i]; }
        else {
            sm::store<18>(O.W + 18 * (size_t)onext, a1);
            O.photo[onext] = p; O.feature[onext] = f;
            onext++;
        }
        double t18[18];
        sm::mtm<6, 6, 3>(J.J2, WD, t18); for (int i = 0; i < 18; i++) Wa[i] += t18[i];   // C2_p^T W D_f
        sm::mtm<6, 6, 3>(J.J3, WD, t18); for (int i = 0; i < 18; i++) Wb[i] += t18[i];   // C3_p^T W D_f
        // (a,a) += C2_p^T W C2_f + ^T ;  (b,b) += C3_p^T W C3_f + ^T
        sm::mtm<6, 6, 6>(J.J2, WC2, T);
        for (int r = 0; r < 6; r++) for (int q = 0; q < 6; q++) Uaa[6 * r + q] += T[6 * r + q] + T[6 * q + r];
        sm::mtm<6, 6, 6>(J.J3, WC3, T);
        for (int r = 0; r < 6; r++) for (int q = 0; q < 6; q++) Ubb[6 * r + q] += T[6 * r + q] + T[6 * q + r];
        // (a,b): C2_p^T W C3_f is block (a,b); C3_p^T W C2_f is block (b,a)
        sm::mtm<6, 6, 6>(J.J2, WC3, T);
        for (int r = 0; r < 6; r++) for (int q = 0; q < 6; q++)
            Uab[6 * r + q] += (a < b) ? T[6 * r + q] : (a > b ? T[6 * q + r] : 0.0);
        sm::mtm<6, 6, 6>(J.J3, WC2, T);
        for (int r = 0; r < 6; r++) for (int q = 0; q < 6; q++)
            Uab[6 * r + q] += (a > b) ? T[6 * r + q] : (a < b ? T[6 * q + r] : 0.0);
        // (p,a): D_p^T W C2_f -> slot p ; (p,b): D_p^T W C3_f -> slot m+p
        sm::mtm<6, 6, 6>(J.J1, WC2, T); add6(Un + 36 * (size_t)p, T, p <= a, p >= a);
        sm::mtm<6, 6, 6>(J.J1, WC3, T); add6(Un + 36 * (size_t)(m + p), T, p <= b, p >= b);
    }
    sm::store<18>(O.W + 18 * (size_t)o0, Wa);
    sm::store<18>(O.W + 18 * (size_t)(o0 + 1), Wb);
    add6(Un + 36 * (size_t)a, Uaa, true, false);
    add6(Un + 36 * (size_t)(m + b), Ubb, true, false);
    add6(Un + 36 * (size_t)b, Uab, true, false);
}

__global__ void k_mono_wend(DMap *__restrict__ out, int K, const int *__restrict__ nWnew)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    out[k].wPtr[out[k].n] = nWnew[k];
}

void exclusive_scan(Context &ctx, const int *in, int *out, int n)
{
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, in, out, n, ctx.stream);
    DevBuf<char> tmp(tmp_bytes, ctx.stream);
    cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, in, out, n, ctx.stream);
}

} // namespace

std::vector<MapHandle> transform_mono_batch(Context &ctx, const std::vector<MapHandle> &in,
                                            const std::vector<int> &newRef, const std::vector<int> &newSca,
                                            const std::vector<int> &newFix)
{
    const int K = (int)in.size();
    if (K == 0) return {};
    cudaStream_t s = ctx.stream;
    ctx.begin("mono.transform");
    const int TB = 128;
    int nl = 0;
    OpMaps A;
    A.build(in, s);
    DevBuf<int> dRef(K, s), dSca(K, s), dFix(K, s), slots(4 * (size_t)K, s);
    dRef.upload(newRef); dSca.upload(newSca); dFix.upload(newFix);
    std::vector<int> init(4 * (size_t)K, INT_MAX);
    slots.upload(init);
    k_mono_find<<<ceil_div(A.totPose, TB), TB, 0, s>>>(A.d.p, A.dPosePre.p, K, A.totPose, dRef.p, dSca.p, slots.p); nl++;
    std::vector<int> hSlots(4 * (size_t)K);
    slots.download(hSlots.data(), hSlots.size());
    CUDA_CHECK(cudaStreamSynchronize(s));
    for (int k = 0; k < K; k++)
        for (int q = 0; q < 4; q++)
            if (hSlots[q * K + k] == INT_MAX)
                throw LsfmError(LSFM_ERR_REF_NOT_FOUND, "mono Transform: Ref/ScaP pose not in the state of the map");
    DevBuf<MonoConst> mc(K, s);
    k_mono_const_fwd<<<ceil_div(K, 64), 64, 0, s>>>(A.d.p, K, slots.p, dFix.p, mc.p); nl++;
    DevBuf<int> uFlag(A.totU + 1, s), uScan(A.totU + 1, s), fCnt(A.totFeat + 1, s), fScan(A.totFeat + 1, s);
    k_mono_count_u<<<ceil_div(A.totU + 1, TB), TB, 0, s>>>(A.d.p, A.dUPre.p, K, A.totU, mc.p, uFlag.p); nl++;
    k_mono_count_w<<<ceil_div(A.totFeat + 1, TB), TB, 0, s>>>(A.d.p, A.dFeatPre.p, K, A.totFeat, mc.p, fCnt.p); nl++;
    exclusive_scan(ctx, uFlag.p, uScan.p, A.totU + 1); nl += 2;
    exclusive_scan(ctx, fCnt.p, fScan.p, A.totFeat + 1); nl += 2;
    DevBuf<int> dSizes(2 * (size_t)K, s);
    k_mono_sizes<<<ceil_div(K, TB), TB, 0, s>>>(A.dUPre.p, A.dFeatPre.p, K, uScan.p, fScan.p, dSizes.p, dSizes.p + K); nl++;
    std::vector<int> hSizes(2 * (size_t)K);
    std::vector<MonoConst> hmc(K);
    dSizes.download(hSizes.data(), hSizes.size());
    mc.download(hmc.data(), K);
    CUDA_CHECK(cudaStreamSynchronize(s));

    std::vector<DMap> shapes(K);
    for (int k = 0; k < K; k++) {
        DMap &o = shapes[k];
        o = A.h[k];
        o.Ref = newRef[k]; o.ScaP = newSca[k]; o.Fix = newFix[k]; o.Sign = hmc[k].signOut;
        o.nU = 2 * A.h[k].m + hSizes[k];
        o.nW = hSizes[K + k];
    }
    std::vector<MapHandle> out = alloc_maps(ctx, shapes);
    OpMaps B;
    B.build(out, s);
    DevBuf<PoseJ3> pj(A.totPose, s);
    k_mono_state<<<ceil_div(A.totPose + A.totFeat, TB), TB, 0, s>>>(A.d.p, B.d.p, A.dPosePre.p, A.dFeatPre.p, K,
                                                                   A.totPose, A.totFeat, mc.p); nl++;
    k_mono_const_inv<<<ceil_div(K, 64), 64, 0, s>>>(B.d.p, K, mc.p); nl++;
    k_mono_posejac<<<ceil_div(A.totPose, 64), 64, 0, s>>>(B.d.p, A.dPosePre.p, K, A.totPose, mc.p, pj.p); nl++;
    k_mono_uinit<<<ceil_div(A.totPose, TB), TB, 0, s>>>(B.d.p, A.dPosePre.p, K, A.totPose, mc.p); nl++;
    k_mono_wend<<<ceil_div(K, TB), TB, 0, s>>>(B.d.p, K, dSizes.p + K); nl++;
    if (A.totU > 0) {
        k_mono_ucong<<<ceil_div(A.totU, 64), 64, 0, s>>>(A.d.p, B.d.p, A.dUPre.p, A.dPosePre.p, K, A.totU, mc.p,
                                                        pj.p, uScan.p); nl++;
    }
    if (A.totFeat > 0) {
        k_mono_wv<<<ceil_div(A.totFeat, 64), 64, 0, s>>>(A.d.p, B.d.p, A.dFeatPre.p, A.dPosePre.p, K, A.totFeat,
                                                        mc.p, pj.p, fScan.p); nl++;
    }
    KERNEL_CHECK();
    ctx.end(0.0, 0.0, nl);
    return out;
}
