// Batched frame transform of (state, information) for stereo maps: rows a2-a5 of SURVEY 8(a).
//
// Reference: CLinearSFMImp::lmj_Transform_PF3DStereo, LinearSFMImp.cpp:349-1924
//   state transform            389-455        Jacobian build             459-683
//   U congruence + routing     686-1268       W/V congruence + routing   1270-1916
//
// The reference materialises dense J1/J2 and multiplies 6x6x6 products for every block.  Here:
//   * J = D + C e_pos^T with D block diagonal and C one block column, so I' = J^T I J splits into
//     four families per block (D^T I D, C^T I D, D^T I C, C^T I C);
//   * feature Jacobians (D_f = Q, C_f = [-Q | dQ (X'-t')]) are recomputed on the fly, pose
//     Jacobians are kept as four 3x3 blocks per pose (structural zeros skipped);
//   * one thread per feature streams that feature's W blocks; sums into shared targets
//     (U'(p,pos), U'(pos,pos)) are warp-aggregated FP64 atomics;
//   * output block lists keep the reference's exact order (Appendix C.1 of SURVEY.md): slots
//     [0,m) = pairs with posID, then surviving U blocks in input order; per feature the new
//     (posID,f) block first, then its surviving W blocks in input order.
#include "ops.h"
#include "geom.cuh"
#include "small_mat.cuh"
#include <cub/cub.cuh>
#include <climits>

namespace {

struct TfConst {
    double R[9], t[3];          // rotation / position of the new base pose in the OLD frame
    double tn[3];               // new value of slot pos: position of the old base in the new frame
    double Q[9], QA[9], QB[9], QG[9];
    int posID;
    int oldRef;
};

// Pose Jacobian blocks: J1 = [[s Q, b1],[0, c1]],  J2 = [[e Q, f2],[0, g2]]
//   ordinary pose: s=+1, e=-1, b1=0;   slot pos: s=-1, e=0, f2=g2=0
struct PoseJac {
    double b1[9], c1[9], f2[9], g2[9];
};

__global__ void k_find_pos(const DMap *__restrict__ in, const int *__restrict__ posePre, int K,
                           int totPose, const int *__restrict__ newRef, int *__restrict__ posID)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= totPose) return;
    int k = seg_find(posePre, K, g);
    int p = g - posePre[k];
    if (in[k].poseNo[p] == -newRef[k]) atomicMin(&posID[k], p);
}

__global__ void k_count_u(const DMap *__restrict__ in, const int *__restrict__ uPre, int K, int totU,
                          const int *__restrict__ posID, int *__restrict__ flag)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g > totU) return;
    if (g == totU) { flag[g] = 0; return; }
    int k = seg_find(uPre, K, g);
    int b = g - uPre[k];
    int pid = posID[k];
    flag[g] = (in[k].Ui[b] != pid && in[k].Uj[b] != pid) ? 1 : 0;
}

__global__ void k_count_w(const DMap *__restrict__ in, const int *__restrict__ featPre, int K,
                          int totFeat, const int *__restrict__ posID, int *__restrict__ cnt)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g > totFeat) return;
    if (g == totFeat) { cnt[g] = 0; return; }
    int k = seg_find(featPre, K, g);
    int f = g - featPre[k];
    const DMap &M = in[k];
    int pid = posID[k];
    int c = 1;
    for (int j = M.wPtr[f]; j < M.wPtr[f + 1]; j++) c += (M.photo[j] != pid);
    cnt[g] = c;
}

__global__ void k_sizes(const int *__restrict__ uPre, const int *__restrict__ featPre, int K,
                        const int *__restrict__ uScan, const int *__restrict__ fScan,
                        int *__restrict__ nSurv, int *__restrict__ nWnew)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    nSurv[k] = uScan[uPre[k + 1]] - uScan[uPre[k]];
    nWnew[k] = fScan[featPre[k + 1]] - fScan[featPre[k]];
}

// one thread per map: constants of the transform + the new value / Jacobian of slot pos
__global__ void k_tf_const(const DMap *__restrict__ in, DMap *__restrict__ out,
                           const int *__restrict__ posePre, int K, const int *__restrict__ posID,
                           const int *__restrict__ nWnew, TfConst *__restrict__ tc,
                           PoseJac *__restrict__ pj)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const DMap &M = in[k];
    out[k].wPtr[M.n] = nWnew[k];
    TfConst c;
    int p = posID[k];
    c.posID = p;
    c.oldRef = M.Ref;
    const double *x = M.poseVal + 6 * (size_t)p;
    c.t[0] = x[0]; c.t[1] = x[1]; c.t[2] = x[2];
    geom::rot_ypr(x[3], x[4], x[5], c.R);
    double rt[3];
    geom::mat3_vec(c.R, c.t, rt);
    c.tn[0] = -rt[0]; c.tn[1] = -rt[1]; c.tn[2] = -rt[2];
    double an[3];
    geom::ypr_of_transpose(c.R, an[0], an[1], an[2]);
    geom::rot_derivs(an[0], an[1], an[2], c.Q, c.QA, c.QB, c.QG);
    double dA[3], dB[3], dG[3];
    geom::dangles(c.QA, c.Q, true, dA);
    geom::dangles(c.QB, c.Q, true, dB);
    geom::dangles(c.QG, c.Q, true, dG);
    double *y = out[k].poseVal + 6 * (size_t)p;
    y[0] = c.tn[0]; y[1] = c.tn[1]; y[2] = c.tn[2]; y[3] = an[0]; y[4] = an[1]; y[5] = an[2];
    PoseJac J;
    double ta[3], tb[3], tg[3];
    geom::mat3_vec(c.QA, c.tn, ta);
    geom::mat3_vec(c.QB, c.tn, tb);
    geom::mat3_vec(c.QG, c.tn, tg);
#pragma unroll
    for (int r = 0; r < 3; r++) {
        J.b1[3 * r + 0] = -ta[r]; J.b1[3 * r + 1] = -tb[r]; J.b1[3 * r + 2] = -tg[r];
        J.c1[3 * r + 0] = dA[r];  J.c1[3 * r + 1] = dB[r];  J.c1[3 * r + 2] = dG[r];
    }
#pragma unroll
    for (int i = 0; i < 9; i++) { J.f2[i] = 0.0; J.g2[i] = 0.0; }
    pj[posePre[k] + p] = J;
    tc[k] = c;
}

// one thread per pose: new pose value + Jacobian blocks (ordinary poses), new stno
__global__ void k_tf_pose(const DMap *__restrict__ in, DMap *__restrict__ out,
                          const int *__restrict__ posePre, int K, int totPose,
                          const TfConst *__restrict__ tc, PoseJac *__restrict__ pj)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= totPose) return;
    int k = seg_find(posePre, K, g);
    int p = g - posePre[k];
    const DMap &M = in[k];
    const TfConst &c = tc[k];
    if (p == c.posID) { out[k].poseNo[p] = -c.oldRef; return; }     // LinearSFMImp.cpp:416-417
    out[k].poseNo[p] = M.poseNo[p];
    const double *x = M.poseVal + 6 * (size_t)p;
    double d[3] = {x[0] - c.t[0], x[1] - c.t[1], x[2] - c.t[2]};
    double t2[3];
    geom::mat3_vec(c.R, d, t2);
    double Rold[9], R3[9], a2[3];
    geom::rot_ypr(x[3], x[4], x[5], Rold);
    geom::mul_abt(Rold, c.R, R3);
    geom::ypr_of(R3, a2[0], a2[1], a2[2]);
    double *y = out[k].poseVal + 6 * (size_t)p;
    y[0] = t2[0]; y[1] = t2[1]; y[2] = t2[2]; y[3] = a2[0]; y[4] = a2[1]; y[5] = a2[2];

    // Jacobian of (old pose) wrt (new pose, new slot pos)     LinearSFMImp.cpp:538-632
    double R2[9], dA2[9], dB2[9], dG2[9], Ri[9], tmp[9];
    geom::rot_derivs(a2[0], a2[1], a2[2], R2, dA2, dB2, dG2);
    geom::mul_abt(R2, c.Q, Ri);
    PoseJac J;
    double v[3];
    geom::mul_abt(dA2, c.Q, tmp); geom::dangles(tmp, Ri, false, v);
    J.c1[0] = v[0]; J.c1[3] = v[1]; J.c1[6] = v[2];
    geom::mul_abt(dB2, c.Q, tmp); geom::dangles(tmp, Ri, false, v);
    J.c1[1] = v[0]; J.c1[4] = v[1]; J.c1[7] = v[2];
    geom::mul_abt(dG2, c.Q, tmp); geom::dangles(tmp, Ri, false, v);
    J.c1[2] = v[0]; J.c1[5] = v[1]; J.c1[8] = v[2];
    geom::mul_abt(R2, c.QA, tmp); geom::dangles(tmp, Ri, false, v);
    J.g2[0] = v[0]; J.g2[3] = v[1]; J.g2[6] = v[2];
    geom::mul_abt(R2, c.QB, tmp); geom::dangles(tmp, Ri, false, v);
    J.g2[1] = v[0]; J.g2[4] = v[1]; J.g2[7] = v[2];
    geom::mul_abt(R2, c.QG, tmp); geom::dangles(tmp, Ri, false, v);
    J.g2[2] = v[0]; J.g2[5] = v[1]; J.g2[8] = v[2];
    double dd[3] = {t2[0] - c.tn[0], t2[1] - c.tn[1], t2[2] - c.tn[2]};
    geom::mat3_vec(c.QA, dd, v); J.f2[0] = v[0]; J.f2[3] = v[1]; J.f2[6] = v[2];
    geom::mat3_vec(c.QB, dd, v); J.f2[1] = v[0]; J.f2[4] = v[1]; J.f2[7] = v[2];
    geom::mat3_vec(c.QG, dd, v); J.f2[2] = v[0]; J.f2[5] = v[1]; J.f2[8] = v[2];
#pragma unroll
    for (int i = 0; i < 9; i++) J.b1[i] = 0.0;
    pj[g] = J;
}

// slots [0,m) of the new U list: pairs with posID (LinearSFMImp.cpp:711-723)
__global__ void k_tf_uinit(DMap *__restrict__ out, const int *__restrict__ posePre, int K,
                           int totPose, const int *__restrict__ posID)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= totPose) return;
    int k = seg_find(posePre, K, g);
    int i = g - posePre[k];
    int pid = posID[k];
    out[k].Ui[i] = i <= pid ? i : pid;
    out[k].Uj[i] = i <= pid ? pid : i;
    double *u = out[k].U + 36 * (size_t)i;      // shared accumulation targets start from zero
#pragma unroll
    for (int q = 0; q < 36; q++) u[q] = 0.0;
}

__device__ __forceinline__ void dense_jac(const TfConst &c, const PoseJac &J, bool isPos,
                                          double *J1, double *J2)
{
    double s = isPos ? -1.0 : 1.0, e = isPos ? 0.0 : -1.0;
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int q = 0; q < 3; q++) {
            J1[6 * r + q] = s * c.Q[3 * r + q];     J1[6 * r + 3 + q] = J.b1[3 * r + q];
            J1[6 * (r + 3) + q] = 0.0;              J1[6 * (r + 3) + 3 + q] = J.c1[3 * r + q];
            J2[6 * r + q] = e * c.Q[3 * r + q];     J2[6 * r + 3 + q] = J.f2[3 * r + q];
            J2[6 * (r + 3) + q] = 0.0;              J2[6 * (r + 3) + 3 + q] = J.g2[3 * r + q];
        }
}

__device__ __forceinline__ void add_block(double *dst, const double *P, bool transpose)
{
#pragma unroll
    for (int r = 0; r < 6; r++)
#pragma unroll
        for (int q = 0; q < 6; q++) atomicAdd(dst + 6 * r + q, transpose ? P[6 * q + r] : P[6 * r + q]);
}

// one thread per old U block (LinearSFMImp.cpp:725-1266). U blocks are few (nU << nW).
__global__ void __launch_bounds__(128)
k_ucong(const DMap *__restrict__ in, DMap *__restrict__ out, const int *__restrict__ uPre,
        const int *__restrict__ posePre, int K, int totU, const TfConst *__restrict__ tc,
        const PoseJac *__restrict__ pj, const int *__restrict__ uScan)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= totU) return;
    int k = seg_find(uPre, K, g);
    int b = g - uPre[k];
    const DMap &M = in[k];
    const TfConst &c = tc[k];
    int pid = c.posID;
    int i = M.Ui[b], j = M.Uj[b];
    double I[36];
    sm::load<36>(M.U + 36 * (size_t)b, I);
    double J1i[36], J2i[36], J1j[36], J2j[36];
    dense_jac(c, pj[posePre[k] + i], i == pid, J1i, J2i);
    dense_jac(c, pj[posePre[k] + j], j == pid, J1j, J2j);
    double *Un = out[k].U;
    double T[36], P[36];

    // C_i^T I C_j -> (pos,pos)
    sm::mtm<6, 6, 6>(J2i, I, T);
    sm::mm<6, 6, 6>(T, J2j, P);
    add_block(Un + 36 * (size_t)pid, P, false);
    if (i != j) add_block(Un + 36 * (size_t)pid, P, true);
    // C_i^T I D_j -> (pos,j), stored in slot j
    sm::mm<6, 6, 6>(T, J1j, P);
    if (j >= pid) add_block(Un + 36 * (size_t)j, P, false);
    if (j <= pid && i != j) add_block(Un + 36 * (size_t)j, P, true);
    // D_i^T I D_j -> (i,j)
    sm::mtm<6, 6, 6>(J1i, I, T);
    sm::mm<6, 6, 6>(T, J1j, P);
    if (i == pid) add_block(Un + 36 * (size_t)j, P, false);
    else if (j == pid) add_block(Un + 36 * (size_t)i, P, false);
    else {
        int slot = M.m + (uScan[g] - uScan[uPre[k]]);
        sm::store<36>(Un + 36 * (size_t)slot, P);
        out[k].Ui[slot] = i;
        out[k].Uj[slot] = j;
    }
    // D_i^T I C_j -> (i,pos), stored in slot i
    sm::mm<6, 6, 6>(T, J2j, P);
    if (i <= pid) add_block(Un + 36 * (size_t)i, P, false);
    if (i >= pid && i != j) add_block(Un + 36 * (size_t)i, P, true);
}

// X = [Xt; Xb] (6 x N).  out = Jp^T X for the block-triangular pose Jacobian [[a,b],[0,c]]:
//   top = a^T Xt, bottom = b^T Xt + c^T Xb
template <int N>
__device__ __forceinline__ void jt_mul(const double *a, double sgn, const double *b, const double *c,
                                       bool has_b, const double *X, double *out)
{
    sm::mtm<3, 3, N>(a, X, out);
#pragma unroll
    for (int i = 0; i < 3 * N; i++) out[i] *= sgn;
    sm::mtm<3, 3, N>(c, X + 3 * N, out + 3 * N);
    if (has_b) sm::mtm_acc<3, 3, N>(b, X, out + 3 * N);
}

// one thread per feature (LinearSFMImp.cpp:1300-1915)
__global__ void __launch_bounds__(128)
k_wvcong(const DMap *__restrict__ in, DMap *__restrict__ out, const int *__restrict__ featPre,
         const int *__restrict__ posePre, int K, int totFeat, const TfConst *__restrict__ tc,
         const PoseJac *__restrict__ pj, const int *__restrict__ fScan)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    bool live = g < totFeat;
    int k = live ? seg_find(featPre, K, g) : 0;
    int f = live ? g - featPre[k] : 0;
    const DMap &M = in[k];
    const DMap &O = out[k];
    const TfConst &c = tc[k];
    const int pid = c.posID;

    double Q[9];
    sm::load<9>(c.Q, Q);
    double Tf[9];                    // columns QA d, QB d, QG d,  d = X' - t'
    double Wpos[18];                 // new block (posID, f)
    double G[36];                    // this feature's share of U'(pos,pos) (before symmetrising)
    double Vn[9];
    int kf = 0, w0 = 0, o0 = 0;
    if (live) {
        const double *x = M.featVal + 3 * (size_t)f;
        double d0[3] = {x[0] - c.t[0], x[1] - c.t[1], x[2] - c.t[2]};
        double xn[3];
        geom::mat3_vec(c.R, d0, xn);
        double *y = O.featVal + 3 * (size_t)f;
        y[0] = xn[0]; y[1] = xn[1]; y[2] = xn[2];
        O.featNo[f] = M.featNo[f];
        double d[3] = {xn[0] - c.tn[0], xn[1] - c.tn[1], xn[2] - c.tn[2]};
        double v[3];
        geom::mat3_vec(c.QA, d, v); Tf[0] = v[0]; Tf[3] = v[1]; Tf[6] = v[2];
        geom::mat3_vec(c.QB, d, v); Tf[1] = v[0]; Tf[4] = v[1]; Tf[7] = v[2];
        geom::mat3_vec(c.QG, d, v); Tf[2] = v[0]; Tf[5] = v[1]; Tf[8] = v[2];

        double V[9], VQ[9], VT[9], M1[9], M2[9];
        sm::load<9>(M.V + 9 * (size_t)f, V);
        sm::mm<3, 3, 3>(V, Q, VQ);
        sm::mm<3, 3, 3>(V, Tf, VT);
        sm::mtm<3, 3, 3>(Q, VQ, Vn);          // V' = Q^T V Q
        sm::mtm<3, 3, 3>(Tf, VQ, M1);         // T^T V Q
        sm::mtm<3, 3, 3>(Tf, VT, M2);         // T^T V T
        sm::store<9>(O.V + 9 * (size_t)f, Vn);
        // W'(pos,f) = C_f^T V D_f = [-V'; M1]
#pragma unroll
        for (int i = 0; i < 9; i++) { Wpos[i] = -Vn[i]; Wpos[9 + i] = M1[i]; }
        // C_f^T V C_f = [[V', -M1^T],[-M1, M2]]  (already symmetric -> goes in full, halve later)
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int q = 0; q < 3; q++) {
                G[6 * r + q] = 0.5 * Vn[3 * r + q];
                G[6 * r + 3 + q] = -0.5 * M1[3 * q + r];
                G[6 * (r + 3) + q] = -0.5 * M1[3 * r + q];
                G[6 * (r + 3) + 3 + q] = 0.5 * M2[3 * r + q];
            }
        w0 = M.wPtr[f];
        kf = M.wPtr[f + 1] - w0;
        o0 = fScan[g] - fScan[featPre[k]];
        O.wPtr[f] = o0;
        O.photo[o0] = pid;
        O.feature[o0] = f;
    } else {
#pragma unroll
        for (int i = 0; i < 36; i++) G[i] = 0.0;
    }

    int kmax = __reduce_max_sync(0xffffffffu, kf);
    int onext = o0 + 1;
    for (int jj = 0; jj < kmax; jj++) {
        bool act = jj < kf;
        int p = 0;
        double A3[36];
        if (act) {
            int j = w0 + jj;
            p = M.photo[j];
            double W[18], WQ[18], WT[18];
            sm::load<18>(M.W + 18 * (size_t)j, W);
            sm::mm<6, 3, 3>(W, Q, WQ);
            sm::mm<6, 3, 3>(W, Tf, WT);
            const PoseJac &J = pj[posePre[k] + p];
            bool isPos = (p == pid);
            double a1[18], a3[18];
            // D_p^T [WQ | WT]
            jt_mul<3>(Q, isPos ? -1.0 : 1.0, J.b1, J.c1, isPos, WQ, a1);
            jt_mul<3>(Q, isPos ? -1.0 : 1.0, J.b1, J.c1, isPos, WT, a3);
            if (isPos) {
#pragma unroll
                for (int i = 0; i < 18; i++) Wpos[i] += a1[i];
            } else {
                sm::store<18>(O.W + 18 * (size_t)onext, a1);
                O.photo[onext] = p;
                O.feature[onext] = f;
                onext++;
                // C_p^T [WQ | WT]  (zero for slot pos)
                double a2[18], a4[18];
                jt_mul<3>(Q, -1.0, J.f2, J.g2, true, WQ, a2);
                jt_mul<3>(Q, -1.0, J.f2, J.g2, true, WT, a4);
#pragma unroll
                for (int i = 0; i < 18; i++) Wpos[i] += a2[i];
                // C_p^T W C_f = [-a2 | a4]
#pragma unroll
                for (int r = 0; r < 6; r++)
#pragma unroll
                    for (int q = 0; q < 3; q++) {
                        G[6 * r + q] -= a2[3 * r + q];
                        G[6 * r + 3 + q] += a4[3 * r + q];
                    }
            }
            // D_p^T W C_f = [-a1 | a3] -> U'(p,pos), slot p
#pragma unroll
            for (int r = 0; r < 6; r++)
#pragma unroll
                for (int q = 0; q < 3; q++) {
                    A3[6 * r + q] = -a1[3 * r + q];
                    A3[6 * r + 3 + q] = a3[3 * r + q];
                }
            if (p > pid) {          // stored as (posID, p): transpose
                double Tt[36];
#pragma unroll
                for (int r = 0; r < 6; r++)
#pragma unroll
                    for (int q = 0; q < 6; q++) Tt[6 * r + q] = A3[6 * q + r];
#pragma unroll
                for (int i = 0; i < 36; i++) A3[i] = Tt[i];
            } else if (p == pid) {  // diagonal block gets X + X^T
                double Tt[36];
#pragma unroll
                for (int r = 0; r < 6; r++)
#pragma unroll
                    for (int q = 0; q < 6; q++) Tt[6 * r + q] = A3[6 * r + q] + A3[6 * q + r];
#pragma unroll
                for (int i = 0; i < 36; i++) A3[i] = Tt[i];
            }
        }
        sm::warp_agg_atomic_add<36>(O.U + 36 * (size_t)p, A3, act);
    }
    if (live) sm::store<18>(O.W + 18 * (size_t)o0, Wpos);
    // U'(pos,pos) += G + G^T
    double S[36];
#pragma unroll
    for (int r = 0; r < 6; r++)
#pragma unroll
        for (int q = 0; q < 6; q++) S[6 * r + q] = G[6 * r + q] + G[6 * q + r];
    sm::warp_agg_atomic_add<36>(O.U + 36 * (size_t)pid, S, live);
}

} // namespace

void OpMaps::build(const std::vector<DMap> &maps, cudaStream_t s)
{
    K = (int)maps.size();
    h = maps;
    posePre.assign(K + 1, 0); featPre.assign(K + 1, 0); uPre.assign(K + 1, 0); wPre.assign(K + 1, 0);
    long long tp = 0, tf = 0, tu = 0, tw = 0;
    for (int k = 0; k < K; k++) {
        tp += h[k].m; tf += h[k].n; tu += h[k].nU; tw += h[k].nW;
        if (tp > INT_MAX || tf > INT_MAX || tu > INT_MAX || tw > INT_MAX)
            throw LsfmError(LSFM_ERR_ARG, "level too large for 32-bit block indices");
        posePre[k + 1] = (int)tp; featPre[k + 1] = (int)tf; uPre[k + 1] = (int)tu; wPre[k + 1] = (int)tw;
    }
    totPose = (int)tp; totFeat = (int)tf; totU = (int)tu; totW = (int)tw;
    d.alloc(K, s); d.upload(h);
    dPosePre.alloc(K + 1, s); dPosePre.upload(posePre);
    dFeatPre.alloc(K + 1, s); dFeatPre.upload(featPre);
    dUPre.alloc(K + 1, s); dUPre.upload(uPre);
    dWPre.alloc(K + 1, s); dWPre.upload(wPre);
}

void OpMaps::build(const std::vector<MapHandle> &maps, cudaStream_t s)
{
    std::vector<DMap> v(maps.size());
    for (size_t i = 0; i < maps.size(); i++) v[i] = maps[i].d;
    build(v, s);
}

std::vector<MapHandle> alloc_maps(Context &ctx, std::vector<DMap> &shapes)
{
    size_t bytes = 0;
    for (auto &s : shapes) {
        bytes += Arena::pad(sizeof(int) * s.m) + Arena::pad(sizeof(double) * 6 * s.m) +
                 Arena::pad(sizeof(int) * s.n) + Arena::pad(sizeof(double) * 3 * s.n) +
                 Arena::pad(sizeof(double) * 36 * (size_t)s.nU) + 2 * Arena::pad(sizeof(int) * s.nU) +
                 Arena::pad(sizeof(double) * 18 * (size_t)s.nW) + 2 * Arena::pad(sizeof(int) * s.nW) +
                 Arena::pad(sizeof(double) * 9 * (size_t)s.n) + Arena::pad(sizeof(int) * (s.n + 1));
    }
    auto arena = std::make_shared<Arena>(bytes, ctx.stream);
    std::vector<MapHandle> out(shapes.size());
    for (size_t i = 0; i < shapes.size(); i++) {
        DMap &s = shapes[i];
        s.poseNo = arena->take<int>(s.m);
        s.poseVal = arena->take<double>(6 * (size_t)s.m);
        s.featNo = arena->take<int>(s.n);
        s.featVal = arena->take<double>(3 * (size_t)s.n);
        s.U = arena->take<double>(36 * (size_t)s.nU);
        s.Ui = arena->take<int>(s.nU);
        s.Uj = arena->take<int>(s.nU);
        s.W = arena->take<double>(18 * (size_t)s.nW);
        s.photo = arena->take<int>(s.nW);
        s.feature = arena->take<int>(s.nW);
        s.V = arena->take<double>(9 * (size_t)s.n);
        s.wPtr = arena->take<int>(s.n + 1);
        out[i].d = s;
        out[i].arena = arena;
    }
    return out;
}

// Byte layout of ONE map inside an arena of its own (same order / padding as alloc_maps): assigns
// the pointers of `s` relative to `base` and returns the size.  Used for device-to-device hand-over
// of a whole map between ranks: both sides derive identical offsets from the shape alone.
size_t map_layout(DMap &s, char *base)
{
    size_t off = 0;
    auto take = [&](size_t bytes) { char *p = base + off; off += Arena::pad(bytes); return p; };
    s.poseNo = (int *)take(sizeof(int) * s.m);
    s.poseVal = (double *)take(sizeof(double) * 6 * (size_t)s.m);
    s.featNo = (int *)take(sizeof(int) * s.n);
    s.featVal = (double *)take(sizeof(double) * 3 * (size_t)s.n);
    s.U = (double *)take(sizeof(double) * 36 * (size_t)s.nU);
    s.Ui = (int *)take(sizeof(int) * s.nU);
    s.Uj = (int *)take(sizeof(int) * s.nU);
    s.W = (double *)take(sizeof(double) * 18 * (size_t)s.nW);
    s.photo = (int *)take(sizeof(int) * s.nW);
    s.feature = (int *)take(sizeof(int) * s.nW);
    s.V = (double *)take(sizeof(double) * 9 * (size_t)s.n);
    s.wPtr = (int *)take(sizeof(int) * (s.n + 1));
    return off;
}

static void exclusive_scan(Context &ctx, const int *in, int *out, int n)
{
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, in, out, n, ctx.stream);
    DevBuf<char> tmp(tmp_bytes, ctx.stream);
    cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, in, out, n, ctx.stream);
}

std::vector<MapHandle> transform_stereo_batch(Context &ctx, const std::vector<MapHandle> &in,
                                              const std::vector<int> &newRef)
{
    const int K = (int)in.size();
    if (K == 0) return {};
    cudaStream_t s = ctx.stream;
    ctx.begin("transform");
    OpMaps A;
    A.build(in, s);
    DevBuf<int> dRef(K, s); dRef.upload(newRef);
    std::vector<int> init(K, INT_MAX);
    DevBuf<int> dPos(K, s); dPos.upload(init);
    const int TB = 256;
    int nl = 0;
    k_find_pos<<<ceil_div(A.totPose, TB), TB, 0, s>>>(A.d.p, A.dPosePre.p, K, A.totPose, dRef.p, dPos.p); nl++;
    DevBuf<int> uFlag(A.totU + 1, s), uScan(A.totU + 1, s), fCnt(A.totFeat + 1, s), fScan(A.totFeat + 1, s);
    k_count_u<<<ceil_div(A.totU + 1, TB), TB, 0, s>>>(A.d.p, A.dUPre.p, K, A.totU, dPos.p, uFlag.p); nl++;
    k_count_w<<<ceil_div(A.totFeat + 1, TB), TB, 0, s>>>(A.d.p, A.dFeatPre.p, K, A.totFeat, dPos.p, fCnt.p); nl++;
    exclusive_scan(ctx, uFlag.p, uScan.p, A.totU + 1); nl += 2;
    exclusive_scan(ctx, fCnt.p, fScan.p, A.totFeat + 1); nl += 2;
    DevBuf<int> dSizes(2 * (size_t)K, s);
    k_sizes<<<ceil_div(K, TB), TB, 0, s>>>(A.dUPre.p, A.dFeatPre.p, K, uScan.p, fScan.p, dSizes.p, dSizes.p + K); nl++;
    std::vector<int> hPos(K), hSizes(2 * (size_t)K);
    dPos.download(hPos.data(), K);
    dSizes.download(hSizes.data(), 2 * (size_t)K);
    CUDA_CHECK(cudaStreamSynchronize(s));
    KERNEL_CHECK();

    std::vector<DMap> shapes(K);
    double bytes = 0.0;
    for (int k = 0; k < K; k++) {
        if (hPos[k] == INT_MAX)
            throw LsfmError(LSFM_ERR_REF_NOT_FOUND, "Transform: pose " + std::to_string(newRef[k]) +
                                                        " is not in the state of the map");
        DMap &o = shapes[k];
        o = A.h[k];
        o.Ref = newRef[k];
        o.nU = A.h[k].m + hSizes[k];
        o.nW = hSizes[K + k];
        bytes += map_bytes(A.h[k]);
    }
    std::vector<MapHandle> out = alloc_maps(ctx, shapes);
    for (int k = 0; k < K; k++) bytes += map_bytes(out[k].d);
    OpMaps B;
    B.build(out, s);
    DevBuf<TfConst> tc(K, s);
    DevBuf<PoseJac> pj(A.totPose, s);
    k_tf_const<<<ceil_div(K, 64), 64, 0, s>>>(A.d.p, B.d.p, A.dPosePre.p, K, dPos.p, dSizes.p + K, tc.p, pj.p); nl++;
    k_tf_pose<<<ceil_div(A.totPose, 128), 128, 0, s>>>(A.d.p, B.d.p, A.dPosePre.p, K, A.totPose, tc.p, pj.p); nl++;
    k_tf_uinit<<<ceil_div(A.totPose, TB), TB, 0, s>>>(B.d.p, A.dPosePre.p, K, A.totPose, dPos.p); nl++;
    if (A.totU > 0) {
        k_ucong<<<ceil_div(A.totU, 128), 128, 0, s>>>(A.d.p, B.d.p, A.dUPre.p, A.dPosePre.p, K, A.totU,
                                                     tc.p, pj.p, uScan.p); nl++;
    }
    if (A.totFeat > 0) {
        k_wvcong<<<ceil_div(A.totFeat, 128), 128, 0, s>>>(A.d.p, B.d.p, A.dFeatPre.p, A.dPosePre.p, K,
                                                         A.totFeat, tc.p, pj.p, fScan.p); nl++;
    }
    KERNEL_CHECK();
    ctx.end(bytes, 0.0, nl);
    return out;
}
