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
//   * the W/V part is one pass over W (k_tf_chunk, transform_chunk.cuh: CTA per 128 features, one
//     thread per W block, per-feature and per-pose sums formed in shared memory / registers);
//     the per-pose sums are applied to U'(p,pos) and U'(pos,pos) once per pose (k_tf_posefin);
//     sums into one shared target are reduced across the warp before the FP64 atomics;
//   * output block lists keep the reference's exact order (Appendix C.1 of SURVEY.md): slots
//     [0,m) = pairs with posID, then surviving U blocks in input order; per feature the new
//     (posID,f) block first, then its surviving W blocks in input order.
#include "ops.h"
#include "geom.cuh"
#include "small_mat.cuh"
#include "det_accum.cuh"
#include <cub/cub.cuh>
#include <climits>
#include <type_traits>

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
    out[k].Uj[i] = i <= pid ? pid : i;          // the block values come from the record reduction
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

// Record writers of the deterministic U' accumulation (det_accum.cuh).  The new U list's slots [0,m)
// (pairs with posID) are shared targets: every old U block, every pose sum and every feature chunk
// adds into them.  Target index = global pose index of the slot; `none` = no contribution.
__device__ __forceinline__ void rec_write(int *__restrict__ key, double *__restrict__ val, size_t r, int target,
                                          const double *P)
{
    key[r] = target;
    double2 *dst = reinterpret_cast<double2 *>(val + 36 * r);
#pragma unroll
    for (int q = 0; q < 18; q++) dst[q] = make_double2(P[2 * q], P[2 * q + 1]);
}

// one thread per old U block (LinearSFMImp.cpp:725-1266): four products per block.  Everything that
// lands in the (pos,pos) slot is summed in the thread and written as ONE record of the map's hot-target
// list (ppKey/ppVal, reduced per map by k_pp_reduce); the other products are records 3g .. 3g+2 of the
// sorted list (targets: slot j; slot i or j; slot i) -- survivors (neither index == posID) are stored
// straight into their own slot.
__global__ void __launch_bounds__(128)
k_ucong(const DMap *__restrict__ in, DMap *__restrict__ out, const int *__restrict__ uPre,
        const int *__restrict__ posePre, int K, int totU, const TfConst *__restrict__ tc,
        const PoseJac *__restrict__ pj, const int *__restrict__ uScan,
        int *__restrict__ rkey, double *__restrict__ rval, int none,
        int *__restrict__ ppKey, double *__restrict__ ppVal)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= totU) return;
    int k = seg_find(uPre, K, g);
    int b = g - uPre[k];
    const DMap &M = in[k];
    const TfConst &c = tc[k];
    int pid = c.posID;
    int i = M.Ui[b], j = M.Uj[b];
    const int base = posePre[k];
    double I[36];
    sm::load<36>(M.U + 36 * (size_t)b, I);
    double J1i[36], J2i[36], J1j[36], J2j[36];
    dense_jac(c, pj[posePre[k] + i], i == pid, J1i, J2i);
    dense_jac(c, pj[posePre[k] + j], j == pid, J1j, J2j);
    double *Un = out[k].U;
    double T[36], P[36], PP[36];
    const size_t r0 = 3 * (size_t)g;
    // a product P (and/or its transpose) for slot `slot`: a record, or a term of PP when slot == pos
    auto emit = [&](size_t r, int slot, bool plain, bool transposed) {
        if (slot == pid) {
#pragma unroll
            for (int a = 0; a < 6; a++)
#pragma unroll
                for (int q = 0; q < 6; q++)
                    PP[6 * a + q] += (plain ? P[6 * a + q] : 0.0) + (transposed ? P[6 * q + a] : 0.0);
            rkey[r] = none;
            return;
        }
        double Q[36];
#pragma unroll
        for (int a = 0; a < 6; a++)
#pragma unroll
            for (int q = 0; q < 6; q++) Q[6 * a + q] = (plain ? P[6 * a + q] : 0.0) + (transposed ? P[6 * q + a] : 0.0);
        rec_write(rkey, rval, r, base + slot, Q);
    };

    // C_i^T I C_j -> (pos,pos)
    sm::mtm<6, 6, 6>(J2i, I, T);
    sm::mm<6, 6, 6>(T, J2j, P);
#pragma unroll
    for (int r = 0; r < 6; r++)
#pragma unroll
        for (int q = 0; q < 6; q++) PP[6 * r + q] = P[6 * r + q] + (i != j ? P[6 * q + r] : 0.0);
    // C_i^T I D_j -> (pos,j), stored in slot j (transposed when j < pos; both orientations when j == pos)
    sm::mm<6, 6, 6>(T, J1j, P);
    emit(r0, j, j >= pid, j <= pid && i != j);
    // D_i^T I D_j -> (i,j)
    sm::mtm<6, 6, 6>(J1i, I, T);
    sm::mm<6, 6, 6>(T, J1j, P);
    if (i == pid) emit(r0 + 1, j, true, false);
    else if (j == pid) emit(r0 + 1, i, true, false);
    else {
        rkey[r0 + 1] = none;
        int slot = M.m + (uScan[g] - uScan[uPre[k]]);
        sm::store<36>(Un + 36 * (size_t)slot, P);
        out[k].Ui[slot] = i;
        out[k].Uj[slot] = j;
    }
    // D_i^T I C_j -> (i,pos), stored in slot i
    sm::mm<6, 6, 6>(T, J2j, P);
    emit(r0 + 2, i, i <= pid, i >= pid && i != j);
    ppKey[g] = k;
    double2 *dst = reinterpret_cast<double2 *>(ppVal + 36 * (size_t)g);
#pragma unroll
    for (int q = 0; q < 18; q++) dst[q] = make_double2(PP[2 * q], PP[2 * q + 1]);
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

#include "transform_chunk.cuh"

// One WARP per pose: gathers the pose's sums SW_p = sum_f W_pf, SWT_p = sum_f W_pf T_f from the records
// the chunk kernel left behind (per chunk: pose bitmap + popcount prefix -> the pose's record slot with
// two loads; chunks strided over the lanes, lane-private partial sums combined by a fixed butterfly:
// bit-identical from run to run), then applies the pose Jacobians ONCE:
//     U'(p,pos) += oriented [-D_p^T SW Q | D_p^T SWT]   ->  record of the sorted list (target: slot p)
//     U'(pos,pos) += G + G^T, G = C_p^T [-SW Q | SWT]   ->  record of the map's hot-target list
__global__ void __launch_bounds__(128)
k_tf_posefin(const int *__restrict__ posePre, int K, int totPose,
             const TfConst *__restrict__ tc, const PoseJac *__restrict__ pj,
             const int *__restrict__ chunkPre, const unsigned *__restrict__ chunkBits, int bitsStride,
             const double *__restrict__ chunkRec, const double *__restrict__ poseAccSlow,
             int *__restrict__ rkey, double *__restrict__ rval, int none,
             int *__restrict__ ppKey, double *__restrict__ ppVal, const int *__restrict__ poseCR)
{
    const int lane = threadIdx.x & 31;
    const int gp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (gp >= totPose) return;                   // whole warps leave together
    const int k = seg_find(posePre, K, gp);
    const int p = gp - posePre[k];
    const int pw = p >> 5;
    const unsigned pbit = 1u << (p & 31);
    double acc[36];
#pragma unroll
    for (int q = 0; q < 36; q++) acc[q] = 0.0;
    // only the chunks between the first and the last one that recorded the pose (k_tf_chunk's running maxima);
    // the start is rounded down to a multiple of 32 chunks: every chunk keeps its lane, the sums keep their order
    int cbeg = chunkPre[k], cend = chunkPre[k];
    {
        const int vmin = poseCR[2 * (size_t)gp], vmax = poseCR[2 * (size_t)gp + 1];
        if (vmax > 0) { cbeg += ((0x7fffffff - vmin) - cbeg) & ~31; cend = vmax; }
    }
    for (int c = cbeg + lane; c < cend; c += 32) {
        const unsigned *gb = chunkBits + (size_t)c * 2 * bitsStride;
        const unsigned w = gb[pw];
        if (w & pbit) {
            const int slot = (int)gb[bitsStride + pw] + __popc(w & (pbit - 1u));
            const double2 *r = reinterpret_cast<const double2 *>(chunkRec + 36 * (32 * (size_t)c + slot));
#pragma unroll
            for (int q = 0; q < 18; q++) { double2 v = r[q]; acc[2 * q] += v.x; acc[2 * q + 1] += v.y; }
        }
    }
#pragma unroll
    for (int q = 0; q < 36; q++) acc[q] = sm::warp_sum(acc[q]);
    if (lane != 0) return;
    const TfConst &c = tc[k];
    const int pid = c.posID;
    const bool isPos = (p == pid);
    double Q[9], SW[18], SWT[18], SWQ[18];
    sm::load<9>(c.Q, Q);
#pragma unroll
    for (int q = 0; q < 18; q++) {
        SW[q] = acc[q] + poseAccSlow[36 * (size_t)gp + q];
        SWT[q] = acc[18 + q] + poseAccSlow[36 * (size_t)gp + 18 + q];
    }
    sm::mm<6, 3, 3>(SW, Q, SWQ);
    const PoseJac &J = pj[gp];
    double a1[18], a3[18];
    jt_mul<3>(Q, isPos ? -1.0 : 1.0, J.b1, J.c1, isPos, SWQ, a1);
    jt_mul<3>(Q, isPos ? -1.0 : 1.0, J.b1, J.c1, isPos, SWT, a3);
    double X[36], G2[36];
#pragma unroll
    for (int q = 0; q < 36; q++) { X[q] = 0.0; G2[q] = 0.0; }
#pragma unroll
    for (int r = 0; r < 6; r++)
#pragma unroll
        for (int q = 0; q < 6; q++) {
            double xrq = (q < 3) ? -a1[3 * r + q] : a3[3 * r + q - 3];
            if (p < pid) X[6 * r + q] = xrq;
            else if (p > pid) X[6 * q + r] = xrq;
            else { G2[6 * r + q] += xrq; G2[6 * q + r] += xrq; }
        }
    if (!isPos) {
        double a2[18], a4[18];
        jt_mul<3>(Q, -1.0, J.f2, J.g2, true, SWQ, a2);
        jt_mul<3>(Q, -1.0, J.f2, J.g2, true, SWT, a4);
#pragma unroll
        for (int r = 0; r < 6; r++)
#pragma unroll
            for (int q = 0; q < 6; q++) {
                double grq = (q < 3) ? -a2[3 * r + q] : a4[3 * r + q - 3];
                G2[6 * r + q] += grq;
                G2[6 * q + r] += grq;
            }
    }
    if (!isPos) rec_write(rkey, rval, (size_t)gp, posePre[k] + p, X); else rkey[gp] = none;
    rec_write(ppKey, ppVal, (size_t)gp, k, G2);
}

// U'(pos,pos) of map k = sum of the map's hot-target records: its U blocks' (run 0), its poses' (run 1)
// and its feature chunks' (run 2).  Each run lies map-contiguously at a WIN-aligned offset, so whole
// windows inside a map's range were pre-summed by det::k_window.  One CTA per map, fixed order.
struct PPRuns { int off[3]; const int *pre[3]; };
__global__ void __launch_bounds__(256)
k_pp_reduce(PPRuns R, const double *__restrict__ val, const double *__restrict__ part,
            const TfConst *__restrict__ tc, DMap *__restrict__ out)
{
    constexpr int NS = 7, WIN = det::WIN;
    __shared__ double sh[NS][36];
    const int k = blockIdx.x, tid = threadIdx.x;
    int lo[3], nHead[3], c0[3], nWin[3], tailBeg[3], nIt[3];
    int nItems = 0;
#pragma unroll
    for (int u = 0; u < 3; u++) {
        const int a = R.off[u] + R.pre[u][k], b = R.off[u] + R.pre[u][k + 1];
        int w0 = (a + WIN - 1) / WIN, w1 = b / WIN, he = b, tb = b;
        if (w1 > w0) { he = w0 * WIN; tb = w1 * WIN; } else { w0 = w1 = 0; }
        lo[u] = a; nHead[u] = he - a; c0[u] = w0; nWin[u] = w1 - w0; tailBeg[u] = tb;
        nIt[u] = (he - a) + (w1 - w0) + (b - tb);
        nItems += nIt[u];
    }
    const int s = tid / 36, q = tid - 36 * s;
    if (s < NS) {
        auto item = [&](int i) -> double {
#pragma unroll
            for (int u = 0; u < 3; u++) {
                if (i < nIt[u]) {
                    if (i < nHead[u]) return val[36 * (size_t)(lo[u] + i) + q];
                    i -= nHead[u];
                    if (i < nWin[u]) return part[36 * (size_t)(c0[u] + i) + q];
                    return val[36 * (size_t)(tailBeg[u] + (i - nWin[u])) + q];
                }
                i -= nIt[u];
            }
            return 0.0;
        };
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int i = s;
        for (; i + 3 * NS < nItems; i += 4 * NS) {
            a0 += item(i); a1 += item(i + NS); a2 += item(i + 2 * NS); a3 += item(i + 3 * NS);
        }
        if (i < nItems) a0 += item(i);
        if (i + NS < nItems) a1 += item(i + NS);
        if (i + 2 * NS < nItems) a2 += item(i + 2 * NS);
        sh[s][q] = (a0 + a1) + (a2 + a3);
    }
    __syncthreads();
    if (tid < 36) {
        double sum = 0.0;
        const int ns = min(NS, nItems);
        for (int i = 0; i < ns; i++) sum += sh[i][tid];
        out[k].U[36 * (size_t)tc[k].posID + tid] = sum;
    }
}

__global__ void k_fill_int(int *__restrict__ a, int n, int v)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < n) a[g] = v;
}

// target t = global pose index = slot t - posePre[k] of map k's new U list
struct ApplyU {
    DMap *out;
    const int *posePre;
    int K;
    __device__ void operator()(int t, int q, double sum, int cnt) const
    {
        if (cnt == 0) return;            // only the (pos,pos) slot has no record here: k_pp_reduce writes it
        const int k = seg_find(posePre, K, t);
        out[k].U[36 * (size_t)(t - posePre[k]) + q] = sum;
    }
};



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
    // one blob: [DMap x K][posePre][featPre][uPre][wPre]  (each prefix K+1 ints, 16-byte aligned)
    size_t oD = 0, szD = sizeof(DMap) * (size_t)K;
    size_t szP = (sizeof(int) * (size_t)(K + 1) + 15) & ~(size_t)15;
    size_t o1 = (szD + 15) & ~(size_t)15, o2 = o1 + szP, o3 = o2 + szP, o4 = o3 + szP, total = o4 + szP;
    std::vector<char> host(total, 0);
    if (K) memcpy(host.data() + oD, h.data(), szD);
    memcpy(host.data() + o1, posePre.data(), sizeof(int) * (K + 1));
    memcpy(host.data() + o2, featPre.data(), sizeof(int) * (K + 1));
    memcpy(host.data() + o3, uPre.data(), sizeof(int) * (K + 1));
    memcpy(host.data() + o4, wPre.data(), sizeof(int) * (K + 1));
    blob.alloc(total, s);
    blob.upload(host.data(), total);
    d.p = (DMap *)(blob.p + oD);
    dPosePre.p = (int *)(blob.p + o1);
    dFeatPre.p = (int *)(blob.p + o2);
    dUPre.p = (int *)(blob.p + o3);
    dWPre.p = (int *)(blob.p + o4);
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
    static const bool dbgT = getenv("LSFM_DEBUG_T") != nullptr;
    auto tq0 = std::chrono::steady_clock::now();
    dPos.download(hPos.data(), K);
    dSizes.download(hSizes.data(), 2 * (size_t)K);
    CUDA_CHECK(cudaStreamSynchronize(s));
    auto tq1 = std::chrono::steady_clock::now();
    ctx.idle_begin();
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
    auto tq2 = std::chrono::steady_clock::now();
    std::vector<MapHandle> out = alloc_maps(ctx, shapes);
    auto tq3 = std::chrono::steady_clock::now();
    for (int k = 0; k < K; k++) bytes += map_bytes(out[k].d);
    OpMaps B;
    B.build(out, s);
    auto tq4 = std::chrono::steady_clock::now();
    DevBuf<TfConst> tc(K, s);
    DevBuf<PoseJac> pj(A.totPose, s);
    ctx.idle_end(0);
    if (dbgT) {
        auto us = [](auto a, auto b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
        fprintf(stderr, "[transform K=%d] wait %.0f us | shapes %.0f alloc_maps %.0f build %.0f tc/pj %.0f us\n", K, us(tq0, tq1),
                us(tq1, tq2), us(tq2, tq3), us(tq3, tq4), us(tq4, std::chrono::steady_clock::now()));
    }
    k_tf_const<<<ceil_div(K, 64), 64, 0, s>>>(A.d.p, B.d.p, A.dPosePre.p, K, dPos.p, dSizes.p + K, tc.p, pj.p); nl++;
    k_tf_pose<<<ceil_div(A.totPose, 128), 128, 0, s>>>(A.d.p, B.d.p, A.dPosePre.p, K, A.totPose, tc.p, pj.p); nl++;
    k_tf_uinit<<<ceil_div(A.totPose, TB), TB, 0, s>>>(B.d.p, A.dPosePre.p, K, A.totPose, dPos.p); nl++;
    // feature chunks of the one-pass W/V kernel (transform_chunk.cuh): CTA per <= 128 consecutive features
    std::vector<tfc::Chunk> chunks;
    std::vector<int> chunkPre(K + 1, 0);
    int maxWords = 1;
    for (int k = 0; k < K; k++) {
        maxWords = std::max(maxWords, (A.h[k].m + 31) / 32);
        // the join that produced the map may have narrowed some chunks (> 30 poses); the features are the same
        const std::vector<int> *cs = in[k].chunkStarts.get();
        if (cs && !cs->empty() && cs->front() == 0 && cs->back() == A.h[k].n) {
            for (size_t i = 0; i + 1 < cs->size(); i++)
                if ((*cs)[i + 1] > (*cs)[i]) chunks.push_back({k, (*cs)[i], (*cs)[i + 1]});
            out[k].chunkStarts = in[k].chunkStarts;
        } else {
            for (int f0 = 0; f0 < A.h[k].n; f0 += tfc::TC_FCH)
                chunks.push_back({k, f0, std::min(A.h[k].n, f0 + tfc::TC_FCH)});
        }
        chunkPre[k + 1] = (int)chunks.size();
    }
    const int nChunks = (int)chunks.size();
    // Deterministic U' accumulation (det_accum.cuh).
    //   sorted list : target = global pose index of the slot, `none` = totPose
    //                 [0, 3 totU) k_ucong | [.., + totPose) k_tf_posefin
    //   hot list    : everything that lands in a map's (pos,pos) slot, map-contiguous by construction:
    //                 run 0 [0, totU) k_ucong | run 1 k_tf_posefin | run 2 the feature chunks;
    //                 runs start at WIN-aligned offsets, key = map index (padding: K)
    const int none = A.totPose;
    const size_t recP = 3 * (size_t)A.totU;
    const size_t nrec = recP + (size_t)A.totPose;
    auto up = [](size_t v) { return (v + det::WIN - 1) / det::WIN * det::WIN; };
    const size_t ppP = up((size_t)A.totU), ppC = ppP + up((size_t)A.totPose), ppN = ppC + up((size_t)nChunks);
    if (nrec > 0x7fffffffull || ppN > 0x7fffffffull)
        throw LsfmError(LSFM_ERR_ARG, "level too large for 32-bit record indices");
    DevBuf<int> rkey(nrec, s), ppKey(ppN, s);
    DevBuf<double> rval(36 * nrec, s), ppVal(36 * ppN, s);
    k_fill_int<<<ceil_div((long long)ppN, TB), TB, 0, s>>>(ppKey.p, (int)ppN, K); nl++;
    if (A.totU > 0) {
        k_ucong<<<ceil_div(A.totU, 128), 128, 0, s>>>(A.d.p, B.d.p, A.dUPre.p, A.dPosePre.p, K, A.totU,
                                                     tc.p, pj.p, uScan.p, rkey.p, rval.p, none, ppKey.p, ppVal.p); nl++;
    }
    DevBuf<tfc::Chunk> dChunks(std::max(nChunks, 1), s);
    DevBuf<int> dChunkPre(K + 1, s);
    const int bitsStride = maxWords;
    DevBuf<unsigned> chunkBits(2 * (size_t)bitsStride * std::max(nChunks, 1), s);
    DevBuf<double> chunkRec(36 * 32 * (size_t)std::max(nChunks, 1), s);
    DevBuf<double> poseAcc(36 * (size_t)A.totPose, s);      // slow-path (chunk overflow) sums only
    DevBuf<int> poseCR(2 * (size_t)std::max(A.totPose, 1), s);   // per pose: first / last chunk with a record of it
    dChunkPre.upload(chunkPre);
    poseAcc.zero();
    poseCR.zero();
    // Chunks with more than TC_CMAX distinct poses can only occur in maps whose join already had to split
    // chunks (landmarks seen by more than 30 poses stay in overflowing chunks): only then the slow path's
    // fixed-point bookkeeping is set up (k_tf_slow_pre).
    static const bool force_ovf = getenv("LSFM_FORCE_OVERFLOW") != nullptr;   // test hook: slow path
    bool maySlow = force_ovf;
    for (int k = 0; k < K; k++) maySlow = maySlow || (bool)in[k].chunkStarts;
    DevBuf<int> pexp(maySlow ? 36 * (size_t)A.totPose : 1, s), pcnt(maySlow ? (size_t)A.totPose : 1, s);
    DevBuf<long long> poseFx(maySlow ? 36 * (size_t)A.totPose : 1, s);
    if (nChunks > 0) {
        dChunks.upload(chunks);
        const int cmaxUse = force_ovf ? 4 : tfc::TC_CMAX;
        if (maySlow) {
            CUDA_CHECK(cudaMemsetAsync(pexp.p, 0x80, sizeof(int) * 36 * (size_t)A.totPose, s));     // < -2000: nothing
            CUDA_CHECK(cudaMemsetAsync(pcnt.p, 0, sizeof(int) * (size_t)A.totPose, s));
            CUDA_CHECK(cudaMemsetAsync(poseFx.p, 0, sizeof(long long) * 36 * (size_t)A.totPose, s));
            const size_t shp = sizeof(unsigned) * (size_t)maxWords;
            ctx.ensure_smem((const void *)tfc::k_tf_slow_pre, shp);
            tfc::k_tf_slow_pre<<<nChunks, 128, shp, s>>>(A.d.p, dChunks.p, A.dPosePre.p, tc.p, cmaxUse, pexp.p, pcnt.p); nl++;
        }
        const size_t shb = tfc::Layout::bytes(maxWords);
        ctx.ensure_smem((const void *)tfc::k_tf_chunk, shb);
        // stage timer of its own: this one kernel carries the W/V bytes of the transform
        //   in : W 144 + photo 4 per block; V 72 + X 24 + id 4 + wPtr 4 per feature
        //   out: W 144 + photo/feature 8 per block; V 72 + X 24 + id 4 + wPtr 4 per feature
        ctx.end(0.0, 0.0, nl);
        nl = 0;
        ctx.begin("transform.wv");
        tfc::k_tf_chunk<<<nChunks, tfc::TC_THREADS, shb, s>>>(A.d.p, B.d.p, dChunks.p, A.dFeatPre.p, A.dPosePre.p,
                                                            tc.p, pj.p, fScan.p, pexp.p, pcnt.p, poseFx.p, cmaxUse,
                                                            chunkBits.p, bitsStride, chunkRec.p, ppKey.p + ppC,
                                                            ppVal.p + 36 * ppC, poseCR.p);
        {
            double wvBytes = 0.0;
            for (int k = 0; k < K; k++)
                wvBytes += 148.0 * A.h[k].nW + 152.0 * out[k].d.nW + 2.0 * 104.0 * A.h[k].n;
            bytes -= wvBytes;
            ctx.end(wvBytes, 0.0, 1);
        }
        ctx.begin("transform");
        if (maySlow) {
            tfc::k_tf_slow_convert<<<ceil_div(36ll * A.totPose, 256), 256, 0, s>>>(36 * A.totPose, pexp.p, pcnt.p, poseFx.p,
                                                                                 poseAcc.p); nl++;
        }
    }
    k_tf_posefin<<<ceil_div(32ll * A.totPose, 128), 128, 0, s>>>(A.dPosePre.p, K, A.totPose, tc.p, pj.p, dChunkPre.p,
                                                              chunkBits.p, bitsStride, chunkRec.p, poseAcc.p,
                                                              rkey.p + recP, rval.p + 36 * recP, none,
                                                              ppKey.p + ppP, ppVal.p + 36 * ppP, poseCR.p); nl++;
    {
        det::Sorted srt;
        nl += det::sort_records(ctx, rkey.p, (int)nrec, none, srt);
        nl += det::reduce<36>(ctx, srt, rval.p, A.totPose, ApplyU{B.d.p, A.dPosePre.p, K});
        // hot targets: window sums over the hot list as it lies (chunk keys are implied by the layout:
        // a window is whole iff it lies inside one map's range, which k_pp_reduce decides from the prefixes)
        const int nwin = (int)(ppN / det::WIN);
        DevBuf<double> part(36 * (size_t)std::max(nwin, 1), s);
        if (nwin > 0) { det::k_window<36><<<nwin, 256, 0, s>>>(ppKey.p, nullptr, (int)ppN, K, ppVal.p, part.p); nl++; }
        PPRuns R;
        R.off[0] = 0; R.off[1] = (int)ppP; R.off[2] = (int)ppC;
        R.pre[0] = A.dUPre.p; R.pre[1] = A.dPosePre.p; R.pre[2] = dChunkPre.p;
        k_pp_reduce<<<K, 256, 0, s>>>(R, ppVal.p, part.p, tc.p, B.d.p); nl++;
    }
    KERNEL_CHECK();
    ctx.end(bytes, 0.0, nl);
    return out;
}
