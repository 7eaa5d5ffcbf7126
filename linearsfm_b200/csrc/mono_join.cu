// Batched monocular join + solve: rows a16-a17 of SURVEY 8(a).
//
// Reference: CLinearSFMImp::lmj_LinearLS_PF3DMono, LinearSFMImp.cpp:7282-7874, then
//            lmj_solveLinearSFMMono, LinearSFMImp.cpp:6756-7041 (gauge handling in solve.cu).
// Differences from the stereo join (join.cu), all restated here:
//   * End (already expressed in Cur's frame) and Cur share TWO poses: Cur's Ref pose is End's slot
//     posID1 (the all-zero pose), Cur's ScaP pose is End's slot posID2; joint poses = End's m1 poses
//     followed by Cur's other poses in order (CurPose2 map, 7383-7409), m = m1 + m2 - 2;
//   * every U / W block touching posID1 is dropped (7482, 7531, 7598, 7645, 7720);
//   * Cur's (posID2,posID2) U block is ADDED to End's last (posID2,posID2) block if End has one
//     (Fl/FlA, 7484-7488, 7533-7542); for a common feature Cur's (posID2,f) W block is added to
//     End's (posID2,f) block if present (7600-7604, 7647-7660);
//   * the angles of the shared ScaP pose are brought to the same 2*pi branch in both maps before
//     they enter the right-hand side (wrap-around with PI = 3.1415926, 7427-7465);
//   * solver gauge: Ref = posID1, Fix row = 6*posID2 + End.Fix, value End.Sign (7797-7801).
// First (correctness) version: one thread per pose / U block / joint feature.
#include "join_common.cuh"
#include "geom.cuh"
#include "small_mat.cuh"
#include <cub/cub.cuh>
#include <climits>

namespace {

using namespace joinc;

struct MJ {                 // per-join constants
    int posID1, posID2;     // End slots of the shared Ref / ScaP poses
    int cRef, cSca;         // Cur slots of its Ref / ScaP poses
    int flU;                // index of End's last kept (posID2,posID2) U block, or -1
    double wrpE[3], wrpC[3];   // wrapped angles of the ScaP pose in End / Cur
};

__global__ void k_mj_find(const DMap *__restrict__ E, const DMap *__restrict__ C,
                          const int *__restrict__ posePreE, const int *__restrict__ posePreC, int K,
                          int totE, int totC, int *__restrict__ slots /* [4][K] */)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < totE) {
        int k = seg_find(posePreE, K, g);
        int p = g - posePreE[k];
        int no = E[k].poseNo[p];
        if (no == -E[k].Ref) atomicMin(&slots[0 * K + k], p);
        if (no == -E[k].ScaP) atomicMin(&slots[1 * K + k], p);
    } else if (g < totE + totC) {
        int gc = g - totE;
        int k = seg_find(posePreC, K, gc);
        int p = gc - posePreC[k];
        int no = C[k].poseNo[p];
        if (no == -C[k].Ref) atomicMin(&slots[2 * K + k], p);
        else if (no == -C[k].ScaP) atomicMin(&slots[3 * K + k], p);       // else-if as in 7385-7394
    }
}

__device__ __forceinline__ void wrap3(double *w1, double *w2)
{
    const double PI = LSFM_PI;
    for (int i = 0; i < 3; i++) {
        if (w1[i] > PI) { int t = (int)(w1[i] / (2 * PI)); w1[i] -= (t + 1) * (2 * PI); }
        if (w1[i] < -PI) { int t = (int)(w1[i] / (2 * PI)); w1[i] -= (t - 1) * (2 * PI); }
        if (w2[i] > PI) { int t = (int)(w2[i] / (2 * PI)); w2[i] -= (t + 1) * (2 * PI); }
        if (w2[i] < -PI) { int t = (int)(w2[i] / (2 * PI)); w2[i] -= (t - 1) * (2 * PI); }
        double err = w2[i] - w1[i];
        if (err > PI) w2[i] -= 2 * PI;
        else if (err < -PI) w2[i] += 2 * PI;
    }
}

__global__ void k_mj_const(const DMap *__restrict__ E, const DMap *__restrict__ C, int K,
                           const int *__restrict__ slots, MJ *__restrict__ mj)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    MJ c;
    c.posID1 = slots[0 * K + k]; c.posID2 = slots[1 * K + k];
    c.cRef = slots[2 * K + k]; c.cSca = slots[3 * K + k];
    c.flU = -1;
    for (int i = 0; i < 3; i++) {
        c.wrpE[i] = E[k].poseVal[6 * (size_t)c.posID2 + 3 + i];
        c.wrpC[i] = C[k].poseVal[6 * (size_t)c.cSca + 3 + i];
    }
    wrap3(c.wrpE, c.wrpC);
    mj[k] = c;
}

// joint index of Cur pose i (CurPose2, 7383-7409)
__device__ __forceinline__ int cur_pose_joint(const MJ &c, int m1, int i)
{
    if (i == c.cRef) return c.posID1;
    if (i == c.cSca) return c.posID2;
    return m1 + i - (i > c.cRef ? 1 : 0) - (i > c.cSca ? 1 : 0);
}

// estimate of a pose with the wrapped ScaP angles substituted
__device__ __forceinline__ void load_pose(const DMap &M, int p, int scaSlot, const double *wr, double *x)
{
    sm::load<6>(M.poseVal + 6 * (size_t)p, x);
    if (p == scaSlot) { x[3] = wr[0]; x[4] = wr[1]; x[5] = wr[2]; }
}

// U bookkeeping: keep flags for End then Cur blocks of every join; last End (posID2,posID2) block
__global__ void k_mj_uflag(const DMap *__restrict__ E, const DMap *__restrict__ C,
                           const int *__restrict__ uPreS /* prefix of nU1+nU2 */, int K, int totS,
                           MJ *__restrict__ mj, int *__restrict__ flag)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g > totS) return;
    if (g == totS) { flag[g] = 0; return; }
    int k = seg_find(uPreS, K, g);
    int b = g - uPreS[k];
    const MJ &c = mj[k];
    if (b < E[k].nU) {
        int i = E[k].Ui[b], j = E[k].Uj[b];
        bool keep = (i != c.posID1 && j != c.posID1);
        flag[g] = keep;
        if (keep && i == c.posID2 && j == c.posID2) atomicMax(&mj[k].flU, b);
    } else {
        flag[g] = 0;        // decided by k_mj_uflag_cur once flU is known
    }
}

__global__ void k_mj_uflag_cur(const DMap *__restrict__ E, const DMap *__restrict__ C,
                               const int *__restrict__ uPreS, int K, int totS, const MJ *__restrict__ mj,
                               int *__restrict__ flag)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= totS) return;
    int k = seg_find(uPreS, K, g);
    int b = g - uPreS[k];
    if (b < E[k].nU) return;
    const MJ &c = mj[k];
    int sb = b - E[k].nU;
    int ci = cur_pose_joint(c, E[k].m, C[k].Ui[sb]), cj = cur_pose_joint(c, E[k].m, C[k].Uj[sb]);
    bool drop = (ci == c.posID1 || cj == c.posID1);
    bool merge = (!drop && ci == c.posID2 && cj == c.posID2 && c.flU >= 0);
    flag[g] = (!drop && !merge) ? 1 : 0;
}

// per joint feature: number of W blocks it keeps
__global__ void k_mj_wcount(const DMap *__restrict__ E, const DMap *__restrict__ C,
                            const int *__restrict__ featPreJ, int K, int totJ, const MJ *__restrict__ mj,
                            const int *__restrict__ curOfJoint, int *__restrict__ cnt)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g > totJ) return;
    if (g == totJ) { cnt[g] = 0; return; }
    int k = seg_find(featPreJ, K, g);
    int jf = g - featPreJ[k];
    const MJ &c = mj[k];
    const DMap &Em = E[k];
    const DMap &Cm = C[k];
    int n = 0;
    bool fl = false;
    if (jf < Em.n)
        for (int j = Em.wPtr[jf]; j < Em.wPtr[jf + 1]; j++) {
            int p = Em.photo[j];
            if (p != c.posID1) { n++; if (p == c.posID2) fl = true; }
        }
    int cf = curOfJoint[g];
    if (cf >= 0)
        for (int j = Cm.wPtr[cf]; j < Cm.wPtr[cf + 1]; j++) {
            int p = cur_pose_joint(c, Em.m, Cm.photo[j]);
            if (p == c.posID1) continue;
            if (p == c.posID2 && fl && jf < Em.n) continue;      // merged into End's block
            n++;
        }
    cnt[g] = n;
}

__global__ void k_mj_sizes(const int *__restrict__ uPreS, const int *__restrict__ featPreJ, int K,
                           const int *__restrict__ uScan, const int *__restrict__ wScan,
                           int *__restrict__ nUJ, int *__restrict__ nWJ)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    nUJ[k] = uScan[uPreS[k + 1]] - uScan[uPreS[k]];
    nWJ[k] = wScan[featPreJ[k + 1]] - wScan[featPreJ[k]];
}

__global__ void k_mj_pose(const DMap *__restrict__ E, const DMap *__restrict__ C, DMap *__restrict__ J,
                          const int *__restrict__ posePreC, int K, int totC, const MJ *__restrict__ mj)
{
    // End poses copy 1:1; Cur poses go to their joint slot unless shared
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= totC) return;
    int k = seg_find(posePreC, K, g);
    int i = g - posePreC[k];
    const MJ &c = mj[k];
    if (i == c.cRef || i == c.cSca) return;
    J[k].poseNo[cur_pose_joint(c, E[k].m, i)] = C[k].poseNo[i];
}

__global__ void k_mj_pose_end(const DMap *__restrict__ E, DMap *__restrict__ J,
                              const int *__restrict__ posePreE, int K, int totE)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= totE) return;
    int k = seg_find(posePreE, K, g);
    int i = g - posePreE[k];
    J[k].poseNo[i] = E[k].poseNo[i];
}

__global__ void __launch_bounds__(64)
k_mj_u(const DMap *__restrict__ E, const DMap *__restrict__ C, DMap *__restrict__ J,
       const int *__restrict__ uPreS, const int *__restrict__ posePreJ, int K, int totS,
       const MJ *__restrict__ mj, const int *__restrict__ flag, const int *__restrict__ uScan,
       double *__restrict__ eP)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= totS) return;
    int k = seg_find(uPreS, K, g);
    int b = g - uPreS[k];
    const MJ &c = mj[k];
    const bool fromE = b < E[k].nU;
    const DMap &S = fromE ? E[k] : C[k];
    const int sb = fromE ? b : b - E[k].nU;
    const int m1 = E[k].m;
    int si = S.Ui[sb], sj = S.Uj[sb];
    int ji = fromE ? si : cur_pose_joint(c, m1, si);
    int jj = fromE ? sj : cur_pose_joint(c, m1, sj);
    if (ji == c.posID1 || jj == c.posID1) return;              // dropped
    double U[36], xi[6], xj[6], y[6];
    sm::load<36>(S.U + 36 * (size_t)sb, U);
    const int scaSlot = fromE ? c.posID2 : c.cSca;
    const double *wr = fromE ? c.wrpE : c.wrpC;
    load_pose(S, sj, scaSlot, wr, xj);
    double *ePk = eP + 6 * (size_t)posePreJ[k];
    if (!flag[g]) {
        // Cur's (posID2,posID2) block merged into End's (7533-7542)
        int dst = uScan[uPreS[k] + c.flU] - uScan[uPreS[k]];
        double *u = J[k].U + 36 * (size_t)dst;
        for (int q = 0; q < 36; q++) atomicAdd(u + q, U[q]);
        sm::mm<6, 6, 1>(U, xj, y);
        for (int q = 0; q < 6; q++) atomicAdd(ePk + 6 * c.posID2 + q, y[q]);
        return;
    }
    int dst = uScan[g] - uScan[uPreS[k]];
    double *u = J[k].U + 36 * (size_t)dst;
    // the merge target may receive atomic adds from a Cur block: add instead of store
    for (int q = 0; q < 36; q++) atomicAdd(u + q, U[q]);
    J[k].Ui[dst] = ji;
    J[k].Uj[dst] = jj;
    sm::mm<6, 6, 1>(U, xj, y);
    for (int q = 0; q < 6; q++) atomicAdd(ePk + 6 * ji + q, y[q]);
    if (si != sj) {
        load_pose(S, si, scaSlot, wr, xi);
        sm::mtm<6, 6, 1>(U, xi, y);
        for (int q = 0; q < 6; q++) atomicAdd(ePk + 6 * jj + q, y[q]);
    }
}

// one thread per joint feature (7584-7758)
__global__ void __launch_bounds__(64)
k_mj_feat(const DMap *__restrict__ E, const DMap *__restrict__ C, DMap *__restrict__ J,
          const int *__restrict__ featPreJ, const int *__restrict__ posePreJ, int K, int totJ,
          const MJ *__restrict__ mj, const int *__restrict__ curOfJoint, const int *__restrict__ wScan,
          double *__restrict__ eP, double *__restrict__ eF)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= totJ) return;
    int k = seg_find(featPreJ, K, g);
    int jf = g - featPreJ[k];
    const MJ &c = mj[k];
    const DMap &Em = E[k];
    const DMap &Cm = C[k];
    const DMap &Jm = J[k];
    const int m1 = Em.m;
    const bool hasE = jf < Em.n;
    const int cf = curOfJoint[g];
    double *ePk = eP + 6 * (size_t)posePreJ[k];
    double ef[3] = {0, 0, 0}, V[9], xe[3] = {0, 0, 0}, xc[3] = {0, 0, 0};
    for (int i = 0; i < 9; i++) V[i] = 0.0;
    int o = wScan[g] - wScan[featPreJ[k]];
    Jm.wPtr[jf] = o;
    int flSlot = -1;                                 // joint position of End's (posID2,f) block
    if (hasE) {
        sm::load<9>(Em.V + 9 * (size_t)jf, V);
        sm::load<3>(Em.featVal + 3 * (size_t)jf, xe);
        sm::mm<3, 3, 1>(V, xe, ef);
        Jm.featNo[jf] = Em.featNo[jf];
        for (int j = Em.wPtr[jf]; j < Em.wPtr[jf + 1]; j++) {
            int p = Em.photo[j];
            if (p == c.posID1) continue;
            if (p == c.posID2) flSlot = o;
            double W[18], y[6], xp[6], t[3];
            sm::load<18>(Em.W + 18 * (size_t)j, W);
            sm::store<18>(Jm.W + 18 * (size_t)o, W);
            Jm.photo[o] = p; Jm.feature[o] = jf;
            sm::mm<6, 3, 1>(W, xe, y);
            for (int q = 0; q < 6; q++) atomicAdd(ePk + 6 * p + q, y[q]);
            load_pose(Em, p, c.posID2, c.wrpE, xp);
            sm::mtm<3, 6, 1>(W, xp, t);
            ef[0] += t[0]; ef[1] += t[1]; ef[2] += t[2];
            o++;
        }
    }
    if (cf >= 0) {
        double Vc[9], t[3];
        sm::load<9>(Cm.V + 9 * (size_t)cf, Vc);
        sm::load<3>(Cm.featVal + 3 * (size_t)cf, xc);
        sm::mm<3, 3, 1>(Vc, xc, t);
        for (int i = 0; i < 9; i++) V[i] += Vc[i];
        ef[0] += t[0]; ef[1] += t[1]; ef[2] += t[2];
        if (!hasE) Jm.featNo[jf] = Cm.featNo[cf];
        for (int j = Cm.wPtr[cf]; j < Cm.wPtr[cf + 1]; j++) {
            int sp = Cm.photo[j];
            int p = cur_pose_joint(c, m1, sp);
            if (p == c.posID1) continue;
            double W[18], y[6], xp[6], t2[3];
            sm::load<18>(Cm.W + 18 * (size_t)j, W);
            sm::mm<6, 3, 1>(W, xc, y);
            load_pose(Cm, sp, c.cSca, c.wrpC, xp);
            sm::mtm<3, 6, 1>(W, xp, t2);
            ef[0] += t2[0]; ef[1] += t2[1]; ef[2] += t2[2];
            for (int q = 0; q < 6; q++) atomicAdd(ePk + 6 * p + q, y[q]);
            if (p == c.posID2 && flSlot >= 0 && hasE) {
                double *w = Jm.W + 18 * (size_t)flSlot;            // own block of this thread: no race
                for (int q = 0; q < 18; q++) w[q] += W[q];
            } else {
                sm::store<18>(Jm.W + 18 * (size_t)o, W);
                Jm.photo[o] = p; Jm.feature[o] = jf;
                o++;
            }
        }
    }
    sm::store<9>(Jm.V + 9 * (size_t)jf, V);
    sm::store<3>(eF + 3 * (size_t)g, ef);
}

__global__ void k_mj_finish(DMap *__restrict__ J, const DMap *__restrict__ E, int K,
                            const MJ *__restrict__ mj, int *__restrict__ gRef, int *__restrict__ gFix,
                            int *__restrict__ gSign)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    J[k].wPtr[J[k].n] = J[k].nW;
    gRef[k] = mj[k].posID1;
    gFix[k] = 6 * mj[k].posID2 + E[k].Fix;           // posFix = pos2 + End.Fix (7798)
    gSign[k] = E[k].Sign;
}

void exclusive_scan(Context &ctx, const int *in, int *out, int n)
{
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, in, out, n, ctx.stream);
    DevBuf<char> tmp(tmp_bytes, ctx.stream);
    cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, in, out, n, ctx.stream);
}

} // namespace

std::vector<MapHandle> join_mono_batch(Context &ctx, const std::vector<MapHandle> &End,
                                       const std::vector<MapHandle> &Cur)
{
    const int K = (int)End.size();
    if (K == 0) return {};
    if ((int)Cur.size() != K) throw LsfmError(LSFM_ERR_ARG, "mono join: End/Cur count mismatch");
    cudaStream_t s = ctx.stream;
    const int TB = 128;
    int nl = 0;
    ctx.begin("mono.join");
    OpMaps E, C;
    E.build(End, s);
    C.build(Cur, s);

    // shared poses
    DevBuf<int> slots(4 * (size_t)K, s);
    std::vector<int> init(4 * (size_t)K, INT_MAX);
    slots.upload(init);
    k_mj_find<<<ceil_div(E.totPose + C.totPose, TB), TB, 0, s>>>(E.d.p, C.d.p, E.dPosePre.p, C.dPosePre.p, K,
                                                                E.totPose, C.totPose, slots.p); nl++;
    DevBuf<MJ> mj(K, s);
    k_mj_const<<<ceil_div(K, 64), 64, 0, s>>>(E.d.p, C.d.p, K, slots.p, mj.p); nl++;

    // common features (same hash join as stereo)
    u64 cap = 1024;
    while (cap < 2ull * (u64)C.totFeat) cap <<= 1;
    DevBuf<u64> hkeys(cap, s);
    DevBuf<int> hvals(cap, s);
    CUDA_CHECK(cudaMemsetAsync(hkeys.p, 0xff, cap * sizeof(u64), s));
    CUDA_CHECK(cudaMemsetAsync(hvals.p, 0x7f, cap * sizeof(int), s));
    DevBuf<int> owner(C.totFeat + 1, s), onlyFlag(C.totFeat + 1, s), onlyScan(C.totFeat + 1, s);
    CUDA_CHECK(cudaMemsetAsync(owner.p, 0xff, (C.totFeat + 1) * sizeof(int), s));
    if (C.totFeat > 0) { k_hash_insert<<<ceil_div(C.totFeat, TB), TB, 0, s>>>(C.d.p, C.dFeatPre.p, K, C.totFeat, hkeys.p, hvals.p, cap - 1); nl++; }
    if (E.totFeat > 0 && C.totFeat > 0) {
        k_hash_probe<<<ceil_div(E.totFeat, TB), TB, 0, s>>>(E.d.p, E.dFeatPre.p, C.dFeatPre.p, K, E.totFeat, hkeys.p,
                                                           hvals.p, cap - 1, owner.p); nl++;
    }
    k_only_flag<<<ceil_div(C.totFeat + 1, TB), TB, 0, s>>>(owner.p, C.totFeat, onlyFlag.p); nl++;
    exclusive_scan(ctx, onlyFlag.p, onlyScan.p, C.totFeat + 1); nl += 2;
    DevBuf<int> dOnly(K, s);
    k_only_count<<<ceil_div(K, TB), TB, 0, s>>>(C.dFeatPre.p, K, onlyScan.p, dOnly.p); nl++;
    std::vector<int> nOnly(K), hSlots(4 * (size_t)K);
    dOnly.download(nOnly.data(), K);
    slots.download(hSlots.data(), hSlots.size());
    CUDA_CHECK(cudaStreamSynchronize(s));
    for (size_t i = 0; i < hSlots.size(); i++)
        if (hSlots[i] == INT_MAX)
            throw LsfmError(LSFM_ERR_REF_NOT_FOUND, "mono join: shared Ref/ScaP pose missing from a map");

    // provisional joint descriptors (features known, block counts not yet) for the counting kernels
    std::vector<DMap> prov(K);
    std::vector<int> uPreS(K + 1, 0);
    for (int k = 0; k < K; k++) {
        memset(&prov[k], 0, sizeof(DMap));
        prov[k].m = E.h[k].m + C.h[k].m - 2;
        prov[k].n = E.h[k].n + nOnly[k];
        uPreS[k + 1] = uPreS[k] + E.h[k].nU + C.h[k].nU;
    }
    OpMaps P;
    P.build(prov, s);
    const int totS = uPreS[K];
    DevBuf<int> dUPreS(K + 1, s); dUPreS.upload(uPreS);
    DevBuf<int> jointOfCur(C.totFeat + 1, s), curOfJoint(P.totFeat + 1, s);
    CUDA_CHECK(cudaMemsetAsync(curOfJoint.p, 0xff, (P.totFeat + 1) * sizeof(int), s));
    if (C.totFeat > 0) {
        k_joint_index<<<ceil_div(C.totFeat, TB), TB, 0, s>>>(E.d.p, C.dFeatPre.p, P.dFeatPre.p, K, C.totFeat, owner.p,
                                                            onlyScan.p, jointOfCur.p, curOfJoint.p); nl++;
    }
    DevBuf<int> uFlag(totS + 1, s), uScan(totS + 1, s), wCnt(P.totFeat + 1, s), wScan(P.totFeat + 1, s);
    k_mj_uflag<<<ceil_div(totS + 1, TB), TB, 0, s>>>(E.d.p, C.d.p, dUPreS.p, K, totS, mj.p, uFlag.p); nl++;
    k_mj_uflag_cur<<<ceil_div(totS, TB), TB, 0, s>>>(E.d.p, C.d.p, dUPreS.p, K, totS, mj.p, uFlag.p); nl++;
    k_mj_wcount<<<ceil_div(P.totFeat + 1, TB), TB, 0, s>>>(E.d.p, C.d.p, P.dFeatPre.p, K, P.totFeat, mj.p,
                                                          curOfJoint.p, wCnt.p); nl++;
    exclusive_scan(ctx, uFlag.p, uScan.p, totS + 1); nl += 2;
    exclusive_scan(ctx, wCnt.p, wScan.p, P.totFeat + 1); nl += 2;
    DevBuf<int> dSizes(2 * (size_t)K, s);
    k_mj_sizes<<<ceil_div(K, TB), TB, 0, s>>>(dUPreS.p, P.dFeatPre.p, K, uScan.p, wScan.p, dSizes.p, dSizes.p + K); nl++;
    std::vector<int> hSizes(2 * (size_t)K);
    dSizes.download(hSizes.data(), hSizes.size());
    CUDA_CHECK(cudaStreamSynchronize(s));

    std::vector<DMap> shapes(K);
    for (int k = 0; k < K; k++) {
        DMap &j = shapes[k];
        memset(&j, 0, sizeof(j));
        j.m = prov[k].m; j.n = prov[k].n;
        j.nU = hSizes[k]; j.nW = hSizes[K + k];
        j.Ref = C.h[k].Ref; j.ScaP = C.h[k].ScaP; j.Fix = C.h[k].Fix; j.Sign = C.h[k].Sign;       // 7365-7369
        j.FRef = E.h[k].FRef; j.FScaP = E.h[k].FScaP; j.FFix = E.h[k].FFix;                        // 7371-7373
    }
    std::vector<MapHandle> out = alloc_maps(ctx, shapes);
    OpMaps J;
    J.build(out, s);
    for (int k = 0; k < K; k++)
        if (out[k].d.nU) CUDA_CHECK(cudaMemsetAsync(out[k].d.U, 0, sizeof(double) * 36 * (size_t)out[k].d.nU, s));
    DevBuf<double> eP(6 * (size_t)J.totPose, s), eF(3 * (size_t)J.totFeat, s);
    eP.zero();
    k_mj_pose_end<<<ceil_div(E.totPose, TB), TB, 0, s>>>(E.d.p, J.d.p, E.dPosePre.p, K, E.totPose); nl++;
    k_mj_pose<<<ceil_div(C.totPose, TB), TB, 0, s>>>(E.d.p, C.d.p, J.d.p, C.dPosePre.p, K, C.totPose, mj.p); nl++;
    if (totS > 0) {
        k_mj_u<<<ceil_div(totS, 64), 64, 0, s>>>(E.d.p, C.d.p, J.d.p, dUPreS.p, J.dPosePre.p, K, totS, mj.p,
                                                uFlag.p, uScan.p, eP.p); nl++;
    }
    if (J.totFeat > 0) {
        k_mj_feat<<<ceil_div(J.totFeat, 64), 64, 0, s>>>(E.d.p, C.d.p, J.d.p, J.dFeatPre.p, J.dPosePre.p, K,
                                                        J.totFeat, mj.p, curOfJoint.p, wScan.p, eP.p, eF.p); nl++;
    }
    DevBuf<int> gRef(K, s), gFix(K, s), gSign(K, s);
    k_mj_finish<<<ceil_div(K, TB), TB, 0, s>>>(J.d.p, E.d.p, K, mj.p, gRef.p, gFix.p, gSign.p); nl++;
    KERNEL_CHECK();
    ctx.end(0.0, 0.0, nl);

    MonoGauge gauge{gRef.p, gFix.p, gSign.p};
    solve_stereo_batch(ctx, J, eP.p, eF.p, nullptr, &gauge);
    return out;
}
