// Batched linear join of map pairs: rows a6-a7 of SURVEY 8(a) (+ the optional per-join objective).
//
// Reference: CLinearSFMImp::lmj_LinearLS_PF3DStereo, LinearSFMImp.cpp:2551-2978
//   common-feature search (std::find, O(n1 n2))     2565-2599   -> device hash join, O(n1+n2)
//   joint layout, U/W/V merge, RHS eP/eF            2606-2930
//   -> lmj_solveLinearSFMStereo                     2966        (solve.cu)
//
// Output ordering is the reference's: poses = End then Cur (+m1); features =
// End's (common ones get V summed) then Cur-only in Cur order; U list = End's then Cur's; per joint
// feature the End blocks then the Cur blocks.
#include "ops.h"
#include "small_mat.cuh"
#include "det_accum.cuh"
#include <cub/cub.cuh>

namespace {

typedef unsigned long long u64;
#define HASH_EMPTY 0xffffffffffffffffull

__device__ __forceinline__ u64 mix64(u64 x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

__global__ void k_hash_insert(const DMap *__restrict__ C, const int *__restrict__ featPreC, int K,
                              int totC, u64 *__restrict__ keys, int *__restrict__ vals, u64 mask)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= totC) return;
    int k = seg_find(featPreC, K, g);
    int c = g - featPreC[k];
    u64 key = ((u64)(unsigned)k << 32) | (unsigned)C[k].featNo[c];
    u64 h = mix64(key) & mask;
    while (true) {
        u64 prev = atomicCAS(&keys[h], HASH_EMPTY, key);
        if (prev == HASH_EMPTY || prev == key) { atomicMin(&vals[h], c); return; }   // first index wins
        h = (h + 1) & mask;
    }
}

// for each End feature: first Cur feature with the same id (LinearSFMImp.cpp:2581-2599)
__global__ void k_hash_probe(const DMap *__restrict__ E, const int *__restrict__ featPreE,
                             const int *__restrict__ featPreC, int K, int totE,
                             const u64 *__restrict__ keys, const int *__restrict__ vals, u64 mask,
                             int *__restrict__ ownerOfCur)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= totE) return;
    int k = seg_find(featPreE, K, g);
    int i = g - featPreE[k];
    u64 key = ((u64)(unsigned)k << 32) | (unsigned)E[k].featNo[i];
    u64 h = mix64(key) & mask;
    while (true) {
        u64 cur = keys[h];
        if (cur == HASH_EMPTY) return;
        if (cur == key) { atomicMax(&ownerOfCur[featPreC[k] + vals[h]], i); return; }  // last i wins (2596)
        h = (h + 1) & mask;
    }
}

__global__ void k_only_flag(const int *__restrict__ ownerOfCur, int totC, int *__restrict__ flag)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g > totC) return;
    flag[g] = (g < totC && ownerOfCur[g] < 0) ? 1 : 0;
}

__global__ void k_only_count(const int *__restrict__ featPreC, int K, const int *__restrict__ scan,
                             int *__restrict__ nOnly)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    nOnly[k] = scan[featPreC[k + 1]] - scan[featPreC[k]];
}

// joint index of every Cur feature (Curfeature2, 2596/2639) and the inverse map joint -> Cur
__global__ void k_joint_index(const DMap *__restrict__ E, const int *__restrict__ featPreC,
                              const int *__restrict__ featPreJ, int K, int totC,
                              const int *__restrict__ ownerOfCur, const int *__restrict__ scan,
                              int *__restrict__ jointOfCur, int *__restrict__ curOfJoint)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= totC) return;
    int k = seg_find(featPreC, K, g);
    int own = ownerOfCur[g];
    int jf = own >= 0 ? own : E[k].n + (scan[g] - scan[featPreC[k]]);
    jointOfCur[g] = jf;
    curOfJoint[featPreJ[k] + jf] = g - featPreC[k];
}

__global__ void k_joint_count(const DMap *__restrict__ E, const DMap *__restrict__ C,
                              const int *__restrict__ featPreJ, int K, int totJ,
                              const int *__restrict__ curOfJoint, int *__restrict__ cnt)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g > totJ) return;
    if (g == totJ) { cnt[g] = 0; return; }
    int k = seg_find(featPreJ, K, g);
    int jf = g - featPreJ[k];
    int c = curOfJoint[g];
    int n = 0;
    if (jf < E[k].n) n += E[k].wPtr[jf + 1] - E[k].wPtr[jf];
    if (c >= 0) n += C[k].wPtr[c + 1] - C[k].wPtr[c];
    cnt[g] = n;
}

__global__ void k_join_pose(const DMap *__restrict__ E, const DMap *__restrict__ C,
                            DMap *__restrict__ J, const int *__restrict__ posePreJ, int K, int totP)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= totP) return;
    int k = seg_find(posePreJ, K, g);
    int p = g - posePreJ[k];
    int m1 = E[k].m;
    J[k].poseNo[p] = p < m1 ? E[k].poseNo[p] : C[k].poseNo[p - m1];
}

// U list = End's then Cur's (+m1); eP += U xhat (and the transposed product for off-diagonal blocks):
// two 6-vector records per block (targets: joint pose rows i and j), added up in a fixed order by
// det::reduce -- no FP64 atomics
__global__ void k_join_u(const DMap *__restrict__ E, const DMap *__restrict__ C, DMap *__restrict__ J,
                         const int *__restrict__ uPreJ, const int *__restrict__ posePreJ, int K, int totU,
                         int *__restrict__ rkey, double *__restrict__ rval, int none)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= totU) return;
    int k = seg_find(uPreJ, K, g);
    int b = g - uPreJ[k];
    const DMap &S = (b < E[k].nU) ? E[k] : C[k];
    int off = (b < E[k].nU) ? 0 : E[k].m;
    int sb = (b < E[k].nU) ? b : b - E[k].nU;
    double U[36];
    sm::load<36>(S.U + 36 * (size_t)sb, U);
    sm::store<36>(J[k].U + 36 * (size_t)b, U);
    int i = S.Ui[sb], j = S.Uj[sb];
    J[k].Ui[b] = i + off;
    J[k].Uj[b] = j + off;
    double xi[6], xj[6], y[6];
    sm::load<6>(S.poseVal + 6 * (size_t)j, xj);
    sm::mm<6, 6, 1>(U, xj, y);
    rkey[2 * (size_t)g] = posePreJ[k] + i + off;
    sm::store<6>(rval + 12 * (size_t)g, y);
    if (i != j) {
        sm::load<6>(S.poseVal + 6 * (size_t)i, xi);
        sm::mtm<6, 6, 1>(U, xi, y);
        rkey[2 * (size_t)g + 1] = posePreJ[k] + j + off;
        sm::store<6>(rval + 12 * (size_t)g + 6, y);
    } else {
        rkey[2 * (size_t)g + 1] = none;
    }
}

struct ApplyEP {
    double *eP;
    __device__ void operator()(int t, int q, double sum, int) const { eP[6 * (size_t)t + q] = sum; }
};

// one thread per joint feature: V merge, labels, CSR start, the V part of eF, and the feature's two
// estimates (End side, Cur side) for the solver's d vectors (2747-2930)
__global__ void __launch_bounds__(256)
k_join_featinit(const DMap *__restrict__ E, const DMap *__restrict__ C, DMap *__restrict__ J,
                const int *__restrict__ featPreJ, int K, int totJ, const int *__restrict__ curOfJoint,
                const int *__restrict__ wScan, double *__restrict__ eF, double *__restrict__ xhat)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= totJ) return;
    int k = seg_find(featPreJ, K, g);
    int jf = g - featPreJ[k];
    const DMap &Em = E[k];
    const DMap &Cm = C[k];
    const DMap &Jm = J[k];
    int c = curOfJoint[g];
    bool hasE = jf < Em.n;
    double ef[3] = {0.0, 0.0, 0.0};
    double V[9];
#pragma unroll
    for (int i = 0; i < 9; i++) V[i] = 0.0;
    double xe[3] = {0, 0, 0}, xc[3] = {0, 0, 0};
    if (hasE) {
        sm::load<9>(Em.V + 9 * (size_t)jf, V);
        sm::load<3>(Em.featVal + 3 * (size_t)jf, xe);
        sm::mm<3, 3, 1>(V, xe, ef);
        Jm.featNo[jf] = Em.featNo[jf];
    }
    if (c >= 0) {
        double Vc[9], t[3];
        sm::load<9>(Cm.V + 9 * (size_t)c, Vc);
        sm::load<3>(Cm.featVal + 3 * (size_t)c, xc);
        sm::mm<3, 3, 1>(Vc, xc, t);
#pragma unroll
        for (int i = 0; i < 9; i++) V[i] += Vc[i];
        ef[0] += t[0]; ef[1] += t[1]; ef[2] += t[2];
        if (!hasE) Jm.featNo[jf] = Cm.featNo[c];
    }
    sm::store<9>(Jm.V + 9 * (size_t)jf, V);
    Jm.wPtr[jf] = wScan[g] - wScan[featPreJ[k]];
    sm::store<3>(eF + 3 * (size_t)g, ef);
    double *x = xhat + 6 * (size_t)g;
    x[0] = xe[0]; x[1] = xe[1]; x[2] = xe[2]; x[3] = xc[0]; x[4] = xc[1]; x[5] = xc[2];
}

// one thread per SOURCE W block (End's blocks, then Cur's, of every join): its slot in the joint map
// (per joint feature: End blocks, then Cur blocks), the integer labels, and the slot itself for the
// value copy that runs later (k_join_w).  Only integers: this is all the S pattern needs.
__global__ void __launch_bounds__(256)
k_join_ints(const DMap *__restrict__ E, const DMap *__restrict__ C, DMap *__restrict__ J,
            const int *__restrict__ wPreJ, const int *__restrict__ featPreJ,
            const int *__restrict__ featPreC, int K, int totW, const int *__restrict__ jointOfCur,
            const int *__restrict__ wScan, int *__restrict__ dstOf)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= totW) return;
    int k = seg_find(wPreJ, K, g);
    int b = g - wPreJ[k];
    const DMap &Em = E[k];
    bool fromE = b < Em.nW;
    const DMap &S = fromE ? Em : C[k];
    int sb = fromE ? b : b - Em.nW;
    int f = S.feature[sb];
    int jf = fromE ? f : jointOfCur[featPreC[k] + f];
    int before = (!fromE && jf < Em.n) ? Em.wPtr[jf + 1] - Em.wPtr[jf] : 0;
    int gj = featPreJ[k] + jf;
    int o = (wScan[gj] - wScan[featPreJ[k]]) + before + (sb - S.wPtr[f]);
    J[k].photo[o] = S.photo[sb] + (fromE ? 0 : Em.m);
    J[k].feature[o] = jf;
    dstOf[g] = o;
}

// W value copy + the W part of eF (eF_f += W_pf^T xhat_p, 2830-2930), one thread per source block.
// A warp whose 32 blocks come from one source map reads them as ONE contiguous 4.6 KB span
// (coalesced), scatters whole blocks to their joint slots (runs of consecutive slots) and passes
// each lane its own block through shared memory for the 6x3 product; the per-feature sums are
// reduced across the lanes of a feature before the atomics.
__global__ void __launch_bounds__(128)
k_join_w(const DMap *__restrict__ E, const DMap *__restrict__ C, DMap *__restrict__ J,
         const int *__restrict__ wPreJ, const int *__restrict__ featPreJ,
         const int *__restrict__ featPreC, int K, int totW, const int *__restrict__ jointOfCur,
         const int *__restrict__ dstOf, double *__restrict__ featSide, double *__restrict__ warpCont)
{
    __shared__ double tile[4][32 * 19];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    bool live = g < totW;
    int gg = live ? g : totW - 1;
    int k = seg_find(wPreJ, K, gg);
    int b = gg - wPreJ[k];
    const DMap &Em = E[k];
    bool fromE = b < Em.nW;
    const DMap &S = fromE ? Em : C[k];
    int sb = fromE ? b : b - Em.nW;
    int f = S.feature[sb];
    int p = S.photo[sb];
    int jf = fromE ? f : jointOfCur[featPreC[k] + f];
    int gj = featPreJ[k] + jf;
    int o = dstOf[gg];
    int k0 = __shfl_sync(full, k, 0);
    int e0 = __shfl_sync(full, (int)fromE, 0);
    bool uni = __all_sync(full, live && k == k0 && (int)fromE == e0);
    double W[18];
    double *dst = J[k].W;
    if (uni) {
        // 16-byte pieces: a block is nine of them (144 bytes, 16-byte aligned in both maps)
        const double2 *src = reinterpret_cast<const double2 *>(S.W + 18 * (size_t)(sb - lane));   // the warp's first block
        double2 *dst2 = reinterpret_cast<double2 *>(dst);
        double *tw = tile[warp];
#pragma unroll
        for (int it = 0; it < 9; it++) {
            int e = it * 32 + lane;
            int blk = e / 9, el = e - 9 * blk;
            double2 v = src[e];
            int ob = __shfl_sync(full, o, blk);
            dst2[9 * (size_t)ob + el] = v;
            tw[blk * 19 + 2 * el] = v.x;
            tw[blk * 19 + 2 * el + 1] = v.y;
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 18; q++) W[q] = tw[lane * 19 + q];
    } else if (live) {
        sm::load<18>(S.W + 18 * (size_t)sb, W);
        sm::store<18>(dst + 18 * (size_t)o, W);
    }
    double t[3] = {0.0, 0.0, 0.0};
    int key = -1 - lane;                       // dead lanes get unique keys
    if (live) {
        double xp[6];
        sm::load<6>(S.poseVal + 6 * (size_t)p, xp);
        sm::mtm<3, 6, 1>(W, xp, t);
        key = 2 * gj + (fromE ? 0 : 1);    // End and Cur runs of one feature stay separate segments: a key
                                            // must never reappear after a different one (a,b,a would be double counted)
    }
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        int okey = __shfl_down_sync(full, key, off);
        bool take = (lane + off < 32) && (okey == key);
#pragma unroll
        for (int i = 0; i < 3; i++) {
            double ov = __shfl_down_sync(full, t[i], off);
            if (take) t[i] += ov;
        }
    }
    // Run heads store their run's sum (no atomics): a run that starts at the first block of its
    // feature in the source map -> featSide[2 gj + side]; a run cut by the warp boundary (lane 0, block
    // in the middle of its feature) -> warpCont[warp].  k_join_ef_fin adds them up per feature in order.
    int pkey = __shfl_up_sync(full, key, 1);
    if (live && (lane == 0 || pkey != key)) {
        double *e = (sb == S.wPtr[f]) ? featSide + 3 * (size_t)(2 * gj + (fromE ? 0 : 1))
                                      : warpCont + 3 * (size_t)(g >> 5);
        e[0] = t[0]; e[1] = t[1]; e[2] = t[2];
    }
}

// eF_f = V part (k_join_featinit) + sum of the End-side runs + sum of the Cur-side runs, fixed order
__global__ void k_join_ef_fin(const DMap *__restrict__ E, const DMap *__restrict__ C,
                              const int *__restrict__ wPreJ, const int *__restrict__ featPreJ, int K, int totJ,
                              const int *__restrict__ curOfJoint, const double *__restrict__ featSide,
                              const double *__restrict__ warpCont, double *__restrict__ eF)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= totJ) return;
    int k = seg_find(featPreJ, K, g);
    int jf = g - featPreJ[k];
    double s0 = eF[3 * (size_t)g], s1 = eF[3 * (size_t)g + 1], s2 = eF[3 * (size_t)g + 2];
    auto side = [&](int ga, int gb, int sd) {
        if (gb <= ga) return;
        const double *h = featSide + 3 * (size_t)(2 * g + sd);
        double a0 = h[0], a1 = h[1], a2 = h[2];
        for (int w = (ga >> 5) + 1; w <= ((gb - 1) >> 5); w++) {
            const double *c = warpCont + 3 * (size_t)w;
            a0 += c[0]; a1 += c[1]; a2 += c[2];
        }
        s0 += a0; s1 += a1; s2 += a2;
    };
    if (jf < E[k].n) side(wPreJ[k] + E[k].wPtr[jf], wPreJ[k] + E[k].wPtr[jf + 1], 0);
    int c = curOfJoint[g];
    if (c >= 0) side(wPreJ[k] + E[k].nW + C[k].wPtr[c], wPreJ[k] + E[k].nW + C[k].wPtr[c + 1], 1);
    eF[3 * (size_t)g] = s0; eF[3 * (size_t)g + 1] = s1; eF[3 * (size_t)g + 2] = s2;
}

__global__ void k_wptr_end(DMap *__restrict__ J, int K)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    J[k].wPtr[J[k].n] = J[k].nW;
}

// F_k = sum over End and Cur of (x|_k - xhat_k)^T I_k (x|_k - xhat_k)   (residual form)
// thread per source block; one atomic per block into obj[k].
__global__ void k_objective(const DMap *__restrict__ E, const DMap *__restrict__ C,
                            const DMap *__restrict__ J, int K, const int *__restrict__ uPreJ,
                            const int *__restrict__ wPreJ, const int *__restrict__ featPreE,
                            const int *__restrict__ featPreC, int totU, int totW, int totFE, int totFC,
                            const int *__restrict__ jointOfCur, double *__restrict__ obj)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    int which = 0;                      // 0:U 1:W 2:V(End) 3:V(Cur)
    int idx = g;
    if (idx >= totU) { idx -= totU; which = 1;
        if (idx >= totW) { idx -= totW; which = 2;
            if (idx >= totFE) { idx -= totFE; which = 3; if (idx >= totFC) return; } } }
    double val = 0.0;
    int k;
    if (which == 0) {
        k = seg_find(uPreJ, K, idx);
        int b = idx - uPreJ[k];
        bool fromE = b < E[k].nU;
        const DMap &S = fromE ? E[k] : C[k];
        int sb = fromE ? b : b - E[k].nU, off = fromE ? 0 : E[k].m;
        int i = S.Ui[sb], j = S.Uj[sb];
        double di[6], dj[6], U[36], y[6];
        for (int q = 0; q < 6; q++) {
            di[q] = J[k].poseVal[6 * (size_t)(i + off) + q] - S.poseVal[6 * (size_t)i + q];
            dj[q] = J[k].poseVal[6 * (size_t)(j + off) + q] - S.poseVal[6 * (size_t)j + q];
        }
        sm::load<36>(S.U + 36 * (size_t)sb, U);
        sm::mm<6, 6, 1>(U, dj, y);
        for (int q = 0; q < 6; q++) val += di[q] * y[q];
        if (i != j) val *= 2.0;
    } else if (which == 1) {
        // W blocks are enumerated in SOURCE order: End's blocks then Cur's blocks of join k
        k = seg_find(wPreJ, K, idx);
        int b = idx - wPreJ[k];
        bool fromE = b < E[k].nW;
        const DMap &S = fromE ? E[k] : C[k];
        int sb = fromE ? b : b - E[k].nW, off = fromE ? 0 : E[k].m;
        int p = S.photo[sb], f = S.feature[sb];
        int jf = fromE ? f : jointOfCur[featPreC[k] + f];
        double dp[6], df[3], W[18], y[6];
        for (int q = 0; q < 6; q++)
            dp[q] = J[k].poseVal[6 * (size_t)(p + off) + q] - S.poseVal[6 * (size_t)p + q];
        for (int q = 0; q < 3; q++)
            df[q] = J[k].featVal[3 * (size_t)jf + q] - S.featVal[3 * (size_t)f + q];
        sm::load<18>(S.W + 18 * (size_t)sb, W);
        sm::mm<6, 3, 1>(W, df, y);
        for (int q = 0; q < 6; q++) val += dp[q] * y[q];
        val *= 2.0;
    } else {
        bool fromE = which == 2;
        const int *pre = fromE ? featPreE : featPreC;
        k = seg_find(pre, K, idx);
        int f = idx - pre[k];
        const DMap &S = fromE ? E[k] : C[k];
        int jf = fromE ? f : jointOfCur[featPreC[k] + f];
        double df[3], V[9], y[3];
        for (int q = 0; q < 3; q++)
            df[q] = J[k].featVal[3 * (size_t)jf + q] - S.featVal[3 * (size_t)f + q];
        sm::load<9>(S.V + 9 * (size_t)f, V);
        sm::mm<3, 3, 1>(V, df, y);
        val = df[0] * y[0] + df[1] * y[1] + df[2] * y[2];
    }
    atomicAdd(&obj[k], val);
}

void exclusive_scan(Context &ctx, const int *in, int *out, int n)
{
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, in, out, n, ctx.stream);
    DevBuf<char> tmp(tmp_bytes, ctx.stream);
    cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, in, out, n, ctx.stream);
}

} // namespace

std::vector<MapHandle> join_stereo_batch(Context &ctx, const std::vector<MapHandle> &End,
                                         const std::vector<MapHandle> &Cur)
{
    const int K = (int)End.size();
    if (K == 0) return {};
    if ((int)Cur.size() != K) throw LsfmError(LSFM_ERR_ARG, "join: End/Cur count mismatch");
    cudaStream_t s = ctx.stream;
    const int TB = 256;
    int nl = 0;
    ctx.begin("join.match");
    OpMaps E, C;
    E.build(End, s);
    C.build(Cur, s);

    // ---- a6: common features (hash join) ----
    u64 cap = 1024;
    while (cap < 2ull * (u64)C.totFeat) cap <<= 1;
    DevBuf<u64> hkeys(cap, s);
    DevBuf<int> hvals(cap, s);
    CUDA_CHECK(cudaMemsetAsync(hkeys.p, 0xff, cap * sizeof(u64), s));
    CUDA_CHECK(cudaMemsetAsync(hvals.p, 0x7f, cap * sizeof(int), s));
    DevBuf<int> owner(C.totFeat + 1, s), onlyFlag(C.totFeat + 1, s), onlyScan(C.totFeat + 1, s);
    CUDA_CHECK(cudaMemsetAsync(owner.p, 0xff, (C.totFeat + 1) * sizeof(int), s));
    if (C.totFeat > 0) {
        k_hash_insert<<<ceil_div(C.totFeat, TB), TB, 0, s>>>(C.d.p, C.dFeatPre.p, K, C.totFeat, hkeys.p, hvals.p, cap - 1); nl++;
    }
    if (E.totFeat > 0 && C.totFeat > 0) {
        k_hash_probe<<<ceil_div(E.totFeat, TB), TB, 0, s>>>(E.d.p, E.dFeatPre.p, C.dFeatPre.p, K, E.totFeat,
                                                           hkeys.p, hvals.p, cap - 1, owner.p); nl++;
    }
    k_only_flag<<<ceil_div(C.totFeat + 1, TB), TB, 0, s>>>(owner.p, C.totFeat, onlyFlag.p); nl++;
    exclusive_scan(ctx, onlyFlag.p, onlyScan.p, C.totFeat + 1); nl += 2;
    DevBuf<int> dOnly(K, s);
    k_only_count<<<ceil_div(K, TB), TB, 0, s>>>(C.dFeatPre.p, K, onlyScan.p, dOnly.p); nl++;
    std::vector<int> nOnly(K);
    dOnly.download(nOnly.data(), K);
    CUDA_CHECK(cudaStreamSynchronize(s));
    ctx.idle_begin();
    KERNEL_CHECK();
    ctx.end(8.0 * (E.totFeat + C.totFeat), 0.0, nl);
    nl = 0;

    // ---- joint shapes (2606-2611) ----
    ctx.begin("join.merge");
    std::vector<DMap> shapes(K);
    double bytes = 0.0;
    for (int k = 0; k < K; k++) {
        DMap &j = shapes[k];
        memset(&j, 0, sizeof(j));
        j.m = E.h[k].m + C.h[k].m;
        j.n = E.h[k].n + nOnly[k];
        j.nU = E.h[k].nU + C.h[k].nU;
        j.nW = E.h[k].nW + C.h[k].nW;
        j.Ref = C.h[k].Ref;             // 2974
        j.FRef = E.h[k].FRef;           // 2624
        bytes += map_bytes(E.h[k]) + map_bytes(C.h[k]);
    }
    std::vector<MapHandle> out = alloc_maps(ctx, shapes);
    OpMaps J;
    J.build(out, s);
    for (int k = 0; k < K; k++) bytes += map_bytes(J.h[k]);

    DevBuf<int> jointOfCur(C.totFeat + 1, s), curOfJoint(J.totFeat + 1, s);
    ctx.idle_end(1);
    CUDA_CHECK(cudaMemsetAsync(curOfJoint.p, 0xff, (J.totFeat + 1) * sizeof(int), s));
    if (C.totFeat > 0) {
        k_joint_index<<<ceil_div(C.totFeat, TB), TB, 0, s>>>(E.d.p, C.dFeatPre.p, J.dFeatPre.p, K, C.totFeat,
                                                            owner.p, onlyScan.p, jointOfCur.p, curOfJoint.p); nl++;
    }
    DevBuf<int> wCnt(J.totFeat + 1, s), wScan(J.totFeat + 1, s);
    k_joint_count<<<ceil_div(J.totFeat + 1, TB), TB, 0, s>>>(E.d.p, C.d.p, J.dFeatPre.p, K, J.totFeat,
                                                            curOfJoint.p, wCnt.p); nl++;
    exclusive_scan(ctx, wCnt.p, wScan.p, J.totFeat + 1); nl += 2;

    // eP: only the U part here; the W xhat_f part is folded into the reduced right-hand side by the
    // solver through the per-feature d vectors (SolveExtra, k_vinv)
    DevBuf<double> eP(6 * (size_t)J.totPose, s), eF(3 * (size_t)J.totFeat, s), xhat(6 * (size_t)J.totFeat, s);
    DevBuf<int> dstOf(std::max(J.totW, 1), s);
    k_join_pose<<<ceil_div(J.totPose, TB), TB, 0, s>>>(E.d.p, C.d.p, J.d.p, J.dPosePre.p, K, J.totPose); nl++;
    {
        DevBuf<int> rkey(2 * (size_t)std::max(J.totU, 1), s);
        DevBuf<double> rval(12 * (size_t)std::max(J.totU, 1), s);
        if (J.totU > 0) {
            k_join_u<<<ceil_div(J.totU, 128), 128, 0, s>>>(E.d.p, C.d.p, J.d.p, J.dUPre.p, J.dPosePre.p, K, J.totU,
                                                          rkey.p, rval.p, J.totPose); nl++;
        }
        det::Sorted srt;
        nl += det::sort_records(ctx, rkey.p, 2 * J.totU, J.totPose, srt);
        nl += det::reduce<6>(ctx, srt, rval.p, J.totPose, ApplyEP{eP.p});
    }
    if (J.totFeat > 0) {
        k_join_featinit<<<ceil_div(J.totFeat, TB), TB, 0, s>>>(E.d.p, C.d.p, J.d.p, J.dFeatPre.p, K, J.totFeat,
                                                              curOfJoint.p, wScan.p, eF.p, xhat.p); nl++;
    }
    if (J.totW > 0) {
        k_join_ints<<<ceil_div(J.totW, TB), TB, 0, s>>>(E.d.p, C.d.p, J.d.p, J.dWPre.p, J.dFeatPre.p, C.dFeatPre.p, K,
                                                       J.totW, jointOfCur.p, wScan.p, dstOf.p); nl++;
    }
    k_wptr_end<<<ceil_div(K, TB), TB, 0, s>>>(J.d.p, K); nl++;
    std::vector<int> hSplit(K);
    for (int k = 0; k < K; k++) hSplit[k] = E.h[k].m;
    DevBuf<int> dSplit(K, s);
    dSplit.upload(hSplit);
    KERNEL_CHECK();
    const double valueBytes = 288.0 * J.totW + 24.0 * J.totFeat;
    ctx.end(bytes - valueBytes, 0.0, nl);

    // ---- a8-a13; the W value copy is queued by the solver once the pattern is with the host ----
    SolveExtra ex;
    ex.xhat = xhat.p;
    ex.split = dSplit.p;
    DevBuf<double> featSide(6 * (size_t)std::max(J.totFeat, 1), s), warpCont(3 * (size_t)(J.totW / 32 + 1), s);
    ex.after_pattern = [&]() {
        ctx.begin("join.values");
        if (J.totW > 0) {
            k_join_w<<<ceil_div(J.totW, 128), 128, 0, s>>>(E.d.p, C.d.p, J.d.p, J.dWPre.p, J.dFeatPre.p, C.dFeatPre.p, K,
                                                          J.totW, jointOfCur.p, dstOf.p, featSide.p, warpCont.p);
            k_join_ef_fin<<<ceil_div(J.totFeat, TB), TB, 0, s>>>(E.d.p, C.d.p, J.dWPre.p, J.dFeatPre.p, K, J.totFeat,
                                                                curOfJoint.p, featSide.p, warpCont.p, eF.p);
        }
        KERNEL_CHECK();
        ctx.end(valueBytes, 0.0, 2);
    };
    // feature chunks: End's features keep their indices in the joint map (Cur's new ones follow), so End's
    // chunk boundaries -- narrowed where chunks saw too many poses -- seed the joint map's
    std::vector<std::shared_ptr<const std::vector<int>>> seed(K), chunksOut;
    for (int k = 0; k < K; k++) {
        if (!End[k].chunkStarts) continue;
        const std::vector<int> &es = *End[k].chunkStarts;
        const int nE = E.h[k].n, nJ = J.h[k].n;
        if (es.empty() || es.front() != 0 || es.back() != nE) continue;
        std::vector<int> st(es.begin(), es.end() - 1);
        for (int f0 = nE; f0 < nJ; f0 += 128) st.push_back(f0);
        st.push_back(nJ);
        seed[k] = std::make_shared<const std::vector<int>>(std::move(st));
    }
    ex.chunkSeed = &seed;
    ex.chunksOut = &chunksOut;
    solve_stereo_batch(ctx, J, eP.p, eF.p, nullptr, nullptr, &ex);
    for (int k = 0; k < K && k < (int)chunksOut.size(); k++) out[k].chunkStarts = chunksOut[k];

    if (ctx.want_objective) {
        ctx.begin("objective");
        DevBuf<double> obj(K, s);
        obj.zero();
        int tot = J.totU + J.totW + E.totFeat + C.totFeat;
        k_objective<<<ceil_div(tot, TB), TB, 0, s>>>(E.d.p, C.d.p, J.d.p, K, J.dUPre.p, J.dWPre.p, E.dFeatPre.p,
                                                    C.dFeatPre.p, J.totU, J.totW, E.totFeat, C.totFeat,
                                                    jointOfCur.p, obj.p);
        std::vector<double> h = obj.to_host();
        ctx.objectives.insert(ctx.objectives.end(), h.begin(), h.end());
        ctx.end(0.0, 0.0, 1);
    }
    return out;
}
