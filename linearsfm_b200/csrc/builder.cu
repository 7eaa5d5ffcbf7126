// Stereo local-map builder: the step upstream of LinearSFM (north_star bullets 1-2).
//
// There is no reference code for it (LinearSFM reads finished localmap_*.txt files); conventions are
// those of builder_math.h.  One local map = two consecutive stereo frames: the map frame is the
// first frame, the state is the pose of the second frame plus the n landmarks seen in both.
//
// k_build_stereo: ONE CTA PER MAP, every map of the run in one launch.  Levenberg-Marquardt on
//     min sum_f |z0_f - h(X_f)|^2 + |z1_f - h(R(a)(X_f - t))|^2   (weights 1/sigma^2):
//   per iteration, one thread per landmark (SoA measurement arrays, coalesced): residuals +
//   Jacobians of its six measurements, its blocks V (3x3), W (6x3), Ub (6x6), gF, gP; the landmark is
//   eliminated on the spot (Schur: S_f = Ub - W V^-1 W^T, e_f = gP - W V^-1 gF) and the 27 numbers
//   of (S_f upper triangle, e_f) are summed over the map by a warp-shuffle + shared-memory
//   reduction (fixed order: deterministic); thread 0 solves the 6x6 system, every thread
//   back-substitutes its landmarks.  After convergence the blocks are evaluated once more at the
//   estimate: that is the (U, W, V) LinearSFM receives.
#include "builder.h"
#include "builder_math.h"
#include <cstring>
#include <vector>

namespace {

constexpr int BT = 128;

struct BuildArgs {
    const int *featPre;        // [K+1]
    const double *z0, *z1;     // [3 tot]
    double *pose;              // [6 K]  in: initial guess, out: estimate
    double *X;                 // [3 tot] in: initial guess (if haveX), out: estimate
    double *W, *V, *U;         // [18 tot], [9 tot], [36 K]  out (W, V double as per-iteration scratch)
    double *gF;                // [3 tot] scratch
    double *Xb;                // [3 tot] scratch: the last accepted landmark estimates
    int *iters, *status;       // [K]
    int haveX;
};

// sum v[0..N) over the CTA; result valid in red[0..N) after the call (all threads must call)
template <int N>
__device__ __forceinline__ void cta_sum(double *v, double (*red)[N], int tid)
{
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int i = 0; i < N; i++) {
        double s = v[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
        if (lane == 0) red[warp][i] = s;
    }
    __syncthreads();
    if (tid < N) {
        double s = red[0][tid];
#pragma unroll
        for (int w = 1; w < BT / 32; w++) s += red[w][tid];
        red[0][tid] = s;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(BT)
k_build_stereo(BuildArgs A, bld::Cam cam, int maxIters, double tol)
{
    __shared__ double red[BT / 32][36];        // reused as [4][28] / [4][21]
    __shared__ double pose[6], poseB[6], dP[6];
    __shared__ bld::PoseLin P;
    __shared__ int bad;
    const int k = blockIdx.x, tid = threadIdx.x;
    const int f0 = A.featPre[k], n = A.featPre[k + 1] - f0;
    if (tid < 6) pose[tid] = A.pose[6 * (size_t)k + tid];
    if (tid == 0) bad = 0;
    if (!A.haveX)
        for (int f = tid; f < n; f += BT) bld::triangulate(cam, A.z0 + 3 * (size_t)(f0 + f), A.X + 3 * (size_t)(f0 + f));
    __syncthreads();
    // Levenberg-Marquardt: every pass linearises at the current estimate; if the cost went up since
    // the last accepted point the step is undone and lambda grows tenfold, otherwise the point is
    // accepted (lambda shrinks tenfold unless the previous pass was a rejection) and a damped step is
    // taken (V + lambda diag V per landmark, S + lambda diag S on the pose).  All control values are
    // CTA-uniform (derived from the reduced sums in shared memory).
    double lam = 1e-3, costPrev = INFINITY;
    bool rejected = false;
    if (tid < 6) poseB[tid] = pose[tid];
    for (int f = tid; f < n; f += BT) {
        const size_t g = (size_t)(f0 + f);
        A.Xb[3 * g] = A.X[3 * g]; A.Xb[3 * g + 1] = A.X[3 * g + 1]; A.Xb[3 * g + 2] = A.X[3 * g + 2];
    }
    int it = 0;
    while (it < maxIters) {
        it++;
        const double lamStep = rejected ? lam : fmax(lam * 0.1, 1e-12);    // lambda of the step if this point is accepted
        if (tid == 0) bld::pose_lin(pose, P);
        __syncthreads();
        double acc[28];
#pragma unroll
        for (int i = 0; i < 28; i++) acc[i] = 0.0;
        for (int f = tid; f < n; f += BT) {
            const size_t g = (size_t)(f0 + f);
            double X[3] = {A.X[3 * g], A.X[3 * g + 1], A.X[3 * g + 2]};
            double V[9], W[18], Ub[36], gF[3], gP[6], Vi[9], WVi[18];
            acc[27] += bld::feature_blocks(cam, P, X, A.z0 + 3 * g, A.z1 + 3 * g, V, W, Ub, gF, gP);
            V[0] *= 1.0 + lamStep; V[4] *= 1.0 + lamStep; V[8] *= 1.0 + lamStep;
            bld::inv3_sym(V, Vi);
#pragma unroll
            for (int i = 0; i < 6; i++)
#pragma unroll
                for (int j = 0; j < 3; j++)
                    WVi[3 * i + j] = W[3 * i] * Vi[j] + W[3 * i + 1] * Vi[3 + j] + W[3 * i + 2] * Vi[6 + j];
            int q = 0;
#pragma unroll
            for (int i = 0; i < 6; i++) {
#pragma unroll
                for (int j = i; j < 6; j++)
                    acc[q++] += Ub[6 * i + j] - (WVi[3 * i] * W[3 * j] + WVi[3 * i + 1] * W[3 * j + 1] + WVi[3 * i + 2] * W[3 * j + 2]);
                acc[21 + i] += gP[i] - (WVi[3 * i] * gF[0] + WVi[3 * i + 1] * gF[1] + WVi[3 * i + 2] * gF[2]);
            }
#pragma unroll
            for (int i = 0; i < 18; i++) A.W[18 * g + i] = W[i];
#pragma unroll
            for (int i = 0; i < 9; i++) A.V[9 * g + i] = Vi[i];
            A.gF[3 * g] = gF[0]; A.gF[3 * g + 1] = gF[1]; A.gF[3 * g + 2] = gF[2];
        }
        cta_sum<28>(acc, reinterpret_cast<double (*)[28]>(&red[0][0]), tid);
        const double cost = red[0][27];
        if (!(cost <= costPrev * (1.0 + 1e-9))) {
            // the last step made it worse: back to the accepted point, more damping
            __syncthreads();
            if (tid < 6) pose[tid] = poseB[tid];
            for (int f = tid; f < n; f += BT) {
                const size_t g = (size_t)(f0 + f);
                A.X[3 * g] = A.Xb[3 * g]; A.X[3 * g + 1] = A.Xb[3 * g + 1]; A.X[3 * g + 2] = A.Xb[3 * g + 2];
            }
            lam *= 10.0;
            rejected = true;
            __syncthreads();
            continue;
        }
        costPrev = cost;
        lam = lamStep;
        rejected = false;
        if (tid == 0) {
            const double *r = &red[0][0];
            double S[36], e[6];
            int q = 0;
            for (int i = 0; i < 6; i++) {
                for (int j = i; j < 6; j++) { S[6 * i + j] = r[q]; S[6 * j + i] = r[q]; q++; }
                e[i] = r[21 + i];
            }
            for (int i = 0; i < 6; i++) S[7 * i] *= 1.0 + lam;
            if (!bld::solve6_spd(S, e)) { bad = 1; for (int i = 0; i < 6; i++) e[i] = 0.0; }
            for (int i = 0; i < 6; i++) dP[i] = e[i];
        }
        if (tid < 6) poseB[tid] = pose[tid];
        __syncthreads();
        if (bad) break;
        for (int f = tid; f < n; f += BT) {
            const size_t g = (size_t)(f0 + f);
            double t[3];
#pragma unroll
            for (int j = 0; j < 3; j++) {
                double s = A.gF[3 * g + j];
#pragma unroll
                for (int i = 0; i < 6; i++) s -= A.W[18 * g + 3 * i + j] * dP[i];
                t[j] = s;
            }
            const double *Vi = A.V + 9 * g;
#pragma unroll
            for (int i = 0; i < 3; i++) {
                const double x = A.X[3 * g + i];
                A.Xb[3 * g + i] = x;                   // the accepted point
                A.X[3 * g + i] = x + (Vi[3 * i] * t[0] + Vi[3 * i + 1] * t[1] + Vi[3 * i + 2] * t[2]);
            }
        }
        double mx = 0.0;
#pragma unroll
        for (int i = 0; i < 6; i++) mx = fmax(mx, fabs(dP[i]));
        __syncthreads();
        if (tid < 6) pose[tid] += dP[tid];
        __syncthreads();
        if (mx < tol) break;
    }
    // information at the estimate
    if (tid == 0) bld::pose_lin(pose, P);
    __syncthreads();
    double accU[21];
#pragma unroll
    for (int i = 0; i < 21; i++) accU[i] = 0.0;
    for (int f = tid; f < n; f += BT) {
        const size_t g = (size_t)(f0 + f);
        double X[3] = {A.X[3 * g], A.X[3 * g + 1], A.X[3 * g + 2]};
        double V[9], W[18], Ub[36], gF[3], gP[6];
        bld::feature_blocks(cam, P, X, A.z0 + 3 * g, A.z1 + 3 * g, V, W, Ub, gF, gP);
        int q = 0;
#pragma unroll
        for (int i = 0; i < 6; i++)
#pragma unroll
            for (int j = i; j < 6; j++) accU[q++] += Ub[6 * i + j];
#pragma unroll
        for (int i = 0; i < 18; i++) A.W[18 * g + i] = W[i];
#pragma unroll
        for (int i = 0; i < 9; i++) A.V[9 * g + i] = V[i];
    }
    cta_sum<21>(accU, reinterpret_cast<double (*)[21]>(&red[0][0]), tid);
    if (tid == 0) {
        const double *r = &red[0][0];
        double *U = A.U + 36 * (size_t)k;
        int q = 0;
        for (int i = 0; i < 6; i++)
            for (int j = i; j < 6; j++) { U[6 * i + j] = r[q]; U[6 * j + i] = r[q]; q++; }
        A.iters[k] = it;
        A.status[k] = bad;
    }
    if (tid < 6) A.pose[6 * (size_t)k + tid] = pose[tid];
}

} // namespace

void build_localmaps_stereo(Context &ctx, const lsfm_stereo_pair *pairs, int K, const lsfm_stereo_cam &cam,
                            int max_iters, double tol, lsfm_map *out, int *iters_done)
{
    if (K <= 0) return;
    if (!(cam.f > 0.0) || !(cam.baseline > 0.0) || !(cam.sigma > 0.0))
        throw LsfmError(LSFM_ERR_ARG, "builder: focal length, baseline and sigma must be positive");
    cudaStream_t s = ctx.stream;
    std::vector<int> featPre(K + 1, 0);
    bool haveX = true;
    for (int k = 0; k < K; k++) {
        if (pairs[k].n <= 0 || !pairs[k].z0 || !pairs[k].z1 || !pairs[k].pose0 || !pairs[k].feat_id)
            throw LsfmError(LSFM_ERR_ARG, "builder: pair " + std::to_string(k) + " is empty or has null arrays");
        featPre[k + 1] = featPre[k] + pairs[k].n;
        if (!pairs[k].X0) haveX = false;
    }
    const size_t tot = (size_t)featPre[K];
    std::vector<double> hz0(3 * tot), hz1(3 * tot), hX(3 * tot, 0.0), hpose(6 * (size_t)K);
    for (int k = 0; k < K; k++) {
        const size_t o = 3 * (size_t)featPre[k], c = 3 * (size_t)pairs[k].n;
        memcpy(hz0.data() + o, pairs[k].z0, sizeof(double) * c);
        memcpy(hz1.data() + o, pairs[k].z1, sizeof(double) * c);
        if (haveX) memcpy(hX.data() + o, pairs[k].X0, sizeof(double) * c);
        memcpy(hpose.data() + 6 * (size_t)k, pairs[k].pose0, sizeof(double) * 6);
    }
    DevBuf<int> dPre(K + 1, s), dIters(K, s), dStatus(K, s);
    DevBuf<double> dz0(3 * tot, s), dz1(3 * tot, s), dX(3 * tot, s), dPose(6 * (size_t)K, s);
    DevBuf<double> dW(18 * tot, s), dV(9 * tot, s), dU(36 * (size_t)K, s), dgF(3 * tot, s), dXb(3 * tot, s);
    dPre.upload(featPre); dz0.upload(hz0); dz1.upload(hz1); dPose.upload(hpose);
    if (haveX) dX.upload(hX);
    BuildArgs A;
    A.featPre = dPre.p; A.z0 = dz0.p; A.z1 = dz1.p; A.pose = dPose.p; A.X = dX.p;
    A.W = dW.p; A.V = dV.p; A.U = dU.p; A.gF = dgF.p; A.Xb = dXb.p; A.iters = dIters.p; A.status = dStatus.p;
    A.haveX = haveX ? 1 : 0;
    bld::Cam c{cam.f, cam.baseline, cam.cx, cam.cy, 1.0 / (cam.sigma * cam.sigma)};
    ctx.begin("builder");
    k_build_stereo<<<K, BT, 0, s>>>(A, c, max_iters, tol);
    KERNEL_CHECK();
    // measurements in (144 B / landmark incl. the guess), state + blocks out (18+9+3 doubles / landmark)
    ctx.end(48.0 * tot + 264.0 * tot, 0.0, 1);
    std::vector<double> hW(18 * tot), hV(9 * tot), hU(36 * (size_t)K);
    std::vector<int> hIters(K), hStatus(K);
    dX.download(hX.data(), 3 * tot); dPose.download(hpose.data(), 6 * (size_t)K);
    dW.download(hW.data(), 18 * tot); dV.download(hV.data(), 9 * tot); dU.download(hU.data(), 36 * (size_t)K);
    dIters.download(hIters.data(), K); dStatus.download(hStatus.data(), K);
    CUDA_CHECK(cudaStreamSynchronize(s));
    for (int k = 0; k < K; k++)
        if (hStatus[k])
            throw LsfmError(LSFM_ERR_NOT_SPD, "builder: pose system of pair " + std::to_string(k) + " is not positive definite");
    auto A_ = [](size_t n, size_t sz) { return malloc((n * sz) != 0 ? (n * sz) : 1); };
    for (int k = 0; k < K; k++) {
        const int n = pairs[k].n;
        const size_t o = (size_t)featPre[k];
        lsfm_map &M = out[k];
        memset(&M, 0, sizeof(M));
        M.Ref = pairs[k].Ref; M.FRef = pairs[k].Ref; M.m = 1; M.n = n; M.nU = 1; M.nW = n; M.r = 6 + 3 * n;
        M.stno = (int *)A_(M.r, sizeof(int));
        M.stVal = (double *)A_(M.r, sizeof(double));
        for (int i = 0; i < 6; i++) { M.stno[i] = -pairs[k].pose_id; M.stVal[i] = hpose[6 * (size_t)k + i]; }
        for (int f = 0; f < n; f++)
            for (int i = 0; i < 3; i++) { M.stno[6 + 3 * f + i] = pairs[k].feat_id[f]; M.stVal[6 + 3 * f + i] = hX[3 * (o + f) + i]; }
        M.U = (double *)A_(36, sizeof(double));
        memcpy(M.U, hU.data() + 36 * (size_t)k, sizeof(double) * 36);
        M.Ui = (int *)A_(1, sizeof(int)); M.Uj = (int *)A_(1, sizeof(int));
        M.Ui[0] = 0; M.Uj[0] = 0;
        M.W = (double *)A_(18 * (size_t)n, sizeof(double));
        memcpy(M.W, hW.data() + 18 * o, sizeof(double) * 18 * (size_t)n);
        M.V = (double *)A_(9 * (size_t)n, sizeof(double));
        memcpy(M.V, hV.data() + 9 * o, sizeof(double) * 9 * (size_t)n);
        M.photo = (int *)A_(n, sizeof(int)); M.feature = (int *)A_(n, sizeof(int)); M.FBlock = (int *)A_(n, sizeof(int));
        for (int f = 0; f < n; f++) { M.photo[f] = 0; M.feature[f] = f; M.FBlock[f] = f; }
        if (iters_done) iters_done[k] = hIters[k];
    }
}
