// Batched solve of the joint linear systems: rows a8-a13 of SURVEY 8(a).
//
// Reference: CLinearSFMImp::lmj_solveLinearSFMStereo, LinearSFMImp.cpp:2119-2378
//   block mask / Sidxij (O(m^2) char mask)       2131-2205  -> sorted unique pair keys (O(nnz))
//   pba_inverseV                                  3022-3042
//   S = U - W V^-1 W^T,  E = ea - W V^-1 eb       2214-2332
//   CSC build + CHOLMOD (AMD, factorize, solve)   2334-2361, 2380-2549 -> block multifrontal
//                                                 Cholesky on the GPU (chol_symbolic.cpp + here)
//   pba_solveFeatures                             2980-3020
#include "ops.h"
#include "chol_symbolic.h"
#include "small_mat.cuh"
#include "det_accum.cuh"
#include <cub/cub.cuh>
#include <thread>
#include <chrono>

namespace {

typedef unsigned long long u64;
__host__ __device__ inline u64 pair_key(int k, int a, int b)
{
    int lo = a < b ? a : b, hi = a < b ? b : a;
    return ((u64)(unsigned)k << 44) | ((u64)(unsigned)lo << 22) | (u64)(unsigned)hi;
}

// ---------------------------------------------------------------------------------------------
// a8: block pattern of S
// ---------------------------------------------------------------------------------------------
// Chunk-level exact dedupe (default path): one CTA per chunk of consecutive features builds the
// chunk's local pose table (bitmap + popcount prefix) and a bitmap over the <= 496 local pose
// pairs; a pair bit is set only if ONE feature is seen by both poses (the reference's smask rule,
// LinearSFMImp.cpp:2156-2173), so the pattern stays bit-exact while ~100x fewer keys reach the sort.
constexpr int PAT_CMAX = 31;
constexpr int PAT_THREADS = 128;
struct FeatChunk { int k, f0, f1; };
// Per chunk, written by pass 0 and reused by pass 1 and by the Schur kernel:
//   [0..30] local pose table (ascending pose index), [31] #distinct poses, [32..47] pair bitmap,
//   [48] mask of the chunk's DENSE local poses (seen by at least half of the chunk's features, at most
//   SCH_HMAX of them: their pose pairs are handled as one dense DMMA contraction), [49] their number
constexpr int CHUNK_INFO_INTS = 64;
constexpr int SCH_HMAX = 16;

// set bit (lo,hi) of join k's upper-triangular pose-pair bitmap (row lo, W words per row); the plain
// read first keeps the hub pairs (set by every chunk of a map) from turning into atomic traffic
__device__ __forceinline__ void bm_set(unsigned *__restrict__ bm, int off, int W, int a, int b)
{
    int lo = a < b ? a : b, hi = a < b ? b : a;
    unsigned *w = bm + off + lo * W + (hi >> 5);
    unsigned bit = 1u << (hi & 31);
    if ((*(volatile unsigned *)w & bit) == 0u) atomicOr(w, bit);
}

__global__ void __launch_bounds__(PAT_THREADS)
k_pat_chunk(const DMap *__restrict__ J, const FeatChunk *__restrict__ chunks,
            unsigned *__restrict__ bm, const int *__restrict__ bmOff,
            int *__restrict__ maxNposes, int pat_cmax, int *__restrict__ chunkInfo,
            int *__restrict__ blkInfo, const int *__restrict__ wPre,
            unsigned *__restrict__ patBits, int bitsStride, unsigned long long *__restrict__ pairFeat,
            const int *__restrict__ posePre, int *__restrict__ poseCR)
{
    extern __shared__ unsigned smu[];
    const FeatChunk ch = chunks[blockIdx.x];
    const DMap &M = J[ch.k];
    const int words = (M.m + 31) >> 5;
    unsigned *bitmap = smu;                          // [words]
    int *prefix = (int *)(bitmap + words);           // [words]
    unsigned *pairBits = (unsigned *)(prefix + words);   // [16]
    int *poses = (int *)(pairBits + 16);             // [PAT_CMAX + 1]
    int *misc = poses + PAT_CMAX + 1;                // [0] nposes
    int *scnt = misc + 4;                            // [32] blocks per local pose
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
    int *ci = chunkInfo + CHUNK_INFO_INTS * (size_t)blockIdx.x;
    const int boff = bmOff[ch.k];
    const int w0 = M.wPtr[ch.f0], w1 = M.wPtr[ch.f1];
    for (int i = tid; i < words; i += nt) bitmap[i] = 0u;
    if (tid < 16) pairBits[tid] = 0u;
    if (tid < 32) scnt[tid] = 0;
    __syncthreads();
    for (int j = w0 + tid; j < w1; j += nt) {
        int p = M.photo[j];
        atomicOr(&bitmap[p >> 5], 1u << (p & 31));
    }
    __syncthreads();
    if (warp == 0) {
        int run = 0;
        for (int base = 0; base < words; base += 32) {
            int c = (base + lane < words) ? __popc(bitmap[base + lane]) : 0;
            int incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            if (base + lane < words) prefix[base + lane] = run + incl - c;
            run += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) misc[0] = run;
    }
    __syncthreads();
    const int nposes = misc[0];
    if (tid == 0) { atomicMax(maxNposes, nposes); ci[31] = nposes; }
    // algorithmic work of the Schur complement: sum over the chunk's features of k_f (k_f + 1) / 2
    // 6x3x6 products (LinearSFMImp.cpp:2275-2318) -- statistics only (FP64 roofline of the Schur kernel)
    if (pairFeat) {
        unsigned long long c = 0;
        for (int f = ch.f0 + tid; f < ch.f1; f += nt) {
            const unsigned long long kf = (unsigned long long)(M.wPtr[f + 1] - M.wPtr[f]);
            c += kf * (kf + 1) / 2;
        }
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (lane == 0 && c) atomicAdd(pairFeat, c);
    }
    // pose bitmap + popcount prefix of the chunk -> global: the E gather (k_e_gather) finds a pose's slot
    // in this chunk with two loads (overflow chunks publish an empty bitmap: they record nothing)
    {
        unsigned *gb = patBits + (size_t)blockIdx.x * 2 * bitsStride;
        for (int i = tid; i < words; i += nt) {
            gb[i] = (nposes <= pat_cmax) ? bitmap[i] : 0u;
            gb[bitsStride + i] = (unsigned)prefix[i];
        }
    }
    if (nposes <= pat_cmax) {
        int *bi = blkInfo + wPre[ch.k];
        for (int f = ch.f0 + tid; f < ch.f1; f += nt) {
            int a0 = M.wPtr[f], a1 = M.wPtr[f + 1];
            for (int a = a0; a < a1; a++) {
                int pa = M.photo[a];
                int sa = prefix[pa >> 5] + __popc(bitmap[pa >> 5] & ((1u << (pa & 31)) - 1u));
                bi[a] = ((f - ch.f0) << 8) | sa;
                atomicAdd(&scnt[sa], 1);
                for (int b = a; b < a1; b++) {
                    int pb = M.photo[b];
                    int sb = prefix[pb >> 5] + __popc(bitmap[pb >> 5] & ((1u << (pb & 31)) - 1u));
                    int i = min(sa, sb), j = max(sa, sb);
                    int idx = i * nposes - (i * (i - 1)) / 2 + (j - i);
                    atomicOr(&pairBits[idx >> 5], 1u << (idx & 31));
                }
            }
        }
        for (int i = tid; i < words; i += nt) {
            unsigned b = bitmap[i];
            int r = prefix[i];
            while (b) {
                int bit = __ffs(b) - 1;
                poses[r] = i * 32 + bit; ci[r] = i * 32 + bit; r++;
                // range of chunks that will hold an E record of the pose (running maxima over zeroed memory):
                // k_e_gather walks only those
                int *cr = poseCR + 2 * (size_t)(posePre[ch.k] + i * 32 + bit);
                atomicMax(cr, 0x7fffffff - (int)blockIdx.x);
                atomicMax(cr + 1, (int)blockIdx.x + 1);
                b &= b - 1;
            }
        }
        __syncthreads();
        // the chunk's pair bitmap stays behind for the Schur kernel ([32..47])
        if (tid < 16) ci[32 + tid] = (int)pairBits[tid];
        if (tid == 0) {
            unsigned dm = 0u;
            int H = 0;
            const int nf = ch.f1 - ch.f0;
            for (int sl = 0; sl < nposes && H < SCH_HMAX; sl++)
                if (2 * scnt[sl] >= nf) { dm |= 1u << sl; H++; }
            ci[48] = (int)dm;
            ci[49] = H;
        }
        // the chunk's distinct pairs -> the join's pose-pair bitmap (exact dedupe across chunks)
        const int npairs = nposes * (nposes + 1) / 2;
        for (int t = tid; t < npairs; t += nt) {
            if (!((pairBits[t >> 5] >> (t & 31)) & 1u)) continue;
            int r = t, i = 0;
            while (r >= nposes - i) { r -= nposes - i; i++; }
            bm_set(bm, boff, words, poses[i], poses[i + r]);
        }
        return;
    }
    // overflow: too many distinct poses in this chunk -> per-feature pairs straight to the bitmap
    for (int f = ch.f0 + tid; f < ch.f1; f += nt) {
        int a0 = M.wPtr[f], a1 = M.wPtr[f + 1];
        for (int a = a0; a < a1; a++)
            for (int b = a; b < a1; b++) bm_set(bm, boff, words, M.photo[a], M.photo[b]);
    }
}

__global__ void k_pat_u(const DMap *__restrict__ J, const int *__restrict__ uPre, int K, int totU,
                        unsigned *__restrict__ bm, const int *__restrict__ bmOff)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= totU) return;
    int k = seg_find(uPre, K, g);
    int b = g - uPre[k];
    bm_set(bm, bmOff[k], (J[k].m + 31) >> 5, J[k].Ui[b], J[k].Uj[b]);
}

// one warp per block row (global pose g): number of pattern blocks in the row
__global__ void k_bm_rowcount(const unsigned *__restrict__ bm, const int *__restrict__ bmOff,
                              const DMap *__restrict__ J, const int *__restrict__ posePre, int K, int totP,
                              int *__restrict__ rowCnt)
{
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (g > totP) return;
    if (g == totP) { if (lane == 0) rowCnt[g] = 0; return; }
    const int k = seg_find(posePre, K, g);
    const int r = g - posePre[k];
    const int W = (J[k].m + 31) >> 5;
    const unsigned *row = bm + bmOff[k] + (size_t)r * W;
    int c = 0;
    for (int w = (r >> 5) + lane; w < W; w += 32) c += __popc(row[w]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) rowCnt[g] = c;
}

// one warp per block row: the row's keys (join << 44 | row << 22 | col) in ascending column order
// at rowPtr[g]; rows are consecutive, so the key list is sorted by (join, row, col)
__global__ void k_bm_emit(const unsigned *__restrict__ bm, const int *__restrict__ bmOff,
                          const DMap *__restrict__ J, const int *__restrict__ posePre, int K, int totP,
                          const int *__restrict__ rowPtr, u64 *__restrict__ keys, int cap)
{
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (g >= totP) return;
    const int k = seg_find(posePre, K, g);
    const int r = g - posePre[k];
    const int W = (J[k].m + 31) >> 5;
    const unsigned *row = bm + bmOff[k] + (size_t)r * W;
    int base = rowPtr[g];
    for (int w0 = (r >> 5); w0 < W; w0 += 32) {
        const int w = w0 + lane;
        unsigned bits = (w < W) ? row[w] : 0u;
        int c = __popc(bits), incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        int o = base + incl - c;
        while (bits) {
            int bit = __ffs(bits) - 1;
            bits &= bits - 1;
            if (o < cap) keys[o] = pair_key(k, r, w * 32 + bit);
            o++;
        }
        base += __shfl_sync(0xffffffffu, incl, 31);
    }
}

__device__ __forceinline__ int find_slot(const u64 *__restrict__ keys, const int *__restrict__ rowPtr,
                                         int gposeRow, u64 key)
{
    int lo = rowPtr[gposeRow], hi = rowPtr[gposeRow + 1];
    while (lo < hi) { int mid = (lo + hi) >> 1; if (keys[mid] < key) lo = mid + 1; else hi = mid; }
    return lo;          // the pattern contains every queried pair by construction
}

// ---------------------------------------------------------------------------------------------
// a9: V^-1 by the cofactor formula Eigen's Matrix3d::inverse uses, symmetrised from the upper
// triangle exactly as pba_inverseV does (LinearSFMImp.cpp:3035-3040)
// ---------------------------------------------------------------------------------------------
// Also writes, per feature, the two 3-vectors the Schur kernels need for the reduced right-hand side:
//     d_side = V^-1 eF_f - xhat_f^side          (side = End / Cur half of the joint pose list)
// so that E_p = eP_p - sum_f W_pf d_side(p): with eP holding only the U part of the reference's eP
// (2747-2790) this equals eP_ref - W V^-1 eF (2246-2332) without the join ever accumulating
// W xhat_f per pose.  xhat == nullptr (stand-alone solve operator, mono): d = V^-1 eF, eP as given.
__global__ void k_vinv(const DMap *__restrict__ J, const int *__restrict__ featPre, int K, int totF,
                       const double *__restrict__ eF, const double *__restrict__ xhat,
                       double *__restrict__ Vinv, double *__restrict__ dvec)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= totF) return;
    int k = seg_find(featPre, K, g);
    int f = g - featPre[k];
    double a[9];
    sm::load<9>(J[k].V + 9 * (size_t)f, a);
    double c00 = a[4] * a[8] - a[5] * a[7];
    double c10 = a[5] * a[6] - a[3] * a[8];     // cofactor of (1,0)... first-column cofactors
    double c20 = a[3] * a[7] - a[4] * a[6];
    double det = a[0] * c00 + a[1] * c10 + a[2] * c20;
    double id = 1.0 / det;
    double i00 = c00 * id;
    double i01 = (a[2] * a[7] - a[1] * a[8]) * id;
    double i02 = (a[1] * a[5] - a[2] * a[4]) * id;
    double i11 = (a[0] * a[8] - a[2] * a[6]) * id;
    double i12 = (a[2] * a[3] - a[0] * a[5]) * id;
    double i22 = (a[0] * a[4] - a[1] * a[3]) * id;
    double *o = Vinv + 9 * (size_t)g;
    o[0] = i00; o[1] = i01; o[2] = i02;
    o[3] = i01; o[4] = i11; o[5] = i12;
    o[6] = i02; o[7] = i12; o[8] = i22;
    double e0 = eF[3 * (size_t)g], e1 = eF[3 * (size_t)g + 1], e2 = eF[3 * (size_t)g + 2];
    double v0 = i00 * e0 + i01 * e1 + i02 * e2;
    double v1 = i01 * e0 + i11 * e1 + i12 * e2;
    double v2 = i02 * e0 + i12 * e1 + i22 * e2;
    double *d = dvec + 6 * (size_t)g;
    if (xhat) {
        const double *x = xhat + 6 * (size_t)g;
        d[0] = v0 - x[0]; d[1] = v1 - x[1]; d[2] = v2 - x[2];
        d[3] = v0 - x[3]; d[4] = v1 - x[4]; d[5] = v2 - x[5];
    } else {
        d[0] = v0; d[1] = v1; d[2] = v2; d[3] = v0; d[4] = v1; d[5] = v2;
    }
}

// ---------------------------------------------------------------------------------------------
// a10: S = U - W V^-1 W^T, E = eP - W V^-1 eF
// ---------------------------------------------------------------------------------------------
__global__ void k_s_from_u(const DMap *__restrict__ J, const int *__restrict__ uPre,
                           const int *__restrict__ posePre, int K, int totU,
                           const u64 *__restrict__ keys, const int *__restrict__ rowPtr,
                           double *__restrict__ S)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= totU) return;
    int k = seg_find(uPre, K, g);
    int b = g - uPre[k];
    int i = J[k].Ui[b], j = J[k].Uj[b];
    int slot = find_slot(keys, rowPtr, posePre[k] + i, pair_key(k, i, j));
    const double *u = J[k].U + 36 * (size_t)b;
    double *s = S + 36 * (size_t)slot;
    // blocks of one list are distinct in stereo; atomics keep the mono duplicate case (+=) valid
#pragma unroll
    for (int q = 0; q < 36; q++) atomicAdd(s + q, u[q]);
}

// v1: one thread per (W block a); loops over the blocks b of the same feature with photo_b >= photo_a
__global__ void __launch_bounds__(128)
k_schur(const DMap *__restrict__ J, const int *__restrict__ wPre, const int *__restrict__ featPre,
        const int *__restrict__ posePre, int K, int totW, const double *__restrict__ Vinv,
        const double *__restrict__ dvec, const int *__restrict__ split, const u64 *__restrict__ keys,
        const int *__restrict__ rowPtr, double *__restrict__ S, double *__restrict__ E)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    bool live = g < totW;
    int k = live ? seg_find(wPre, K, g) : 0;
    int a = live ? g - wPre[k] : 0;
    const DMap &M = J[k];
    int f = live ? M.feature[a] : 0;
    int pa = live ? M.photo[a] : 0;
    double WV[18], y[6];
    if (live) {
        double W[18], Vi[9];
        sm::load<18>(M.W + 18 * (size_t)a, W);
        sm::load<9>(Vinv + 9 * (size_t)(featPre[k] + f), Vi);
        sm::mmt<6, 3, 3>(W, Vi, WV);                       // WV[i][j] = sum_k W[i][k] Vinv[j][k] (2269-2270)
        double d[3];
        sm::load<3>(dvec + 6 * (size_t)(featPre[k] + f) + ((split && pa >= split[k]) ? 3 : 0), d);
        sm::mm<6, 3, 1>(W, d, y);
#pragma unroll
        for (int q = 0; q < 6; q++) y[q] = -y[q];
    }
    sm::warp_agg_atomic_add<6>(E + 6 * (size_t)(posePre[k] + pa), y, live);
    if (!live) return;
    int w0 = M.wPtr[f], w1 = M.wPtr[f + 1];
    for (int b = w0; b < w1; b++) {
        int pb = M.photo[b];
        if (pb < pa) continue;
        double Wb[18], P[36];
        sm::load<18>(M.W + 18 * (size_t)b, Wb);
        sm::mmt<6, 3, 6>(WV, Wb, P);
        int slot = find_slot(keys, rowPtr, posePre[k] + pa, pair_key(k, pa, pb));
        double *s = S + 36 * (size_t)slot;
#pragma unroll
        for (int q = 0; q < 36; q++) atomicAdd(s + q, -P[q]);
    }
}

// ---- order-independent accumulation of S ------------------------------------------------------
// The chunks' contributions to one S block arrive in an order the hardware chooses.  FP64 atomics would
// make the last bits of S differ from run to run; instead every contribution is converted to 64-bit
// FIXED POINT and added with integer atomics, which are exact and therefore order-independent.  The
// scale of entry (6i+r, 6j+c) comes from a bound that holds for every partial sum: the subtracted
// matrix D = sum_f W V^-1 W^T is positive semi-definite with D <= U on the diagonal (S = U - D must be
// positive definite), so |D(ir,jc)| <= sqrt(U(ir,ir) U(jc,jc)).  With ex(x) = ilogb(x) + 1 the entry is
// below 2^eb, eb = ceil((ex_ir + ex_jc) / 2), and is stored as round(x 2^(60 - eb)): |sum| < 2^60, the
// quantum is 2^-60 of the bound (double's own quantum is 2^-53 of the value).
__device__ __forceinline__ int row_exp_of(double u) { return (u > 0.0 && u < 1e300) ? ilogb(u) + 1 : 0; }
__device__ __forceinline__ int fx_shift(int ea, int eb_) { return 60 - ((ea + eb_ + 1) >> 1); }
__device__ __forceinline__ double pow2(int e) { return __longlong_as_double((long long)(1023 + e) << 52); }

// Fixed-point accumulation of E for the slow path below: the contributions y = W_pf d_f have no a-priori
// bound, so a pre-pass over the overflow chunks (k_schur_slow_pre) records, per (pose, row), the largest
// exponent of a contribution and, per pose, how many there are; every partial sum then stays below
// 2^(emax + ceil(log2(count + 1))) and is accumulated as round(y 2^shift) with integer atomics.
struct SlowFx {
    const int *sexp;            // [6 totPose] exponents of the diagonal of U (S scale)
    long long *Sfx;             // [36 nuis]
    int *eexp;                  // [6 totPose] largest contribution exponent per (pose, row); < -2000: none
    int *ecnt;                  // [totPose]   contributions per pose
    long long *Efx;             // [6 totPose]
};
__device__ __forceinline__ int e_shift(int emax, int cnt) { return 60 - emax - (32 - __clz(cnt)); }

// Slow path for one W block, used when a chunk sees too many distinct poses: exact integer atomics into the
// fixed-point accumulators of S and E (order-independent, like the fast path's flush).
__device__ __noinline__ void schur_block_slow(const DMap &M, int k, int a, const int *__restrict__ featPre,
                                 const int *__restrict__ posePre, const double *__restrict__ Vinv,
                                 const double *__restrict__ dvec, const int *__restrict__ split,
                                 const u64 *__restrict__ keys,
                                 const int *__restrict__ rowPtr, const SlowFx fx)
{
    int f = M.feature[a], pa = M.photo[a];
    double W[18], Vi[9], WV[18], d[3], y[6];
    sm::load<18>(M.W + 18 * (size_t)a, W);
    sm::load<9>(Vinv + 9 * (size_t)(featPre[k] + f), Vi);
    sm::mmt<6, 3, 3>(W, Vi, WV);
    sm::load<3>(dvec + 6 * (size_t)(featPre[k] + f) + ((split && pa >= split[k]) ? 3 : 0), d);
    sm::mm<6, 3, 1>(W, d, y);
    const size_t gpa = (size_t)(posePre[k] + pa);
    const int cnt = fx.ecnt[gpa];
    for (int q = 0; q < 6; q++) {
        const int em = fx.eexp[6 * gpa + q];
        if (em < -2000) continue;                                       // every contribution to this row is 0
        const long long v = __double2ll_rn(-y[q] * pow2(e_shift(em, cnt)));
        atomicAdd(reinterpret_cast<unsigned long long *>(fx.Efx) + 6 * gpa + q, (unsigned long long)v);
    }
    int ea[6];
    for (int q = 0; q < 6; q++) ea[q] = fx.sexp[6 * gpa + q];
    for (int b = M.wPtr[f]; b < M.wPtr[f + 1]; b++) {
        int pb = M.photo[b];
        if (pb < pa) continue;
        double Wb[18], P[36];
        sm::load<18>(M.W + 18 * (size_t)b, Wb);
        sm::mmt<6, 3, 6>(WV, Wb, P);
        int slot = find_slot(keys, rowPtr, posePre[k] + pa, pair_key(k, pa, pb));
        unsigned long long *sp = reinterpret_cast<unsigned long long *>(fx.Sfx) + 36 * (size_t)slot;
        const size_t gpb = (size_t)(posePre[k] + pb);
        for (int r = 0; r < 6; r++)
            for (int c = 0; c < 6; c++) {
                const long long v = __double2ll_rn(P[6 * r + c] * pow2(fx_shift(ea[r], fx.sexp[6 * gpb + c])));
                atomicAdd(sp + 6 * r + c, (unsigned long long)v);
            }
    }
}

// pre-pass of the slow path: exponent / count bookkeeping of the E contributions of the overflow chunks
__global__ void __launch_bounds__(128)
k_schur_slow_pre(const DMap *__restrict__ J, const FeatChunk *__restrict__ chunks, const int *__restrict__ chunkInfo,
                 int pat_cmax, const int *__restrict__ featPre, const int *__restrict__ posePre,
                 const double *__restrict__ dvec, const int *__restrict__ split, int *__restrict__ eexp,
                 int *__restrict__ ecnt)
{
    const int np = chunkInfo[CHUNK_INFO_INTS * (size_t)blockIdx.x + 31];
    if (np <= 31 && np <= pat_cmax) return;
    const FeatChunk ch = chunks[blockIdx.x];
    const DMap &M = J[ch.k];
    const int k = ch.k;
    for (int a = M.wPtr[ch.f0] + threadIdx.x; a < M.wPtr[ch.f1]; a += blockDim.x) {
        const int f = M.feature[a], pa = M.photo[a];
        double W[18], d[3], y[6];
        sm::load<18>(M.W + 18 * (size_t)a, W);
        sm::load<3>(dvec + 6 * (size_t)(featPre[k] + f) + ((split && pa >= split[k]) ? 3 : 0), d);
        sm::mm<6, 3, 1>(W, d, y);
        const size_t gpa = (size_t)(posePre[k] + pa);
        atomicAdd(ecnt + gpa, 1);
        for (int q = 0; q < 6; q++)
            if (y[q] != 0.0 && fabs(y[q]) < 1e300) atomicMax(eexp + 6 * gpa + q, ilogb(y[q]) + 1);
    }
}

// E += the slow path's fixed-point sums (one thread per (pose, row))
__global__ void k_e_convert(int n6, const int *__restrict__ eexp, const int *__restrict__ ecnt,
                            const long long *__restrict__ Efx, double *__restrict__ E)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n6) return;
    const int em = eexp[g];
    if (em < -2000) return;
    E[g] += (double)Efx[g] * pow2(-e_shift(em, ecnt[g / 6]));
}

constexpr int SCH_FCHUNK = 128;        // features per chunk (pattern + Schur kernels)

// one thread per (pose, row): exponent of the diagonal entry U(ir,ir), read from S after k_s_from_u
__global__ void k_row_exp(const u64 *__restrict__ keys, const int *__restrict__ rowPtr,
                          const DMap *__restrict__ J, const int *__restrict__ posePre, int K, int totP,
                          const double *__restrict__ S, int *__restrict__ sexp)
{
    int g6 = blockIdx.x * blockDim.x + threadIdx.x;
    int g = g6 / 6, r = g6 - 6 * g;
    if (g >= totP) return;
    int k = seg_find(posePre, K, g);
    int p = g - posePre[k];
    int slot = rowPtr[g];
    int e = 0;
    if (slot < rowPtr[g + 1] && keys[slot] == pair_key(k, p, p)) e = row_exp_of(S[36 * (size_t)slot + 7 * r]);
    sexp[g6] = e;
}

// S -= fixed-point sums (one thread per entry)
__global__ void k_s_convert(const u64 *__restrict__ keys, int nuis, const int *__restrict__ posePre,
                            const int *__restrict__ sexp, const long long *__restrict__ Sfx,
                            double *__restrict__ S)
{
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t t = g / 36;
    if (t >= (size_t)nuis) return;
    int q = (int)(g - 36 * t), r = q / 6, c = q - 6 * r;
    const long long v = Sfx[g];
    if (v == 0) return;
    const u64 key = keys[t];
    const int k = (int)(key >> 44), lo = (int)((key >> 22) & ((1u << 22) - 1)), hi = (int)(key & ((1u << 22) - 1));
    const int sh = fx_shift(sexp[6 * (size_t)(posePre[k] + lo) + r], sexp[6 * (size_t)(posePre[k] + hi) + c]);
    S[g] -= (double)v * pow2(-sh);
}

// E_p -= sum over the chunks of the pose's join of the chunk's share (record [chunk][slot][6]); one warp
// per pose, chunks strided over the lanes, lane-private sums, fixed butterfly: bit-identical every run
__global__ void __launch_bounds__(128)
k_e_gather(const int *__restrict__ posePre, int K, int totP, const int *__restrict__ chunkPre,
           const unsigned *__restrict__ patBits, int bitsStride, const double *__restrict__ Erec,
           double *__restrict__ E, const int *__restrict__ poseCR)
{
    const int lane = threadIdx.x & 31;
    const int gp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (gp >= totP) return;
    const int k = seg_find(posePre, K, gp);
    const int p = gp - posePre[k];
    const int pw = p >> 5;
    const unsigned pbit = 1u << (p & 31);
    double acc[6] = {0, 0, 0, 0, 0, 0};
    // only the chunks between the first and the last one with a record of the pose; the start is rounded down to
    // a multiple of 32 chunks, so every chunk keeps its lane and the sums their order
    int cbeg = chunkPre[k], cend = chunkPre[k];
    {
        const int vmin = poseCR[2 * (size_t)gp], vmax = poseCR[2 * (size_t)gp + 1];
        if (vmax > 0) { cbeg += ((0x7fffffff - vmin) - cbeg) & ~31; cend = vmax; }
    }
    for (int c = cbeg + lane; c < cend; c += 32) {
        const unsigned *gb = patBits + (size_t)c * 2 * bitsStride;
        const unsigned w = gb[pw];
        if (w & pbit) {
            const int slot = (int)gb[bitsStride + pw] + __popc(w & (pbit - 1u));
            const double2 *r = reinterpret_cast<const double2 *>(Erec + 6 * (32 * (size_t)c + slot));
            const double2 v0 = r[0], v1 = r[1], v2 = r[2];
            acc[0] += v0.x; acc[1] += v0.y; acc[2] += v1.x; acc[3] += v1.y; acc[4] += v2.x; acc[5] += v2.y;
        }
    }
#pragma unroll
    for (int q = 0; q < 6; q++) acc[q] = sm::warp_sum(acc[q]);
    if (lane < 6) {
        double v = acc[0];
#pragma unroll
        for (int q = 1; q < 6; q++) if (lane == q) v = acc[q];
        E[6 * (size_t)gp + lane] -= v;
    }
}
} // namespace
#include "schur_pipe.cuh"
#include "schur_dense.cuh"
#include "schur_lock.cuh"
namespace {

// ---------------------------------------------------------------------------------------------
// a17 (mono): gauge rows.  The reference deletes the six rows/columns of the zero pose and the
// scalar row `Fix` from the CSC (pba_constructCSSGN, LinearSFMImp.cpp:7136-7170) and re-inserts
// zeros afterwards (7010-7026).  Here the rows/columns stay in the block system but are replaced by
// identity rows with a zero right-hand side -- same solution, block structure intact.
// ---------------------------------------------------------------------------------------------
__global__ void k_pat_diag(const DMap *__restrict__ J, const int *__restrict__ posePre, int K, int totP,
                           unsigned *__restrict__ bm, const int *__restrict__ bmOff)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= totP) return;
    int k = seg_find(posePre, K, g);
    int p = g - posePre[k];
    bm_set(bm, bmOff[k], (J[k].m + 31) >> 5, p, p);
}

__global__ void k_mono_gauge(const u64 *__restrict__ keys, int nuis, const int *__restrict__ refPose,
                             const int *__restrict__ fixScalar, double *__restrict__ S)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nuis) return;
    u64 key = keys[g];
    int k = (int)(key >> 44), lo = (int)((key >> 22) & ((1u << 22) - 1)), hi = (int)(key & ((1u << 22) - 1));
    int ref = refPose[k], fp = fixScalar[k] / 6, fr = fixScalar[k] % 6;
    double *s = S + 36 * (size_t)g;
    if (lo == ref || hi == ref) {
        for (int q = 0; q < 36; q++) s[q] = 0.0;
        if (lo == hi) for (int q = 0; q < 6; q++) s[7 * q] = 1.0;
        return;
    }
    if (lo == fp) for (int q = 0; q < 6; q++) s[6 * fr + q] = 0.0;
    if (hi == fp) for (int q = 0; q < 6; q++) s[6 * q + fr] = 0.0;
    if (lo == fp && hi == fp) s[7 * fr] = 1.0;
}

__global__ void k_mono_gauge_rhs(const int *__restrict__ posePre, int K, const int *__restrict__ refPose,
                                 const int *__restrict__ fixScalar, double *__restrict__ E)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    double *e = E + 6 * (size_t)posePre[k];
    for (int q = 0; q < 6; q++) e[6 * refPose[k] + q] = 0.0;
    e[fixScalar[k]] = 0.0;
}

__global__ void k_mono_fix(DMap *__restrict__ J, int K, const int *__restrict__ fixScalar,
                           const int *__restrict__ sign)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    J[k].poseVal[fixScalar[k]] = (double)sign[k];          // stVal[Fix] = Sign (7026)
}

// ---------------------------------------------------------------------------------------------
// a11-a12: numeric block multifrontal Cholesky with the right-hand side carried as an extra row
// ---------------------------------------------------------------------------------------------
__global__ void k_front_assemble(const SlotMap *__restrict__ slot, int nslot,
                                 const SnodeDesc *__restrict__ sn, const double *__restrict__ S,
                                 double *__restrict__ fronts)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    int i = g / 36, e = g - 36 * i;
    if (i >= nslot) return;
    SlotMap m = slot[i];
    const SnodeDesc &d = sn[m.sn];
    int ld = 6 * (d.ncols + d.nstruct) + 1;
    int x = e / 6, y = e - 6 * x;                       // front block entry (x,y)
    double v = m.transpose ? S[36 * (size_t)i + 6 * y + x] : S[36 * (size_t)i + 6 * x + y];
    fronts[d.frontOff + (size_t)(6 * m.lcol + y) * ld + 6 * m.lrow + x] = v;
}

__global__ void k_front_rhs(const int *__restrict__ poseSn, const int *__restrict__ poseLcol, int totP,
                            const SnodeDesc *__restrict__ sn, const double *__restrict__ E,
                            double *__restrict__ fronts)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    int p = g / 6, q = g - 6 * p;
    if (p >= totP) return;
    const SnodeDesc &d = sn[poseSn[p]];
    int fs = 6 * (d.ncols + d.nstruct), ld = fs + 1;
    fronts[d.frontOff + (size_t)(6 * poseLcol[p] + q) * ld + fs] = E[6 * (size_t)p + q];
}

// Dependency flags of the single-launch ("persistent") factor / back-solve kernels: one CTA per front,
// fronts in assembly-tree level order, so every front a CTA waits for belongs to a CTA with a SMALLER block
// index -- already finished or resident (the hardware hands out blocks in index order): no deadlock.  The
// spin is bounded anyway (~2 s): on a timeout the error flag is raised instead of hanging the device.
__device__ __forceinline__ void front_wait(const int *flag, int *errflag)
{
    const long long t0 = clock64();
    while (*(volatile const int *)flag == 0) {
        __nanosleep(100);
        if (clock64() - t0 > (1ll << 32)) { atomicOr(errflag, 4); break; }
    }
}
__device__ __forceinline__ void front_signal(int *flag)
{
    __syncthreads();
    if (threadIdx.x == 0) { __threadfence(); atomicExch(flag, 1); }
}

// Factor a panel held in (shared or global) memory: P is [pc][ldp] column-major, rows 0..pc-1 hold the
// diagonal block (lower triangle), rows pc..rows-1 everything below it.  6x6 diagonal blocks by one
// thread in registers (factor + its inverse), row solves and the in-panel updates by the whole CTA.
__device__ __forceinline__ void factor_panel(double *P, const int ldp, const int rows, const int pc, double *Li,
                                             int *errflag, const int tid, const int nt)
{
    const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
    for (int b0 = 0; b0 < pc; b0 += 6) {
        if (tid == 0) {
            // 6x6 Cholesky of the diagonal block and the inverse of its factor, in registers
            double a[6][6], li[6][6];
#pragma unroll
            for (int j = 0; j < 6; j++)
#pragma unroll
                for (int i = j; i < 6; i++) a[i][j] = P[(b0 + j) * ldp + b0 + i];
            bool bad = false;
#pragma unroll
            for (int j = 0; j < 6; j++) {
                double sdiag = a[j][j];
#pragma unroll
                for (int q = 0; q < j; q++) sdiag = fma(-a[j][q], a[j][q], sdiag);
                if (!(sdiag > 0.0)) { bad = true; sdiag = 1.0; }
                double djj = sqrt(sdiag);
                double inv = 1.0 / djj;
                a[j][j] = djj;
                li[j][j] = inv;
#pragma unroll
                for (int i = j + 1; i < 6; i++) {
                    double v = a[i][j];
#pragma unroll
                    for (int q = 0; q < j; q++) v = fma(-a[i][q], a[j][q], v);
                    a[i][j] = v * inv;
                }
            }
            // li = L^-1 (lower): li[i][j] = -li[i][i] * sum_{q=j}^{i-1} L[i][q] li[q][j]
#pragma unroll
            for (int j = 0; j < 6; j++)
#pragma unroll
                for (int i = j + 1; i < 6; i++) {
                    double v = 0.0;
#pragma unroll
                    for (int q = j; q < i; q++) v = fma(a[i][q], li[q][j], v);
                    li[i][j] = -li[i][i] * v;
                }
            if (bad) atomicOr(errflag, 1);
#pragma unroll
            for (int j = 0; j < 6; j++)
#pragma unroll
                for (int i = j; i < 6; i++) {
                    P[(b0 + j) * ldp + b0 + i] = a[i][j];
                    Li[6 * i + j] = li[i][j];
                }
        }
        __syncthreads();
        // rows below the diagonal block: X = A L^-T, no dependent chain per row
        for (int rr = b0 + 6 + tid; rr < rows; rr += nt) {
            double av[6], x[6];
#pragma unroll
            for (int q = 0; q < 6; q++) av[q] = P[(b0 + q) * ldp + rr];
#pragma unroll
            for (int q = 0; q < 6; q++) {
                double v = 0.0;
#pragma unroll
                for (int p = 0; p <= q; p++) v = fma(av[p], Li[6 * q + p], v);
                x[q] = v;
            }
#pragma unroll
            for (int q = 0; q < 6; q++) P[(b0 + q) * ldp + rr] = x[q];
        }
        __syncthreads();
        // remaining panel columns
        const int rem = pc - b0 - 6;
        if (rem > 0) {
            for (int c = b0 + 6 + warp; c < pc; c += nw) {
                double lc[6];
#pragma unroll
                for (int q = 0; q < 6; q++) lc[q] = P[(b0 + q) * ldp + c];
                for (int rr = c + lane; rr < rows; rr += 32) {
                    double v = P[c * ldp + rr];
#pragma unroll
                    for (int q = 0; q < 6; q++) v -= P[(b0 + q) * ldp + rr] * lc[q];
                    P[c * ldp + rr] = v;
                }
            }
            __syncthreads();
        }
    }
}

// one CTA per front of the level: extend-add the children, factor the leading ncols block columns,
// leave the Schur complement (+ updated rhs row) in place for the parent.
// The factorisation is blocked: a panel of up to `pcMax` columns (all rows below, rhs row included)
// is staged in shared memory, factored there (6x6 diagonal blocks by one warp, row solves and the
// in-panel updates by the whole CTA), applied ONCE to everything right of it in global memory
// (left-looking rank-pc update, panel read from shared memory) and written back.
// GP = true: fronts too tall for a shared-memory panel (> ~2100 rows: loop-closure-heavy or very large
// systems) keep the panel in a per-front global scratch area instead -- same code, slower, no failure.
template <bool GP>
__global__ void __launch_bounds__(512)
k_front_factor(const int *__restrict__ levelSn, const SnodeDesc *__restrict__ sn,
               const int *__restrict__ childIdx, const int *__restrict__ relIdx,
               double *__restrict__ fronts, int *__restrict__ errflag, int pcMax,
               double *__restrict__ panelG, size_t panelStride, int *__restrict__ done)
{
    extern __shared__ double Psh[];               // [pc][ldp] column-major panel
    double *P = GP ? panelG + (size_t)blockIdx.x * panelStride : Psh;
    __shared__ double Li[36];                     // inverse of the current diagonal block's factor
    const int sid = levelSn[blockIdx.x];
    const SnodeDesc d = sn[sid];
    if (done) {                                   // single-launch mode: wait for the children's fronts
        if (threadIdx.x == 0) {
            for (int ci = 0; ci < d.nchild; ci++) front_wait(done + childIdx[d.childOff + ci], errflag);
            __threadfence();
        }
        __syncthreads();
    }
    double *F = fronts + d.frontOff;
    const int fs = 6 * (d.ncols + d.nstruct), ld = fs + 1, nc = 6 * d.ncols;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;

    // extend-add of the children's update matrices (lower triangle + rhs row).  A warp walks one
    // child column at a time, four 32-row pieces per pass: all loads of a pass (child values and
    // the parent entries they land on) are issued before the first add, so the L2 latency of this
    // scattered read-modify-write is paid once per pass instead of once per entry.
    for (int ci = 0; ci < d.nchild; ci++) {
        const SnodeDesc c = sn[childIdx[d.childOff + ci]];
        const double *Fc = fronts + c.frontOff;
        const int ncc = 6 * c.ncols, us = 6 * c.nstruct, ldc = 6 * (c.ncols + c.nstruct) + 1;
        const int *rel = relIdx + c.structOff;
        const int rows = us + 1;
        for (int cc = 2 * warp; cc < us; cc += 2 * nw) {
            const bool has2 = (cc + 1 < us);
            const int pcx0 = 6 * rel[cc / 6] + (cc % 6);
            const int pcx1 = has2 ? 6 * rel[(cc + 1) / 6] + ((cc + 1) % 6) : pcx0;
            const double *src0 = Fc + (size_t)(ncc + cc) * ldc + ncc;
            const double *src1 = src0 + ldc;
            double *dst0 = F + (size_t)pcx0 * ld;
            double *dst1 = F + (size_t)pcx1 * ld;
            for (int r0 = cc + lane; r0 < rows; r0 += 128) {
                double cv0[4], pv0[4], cv1[4], pv1[4];
                int pr[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int r = r0 + 32 * j;
                    pr[j] = -1;
                    if (r < rows) {
                        pr[j] = (r == us) ? fs : 6 * rel[r / 6] + (r % 6);
                        cv0[j] = __ldcg(src0 + r);
                        pv0[j] = dst0[pr[j]];
                        if (has2 && r > cc) { cv1[j] = __ldcg(src1 + r); pv1[j] = dst1[pr[j]]; }
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if (pr[j] >= 0) {
                        dst0[pr[j]] = pv0[j] + cv0[j];
                        if (has2 && r0 + 32 * j > cc) dst1[pr[j]] = pv1[j] + cv1[j];
                    }
            }
        }
        __syncthreads();
    }

    for (int p0 = 0; p0 < nc; p0 += pcMax) {
        const int pc = min(pcMax, nc - p0);
        const int rows = fs + 1 - p0;             // front rows p0..fs (rhs row last)
        const int ldp = rows;
        for (int t = tid; t < pc * rows; t += nt) {
            int q = t / rows, rr = t - q * rows;
            P[q * ldp + rr] = (rr >= q) ? F[(size_t)(p0 + q) * ld + p0 + rr] : 0.0;
        }
        __syncthreads();
        factor_panel(P, ldp, rows, pc, Li, errflag, tid, nt);
        // write the factored panel back
        for (int t = tid; t < pc * rows; t += nt) {
            int q = t / rows, rr = t - q * rows;
            if (rr >= q) F[(size_t)(p0 + q) * ld + p0 + rr] = P[q * ldp + rr];
        }
        // rank-pc update of everything right of the panel (lower triangle + rhs row): a warp takes
        // two columns and four 32-row pieces at a time -- eight independent FMA chains per lane, the
        // global loads of a pass issued together, six shared loads per eight FMAs
        for (int c = p0 + pc + 2 * warp; c < fs; c += 2 * nw) {
            const bool has2 = (c + 1 < fs);
            const int cc0 = c - p0, cc1 = has2 ? cc0 + 1 : cc0;
            double *F0 = F + (size_t)c * ld;
            double *F1 = F + (size_t)(c + 1) * ld;
            for (int r0 = c + lane; r0 <= fs; r0 += 128) {
                double v0[4], v1[4];
                int rr[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int r = r0 + 32 * j;
                    const bool ok0 = r <= fs, ok1 = has2 && ok0 && r > c;
                    rr[j] = ok0 ? r - p0 : cc0;      // idle pieces read a valid panel entry, result dropped
                    v0[j] = ok0 ? F0[r] : 0.0;
                    v1[j] = ok1 ? F1[r] : 0.0;
                }
                for (int q = 0; q < pc; q++) {
                    const double a0 = -P[q * ldp + cc0], a1 = -P[q * ldp + cc1];
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const double x = P[q * ldp + rr[j]];
                        v0[j] = fma(x, a0, v0[j]);
                        v1[j] = fma(x, a1, v1[j]);
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int r = r0 + 32 * j;
                    if (r <= fs) F0[r] = v0[j];
                    if (has2 && r <= fs && r > c) F1[r] = v1[j];
                }
            }
        }
        __syncthreads();
    }
    if (done) front_signal(done + sid);
}

// top-down: x_J = L11^-T (y_J - L21^T x_struct); xperm holds the solution in elimination order.
// L11 is solved in column blocks of <= BS_PC staged in shared memory (the per-column chain of a
// triangular solve is latency-bound: shared memory instead of L2 for every step).
constexpr int BS_PC = 48;
constexpr int BS_TS = BS_PC + 1;      // padded stride: a lane walks a ROW of the column-major triangle
// PRE = true (big fronts): y_J - L21^T x_struct was formed by k_bf_back_gemv (many CTAs) and waits in
// the node's own slice of xperm; only the triangular solve with L11 is left for this CTA.
template <bool PRE>
__global__ void __launch_bounds__(512)
k_front_backsolve(const int *__restrict__ levelSn, const SnodeDesc *__restrict__ sn,
                  const int *__restrict__ structIdx, const double *__restrict__ fronts,
                  double *__restrict__ xperm, int *__restrict__ done, int *__restrict__ errflag)
{
    // single-launch mode (done != nullptr): the list is walked from its END (parents first)
    const int sid = done ? levelSn[gridDim.x - 1 - blockIdx.x] : levelSn[blockIdx.x];
    const SnodeDesc d = sn[sid];
    if (done) {
        if (threadIdx.x == 0 && d.parent >= 0) { front_wait(done + d.parent, errflag); __threadfence(); }
        __syncthreads();
    }
    const double *F = fronts + d.frontOff;
    const int fs = 6 * (d.ncols + d.nstruct), ld = fs + 1, nc = 6 * d.ncols, us = PRE ? 0 : 6 * d.nstruct;
    extern __shared__ double sh[];
    double *xs = sh;            // us entries: solution at the struct rows
    double *t = sh + us;        // nc entries
    double *T = t + nc;         // [BS_PC][BS_PC] triangle of the current column block (column-major)
    const int tid = threadIdx.x, nt = blockDim.x, warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
    double *xj = xperm + 6 * (size_t)d.poseOff;
    for (int i = tid; i < us; i += nt) xs[i] = __ldcg(xj + 6 * structIdx[d.structOff + i / 6] + (i % 6));
    if (PRE)
        for (int c = tid; c < nc; c += nt) t[c] = xj[6 * (size_t)d.first + c];
    __syncthreads();
    // four columns per warp and pass: their loads are independent, the L2 latency is paid once
    for (int cb = 4 * warp; cb < nc && !PRE; cb += 4 * nw) {
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        for (int r = lane; r < us; r += 32) {
            const double x = xs[r];
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (cb + j < nc) acc[j] = fma(F[(size_t)(cb + j) * ld + nc + r], x, acc[j]);
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            double a = sm::warp_sum(acc[j]);
            if (lane == 0 && cb + j < nc) t[cb + j] = F[(size_t)(cb + j) * ld + fs] - a;
        }
    }
    __syncthreads();
    for (int c1 = nc; c1 > 0; c1 -= BS_PC) {
        const int c0 = max(0, c1 - BS_PC), pb = c1 - c0;
        // contributions of the already solved columns [c1, nc)
        if (c1 < nc)
            for (int cb = c0 + 4 * warp; cb < c1; cb += 4 * nw) {
                double acc[4] = {0.0, 0.0, 0.0, 0.0};
                for (int r = c1 + lane; r < nc; r += 32) {
                    const double x = t[r];
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        if (cb + j < c1) acc[j] = fma(F[(size_t)(cb + j) * ld + r], x, acc[j]);
                }
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    double a = sm::warp_sum(acc[j]);
                    if (lane == 0 && cb + j < c1) t[cb + j] -= a;
                }
            }
        for (int e = tid; e < pb * pb; e += nt) {
            int c = e / pb, r = e - c * pb;
            T[c * BS_TS + r] = (r >= c) ? F[(size_t)(c0 + c) * ld + c0 + r] : 0.0;
        }
        __syncthreads();
        if (warp == 0) {
            // column-oriented back substitution: lane owns columns lane, lane+32; once x_r is known
            // every lane removes it from its own right-hand sides (no reduction on the chain)
            double tv[2], idg[2];
#pragma unroll
            for (int u = 0; u < 2; u++) {
                int c = lane + 32 * u;
                tv[u] = c < pb ? t[c0 + c] : 0.0;
                idg[u] = c < pb ? 1.0 / T[c * BS_TS + c] : 0.0;
            }
            for (int r = pb - 1; r >= 0; r--) {
                double mine = (r >= 32) ? tv[1] * idg[1] : tv[0] * idg[0];
                double xr = __shfl_sync(0xffffffffu, mine, r & 31);
                if (lane == (r & 31)) { if (r >= 32) tv[1] = xr; else tv[0] = xr; }
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    int c = lane + 32 * u;
                    if (c < r) tv[u] = fma(-T[c * BS_TS + r], xr, tv[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < 2; u++) {
                int c = lane + 32 * u;
                if (c < pb) t[c0 + c] = tv[u];
            }
        }
        __syncthreads();
    }
    for (int c = tid; c < nc; c += nt) xj[6 * (size_t)d.first + c] = t[c];
    if (done) front_signal(done + sid);
}

} // namespace
#include "chol_big.cuh"
namespace {
__global__ void k_unpermute(DMap *__restrict__ J, const int *__restrict__ posePre, int K, int totP,
                            const int *__restrict__ perm, const double *__restrict__ xperm)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    int gp = g / 6, q = g - 6 * gp;
    if (gp >= totP) return;
    int k = seg_find(posePre, K, gp);
    J[k].poseVal[6 * (size_t)perm[gp] + q] = xperm[6 * (size_t)gp + q];
}

// a13: x_f = V^-1 (eF_f - sum_p W_pf^T x_p)      (LinearSFMImp.cpp:2980-3020)
__global__ void k_backsub(DMap *__restrict__ J, const int *__restrict__ featPre, int K, int totF,
                          const double *__restrict__ Vinv, const double *__restrict__ eF)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= totF) return;
    int k = seg_find(featPre, K, g);
    int f = g - featPre[k];
    const DMap &M = J[k];
    double acc[3] = {0, 0, 0};
    for (int j = M.wPtr[f]; j < M.wPtr[f + 1]; j++) {
        double W[18], xp[6], t[3];
        // 16-byte loads: W blocks (144 bytes) and pose rows (48 bytes) are 16-byte aligned in the arena
        const double2 *w2 = reinterpret_cast<const double2 *>(M.W + 18 * (size_t)j);
        const double2 *x2 = reinterpret_cast<const double2 *>(M.poseVal + 6 * (size_t)M.photo[j]);
#pragma unroll
        for (int q = 0; q < 9; q++) { const double2 v = w2[q]; W[2 * q] = v.x; W[2 * q + 1] = v.y; }
#pragma unroll
        for (int q = 0; q < 3; q++) { const double2 v = x2[q]; xp[2 * q] = v.x; xp[2 * q + 1] = v.y; }
        sm::mtm<3, 6, 1>(W, xp, t);
        acc[0] += t[0]; acc[1] += t[1]; acc[2] += t[2];
    }
    double r[3] = {eF[3 * (size_t)g] - acc[0], eF[3 * (size_t)g + 1] - acc[1], eF[3 * (size_t)g + 2] - acc[2]};
    double Vi[9], x[3];
    sm::load<9>(Vinv + 9 * (size_t)g, Vi);
    sm::mm<3, 3, 1>(Vi, r, x);
    double *o = M.featVal + 3 * (size_t)f;
    o[0] = x[0]; o[1] = x[1]; o[2] = x[2];
}

void exclusive_scan(Context &ctx, const int *in, int *out, int n)
{
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, in, out, n, ctx.stream);
    DevBuf<char> tmp(tmp_bytes, ctx.stream);
    cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, in, out, n, ctx.stream);
}

} // namespace

void solve_stereo_batch(Context &ctx, const OpMaps &J, const double *eP, const double *eF,
                        SolveDebug *dbg, const MonoGauge *gauge, const SolveExtra *ex)
{
    const int K = J.K;
    cudaStream_t s = ctx.stream;
    const int TB = 256;
    int nl = 0;

    // ---- a8: pattern ----
    ctx.begin("solve.pattern");
    // feature chunks shared by the pattern and the Schur kernels
    std::vector<FeatChunk> chunks;
    int maxWords = 1;
    for (int k = 0; k < K; k++) maxWords = std::max(maxWords, (J.h[k].m + 31) / 32);
    // chunk boundaries per join: start indices ending with n (seeded by the caller or SCH_FCHUNK wide)
    std::vector<std::vector<int>> starts(K);
    bool customChunks = false;
    for (int k = 0; k < K; k++) {
        const int n = J.h[k].n;
        const std::vector<int> *seed = (ex && ex->chunkSeed && (*ex->chunkSeed)[k]) ? (*ex->chunkSeed)[k].get() : nullptr;
        if (seed && !seed->empty() && seed->front() == 0 && seed->back() == n) { starts[k] = *seed; customChunks = true; }
        else {
            for (int f0 = 0; f0 < n; f0 += SCH_FCHUNK) starts[k].push_back(f0);
            starts[k].push_back(n);
        }
    }
    int nChunks = 0;
    DevBuf<FeatChunk> dChunks;
    int maxNposes = 1 << 30;             // max distinct poses of any chunk (measured by k_pat_chunk)
    DevBuf<int> dMaxNp(4, s);            // [0] max #poses per chunk, [2..3] pair-feature count (stage timing only)
    DevBuf<int> chunkInfo, blkInfo((size_t)std::max(J.totW, 1), s);
    // per chunk: pose bitmap + popcount prefix (the E gather's index) and the chunks of every join
    const int bitsStride = maxWords;
    DevBuf<unsigned> patBits;
    std::vector<int> chunkPre;
    DevBuf<int> dChunkPre(K + 1, s);
    auto build_chunks = [&]() {
        chunks.clear();
        chunkPre.assign(K + 1, 0);
        for (int k = 0; k < K; k++) {
            for (size_t i = 0; i + 1 < starts[k].size(); i++)
                if (starts[k][i + 1] > starts[k][i]) chunks.push_back({k, starts[k][i], starts[k][i + 1]});
            chunkPre[k + 1] = (int)chunks.size();
        }
        nChunks = (int)chunks.size();
        dChunks.alloc(nChunks, s);
        dChunks.upload(chunks);
        chunkInfo.alloc((size_t)CHUNK_INFO_INTS * std::max(nChunks, 1), s);
        patBits.alloc(2 * (size_t)bitsStride * std::max(nChunks, 1), s);
        dChunkPre.upload(chunkPre);
    };
    build_chunks();
    // test hook: LSFM_FORCE_OVERFLOW=1 sends every chunk with > 4 poses down the overflow paths
    static const bool force_ovf = getenv("LSFM_FORCE_OVERFLOW") != nullptr;
    const int pat_cmax_used = force_ovf ? 4 : PAT_CMAX;
    // Pose-pair bitmap per join (upper triangle, (m+31)/32 words per row): every chunk ORs its
    // distinct pairs in, the U blocks (and the mono gauge diagonal) follow; row popcounts + one scan
    // give the block-CRS row pointers and the keys come out already sorted -- no key sort, no
    // intermediate count on the host.
    std::vector<int> bmOff(K + 1, 0);
    long long capBlocks = 0;
    for (int k = 0; k < K; k++) {
        long long mk = J.h[k].m;
        long long wordsK = mk * ((mk + 31) / 32);
        if (bmOff[k] + wordsK > 0x7fffffffLL) throw LsfmError(LSFM_ERR_ARG, "pattern bitmap too large");
        bmOff[k + 1] = bmOff[k] + (int)wordsK;
        capBlocks += mk * (mk + 1) / 2;
    }
    DevBuf<int> dBmOff(K + 1, s);
    dBmOff.upload(bmOff);
    DevBuf<unsigned> bm((size_t)std::max(bmOff[K], 1), s);
    bm.zero();
    DevBuf<int> poseCR(2 * (size_t)std::max(J.totPose, 1), s);   // per pose: first / last chunk that records it
    poseCR.zero();
    auto pattern_chunk_pass = [&]() {
        if (nChunks == 0) return;
        size_t shb = sizeof(int) * (2 * (size_t)maxWords + 16 + PAT_CMAX + 1 + 4 + 32);
        if (shb > 48 * 1024)
            CUDA_CHECK(cudaFuncSetAttribute(k_pat_chunk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shb));
        dMaxNp.zero();
        poseCR.zero();
        k_pat_chunk<<<nChunks, PAT_THREADS, shb, s>>>(J.d.p, dChunks.p, bm.p, dBmOff.p, dMaxNp.p, pat_cmax_used,
                                                      chunkInfo.p, blkInfo.p, J.dWPre.p, patBits.p, bitsStride,
                                                      ctx.timing ? (unsigned long long *)(dMaxNp.p + 2) : nullptr,
                                                      J.dPosePre.p, poseCR.p); nl++;
    };
    pattern_chunk_pass();
    if (gauge) {   // mono: the zero pose has no block at all after the join; keep every diagonal
        k_pat_diag<<<ceil_div(J.totPose, TB), TB, 0, s>>>(J.d.p, J.dPosePre.p, K, J.totPose, bm.p, dBmOff.p); nl++;
    }
    if (J.totU > 0) { k_pat_u<<<ceil_div(J.totU, TB), TB, 0, s>>>(J.d.p, J.dUPre.p, K, J.totU, bm.p, dBmOff.p); nl++; }
    DevBuf<int> rowCnt(J.totPose + 1, s), rowPtr(J.totPose + 1, s);
    k_bm_rowcount<<<ceil_div((J.totPose + 1) * 32, TB), TB, 0, s>>>(bm.p, dBmOff.p, J.d.p, J.dPosePre.p, K, J.totPose, rowCnt.p); nl++;
    exclusive_scan(ctx, rowCnt.p, rowPtr.p, J.totPose + 1); nl += 2;
    // the number of blocks is not known on the host yet: keys are produced / copied up to a bound
    // (exact for small joins, generous for large ones); the rare overshoot takes a second copy
    int keyCap = (int)std::min<long long>(capBlocks, 64LL * J.totPose + 4096);
    if (force_ovf) keyCap = std::min(keyCap, 16);     // test hook: exercise the second-copy path
    DevBuf<u64> keys((size_t)std::max(keyCap, 1), s);
    k_bm_emit<<<ceil_div(J.totPose * 32, TB), TB, 0, s>>>(bm.p, dBmOff.p, J.d.p, J.dPosePre.p, K, J.totPose, rowPtr.p, keys.p, keyCap); nl++;
    // one synchronisation for everything the host needs: keys, row pointers (#blocks = last entry)
    int nuis = 0;
    char *pin = ctx.pinDown.need(sizeof(u64) * ((size_t)keyCap + 1) + sizeof(int) * ((size_t)J.totPose + 8));
    u64 *hKeys = (u64 *)pin;
    int *hRowPtr = (int *)(pin + sizeof(u64) * ((size_t)keyCap + 1));
    int *hMaxNp = hRowPtr + J.totPose + 1;
    hMaxNp[0] = hMaxNp[2] = hMaxNp[3] = 0;
    if (nChunks > 0) CUDA_CHECK(cudaMemcpyAsync(hMaxNp, dMaxNp.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, s));
    if (keyCap) CUDA_CHECK(cudaMemcpyAsync(hKeys, keys.p, sizeof(u64) * keyCap, cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaMemcpyAsync(hRowPtr, rowPtr.p, sizeof(int) * (J.totPose + 1), cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    ctx.idle_begin();
    nuis = hRowPtr[J.totPose];
    if (nChunks > 0) maxNposes = *hMaxNp;
    // Chunks that see more than PAT_CMAX distinct poses take the slow thread-per-block paths (with FP64
    // atomics) in the pattern, Schur and Transform kernels.  Split THOSE chunks in four (by features, down to 8
    // features) and repeat the chunk pass -- the pose-pair bitmap only gains bits it already has, chunkInfo /
    // blkInfo / patBits are rewritten.  The limit is one below the kernels' capacity: the Transform of the
    // joined map adds one pose (the new origin) to every feature.  What still overflows at 8 features (a
    // landmark seen by more than 30 poses) stays on the slow path.
    static const bool no_split = getenv("LSFM_NO_CHUNK_SPLIT") != nullptr;
    constexpr int SPLIT_LIMIT = PAT_CMAX - 1;
    for (int round = 0; round < 3 && nChunks > 0 && maxNposes > SPLIT_LIMIT && !force_ovf && !no_split; round++) {
        std::vector<int> np((size_t)nChunks);
        CUDA_CHECK(cudaMemcpy2DAsync(np.data(), sizeof(int), chunkInfo.p + 31, sizeof(int) * CHUNK_INFO_INTS, sizeof(int),
                                     (size_t)nChunks, cudaMemcpyDeviceToHost, s));
        CUDA_CHECK(cudaStreamSynchronize(s));
        bool any = false;
        std::vector<std::vector<int>> ns(K);
        for (int c = 0; c < nChunks; c++) {
            const FeatChunk &ch = chunks[c];
            const int w = ch.f1 - ch.f0;
            if (np[c] > SPLIT_LIMIT && w > 8) {
                const int part = std::max(8, (w + 3) / 4);
                for (int f = ch.f0; f < ch.f1; f += part) ns[ch.k].push_back(f);
                any = true;
            } else ns[ch.k].push_back(ch.f0);
        }
        if (!any) break;
        for (int k = 0; k < K; k++) { ns[k].push_back(J.h[k].n); starts[k].swap(ns[k]); }
        customChunks = true;
        build_chunks();
        pattern_chunk_pass();
        CUDA_CHECK(cudaMemcpyAsync(hMaxNp, dMaxNp.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, s));
        CUDA_CHECK(cudaStreamSynchronize(s));
        maxNposes = *hMaxNp;
    }
    if (getenv("LSFM_DEBUG") && nChunks > 0 && maxNposes > PAT_CMAX) {
        std::vector<int> np((size_t)nChunks);
        CUDA_CHECK(cudaMemcpy2DAsync(np.data(), sizeof(int), chunkInfo.p + 31, sizeof(int) * CHUNK_INFO_INTS, sizeof(int),
                                     (size_t)nChunks, cudaMemcpyDeviceToHost, s));
        CUDA_CHECK(cudaStreamSynchronize(s));
        int nov = 0, f8 = 0;
        for (int c = 0; c < nChunks; c++)
            if (np[c] > PAT_CMAX) { nov++; f8 += (chunks[c].f1 - chunks[c].f0 <= 8); }
        fprintf(stderr, "    [chunks] K=%d nChunks=%d maxNposes=%d: %d chunks stay on the slow path (%d of them <= 8 features)\n",
                K, nChunks, maxNposes, nov, f8);
    }
    if (ex && ex->chunksOut) {
        ex->chunksOut->assign(K, nullptr);
        if (customChunks)
            for (int k = 0; k < K; k++)
                if ((int)starts[k].size() != (J.h[k].n + SCH_FCHUNK - 1) / SCH_FCHUNK + 1)
                    (*ex->chunksOut)[k] = std::make_shared<const std::vector<int>>(starts[k]);
    }
    // FP64 work of the Schur stage: 216 flop per (pose pair, feature) product, 108 per block for W V^-1,
    // 36 per block for the reduced right-hand side
    const double schurFlops = 216.0 * (double)(((unsigned long long)(unsigned)hMaxNp[3] << 32) | (unsigned)hMaxNp[2]) +
                              144.0 * J.totW;
    if (nuis > keyCap) {                  // bound overshot: emit and fetch again with the exact size
        keys.alloc((size_t)nuis, s);
        k_bm_emit<<<ceil_div(J.totPose * 32, TB), TB, 0, s>>>(bm.p, dBmOff.p, J.d.p, J.dPosePre.p, K, J.totPose, rowPtr.p, keys.p, nuis); nl++;
        pin = ctx.pinDown.need(sizeof(u64) * ((size_t)nuis + 1) + sizeof(int) * ((size_t)J.totPose + 4));
        hKeys = (u64 *)pin;
        hRowPtr = (int *)(pin + sizeof(u64) * ((size_t)nuis + 1));
        CUDA_CHECK(cudaMemcpyAsync(hKeys, keys.p, sizeof(u64) * nuis, cudaMemcpyDeviceToHost, s));
        CUDA_CHECK(cudaMemcpyAsync(hRowPtr, rowPtr.p, sizeof(int) * (J.totPose + 1), cudaMemcpyDeviceToHost, s));
        CUDA_CHECK(cudaStreamSynchronize(s));
    }
    KERNEL_CHECK();
    ctx.end(8.0 * nuis + 4.0 * bmOff[K], 0.0, nl);
    nl = 0;

    // The pattern is on its way to the host symbolic phase; work the caller deferred until now
    // (the join's value copy) is queued here so that it runs while the host analyses the pattern.
    ctx.idle_end(2);
    if (ex && ex->after_pattern) ex->after_pattern();

    // ---- a9-a10: V^-1, S, E ----
    ctx.begin("solve.schur_prep");
    DevBuf<double> Vinv(9 * (size_t)J.totFeat, s), S(36 * (size_t)nuis, s), E(6 * (size_t)J.totPose, s);
    DevBuf<double> dvec(6 * (size_t)std::max(J.totFeat, 1), s);
    const double *xhat = ex ? ex->xhat : nullptr;
    const int *split = ex ? ex->split : nullptr;
    S.zero();
    CUDA_CHECK(cudaMemcpyAsync(E.p, eP, sizeof(double) * 6 * (size_t)J.totPose, cudaMemcpyDeviceToDevice, s));
    if (J.totFeat > 0) { k_vinv<<<ceil_div(J.totFeat, TB), TB, 0, s>>>(J.d.p, J.dFeatPre.p, K, J.totFeat, eF, xhat, Vinv.p, dvec.p); nl++; }
    if (J.totU > 0) {
        k_s_from_u<<<ceil_div(J.totU, TB), TB, 0, s>>>(J.d.p, J.dUPre.p, J.dPosePre.p, K, J.totU, keys.p, rowPtr.p, S.p); nl++;
    }
    // fixed-point accumulators of S (order-independent integer atomics) and their per-row exponents
    DevBuf<long long> Sfx(36 * (size_t)std::max(nuis, 1), s);
    DevBuf<int> sexp(6 * (size_t)std::max(J.totPose, 1), s);
    DevBuf<double> Erec(6 * 32 * (size_t)std::max(nChunks, 1), s);
    CUDA_CHECK(cudaMemsetAsync(Sfx.p, 0, sizeof(long long) * 36 * (size_t)std::max(nuis, 1), s));
    k_row_exp<<<ceil_div(6ll * J.totPose, TB), TB, 0, s>>>(keys.p, rowPtr.p, J.d.p, J.dPosePre.p, K, J.totPose, S.p, sexp.p); nl++;
    // slow path (chunks with more than PAT_CMAX poses): fixed-point E accumulators + the pre-pass that sizes them
    const bool anySlow = nChunks > 0 && (maxNposes > PAT_CMAX || maxNposes > pat_cmax_used);
    DevBuf<int> eexp(anySlow ? 6 * (size_t)J.totPose : 1, s), ecnt(anySlow ? (size_t)J.totPose : 1, s);
    DevBuf<long long> Efx(anySlow ? 6 * (size_t)J.totPose : 1, s);
    if (anySlow) {
        CUDA_CHECK(cudaMemsetAsync(eexp.p, 0x80, sizeof(int) * 6 * (size_t)J.totPose, s));      // 0x80808080 < -2000
        CUDA_CHECK(cudaMemsetAsync(ecnt.p, 0, sizeof(int) * (size_t)J.totPose, s));
        CUDA_CHECK(cudaMemsetAsync(Efx.p, 0, sizeof(long long) * 6 * (size_t)J.totPose, s));
        k_schur_slow_pre<<<nChunks, 128, 0, s>>>(J.d.p, dChunks.p, chunkInfo.p, pat_cmax_used, J.dFeatPre.p, J.dPosePre.p,
                                                 dvec.p, split, eexp.p, ecnt.p); nl++;
    }
    const SlowFx slowFx{sexp.p, Sfx.p, eexp.p, ecnt.p, Efx.p};
    ctx.end(72.0 * J.totFeat * 2 + 576.0 * J.totU, 0.0, nl);
    nl = 0;
    ctx.begin("solve.schur");
    if (J.totW > 0) {
        static const bool use_v1 = getenv("LSFM_SCHUR_V1") != nullptr;
        if (use_v1) {
            k_schur<<<ceil_div(J.totW, 128), 128, 0, s>>>(J.d.p, J.dWPre.p, J.dFeatPre.p, J.dPosePre.p, K, J.totW,
                                                         Vinv.p, dvec.p, split, keys.p, rowPtr.p, S.p, E.p); nl++;
        } else {
            // pipelined kernel; instantiation chosen from the measured max #distinct poses per chunk
            auto launch = [&](auto kern, size_t shb, int threads) {
                if (shb > 48 * 1024) {
                    CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shb));
                    CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                                    (int)cudaSharedmemCarveoutMaxShared));
                }
                kern<<<nChunks, threads, shb, s>>>(J.d.p, dChunks.p, chunkInfo.p, blkInfo.p, pat_cmax_used, J.dWPre.p,
                                                  J.dFeatPre.p, J.dPosePre.p, Vinv.p, dvec.p, split, keys.p, rowPtr.p, S.p, E.p,
                                                  sexp.p, Sfx.p, Erec.p, slowFx);
            };
            static const bool force_ovf2 = getenv("LSFM_FORCE_OVERFLOW") != nullptr;
            // LSFM_SCHUR_DENSE=1: the chunk's dense pose pairs as one DMMA contraction (schur_dense.cuh).
            // Parity-green, but on B200 the DMMA and DFMA peaks are equal (37 vs 33 TF/s measured,
            // profiles/r2_dmma_ubench.txt) and the kernel is bound by its per-batch barriers, not by the
            // FP64 pipe (profiles/r2_ncu_schur_dense.md): it is slower than the register-tiled scalar
            // kernel at every tree level, so the latter stays the default.
            static const bool use_dense = getenv("LSFM_SCHUR_DENSE") != nullptr;
            static const bool use_pipe = getenv("LSFM_SCHUR_PIPE") != nullptr;    // previous per-lane-list kernel
            static const int lock_t = getenv("LSFM_SCHUR_LOCK_T") ? atoi(getenv("LSFM_SCHUR_LOCK_T")) : 0;
            // smallest per-level maximum of poses per chunk from which the lock-step kernel is used
            static const int lock_min = getenv("LSFM_SCHUR_LOCK_MIN") ? std::max(2, atoi(getenv("LSFM_SCHUR_LOCK_MIN"))) : 3;   // measured: level 2 0.35 -> 0.29 ms, level 1 0.26 -> 0.25, level 0 (two poses) stays with k_schur_pipe
            if (use_dense) {
                if (maxNposes <= 16 || force_ovf2)
                    launch(schur_dense::k_schur_dense<16, 160, 256, 2, 8>, schur_dense::Layout<16, 160, 256>::bytes(), 256);
                else
                    launch(schur_dense::k_schur_dense<31, 248, 512, 4, 16>, schur_dense::Layout<31, 248, 512>::bytes(), 512);
            } else if (maxNposes <= lock_min - 1 || force_ovf2)
                launch(schur_pipe::k_schur_pipe<8, 64, 32, 128, 1>, schur_pipe::Layout<8, 64, 32>::bytes(), 128);
            else if (!use_pipe)
                // 3-31 poses per chunk: lock-step kernel (128 threads x 4 CTAs/SM up to 17 poses, 256 x 2 above)
                // (LSFM_SCHUR_LOCK_T: 0 = pick by the level's maximum, 128 / 256 = force one instantiation)
                if (lock_t == 128 || (lock_t == 0 && maxNposes <= 17))
                    launch(schur_lock::k_schur_lock<96, 32, 128, 4>, schur_lock::Layout<96, 32>::bytes(), 128);
                else
                    launch(schur_lock::k_schur_lock<224, 32, 256, 2>, schur_lock::Layout<224, 32>::bytes(), 256);
            else if (maxNposes <= 16)
                launch(schur_pipe::k_schur_pipe<16, 128, 32, 256, 1>, schur_pipe::Layout<16, 128, 32>::bytes(), 256);
            else
                launch(schur_pipe::k_schur_pipe<31, 248, 32, 512, 1>, schur_pipe::Layout<31, 248, 32>::bytes(), 512);
            nl++;
            KERNEL_CHECK();
            ctx.end(152.0 * J.totW + 96.0 * J.totFeat + 288.0 * nuis + 48.0 * J.totPose, schurFlops, nl);
            nl = 0;
            ctx.begin("solve.schur_fin");
            // S -= the fixed-point sums; E_p -= the chunks' shares, gathered per pose in a fixed order
            k_s_convert<<<ceil_div(36ll * nuis, TB), TB, 0, s>>>(keys.p, nuis, J.dPosePre.p, sexp.p, Sfx.p, S.p); nl++;
            if (anySlow) { k_e_convert<<<ceil_div(6ll * J.totPose, TB), TB, 0, s>>>(6 * J.totPose, eexp.p, ecnt.p, Efx.p, E.p); nl++; }
            k_e_gather<<<ceil_div(32ll * J.totPose, 128), 128, 0, s>>>(J.dPosePre.p, K, J.totPose, dChunkPre.p, patBits.p,
                                                                    bitsStride, Erec.p, E.p, poseCR.p); nl++;
        }
    }
    KERNEL_CHECK();
    // (algorithmic bytes of the Schur kernel: every W block + its feature's V^-1 and eF read once,
    // S and E written once)
    ctx.end(576.0 * nuis, 0.0, nl);
    nl = 0;
    if (gauge && nuis > 0) {
        k_mono_gauge<<<ceil_div(nuis, TB), TB, 0, s>>>(keys.p, nuis, gauge->refPose, gauge->fixScalar, S.p);
        k_mono_gauge_rhs<<<ceil_div(K, TB), TB, 0, s>>>(J.dPosePre.p, K, gauge->refPose, gauge->fixScalar, E.p);
    }
    ctx.begin("solve.symbolic");

    // ---- symbolic on the host while the GPU accumulates S ----
    std::vector<int> mvec(K), sOff(K + 1);
    for (int k = 0; k < K; k++) { mvec[k] = J.h[k].m; sOff[k] = hRowPtr[J.posePre[k]]; }
    sOff[K] = nuis;
    BatchSymbolic sym;
    static const bool dbg_time = getenv("LSFM_DEBUG") != nullptr;
    auto tsym0 = std::chrono::steady_clock::now();
    try {
        build_symbolic(K, mvec, J.posePre, hKeys, (size_t)nuis, sOff, sym, (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency())));
    } catch (const std::exception &e) { throw LsfmError(LSFM_ERR_ARG, std::string("symbolic: ") + e.what()); }
    if (dbg_time) {
        double hms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tsym0).count();
        bool schur_done = cudaStreamQuery(s) == cudaSuccess;
        fprintf(stderr, "    [symbolic] K=%d totPose=%d nuis=%d host %.3f ms, Schur kernel %s when it finished\n", K,
                J.totPose, nuis, hms, schur_done ? "ALREADY DONE (GPU idle)" : "still running");
    }
    ctx.end(8.0 * nuis, 0.0, nl);
    nl = 0;

    // ---- numeric ----
    ctx.begin("solve.cholesky");
    int ns = (int)sym.sn.size();
    DevBuf<SnodeDesc> dSn(ns, s); dSn.upload(sym.sn);
    DevBuf<int> dStruct(sym.structIdx.size(), s); dStruct.upload(sym.structIdx);
    DevBuf<int> dRel(sym.relIdx.size(), s); dRel.upload(sym.relIdx);
    DevBuf<int> dChild(sym.childIdx.size(), s); dChild.upload(sym.childIdx);
    DevBuf<int> dLevelSn(sym.levelSn.size(), s);        // filled below (small fronts first, then the big ones)
    DevBuf<SlotMap> dSlot(sym.slot.size(), s); dSlot.upload(sym.slot);
    DevBuf<int> dPoseSn(J.totPose, s); dPoseSn.upload(sym.poseSn);
    DevBuf<int> dPoseLcol(J.totPose, s); dPoseLcol.upload(sym.poseLcol);
    DevBuf<int> dPerm(J.totPose, s); dPerm.upload(sym.perm);
    DevBuf<double> fronts((size_t)sym.frontDoubles, s);
    fronts.zero();
    DevBuf<double> xperm(6 * (size_t)J.totPose, s);
    struct { int *p; } err = {ctx.error_flag()};
    if (nuis > 0) { k_front_assemble<<<ceil_div(36ll * nuis, TB), TB, 0, s>>>(dSlot.p, nuis, dSn.p, S.p, fronts.p); nl++; }
    k_front_rhs<<<ceil_div(6ll * J.totPose, TB), TB, 0, s>>>(dPoseSn.p, dPoseLcol.p, J.totPose, dSn.p, E.p, fronts.p); nl++;
    int nLevels = (int)sym.levelPtr.size() - 1;
    // Fronts of bigfront::BF_MIN_FS or more scalar rows (loop closures, dense overlap) are factored by the
    // whole GPU (chol_big.cuh: multi-CTA panels + DMMA trailing updates); the others one CTA per front.
    // Per level: the small fronts first in the level's list, then the big ones.
    static const bool no_big = getenv("LSFM_NO_BIGFRONT") != nullptr;      // test hook: one CTA per front always
    // (test hook: LSFM_BIGFRONT_MIN_FS lowers the size from which a front takes the multi-CTA path)
    static const int big_min_fs = getenv("LSFM_BIGFRONT_MIN_FS") ? std::max(48, atoi(getenv("LSFM_BIGFRONT_MIN_FS")))
                                                                 : bigfront::BF_MIN_FS;
    std::vector<int> lvlList(sym.levelSn.size()), nSmall(nLevels, 0);
    int maxSmallFdim = 1;
    for (int l = 0; l < nLevels; l++) {
        int a = sym.levelPtr[l], b = sym.levelPtr[l + 1], w = a;
        auto is_big = [&](int sid) {
            const SnodeDesc &d = sym.sn[sid];
            return !no_big && 6 * (d.ncols + d.nstruct) >= big_min_fs;
        };
        for (int i = a; i < b; i++) if (!is_big(sym.levelSn[i])) lvlList[w++] = sym.levelSn[i];
        nSmall[l] = w - a;
        for (int i = a; i < a + nSmall[l]; i++) {
            const SnodeDesc &d = sym.sn[lvlList[i]];
            maxSmallFdim = std::max(maxSmallFdim, d.ncols + d.nstruct);
        }
        for (int i = a; i < b; i++) if (is_big(sym.levelSn[i])) lvlList[w++] = sym.levelSn[i];
    }
    dLevelSn.upload(lvlList);
    // panel width: 48 columns (8 pose blocks) unless the tallest small front of the batch needs a narrower
    // one; fronts whose panel would be narrower than two pose blocks keep it in global memory instead
    const size_t frows = 6 * (size_t)maxSmallFdim + 2;
    int pcMax = (int)std::min<size_t>(48, ((size_t)200 * 1024 / (8 * frows)) / 6 * 6);
    static const bool force_gp = getenv("LSFM_FORCE_GLOBAL_PANEL") != nullptr;      // test hook
    const bool globalPanel = pcMax < 12 || force_gp;
    if (globalPanel) pcMax = 48;
    const size_t panelStride = (size_t)pcMax * frows;
    const size_t shf = globalPanel ? 0 : sizeof(double) * panelStride;
    int maxCnt = 1;
    for (int l = 0; l < nLevels; l++) maxCnt = std::max(maxCnt, nSmall[l]);
    DevBuf<double> panelG(globalPanel ? panelStride * (size_t)maxCnt : 1, s);
    if (shf > 48 * 1024)
        CUDA_CHECK(cudaFuncSetAttribute(k_front_factor<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shf));
    const size_t shSyrk = sizeof(double) * 2 * bigfront::BF_PC * bigfront::BF_LDS;
    const size_t shPanel = sizeof(double) * bigfront::BF_PC * (bigfront::BF_PC + bigfront::BF_RS + 1);
    bool syrkAttr = false;
    int maxBig = 0;
    for (int l = 0; l < nLevels; l++) maxBig = std::max(maxBig, sym.levelPtr[l + 1] - sym.levelPtr[l] - nSmall[l]);
    DevBuf<double> diagScratch((size_t)std::max(maxBig, 1) * bigfront::BF_PC * bigfront::BF_PC, s);
    // No big front anywhere (sequential scenes: fronts <= ~200 rows): ONE launch for the whole assembly tree of
    // every join of the batch instead of one per tree level -- a CTA per front in level order, children /
    // parents awaited through flags (front_wait / front_signal).  LSFM_CHOL_LEVELS=1: per-level launches.
    static const bool per_level = getenv("LSFM_CHOL_LEVELS") != nullptr;
    int totSmall = 0;
    for (int l = 0; l < nLevels; l++) totSmall += nSmall[l];
    const bool oneLaunch = !per_level && !globalPanel && totSmall == ns && ns > 0;
    DevBuf<int> doneF(oneLaunch ? 2 * (size_t)ns : 1, s);
    if (oneLaunch) {
        CUDA_CHECK(cudaMemsetAsync(doneF.p, 0, sizeof(int) * 2 * (size_t)ns, s));
        const int thr = (ns <= 8 * ctx.num_sms) ? 512 : 256;
        k_front_factor<false><<<ns, thr, shf, s>>>(dLevelSn.p, dSn.p, dChild.p, dRel.p, fronts.p, err.p, pcMax, nullptr, 0,
                                                   doneF.p);
        nl++;
    }
    for (int l = 0; l < nLevels && !oneLaunch; l++) {
        const int cnt = nSmall[l];
        const int nBig = sym.levelPtr[l + 1] - sym.levelPtr[l] - cnt;
        if (cnt > 0) {
            // few fronts on the level (the top of the assembly tree): twice the threads per front --
            // the SMs are idle anyway and the extend-add / trailing update scale with the warps
            const int thr = (cnt <= ctx.num_sms) ? 512 : 256;
            if (globalPanel)
                k_front_factor<true><<<cnt, thr, 0, s>>>(dLevelSn.p + sym.levelPtr[l], dSn.p, dChild.p, dRel.p, fronts.p, err.p,
                                                        pcMax, panelG.p, panelStride, nullptr);
            else
                k_front_factor<false><<<cnt, thr, shf, s>>>(dLevelSn.p + sym.levelPtr[l], dSn.p, dChild.p, dRel.p, fronts.p,
                                                           err.p, pcMax, nullptr, 0, nullptr);
            nl++;
        }
        if (nBig > 0) {
            const int *bigSn = dLevelSn.p + sym.levelPtr[l] + cnt;
            int maxFd = 0, maxNc = 0, maxFs = 0;
            bool anyChild = false;
            for (int i = 0; i < nBig; i++) {
                const SnodeDesc &d = sym.sn[lvlList[sym.levelPtr[l] + cnt + i]];
                maxFd = std::max(maxFd, d.ncols + d.nstruct);
                maxNc = std::max(maxNc, 6 * d.ncols);
                maxFs = std::max(maxFs, 6 * (d.ncols + d.nstruct));
                anyChild = anyChild || d.nchild > 0;
            }
            if (!syrkAttr) {
                CUDA_CHECK(cudaFuncSetAttribute(bigfront::k_bf_syrk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shSyrk));
                CUDA_CHECK(cudaFuncSetAttribute(bigfront::k_bf_panel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shPanel));
                syrkAttr = true;
            }
            if (anyChild) {
                bigfront::k_bf_extend<<<dim3(maxFd, nBig), 256, 0, s>>>(bigSn, dSn.p, dChild.p, dRel.p, fronts.p);
                nl++;
            }
            for (int p0 = 0; p0 < maxNc; p0 += bigfront::BF_PC) {
                const int below = maxFs + 1 - p0;          // upper bound of the rows below any front's panel
                bigfront::k_bf_panel<<<dim3(std::max(1, ceil_div(below, bigfront::BF_RS)), nBig), bigfront::BF_PT, shPanel, s>>>(
                    bigSn, dSn.p, fronts.p, err.p, p0, diagScratch.p);
                const int nT = ceil_div(below, bigfront::BF_T);
                bigfront::k_bf_syrk<<<dim3(nT * (nT + 1) / 2, nBig), 128, shSyrk, s>>>(bigSn, dSn.p, fronts.p, p0, diagScratch.p);
                nl += 2;
            }
        }
    }
    size_t shb = sizeof(double) * (6 * (size_t)(maxSmallFdim + 1) + BS_PC * BS_TS);
    if (shb > 220 * 1024)
        throw LsfmError(LSFM_ERR_ARG, "a Cholesky front of " + std::to_string(maxSmallFdim) +
                                          " pose blocks exceeds the back-solve's shared-memory capacity (~4400 blocks)");
    if (shb > 48 * 1024)
        CUDA_CHECK(cudaFuncSetAttribute(k_front_backsolve<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shb));
    if (oneLaunch) {
        const int thr = (ns <= 8 * ctx.num_sms) ? 512 : 256;
        k_front_backsolve<false><<<ns, thr, shb, s>>>(dLevelSn.p, dSn.p, dStruct.p, fronts.p, xperm.p, doneF.p + ns, err.p);
        nl++;
    }
    for (int l = nLevels - 1; l >= 0 && !oneLaunch; l--) {
        const int cnt = nSmall[l];
        const int nBig = sym.levelPtr[l + 1] - sym.levelPtr[l] - cnt;
        if (cnt > 0) {
            const int thr = (cnt <= ctx.num_sms) ? 512 : 256;
            k_front_backsolve<false><<<cnt, thr, shb, s>>>(dLevelSn.p + sym.levelPtr[l], dSn.p, dStruct.p, fronts.p, xperm.p,
                                                           nullptr, err.p);
            nl++;
        }
        if (nBig > 0) {
            const int *bigSn = dLevelSn.p + sym.levelPtr[l] + cnt;
            int maxNc = 0;
            for (int i = 0; i < nBig; i++) maxNc = std::max(maxNc, 6 * sym.sn[lvlList[sym.levelPtr[l] + cnt + i]].ncols);
            bigfront::k_bf_back_gemv<<<dim3(ceil_div(maxNc, 8), nBig), 256, 0, s>>>(bigSn, dSn.p, dStruct.p, fronts.p, xperm.p);
            const size_t shbig = sizeof(double) * ((size_t)maxNc + BS_PC * BS_TS);
            k_front_backsolve<true><<<nBig, 512, shbig, s>>>(bigSn, dSn.p, dStruct.p, fronts.p, xperm.p, nullptr, err.p);
            nl += 2;
        }
    }
    k_unpermute<<<ceil_div(6ll * J.totPose, TB), TB, 0, s>>>(J.d.p, J.dPosePre.p, K, J.totPose, dPerm.p, xperm.p); nl++;
    // the not-SPD flag is sticky in the context and checked once per API call (no sync here)
    KERNEL_CHECK();
    ctx.end(0.0, sym.flops, nl);
    nl = 0;

    // ---- a13 ----
    ctx.begin("solve.backsub");
    if (J.totFeat > 0) { k_backsub<<<ceil_div(J.totFeat, TB), TB, 0, s>>>(J.d.p, J.dFeatPre.p, K, J.totFeat, Vinv.p, eF); nl++; }
    if (gauge) { k_mono_fix<<<ceil_div(K, TB), TB, 0, s>>>(J.d.p, K, gauge->fixScalar, gauge->sign); nl++; }
    KERNEL_CHECK();
    ctx.end(144.0 * J.totW + 96.0 * J.totFeat + 8.0 * (6.0 * J.totPose + 3.0 * J.totFeat), 0.0, nl);

    if (dbg) {
        int m0 = J.h[0].m, n0 = sOff[1];
        dbg->rowptr.assign(hRowPtr, hRowPtr + m0 + 1);
        dbg->colidx.resize(n0);
        for (int i = 0; i < n0; i++) dbg->colidx[i] = (int)(hKeys[i] & ((1ull << 22) - 1));
        dbg->S.resize(36 * (size_t)n0);
        dbg->E.resize(6 * (size_t)m0);
        CUDA_CHECK(cudaMemcpyAsync(dbg->S.data(), S.p, sizeof(double) * 36 * (size_t)n0, cudaMemcpyDeviceToHost, s));
        CUDA_CHECK(cudaMemcpyAsync(dbg->E.data(), E.p, sizeof(double) * 6 * (size_t)m0, cudaMemcpyDeviceToHost, s));
        CUDA_CHECK(cudaStreamSynchronize(s));
        dbg->perm.assign(sym.perm.begin(), sym.perm.begin() + m0);
        dbg->chol_flops = sym.flops;
    }
}
