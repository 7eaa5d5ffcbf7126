// k_tf_chunk: the W/V part of the frame transform as ONE pass over W (included by transform.cu).
//
//   LinearSFMImp.cpp:1300-1915 -- for a map re-expressed in the frame of its pose `pos`:
//     per feature f : X' = R (X - t), V' = Q^T V Q, new block W'(pos,f) = [-V'; T_f^T V Q] + sum_p wadd_pf,
//                     U'(pos,pos) += C_f^T V C_f
//     per block (p,f): a1 = D_p^T W Q  (kept, unless p == pos: folded into W'(pos,f)),
//                     wadd = C_p^T W Q (p != pos)  or  D_pos^T W Q (p == pos)
//     per pose p    : SW_p = sum_f W_pf, SWT_p = sum_f W_pf T_f  (k_tf_posefin applies the pose
//                     Jacobians once per pose: U'(p,pos), U'(pos,pos))
//
// One CTA (128 threads) per chunk of TC_FCH consecutive features of one map.  Prologue: local pose
// table of the chunk (bitmap + popcount prefix), per-feature prep (thread per feature: X', V', d_f,
// the V part of W'(pos,f) -> shared accumulator, C_f^T V C_f -> CTA sum).  Then the chunk's W blocks
// are walked in batches of 128, ONE THREAD PER BLOCK with the whole 6x3 block in registers:
//     A   nine 16-byte loads of the block (-> shared tile); T_f from d_f; W T_f -> shared tile; X = W Q;
//         a1 -> its output slot directly (nine 16-byte stores); wadd -> shared tile;
//         the block joins the batch list of its local pose
//     A'  pose sums of the WARP's 32 blocks on the FP64 tensor cores: [slot x block] indicator (slots
//         exchanged by shuffles) times the [block x 36] W | W T_f rows of the warp, mma.sync m8n8k4 (DMMA),
//         five independent column tiles per tile of eight local poses; the warp's partial sums replace
//         its (consumed) W / W T_f rows in shared memory
//     B   thread per (local pose, element): adds the four warps' partial sums, in warp order, to its
//         accumulator (in REGISTERS for the whole chunk);
//         thread per (feature, element pair): the feature's wadd rows of this batch -> W'(pos,f)
// Two barriers per batch; W is read once from HBM and written once.  At the end the pose sums are
// written as one record per (chunk, local pose) next to the chunk's pose bitmap; k_tf_posefin gathers
// them per pose in a fixed order (no FP64 atomics: bit-identical results run to run).
// Chunks with more than 31 distinct poses take a slow path (thread per block, global atomics).
#pragma once

namespace tfc {

constexpr int TC_FCH = 128;       // features per chunk
constexpr int TC_THREADS = 128;
constexpr int TC_BATCH = 128;     // W blocks per batch (one per thread)
constexpr int TC_CMAX = 31;       // distinct poses per chunk on the fast path
constexpr int TC_LD = 18;         // dense tile rows: written / read per block as nine 16-byte accesses (conflict-free)
constexpr int TC_ACC = (TC_CMAX * 36 + TC_THREADS - 1) / TC_THREADS;   // pose-sum accumulators per thread

struct Chunk { int k, f0, f1; };

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

struct Layout {
    static constexpr int WA = 0;                                           // [BATCH][LD] double: wadd rows
    static constexpr int WT = WA + TC_BATCH * TC_LD * 8;                   // [BATCH][LD] double: W T_f
    static constexpr int Wr = WT + TC_BATCH * TC_LD * 8;                   // [BATCH][LD] double: the W blocks themselves
    static constexpr int dF = Wr + TC_BATCH * TC_LD * 8;                   // [FCH][3] double: X'_f - t'
    static constexpr int Jt = dF + TC_FCH * 3 * 8;                         // [CMAX][27] double: c1, Bm, Cm per local pose
    static constexpr int Cst = Jt + TC_CMAX * 27 * 8;                      // Q, QA, QB, QG (36 doubles)
    static constexpr int wptr = Cst + 36 * 8;                              // [FCH+4] int
    static constexpr int optr = wptr + (TC_FCH + 4) * 4;                   // [FCH+4] int
    static constexpr int pidPos = optr + (TC_FCH + 4) * 4;                 // [FCH] int
    static constexpr int lcnt = (pidPos + TC_FCH * 4 + 15) / 16 * 16;                       // (1 KB, unused)
    static constexpr int poses = lcnt + 2 * 4 * 32 * 4;                    // [32] int
    static constexpr int misc = poses + 32 * 4;                            // [4] int
    static constexpr int bitmap = misc + 16;                               // [words] unsigned + [words] int
    static size_t bytes(int words) { return (size_t)bitmap + 8 * (size_t)words + 16; }
};

// bottom half of J^T X for the block-triangular pose Jacobian [[a,b],[0,c]]: b^T Xtop + c^T Xbot
__device__ __forceinline__ void jt_bottom(const double *__restrict__ b, const double *__restrict__ c,
                                          bool has_b, const double *X, double *out)
{
    double cc[9];
#pragma unroll
    for (int i = 0; i < 9; i++) cc[i] = c[i];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 3; k++) s = fma(cc[3 * k + i], X[9 + 3 * k + j], s);
            out[3 * i + j] = s;
        }
    if (has_b) {
#pragma unroll
        for (int i = 0; i < 9; i++) cc[i] = b[i];
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) {
                double s = out[3 * i + j];
#pragma unroll
                for (int k = 0; k < 3; k++) s = fma(cc[3 * k + i], X[3 * k + j], s);
                out[3 * i + j] = s;
            }
    }
}

__global__ void __launch_bounds__(TC_THREADS, 3)
k_tf_chunk(const DMap *__restrict__ in, DMap *__restrict__ out, const Chunk *__restrict__ chunks,
           const int *__restrict__ featPre, const int *__restrict__ posePre,
           const TfConst *__restrict__ tc, const PoseJac *__restrict__ pj, const int *__restrict__ fScan,
           const int *__restrict__ pexp, const int *__restrict__ pcnt, long long *__restrict__ poseFx, int cmaxUse,
           unsigned *__restrict__ chunkBits, int bitsStride, double *__restrict__ chunkRec,
           int *__restrict__ ppKey, double *__restrict__ ppVal, int *__restrict__ poseCR)
{
    typedef Layout L;
    extern __shared__ __align__(16) unsigned char smraw[];
    double *WAt = (double *)(smraw + L::WA);
    double *WTt = (double *)(smraw + L::WT);
    double *Wrt = (double *)(smraw + L::Wr);
    double *dFs = (double *)(smraw + L::dF);
    double *Jt = (double *)(smraw + L::Jt);
    double *Cst = (double *)(smraw + L::Cst);
    int *wptr = (int *)(smraw + L::wptr);
    int *optr = (int *)(smraw + L::optr);
    int *pidPos = (int *)(smraw + L::pidPos);
    int *poses = (int *)(smraw + L::poses);
    int *misc = (int *)(smraw + L::misc);
    unsigned *bitmap = (unsigned *)(smraw + L::bitmap);

    const Chunk ch = chunks[blockIdx.x];
    const int k = ch.k;
    const DMap &M = in[k];
    const DMap &O = out[k];
    const TfConst &c = tc[k];
    const int pid = c.posID;
    const int words = (M.m + 31) >> 5;
    int *prefix = (int *)(bitmap + words);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nfeat = ch.f1 - ch.f0;
    const int gf0 = featPre[k] + ch.f0;
    const PoseJac *pjk = pj + posePre[k];

    // ---------------- prologue ----------------
    {
        const int fbase = fScan[featPre[k]];
        for (int i = tid; i <= nfeat; i += TC_THREADS) {
            wptr[i] = M.wPtr[ch.f0 + i];
            optr[i] = fScan[gf0 + i] - fbase;
        }
    }
    for (int i = tid; i < nfeat; i += TC_THREADS) pidPos[i] = 0x7fffffff;
    for (int i = tid; i < words; i += TC_THREADS) bitmap[i] = 0u;
    if (tid < 9) {
        Cst[tid] = c.Q[tid]; Cst[9 + tid] = c.QA[tid]; Cst[18 + tid] = c.QB[tid]; Cst[27 + tid] = c.QG[tid];
    }
    __syncthreads();
    const int w0 = wptr[0], w1 = wptr[nfeat];
    for (int j0 = w0 + tid; j0 < w1; j0 += 4 * TC_THREADS) {       // (four loads in flight per thread)
        int pp[4];
#pragma unroll
        for (int u = 0; u < 4; u++) pp[u] = (j0 + u * TC_THREADS < w1) ? M.photo[j0 + u * TC_THREADS] : -1;
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int p = pp[u], j = j0 + u * TC_THREADS;
            if (p < 0) continue;
            atomicOr(&bitmap[p >> 5], 1u << (p & 31));
            if (p == pid) {                     // position of the pos block inside its feature
                const int fl = M.feature[j] - ch.f0;
                pidPos[fl] = j - wptr[fl];
            }
        }
    }
    if (tid < nfeat) O.wPtr[ch.f0 + tid] = optr[tid];
    __syncthreads();
    if (warp == 0) {
        int run = 0;
        for (int base = 0; base < words; base += 32) {
            int cc = (base + lane < words) ? __popc(bitmap[base + lane]) : 0;
            int incl = cc;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            if (base + lane < words) prefix[base + lane] = run + incl - cc;
            run += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) misc[0] = run;
    }
    __syncthreads();
    const int nposes = misc[0];
    const bool fast = nposes <= cmaxUse;
    // the chunk's pose bitmap + popcount prefix go to global memory: k_tf_posefin finds a pose's record
    // slot in this chunk with two loads (slow-path chunks publish an empty bitmap: nothing recorded)
    {
        unsigned *gb = chunkBits + (size_t)blockIdx.x * 2 * bitsStride;
        for (int i = tid; i < words; i += TC_THREADS) {
            gb[i] = fast ? bitmap[i] : 0u;
            gb[bitsStride + i] = (unsigned)prefix[i];
        }
    }
    if (fast)
        for (int i = tid; i < words; i += TC_THREADS) {
            unsigned b = bitmap[i];
            int r = prefix[i];
            while (b) {
                int bit = __ffs(b) - 1;
                poses[r++] = i * 32 + bit;
                // range of chunks that hold a record of the pose (both as running maxima over zeroed memory):
                // k_tf_posefin walks only those
                int *cr = poseCR + 2 * (size_t)(posePre[k] + i * 32 + bit);
                atomicMax(cr, 0x7fffffff - (int)blockIdx.x);
                atomicMax(cr + 1, (int)blockIdx.x + 1);
                b &= b - 1;
            }
        }
    double Q[9];
#pragma unroll
    for (int i = 0; i < 9; i++) Q[i] = Cst[i];
    // per-feature prep: X', V', d_f, the V part of W'(pos,f) and of U'(pos,pos)
    {
        double S21[21];
#pragma unroll
        for (int i = 0; i < 21; i++) S21[i] = 0.0;
        if (tid < nfeat) {
            const int f = ch.f0 + tid;
            const double *x = M.featVal + 3 * (size_t)f;
            double d0[3] = {x[0] - c.t[0], x[1] - c.t[1], x[2] - c.t[2]};
            double xn[3];
            geom::mat3_vec(c.R, d0, xn);
            double *y = O.featVal + 3 * (size_t)f;
            y[0] = xn[0]; y[1] = xn[1]; y[2] = xn[2];
            O.featNo[f] = M.featNo[f];
            double d[3] = {xn[0] - c.tn[0], xn[1] - c.tn[1], xn[2] - c.tn[2]};
            dFs[3 * tid] = d[0]; dFs[3 * tid + 1] = d[1]; dFs[3 * tid + 2] = d[2];
            double Tf[9], v[3];
            geom::mat3_vec(Cst + 9, d, v);  Tf[0] = v[0]; Tf[3] = v[1]; Tf[6] = v[2];
            geom::mat3_vec(Cst + 18, d, v); Tf[1] = v[0]; Tf[4] = v[1]; Tf[7] = v[2];
            geom::mat3_vec(Cst + 27, d, v); Tf[2] = v[0]; Tf[5] = v[1]; Tf[8] = v[2];
            double V[9], VQ[9], VT[9], Vn[9], M1[9], M2[9];
            sm::load<9>(M.V + 9 * (size_t)f, V);
            sm::mm<3, 3, 3>(V, Q, VQ);
            sm::mm<3, 3, 3>(V, Tf, VT);
            sm::mtm<3, 3, 3>(Q, VQ, Vn);
            sm::mtm<3, 3, 3>(Tf, VQ, M1);
            sm::mtm<3, 3, 3>(Tf, VT, M2);
            sm::store<9>(O.V + 9 * (size_t)f, Vn);
            const int o0 = optr[tid];
            O.photo[o0] = pid;
            O.feature[o0] = f;
            // the (posID,f) block starts as C_f^T V D_f = [-V'; M1]; the batches add the W terms
            // (read-modify-write by threads of this CTA only, ordered by the CTA barriers)
            double *wp = O.W + 18 * (size_t)o0;
#pragma unroll
            for (int i = 0; i < 9; i++) { wp[i] = -Vn[i]; wp[9 + i] = M1[i]; }
            // upper triangle of C_f^T V C_f = [[V', -M1^T],[-M1, M2]]
            int q = 0;
#pragma unroll
            for (int r = 0; r < 6; r++)
#pragma unroll
                for (int cc = r; cc < 6; cc++) {
                    double s;
                    if (r < 3 && cc < 3) s = Vn[3 * r + cc];
                    else if (r < 3) s = -M1[3 * (cc - 3) + r];
                    else s = M2[3 * (r - 3) + (cc - 3)];
                    S21[q++] = s;
                }
        }
        // CTA sum of the 21 values (scratch: the WA/WT tiles, 21 x 128 doubles)
        double *scr = WAt;
#pragma unroll
        for (int i = 0; i < 21; i++) scr[i * TC_FCH + tid] = S21[i];
        __syncthreads();
        if (fast)                                   // Jacobian blocks of the chunk's poses -> shared
            for (int i = tid; i < nposes * 27; i += TC_THREADS) {
                const int slot = i / 27, e = i - 27 * slot;
                const int p = poses[slot];
                const PoseJac &J = pjk[p];
                const bool isPos = (p == pid);
                Jt[i] = e < 9 ? J.c1[e] : e < 18 ? (isPos ? J.b1[e - 9] : J.f2[e - 9]) : (isPos ? J.c1[e - 18] : J.g2[e - 18]);
            }
        {
            const int i = tid >> 2, part = tid & 3;           // 32 groups of 4 lanes, 21 in use
            double s = 0.0;
            if (i < 21)
                for (int t = 0; t < 32; t++) s += scr[i * TC_FCH + part * 32 + t];
            s += __shfl_down_sync(0xffffffffu, s, 2, 4);
            s += __shfl_down_sync(0xffffffffu, s, 1, 4);
            if (part == 0 && i < 21) {
                // the chunk's share of U'(pos,pos): one record per chunk (deterministic reduction later)
                int r = 0, rem = i;
                while (rem >= 6 - r) { rem -= 6 - r; r++; }
                int cc = r + rem;
                double *u = ppVal + 36 * (size_t)blockIdx.x;
                u[6 * r + cc] = s;
                if (cc != r) u[6 * cc + r] = s;
                if (i == 0) ppKey[blockIdx.x] = k;              // hot-target list: key = map index
            }
        }
        __syncthreads();
    }

    double acc[TC_ACC];                                             // pose sums: thread owns (local pose, element) items
#pragma unroll
    for (int u = 0; u < TC_ACC; u++) acc[u] = 0.0;
    const int nitems = fast ? nposes * 36 : 0;
    const int nMt = fast ? (nposes + 7) >> 3 : 0;                   // tiles of eight local poses
    const int fg = lane >> 2, ft = lane & 3;                        // fragment row group / position in the group
    // the warp's partial pose sums [slot][36] of a batch live where its W / W T_f rows were: slots 0-15 in the
    // W rows, 16-31 in the W T_f rows (32 x 18 doubles each)
    double *stW = Wrt + 18 * 32 * warp, *stT = WTt + 18 * 32 * warp;
    const double *Wg = M.W;

    // ---------------- batches: one thread per W block ----------------
    // phase A for one block; FAST: chunk-local pose slots (lists, Jacobians from shared memory, W
    // through the warp's private part of the Wr tile), else global Jacobians / atomics
    auto phase_a = [&](auto fast_tag, const int j, const int p, const int fb, const int slot) {
        constexpr bool FAST = decltype(fast_tag)::value;
        const bool isPos = (p == pid);
        double W[18];
        {
            const double2 *src = FAST ? reinterpret_cast<const double2 *>(Wrt + 18 * tid)
                                      : reinterpret_cast<const double2 *>(Wg + 18 * (size_t)j);
#pragma unroll
            for (int i = 0; i < 9; i++) { double2 v = src[i]; W[2 * i] = v.x; W[2 * i + 1] = v.y; }
        }
        {
            // T_f = [QA d | QB d | QG d];  W T_f -> tile (or straight to the pose sums on the slow path)
            const double d0 = dFs[3 * fb], d1 = dFs[3 * fb + 1], d2 = dFs[3 * fb + 2];
            double Tf[9];
#pragma unroll
            for (int cc = 0; cc < 3; cc++)
#pragma unroll
                for (int r = 0; r < 3; r++)
                    Tf[3 * r + cc] = fma(Cst[9 + 9 * cc + 3 * r + 2], d2,
                                         fma(Cst[9 + 9 * cc + 3 * r + 1], d1, Cst[9 + 9 * cc + 3 * r] * d0));
            double WT[18];
            sm::mm<6, 3, 3>(W, Tf, WT);
            if (FAST) {
                double2 *row = reinterpret_cast<double2 *>(WTt + 18 * tid);
#pragma unroll
                for (int i = 0; i < 9; i++) row[i] = make_double2(WT[2 * i], WT[2 * i + 1]);
            } else {
                // slow path: exact fixed-point integer atomics (scale from k_tf_slow_pre's exponent / count
                // bookkeeping): order-independent, same bits every run
                const size_t gp = (size_t)(posePre[k] + p);
                unsigned long long *pa = reinterpret_cast<unsigned long long *>(poseFx) + 36 * gp;
                const int bits = 32 - __clz(pcnt[gp]);
#pragma unroll
                for (int i = 0; i < 36; i++) {
                    const int em = pexp[36 * gp + i];
                    if (em < -2000) continue;
                    const double v = i < 18 ? W[i] : WT[i - 18];
                    const int sh = 60 - em - bits;
                    atomicAdd(pa + i, (unsigned long long)__double2ll_rn(v * __longlong_as_double((long long)(1023 + sh) << 52)));
                }
            }
        }
        double X[18];
        sm::mm<6, 3, 3>(W, Q, X);                                   // W Q
        double P[9], B1[9];
        sm::mtm<3, 3, 3>(Q, X, P);                                  // Q^T Xtop
        const PoseJac &J = pjk[p];
        const double *jc1 = FAST ? Jt + 27 * slot : J.c1;
        const double *jB = FAST ? Jt + 27 * slot + 9 : (isPos ? J.b1 : J.f2);
        const double *jC = FAST ? Jt + 27 * slot + 18 : (isPos ? J.c1 : J.g2);
        if (!isPos) {
            jt_bottom(jc1, jc1, false, X, B1);                      // a1 = D_p^T W Q = [P; c1^T Xbot]
            const int jrel = j - wptr[fb];
            const int o = optr[fb] + 1 + jrel - ((pidPos[fb] < jrel) ? 1 : 0);
            double2 *dst = reinterpret_cast<double2 *>(O.W + 18 * (size_t)o);
            dst[0] = make_double2(P[0], P[1]); dst[1] = make_double2(P[2], P[3]);
            dst[2] = make_double2(P[4], P[5]); dst[3] = make_double2(P[6], P[7]);
            dst[4] = make_double2(P[8], B1[0]); dst[5] = make_double2(B1[1], B1[2]);
            dst[6] = make_double2(B1[3], B1[4]); dst[7] = make_double2(B1[5], B1[6]);
            dst[8] = make_double2(B1[7], B1[8]);
            O.photo[o] = p;
            O.feature[o] = ch.f0 + fb;
        }
        // wadd = [-P; Bm^T Xtop + Cm^T Xbot]: C_p^T W Q, or D_pos^T W Q for the pos block
        jt_bottom(jB, jC, true, X, B1);
        {
            double2 *row = reinterpret_cast<double2 *>(WAt + 18 * tid);
            row[0] = make_double2(-P[0], -P[1]); row[1] = make_double2(-P[2], -P[3]);
            row[2] = make_double2(-P[4], -P[5]); row[3] = make_double2(-P[6], -P[7]);
            row[4] = make_double2(-P[8], B1[0]); row[5] = make_double2(B1[1], B1[2]);
            row[6] = make_double2(B1[3], B1[4]); row[7] = make_double2(B1[5], B1[6]);
            row[8] = make_double2(B1[7], B1[8]);
        }
    };

    int pNext = 0, fNext = 0;                                       // the block's pose / feature, one batch ahead
    if (w0 + tid < w1) { pNext = M.photo[w0 + tid]; fNext = M.feature[w0 + tid]; }
    for (int jb = w0; jb < w1; jb += TC_BATCH) {
        const int j = jb + tid;
        const int pCur = pNext, fb = fNext - ch.f0;
        if (j + TC_BATCH < w1) { pNext = M.photo[j + TC_BATCH]; fNext = M.feature[j + TC_BATCH]; }
        const int jl = min(jb + TC_BATCH, w1) - 1;                  // last block of the batch
        if (j == jb) misc[1] = fb;                                  // first / last feature of the batch
        if (j == jl) misc[2] = fb;
        if (fast) {
            // the warp's 32 blocks = one contiguous 4.6 KB span: coalesced 16-byte async copies into
            // the warp's private part of the Wr tile; the next batch is pulled into L2 meanwhile
            const int jw = jb + 32 * warp;
            const int nv = min(32, w1 - jw);
            if (nv > 0) {
                const double *srcw = Wg + 18 * (size_t)jw;
                double *dstw = Wrt + 18 * 32 * warp;
                for (int q = lane; q < nv * 9; q += 32) cp_async16(dstw + 2 * q, srcw + 2 * q);
            }
            cp_async_commit();
            if (j + TC_BATCH < w1) {
                const char *nx = reinterpret_cast<const char *>(Wg + 18 * (size_t)(j + TC_BATCH));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nx));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + 128));
            }
            cp_async_wait_all();
            __syncwarp();
            int slot = 0xff;                                        // no block (last batch): no slot
            if (j < w1) {
                slot = prefix[pCur >> 5] + __popc(bitmap[pCur >> 5] & ((1u << (pCur & 31)) - 1u));
                phase_a(std::true_type(), j, pCur, fb, slot);
            } else {
                // ... and zero rows in the W / W T_f tiles (0 x stale bits must not turn into a NaN in the
                // tensor-core sums)
                double2 *r0 = reinterpret_cast<double2 *>(Wrt + 18 * tid), *r1 = reinterpret_cast<double2 *>(WTt + 18 * tid);
#pragma unroll
                for (int i = 0; i < 9; i++) { r0[i] = make_double2(0.0, 0.0); r1[i] = make_double2(0.0, 0.0); }
            }
            __syncwarp();
            // pose sums of the warp's blocks: P [slot][36] = Ind [slot][block] x (W | W T_f) [block][36] as m8n8k4
            // DMMAs.  A K step takes the warp's blocks 8 i + {0,2,4,6} (+1 for the odd step) so that the B loads of a
            // half warp fall into distinct banks; the summation order is fixed by the code (same bits every run).
            if (nv > 0) {
                double c[4][5][2];
#pragma unroll
                for (int mt = 0; mt < 4; mt++)
#pragma unroll
                    for (int nt = 0; nt < 5; nt++) c[mt][nt][0] = c[mt][nt][1] = 0.0;
                const int npair = (nv + 7) >> 3;
                for (int i = 0; i < npair; i++) {
                    const int s0 = __shfl_sync(0xffffffffu, slot, 8 * i + 2 * ft);
                    const int s1 = __shfl_sync(0xffffffffu, slot, 8 * i + 2 * ft + 1);
                    double b0[5], b1[5];
#pragma unroll
                    for (int nt = 0; nt < 5; nt++) {
                        const int el = 8 * nt + fg;
                        const double *src = (el < 18 ? stW + el : stT + (el - 18)) + TC_LD * (8 * i + 2 * ft);
                        const bool ok = (nt < 4) || (fg < 4);
                        b0[nt] = ok ? src[0] : 0.0;
                        b1[nt] = ok ? src[TC_LD] : 0.0;
                    }
#pragma unroll
                    for (int mt = 0; mt < 4; mt++)
                        if (mt < nMt) {
                            const double a0 = __hiloint2double((s0 == 8 * mt + fg) ? 0x3ff00000 : 0, 0);
                            const double a1 = __hiloint2double((s1 == 8 * mt + fg) ? 0x3ff00000 : 0, 0);
#pragma unroll
                            for (int nt = 0; nt < 5; nt++) dmma(c[mt][nt][0], c[mt][nt][1], a0, b0[nt]);
#pragma unroll
                            for (int nt = 0; nt < 5; nt++) dmma(c[mt][nt][0], c[mt][nt][1], a1, b1[nt]);
                        }
                }
                __syncwarp();                                       // every lane has read its operands
#pragma unroll
                for (int mt = 0; mt < 4; mt++)
                    if (mt < nMt) {
                        double *row = ((mt < 2) ? stW : stT) + 36 * (8 * (mt & 1) + fg) + 2 * ft;
#pragma unroll
                        for (int nt = 0; nt < 5; nt++)
                            if (nt < 4 || ft < 2)
                                *reinterpret_cast<double2 *>(row + 8 * nt) = make_double2(c[mt][nt][0], c[mt][nt][1]);
                    }
            }
        } else if (j < w1) {
            phase_a(std::false_type(), j, pCur, fb, 0);
        }
        __syncthreads();
        // W'(pos,f) of the batch's features is a read-modify-write of global memory (the rows were
        // initialised by the prologue / earlier batches of THIS CTA): issue this thread's first two loads
        // now, the pose sums below hide their latency
        const int fLo = misc[1], nfb = misc[2] - fLo + 1;
        // (16-byte pieces: an item is a feature's element pair (2 el, 2 el + 1); rows are 144 bytes, aligned)
        double2 oldw[2] = {make_double2(0.0, 0.0), make_double2(0.0, 0.0)};
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int e = tid + u * TC_THREADS;
            if (e < nfb * 9) {
                const int fl = e / 9, el = e - 9 * fl;
                oldw[u] = reinterpret_cast<const double2 *>(O.W + 18 * (size_t)optr[fLo + fl])[el];
            }
        }
        // pose sums: the warps' partial sums, added in warp order
        {
            const int nwa = (jl - jb + 32) >> 5;                    // warps that held a block
#pragma unroll
            for (int u = 0; u < TC_ACC; u++) {
                const int it = tid + u * TC_THREADS;
                if (it < nitems) {
                    const int slot = it / 36, el = it - 36 * slot;
                    const double *src = ((slot < 16) ? Wrt : WTt) + 36 * (slot & 15) + el;
                    double sum = acc[u];
                    for (int w = 0; w < nwa; w++) sum += src[18 * 32 * w];
                    acc[u] = sum;
                }
            }
        }
        // W'(pos,f) += the feature's wadd rows of this batch
        {
            int u = 0;
            for (int e = tid; e < nfb * 9; e += TC_THREADS, u++) {
                const int fl = e / 9, el = e - 9 * fl;
                const int fx = fLo + fl;
                const int j0 = max(wptr[fx], jb) - jb, j1 = min(wptr[fx + 1], jl + 1) - jb;
                double s0 = 0.0, s1 = 0.0;
                for (int b = j0; b < j1; b++) {
                    const double2 v = reinterpret_cast<const double2 *>(WAt + b * TC_LD)[el];
                    s0 += v.x; s1 += v.y;
                }
                double2 *dst = reinterpret_cast<double2 *>(O.W + 18 * (size_t)optr[fx]) + el;
                const double2 o = (u == 0) ? oldw[0] : (u == 1) ? oldw[1] : *dst;
                *dst = make_double2(o.x + s0, o.y + s1);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < TC_ACC; u++) {
        const int it = tid + u * TC_THREADS;
        if (it < nitems) chunkRec[36 * 32 * (size_t)blockIdx.x + it] = acc[u];   // [chunk][slot][36]
    }
}

// Pre-pass of the slow path (chunks with more than cmaxUse distinct poses): per (pose, element) the largest
// exponent of a contribution to SW / SWT and per pose the number of contributions, so that the chunk kernel can
// accumulate them in fixed point.  Launched only when a map carries split chunks (or the test hook is on).
__global__ void __launch_bounds__(128)
k_tf_slow_pre(const DMap *__restrict__ in, const Chunk *__restrict__ chunks, const int *__restrict__ posePre,
              const TfConst *__restrict__ tc, int cmaxUse, int *__restrict__ pexp, int *__restrict__ pcnt)
{
    extern __shared__ unsigned bm[];               // [words] pose bitmap of the chunk
    __shared__ int cntSh;
    const Chunk ch = chunks[blockIdx.x];
    const DMap &M = in[ch.k];
    const TfConst &c = tc[ch.k];
    const int k = ch.k, words = (M.m + 31) >> 5, tid = threadIdx.x;
    const int w0 = M.wPtr[ch.f0], w1 = M.wPtr[ch.f1];
    for (int i = tid; i < words; i += 128) bm[i] = 0u;
    if (tid == 0) cntSh = 0;
    __syncthreads();
    for (int j = w0 + tid; j < w1; j += 128) { const int p = M.photo[j]; atomicOr(&bm[p >> 5], 1u << (p & 31)); }
    __syncthreads();
    int pc = 0;
    for (int i = tid; i < words; i += 128) pc += __popc(bm[i]);
    if (pc) atomicAdd(&cntSh, pc);
    __syncthreads();
    if (cntSh <= cmaxUse) return;
    for (int j = w0 + tid; j < w1; j += 128) {
        const int p = M.photo[j], f = M.feature[j];
        const double *x = M.featVal + 3 * (size_t)f;
        const double d0[3] = {x[0] - c.t[0], x[1] - c.t[1], x[2] - c.t[2]};
        double xn[3];
        geom::mat3_vec(c.R, d0, xn);
        const double d[3] = {xn[0] - c.tn[0], xn[1] - c.tn[1], xn[2] - c.tn[2]};
        double Tf[9], v[3];
        geom::mat3_vec(c.QA, d, v); Tf[0] = v[0]; Tf[3] = v[1]; Tf[6] = v[2];
        geom::mat3_vec(c.QB, d, v); Tf[1] = v[0]; Tf[4] = v[1]; Tf[7] = v[2];
        geom::mat3_vec(c.QG, d, v); Tf[2] = v[0]; Tf[5] = v[1]; Tf[8] = v[2];
        double W[18], WT[18];
        sm::load<18>(M.W + 18 * (size_t)j, W);
        sm::mm<6, 3, 3>(W, Tf, WT);
        const size_t gp = (size_t)(posePre[k] + p);
        atomicAdd(pcnt + gp, 1);
        for (int i = 0; i < 36; i++) {
            const double val = i < 18 ? W[i] : WT[i - 18];
            // +1: the chunk kernel forms T_f with fused multiply-adds, its W T_f can exceed this one by an ulp
            if (val != 0.0 && fabs(val) < 1e300) atomicMax(pexp + 36 * gp + i, ilogb(val) + 2);
        }
    }
}

// poseAcc = the slow path's fixed-point sums (one thread per (pose, element)); zero where nothing was added
__global__ void k_tf_slow_convert(int n36, const int *__restrict__ pexp, const int *__restrict__ pcnt,
                                  const long long *__restrict__ poseFx, double *__restrict__ poseAcc)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n36) return;
    const int em = pexp[g];
    double v = 0.0;
    if (em >= -2000) {
        const int sh = 60 - em - (32 - __clz(pcnt[g / 36]));
        v = (double)poseFx[g] * __longlong_as_double((long long)(1023 - sh) << 52);
    }
    poseAcc[g] = v;
}

} // namespace tfc
