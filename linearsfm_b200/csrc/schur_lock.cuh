// k_schur_lock: lock-step Schur complement kernel for the upper tree levels (included by solve.cu).
//
//   S(p,q) -= sum_f W_pf V_f^-1 W_qf^T ,  E_p -= sum_f W_pf d_f      (LinearSFMImp.cpp:2246-2332;
//   d_f = V_f^-1 eF_f - xhat_f of the pose's side of the join, see k_vinv)
//
// Same decomposition as k_schur_pipe (schur_pipe.cuh): one CTA per chunk of SCH_FCHUNK consecutive
// features of one join; the chunk's <= 31 distinct poses have local indices (left behind by the pattern
// kernel); a thread owns one pose-pair block of S in REGISTERS for the whole chunk while the chunk's W
// blocks stream through a double-buffered cp.async raw stage.  What differs is the walk over the
// features.  At the upper levels a chunk sees 12-23 poses and 60-77 % of its (pair, feature)
// combinations are populated (hub poses of the ancestors' transforms are seen by every feature), so
// instead of every lane chasing ITS OWN list of features (divergent search, 32 distinct operand
// addresses per shared load, one 512-thread CTA per SM) all lanes step through the batch's features
// TOGETHER:
//   * an absent (pose, feature) block is an all-zero block (index RB of every operand buffer), so the
//     108-FMA body has no branch and adds exact zeros where the reference adds nothing;
//   * pairs are laid out diagonal first, then row-major (i, j>i): the lanes of a warp read the SAME
//     W V^-1 block (broadcast) and CONSECUTIVE W blocks of the feature, ~3 shared-memory wavefronts per
//     pair of 16-byte operand loads instead of 8;
//   * the pose's share of E is carried by nposes extra "E items" (18 FMAs per feature instead of 108)
//     so that the pair lanes run one branch-free body;
//   * when the chunk's P pairs fit THREADS several times, the features of a batch are dealt to NS feature
//     slices of the pairs and ceil(NS / 3) slices of the E items (thread = (slice, item)); the slices'
//     partial sums are added in slice order through shared memory (fixed order: same bits every run);
//   * chunks with more items than threads (> 21 poses: ~1 % of the root level) take several passes
//     over their features, one item window per pass;
//   * top levels: 256 threads, 128 registers, ~108 KB of shared memory: TWO CTAs per SM, so one chunk's
//     barrier / re-layout phases overlap the other's arithmetic; middle levels (<= 17 poses per chunk,
//     little arithmetic per chunk: the chain of dependent global round trips of prologue, batches and
//     flush dominates): 128 threads and a 96-block stage, FOUR CTAs per SM.
// Flush: identical to k_schur_pipe (fixed-point integer atomics into S, E records per chunk and pose).
#pragma once

namespace schur_lock {

using schur_pipe::cp_async16;
using schur_pipe::cp_async8;
using schur_pipe::cp_async4;
using schur_pipe::cp_async_commit;
using schur_pipe::cp_async_wait_all;

template <int RB, int NBMAX>
struct Layout {
    // byte offsets; the first three areas are contiguous: the slice reduction reuses them as scratch
    static constexpr int rawW = 0;                                   // [2][(RB+1)*18] double, block RB = 0
    static constexpr int WVsm = rawW + 2 * (RB + 1) * 18 * 8;        // [(RB+1)*18] double, block RB = 0
    static constexpr int rawVi = WVsm + (RB + 1) * 18 * 8;           // [2][NBMAX*9+1] double
    static constexpr int rawEf = rawVi + 2 * (NBMAX * 9 + 1) * 8;    // [2][NBMAX*6] double
    static constexpr int rawPh = rawEf + 2 * (NBMAX * 6) * 8;        // [2][RB] int
    static constexpr int blkOf = (rawPh + 2 * RB * 4 + 15) / 16 * 16;   // [2][NBMAX*32] unsigned char
    static constexpr int poses = blkOf + 2 * NBMAX * 32;             // [32] int
    static constexpr int wptr = poses + 32 * 4;                      // [SCH_FCHUNK+1] int
    static constexpr int end = wptr + (SCH_FCHUNK + 1 + 3) / 4 * 16;
    static constexpr int scratch = WVsm + (RB + 1) * 18 * 8;         // bytes usable by the slice reduction
    static size_t bytes() { return (size_t)end + 16; }
};

template <int RB, int NBMAX, int THREADS, int MINCTA>
__global__ void __launch_bounds__(THREADS, MINCTA)
k_schur_lock(const DMap *__restrict__ J, const FeatChunk *__restrict__ chunks,
             const int *__restrict__ chunkInfo, const int *__restrict__ blkInfo, int pat_cmax,
             const int *__restrict__ wPre, const int *__restrict__ featPre, const int *__restrict__ posePre,
             const double *__restrict__ Vinv, const double *__restrict__ dvec, const int *__restrict__ split,
             const u64 *__restrict__ keys, const int *__restrict__ rowPtr,
             double *__restrict__ S, double *__restrict__ E,
             const int *__restrict__ sexp, long long *__restrict__ Sfx, double *__restrict__ Erec,
             const SlowFx slow)
{
    static_assert(RB >= 62 && RB <= 254, "block budget (a feature has <= 31 blocks; indices are bytes)");
    static_assert(Layout<RB, NBMAX>::scratch >= (THREADS - 3) * 36 * 8, "slice reduction scratch");
    typedef Layout<RB, NBMAX> L;
    extern __shared__ __align__(16) unsigned char smraw[];
    double *rawW = (double *)(smraw + L::rawW);
    double *WVsm = (double *)(smraw + L::WVsm);
    double *rawVi = (double *)(smraw + L::rawVi);
    double *rawEf = (double *)(smraw + L::rawEf);
    int *rawPh = (int *)(smraw + L::rawPh);
    unsigned char *blkOf = (unsigned char *)(smraw + L::blkOf);
    int *poses = (int *)(smraw + L::poses);
    int *wptr = (int *)(smraw + L::wptr);
    const FeatChunk ch = chunks[blockIdx.x];
    const DMap &M = J[ch.k];
    const int k = ch.k;
    const int tid = threadIdx.x;
    const int nfeat = ch.f1 - ch.f0;
    const int *ci = chunkInfo + CHUNK_INFO_INTS * (size_t)blockIdx.x;

    for (int i = tid; i <= nfeat; i += THREADS) wptr[i] = M.wPtr[ch.f0 + i];
    if (tid < 31) poses[tid] = ci[tid];
    const int nposes = ci[31];
    // the all-zero block of the three operand buffers
    if (tid < 18) {
        rawW[RB * 18 + tid] = 0.0;
        rawW[(RB + 1) * 18 + RB * 18 + tid] = 0.0;
        WVsm[RB * 18 + tid] = 0.0;
    }
    __syncthreads();
    const int w0 = wptr[0], w1 = wptr[nfeat];
    if (nposes == 0) return;
    if (nposes > 31 || nposes > pat_cmax) {       // the pattern kernel's overflow chunks
        for (int a = w0 + tid; a < w1; a += THREADS)
            schur_block_slow(M, k, a, featPre, posePre, Vinv, dvec, split, keys, rowPtr, slow);
        return;
    }
    const int P = nposes * (nposes + 1) / 2;
    // Items: NS feature slices of the P pairs, then NSE feature slices of the nposes E items.  An E item costs
    // about a third of a pair item per feature (lookup + 18 FMAs against 108), so NSE = ceil(NS / 3) keeps the
    // E lanes from becoming the stragglers of a batch without spending pair lanes on them.
    int NS = 1, NSE = 1, npass = 1;
    if (P + nposes <= THREADS) {
        NS = min((THREADS - nposes) / P, NBMAX);
        while (NS > 1 && NS * P + ((NS + 2) / 3) * nposes > THREADS) NS--;
        NSE = (NS + 2) / 3;
    } else npass = (P + nposes + THREADS - 1) / THREADS;
    const int nPairItems = NS * P;

    const double *Wg = M.W;
    const int *Pg = blkInfo + wPre[k];
    const double *Vg = Vinv + 9 * (size_t)(featPre[k] + ch.f0);
    const double *Eg = dvec + 6 * (size_t)(featPre[k] + ch.f0);

    // end of the batch that starts at feature fa (uniform across the CTA)
    auto batch_end = [&](int fa) {
        int target = wptr[fa] + RB;
        int lo = fa + 1, hi = min(nfeat, fa + NBMAX);
        while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (wptr[mid] <= target) lo = mid; else hi = mid - 1; }
        return lo;
    };
    // raw stage of the batch [fa, fe) into buffer `buf` (asynchronous)
    auto issue = [&](int fa, int fe, int buf) {
        const int nbf = fe - fa;
        const int b0 = wptr[fa], nblk = wptr[fe] - b0;
        double *dW = rawW + buf * ((RB + 1) * 18);
        const double *sW = Wg + 18 * (size_t)b0;
        for (int c = tid; c < nblk * 9; c += THREADS) cp_async16(dW + 2 * c, sW + 2 * c);
        int *dP = rawPh + buf * RB;
        for (int c = tid; c < nblk; c += THREADS) cp_async4(dP + c, Pg + b0 + c);
        double *dV = rawVi + buf * (NBMAX * 9 + 1);
        for (int c = tid; c < nbf * 9; c += THREADS) cp_async8(dV + c, Vg + 9 * (size_t)fa + c);
        double *dE = rawEf + buf * (NBMAX * 6);
        for (int c = tid; c < nbf * 6; c += THREADS) cp_async8(dE + c, Eg + 6 * (size_t)fa + c);
        cp_async_commit();
    };

    for (int pass = 0; pass < npass; pass++) {
        // this thread's item: a pose pair (kind 1; diagonal pairs first, then (i, j>i) row by row), a
        // pose's share of E (kind 2), or nothing
        int kind = 0, pi = 0, pj = 0, slice = 0, nsl = 1;   // nsl: slices of this item's kind
        const int item = tid + pass * THREADS;        // [0, NS P): slice-major pairs; then slice-major E items
        if (item < nPairItems) {
            kind = 1; nsl = NS;
            slice = item / P;
            int t = item - slice * P;
            if (t < nposes) { pi = t; pj = t; }
            else {
                t -= nposes;
                int i = 0;
                while (t >= nposes - 1 - i) { t -= nposes - 1 - i; i++; }
                pi = i; pj = i + 1 + t;
            }
        } else if (item - nPairItems < NSE * nposes) {
            kind = 2; nsl = NSE;
            slice = (item - nPairItems) / nposes;
            pi = pj = (item - nPairItems) - slice * nposes;
        }
        double acc[36];
#pragma unroll
        for (int q = 0; q < 36; q++) acc[q] = 0.0;
        const int sideOff = (split && poses[pi] >= split[k]) ? 3 : 0;   // E lanes: End / Cur side's d vector

        __syncthreads();                                   // previous pass done with every buffer
        for (int i = tid; i < 2 * NBMAX * 32; i += THREADS) blkOf[i] = (unsigned char)RB;
        int fa = 0, fe = batch_end(0);
        issue(fa, fe, 0);
        for (int bi = 0; fa < nfeat; bi++) {
            const int buf = bi & 1;
            const int nbf = fe - fa;
            const int b0 = wptr[fa], nblk = wptr[fe] - b0;
            const int fa2 = fe, fe2 = (fa2 < nfeat) ? batch_end(fa2) : fa2;
            cp_async_wait_all();
            __syncthreads();                               // this batch landed; previous update done
            if (fa2 < nfeat) issue(fa2, fe2, buf ^ 1);
            if (bi > 0)
                for (int i = tid; i < NBMAX * 32; i += THREADS) blkOf[(buf ^ 1) * NBMAX * 32 + i] = (unsigned char)RB;
            // re-layout: one thread per (block, row) writes the row of W V^-1; W stays in the raw buffer
            const double *rW = rawW + buf * ((RB + 1) * 18);
            const int *rP = rawPh + buf * RB;
            const double *rV = rawVi + buf * (NBMAX * 9 + 1);
            unsigned char *bo = blkOf + buf * NBMAX * 32;
            for (int e = tid; e < nblk * 6; e += THREADS) {
                int blk = e / 6, r = e - 6 * blk;
                int info = rP[blk];
                int fb = (info >> 8) - fa, slot = info & 255;
                const double *wr = rW + 18 * blk + 3 * r;
                const double *vi = rV + 9 * fb;
                double w0_ = wr[0], w1_ = wr[1], w2_ = wr[2];
                double *dv = WVsm + blk * 18 + 3 * r;
                dv[0] = w0_ * vi[0] + w1_ * vi[1] + w2_ * vi[2];
                dv[1] = w0_ * vi[3] + w1_ * vi[4] + w2_ * vi[5];
                dv[2] = w0_ * vi[6] + w1_ * vi[7] + w2_ * vi[8];
                if (r == 0) bo[fb * 32 + slot] = (unsigned char)blk;
            }
            __syncthreads();
            if (kind == 1) {
                int ba = slice < nbf ? bo[slice * 32 + pi] : RB, bb = slice < nbf ? bo[slice * 32 + pj] : RB;
                for (int fb = slice; fb < nbf; fb += NS) {
                    const double2 *wv2 = reinterpret_cast<const double2 *>(WVsm + ba * 18);
                    const double2 *w2 = reinterpret_cast<const double2 *>(rW + bb * 18);
                    const int fn = fb + NS;                // next feature's block indices, ahead of the FMAs
                    if (fn < nbf) { ba = bo[fn * 32 + pi]; bb = bo[fn * 32 + pj]; }
                    double b[18];
#pragma unroll
                    for (int q = 0; q < 9; q++) { double2 v = w2[q]; b[2 * q] = v.x; b[2 * q + 1] = v.y; }
#pragma unroll
                    for (int rp = 0; rp < 3; rp++) {       // two rows of W V^-1 per three 16-byte loads
                        const double2 x0 = wv2[3 * rp], x1 = wv2[3 * rp + 1], x2 = wv2[3 * rp + 2];
                        const double av[6] = {x0.x, x0.y, x1.x, x1.y, x2.x, x2.y};
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            const int r = 2 * rp + h;
                            const double a0 = av[3 * h], a1 = av[3 * h + 1], a2 = av[3 * h + 2];
#pragma unroll
                            for (int c = 0; c < 6; c++)
                                acc[6 * r + c] = fma(a2, b[3 * c + 2], fma(a1, b[3 * c + 1], fma(a0, b[3 * c], acc[6 * r + c])));
                        }
                    }
                }
            } else if (kind == 2) {
                const double *rE = rawEf + buf * (NBMAX * 6) + sideOff;
                for (int fb = slice; fb < nbf; fb += NSE) {
                    const int bb = bo[fb * 32 + pi];
                    if (bb == RB) continue;
                    const double2 *w2 = reinterpret_cast<const double2 *>(rW + bb * 18);
                    const double e0 = rE[6 * fb], e1 = rE[6 * fb + 1], e2 = rE[6 * fb + 2];
                    double b[18];
#pragma unroll
                    for (int q = 0; q < 9; q++) { double2 v = w2[q]; b[2 * q] = v.x; b[2 * q + 1] = v.y; }
#pragma unroll
                    for (int r = 0; r < 6; r++)
                        acc[r] = fma(b[3 * r + 2], e2, fma(b[3 * r + 1], e1, fma(b[3 * r], e0, acc[r])));
                }
            }
            fa = fa2; fe = fe2;
        }
        // partial blocks of slices 1.. go through shared memory ([entry][thread]: conflict-free) and are
        // added by the slice-0 thread of the pair in slice order
        if (NS > 1) {                                      // (NSE > 1 implies NS > 1)
            __syncthreads();                               // every operand buffer is free now
            double *red = (double *)smraw;
            // scratch rows: the pair items of slices 1.., then the E items of slices 1..
            const int nredP = (NS - 1) * P, nred = nredP + (NSE - 1) * nposes;
            const int width = kind == 1 ? P : nposes;                     // items per slice of this kind
            const int base = kind == 1 ? item - P : nredP + (item - nPairItems) - nposes;   // row of a slice > 0 item
            if (kind != 0 && slice > 0) {
#pragma unroll
                for (int q = 0; q < 36; q++) red[q * nred + base] = acc[q];
            }
            __syncthreads();
            if (kind != 0 && slice == 0) {
                const int first = kind == 1 ? item : nredP + (item - nPairItems);   // this item's row in slice 1
                for (int sl = 1; sl < nsl; sl++) {
#pragma unroll
                    for (int q = 0; q < 36; q++) acc[q] += red[q * nred + first + (sl - 1) * width];
                }
            }
        }
        if (kind == 2 && slice == 0) {
            // the pose's share of E from this chunk: a record, gathered per pose by k_e_gather
            double *e = Erec + 6 * (32 * (size_t)blockIdx.x + pi);
#pragma unroll
            for (int q = 0; q < 6; q++) e[q] = acc[q];
        } else if (kind == 1 && slice == 0) {
            // the pattern kernel's bitmap of the chunk's pairs (row-major upper triangle incl. diagonal)
            // decides which pairs exist
            const int i = pi, j = pj;
            const int idx = i * nposes - (i * (i - 1)) / 2 + (j - i);
            if (((unsigned)ci[32 + (idx >> 5)] >> (idx & 31)) & 1u) {
                const int gi = poses[i], gj = poses[j];
                const int slot = find_slot(keys, rowPtr, posePre[k] + gi, pair_key(k, gi, gj));
                unsigned long long *sp = reinterpret_cast<unsigned long long *>(Sfx) + 36 * (size_t)slot;
                int ei[6], ej[6];
#pragma unroll
                for (int q = 0; q < 6; q++) {
                    ei[q] = sexp[6 * (size_t)(posePre[k] + gi) + q];
                    ej[q] = sexp[6 * (size_t)(posePre[k] + gj) + q];
                }
                // fixed point, integer atomics: exact, hence independent of the order the chunks arrive in
                if (i == j) {
#pragma unroll
                    for (int r = 0; r < 6; r++)
#pragma unroll
                        for (int c = r; c < 6; c++) {
                            const long long v = __double2ll_rn(acc[6 * r + c] * pow2(fx_shift(ei[r], ej[c])));
                            atomicAdd(sp + 6 * r + c, (unsigned long long)v);
                            if (c > r) atomicAdd(sp + 6 * c + r, (unsigned long long)v);
                        }
                } else {
#pragma unroll
                    for (int r = 0; r < 6; r++)
#pragma unroll
                        for (int c = 0; c < 6; c++) {
                            const long long v = __double2ll_rn(acc[6 * r + c] * pow2(fx_shift(ei[r], ej[c])));
                            atomicAdd(sp + 6 * r + c, (unsigned long long)v);
                        }
                }
            }
        }
    }
}

} // namespace schur_lock
