// k_schur_dense: Schur complement of the reduced camera system with the dense part of every feature
// chunk on the FP64 tensor cores (included by solve.cu).
//
//   S(p,q) -= sum_f W_pf V_f^-1 W_qf^T ,  E_p -= sum_f W_pf d_f          (LinearSFMImp.cpp:2246-2332)
//
// One CTA per chunk of SCH_FCHUNK consecutive features of one join, as in schur_pipe.cuh (same raw
// stage: double-buffered cp.async batches of the chunk's contiguous W blocks / V^-1 / d vectors, W V^-1
// rows formed once per block).  What differs is who multiplies:
//   * The chunk's DENSE local poses (seen by at least half of its features; picked by the pattern
//     kernel, at most SCH_HMAX = 16) are the former frame origins every feature of the chunk is linked
//     to after a few Transforms: their pose pairs are ~75 % of all (pair, feature)
//     products at the top of the tree and nearly all of them at the bottom.  For them the update is one
//     dense contraction  S_DD -= A B^T,  A = [W_df] (6H x 3F), B = [W_df V_f^-1] (6H x 3F), absent
//     (pose, feature) combinations being zero.  It runs as mma.sync.m8n8k4.f64 (DMMA): a warp owns a
//     16 x 32 super-tile of S_DD (2 x 4 accumulator tiles in registers), k runs over (feature, column)
//     pairs with no padding (three k-steps per four features), and the operand fragments are read
//     straight from the block-major staging buffers through the per-batch (feature, pose) -> block
//     table -- no re-layout pass, no zero fill.  With few super-tiles the features of a batch are dealt
//     to several warps per super-tile (k-slices), combined in a fixed order at the end.
//   * Pairs with a sparse pose keep the thread-per-pair scheme of schur_pipe.cuh (one 6x6 accumulator
//     block per thread, every lane walking its own list of features that see both poses).
//   * E: one thread per (local pose, row).
// Flush: fixed-point integer atomics into S (exact, order-independent; solve.cu) and one 6-vector
// record per (chunk, local pose) for E.
#pragma once

namespace schur_dense {

using schur_pipe::cp_async16;
using schur_pipe::cp_async8;
using schur_pipe::cp_async4;
using schur_pipe::cp_async_commit;
using schur_pipe::cp_async_wait_all;

constexpr int LD = 18;

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int CMAX, int MAXBLK, int NBMAX, int THREADS>
struct Layout {
    static constexpr int NW = THREADS / 32;
    static constexpr int stageBytes = NW * 8 * 64 * 8;          // accumulator staging of the flush
    static constexpr int rawWBytes = 2 * MAXBLK * 18 * 8;
    static constexpr int rawW = 0;                                   // [2][MAXBLK*18] double (later: staging)
    static constexpr int rawVi = rawW + (rawWBytes > stageBytes ? rawWBytes : stageBytes);   // [2][NBMAX*9+1] double
    static constexpr int rawEf = rawVi + 2 * (NBMAX * 9 + 1) * 8;    // [2][NBMAX*6] double (d vectors)
    static constexpr int rawPh = rawEf + 2 * (NBMAX * 6) * 8;        // [2][MAXBLK] int (block infos)
    static constexpr int WVsm = (rawPh + 2 * MAXBLK * 4 + 15) / 16 * 16;   // [MAXBLK][LD] double: W V^-1
    static constexpr int present = WVsm + MAXBLK * LD * 8;           // [NBMAX] unsigned
    static constexpr int poses = present + NBMAX * 4;                // [32] int
    static constexpr int wptr = poses + 32 * 4;                      // [SCH_FCHUNK+4] int
    static constexpr int blkOf = wptr + (SCH_FCHUNK + 4) * 4;        // [NBMAX][32] unsigned char (255: absent)
    static constexpr int dslot = blkOf + NBMAX * 32;                 // [16] int: dense index -> local pose slot
    static constexpr int dexp = dslot + 16 * 4;                      // [16][6] int: row exponents of the dense poses
    static constexpr int ddSlot = dexp + 16 * 6 * 4;                 // [136] int: S slot of every dense pair (-1: not in the pattern)
    static constexpr int spairs = ddSlot + 136 * 4;                  // [496] unsigned short: sparse pairs (i << 8 | j)
    static constexpr int stl = spairs + 496 * 2;                     // [16] int: active super-tiles (si << 8 | sj)
    static constexpr int misc = stl + 16 * 4;                        // [4] int
    static constexpr int end = misc + 16;
    static size_t bytes() { return (size_t)end + 16; }
};

template <int CMAX, int MAXBLK, int NBMAX, int THREADS, int HCAP>
__global__ void __launch_bounds__(THREADS, 512 / THREADS)
k_schur_dense(const DMap *__restrict__ J, const FeatChunk *__restrict__ chunks,
              const int *__restrict__ chunkInfo, const int *__restrict__ blkInfo, int pat_cmax,
              const int *__restrict__ wPre, const int *__restrict__ featPre, const int *__restrict__ posePre,
              const double *__restrict__ Vinv, const double *__restrict__ dvec, const int *__restrict__ split,
              const u64 *__restrict__ keys, const int *__restrict__ rowPtr,
              double *__restrict__ S, double *__restrict__ E,
              const int *__restrict__ sexp, long long *__restrict__ Sfx, double *__restrict__ Erec)
{
    static_assert(MAXBLK >= 2 * CMAX && MAXBLK <= 254, "block budget");
    static_assert(NBMAX == 32, "a batch is eight groups of four features");
    typedef Layout<CMAX, MAXBLK, NBMAX, THREADS> L;
    constexpr int NW = THREADS / 32;
    extern __shared__ __align__(16) unsigned char smraw[];
    double *rawW = (double *)(smraw + L::rawW);
    double *rawVi = (double *)(smraw + L::rawVi);
    double *rawEf = (double *)(smraw + L::rawEf);
    int *rawPh = (int *)(smraw + L::rawPh);
    double *WVsm = (double *)(smraw + L::WVsm);
    unsigned *present = (unsigned *)(smraw + L::present);
    int *poses = (int *)(smraw + L::poses);
    int *wptr = (int *)(smraw + L::wptr);
    unsigned char *blkOf = (unsigned char *)(smraw + L::blkOf);
    int *dslot = (int *)(smraw + L::dslot);
    int *dexp = (int *)(smraw + L::dexp);
    int *ddSlot = (int *)(smraw + L::ddSlot);
    unsigned short *spairs = (unsigned short *)(smraw + L::spairs);
    int *stl = (int *)(smraw + L::stl);
    int *misc = (int *)(smraw + L::misc);

    const FeatChunk ch = chunks[blockIdx.x];
    const DMap &M = J[ch.k];
    const int k = ch.k;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nfeat = ch.f1 - ch.f0;
    const int *ci = chunkInfo + CHUNK_INFO_INTS * (size_t)blockIdx.x;

    for (int i = tid; i <= nfeat; i += THREADS) wptr[i] = M.wPtr[ch.f0 + i];
    if (tid < 32) poses[tid] = (tid < CMAX) ? ci[tid] : 0;
    const int nposes = ci[31];
    // the first HCAP dense poses of the pattern kernel's pick (the rest are treated as sparse: with
    // HCAP = 8 / 16 the dense part has at most 5 / 12 super-tiles, one per warp)
    unsigned denseMask = (unsigned)ci[48];
    int H = ci[49];
    while (H > HCAP) { denseMask &= ~(1u << (31 - __clz(denseMask))); H--; }
    if (tid == 0) misc[0] = 0;
    __syncthreads();
    const int w0 = wptr[0], w1 = wptr[nfeat];
    if (nposes == 0) return;
    if (nposes > CMAX || nposes > pat_cmax) {   // not expected: the host picks CMAX from the measured maximum
        for (int a = w0 + tid; a < w1; a += THREADS)
            schur_block_slow(M, k, a, featPre, posePre, Vinv, dvec, split, keys, rowPtr, S, E);
        return;
    }
    const int gbase = posePre[k];

    // ---- dense poses: slots, row exponents, S slots of their pairs ----
    if (tid < 16) {
        // tid-th set bit of denseMask
        unsigned m = denseMask;
        for (int q = 0; q < tid; q++) m &= m - 1;
        dslot[tid] = (tid < H) ? (__ffs(m) - 1) : -1;
    }
    __syncthreads();
    if (tid < 6 * H) dexp[tid] = sexp[6 * (size_t)(gbase + poses[dslot[tid / 6]]) + (tid % 6)];
    const int nDD = H * (H + 1) / 2;
    for (int t = tid; t < nDD; t += THREADS) {
        int r = t, di = 0;
        while (r >= H - di) { r -= H - di; di++; }
        const int dj = di + r;
        const int si = dslot[di], sj = dslot[dj];          // si < sj or equal (slots ascend with d)
        const int idx = si * nposes - (si * (si - 1)) / 2 + (sj - si);
        int slot = -1;
        if (((unsigned)ci[32 + (idx >> 5)] >> (idx & 31)) & 1u) {
            const int gi = poses[si], gj = poses[sj];
            slot = find_slot(keys, rowPtr, gbase + gi, pair_key(k, gi, gj));
        }
        ddSlot[t] = slot;
    }
    // ---- sparse pairs (at least one pose outside the dense set) that occur in the chunk ----
    const int npairsAll = nposes * (nposes + 1) / 2;
    for (int t = tid; t < npairsAll; t += THREADS) {
        if (!(((unsigned)ci[32 + (t >> 5)] >> (t & 31)) & 1u)) continue;
        int r = t, i = 0;
        while (r >= nposes - i) { r -= nposes - i; i++; }
        const int j = i + r;
        if (((denseMask >> i) & 1u) && ((denseMask >> j) & 1u)) continue;
        spairs[atomicAdd(&misc[0], 1)] = (unsigned short)((i << 8) | j);
    }
    // ---- super-tiles of the dense part: 16 x 32 outputs each (2 x 4 tiles of 8 x 8) ----
    const int nT = (6 * H + 7) >> 3;
    if (tid == 0) {
        int n = 0;
        for (int si = 0; 2 * si < nT; si++)
            for (int sj = 0; 4 * sj < nT; sj++)
                if (min(4 * sj + 3, nT - 1) >= 2 * si) stl[n++] = (si << 8) | sj;
        misc[1] = n;
    }
    __syncthreads();
    const int nSparse = misc[0];
    const int nST = misc[1];
    const int nKS = (nST > 0 && nST <= NW) ? NW / nST : 1;
    const int nTasks = nST * nKS;                   // <= NW by construction (HCAP)

    // the warp's task and, per lane, the rows of its 2 A fragments and 4 B fragments, packed as
    // (local pose slot | 3 * row-in-block << 8); -1: padding row beyond 6H
    int aInfo[2], bInfo[4], tMask = 0, tSlice = 0;
    double acc[8][2];
#pragma unroll
    for (int a = 0; a < 2; a++) aInfo[a] = -1;
#pragma unroll
    for (int b = 0; b < 4; b++) bInfo[b] = -1;
#pragma unroll
    for (int q = 0; q < 8; q++) { acc[q][0] = 0.0; acc[q][1] = 0.0; }
    if (warp < nTasks) {
        const int st = stl[warp % nST];
        tSlice = warp / nST;
        const int si = st >> 8, sj = st & 255;
#pragma unroll
        for (int a = 0; a < 2; a++) {
            const int row = (2 * si + a) * 8 + (lane >> 2);
            if (row < 6 * H) aInfo[a] = dslot[row / 6] | (((row % 6) * 3) << 8);
        }
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int col = (4 * sj + b) * 8 + (lane >> 2);
            if (col < 6 * H) bInfo[b] = dslot[col / 6] | (((col % 6) * 3) << 8);
        }
#pragma unroll
        for (int a = 0; a < 2; a++)
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int ti = 2 * si + a, tj = 4 * sj + b;
                if (ti < nT && tj < nT && tj >= ti) tMask |= 1 << (a * 4 + b);
            }
    }
    // the thread's sparse pair (dealt from the last thread downwards: the first warps carry the
    // dense tasks) and its E row
    int pi = -1, pj = -1;
    {
        const int sp = THREADS - 1 - tid;
        if (sp < nSparse) { const int v = spairs[sp]; pi = v >> 8; pj = v & 255; }
    }
    double sacc[36];
#pragma unroll
    for (int q = 0; q < 36; q++) sacc[q] = 0.0;
    const int eSlot = tid / 6, eRow = tid - 6 * eSlot;
    const bool eLive = eSlot < nposes;
    const int eSide = (eLive && split && poses[eSlot] >= split[k]) ? 3 : 0;
    double eacc = 0.0;

    const double *Wg = M.W;
    const int *Pg = blkInfo + wPre[k];
    const double *Vg = Vinv + 9 * (size_t)(featPre[k] + ch.f0);
    const double *Eg = dvec + 6 * (size_t)(featPre[k] + ch.f0);

    auto batch_end = [&](int fa) {
        int target = wptr[fa] + MAXBLK;
        int lo = fa + 1, hi = min(nfeat, fa + NBMAX);
        while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (wptr[mid] <= target) lo = mid; else hi = mid - 1; }
        return lo;
    };
    auto issue = [&](int fa, int fe, int buf) {
        const int nbf = fe - fa;
        const int b0 = wptr[fa], nblk = wptr[fe] - b0;
        double *dW = rawW + buf * (MAXBLK * 18);
        const double *sW = Wg + 18 * (size_t)b0;
        for (int c = tid; c < nblk * 9; c += THREADS) cp_async16(dW + 2 * c, sW + 2 * c);
        int *dP = rawPh + buf * MAXBLK;
        for (int c = tid; c < nblk; c += THREADS) cp_async4(dP + c, Pg + b0 + c);
        double *dV = rawVi + buf * (NBMAX * 9 + 1);
        for (int c = tid; c < nbf * 9; c += THREADS) cp_async8(dV + c, Vg + 9 * (size_t)fa + c);
        double *dE = rawEf + buf * (NBMAX * 6);
        for (int c = tid; c < nbf * 6; c += THREADS) cp_async8(dE + c, Eg + 6 * (size_t)fa + c);
        cp_async_commit();
    };

    int fa = 0, fe = batch_end(0);
    issue(fa, fe, 0);
    for (int bi = 0; fa < nfeat; bi++) {
        const int buf = bi & 1;
        const int nbf = fe - fa;
        const int b0 = wptr[fa], nblk = wptr[fe] - b0;
        const int fa2 = fe, fe2 = (fa2 < nfeat) ? batch_end(fa2) : fa2;
        cp_async_wait_all();
        __syncthreads();                               // this batch landed; previous batch's readers are done
        if (fa2 < nfeat) issue(fa2, fe2, buf ^ 1);
        for (int i = tid; i < NBMAX * 8; i += THREADS) reinterpret_cast<unsigned *>(blkOf)[i] = 0xffffffffu;
        if (tid < NBMAX) present[tid] = 0u;
        __syncthreads();
        // W V^-1 rows (one thread per (block, row)) and the (feature, pose) -> block table of the batch
        const double *rW = rawW + buf * (MAXBLK * 18);
        const int *rP = rawPh + buf * MAXBLK;
        const double *rV = rawVi + buf * (NBMAX * 9 + 1);
        for (int e = tid; e < nblk * 6; e += THREADS) {
            int blk = e / 6, r = e - 6 * blk;
            int info = rP[blk];
            int fb = (info >> 8) - fa, slot = info & 255;
            const double *wr = rW + 18 * blk + 3 * r;
            const double *vi = rV + 9 * fb;
            double w0_ = wr[0], w1_ = wr[1], w2_ = wr[2];
            double *dv = WVsm + blk * LD + 3 * r;
            dv[0] = w0_ * vi[0] + w1_ * vi[1] + w2_ * vi[2];
            dv[1] = w0_ * vi[3] + w1_ * vi[4] + w2_ * vi[5];
            dv[2] = w0_ * vi[6] + w1_ * vi[7] + w2_ * vi[8];
            if (r == 0) {
                blkOf[fb * 32 + slot] = (unsigned char)blk;
                atomicOr(&present[fb], 1u << slot);
            }
        }
        __syncthreads();
        const double *rE = rawEf + buf * (NBMAX * 6);

        // ---- dense part: DMMA over (feature, column) pairs, three k-steps per four features ----
        if (tMask != 0) {                                              // warp-uniform
            const int ngroups = (nbf + 3) >> 2;
            const int lk = lane & 3;
            for (int g = tSlice; g < ngroups; g += nKS) {
#pragma unroll
                for (int step = 0; step < 3; step++) {
                    const int kk = 4 * step + lk;                      // 0..11 within the group
                    const int fo = (kk * 11) >> 5;                     // kk / 3
                    const int x = kk - 3 * fo;
                    const unsigned char *bo = blkOf + (4 * g + fo) * 32;
                    double av[2], bv[4];
#pragma unroll
                    for (int a = 0; a < 2; a++) {
                        av[a] = 0.0;
                        if (aInfo[a] >= 0) {
                            const int blk = bo[aInfo[a] & 255];
                            if (blk != 255) av[a] = rW[blk * 18 + (aInfo[a] >> 8) + x];
                        }
                    }
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        bv[b] = 0.0;
                        if (bInfo[b] >= 0) {
                            const int blk = bo[bInfo[b] & 255];
                            if (blk != 255) bv[b] = WVsm[blk * LD + (bInfo[b] >> 8) + x];
                        }
                    }
#pragma unroll
                    for (int a = 0; a < 2; a++)
#pragma unroll
                        for (int b = 0; b < 4; b++)
                            if (tMask & (1 << (a * 4 + b)))            // warp-uniform
                                dmma(acc[a * 4 + b][0], acc[a * 4 + b][1], av[a], bv[b]);
                }
            }
        }
        // ---- sparse pairs: every lane walks its own list of features that see both poses ----
        {
            const int mi = pi < 0 ? 0 : pi, mj = pi < 0 ? 0 : pj;
            auto next_hit = [&](int fb) {
                while (fb < nbf) {
                    const unsigned pr = present[fb];
                    if ((pr >> mi) & (pr >> mj) & 1u) break;
                    fb++;
                }
                return fb;
            };
            int fb = pi < 0 ? nbf : next_hit(0);
            while (__any_sync(0xffffffffu, fb < nbf)) {
                if (fb < nbf) {
                    const double2 *wv2 = reinterpret_cast<const double2 *>(WVsm + (int)blkOf[fb * 32 + pi] * LD);
                    const double2 *w2 = reinterpret_cast<const double2 *>(rW + (int)blkOf[fb * 32 + pj] * 18);
                    double b[18];
#pragma unroll
                    for (int q = 0; q < 9; q++) { double2 v = w2[q]; b[2 * q] = v.x; b[2 * q + 1] = v.y; }
                    const int cmin = (pi == pj) ? 1 : 0;               // diagonal pair: upper triangle only
#pragma unroll
                    for (int rp = 0; rp < 3; rp++) {
                        const double2 x0 = wv2[3 * rp], x1 = wv2[3 * rp + 1], x2 = wv2[3 * rp + 2];
                        const double av[6] = {x0.x, x0.y, x1.x, x1.y, x2.x, x2.y};
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            const int r = 2 * rp + h;
                            const double a0 = av[3 * h], a1 = av[3 * h + 1], a2 = av[3 * h + 2];
#pragma unroll
                            for (int c = 0; c < 6; c++)
                                if (c >= r * cmin)
                                    sacc[6 * r + c] = fma(a2, b[3 * c + 2], fma(a1, b[3 * c + 1], fma(a0, b[3 * c], sacc[6 * r + c])));
                        }
                    }
                    fb = next_hit(fb + 1);
                }
            }
        }
        // ---- E: thread per (local pose, row) ----
        if (eLive) {
            for (int fb = 0; fb < nbf; fb++) {
                if (!((present[fb] >> eSlot) & 1u)) continue;
                const double *wr = rW + (int)blkOf[fb * 32 + eSlot] * 18 + 3 * eRow;
                const double *d = rE + 6 * fb + eSide;
                eacc = fma(wr[2], d[2], fma(wr[1], d[1], fma(wr[0], d[0], eacc)));
            }
        }
        fa = fa2; fe = fe2;
    }
    __syncthreads();                                    // everybody is done with the raw buffers

    // ---- flush ----
    if (eLive) Erec[6 * (32 * (size_t)blockIdx.x + eSlot) + eRow] = eacc;
    // dense accumulators -> staging [task][tile][8][8]
    double *stage = rawW;
    if (warp < nTasks) {
#pragma unroll
        for (int q = 0; q < 8; q++)
            if (tMask & (1 << q))
                *reinterpret_cast<double2 *>(stage + ((size_t)warp * 8 + q) * 64 + (lane >> 2) * 8 + 2 * (lane & 3)) =
                    make_double2(acc[q][0], acc[q][1]);
    }
    __syncthreads();
    // dense pairs: every (pair, entry) sums its k-slices in a fixed order; fixed-point integer atomics
    for (int e = tid; e < nDD * 36; e += THREADS) {
        const int t = e / 36, q = e - 36 * t;
        const int slot = ddSlot[t];
        if (slot < 0) continue;
        int rr = t, di = 0;
        while (rr >= H - di) { rr -= H - di; di++; }
        const int dj = di + rr;
        const int r = q / 6, c = q - 6 * r;
        if (di == dj && c < r) continue;                 // mirrored from (c, r)
        const int row = 6 * di + r, col = 6 * dj + c;
        const int ti = row >> 3, tj = col >> 3;
        const int si = ti >> 1, sj = tj >> 2;
        int stIdx = 0;
        while (stl[stIdx] != ((si << 8) | sj)) stIdx++;
        const int tile = (ti & 1) * 4 + (tj & 3);
        double v = 0.0;
        for (int sl = 0; sl < nKS; sl++)
            v += stage[((size_t)(sl * nST + stIdx) * 8 + tile) * 64 + (row & 7) * 8 + (col & 7)];
        const long long fx = __double2ll_rn(v * pow2(fx_shift(dexp[6 * di + r], dexp[6 * dj + c])));
        unsigned long long *sp = reinterpret_cast<unsigned long long *>(Sfx) + 36 * (size_t)slot;
        atomicAdd(sp + 6 * r + c, (unsigned long long)fx);
        if (di == dj && c > r) atomicAdd(sp + 6 * c + r, (unsigned long long)fx);
    }
    // sparse pairs: straight from the registers
    if (pi >= 0) {
        const int gi = poses[pi], gj = poses[pj];
        const int slot = find_slot(keys, rowPtr, gbase + gi, pair_key(k, gi, gj));
        unsigned long long *sp = reinterpret_cast<unsigned long long *>(Sfx) + 36 * (size_t)slot;
        int ei[6], ej[6];
#pragma unroll
        for (int q = 0; q < 6; q++) {
            ei[q] = sexp[6 * (size_t)(gbase + gi) + q];
            ej[q] = sexp[6 * (size_t)(gbase + gj) + q];
        }
#pragma unroll
        for (int r = 0; r < 6; r++)
#pragma unroll
            for (int c = 0; c < 6; c++) {
                if (pi == pj && c < r) continue;
                const long long fx = __double2ll_rn(sacc[6 * r + c] * pow2(fx_shift(ei[r], ej[c])));
                atomicAdd(sp + 6 * r + c, (unsigned long long)fx);
                if (pi == pj && c > r) atomicAdd(sp + 6 * c + r, (unsigned long long)fx);
            }
    }
}

} // namespace schur_dense
