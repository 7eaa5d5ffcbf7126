// k_schur_dense: Schur complement of the reduced camera system with the dense part of every feature
// chunk on the FP64 tensor cores (included by solve.cu).
//
//   S(p,q) -= sum_f W_pf V_f^-1 W_qf^T ,  E_p -= sum_f W_pf d_f          (LinearSFMImp.cpp:2246-2332)
//
// One CTA per chunk of SCH_FCHUNK consecutive features of one join.  The chunk's W blocks stream through
// shared memory in batches (block budget MAXBLK, at most 32 features), double buffered.
//
// Warp specialisation, producer / consumer over the two stage buffers with named barriers:
//   * STAGE warps (the last NWS): cp.async the batch's contiguous W span / block infos / V^-1 / d vectors,
//     form the W V^-1 rows and the batch's (feature, pose) -> block table, signal READY[buf]; wait for
//     FREE[buf] before refilling it.  Loads of batch b+1 are in flight while batch b is multiplied.
//   * COMPUTE warps wait for READY[buf], multiply, signal FREE[buf].  They never wait for memory and
//     never meet the stage warps at a CTA-wide barrier inside the loop.
// Who multiplies what:
//   * The chunk's DENSE local poses (seen by at least half of its features; picked by the pattern
//     kernel) are the former frame origins every feature of the chunk is linked to after a few
//     Transforms: their pose pairs are ~75 % of all (pair, feature) products at the top
//     of the tree and nearly all of them at the bottom.  For them the update is one dense contraction
//     S_DD -= A B^T,  A = [W_df] (6H x 3F), B = [W_df V_f^-1] (6H x 3F), absent (pose, feature)
//     combinations being zero.  It runs as mma.sync.m8n8k4.f64 (DMMA): a task is a 16 x 32 super-tile of
//     S_DD (2 x 4 accumulator tiles in registers) times a slice of the batch's features; k runs over
//     (feature, column) pairs without padding (three k-steps per four features); the operand fragments
//     are read straight from the block-major stage buffers through the (feature, pose) -> block table --
//     no re-layout pass, no zero fill.  Super-tiles with many active tiles get more feature slices, so
//     that the compute warps carry about the same number of DMMAs.
//   * Pairs with a sparse pose: one 6x6 accumulator block per thread, every lane walking its own list
//     of features that see both poses.
//   * E: one thread per (local pose, row).
// Flush: fixed-point integer atomics into S (exact, order-independent; solve.cu) and one 6-vector
// record per (chunk, local pose) for E.
#pragma once

namespace schur_dense {

using schur_pipe::cp_async16;
using schur_pipe::cp_async8;
using schur_pipe::cp_async4;
using schur_pipe::cp_async_commit;

constexpr int LD = 18;
constexpr int NB = 32;            // features per batch at most (eight groups of four)

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

template <int CMAX, int MAXBLK, int THREADS>
struct Layout {
    static constexpr int NW = THREADS / 32;
    static constexpr int stageBytes = NW * 8 * 64 * 8;               // accumulator staging of the flush
    static constexpr int rawWBytes = 2 * MAXBLK * 18 * 8;
    static constexpr int rawW = 0;                                   // [2][MAXBLK*18] double (later: staging)
    static constexpr int rawVi = rawW + (rawWBytes > stageBytes ? rawWBytes : stageBytes);   // [2][NB*9+1] double
    static constexpr int rawEf = rawVi + 2 * (NB * 9 + 1) * 8;       // [2][NB*6] double (d vectors)
    static constexpr int rawPh = rawEf + 2 * (NB * 6) * 8;           // [2][MAXBLK] int (block infos)
    static constexpr int WVsm = (rawPh + 2 * MAXBLK * 4 + 15) / 16 * 16;   // [2][MAXBLK][LD] double: W V^-1
    static constexpr int present = WVsm + 2 * MAXBLK * LD * 8;       // [2][NB] unsigned
    static constexpr int blkOf = present + 2 * NB * 4;               // [2][NB][32] unsigned char (255: absent)
    static constexpr int poses = blkOf + 2 * NB * 32;                // [32] int
    static constexpr int wptr = poses + 32 * 4;                      // [SCH_FCHUNK+4] int
    static constexpr int bend = wptr + (SCH_FCHUNK + 4) * 4;         // [SCH_FCHUNK+4] int: first feature of every batch
    static constexpr int dslot = bend + (SCH_FCHUNK + 4) * 4;        // [16] int: dense index -> local pose slot
    static constexpr int dexp = dslot + 16 * 4;                      // [16][6] int: row exponents of the dense poses
    static constexpr int ddSlot = dexp + 16 * 6 * 4;                 // [136] int: S slot of every dense pair (-1: not in the pattern)
    static constexpr int spairs = ddSlot + 136 * 4;                  // [496] unsigned short: sparse pairs (i << 8 | j)
    static constexpr int stl = spairs + 496 * 2;                     // [16] int: active super-tiles (si << 8 | sj)
    static constexpr int stFirst = stl + 16 * 4;                     // [16] int: first task of the super-tile
    static constexpr int stN = stFirst + 16 * 4;                     // [16] int: its number of feature slices
    static constexpr int misc = stN + 16 * 4;                        // [4] int
    static constexpr int end = misc + 16;
    static size_t bytes() { return (size_t)end + 16; }
};

// NWS stage warps, the rest compute.  HCAP: at most that many dense poses (more are treated as sparse)
// so that the dense part has no more super-tiles than there are compute warps.
template <int CMAX, int MAXBLK, int THREADS, int NWS, int HCAP>
__global__ void __launch_bounds__(THREADS, 512 / THREADS)
k_schur_dense(const DMap *__restrict__ J, const FeatChunk *__restrict__ chunks,
              const int *__restrict__ chunkInfo, const int *__restrict__ blkInfo, int pat_cmax,
              const int *__restrict__ wPre, const int *__restrict__ featPre, const int *__restrict__ posePre,
              const double *__restrict__ Vinv, const double *__restrict__ dvec, const int *__restrict__ split,
              const u64 *__restrict__ keys, const int *__restrict__ rowPtr,
              double *__restrict__ S, double *__restrict__ E,
              const int *__restrict__ sexp, long long *__restrict__ Sfx, double *__restrict__ Erec,
             const SlowFx slow)
{
    static_assert(MAXBLK >= 2 * CMAX && MAXBLK <= 254, "block budget");
    typedef Layout<CMAX, MAXBLK, THREADS> L;
    constexpr int NW = THREADS / 32, NWC = NW - NWS;
    constexpr int TC = 32 * NWC, TS = 32 * NWS;          // compute / stage threads
    constexpr int BAR_READY = 1, BAR_FREE = 3, BAR_STAGE = 5;
    extern __shared__ __align__(16) unsigned char smraw[];
    double *rawW = (double *)(smraw + L::rawW);
    double *rawVi = (double *)(smraw + L::rawVi);
    double *rawEf = (double *)(smraw + L::rawEf);
    int *rawPh = (int *)(smraw + L::rawPh);
    double *WVsm = (double *)(smraw + L::WVsm);
    unsigned *present = (unsigned *)(smraw + L::present);
    unsigned char *blkOf = (unsigned char *)(smraw + L::blkOf);
    int *poses = (int *)(smraw + L::poses);
    int *wptr = (int *)(smraw + L::wptr);
    int *bend = (int *)(smraw + L::bend);
    int *dslot = (int *)(smraw + L::dslot);
    int *dexp = (int *)(smraw + L::dexp);
    int *ddSlot = (int *)(smraw + L::ddSlot);
    unsigned short *spairs = (unsigned short *)(smraw + L::spairs);
    int *stl = (int *)(smraw + L::stl);
    int *stFirst = (int *)(smraw + L::stFirst);
    int *stN = (int *)(smraw + L::stN);
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
    unsigned denseMask = (unsigned)ci[48];
    int H = ci[49];
    while (H > HCAP) { denseMask &= ~(1u << (31 - __clz(denseMask))); H--; }
    if (tid == 0) misc[0] = 0;
    __syncthreads();
    const int w0 = wptr[0], w1 = wptr[nfeat];
    if (nposes == 0) return;
    if (nposes > CMAX || nposes > pat_cmax) {   // not expected: the host picks CMAX from the measured maximum
        for (int a = w0 + tid; a < w1; a += THREADS)
            schur_block_slow(M, k, a, featPre, posePre, Vinv, dvec, split, keys, rowPtr, slow);
        return;
    }
    const int gbase = posePre[k];

    // ---- prologue: batches, dense poses, sparse pairs, tasks ----
    if (tid == 0) {
        // batch boundaries: consecutive features while their blocks fit MAXBLK, at most NB features
        int nb = 0, fa = 0;
        while (fa < nfeat) {
            bend[nb++] = fa;
            int fe = fa + 1;
            while (fe < nfeat && fe - fa < NB && wptr[fe + 1] - wptr[fa] <= MAXBLK) fe++;
            fa = fe;
        }
        bend[nb] = nfeat;
        misc[2] = nb;
    }
    if (tid >= 32 && tid < 48) {
        const int d = tid - 32;                 // d-th set bit of denseMask
        unsigned m = denseMask;
        for (int q = 0; q < d; q++) m &= m - 1;
        dslot[d] = (d < H) ? (__ffs(m) - 1) : -1;
    }
    const int nT = (6 * H + 7) >> 3;            // 8-row tiles of the dense part
    if (tid == 64) {
        // active 16 x 32 super-tiles (those holding a tile on or above the block diagonal) and their
        // numbers of active tiles; the compute warps are dealt out as feature slices, most-loaded first
        int n = 0, cnt[16], ns[16];
        for (int si = 0; 2 * si < nT; si++)
            for (int sj = 0; 4 * sj < nT; sj++)
                if (min(4 * sj + 3, nT - 1) >= 2 * si) {
                    int c = 0;
                    for (int a = 0; a < 2; a++)
                        for (int b = 0; b < 4; b++) {
                            const int ti = 2 * si + a, tj = 4 * sj + b;
                            if (ti < nT && tj < nT && tj >= ti) c++;
                        }
                    cnt[n] = c; ns[n] = 1;
                    stl[n++] = (si << 8) | sj;
                }
        int spare = NWC - n;
        while (spare > 0 && n > 0) {
            int best = 0;
            for (int i = 1; i < n; i++)
                if (cnt[i] * ns[best] > cnt[best] * ns[i]) best = i;       // largest tiles per slice
            if (ns[best] >= 8) break;
            ns[best]++; spare--;
        }
        int first = 0;
        for (int i = 0; i < n; i++) { stFirst[i] = first; stN[i] = ns[i]; first += ns[i]; }
        misc[1] = n;
        misc[3] = first;
    }
    __syncthreads();
    if (tid < 6 * H) dexp[tid] = sexp[6 * (size_t)(gbase + poses[dslot[tid / 6]]) + (tid % 6)];
    const int nDD = H * (H + 1) / 2;
    for (int t = tid; t < nDD; t += THREADS) {
        int r = t, di = 0;
        while (r >= H - di) { r -= H - di; di++; }
        const int dj = di + r;
        const int si = dslot[di], sj = dslot[dj];          // slots ascend with the dense index
        const int idx = si * nposes - (si * (si - 1)) / 2 + (sj - si);
        int slot = -1;
        if (((unsigned)ci[32 + (idx >> 5)] >> (idx & 31)) & 1u) {
            const int gi = poses[si], gj = poses[sj];
            slot = find_slot(keys, rowPtr, gbase + gi, pair_key(k, gi, gj));
        }
        ddSlot[t] = slot;
    }
    const int npairsAll = nposes * (nposes + 1) / 2;
    for (int t = tid; t < npairsAll; t += THREADS) {
        if (!(((unsigned)ci[32 + (t >> 5)] >> (t & 31)) & 1u)) continue;
        int r = t, i = 0;
        while (r >= nposes - i) { r -= nposes - i; i++; }
        const int j = i + r;
        if (((denseMask >> i) & 1u) && ((denseMask >> j) & 1u)) continue;
        spairs[atomicAdd(&misc[0], 1)] = (unsigned short)((i << 8) | j);
    }
    __syncthreads();
    const int nSparse = misc[0];
    const int nST = misc[1];
    const int nBatches = misc[2];
    const int nTasks = misc[3];

    const double *Wg = M.W;
    const int *Pg = blkInfo + wPre[k];
    const double *Vg = Vinv + 9 * (size_t)(featPre[k] + ch.f0);
    const double *Eg = dvec + 6 * (size_t)(featPre[k] + ch.f0);

    if (warp >= NWC) {
        // =========================== STAGE warps ===========================
        const int st = tid - TC;
        auto issue = [&](int b) {
            const int buf = b & 1;
            const int fa = bend[b], fe = bend[b + 1];
            const int nbf = fe - fa;
            const int b0 = wptr[fa], nblk = wptr[fe] - b0;
            double *dW = rawW + buf * (MAXBLK * 18);
            const double *sW = Wg + 18 * (size_t)b0;
            for (int c = st; c < nblk * 9; c += TS) cp_async16(dW + 2 * c, sW + 2 * c);
            int *dP = rawPh + buf * MAXBLK;
            for (int c = st; c < nblk; c += TS) cp_async4(dP + c, Pg + b0 + c);
            double *dV = rawVi + buf * (NB * 9 + 1);
            for (int c = st; c < nbf * 9; c += TS) cp_async8(dV + c, Vg + 9 * (size_t)fa + c);
            double *dE = rawEf + buf * (NB * 6);
            for (int c = st; c < nbf * 6; c += TS) cp_async8(dE + c, Eg + 6 * (size_t)fa + c);
            cp_async_commit();
        };
        issue(0);
        if (nBatches > 1) issue(1);
        for (int b = 0; b < nBatches; b++) {
            const int buf = b & 1;
            if (b + 1 < nBatches) cp_async_wait<1>(); else cp_async_wait<0>();
            // tables of this buffer: free since FREE[buf] of batch b-2 (awaited before its refill)
            for (int i = st; i < NB * 8; i += TS) reinterpret_cast<unsigned *>(blkOf + buf * NB * 32)[i] = 0xffffffffu;
            if (st < NB) present[buf * NB + st] = 0u;
            bar_sync(BAR_STAGE, TS);                   // every stage thread's copies have landed; tables reset
            const int fa = bend[b];
            const int nblk = wptr[bend[b + 1]] - wptr[fa];
            const double *rW = rawW + buf * (MAXBLK * 18);
            const int *rP = rawPh + buf * MAXBLK;
            const double *rV = rawVi + buf * (NB * 9 + 1);
            double *wv = WVsm + buf * (MAXBLK * LD);
            unsigned char *bo = blkOf + buf * NB * 32;
            for (int e = st; e < nblk * 6; e += TS) {
                int blk = e / 6, r = e - 6 * blk;
                int info = rP[blk];
                int fb = (info >> 8) - fa, slot = info & 255;
                const double *wr = rW + 18 * blk + 3 * r;
                const double *vi = rV + 9 * fb;
                double w0_ = wr[0], w1_ = wr[1], w2_ = wr[2];
                double *dv = wv + blk * LD + 3 * r;
                dv[0] = w0_ * vi[0] + w1_ * vi[1] + w2_ * vi[2];
                dv[1] = w0_ * vi[3] + w1_ * vi[4] + w2_ * vi[5];
                dv[2] = w0_ * vi[6] + w1_ * vi[7] + w2_ * vi[8];
                if (r == 0) {
                    bo[fb * 32 + slot] = (unsigned char)blk;
                    atomicOr(&present[buf * NB + fb], 1u << slot);
                }
            }
            __threadfence_block();
            bar_arrive(BAR_READY + buf, THREADS);      // batch b is ready for the compute warps
            if (b + 2 < nBatches) {
                bar_sync(BAR_FREE + buf, THREADS);     // compute warps are done with batch b
                issue(b + 2);
            }
        }
    } else {
        // =========================== COMPUTE warps ===========================
        // the warp's dense task: super-tile + feature slice; per lane the rows of its 2 A and 4 B
        // fragments, packed as (local pose slot | 3 * row-in-block << 8); -1: padding row beyond 6H
        int aInfo[2], bInfo[4], tMask = 0, tSlice = 0, tNS = 1;
        double acc[8][2];
#pragma unroll
        for (int a = 0; a < 2; a++) aInfo[a] = -1;
#pragma unroll
        for (int b = 0; b < 4; b++) bInfo[b] = -1;
#pragma unroll
        for (int q = 0; q < 8; q++) { acc[q][0] = 0.0; acc[q][1] = 0.0; }
        if (warp < nTasks) {
            int sti = 0;
            while (sti + 1 < nST && stFirst[sti + 1] <= warp) sti++;
            tSlice = warp - stFirst[sti];
            tNS = stN[sti];
            const int stv = stl[sti];
            const int si = stv >> 8, sj = stv & 255;
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
        // the thread's sparse pair (dealt from the last compute thread downwards: the first warps carry
        // the heaviest dense tasks) and its E row
        int pi = -1, pj = -1;
        {
            const int sp = TC - 1 - tid;
            if (sp < nSparse) { const int v = spairs[sp]; pi = v >> 8; pj = v & 255; }
        }
        double sacc[36];
#pragma unroll
        for (int q = 0; q < 36; q++) sacc[q] = 0.0;
        const int eSlot = tid / 6, eRow = tid - 6 * eSlot;
        const bool eLive = eSlot < nposes;
        const int eSide = (eLive && split && poses[eSlot] >= split[k]) ? 3 : 0;
        double eacc = 0.0;

        for (int b = 0; b < nBatches; b++) {
            const int buf = b & 1;
            const int nbf = bend[b + 1] - bend[b];
            bar_sync(BAR_READY + buf, THREADS);
            const double *rW = rawW + buf * (MAXBLK * 18);
            const double *wv = WVsm + buf * (MAXBLK * LD);
            const unsigned char *bo0 = blkOf + buf * NB * 32;
            const unsigned *prs = present + buf * NB;
            const double *rE = rawEf + buf * (NB * 6);
            // ---- dense part: DMMA over (feature, column) pairs, three k-steps per four features ----
            if (tMask != 0) {                                              // warp-uniform
                const int ngroups = (nbf + 3) >> 2;
                const int lk = lane & 3;
                for (int g = tSlice; g < ngroups; g += tNS) {
#pragma unroll
                    for (int step = 0; step < 3; step++) {
                        const int kk = 4 * step + lk;                      // 0..11 within the group
                        const int fo = (kk * 11) >> 5;                     // kk / 3
                        const int x = kk - 3 * fo;
                        const unsigned char *bo = bo0 + (4 * g + fo) * 32;
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
                        for (int bb = 0; bb < 4; bb++) {
                            bv[bb] = 0.0;
                            if (bInfo[bb] >= 0) {
                                const int blk = bo[bInfo[bb] & 255];
                                if (blk != 255) bv[bb] = wv[blk * LD + (bInfo[bb] >> 8) + x];
                            }
                        }
#pragma unroll
                        for (int a = 0; a < 2; a++)
#pragma unroll
                            for (int bb = 0; bb < 4; bb++)
                                if (tMask & (1 << (a * 4 + bb)))            // warp-uniform
                                    dmma(acc[a * 4 + bb][0], acc[a * 4 + bb][1], av[a], bv[bb]);
                    }
                }
            }
            // ---- sparse pairs: every lane walks its own list of features that see both poses ----
            {
                const int mi = pi < 0 ? 0 : pi, mj = pi < 0 ? 0 : pj;
                auto next_hit = [&](int fb) {
                    while (fb < nbf) {
                        const unsigned pr = prs[fb];
                        if ((pr >> mi) & (pr >> mj) & 1u) break;
                        fb++;
                    }
                    return fb;
                };
                int fb = pi < 0 ? nbf : next_hit(0);
                while (__any_sync(0xffffffffu, fb < nbf)) {
                    if (fb < nbf) {
                        const double2 *wv2 = reinterpret_cast<const double2 *>(wv + (int)bo0[fb * 32 + pi] * LD);
                        const double2 *w2 = reinterpret_cast<const double2 *>(rW + (int)bo0[fb * 32 + pj] * 18);
                        double bq[18];
#pragma unroll
                        for (int q = 0; q < 9; q++) { double2 v = w2[q]; bq[2 * q] = v.x; bq[2 * q + 1] = v.y; }
                        const int cmin = (pi == pj) ? 1 : 0;               // diagonal pair: upper triangle only
#pragma unroll
                        for (int rp = 0; rp < 3; rp++) {
                            const double2 x0 = wv2[3 * rp], x1 = wv2[3 * rp + 1], x2 = wv2[3 * rp + 2];
                            const double avv[6] = {x0.x, x0.y, x1.x, x1.y, x2.x, x2.y};
#pragma unroll
                            for (int h = 0; h < 2; h++) {
                                const int r = 2 * rp + h;
                                const double a0 = avv[3 * h], a1 = avv[3 * h + 1], a2 = avv[3 * h + 2];
#pragma unroll
                                for (int c = 0; c < 6; c++)
                                    if (c >= r * cmin)
                                        sacc[6 * r + c] = fma(a2, bq[3 * c + 2], fma(a1, bq[3 * c + 1], fma(a0, bq[3 * c], sacc[6 * r + c])));
                            }
                        }
                        fb = next_hit(fb + 1);
                    }
                }
            }
            // ---- E: thread per (local pose, row) ----
            if (eLive) {
                for (int fb = 0; fb < nbf; fb++) {
                    if (!((prs[fb] >> eSlot) & 1u)) continue;
                    const double *wr = rW + (int)bo0[fb * 32 + eSlot] * 18 + 3 * eRow;
                    const double *d = rE + 6 * fb + eSide;
                    eacc = fma(wr[2], d[2], fma(wr[1], d[1], fma(wr[0], d[0], eacc)));
                }
            }
            if (b + 2 < nBatches) bar_arrive(BAR_FREE + buf, THREADS);
        }
        // ---- flush of what lives in registers: E record, dense accumulators -> staging, sparse pairs ----
        if (eLive) Erec[6 * (32 * (size_t)blockIdx.x + eSlot) + eRow] = eacc;
        bar_sync(BAR_STAGE + 1, TC);                   // every compute warp has left the stage buffers
        double *stage = rawW;
        if (warp < nTasks) {
#pragma unroll
            for (int q = 0; q < 8; q++)
                if (tMask & (1 << q))
                    *reinterpret_cast<double2 *>(stage + ((size_t)warp * 8 + q) * 64 + (lane >> 2) * 8 + 2 * (lane & 3)) =
                        make_double2(acc[q][0], acc[q][1]);
        }
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
    // ---- sparse pairs beyond the compute threads' capacity (chunks whose features see many different
    // pose subsets): one warp per pair straight from global memory, lanes over the features, fixed
    // butterfly -- slow, rare, same result every run ----
    for (int sp = TC + warp; sp < nSparse; sp += NW) {
        const int v = spairs[sp];
        const int si = v >> 8, sj = v & 255;
        double pa[36];
#pragma unroll
        for (int q = 0; q < 36; q++) pa[q] = 0.0;
        for (int f = lane; f < nfeat; f += 32) {
            int bi_ = -1, bj_ = -1;
            for (int a = wptr[f]; a < wptr[f + 1]; a++) {
                const int sl = Pg[a] & 255;
                if (sl == si) bi_ = a;
                if (sl == sj) bj_ = a;
            }
            if (bi_ < 0 || bj_ < 0) continue;
            double Wi[18], Wj[18], Vi[9], WV[18];
            sm::load<18>(Wg + 18 * (size_t)bi_, Wi);
            sm::load<18>(Wg + 18 * (size_t)bj_, Wj);
            sm::load<9>(Vg + 9 * (size_t)f, Vi);
            sm::mmt<6, 3, 3>(Wi, Vi, WV);
#pragma unroll
            for (int r = 0; r < 6; r++)
#pragma unroll
                for (int c = 0; c < 6; c++)
                    pa[6 * r + c] = fma(WV[3 * r + 2], Wj[3 * c + 2], fma(WV[3 * r + 1], Wj[3 * c + 1], fma(WV[3 * r], Wj[3 * c], pa[6 * r + c])));
        }
#pragma unroll
        for (int q = 0; q < 36; q++) pa[q] = sm::warp_sum(pa[q]);
        if (lane == 0) {
            const int gi = poses[si], gj = poses[sj];
            const int slot = find_slot(keys, rowPtr, gbase + gi, pair_key(k, gi, gj));
            unsigned long long *sp_ = reinterpret_cast<unsigned long long *>(Sfx) + 36 * (size_t)slot;
            for (int r = 0; r < 6; r++)
                for (int c = 0; c < 6; c++) {
                    if (si == sj && c < r) continue;
                    const int sh = fx_shift(sexp[6 * (size_t)(gbase + gi) + r], sexp[6 * (size_t)(gbase + gj) + c]);
                    const long long fx = __double2ll_rn(pa[6 * r + c] * pow2(sh));
                    atomicAdd(sp_ + 6 * r + c, (unsigned long long)fx);
                    if (si == sj && c > r) atomicAdd(sp_ + 6 * c + r, (unsigned long long)fx);
                }
        }
    }
    __syncthreads();
    // ---- dense pairs: every (pair, entry) sums its feature slices in a fixed order; fixed-point atomics ----
    const double *stage = rawW;
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
        const int stKey = ((ti >> 1) << 8) | (tj >> 2);
        int sti = 0;
        while (stl[sti] != stKey) sti++;
        const int tile = (ti & 1) * 4 + (tj & 3);
        const int t0 = stFirst[sti], ns = stN[sti];
        double v = 0.0;
        for (int sl = 0; sl < ns; sl++)
            v += stage[((size_t)(t0 + sl) * 8 + tile) * 64 + (row & 7) * 8 + (col & 7)];
        const long long fx = __double2ll_rn(v * pow2(fx_shift(dexp[6 * di + r], dexp[6 * dj + c])));
        unsigned long long *sp = reinterpret_cast<unsigned long long *>(Sfx) + 36 * (size_t)slot;
        atomicAdd(sp + 6 * r + c, (unsigned long long)fx);
        if (di == dj && c > r) atomicAdd(sp + 6 * c + r, (unsigned long long)fx);
    }
}

} // namespace schur_dense
