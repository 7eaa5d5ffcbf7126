// k_schur_pipe: register-tiled, software-pipelined Schur complement kernel (included by solve.cu).
//
//   S(p,q) -= sum_f W_pf V_f^-1 W_qf^T ,  E_p -= sum_f W_pf d_f      (LinearSFMImp.cpp:2246-2332;
//   d_f = V_f^-1 eF_f - xhat_f of the pose's side of the join, see k_vinv)
//
// One CTA per chunk of SCH_FCHUNK consecutive features of one join.  The <= CMAX distinct poses the
// chunk touches have local indices; the pattern kernel (k_pat_chunk, pass 0) already built that
// table and left it in chunkInfo (pose list) and blkInfo (per W block: feature-in-chunk << 8 | local
// pose), so the prologue here is two small loads.  Each thread owns two pose-pair slots and keeps
// their 6x6 blocks in REGISTERS for the whole chunk; the chunk's features stream through shared
// memory in batches of NB:
//     raw stage   : cp.async (LDGSTS) of the batch's contiguous W blocks / block infos / V^-1 / eF
//                   into a DOUBLE-BUFFERED raw area -- issued one batch ahead, so HBM latency
//                   overlaps the arithmetic of the previous batch;
//     re-layout   : one thread per (block, row): the row of W V^-1 -> a dense tile (W itself stays in
//                   the raw buffer; both operands are read with 16-byte shared loads);
//     pair update : thread (i,j) adds W V^-1|_i * W^T|_j for every feature that sees both poses.
// One flush per touched pair and chunk: 36 fixed-point integer atomics into the S block (exact, hence
// independent of the order in which the chunks arrive; scale from the diagonal of U, see solve.cu), and
// for the diagonal pairs the pose's share of E as a 6-vector record that k_e_gather adds up per pose
// in a fixed order -- no FP64 atomics, results are bit-identical from run to run.  Three instantiations trade registers /
// shared memory for resident CTAs: (CMAX 8, 64 thr) x4-5 per SM for the lower tree levels,
// (16, 128 thr) x2, (31, 256 thr) x1 for the top levels.
#pragma once

namespace schur_pipe {

constexpr int LD = 18;           // dense 144-byte rows: operands are read as nine 16-byte loads
// lower-triangle entries of a diagonal pair's accumulator block that carry its share of E
__device__ constexpr int EIDX[6] = {6, 12, 13, 18, 19, 20};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::); }

// Batches are sized by a BLOCK budget, not a feature count: a batch takes consecutive features
// while their blocks fit into MAXBLK (a feature has <= CMAX <= MAXBLK blocks) and up to NBMAX features.  The lower tree levels (2-4 blocks per feature) therefore run a
// chunk in 4-6 batches instead of 16, the top levels (~14 blocks per feature) in ~8.
template <int CMAX, int MAXBLK, int NBMAX>
struct Layout {
    // all offsets in bytes, 16-byte aligned where cp.async needs it
    static constexpr int rawW = 0;                                   // [2][MAXBLK*18] double
    static constexpr int rawVi = rawW + 2 * MAXBLK * 18 * 8;         // [2][NBMAX*9+1] double (+1: 16B pad)
    static constexpr int rawEf = rawVi + 2 * (NBMAX * 9 + 1) * 8;    // [2][NBMAX*6] double (d vectors)
    static constexpr int rawPh = rawEf + 2 * (NBMAX * 6) * 8;        // [2][MAXBLK] int (block infos)
    static constexpr int WVsm = (rawPh + 2 * MAXBLK * 4 + 15) / 16 * 16;   // [MAXBLK][LD] double: W V^-1 (W itself is read from the raw buffer)
    static constexpr int present = WVsm + MAXBLK * LD * 8;           // [2][NBMAX] unsigned (by raw buffer)
    static constexpr int poses = present + 2 * NBMAX * 4;            // [CMAX+1] int
    static constexpr int wptr = poses + (CMAX + 1 + 3) / 4 * 16;     // [SCH_FCHUNK+1] int
    static constexpr int blkOf = wptr + (SCH_FCHUNK + 1 + 3) / 4 * 16;   // [NBMAX][32] unsigned char
    static constexpr int end = blkOf + NBMAX * 32;
    static size_t bytes() { return (size_t)end + 16; }
};

template <int CMAX, int MAXBLK, int NBMAX, int THREADS, int SLOTS>
__global__ void __launch_bounds__(THREADS, (SLOTS == 1 ? 512 : 256) / THREADS)
k_schur_pipe(const DMap *__restrict__ J, const FeatChunk *__restrict__ chunks,
             const int *__restrict__ chunkInfo, const int *__restrict__ blkInfo, int pat_cmax,
             const int *__restrict__ wPre, const int *__restrict__ featPre, const int *__restrict__ posePre,
             const double *__restrict__ Vinv, const double *__restrict__ dvec, const int *__restrict__ split,
             const u64 *__restrict__ keys, const int *__restrict__ rowPtr,
             double *__restrict__ S, double *__restrict__ E,
             const int *__restrict__ sexp, long long *__restrict__ Sfx, double *__restrict__ Erec,
             const SlowFx slow)
{
    static_assert(MAXBLK >= 2 * CMAX && MAXBLK <= 255, "block budget");
    static_assert(SLOTS == 1, "one pair slot per thread");
    typedef Layout<CMAX, MAXBLK, NBMAX> L;
    constexpr int QBLK = MAXBLK;           // a batch ends at the last feature whose blocks still fit
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
    const FeatChunk ch = chunks[blockIdx.x];
    const DMap &M = J[ch.k];
    const int k = ch.k;
    const int tid = threadIdx.x;
    const int nfeat = ch.f1 - ch.f0;
    const int *ci = chunkInfo + CHUNK_INFO_INTS * (size_t)blockIdx.x;

    for (int i = tid; i <= nfeat; i += THREADS) wptr[i] = M.wPtr[ch.f0 + i];
    if (tid < CMAX) poses[tid] = ci[tid];
    const int nposes = ci[31];
    __syncthreads();
    const int w0 = wptr[0], w1 = wptr[nfeat];
    if (nposes == 0) return;
    if (nposes > CMAX || nposes > pat_cmax) {   // not expected: the host picks CMAX from the measured maximum
        for (int a = w0 + tid; a < w1; a += THREADS)
            schur_block_slow(M, k, a, featPre, posePre, Vinv, dvec, split, keys, rowPtr, slow);
        return;
    }
    const int npairs = nposes * (nposes + 1) / 2;
    // Replicas: when the chunk has few pairs, rep threads share one pair (features dealt round-robin).
    // rep is a power of two and the replicas of a pair sit in ADJACENT lanes, so that their partial
    // blocks are combined by a fixed shuffle butterfly before the flush (no atomics, same bits every run).
    int rep = 1;
    while (2 * rep <= min(min(32, NBMAX), THREADS / npairs)) rep *= 2;
    // pair slots: the nposes diagonal pairs first (so they share warps: they run a different body),
    // then the off-diagonal pairs (i<j) row by row
    int pi[SLOTS], pj[SLOTS], pr0[SLOTS];
#pragma unroll
    for (int u = 0; u < SLOTS; u++) {
        pi[u] = -1; pj[u] = -1; pr0[u] = 0;
        int q = tid + u * THREADS;
        int t = q / rep, r = q - t * rep;
        if (t < npairs) {
            pr0[u] = r;
            if (t < nposes) { pi[u] = t; pj[u] = t; }
            else {
                t -= nposes;
                int i = 0;
                while (t >= nposes - 1 - i) { t -= nposes - 1 - i; i++; }
                pi[u] = i; pj[u] = i + 1 + t;
            }
        }
    }
    // acc: the 6x6 block.  Diagonal pairs only accumulate the upper triangle (the block is
    // symmetric) and keep their share of E in six otherwise unused lower-triangle entries.
    double acc[SLOTS][36];
#pragma unroll
    for (int u = 0; u < SLOTS; u++) {
#pragma unroll
        for (int q = 0; q < 36; q++) acc[u][q] = 0.0;
    }
    bool touched[SLOTS];
    int sideOff[SLOTS];          // diagonal pairs: which of the feature's two d vectors (End / Cur side)
#pragma unroll
    for (int u = 0; u < SLOTS; u++) {
        touched[u] = false;
        sideOff[u] = (split && pi[u] >= 0 && poses[pi[u]] >= split[k]) ? 3 : 0;
    }

    const double *Wg = M.W;
    const int *Pg = blkInfo + wPre[k];
    const double *Vg = Vinv + 9 * (size_t)(featPre[k] + ch.f0);
    const double *Eg = dvec + 6 * (size_t)(featPre[k] + ch.f0);

    // end of the batch that starts at feature fa (uniform across the CTA)
    auto batch_end = [&](int fa) {
        int target = wptr[fa] + QBLK;
        int lo = fa + 1, hi = min(nfeat, fa + NBMAX);
        while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (wptr[mid] <= target) lo = mid; else hi = mid - 1; }
        return lo;
    };
    // raw stage of the batch [fa, fe) into buffer `buf` (asynchronous)
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

    for (int i = tid; i < 2 * NBMAX; i += THREADS) present[i] = 0u;
    int fa = 0, fe = batch_end(0);
    issue(fa, fe, 0);
    for (int bi = 0; fa < nfeat; bi++) {
        const int buf = bi & 1;
        const int nbf = fe - fa;
        const int b0 = wptr[fa], nblk = wptr[fe] - b0;
        const int fa2 = fe, fe2 = (fa2 < nfeat) ? batch_end(fa2) : fa2;
        cp_async_wait_all();
        __syncthreads();                               // this batch landed; previous pair update done
        for (int i = tid; i < NBMAX; i += THREADS) present[(buf ^ 1) * NBMAX + i] = 0u;   // for the next batch
        if (fa2 < nfeat) issue(fa2, fe2, buf ^ 1);
        // re-layout raw -> padded rows, one thread per (block, row); the row of W V^-1 on the way
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
                atomicOr(&present[buf * NBMAX + fb], 1u << slot);
            }
        }
        __syncthreads();
        const double *rE = rawEf + buf * (NBMAX * 6);
        {
            // Every lane walks ITS OWN list of features that see both poses of its pair: the search for
            // the next hit is a cheap divergent scan, the 108-FMA update below runs with all lanes that
            // still have work (before: the warp executed the body whenever ANY lane hit, at ~20 % density)
            constexpr int u = 0;
            const unsigned *prs = present + buf * NBMAX;
            const int mi = pi[u] < 0 ? 0 : pi[u], mj = pi[u] < 0 ? 0 : pj[u];
            auto next_hit = [&](int fb) {
                while (fb < nbf) {
                    const unsigned pr = prs[fb];
                    if ((pr >> mi) & (pr >> mj) & 1u) break;
                    fb += rep;
                }
                return fb;
            };
            int fb = pi[u] < 0 ? nbf : next_hit(pr0[u]);
            while (__any_sync(0xffffffffu, fb < nbf)) {
              if (fb < nbf) {
                touched[u] = true;

                const double2 *wv2 = reinterpret_cast<const double2 *>(WVsm + (int)blkOf[fb * 32 + pi[u]] * LD);
                const double2 *w2 = reinterpret_cast<const double2 *>(rW + (int)blkOf[fb * 32 + pj[u]] * 18);
                double b[18];
#pragma unroll
                for (int q = 0; q < 9; q++) { double2 v = w2[q]; b[2 * q] = v.x; b[2 * q + 1] = v.y; }
                if (pi[u] == pj[u]) {
                    const double *ef = rE + 6 * fb + sideOff[u];
                    const double e0 = ef[0], e1 = ef[1], e2 = ef[2];
#pragma unroll
                    for (int rp = 0; rp < 3; rp++) {            // two rows of W V^-1 per three 16-byte loads
                        const double2 x0 = wv2[3 * rp], x1 = wv2[3 * rp + 1], x2 = wv2[3 * rp + 2];
                        const double av[6] = {x0.x, x0.y, x1.x, x1.y, x2.x, x2.y};
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            const int r = 2 * rp + h;
                            const double a0 = av[3 * h], a1 = av[3 * h + 1], a2 = av[3 * h + 2];
#pragma unroll
                            for (int c = r; c < 6; c++)
                                acc[u][6 * r + c] = fma(a2, b[3 * c + 2], fma(a1, b[3 * c + 1], fma(a0, b[3 * c], acc[u][6 * r + c])));
                            acc[u][EIDX[r]] = fma(b[3 * r + 2], e2, fma(b[3 * r + 1], e1, fma(b[3 * r], e0, acc[u][EIDX[r]])));
                        }
                    }
                } else {
#pragma unroll
                    for (int rp = 0; rp < 3; rp++) {
                        const double2 x0 = wv2[3 * rp], x1 = wv2[3 * rp + 1], x2 = wv2[3 * rp + 2];
                        const double av[6] = {x0.x, x0.y, x1.x, x1.y, x2.x, x2.y};
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            const int r = 2 * rp + h;
                            const double a0 = av[3 * h], a1 = av[3 * h + 1], a2 = av[3 * h + 2];
#pragma unroll
                            for (int c = 0; c < 6; c++)     // three chained FMAs (not mul/fma/fma + add)
                                acc[u][6 * r + c] = fma(a2, b[3 * c + 2], fma(a1, b[3 * c + 1], fma(a0, b[3 * c], acc[u][6 * r + c])));
                        }
                    }
                }
                fb = next_hit(fb + rep);
              }
            }
        }
        fa = fa2; fe = fe2;
    }
    // combine the replicas of a pair (adjacent lanes, rep a power of two <= 32): after the butterfly
    // every replica holds the same bits; replica 0 writes the pair's record
    if (rep > 1) {
        for (int o = 1; o < rep; o <<= 1) {
#pragma unroll
            for (int q = 0; q < 36; q++) acc[0][q] += __shfl_xor_sync(0xffffffffu, acc[0][q], o);
        }
    }
    (void)touched;
    if (pi[0] >= 0 && pr0[0] == 0) {
        // the pattern kernel's bitmap of the chunk's pairs (row-major upper triangle incl. diagonal)
        // decides which pairs exist
        const int i = pi[0], j = pj[0];
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
                        const long long v = __double2ll_rn(acc[0][6 * r + c] * pow2(fx_shift(ei[r], ej[c])));
                        atomicAdd(sp + 6 * r + c, (unsigned long long)v);
                        if (c > r) atomicAdd(sp + 6 * c + r, (unsigned long long)v);
                    }
                // the pose's share of E from this chunk: a record, gathered per pose by k_e_gather
                double *e = Erec + 6 * (32 * (size_t)blockIdx.x + i);
#pragma unroll
                for (int q = 0; q < 6; q++) e[q] = acc[0][EIDX[q]];
            } else {
#pragma unroll
                for (int r = 0; r < 6; r++)
#pragma unroll
                    for (int c = 0; c < 6; c++) {
                        const long long v = __double2ll_rn(acc[0][6 * r + c] * pow2(fx_shift(ei[r], ej[c])));
                        atomicAdd(sp + 6 * r + c, (unsigned long long)v);
                    }
            }
        }
    }
}

} // namespace schur_pipe
