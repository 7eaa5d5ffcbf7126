// Big fronts of the multifrontal Cholesky (included by solve.cu): fronts with BF_MIN_FS or more scalar
// rows -- loop-closure / dense-overlap systems, where a supernode's front reaches 600 .. 2000 rows --
// are factored by the WHOLE GPU instead of one CTA (replaces cholmod_factorize's dense supernodal
// kernels, LinearSFMImp.cpp:2444).  Right-looking, panel width 48 columns (8 pose blocks):
//
//   k_bf_extend  extend-add of the children's update matrices.  A CTA owns one block column of the
//                parent and walks the children in their fixed order: no two CTAs touch the same entry,
//                no atomics, bit-identical results.
//   k_bf_panel   one CTA per slab of 128 rows below the panel's diagonal block: every CTA factors the
//                48 x 48 diagonal block (redundantly) and solves ITS rows against it in shared memory
//                (the panel routine of the one-CTA kernel, factor_panel).
//   k_bf_syrk    trailing update C -= P P^T (lower triangle + the right-hand-side row) in 64 x 64 tiles on
//                the FP64 tensor cores: mma.sync m8n8k4 (DMMA), a warp owns a 32 x 32 sub-tile as 4 x 4
//                fragments, operands staged k-major in shared memory (stride 68: conflict-free fragment
//                loads), K = 48 in twelve steps.
// The fronts stay in their column-major global layout (leading dimension fs + 1, right-hand side as the
// last row); the small fronts of the same tree level keep the one-CTA-per-front kernel.  All big fronts
// of a level share each launch (blockIdx.y = front).
#pragma once

namespace bigfront {

constexpr int BF_MIN_FS = 288;      // scalar rows (48 pose blocks) from which a front takes this path (measured: 480 / 288 / 192 -> 26.7 / 22.5 / 20.8 ms of Cholesky on the loop-closure scene)
constexpr int BF_PC = 48;           // panel width
constexpr int BF_RS = 128;          // rows per k_bf_panel CTA (one per thread)
constexpr int BF_T = 64;            // trailing tile
constexpr int BF_LDS = 68;          // shared-memory stride of the k-major operand tiles

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// ---- extend-add: grid (parent block column, front); the CTA's eight warps split the rows ----
__global__ void __launch_bounds__(256)
k_bf_extend(const int *__restrict__ bigSn, const SnodeDesc *__restrict__ sn,
            const int *__restrict__ childIdx, const int *__restrict__ relIdx, double *__restrict__ fronts)
{
    const SnodeDesc d = sn[bigSn[blockIdx.y]];
    const int fdim = d.ncols + d.nstruct, fs = 6 * fdim, ld = fs + 1;
    const int Jb = blockIdx.x;
    if (Jb >= fdim) return;
    double *F = fronts + d.frontOff;
    const int tid = threadIdx.x;
    for (int ci = 0; ci < d.nchild; ci++) {
        const SnodeDesc c = sn[childIdx[d.childOff + ci]];
        const int *rel = relIdx + c.structOff;
        // the child's struct index that lands on parent block column Jb (rel ascending)
        int lo = 0, hi = c.nstruct;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (rel[mid] < Jb) lo = mid + 1; else hi = mid; }
        if (lo >= c.nstruct || rel[lo] != Jb) continue;
        const double *Fc = fronts + c.frontOff;
        const int ncc = 6 * c.ncols, us = 6 * c.nstruct, ldc = 6 * (c.ncols + c.nstruct) + 1;
        // the six columns of the block together: entry e = (row r, column q); four entries per thread
        // and pass, all loads before the first add (the scattered read-modify-write pays its L2
        // latency once per pass)
        const int cc0 = 6 * lo;
        const int rows = us + 1 - cc0;                      // child rows cc0 .. us (rhs row last)
        for (int e0 = tid; e0 < 6 * rows; e0 += 4 * 256) {
            double cv[4], pv[4];
            double *dp[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int e = e0 + 256 * u;
                dp[u] = nullptr;
                if (e < 6 * rows) {
                    const int q = e / rows, r = cc0 + (e - q * rows);
                    if (r >= cc0 + q) {                     // lower triangle of the child's update matrix
                        const int pr = (r == us) ? fs : 6 * rel[r / 6] + (r % 6);
                        dp[u] = F + (size_t)(6 * Jb + q) * ld + pr;
                        cv[u] = Fc[(size_t)(ncc + cc0 + q) * ldc + ncc + r];
                        pv[u] = *dp[u];
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (dp[u]) *dp[u] = pv[u] + cv[u];
        }
        __syncthreads();                                    // children in their fixed order
    }
}

// ---- panel: diagonal block + triangular solve of the rows below; grid (slab, front) ----
// The CTA stages the 48 x 48 diagonal block and ITS slab of BF_RS rows below it in shared memory and runs
// the same panel routine as the one-CTA-per-front kernel on them (factor_panel, solve.cu: 6x6 diagonal
// blocks + their inverses by one thread, row solves and in-panel updates by the whole CTA).  Every CTA of
// the panel factors the diagonal block again -- 18 k flops, cheaper than a launch plus a grid-wide
// dependency; CTA 0 writes it back.
constexpr int BF_PT = 256;
__global__ void __launch_bounds__(BF_PT)
k_bf_panel(const int *__restrict__ bigSn, const SnodeDesc *__restrict__ sn, double *__restrict__ fronts,
           int *__restrict__ errflag, int p0, double *__restrict__ diagOut)
{
    const SnodeDesc d = sn[bigSn[blockIdx.y]];
    const int fs = 6 * (d.ncols + d.nstruct), ld = fs + 1, nc = 6 * d.ncols;
    if (p0 >= nc) return;
    const int pc = min(BF_PC, nc - p0);
    const int below0 = p0 + pc;                   // first row below the diagonal block
    const int nbelow = fs + 1 - below0;           // rows below (the right-hand-side row is the last one)
    const int slab0 = blockIdx.x * BF_RS;
    if (slab0 >= nbelow && blockIdx.x > 0) return;
    const int srows = max(0, min(BF_RS, nbelow - slab0));
    double *F = fronts + d.frontOff;
    extern __shared__ double P[];                 // [pc][ldp] column-major: diagonal block rows, then the slab
    __shared__ double Li[36];
    const int tid = threadIdx.x;
    const int rows = pc + srows, ldp = rows | 1;  // odd leading dimension: a lane walking a row hits every bank
    for (int t = tid; t < pc * rows; t += BF_PT) {
        const int q = t / rows, rr = t - q * rows;
        const int gr = rr < pc ? p0 + rr : below0 + slab0 + (rr - pc);
        P[q * ldp + rr] = (rr >= q) ? F[(size_t)(p0 + q) * ld + gr] : 0.0;
    }
    __syncthreads();
    factor_panel(P, ldp, rows, pc, Li, errflag, tid, BF_PT);
    // The factored diagonal block must NOT go back into the front here: CTAs of this launch that start later
    // still have to read the unfactored block.  CTA 0 parks it in diagOut[front]; k_bf_syrk (next launch)
    // copies it into place.
    double *dg = diagOut + (size_t)blockIdx.y * (BF_PC * BF_PC);
    for (int t = tid; t < pc * rows; t += BF_PT) {
        const int q = t / rows, rr = t - q * rows;
        if (rr < q) continue;
        if (rr < pc) {
            if (blockIdx.x == 0) dg[q * BF_PC + rr] = P[q * ldp + rr];
            continue;
        }
        F[(size_t)(p0 + q) * ld + below0 + slab0 + (rr - pc)] = P[q * ldp + rr];
    }
}

// ---- trailing update on the FP64 tensor cores; grid (lower-triangle tile, front), 128 threads ----
__global__ void __launch_bounds__(128)
k_bf_syrk(const int *__restrict__ bigSn, const SnodeDesc *__restrict__ sn, double *__restrict__ fronts, int p0,
          const double *__restrict__ diagIn)
{
    const SnodeDesc d = sn[bigSn[blockIdx.y]];
    const int fs = 6 * (d.ncols + d.nstruct), ld = fs + 1, nc = 6 * d.ncols;
    if (p0 >= nc) return;
    const int pc = min(BF_PC, nc - p0);
    if (blockIdx.x == 0) {                        // the panel's factored diagonal block, parked by k_bf_panel
        const double *dg = diagIn + (size_t)blockIdx.y * (BF_PC * BF_PC);
        double *Fd = fronts + d.frontOff;
        for (int t = threadIdx.x; t < pc * pc; t += 128) {
            const int q = t / pc, rr = t - q * pc;
            if (rr >= q) Fd[(size_t)(p0 + q) * ld + p0 + rr] = dg[q * BF_PC + rr];
        }
    }
    const int t0 = p0 + pc;                       // first trailing row / column
    const int nrows = fs + 1 - t0;                // trailing rows incl. the right-hand-side row
    const int nT = (nrows + BF_T - 1) / BF_T;
    // tile (I, J), I >= J, from the linear index: row-major over the lower triangle
    int I = 0, rem = blockIdx.x;
    if (rem >= nT * (nT + 1) / 2) return;
    I = (int)((sqrt(8.0 * rem + 1.0) - 1.0) * 0.5);
    while ((I + 1) * (I + 2) / 2 <= rem) I++;
    while (I * (I + 1) / 2 > rem) I--;
    const int J = rem - I * (I + 1) / 2;
    double *F = fronts + d.frontOff;
    extern __shared__ double sh[];
    double *As = sh;                              // [BF_PC][BF_LDS]: P[row tile I][q], k-major
    double *Bs = sh + BF_PC * BF_LDS;             // [BF_PC][BF_LDS]: P[row tile J][q]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rI = t0 + BF_T * I, rJ = t0 + BF_T * J;
    for (int e = tid; e < BF_PC * BF_T; e += 128) {
        const int q = e / BF_T, x = e - q * BF_T;
        const double *col = F + (size_t)(p0 + q) * ld;
        As[q * BF_LDS + x] = (q < pc && rI + x <= fs) ? col[rI + x] : 0.0;
        Bs[q * BF_LDS + x] = (q < pc && rJ + x <= fs) ? col[rJ + x] : 0.0;
    }
    __syncthreads();
    const int wr = 32 * (warp >> 1), wc = 32 * (warp & 1);      // the warp's 32 x 32 sub-tile
    // a diagonal tile's upper sub-tile (rows 0..31, cols 32..63) lies above the diagonal: nothing to do
    if (I == J && wr + 31 < wc) return;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
    const int fr = lane >> 2, fk = lane & 3;
#pragma unroll 4
    for (int ks = 0; ks < BF_PC / 4; ks++) {
        double a[4], b[4];
        const double *ap = As + (4 * ks + fk) * BF_LDS + wr + fr;
        const double *bp = Bs + (4 * ks + fk) * BF_LDS + wc + fr;
#pragma unroll
        for (int i = 0; i < 4; i++) { a[i] = ap[8 * i]; b[i] = bp[8 * i]; }
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    // C(r, c) -= acc for r >= c, r <= fs (right-hand-side row included), c < fs
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int r = rI + wr + 8 * i + fr;
                const int c = rJ + wc + 8 * j + 2 * fk + h;
                if (r >= c && r <= fs && c < fs) F[(size_t)c * ld + r] -= acc[i][j][h];
            }
}

// ---- back-solve, first half: t = y_J - L21^T x_struct; grid (8 columns per CTA, front), a warp per column.
// t lands in the node's own slice of xperm, where k_front_backsolve<true> finishes with L11.
__global__ void __launch_bounds__(256)
k_bf_back_gemv(const int *__restrict__ bigSn, const SnodeDesc *__restrict__ sn,
               const int *__restrict__ structIdx, const double *__restrict__ fronts, double *__restrict__ xperm)
{
    const SnodeDesc d = sn[bigSn[blockIdx.y]];
    const int fs = 6 * (d.ncols + d.nstruct), ld = fs + 1, nc = 6 * d.ncols, us = 6 * d.nstruct;
    const int lane = threadIdx.x & 31, c = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (c >= nc) return;
    const double *F = fronts + d.frontOff;
    double *xj = xperm + 6 * (size_t)d.poseOff;
    const double *col = F + (size_t)c * ld + nc;
    const int *st = structIdx + d.structOff;
    // fixed order: lane-private partial sums over rows lane, lane + 32, ... then one butterfly
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int r = lane;
    for (; r + 96 < us; r += 128) {
        const double x0 = xj[6 * st[r / 6] + r % 6], x1 = xj[6 * st[(r + 32) / 6] + (r + 32) % 6];
        const double x2 = xj[6 * st[(r + 64) / 6] + (r + 64) % 6], x3 = xj[6 * st[(r + 96) / 6] + (r + 96) % 6];
        a0 = fma(col[r], x0, a0); a1 = fma(col[r + 32], x1, a1);
        a2 = fma(col[r + 64], x2, a2); a3 = fma(col[r + 96], x3, a3);
    }
    for (; r < us; r += 32) a0 = fma(col[r], xj[6 * st[r / 6] + r % 6], a0);
    double a = (a0 + a1) + (a2 + a3);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) xj[6 * (size_t)d.first + c] = F[(size_t)c * ld + fs] - a;
}

} // namespace bigfront
