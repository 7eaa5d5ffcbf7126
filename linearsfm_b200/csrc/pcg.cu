// Block-Jacobi preconditioned conjugate gradients on the reduced camera system S x = E -- the
// cross-check of the multifrontal Cholesky that BASELINE.json's north_star asks for (the reference
// solves the same system with CHOLMOD, LinearSFMImp.cpp:2380-2449; it has no iterative solver).
// Test-time tool: C ABI lsfm_pcg_block(), driven by tests/test_gpu_pcg.py on the S / E captured from a
// solve.  S is symmetric, given as the upper block triangle in block CRS (rowptr[m+1], colidx, 36
// doubles per block row-major, diagonal blocks stored full) -- the layout of lsfm_debug_last_solve().
#include "device.h"
#include "small_mat.cuh"
#include "../../include/linearsfm_b200.h"
#include <vector>
#include <cmath>

Context *lsfm_internal_ctx();      // capi.cu

namespace {

// y = S x for the symmetric matrix given by its upper block triangle: thread per stored block
__global__ void k_bsr_symv(int m, int nnzb, const int *__restrict__ rowOf, const int *__restrict__ colidx,
                           const double *__restrict__ S, const double *__restrict__ x, double *__restrict__ y)
{
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nnzb) return;
    const int i = rowOf[b], j = colidx[b];
    double A[36], xi[6], xj[6], t[6];
    sm::load<36>(S + 36 * (size_t)b, A);
    sm::load<6>(x + 6 * (size_t)j, xj);
    sm::mm<6, 6, 1>(A, xj, t);
#pragma unroll
    for (int q = 0; q < 6; q++) atomicAdd(y + 6 * (size_t)i + q, t[q]);
    if (i != j) {
        sm::load<6>(x + 6 * (size_t)i, xi);
        sm::mtm<6, 6, 1>(A, xi, t);
#pragma unroll
        for (int q = 0; q < 6; q++) atomicAdd(y + 6 * (size_t)j + q, t[q]);
    }
}

// inverse of every 6x6 diagonal block (Gauss-Jordan with partial pivoting; the blocks are SPD)
__global__ void k_diag_inverse(int m, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                               const double *__restrict__ S, double *__restrict__ Minv)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    double A[6][12];
    const int b = rowptr[i];          // first block of the row is the diagonal (upper triangle, sorted)
    const bool ok = b < rowptr[i + 1] && colidx[b] == i;
    for (int r = 0; r < 6; r++)
        for (int c = 0; c < 6; c++) {
            A[r][c] = ok ? S[36 * (size_t)b + 6 * r + c] : (r == c ? 1.0 : 0.0);
            A[r][6 + c] = (r == c) ? 1.0 : 0.0;
        }
    for (int c = 0; c < 6; c++) {
        int p = c;
        for (int r = c + 1; r < 6; r++) if (fabs(A[r][c]) > fabs(A[p][c])) p = r;
        if (p != c) for (int q = 0; q < 12; q++) { double t = A[c][q]; A[c][q] = A[p][q]; A[p][q] = t; }
        const double inv = 1.0 / A[c][c];
        for (int q = 0; q < 12; q++) A[c][q] *= inv;
        for (int r = 0; r < 6; r++) {
            if (r == c) continue;
            const double f = A[r][c];
            for (int q = 0; q < 12; q++) A[r][q] -= f * A[c][q];
        }
    }
    for (int r = 0; r < 6; r++)
        for (int c = 0; c < 6; c++) Minv[36 * (size_t)i + 6 * r + c] = A[r][6 + c];
}

__global__ void k_prec(int m, const double *__restrict__ Minv, const double *__restrict__ r, double *__restrict__ z)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    double A[36], v[6], t[6];
    sm::load<36>(Minv + 36 * (size_t)i, A);
    sm::load<6>(r + 6 * (size_t)i, v);
    sm::mm<6, 6, 1>(A, v, t);
    sm::store<6>(z + 6 * (size_t)i, t);
}

__global__ void k_dot(int n, const double *__restrict__ a, const double *__restrict__ b, double *__restrict__ out)
{
    __shared__ double sh[8];
    double s = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s += a[i] * b[i];
    s = sm::warp_sum(s);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += sh[w];
        atomicAdd(out, t);
    }
}

// y = a x + b y
__global__ void k_axpby(int n, double a, const double *__restrict__ x, double b, double *__restrict__ y)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = a * x[i] + b * y[i];
}

} // namespace

extern "C" int lsfm_pcg_block(int m, const int *rowptr, const int *colidx, const double *S, const double *E,
                              double *x, double tol, int max_iters, int *iters_done, double *rel_residual)
{
    if (m <= 0 || !rowptr || !colidx || !S || !E || !x) return LSFM_ERR_ARG;
    if (lsfm_device_count() <= 0) return LSFM_ERR_NO_DEVICE;
    try {
        Context *ctx = lsfm_internal_ctx();
        if (!ctx) return LSFM_ERR_CUDA;
        cudaStream_t s = ctx->stream;
        const int nnzb = rowptr[m], n = 6 * m;
        std::vector<int> rowOf(nnzb);
        for (int i = 0; i < m; i++)
            for (int b = rowptr[i]; b < rowptr[i + 1]; b++) rowOf[b] = i;
        DevBuf<int> dRowptr(m + 1, s), dCol(std::max(nnzb, 1), s), dRowOf(std::max(nnzb, 1), s);
        DevBuf<double> dS(36 * (size_t)std::max(nnzb, 1), s), dMinv(36 * (size_t)m, s);
        DevBuf<double> r(n, s), z(n, s), p(n, s), Ap(n, s), xv(n, s), dots(4, s);
        dRowptr.upload(rowptr, m + 1);
        dCol.upload(colidx, nnzb); dRowOf.upload(rowOf.data(), nnzb);
        dS.upload(S, 36 * (size_t)nnzb);
        r.upload(E, n);                              // x0 = 0: r = E
        xv.zero();
        const int TB = 128;
        k_diag_inverse<<<(m + TB - 1) / TB, TB, 0, s>>>(m, dRowptr.p, dCol.p, dS.p, dMinv.p);
        auto dot = [&](const double *a, const double *b) {
            dots.zero();
            k_dot<<<64, 256, 0, s>>>(n, a, b, dots.p);
            double h = 0.0;
            CUDA_CHECK(cudaMemcpyAsync(&h, dots.p, sizeof(double), cudaMemcpyDeviceToHost, s));
            CUDA_CHECK(cudaStreamSynchronize(s));
            return h;
        };
        const double e2 = dot(r.p, r.p);
        k_prec<<<(m + TB - 1) / TB, TB, 0, s>>>(m, dMinv.p, r.p, z.p);
        CUDA_CHECK(cudaMemcpyAsync(p.p, z.p, sizeof(double) * n, cudaMemcpyDeviceToDevice, s));
        double rz = dot(r.p, z.p);
        int it = 0;
        double rr = e2;
        for (; it < max_iters && rr > tol * tol * e2; it++) {
            Ap.zero();
            k_bsr_symv<<<(nnzb + TB - 1) / TB, TB, 0, s>>>(m, nnzb, dRowOf.p, dCol.p, dS.p, p.p, Ap.p);
            const double pAp = dot(p.p, Ap.p);
            if (!(pAp > 0.0)) break;
            const double alpha = rz / pAp;
            k_axpby<<<(n + 255) / 256, 256, 0, s>>>(n, alpha, p.p, 1.0, xv.p);
            k_axpby<<<(n + 255) / 256, 256, 0, s>>>(n, -alpha, Ap.p, 1.0, r.p);
            rr = dot(r.p, r.p);
            k_prec<<<(m + TB - 1) / TB, TB, 0, s>>>(m, dMinv.p, r.p, z.p);
            const double rz2 = dot(r.p, z.p);
            k_axpby<<<(n + 255) / 256, 256, 0, s>>>(n, 1.0, z.p, rz2 / rz, p.p);
            rz = rz2;
        }
        CUDA_CHECK(cudaMemcpyAsync(x, xv.p, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
        CUDA_CHECK(cudaStreamSynchronize(s));
        if (iters_done) *iters_done = it;
        if (rel_residual) *rel_residual = std::sqrt(rr / (e2 > 0.0 ? e2 : 1.0));
        return LSFM_OK;
    } catch (const LsfmError &e) {
        return e.code;
    }
}
