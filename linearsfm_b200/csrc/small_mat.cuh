// Tiny dense row-major matrix helpers (fully unrolled, registers) + warp-aggregated FP64 atomics.
#pragma once
#include <cuda_runtime.h>

namespace sm {

// C[M x N] = A[M x K] * B[K x N]
template <int M, int K, int N>
__device__ __forceinline__ void mm(const double *A, const double *B, double *C)
{
#pragma unroll
    for (int i = 0; i < M; i++)
#pragma unroll
        for (int j = 0; j < N; j++) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < K; k++) s = fma(A[i * K + k], B[k * N + j], s);
            C[i * N + j] = s;
        }
}

// C[M x N] = A^T * B, A is [K x M], B is [K x N]
template <int M, int K, int N>
__device__ __forceinline__ void mtm(const double *A, const double *B, double *C)
{
#pragma unroll
    for (int i = 0; i < M; i++)
#pragma unroll
        for (int j = 0; j < N; j++) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < K; k++) s = fma(A[k * M + i], B[k * N + j], s);
            C[i * N + j] = s;
        }
}

// C[M x N] += A^T * B
template <int M, int K, int N>
__device__ __forceinline__ void mtm_acc(const double *A, const double *B, double *C)
{
#pragma unroll
    for (int i = 0; i < M; i++)
#pragma unroll
        for (int j = 0; j < N; j++) {
            double s = C[i * N + j];
#pragma unroll
            for (int k = 0; k < K; k++) s = fma(A[k * M + i], B[k * N + j], s);
            C[i * N + j] = s;
        }
}

// C[M x N] = A * B^T, A is [M x K], B is [N x K]
template <int M, int K, int N>
__device__ __forceinline__ void mmt(const double *A, const double *B, double *C)
{
#pragma unroll
    for (int i = 0; i < M; i++)
#pragma unroll
        for (int j = 0; j < N; j++) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < K; k++) s = fma(A[i * K + k], B[j * K + k], s);
            C[i * N + j] = s;
        }
}

template <int N> __device__ __forceinline__ void load(const double *__restrict__ g, double *r)
{
#pragma unroll
    for (int i = 0; i < N; i++) r[i] = g[i];
}

template <int N> __device__ __forceinline__ void store(double *__restrict__ g, const double *r)
{
#pragma unroll
    for (int i = 0; i < N; i++) g[i] = r[i];
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Add vals[0..N) of every lane with active==true into dst[0..N).  MUST be called by all 32 lanes
// of a converged warp (callers keep their loops warp-uniform and pass active=false for idle
// lanes).  When all active lanes target the same destination the N values are butterfly-reduced
// first and one lane issues N atomics instead of 32 N (the "hub pose" case: every feature of a map
// is linked to the same pose block).
template <int N>
__device__ __forceinline__ void warp_agg_atomic_add(double *dst, const double *vals, bool active)
{
    const unsigned full = 0xffffffffu;
    unsigned amask = __ballot_sync(full, active);
    if (amask == 0u) return;
    int leader = __ffs(amask) - 1;
    unsigned long long key = (unsigned long long)dst;
    unsigned long long k0 = __shfl_sync(full, key, leader);
    bool uniform = __all_sync(full, !active || key == k0);
    if (uniform) {
#pragma unroll
        for (int i = 0; i < N; i++) {
            double s = warp_sum(active ? vals[i] : 0.0);
            if ((int)(threadIdx.x & 31) == leader) atomicAdd(dst + i, s);
        }
    } else if (active) {
#pragma unroll
        for (int i = 0; i < N; i++) atomicAdd(dst + i, vals[i]);
    }
}

} // namespace sm
