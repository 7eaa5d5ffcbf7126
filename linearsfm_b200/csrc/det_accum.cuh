// Deterministic scatter-add: the replacement for FP64 atomics on shared accumulation targets.
//
// Several stages of the hot path add many small blocks into few targets (U'(pos,pos) and U'(p,pos) of
// the frame transform, LinearSFMImp.cpp:819-857 / 1386-1423 / 1606-1660; eP of the join, 2658-2790;
// the reduced camera matrix S, 2246-2332).  FP64 atomics make the order of those additions -- and
// therefore the last bits of the result -- vary from run to run, and the hierarchical chain amplifies
// that.  Here every producer thread instead WRITES its contribution as a record
//     key[r] = target index in [0, ntargets) (== ntargets: no contribution),  val[NV r .. NV r + NV)
// at a position r that depends only on the thread's index; one stable radix sort of (key, r) groups
// the records of a target in production order, and one CTA per target adds them up in a fixed
// association (NS interleaved streams, four partial sums each, combined in a fixed order).  The
// result is bit-identical from run to run and independent of the launch configuration of the
// producers' hardware scheduling.
#pragma once
#include "device.h"
#include <cub/cub.cuh>

namespace det {

struct Sorted {
    DevBuf<int> key, idx;
    int n = 0;
};

static __global__ void k_iota(int *__restrict__ a, int n)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < n) a[g] = g;
}

// stable sort of (key[r], r) by key; keys are in [0, ntargets], ntargets = "no contribution".
// Returns the number of kernel launches (for the launch statistics).
inline int sort_records(Context &ctx, const int *key, int n, int ntargets, Sorted &out)
{
    cudaStream_t s = ctx.stream;
    out.n = n;
    out.key.alloc((size_t)std::max(n, 1), s);
    out.idx.alloc((size_t)std::max(n, 1), s);
    if (n == 0) return 0;
    DevBuf<int> iota((size_t)n, s);
    k_iota<<<(n + 255) / 256, 256, 0, s>>>(iota.p, n);
    int bits = 1;
    while ((1ll << bits) <= (long long)ntargets) bits++;
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, key, out.key.p, iota.p, out.idx.p, n, 0, bits, s);
    DevBuf<char> tmp(tb, s);
    cub::DeviceRadixSort::SortPairs(tmp.p, tb, key, out.key.p, iota.p, out.idx.p, n, 0, bits, s);
    return 2 + (bits + 7) / 8;
}

// One CTA per target t: sum of val[NV idx .. ) over the records with key == t, in sorted (= production)
// order with a fixed association; then apply(t, q, sum_q, count) for every q < NV by one thread each.
template <int NV, class Apply>
__global__ void __launch_bounds__(256)
k_reduce(const int *__restrict__ skey, const int *__restrict__ sidx, int n,
         const double *__restrict__ val, int ntargets, Apply apply)
{
    constexpr int NS = 252 / NV;               // interleaved streams (7 for NV = 36, 42 for NV = 6)
    __shared__ double part[NS][NV];
    __shared__ int seg[2];
    const int t = blockIdx.x;
    const int tid = threadIdx.x;
    if (tid < 2) {
        const int want = t + tid;              // lower_bound(skey, want)
        int lo = 0, hi = n;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (skey[mid] < want) lo = mid + 1; else hi = mid; }
        seg[tid] = lo;
    }
    __syncthreads();
    const int lo = seg[0], hi = seg[1];
    const int s = tid / NV, q = tid - s * NV;
    if (s < NS) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int r = lo + s;
        for (; r + 3 * NS < hi; r += 4 * NS) {
            const int i0 = sidx[r], i1 = sidx[r + NS], i2 = sidx[r + 2 * NS], i3 = sidx[r + 3 * NS];
            a0 += val[(size_t)NV * i0 + q];
            a1 += val[(size_t)NV * i1 + q];
            a2 += val[(size_t)NV * i2 + q];
            a3 += val[(size_t)NV * i3 + q];
        }
        if (r < hi) a0 += val[(size_t)NV * sidx[r] + q];
        if (r + NS < hi) a1 += val[(size_t)NV * sidx[r + NS] + q];
        if (r + 2 * NS < hi) a2 += val[(size_t)NV * sidx[r + 2 * NS] + q];
        part[s][q] = (a0 + a1) + (a2 + a3);
    }
    __syncthreads();
    if (tid < NV) {
        double sum = 0.0;
        const int ns = min(NS, hi - lo);
        for (int i = 0; i < ns; i++) sum += part[i][tid];
        apply(t, tid, sum, hi - lo);
    }
}

template <int NV, class Apply>
inline void reduce(Context &ctx, const Sorted &S, const double *val, int ntargets, Apply apply)
{
    if (ntargets <= 0) return;
    k_reduce<NV, Apply><<<ntargets, 256, 0, ctx.stream>>>(S.key.p, S.idx.p, S.n, val, ntargets, apply);
}

} // namespace det
