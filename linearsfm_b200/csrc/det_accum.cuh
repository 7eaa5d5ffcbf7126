// Deterministic scatter-add: the replacement for FP64 atomics on shared accumulation targets.
//
// Several stages of the hot path add many small blocks into few targets (U'(pos,pos) and U'(p,pos) of
// the frame transform, LinearSFMImp.cpp:819-857 / 1386-1423 / 1606-1660; eP of the join, 2658-2790).
// FP64 atomics make the order of those additions -- and therefore the last bits of the result -- vary
// from run to run, and the hierarchical chain amplifies that.  Here every
// producer thread instead WRITES its contribution as a record
//     key[r] = target index in [0, ntargets) (== ntargets: no contribution),  val[NV r .. NV r + NV)
// at a position r that depends only on the thread's index.  Records are brought into target order
// either by construction (producers enumerated target-contiguously: `reduce_sorted` on the records as
// they lie, sidx == nullptr) or by one stable radix sort of (key, r); one CTA per target then adds its
// records up in a fixed association (NS interleaved streams, four partial sums each, combined in a
// fixed order).  Long runs (the one hot target per map) are pre-summed per aligned window of WIN
// records by `k_window`, so the target's CTA reads one partial per window.  The result is bit-identical
// from run to run: nothing depends on the hardware's scheduling.
#pragma once
#include "device.h"
#include <cub/cub.cuh>

namespace det {

constexpr int WIN = 256;     // window length of the first-level sums

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

// record at sorted position r: sidx == nullptr means the records already lie in target order
__device__ __forceinline__ int rec_at(const int *__restrict__ sidx, int r) { return sidx ? sidx[r] : r; }

// First level: window c = sorted positions [WIN c, WIN c + WIN).  If all of them belong to one target
// the window's sum goes to part[NV c ..) (windows that mix targets are read record by record later).
template <int NV>
__global__ void __launch_bounds__(256)
k_window(const int *__restrict__ skey, const int *__restrict__ sidx, int n, int none,
         const double *__restrict__ val, double *__restrict__ part)
{
    constexpr int NS = 252 / NV;
    __shared__ double sh[NS][NV];
    const int c = blockIdx.x, lo = c * WIN, hi = lo + WIN;
    if (!(hi <= n && skey[lo] == skey[hi - 1] && skey[lo] != none)) return;      // CTA-uniform
    const int tid = threadIdx.x, s = tid / NV, q = tid - s * NV;
    if (s < NS) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int r = lo + s;
        for (; r + 3 * NS < hi; r += 4 * NS) {
            a0 += val[(size_t)NV * rec_at(sidx, r) + q];
            a1 += val[(size_t)NV * rec_at(sidx, r + NS) + q];
            a2 += val[(size_t)NV * rec_at(sidx, r + 2 * NS) + q];
            a3 += val[(size_t)NV * rec_at(sidx, r + 3 * NS) + q];
        }
        if (r < hi) a0 += val[(size_t)NV * rec_at(sidx, r) + q];
        if (r + NS < hi) a1 += val[(size_t)NV * rec_at(sidx, r + NS) + q];
        if (r + 2 * NS < hi) a2 += val[(size_t)NV * rec_at(sidx, r + 2 * NS) + q];
        sh[s][q] = (a0 + a1) + (a2 + a3);
    }
    __syncthreads();
    if (tid < NV) {
        double sum = 0.0;
#pragma unroll
        for (int i = 0; i < NS; i++) sum += sh[i][tid];
        part[(size_t)NV * c + tid] = sum;
    }
}

// One CTA per target t: its records are the sorted positions [lo, hi) with key == t.  Items of the sum,
// in order: records before the first whole window, the whole windows' partial sums, the records after.
template <int NV, class Apply>
__global__ void __launch_bounds__(256)
k_reduce(const int *__restrict__ skey, const int *__restrict__ sidx, int n,
         const double *__restrict__ val, const double *__restrict__ part, int ntargets, Apply apply)
{
    constexpr int NS = 252 / NV;               // interleaved streams (7 for NV = 36, 42 for NV = 6)
    __shared__ double sh[NS][NV];
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
    int c0 = (lo + WIN - 1) / WIN, c1 = hi / WIN;          // whole windows inside [lo, hi): c0 .. c1-1
    int headEnd = hi, tailBeg = hi;
    if (c1 > c0) { headEnd = c0 * WIN; tailBeg = c1 * WIN; } else { c0 = c1 = 0; }
    const int nHead = headEnd - lo, nWin = c1 - c0, nTail = hi - tailBeg;
    const int nItems = nHead + nWin + nTail;
    const int s = tid / NV, q = tid - s * NV;
    if (s < NS) {
        auto item = [&](int i) -> double {
            if (i < nHead) return val[(size_t)NV * rec_at(sidx, lo + i) + q];
            i -= nHead;
            if (i < nWin) return part[(size_t)NV * (c0 + i) + q];
            return val[(size_t)NV * rec_at(sidx, tailBeg + (i - nWin)) + q];
        };
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int i = s;
        for (; i + 3 * NS < nItems; i += 4 * NS) {
            a0 += item(i); a1 += item(i + NS); a2 += item(i + 2 * NS); a3 += item(i + 3 * NS);
        }
        if (i < nItems) a0 += item(i);
        if (i + NS < nItems) a1 += item(i + NS);
        if (i + 2 * NS < nItems) a2 += item(i + 2 * NS);
        sh[s][q] = (a0 + a1) + (a2 + a3);
    }
    __syncthreads();
    if (tid < NV) {
        double sum = 0.0;
        const int ns = min(NS, nItems);
        for (int i = 0; i < ns; i++) sum += sh[i][tid];
        apply(t, tid, sum, hi - lo);
    }
}

// keys sorted (by sort_records, or by construction with sidx == nullptr): window sums, then targets.
// Returns the number of launches.
template <int NV, class Apply>
inline int reduce_sorted(Context &ctx, const int *skey, const int *sidx, int n, const double *val, int ntargets,
                         Apply apply)
{
    if (ntargets <= 0) return 0;
    cudaStream_t s = ctx.stream;
    const int nwin = n / WIN;
    DevBuf<double> part((size_t)NV * std::max(nwin, 1), s);
    int nl = 0;
    if (nwin > 0) { k_window<NV><<<nwin, 256, 0, s>>>(skey, sidx, n, ntargets, val, part.p); nl++; }
    k_reduce<NV, Apply><<<ntargets, 256, 0, s>>>(skey, sidx, n, val, part.p, ntargets, apply); nl++;
    return nl;
}

template <int NV, class Apply>
inline int reduce(Context &ctx, const Sorted &S, const double *val, int ntargets, Apply apply)
{
    return reduce_sorted<NV, Apply>(ctx, S.key.p, S.idx.p, S.n, val, ntargets, apply);
}

} // namespace det
