// Shared device code of the stereo and mono joins: hash join of feature ids (first-index semantics of
// std::find, LinearSFMImp.cpp:2581-2599 / 7322-7340) and the joint feature numbering.
#pragma once
#include "ops.h"

namespace joinc {

typedef unsigned long long u64;
#define HASH_EMPTY 0xffffffffffffffffull

__device__ __forceinline__ u64 mix64(u64 x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

__global__ void k_hash_insert(const DMap *__restrict__ C, const int *__restrict__ featPreC, int K,
                              int totC, u64 *__restrict__ keys, int *__restrict__ vals, u64 mask)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= totC) return;
    int k = seg_find(featPreC, K, g);
    int c = g - featPreC[k];
    u64 key = ((u64)(unsigned)k << 32) | (unsigned)C[k].featNo[c];
    u64 h = mix64(key) & mask;
    while (true) {
        u64 prev = atomicCAS(&keys[h], HASH_EMPTY, key);
        if (prev == HASH_EMPTY || prev == key) { atomicMin(&vals[h], c); return; }   // first index wins
        h = (h + 1) & mask;
    }
}

// for each End feature: first Cur feature with the same id (LinearSFMImp.cpp:2581-2599)
__global__ void k_hash_probe(const DMap *__restrict__ E, const int *__restrict__ featPreE,
                             const int *__restrict__ featPreC, int K, int totE,
                             const u64 *__restrict__ keys, const int *__restrict__ vals, u64 mask,
                             int *__restrict__ ownerOfCur)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= totE) return;
    int k = seg_find(featPreE, K, g);
    int i = g - featPreE[k];
    u64 key = ((u64)(unsigned)k << 32) | (unsigned)E[k].featNo[i];
    u64 h = mix64(key) & mask;
    while (true) {
        u64 cur = keys[h];
        if (cur == HASH_EMPTY) return;
        if (cur == key) { atomicMax(&ownerOfCur[featPreC[k] + vals[h]], i); return; }  // last i wins (2596)
        h = (h + 1) & mask;
    }
}

__global__ void k_only_flag(const int *__restrict__ ownerOfCur, int totC, int *__restrict__ flag)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g > totC) return;
    flag[g] = (g < totC && ownerOfCur[g] < 0) ? 1 : 0;
}

__global__ void k_only_count(const int *__restrict__ featPreC, int K, const int *__restrict__ scan,
                             int *__restrict__ nOnly)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    nOnly[k] = scan[featPreC[k + 1]] - scan[featPreC[k]];
}

// joint index of every Cur feature (Curfeature2, 2596/2639) and the inverse map joint -> Cur
__global__ void k_joint_index(const DMap *__restrict__ E, const int *__restrict__ featPreC,
                              const int *__restrict__ featPreJ, int K, int totC,
                              const int *__restrict__ ownerOfCur, const int *__restrict__ scan,
                              int *__restrict__ jointOfCur, int *__restrict__ curOfJoint)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= totC) return;
    int k = seg_find(featPreC, K, g);
    int own = ownerOfCur[g];
    int jf = own >= 0 ? own : E[k].n + (scan[g] - scan[featPreC[k]]);
    jointOfCur[g] = jf;
    curOfJoint[featPreJ[k] + jf] = g - featPreC[k];
}


} // namespace joinc
