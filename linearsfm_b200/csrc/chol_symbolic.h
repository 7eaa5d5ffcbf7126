// Host-side symbolic phase of the block multifrontal Cholesky that replaces CHOLMOD
// (cholmod_amd + cholmod_analyze_p, LinearSFMImp.cpp:2380-2449): ordering, supernodes, fronts.
#pragma once
#include <vector>
#include <cstdint>

// LSFM-ND block ordering (spec in DESIGN.md "Ordering"; the oracle's independent twin is
// oracle/cholmod_shim.c:lsfm_nd_order).  Input: symmetric block graph as CSR adjacency WITHOUT
// self loops (ptr[m+1], adj sorted ascending).  Output: perm[k] = original block eliminated k-th,
// and `nodes` = boundaries (size nn+1) of the dissection nodes in elimination order; every node
// becomes one (relaxed) supernode.
void lsfm_nd_order(int m, const int *ptr, const int *adj, std::vector<int> &perm,
                   std::vector<int> &nodes);
// LSFM-MD (exact minimum degree, tie: smallest index; relaxed supernodes) and the library's rule:
// LSFM-ND unless one of its separators exceeds 64 poses (loop closures, dense overlap), then LSFM-MD.
// The oracle's twins: oracle/cholmod_shim.c:lsfm_md_order / lsfm_shim_order.
void lsfm_md_order(int m, const int *ptr, const int *adj, std::vector<int> &perm,
                   std::vector<int> &nodes);
void lsfm_order(int m, const int *ptr, const int *adj, std::vector<int> &perm, std::vector<int> &nodes);

struct SnodeDesc {
    int join;            // which join (map) of the batch
    int first, ncols;    // block columns [first, first+ncols) in the join's elimination order
    int structOff, nstruct;   // rows below the diagonal block (global offset into structIdx)
    long long frontOff;  // offset (doubles) of the (6 fdim + 1) x (6 fdim) column-major front
    int parent;          // global supernode id or -1
    int childOff, nchild;
    int poseOff;         // global pose offset of the join (posePre[join])
    int level;
};

struct SlotMap { int sn; short transpose; short pad; int lrow, lcol; };   // S block -> front block

struct BatchSymbolic {
    std::vector<SnodeDesc> sn;
    std::vector<int> structIdx;     // permuted block indices, ascending
    std::vector<int> relIdx;        // position of structIdx[i] inside the PARENT's front (blocks)
    std::vector<int> childIdx;      // children (global ids) grouped per supernode
    std::vector<int> levelPtr, levelSn;
    std::vector<SlotMap> slot;      // one per S block (global slot order)
    std::vector<int> poseSn, poseLcol;   // per global pose: owning supernode, local block column
    std::vector<int> perm;          // per global pose: perm (local original index eliminated k-th)
    long long frontDoubles = 0;
    double flops = 0.0;             // sum_j c_j^2 over scalar columns of L (CHOLMOD's fl convention)
    int maxFdim = 0;
};

// keys: sorted unique (join<<44 | row<<22 | col) with row<=col; sOff[K+1] offsets per join;
// m[k] poses per join; posePre[K+1].
void build_symbolic(int K, const std::vector<int> &m, const std::vector<int> &posePre,
                    const unsigned long long *keys, size_t nkeys, const std::vector<int> &sOff,
                    BatchSymbolic &out, int nthreads);
