// Symbolic phase of the block multifrontal Cholesky (host, integer work).
//
// Replaces, for the reduced camera system S of every join:
//   cholmod_amd on the m x m block pattern          LinearSFMImp.cpp:2413
//   x6 scalar permutation + cholmod_analyze_p       LinearSFMImp.cpp:2425-2440
// CHOLMOD itself is not part of the reference tree; the ordering implemented here is
// the documented LSFM-ND rule, chosen because it yields a shallow, wide assembly tree (fronts of one
// level are factorised by one segmented launch) instead of the chain a minimum-degree ordering
// produces on the band+arrow structure of sequential map joining.
#include "chol_symbolic.h"
#include <functional>
#include <algorithm>
#include <thread>
#include <stdexcept>
#include <string>

// LSFM-ND (spec, also in DESIGN.md):
//   m <= 32: identity, one node.
//   dissect(L) on an ascending vertex list:
//     |L| <= 8            -> emit L (one node)
//     A = first |L|/2 vertices, B = the rest; cross edges = edges of the graph between A and B;
//     separator S = greedy vertex cover of the cross edges: repeatedly take the vertex with the
//       most uncovered cross edges (tie: smallest index) until none is left;
//     dissect(A \ S), dissect(B \ S), then emit S ascending (one node).
// The map-joining index order is [End poses, Cur poses] recursively, so index bisection follows the
// merge tree; the vertex cover puts the few "hub" poses (former frame origins, adjacent to a whole
// sub-map) into the separators instead of their many neighbours.
// `sepLimit` > 0: give up (return false, outputs unspecified) as soon as a separator exceeds it
static bool nd_order_impl(int m, const int *ptr, const int *adj, std::vector<int> &perm,
                          std::vector<int> &nodes, int sepLimit);
void lsfm_nd_order(int m, const int *ptr, const int *adj, std::vector<int> &perm,
                   std::vector<int> &nodes)
{
    nd_order_impl(m, ptr, adj, perm, nodes, 0);
}
static bool nd_order_impl(int m, const int *ptr, const int *adj, std::vector<int> &perm,
                          std::vector<int> &nodes, int sepLimit)
{
    perm.clear(); nodes.clear();
    perm.reserve(m);
    nodes.push_back(0);
    if (m <= 32) {
        for (int i = 0; i < m; i++) perm.push_back(i);
        if (m > 0) nodes.push_back(m);
        return true;
    }
    struct Item { std::vector<int> L; bool emit; };
    std::vector<Item> stack;
    {
        std::vector<int> all(m);
        for (int i = 0; i < m; i++) all[i] = i;
        stack.push_back({std::move(all), false});
    }
    std::vector<char> side(m, 0);        // 1 = in A, 2 = in B, 3 = chosen for the separator
    std::vector<int> cnt(m, 0);
    while (!stack.empty()) {
        Item it = std::move(stack.back());
        stack.pop_back();
        const std::vector<int> &L = it.L;
        if (L.empty()) continue;
        if (it.emit || L.size() <= 8) {
            perm.insert(perm.end(), L.begin(), L.end());
            nodes.push_back((int)perm.size());
            continue;
        }
        size_t h = L.size() / 2;
        for (size_t i = 0; i < L.size(); i++) side[L[i]] = i < h ? 1 : 2;
        long long uncovered = 0;
        // cross edges: L is ascending, so A lies below mid = L[h] and B at or above it, and the adjacency
        // lists are ascending -- only the neighbours inside [mid, L.back()] (for a vertex of A) or inside
        // [L.front(), mid) (for a vertex of B) can be on the other side: two binary searches per vertex
        // instead of a walk over the whole list (same counts, same ordering; 1.4 -> 0.4 ms at m = 3499)
        const int lo = L.front(), mid = L[h], hi = L.back();
        for (size_t i = 0; i < L.size(); i++) {
            const int v = L[i];
            const int *b = adj + ptr[v], *e = adj + ptr[v + 1];
            int c = 0;
            if (i < h) {
                for (const int *q = std::lower_bound(b, e, mid); q < e && *q <= hi; ++q) c += (side[*q] == 2);
            } else {
                for (const int *q = std::lower_bound(b, e, lo); q < e && *q < mid; ++q) c += (side[*q] == 1);
            }
            cnt[v] = c;
            uncovered += c;
        }
        uncovered /= 2;
        std::vector<int> sep, cand;              // only endpoints of cross edges can enter the cover
        for (int v : L) if (cnt[v] > 0) cand.push_back(v);
        while (uncovered > 0) {
            int best = -1, bc = 0;
            for (int v : cand)
                if (cnt[v] > bc) { bc = cnt[v]; best = v; }     // ascending scan: ties keep the smallest index
            {
                const bool inA = side[best] == 1;
                const char other = inA ? 2 : 1;
                const int *b = adj + ptr[best], *e = adj + ptr[best + 1];
                for (const int *q = std::lower_bound(b, e, inA ? mid : lo); q < e && (inA ? *q <= hi : *q < mid); ++q)
                    if (side[*q] == other) cnt[*q]--;
            }
            uncovered -= bc;
            cnt[best] = 0;
            side[best] = 3;
            sep.push_back(best);
            if (sepLimit > 0 && (int)sep.size() > sepLimit) return false;     // (the cover loop is O(|S| |L|))
        }
        if (sepLimit > 0 && (int)sep.size() > sepLimit) return false;
        std::sort(sep.begin(), sep.end());
        std::vector<int> A, B;
        for (size_t i = 0; i < L.size(); i++) {
            int v = L[i];
            if (side[v] == 1) A.push_back(v); else if (side[v] == 2) B.push_back(v);
        }
        for (int v : L) { side[v] = 0; cnt[v] = 0; }
        // elimination order: A-subtree, B-subtree, separator  => push in reverse
        stack.push_back({std::move(sep), true});
        stack.push_back({std::move(B), false});
        stack.push_back({std::move(A), false});
    }
    return true;
}

// LSFM-MD: exact minimum degree on the block graph, the fallback for patterns on which the index
// bisection of LSFM-ND has no small separator (loop closures / multi-lap trajectories / dense aerial
// overlap: poses far apart in the index order share features, the cross edges of an index cut need a
// vertex cover of hundreds of poses and the fill explodes).  Spec (the oracle's twin restates it):
//   adjacency sets as bit rows; repeat m times: v = the remaining vertex of minimum degree (tie:
//   smallest index); emit v; every neighbour u of v gets adj(u) |= adj(v) minus {u, v}.
// Supernodes: consecutive columns k, k+1 are merged when v_{k+1} is a neighbour of v_k at its elimination
// and the structure grows by at most MD_RELAX blocks (fundamental supernodes: 0), at most MD_MAXW
// columns per node.
static const int MD_RELAX = 2, MD_MAXW = 96;
void lsfm_md_order(int m, const int *ptr, const int *adj, std::vector<int> &perm,
                   std::vector<int> &nodes)
{
    typedef unsigned long long W;
    const int nw = (m + 63) / 64;
    std::vector<W> bits((size_t)m * nw, 0);
    const int INF = 1 << 30;
    std::vector<int> deg(m);              // degree of the remaining vertices, INF once eliminated
    for (int v = 0; v < m; v++) {
        W *r = &bits[(size_t)v * nw];
        for (int p = ptr[v]; p < ptr[v + 1]; p++) r[adj[p] >> 6] |= 1ull << (adj[p] & 63);
        deg[v] = ptr[v + 1] - ptr[v];
    }
    std::vector<int> colcnt(m), nb, nzw;
    std::vector<char> nextIsNb(m, 0);
    perm.clear(); nodes.clear();
    perm.reserve(m);
    // degree buckets as bit rows over the vertices (smallest index of a bucket = first set bit); rows are
    // created on first use: the degrees that occur are far fewer than m
    std::vector<std::vector<W>> bucket(m + 1);
    std::vector<int> bcount(m + 1, 0);
    auto bucket_set = [&](int d, int v) {
        if (bucket[d].empty()) bucket[d].assign(nw, 0);
        bucket[d][v >> 6] |= 1ull << (v & 63);
        bcount[d]++;
    };
    auto bucket_clear = [&](int d, int v) { bucket[d][v >> 6] &= ~(1ull << (v & 63)); bcount[d]--; };
    for (int v = 0; v < m; v++) bucket_set(deg[v], v);
    int mind = 0;
    for (int k = 0; k < m; k++) {
        // minimum degree, smallest index among the ties
        while (bcount[mind] == 0) mind++;
        const int bd = mind;
        int v = 0;
        {
            const W *b = bucket[bd].data();
            int w = 0;
            while (b[w] == 0) w++;
            v = w * 64 + __builtin_ctzll(b[w]);
        }
        bucket_clear(bd, v);
        deg[v] = INF;
        perm.push_back(v);
        colcnt[k] = bd;
        const W *rv = &bits[(size_t)v * nw];
        nb.clear(); nzw.clear();
        for (int w = 0; w < nw; w++) {
            W x = rv[w];
            if (x) nzw.push_back(w);
            while (x) { int b = __builtin_ctzll(x); nb.push_back(w * 64 + b); x &= x - 1; }
        }
        // adj(u) |= adj(v) \ {u, v}: only the words adj(v) occupies change; the degree follows the new bits
        const W vbit = 1ull << (v & 63);
        const int vw = v >> 6;
        for (int u : nb) {
            W *ru = &bits[(size_t)u * nw];
            int d = deg[u] - 1;                            // v leaves
            ru[vw] &= ~vbit;
            const W ubit = 1ull << (u & 63);
            const int uw = u >> 6;
            for (int w : nzw) {
                W add = rv[w] & ~ru[w];
                if (w == uw) add &= ~ubit;
                d += __builtin_popcountll(add);
                ru[w] |= add;
            }
            if (d != deg[u]) {
                bucket_clear(deg[u], u);
                bucket_set(d, u);
                deg[u] = d;
                if (d < mind) mind = d;
            }
        }
    }
    // was the vertex emitted at k+1 a neighbour of the one emitted at k (at ITS elimination)?  Row v keeps
    // its neighbourhood at elimination time (nothing touches a row once its vertex is gone).
    for (int k = 0; k + 1 < m; k++) {
        const W *rv = &bits[(size_t)perm[k] * nw];
        const int u = perm[k + 1];
        nextIsNb[k] = (rv[u >> 6] >> (u & 63)) & 1ull;
    }
    nodes.push_back(0);
    int width = 0;
    for (int k = 0; k < m; k++) {
        width++;
        bool merge = k + 1 < m && nextIsNb[k] && width < MD_MAXW &&
                     colcnt[k + 1] + 1 - colcnt[k] >= 0 && colcnt[k + 1] + 1 - colcnt[k] <= MD_RELAX;
        if (!merge) { nodes.push_back(k + 1); width = 0; }
    }
}

// The ordering rule of the library: LSFM-ND, unless one of its separators has more than
// ND_MAX_SEPARATOR poses -- then LSFM-MD.
static const int ND_MAX_SEPARATOR = 64;
void lsfm_order(int m, const int *ptr, const int *adj, std::vector<int> &perm, std::vector<int> &nodes)
{
    if (!nd_order_impl(m, ptr, adj, perm, nodes, ND_MAX_SEPARATOR)) lsfm_md_order(m, ptr, adj, perm, nodes);
}

namespace {

struct JoinSym {
    int m = 0;
    // m <= 32 (the thousands of joins of the lower tree levels): identity ordering, one dense supernode without
    // struct rows, children or parent.  Nothing is stored for such a join -- build_symbolic writes its
    // descriptor, pose entries and slots directly (a tree level of 1,749 two-pose joins took 0.64 ms of
    // vector allocations, longer than the Schur kernel it is meant to hide behind)
    bool small = false;
    int nsn() const { return small ? 1 : (int)structs.size(); }
    std::vector<int> perm, iperm, nodes, snOf;
    std::vector<std::vector<int>> structs, children;
    std::vector<int> parent, level;
};

typedef unsigned long long u64;
const u64 M22 = (1ull << 22) - 1;

void analyse_join(int m, const u64 *keys, int nk, JoinSym &J)
{
    J.m = m;
    if (m <= 32) { J.small = true; return; }
    // symmetric adjacency without self loops
    std::vector<int> ptr(m + 1, 0);
    for (int i = 0; i < nk; i++) {
        int a = (int)((keys[i] >> 22) & M22), b = (int)(keys[i] & M22);
        if (a != b) { ptr[a + 1]++; ptr[b + 1]++; }
    }
    for (int i = 0; i < m; i++) ptr[i + 1] += ptr[i];
    std::vector<int> adj(ptr[m]), fill(ptr.begin(), ptr.end() - 1);
    for (int i = 0; i < nk; i++) {
        int a = (int)((keys[i] >> 22) & M22), b = (int)(keys[i] & M22);
        if (a != b) { adj[fill[a]++] = b; adj[fill[b]++] = a; }
    }
    // the keys are sorted by (row, col) with row <= col: vertex v first receives its smaller
    // neighbours (rows a < v, in ascending a), then the larger ones (row v, ascending col) -- every
    // adjacency list is already ascending, no sort needed
    lsfm_order(m, ptr.data(), adj.data(), J.perm, J.nodes);
    if ((int)J.perm.size() != m) throw std::runtime_error("ordering lost vertices");
    J.iperm.assign(m, 0);
    for (int k = 0; k < m; k++) J.iperm[J.perm[k]] = k;
    int ns = (int)J.nodes.size() - 1;
    J.snOf.assign(m, 0);
    for (int s = 0; s < ns; s++)
        for (int c = J.nodes[s]; c < J.nodes[s + 1]; c++) J.snOf[c] = s;
    J.structs.assign(ns, {});
    J.children.assign(ns, {});
    J.parent.assign(ns, -1);
    J.level.assign(ns, 0);
    std::vector<int> mark(m, -1);
    for (int s = 0; s < ns; s++) {
        int end = J.nodes[s + 1];
        std::vector<int> &st = J.structs[s];
        for (int c = J.nodes[s]; c < end; c++) {
            int v = J.perm[c];
            for (int p = ptr[v]; p < ptr[v + 1]; p++) {
                int r = J.iperm[adj[p]];
                if (r >= end && mark[r] != s) { mark[r] = s; st.push_back(r); }
            }
        }
        for (int ch : J.children[s])
            for (int r : J.structs[ch])
                if (r >= end && mark[r] != s) { mark[r] = s; st.push_back(r); }
        std::sort(st.begin(), st.end());
        if (!st.empty()) {
            int par = J.snOf[st[0]];
            J.parent[s] = par;
            J.children[par].push_back(s);
        }
        int lev = 0;
        for (int ch : J.children[s]) lev = std::max(lev, J.level[ch] + 1);
        J.level[s] = lev;
    }
}

} // namespace

void build_symbolic(int K, const std::vector<int> &m, const std::vector<int> &posePre,
                    const u64 *keys, size_t nkeys, const std::vector<int> &sOff,
                    BatchSymbolic &out, int nthreads)
{
    std::vector<JoinSym> js(K);
    const int nthreads0 = std::min(nthreads, 8);
    nthreads = std::max(1, std::min(nthreads, K));
    // threads pay only for joins that need a graph (m > 32); the rest is a few nanoseconds per key
    long long work = 0;
    for (int k = 0; k < K; k++) if (m[k] > 32) work += sOff[k + 1] - sOff[k];
    if (work < 20000) nthreads = 1;
    std::vector<std::string> errs(nthreads);
    auto worker = [&](int tid) {
        try {
            for (int k = tid; k < K; k += nthreads)
                analyse_join(m[k], keys + sOff[k], sOff[k + 1] - sOff[k], js[k]);
        } catch (const std::exception &e) { errs[tid] = e.what(); }
    };
    if (nthreads == 1) worker(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; t++) th.emplace_back(worker, t);
        for (auto &t : th) t.join();
    }
    for (auto &e : errs) if (!e.empty()) throw std::runtime_error(e);

    out = BatchSymbolic();
    int totPose = posePre[K];
    out.poseSn.assign(totPose, -1);
    out.poseLcol.assign(totPose, 0);
    out.perm.assign(totPose, 0);
    out.slot.resize(nkeys);
    std::vector<int> snBase(K + 1, 0);
    for (int k = 0; k < K; k++) snBase[k + 1] = snBase[k] + js[k].nsn();
    int nsTot = snBase[K];
    out.sn.resize(nsTot);
    // per-join offsets into the flat arrays (struct rows, children, fronts), so that the joins -- and below the
    // key ranges of the slot map -- can be filled in by several host threads: the symbolic phase of the top tree
    // levels (K = 1..3 joins of thousands of poses) is longer than the GPU work it overlaps with
    std::vector<long long> soK(K + 1, 0), coK(K + 1, 0), foK(K + 1, 0);
    int maxLevel = 0;
    for (int k = 0; k < K; k++) {
        long long so = soK[k], co = coK[k], fo = foK[k];
        if (js[k].small) {
            const long long fs = 6 * (long long)js[k].m;
            fo += (fs + 1) * fs;
            fo = (fo + 1) & ~1ll;
        }
        for (size_t s = 0; s < js[k].structs.size(); s++) {
            so += (long long)js[k].structs[s].size();
            co += (long long)js[k].children[s].size();
            maxLevel = std::max(maxLevel, js[k].level[s]);
            const long long fs = 6 * (long long)((js[k].nodes[s + 1] - js[k].nodes[s]) + (long long)js[k].structs[s].size());
            fo += (fs + 1) * fs;
            fo = (fo + 1) & ~1ll;
        }
        soK[k + 1] = so; coK[k + 1] = co; foK[k + 1] = fo;
    }
    out.structIdx.resize(soK[K]);
    out.relIdx.assign(soK[K], -1);
    out.childIdx.resize(coK[K]);
    const long long fo = foK[K];
    std::vector<double> flopsK(K, 0.0);
    std::vector<int> maxFdimK(K, 0);
    // tasks: run fn(0..n-1) on up to `nthreads` threads (static interleave); exceptions are re-thrown
    const int poolThreads = std::max(1, nthreads0);
    auto parallel_for = [&](int n, bool big, const std::function<void(int)> &fn) {
        const int T = big ? std::min(poolThreads, n) : 1;
        if (T <= 1) { for (int i = 0; i < n; i++) fn(i); return; }
        std::vector<std::string> er(T);
        auto w = [&](int tid) {
            try { for (int i = tid; i < n; i += T) fn(i); } catch (const std::exception &e) { er[tid] = e.what(); }
        };
        std::vector<std::thread> th;
        for (int t = 1; t < T; t++) th.emplace_back(w, t);
        w(0);
        for (auto &t : th) t.join();
        for (auto &e : er) if (!e.empty()) throw std::runtime_error(e);
    };
    auto fill_join = [&](int k) {
        JoinSym &J = js[k];
        long long so = soK[k], co = coK[k], fo = foK[k];
        if (J.small) {
            SnodeDesc &d = out.sn[snBase[k]];
            d.join = k;
            d.first = 0;
            d.ncols = J.m;
            d.structOff = (int)so;
            d.nstruct = 0;
            d.parent = -1;
            d.childOff = (int)co;
            d.nchild = 0;
            d.poseOff = posePre[k];
            d.level = 0;
            d.frontOff = fo;
            maxFdimK[k] = J.m;
            const double nc = 6.0 * J.m;
            for (int q = 0; q < 6 * J.m; q++) { double c = nc - q; flopsK[k] += c * c; }
            for (int c = 0; c < J.m; c++) {
                out.poseSn[posePre[k] + c] = snBase[k];
                out.poseLcol[posePre[k] + c] = c;
                out.perm[posePre[k] + c] = c;
            }
            return;
        }
        int ns = (int)J.structs.size();
        for (int s = 0; s < ns; s++) {
            SnodeDesc &d = out.sn[snBase[k] + s];
            d.join = k;
            d.first = J.nodes[s];
            d.ncols = J.nodes[s + 1] - J.nodes[s];
            d.structOff = (int)so;
            d.nstruct = (int)J.structs[s].size();
            std::copy(J.structs[s].begin(), J.structs[s].end(), out.structIdx.begin() + so);
            so += d.nstruct;
            d.parent = J.parent[s] < 0 ? -1 : snBase[k] + J.parent[s];
            d.childOff = (int)co;
            d.nchild = (int)J.children[s].size();
            for (int ch : J.children[s]) out.childIdx[co++] = snBase[k] + ch;
            d.poseOff = posePre[k];
            d.level = J.level[s];
            long long fdim = d.ncols + d.nstruct;
            long long fs = 6 * fdim;
            d.frontOff = fo;
            fo += (fs + 1) * fs;
            fo = (fo + 1) & ~1ll;
            maxFdimK[k] = std::max(maxFdimK[k], (int)fdim);
            double nc = 6.0 * d.ncols, nr = 6.0 * d.nstruct;
            for (int q = 0; q < 6 * d.ncols; q++) { double c = (nc - q) + nr; flopsK[k] += c * c; }
        }
        // relative indices of every supernode's struct inside its parent's front
        for (int s = 0; s < ns; s++) {
            int par = J.parent[s];
            if (par < 0) continue;
            const SnodeDesc &d = out.sn[snBase[k] + s];
            const std::vector<int> &ps = J.structs[par];
            int pf = J.nodes[par], pe = J.nodes[par + 1], pn = pe - pf;
            size_t q = 0;
            for (int i = 0; i < d.nstruct; i++) {
                int r = out.structIdx[d.structOff + i];
                int pos;
                if (r < pe) pos = r - pf;
                else {
                    while (q < ps.size() && ps[q] < r) q++;
                    if (q >= ps.size() || ps[q] != r) throw std::runtime_error("symbolic: struct not nested");
                    pos = pn + (int)q;
                }
                out.relIdx[d.structOff + i] = pos;
            }
        }
        for (int c = 0; c < J.m; c++) {
            int s = J.snOf[c];
            out.poseSn[posePre[k] + J.perm[c]] = snBase[k] + s;
            out.poseLcol[posePre[k] + J.perm[c]] = c - J.nodes[s];
            out.perm[posePre[k] + c] = J.perm[c];
        }
    };
    parallel_for(K, work >= 20000, fill_join);
    for (int k = 0; k < K; k++) { out.maxFdim = std::max(out.maxFdim, maxFdimK[k]); out.flops += flopsK[k]; }
    // S slot -> front block: pieces of at most SLOT_PIECE keys of one join
    const int SLOT_PIECE = 8192;
    struct Piece { int k, i0, i1; };
    std::vector<Piece> pieces;
    for (int k = 0; k < K; k++)
        for (int i = sOff[k]; i < sOff[k + 1]; i += SLOT_PIECE) pieces.push_back({k, i, std::min(sOff[k + 1], i + SLOT_PIECE)});
    auto fill_slots = [&](int pi) {
        const Piece &P = pieces[pi];
        const int k = P.k;
        const JoinSym &J = js[k];
        if (J.small) {
            // identity ordering, one front: S(a, b) with a <= b is the front's block (row b, col a), transposed
            for (int i = P.i0; i < P.i1; i++) {
                SlotMap &sm = out.slot[i];
                sm.sn = snBase[k];
                sm.lcol = (int)((keys[i] >> 22) & M22);
                sm.lrow = (int)(keys[i] & M22);
                sm.pad = 0;
                sm.transpose = 1;
            }
            return;
        }
        for (int i = P.i0; i < P.i1; i++) {
            int a = (int)((keys[i] >> 22) & M22), b = (int)(keys[i] & M22);
            int pa = J.iperm[a], pb = J.iperm[b];
            int c = std::min(pa, pb), r = std::max(pa, pb);
            int s = J.snOf[c];
            SlotMap &sm = out.slot[i];
            sm.sn = snBase[k] + s;
            sm.lcol = c - J.nodes[s];
            sm.pad = 0;
            if (r < J.nodes[s + 1]) sm.lrow = r - J.nodes[s];
            else {
                const std::vector<int> &st = J.structs[s];
                auto it = std::lower_bound(st.begin(), st.end(), r);
                if (it == st.end() || *it != r) throw std::runtime_error("symbolic: S block outside front");
                sm.lrow = (J.nodes[s + 1] - J.nodes[s]) + (int)(it - st.begin());
            }
            // front block (row r, col c) = S_full(perm[r], perm[c]); slot holds S_full(a, b)
            sm.transpose = (pa == r) ? 0 : 1;
            // diagonal blocks: the reference reads only the UPPER triangle (LinearSFMImp.cpp:2221-2230,
            // 2474-2481); the front keeps the lower triangle, so take it from the transposed block
            if (a == b) sm.transpose = 1;
        }
    };
    parallel_for((int)pieces.size(), work >= 20000, fill_slots);
    out.frontDoubles = fo;
    // level sets
    out.levelPtr.assign(maxLevel + 2, 0);
    for (auto &d : out.sn) out.levelPtr[d.level + 1]++;
    for (int l = 0; l <= maxLevel; l++) out.levelPtr[l + 1] += out.levelPtr[l];
    out.levelSn.resize(nsTot);
    std::vector<int> fillp(out.levelPtr.begin(), out.levelPtr.end() - 1);
    for (int s = 0; s < nsTot; s++) out.levelSn[fillp[out.sn[s].level]++] = s;
}
