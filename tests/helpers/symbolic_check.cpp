// Host unit check of the symbolic phase of the block multifrontal Cholesky (linearsfm_b200/csrc/chol_symbolic.cpp,
// the replacement of cholmod_amd + cholmod_analyze_p, LinearSFMImp.cpp:2413-2440).  No GPU, no oracle: every
// property is checked against an independent dense boolean elimination of the permuted block pattern.
//
//   symbolic_check <seed> <cases>      prints "violations N" and exits with N != 0
//
// Checked per batch (random mixes of small joins, m <= 32 = the fast path without per-join storage, and large
// joins with band + hub + random patterns):
//   P1  perm is a permutation per join; poseSn / poseLcol agree with the supernode column ranges
//   P2  small joins: exactly one supernode, identity order, no struct rows, no parent
//   P3  every column's true fill pattern (dense elimination) lies inside its supernode's front rows
//   P4  relIdx places every struct row of a child at the same block row inside the parent's front
//   P5  slot map: S(a, b) lands on front block (row max(pa, pb), col min(pa, pb)) of the owning supernode,
//       transposed exactly when the slot's (row, col) order differs from the front's, diagonal blocks transposed
//   P6  fronts are disjoint, (6 f + 1) x 6 f doubles each, even offsets; frontDoubles is the end
//   P7  levels: a parent's level exceeds its children's; levelPtr / levelSn partition the supernodes by level
//   P8  flops = sum over scalar columns of (column count)^2, maxFdim = the largest front
#include "chol_symbolic.h"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>

typedef unsigned long long u64;
static int violations = 0;
#define CHECK(c, ...) do { if (!(c)) { if (violations < 20) { printf("VIOLATION %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); } violations++; } } while (0)

static void check_batch(int K, const std::vector<int> &m, const std::vector<u64> &keys, const std::vector<int> &sOff,
                        int nthreads)
{
    std::vector<int> posePre(K + 1, 0);
    for (int k = 0; k < K; k++) posePre[k + 1] = posePre[k] + m[k];
    BatchSymbolic sym;
    build_symbolic(K, m, posePre, keys.data(), keys.size(), sOff, sym, nthreads);
    const int ns = (int)sym.sn.size();
    // supernodes of a join are contiguous and in elimination order
    std::vector<int> snFirst(K + 1, ns);
    for (int s = ns - 1; s >= 0; s--) snFirst[sym.sn[s].join] = s;
    long long frontEnd = 0;
    double flops = 0.0;
    int maxF = 0;
    for (int s = 0; s < ns; s++) {
        const SnodeDesc &d = sym.sn[s];
        const long long fs = 6ll * (d.ncols + d.nstruct);
        CHECK(d.frontOff == frontEnd, "front %d offset %lld expected %lld", s, d.frontOff, frontEnd);        // P6
        CHECK((d.frontOff & 1) == 0, "front %d offset odd", s);
        frontEnd = (d.frontOff + (fs + 1) * fs + 1) & ~1ll;
        maxF = std::max(maxF, d.ncols + d.nstruct);
        for (int q = 0; q < 6 * d.ncols; q++) { double c = (6.0 * d.ncols - q) + 6.0 * d.nstruct; flops += c * c; }
        CHECK(d.poseOff == posePre[d.join], "poseOff of supernode %d", s);
        if (d.parent >= 0) {                                                                                  // P7
            CHECK(sym.sn[d.parent].join == d.join && d.parent > s, "parent of %d", s);
            CHECK(sym.sn[d.parent].level > d.level, "level of parent of %d", s);
        }
        for (int c = 0; c < d.nchild; c++) CHECK(sym.sn[sym.childIdx[d.childOff + c]].parent == s, "child list of %d", s);
    }
    CHECK(sym.frontDoubles == frontEnd, "frontDoubles %lld vs %lld", sym.frontDoubles, frontEnd);
    CHECK(sym.maxFdim == maxF, "maxFdim %d vs %d", sym.maxFdim, maxF);                                        // P8
    CHECK(std::fabs(sym.flops - flops) <= 1e-9 * std::max(1.0, flops), "flops %.17g vs %.17g", sym.flops, flops);
    {
        std::vector<int> seen(ns, 0);
        CHECK((int)sym.levelSn.size() == ns && sym.levelPtr.front() == 0 && sym.levelPtr.back() == ns, "level sets");
        for (size_t l = 0; l + 1 < sym.levelPtr.size(); l++)
            for (int i = sym.levelPtr[l]; i < sym.levelPtr[l + 1]; i++) {
                CHECK(sym.sn[sym.levelSn[i]].level == (int)l, "levelSn[%d]", i);
                seen[sym.levelSn[i]]++;
            }
        for (int s = 0; s < ns; s++) CHECK(seen[s] == 1, "supernode %d in %d level sets", s, seen[s]);
    }
    for (int k = 0; k < K; k++) {
        const int mk = m[k], p0 = posePre[k];
        const int s0 = snFirst[k], s1 = (k + 1 < K) ? snFirst[k + 1] : ns;
        // P1
        std::vector<int> ip(mk, -1);
        for (int c = 0; c < mk; c++) {
            const int v = sym.perm[p0 + c];
            CHECK(v >= 0 && v < mk && ip[v] < 0, "perm of join %d", k);
            if (v >= 0 && v < mk) ip[v] = c;
        }
        std::vector<int> snOfCol(mk, -1);
        int cols = 0;
        for (int s = s0; s < s1; s++) {
            CHECK(sym.sn[s].first == cols, "column ranges of join %d", k);
            for (int c = 0; c < sym.sn[s].ncols; c++) snOfCol[cols++] = s;
        }
        CHECK(cols == mk, "join %d: supernodes cover %d of %d columns", k, cols, mk);
        for (int v = 0; v < mk; v++) {
            const int c = ip[v];
            CHECK(sym.poseSn[p0 + v] == snOfCol[c] && sym.poseLcol[p0 + v] == c - sym.sn[snOfCol[c]].first, "poseSn / poseLcol of join %d", k);
        }
        if (mk <= 32) {                                                                                       // P2
            CHECK(s1 - s0 == 1, "small join %d has %d supernodes", k, s1 - s0);
            CHECK(sym.sn[s0].nstruct == 0 && sym.sn[s0].parent < 0 && sym.sn[s0].nchild == 0 && sym.sn[s0].level == 0, "small join %d", k);
            for (int c = 0; c < mk; c++) CHECK(sym.perm[p0 + c] == c, "small join %d: order not the identity", k);
        }
        // dense boolean elimination of the permuted pattern (lower triangle: L[r][c], r > c)
        std::vector<std::vector<char>> L(mk, std::vector<char>(mk, 0));
        for (int i = sOff[k]; i < sOff[k + 1]; i++) {
            const int a = (int)((keys[i] >> 22) & 0x3fffff), b = (int)(keys[i] & 0x3fffff);
            if (a == b) continue;
            const int r = std::max(ip[a], ip[b]), c = std::min(ip[a], ip[b]);
            L[r][c] = 1;
        }
        for (int c = 0; c < mk; c++) {
            int first = -1;
            for (int r = c + 1; r < mk; r++)
                if (L[r][c]) {
                    if (first < 0) { first = r; continue; }
                    L[r][first] = 1;             // fill: the column's pattern is passed to its parent column
                }
        }
        // P3: rows of column c are inside the front of its supernode
        for (int s = s0; s < s1; s++) {
            const SnodeDesc &d = sym.sn[s];
            std::vector<char> inFront(mk, 0);
            for (int c = d.first; c < d.first + d.ncols; c++) inFront[c] = 1;
            int prev = d.first + d.ncols - 1;
            for (int i = 0; i < d.nstruct; i++) {
                const int r = sym.structIdx[d.structOff + i];
                CHECK(r > prev && r < mk, "struct rows of supernode %d not ascending / in range", s);
                prev = r;
                if (r >= 0 && r < mk) inFront[r] = 1;
            }
            for (int c = d.first; c < d.first + d.ncols; c++)
                for (int r = c + 1; r < mk; r++)
                    CHECK(!L[r][c] || inFront[r], "join %d: fill (%d,%d) outside the front of supernode %d", k, r, c, s);
            // P4
            if (d.parent >= 0) {
                const SnodeDesc &p = sym.sn[d.parent];
                CHECK(d.nstruct > 0 && snOfCol[sym.structIdx[d.structOff]] == d.parent, "parent of %d is not the owner of its first struct row", s);
                for (int i = 0; i < d.nstruct; i++) {
                    const int r = sym.structIdx[d.structOff + i], rel = sym.relIdx[d.structOff + i];
                    const int back = rel < p.ncols ? p.first + rel : sym.structIdx[p.structOff + rel - p.ncols];
                    CHECK(rel >= 0 && rel < p.ncols + p.nstruct && back == r, "relIdx of supernode %d row %d", s, i);
                }
            } else CHECK(d.nstruct == 0, "root supernode %d has struct rows", s);
        }
        // P5
        for (int i = sOff[k]; i < sOff[k + 1]; i++) {
            const int a = (int)((keys[i] >> 22) & 0x3fffff), b = (int)(keys[i] & 0x3fffff);
            const int pa = ip[a], pb = ip[b], c = std::min(pa, pb), r = std::max(pa, pb);
            const SlotMap &sm = sym.slot[i];
            const SnodeDesc &d = sym.sn[snOfCol[c]];
            CHECK(sm.sn == snOfCol[c] && sm.lcol == c - d.first, "slot %d: supernode / column", i);
            const int row = sm.lrow < d.ncols ? d.first + sm.lrow
                                              : (sm.lrow - d.ncols < d.nstruct ? sym.structIdx[d.structOff + sm.lrow - d.ncols] : -1);
            CHECK(row == r, "slot %d: row %d expected %d", i, row, r);
            CHECK(sm.transpose == ((a == b || pa != r) ? 1 : 0), "slot %d: transpose flag", i);
        }
    }
}

int main(int argc, char **argv)
{
    const unsigned seed = argc > 1 ? (unsigned)atoi(argv[1]) : 1u;
    const int cases = argc > 2 ? atoi(argv[2]) : 60;
    std::mt19937 rng(seed);
    for (int t = 0; t < cases; t++) {
        const int mode = t % 5;      // 0 small only, 1 mixed, 2 large, 3 tiny / empty, 4 one tree level of equal small joins
        int K = 1 + (int)(rng() % 24);
        if (mode == 4) K = 300 + (int)(rng() % 300);
        if (mode == 2) K = 1 + (int)(rng() % 4);
        std::vector<int> m(K), sOff(K + 1, 0);
        std::vector<u64> keys;
        const int eq = 2 << (rng() % 4);
        for (int k = 0; k < K; k++) {
            int mk = mode == 0 ? (int)(rng() % 33) : mode == 1 ? (int)(rng() % 90) : mode == 2 ? 33 + (int)(rng() % 260)
                   : mode == 3 ? (int)(rng() % 4) : eq;
            m[k] = mk;
            const int band = 1 + (int)(rng() % 6), nh = (int)(rng() % 4);
            const double dens = (rng() % 100) / 2000.0;
            std::vector<int> hubs;
            for (int h = 0; h < nh && mk > 0; h++) hubs.push_back((int)(rng() % mk));
            for (int a = 0; a < mk; a++)
                for (int b = a; b < mk; b++) {
                    bool on = a == b || b - a <= band || (rng() % 100000) / 100000.0 < dens;
                    for (int h : hubs) on = on || a == h || b == h;
                    if (mode == 4) on = true;
                    if (on) keys.push_back(((u64)k << 44) | ((u64)a << 22) | (u64)b);
                }
            sOff[k + 1] = (int)keys.size();
        }
        check_batch(K, m, keys, sOff, 1 + (int)(rng() % 8));
    }
    // two large joins: a sequential chain with former-origin hubs (index bisection, LSFM-ND) and the same chain
    // with loop closures between distant indices (no small index separator: LSFM-MD)
    for (int closed = 0; closed < 2; closed++) {
        const int mk = closed ? 900 : 1500;
        std::vector<std::vector<char>> on(mk, std::vector<char>(mk, 0));
        for (int a = 0; a < mk; a++) for (int b = a; b < mk && b <= a + 5; b++) on[a][b] = 1;
        for (int h = 0; h < mk; h += 128) for (int b = h; b < std::min(mk, h + 128); b++) on[h][b] = 1;
        if (closed)
            for (int e = 0; e < 4 * mk; e++) {
                int a = (int)(rng() % mk), b = (int)((a + mk / 3 + rng() % 40) % mk);
                on[std::min(a, b)][std::max(a, b)] = 1;
            }
        std::vector<u64> keys;
        for (int a = 0; a < mk; a++) for (int b = a; b < mk; b++) if (on[a][b]) keys.push_back(((u64)a << 22) | (u64)b);
        check_batch(1, {mk}, keys, {0, (int)keys.size()}, 4);
    }
    printf("violations %d\n", violations);
    return violations != 0;
}
