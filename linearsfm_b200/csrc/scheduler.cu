// Merge-tree scheduler: row a14 of SURVEY 8(a).
//
// Reference: CLinearSFMImp::lmj_PF3D_Divide_ConquerStereo, LinearSFMImp.cpp:1926-2099 -- a
// sequential loop over levels and pairs.  Here every level is ONE batch: all End maps of the level
// are re-expressed in their partner's frame by one segmented Transform, all pairs are joined and
// solved by one segmented Join/Solve, and the re-base of the odd outputs (1997-2025) is one more
// segmented Transform.  Pairing, leftover handling (1940-1948) and the re-base rule are the
// reference's, evaluated on GLOBAL output indices so that a rank owning a power-of-two aligned
// slice of the leaves reproduces exactly what the sequential scheduler would do with that slice.
#include "scheduler.h"
#include <cstdio>
#include <cstdlib>
#include <chrono>

std::vector<MapHandle> transform_or_pass(Context &ctx, const std::vector<MapHandle> &in,
                                         const std::vector<int> &newRef)
{
    // maps already in the requested frame are passed through (LinearSFMImp.cpp:352-355)
    std::vector<MapHandle> todo;
    std::vector<int> refs, where;
    for (size_t i = 0; i < in.size(); i++)
        if (in[i].d.Ref != newRef[i]) { todo.push_back(in[i]); refs.push_back(newRef[i]); where.push_back((int)i); }
    std::vector<MapHandle> done = transform_stereo_batch(ctx, todo, refs);
    std::vector<MapHandle> out = in;
    for (size_t j = 0; j < where.size(); j++) out[where[j]] = done[j];
    return out;
}

std::vector<MapHandle> solve_tree_stereo(Context &ctx, std::vector<MapHandle> level, bool verbose,
                                         int first_index, int max_levels)
{
    int L = 0;
    long long base = first_index;             // global index of level[0] at the current level
    // max_levels < 0: run to the root.  max_levels >= 0: run exactly that many levels, even on a
    // single map (a leftover still goes through the re-base rule of its level).
    while ((max_levels < 0 && level.size() > 1) || (max_levels >= 0 && L < max_levels && !level.empty())) {
        int count = (int)level.size();
        int npairs = count / 2;
        bool leftover = (count % 2) != 0;
        if (base % 2 != 0) throw LsfmError(LSFM_ERR_ARG, "tree slice must start at an even index");
        std::vector<MapHandle> E(npairs), C(npairs);
        std::vector<int> refs(npairs);
        for (int i = 0; i < npairs; i++) {
            E[i] = level[2 * i];
            C[i] = level[2 * i + 1];
            refs[i] = C[i].d.Ref;
            if (verbose) {
                printf("Join Level %d Local Map %lld\n", L, base + 2 * i + 1);
                printf("Join Level %d Local Map %lld\n", L, base + 2 * i + 2);
                printf("Generate Level %d Local Map %lld\n\n", L + 1, base / 2 + i + 1);
            }
        }
        if (verbose && leftover) {
            printf("Join Level %d Local Map %lld\n", L, base + count);
            printf("Generate Level %d Local Map %lld\n\n", L + 1, base / 2 + npairs + 1);
        }
        static const bool dbg = getenv("LSFM_DEBUG") != nullptr;
        auto now = [&]() { if (dbg) cudaStreamSynchronize(ctx.stream); return std::chrono::steady_clock::now(); };
        auto t0 = now();
        std::vector<MapHandle> Et = transform_or_pass(ctx, E, refs);
        E.clear();
        auto t1 = now();
        std::vector<MapHandle> next = join_stereo_batch(ctx, Et, C);
        Et.clear(); C.clear();
        auto t2 = now();
        if (leftover) next.push_back(level[count - 1]);
        level.clear();
        base /= 2;
        // re-base every output with even (index+1) whose Ref is ahead of its first frame (1997-2025)
        std::vector<MapHandle> rb;
        std::vector<int> rbRef, rbIdx;
        std::vector<int> refAfter(next.size());
        for (size_t i = 0; i < next.size(); i++) {
            long long gi = base + (long long)i;
            refAfter[i] = next[i].d.Ref;
            if ((gi + 1) % 2 == 0 && next[i].d.Ref > next[i].d.FRef) {
                rb.push_back(next[i]); rbRef.push_back(next[i].d.FRef); rbIdx.push_back((int)i);
                refAfter[i] = next[i].d.FRef;
            }
        }
        // The even outputs are the End maps of the NEXT level; their Transform into the partner's
        // frame (1964) is independent of the re-base of the odd outputs, so both go into one
        // segmented launch set (the next level then finds End.Ref == Cur.Ref and passes through).
        bool more = (max_levels < 0) ? (next.size() > 1) : (L + 1 < max_levels);
        if (more)
            for (size_t i = 0; i + 1 < next.size(); i += 2)
                if (next[i].d.Ref != refAfter[i + 1]) {
                    rb.push_back(next[i]); rbRef.push_back(refAfter[i + 1]); rbIdx.push_back((int)i);
                }
        if (!rb.empty()) {
            std::vector<MapHandle> done = transform_stereo_batch(ctx, rb, rbRef);
            for (size_t j = 0; j < rbIdx.size(); j++) next[rbIdx[j]] = done[j];
        }
        if (dbg) {
            auto t3 = now();
            try { ctx.check_errors(); } catch (const LsfmError &e) {
                fprintf(stderr, "[level %2d] %s\n", L, e.what());
                throw;
            }
            auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
            fprintf(stderr, "[level %2d] pairs %5d  transform %7.3f ms  join+solve %7.3f ms  rebase(%zu) %7.3f ms\n", L,
                    npairs, ms(t0, t1), ms(t1, t2), rb.size(), ms(t2, t3));
        }
        level = std::move(next);
        L++;
    }
    if (getenv("LSFM_DEBUG")) {
        fprintf(stderr, "host work between a size read-back and the next launch (GPU idle): transform %.3f ms, join %.3f ms, pattern %.3f ms\n",
                ctx.idle_ms[0], ctx.idle_ms[1], ctx.idle_ms[2]);
        ctx.idle_ms[0] = ctx.idle_ms[1] = ctx.idle_ms[2] = 0.0;
        fprintf(stderr, "device pool: %lld cudaMalloc calls so far, %.1f MB reserved\n", DevicePool::get().misses,
                DevicePool::get().reserved / 1e6);
    }
    return level;
}

MapHandle final_rebase_stereo(Context &ctx, const MapHandle &root)
{
    // final re-base to the first frame (LinearSFMImp.cpp:2039-2063)
    if (root.d.Ref > root.d.FRef) {
        std::vector<MapHandle> in{root};
        std::vector<int> ref{root.d.FRef};
        return transform_stereo_batch(ctx, in, ref)[0];
    }
    return root;
}

// ---------------------------------------------------------------------------------------------
// Monocular merge tree: lmj_PF3D_Divide_ConquerMono, LinearSFMImp.cpp:6511-6658.  Same pairing and
// re-base rule as stereo; the Transform takes (Ref, ScaP, Fix) of the partner / of the first frame,
// and is skipped when both Ref and ScaP already match (3176).
// ---------------------------------------------------------------------------------------------
static std::vector<MapHandle> transform_mono_or_pass(Context &ctx, const std::vector<MapHandle> &in,
                                                     const std::vector<int> &r, const std::vector<int> &sc,
                                                     const std::vector<int> &fx)
{
    std::vector<MapHandle> todo;
    std::vector<int> r2, s2, f2, where;
    for (size_t i = 0; i < in.size(); i++)
        if (!(in[i].d.Ref == r[i] && in[i].d.ScaP == sc[i])) {
            todo.push_back(in[i]); r2.push_back(r[i]); s2.push_back(sc[i]); f2.push_back(fx[i]);
            where.push_back((int)i);
        }
    std::vector<MapHandle> done = transform_mono_batch(ctx, todo, r2, s2, f2);
    std::vector<MapHandle> out = in;
    for (size_t j = 0; j < where.size(); j++) out[where[j]] = done[j];
    return out;
}

std::vector<MapHandle> solve_tree_mono(Context &ctx, std::vector<MapHandle> level, bool verbose)
{
    int L = 0;
    while (level.size() > 1) {
        int count = (int)level.size();
        int npairs = count / 2;
        bool leftover = (count % 2) != 0;
        std::vector<MapHandle> E(npairs), C(npairs);
        std::vector<int> r(npairs), sc(npairs), fx(npairs);
        for (int i = 0; i < npairs; i++) {
            E[i] = level[2 * i];
            C[i] = level[2 * i + 1];
            r[i] = C[i].d.Ref; sc[i] = C[i].d.ScaP; fx[i] = C[i].d.Fix;          // 6549
            if (verbose) {
                printf("Join Level %d Local Map %d\n", L, 2 * i + 1);
                printf("Join Level %d Local Map %d\n", L, 2 * i + 2);
                printf("Generate Level %d Local Map %d\n\n", L + 1, i + 1);
            }
        }
        if (verbose && leftover) {
            printf("Join Level %d Local Map %d\n", L, count);
            printf("Generate Level %d Local Map %d\n\n", L + 1, npairs + 1);
        }
        std::vector<MapHandle> Et = transform_mono_or_pass(ctx, E, r, sc, fx);
        E.clear();
        std::vector<MapHandle> next = join_mono_batch(ctx, Et, C);
        Et.clear(); C.clear();
        if (leftover) next.push_back(level[count - 1]);
        level.clear();
        std::vector<MapHandle> rb;
        std::vector<int> rr, rs, rf, idx;
        for (size_t i = 0; i < next.size(); i++)
            if ((i + 1) % 2 == 0 && next[i].d.Ref > next[i].d.FRef) {           // 6576-6597
                rb.push_back(next[i]); rr.push_back(next[i].d.FRef); rs.push_back(next[i].d.FScaP);
                rf.push_back(next[i].d.FFix); idx.push_back((int)i);
            }
        if (!rb.empty()) {
            std::vector<MapHandle> done = transform_mono_or_pass(ctx, rb, rr, rs, rf);
            for (size_t j = 0; j < idx.size(); j++) next[idx[j]] = done[j];
        }
        level = std::move(next);
        L++;
    }
    if (!level.empty() && level[0].d.Ref > level[0].d.FRef) {                    // 6613-6630
        std::vector<MapHandle> in{level[0]};
        level[0] = transform_mono_or_pass(ctx, in, {level[0].d.FRef}, {level[0].d.FScaP}, {level[0].d.FFix})[0];
    }
    return level;
}
