// Host-side interfaces of the batched ("segmented") device operators of the hot path.
// Each operator processes EVERY map / join of one merge-tree level in a handful of launches.
#pragma once
#include <memory>
#include "device.h"
#include <vector>
#include <functional>

// Device view of the maps an operator works on: descriptor array + prefix sums of their sizes.
struct OpMaps {
    int K = 0;
    std::vector<DMap> h;
    std::vector<int> posePre, featPre, uPre, wPre;     // K+1 prefix sums
    // descriptor array and the four prefix arrays live in ONE device blob filled by ONE H2D copy
    template <class T> struct View { T *p = nullptr; };
    DevBuf<char> blob;
    View<DMap> d;
    View<int> dPosePre, dFeatPre, dUPre, dWPre;
    int totPose = 0, totFeat = 0, totU = 0, totW = 0;
    void build(const std::vector<MapHandle> &maps, cudaStream_t s);
    void build(const std::vector<DMap> &maps, cudaStream_t s);
};

// R(map) of SURVEY 8(d): bytes of one map's arrays (stVal+stno, U+Ui+Uj, W+photo+feature, V+FBlock)
static inline double map_bytes(const DMap &d)
{
    return 12.0 * (6.0 * d.m + 3.0 * d.n) + 296.0 * d.nU + 152.0 * d.nW + 76.0 * d.n;
}

// Allocate one arena holding `shapes.size()` maps with the given sizes; fills pointers of `out`.
std::vector<MapHandle> alloc_maps(Context &ctx, std::vector<DMap> &shapes);
size_t map_layout(DMap &shape, char *base);

// a2-a5: re-express each map in the frame of pose `newRef[k]` (LinearSFMImp.cpp:349-1924).
std::vector<MapHandle> transform_stereo_batch(Context &ctx, const std::vector<MapHandle> &in,
                                              const std::vector<int> &newRef);

// a15: monocular frame transform (LinearSFMImp.cpp:3173-6509): Ref/ScaP/Fix of the target frame per map
std::vector<MapHandle> transform_mono_batch(Context &ctx, const std::vector<MapHandle> &in,
                                            const std::vector<int> &newRef, const std::vector<int> &newSca,
                                            const std::vector<int> &newFix);
// a16-a17: monocular join + solve (LinearSFMImp.cpp:7282-7874, 6756-7041)
std::vector<MapHandle> join_mono_batch(Context &ctx, const std::vector<MapHandle> &End,
                                       const std::vector<MapHandle> &Cur);

// a6-a13: linear join of End[k] (already in Cur[k]'s frame) with Cur[k]
// (LinearSFMImp.cpp:2551-2978 + 2119-2378). Returns the joint maps.
std::vector<MapHandle> join_stereo_batch(Context &ctx, const std::vector<MapHandle> &End,
                                         const std::vector<MapHandle> &Cur);

// a8-a13 alone: solve the joint system of every map given the right-hand sides
// (LinearSFMImp.cpp:2119-2378). eP: concatenated 6m per map, eF: concatenated 3n per map.
// Writes poseVal/featVal of the maps in place.
struct SolveDebug;   // optional capture of pattern / ordering for the parity tests
// mono gauge (device arrays, one entry per join): local index of the all-zero Ref pose, local scalar
// row of the pinned ScaP translation component, and its value Sign (LinearSFMImp.cpp:7797-7801)
struct MonoGauge { const int *refPose; const int *fixScalar; const int *sign; };
// Optional hooks of the stereo join (join.cu):
//   after_pattern : called once the S pattern is on its way to the host; queues device work that
//                   the pattern did not need (the join's W/V value copy and eF), so that it overlaps
//                   the host symbolic phase.
//   xhat[6 totFeat]: per joint feature the End-side and Cur-side estimates (zeros where absent);
//   split[K]       : number of End poses of every join (joint pose p < split -> End side).
//   With these, eP holds only the U part of the reference's eP and the Schur kernel folds the
//   W xhat_f part into the reduced right-hand side (see k_vinv in solve.cu).
struct SolveExtra {
    std::function<void()> after_pattern;
    const double *xhat = nullptr;
    const int *split = nullptr;
    // feature chunks: optional per-join seed (start indices ending with n, nullptr entries = default) and the
    // lists the pattern stage ended up with (one per join; nullptr when every chunk is the default width)
    const std::vector<std::shared_ptr<const std::vector<int>>> *chunkSeed = nullptr;
    std::vector<std::shared_ptr<const std::vector<int>>> *chunksOut = nullptr;
};
void solve_stereo_batch(Context &ctx, const OpMaps &J, const double *eP, const double *eF,
                        SolveDebug *dbg, const MonoGauge *gauge = nullptr, const SolveExtra *ex = nullptr);

struct SolveDebug {
    // for the first map of the batch only
    std::vector<int> rowptr, colidx;          // block CRS of S (reference's Sidxij, 2190-2205)
    std::vector<double> S;                    // 36 per block, row-major, diagonal blocks full
    std::vector<double> E;                    // reduced rhs
    std::vector<int> perm;                    // block elimination ordering
    std::vector<int> snode_first, snode_parent;
    double chol_flops = 0.0;
};
