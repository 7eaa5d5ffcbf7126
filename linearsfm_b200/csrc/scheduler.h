#pragma once
#include "ops.h"

std::vector<MapHandle> transform_or_pass(Context &ctx, const std::vector<MapHandle> &in,
                                         const std::vector<int> &newRef);
// Runs the merge tree on `level` (leaf maps or any intermediate level). Returns the remaining maps
// (one map unless max_levels stops earlier). first_index = global 0-based index of level[0].
std::vector<MapHandle> solve_tree_stereo(Context &ctx, std::vector<MapHandle> level, bool verbose,
                                         int first_index, int max_levels);
MapHandle final_rebase_stereo(Context &ctx, const MapHandle &root);

// mono merge tree incl. the final re-base (LinearSFMImp.cpp:6511-6658)
std::vector<MapHandle> solve_tree_mono(Context &ctx, std::vector<MapHandle> level, bool verbose);
