// Stereo local-map builder: two-view bundle adjustment of every pair of consecutive
// stereo frames -> the local maps (state + block information U, W, V) LinearSFM joins.
#pragma once
#include "device.h"
#include "../../include/linearsfm_b200.h"

// Builds K local maps on the device (one CTA per map, all maps in one launch) and returns them as
// host lsfm_map structs (malloc'd, lsfm_free_map releases them).  iters_done may be null.
void build_localmaps_stereo(Context &ctx, const lsfm_stereo_pair *pairs, int K, const lsfm_stereo_cam &cam,
                            int max_iters, double tol, lsfm_map *out, int *iters_done);
