#pragma once
#include "ops.h"
#include "../../include/linearsfm_b200.h"

std::vector<MapHandle> upload_maps(Context &ctx, const lsfm_map *maps, int K, bool validate);
void download_map(Context &ctx, const MapHandle &h, lsfm_map *out);
void download_state(Context &ctx, const MapHandle &h, int *stno, double *stVal);
void free_host_map(lsfm_map *m);
// forget per-device state (the upload-complete event) when the context goes away / moves to another GPU
void mapio_shutdown();
