// Cell-sorted, shared-memory-tiled kernels for the 3D uGIMP step (filled in below).
#pragma once
#include "mpm_types.cuh"

struct TiledState {
    int enabled;
};
static inline void tiled_state_init(TiledState &t) { t.enabled = 0; }
static inline void tiled_state_free(TiledState &t) {}
static inline void tiled_on_upload(TiledState &t) {}
