// pm_kernels_tiled.cuh — placeholder until the TMA-staged tiled pressure sweep lands.
#pragma once
#include <string>
#include "pm_common.cuh"
struct TiledPlan { int sweeps = 1; };
static inline bool tiled_supported(const pm_config&, const KP&) { return false; }
static inline bool tiled_create(TiledPlan*, const pm_config&, const KP&, double*, double*, double*, int, std::string* e) { *e = "not built"; return false; }
static inline void tiled_destroy(TiledPlan*) {}
