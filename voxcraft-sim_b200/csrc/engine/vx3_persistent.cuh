// On-chip persistent step kernel for a single small collision-free body (see DESIGN.md "persistent path").
#pragma once
#include <cuda_runtime.h>

#include <vector>

#include "vx3_kernels.cuh"

namespace vx3 {

struct PersistentPlan {
    bool ok = false;
};

inline void persistent_plan(PersistentPlan &p, const std::vector<SimC> &, bool, bool, bool, const cudaDeviceProp &) { p.ok = false; }
inline int persistent_run(PersistentPlan &, const Dev &, cudaStream_t, long long, bool, long long *) { return -1; }
inline void persistent_free(PersistentPlan &) {}

} // namespace vx3
