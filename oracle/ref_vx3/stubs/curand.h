// TEST INFRASTRUCTURE — host stand-in for cuRAND: VX3.cuh's random() (used only to pick the ONE link whose strain the
// reference samples for divergence, VX3_VoxelyzeKernel.cu:271-278) and VX3_dictionary's hash seed.
#pragma once
#include <cstdint>
struct curandState_t { uint64_t s; };
inline void curand_init(unsigned long long seed, unsigned long long, unsigned long long, curandState_t *st) { st->s = seed * 6364136223846793005ull + 1442695040888963407ull; }
inline unsigned int curand(curandState_t *st) { st->s = st->s * 6364136223846793005ull + 1442695040888963407ull; return (unsigned int)(st->s >> 33); }
