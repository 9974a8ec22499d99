// Host-side material constants of a model: everything the step loop reads from a voxel or link material, produced
// straight into the ABI records (vx3_voxel_material / vx3_link_material) by three pure functions.
//
// Design (own code, not the reference's class tree): a material is (a) a stress-strain law — a StressLaw value built
// by one of three factories and combined for links by `in_series` — and (b) a handful of scalars.  Nothing is
// mutated in place and there is no dependent-material bookkeeping: the builder calls voxel_constants() once per palette
// entry and link_constants() once per distinct (material, material) pair, after every input is final.
//
// The VALUES have to equal what the reference's host library computes (a model built here is compared bit for bit
// with one exported from CVX_Sim::Import, tests/test_oracle_vs_ref.py), so each expression keeps the reference's
// precision — float members, double where C++ promotes — and cites the line that fixes it:
//   stress-strain models  src/old/VX_Material.cpp:300-487      eHat               src/old/VX_Material.cpp:542-549
//   voxel constants       src/old/VX_MaterialVoxel.cpp:57-79    link blend + beam  src/old/VX_MaterialLink.cpp:45-141
#pragma once
#include <vector>

#include "../../../include/vx3_abi.h"

namespace vx3 {

// Piece-wise linear, monotone stress-strain law through the origin.  eps[0] = sig[0] = 0 always; a law with one segment
// is "linear" (Hooke up to an optional failure point).  -1 in a yield / failure field means "none".
struct StressLaw {
    std::vector<float> eps, sig;
    bool linear = true;
    float E = 0;                              // slope of the first segment
    float yield_sig = -1, yield_eps = -1;
    float fail_sig = -1, fail_eps = -1;

    // MatModel 0 / 1 (src/VXA/VX_Object.cpp:1395): Hooke with modulus E; fail_stress < 0 = never fails
    static bool hooke(float E, float fail_stress, StressLaw *out);
    // MatModel 2: elastic slope E to the yield stress, plastic slope Ep beyond it
    static bool bilinear(float E, float Ep, float yield_stress, float fail_stress, StressLaw *out);
    // MatModel 3: n tabulated points (a leading (0,0) is optional); yield point by the 0.2 % offset rule
    static bool tabulated(int n, const float *strain, const float *stress, StressLaw *out);
    // two half-links in series (springs in series segment by segment); fail_stress = the weaker material's
    static StressLaw in_series(const StressLaw &a, const StressLaw &b, float fail_stress);

    float tangent(float strain) const;       // local slope d(sigma)/d(epsilon); 0 beyond failure
    float strain_for(float stress) const;    // inverse of the law
};

// What the builder knows about one palette material before anything is derived.
struct MaterialInput {
    StressLaw law;
    float nu = 0, rho = 1, cte = 0, mu_static = 0, mu_kinetic = 0;
    float zeta_internal = 1, zeta_global = 0, zeta_collision = 0;
    float grav_mult = 0;
    double nom_size = 0.001;
    int r = -1, g = -1, b = -1, a = -1;
    // identity / VX3 additions, passed through
    int matid = 0, fixed = 0, sticky = 0, is_target = 0, is_measured = 1, is_pacemaker = 0, is_electrical_active = 0;
    double cilia = 0, pacemaker_period = 0, signal_value_decay = 0.9, signal_time_delay = 0.0, inactive_period = 0.05;
    double remove_after_s = 0, thermal_on_after_s = 0, cilia_on_after_s = 0;
};

float clamp_poisson(float nu);   // [0, 0.5)
float clamp_density(float rho);  // > 0
int clamp_colour(int c);         // 0..255

// Fills `out` (all derived members included).  The law's points are appended to `pool` (two vectors per material), which
// must outlive `out`: out.strain_data / out.stress_data point into it.
void voxel_constants(const MaterialInput &m, vx3_voxel_material *out, std::vector<std::vector<float>> *pool);
// The blended material of a link between voxels of materials a and b, and its beam constants.
void link_constants(const vx3_voxel_material &a, const vx3_voxel_material &b, int index_a, int index_b, vx3_link_material *out,
                    std::vector<std::vector<float>> *pool);

} // namespace vx3
