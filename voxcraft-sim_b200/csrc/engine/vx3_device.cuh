// Device-side data model of a batch: structure-of-arrays with integer indices, all simulations of the
// batch concatenated (voxel g = voff[sim] + local index; link slot g = loff[sim] + local index).
//
// Replaces the reference's array-of-fat-structs with raw device pointers
// (VX3_Voxel src/VX3/VX3_Voxel.h:268-314, VX3_Link src/VX3/VX3_Link.h:209-245, materials
// src/VX3/VX3_Material.h:107-166 / VX3_MaterialVoxel.h:50-59 / VX3_MaterialLink.h:24-37).
#pragma once
#include <cstdint>

#include "../../../include/vx3_abi.h"
#include "vx3_math.cuh"

namespace vx3 {

// device-only voxel flag bits (above the reference's boolStates bits, which keep their values)
#define VXF_ENABLE_ATTACH (1 << 8) // VX3_Voxel::enableAttach (VX3_Voxel.h:302)
#define VXF_REMOVED (1 << 9)       // VX3_Voxel::removed
#define VXF_BOOLSTATE_MASK 0xFF

// link state word = vx3_state_view.link_flags bits + axis + transient bit
#define LKS_VALID VX3_LINKSTATE_LOCAL_VELOCITY_VALID
#define LKS_SMALL VX3_LINKSTATE_SMALL_ANGLE
#define LKS_DETACHED VX3_LINKSTATE_DETACHED
#define LKS_REMOVED VX3_LINKSTATE_REMOVED
#define LKS_AXIS_SHIFT 4
#define LKS_AXIS_MASK (3 << LKS_AXIS_SHIFT)
#define LKS_JUST_CREATED (1 << 6) // attached during the current step (cleared by the next link pass)
#define LKS_FAILED (1 << 7)       // on the simulation's failed-link list of this step (EnableDetach): detached by k_resolve_detach
#define LKS_NEWLINK_SHIFT VX3_LINKSTATE_NEWLINK_SHIFT
#define LKS_PUBLIC_MASK (~(LKS_AXIS_MASK | LKS_JUST_CREATED | LKS_FAILED))

#define VX3_DEV_MAX_TOKENS 128 // per-voxel programs (force field, attach conditions): up to this many tokens run in the small evaluator frame
#define VX3_MAX_PARTNERS 96    // contact partners of one voxel inside the collision envelope
#define VX3_CELL_SLOTS 8       // voxels a grid bucket holds inline (one 32-byte sector)

// Voxel material constants the step reads (all derived values precomputed in the reference's arithmetic).
struct VoxMatC {
    double nomSize;
    double size[3];        // mat->size() = nomSize*extScale
    double cilia;
    double thermal_on_after, cilia_on_after, remove_after;
    float alphaCTE, muStatic, muKinetic, massInverse;
    float mass, momentInertiaInverse, globalDampT, globalDampR; // zetaGlobal*_2xSqMxExS, zetaGlobal*_2xSqIxExSxSxS
    float colDampT, penStiff, gravityForce, dampMultNum;        // zetaCollision*_2xSqMxExS, (float)(2*E*nomSize), -mass*9.80665f*gravMult, 2*_sqrtMass*zetaInternal
    float E;
    int32_t fixed, sticky, is_target, is_measured, matid;
    int32_t self_lmat; // link material of a (this,this) pair: used by attach (global link-material index), -1 if none
    float nu;          // poissonsRatio()
};

// What the link pass needs of an end voxel's material (one record per VoxMatC entry, same index)
struct VoxMatL {
    double size[3];
    double thermal_on_after;
    float alphaCTE, dampMultNum;
    int32_t fixed, _pad;
};

// Link material constants (VX3_MaterialLink + the VX3_Material stress model).
struct LinkMatC {
    float E, nu, eHat, epsilonFail;
    float a1, a2, b1, b2, b3;
    float sqA1, sqA2xIp, sqB1, sqB2xFMp, sqB3xIp;
    int32_t linear;
    int32_t data_off, n_data; // into the strain/stress pool (device layout: duplicated leading 0)
    int32_t _pad;
};

// Signal constants of a voxel material (VX3_Material.h:151-161), same index as VoxMatC
struct SigMatC {
    double pacemaker_period, value_decay, time_delay, inactive_period;
    int32_t is_pacemaker, _pad;
};

struct ExtC { // VX3_External
    int32_t dof;
    float force[3], moment[3];
    double translation[3];
    double rotq[4];
};

// per-simulation constants
struct SimC {
    int32_t voff, nvox, loff, lcap, nhostlinks;
    int32_t vary_temp, enable_expansion;
    int32_t enable_collision, enable_attach, enable_detach, enable_cilia;
    int32_t safety_guard;
    int32_t has_ff, has_attach_cond; // any force-field / attach-condition program present
    int32_t prog_off[VX3_PROG_COUNT], prog_n[VX3_PROG_COUNT];
    int32_t tgt_off, ntgt;
    int32_t chunk_off, nchunks; // CoM reduction chunks
    int32_t cand_off, cand_cap; // this simulation's region of the attach-candidate array (cand_cap is a power of two; 0 = cannot attach)
    int32_t fail_off, fail_cap; // this simulation's region of the failed-link list (EnableDetach)
    int32_t secondary_experiment, enable_signals;
    double reinit_after; // ReinitializeInitialPositionAfterThisManySeconds
    double temp_amp, temp_period;
    double vox_size, pair_radius; // MaxDistInVoxelLengthsToCountAsPair * voxSize (0 = closeness off)
    double cell_inv;              // 1 / collision grid cell edge
    double dt_frac, optimal_dt;
    int32_t dt_from_state, _pad_dt; // a model link's material has nu != 0: recommendedTimeStep() depends on the rest lengths at the first
                                    // step's temperatures (k_optimal_dt), not only on the model
};

// bits of SimD::hot_flags (constants of the simulation mirrored next to the per-step scalars)
#define SHF_THERMAL (1 << 0)     // VaryTempEnabled && TempPeriod > 0
#define SHF_EXPANSION (1 << 1)   // EnableExpansion
#define SHF_CILIA (1 << 2)
#define SHF_FORCE_FIELD (1 << 3)
#define SHF_ATTACH_COND (1 << 4)
#define SHF_SIGNALS (1 << 5)
#define SHF_DETACH (1 << 6)      // EnableDetach: the link pass lists the links whose failure strain is passed
#define SHF_SECONDARY (1 << 7)   // SecondaryExperiment
#define SHF_STOP_PROG (1 << 8)   // a stop-condition program is present

// per-simulation dynamic scalars (VX3_VoxelyzeKernel members that change during the run).  The first 48 bytes are the
// "hot" block every link / voxel of the simulation needs each step; the streaming kernels prefetch it with three
// 16-byte cp.async, so its layout and the 16-byte alignment matter.
struct alignas(16) SimD {
    double t;          // currentTime
    int32_t status;    // vx3_status
    int32_t diverged;  // set by the link pass of the current step
    float dt;          // step in use (float, VX3_VoxelyzeKernel.cu:237,253)
    int32_t hot_flags; // SHF_* (constant)
    double temp_amp;   // TempAmplitude (constant)
    double temp_period; // TempPeriod (constant)
    int32_t topo_epoch; // bumped whenever a link of the simulation is created, detached or removed (validates Dev::nbcache entries); starts at 1
    int32_t _hot_pad;
    long long steps;   // CurStepCount
    int32_t link_cnt;  // d_v_links.size()
    int32_t collision_count;
    int32_t nsurface;
    int32_t angle_samples;
    int32_t num_close_pairs;
    int32_t err;
    int32_t attach_events, detach_events;
    double com[3], com_hist[2][3], com0[3];
    double recent_angle, target_closeness, fitness;
    double total_dist;
    int32_t n_measured;
    int32_t initpos_reinitialized; // InitialPositionReinitialized
    int32_t cand_count;            // attach candidates appended by the contact phase of this step
    int32_t fail_count;            // failed links appended by the link pass of this step
    int32_t cand_peak, fail_peak;  // largest counts a step has produced (diagnostics)
};

struct Chunk { int32_t sim, vstart, vcount, _pad; };

// Everything the contact phase reads of a voxel — as itself or as somebody's candidate partner — in one aligned 64-byte
// record, so that a candidate costs one memory round trip instead of a chain (cell -> material index -> material -> pose).
struct alignas(64) ContactRec {
    int32_t cx, cy, cz, bucket; // grid cell and hash bucket; bucket < 0: not in the grid (interior / removed / not colliding)
    double px, py, pz;          // position
    double bs;                  // baseSizeAverage() at this step's temperature (VX3_Voxel.h:101-104)
    int32_t sim, mat;           // simulation, global voxel-material index
    int32_t fixed, _pad;
};

// One slot of a grid bucket: the voxel and its position rounded to float — enough for a conservative first cut of the envelope
// test without touching the voxel's 64-byte ContactRec (a bucket's 8 slots are one 128-byte line).
struct alignas(16) CellItem {
    float x, y, z;
    int32_t v;
};

struct Cand { // attach candidate (VX3_VoxelyzeKernel.cu:729-812), sorted by (hi, lo) before resolution
    unsigned long long key; // hi<<32 | lo (global voxel indices)
    int32_t info;           // dir1 | dir2<<3 | axis<<6 | reverse<<8
    int32_t _pad;
};

// Blocked SoA ("AoSoA") index maps of the three hot record arrays: the planes of 32 consecutive items are contiguous, so
// a warp's access to one plane is one coalesced 512-byte run AND a tile's whole record is one contiguous block
// (2.5 KB of link history, 3 KB of link end forces, 1.5 KB of momenta per 32 items).  Plain plane-major SoA coalesces just as
// well but scatters every tile over as many DRAM pages / TLB entries as there are planes (measured: profiles/).
VXHD size_t idx_lh(int p, size_t g) { return ((g >> 5) * 5 + p) * 32 + (g & 31); }
VXHD size_t idx_mo(int p, size_t v) { return ((v >> 5) * 3 + p) * 32 + (v & 31); }
VXHD size_t idx_lf(int k, size_t g) { return ((g >> 5) * 6 + k) * 32 + (g & 31); }

// All device arrays of a batch (plain pointers; owned by the host Batch object).
// Receive side of a slab's halo exchange as the link pass sees it (vx3_halo.cuh; device-resident, Dev::hin): the link kernel of a slab
// batch reads its ghost voxels' poses straight from the receive buffers.
struct HaloIn {
    int face_tile0;              // link tiles [0, face_tile0) hold no link with a ghost end
    int n[2];                    // records expected from the lower / upper neighbour
    const int32_t *idx[2];       // my ghost voxels, in the neighbour's send order
    const double *buf[2];        // [2 parities][n][8] doubles, written by the neighbour
    const unsigned int *flag[2]; // [2 parities] send numbers published by the neighbour (32 words apart)
    unsigned int *seq;           // {my sends so far, the send I have collected, ...} (Halo::seq)
    unsigned int *state;         // [0] the send number some warp has undertaken to wait for, [32] the send number known to be in (halo_arrival_wait)
    const int32_t *ghost_row;    // [nvox] -1: an owned voxel; else side << 30 | position in that side's receive buffer
    int *err;
    long long spin_cycles;
};

// Send side, as the voxel pass sees it (Dev::hout): a face voxel's new pose record goes into its own row AND straight into the
// neighbour's receive buffer (peer memory); the last CTA of the pass to finish publishes the send number.  The voxel's position
// in the neighbour's buffer is vc4[v].w - 1 = side << 30 | position (0: not a face voxel).
struct HaloOut {
    double *buf[2];          // the neighbours' receive buffers for my side: [2 parities][n][8] doubles (peer memory)
    unsigned int *flag[2];   // the neighbours' [2 parities] send-number words
    int n[2];
    unsigned int *seq;       // {my sends so far, ...} (Halo::seq)
    unsigned int *count;     // arrival counter of the pass's CTAs
};

struct Dev {
    int32_t nsims, nvox, nlinkslots, nchunks;
    const HaloIn *hin;   // slab batches whose link pass reads the receive buffers itself, else NULL
    const HaloOut *hout; // slab batches whose voxel pass sends, else NULL
    unsigned int *vox_count; // arrival counter of the voxel pass's CTAs (the last one does the end-of-step bookkeeping, k_voxels)
    const SimC *simc;
    SimD *simd;
    const VoxMatC *vmat_tab;
    const VoxMatL *vmatl_tab;
    const LinkMatC *lmat_tab;
    int32_t n_vmats, n_lmats;
    const float *strain_pool, *stress_pool;
    const vx3_token *tokens;
    const ExtC *exts;
    const Chunk *chunks;
    const int32_t *targets;
    // voxels
    int32_t vstride, lstride; // plane strides of the voxel / link-slot SoA planes (multiples of 32 elements)
    double *pose;     // [nvox][8]: pos xyz, orient wxyz, {float temperature of the coming step, float previousDt}
    double2 *mom2;    // [vstride/32][3][32] (idx_mo): {linMom.x, linMom.y}, {linMom.z, angMom.x}, {angMom.y, angMom.z}
    int32_t *vflags;  // boolStates | VXF_*
    const int32_t *vmat; // global voxel-material index
    const int32_t *vsim;
    const int4 *vc4;     // {vmat, vsim, vext, halo send position + 1 or 0}: the voxel's constant indices in one 16-byte record (streaming voxel pass)
    const double *phase;
    float *tempe;     // temperature the last executed step used (VX3_Voxel::temp)
    int32_t *vlinks;  // [nvox][6] global link slot or -1
    const int32_t *vext;
    const int16_t *ixyz; // [nvox][3]
    double *contact;  // [nvox][3] (NULL if no sim collides)
    const double *base_cilia, *shift_cilia; // [nvox][3] or NULL
    double *initpos;  // [nvox][3]
    // signals (NULL unless a simulation has EnableSignals): [nvox][6] = localSignal, localSignaldt, inactiveUntil,
    // packmakerNextPulse, d_signal.value, d_signal.activeTime (VX3_Voxel.h:304-309); sprop[nvox] = the value a voxel
    // propagates to its neighbours in the current step (0 = none), see k_signals
    double *sig;
    double *sprop;
    const SigMatC *smat_tab;
    // links
    int2 *lends;      // (vneg, vpos) global voxel indices; x<0 = empty pool slot
    int32_t *lstate;
    int32_t *lmat;
    int4 *lc4;        // {vneg, vpos, lmat, sim}: the link's constant indices in one 16-byte record (streaming link pass)
    double2 *lh2;     // [lstride/32][5][32] (idx_lh): {pos2.x, pos2.y}, {pos2.z, a1v.x}, {a1v.y, a1v.z}, {a2v.x, a2v.y}, {a2v.z, currentRestLength}
    float4 *lstrain;  // strain, maxStrain, strainOffset, _stress
    float2 *larea;    // currentTransverseArea, currentTransverseStrainSum
    // link end forces [lstride/32][6][32] (idx_lf): {Fneg.x, Fneg.y}, {Fneg.z, Mneg.x}, {Mneg.y, Mneg.z}, then the same three for
    // the positive end.  Written coalesced by the link pass, gathered by the voxel pass through vlinks.  (Storing them by
    // receiving voxel and direction instead makes the voxel pass stream, but a direction without a link leaves a 16-byte hole
    // in its sector, and every partially written sector costs an ECC read-modify-write at eviction: measured 2.5x slower.)
    double2 *lf2;
    // collision grid (hashed uniform grid, per-bucket lists)
    int32_t hmask;
    // bucket b: cell_cnt[b] voxels; the first VX3_CELL_SLOTS of them inline in cell_items[b][] (voxel + float position), the rest (rare) chained
    // through cell_ovf[b] (voxel + 1, 0 = none) / cell_next[].  cell_cnt and cell_ovf are one allocation, zeroed every step
    int32_t *cell_cnt, *cell_ovf, *cell_next;
    CellItem *cell_items;
    struct ContactRec *crec; // [nvox] what the contact phase needs of a voxel, in one 64-byte record (written by k_grid_build)
    int32_t *uf;      // [nvox] union-find parents over the voxels (NULL unless a simulation can attach), see uf_find
    float4 *pcache;   // [nvox] {poissonsStrain() x, y, z, validity stamp as int bits} (VX3_Voxel.h:287-288), for the transverse info of a link
                      // the attach phase creates; NULL unless an attaching simulation has a material with nu != 0.  Stamp: -2 invalid,
                      // -1 valid since import (every strain was zero), s >= 0 computed during step s (resolve_link_transverse)
    int2 *nbcache;    // [nvox][8] {partner, topo_epoch << 1 | answer}: what within_five_links said about (voxel, partner) while the link graph
                      // was at that epoch — a settled pile asks the same questions every step (allocated with vnb)
    int32_t *vnb;     // [nvox][8] the voxel at the far end of each of the six link slots (-1: none; two pad entries) — the link graph as
                      // an adjacency table for within_five_links (allocated with uf; kept current by attach / detach / removal)
    Cand *cands;        // per-simulation regions (SimC::cand_off / cand_cap), counts in SimD::cand_count
    int32_t *fail_list; // per-simulation regions (SimC::fail_off / fail_cap) of link slots, counts in SimD::fail_count
    // CoM partials [nchunks][6]: sum m*x, m*y, m*z, m, sum dist, n_measured
    double *com_part;
#ifdef __CUDACC__
    __device__ __forceinline__ double2 *lh(int p, int g) const { return lh2 + idx_lh(p, g); }
    __device__ __forceinline__ double2 *mo(int p, int v) const { return mom2 + idx_mo(p, v); }
    __device__ __forceinline__ double2 *lf(int k, int g) const { return lf2 + idx_lf(k, g); }
    __device__ __forceinline__ double &lrest(int g) const { return lh2[idx_lh(4, g)].y; }
#endif
};

} // namespace vx3
