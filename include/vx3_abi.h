/*
 * vx3_abi.h — C ABI of the B200-native voxel physics engine.
 *
 * This is the drop-in boundary for the reference's VX3_VoxelyzeKernel step loop
 * (seam (iii) of SURVEY.md §8(b)).  Every entry point cites the reference
 * interface it replaces (paths relative to the voxcraft-sim source tree).
 *
 * Plain C: pointers, sizes and PODs only.  No torch / CUDA types in any signature.
 * All model arrays are HOST pointers owned by the caller; the library copies what
 * it needs during vx3_batch_create() and never touches them afterwards.
 */
#ifndef VX3_ABI_H
#define VX3_ABI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VX3_ABI_VERSION 1

/* ------------------------------------------------------------------ enums -- */

/* src/Utils/VX3_MathTree.h:4-29 (VX3_MathTreeOperator) — same numeric order. */
enum vx3_op {
    VX3_OP_END = 0, VX3_OP_CONST, VX3_OP_E, VX3_OP_PI, VX3_OP_VAR, VX3_OP_ADD, VX3_OP_SUB,
    VX3_OP_MUL, VX3_OP_DIV, VX3_OP_POW, VX3_OP_SQRT, VX3_OP_SIN, VX3_OP_COS, VX3_OP_TAN,
    VX3_OP_ATAN, VX3_OP_LOG, VX3_OP_INT, VX3_OP_ABS, VX3_OP_NOT, VX3_OP_GREATERTHAN,
    VX3_OP_LESSTHAN, VX3_OP_AND, VX3_OP_OR, VX3_OP_NORMALCDF
};

/* src/old/types.h:33-39 (voxFlags).  NB: the bit named SURFACE means *interior*. */
#define VX3_VOX_SURFACE               (1 << 1)
#define VX3_VOX_FLOOR_ENABLED         (1 << 2)
#define VX3_VOX_FLOOR_STATIC_FRICTION (1 << 3)
#define VX3_VOX_COLLISIONS_ENABLED    (1 << 5)
/* not a reference bit: halo copy of a voxel another rank owns (slab decomposition of one body, vx3_batch_halo_*):
 * never integrated, not counted in the centre of mass, its pose record is overwritten by the halo exchange */
#define VX3_VOX_GHOST                 (1 << 7)

/* src/old/types.h:6-8 (linkFlags) */
#define VX3_LINK_LOCAL_VELOCITY_VALID (1 << 0)

/* src/old/types.h:47-55 (dofComponent) */
#define VX3_DOF_X_TRANSLATE (1 << 0)
#define VX3_DOF_Y_TRANSLATE (1 << 1)
#define VX3_DOF_Z_TRANSLATE (1 << 2)
#define VX3_DOF_X_ROTATE    (1 << 3)
#define VX3_DOF_Y_ROTATE    (1 << 4)
#define VX3_DOF_Z_ROTATE    (1 << 5)

/* linkDirection (src/old/types.h:13-20): X_POS=0 X_NEG=1 Y_POS=2 Y_NEG=3 Z_POS=4 Z_NEG=5
 * linkAxis      (src/old/types.h:41-45): X=0 Y=1 Z=2 */

#define VX3_MAX_TOKENS 1024 /* src/VX3/VX3_VoxelyzeKernel.cuh:113-119 */

/* program slots (the ten token programs poked by readVXD,
 * src/VX3/VX3_SimulationManager.cu:326-357) */
enum vx3_program_slot {
    VX3_PROG_STOP = 0,   /* StopConditionFormula */
    VX3_PROG_FITNESS,    /* FitnessFunction      */
    VX3_PROG_FORCE_X, VX3_PROG_FORCE_Y, VX3_PROG_FORCE_Z, /* ForceField */
    VX3_PROG_ATTACH_0, VX3_PROG_ATTACH_1, VX3_PROG_ATTACH_2, VX3_PROG_ATTACH_3,
    VX3_PROG_ATTACH_4,   /* AttachCondition.Condition_0..4 */
    VX3_PROG_COUNT
};

/* per-simulation status after run/step */
enum vx3_status {
    VX3_SIM_RUNNING = 0,  /* still stepping                                        */
    VX3_SIM_STOPPED = 1,  /* StopConditionMet (VX3_SimulationManager.cu:63)         */
    VX3_SIM_DIVERGED = 2, /* doTimeStep returned false (:65-69)                     */
    VX3_SIM_STEP_CAP = 3  /* hit the 1,000,000-step cap (:62)                       */
};

/* error codes (return values; 0 = ok) */
#define VX3_OK 0
#define VX3_ERR_INVALID -1   /* bad argument / malformed model    */
#define VX3_ERR_CUDA -2      /* CUDA runtime error (see vx3_last_error) */
#define VX3_ERR_NO_DEVICE -3 /* no usable sm_100 device — there is NO CPU fallback */
#define VX3_ERR_CAPACITY -4  /* link pool exhausted during attach */

/* ------------------------------------------------------------------ PODs --- */

/* One token of a compiled math tree.  src/Utils/VX3_MathTree.h:30-37 */
typedef struct vx3_token {
    int32_t op; /* enum vx3_op */
    int32_t _pad;
    double value;
} vx3_token;

/* A token program in the reference's reverse-BFS order, terminated by VX3_OP_END
 * (ParseMathTree, src/VX3/VX3_SimulationManager.cu:157-275).  n == 0 means "tag
 * absent": the reference then evaluates an all-mtEND buffer and reads an
 * uninitialised value (VX3_MathTree.h:57-58); this ABI defines the result instead:
 * stop = never, fitness = 0, attach condition = true.  (Force-field slots default
 * to the constant 0, VX3_ForceField.h:10-20.) */
typedef struct vx3_program {
    int32_t n;
    int32_t _pad;
    const vx3_token *tok;
} vx3_program;

/* Voxel material: the fields of VX3_Material + VX3_MaterialVoxel the step loop reads.
 * src/VX3/VX3_Material.h:107-166, src/VX3/VX3_MaterialVoxel.h:50-59.  All derived
 * ("_x") members are supplied by the host model builder exactly as the reference
 * host copies them from CVX_MaterialVoxel (VX3_MaterialVoxel.cu:4-11). */
typedef struct vx3_voxel_material {
    int32_t matid;
    int32_t fixed, sticky, is_target, is_measured, linear;
    int32_t is_pacemaker, is_electrical_active;
    int32_t r, g, b, a;
    float E, sigmaYield, sigmaFail, epsilonYield, epsilonFail;
    float nu, rho, alphaCTE, muStatic, muKinetic;
    float zetaInternal, zetaGlobal, zetaCollision;
    float eHat;
    /* VX3_MaterialVoxel */
    float gravMult, mass, massInverse, sqrtMass, firstMoment, momentInertia, momentInertiaInverse;
    float _2xSqMxExS, _2xSqIxExSxSxS;
    int32_t n_data;           /* host strainData.size() (first point is 0,0)      */
    const float *strain_data; /* host arrays; the engine prepends the duplicated  */
    const float *stress_data; /* leading 0 exactly like syncVectors (VX3_Material.cu:463-477) */
    double nomSize;
    double extScale[3];
    double cilia;
    double pacemaker_period, signal_value_decay, signal_time_delay, inactive_period;
    double remove_after_s, thermal_on_after_s, cilia_on_after_s;
} vx3_voxel_material;

/* Link material: VX3_MaterialLink (src/VX3/VX3_MaterialLink.h:24-37) — a voxel material
 * (blended, VX3_MaterialLink.cu:53-127) plus the beam constants (:129-149). */
typedef struct vx3_link_material {
    vx3_voxel_material m;
    int32_t vox1_mat, vox2_mat; /* indices into voxel_mats */
    float a1, a2, b1, b2, b3;
    float sqA1, sqA2xIp, sqB1, sqB2xFMp, sqB3xIp;
} vx3_link_material;

/* VX3_External (src/VX3/VX3_External.h:69-73) */
typedef struct vx3_external {
    int32_t dof_fixed; /* dofObject bits */
    float force[3], moment[3];
    double translation[3];
    double rotation[3];
    double rotation_q[4]; /* w,x,y,z: cached _extRotationQ */
} vx3_external;

/* Scalar options of one simulation: the VX3_VoxelyzeKernel fields set by the ctor
 * (src/VX3/VX3_VoxelyzeKernel.cu:93-101) and by readVXD
 * (src/VX3/VX3_SimulationManager.cu:321-376). */
typedef struct vx3_sim_options {
    double vox_size;
    double dt_frac;
    int32_t temp_enabled, vary_temp_enabled;
    double temp_base, temp_amplitude, temp_period;
    int32_t enable_collision; /* default 1 (true) */
    int32_t enable_attach, enable_detach;
    double watch_distance, bounding_radius;
    int32_t safety_guard; /* default 500 */
    int32_t record_step_size, record_link, record_voxel;
    int32_t save_position_of_all_voxels;
    double max_dist_in_voxel_lengths_to_count_as_pair;
    int32_t enable_cilia, enable_signals;
    int32_t secondary_experiment;
    double reinit_initial_position_after_s;
    int32_t enable_expansion;
    int32_t _pad;
} vx3_sim_options;

/* One simulation's host model, flat structure-of-arrays with integer indices.
 * Replaces the deep copy of the CVX_Sim object graph done by
 * VX3_VoxelyzeKernel::VX3_VoxelyzeKernel(CVX_Sim*) (src/VX3/VX3_VoxelyzeKernel.cu:27-105).
 * Voxel order = In->Vx.voxelsList, link order = In->Vx.linksList.
 * Arrays marked [opt] may be NULL (then the stated default is used). */
typedef struct vx3_model_desc {
    char name[256]; /* vxa_filename */

    int32_t n_voxel_mats;
    int32_t n_link_mats;
    const vx3_voxel_material *voxel_mats;
    const vx3_link_material *link_mats;

    int32_t n_voxels;
    int32_t n_links;
    int32_t n_externals;
    int32_t link_capacity; /* pool size for attach-created links; 0 = library picks */

    /* voxels (VX3_Voxel::VX3_Voxel(CVX_Voxel*, ...), src/VX3/VX3_Voxel.cu:10-50) */
    const int16_t *ix, *iy, *iz;
    const int32_t *vox_mat;       /* index into voxel_mats                         */
    const double *pos;            /* [n_voxels][3]                                 */
    const double *orient;         /* [n_voxels][4] w,x,y,z  [opt: identity]        */
    const double *lin_mom;        /* [n_voxels][3]          [opt: 0]               */
    const double *ang_mom;        /* [n_voxels][3]          [opt: 0]               */
    const int32_t *vox_flags;     /* boolStates                                    */
    const float *temp;            /* tempe                  [opt: 0]               */
    const double *phase_offset;   /*                        [opt: 0]               */
    const int32_t *vox_links;     /* [n_voxels][6] link index per linkDirection, -1 = none */
    const int32_t *vox_ext;       /* index into externals, -1 = none [opt: none]   */
    const double *base_cilia;     /* [n_voxels][3]          [opt: 0]               */
    const double *shift_cilia;    /* [n_voxels][3]          [opt: 0]               */
    const vx3_external *externals;

    /* links (VX3_Link::VX3_Link(CVX_Link*, ...), src/VX3/VX3_Link.cu:6-30) */
    const int32_t *link_vneg, *link_vpos; /* voxel indices                         */
    const int32_t *link_axis;             /* linkAxis                              */
    const int32_t *link_mat;              /* index into link_mats                  */
    /* link state [opt: CVX_Link::reset() values, src/old/VX_Link.cpp:56-70] */
    const double *link_pos2, *link_angle1v, *link_angle2v; /* [n_links][3]         */
    const float *link_strain, *link_max_strain, *link_strain_offset, *link_stress;
    const int32_t *link_flags;       /* boolStates                                 */
    const int32_t *link_small_angle; /* 0/1                                        */
    const double *link_rest_length;
    const float *link_transverse_area, *link_transverse_strain_sum;
    const float *link_strain_ratio;

    vx3_sim_options opt;
    vx3_program prog[VX3_PROG_COUNT];
} vx3_model_desc;

/* Result of one simulation: the fields collectResults reads back
 * (src/VX3/VX3_SimulationManager.cu:428-470) + status. */
typedef struct vx3_result {
    char name[256];
    int32_t status; /* enum vx3_status */
    int32_t num_voxel;
    int32_t num_measured_voxel;
    int32_t num_close_pairs;
    int64_t steps;  /* CurStepCount */
    int32_t num_links; /* d_v_links.size() incl. attach-created */
    int32_t collision_count;
    double current_time;
    double fitness_score; /* NaN when diverged: keeps the report's "NaN sorts last" rule */
    double vox_size;
    double initial_com[3];
    double current_com[3];
    double total_distance_of_all_voxels;
    double recent_angle;
    double target_closeness;
    double dt; /* the float step actually used, widened */
} vx3_result;

/* Host-side SoA dump of one simulation's live state (parity hook).  The caller
 * provides the buffers (NULL = skip that field); capacities are in elements. */
typedef struct vx3_state_view {
    int32_t n_voxels;  /* in: capacity, out: count */
    int32_t n_links;   /* in: capacity, out: live link count (incl. attached) */
    double *pos;       /* [n_voxels][3] */
    double *orient;    /* [n_voxels][4] */
    double *lin_mom;   /* [n_voxels][3] */
    double *ang_mom;   /* [n_voxels][3] */
    int32_t *vox_flags;
    float *temp;
    int32_t *vox_links; /* [n_voxels][6] */
    double *contact_force; /* [n_voxels][3]: pending contact force (debug)            */
    int32_t *link_vneg, *link_vpos, *link_axis, *link_mat;
    double *link_pos2, *link_angle1v, *link_angle2v; /* [n_links][3] */
    double *link_force_neg, *link_force_pos, *link_moment_neg, *link_moment_pos; /* [n_links][3] */
    float *link_strain, *link_max_strain, *link_strain_offset, *link_stress;
    int32_t *link_flags;    /* bit0 LOCAL_VELOCITY_VALID, bit1 smallAngle, bit2 isDetached,
                               bit3 removed, bits 8.. isNewLink countdown */
    double *link_rest_length;
    double *signal;         /* [n_voxels][6]: localSignal, localSignaldt, inactiveUntil, packmakerNextPulse,
                               d_signal.value, d_signal.activeTime (VX3_Voxel.h:304-309); zeros when EnableSignals=0 */
} vx3_state_view;

#define VX3_LINKSTATE_LOCAL_VELOCITY_VALID (1 << 0)
#define VX3_LINKSTATE_SMALL_ANGLE (1 << 1)
#define VX3_LINKSTATE_DETACHED (1 << 2)
#define VX3_LINKSTATE_REMOVED (1 << 3)
#define VX3_LINKSTATE_NEWLINK_SHIFT 8

typedef struct vx3_run_opts {
    int64_t max_steps;      /* 0 = the reference cap, 1,000,000 (VX3_SimulationManager.cu:62) */
    int32_t steps_per_launch; /* 0 = library picks; granularity of host visibility      */
    int32_t emit_history;   /* honour RecordStepSize and stream frames to the callback  */
} vx3_run_opts;

/* History sink: receives, byte for byte, what the reference's CUDA_Simulation kernel writes with device printf
 * (src/VX3/VX3_SimulationManager.cu:11-121): the "Simulation %d runs" line (:25), the {{{setting}}} header when
 * RecordStepSize > 0 (:40-50), "real_stepsize: ..." (:56-58; preceded / followed by recommendedTimeStep's "WARNING: No links."
 * for a model without links), the voxel / link frames (:70-114), "Diverged" (:65-69) and the "ends" line (:118-119).  The
 * device index in those lines is the batch's device, the simulation index the model's position in the batch.  Called on the
 * caller's thread. */
typedef void (*vx3_history_cb)(void *user, int sim, const char *bytes, size_t n);

typedef struct vx3_batch vx3_batch; /* opaque: one batch = one device = one stream */

/* -------------------------------------------------------------- entry points */

/* Build the device-resident batch from n host models on CUDA device `device`.
 * Replaces VX3_SimulationManager::readVXD's per-file "VX3_VoxelyzeKernel h_d_tmp(&MainSim)"
 * + cudaMemcpy (src/VX3/VX3_SimulationManager.cu:277-381) and the device-side init
 * at the top of CUDA_Simulation (syncVectors, saveInitialPosition, registerTargets,
 * updateCurrentCenterOfMass, InitializeCenterOfMass; :20-24,54-55). */
int vx3_batch_create(int device, const vx3_model_desc *models, int n, vx3_batch **out);

/* Run every simulation to its stop condition, divergence or the step cap.
 * Replaces startKernel + the CUDA_Simulation loop
 * (src/VX3/VX3_SimulationManager.cu:409-426, :62-115) and the closing
 * updateCurrentCenterOfMass + computeFitness (:116-117). */
int vx3_batch_run(vx3_batch *b, const vx3_run_opts *opts, vx3_history_cb cb, void *user);

/* Advance every still-running simulation by exactly k calls of doTimeStep
 * (src/VX3/VX3_VoxelyzeKernel.cu:237-359) without evaluating the stop condition.
 * Parity / benchmark hook. */
int vx3_batch_step(vx3_batch *b, int64_t k);

/* Same as vx3_batch_step but with an explicit dt (float, as doTimeStep(float dt));
 * dt < 0 selects DtFrac * recommendedTimeStep() like the reference default. */
int vx3_batch_step_dt(vx3_batch *b, int64_t k, float dt);

/* Block until all queued work of the batch has finished. */
int vx3_batch_sync(vx3_batch *b);

/* Copy one simulation's live state to host buffers (parity hook; no reference twin —
 * the reference reads VX3_Voxel structs back wholesale, VX3_SimulationManager.cu:448). */
int vx3_batch_state(vx3_batch *b, int sim, vx3_state_view *view);

/* Refresh center of mass + fitness and fill out[0..n).  Replaces collectResults
 * (src/VX3/VX3_SimulationManager.cu:428-470).  Order = model order (unsorted). */
int vx3_batch_results(vx3_batch *b, vx3_result *out);

/* Optional per-voxel outputs of collectResults (SavePositionOfAllVoxels):
 * init_pos / pos are [n_voxels][3], mats is matid per voxel.  NULL = skip.
 * sim = -1: all simulations of the batch, concatenated in model order (buffers sized for the batch's total voxel count). */
int vx3_batch_positions(vx3_batch *b, int sim, double *init_pos, double *pos, int32_t *mats);

/* recommendedTimeStep() of simulation `sim` (src/VX3/VX3_VoxelyzeKernel.cu:184-217). */
int vx3_batch_recommended_dt(vx3_batch *b, int sim, double *out);

/* Device-time of the last vx3_batch_step/_run call in milliseconds, measured with CUDA
 * events on the batch's own stream (bench hook), and the number of kernel launches
 * it issued. */
int vx3_batch_last_timing(vx3_batch *b, double *ms, int64_t *launches);

/* Bench hooks.  set_profiling(on): bracket every kernel launch of the following step/run calls with CUDA events
 * on the batch's stream and accumulate the device time per kernel; use_persistent = 0 forces the streaming
 * kernels even where the on-chip persistent kernel applies.  kernel_stats(index): name, accumulated
 * milliseconds and launch count of kernel `index` (0, 1, ...); returns 1 past the last kernel. */
int vx3_batch_set_profiling(vx3_batch *b, int on, int use_persistent);
int vx3_batch_kernel_stats(vx3_batch *b, int index, char *name, int name_cap, double *total_ms, int64_t *launches);

/* Fused step (voxcraft-sim_b200/csrc/engine/vx3_fused.cuh), OPT-IN with environment VX3_FUSED=1 at batch creation: batches
 * with a fixed link topology (no collisions / attach / detach / SecondaryExperiment / signals) are cut into spatial blocks
 * and stored block by block; links interior to a block are evaluated and their voxels integrated by one CTA, the end
 * forces passing through shared memory instead of HBM.  The ABI keeps the model's numbering and the results are
 * bit-identical to the two-pass kernels.  One visible difference: the link force / moment arrays of vx3_batch_state are
 * those of the LAST step of the last vx3_batch_step* call (only that step stores interior forces to HBM); after
 * vx3_batch_run they are unspecified for a simulation that stopped before the end of a launch chunk.
 *   set_fused(on = 0): run the two-pass kernels on a batch created with VX3_FUSED=1 (same storage order).
 *   fused_info: out4 = {active (0/1), blocks, links interior to a block, links of the face pre-pass}. */
int vx3_batch_set_fused(vx3_batch *b, int on);
int vx3_batch_fused_info(vx3_batch *b, int32_t *out4);
/* Host-only check of the block partition of `n` models (no device needed; test hook).  max_block_voxels <= 0: the
 * engine's default.  out8 = {blocks, interior links, face links, live links, voxels, largest block, voxels covered
 * exactly once (must equal voxels), links covered exactly once (must equal live links)}.  Returns VX3_ERR_INVALID when
 * the models do not fit the block model. */
int vx3_fused_plan_check(const vx3_model_desc *models, int n, int max_block_voxels, int64_t *out8);

/* --- slab decomposition of ONE body over the GPUs of a box (BASELINE config 5; no reference twin: the reference only
 * spreads independent files over devices, src/Executables/vx3_node_worker.cu:88-93).  Each rank creates a batch of one
 * sub-model: the voxels of its slab, VX3_VOX_GHOST copies of the neighbour slabs' face voxels and every link with an
 * owned end (host logic: voxcraft-sim_b200/parallel.py).  side 0 = lower neighbour, 1 = upper neighbour.
 *   halo_setup    the face voxels this rank sends to that neighbour and the ghost voxels it receives from it (voxel
 *                 indices of the sub-model; both ranks list them in the same order)
 *   halo_export   64-byte CUDA IPC handle of this rank's receive block for that side
 *   halo_connect  the handle the neighbour exported for its side facing this rank (+ its n_recv as a consistency check)
 * After connect, every step ends with the exchange of the face voxels' 64-byte pose records inside the step stream
 * (peer stores over NVLink + step-number flags, no host round trip).  All ranks must step in lock step with the same
 * explicit dt (vx3_batch_step_dt).  com_sums: raw centre-of-mass sums over OWNED voxels, to be added across ranks. */
int vx3_batch_halo_setup(vx3_batch *b, int side, int n_send, const int32_t *send_vox, int n_recv, const int32_t *recv_vox);
int vx3_batch_halo_export(vx3_batch *b, int side, void *handle64);
int vx3_batch_halo_connect(vx3_batch *b, int side, const void *peer_handle64, int peer_n_recv);
int vx3_batch_halo_connect_local(vx3_batch *b, int side, vx3_batch *peer); /* both slabs driven by this process */
int vx3_batch_com_sums(vx3_batch *b, int sim, double *out6);
/* Event counters of one simulation (test / diagnostics hook; no reference twin): out8 = {attach events, detach events, live link
 * slots in use, 0 (reserved), largest number of attach candidates one step produced, largest number of links one step put on the
 * failed list, 0, 0}. */
int vx3_batch_counters(vx3_batch *b, int sim, int64_t *out8);
/* Self-check of the depth-5 neighbour search (test hook; is_neighbor, VX3_VoxelyzeKernel.cu:651-680): n_pairs pseudo-random voxel
 * pairs of simulation `sim` — a voxel and one picked within its index neighbourhood, so that most pairs are a few links apart — are
 * put to the bounded two-sided search the contact phase uses and to the reference's path walk; *mismatches = pairs on which the
 * two disagree, *positives = pairs found within five links.  VX3_ERR_INVALID when the batch holds no adjacency table (nothing in it
 * attaches). */
int vx3_batch_check_neighbor_search(vx3_batch *b, int sim, int n_pairs, unsigned seed, int *mismatches, int *positives);
/* Queue k steps (explicit dt, or dt < 0 like vx3_batch_step) without waiting; vx3_batch_sync waits.  For one host thread
 * that drives several batches whose step streams wait for each other (slabs of a decomposed body): queue the slabs in
 * rounds of a few dozen steps — a round that overflows the driver's launch queue (~1000 launches) blocks the host on one
 * slab while that slab waits for a neighbour whose round has not been queued yet. */
int vx3_batch_step_async(vx3_batch *b, int64_t k, float dt);

/* Sort results like sortResults (src/VX3/VX3_SimulationManager.cu:472,
 * VX3_SimulationResult.h:26-33): fitness descending, NaN last.  Host-only. */
void vx3_sort_results(vx3_result *r, int n);

/* Destroying a batch returns its device arena, pinned staging buffer, stream and events to a small per-process cache
 * (at most two idle sets per device) that the next vx3_batch_create on that device reuses: a worker that evaluates one
 * batch after another pays cudaMalloc / cudaMallocHost / cudaFree once.  (The reference leaks by design instead:
 * src/old/VX3_MemoryCleaner.h:14-18, src/VX3/VX3_VoxelyzeKernel.cu:522.) */
void vx3_batch_destroy(vx3_batch *b);
/* Frees every idle cached arena / staging buffer / stream now. */
void vx3_engine_trim(void);

/* Thread-local description of the last error returned on this thread. */
const char *vx3_last_error(void);

int vx3_abi_version(void);
/* sizeof() of a struct of this header / vx3_model.h by name (binding self-check); 0 = unknown name. */
size_t vx3_abi_sizeof(const char *struct_name);

#ifdef __cplusplus
}
#endif
#endif /* VX3_ABI_H */
