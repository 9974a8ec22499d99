/*
 * vx3_model.h — host-side model builder (C ABI).
 *
 * Produces the flat vx3_model_desc consumed by vx3_batch_create() from the same
 * inputs the reference feeds CVX_Sim::Import (src/VXA/VX_Sim.cpp:59-153): a palette
 * of VXC materials, a lattice structure, environment and simulator settings.
 * It reproduces the reference's voxel order, link order, link-material table and
 * every derived material constant bit-for-bit, so a
 * model built here equals the one exported from a CVX_Sim.
 *
 * Host code only: nothing in this header touches the GPU.
 */
#ifndef VX3_MODEL_H
#define VX3_MODEL_H

#include "vx3_abi.h"

#ifdef __cplusplus
extern "C" {
#endif

/* CVXC_Material <Mechanical> block (src/VXA/VX_Object.cpp:1395-1460) + <Display>.
 * Defaults (vx3_material_params_default) = the parser's "tag absent" values. */
typedef struct vx3_material_params {
    int32_t mat_model; /* 0 linear, 1 linear+fail, 2 bilinear, 3 data (MDL_*) */
    int32_t n_data;    /* MDL_DATA only */
    const double *strain_data, *stress_data;
    double elastic_mod, plastic_mod, yield_stress, fail_stress, fail_strain;
    double density, poissons_ratio, cte, material_temp_phase, u_static, u_dynamic;
    int32_t is_pacemaker, is_measured, is_electrical_active, is_target, fixed, sticky;
    double pacemaker_period, signal_value_decay, signal_time_delay, inactive_period;
    double remove_after_s, thermal_on_after_s, cilia_on_after_s;
    double cilia;
    double red, green, blue, alpha; /* 0..1 */
} vx3_material_params;

/* CVX_Environment (src/old/VX_Environment.cpp:22-37,101-178) + the CVX_Sim fields Import
 * reads (src/VXA/VX_Sim.cpp:66-83,148; parse :185-294). */
typedef struct vx3_env_params {
    int32_t grav_enabled; /* default 1 */
    double grav_acc;      /* default -9.81 */
    int32_t floor_enabled; /* default 1 */
    int32_t temp_enabled; /* default 1 (CVX_Environment ctor) */
    double temp_base;     /* 25 */
    double temp_amplitude; /* 0 */
    int32_t vary_temp_enabled; /* 0 */
    double temp_period;   /* 0.1 */
    double bond_damping_z, col_damping_z, slow_damping_z; /* CVX_Sim: 0/0/0 when <Damping> absent */
    int32_t volume_effects_enabled; /* Features.VolumeEffectsEnabled, default 0 → all nu forced to 0 */
    int32_t self_col_enabled; /* Collisions.SelfColEnabled: sets COLLISIONS_ENABLED voxel flag */
} vx3_env_params;

typedef struct vx3_builder vx3_builder;

void vx3_material_params_default(vx3_material_params *p);
void vx3_env_params_default(vx3_env_params *p);
void vx3_sim_options_default(vx3_sim_options *o); /* the VX3 readVXD defaults */

/* lattice_dim = VXC Lattice_Dim = voxel size (m). */
vx3_builder *vx3_builder_create(double lattice_dim);
void vx3_builder_destroy(vx3_builder *b);

/* Appends a palette material; returns its 1-based palette index (= matid), <0 on error. */
int vx3_builder_add_material(vx3_builder *b, const vx3_material_params *p);
int vx3_builder_set_env(vx3_builder *b, const vx3_env_params *e);
int vx3_builder_set_options(vx3_builder *b, const vx3_sim_options *o);
int vx3_builder_set_name(vx3_builder *b, const char *name);
/* Token program in reference order (see vx3_program). */
int vx3_builder_set_program(vx3_builder *b, int slot, const vx3_token *tok, int n);
/* mat: nx*ny*nz palette indices (0 = empty), x fastest then y then z
 * (src/VXA/VX_Object.cpp:1826-1829).  phase_offset / base_cilia / shift_cilia are per
 * lattice CELL (same indexing; cilia arrays are [cell][3]) or NULL.  The reference
 * consumes the per-cell layers only for filled cells (:1852-1957). */
int vx3_builder_set_structure(vx3_builder *b, int nx, int ny, int nz, const uint8_t *mat,
                              const double *phase_offset, const double *base_cilia,
                              const double *shift_cilia);
/* Attach a VX3_External (fixed DOFs, prescribed displacement, external force/moment) to voxel
 * `voxel_index` (index in build order = lattice scan order over filled cells).  This is what the
 * boundary-condition loop of Import produces per touched voxel (src/VXA/VX_Sim.cpp:109-143). */
int vx3_builder_set_external(vx3_builder *b, int voxel_index, const vx3_external *e);

/* Runs the Import sequence and returns the flat model.  The pointer and every array it
 * references stay valid until the builder is destroyed or built again. */
const vx3_model_desc *vx3_builder_build(vx3_builder *b);

/* VXA / VXD front end.  Parses the VXA text, applies the VXD overrides (children of <VXD> carrying
 * replace="VXA.path" replace that subtree of the base VXA, src/Utils/ctool.h:49-57) and returns a builder ready
 * for vx3_builder_build().  Replaces CVX_Sim::ReadVXA (src/VXA/VX_Sim.cpp:155-294), CVX_Environment::ReadXML
 * (src/old/VX_Environment.cpp:101-178), CVX_Object::ReadXML (src/VXA/VX_Object.cpp), the VX3 tag reads and
 * ParseMathTree of readVXD (src/VX3/VX3_SimulationManager.cu:157-376).  NULL on error (vx3_model_last_error). */
vx3_builder *vx3_vxa_parse(const char *vxa_xml, const char *vxd_xml /* may be NULL */, const char *name /* may be NULL */);
vx3_builder *vx3_vxa_load(const char *vxa_path, const char *vxd_path /* may be NULL */);

/* Host recommendedTimeStep (src/VX3/VX3_VoxelyzeKernel.cu:184-217) on a flat model. */
double vx3_model_recommended_dt(const vx3_model_desc *m);
/* OptimalDt as the reference's first doTimeStep(dt < 0) evaluates it (VX3_VoxelyzeKernel.cu:240-247): recommendedTimeStep() AFTER that
 * step's updateTemperature, i.e. with the rest lengths at the t = 0 temperatures — differs from vx3_model_recommended_dt only when a
 * link material has nu != 0 (stiffness eHat * area / ((1 + strain) * restLength), VX3_Link.cu:268-277).  What a run with dt < 0 uses. */
double vx3_model_first_step_dt(const vx3_model_desc *m);

const char *vx3_model_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* VX3_MODEL_H */
