"""ctypes mirror of include/vx3_abi.h and include/vx3_model.h.

Plumbing only: every struct here is a field-for-field copy of the C header (checked against
the compiled library by ``vx3_abi_sizeof`` in tests/test_abi.py).  No physics lives in Python.
"""
import ctypes as C

VOX_GHOST = 1 << 7  # include/vx3_abi.h VX3_VOX_GHOST
VX3_PROG_COUNT = 10
(PROG_STOP, PROG_FITNESS, PROG_FORCE_X, PROG_FORCE_Y, PROG_FORCE_Z, PROG_ATTACH_0, PROG_ATTACH_1, PROG_ATTACH_2,
 PROG_ATTACH_3, PROG_ATTACH_4) = range(10)

OPS = ["END", "CONST", "E", "PI", "VAR", "ADD", "SUB", "MUL", "DIV", "POW", "SQRT", "SIN", "COS", "TAN", "ATAN", "LOG",
       "INT", "ABS", "NOT", "GREATERTHAN", "LESSTHAN", "AND", "OR", "NORMALCDF"]
OP = {name: i for i, name in enumerate(OPS)}
VARS = {"x": 0, "y": 1, "z": 2, "hit": 3, "t": 4, "angle": 5, "targetCloseness": 6, "numClosePairs": 7, "num_voxel": 8}

VOX_SURFACE = 1 << 1
VOX_FLOOR_ENABLED = 1 << 2
VOX_FLOOR_STATIC_FRICTION = 1 << 3
VOX_COLLISIONS_ENABLED = 1 << 5
LINKSTATE_LOCAL_VELOCITY_VALID = 1 << 0
LINKSTATE_SMALL_ANGLE = 1 << 1
LINKSTATE_DETACHED = 1 << 2
LINKSTATE_REMOVED = 1 << 3
LINKSTATE_NEWLINK_SHIFT = 8

SIM_RUNNING, SIM_STOPPED, SIM_DIVERGED, SIM_STEP_CAP = range(4)

i32, i64, f32, f64 = C.c_int32, C.c_int64, C.c_float, C.c_double
P = C.POINTER


class Token(C.Structure):
    _fields_ = [("op", i32), ("_pad", i32), ("value", f64)]


class Program(C.Structure):
    _fields_ = [("n", i32), ("_pad", i32), ("tok", P(Token))]


class VoxelMaterial(C.Structure):
    _fields_ = [("matid", i32), ("fixed", i32), ("sticky", i32), ("is_target", i32), ("is_measured", i32), ("linear", i32),
                ("is_pacemaker", i32), ("is_electrical_active", i32), ("r", i32), ("g", i32), ("b", i32), ("a", i32),
                ("E", f32), ("sigmaYield", f32), ("sigmaFail", f32), ("epsilonYield", f32), ("epsilonFail", f32),
                ("nu", f32), ("rho", f32), ("alphaCTE", f32), ("muStatic", f32), ("muKinetic", f32),
                ("zetaInternal", f32), ("zetaGlobal", f32), ("zetaCollision", f32), ("eHat", f32),
                ("gravMult", f32), ("mass", f32), ("massInverse", f32), ("sqrtMass", f32), ("firstMoment", f32),
                ("momentInertia", f32), ("momentInertiaInverse", f32), ("_2xSqMxExS", f32), ("_2xSqIxExSxSxS", f32),
                ("n_data", i32), ("strain_data", P(f32)), ("stress_data", P(f32)),
                ("nomSize", f64), ("extScale", f64 * 3), ("cilia", f64),
                ("pacemaker_period", f64), ("signal_value_decay", f64), ("signal_time_delay", f64), ("inactive_period", f64),
                ("remove_after_s", f64), ("thermal_on_after_s", f64), ("cilia_on_after_s", f64)]


class LinkMaterial(C.Structure):
    _fields_ = [("m", VoxelMaterial), ("vox1_mat", i32), ("vox2_mat", i32),
                ("a1", f32), ("a2", f32), ("b1", f32), ("b2", f32), ("b3", f32),
                ("sqA1", f32), ("sqA2xIp", f32), ("sqB1", f32), ("sqB2xFMp", f32), ("sqB3xIp", f32)]


class External(C.Structure):
    _fields_ = [("dof_fixed", i32), ("force", f32 * 3), ("moment", f32 * 3), ("translation", f64 * 3),
                ("rotation", f64 * 3), ("rotation_q", f64 * 4)]


class SimOptions(C.Structure):
    _fields_ = [("vox_size", f64), ("dt_frac", f64), ("temp_enabled", i32), ("vary_temp_enabled", i32),
                ("temp_base", f64), ("temp_amplitude", f64), ("temp_period", f64),
                ("enable_collision", i32), ("enable_attach", i32), ("enable_detach", i32),
                ("watch_distance", f64), ("bounding_radius", f64), ("safety_guard", i32),
                ("record_step_size", i32), ("record_link", i32), ("record_voxel", i32),
                ("save_position_of_all_voxels", i32), ("max_dist_in_voxel_lengths_to_count_as_pair", f64),
                ("enable_cilia", i32), ("enable_signals", i32), ("secondary_experiment", i32),
                ("reinit_initial_position_after_s", f64), ("enable_expansion", i32), ("_pad", i32)]


class ModelDesc(C.Structure):
    _fields_ = [("name", C.c_char * 256),
                ("n_voxel_mats", i32), ("n_link_mats", i32), ("voxel_mats", P(VoxelMaterial)), ("link_mats", P(LinkMaterial)),
                ("n_voxels", i32), ("n_links", i32), ("n_externals", i32), ("link_capacity", i32),
                ("ix", P(C.c_int16)), ("iy", P(C.c_int16)), ("iz", P(C.c_int16)), ("vox_mat", P(i32)),
                ("pos", P(f64)), ("orient", P(f64)), ("lin_mom", P(f64)), ("ang_mom", P(f64)),
                ("vox_flags", P(i32)), ("temp", P(f32)), ("phase_offset", P(f64)), ("vox_links", P(i32)), ("vox_ext", P(i32)),
                ("base_cilia", P(f64)), ("shift_cilia", P(f64)), ("externals", P(External)),
                ("link_vneg", P(i32)), ("link_vpos", P(i32)), ("link_axis", P(i32)), ("link_mat", P(i32)),
                ("link_pos2", P(f64)), ("link_angle1v", P(f64)), ("link_angle2v", P(f64)),
                ("link_strain", P(f32)), ("link_max_strain", P(f32)), ("link_strain_offset", P(f32)), ("link_stress", P(f32)),
                ("link_flags", P(i32)), ("link_small_angle", P(i32)), ("link_rest_length", P(f64)),
                ("link_transverse_area", P(f32)), ("link_transverse_strain_sum", P(f32)), ("link_strain_ratio", P(f32)),
                ("opt", SimOptions), ("prog", Program * VX3_PROG_COUNT)]


class Result(C.Structure):
    _fields_ = [("name", C.c_char * 256), ("status", i32), ("num_voxel", i32), ("num_measured_voxel", i32),
                ("num_close_pairs", i32), ("steps", i64), ("num_links", i32), ("collision_count", i32),
                ("current_time", f64), ("fitness_score", f64), ("vox_size", f64), ("initial_com", f64 * 3),
                ("current_com", f64 * 3), ("total_distance_of_all_voxels", f64), ("recent_angle", f64),
                ("target_closeness", f64), ("dt", f64)]


class StateView(C.Structure):
    _fields_ = [("n_voxels", i32), ("n_links", i32),
                ("pos", P(f64)), ("orient", P(f64)), ("lin_mom", P(f64)), ("ang_mom", P(f64)),
                ("vox_flags", P(i32)), ("temp", P(f32)), ("vox_links", P(i32)), ("contact_force", P(f64)),
                ("link_vneg", P(i32)), ("link_vpos", P(i32)), ("link_axis", P(i32)), ("link_mat", P(i32)),
                ("link_pos2", P(f64)), ("link_angle1v", P(f64)), ("link_angle2v", P(f64)),
                ("link_force_neg", P(f64)), ("link_force_pos", P(f64)), ("link_moment_neg", P(f64)), ("link_moment_pos", P(f64)),
                ("link_strain", P(f32)), ("link_max_strain", P(f32)), ("link_strain_offset", P(f32)), ("link_stress", P(f32)),
                ("link_flags", P(i32)), ("link_rest_length", P(f64)), ("signal", P(f64))]


class RunOpts(C.Structure):
    _fields_ = [("max_steps", i64), ("steps_per_launch", i32), ("emit_history", i32)]


HISTORY_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_int, P(C.c_char), C.c_size_t)


class MaterialParams(C.Structure):
    _fields_ = [("mat_model", i32), ("n_data", i32), ("strain_data", P(f64)), ("stress_data", P(f64)),
                ("elastic_mod", f64), ("plastic_mod", f64), ("yield_stress", f64), ("fail_stress", f64), ("fail_strain", f64),
                ("density", f64), ("poissons_ratio", f64), ("cte", f64), ("material_temp_phase", f64),
                ("u_static", f64), ("u_dynamic", f64),
                ("is_pacemaker", i32), ("is_measured", i32), ("is_electrical_active", i32), ("is_target", i32),
                ("fixed", i32), ("sticky", i32),
                ("pacemaker_period", f64), ("signal_value_decay", f64), ("signal_time_delay", f64), ("inactive_period", f64),
                ("remove_after_s", f64), ("thermal_on_after_s", f64), ("cilia_on_after_s", f64), ("cilia", f64),
                ("red", f64), ("green", f64), ("blue", f64), ("alpha", f64)]


class EnvParams(C.Structure):
    _fields_ = [("grav_enabled", i32), ("grav_acc", f64), ("floor_enabled", i32), ("temp_enabled", i32),
                ("temp_base", f64), ("temp_amplitude", f64), ("vary_temp_enabled", i32), ("temp_period", f64),
                ("bond_damping_z", f64), ("col_damping_z", f64), ("slow_damping_z", f64),
                ("volume_effects_enabled", i32), ("self_col_enabled", i32)]


STRUCTS = {"vx3_token": Token, "vx3_program": Program, "vx3_voxel_material": VoxelMaterial,
           "vx3_link_material": LinkMaterial, "vx3_external": External, "vx3_sim_options": SimOptions,
           "vx3_model_desc": ModelDesc, "vx3_result": Result, "vx3_state_view": StateView, "vx3_run_opts": RunOpts,
           "vx3_material_params": MaterialParams, "vx3_env_params": EnvParams}


def declare_model_api(lib):
    """argtypes/restypes for include/vx3_model.h."""
    vp = C.c_void_p
    lib.vx3_material_params_default.argtypes = [P(MaterialParams)]
    lib.vx3_env_params_default.argtypes = [P(EnvParams)]
    lib.vx3_sim_options_default.argtypes = [P(SimOptions)]
    lib.vx3_builder_create.argtypes = [f64]
    lib.vx3_builder_create.restype = vp
    lib.vx3_builder_destroy.argtypes = [vp]
    lib.vx3_builder_add_material.argtypes = [vp, P(MaterialParams)]
    lib.vx3_builder_set_env.argtypes = [vp, P(EnvParams)]
    lib.vx3_builder_set_options.argtypes = [vp, P(SimOptions)]
    lib.vx3_builder_set_name.argtypes = [vp, C.c_char_p]
    lib.vx3_builder_set_program.argtypes = [vp, C.c_int, P(Token), C.c_int]
    lib.vx3_builder_set_structure.argtypes = [vp, C.c_int, C.c_int, C.c_int, P(C.c_uint8), P(f64), P(f64), P(f64)]
    lib.vx3_builder_set_external.argtypes = [vp, C.c_int, P(External)]
    lib.vx3_builder_build.argtypes = [vp]
    lib.vx3_builder_build.restype = P(ModelDesc)
    lib.vx3_vxa_parse.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p]
    lib.vx3_vxa_parse.restype = vp
    lib.vx3_vxa_load.argtypes = [C.c_char_p, C.c_char_p]
    lib.vx3_vxa_load.restype = vp
    lib.vx3_model_recommended_dt.argtypes = [P(ModelDesc)]
    lib.vx3_model_recommended_dt.restype = f64
    lib.vx3_model_first_step_dt.argtypes = [P(ModelDesc)]
    lib.vx3_model_first_step_dt.restype = f64
    lib.vx3_model_last_error.restype = C.c_char_p
    return lib


def declare_engine_api(lib):
    """argtypes/restypes for include/vx3_abi.h."""
    vp = C.c_void_p
    lib.vx3_batch_create.argtypes = [C.c_int, P(ModelDesc), C.c_int, P(vp)]
    lib.vx3_batch_run.argtypes = [vp, P(RunOpts), HISTORY_CB, vp]
    lib.vx3_batch_step.argtypes = [vp, i64]
    lib.vx3_batch_step_dt.argtypes = [vp, i64, f32]
    lib.vx3_batch_sync.argtypes = [vp]
    lib.vx3_batch_state.argtypes = [vp, C.c_int, P(StateView)]
    lib.vx3_batch_results.argtypes = [vp, P(Result)]
    lib.vx3_batch_positions.argtypes = [vp, C.c_int, P(f64), P(f64), P(i32)]
    lib.vx3_batch_recommended_dt.argtypes = [vp, C.c_int, P(f64)]
    lib.vx3_batch_last_timing.argtypes = [vp, P(f64), P(i64)]
    lib.vx3_batch_set_profiling.argtypes = [vp, C.c_int, C.c_int]
    lib.vx3_batch_kernel_stats.argtypes = [vp, C.c_int, C.c_char_p, C.c_int, P(f64), P(i64)]
    lib.vx3_batch_set_fused.argtypes = [vp, C.c_int]
    lib.vx3_batch_fused_info.argtypes = [vp, P(i32)]
    lib.vx3_fused_plan_check.argtypes = [P(ModelDesc), C.c_int, C.c_int, P(i64)]
    lib.vx3_batch_halo_setup.argtypes = [vp, C.c_int, C.c_int, P(i32), C.c_int, P(i32)]
    lib.vx3_batch_halo_export.argtypes = [vp, C.c_int, vp]
    lib.vx3_batch_halo_connect.argtypes = [vp, C.c_int, vp, C.c_int]
    lib.vx3_batch_com_sums.argtypes = [vp, C.c_int, P(f64)]
    lib.vx3_batch_counters.argtypes = [vp, C.c_int, P(C.c_int64)]
    lib.vx3_batch_check_neighbor_search.argtypes = [vp, C.c_int, C.c_int, C.c_uint, P(C.c_int), P(C.c_int)]
    lib.vx3_batch_halo_connect_local.argtypes = [vp, C.c_int, vp]
    lib.vx3_batch_step_async.argtypes = [vp, i64, f32]
    lib.vx3_abi_sizeof.argtypes = [C.c_char_p]
    lib.vx3_abi_sizeof.restype = C.c_size_t
    lib.vx3_sort_results.argtypes = [P(Result), C.c_int]
    lib.vx3_sort_results.restype = None
    lib.vx3_batch_destroy.argtypes = [vp]
    lib.vx3_batch_destroy.restype = None
    lib.vx3_engine_trim.argtypes = []
    lib.vx3_engine_trim.restype = None
    lib.vx3_last_error.restype = C.c_char_p
    lib.vx3_abi_version.restype = C.c_int
    return lib


ENGINE_SYMBOLS = ["vx3_batch_create", "vx3_batch_run", "vx3_batch_step", "vx3_batch_step_dt", "vx3_batch_sync",
                  "vx3_batch_state", "vx3_batch_results", "vx3_batch_positions", "vx3_batch_recommended_dt",
                  "vx3_batch_last_timing", "vx3_batch_set_profiling", "vx3_batch_kernel_stats", "vx3_batch_set_fused",
                  "vx3_batch_fused_info", "vx3_fused_plan_check", "vx3_batch_halo_setup",
                  "vx3_batch_halo_export", "vx3_batch_halo_connect", "vx3_batch_halo_connect_local", "vx3_batch_com_sums",
                  "vx3_batch_counters", "vx3_batch_check_neighbor_search", "vx3_batch_step_async", "vx3_abi_sizeof", "vx3_sort_results", "vx3_batch_destroy", "vx3_engine_trim", "vx3_last_error",
                  "vx3_abi_version"]
WORKER_SYMBOLS = ["vx3_worker_run_vxt", "vx3_worker_run_files", "vx3_write_report", "vx3_write_report_positions"]
MODEL_SYMBOLS = ["vx3_material_params_default", "vx3_env_params_default", "vx3_sim_options_default", "vx3_builder_create",
                 "vx3_builder_destroy", "vx3_builder_add_material", "vx3_builder_set_env", "vx3_builder_set_options",
                 "vx3_builder_set_name", "vx3_builder_set_program", "vx3_builder_set_structure", "vx3_builder_set_external",
                 "vx3_builder_build", "vx3_vxa_parse", "vx3_vxa_load", "vx3_model_recommended_dt", "vx3_model_first_step_dt", "vx3_model_last_error"]
