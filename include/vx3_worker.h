/*
 * vx3_worker.h — the batch manager / node worker as a library call.
 *
 * Replaces the host side of vx3_node_worker (src/Executables/vx3_node_worker.cu:32-141) and
 * VX3_SimulationManager::start / readVXD / collectResults / sortResults (src/VX3/VX3_SimulationManager.cu:138-155,
 * 277-381, 428-472): reads the .vxt task file, shards the VXD files over the devices round-robin
 * (file i -> device i % nDevices), runs one engine batch per device (one host thread each), streams the .history
 * text to stdout, gathers and sorts the results (fitness descending, NaN last) and writes the .vxr report.
 */
#ifndef VX3_WORKER_H
#define VX3_WORKER_H

#include "vx3_abi.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vx3_worker_opts {
    int32_t n_devices;    /* 0 = every visible CUDA device (cudaGetDeviceCount, like the reference) */
    int32_t emit_history; /* 1 = honour RecordStepSize and write frames to stdout */
    int64_t max_steps;    /* 0 = the reference cap (1,000,000) */
    int32_t verbose;
    int32_t _pad;
} vx3_worker_opts;

/* Whole vx3_node_worker job: `vxt_path` names base VXA, input dir and VXD files; the report goes to `vxr_path`. */
int vx3_worker_run_vxt(const char *vxt_path, const char *vxr_path, const vx3_worker_opts *opts);

/* Same with the task given directly.  results (optional, capacity n) receives the sorted records. */
int vx3_worker_run_files(const char *base_vxa, const char *input_dir, const char *const *vxd_files, int n, const char *vxr_path,
                         const vx3_worker_opts *opts, vx3_result *results);

/* Report writer alone (report.inputdir / bestfit / detail.<name>..., src/Executables/vx3_node_worker.cu:98-141). */
int vx3_write_report(const char *vxr_path, const char *input_dir, const vx3_result *sorted, int n);

/* Per-voxel outputs of one result (collectResults, src/VX3/VX3_SimulationManager.cu:445-466), written by the report as
 * <init_pos> / <pos> ("x,y,z;" with std::to_string = %f) and <mats> ("id;") when the simulation's VXA sets
 * SavePositionOfAllVoxels (src/Executables/vx3_node_worker.cu:122-139).  n_voxels = 0: nothing to write for that result. */
typedef struct vx3_voxel_positions {
    int32_t n_voxels;
    int32_t _pad;
    const double *init_pos; /* [n_voxels][3] */
    const double *pos;      /* [n_voxels][3] */
    const int32_t *mats;    /* [n_voxels] matid */
} vx3_voxel_positions;

/* vx3_write_report with the per-voxel entries: positions is NULL or one record per result, in the same (sorted) order. */
int vx3_write_report_positions(const char *vxr_path, const char *input_dir, const vx3_result *sorted, const vx3_voxel_positions *positions, int n);

#ifdef __cplusplus
}
#endif
#endif /* VX3_WORKER_H */
