// ============================================================================
// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// C harness around the reference's OWN VX3 step loop, compiled for the host by oracle/Makefile (target ref_vx3)
// from the unmodified sources under /root/reference/src/VX3 + src/Utils (see vxhost.h for how) into
// oracle/_ref/libvxref_vx3.so.  It exists to PIN oracle/vx3_oracle.cpp on everything the reference's CPU library
// (src/old) does not have: per-voxel phase actuation, the all-pairs collision sweep, attach / detach, isNewLink,
// static-friction angMom=0, cilia, signals, SecondaryExperiment, the math-tree evaluator, CoM / angle / closeness /
// fitness, and the history frames CUDA_Simulation prints.
//
// The harness itself only does what VX3_SimulationManager::readVXD does around the kernel object
// (src/VX3/VX3_SimulationManager.cu:277-381): load the VXA through CVX_Sim, construct VX3_VoxelyzeKernel(CVX_Sim*),
// poke the VX3-only settings, then call the reference's functions:
//   vx3ref_init            = the prologue of CUDA_Simulation (:20-24, :54-55)
//   vx3ref_step            = VX3_VoxelyzeKernel::doTimeStep (VX3_VoxelyzeKernel.cu:237-359), k times
//   vx3ref_run_simulation  = the CUDA_Simulation kernel itself (:11-121), device printf captured
//   vx3ref_eval            = VX3_MathTree::eval (src/Utils/VX3_MathTree.h:50-192)
// The VX3-only settings come from the same flat vx3_model_desc the product consumes (options + token programs):
// the reference parses them with boost::property_tree (absent here); the VXA text gives CVX_Sim everything else.
// Programs the VXA leaves out are poked as the constants include/vx3_abi.h defines for "tag absent"
// (stop = 0, fitness = 0, attach condition = 1) — the reference reads an uninitialised value there.
// ============================================================================
#include "../../include/vx3_abi.h"

#include "VX3_MemoryCleaner.h"
#include "VX3_VoxelyzeKernel.cuh"
#include "VX_Environment.h"
#include "VX_Object.h"
#include "VX_Sim.h"

#include <map>
#include <new>
#include <unordered_set>

// ---- allocation registry: the reference's device containers run `delete main` on members that are uninitialised
// stack garbage when a VX3_Voxel / VX3_Material* temporary is constructed on the HOST (VX3_vector.cuh:48,
// VX3_queue.cuh:18 "Never called, since we copy the mem to GPU" — here they are called).  Every block handed out by
// operator new in this library is registered; operator delete ignores anything else.  Hidden visibility (Makefile)
// keeps these replacements private to this .so.
namespace {
template <class T> struct MallocAlloc {
    typedef T value_type;
    MallocAlloc() {}
    template <class U> MallocAlloc(const MallocAlloc<U> &) {}
    T *allocate(size_t n) { return (T *)malloc(n * sizeof(T)); }
    void deallocate(T *p, size_t) { free(p); }
    template <class U> bool operator==(const MallocAlloc<U> &) const { return true; }
    template <class U> bool operator!=(const MallocAlloc<U> &) const { return false; }
};
typedef std::unordered_set<void *, std::hash<void *>, std::equal_to<void *>, MallocAlloc<void *>> BlockSet;
BlockSet &blocks() { static BlockSet *s = new (malloc(sizeof(BlockSet))) BlockSet(); return *s; }
} // namespace
void *operator new(size_t n) {
    void *p = malloc(n ? n : 1);
    if (!p) throw std::bad_alloc();
    blocks().insert(p);
    return p;
}
void *operator new[](size_t n) { return operator new(n); }
void operator delete(void *p) noexcept {
    if (!p) return;
    auto it = blocks().find(p);
    if (it == blocks().end()) return; // garbage pointer, or a malloc() block of the reference: left alone (leaks by design)
    blocks().erase(it);
    free(p);
}
void operator delete[](void *p) noexcept { operator delete(p); }
void operator delete(void *p, size_t) noexcept { operator delete(p); }
void operator delete[](void *p, size_t) noexcept { operator delete(p); }

// src/old/VX3_MemoryCleaner.cu defines these; that file is host daemon plumbing and is not compiled here
bool VX3_MemoryCleaner_running = true;
boost::mutex MemoryCleaner_mutex;
std::vector<void *> MemoryCleaner_toBeFreedCUDAPointer;

// the reference's simulation kernel, compiled from VX3_SimulationManager.cu:11-121 by the Makefile
void CUDA_Simulation(VX3_VoxelyzeKernel *d_voxelyze_3, int num_simulation, int device_index);

struct vx3ref {
    CVX_Environment *env = nullptr;
    CVX_Sim *sim = nullptr;
    CVX_Object *obj = nullptr;
    VX3_VoxelyzeKernel *k = nullptr;
    std::string msg, out;
    bool diverged = false, inited = false;
    float lastDt = 0;
};

template <class T> static T *zero_new() { return new (calloc(1, sizeof(T))) T(); }

static void poke_program(VX3_MathTreeToken *dst, const vx3_program &p, bool have_default, double default_const) {
    if (p.n > 0 && p.tok) {
        for (int i = 0; i < p.n && i < 1024; i++) {
            dst[i].op = (VX3_MathTreeOperator)p.tok[i].op;
            dst[i].value = p.tok[i].value;
        }
    } else if (have_default) {
        dst[0].op = mtCONST;
        dst[0].value = default_const;
        dst[1].op = mtEND;
    }
}

extern "C" {
#define VXAPI __attribute__((visibility("default")))

VXAPI vx3ref *vx3ref_create(const char *vxa_path, const vx3_model_desc *d) {
    vx3ref *h = new vx3ref();
    h->env = zero_new<CVX_Environment>();
    h->sim = zero_new<CVX_Sim>();
    h->obj = zero_new<CVX_Object>();
    h->env->pObj = h->obj;
    h->sim->pEnv = h->env;
    if (!h->sim->LoadVXAFile(vxa_path, &h->msg)) return h;
    h->sim->Import(NULL, NULL, &h->msg);
    if (h->sim->Vx.voxelCount() <= 0) return h;
    // per-voxel externals of the flat model (ModelSpec.set_external has no VXA spelling): applied through the reference's
    // CVX_External API exactly like the boundary-condition loop of Import (src/VXA/VX_Sim.cpp:109-143)
    if (d->vox_ext && d->externals && d->n_voxels == h->sim->Vx.voxelCount()) {
        for (int i = 0; i < d->n_voxels; i++) {
            if (d->vox_ext[i] < 0) continue;
            const vx3_external &e = d->externals[d->vox_ext[i]];
            CVX_External *x = h->sim->Vx.voxel(i)->external();
            const dofComponent dofs[6] = {X_TRANSLATE, Y_TRANSLATE, Z_TRANSLATE, X_ROTATE, Y_ROTATE, Z_ROTATE};
            for (int c = 0; c < 6; c++)
                if (e.dof_fixed & dofs[c]) x->setDisplacement(dofs[c], c < 3 ? e.translation[c] : e.rotation[c - 3]);
            x->setForce(e.force[0], e.force[1], e.force[2]);
            x->setMoment(e.moment[0], e.moment[1], e.moment[2]);
        }
    }
    VX3_VoxelyzeKernel *k = new (calloc(1, sizeof(VX3_VoxelyzeKernel))) VX3_VoxelyzeKernel(h->sim);
    strncpy(k->vxa_filename, d->name, sizeof(k->vxa_filename) - 1);
    const vx3_sim_options &o = d->opt; // readVXD, VX3_SimulationManager.cu:321-365
    poke_program(k->StopConditionFormula, d->prog[VX3_PROG_STOP], true, 0.0);
    k->EnableCollision = o.enable_collision != 0;
    k->enableAttach = o.enable_attach != 0;
    k->enableDetach = o.enable_detach != 0;
    k->watchDistance = o.watch_distance;
    k->boundingRadius = o.bounding_radius;
    k->SafetyGuard = o.safety_guard;
    for (int c = 0; c < 5; c++) poke_program(k->AttachCondition[c], d->prog[VX3_PROG_ATTACH_0 + c], true, 1.0);
    k->RecordStepSize = o.record_step_size;
    k->RecordLink = o.record_link;
    k->RecordVoxel = o.record_voxel;
    poke_program(k->fitness_function, d->prog[VX3_PROG_FITNESS], true, 0.0);
    poke_program(k->force_field.token_x_forcefield, d->prog[VX3_PROG_FORCE_X], false, 0.0);
    poke_program(k->force_field.token_y_forcefield, d->prog[VX3_PROG_FORCE_Y], false, 0.0);
    poke_program(k->force_field.token_z_forcefield, d->prog[VX3_PROG_FORCE_Z], false, 0.0);
    k->SavePositionOfAllVoxels = o.save_position_of_all_voxels;
    k->MaxDistInVoxelLengthsToCountAsPair = o.max_dist_in_voxel_lengths_to_count_as_pair;
    k->EnableCilia = o.enable_cilia;
    k->EnableSignals = o.enable_signals;
    k->SecondaryExperiment = o.secondary_experiment;
    k->ReinitializeInitialPositionAfterThisManySeconds = o.reinit_initial_position_after_s;
    k->EnableExpansion = o.enable_expansion;
    h->k = k;
    return h;
}

VXAPI int vx3ref_ok(vx3ref *h) { return h && h->k != nullptr; }
VXAPI const char *vx3ref_message(vx3ref *h) { return h->msg.c_str(); }

// the prologue of CUDA_Simulation (VX3_SimulationManager.cu:20-24, 54-55)
VXAPI void vx3ref_init(vx3ref *h) {
    VX3_VoxelyzeKernel *k = h->k;
    k->syncVectors();
    k->saveInitialPosition();
    k->isSurfaceChanged = true;
    k->registerTargets();
    k->updateCurrentCenterOfMass();
    k->InitializeCenterOfMass();
    h->inited = true;
}

VXAPI double vx3ref_recommended_dt(vx3ref *h) { return h->k->recommendedTimeStep(); }

// k calls of doTimeStep(dt); stops at the first call that returns false (diverged)
VXAPI long vx3ref_step(vx3ref *h, long n, float dt) {
    if (!h->inited) vx3ref_init(h);
    long done = 0;
    for (; done < n && !h->diverged; done++) {
        if (!h->k->doTimeStep(dt)) {
            h->diverged = true;
            break;
        }
    }
    h->lastDt = dt;
    return done;
}

// runs the reference's CUDA_Simulation kernel (one thread) on a freshly created handle; returns the captured stdout size
VXAPI long vx3ref_run_simulation(vx3ref *h) {
    vxhost_out().clear();
    vxhost_capture() = true;
    vxhost_launch(CUDA_Simulation, dim3(1), dim3(1), h->k, 1, 0);
    vxhost_capture() = false;
    h->out.swap(vxhost_out());
    h->inited = true;
    return (long)h->out.size();
}
VXAPI const char *vx3ref_output(vx3ref *h) { return h->out.c_str(); }

VXAPI double vx3ref_eval(const vx3_token *tok, int n, const double *v) {
    static VX3_MathTreeToken buf[1024];
    for (int i = 0; i < 1024; i++) buf[i] = VX3_MathTreeToken();
    for (int i = 0; i < n && i < 1024; i++) {
        buf[i].op = (VX3_MathTreeOperator)tok[i].op;
        buf[i].value = tok[i].value;
    }
    return VX3_MathTree::eval(v[0], v[1], v[2], v[3], v[4], v[5], v[6], (int)v[7], (int)v[8], buf);
}

VXAPI int vx3ref_counts(vx3ref *h, int *n_voxels, int *n_links, int *n_surface, int *n_link_mats) {
    VX3_VoxelyzeKernel *k = h->k;
    if (n_voxels) *n_voxels = k->num_d_voxels;
    if (n_links) *n_links = h->inited ? (int)k->d_v_links.size() : k->num_d_links;
    if (n_surface) *n_surface = k->num_d_surface_voxels;
    if (n_link_mats) *n_link_mats = h->inited ? (int)k->d_v_linkMats.size() : k->num_d_linkMats;
    return 0;
}

VXAPI int vx3ref_surface(vx3ref *h, int *out, int cap) {
    VX3_VoxelyzeKernel *k = h->k;
    if (!k->d_surface_voxels) return 0;
    for (int i = 0; i < k->num_d_surface_voxels && i < cap; i++) out[i] = (int)(k->d_surface_voxels[i] - k->d_voxels);
    return k->num_d_surface_voxels;
}

// the fields collectResults reads (VX3_SimulationManager.cu:428-470); refresh = the closing
// updateCurrentCenterOfMass + computeFitness of CUDA_Simulation (:116-117)
VXAPI int vx3ref_result(vx3ref *h, vx3_result *r, int refresh) {
    VX3_VoxelyzeKernel *k = h->k;
    if (refresh) {
        k->updateCurrentCenterOfMass();
        k->computeFitness();
    }
    memset(r, 0, sizeof(*r));
    strncpy(r->name, k->vxa_filename, sizeof(r->name) - 1);
    r->status = h->diverged ? VX3_SIM_DIVERGED : VX3_SIM_RUNNING;
    r->num_voxel = k->num_d_voxels;
    r->num_close_pairs = k->numClosePairs;
    r->steps = (int64_t)k->CurStepCount;
    r->num_links = (int)k->d_v_links.size();
    r->collision_count = k->collisionCount;
    r->current_time = k->currentTime;
    r->fitness_score = k->fitness_score;
    r->vox_size = k->voxSize;
    r->initial_com[0] = k->initialCenterOfMass.x; r->initial_com[1] = k->initialCenterOfMass.y; r->initial_com[2] = k->initialCenterOfMass.z;
    r->current_com[0] = k->currentCenterOfMass.x; r->current_com[1] = k->currentCenterOfMass.y; r->current_com[2] = k->currentCenterOfMass.z;
    r->recent_angle = k->recentAngle;
    r->target_closeness = k->targetCloseness;
    r->dt = h->lastDt;
    for (int j = 0; j < k->num_d_voxels; j++) {
        if (k->d_voxels[j].isMeasured) {
            r->num_measured_voxel++;
            Vec3D<> a(k->d_voxels[j].pos.x, k->d_voxels[j].pos.y, k->d_voxels[j].pos.z);
            Vec3D<> b(k->d_initialPosition[j].x, k->d_initialPosition[j].y, k->d_initialPosition[j].z);
            r->total_distance_of_all_voxels += a.Dist(b);
        }
    }
    return 0;
}

static void put3(double *dst, size_t i, const VX3_Vec3D<double> &v) {
    if (dst) { dst[3 * i] = v.x; dst[3 * i + 1] = v.y; dst[3 * i + 2] = v.z; }
}

VXAPI int vx3ref_state(vx3ref *h, vx3_state_view *w) {
    VX3_VoxelyzeKernel *k = h->k;
    const int nv = k->num_d_voxels, nl = (int)k->d_v_links.size();
    if (w->n_voxels < nv || w->n_links < nl) {
        w->n_voxels = nv;
        w->n_links = nl;
        return -1;
    }
    w->n_voxels = nv;
    w->n_links = nl;
    std::map<VX3_Link *, int> lindex;
    for (int i = 0; i < nl; i++) lindex[k->d_v_links[i]] = i;
    std::map<VX3_MaterialLink *, int> mindex;
    for (int i = 0; i < (int)k->d_v_linkMats.size(); i++) mindex[k->d_v_linkMats[i]] = i;
    for (int i = 0; i < nv; i++) {
        VX3_Voxel &v = k->d_voxels[i];
        put3(w->pos, i, v.pos);
        put3(w->lin_mom, i, v.linMom);
        put3(w->ang_mom, i, v.angMom);
        put3(w->contact_force, i, v.contactForce);
        if (w->orient) { w->orient[4 * i] = v.orient.w; w->orient[4 * i + 1] = v.orient.x; w->orient[4 * i + 2] = v.orient.y; w->orient[4 * i + 3] = v.orient.z; }
        if (w->vox_flags) w->vox_flags[i] = (int)v.boolStates;
        if (w->temp) w->temp[i] = v.tempe;
        if (w->vox_links)
            for (int d = 0; d < 6; d++) w->vox_links[6 * i + d] = v.links[d] ? lindex.at(v.links[d]) : -1;
        if (w->signal) {
            double *o = w->signal + 6 * (size_t)i;
            o[0] = v.localSignal; o[1] = v.localSignaldt; o[2] = v.inactiveUntil; o[3] = v.packmakerNextPulse; o[4] = v.d_signal.value; o[5] = v.d_signal.activeTime;
        }
    }
    for (int i = 0; i < nl; i++) {
        VX3_Link &l = *k->d_v_links[i];
        if (w->link_vneg) w->link_vneg[i] = (int)(l.pVNeg - k->d_voxels);
        if (w->link_vpos) w->link_vpos[i] = (int)(l.pVPos - k->d_voxels);
        if (w->link_axis) w->link_axis[i] = (int)l.axis;
        if (w->link_mat) w->link_mat[i] = mindex.at(l.mat);
        put3(w->link_pos2, i, l.pos2);
        put3(w->link_angle1v, i, l.angle1v);
        put3(w->link_angle2v, i, l.angle2v);
        put3(w->link_force_neg, i, l.forceNeg);
        put3(w->link_force_pos, i, l.forcePos);
        put3(w->link_moment_neg, i, l.momentNeg);
        put3(w->link_moment_pos, i, l.momentPos);
        if (w->link_strain) w->link_strain[i] = l.strain;
        if (w->link_max_strain) w->link_max_strain[i] = l.maxStrain;
        if (w->link_strain_offset) w->link_strain_offset[i] = l.strainOffset;
        if (w->link_stress) w->link_stress[i] = l._stress;
        if (w->link_flags)
            w->link_flags[i] = ((l.boolStates & LOCAL_VELOCITY_VALID) ? VX3_LINKSTATE_LOCAL_VELOCITY_VALID : 0) |
                               (l.smallAngle ? VX3_LINKSTATE_SMALL_ANGLE : 0) | (l.isDetached ? VX3_LINKSTATE_DETACHED : 0) |
                               (l.removed ? VX3_LINKSTATE_REMOVED : 0) | (l.isNewLink << VX3_LINKSTATE_NEWLINK_SHIFT);
        if (w->link_rest_length) w->link_rest_length[i] = l.currentRestLength;
    }
    return 0;
}

// per-voxel extras the state view does not carry: enableAttach, removed (for the VX3-only flag bits)
VXAPI int vx3ref_voxel_extras(vx3ref *h, int32_t *enable_attach, int32_t *removed) {
    for (int i = 0; i < h->k->num_d_voxels; i++) {
        if (enable_attach) enable_attach[i] = h->k->d_voxels[i].enableAttach;
        if (removed) removed[i] = h->k->d_voxels[i].removed;
    }
    return 0;
}

} // extern "C"
