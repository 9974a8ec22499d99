// ============================================================================
// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// Thin C shim around the UNMODIFIED reference CPU library (src/old + src/VXA), compiled
// in place from /root/reference by oracle/Makefile into oracle/_ref/libvxref.so.
// No reference source is copied into this repository; this file only calls the
// reference's public classes:
//   * load a VXA through CVX_Sim::LoadVXAFile + Import — the same wiring as
//     src/VX3/VX3_SimulationManager.cu:300-308;
//   * export the resulting host object graph as a flat vx3_model_desc — exactly the
//     conversion a reference maintainer adds to call the B200 engine instead of
//     "VX3_VoxelyzeKernel h_d_tmp(&MainSim)" (field list mirrors
//     the VX3_* host constructors: VX3_Voxel.cu:10-50, VX3_Link.cu:6-30,
//     VX3_Material.cu:4-30, VX3_MaterialVoxel.cu:4-11, VX3_MaterialLink.cu:4-22);
//   * step with CVoxelyze::doTimeStep (src/old/Voxelyze.cpp:251-284) and dump state.
// ============================================================================
#include "../include/vx3_abi.h"

#include "VX_Environment.h"
#include "VX_Link.h"
#include "VX_MaterialLink.h"
#include "VX_MaterialVoxel.h"
#include "VX_Object.h"
#include "VX_Sim.h"
#include "VX_Voxel.h"
#include "Voxelyze.h"

#include <cstring>
#include <map>
#include <string>
#include <vector>

struct vxref {
    // zero-initialised storage: CVX_Sim's ctor leaves the damping ratios unset
    CVX_Environment *env;
    CVX_Sim *sim;
    CVX_Object *obj;
    std::string msg;
    // flat export (owns memory)
    vx3_model_desc d;
    std::vector<vx3_voxel_material> vmats;
    std::vector<vx3_link_material> lmats;
    std::vector<std::vector<float>> data;
    std::vector<int16_t> ix, iy, iz;
    std::vector<int32_t> vmat, vflags, vlinks, vext, lneg, lpos, laxis, lmat, lflags, lsmall;
    std::vector<double> pos, orient, linmom, angmom, phase, bcil, scil, lpos2, la1v, la2v, lrest;
    std::vector<float> temp, lstrain, lmaxstrain, loffset, lstress, larea, ltsum, lratio;
    std::vector<vx3_external> exts;
};

template <class T> static T *zero_new() {
    void *p = calloc(1, sizeof(T));
    return new (p) T();
}

static void fill_mat(CVX_MaterialVoxel *p, vx3_voxel_material &o, std::vector<std::vector<float>> &keep) {
    memset(&o, 0, sizeof(o));
    o.matid = p->matid;
    o.fixed = p->fixed; o.sticky = p->sticky; o.is_target = p->isTarget; o.is_measured = p->isMeasured;
    o.linear = p->linear; o.is_pacemaker = p->isPaceMaker; o.is_electrical_active = p->isElectricalActive;
    o.r = p->r; o.g = p->g; o.b = p->b; o.a = p->a;
    o.E = p->E; o.sigmaYield = p->sigmaYield; o.sigmaFail = p->sigmaFail; o.epsilonYield = p->epsilonYield; o.epsilonFail = p->epsilonFail;
    o.nu = p->nu; o.rho = p->rho; o.alphaCTE = p->alphaCTE; o.muStatic = p->muStatic; o.muKinetic = p->muKinetic;
    o.zetaInternal = p->zetaInternal; o.zetaGlobal = p->zetaGlobal; o.zetaCollision = p->zetaCollision;
    o.eHat = p->_eHat;
    o.gravMult = p->gravMult; o.mass = p->_mass; o.massInverse = p->_massInverse; o.sqrtMass = p->_sqrtMass;
    o.firstMoment = p->_firstMoment; o.momentInertia = p->_momentInertia; o.momentInertiaInverse = p->_momentInertiaInverse;
    o._2xSqMxExS = p->_2xSqMxExS; o._2xSqIxExSxSxS = p->_2xSqIxExSxSxS;
    keep.push_back(p->strainData);
    o.strain_data = keep.back().data();
    keep.push_back(p->stressData);
    o.stress_data = keep.back().data();
    o.n_data = (int)p->strainData.size();
    o.nomSize = p->nomSize;
    o.extScale[0] = p->extScale.x; o.extScale[1] = p->extScale.y; o.extScale[2] = p->extScale.z;
    o.cilia = p->Cilia;
    o.pacemaker_period = p->PaceMakerPeriod;
    o.signal_value_decay = p->signalValueDecay;
    o.signal_time_delay = p->signalTimeDelay;
    o.inactive_period = p->inactivePeriod;
    o.remove_after_s = p->RemoveFromSimulationAfterThisManySeconds;
    o.thermal_on_after_s = p->TurnOnThermalExpansionAfterThisManySeconds;
    o.cilia_on_after_s = p->TurnOnCiliaAfterThisManySeconds;
}

extern "C" {

vxref *vxref_load_vxa(const char *path) {
    vxref *h = new vxref();
    h->env = zero_new<CVX_Environment>();
    h->sim = zero_new<CVX_Sim>();
    h->obj = zero_new<CVX_Object>();
    h->env->pObj = h->obj;
    h->sim->pEnv = h->env;
    if (!h->sim->LoadVXAFile(path, &h->msg)) return h; // caller checks vxref_ok
    h->sim->Import(NULL, NULL, &h->msg);
    return h;
}

int vxref_ok(vxref *h) { return h && h->sim->Vx.voxelCount() > 0; }
const char *vxref_message(vxref *h) { return h->msg.c_str(); }

// DtFrac and the <Damping> values parsed by ReadVXA
double vxref_dtfrac(vxref *h) { return h->sim->DtFrac; }

const vx3_model_desc *vxref_export(vxref *h) {
    CVoxelyze &Vx = h->sim->Vx;
    vx3_model_desc &d = h->d;
    memset(&d, 0, sizeof(d));
    h->data.clear();
    h->data.reserve(2 * (Vx.voxelMats.size() + Vx.linkMats.size()) + 4);

    std::map<CVX_MaterialVoxel *, int> vmIndex;
    h->vmats.resize(Vx.voxelMats.size());
    for (size_t i = 0; i < Vx.voxelMats.size(); i++) {
        vmIndex[Vx.voxelMats[i]] = (int)i;
        fill_mat(Vx.voxelMats[i], h->vmats[i], h->data);
    }
    std::map<CVX_MaterialLink *, int> lmIndex;
    h->lmats.resize(Vx.linkMats.size());
    {
        int i = 0;
        for (CVX_MaterialLink *p : Vx.linkMats) {
            lmIndex[p] = i;
            vx3_link_material &o = h->lmats[i];
            fill_mat(p, o.m, h->data);
            o.vox1_mat = vmIndex[p->vox1Mat];
            o.vox2_mat = vmIndex[p->vox2Mat];
            o.a1 = p->_a1; o.a2 = p->_a2; o.b1 = p->_b1; o.b2 = p->_b2; o.b3 = p->_b3;
            o.sqA1 = p->_sqA1; o.sqA2xIp = p->_sqA2xIp; o.sqB1 = p->_sqB1; o.sqB2xFMp = p->_sqB2xFMp; o.sqB3xIp = p->_sqB3xIp;
            i++;
        }
    }
    const int nV = Vx.voxelCount(), nL = Vx.linkCount();
    std::map<CVX_Voxel *, int> vIndex;
    std::map<CVX_Link *, int> lIndex;
    for (int i = 0; i < nV; i++) vIndex[Vx.voxelsList[i]] = i;
    for (int i = 0; i < nL; i++) lIndex[Vx.linksList[i]] = i;

    h->ix.resize(nV); h->iy.resize(nV); h->iz.resize(nV); h->vmat.resize(nV); h->vflags.resize(nV);
    h->vlinks.assign(6 * (size_t)nV, -1); h->vext.assign(nV, -1);
    h->pos.resize(3 * (size_t)nV); h->orient.resize(4 * (size_t)nV); h->linmom.resize(3 * (size_t)nV); h->angmom.resize(3 * (size_t)nV);
    h->phase.resize(nV); h->bcil.resize(3 * (size_t)nV); h->scil.resize(3 * (size_t)nV); h->temp.resize(nV);
    h->exts.clear();
    for (int i = 0; i < nV; i++) {
        CVX_Voxel *p = Vx.voxelsList[i];
        h->ix[i] = p->ix; h->iy[i] = p->iy; h->iz[i] = p->iz;
        h->vmat[i] = vmIndex[p->mat];
        h->pos[3 * i] = p->pos.x; h->pos[3 * i + 1] = p->pos.y; h->pos[3 * i + 2] = p->pos.z;
        h->orient[4 * i] = p->orient.w; h->orient[4 * i + 1] = p->orient.x; h->orient[4 * i + 2] = p->orient.y; h->orient[4 * i + 3] = p->orient.z;
        h->linmom[3 * i] = p->linMom.x; h->linmom[3 * i + 1] = p->linMom.y; h->linmom[3 * i + 2] = p->linMom.z;
        h->angmom[3 * i] = p->angMom.x; h->angmom[3 * i + 1] = p->angMom.y; h->angmom[3 * i + 2] = p->angMom.z;
        h->vflags[i] = p->boolStates;
        h->temp[i] = p->temp;
        h->phase[i] = p->phaseOffset;
        h->bcil[3 * i] = p->baseCiliaForce.x; h->bcil[3 * i + 1] = p->baseCiliaForce.y; h->bcil[3 * i + 2] = p->baseCiliaForce.z;
        h->scil[3 * i] = p->shiftCiliaForce.x; h->scil[3 * i + 1] = p->shiftCiliaForce.y; h->scil[3 * i + 2] = p->shiftCiliaForce.z;
        for (int k = 0; k < 6; k++)
            if (p->links[k]) h->vlinks[6 * (size_t)i + k] = lIndex[p->links[k]];
        if (p->ext) {
            vx3_external e;
            memset(&e, 0, sizeof(e));
            e.dof_fixed = p->ext->dofFixed;
            e.force[0] = p->ext->extForce.x; e.force[1] = p->ext->extForce.y; e.force[2] = p->ext->extForce.z;
            e.moment[0] = p->ext->extMoment.x; e.moment[1] = p->ext->extMoment.y; e.moment[2] = p->ext->extMoment.z;
            e.translation[0] = p->ext->extTranslation.x; e.translation[1] = p->ext->extTranslation.y; e.translation[2] = p->ext->extTranslation.z;
            e.rotation[0] = p->ext->extRotation.x; e.rotation[1] = p->ext->extRotation.y; e.rotation[2] = p->ext->extRotation.z;
            Quat3D<double> q = p->ext->rotationQuat();
            e.rotation_q[0] = q.w; e.rotation_q[1] = q.x; e.rotation_q[2] = q.y; e.rotation_q[3] = q.z;
            h->vext[i] = (int)h->exts.size();
            h->exts.push_back(e);
        }
    }
    h->lneg.resize(nL); h->lpos.resize(nL); h->laxis.resize(nL); h->lmat.resize(nL); h->lflags.resize(nL); h->lsmall.resize(nL);
    h->lpos2.resize(3 * (size_t)nL); h->la1v.resize(3 * (size_t)nL); h->la2v.resize(3 * (size_t)nL); h->lrest.resize(nL);
    h->lstrain.resize(nL); h->lmaxstrain.resize(nL); h->loffset.resize(nL); h->lstress.resize(nL);
    h->larea.resize(nL); h->ltsum.resize(nL); h->lratio.resize(nL);
    for (int i = 0; i < nL; i++) {
        CVX_Link *p = Vx.linksList[i];
        h->lneg[i] = vIndex[p->pVNeg]; h->lpos[i] = vIndex[p->pVPos];
        h->laxis[i] = (int)p->axis;
        h->lmat[i] = lmIndex[p->mat];
        h->lflags[i] = p->boolStates;
        h->lsmall[i] = p->smallAngle ? 1 : 0;
        h->lpos2[3 * i] = p->pos2.x; h->lpos2[3 * i + 1] = p->pos2.y; h->lpos2[3 * i + 2] = p->pos2.z;
        h->la1v[3 * i] = p->angle1v.x; h->la1v[3 * i + 1] = p->angle1v.y; h->la1v[3 * i + 2] = p->angle1v.z;
        h->la2v[3 * i] = p->angle2v.x; h->la2v[3 * i + 1] = p->angle2v.y; h->la2v[3 * i + 2] = p->angle2v.z;
        h->lrest[i] = p->currentRestLength;
        h->lstrain[i] = p->strain; h->lmaxstrain[i] = p->maxStrain; h->loffset[i] = p->strainOffset; h->lstress[i] = p->_stress;
        h->larea[i] = p->currentTransverseArea; h->ltsum[i] = p->currentTransverseStrainSum; h->lratio[i] = p->strainRatio;
    }

    strncpy(d.name, "ref", sizeof(d.name) - 1);
    d.n_voxel_mats = (int)h->vmats.size(); d.n_link_mats = (int)h->lmats.size();
    d.voxel_mats = h->vmats.data(); d.link_mats = h->lmats.data();
    d.n_voxels = nV; d.n_links = nL; d.n_externals = (int)h->exts.size();
    d.ix = h->ix.data(); d.iy = h->iy.data(); d.iz = h->iz.data(); d.vox_mat = h->vmat.data();
    d.pos = h->pos.data(); d.orient = h->orient.data(); d.lin_mom = h->linmom.data(); d.ang_mom = h->angmom.data();
    d.vox_flags = h->vflags.data(); d.temp = h->temp.data(); d.phase_offset = h->phase.data();
    d.vox_links = h->vlinks.data(); d.vox_ext = h->vext.data(); d.base_cilia = h->bcil.data(); d.shift_cilia = h->scil.data();
    d.externals = h->exts.data();
    d.link_vneg = h->lneg.data(); d.link_vpos = h->lpos.data(); d.link_axis = h->laxis.data(); d.link_mat = h->lmat.data();
    d.link_pos2 = h->lpos2.data(); d.link_angle1v = h->la1v.data(); d.link_angle2v = h->la2v.data();
    d.link_strain = h->lstrain.data(); d.link_max_strain = h->lmaxstrain.data(); d.link_strain_offset = h->loffset.data();
    d.link_stress = h->lstress.data(); d.link_flags = h->lflags.data(); d.link_small_angle = h->lsmall.data();
    d.link_rest_length = h->lrest.data(); d.link_transverse_area = h->larea.data();
    d.link_transverse_strain_sum = h->ltsum.data(); d.link_strain_ratio = h->lratio.data();

    // VX3_VoxelyzeKernel ctor, src/VX3/VX3_VoxelyzeKernel.cu:29,93-101
    memset(&d.opt, 0, sizeof(d.opt));
    d.opt.vox_size = Vx.voxSize;
    d.opt.dt_frac = h->sim->DtFrac;
    d.opt.temp_enabled = h->env->IsTempEnabled();
    d.opt.vary_temp_enabled = h->env->IsTempVaryEnabled();
    d.opt.temp_base = h->env->GetTempBase();
    d.opt.temp_amplitude = h->env->GetTempAmplitude();
    d.opt.temp_period = h->env->GetTempPeriod();
    // readVXD defaults, src/VX3/VX3_SimulationManager.cu:328-367 (the shim does not parse VX3 tags)
    d.opt.enable_collision = 1;
    d.opt.watch_distance = 1.0;
    d.opt.bounding_radius = 0.75;
    d.opt.safety_guard = 500;
    d.opt.record_voxel = 1;
    return &d;
}

double vxref_recommended_dt(vxref *h) { return h->sim->Vx.recommendedTimeStep(); }

long vxref_step(vxref *h, long k, float dt) {
    long i = 0;
    for (; i < k; i++)
        if (!h->sim->Vx.doTimeStep(dt)) break;
    return i;
}

int vxref_state(vxref *h, vx3_state_view *w) {
    CVoxelyze &Vx = h->sim->Vx;
    int nv = Vx.voxelCount(), nl = Vx.linkCount();
    if (w->n_voxels < nv || w->n_links < nl) return -1;
    w->n_voxels = nv;
    w->n_links = nl;
    for (int i = 0; i < nv; i++) {
        CVX_Voxel *p = Vx.voxelsList[i];
        if (w->pos) { w->pos[3 * i] = p->pos.x; w->pos[3 * i + 1] = p->pos.y; w->pos[3 * i + 2] = p->pos.z; }
        if (w->orient) { w->orient[4 * i] = p->orient.w; w->orient[4 * i + 1] = p->orient.x; w->orient[4 * i + 2] = p->orient.y; w->orient[4 * i + 3] = p->orient.z; }
        if (w->lin_mom) { w->lin_mom[3 * i] = p->linMom.x; w->lin_mom[3 * i + 1] = p->linMom.y; w->lin_mom[3 * i + 2] = p->linMom.z; }
        if (w->ang_mom) { w->ang_mom[3 * i] = p->angMom.x; w->ang_mom[3 * i + 1] = p->angMom.y; w->ang_mom[3 * i + 2] = p->angMom.z; }
        if (w->vox_flags) w->vox_flags[i] = p->boolStates;
        if (w->temp) w->temp[i] = p->temp;
    }
    for (int i = 0; i < nl; i++) {
        CVX_Link *p = Vx.linksList[i];
        if (w->link_pos2) { w->link_pos2[3 * i] = p->pos2.x; w->link_pos2[3 * i + 1] = p->pos2.y; w->link_pos2[3 * i + 2] = p->pos2.z; }
        if (w->link_angle1v) { w->link_angle1v[3 * i] = p->angle1v.x; w->link_angle1v[3 * i + 1] = p->angle1v.y; w->link_angle1v[3 * i + 2] = p->angle1v.z; }
        if (w->link_angle2v) { w->link_angle2v[3 * i] = p->angle2v.x; w->link_angle2v[3 * i + 1] = p->angle2v.y; w->link_angle2v[3 * i + 2] = p->angle2v.z; }
        if (w->link_force_neg) { w->link_force_neg[3 * i] = p->forceNeg.x; w->link_force_neg[3 * i + 1] = p->forceNeg.y; w->link_force_neg[3 * i + 2] = p->forceNeg.z; }
        if (w->link_force_pos) { w->link_force_pos[3 * i] = p->forcePos.x; w->link_force_pos[3 * i + 1] = p->forcePos.y; w->link_force_pos[3 * i + 2] = p->forcePos.z; }
        if (w->link_moment_neg) { w->link_moment_neg[3 * i] = p->momentNeg.x; w->link_moment_neg[3 * i + 1] = p->momentNeg.y; w->link_moment_neg[3 * i + 2] = p->momentNeg.z; }
        if (w->link_moment_pos) { w->link_moment_pos[3 * i] = p->momentPos.x; w->link_moment_pos[3 * i + 1] = p->momentPos.y; w->link_moment_pos[3 * i + 2] = p->momentPos.z; }
        if (w->link_strain) w->link_strain[i] = p->strain;
        if (w->link_max_strain) w->link_max_strain[i] = p->maxStrain;
        if (w->link_strain_offset) w->link_strain_offset[i] = p->strainOffset;
        if (w->link_stress) w->link_stress[i] = p->_stress;
        if (w->link_flags)
            w->link_flags[i] = ((p->boolStates & LOCAL_VELOCITY_VALID) ? VX3_LINKSTATE_LOCAL_VELOCITY_VALID : 0) |
                               (p->smallAngle ? VX3_LINKSTATE_SMALL_ANGLE : 0);
        if (w->link_rest_length) w->link_rest_length[i] = p->currentRestLength;
    }
    return 0;
}

void vxref_destroy(vxref *h) {
    if (!h) return;
    // objects live in calloc'd storage; the reference leaks by design, so do we (test process only)
    delete h;
}

} // extern "C"
