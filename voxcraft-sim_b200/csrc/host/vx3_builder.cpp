// Host model builder: palette + lattice + environment  ->  flat vx3_model_desc.
//
// Re-states the reference's CVX_Sim::Import sequence (src/VXA/VX_Sim.cpp:59-153) and the
// CVoxelyze voxel/link construction it drives (src/old/Voxelyze.cpp:439-461 addVoxel,
// :507-539 addLink, :626-641 combinedMaterial, :644-670 setVoxelSize) on flat arrays.
#include "../../../include/vx3_model.h"
#include "vx3_materials.h"

#include <cfloat>
#include <cmath>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

using namespace vx3;

static thread_local std::string g_model_err;
extern "C" const char *vx3_model_last_error(void) { return g_model_err.c_str(); }
void vx3_model_set_error(const std::string &msg) { g_model_err = msg; }

struct vx3_builder {
    double latDim = 0.001;
    std::vector<vx3_material_params> palette;
    std::vector<std::vector<double>> paletteStrain, paletteStress;
    vx3_env_params env;
    vx3_sim_options opt;
    std::string name;
    std::vector<vx3_token> prog[VX3_PROG_COUNT];
    bool progSet[VX3_PROG_COUNT] = {};
    int nx = 0, ny = 0, nz = 0;
    std::vector<uint8_t> cells;
    std::vector<double> cellPhase, cellBaseCilia, cellShiftCilia;
    struct Ext {
        int voxel;
        vx3_external e;
    };
    std::vector<Ext> exts;

    // ---- build outputs (own the memory referenced by `desc`) ----
    vx3_model_desc desc;
    std::vector<MaterialInput> vmats; // palette materials as configured (inputs of voxel_constants)
    std::vector<vx3_voxel_material> o_vmats;
    std::vector<vx3_link_material> o_lmats;
    std::vector<std::vector<float>> o_data; // strain/stress arrays kept alive
    std::vector<int16_t> ix, iy, iz;
    std::vector<int32_t> vmat, vflags, vlinks, vext;
    std::vector<double> pos, orient, linmom, angmom, phase, bcil, scil;
    std::vector<float> temp;
    std::vector<vx3_external> o_exts;
    std::vector<int32_t> lneg, lpos, laxis, lmat, lflags, lsmall;
    std::vector<double> lpos2, la1v, la2v, lrest;
    std::vector<float> lstrain, lmaxstrain, loffset, lstress, larea, ltsum, lratio;
};

extern "C" void vx3_material_params_default(vx3_material_params *p) {
    // "tag absent" values of CVXC_Material::ReadXML, src/VXA/VX_Object.cpp:1385-1448
    memset(p, 0, sizeof(*p));
    p->mat_model = 0;
    p->signal_value_decay = 0.9;
    p->signal_time_delay = 0.03;
    p->inactive_period = 0.03;
    p->is_measured = 1;
    p->red = p->green = p->blue = 0.5;
    p->alpha = 1.0;
}

extern "C" void vx3_env_params_default(vx3_env_params *e) {
    // CVX_Environment ctor src/old/VX_Environment.cpp:22-37; damping: zeroed CVX_Sim
    memset(e, 0, sizeof(*e));
    e->grav_enabled = 1;
    e->grav_acc = -9.81;
    e->floor_enabled = 1;
    e->temp_enabled = 1;
    e->temp_base = 25;
    e->temp_amplitude = 0;
    e->vary_temp_enabled = 0;
    e->temp_period = 0.1;
}

extern "C" void vx3_sim_options_default(vx3_sim_options *o) {
    // src/VX3/VX3_SimulationManager.cu:328-367 defaults; DtFrac default src/VXA/VX_Sim.cpp ReadVXA (0.9)
    memset(o, 0, sizeof(*o));
    o->dt_frac = 0.9;
    o->enable_collision = 1;
    o->watch_distance = 1.0;
    o->bounding_radius = 0.75;
    o->safety_guard = 500;
    o->record_voxel = 1;
}

extern "C" vx3_builder *vx3_builder_create(double lattice_dim) {
    auto *b = new vx3_builder();
    b->latDim = lattice_dim;
    vx3_env_params_default(&b->env);
    vx3_sim_options_default(&b->opt);
    b->name = "model";
    return b;
}
extern "C" void vx3_builder_destroy(vx3_builder *b) { delete b; }

extern "C" int vx3_builder_add_material(vx3_builder *b, const vx3_material_params *p) {
    if (!b || !p) return VX3_ERR_INVALID;
    b->palette.push_back(*p);
    b->paletteStrain.emplace_back();
    b->paletteStress.emplace_back();
    if (p->mat_model == 3 && p->n_data > 0 && p->strain_data && p->stress_data) {
        b->paletteStrain.back().assign(p->strain_data, p->strain_data + p->n_data);
        b->paletteStress.back().assign(p->stress_data, p->stress_data + p->n_data);
    }
    return (int)b->palette.size();
}
extern "C" int vx3_builder_set_env(vx3_builder *b, const vx3_env_params *e) {
    if (!b || !e) return VX3_ERR_INVALID;
    b->env = *e;
    return VX3_OK;
}
extern "C" int vx3_builder_set_options(vx3_builder *b, const vx3_sim_options *o) {
    if (!b || !o) return VX3_ERR_INVALID;
    b->opt = *o;
    return VX3_OK;
}
extern "C" int vx3_builder_set_name(vx3_builder *b, const char *name) {
    if (!b || !name) return VX3_ERR_INVALID;
    b->name = name;
    return VX3_OK;
}
extern "C" int vx3_builder_set_program(vx3_builder *b, int slot, const vx3_token *tok, int n) {
    if (!b || slot < 0 || slot >= VX3_PROG_COUNT || n < 0 || n > VX3_MAX_TOKENS || (n && !tok)) return VX3_ERR_INVALID;
    b->prog[slot].assign(tok, tok + n);
    b->progSet[slot] = n > 0;
    return VX3_OK;
}
extern "C" int vx3_builder_set_structure(vx3_builder *b, int nx, int ny, int nz, const uint8_t *mat, const double *phase_offset,
                                         const double *base_cilia, const double *shift_cilia) {
    if (!b || nx <= 0 || ny <= 0 || nz <= 0 || !mat) return VX3_ERR_INVALID;
    if (nx > 32767 || ny > 32767 || nz > 32767) return VX3_ERR_INVALID; // reference indices are short
    size_t n = (size_t)nx * ny * nz;
    b->nx = nx;
    b->ny = ny;
    b->nz = nz;
    b->cells.assign(mat, mat + n);
    b->cellPhase.clear();
    b->cellBaseCilia.clear();
    b->cellShiftCilia.clear();
    if (phase_offset) b->cellPhase.assign(phase_offset, phase_offset + n);
    if (base_cilia) b->cellBaseCilia.assign(base_cilia, base_cilia + 3 * n);
    if (shift_cilia) b->cellShiftCilia.assign(shift_cilia, shift_cilia + 3 * n);
    return VX3_OK;
}
extern "C" int vx3_builder_set_external(vx3_builder *b, int voxel_index, const vx3_external *e) {
    if (!b || !e || voxel_index < 0) return VX3_ERR_INVALID;
    b->exts.push_back({voxel_index, *e});
    return VX3_OK;
}

// One palette entry -> the inputs of voxel_constants().  Same outcome as Vx.addMaterial() defaults (E = 1e6 Pa, rho = 1e3,
// Voxelyze.h:88) overridden by CVX_Sim::CopyMat (src/VXA/VX_Sim.cpp:368-420): a model whose parameters fail validation
// leaves the default law in place, like the reference's setters that return false.
static MaterialInput palette_material(const vx3_material_params &o, const std::vector<double> &sd, const std::vector<double> &ss, int matid,
                                      const vx3_env_params &env, float grav_mult) {
    MaterialInput n;
    StressLaw::hooke(1e6f, -1, &n.law);
    n.grav_mult = grav_mult;
    n.matid = matid;
    n.is_pacemaker = o.is_pacemaker != 0;
    n.pacemaker_period = o.pacemaker_period;
    n.is_electrical_active = o.is_electrical_active != 0;
    n.signal_value_decay = o.signal_value_decay;
    n.signal_time_delay = o.signal_time_delay;
    n.inactive_period = o.inactive_period;
    n.is_measured = o.is_measured;
    n.remove_after_s = o.remove_after_s;
    n.thermal_on_after_s = o.thermal_on_after_s;
    n.cilia_on_after_s = o.cilia_on_after_s;
    n.is_target = o.is_target != 0;
    n.fixed = o.fixed != 0;
    n.sticky = o.sticky != 0;
    n.cilia = o.cilia;
    // GetRedi() = (int)(Red*255), src/VXA/VX_Object.h:419-425
    n.r = clamp_colour((int)(o.red * 255)); n.g = clamp_colour((int)(o.green * 255));
    n.b = clamp_colour((int)(o.blue * 255)); n.a = clamp_colour((int)(o.alpha * 255));
    StressLaw law;
    bool ok = false;
    switch (o.mat_model) {
    case 0: ok = StressLaw::hooke((float)o.elastic_mod, -1, &law); break;
    case 1: ok = StressLaw::hooke((float)o.elastic_mod, (float)o.fail_stress, &law); break;
    case 2: ok = StressLaw::bilinear((float)o.elastic_mod, (float)o.plastic_mod, (float)o.yield_stress, (float)o.fail_stress, &law); break;
    case 3: {
        std::vector<float> eps(sd.begin(), sd.end()), sig(ss.begin(), ss.end());
        ok = !eps.empty() && StressLaw::tabulated((int)eps.size(), eps.data(), sig.data(), &law);
        break;
    }
    }
    if (ok) n.law = law;
    n.nu = clamp_poisson((float)o.poissons_ratio);
    n.rho = clamp_density((float)o.density);
    n.cte = (float)o.cte;
    auto nonneg = [](float v) { return v <= 0 ? 0.0f : v; };
    n.mu_static = nonneg((float)o.u_static);
    n.mu_kinetic = nonneg((float)o.u_dynamic);
    n.zeta_global = nonneg((float)env.slow_damping_z);
    n.zeta_internal = nonneg((float)env.bond_damping_z);
    n.zeta_collision = nonneg((float)env.col_damping_z);
    return n;
}

extern "C" const vx3_model_desc *vx3_builder_build(vx3_builder *b) {
    if (!b) return nullptr;
    if (b->cells.empty()) {
        g_model_err = "no structure set";
        return nullptr;
    }
    const vx3_env_params &env = b->env;
    const int nPal = (int)b->palette.size();
    const double voxSize = b->latDim; // Vx.setVoxelSize(LocalVXC.GetLatDimEnv().x), VX_Sim.cpp:89

    // ---- gravity (VX_Sim.cpp:66-67,485; Voxelyze.cpp setGravity) ----
    float gravArg = (float)env.grav_acc;              // SetGravityAccel(float grav)
    float grav = (float)(-gravArg / 9.80665);         // Vx.setGravity(-grav/9.80665)
    const bool floorOn = env.floor_enabled != 0;

    // ---- materials (VX_Sim.cpp:72-86) then setVoxelSize (:89) ----
    b->vmats.clear();
    for (int i = 0; i < nPal; i++) {
        MaterialInput m = palette_material(b->palette[i], b->paletteStrain[i], b->paletteStress[i], i + 1, env, grav);
        m.nom_size = voxSize <= 0 ? (double)FLT_MIN : voxSize; // Vx.setVoxelSize, VX_MaterialVoxel.cpp:82-87
        // EnableVolumeEffects at the end of Import (VX_Sim.cpp:148,445-458): Poisson's ratio is forced to 0 unless the feature is on
        if (!env.volume_effects_enabled) m.nu = 0.0f;
        b->vmats.push_back(m);
    }
    b->o_data.clear();
    b->o_data.reserve(2 * (size_t)nPal * (nPal + 2) + 4); // the records point into o_data: it must never reallocate
    b->o_vmats.resize(nPal);
    for (int i = 0; i < nPal; i++) voxel_constants(b->vmats[i], &b->o_vmats[i], &b->o_data);

    // ---- voxels + links in lattice scan order (VX_Sim.cpp:92-107; Voxelyze.cpp:439-461) ----
    const int nx = b->nx, ny = b->ny, nz = b->nz;
    const size_t nCells = (size_t)nx * ny * nz;
    std::vector<int32_t> cell2vox(nCells, -1);
    b->ix.clear(); b->iy.clear(); b->iz.clear(); b->vmat.clear(); b->vflags.clear(); b->vlinks.clear(); b->vext.clear();
    b->pos.clear(); b->orient.clear(); b->linmom.clear(); b->angmom.clear(); b->phase.clear(); b->bcil.clear(); b->scil.clear();
    b->temp.clear();
    b->lneg.clear(); b->lpos.clear(); b->laxis.clear(); b->lmat.clear();
    std::vector<std::pair<int, int>> lmatKey; // (vox1Mat, vox2Mat) in creation order

    auto combined = [&](int m1, int m2) -> int { // Voxelyze.cpp:626-641
        for (size_t k = 0; k < lmatKey.size(); k++)
            if ((lmatKey[k].first == m1 && lmatKey[k].second == m2) || (lmatKey[k].first == m2 && lmatKey[k].second == m1)) return (int)k;
        lmatKey.push_back({m1, m2});
        return (int)lmatKey.size() - 1;
    };

    for (size_t c = 0; c < nCells; c++) {
        int matIndex = (int)b->cells[c] - 1;
        if (matIndex < 0) continue;
        if (matIndex >= nPal) {
            g_model_err = "structure references a material that is not in the palette";
            return nullptr;
        }
        int x = (int)(c % nx), y = (int)((c / nx) % ny), z = (int)(c / ((size_t)nx * ny));
        int v = (int)b->ix.size();
        cell2vox[c] = v;
        b->ix.push_back((int16_t)x); b->iy.push_back((int16_t)y); b->iz.push_back((int16_t)z);
        b->vmat.push_back(matIndex);
        b->pos.push_back(x * voxSize); b->pos.push_back(y * voxSize); b->pos.push_back(z * voxSize);
        b->orient.push_back(1); b->orient.push_back(0); b->orient.push_back(0); b->orient.push_back(0);
        for (int k = 0; k < 3; k++) { b->linmom.push_back(0); b->angmom.push_back(0); }
        // CVX_Voxel::reset(): FLOOR_STATIC_FRICTION set (VX_Voxel.cpp:48-57); enableFloor(floor)
        b->vflags.push_back(VX3_VOX_FLOOR_STATIC_FRICTION | (floorOn ? VX3_VOX_FLOOR_ENABLED : 0));
        b->temp.push_back(0.0f);
        // per-voxel layers are consumed in voxel order from the cell arrays (filled cells only)
        b->phase.push_back(b->cellPhase.empty() ? 0.0 : b->cellPhase[c]);
        for (int k = 0; k < 3; k++) {
            b->bcil.push_back(b->cellBaseCilia.empty() ? 0.0 : b->cellBaseCilia[3 * c + k]);
            b->scil.push_back(b->cellShiftCilia.empty() ? 0.0 : b->cellShiftCilia[3 * c + k]);
        }
        for (int k = 0; k < 6; k++) b->vlinks.push_back(-1);
        b->vext.push_back(-1);
        // addLink for directions 0..5: only X-, Y-, Z- neighbours exist yet (earlier in scan order)
        const int dx[3] = {1, 0, 0}, dy[3] = {0, 1, 0}, dz[3] = {0, 0, 1};
        for (int ax = 0; ax < 3; ax++) {
            int qx = x - dx[ax], qy = y - dy[ax], qz = z - dz[ax];
            if (qx < 0 || qy < 0 || qz < 0) continue;
            int nb = cell2vox[(size_t)qx + (size_t)nx * (qy + (size_t)ny * qz)];
            if (nb < 0) continue;
            int li = (int)b->lneg.size();
            // CVX_Link ctor (VX_Link.cpp:21-55): voxel1 = new voxel has index+1 -> reverseOrder: pVNeg = neighbour
            b->lneg.push_back(nb);
            b->lpos.push_back(v);
            b->laxis.push_back(ax);
            b->lmat.push_back(combined(matIndex, b->vmat[nb])); // combinedMaterial(voxel1->material(), voxel2->material())
            b->vlinks[6 * (size_t)v + 2 * ax + 1] = li;          // this voxel: negative direction
            b->vlinks[6 * (size_t)nb + 2 * ax + 0] = li;         // neighbour: positive direction
        }
    }
    const int nV = (int)b->ix.size(), nL = (int)b->lneg.size();
    // updateSurface (VX_Voxel.cpp:375-380): the SURFACE bit means interior
    for (int v = 0; v < nV; v++) {
        bool interior = true;
        for (int k = 0; k < 6; k++) if (b->vlinks[6 * (size_t)v + k] < 0) interior = false;
        // the CPU lib only runs updateSurface from addLinkInfo/removeLinkInfo, so a voxel without links keeps bit clear
        if (interior) b->vflags[v] |= VX3_VOX_SURFACE;
    }

    // ---- externals ----
    b->o_exts.clear();
    for (auto &e : b->exts) {
        if (e.voxel >= nV) {
            g_model_err = "external refers to a voxel index out of range";
            return nullptr;
        }
        if (b->vext[e.voxel] >= 0) b->o_exts[b->vext[e.voxel]] = e.e;
        else {
            b->vext[e.voxel] = (int)b->o_exts.size();
            b->o_exts.push_back(e.e);
        }
    }

    // ---- initial temperature (VX_Sim.cpp:146-147,467-483; VX_Environment.cpp:327-346) ----
    if (env.temp_enabled) {
        double CurTemp;
        float ret;
        if (env.vary_temp_enabled) {
            if (env.temp_period == 0) ret = 0.0f;
            else {
                CurTemp = env.temp_base + env.temp_amplitude * sin(2 * 3.1415926 / env.temp_period * 0.0);
                ret = (float)CurTemp;
            }
        } else {
            CurTemp = env.temp_base + env.temp_amplitude;
            ret = (float)CurTemp;
        }
        float t = (float)(ret - env.temp_base);
        for (int v = 0; v < nV; v++) b->temp[v] = t;
    }

    // ---- link materials: one per distinct material pair, in creation order (Voxelyze.cpp:626-641) ----
    b->o_lmats.resize(lmatKey.size());
    for (size_t k = 0; k < lmatKey.size(); k++)
        link_constants(b->o_vmats[lmatKey[k].first], b->o_vmats[lmatKey[k].second], lmatKey[k].first, lmatKey[k].second, &b->o_lmats[k], &b->o_data);

    // ---- link state = CVX_Link::reset() (VX_Link.cpp:56-70) ----
    b->lpos2.assign(3 * (size_t)nL, 0.0); b->la1v.assign(3 * (size_t)nL, 0.0); b->la2v.assign(3 * (size_t)nL, 0.0);
    b->lstrain.assign(nL, 0.0f); b->lmaxstrain.assign(nL, 0.0f); b->loffset.assign(nL, 0.0f); b->lstress.assign(nL, 0.0f);
    b->lflags.assign(nL, 0); b->lsmall.assign(nL, 1);
    b->lrest.resize(nL); b->larea.resize(nL); b->ltsum.assign(nL, 0.0f); b->lratio.resize(nL);
    for (int l = 0; l < nL; l++) {
        const vx3_voxel_material &mn = b->o_vmats[b->vmat[b->lneg[l]]], &mp = b->o_vmats[b->vmat[b->lpos[l]]];
        int ax = b->laxis[l];
        b->lratio[l] = mp.E / mn.E;
        // baseSize(axis) = mat->size()[axis]*(1+temp*alphaCTE), VX_Voxel.h:91 (bracket is float)
        double bn = (mn.nomSize * mn.extScale[ax]) * (1 + b->temp[b->lneg[l]] * mn.alphaCTE);
        double bp = (mp.nomSize * mp.extScale[ax]) * (1 + b->temp[b->lpos[l]] * mp.alphaCTE);
        b->lrest[l] = 0.5 * (bn + bp);
        float sn = (float)mn.nomSize, sp = (float)mp.nomSize; // transverseArea with zero strain
        b->larea[l] = 0.5f * (sn * sn + sp * sp);
    }

    // ---- export ----

    vx3_model_desc &d = b->desc;
    memset(&d, 0, sizeof(d));
    strncpy(d.name, b->name.c_str(), sizeof(d.name) - 1);
    d.n_voxel_mats = (int)b->o_vmats.size();
    d.n_link_mats = (int)b->o_lmats.size();
    d.voxel_mats = b->o_vmats.data();
    d.link_mats = b->o_lmats.data();
    d.n_voxels = nV;
    d.n_links = nL;
    d.n_externals = (int)b->o_exts.size();
    d.link_capacity = 0;
    d.ix = b->ix.data(); d.iy = b->iy.data(); d.iz = b->iz.data();
    d.vox_mat = b->vmat.data();
    d.pos = b->pos.data(); d.orient = b->orient.data(); d.lin_mom = b->linmom.data(); d.ang_mom = b->angmom.data();
    d.vox_flags = b->vflags.data(); d.temp = b->temp.data(); d.phase_offset = b->phase.data();
    d.vox_links = b->vlinks.data(); d.vox_ext = b->vext.data();
    d.base_cilia = b->bcil.data(); d.shift_cilia = b->scil.data();
    d.externals = b->o_exts.data();
    d.link_vneg = b->lneg.data(); d.link_vpos = b->lpos.data(); d.link_axis = b->laxis.data(); d.link_mat = b->lmat.data();
    d.link_pos2 = b->lpos2.data(); d.link_angle1v = b->la1v.data(); d.link_angle2v = b->la2v.data();
    d.link_strain = b->lstrain.data(); d.link_max_strain = b->lmaxstrain.data();
    d.link_strain_offset = b->loffset.data(); d.link_stress = b->lstress.data();
    d.link_flags = b->lflags.data(); d.link_small_angle = b->lsmall.data();
    d.link_rest_length = b->lrest.data();
    d.link_transverse_area = b->larea.data(); d.link_transverse_strain_sum = b->ltsum.data();
    d.link_strain_ratio = b->lratio.data();

    d.opt = b->opt;
    d.opt.vox_size = voxSize;
    // VX3_VoxelyzeKernel ctor copies these from the environment (VX3_VoxelyzeKernel.cu:97-101)
    d.opt.temp_enabled = env.temp_enabled;
    d.opt.vary_temp_enabled = env.vary_temp_enabled;
    d.opt.temp_base = env.temp_base;
    d.opt.temp_amplitude = env.temp_amplitude;
    d.opt.temp_period = env.temp_period;
    for (int s = 0; s < VX3_PROG_COUNT; s++) {
        d.prog[s].n = (int)b->prog[s].size();
        d.prog[s].tok = b->prog[s].empty() ? nullptr : b->prog[s].data();
    }
    return &d;
}

extern "C" double vx3_model_recommended_dt(const vx3_model_desc *m) {
    // VX3_VoxelyzeKernel::recommendedTimeStep, src/VX3/VX3_VoxelyzeKernel.cu:184-217
    if (!m) return 0.0;
    double MaxFreq2 = 0.0f;
    for (int i = 0; i < m->n_links; i++) {
        const vx3_link_material &lm = m->link_mats[m->link_mat[i]];
        double m1 = m->voxel_mats[m->vox_mat[m->link_vneg[i]]].mass, m2 = m->voxel_mats[m->vox_mat[m->link_vpos[i]]].mass;
        float stiff;
        if (lm.m.nu == 0.0f) stiff = lm.a1; // isXyzIndependent
        else {
            float strain = m->link_strain ? m->link_strain[i] : 0.0f;
            stiff = (float)(lm.m.eHat * m->link_transverse_area[i] / ((strain + 1) * m->link_rest_length[i]));
        }
        double thisMaxFreq2 = stiff / (m1 < m2 ? m1 : m2);
        if (thisMaxFreq2 > MaxFreq2) MaxFreq2 = thisMaxFreq2;
    }
    if (MaxFreq2 <= 0.0f) {
        for (int i = 0; i < m->n_voxels; i++) {
            const vx3_voxel_material &vm = m->voxel_mats[m->vox_mat[i]];
            double thisMaxFreq2 = vm.E * vm.nomSize / vm.mass;
            if (thisMaxFreq2 > MaxFreq2) MaxFreq2 = thisMaxFreq2;
        }
    }
    if (MaxFreq2 <= 0.0f) return 0.0f;
    return 1.0f / (6.283185f * sqrt(MaxFreq2));
}

// recommendedTimeStep() (VX3_VoxelyzeKernel.cu:184-217) where it depends on the state.  For a link whose material has nu != 0 the
// stiffness is eHat * transverse area / ((1 + strain) * rest length) (VX3_Link::axialStiffness, VX3_Link.cu:268-277), and the
// reference evaluates OptimalDt ONCE, in its first doTimeStep(dt < 0), AFTER that step's updateTemperature (:240-247): the rest
// lengths are those at the t = 0 temperatures, which per-voxel phase offsets make non-zero — not the model's.  Same mixed
// precision as the reference; vx3_model_recommended_dt keeps answering for the model as imported.  Equal to it when no link
// material has nu != 0.
extern "C" double vx3_model_first_step_dt(const vx3_model_desc *mp) {
    if (!mp) return 0.0;
    const vx3_model_desc &m = *mp;
    const bool vary = m.opt.vary_temp_enabled && m.opt.temp_period > 0;
    std::vector<float> te((size_t)m.n_voxels);
    for (int i = 0; i < m.n_voxels; i++) {
        const vx3_voxel_material &vm = m.voxel_mats[m.vox_mat[i]];
        te[(size_t)i] = m.temp ? m.temp[i] : 0.0f;
        if (!vary || vm.thermal_on_after_s > 0.0 || vm.fixed) continue; // gpu_update_temperature (:625-650) at currentTime = 0
        double cur = m.opt.temp_amplitude * sin(2 * 3.1415926f * (0.0 / m.opt.temp_period + (m.phase_offset ? m.phase_offset[i] : 0.0)));
        if (!m.opt.enable_expansion && cur > 0) cur = 0;
        te[(size_t)i] = (float)cur;
    }
    double MaxFreq2 = 0.0f;
    for (int i = 0; i < m.n_links; i++) {
        const vx3_link_material &lm = m.link_mats[m.link_mat[i]];
        const int vn = m.link_vneg[i], vp = m.link_vpos[i], axis = m.link_axis[i];
        const vx3_voxel_material &mn = m.voxel_mats[m.vox_mat[vn]], &mp = m.voxel_mats[m.vox_mat[vp]];
        float stiff;
        if (lm.m.nu == 0.0f) stiff = lm.a1;
        else {
            const double rest = 0.5 * ((mn.nomSize * mn.extScale[axis]) * (1 + te[(size_t)vn] * mn.alphaCTE) + (mp.nomSize * mp.extScale[axis]) * (1 + te[(size_t)vp] * mp.alphaCTE));
            const float strain = m.link_strain ? m.link_strain[i] : 0.0f;
            stiff = (float)(lm.m.eHat * m.link_transverse_area[i] / ((strain + 1) * rest));
        }
        const double m1 = mn.mass, m2 = mp.mass;
        const double f2 = stiff / (m1 < m2 ? m1 : m2);
        if (f2 > MaxFreq2) MaxFreq2 = f2;
    }
    if (MaxFreq2 <= 0.0f) return vx3_model_recommended_dt(&m);
    return 1.0f / (6.283185f * sqrt(MaxFreq2));
}
