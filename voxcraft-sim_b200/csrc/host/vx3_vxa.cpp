// VXA / VXD front end: XML text -> vx3_builder (then vx3_builder_build -> vx3_model_desc).
//
// Follows the reference readers tag for tag, including their "tag absent" defaults:
//   VXD merge            src/Utils/ctool.h:49-57 (children of <VXD> carrying replace="VXA.a.b" replace that subtree)
//   <Simulator>          src/VXA/VX_Sim.cpp:185-294 (Integration, Damping, Collisions, Features, StopCondition)
//   <Environment>        src/old/VX_Environment.cpp:101-178 (Boundary_Conditions, Gravity, Thermal)
//   <VXC>                src/VXA/VX_Object.cpp: Lattice :446-500, Palette/Material :1373-1460, Structure :1760-1957
//   VX3 additions        src/VX3/VX3_SimulationManager.cu:321-376 (AttachDetach, RecordHistory, FitnessFunction, ForceField, ...)
//   math trees           ParseMathTree, src/VX3/VX3_SimulationManager.cu:157-275 (reverse BFS token order)
//   boundary conditions  CVX_Sim::Import, src/VXA/VX_Sim.cpp:109-143; CP_Box::IsTouching src/old/VX_FRegion.cpp:420-429
#include <cmath>
#include <cstring>
#include <deque>
#include <string>
#include <vector>

#include "../../../include/vx3_model.h"
#include "vx3_xml.h"

using namespace vx3;

void vx3_model_set_error(const std::string &msg); // vx3_builder.cpp

namespace {

const char *kOpNames[] = {"mtEND", "mtCONST", "mtE", "mtPI", "mtVAR", "mtADD", "mtSUB", "mtMUL", "mtDIV", "mtPOW", "mtSQRT", "mtSIN",
                          "mtCOS", "mtTAN", "mtATAN", "mtLOG", "mtINT", "mtABS", "mtNOT", "mtGREATERTHAN", "mtLESSTHAN", "mtAND", "mtOR",
                          "mtNORMALCDF"};
const char *kVarNames[] = {"x", "y", "z", "hit", "t", "angle", "targetCloseness", "numClosePairs", "num_voxel"};

std::string rtrim(const std::string &s) {
    size_t e = s.size();
    while (e > 0 && strchr(" \t\r\n", s[e - 1])) e--;
    return s.substr(0, e);
}

// ParseMathTree: BFS over the subtree, every visited child pushed on a stack under an mtEND sentinel, then popped.
bool parse_math_tree(const XNode *root, std::vector<vx3_token> &out, std::string &err) {
    out.clear();
    if (!root) return true; // tag absent: empty program (defined behaviour, see vx3_abi.h)
    std::deque<const XNode *> frontier{root};
    std::vector<std::pair<std::string, std::string>> stack{{"mtEND", ""}};
    while (!frontier.empty()) {
        const XNode *t = frontier.front();
        frontier.pop_front();
        for (auto &k : t->kids) {
            stack.emplace_back(rtrim(k->name), rtrim(k->text));
            frontier.push_back(k.get());
        }
    }
    if (stack.size() == 1) return true; // element present but empty: treat like absent
    for (size_t i = stack.size(); i-- > 0;) {
        vx3_token tk;
        memset(&tk, 0, sizeof(tk));
        int op = -1;
        for (int k = 0; k < 24; k++)
            if (stack[i].first == kOpNames[k]) op = k;
        if (op < 0) { err = "math tree: operation <" + stack[i].first + "> is not implemented"; return false; }
        tk.op = op;
        if (op == VX3_OP_VAR) {
            int v = -1;
            std::string name = xml_trim(stack[i].second);
            for (int k = 0; k < 9; k++)
                if (name == kVarNames[k]) v = k;
            if (v < 0) { err = "math tree: no such variable '" + name + "'"; return false; }
            tk.value = v;
        } else if (op == VX3_OP_CONST) {
            char *end = nullptr;
            tk.value = strtod(stack[i].second.c_str(), &end);
            if (end == stack[i].second.c_str()) { err = "math tree: mtCONST with no number"; return false; }
        }
        out.push_back(tk);
    }
    if (out.size() > VX3_MAX_TOKENS) { err = "math tree: token size overflow"; return false; }
    return true;
}

double get_d(const XNode *p, const char *tag, double dflt) {
    double v;
    return xml_get_double(p, tag, &v) ? v : dflt;
}
int get_i(const XNode *p, const char *tag, int dflt) {
    int v;
    return xml_get_int(p, tag, &v) ? v : dflt;
}
bool get_b(const XNode *p, const char *tag, bool dflt) {
    bool v;
    return xml_get_bool(p, tag, &v) ? v : dflt;
}

struct Region { // CVX_FRegion with a CP_Box primitive
    int prim = 0;
    double X = 0, Y = 0, Z = 0, dX = 0, dY = 0, dZ = 0, Radius = 0;
    int dof = 0;
    double force[3] = {0, 0, 0}, torque[3] = {0, 0, 0}, disp[3] = {0, 0, 0}, adisp[3] = {0, 0, 0};
};

std::vector<double> split_csv(const std::string &s) {
    std::vector<double> out;
    size_t pos = 0;
    while (pos <= s.size()) {
        size_t c = s.find(',', pos);
        std::string item = s.substr(pos, c == std::string::npos ? std::string::npos : c - pos);
        out.push_back(atof(item.c_str()));
        if (c == std::string::npos) break;
        pos = c + 1;
    }
    return out;
}

vx3_builder *fail(vx3_builder *b, const std::string &msg) {
    if (b) vx3_builder_destroy(b);
    vx3_model_set_error(msg);
    return nullptr;
}

} // namespace

extern "C" vx3_builder *vx3_vxa_parse(const char *vxa_xml, const char *vxd_xml, const char *name) {
    if (!vxa_xml) return fail(nullptr, "vxa text is NULL");
    std::string err;
    std::unique_ptr<XNode> doc = xml_parse(vxa_xml, &err);
    if (!doc) return fail(nullptr, "VXA: " + err);
    if (!doc->child("VXA")) return fail(nullptr, "VXA: no <VXA> root element");
    if (vxd_xml && *vxd_xml) { // ctool::ptree_merge
        std::unique_ptr<XNode> vxd = xml_parse(vxd_xml, &err);
        if (!vxd) return fail(nullptr, "VXD: " + err);
        const XNode *src = vxd->child("VXD") ? vxd->child("VXD") : vxd.get();
        for (auto &k : src->kids) {
            const std::string *rep = k->attr("replace");
            if (rep && !rep->empty()) doc->put(*rep, k->clone());
        }
    }
    const XNode *vxa = doc->child("VXA");
    const XNode *sim = vxa->child("Simulator");
    const XNode *envx = vxa->child("Environment");
    const XNode *vxc = vxa->child("VXC") ? vxa->child("VXC") : vxa->child("DMF");
    if (!vxc) return fail(nullptr, "VXA: no <VXC> element");

    // ---- lattice ----
    const XNode *lat = vxc->child("Lattice");
    double latDim = get_d(lat, "Lattice_Dim", 0.001);
    latDim *= get_d(lat, "X_Dim_Adj", 1.0);
    vx3_builder *b = vx3_builder_create(latDim);
    if (name) vx3_builder_set_name(b, name);

    // ---- palette ----
    const XNode *pal = vxc->child("Palette");
    std::vector<std::vector<double>> keepStrain, keepStress;
    if (pal) {
        auto mats = pal->children("Material");
        keepStrain.resize(mats.size());
        keepStress.resize(mats.size());
        size_t mi = 0;
        for (const XNode *m : mats) {
            vx3_material_params p;
            vx3_material_params_default(&p);
            const XNode *disp = m->child("Display");
            if (disp) {
                p.red = get_d(disp, "Red", 0.5); p.green = get_d(disp, "Green", 0.5); p.blue = get_d(disp, "Blue", 0.5); p.alpha = get_d(disp, "Alpha", 1.0);
            }
            if (m->child("MatType") && get_i(m, "MatType", 0) != 0) return fail(b, "VXA: only SINGLE materials are supported");
            const XNode *me = m->child("Mechanical");
            if (me) {
                p.mat_model = get_i(me, "MatModel", 0);
                const XNode *ss = me->child("SSData");
                if (ss) {
                    int n = get_i(ss, "NumDataPts", 0);
                    const XNode *sd = ss->child("StrainData"), *st = ss->child("StressData");
                    if (sd) for (const XNode *e : sd->children("Strain")) keepStrain[mi].push_back(atof(e->text.c_str()));
                    if (st) for (const XNode *e : st->children("Stress")) keepStress[mi].push_back(atof(e->text.c_str()));
                    if ((int)keepStrain[mi].size() > n) keepStrain[mi].resize(n);
                    if ((int)keepStress[mi].size() > n) keepStress[mi].resize(n);
                    if (keepStrain[mi].size() == keepStress[mi].size() && !keepStrain[mi].empty()) {
                        p.n_data = (int)keepStrain[mi].size();
                        p.strain_data = keepStrain[mi].data();
                        p.stress_data = keepStress[mi].data();
                    }
                }
                p.is_pacemaker = get_b(me, "isPaceMaker", false);
                p.pacemaker_period = get_d(me, "PaceMakerPeriod", 0);
                p.signal_value_decay = get_d(me, "signalValueDecay", 0.9);
                p.signal_time_delay = get_d(me, "signalTimeDelay", 0.03);
                p.inactive_period = get_d(me, "inactivePeriod", 0.03);
                p.is_measured = get_i(me, "isMeasured", 1);
                p.remove_after_s = get_d(me, "RemoveFromSimulationAfterThisManySeconds", 0.0);
                p.thermal_on_after_s = get_d(me, "TurnOnThermalExpansionAfterThisManySeconds", 0.0);
                p.cilia_on_after_s = get_d(me, "TurnOnCiliaAfterThisManySeconds", 0.0);
                p.is_electrical_active = get_b(me, "isElectricalActive", false);
                p.is_target = get_b(me, "isTarget", false);
                p.fixed = get_i(me, "Fixed", 0);
                p.sticky = get_i(me, "Sticky", 0);
                p.cilia = get_d(me, "Cilia", 0);
                p.elastic_mod = get_d(me, "Elastic_Mod", 0);
                p.plastic_mod = get_d(me, "Plastic_Mod", 0);
                p.yield_stress = get_d(me, "Yield_Stress", 0);
                p.fail_stress = get_d(me, "Fail_Stress", 0);
                p.fail_strain = get_d(me, "Fail_Strain", 0);
                p.density = get_d(me, "Density", 0);
                p.poissons_ratio = get_d(me, "Poissons_Ratio", 0);
                p.cte = get_d(me, "CTE", 0);
                p.material_temp_phase = get_d(me, "MaterialTempPhase", 0);
                p.u_static = get_d(me, "uStatic", 0);
                p.u_dynamic = get_d(me, "uDynamic", 0);
            }
            if (vx3_builder_add_material(b, &p) < 0) return fail(b, "VXA: bad material");
            mi++;
        }
    }

    // ---- structure ----
    const XNode *st = vxc->child("Structure");
    if (!st) return fail(b, "VXA: no <Structure>");
    const std::string *comp = st->attr("Compression");
    if (!comp || *comp != "ASCII_READABLE") return fail(b, "VXA: only Compression=\"ASCII_READABLE\" structures are supported");
    const int nx = get_i(st, "X_Voxels", 1), ny = get_i(st, "Y_Voxels", 1), nz = get_i(st, "Z_Voxels", 1);
    if (nx <= 0 || ny <= 0 || nz <= 0) return fail(b, "VXA: bad structure dimensions");
    const size_t nxy = (size_t)nx * ny, ncell = nxy * nz;
    std::vector<uint8_t> cells(ncell, 0);
    const XNode *data = st->child("Data");
    if (!data) return fail(b, "VXA: no <Data> in <Structure>");
    auto layers = data->children("Layer");
    if ((int)layers.size() < nz) return fail(b, "VXA: Voxel layer data not present or does not match expected size.");
    for (int z = 0; z < nz; z++) {
        const std::string &raw = layers[z]->text;
        if (raw.size() != nxy) return fail(b, "VXA: Voxel layer data not present or does not match expected size.");
        for (size_t k = 0; k < nxy; k++) cells[nxy * z + k] = (uint8_t)(raw[k] - 48);
    }
    std::vector<double> phase, bcil, scil;
    if (const XNode *po = st->child("PhaseOffset")) {
        auto pl = po->children("Layer");
        phase.assign(ncell, 0.0);
        for (int z = 0; z < nz && z < (int)pl.size(); z++) {
            std::vector<double> v = split_csv(pl[z]->text);
            for (size_t k = 0; k < nxy && k < v.size(); k++) phase[nxy * z + k] = v[k];
        }
    }
    auto cilia_layers = [&](const char *tag, std::vector<double> &out) {
        const XNode *c = st->child(tag);
        if (!c) return;
        auto cl = c->children("Layer");
        out.assign(3 * ncell, 0.0);
        for (int z = 0; z < nz && z < (int)cl.size(); z++) {
            std::vector<double> v = split_csv(cl[z]->text);
            for (size_t k = 0; k < nxy; k++)
                if (v.size() > 3 * k + 2) // shorter layers are padded with zeros (VX_Object.cpp:1884-1890)
                    for (int a = 0; a < 3; a++) out[3 * (nxy * z + k) + a] = v[3 * k + a];
        }
    };
    cilia_layers("BaseCiliaForce", bcil);
    cilia_layers("ShiftCiliaForce", scil);
    if (vx3_builder_set_structure(b, nx, ny, nz, cells.data(), phase.empty() ? nullptr : phase.data(), bcil.empty() ? nullptr : bcil.data(),
                                  scil.empty() ? nullptr : scil.data()) != VX3_OK)
        return fail(b, "VXA: bad structure");

    // ---- environment + classic simulator settings ----
    vx3_env_params e;
    vx3_env_params_default(&e);
    std::vector<Region> regions;
    if (envx) {
        if (const XNode *bc = envx->child("Boundary_Conditions")) {
            for (const XNode *fr : bc->children("FRegion")) {
                Region r;
                r.prim = get_i(fr, "PrimType", -1);
                r.X = get_d(fr, "X", 0); r.Y = get_d(fr, "Y", 0); r.Z = get_d(fr, "Z", 0);
                r.dX = get_d(fr, "dX", 0); r.dY = get_d(fr, "dY", 0); r.dZ = get_d(fr, "dZ", 0);
                r.Radius = get_d(fr, "Radius", 0);
                int dof;
                if (xml_get_int(fr, "DofFixed", &dof)) r.dof = dof & 0x3F;
                else r.dof = get_b(fr, "Fixed", false) ? 0x3F : 0;
                r.force[0] = get_d(fr, "ForceX", 0); r.force[1] = get_d(fr, "ForceY", 0); r.force[2] = get_d(fr, "ForceZ", 0);
                r.torque[0] = get_d(fr, "TorqueX", 0); r.torque[1] = get_d(fr, "TorqueY", 0); r.torque[2] = get_d(fr, "TorqueZ", 0);
                r.disp[0] = get_d(fr, "DisplaceX", 0); r.disp[1] = get_d(fr, "DisplaceY", 0); r.disp[2] = get_d(fr, "DisplaceZ", 0);
                r.adisp[0] = get_d(fr, "AngDisplaceX", 0); r.adisp[1] = get_d(fr, "AngDisplaceY", 0); r.adisp[2] = get_d(fr, "AngDisplaceZ", 0);
                if (r.prim != 0) return fail(b, "VXA: only box boundary-condition regions (PrimType 0) are supported");
                regions.push_back(r);
            }
        }
        if (const XNode *g = envx->child("Gravity")) {
            e.grav_enabled = get_b(g, "GravEnabled", false);
            e.grav_acc = get_d(g, "GravAcc", -9.81);
            e.floor_enabled = get_b(g, "FloorEnabled", false);
        }
        if (const XNode *t = envx->child("Thermal")) {
            e.temp_enabled = get_b(t, "TempEnabled", false);
            e.temp_base = get_d(t, "TempBase", 25);
            double amp;
            if (xml_get_double(t, "TempAmplitude", &amp)) e.temp_amplitude = amp;
            else if (xml_get_double(t, "TempAmp", &amp)) e.temp_amplitude = amp - e.temp_base;
            else e.temp_amplitude = 0;
            e.vary_temp_enabled = get_b(t, "VaryTempEnabled", false);
            e.temp_period = get_d(t, "TempPeriod", 0.1);
        }
    }
    vx3_sim_options o;
    vx3_sim_options_default(&o);
    if (sim) {
        if (const XNode *in = sim->child("Integration")) o.dt_frac = get_d(in, "DtFrac", 0.9);
        if (const XNode *d = sim->child("Damping")) { // values pass through a float in the reference (VX_Sim.cpp:198-211)
            e.bond_damping_z = (float)get_d(d, "BondDampingZ", 0.1);
            e.col_damping_z = (float)get_d(d, "ColDampingZ", 1.0);
            e.slow_damping_z = (float)get_d(d, "SlowDampingZ", 1.0);
        }
        if (const XNode *c = sim->child("Collisions")) e.self_col_enabled = get_b(c, "SelfColEnabled", false);
        if (const XNode *f = sim->child("Features")) e.volume_effects_enabled = get_b(f, "VolumeEffectsEnabled", false);
    }
    vx3_builder_set_env(b, &e);

    // ---- VX3 additions (readVXD) ----
    const XNode *ad = sim ? sim->child("AttachDetach") : nullptr;
    o.enable_collision = get_b(ad, "EnableCollision", true);
    o.enable_attach = get_b(ad, "EnableAttach", false);
    o.enable_detach = get_b(ad, "EnableDetach", false);
    o.watch_distance = get_d(ad, "watchDistance", 1.0);
    o.bounding_radius = get_d(ad, "boundingRadius", 0.75);
    o.safety_guard = get_i(ad, "SafetyGuard", 500);
    const XNode *rh = sim ? sim->child("RecordHistory") : nullptr;
    o.record_step_size = get_i(rh, "RecordStepSize", 0);
    o.record_link = get_i(rh, "RecordLink", 0);
    o.record_voxel = get_i(rh, "RecordVoxel", 1);
    o.save_position_of_all_voxels = get_i(sim, "SavePositionOfAllVoxels", 0);
    o.max_dist_in_voxel_lengths_to_count_as_pair = get_d(sim, "MaxDistInVoxelLengthsToCountAsPair", 0);
    o.enable_cilia = get_i(sim, "EnableCilia", 0);
    o.enable_signals = get_i(sim, "EnableSignals", 0);
    o.secondary_experiment = get_i(sim, "SecondaryExperiment", 0);
    o.reinit_initial_position_after_s = get_d(sim, "ReinitializeInitialPositionAfterThisManySeconds", 0.0);
    o.enable_expansion = get_i(sim, "EnableExpansion", 0);
    vx3_builder_set_options(b, &o);

    struct { int slot; const char *path; } progs[] = {
        {VX3_PROG_STOP, "StopCondition.StopConditionFormula"}, {VX3_PROG_FITNESS, "FitnessFunction"},
        {VX3_PROG_FORCE_X, "ForceField.x_forcefield"}, {VX3_PROG_FORCE_Y, "ForceField.y_forcefield"}, {VX3_PROG_FORCE_Z, "ForceField.z_forcefield"},
        {VX3_PROG_ATTACH_0, "AttachDetach.AttachCondition.Condition_0"}, {VX3_PROG_ATTACH_1, "AttachDetach.AttachCondition.Condition_1"},
        {VX3_PROG_ATTACH_2, "AttachDetach.AttachCondition.Condition_2"}, {VX3_PROG_ATTACH_3, "AttachDetach.AttachCondition.Condition_3"},
        {VX3_PROG_ATTACH_4, "AttachDetach.AttachCondition.Condition_4"}};
    for (auto &pg : progs) {
        std::vector<vx3_token> tok;
        if (!parse_math_tree(sim ? sim->path(pg.path) : nullptr, tok, err)) return fail(b, "VXA: " + err);
        if (!tok.empty()) vx3_builder_set_program(b, pg.slot, tok.data(), (int)tok.size());
    }

    // ---- boundary conditions -> per-voxel externals (CVX_Sim::Import :109-143) ----
    if (!regions.empty()) {
        const double ws[3] = {latDim * nx, latDim * ny, latDim * nz}; // GetWorkSpace() for a rectangular lattice
        const double half[3] = {latDim / 2.0, latDim / 2.0, latDim / 2.0};
        auto touching = [&](const Region &r, const double p[3]) { // CP_Box::IsTouching(P, Dist, Envelope)
            const double ps[3] = {p[0] / ws[0], p[1] / ws[1], p[2] / ws[2]};
            const double ds[3] = {half[0] / ws[0], half[1] / ws[1], half[2] / ws[2]};
            return ps[0] + ds[0] > r.X && ps[0] - ds[0] < r.X + r.dX && ps[1] + ds[1] > r.Y && ps[1] - ds[1] < r.Y + r.dY &&
                   ps[2] + ds[2] > r.Z && ps[2] - ds[2] < r.Z + r.dZ;
        };
        std::vector<int> sizes(regions.size(), 0); // GetNumTouching
        for (size_t c = 0; c < ncell; c++) {
            if (!cells[c]) continue;
            const int x = (int)(c % nx), y = (int)((c / nx) % ny), z = (int)(c / nxy);
            const double p[3] = {(x + 0.5) * latDim, (y + 0.5) * latDim, (z + 0.5) * latDim};
            for (size_t j = 0; j < regions.size(); j++)
                if (touching(regions[j], p)) sizes[j]++;
        }
        int vox = 0;
        for (size_t c = 0; c < ncell; c++) {
            if (!cells[c]) continue;
            const int x = (int)(c % nx), y = (int)((c / nx) % ny), z = (int)(c / nxy);
            const double p[3] = {x * latDim + latDim / 2, y * latDim + latDim / 2, z * latDim + latDim / 2};
            vx3_external ex;
            memset(&ex, 0, sizeof(ex));
            ex.rotation_q[0] = 1.0;
            bool any = false;
            for (size_t j = 0; j < regions.size(); j++) {
                const Region &r = regions[j];
                if (!touching(r, p)) continue;
                any = true;
                for (int a = 0; a < 3; a++) {
                    if (r.dof & (1 << a)) { ex.dof_fixed |= (1 << a); ex.translation[a] = r.disp[a]; }
                    if (r.dof & (8 << a)) { ex.dof_fixed |= (8 << a); ex.rotation[a] = r.adisp[a]; }
                }
                // addForce((Vec3D<float>)(force() + Force/Sizes[j])): adds the running sum again, as the reference does
                for (int a = 0; a < 3; a++) {
                    ex.force[a] += (float)((double)ex.force[a] + r.force[a] * (1.0 / sizes[j]));
                    ex.moment[a] += (float)((double)ex.moment[a] + r.torque[a] * (1.0 / sizes[j]));
                }
            }
            if (any) {
                if (ex.rotation[0] != 0 || ex.rotation[1] != 0 || ex.rotation[2] != 0) { // Quat3D(rotation vector)
                    const double tx = ex.rotation[0] * 0.5, ty = ex.rotation[1] * 0.5, tz = ex.rotation[2] * 0.5;
                    const double m2 = tx * tx + ty * ty + tz * tz;
                    double w, s;
                    if (m2 * m2 < 5.328e-15) { w = 1.0 - 0.5 * m2; s = 1.0 - m2 / 6.0; }
                    else { const double m = sqrt(m2); w = cos(m); s = sin(m) / m; }
                    ex.rotation_q[0] = w; ex.rotation_q[1] = tx * s; ex.rotation_q[2] = ty * s; ex.rotation_q[3] = tz * s;
                }
                vx3_builder_set_external(b, vox, &ex);
            }
            vox++;
        }
    }
    return b;
}

extern "C" vx3_builder *vx3_vxa_load(const char *vxa_path, const char *vxd_path) {
    if (!vxa_path) return fail(nullptr, "vxa path is NULL");
    std::string vxa, vxd;
    if (!xml_read_file(vxa_path, &vxa)) return fail(nullptr, std::string("cannot read ") + vxa_path);
    if (vxd_path && *vxd_path && !xml_read_file(vxd_path, &vxd)) return fail(nullptr, std::string("cannot read ") + vxd_path);
    // vxa_filename = file.filename() of the VXD (VX3_SimulationManager.cu:321), or of the VXA when there is none
    std::string name = (vxd_path && *vxd_path) ? vxd_path : vxa_path;
    size_t slash = name.find_last_of('/');
    if (slash != std::string::npos) name = name.substr(slash + 1);
    return vx3_vxa_parse(vxa.c_str(), vxd.empty() ? nullptr : vxd.c_str(), name.c_str());
}
