// See vx3_materials.h.  Precision notes: every `float` local below is a float in the reference as well, and a double
// literal in an expression promotes exactly where it does there; sqrtf / sqrt are spelled out because the reference's
// unqualified sqrt() resolves to the float overload for float arguments (that decides the last bit of a link's nu).
#include "vx3_materials.h"

#include <cfloat>
#include <cmath>
#include <cstring>

namespace vx3 {

namespace {

// the strain a law assigns to "no failure stress given" (src/old/VX_Material.cpp:381-383, :418-419)
const float kOpenEndedLinearStress = 1000000.0f;

void two_point(StressLaw *w, float e1, float s1) {
    w->eps = {0.0f, e1};
    w->sig = {0.0f, s1};
}

// 0.2 % strain-offset rule on the tabulated points (src/old/VX_Material.cpp:453-487).  The reference scans the segments
// that start at points 1 .. size-3 — the last segment is never tested — and falls back to the failure point.
void offset_yield(StressLaw *w) {
    const float slope0 = w->E;
    const float icept0 = (-0.2f / 100 * slope0);
    const int last_start = (int)w->eps.size() - 3;
    for (int i = 1; i <= last_start; i++) {
        const float x1 = w->eps[i], x2 = w->eps[i + 1], y1 = w->sig[i], y2 = w->sig[i + 1];
        const float slope = (y2 - y1) / (x2 - x1);
        const float icept = y1 - slope * x1;
        if (slope0 == slope) continue;
        const float x = (icept - icept0) / (slope0 - slope);
        if (x > x1 && x < x2) {
            const float frac = (x - x1) / (x2 - x1);
            w->yield_sig = y1 + frac * (y2 - y1);
            w->yield_eps = x;
            return;
        }
    }
    w->yield_sig = w->fail_sig;
    w->yield_eps = w->fail_eps;
}

bool past_failure(const StressLaw &w, float strain) { return w.fail_eps != -1.0f && strain > w.fail_eps; }

// index of the segment [i-1, i] a strain falls on, for strains beyond the first segment (the last segment is open-ended)
int segment_of(const std::vector<float> &x, float v) {
    const int n = (int)x.size();
    for (int i = 2; i < n; i++)
        if (v <= x[i] || i == n - 1) return i;
    return -1;
}

float e_hat(float E, float nu) { return E / ((1 - 2 * nu) * (1 + nu)); } // src/old/VX_Material.cpp:544

void put_law(const StressLaw &w, vx3_voxel_material *o, std::vector<std::vector<float>> *pool) {
    o->linear = w.linear;
    o->E = w.E;
    o->sigmaYield = w.yield_sig;
    o->sigmaFail = w.fail_sig;
    o->epsilonYield = w.yield_eps;
    o->epsilonFail = w.fail_eps;
    pool->push_back(w.eps);
    o->strain_data = pool->back().data();
    pool->push_back(w.sig);
    o->stress_data = pool->back().data();
    o->n_data = (int)w.eps.size();
}

// mass / inertia / damping constants of a cube of edge `size` (src/old/VX_MaterialVoxel.cpp:57-79)
void cube_constants(vx3_voxel_material *o) {
    const double size = o->nomSize;
    const double volume = size * size * size;
    o->mass = (float)(volume * o->rho);
    o->momentInertia = (float)(o->mass * size * size / 6.0f);
    o->firstMoment = (float)(o->mass * size / 2.0f);
    if (volume == 0 || o->mass == 0 || o->momentInertia == 0) {
        o->massInverse = o->sqrtMass = o->momentInertiaInverse = o->_2xSqMxExS = o->_2xSqIxExSxSxS = 0.0f;
        return;
    }
    o->massInverse = 1.0f / o->mass;
    o->sqrtMass = sqrtf(o->mass);
    o->momentInertiaInverse = 1.0f / o->momentInertia;
    o->_2xSqMxExS = (float)(2.0f * sqrt(o->mass * o->E * size));
    o->_2xSqIxExSxSxS = (float)(2.0f * sqrt(o->momentInertia * o->E * size * size * size));
}

} // namespace

// ------------------------------------------------------------------ stress-strain laws
bool StressLaw::hooke(float E, float fail_stress, StressLaw *out) { // src/old/VX_Material.cpp:372-401
    if (E <= 0) return false;
    if (fail_stress != -1.0f && fail_stress <= 0) return false;
    const bool open_ended = fail_stress == -1;
    const float end_sig = open_ended ? kOpenEndedLinearStress : fail_stress;
    const float end_eps = end_sig / E;
    StressLaw w;
    two_point(&w, end_eps, end_sig);
    w.linear = true;
    w.E = E;
    w.yield_sig = w.fail_sig = fail_stress;
    w.yield_eps = w.fail_eps = open_ended ? -1 : end_eps;
    *out = w;
    return true;
}

bool StressLaw::bilinear(float E, float Ep, float yield_stress, float fail_stress, StressLaw *out) { // :406-450
    if (E <= 0 || Ep <= 0 || Ep >= E || yield_stress <= 0) return false;
    if (fail_stress != -1.0f && fail_stress <= yield_stress) return false;
    const float yield_eps = yield_stress / E;
    const float end_sig = fail_stress == -1 ? 3 * yield_stress : fail_stress;
    const float icept = yield_stress - Ep * yield_eps; // the plastic line sigma = Ep*eps + icept
    const float end_eps = (end_sig - icept) / Ep;
    StressLaw w;
    w.eps = {0.0f, yield_eps, end_eps};
    w.sig = {0.0f, yield_stress, end_sig};
    w.linear = false;
    w.E = E;
    w.yield_sig = yield_stress;
    w.yield_eps = yield_eps;
    w.fail_sig = fail_stress;
    w.fail_eps = fail_stress == -1.0f ? -1.0f : end_eps;
    *out = w;
    return true;
}

bool StressLaw::tabulated(int n, const float *strain, const float *stress, StressLaw *out) { // :300-367
    if (n > 0 && strain[0] == 0 && stress[0] == 0) { // an explicit origin is implied anyway
        strain++;
        stress++;
        n--;
    }
    if (n <= 0 || strain[0] <= 0 || stress[0] <= 0) return false;
    StressLaw w;
    w.eps.assign(1, 0.0f);
    w.sig.assign(1, 0.0f);
    float prev = 0.0f;
    for (int i = 0; i < n; i++) { // strains strictly ascending (the reference's slope test compares with 0/0 and never fires, :336)
        if (strain[i] <= prev) return false;
        prev = strain[i];
        w.eps.push_back(strain[i]);
        w.sig.push_back(stress[i]);
    }
    w.E = w.sig[1] / w.eps[1];
    w.fail_sig = w.sig.back();
    w.fail_eps = w.eps.back();
    w.linear = n == 1;
    if (n <= 2) {
        w.yield_sig = w.sig[1];
        w.yield_eps = w.eps[1];
    } else
        offset_yield(&w);
    *out = w;
    return true;
}

float StressLaw::tangent(float strain) const { // src/old/VX_Material.cpp:231-243
    if (past_failure(*this, strain)) return 0.0f;
    if (strain <= eps[1] || linear) return E;
    const int i = segment_of(eps, strain);
    return i < 0 ? 0.0f : (sig[i] - sig[i - 1]) / (eps[i] - eps[i - 1]);
}

float StressLaw::strain_for(float stress) const { // CVX_Material::strain(float stress)
    if (stress <= sig[1] || linear) return stress / E;
    const int i = segment_of(sig, stress);
    if (i < 0) return 0.0f;
    const float frac = (stress - sig[i - 1]) / (sig[i] - sig[i - 1]);
    return eps[i - 1] + frac * (eps[i] - eps[i - 1]);
}

StressLaw StressLaw::in_series(const StressLaw &a, const StressLaw &b, float fail_stress) { // src/old/VX_MaterialLink.cpp:72-104
    StressLaw w;
    if (a.linear && b.linear) {
        if (!hooke(2.0f * a.E * b.E / (a.E + b.E), fail_stress, &w)) hooke(1.0f, -1, &w); // (the reference keeps its cleared default, E = 1)
        return w;
    }
    // merge the two break-point lists in ascending strain while both still have points; each merged segment gets the
    // series stiffness of the two tangents just below its upper end
    std::vector<float> ms(1, 0.0f), mt(1, 0.0f);
    size_t ia = 1, ib = 1;
    while (ia < a.eps.size() && ib < b.eps.size()) {
        const float x = b.eps[ib] < a.eps[ia] ? b.eps[ib] : a.eps[ia];
        if (x == a.eps[ia]) ia++;
        if (x == b.eps[ib]) ib++;
        const float ka = a.tangent(x - FLT_EPSILON), kb = b.tangent(x - FLT_EPSILON);
        const float k = 2.0f * ka * kb / (ka + kb);
        const float x0 = ms.back(), y0 = mt.back();
        ms.push_back(x);
        mt.push_back(y0 + k * (x - x0));
    }
    if (!tabulated((int)ms.size(), ms.data(), mt.data(), &w)) hooke(1.0f, -1, &w);
    w.fail_sig = fail_stress;
    w.fail_eps = fail_stress == -1.0f ? -1.0f : w.strain_for(fail_stress);
    return w;
}

// ------------------------------------------------------------------ scalar clamps (src/old/VX_Material.cpp:245-280,489-502)
float clamp_poisson(float nu) {
    if (nu < 0) nu = 0;
    if (nu >= 0.5) nu = 0.5 - FLT_EPSILON * 2;
    return nu;
}
float clamp_density(float rho) { return rho <= 0 ? FLT_MIN : rho; }
int clamp_colour(int c) { return c > 255 ? 255 : (c < 0 ? 0 : c); }

// ------------------------------------------------------------------ records
void voxel_constants(const MaterialInput &m, vx3_voxel_material *o, std::vector<std::vector<float>> *pool) {
    memset(o, 0, sizeof(*o));
    o->matid = m.matid;
    o->fixed = m.fixed; o->sticky = m.sticky; o->is_target = m.is_target; o->is_measured = m.is_measured;
    o->is_pacemaker = m.is_pacemaker; o->is_electrical_active = m.is_electrical_active;
    o->r = m.r; o->g = m.g; o->b = m.b; o->a = m.a;
    put_law(m.law, o, pool);
    o->nu = m.nu; o->rho = m.rho; o->alphaCTE = m.cte; o->muStatic = m.mu_static; o->muKinetic = m.mu_kinetic;
    o->zetaInternal = m.zeta_internal; o->zetaGlobal = m.zeta_global; o->zetaCollision = m.zeta_collision;
    o->eHat = e_hat(o->E, o->nu);
    o->gravMult = m.grav_mult;
    o->nomSize = m.nom_size;
    o->extScale[0] = o->extScale[1] = o->extScale[2] = 1.0;
    cube_constants(o);
    o->cilia = m.cilia;
    o->pacemaker_period = m.pacemaker_period;
    o->signal_value_decay = m.signal_value_decay;
    o->signal_time_delay = m.signal_time_delay;
    o->inactive_period = m.inactive_period;
    o->remove_after_s = m.remove_after_s;
    o->thermal_on_after_s = m.thermal_on_after_s;
    o->cilia_on_after_s = m.cilia_on_after_s;
}

static StressLaw law_of(const vx3_voxel_material &v) {
    StressLaw w;
    if (v.n_data < 2 || !v.strain_data || !v.stress_data) { // a record without its points: Hooke with the record's modulus
        if (!StressLaw::hooke(v.E, v.sigmaFail, &w)) StressLaw::hooke(1.0f, -1, &w);
        return w;
    }
    w.eps.assign(v.strain_data, v.strain_data + v.n_data);
    w.sig.assign(v.stress_data, v.stress_data + v.n_data);
    w.linear = v.linear != 0;
    w.E = v.E;
    w.yield_sig = v.sigmaYield; w.yield_eps = v.epsilonYield;
    w.fail_sig = v.sigmaFail; w.fail_eps = v.epsilonFail;
    return w;
}

void link_constants(const vx3_voxel_material &a, const vx3_voxel_material &b, int ia, int ib, vx3_link_material *out,
                    std::vector<std::vector<float>> *pool) {
    // the blended material (src/old/VX_MaterialLink.cpp:45-118): plain averages for the scalars, springs in series for
    // the law, the weaker failure stress, and a Poisson's ratio chosen so that eHat is the series value too
    MaterialInput m; // identity fields keep the defaults of a material nobody configured (is_measured = 0 below)
    m.is_measured = 0;
    m.nom_size = 0.5 * (a.nomSize + b.nomSize);
    m.r = (int)(0.5 * (a.r + b.r)); m.g = (int)(0.5 * (a.g + b.g)); m.b = (int)(0.5 * (a.b + b.b)); m.a = (int)(0.5 * (a.a + b.a));
    m.rho = 0.5f * (a.rho + b.rho);
    m.cte = 0.5f * (a.alphaCTE + b.alphaCTE);
    m.mu_static = 0.5f * (a.muStatic + b.muStatic);
    m.mu_kinetic = 0.5f * (a.muKinetic + b.muKinetic);
    m.zeta_internal = 0.5f * (a.zetaInternal + b.zetaInternal);
    m.zeta_global = 0.5f * (a.zetaGlobal + b.zetaGlobal);
    m.zeta_collision = 0.5f * (a.zetaCollision + b.zetaCollision);
    const float fa = a.sigmaFail, fb = b.sigmaFail;
    const float weaker = fa == -1.0f ? fb : (fb == -1.0f ? fa : (fa < fb ? fa : fb));
    m.law = StressLaw::in_series(law_of(a), law_of(b), weaker);
    if (a.nu == 0 && b.nu == 0) m.nu = 0;
    else { // eHat = E/((1-2nu)(1+nu))  ->  (nu + 1/4)^2 = (eHat - E)/(2 eHat) + 1/16   (:108-116)
        const float series_ehat = 2 * a.eHat * b.eHat / (a.eHat + b.eHat);
        const float series_e = m.law.E;
        const float sq = (series_ehat - series_e) / (2 * series_ehat) + 0.0625;
        m.nu = sqrtf(sq) - 0.25;
    }
    voxel_constants(m, &out->m, pool);
    out->vox1_mat = ia;
    out->vox2_mat = ib;
    // Euler-Bernoulli beam of square section L x L and length L (:120-141)
    const float E = out->m.E, L = (float)m.nom_size;
    out->a1 = E * L;                                  // EA/L
    out->a2 = E * L * L * L / (12.0f * (1 + m.nu));   // GJ/L
    out->b1 = E * L;                                  // 12EI/L^3
    out->b2 = E * L * L / 2.0f;                       // 6EI/L^2
    out->b3 = E * L * L * L / 6.0f;                   // 2EI/L
    out->sqA1 = sqrtf(out->a1);
    out->sqA2xIp = sqrtf(out->a2 * L * L / 6.0f);
    out->sqB1 = sqrtf(out->b1);
    out->sqB2xFMp = sqrtf(out->b2 * L / 2.0f);
    out->sqB3xIp = sqrtf(out->b3 * L * L / 6.0f);
}

} // namespace vx3
