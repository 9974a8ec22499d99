// ============================================================================
// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// CPU restatement (single-threaded C++) of the reference's VX3 step loop, used only
// as the parity checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
// leg.  Nothing under voxcraft-sim_b200/ links, imports or calls this file.
//
// It follows the reference's device code file by file (each function cites the
// file:line it restates), on the same flat vx3_model_desc the product consumes:
//   step driver      src/VX3/VX3_VoxelyzeKernel.cu:237-359, VX3_SimulationManager.cu:11-121
//   temperature      src/VX3/VX3_VoxelyzeKernel.cu:219-235,625-650
//   link forces      src/VX3/VX3_Link.cu:58-260
//   voxel integrate  src/VX3/VX3_Voxel.cu:162-277,350-426
//   materials        src/VX3/VX3_Material.cu:90-124, VX3_MaterialLink.cu:53-149
//   collisions       src/VX3/VX3_VoxelyzeKernel.cu:651-843, VX3_Collision.cu:3-31
//   detach           src/VX3/VX3_VoxelyzeKernel.cu:946-968
//   math tree        src/Utils/VX3_MathTree.h:50-192
//   vector / quat    src/Utils/VX3_Vec3D.h, src/Utils/VX3_Quat3D.h
//
// PARITY PINNING: the reference ships no golden vectors for this path.
// This restatement is pinned against the reference's own CPU implementation
// (src/old, compiled unmodified into oracle/_ref by oracle/Makefile): with
// cpu_lib_mode=1 it reproduces CVoxelyze::doTimeStep bit-for-bit on the shared
// feature subset (tests/test_oracle_vs_ref.py, tests/golden/*.json).  The VX3-only
// behaviours (listed next to `cpu_lib_mode` below) have no compilable reference here
// (the VX3 CUDA path does not build on CUDA 12) — for those, parity is pinned only by
// analytic known-answer tests: "parity unpinned by reference" applies to them.
//
// Where the reference is racy (contact-force accumulation, attach slot claims) the
// canonical sequential order of SURVEY.md Appendix A.7 is used.
// Build: g++ -O2 -ffp-contract=off  (no FMA contraction: matches the reference x86-64 build)
// ============================================================================
#include "../include/vx3_abi.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

// ------------------------------------------------------------------ vec / quat
struct V3 {
    double x = 0, y = 0, z = 0;
    V3() {}
    V3(double a, double b, double c) : x(a), y(b), z(c) {}
    V3 operator+(const V3 &v) const { return V3(x + v.x, y + v.y, z + v.z); }
    V3 operator-(const V3 &v) const { return V3(x - v.x, y - v.y, z - v.z); }
    V3 operator-() const { return V3(-x, -y, -z); }
    V3 operator*(double f) const { return V3(f * x, f * y, f * z); } // VX3_Vec3D.h:57
    V3 operator/(double f) const { // VX3_Vec3D.h:59: multiply by the reciprocal
        double Inv = 1.0 / f;
        return V3(Inv * x, Inv * y, Inv * z);
    }
    V3 &operator+=(const V3 &v) { x += v.x; y += v.y; z += v.z; return *this; }
    V3 &operator-=(const V3 &v) { x -= v.x; y -= v.y; z -= v.z; return *this; }
    bool operator==(const V3 &v) const { return x == v.x && y == v.y && z == v.z; }
    double Dot(const V3 &v) const { return x * v.x + y * v.y + z * v.z; }
    double Length2() const { return x * x + y * y + z * z; }
    double Length() const { return sqrt(x * x + y * y + z * z); }
    double Dist2(const V3 &v) const { return (v.x - x) * (v.x - x) + (v.y - y) * (v.y - y) + (v.z - z) * (v.z - z); }
    double Dist(const V3 &v) const { return sqrt(Dist2(v)); }
    V3 Normalized() const { // VX3_Vec3D.h:94
        double l = sqrt(x * x + y * y + z * z);
        return l > 0 ? (*this) / l : (*this);
    }
    void NormalizeFast() { // VX3_Vec3D.h:83
        double l = sqrt(x * x + y * y + z * z);
        if (l > 0) {
            double li = 1.0 / l;
            x *= li; y *= li; z *= li;
        }
    }
    V3 Abs() const { return V3(x >= 0 ? x : -x, y >= 0 ? y : -y, z >= 0 ? z : -z); }
};
inline V3 operator*(double f, const V3 &v) { return v * f; }

const double Q_PI = 3.14159265358979;
const double DBL_EPSILONx24 = 5.328e-15;
const double DISCARD_ANGLE_RAD = 1e-7;
const double SMALL_ANGLE_RAD = 1.732e-2;
const double SLTHRESH_ACOS2SQRT = 2.4e-3;

// ---- libm error model (test hook, vx3o_set_libm_jitter): the GPU's libdevice sin / cos / acos are within 1 ulp of the
// correctly rounded result, glibc's too, but not identically rounded.  With a non-zero seed every sin / cos / acos result
// of the physics path (and every transcendental of a math-tree program) is moved by -1, 0 or +1 ulp at (seeded) random: running a few such replicas next to the exact one
// gives the envelope inside which ANY 1-ulp libm's trajectory must lie — the yardstick of the GPU tolerance gates.
static unsigned long long g_libm_jitter = 0; // 0 = off
static int g_libm_jitter_mode = 0;           // +1 / -1: every result moved one ulp up / down; 0: random per call
static inline double jitter_ulp(double v) {
    if (!g_libm_jitter) return v;
    g_libm_jitter += 0x9E3779B97F4A7C15ull;
    unsigned long long z = g_libm_jitter;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    // seeds 1 and 2 model a libm that is CONSISTENTLY off in one direction (always +1 / always -1 ulp): a systematic bias adds up
    // linearly over the steps where random jitter adds up like a random walk
    if (g_libm_jitter_mode) return nextafter(v, g_libm_jitter_mode > 0 ? INFINITY : -INFINITY);
    const unsigned r = (unsigned)(z >> 33) % 3u;
    return r == 0 ? v : nextafter(v, r == 1 ? INFINITY : -INFINITY);
}
static inline float jitter_ulp_f(float v) { // the same model for a float result (powf of the poissons strain): CUDA documents 4 ulp for powf,
                                            // glibc rounds correctly; the model moves the result by two float ulps
    if (!g_libm_jitter) return v;
    const double d = jitter_ulp((double)v); // draws the direction: one DOUBLE ulp up / down / unchanged ...
    if (d == (double)v) return v;
    const float dir = d > (double)v ? INFINITY : -INFINITY;
    return nextafterf(nextafterf(v, dir), dir); // ... applied as two FLOAT ulps
}
static inline double m_sin(double x) { return jitter_ulp(sin(x)); }
static inline double m_cos(double x) { return jitter_ulp(cos(x)); }
static inline double m_acos(double x) { return jitter_ulp(acos(x)); }

struct Q4 {
    double w = 1, x = 0, y = 0, z = 0;
    Q4() {}
    Q4(double a, double b, double c, double d) : w(a), x(b), y(c), z(d) {}
    explicit Q4(const V3 &rv) { FromRotationVector(rv); }
    Q4 operator*(const Q4 &f) const { // VX3_Quat3D.h:196-201
        return Q4(w * f.w - x * f.x - y * f.y - z * f.z, w * f.x + x * f.w + y * f.z - z * f.y,
                  w * f.y - x * f.z + y * f.w + z * f.x, w * f.z + x * f.y - y * f.x + z * f.w);
    }
    Q4 Conjugate() const { return Q4(w, -x, -y, -z); }
    double Angle() const { return 2.0 * m_acos(w > 1 ? 1 : w); }
    double AngleDegrees() const { return Angle() * 57.29577951308232; }
    V3 ToRotationVector() const { // VX3_Quat3D.h:344-359
        if (w >= 1.0 || w <= -1.0) return V3(0, 0, 0);
        double squareLength = 1.0 - w * w;
        if (squareLength < SLTHRESH_ACOS2SQRT) return V3(x, y, z) * 2.0 * sqrt((2 - 2 * w) / squareLength);
        else return V3(x, y, z) * 2.0 * m_acos(w) / sqrt(squareLength);
    }
    void FromRotationVector(const V3 &VecIn) { // VX3_Quat3D.h:361-377
        V3 theta = VecIn / 2;
        double s, thetaMag2 = theta.Length2();
        if (thetaMag2 * thetaMag2 < DBL_EPSILONx24) {
            w = 1.0 - 0.5 * thetaMag2;
            s = 1.0 - thetaMag2 / 6.0;
        } else {
            double thetaMag = sqrt(thetaMag2);
            w = m_cos(thetaMag);
            s = m_sin(thetaMag) / thetaMag;
        }
        x = theta.x * s;
        y = theta.y * s;
        z = theta.z * s;
    }
    void FromAngleToPosX(const V3 &RotateFrom) { // VX3_Quat3D.h:384-432
        if (V3(0, 0, 0) == RotateFrom) return;
        double YoverX = RotateFrom.y / RotateFrom.x;
        double ZoverX = RotateFrom.z / RotateFrom.x;
        if (YoverX < SMALL_ANGLE_RAD && YoverX > -SMALL_ANGLE_RAD && ZoverX < SMALL_ANGLE_RAD && ZoverX > -SMALL_ANGLE_RAD) {
            x = 0;
            y = 0.5 * ZoverX;
            z = -0.5 * YoverX;
            w = 1 + 0.5 * (-y * y - z * z);
            return;
        }
        V3 RotFromNorm = RotateFrom;
        RotFromNorm.NormalizeFast();
        double theta = m_acos(RotFromNorm.x);
        if (theta > Q_PI - DISCARD_ANGLE_RAD) {
            w = 0; x = 0; y = 1; z = 0;
            return;
        }
        const double AxisMagInv = 1.0 / sqrt(RotFromNorm.z * RotFromNorm.z + RotFromNorm.y * RotFromNorm.y);
        const double a = 0.5 * theta;
        const double s = m_sin(a);
        w = m_cos(a);
        x = 0;
        y = RotFromNorm.z * AxisMagInv * s;
        z = -RotFromNorm.y * AxisMagInv * s;
    }
    V3 RotateVec3D(const V3 &f) const { // VX3_Quat3D.h:434-443
        double fx = f.x, fy = f.y, fz = f.z;
        double tw = fx * x + fy * y + fz * z;
        double tx = fx * w - fy * z + fz * y;
        double ty = fx * z + fy * w - fz * x;
        double tz = -fx * y + fy * x + fz * w;
        return V3(w * tx + x * tw + y * tz - z * ty, w * ty - x * tz + y * tw + z * tx, w * tz + x * ty - y * tx + z * tw);
    }
    V3 RotateVec3DInv(const V3 &f) const { // VX3_Quat3D.h:459-469
        double fx = f.x, fy = f.y, fz = f.z;
        double tw = x * fx + y * fy + z * fz;
        double tx = w * fx - y * fz + z * fy;
        double ty = w * fy + x * fz - z * fx;
        double tz = w * fz - x * fy + y * fx;
        return V3(tw * x + tx * w + ty * z - tz * y, tw * y - tx * z + ty * w + tz * x, tw * z + tx * y - ty * x + tz * w);
    }
};

// ------------------------------------------------------------------ math tree
// VX3_MathTree::eval, src/Utils/VX3_MathTree.h:50-192
double mt_eval(const std::vector<vx3_token> &buff, double x, double y, double z, double hit, double t, double angle,
               double closeness, int numClosePairs, int num_voxel, bool *ok = nullptr) {
    double values[1024];
    int values_cursor = 0, process_cursor = 0;
    if (ok) *ok = true;
    for (int i = 0; i < 1024 && i < (int)buff.size(); i++) {
        const double v = buff[i].value;
        double &out = values[values_cursor];
        double *p = &values[process_cursor];
        switch (buff[i].op) {
        case VX3_OP_END: return values[process_cursor];
        case VX3_OP_CONST: out = v; break;
        case VX3_OP_E: out = 2.71828182845904523536; break;
        case VX3_OP_PI: out = 3.14159265358979323846; break;
        case VX3_OP_VAR:
            if (v < 0.5) out = x;
            else if (v < 1.5) out = y;
            else if (v < 2.5) out = z;
            else if (v < 3.5) out = hit;
            else if (v < 4.5) out = t;
            else if (v < 5.5) out = angle;
            else if (v < 6.5) out = closeness;
            else if (v < 7.5) out = numClosePairs;
            else if (v < 8.5) out = num_voxel;
            break;
        case VX3_OP_SIN: out = jitter_ulp(sin(p[0])); process_cursor++; break;
        case VX3_OP_COS: out = jitter_ulp(cos(p[0])); process_cursor++; break;
        case VX3_OP_TAN: out = jitter_ulp(tan(p[0])); process_cursor++; break;
        case VX3_OP_ATAN: out = jitter_ulp(atan(p[0])); process_cursor++; break;
        case VX3_OP_LOG: out = jitter_ulp(log(p[0])); process_cursor++; break;
        case VX3_OP_INT: out = rint(p[0]); process_cursor++; break;
        case VX3_OP_NORMALCDF: out = jitter_ulp(0.5 * erfc(-p[0] * M_SQRT1_2)); process_cursor++; break; // normcdf
        case VX3_OP_ADD: out = p[1] + p[0]; process_cursor += 2; break;
        case VX3_OP_SUB: out = p[1] - p[0]; process_cursor += 2; break;
        case VX3_OP_MUL: out = p[1] * p[0]; process_cursor += 2; break;
        case VX3_OP_DIV: out = p[1] / p[0]; process_cursor += 2; break;
        case VX3_OP_POW: out = jitter_ulp(pow(p[1], p[0])); process_cursor += 2; break;
        case VX3_OP_SQRT: out = sqrt(p[0]); process_cursor++; break;
        case VX3_OP_ABS: out = fabs(p[0]); process_cursor++; break;
        case VX3_OP_NOT: out = !p[0]; process_cursor++; break;
        case VX3_OP_GREATERTHAN: out = p[1] > p[0]; process_cursor += 2; break;
        case VX3_OP_LESSTHAN: out = p[1] < p[0]; process_cursor += 2; break;
        case VX3_OP_AND: out = p[1] && p[0]; process_cursor += 2; break;
        case VX3_OP_OR: out = p[1] || p[0]; process_cursor += 2; break;
        default:
            if (ok) *ok = false;
            return -1;
        }
        if (process_cursor > values_cursor) {
            if (ok) *ok = false;
            return -1;
        }
        values_cursor++;
    }
    if (ok) *ok = false;
    return -1;
}

// ------------------------------------------------------------------ materials
struct OMat { // VX3_MaterialVoxel / VX3_MaterialLink view
    vx3_voxel_material m;
    std::vector<float> strainData, stressData; // device layout: duplicated leading 0 (VX3_Material.cu:463-477)
    int vox1 = -1, vox2 = -1;
    float a1 = 0, a2 = 0, b1 = 0, b2 = 0, b3 = 0, sqA1 = 0, sqA2xIp = 0, sqB1 = 0, sqB2xFMp = 0, sqB3xIp = 0;
    bool removedFlag = false; // VX3_Material::removed (SecondaryExperiment)

    bool isFailed(float strain) const { return m.epsilonFail != -1.0f && strain > m.epsilonFail; }
    float stress(float strain, float transverseStrainSum = 0.0f, bool forceLinear = false) const { // VX3_Material.cu:90-124
        if (isFailed(strain)) return 0.0f;
        if (strain <= strainData[1] || m.linear || forceLinear) {
            if (m.nu == 0.0f) return m.E * strain;
            else return m.eHat * ((1 - m.nu) * strain + m.nu * transverseStrainSum);
        }
        int DataCount = (int)strainData.size();
        for (int i = 2; i < DataCount; i++) {
            if (strain <= strainData[i] || i == DataCount - 1) {
                float Perc = (strain - strainData[i - 1]) / (strainData[i] - strainData[i - 1]);
                float basicStress = stressData[i - 1] + Perc * (stressData[i] - stressData[i - 1]);
                if (m.nu == 0.0f) return basicStress;
                else {
                    float modulus = (stressData[i] - stressData[i - 1]) / (strainData[i] - strainData[i - 1]);
                    float modulusHat = modulus / ((1 - 2 * m.nu) * (1 + m.nu));
                    float effectiveStrain = basicStress / modulus;
                    float effectiveTransverseStrainSum = transverseStrainSum * (effectiveStrain / strain);
                    return modulusHat * ((1 - m.nu) * effectiveStrain + m.nu * effectiveTransverseStrainSum);
                }
            }
        }
        return 0.0f;
    }
    // VX3_MaterialVoxel.h accessors
    float globalDampingTranslateC() const { return m.zetaGlobal * m._2xSqMxExS; }
    float globalDampingRotateC() const { return m.zetaGlobal * m._2xSqIxExSxSxS; }
    float collisionDampingTranslateC() const { return m.zetaCollision * m._2xSqMxExS; }
    float penetrationStiffness() const { return (float)(2 * m.E * m.nomSize); }
    float gravityForce() const { return -m.mass * 9.80665f * m.gravMult; }
};

void load_mat(const vx3_voxel_material &in, OMat &o, bool cpuLayout) {
    o.m = in;
    o.strainData.clear();
    o.stressData.clear();
    if (!cpuLayout) { // syncVectors pushes a 0 and then the host data which already starts with 0
        o.strainData.push_back(0.0f);
        o.stressData.push_back(0.0f);
    }
    for (int i = 0; i < in.n_data; i++) {
        o.strainData.push_back(in.strain_data[i]);
        o.stressData.push_back(in.stress_data[i]);
    }
    while (o.strainData.size() < 2) { // guard against malformed models
        o.strainData.push_back(0.0f);
        o.stressData.push_back(0.0f);
    }
    o.m.strain_data = nullptr;
    o.m.stress_data = nullptr;
}

// ------------------------------------------------------------------ sim objects
struct OExt {
    vx3_external e;
    bool isFixed(int dof) const { return (e.dof_fixed & dof) != 0; }
    bool isFixedAll() const { return (e.dof_fixed & 0x3F) == 0x3F; }
    bool isFixedAnyRotation() const { return isFixed(VX3_DOF_X_ROTATE) || isFixed(VX3_DOF_Y_ROTATE) || isFixed(VX3_DOF_Z_ROTATE); }
    bool isFixedAllRotation() const { return isFixed(VX3_DOF_X_ROTATE) && isFixed(VX3_DOF_Y_ROTATE) && isFixed(VX3_DOF_Z_ROTATE); }
};

struct OVoxel {
    int mat = 0, ix = 0, iy = 0, iz = 0, ext = -1;
    int links[6] = {-1, -1, -1, -1, -1, -1};
    V3 pos, linMom, angMom;
    Q4 orient;
    int boolStates = 0;
    float tempe = 0, previousDt = 0;
    double phaseOffset = 0;
    V3 contactForce, baseCiliaForce, shiftCiliaForce, CiliaForce;
    bool enableAttach = true;
    bool removed = false;
    // signals (VX3_Voxel.h:304-309): d_signal {value, activeTime}, localSignal(+dt), inactiveUntil, packmakerNextPulse
    double localSignal = 0, localSignaldt = 0, inactiveUntil = 0, packmakerNextPulse = 0;
    double sigValue = 0, sigActiveTime = 0;
    // cached poissonsStrain() and its flag (VX3_Voxel.h:287-288).  VX3 never refreshes a link's transverse info while stepping
    // (VX3_Link.cu:146-150, commented out), so the cache only matters when a link is CREATED: the attach ctor's reset()
    float pStrain[3] = {0, 0, 0};
    bool poissonsStrainInvalid = true;
};

struct OLink {
    int vNeg = 0, vPos = 0, axis = 0, mat = 0;
    V3 forceNeg, forcePos, momentNeg, momentPos;
    float strain = 0, maxStrain = 0, strainOffset = 0;
    int boolStates = 0;
    float strainRatio = 1;
    V3 pos2, angle1v, angle2v;
    Q4 angle1, angle2;
    bool smallAngle = true;
    double currentRestLength = 0;
    float currentTransverseArea = 0, currentTransverseStrainSum = 0;
    float _stress = 0;
    int isNewLink = 0;
    bool isDetached = false, removed = false;
};

const float HYSTERESIS_FACTOR = 1.2f, SA_BOND_BEND_RAD = 0.05f, SA_BOND_EXT_PERC = 0.50f; // src/old/types.h:57-59
const double COLLISION_ENVELOPE_RADIUS = 0.625;                                            // VX3_Collision.h:5

} // namespace

struct vx3o_sim {
    // cpu_lib_mode = 1 switches the VX3-only behaviours OFF so the step equals the reference CPU
    // library's CVoxelyze::doTimeStep (src/old/Voxelyze.cpp:251-284) — used to pin this file:
    //   * dt*massInverse / dt*momentInertiaInverse evaluated in float (VX_Voxel.cpp:163 float dt)
    //   * no angMom=0 under static friction (VX3_Voxel.cu:259-264 is VX3-only)
    //   * no per-step temperature, collisions, attach/detach, force field, CoM sampling
    //   * material data arrays without the duplicated leading 0
    int cpu_lib_mode = 0;
    bool importing = false; // vx3o_create is building the model's own links (resetLink)
    std::string name;
    vx3_sim_options opt;
    std::vector<OMat> vmats, lmats;
    std::vector<OVoxel> vox;
    std::vector<OLink> links;
    std::vector<OExt> exts;
    std::vector<vx3_token> prog[VX3_PROG_COUNT];
    std::vector<int> surface;
    std::vector<V3> initialPosition;
    std::vector<int> targets;

    double currentTime = 0, OptimalDt = 0;
    unsigned long CurStepCount = 0;
    V3 currentCenterOfMass, initialCenterOfMass, comHistory[2];
    int angleSampleTimes = 0;
    double recentAngle = 0, targetCloseness = 0, fitness_score = 0;
    int numClosePairs = 0, collisionCount = 0;
    bool isSurfaceChanged = false, InitialPositionReinitialized = false;
    int status = VX3_SIM_RUNNING;
    float lastDt = 0;
    long attachEvents = 0, detachEvents = 0;

    // ---- helpers mirroring VX3_Voxel.h ----
    const OMat &vm(const OVoxel &v) const { return vmats[v.mat]; }
    double baseSizeAxis(const OVoxel &v, int axis) const { // VX3_Voxel.h:95-98: bracket is float
        const OMat &m = vm(v);
        return (m.m.nomSize * m.m.extScale[axis]) * (1 + v.tempe * m.m.alphaCTE);
    }
    double baseSizeAverage(const OVoxel &v) const { // VX3_Voxel.h:101-104
        const OMat &m = vm(v);
        float br = 1 + v.tempe * m.m.alphaCTE;
        V3 b = V3(m.m.nomSize * m.m.extScale[0], m.m.nomSize * m.m.extScale[1], m.m.nomSize * m.m.extScale[2]) * br;
        return (b.x + b.y + b.z) / 3.0f;
    }
    float floorPenetration(const OVoxel &v) const { // VX3_Voxel.h:185-187
        return (float)(baseSizeAverage(v) / 2 - vm(v).m.nomSize / 2 - v.pos.z);
    }
    V3 velocity(const OVoxel &v) const { return v.linMom * vm(v).m.massInverse; }
    V3 angularVelocity(const OVoxel &v) const { return v.angMom * vm(v).m.momentInertiaInverse; }
    float dampingMultiplier(const OVoxel &v) const { // VX3_Voxel.h:206-208 (all float)
        const OMat &m = vm(v);
        return 2 * m.m.sqrtMass * m.m.zetaInternal / v.previousDt;
    }
    static bool isNegative(int dir) { return dir % 2 == 1; }

    // ---- link ----
    static V3 toAxisX(int axis, const V3 &v) { // VX3_Link.h:135-144
        switch (axis) {
        case 1: return V3(v.y, -v.x, v.z);
        case 2: return V3(v.z, v.y, -v.x);
        default: return v;
        }
    }
    static Q4 toAxisX(int axis, const Q4 &q) { // VX3_Link.h:146-155
        switch (axis) {
        case 1: return Q4(q.w, q.y, -q.x, q.z);
        case 2: return Q4(q.w, q.z, q.y, -q.x);
        default: return q;
        }
    }
    static void toAxisOriginal(int axis, V3 *pV) { // VX3_Link.h:157-174
        switch (axis) {
        case 1: { double tmp = pV->y; pV->y = pV->x; pV->x = -tmp; break; }
        case 2: { double tmp = pV->z; pV->z = pV->x; pV->x = -tmp; break; }
        default: break;
        }
    }
    void updateRestLength(OLink &l) { // VX3_Link.cu:80-83
        l.currentRestLength = 0.5 * (baseSizeAxis(vox[l.vNeg], l.axis) + baseSizeAxis(vox[l.vPos], l.axis));
    }
    float axialStrain(const OLink &l, bool positiveEnd) const { // VX3_Link.cu:72-74
        return positiveEnd ? 2.0f * l.strain * l.strainRatio / (1.0f + l.strainRatio) : 2.0f * l.strain / (1.0f + l.strainRatio);
    }
    void voxelStrain(const OVoxel &v, bool poissons, float out[3]) const { // VX3_Voxel::strain, VX3_Voxel.cu:428-468
        float intStrRet[3] = {0, 0, 0};
        int numBondAxis[3] = {0, 0, 0};
        bool tension[3] = {false, false, false};
        for (int i = 0; i < 6; i++)
            if (v.links[i] >= 0) {
                const int axis = i / 2;
                intStrRet[axis] += axialStrain(links[v.links[i]], isNegative(i));
                numBondAxis[axis]++;
            }
        for (int i = 0; i < 3; i++) {
            if (numBondAxis[i] == 2) intStrRet[i] *= 0.5f;
            if (poissons)
                tension[i] = (numBondAxis[i] == 2) ||
                             (v.ext >= 0 && (numBondAxis[i] == 1 && (exts[v.ext].isFixed(1 << i) || (float)exts[v.ext].e.force[i] != 0)));
        }
        if (poissons && !(tension[0] && tension[1] && tension[2])) {
            float add = 0;
            for (int i = 0; i < 3; i++)
                if (tension[i]) add += intStrRet[i];
            const float value = jitter_ulp_f(powf(1.0f + add, -vm(v).m.nu)) - 1.0f;
            for (int i = 0; i < 3; i++)
                if (!tension[i]) intStrRet[i] = value;
        }
        for (int i = 0; i < 3; i++) out[i] = intStrRet[i];
    }
    const float *poissonsStrain(OVoxel &v) { // VX3_Voxel.cu:470-476
        if (v.poissonsStrainInvalid) {
            voxelStrain(v, true, v.pStrain);
            v.poissonsStrainInvalid = false;
        }
        return v.pStrain;
    }
    float transverseStrainSum(OVoxel &v, int axis) { // VX3_Voxel.cu:478-494
        if (vm(v).m.nu == 0) return 0;
        const float *ps = poissonsStrain(v);
        switch (axis) {
        case 0: return ps[1] + ps[2];
        case 1: return ps[0] + ps[2];
        case 2: return ps[0] + ps[1];
        default: return 0.0f;
        }
    }
    float transverseArea(OVoxel &v, int axis) { // VX3_Voxel.cu:496-513
        float size = (float)vm(v).m.nomSize;
        if (vm(v).m.nu == 0) return size * size;
        const float *p = poissonsStrain(v);
        const double ps[3] = {p[0], p[1], p[2]};
        switch (axis) {
        case 0: return (float)(size * size * (1 + ps[1]) * (1 + ps[2]));
        case 1: return (float)(size * size * (1 + ps[0]) * (1 + ps[2]));
        case 2: return (float)(size * size * (1 + ps[0]) * (1 + ps[1]));
        default: return size * size;
        }
    }
    void resetLink(OLink &l) { // VX3_Link.cu:58-70
        l.pos2 = l.angle1v = l.angle2v = V3();
        l.angle1 = l.angle2 = Q4();
        l.forceNeg = l.forcePos = l.momentNeg = l.momentPos = V3();
        l.strain = l.maxStrain = l.strainOffset = l._stress = 0.0f;
        l.strainRatio = vm(vox[l.vPos]).m.E / vm(vox[l.vNeg]).m.E;
        l.smallAngle = true;
        l.boolStates &= ~VX3_LINK_LOCAL_VELOCITY_VALID;
        updateRestLength(l);
        // updateTransverseInfo (VX3_Link.cu:85-88).  At import every strain is zero (area = size^2, sum = 0, and the end voxels'
        // caches become valid with zeros); for a link the attach phase creates the end voxels' CURRENT poissons strains go in —
        // this link (strain 0) already sits in their slots — and stay frozen for the link's life
        if (importing) { // host import (CVX_Link::reset with every strain zero): the same calls, whose answers are size^2 and 0
            for (int vi : {l.vNeg, l.vPos})
                if (vm(vox[vi]).m.nu != 0) {
                    vox[vi].pStrain[0] = vox[vi].pStrain[1] = vox[vi].pStrain[2] = 0.0f;
                    vox[vi].poissonsStrainInvalid = false;
                }
            const float sn = (float)vm(vox[l.vNeg]).m.nomSize, sp = (float)vm(vox[l.vPos]).m.nomSize;
            l.currentTransverseArea = 0.5f * (sn * sn + sp * sp);
            l.currentTransverseStrainSum = 0.0f;
            return;
        }
        l.currentTransverseArea = 0.5f * (transverseArea(vox[l.vNeg], l.axis) + transverseArea(vox[l.vPos], l.axis));
        l.currentTransverseStrainSum = 0.5f * (transverseStrainSum(vox[l.vNeg], l.axis) + transverseStrainSum(vox[l.vPos], l.axis));
    }

    void orientLink(OLink &l) { // VX3_Link.cu:90-133
        const OVoxel &vp = vox[l.vPos], &vn = vox[l.vNeg];
        V3 _pos2 = vp.pos - vn.pos;
        l.pos2 = toAxisX(l.axis, _pos2);
        l.angle1 = toAxisX(l.axis, vn.orient);
        l.angle2 = toAxisX(l.axis, vp.orient);
        Q4 totalRot = l.angle1.Conjugate();
        l.pos2 = totalRot.RotateVec3D(l.pos2);
        l.angle2 = totalRot * l.angle2;
        l.angle1 = Q4();
        float SmallTurn = (float)((fabs(l.pos2.z) + fabs(l.pos2.y)) / l.pos2.x);
        float ExtendPerc = (float)(fabs(1 - l.pos2.x / l.currentRestLength));
        if (!l.smallAngle && SmallTurn < SA_BOND_BEND_RAD && ExtendPerc < SA_BOND_EXT_PERC) {
            l.smallAngle = true;
            l.boolStates &= ~VX3_LINK_LOCAL_VELOCITY_VALID;
        } else if (l.smallAngle && (SmallTurn > HYSTERESIS_FACTOR * SA_BOND_BEND_RAD || ExtendPerc > HYSTERESIS_FACTOR * SA_BOND_EXT_PERC)) {
            l.smallAngle = false;
            l.boolStates &= ~VX3_LINK_LOCAL_VELOCITY_VALID;
        }
        if (l.smallAngle) {
            l.pos2.x -= l.currentRestLength;
        } else {
            l.angle1.FromAngleToPosX(l.pos2);
            totalRot = l.angle1 * totalRot;
            l.angle2 = l.angle1 * l.angle2;
            l.pos2 = V3(l.pos2.Length() - l.currentRestLength, 0, 0);
        }
        l.angle1v = l.angle1.ToRotationVector();
        l.angle2v = l.angle2.ToRotationVector();
    }

    float updateStrain(OLink &l, float axialStrain) { // VX3_Link.cu:220-260
        const OMat &mat = lmats[l.mat];
        l.strain = axialStrain;
        if (mat.m.linear) {
            if (axialStrain > l.maxStrain) l.maxStrain = axialStrain;
            return mat.stress(axialStrain, l.currentTransverseStrainSum);
        } else {
            float returnStress;
            if (axialStrain > l.maxStrain) {
                l.maxStrain = axialStrain;
                returnStress = mat.stress(axialStrain, l.currentTransverseStrainSum);
                if (mat.m.nu != 0.0f) l.strainOffset = l.maxStrain - mat.stress(axialStrain) / (mat.m.eHat * (1 - mat.m.nu));
                else l.strainOffset = l.maxStrain - returnStress / mat.m.E;
            } else {
                float relativeStrain = axialStrain - l.strainOffset;
                if (mat.m.nu != 0.0f) returnStress = mat.stress(relativeStrain, l.currentTransverseStrainSum, true);
                else returnStress = mat.m.E * relativeStrain;
            }
            return returnStress;
        }
    }

    void updateForces(OLink &l) { // VX3_Link.cu:135-218
        const OMat &mat = lmats[l.mat];
        V3 oldPos2 = l.pos2, oldAngle1v = l.angle1v, oldAngle2v = l.angle2v;
        orientLink(l);
        V3 dPos2 = 0.5 * (l.pos2 - oldPos2);
        V3 dAngle1 = 0.5 * (l.angle1v - oldAngle1v);
        V3 dAngle2 = 0.5 * (l.angle2v - oldAngle2v);
        // updateTransverseInfo() is commented out in VX3 (VX3_Link.cu:147-150); with nu==0 the CPU lib
        // does not call it either (src/old/VX_Link.cpp:154), which is the only case cpu_lib_mode supports.
        l._stress = updateStrain(l, (float)(l.pos2.x / l.currentRestLength));
        if (mat.isFailed(l.maxStrain)) {
            l.forceNeg = l.forcePos = l.momentNeg = l.momentPos = V3(0, 0, 0);
            return;
        }
        float b1 = mat.b1, b2 = mat.b2, b3 = mat.b3, a2 = mat.a2;
        const V3 &pos2 = l.pos2, &angle1v = l.angle1v, &angle2v = l.angle2v;
        l.forceNeg = V3(l._stress * l.currentTransverseArea, b1 * pos2.y - b2 * (angle1v.z + angle2v.z),
                        b1 * pos2.z + b2 * (angle1v.y + angle2v.y));
        l.forcePos = -l.forceNeg;
        l.momentNeg = V3(a2 * (angle2v.x - angle1v.x), -b2 * pos2.z - b3 * (2 * angle1v.y + angle2v.y),
                         b2 * pos2.y - b3 * (2 * angle1v.z + angle2v.z));
        l.momentPos = V3(a2 * (angle1v.x - angle2v.x), -b2 * pos2.z - b3 * (angle1v.y + 2 * angle2v.y),
                         b2 * pos2.y - b3 * (angle1v.z + 2 * angle2v.z));
        if (l.boolStates & VX3_LINK_LOCAL_VELOCITY_VALID) {
            float sqA1 = mat.sqA1, sqA2xIp = mat.sqA2xIp, sqB1 = mat.sqB1, sqB2xFMp = mat.sqB2xFMp, sqB3xIp = mat.sqB3xIp;
            V3 posCalc(sqA1 * dPos2.x, sqB1 * dPos2.y - sqB2xFMp * (dAngle1.z + dAngle2.z),
                       sqB1 * dPos2.z + sqB2xFMp * (dAngle1.y + dAngle2.y));
            float dmN = dampingMultiplier(vox[l.vNeg]), dmP = dampingMultiplier(vox[l.vPos]);
            l.forceNeg += dmN * posCalc;
            l.forcePos -= dmP * posCalc;
            l.momentNeg -= 0.5 * dmN *
                           V3(-sqA2xIp * (dAngle2.x - dAngle1.x), sqB2xFMp * dPos2.z + sqB3xIp * (2 * dAngle1.y + dAngle2.y),
                              -sqB2xFMp * dPos2.y + sqB3xIp * (2 * dAngle1.z + dAngle2.z));
            l.momentPos -= 0.5 * dmP *
                           V3(sqA2xIp * (dAngle2.x - dAngle1.x), sqB2xFMp * dPos2.z + sqB3xIp * (dAngle1.y + 2 * dAngle2.y),
                              -sqB2xFMp * dPos2.y + sqB3xIp * (dAngle1.z + 2 * dAngle2.z));
        } else
            l.boolStates |= VX3_LINK_LOCAL_VELOCITY_VALID;
        if (!l.smallAngle) {
            l.forceNeg = l.angle1.RotateVec3DInv(l.forceNeg);
            l.momentNeg = l.angle1.RotateVec3DInv(l.momentNeg);
        }
        l.forcePos = l.angle2.RotateVec3DInv(l.forcePos);
        l.momentPos = l.angle2.RotateVec3DInv(l.momentPos);
        toAxisOriginal(l.axis, &l.forceNeg);
        toAxisOriginal(l.axis, &l.forcePos);
        toAxisOriginal(l.axis, &l.momentNeg);
        toAxisOriginal(l.axis, &l.momentPos);
        if (l.isNewLink) { // VX3_Link.cu:206-213
            l.forceNeg = l.forceNeg * 0.01;
            l.forcePos = l.forcePos * 0.01;
            l.momentNeg = l.momentNeg * 0.01;
            l.momentPos = l.momentPos * 0.01;
            l.isNewLink -= 1;
        }
    }

    // ---- voxel ----
    V3 force(OVoxel &v) { // VX3_Voxel.cu:350-380
        const OMat &m = vm(v);
        V3 totalForce(0, 0, 0);
        for (int i = 0; i < 6; i++) {
            if (v.links[i] >= 0) {
                const OLink &l = links[v.links[i]];
                totalForce += isNegative(i) ? l.forcePos : l.forceNeg; // link->force(isNegative(dir))
            }
        }
        totalForce = v.orient.RotateVec3D(totalForce);
        if (v.ext >= 0) { // external()->force() is Vec3D<float>, added component-wise
            const vx3_external &e = exts[v.ext].e;
            totalForce += V3(e.force[0], e.force[1], e.force[2]);
        }
        totalForce -= velocity(v) * m.globalDampingTranslateC();
        totalForce.z += m.gravityForce();
        if (!cpu_lib_mode) {
            totalForce -= v.contactForce;
            v.contactForce = V3();
            totalForce += v.CiliaForce * m.m.cilia;
            v.CiliaForce = V3();
        }
        return totalForce;
    }
    V3 moment(OVoxel &v) { // VX3_Voxel.cu:382-397
        const OMat &m = vm(v);
        V3 totalMoment(0, 0, 0);
        for (int i = 0; i < 6; i++) {
            if (v.links[i] >= 0) {
                const OLink &l = links[v.links[i]];
                totalMoment += isNegative(i) ? l.momentPos : l.momentNeg;
            }
        }
        totalMoment = v.orient.RotateVec3D(totalMoment);
        if (v.ext >= 0) {
            const vx3_external &e = exts[v.ext].e;
            totalMoment += V3(e.moment[0], e.moment[1], e.moment[2]);
        }
        totalMoment -= angularVelocity(v) * m.globalDampingRotateC();
        return totalMoment;
    }
    void floorForce(OVoxel &v, float dt, V3 *pTotalForce) { // VX3_Voxel.cu:399-426
        (void)dt;
        const OMat &m = vm(v);
        float CurPenetration = floorPenetration(v);
        if (CurPenetration >= 0) {
            V3 vel = velocity(v);
            V3 horizontalVel(vel.x, vel.y, 0);
            float normalForce = m.penetrationStiffness() * CurPenetration;
            pTotalForce->z += normalForce - m.collisionDampingTranslateC() * vel.z;
            if (v.boolStates & VX3_VOX_FLOOR_STATIC_FRICTION) {
                float surfaceForceSq = (float)(pTotalForce->x * pTotalForce->x + pTotalForce->y * pTotalForce->y);
                float frictionForceSq = (m.m.muStatic * normalForce) * (m.m.muStatic * normalForce);
                if (surfaceForceSq > frictionForceSq) v.boolStates &= ~VX3_VOX_FLOOR_STATIC_FRICTION;
            } else {
                *pTotalForce -= m.m.muKinetic * normalForce * horizontalVel.Normalized();
            }
        } else
            v.boolStates &= ~VX3_VOX_FLOOR_STATIC_FRICTION;
    }
    double evalProg(int slot, double x, double y, double z, double dflt) const {
        if (prog[slot].empty()) return dflt; // defined behaviour for "tag absent" (vx3_abi.h, vx3_program)
        return mt_eval(prog[slot], x, y, z, collisionCount, currentTime, recentAngle, targetCloseness, numClosePairs, (int)vox.size());
    }
    void timeStep(OVoxel &v, double dt, float dtF) { // VX3_Voxel.cu:162-277
        const OMat &m = vm(v);
        v.previousDt = dt;
        if (dt == 0.0f) return;
        const bool floorEnabled = (v.boolStates & VX3_VOX_FLOOR_ENABLED) != 0;
        if (v.ext >= 0 && exts[v.ext].isFixedAll()) {
            const vx3_external &e = exts[v.ext].e;
            double s = m.m.nomSize;
            v.pos = V3(v.ix * s, v.iy * s, v.iz * s) + V3(e.translation[0], e.translation[1], e.translation[2]);
            v.orient = Q4(e.rotation_q[0], e.rotation_q[1], e.rotation_q[2], e.rotation_q[3]);
            v.linMom = v.angMom = V3(0, 0, 0);
            return;
        }
        V3 curForce = force(v);
        if (!cpu_lib_mode) { // force field, VX3_Voxel.cu:178-184 (absent tags default to the constant 0)
            const double px = v.pos.x, py = v.pos.y, pz = v.pos.z;
            curForce.x += evalProg(VX3_PROG_FORCE_X, px, py, pz, 0.0);
            curForce.y += evalProg(VX3_PROG_FORCE_Y, px, py, pz, 0.0);
            curForce.z += evalProg(VX3_PROG_FORCE_Z, px, py, pz, 0.0);
        }
        V3 fricForce = curForce;
        if (floorEnabled) floorForce(v, dtF, &curForce);
        fricForce = curForce - fricForce;
        v.linMom += curForce * dt;
        // VX3: double dt * float massInverse; CPU lib: float dt * float massInverse (VX_Voxel.cpp:186)
        V3 translate = cpu_lib_mode ? v.linMom * (double)(dtF * m.m.massInverse) : v.linMom * (dt * m.m.massInverse);
        if (floorEnabled && floorPenetration(v) >= 0) {
            double work = fricForce.x * translate.x + fricForce.y * translate.y;
            double hKe = 0.5 * m.m.massInverse * (v.linMom.x * v.linMom.x + v.linMom.y * v.linMom.y);
            if (hKe + work <= 0) v.boolStates |= VX3_VOX_FLOOR_STATIC_FRICTION;
            if (v.boolStates & VX3_VOX_FLOOR_STATIC_FRICTION) {
                v.linMom.x = v.linMom.y = 0;
                translate.x = translate.y = 0;
            }
        } else
            v.boolStates &= ~VX3_VOX_FLOOR_STATIC_FRICTION;
        v.pos += translate;
        V3 curMoment = moment(v);
        v.angMom += curMoment * dt;
        V3 rv = cpu_lib_mode ? v.angMom * (double)(dtF * m.m.momentInertiaInverse) : v.angMom * (dt * m.m.momentInertiaInverse);
        v.orient = Q4(rv) * v.orient;
        if (v.ext >= 0) {
            const OExt &x = exts[v.ext];
            const vx3_external &e = x.e;
            double size = m.m.nomSize;
            if (x.isFixed(VX3_DOF_X_TRANSLATE)) { v.pos.x = v.ix * size + e.translation[0]; v.linMom.x = 0; }
            if (x.isFixed(VX3_DOF_Y_TRANSLATE)) { v.pos.y = v.iy * size + e.translation[1]; v.linMom.y = 0; }
            if (x.isFixed(VX3_DOF_Z_TRANSLATE)) { v.pos.z = v.iz * size + e.translation[2]; v.linMom.z = 0; }
            if (x.isFixedAnyRotation()) {
                if (x.isFixedAllRotation()) {
                    v.orient = Q4(e.rotation_q[0], e.rotation_q[1], e.rotation_q[2], e.rotation_q[3]);
                    v.angMom = V3();
                } else {
                    V3 tmpRotVec = v.orient.ToRotationVector();
                    if (x.isFixed(VX3_DOF_X_ROTATE)) { tmpRotVec.x = 0; v.angMom.x = 0; }
                    if (x.isFixed(VX3_DOF_Y_ROTATE)) { tmpRotVec.y = 0; v.angMom.y = 0; }
                    if (x.isFixed(VX3_DOF_Z_ROTATE)) { tmpRotVec.z = 0; v.angMom.z = 0; }
                    v.orient.FromRotationVector(tmpRotVec);
                }
            }
        }
        if (!cpu_lib_mode) { // VX3_Voxel.cu:259-264
            if (floorEnabled && floorPenetration(v) >= 0) {
                if (v.boolStates & VX3_VOX_FLOOR_STATIC_FRICTION) v.angMom = V3(0, 0, 0);
            }
        }
        v.poissonsStrainInvalid = true; // VX3_Voxel.cu:266 (not reached by a voxel that returns early: dt = 0, all DOFs fixed)
    }

    // ---- signals (VX3_Voxel.cu:279-348).  The reference runs these at the end of every voxel's timeStep with the
    // voxels in parallel (racy: a voxel writes its neighbours' state).  Canonical order here and on the GPU: the
    // voxels take their turns in ascending voxel index.
    void receiveSignal(OVoxel &v, double signalValue, double activeTime, bool force) { // :315-335
        const OMat &m = vm(v);
        if (!force) {
            if (v.inactiveUntil > activeTime) return;
        }
        if (signalValue < 0.1) return;
        v.inactiveUntil = activeTime + m.m.inactive_period;
        v.localSignal = signalValue;
        v.sigValue = signalValue * m.m.signal_value_decay;
        if (v.sigValue < 0.1) v.sigValue = 0;
        v.sigActiveTime = activeTime;
    }
    void propagateSignal(int vi, double t) { // :336-362
        OVoxel &v = vox[vi];
        const OMat &m = vm(v);
        if (v.sigActiveTime > t) return;
        if (v.sigValue < 0.1) return;
        for (int i = 0; i < 6; i++) {
            if (v.links[i] >= 0) {
                const OLink &l = links[v.links[i]];
                const int other = (l.vNeg == vi) ? l.vPos : l.vNeg;
                receiveSignal(vox[other], v.sigValue, t + m.m.signal_time_delay, false);
            }
        }
        v.sigValue = 0;
        v.sigActiveTime = 0;
        v.inactiveUntil = t + 2 * m.m.signal_time_delay + m.m.inactive_period;
    }
    void packMaker(OVoxel &v, double t) { // :304-313
        const OMat &m = vm(v);
        if (!m.m.is_pacemaker) return;
        if (v.packmakerNextPulse > t) return;
        receiveSignal(v, 100, t, true);
        v.packmakerNextPulse = t + m.m.pacemaker_period;
    }
    void localSignalDecay(OVoxel &v, double t) { // :291-302
        if (v.localSignaldt > t) return;
        if (v.localSignal < 0.1) v.localSignal = 0;
        else {
            v.localSignal = v.localSignal * 0.9;
            v.localSignaldt = t + 0.01;
        }
    }
    // does timeStep reach its end for this voxel (:162-174 early returns; gpu_update_voxels :586-589 skips)?
    bool runsSignals(const OVoxel &v, double dt) const {
        if (v.removed || vm(v).m.fixed || dt == 0.0) return false;
        if (v.ext >= 0 && exts[v.ext].isFixedAll()) return false;
        return true;
    }

    // ---- kernel-level ----
    double recommendedTimeStep() { // VX3_VoxelyzeKernel.cu:184-217
        double MaxFreq2 = 0.0f;
        // NB iterates d_links[0..num_d_links): the host-built links only, not attach-created ones
        for (int i = 0; i < nHostLinks; i++) {
            OLink &l = links[i];
            const OMat &lm = lmats[l.mat];
            double m1 = vm(vox[l.vNeg]).m.mass, m2 = vm(vox[l.vPos]).m.mass;
            float stiff;
            if (lm.m.nu == 0.0f) stiff = lm.a1;
            else {
                updateRestLength(l);
                stiff = (float)(lm.m.eHat * l.currentTransverseArea / ((l.strain + 1) * l.currentRestLength));
            }
            double thisMaxFreq2 = stiff / (m1 < m2 ? m1 : m2);
            if (thisMaxFreq2 > MaxFreq2) MaxFreq2 = thisMaxFreq2;
        }
        if (MaxFreq2 <= 0.0f) {
            for (auto &v : vox) {
                const OMat &m = vm(v);
                double thisMaxFreq2 = m.m.E * m.m.nomSize / m.m.mass;
                if (thisMaxFreq2 > MaxFreq2) MaxFreq2 = thisMaxFreq2;
            }
        }
        if (MaxFreq2 <= 0.0f) return 0.0f;
        return 1.0f / (6.283185f * sqrt(MaxFreq2));
    }
    int nHostLinks = 0;

    void updateTemperature() { // VX3_VoxelyzeKernel.cu:219-235,625-650
        if (!opt.vary_temp_enabled || !(opt.temp_period > 0)) return;
        for (auto &v : vox) {
            const OMat &m = vm(v);
            if (v.removed) continue;
            if (m.m.thermal_on_after_s > currentTime) continue;
            if (m.m.fixed) continue;
            double currentTemperature = opt.temp_amplitude * m_sin(2 * 3.1415926f * (currentTime / opt.temp_period + v.phaseOffset));
            if (!opt.enable_expansion) {
                if (currentTemperature > 0) currentTemperature = 0;
            }
            v.tempe = (float)currentTemperature; // setTemperature(float)
            for (int i = 0; i < 6; i++)
                if (v.links[i] >= 0) updateRestLength(links[v.links[i]]);
        }
    }
    void updateCurrentCenterOfMass() { // VX3_VoxelyzeKernel.cu:477-493
        double TotalMass = 0;
        V3 Sum(0, 0, 0);
        for (auto &v : vox) {
            const OMat &m = vm(v);
            if (!m.m.is_measured) continue;
            double ThisMass = m.m.mass;
            Sum += v.pos * ThisMass;
            TotalMass += ThisMass;
        }
        if (TotalMass == 0) {
            currentCenterOfMass = V3();
            return;
        }
        currentCenterOfMass = Sum / TotalMass;
    }
    void updateSurfaceFlag(OVoxel &v) { // VX3_Voxel.cu:515-524 (bit named SURFACE means interior)
        bool interior = true;
        for (int i = 0; i < 6; i++) {
            if (v.links[i] < 0) interior = false;
            else if (links[v.links[i]].isDetached) interior = false;
        }
        if (interior) v.boolStates |= VX3_VOX_SURFACE;
        else v.boolStates &= ~VX3_VOX_SURFACE;
    }
    void regenerateSurfaceVoxels() { // VX3_VoxelyzeKernel.cu:495-513
        surface.clear();
        for (int i = 0; i < (int)vox.size(); i++) {
            updateSurfaceFlag(vox[i]);
            if (!(vox[i].boolStates & VX3_VOX_SURFACE) && !vox[i].removed) surface.push_back(i);
        }
    }
    bool is_neighbor(int voxel1, int voxel2, int incoming_link, int depth) const { // VX3_VoxelyzeKernel.cu:651-680
        if (voxel1 == voxel2) return true;
        if (depth <= 0) return false;
        const OVoxel &v1 = vox[voxel1];
        for (int i = 0; i < 6; i++) {
            int li = v1.links[i];
            if (li >= 0 && li != incoming_link) {
                const OLink &l = links[li];
                int other = (l.vNeg == voxel1) ? l.vPos : l.vNeg;
                if (is_neighbor(other, voxel2, li, depth - 1)) return true;
            }
        }
        return false;
    }
    int combinedMaterial(int mat1, int mat2); // VX3_VoxelyzeKernel.cu:515-528
    void handle_collision_attachment(int i1, int i2); // VX3_VoxelyzeKernel.cu:682-831
    void updateAttach() { // VX3_VoxelyzeKernel.cu:401-459,833-843 in the canonical order
        const int S = (int)surface.size();
        // The pair order is the reference's (first ascending, second < first).  Positions, temperatures and materials do not
        // change during the sweep, so the axis tests that open handle_collision_attachment (:687-695, all side-effect free) are
        // evaluated here on flat copies — same expressions, same comparisons — and only the pairs that pass them enter the
        // function (which repeats them).  Pure speed-up of the checker: 4e8 pair tests per step for config 4 at full size.
        std::vector<double> px(S), py(S), pz(S), bs(S);
        for (int i = 0; i < S; i++) {
            const OVoxel &v = vox[surface[i]];
            px[i] = v.pos.x; py[i] = v.pos.y; pz[i] = v.pos.z;
            bs[i] = baseSizeAverage(v);
        }
        for (int first = 0; first < S; first++) {
            const double fx = px[first], fy = py[first], fz = pz[first], fb = bs[first];
            for (int second = 0; second < first; second++) {
                const double w = (fb + bs[second]) * COLLISION_ENVELOPE_RADIUS;
                const double dx = fx - px[second];
                if (dx > w || dx < -w) continue;
                const double dy = fy - py[second];
                if (dy > w || dy < -w) continue;
                const double dz = fz - pz[second];
                if (dz > w || dz < -w) continue;
                if (vox[surface[first]].removed || vox[surface[second]].removed) continue;
                handle_collision_attachment(surface[first], surface[second]);
            }
        }
    }
    void updateDetach() { // VX3_VoxelyzeKernel.cu:461-475,946-968
        for (int li = 0; li < (int)links.size(); li++) {
            OLink &t = links[li];
            if (t.removed || t.isDetached) continue;
            if (lmats[t.mat].isFailed(t.maxStrain)) {
                t.isDetached = true;
                for (int i = 0; i < 6; i++) {
                    if (vox[t.vNeg].links[i] == li) vox[t.vNeg].links[i] = -1;
                    if (vox[t.vPos].links[i] == li) vox[t.vPos].links[i] = -1;
                }
                isSurfaceChanged = true;
                detachEvents++;
            }
        }
    }
    void computeTargetCloseness() { // VX3_VoxelyzeKernel.cu:545-563
        if (opt.max_dist_in_voxel_lengths_to_count_as_pair == 0) return;
        double R = opt.max_dist_in_voxel_lengths_to_count_as_pair * opt.vox_size;
        double ret = 0;
        numClosePairs = 0;
        for (size_t i = 0; i < targets.size(); i++)
            for (size_t j = i + 1; j < targets.size(); j++) {
                double distance = vox[targets[i]].pos.Dist(vox[targets[j]].pos);
                if (distance < R) numClosePairs++;
                ret += 1 / distance;
            }
        targetCloseness = ret;
    }
    void removeVoxels() { // VX3_VoxelyzeKernel.cu:365-399
        for (int i = 0; i < (int)vmats.size(); i++) {
            OMat &m = vmats[i];
            if (!m.removedFlag && m.m.remove_after_s > 0 && m.m.remove_after_s < currentTime) {
                for (int j = 0; j < (int)vox.size(); j++) {
                    OVoxel &v = vox[j];
                    if (v.mat == i && !v.removed) {
                        v.removed = true;
                        for (int k = 0; k < 6; k++) {
                            int li = v.links[k];
                            if (li < 0) continue;
                            links[li].removed = true;
                            int nb = (links[li].vNeg == j) ? links[li].vPos : links[li].vNeg;
                            for (int q = 0; q < 6; q++)
                                if (vox[nb].links[q] == li) {
                                    vox[nb].links[q] = -1;
                                    break;
                                }
                            v.links[k] = -1;
                        }
                    }
                }
                m.removedFlag = true;
                isSurfaceChanged = true;
            }
        }
    }
    void saveInitialPosition() {
        initialPosition.resize(vox.size());
        for (size_t i = 0; i < vox.size(); i++) initialPosition[i] = vox[i].pos;
    }

    bool doTimeStep(float dt) { // VX3_VoxelyzeKernel.cu:237-359
        if (!cpu_lib_mode) updateTemperature();
        CurStepCount++;
        if (dt == 0) return true;
        else if (dt < 0) {
            if (!OptimalDt) OptimalDt = recommendedTimeStep();
            if (OptimalDt < 1e-10) OptimalDt = 1e-10;
            dt = opt.dt_frac * OptimalDt;
        }
        lastDt = dt;
        bool Diverged = false;
        for (auto &l : links) { // gpu_update_links :566-581
            if (l.removed) continue;
            if (vm(vox[l.vPos]).m.fixed && vm(vox[l.vNeg]).m.fixed) continue;
            if (l.isDetached) continue;
            updateForces(l);
            // the reference samples ONE random link per step (:273-280); this engine defines the
            // check over every link (superset; equals the CPU library, Voxelyze.cpp:265)
            if (l.strain > 100) Diverged = true;
        }
        if (Diverged) return false;
        if (!cpu_lib_mode) {
            if (isSurfaceChanged) {
                isSurfaceChanged = false;
                regenerateSurfaceVoxels();
            }
            if (opt.enable_attach || opt.enable_collision) updateAttach();
            if (opt.enable_detach) updateDetach();
            if (opt.enable_cilia) { // gpu_update_cilia_force :846-859
                for (int si : surface) {
                    OVoxel &v = vox[si];
                    const OMat &m = vm(v);
                    if (v.removed || m.m.cilia == 0 || m.m.cilia_on_after_s > currentTime) continue;
                    v.CiliaForce = v.orient.RotateVec3D(v.baseCiliaForce + v.localSignal * v.shiftCiliaForce);
                }
            }
        }
        const double dtD = dt;
        for (auto &v : vox) { // gpu_update_voxels :582-623
            if (cpu_lib_mode) {
                timeStep(v, dtD, dt);
                continue;
            }
            if (v.removed) continue;
            if (vm(v).m.fixed) continue;
            timeStep(v, dtD, dt);
            v.enableAttach = false;
            bool all = true;
            for (int c = 0; c < 5 && all; c++) all = evalProg(VX3_PROG_ATTACH_0 + c, v.pos.x, v.pos.y, v.pos.z, 1.0) > 0;
            if (all) v.enableAttach = true;
            if (opt.enable_signals && runsSignals(v, dtD)) { // VX3_Voxel.cu:270-275
                const int vi = (int)(&v - &vox[0]);
                propagateSignal(vi, currentTime);
                packMaker(v, currentTime);
                localSignalDecay(v, currentTime);
            }
        }
        if (!cpu_lib_mode) {
            int CycleStep = int(opt.temp_period / dt);
            if (CycleStep > 0 && CurStepCount % CycleStep == 0) { // reference divides by zero when TempPeriod < dt
                angleSampleTimes++;
                comHistory[0] = comHistory[1];
                comHistory[1] = currentCenterOfMass;
                updateCurrentCenterOfMass();
                V3 A = comHistory[0], B = comHistory[1], C = currentCenterOfMass;
                if (B == C || A == B || angleSampleTimes < 3) recentAngle = 0;
                else recentAngle = acos((B - A).Dot(C - B) / (B.Dist(A) * C.Dist(B)));
                computeTargetCloseness();
            }
            if (opt.secondary_experiment) {
                removeVoxels();
                if (!InitialPositionReinitialized && opt.reinit_initial_position_after_s < currentTime) {
                    InitialPositionReinitialized = true;
                    initialCenterOfMass = currentCenterOfMass;
                    saveInitialPosition();
                }
            }
        }
        currentTime += dt;
        return true;
    }
    bool StopConditionMet() const { // VX3_VoxelyzeKernel.cu:162-182
        if (prog[VX3_PROG_STOP].empty()) return false;
        return mt_eval(prog[VX3_PROG_STOP], currentCenterOfMass.x, currentCenterOfMass.y, currentCenterOfMass.z, collisionCount,
                       currentTime, recentAngle, targetCloseness, numClosePairs, (int)vox.size()) > 0;
    }
    void computeFitness() { // VX3_VoxelyzeKernel.cu:530-534
        V3 offset = currentCenterOfMass - initialCenterOfMass;
        if (prog[VX3_PROG_FITNESS].empty()) {
            fitness_score = 0;
            return;
        }
        fitness_score = mt_eval(prog[VX3_PROG_FITNESS], offset.x, offset.y, offset.z, collisionCount, currentTime, recentAngle,
                                targetCloseness, numClosePairs, (int)vox.size());
    }
};

int vx3o_sim::combinedMaterial(int mat1, int mat2) {
    for (int i = 0; i < (int)lmats.size(); i++)
        if ((lmats[i].vox1 == mat1 && lmats[i].vox2 == mat2) || (lmats[i].vox1 == mat2 && lmats[i].vox2 == mat1)) return i;
    // VX3_MaterialLink(mat1, mat2) -> updateAll, VX3_MaterialLink.cu:53-127 (linear pair only: the
    // device setModel() asserts false, VX3_Material.cu:234) + updateDerived :129-149
    const vx3_voxel_material &a = vmats[mat1].m, &b = vmats[mat2].m;
    OMat n;
    memset(&n.m, 0, sizeof(n.m));
    n.vox1 = mat1;
    n.vox2 = mat2;
    n.m.nomSize = 0.5 * (a.nomSize + b.nomSize);
    n.m.rho = 0.5f * (a.rho + b.rho);
    n.m.alphaCTE = 0.5f * (a.alphaCTE + b.alphaCTE);
    n.m.muStatic = 0.5f * (a.muStatic + b.muStatic);
    n.m.muKinetic = 0.5f * (a.muKinetic + b.muKinetic);
    n.m.zetaInternal = 0.5f * (a.zetaInternal + b.zetaInternal);
    n.m.zetaGlobal = 0.5f * (a.zetaGlobal + b.zetaGlobal);
    n.m.zetaCollision = 0.5f * (a.zetaCollision + b.zetaCollision);
    n.m.extScale[0] = n.m.extScale[1] = n.m.extScale[2] = 1.0;
    float stressFail, f1 = a.sigmaFail, f2 = b.sigmaFail;
    if (f1 == -1.0f) stressFail = f2;
    else if (f2 == -1.0f) stressFail = f1;
    else stressFail = f1 < f2 ? f1 : f2;
    { // setModelLinear, VX3_Material.cu:251-278
        float youngsModulus = 2.0f * a.E * b.E / (a.E + b.E), failureStress = stressFail;
        float tmpfailureStress = failureStress;
        if (tmpfailureStress == -1) tmpfailureStress = 1000000;
        float tmpfailStrain = tmpfailureStress / youngsModulus;
        n.strainData = {0.0f, tmpfailStrain};
        n.stressData = {0.0f, tmpfailureStress};
        n.m.linear = 1;
        n.m.E = youngsModulus;
        n.m.sigmaYield = failureStress;
        n.m.sigmaFail = failureStress;
        n.m.epsilonYield = (failureStress == -1) ? -1 : tmpfailStrain;
        n.m.epsilonFail = (failureStress == -1) ? -1 : tmpfailStrain;
    }
    if (a.nu == 0 && b.nu == 0) n.m.nu = 0;
    else {
        float tmpEHat = 2 * a.eHat * b.eHat / (a.eHat + b.eHat);
        float tmpE = n.m.E;
        float c2 = (tmpEHat - tmpE) / (2 * tmpEHat) + 0.0625;
        n.m.nu = sqrt(c2) - 0.25;
    }
    n.m.eHat = n.m.E / ((1 - 2 * n.m.nu) * (1 + n.m.nu));
    float L = (float)n.m.nomSize, E = n.m.E, nu = n.m.nu;
    n.a1 = E * L;
    n.a2 = E * L * L * L / (12.0f * (1 + nu));
    n.b1 = E * L;
    n.b2 = E * L * L / 2.0f;
    n.b3 = E * L * L * L / 6.0f;
    n.sqA1 = sqrtf(n.a1);
    n.sqA2xIp = sqrtf(n.a2 * L * L / 6.0f);
    n.sqB1 = sqrtf(n.b1);
    n.sqB2xFMp = sqrtf(n.b2 * L / 2.0f);
    n.sqB3xIp = sqrtf(n.b3 * L * L / 6.0f);
    lmats.push_back(n);
    return (int)lmats.size() - 1;
}

void vx3o_sim::handle_collision_attachment(int i1, int i2) {
    OVoxel &voxel1 = vox[i1], &voxel2 = vox[i2];
    const OMat &m1 = vm(voxel1), &m2 = vm(voxel2);
    if (m1.m.fixed && m2.m.fixed) return;
    V3 diff = voxel1.pos - voxel2.pos;
    double watchDistance = (baseSizeAverage(voxel1) + baseSizeAverage(voxel2)) * COLLISION_ENVELOPE_RADIUS;
    if (diff.x > watchDistance || diff.x < -watchDistance) return;
    if (diff.y > watchDistance || diff.y < -watchDistance) return;
    if (diff.z > watchDistance || diff.z < -watchDistance) return;
    if (diff.Length() > watchDistance) return;
    if (is_neighbor(i1, i2, -1, 1)) return;
    V3 cache1, cache2;
    if (opt.enable_collision) { // VX3_Collision.cu:3-31
        double penetrationStiff = 2.0f / (1.0f / m1.penetrationStiffness() + 1.0f / m2.penetrationStiffness());
        double dampingC = 0.5f * (m1.collisionDampingTranslateC() + m2.collisionDampingTranslateC());
        V3 offset = voxel2.pos - voxel1.pos;
        double NomDist = (double)((baseSizeAverage(voxel1) + baseSizeAverage(voxel2)) * COLLISION_ENVELOPE_RADIUS);
        double RelDist = NomDist - offset.Length();
        V3 force;
        if (RelDist > 0) {
            V3 unit = offset.Normalized();
            double relativeVelocity = velocity(voxel1).Dot(unit) - velocity(voxel2).Dot(unit);
            force = unit * (penetrationStiff * RelDist + dampingC * relativeVelocity);
        } else
            force = V3(0, 0, 0);
        cache1 = force;
        cache2 = -force;
        voxel1.contactForce += cache1;
        voxel2.contactForce += cache2;
        if ((m1.m.is_target && !m2.m.is_target) || (m2.m.is_target && !m1.m.is_target)) {
            collisionCount++;
            if (opt.enable_signals) { // :719-725: the non-target voxel of the pair fires
                if (m1.m.is_target) receiveSignal(voxel2, 100, currentTime, true);
                else receiveSignal(voxel1, 100, currentTime, true);
            }
        }
    }
    if (!voxel1.enableAttach || !voxel2.enableAttach) return;
    if (m1.m.fixed || m2.m.fixed) return;
    if (voxel1.mat != voxel2.mat) return;
    if (!m1.m.sticky) return;
    // NB: the reference calls handle_collision_attachment for every pair even when enableAttach (the
    // kernel-level flag) is off — the per-voxel flag + sticky material gate it (:729-740).
    if (is_neighbor(i1, i2, -1, 5)) return;
    int link_dir_1, link_dir_2, link_axis;
    V3 e = voxel1.pos - voxel2.pos;
    V3 ea = voxel1.orient.RotateVec3DInv(-e);
    bool reverseOrder = false;
    V3 f = ea.Abs();
    if (f.x >= f.y && f.x >= f.z) {
        link_axis = 0;
        if (ea.x < 0) { link_dir_1 = 1; link_dir_2 = 0; reverseOrder = true; }
        else { link_dir_1 = 0; link_dir_2 = 1; }
    } else if (f.y >= f.x && f.y >= f.z) {
        link_axis = 1;
        if (ea.y < 0) { link_dir_1 = 3; link_dir_2 = 2; reverseOrder = true; }
        else { link_dir_1 = 2; link_dir_2 = 3; }
    } else {
        link_axis = 2;
        if (ea.z < 0) { link_dir_1 = 5; link_dir_2 = 4; reverseOrder = true; }
        else { link_dir_1 = 4; link_dir_2 = 5; }
    }
    if (voxel1.links[link_dir_1] < 0 && voxel2.links[link_dir_2] < 0) {
        // VX3_Link(voxelA, dirA, voxelB, dirB, axis): pVNeg = voxelB, pVPos = voxelA (VX3_Link.cu:31-56)
        OLink L;
        int li = (int)links.size();
        voxel1.links[link_dir_1] = li;
        voxel2.links[link_dir_2] = li;
        L.axis = link_axis;
        if (reverseOrder) { L.vNeg = i2; L.vPos = i1; }
        else { L.vNeg = i1; L.vPos = i2; }
        L.mat = combinedMaterial(vox[L.vPos].mat, vox[L.vNeg].mat);
        L.boolStates = 0;
        links.push_back(L);
        resetLink(links.back());
        links.back().isNewLink = opt.safety_guard;
        isSurfaceChanged = true;
        attachEvents++;
        OVoxel &a = vox[i1], &b2 = vox[i2];
        a.contactForce -= cache1;
        b2.contactForce -= cache2;
    }
}

// ================================================================== C interface
extern "C" {

vx3o_sim *vx3o_create(const vx3_model_desc *m, int cpu_lib_mode) {
    if (!m) return nullptr;
    vx3o_sim *s = new vx3o_sim();
    s->cpu_lib_mode = cpu_lib_mode;
    s->name = m->name;
    s->opt = m->opt;
    s->vmats.resize(m->n_voxel_mats);
    for (int i = 0; i < m->n_voxel_mats; i++) load_mat(m->voxel_mats[i], s->vmats[i], cpu_lib_mode != 0);
    s->lmats.resize(m->n_link_mats);
    for (int i = 0; i < m->n_link_mats; i++) {
        const vx3_link_material &lm = m->link_mats[i];
        OMat &o = s->lmats[i];
        load_mat(lm.m, o, cpu_lib_mode != 0);
        o.vox1 = lm.vox1_mat; o.vox2 = lm.vox2_mat;
        o.a1 = lm.a1; o.a2 = lm.a2; o.b1 = lm.b1; o.b2 = lm.b2; o.b3 = lm.b3;
        o.sqA1 = lm.sqA1; o.sqA2xIp = lm.sqA2xIp; o.sqB1 = lm.sqB1; o.sqB2xFMp = lm.sqB2xFMp; o.sqB3xIp = lm.sqB3xIp;
    }
    s->exts.resize(m->n_externals);
    for (int i = 0; i < m->n_externals; i++) s->exts[i].e = m->externals[i];
    s->vox.resize(m->n_voxels);
    for (int i = 0; i < m->n_voxels; i++) {
        OVoxel &v = s->vox[i];
        v.mat = m->vox_mat[i];
        v.ix = m->ix[i]; v.iy = m->iy[i]; v.iz = m->iz[i];
        v.pos = V3(m->pos[3 * i], m->pos[3 * i + 1], m->pos[3 * i + 2]);
        if (m->orient) v.orient = Q4(m->orient[4 * i], m->orient[4 * i + 1], m->orient[4 * i + 2], m->orient[4 * i + 3]);
        if (m->lin_mom) v.linMom = V3(m->lin_mom[3 * i], m->lin_mom[3 * i + 1], m->lin_mom[3 * i + 2]);
        if (m->ang_mom) v.angMom = V3(m->ang_mom[3 * i], m->ang_mom[3 * i + 1], m->ang_mom[3 * i + 2]);
        v.boolStates = m->vox_flags[i];
        v.tempe = m->temp ? m->temp[i] : 0.0f;
        v.phaseOffset = m->phase_offset ? m->phase_offset[i] : 0.0;
        for (int k = 0; k < 6; k++) v.links[k] = m->vox_links[6 * i + k];
        v.ext = m->vox_ext ? m->vox_ext[i] : -1;
        if (m->base_cilia) v.baseCiliaForce = V3(m->base_cilia[3 * i], m->base_cilia[3 * i + 1], m->base_cilia[3 * i + 2]);
        if (m->shift_cilia) v.shiftCiliaForce = V3(m->shift_cilia[3 * i], m->shift_cilia[3 * i + 1], m->shift_cilia[3 * i + 2]);
    }
    s->links.resize(m->n_links);
    s->nHostLinks = m->n_links;
    for (int i = 0; i < m->n_links; i++) {
        OLink &l = s->links[i];
        l.vNeg = m->link_vneg[i]; l.vPos = m->link_vpos[i]; l.axis = m->link_axis[i]; l.mat = m->link_mat[i];
        s->importing = true;
        s->resetLink(l);
        s->importing = false;
        if (m->link_pos2) l.pos2 = V3(m->link_pos2[3 * i], m->link_pos2[3 * i + 1], m->link_pos2[3 * i + 2]);
        if (m->link_angle1v) l.angle1v = V3(m->link_angle1v[3 * i], m->link_angle1v[3 * i + 1], m->link_angle1v[3 * i + 2]);
        if (m->link_angle2v) l.angle2v = V3(m->link_angle2v[3 * i], m->link_angle2v[3 * i + 1], m->link_angle2v[3 * i + 2]);
        if (m->link_strain) l.strain = m->link_strain[i];
        if (m->link_max_strain) l.maxStrain = m->link_max_strain[i];
        if (m->link_strain_offset) l.strainOffset = m->link_strain_offset[i];
        if (m->link_stress) l._stress = m->link_stress[i];
        if (m->link_flags) l.boolStates = m->link_flags[i];
        if (m->link_small_angle) l.smallAngle = m->link_small_angle[i] != 0;
        if (m->link_rest_length) l.currentRestLength = m->link_rest_length[i];
        if (m->link_transverse_area) l.currentTransverseArea = m->link_transverse_area[i];
        if (m->link_transverse_strain_sum) l.currentTransverseStrainSum = m->link_transverse_strain_sum[i];
        if (m->link_strain_ratio) l.strainRatio = m->link_strain_ratio[i];
    }
    for (int p = 0; p < VX3_PROG_COUNT; p++)
        if (m->prog[p].n > 0 && m->prog[p].tok) s->prog[p].assign(m->prog[p].tok, m->prog[p].tok + m->prog[p].n);
    // device-side init at the top of CUDA_Simulation, VX3_SimulationManager.cu:20-24,54-55
    s->saveInitialPosition();
    s->isSurfaceChanged = true;
    for (int i = 0; i < (int)s->vox.size(); i++)
        if (s->vm(s->vox[i]).m.is_target) s->targets.push_back(i);
    s->updateCurrentCenterOfMass();
    s->initialCenterOfMass = s->currentCenterOfMass;
    return s;
}

void vx3o_destroy(vx3o_sim *s) { delete s; }

double vx3o_recommended_dt(vx3o_sim *s) { return s->recommendedTimeStep(); }

// k calls of doTimeStep(dt) without the stop-condition test; returns steps done (stops early on divergence)
long vx3o_step(vx3o_sim *s, long k, float dt) {
    long done = 0;
    for (; done < k; done++) {
        if (s->status != VX3_SIM_RUNNING) break;
        if (!s->doTimeStep(dt)) {
            s->status = VX3_SIM_DIVERGED;
            break;
        }
    }
    return done;
}

// the CUDA_Simulation loop, VX3_SimulationManager.cu:62-117 (history emission is in vx3o_history_frame)
long vx3o_run(vx3o_sim *s, long max_steps) {
    if (max_steps <= 0) max_steps = 1000000;
    long j = 0;
    for (; j < max_steps; j++) {
        if (s->StopConditionMet()) {
            s->status = VX3_SIM_STOPPED;
            break;
        }
        if (!s->doTimeStep(-1.0f)) {
            s->status = VX3_SIM_DIVERGED;
            break;
        }
    }
    if (j == max_steps && s->status == VX3_SIM_RUNNING) s->status = VX3_SIM_STEP_CAP;
    s->updateCurrentCenterOfMass();
    s->computeFitness();
    return j;
}

int vx3o_result(vx3o_sim *s, vx3_result *r, int refresh) {
    if (refresh) {
        s->updateCurrentCenterOfMass();
        s->computeFitness();
    }
    memset(r, 0, sizeof(*r));
    strncpy(r->name, s->name.c_str(), sizeof(r->name) - 1);
    r->status = s->status;
    r->num_voxel = (int)s->vox.size();
    r->num_close_pairs = s->numClosePairs;
    r->steps = (int64_t)s->CurStepCount;
    r->num_links = (int)s->links.size();
    r->collision_count = s->collisionCount;
    r->current_time = s->currentTime;
    r->fitness_score = s->status == VX3_SIM_DIVERGED ? NAN : s->fitness_score;
    r->vox_size = s->opt.vox_size;
    r->initial_com[0] = s->initialCenterOfMass.x; r->initial_com[1] = s->initialCenterOfMass.y; r->initial_com[2] = s->initialCenterOfMass.z;
    r->current_com[0] = s->currentCenterOfMass.x; r->current_com[1] = s->currentCenterOfMass.y; r->current_com[2] = s->currentCenterOfMass.z;
    r->recent_angle = s->recentAngle;
    r->target_closeness = s->targetCloseness;
    r->dt = s->lastDt;
    // collectResults, VX3_SimulationManager.cu:455-466
    for (size_t j = 0; j < s->vox.size(); j++) {
        if (s->vm(s->vox[j]).m.is_measured) {
            r->num_measured_voxel++;
            r->total_distance_of_all_voxels += s->vox[j].pos.Dist(s->initialPosition[j]);
        }
    }
    return 0;
}

static void put3(double *dst, size_t i, const V3 &v) {
    if (dst) { dst[3 * i] = v.x; dst[3 * i + 1] = v.y; dst[3 * i + 2] = v.z; }
}

int vx3o_state(vx3o_sim *s, vx3_state_view *w) {
    int nv = (int)s->vox.size(), nl = (int)s->links.size();
    if (w->n_voxels < nv || w->n_links < nl) {
        w->n_voxels = nv;
        w->n_links = nl;
        return -1;
    }
    w->n_voxels = nv;
    w->n_links = nl;
    for (int i = 0; i < nv; i++) {
        const OVoxel &v = s->vox[i];
        put3(w->pos, i, v.pos);
        put3(w->lin_mom, i, v.linMom);
        put3(w->ang_mom, i, v.angMom);
        put3(w->contact_force, i, v.contactForce);
        if (w->orient) { w->orient[4 * i] = v.orient.w; w->orient[4 * i + 1] = v.orient.x; w->orient[4 * i + 2] = v.orient.y; w->orient[4 * i + 3] = v.orient.z; }
        if (w->vox_flags) w->vox_flags[i] = v.boolStates;
        if (w->temp) w->temp[i] = v.tempe;
        if (w->vox_links) for (int k = 0; k < 6; k++) w->vox_links[6 * i + k] = v.links[k];
        if (w->signal) {
            double *o = w->signal + 6 * (size_t)i;
            o[0] = v.localSignal; o[1] = v.localSignaldt; o[2] = v.inactiveUntil; o[3] = v.packmakerNextPulse; o[4] = v.sigValue; o[5] = v.sigActiveTime;
        }
    }
    for (int i = 0; i < nl; i++) {
        const OLink &l = s->links[i];
        if (w->link_vneg) w->link_vneg[i] = l.vNeg;
        if (w->link_vpos) w->link_vpos[i] = l.vPos;
        if (w->link_axis) w->link_axis[i] = l.axis;
        if (w->link_mat) w->link_mat[i] = l.mat;
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
            w->link_flags[i] = ((l.boolStates & VX3_LINK_LOCAL_VELOCITY_VALID) ? VX3_LINKSTATE_LOCAL_VELOCITY_VALID : 0) |
                               (l.smallAngle ? VX3_LINKSTATE_SMALL_ANGLE : 0) | (l.isDetached ? VX3_LINKSTATE_DETACHED : 0) |
                               (l.removed ? VX3_LINKSTATE_REMOVED : 0) | (l.isNewLink << VX3_LINKSTATE_NEWLINK_SHIFT);
        if (w->link_rest_length) w->link_rest_length[i] = l.currentRestLength;
    }
    return 0;
}

int vx3o_counts(vx3o_sim *s, int *n_voxels, int *n_links, int *n_surface, long *attach_events, long *detach_events) {
    if (n_voxels) *n_voxels = (int)s->vox.size();
    if (n_links) *n_links = (int)s->links.size();
    if (n_surface) *n_surface = (int)s->surface.size();
    if (attach_events) *attach_events = s->attachEvents;
    if (detach_events) *detach_events = s->detachEvents;
    return 0;
}

// surface voxel list (indices) as of the last regenerateSurfaceVoxels
int vx3o_surface(vx3o_sim *s, int *out, int cap) {
    int n = (int)s->surface.size();
    for (int i = 0; i < n && i < cap; i++) out[i] = s->surface[i];
    return n;
}

// libm error model: seed != 0 moves every sin / cos / acos result of the physics path by -1 / 0 / +1 ulp (see jitter_ulp); 0 = exact
void vx3o_set_libm_jitter(unsigned long long seed) {
    g_libm_jitter = seed;
    g_libm_jitter_mode = seed == 1 ? 1 : (seed == 2 ? -1 : 0);
}

double vx3o_eval(const vx3_token *tok, int n, const double *vars9) {
    std::vector<vx3_token> p(tok, tok + n);
    return mt_eval(p, vars9[0], vars9[1], vars9[2], vars9[3], vars9[4], vars9[5], vars9[6], (int)vars9[7], (int)vars9[8]);
}

} // extern "C"
