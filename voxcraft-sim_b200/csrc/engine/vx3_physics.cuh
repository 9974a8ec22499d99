// Per-link and per-voxel physics on register-resident values.  Shared by the streaming kernels
// (vx3_kernels.cuh) and the on-chip persistent kernel (vx3_persistent.cuh), so both paths execute the
// same arithmetic.  Casts follow SURVEY.md Appendix A.1 (fp64 kinematics, fp32 strain/stress/constants).
//
//   material stress      src/VX3/VX3_Material.cu:90-124
//   link forces          src/VX3/VX3_Link.cu:90-260
//   voxel integration    src/VX3/VX3_Voxel.cu:162-277,350-426
//   temperature          src/VX3/VX3_VoxelyzeKernel.cu:625-650, VX3_Voxel.h:95-104
//   math tree            src/Utils/VX3_MathTree.h:50-192
#pragma once
#include "vx3_device.cuh"

namespace vx3 {

#define VX3_HYSTERESIS_FACTOR 1.2f // src/old/types.h:57-59
#define VX3_SA_BOND_BEND_RAD 0.05f
#define VX3_SA_BOND_EXT_PERC 0.50f
#define VX3_COLLISION_ENVELOPE_RADIUS 0.625 // VX3_Collision.h:5

// ------------------------------------------------------------------ math tree
// vars: x y z hit t angle closeness numClosePairs num_voxel
template <int MAXTOK>
__device__ __noinline__ double mt_eval(const vx3_token *tok, int n, const double *vars, bool *ok) {
    double values[MAXTOK];
    int vc = 0, pc = 0;
    *ok = true;
    for (int i = 0; i < n && i < MAXTOK; i++) {
        const int op = tok[i].op;
        const double v = tok[i].value;
        double out = 0.0;
        const double p0 = values[pc], p1 = (pc + 1 < MAXTOK) ? values[pc + 1] : 0.0;
        switch (op) {
        case VX3_OP_END: return values[pc];
        case VX3_OP_CONST: out = v; break;
        case VX3_OP_E: out = 2.71828182845904523536; break;
        case VX3_OP_PI: out = 3.14159265358979323846; break;
        case VX3_OP_VAR: {
            int k = (int)(v + 0.5);
            out = (k >= 0 && k <= 8) ? vars[k] : values[vc];
            break;
        }
        case VX3_OP_SIN: out = sin(p0); pc++; break;
        case VX3_OP_COS: out = cos(p0); pc++; break;
        case VX3_OP_TAN: out = tan(p0); pc++; break;
        case VX3_OP_ATAN: out = atan(p0); pc++; break;
        case VX3_OP_LOG: out = log(p0); pc++; break;
        case VX3_OP_INT: out = rint(p0); pc++; break;
        case VX3_OP_NORMALCDF: out = normcdf(p0); pc++; break;
        case VX3_OP_ADD: out = p1 + p0; pc += 2; break;
        case VX3_OP_SUB: out = p1 - p0; pc += 2; break;
        case VX3_OP_MUL: out = p1 * p0; pc += 2; break;
        case VX3_OP_DIV: out = p1 / p0; pc += 2; break;
        case VX3_OP_POW: out = pow(p1, p0); pc += 2; break;
        case VX3_OP_SQRT: out = sqrt(p0); pc++; break;
        case VX3_OP_ABS: out = fabs(p0); pc++; break;
        case VX3_OP_NOT: out = !p0; pc++; break;
        case VX3_OP_GREATERTHAN: out = p1 > p0; pc += 2; break;
        case VX3_OP_LESSTHAN: out = p1 < p0; pc += 2; break;
        case VX3_OP_AND: out = p1 && p0; pc += 2; break;
        case VX3_OP_OR: out = p1 || p0; pc += 2; break;
        default: *ok = false; return -1;
        }
        if (pc > vc) { *ok = false; return -1; }
        values[vc++] = out;
    }
    *ok = false;
    return -1;
}

// ------------------------------------------------------------------ temperature
// gpu_update_temperature (VX3_VoxelyzeKernel.cu:625-650): returns the voxel's tempe for the step at time t.
// `stored` is the value kept from earlier steps (used when this voxel is skipped).
__device__ __forceinline__ bool thermal_active(const SimC &S, const VoxMatC &m, int vflags, double t) {
    if (!S.vary_temp || !(S.temp_period > 0)) return false;
    if (vflags & VXF_REMOVED) return false;
    if (m.thermal_on_after > t) return false;
    if (m.fixed) return false;
    return true;
}
__device__ __forceinline__ float voxel_temperature(const SimC &S, double t, double phase) {
    double cur = S.temp_amp * sin(2 * 3.1415926f * (t / S.temp_period + phase));
    if (!S.enable_expansion) {
        if (cur > 0) cur = 0;
    }
    return (float)cur;
}
__device__ __forceinline__ float voxel_temperature(double temp_amp, double temp_period, bool enable_expansion, double t, double phase) {
    double cur = temp_amp * sin(2 * 3.1415926f * (t / temp_period + phase));
    if (!enable_expansion) {
        if (cur > 0) cur = 0;
    }
    return (float)cur;
}
__device__ __forceinline__ double base_size_axis(const VoxMatC &m, float tempe, int axis) { // VX3_Voxel.h:95-98
    return m.size[axis] * (1 + tempe * m.alphaCTE);
}
__device__ __forceinline__ double base_size_average(const VoxMatC &m, float tempe) { // VX3_Voxel.h:101-104
    float br = 1 + tempe * m.alphaCTE;
    double bx = br * m.size[0], by = br * m.size[1], bz = br * m.size[2];
    return (bx + by + bz) / 3.0f;
}
__device__ __forceinline__ float floor_penetration(const VoxMatC &m, float tempe, double z) { // VX3_Voxel.h:185-187
    return (float)(base_size_average(m, tempe) / 2 - m.nomSize / 2 - z);
}

// ------------------------------------------------------------------ material
__device__ __forceinline__ bool mat_failed(const LinkMatC &m, float strain) { return m.epsilonFail != -1.0f && strain > m.epsilonFail; }

__device__ float mat_stress(const LinkMatC &m, const float *__restrict__ strainData, const float *__restrict__ stressData, float strain,
                            float transverseStrainSum, bool forceLinear) { // VX3_Material.cu:90-124
    if (mat_failed(m, strain)) return 0.0f;
    if (m.linear || forceLinear || strain <= strainData[m.data_off + 1]) {
        if (m.nu == 0.0f) return m.E * strain;
        else return m.eHat * ((1 - m.nu) * strain + m.nu * transverseStrainSum);
    }
    const float *sd = strainData + m.data_off, *ss = stressData + m.data_off;
    const int DataCount = m.n_data;
    for (int i = 2; i < DataCount; i++) {
        if (strain <= sd[i] || i == DataCount - 1) {
            float Perc = (strain - sd[i - 1]) / (sd[i] - sd[i - 1]);
            float basicStress = ss[i - 1] + Perc * (ss[i] - ss[i - 1]);
            if (m.nu == 0.0f) return basicStress;
            else {
                float modulus = (ss[i] - ss[i - 1]) / (sd[i] - sd[i - 1]);
                float modulusHat = modulus / ((1 - 2 * m.nu) * (1 + m.nu));
                float effectiveStrain = basicStress / modulus;
                float effectiveTransverseStrainSum = transverseStrainSum * (effectiveStrain / strain);
                return modulusHat * ((1 - m.nu) * effectiveStrain + m.nu * effectiveTransverseStrainSum);
            }
        }
    }
    return 0.0f;
}

// ------------------------------------------------------------------ link
struct LinkRegs {
    V3 pos2, angle1v, angle2v;
    float strain, maxStrain, strainOffset, stress;
    float area, tsum;
    double rest;
    int state;
};
struct LinkOut {
    V3 forceNeg, momentNeg, forcePos, momentPos;
};

__device__ __forceinline__ float link_update_strain(LinkRegs &L, const LinkMatC &m, const float *sd, const float *ss, float axialStrain) { // VX3_Link.cu:220-260
    L.strain = axialStrain;
    if (m.linear) {
        if (axialStrain > L.maxStrain) L.maxStrain = axialStrain;
        return mat_stress(m, sd, ss, axialStrain, L.tsum, false);
    } else {
        float returnStress;
        if (axialStrain > L.maxStrain) {
            L.maxStrain = axialStrain;
            returnStress = mat_stress(m, sd, ss, axialStrain, L.tsum, false);
            if (m.nu != 0.0f) L.strainOffset = L.maxStrain - mat_stress(m, sd, ss, axialStrain, 0.0f, false) / (m.eHat * (1 - m.nu));
            else L.strainOffset = L.maxStrain - returnStress / m.E;
        } else {
            float relativeStrain = axialStrain - L.strainOffset;
            if (m.nu != 0.0f) returnStress = mat_stress(m, sd, ss, relativeStrain, L.tsum, true);
            else returnStress = m.E * relativeStrain;
        }
        return returnStress;
    }
}

// VX3_Link::updateForces (VX3_Link.cu:135-218) incl. orientLink (:90-133), in three stages so that a kernel can run the
// (rare, expensive) large-angle branch of orientLink for a whole tile in a few dense warps instead of in every warp:
//   link_stage_a      orientLink up to the small/large decision (:91-110) and the small-angle case (:111-114)
//   link_stage_large  the large-angle case (:115-125) + angle1's rotation vector — pure function of (pos2, angle2, rest)
//   link_stage_c      rotation vector of angle2, strain/stress, beam forces, damping, rotation back (:127-218)
// link_update_forces runs the three in sequence.  dmN/dmP = dampingMultiplier() of the two end voxels (VX3_Voxel.h:206-208).
struct LinkMid {
    V3 pos2, angle1v;
    Q4 angle1, angle2;
    bool small;
};

__device__ __forceinline__ void link_stage_a(LinkRegs &L, const V3 &posN, const Q4 &qN, const V3 &posP, const Q4 &qP, LinkMid &m) {
    const int axis = (L.state & LKS_AXIS_MASK) >> LKS_AXIS_SHIFT;
    V3 pos2 = toAxisX(axis, posP - posN);
    Q4 angle1 = toAxisX(axis, qN);
    Q4 angle2 = toAxisX(axis, qP);
    Q4 totalRot = angle1.Conjugate();
    pos2 = totalRot.RotateVec3D(pos2);
    angle2 = totalRot * angle2;
    angle1 = Q4();
    bool smallAngle = (L.state & LKS_SMALL) != 0;
    const float SmallTurn = (float)((fabs(pos2.z) + fabs(pos2.y)) / pos2.x);
    const float ExtendPerc = (float)(fabs(1 - pos2.x / L.rest));
    if (!smallAngle && SmallTurn < VX3_SA_BOND_BEND_RAD && ExtendPerc < VX3_SA_BOND_EXT_PERC) {
        smallAngle = true;
        L.state &= ~LKS_VALID;
    } else if (smallAngle && (SmallTurn > VX3_HYSTERESIS_FACTOR * VX3_SA_BOND_BEND_RAD || ExtendPerc > VX3_HYSTERESIS_FACTOR * VX3_SA_BOND_EXT_PERC)) {
        smallAngle = false;
        L.state &= ~LKS_VALID;
    }
    if (smallAngle) {
        pos2.x -= L.rest;
        L.state |= LKS_SMALL;
    } else
        L.state &= ~LKS_SMALL;
    m.pos2 = pos2;
    m.angle1 = angle1;
    m.angle2 = angle2;
    m.angle1v = V3(0, 0, 0); // ToRotationVector of the identity
    m.small = smallAngle;
}

__device__ __forceinline__ void link_stage_large(V3 &pos2, Q4 &angle1, Q4 &angle2, V3 &angle1v, double rest) {
    angle1 = Q4();
    double len;
    const bool have_len = angle1.FromAngleToPosX(pos2, &len);
    angle2 = angle1 * angle2;
    pos2 = V3((have_len ? len : pos2.Length()) - rest, 0, 0);
    angle1v = angle1.ToRotationVector();
}

__device__ __forceinline__ void link_stage_c(LinkRegs &L, const LinkMid &mid, const LinkMatC &m, const float *sd, const float *ss, float dmN, float dmP,
                                             LinkOut &o) {
    const int axis = (L.state & LKS_AXIS_MASK) >> LKS_AXIS_SHIFT;
    const V3 oldPos2 = L.pos2, oldAngle1v = L.angle1v, oldAngle2v = L.angle2v;
    const V3 pos2 = mid.pos2;
    const Q4 angle1 = mid.angle1, angle2 = mid.angle2;
    const bool smallAngle = mid.small;
    const V3 angle1v = mid.angle1v;
    const V3 angle2v = angle2.ToRotationVector();
    L.pos2 = pos2;
    L.angle1v = angle1v;
    L.angle2v = angle2v;
    // ---- updateForces ----
    const V3 dPos2 = 0.5 * (pos2 - oldPos2);
    const V3 dAngle1 = 0.5 * (angle1v - oldAngle1v);
    const V3 dAngle2 = 0.5 * (angle2v - oldAngle2v);
    L.stress = link_update_strain(L, m, sd, ss, (float)(pos2.x / L.rest));
    if (mat_failed(m, L.maxStrain)) {
        o.forceNeg = o.forcePos = o.momentNeg = o.momentPos = V3(0, 0, 0);
        return;
    }
    const float b1 = m.b1, b2 = m.b2, b3 = m.b3, a2 = m.a2;
    V3 forceNeg(L.stress * L.area, b1 * pos2.y - b2 * (angle1v.z + angle2v.z), b1 * pos2.z + b2 * (angle1v.y + angle2v.y));
    V3 forcePos = -forceNeg;
    V3 momentNeg(a2 * (angle2v.x - angle1v.x), -b2 * pos2.z - b3 * (2 * angle1v.y + angle2v.y), b2 * pos2.y - b3 * (2 * angle1v.z + angle2v.z));
    V3 momentPos(a2 * (angle1v.x - angle2v.x), -b2 * pos2.z - b3 * (angle1v.y + 2 * angle2v.y), b2 * pos2.y - b3 * (angle1v.z + 2 * angle2v.z));
    if (L.state & LKS_VALID) {
        const float sqA1 = m.sqA1, sqA2xIp = m.sqA2xIp, sqB1 = m.sqB1, sqB2xFMp = m.sqB2xFMp, sqB3xIp = m.sqB3xIp;
        const V3 posCalc(sqA1 * dPos2.x, sqB1 * dPos2.y - sqB2xFMp * (dAngle1.z + dAngle2.z), sqB1 * dPos2.z + sqB2xFMp * (dAngle1.y + dAngle2.y));
        forceNeg += dmN * posCalc;
        forcePos -= dmP * posCalc;
        momentNeg -= 0.5 * dmN *
                     V3(-sqA2xIp * (dAngle2.x - dAngle1.x), sqB2xFMp * dPos2.z + sqB3xIp * (2 * dAngle1.y + dAngle2.y),
                        -sqB2xFMp * dPos2.y + sqB3xIp * (2 * dAngle1.z + dAngle2.z));
        momentPos -= 0.5 * dmP *
                     V3(sqA2xIp * (dAngle2.x - dAngle1.x), sqB2xFMp * dPos2.z + sqB3xIp * (dAngle1.y + 2 * dAngle2.y),
                        -sqB2xFMp * dPos2.y + sqB3xIp * (dAngle1.z + 2 * dAngle2.z));
    } else
        L.state |= LKS_VALID;
    if (!smallAngle) {
        forceNeg = angle1.RotateVec3DInv(forceNeg);
        momentNeg = angle1.RotateVec3DInv(momentNeg);
    }
    forcePos = angle2.RotateVec3DInv(forcePos);
    momentPos = angle2.RotateVec3DInv(momentPos);
    forceNeg = toAxisOriginal(axis, forceNeg);
    forcePos = toAxisOriginal(axis, forcePos);
    momentNeg = toAxisOriginal(axis, momentNeg);
    momentPos = toAxisOriginal(axis, momentPos);
    const int newLink = (unsigned)L.state >> LKS_NEWLINK_SHIFT;
    if (newLink) { // VX3_Link.cu:206-213
        forceNeg = forceNeg * 0.01;
        forcePos = forcePos * 0.01;
        momentNeg = momentNeg * 0.01;
        momentPos = momentPos * 0.01;
        L.state = (L.state & ((1 << LKS_NEWLINK_SHIFT) - 1)) | ((newLink - 1) << LKS_NEWLINK_SHIFT);
    }
    o.forceNeg = forceNeg;
    o.forcePos = forcePos;
    o.momentNeg = momentNeg;
    o.momentPos = momentPos;
}

__device__ __forceinline__ void link_update_forces(LinkRegs &L, const LinkMatC &m, const float *sd, const float *ss, const V3 &posN,
                                                   const Q4 &qN, const V3 &posP, const Q4 &qP, float dmN, float dmP, LinkOut &o) {
    LinkMid mid;
    link_stage_a(L, posN, qN, posP, qP, mid);
    if (!mid.small) link_stage_large(mid.pos2, mid.angle1, mid.angle2, mid.angle1v, L.rest);
    link_stage_c(L, mid, m, sd, ss, dmN, dmP, o);
}

// ------------------------------------------------------------------ voxel
struct VoxRegs {
    V3 pos, linMom, angMom;
    Q4 orient;
    int flags;
};

// VX3_Voxel::floorForce (VX3_Voxel.cu:399-426)
__device__ __forceinline__ void voxel_floor_force(VoxRegs &v, const VoxMatC &m, float tempe, V3 &F) {
    const float CurPenetration = floor_penetration(m, tempe, v.pos.z);
    if (CurPenetration >= 0) {
        const V3 vel = v.linMom * m.massInverse;
        const V3 horizontalVel(vel.x, vel.y, 0);
        const float normalForce = m.penStiff * CurPenetration;
        F.z += normalForce - m.colDampT * vel.z;
        if (v.flags & VX3_VOX_FLOOR_STATIC_FRICTION) {
            const float surfaceForceSq = (float)(F.x * F.x + F.y * F.y);
            const float frictionForceSq = (m.muStatic * normalForce) * (m.muStatic * normalForce);
            if (surfaceForceSq > frictionForceSq) v.flags &= ~VX3_VOX_FLOOR_STATIC_FRICTION;
        } else {
            F -= m.muKinetic * normalForce * horizontalVel.Normalized();
        }
    } else
        v.flags &= ~VX3_VOX_FLOOR_STATIC_FRICTION;
}

// VX3_Voxel::timeStep (VX3_Voxel.cu:162-277) for dt != 0, in the three parts its data flow falls into, so that a kernel
// can run them in different warps (the on-chip persistent kernel does; the streaming kernel calls voxel_time_step):
//   voxel_step_translate  force() + force field + floor + linear integration + translational DOF fixes (:175-218, :231-244 x/y/z)
//                         reads the OLD orientation; owns pos, linMom, flags
//   voxel_step_rotate     moment() + angular integration + rotational DOF fixes (:220-256); owns orient, angMom
//   voxel_step_join       on the floor in static friction -> angMom = 0 (:259-264): needs the translate part's new pos.z / flags
// linkF/linkM = sums of the incident links' forces and moments in the voxel's local frame (force()/moment() :350-397);
// contact/ciliaF = pending contactForce and CiliaForce*mat->Cilia; ff = force-field value at the pre-step position.
__device__ __forceinline__ bool voxel_fixed_all(const ExtC *ext) { return ext && (ext->dof & 0x3F) == 0x3F; }

__device__ __forceinline__ void voxel_step_translate(V3 &pos, V3 &linMom, int &flags, const Q4 &orient, const VoxMatC &m, const ExtC *ext, int ix,
                                                     int iy, int iz, float tempe, const V3 &linkF, const V3 &contact, const V3 &ciliaF,
                                                     const V3 &ff, double dt) {
    VoxRegs v; // the floor helpers work on a VoxRegs view
    v.pos = pos; v.linMom = linMom; v.flags = flags; v.orient = orient;
    const bool floorEnabled = (v.flags & VX3_VOX_FLOOR_ENABLED) != 0;
    if (voxel_fixed_all(ext)) {
        const double s = m.nomSize;
        pos = V3(ix * s, iy * s, iz * s) + V3(ext->translation[0], ext->translation[1], ext->translation[2]);
        linMom = V3();
        return;
    }
    // force()
    V3 curForce = orient.RotateVec3D(linkF);
    if (ext) curForce += V3(ext->force[0], ext->force[1], ext->force[2]);
    curForce -= (v.linMom * m.massInverse) * m.globalDampT;
    curForce.z += m.gravityForce;
    curForce -= contact;
    curForce += ciliaF;
    curForce.x += ff.x;
    curForce.y += ff.y;
    curForce.z += ff.z;
    V3 fricForce = curForce;
    if (floorEnabled) voxel_floor_force(v, m, tempe, curForce);
    fricForce = curForce - fricForce;
    v.linMom += curForce * dt;
    V3 translate = v.linMom * (dt * m.massInverse);
    if (floorEnabled && floor_penetration(m, tempe, v.pos.z) >= 0) {
        const double work = fricForce.x * translate.x + fricForce.y * translate.y;
        const double hKe = 0.5 * m.massInverse * (v.linMom.x * v.linMom.x + v.linMom.y * v.linMom.y);
        if (hKe + work <= 0) v.flags |= VX3_VOX_FLOOR_STATIC_FRICTION;
        if (v.flags & VX3_VOX_FLOOR_STATIC_FRICTION) {
            v.linMom.x = v.linMom.y = 0;
            translate.x = translate.y = 0;
        }
    } else
        v.flags &= ~VX3_VOX_FLOOR_STATIC_FRICTION;
    v.pos += translate;
    if (ext) {
        const int dof = ext->dof;
        const double size = m.nomSize;
        if (dof & VX3_DOF_X_TRANSLATE) { v.pos.x = ix * size + ext->translation[0]; v.linMom.x = 0; }
        if (dof & VX3_DOF_Y_TRANSLATE) { v.pos.y = iy * size + ext->translation[1]; v.linMom.y = 0; }
        if (dof & VX3_DOF_Z_TRANSLATE) { v.pos.z = iz * size + ext->translation[2]; v.linMom.z = 0; }
    }
    pos = v.pos;
    linMom = v.linMom;
    flags = v.flags;
}

__device__ __forceinline__ void voxel_step_rotate(Q4 &orient, V3 &angMom, const VoxMatC &m, const ExtC *ext, const V3 &linkM, double dt) {
    if (voxel_fixed_all(ext)) {
        orient = Q4(ext->rotq[0], ext->rotq[1], ext->rotq[2], ext->rotq[3]);
        angMom = V3();
        return;
    }
    // moment()
    V3 curMoment = orient.RotateVec3D(linkM);
    if (ext) curMoment += V3(ext->moment[0], ext->moment[1], ext->moment[2]);
    curMoment -= (angMom * m.momentInertiaInverse) * m.globalDampR;
    angMom += curMoment * dt;
    orient = Q4(angMom * (dt * m.momentInertiaInverse)) * orient;
    if (ext) {
        const int dof = ext->dof;
        const int rot = dof & (VX3_DOF_X_ROTATE | VX3_DOF_Y_ROTATE | VX3_DOF_Z_ROTATE);
        if (rot) {
            if (rot == (VX3_DOF_X_ROTATE | VX3_DOF_Y_ROTATE | VX3_DOF_Z_ROTATE)) {
                orient = Q4(ext->rotq[0], ext->rotq[1], ext->rotq[2], ext->rotq[3]);
                angMom = V3();
            } else {
                V3 tmpRotVec = orient.ToRotationVector();
                if (dof & VX3_DOF_X_ROTATE) { tmpRotVec.x = 0; angMom.x = 0; }
                if (dof & VX3_DOF_Y_ROTATE) { tmpRotVec.y = 0; angMom.y = 0; }
                if (dof & VX3_DOF_Z_ROTATE) { tmpRotVec.z = 0; angMom.z = 0; }
                orient.FromRotationVector(tmpRotVec);
            }
        }
    }
}

// true = the voxel rests on the floor in static friction after this step: its angular momentum is cleared (VX3_Voxel.cu:259-264)
__device__ __forceinline__ bool voxel_step_join(const V3 &pos, int flags, const VoxMatC &m, const ExtC *ext, float tempe) {
    if (voxel_fixed_all(ext)) return false; // timeStep returned before (:166-173)
    const bool floorEnabled = (flags & VX3_VOX_FLOOR_ENABLED) != 0;
    return floorEnabled && floor_penetration(m, tempe, pos.z) >= 0 && (flags & VX3_VOX_FLOOR_STATIC_FRICTION);
}

__device__ __forceinline__ void voxel_time_step(VoxRegs &v, const VoxMatC &m, const ExtC *ext, int ix, int iy, int iz, float tempe,
                                                const V3 &linkF, const V3 &linkM, const V3 &contact, const V3 &ciliaF, const V3 &ff,
                                                double dt) {
    const Q4 orient0 = v.orient;
    voxel_step_translate(v.pos, v.linMom, v.flags, orient0, m, ext, ix, iy, iz, tempe, linkF, contact, ciliaF, ff, dt);
    voxel_step_rotate(v.orient, v.angMom, m, ext, linkM, dt);
    if (voxel_step_join(v.pos, v.flags, m, ext, tempe)) v.angMom = V3(0, 0, 0);
}

} // namespace vx3
