// See vx3_materials.h.  Arithmetic is kept cast-for-cast with the reference host library
// (float members, double where the reference promotes) — do not "simplify" expressions.
#include "vx3_materials.h"

namespace vx3 {

VoxelMat::VoxelMat(float youngsModulus, float density, double nominalSize) {
    // CVX_Material(float,float) src/old/VX_Material.cpp:14-21 then
    // CVX_MaterialVoxel::initialize src/old/VX_MaterialVoxel.cpp:29-34
    clear();
    rho = density;
    setModelLinear(youngsModulus);
    updateDerived();
    nomSize = nominalSize;
    gravMult = 0.0f;
    updateDerived();
}

void VoxelMat::clear() { // src/old/VX_Material.cpp:71-90
    r = g = b = a = -1;
    nu = 0.0f;
    rho = 1.0f;
    alphaCTE = 0.0f;
    muStatic = 0.0f;
    muKinetic = 0.0f;
    zetaInternal = 1.0f;
    zetaGlobal = 0.0f;
    zetaCollision = 0.0f;
    extScale[0] = extScale[1] = extScale[2] = 1.0;
    setModelLinear(1.0);
    updateDerived();
}

float VoxelMat::stress(float strain, float transverseStrainSum, bool forceLinear) const {
    // src/old/VX_Material.cpp stress(): host twin of src/VX3/VX3_Material.cu:90-124 (no duplicated 0)
    if (isFailed(strain)) return 0.0f;
    if (strain <= strainData[1] || linear || forceLinear) {
        if (nu == 0.0f) return E * strain;
        else return eHat * ((1 - nu) * strain + nu * transverseStrainSum);
    }
    int DataCount = (int)strainData.size();
    for (int i = 2; i < DataCount; i++) {
        if (strain <= strainData[i] || i == DataCount - 1) {
            float Perc = (strain - strainData[i - 1]) / (strainData[i] - strainData[i - 1]);
            float basicStress = stressData[i - 1] + Perc * (stressData[i] - stressData[i - 1]);
            if (nu == 0.0f) return basicStress;
            else {
                float modulus = (stressData[i] - stressData[i - 1]) / (strainData[i] - strainData[i - 1]);
                float modulusHat = modulus / ((1 - 2 * nu) * (1 + nu));
                float effectiveStrain = basicStress / modulus;
                float effectiveTransverseStrainSum = transverseStrainSum * (effectiveStrain / strain);
                return modulusHat * ((1 - nu) * effectiveStrain + nu * effectiveTransverseStrainSum);
            }
        }
    }
    return 0.0f;
}

float VoxelMat::modulus(float strain) const { // src/old/VX_Material.cpp:231-243
    if (isFailed(strain)) return 0.0f;
    if (strain <= strainData[1] || linear) return E;
    int DataCount = (int)strainData.size();
    for (int i = 2; i < DataCount; i++) {
        if (strain <= strainData[i] || i == DataCount - 1)
            return (stressData[i] - stressData[i - 1]) / (strainData[i] - strainData[i - 1]);
    }
    return 0.0f;
}

float VoxelMat::strainAt(float stress) const { // CVX_Material::strain(float stress)
    if (stress <= stressData[1] || linear) return stress / E;
    int DataCount = (int)strainData.size();
    for (int i = 2; i < DataCount; i++) {
        if (stress <= stressData[i] || i == DataCount - 1) {
            float Perc = (stress - stressData[i - 1]) / (stressData[i] - stressData[i - 1]);
            return strainData[i - 1] + Perc * (strainData[i] - strainData[i - 1]);
        }
    }
    return 0.0f;
}

static int clamp255(int v) { return v > 255 ? 255 : (v < 0 ? 0 : v); }
void VoxelMat::setColor(int red, int green, int blue, int alpha) { // src/old/VX_Material.cpp:245-280
    r = clamp255(red);
    g = clamp255(green);
    b = clamp255(blue);
    a = clamp255(alpha);
}

bool VoxelMat::setModel(int dataPointCount, const float *pStrainValues, const float *pStressValues) {
    // src/old/VX_Material.cpp:300-367
    if (dataPointCount > 0 && *pStrainValues == 0 && *pStressValues == 0) {
        pStrainValues++;
        pStressValues++;
        dataPointCount--;
    }
    if (dataPointCount <= 0) return false;
    if (*pStrainValues <= 0 || *pStressValues <= 0) return false;

    std::vector<float> tmpStrainData, tmpStressData;
    tmpStrainData.push_back(0);
    tmpStressData.push_back(0);
    float sweepStrain = 0.0f, sweepStress = 0.0f;
    for (int i = 0; i < dataPointCount; i++) {
        float thisStrain = pStrainValues[i];
        float thisStress = pStressValues[i];
        if (thisStrain <= sweepStrain) return false;
        // NB the reference compares against tmpStressData[0]/tmpStrainData[0] = 0/0 = NaN, so this
        // slope check never fires (:336); kept as a no-op for fidelity.
        if (i > 0 && (thisStress - sweepStress) / (thisStrain - sweepStrain) > tmpStressData[0] / tmpStrainData[0]) return false;
        sweepStrain = thisStrain;
        sweepStress = thisStress;
        tmpStrainData.push_back(thisStrain);
        tmpStressData.push_back(thisStress);
    }
    strainData = tmpStrainData;
    stressData = tmpStressData;
    E = stressData[1] / strainData[1];
    sigmaFail = stressData[stressData.size() - 1];
    epsilonFail = strainData[strainData.size() - 1];
    linear = (dataPointCount == 1);
    if (dataPointCount == 1 || dataPointCount == 2) {
        sigmaYield = stressData[1];
        epsilonYield = strainData[1];
    } else {
        setYieldFromData();
    }
    return updateDerived();
}

bool VoxelMat::setModelLinear(float youngsModulus, float failureStress) { // src/old/VX_Material.cpp:372-401
    if (youngsModulus <= 0) return false;
    if (failureStress != -1.0f && failureStress <= 0) return false;
    float tmpfailureStress = failureStress;
    if (tmpfailureStress == -1) tmpfailureStress = 1000000;
    float tmpfailStrain = tmpfailureStress / youngsModulus;
    strainData.clear();
    stressData.clear();
    strainData.push_back(0);
    stressData.push_back(0);
    strainData.push_back(tmpfailStrain);
    stressData.push_back(tmpfailureStress);
    linear = true;
    E = youngsModulus;
    sigmaYield = failureStress;
    sigmaFail = failureStress;
    epsilonYield = (failureStress == -1) ? -1 : tmpfailStrain;
    epsilonFail = (failureStress == -1) ? -1 : tmpfailStrain;
    return updateDerived();
}

bool VoxelMat::setModelBilinear(float youngsModulus, float plasticModulus, float yieldStress, float failureStress) {
    // src/old/VX_Material.cpp:406-450
    if (youngsModulus <= 0) return false;
    if (plasticModulus <= 0 || plasticModulus >= youngsModulus) return false;
    if (yieldStress <= 0) return false;
    if (failureStress != -1.0f && failureStress <= yieldStress) return false;
    float yieldStrain = yieldStress / youngsModulus;
    float tmpfailureStress = failureStress;
    if (tmpfailureStress == -1) tmpfailureStress = 3 * yieldStress;
    float tM = plasticModulus;
    float tB = yieldStress - tM * yieldStrain;
    float tmpfailStrain = (tmpfailureStress - tB) / tM;
    strainData.clear();
    strainData.push_back(0);
    strainData.push_back(yieldStrain);
    strainData.push_back(tmpfailStrain);
    stressData.clear();
    stressData.push_back(0);
    stressData.push_back(yieldStress);
    stressData.push_back(tmpfailureStress);
    linear = false;
    E = youngsModulus;
    sigmaYield = yieldStress;
    sigmaFail = failureStress;
    epsilonYield = yieldStrain;
    epsilonFail = failureStress == -1.0f ? -1.0f : tmpfailStrain;
    return updateDerived();
}

bool VoxelMat::setYieldFromData(float percentStrainOffset) { // src/old/VX_Material.cpp:453-487
    sigmaYield = -1.0f;
    epsilonYield = -1.0f;
    float oM = E;
    float oB = (-percentStrainOffset / 100 * oM);
    int dataPoints = (int)strainData.size() - 1;
    for (int i = 1; i < dataPoints - 1; i++) {
        float x1 = strainData[i], x2 = strainData[i + 1];
        float y1 = stressData[i], y2 = stressData[i + 1];
        float tM = (y2 - y1) / (x2 - x1);
        float tB = y1 - tM * x1;
        if (oM != tM) {
            float xIntersect = (tB - oB) / (oM - tM);
            if (xIntersect > x1 && xIntersect < x2) {
                float percentBetweenPoints = (xIntersect - x1) / (x2 - x1);
                sigmaYield = y1 + percentBetweenPoints * (y2 - y1);
                epsilonYield = xIntersect;
                return true;
            }
        }
    }
    sigmaYield = sigmaFail;
    epsilonYield = epsilonFail;
    return false;
}

void VoxelMat::setPoissonsRatio(float poissonsRatio) { // src/old/VX_Material.cpp:489-495
    if (poissonsRatio < 0) poissonsRatio = 0;
    if (poissonsRatio >= 0.5) poissonsRatio = 0.5 - FLT_EPSILON * 2;
    nu = poissonsRatio;
    updateDerived();
}

void VoxelMat::setDensity(float density) { // src/old/VX_Material.cpp:497-502
    if (density <= 0) density = FLT_MIN;
    rho = density;
    updateDerived();
}

bool VoxelMat::setNominalSize(double size) { // src/old/VX_MaterialVoxel.cpp:82-87
    if (size <= 0) size = FLT_MIN;
    nomSize = size;
    return updateDerived();
}

bool VoxelMat::updateDerived() {
    // CVX_Material::updateDerived src/old/VX_Material.cpp:542-549 (dependents are refreshed by the
    // builder after all voxel materials are final — updateAll is a pure function of the two voxel materials)
    eHat = E / ((1 - 2 * nu) * (1 + nu));
    // CVX_MaterialVoxel::updateDerived src/old/VX_MaterialVoxel.cpp:57-79
    double volume = nomSize * nomSize * nomSize;
    mass = (float)(volume * rho);
    momentInertia = (float)(mass * nomSize * nomSize / 6.0f);
    firstMoment = (float)(mass * nomSize / 2.0f);
    if (volume == 0 || mass == 0 || momentInertia == 0) {
        massInverse = sqrtMass = momentInertiaInverse = c2xSqMxExS = c2xSqIxExSxSxS = 0.0f;
        return false;
    }
    massInverse = 1.0f / mass;
    sqrtMass = sqrtf(mass);
    momentInertiaInverse = 1.0f / momentInertia;
    c2xSqMxExS = (float)(2.0f * sqrt(mass * E * nomSize));
    c2xSqIxExSxSxS = (float)(2.0f * sqrt(momentInertia * E * nomSize * nomSize * nomSize));
    return true;
}

bool LinkMat::updateAll(const VoxelMat &m1, const VoxelMat &m2) { // src/old/VX_MaterialLink.cpp:45-118
    nomSize = 0.5 * (m1.nomSize + m2.nomSize);
    r = (int)(0.5 * (m1.r + m2.r));
    g = (int)(0.5 * (m1.g + m2.g));
    b = (int)(0.5 * (m1.b + m2.b));
    a = (int)(0.5 * (m1.a + m2.a));
    rho = 0.5f * (m1.rho + m2.rho);
    alphaCTE = 0.5f * (m1.alphaCTE + m2.alphaCTE);
    muStatic = 0.5f * (m1.muStatic + m2.muStatic);
    muKinetic = 0.5f * (m1.muKinetic + m2.muKinetic);
    zetaInternal = 0.5f * (m1.zetaInternal + m2.zetaInternal);
    zetaGlobal = 0.5f * (m1.zetaGlobal + m2.zetaGlobal);
    zetaCollision = 0.5f * (m1.zetaCollision + m2.zetaCollision);
    extScale[0] = extScale[1] = extScale[2] = 1.0;

    float stressFail = -1.0f, f1 = m1.sigmaFail, f2 = m2.sigmaFail;
    if (f1 == -1.0f) stressFail = f2;
    else if (f2 == -1.0f) stressFail = f1;
    else stressFail = f1 < f2 ? f1 : f2;

    if (m1.linear && m2.linear) setModelLinear(2.0f * m1.E * m2.E / (m1.E + m2.E), stressFail);
    else {
        std::vector<float> newStressValues, newStrainValues;
        newStressValues.push_back(0.0f);
        newStrainValues.push_back(0.0f);
        int dataIt1 = 1, dataIt2 = 1;
        while (dataIt1 < (int)m1.strainData.size() && dataIt2 < (int)m2.strainData.size()) {
            float strain = FLT_MAX;
            if (dataIt1 < (int)m1.strainData.size()) strain = m1.strainData[dataIt1];
            if (dataIt2 < (int)m2.strainData.size() && m2.strainData[dataIt2] < strain) strain = m2.strainData[dataIt2];
            if (strain == m1.strainData[dataIt1]) dataIt1++;
            if (strain == m2.strainData[dataIt2]) dataIt2++;
            float modulus1 = m1.modulus(strain - FLT_EPSILON);
            float modulus2 = m2.modulus(strain - FLT_EPSILON);
            float thisModulus = 2.0f * modulus1 * modulus2 / (modulus1 + modulus2);
            int lastDataIndex = (int)newStrainValues.size() - 1;
            newStrainValues.push_back(strain);
            newStressValues.push_back(newStressValues[lastDataIndex] + thisModulus * (strain - newStrainValues[lastDataIndex]));
        }
        setModel((int)newStrainValues.size(), &newStrainValues[0], &newStressValues[0]);
        sigmaFail = stressFail;
        epsilonFail = stressFail == -1.0f ? -1.0f : strainAt(stressFail);
    }

    if (m1.nu == 0 && m2.nu == 0) nu = 0;
    else {
        float tmpEHat = 2 * m1.eHat * m2.eHat / (m1.eHat + m2.eHat);
        float tmpE = E;
        float c2 = (tmpEHat - tmpE) / (2 * tmpEHat) + 0.0625;
        nu = sqrt(c2) - 0.25;
    }
    return updateDerived();
}

bool LinkMat::updateDerived() { // src/old/VX_MaterialLink.cpp:120-141
    VoxelMat::updateDerived();
    float L = (float)nomSize;
    a1 = E * L;
    a2 = E * L * L * L / (12.0f * (1 + nu));
    b1 = E * L;
    b2 = E * L * L / 2.0f;
    b3 = E * L * L * L / 6.0f;
    sqA1 = sqrt(a1);
    sqA2xIp = sqrt(a2 * L * L / 6.0f);
    sqB1 = sqrt(b1);
    sqB2xFMp = sqrt(b2 * L / 2.0f);
    sqB3xIp = sqrt(b3 * L * L / 6.0f);
    return true;
}

} // namespace vx3
