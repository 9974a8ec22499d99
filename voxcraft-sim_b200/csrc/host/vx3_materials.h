// Host-side material model: computes every derived constant the step loop reads, in the
// same arithmetic (float/double mix, operation order) as the reference host library so
// the flat model equals one exported from a CVX_Sim bit-for-bit.
//
// Follows: src/old/VX_Material.cpp (model setup :300-470, setters :489-540, updateDerived :542-549),
//          src/old/VX_MaterialVoxel.cpp:57-87, src/old/VX_MaterialLink.cpp:45-141.
#pragma once
#include <cfloat>
#include <cmath>
#include <string>
#include <vector>

namespace vx3 {

struct VoxelMat {
    // identity / VX3 additions (src/old/VX_Material.h, copied by CVX_Sim::CopyMat, src/VXA/VX_Sim.cpp:368-420)
    int matid = 0;
    bool fixed = false, sticky = false, isTarget = false, isPaceMaker = false, isElectricalActive = false;
    int isMeasured = 1;
    double Cilia = 0, PaceMakerPeriod = 0, signalValueDecay = 0.9, signalTimeDelay = 0.0, inactivePeriod = 0.05;
    double RemoveAfter = 0, ThermalOnAfter = 0, CiliaOnAfter = 0;
    int r = -1, g = -1, b = -1, a = -1;
    // CVX_Material
    bool linear = true;
    float E = 0, sigmaYield = 0, sigmaFail = 0, epsilonYield = 0, epsilonFail = 0;
    std::vector<float> strainData, stressData;
    float nu = 0, rho = 1, alphaCTE = 0, muStatic = 0, muKinetic = 0;
    float zetaInternal = 1, zetaGlobal = 0, zetaCollision = 0;
    double extScale[3] = {1, 1, 1};
    float eHat = 0;
    // CVX_MaterialVoxel
    double nomSize = 0.001;
    float gravMult = 0;
    float mass = 0, massInverse = 0, sqrtMass = 0, firstMoment = 0, momentInertia = 0, momentInertiaInverse = 0;
    float c2xSqMxExS = 0, c2xSqIxExSxSxS = 0;

    VoxelMat() { clear(); }
    VoxelMat(float youngsModulus, float density, double nominalSize);
    virtual ~VoxelMat() {}

    void clear();
    bool setModel(int n, const float *strain, const float *stress);
    bool setModelLinear(float youngsModulus, float failureStress = -1);
    bool setModelBilinear(float youngsModulus, float plasticModulus, float yieldStress, float failureStress = -1);
    bool setYieldFromData(float percentStrainOffset = 0.2f);
    void setColor(int red, int green, int blue, int alpha);
    void setPoissonsRatio(float v);
    void setDensity(float v);
    void setStaticFriction(float v) { muStatic = v <= 0 ? 0 : v; }
    void setKineticFriction(float v) { muKinetic = v <= 0 ? 0 : v; }
    void setInternalDamping(float z) { zetaInternal = z <= 0 ? 0 : z; }
    void setGlobalDamping(float z) { zetaGlobal = z <= 0 ? 0 : z; }
    void setCollisionDamping(float z) { zetaCollision = z <= 0 ? 0 : z; }
    bool setNominalSize(double size);

    bool isFailed(float strain) const { return epsilonFail != -1.0f && strain > epsilonFail; }
    float stress(float strain, float transverseStrainSum = 0.0f, bool forceLinear = false) const;
    float modulus(float strain) const;
    float strainAt(float stress) const;

    virtual bool updateDerived(); // material + voxel-material level
};

struct LinkMat : VoxelMat {
    LinkMat() { isMeasured = 0; } // CVX_Material member default (src/old/VX_Material.h:176); never set for link materials
    int vox1 = -1, vox2 = -1; // indices into the voxel material table
    float a1 = 0, a2 = 0, b1 = 0, b2 = 0, b3 = 0, sqA1 = 0, sqA2xIp = 0, sqB1 = 0, sqB2xFMp = 0, sqB3xIp = 0;
    bool updateAll(const VoxelMat &m1, const VoxelMat &m2);
    bool updateDerived() override;
};

} // namespace vx3
