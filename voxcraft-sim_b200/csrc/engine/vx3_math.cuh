// Device/host 3-vector and quaternion arithmetic for the step kernels.
//
// Formulas follow the reference's templates so results agree operation for operation:
//   src/Utils/VX3_Vec3D.h  (operator/ multiplies by the reciprocal :59-63, Normalized :94, NormalizeFast :83)
//   src/Utils/VX3_Quat3D.h (operator* :196-201, ToRotationVector :344-359, FromRotationVector :361-377,
//                           FromAngleToPosX :384-432, RotateVec3D :434-443, RotateVec3DInv :459-469)
// Everything is fp64; the fp32 casts of the reference live at the call sites (vx3_physics.cuh).
#pragma once
#include <cmath>

#ifdef __CUDACC__
#define VXHD __host__ __device__ __forceinline__
#else
#define VXHD inline
#endif

namespace vx3 {

struct V3 {
    double x, y, z;
    VXHD V3() : x(0), y(0), z(0) {}
    VXHD V3(double a, double b, double c) : x(a), y(b), z(c) {}
    VXHD V3 operator+(const V3 &v) const { return V3(x + v.x, y + v.y, z + v.z); }
    VXHD V3 operator-(const V3 &v) const { return V3(x - v.x, y - v.y, z - v.z); }
    VXHD V3 operator-() const { return V3(-x, -y, -z); }
    VXHD V3 operator*(double f) const { return V3(f * x, f * y, f * z); }
    VXHD V3 operator/(double f) const {
        double inv = 1.0 / f;
        return V3(inv * x, inv * y, inv * z);
    }
    VXHD V3 &operator+=(const V3 &v) { x += v.x; y += v.y; z += v.z; return *this; }
    VXHD V3 &operator-=(const V3 &v) { x -= v.x; y -= v.y; z -= v.z; return *this; }
    VXHD bool operator==(const V3 &v) const { return x == v.x && y == v.y && z == v.z; }
    VXHD double Dot(const V3 &v) const { return x * v.x + y * v.y + z * v.z; }
    VXHD double Length2() const { return x * x + y * y + z * z; }
    VXHD double Length() const { return sqrt(x * x + y * y + z * z); }
    VXHD double Dist2(const V3 &v) const { return (v.x - x) * (v.x - x) + (v.y - y) * (v.y - y) + (v.z - z) * (v.z - z); }
    VXHD double Dist(const V3 &v) const { return sqrt(Dist2(v)); }
    VXHD V3 Normalized() const {
        double l = sqrt(x * x + y * y + z * z);
        return l > 0 ? (*this) / l : (*this);
    }
    VXHD void NormalizeFast() {
        double l = sqrt(x * x + y * y + z * z);
        if (l > 0) {
            double li = 1.0 / l;
            x *= li; y *= li; z *= li;
        }
    }
    VXHD V3 Abs() const { return V3(x >= 0 ? x : -x, y >= 0 ? y : -y, z >= 0 ? z : -z); }
};
VXHD V3 operator*(double f, const V3 &v) { return v * f; }

#define VX3_Q_PI 3.14159265358979
#define VX3_DBL_EPSILONx24 5.328e-15
#define VX3_DISCARD_ANGLE_RAD 1e-7
#define VX3_SMALL_ANGLE_RAD 1.732e-2
#define VX3_SLTHRESH_ACOS2SQRT 2.4e-3

// sin and cos of one argument: on the device one libdevice call (shared argument reduction; same results as sin() and cos(),
// one dependent chain instead of two on the latency-bound paths that call it)
VXHD void sin_cos(double a, double &s, double &c) {
#ifdef __CUDA_ARCH__
    sincos(a, &s, &c);
#else
    s = sin(a);
    c = cos(a);
#endif
}

struct Q4 {
    double w, x, y, z;
    VXHD Q4() : w(1), x(0), y(0), z(0) {}
    VXHD Q4(double a, double b, double c, double d) : w(a), x(b), y(c), z(d) {}
    VXHD explicit Q4(const V3 &rv) { FromRotationVector(rv); }
    VXHD Q4 operator*(const Q4 &f) const {
        return Q4(w * f.w - x * f.x - y * f.y - z * f.z, w * f.x + x * f.w + y * f.z - z * f.y,
                  w * f.y - x * f.z + y * f.w + z * f.x, w * f.z + x * f.y - y * f.x + z * f.w);
    }
    VXHD Q4 Conjugate() const { return Q4(w, -x, -y, -z); }
    VXHD double Angle() const { return 2.0 * acos(w > 1 ? 1 : w); }
    VXHD V3 ToRotationVector() const {
        if (w >= 1.0 || w <= -1.0) return V3(0, 0, 0);
        double squareLength = 1.0 - w * w;
        if (squareLength < VX3_SLTHRESH_ACOS2SQRT) return V3(x, y, z) * 2.0 * sqrt((2 - 2 * w) / squareLength);
        else return V3(x, y, z) * 2.0 * acos(w) / sqrt(squareLength);
    }
    VXHD void FromRotationVector(const V3 &VecIn) {
        V3 theta = VecIn / 2;
        double s, thetaMag2 = theta.Length2();
        if (thetaMag2 * thetaMag2 < VX3_DBL_EPSILONx24) {
            w = 1.0 - 0.5 * thetaMag2;
            s = 1.0 - thetaMag2 / 6.0;
        } else {
            double thetaMag = sqrt(thetaMag2), sn, cs;
            sin_cos(thetaMag, sn, cs);
            w = cs;
            s = sn / thetaMag;
        }
        x = theta.x * s;
        y = theta.y * s;
        z = theta.z * s;
    }
    // len_out (optional): |RotateFrom| when the general branch computed it for the normalisation (the caller needs the same
    // sqrt(x*x + y*y + z*z) right after, VX3_Link.cu:119: one IEEE sqrt sequence instead of two on the large-angle chain);
    // returns whether it did
    VXHD bool FromAngleToPosX(const V3 &RotateFrom, double *len_out = nullptr) {
        if (V3(0, 0, 0) == RotateFrom) return false;
        double YoverX = RotateFrom.y / RotateFrom.x;
        double ZoverX = RotateFrom.z / RotateFrom.x;
        if (YoverX < VX3_SMALL_ANGLE_RAD && YoverX > -VX3_SMALL_ANGLE_RAD && ZoverX < VX3_SMALL_ANGLE_RAD && ZoverX > -VX3_SMALL_ANGLE_RAD) {
            x = 0;
            y = 0.5 * ZoverX;
            z = -0.5 * YoverX;
            w = 1 + 0.5 * (-y * y - z * z);
            return false;
        }
        V3 RotFromNorm = RotateFrom;
        { // NormalizeFast(), keeping the length
            const double l = sqrt(RotFromNorm.x * RotFromNorm.x + RotFromNorm.y * RotFromNorm.y + RotFromNorm.z * RotFromNorm.z);
            if (len_out) *len_out = l;
            if (l > 0) {
                const double li = 1.0 / l;
                RotFromNorm.x *= li; RotFromNorm.y *= li; RotFromNorm.z *= li;
            }
        }
        double theta = acos(RotFromNorm.x);
        if (theta > VX3_Q_PI - VX3_DISCARD_ANGLE_RAD) {
            w = 0; x = 0; y = 1; z = 0;
            return true;
        }
        const double AxisMagInv = 1.0 / sqrt(RotFromNorm.z * RotFromNorm.z + RotFromNorm.y * RotFromNorm.y);
        const double a = 0.5 * theta;
        double s, c;
        sin_cos(a, s, c);
        w = c;
        x = 0;
        y = RotFromNorm.z * AxisMagInv * s;
        z = -RotFromNorm.y * AxisMagInv * s;
        return true;
    }
    VXHD V3 RotateVec3D(const V3 &f) const {
        double fx = f.x, fy = f.y, fz = f.z;
        double tw = fx * x + fy * y + fz * z;
        double tx = fx * w - fy * z + fz * y;
        double ty = fx * z + fy * w - fz * x;
        double tz = -fx * y + fy * x + fz * w;
        return V3(w * tx + x * tw + y * tz - z * ty, w * ty - x * tz + y * tw + z * tx, w * tz + x * ty - y * tx + z * tw);
    }
    VXHD V3 RotateVec3DInv(const V3 &f) const {
        double fx = f.x, fy = f.y, fz = f.z;
        double tw = x * fx + y * fy + z * fz;
        double tx = w * fx - y * fz + z * fy;
        double ty = w * fy + x * fz - z * fx;
        double tz = w * fz - x * fy + y * fx;
        return V3(tw * x + tx * w + ty * z - tz * y, tw * y - tx * z + ty * w + tz * x, tw * z + tx * y - ty * x + tz * w);
    }
};

// VX3_Link.h:135-174: rotate a vector / quaternion between the link's axis and the X axis
VXHD V3 toAxisX(int axis, const V3 &v) {
    switch (axis) {
    case 1: return V3(v.y, -v.x, v.z);
    case 2: return V3(v.z, v.y, -v.x);
    default: return v;
    }
}
VXHD Q4 toAxisX(int axis, const Q4 &q) {
    switch (axis) {
    case 1: return Q4(q.w, q.y, -q.x, q.z);
    case 2: return Q4(q.w, q.z, q.y, -q.x);
    default: return q;
    }
}
VXHD V3 toAxisOriginal(int axis, const V3 &v) {
    switch (axis) {
    case 1: return V3(-v.y, v.x, v.z);
    case 2: return V3(-v.z, v.y, v.x);
    default: return v;
    }
}

} // namespace vx3
