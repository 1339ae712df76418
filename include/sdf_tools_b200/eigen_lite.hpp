// include/sdf_tools_b200/eigen_lite.hpp -- the handful of Eigen types the SDF-build subset of the sdf_tools API
// mentions (Isometry3d, Vector3d, Vector4d, Matrix4d).  Real Eigen is used when it is installed; this stand-in
// only exists because the build image has none.  Row-major 4x4 homogeneous transform, doubles.
#pragma once
#if __has_include(<Eigen/Geometry>)
#include <Eigen/Geometry>
#define SDF_B200_HAVE_EIGEN 1
#else
#define SDF_B200_HAVE_EIGEN 0
#include <array>
#include <cmath>
#include <cstddef>

namespace Eigen {

struct Vector3d {
    double v[3]{0, 0, 0};
    Vector3d() = default;
    Vector3d(double x, double y, double z) : v{x, y, z} {}
    double x() const { return v[0]; }
    double y() const { return v[1]; }
    double z() const { return v[2]; }
    double operator()(int i) const { return v[i]; }
    double& operator()(int i) { return v[i]; }
};

struct Vector4d {
    double v[4]{0, 0, 0, 0};
    Vector4d() = default;
    Vector4d(double x, double y, double z, double w) : v{x, y, z, w} {}
    double operator()(int i) const { return v[i]; }
    double& operator()(int i) { return v[i]; }
};

struct Matrix4d {
    double m[4][4]{};
    double operator()(int r, int c) const { return m[r][c]; }
    double& operator()(int r, int c) { return m[r][c]; }
    static Matrix4d Identity() { Matrix4d I; for (int i = 0; i < 4; ++i) I.m[i][i] = 1.0; return I; }
};

// Rigid transform: rotation R (3x3) + translation t.
class Isometry3d {
    Matrix4d m_ = Matrix4d::Identity();
public:
    Isometry3d() = default;
    explicit Isometry3d(const Matrix4d& m) : m_(m) {}
    static Isometry3d Identity() { return Isometry3d(); }
    const Matrix4d& matrix() const { return m_; }
    Matrix4d& matrix() { return m_; }
    Vector3d translation() const { return Vector3d(m_(0, 3), m_(1, 3), m_(2, 3)); }
    void setTranslation(double x, double y, double z) { m_(0, 3) = x; m_(1, 3) = y; m_(2, 3) = z; }
    Isometry3d inverse() const {            // [R t]^-1 = [R^T  -R^T t]
        Isometry3d r;
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m_(i, j) = m_(j, i);
        for (int i = 0; i < 3; ++i) r.m_(i, 3) = -(r.m_(i, 0) * m_(0, 3) + r.m_(i, 1) * m_(1, 3) + r.m_(i, 2) * m_(2, 3));
        return r;
    }
    Vector4d operator*(const Vector4d& p) const {
        Vector4d o;
        for (int i = 0; i < 4; ++i) o(i) = m_(i, 0) * p(0) + m_(i, 1) * p(1) + m_(i, 2) * p(2) + m_(i, 3) * p(3);
        return o;
    }
    Vector3d operator*(const Vector3d& p) const {
        const Vector4d o = (*this) * Vector4d(p.x(), p.y(), p.z(), 1.0);
        return Vector3d(o(0), o(1), o(2));
    }
    // rotate a direction (no translation)
    Vector3d rotate(const Vector3d& d) const {
        return Vector3d(m_(0, 0) * d.x() + m_(0, 1) * d.y() + m_(0, 2) * d.z(),
                        m_(1, 0) * d.x() + m_(1, 1) * d.y() + m_(1, 2) * d.z(),
                        m_(2, 0) * d.x() + m_(2, 1) * d.y() + m_(2, 2) * d.z());
    }
};

}  // namespace Eigen
#endif
