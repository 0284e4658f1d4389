/* EigenUtil.h -- minimal stand-ins for the Eigen types on the reference's tracking interface
 * (ref EigenUtil.h:10-16).  Eigen is not available (and not needed: the 6x6 solve and the SE(3)
 * exponential run on the device).  Column-major storage like Eigen, so .data() round-trips through
 * code written for Eigen::Matrix4f (ref Application.cpp:76,81 hand it to glm / float4x4).
 */
#ifndef VH_EIGEN_UTIL_H
#define VH_EIGEN_UTIL_H

template <int R, int Ccols>
struct VhMatrix {
    float v[R * Ccols];                                          /* column-major */
    float* data() { return v; }
    const float* data() const { return v; }
    float& operator()(int r, int c) { return v[c * R + r]; }
    float operator()(int r, int c) const { return v[c * R + r]; }
    float& operator()(int i) { return v[i]; }
    float operator()(int i) const { return v[i]; }
    float& operator[](int i) { return v[i]; }
    float operator[](int i) const { return v[i]; }
    void setZero() { for (float& x : v) x = 0.f; }
    void setIdentity() { setZero(); for (int i = 0; i < (R < Ccols ? R : Ccols); ++i) (*this)(i, i) = 1.f; }
    static VhMatrix Identity() { VhMatrix m; m.setIdentity(); return m; }
    static VhMatrix Zero() { VhMatrix m; m.setZero(); return m; }
    static constexpr int rows() { return R; }
    static constexpr int cols() { return Ccols; }
};

using Matrix6x6f = VhMatrix<6, 6>;
using Matrix4x4f = VhMatrix<4, 4>;
using Matrix3x3f = VhMatrix<3, 3>;
using Vector6f = VhMatrix<6, 1>;
using Vector4f = VhMatrix<4, 1>;

#endif
