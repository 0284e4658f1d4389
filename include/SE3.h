/* SE3.h -- ref SE3.h:6-8.  Closed-form fp64 exponential / logarithm of the twist (v, omega)
 * (the reference calls Eigen's matrix exp()/log(), SE3.cpp:4-19). */
#ifndef SE3_H
#define SE3_H
#include "EigenUtil.h"

Matrix4x4f SE3Exp(const Vector6f& twist);
Vector6f SE3Log(const Matrix4x4f& transform);
Vector6f updateTransform(const Vector6f& perturbation, const Vector6f prev_estimate);

#endif
