/* Solver.h -- ref Solver.h:19-67.  Same public surface (BuildLinearSystem, getTransform, getError,
 * SolveJacobianSystem, PrintSystem); the work is one fused reduction kernel + an on-device 6x6 solve
 * instead of a 7.4 MB Jacobian, cublasSgemv, cublasSsyrk and a host Eigen inverse (Solver.cpp:74-111).
 */
#ifndef SOLVER_H
#define SOLVER_H

#include "EigenUtil.h"
#include "SE3.h"
#include "vh/abi.h"

class Solver {
public:
    unsigned int numIters = 10;

    explicit Solver(vh_context* ctx = nullptr);     /* nullptr: creates a small tracking-only context */
    ~Solver();
    Solver(const Solver&) = delete;
    Solver& operator=(const Solver&) = delete;

    /* ref Solver.cpp:48-124: normal equations from the stored correspondences, solve, SE(3) update. */
    void BuildLinearSystem(const float4* d_input, const float4* d_correspondences, const float4* d_correspondenceNormals,
                           const float* d_residuals, int width, int height);
    void PrintSystem();
    void SolveJacobianSystem(const Matrix6x6f& JTJ, const Vector6f& JTr);   /* ref :126-139, host LDLT variant */
    Matrix4x4f getTransform();                      /* ref Solver.h:32: SE3Exp(estimate) */
    double getError() { return TotalError; }
    vh_context* context() const { return ctx_; }
    void setStream(vh_stream s) { stream_ = s; }

private:
    vh_context* ctx_;
    bool ownsCtx_;
    vh_stream stream_;
    vh_icp_system* d_system_;
    Vector6f update, estimate;
    bool solution_exists = false;
    double TotalError = 0.0;
};

#endif
