// host_classes.cpp -- SDF_Hashtable / CameraTracking / Solver / SE3 host classes on top of the handle API.
// Mirrors the sequencing of ref SDF_Hashtable.cpp:11-89, CameraTracking.cpp:26-69,117-143,
// Solver.cpp:48-124,158-200 and SE3.cpp:4-26 without their OpenGL, Eigen, cuBLAS and console I/O.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <cuda_runtime.h>

#include "CameraTracking.h"
#include "SDF_Hashtable.h"
#include "SE3.h"
#include "Solver.h"

namespace {

[[noreturn]] void die(const char* what) {   // the reference's error contract: print, reset, exit (helper_cuda.h:965-981)
    std::fprintf(stderr, "vh host error: %s: %s\n", what, vh_last_error());
    cudaDeviceReset();
    std::exit(EXIT_FAILURE);
}
#define VH_MUST(expr) do { if ((expr) != VH_OK) die(#expr); } while (0)

void soTerms(const double* w, double& A, double& B, double& C) {
    double t2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], th = std::sqrt(t2);
    if (th < 1e-6) { A = 1.0 - t2 / 6.0; B = 0.5 - t2 / 24.0; C = 1.0 / 6.0 - t2 / 120.0; }
    else { A = std::sin(th) / th; B = (1.0 - std::cos(th)) / t2; C = (th - std::sin(th)) / (t2 * th); }
}
void hat2(const double* w, double* K, double* K2) {
    const double k[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
    for (int i = 0; i < 9; ++i) K[i] = k[i];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) K2[i * 3 + j] = K[i * 3] * K[j] + K[i * 3 + 1] * K[3 + j] + K[i * 3 + 2] * K[6 + j];
}

vh_context* makeTrackingContext(int w, int h) {
    vh_config cfg;
    vh_default_config(&cfg);
    cfg.width = w; cfg.height = h;
    cfg.table.numBuckets = 16; cfg.table.numVoxelBlocks = 16;     // tracking needs no table
    vh_context* c = nullptr;
    VH_MUST(vh_create(&cfg, &c));
    return c;
}

}  // namespace

// ---- SE(3) -------------------------------------------------------------------------------------------
Matrix4x4f SE3Exp(const Vector6f& twist) {                        // ref SE3.cpp:4-11
    double v[3] = {twist(0), twist(1), twist(2)}, w[3] = {twist(3), twist(4), twist(5)};
    double A, B, C, K[9], K2[9];
    soTerms(w, A, B, C);
    hat2(w, K, K2);
    Matrix4x4f M = Matrix4x4f::Identity();
    for (int i = 0; i < 3; ++i) {
        double t = 0;
        for (int j = 0; j < 3; ++j) {
            double I = i == j ? 1.0 : 0.0;
            M(i, j) = (float)(I + A * K[i * 3 + j] + B * K2[i * 3 + j]);
            t += (I + B * K[i * 3 + j] + C * K2[i * 3 + j]) * v[j];
        }
        M(i, 3) = (float)t;
    }
    return M;
}

Vector6f SE3Log(const Matrix4x4f& T) {                            // ref SE3.cpp:14-19
    double tr = (double)T(0, 0) + T(1, 1) + T(2, 2);
    double cs = std::fmin(1.0, std::fmax(-1.0, (tr - 1.0) * 0.5)), th = std::acos(cs);
    double f = th < 1e-6 ? 0.5 + th * th / 12.0 : th / (2.0 * std::sin(th));
    double w[3] = {f * ((double)T(2, 1) - T(1, 2)), f * ((double)T(0, 2) - T(2, 0)), f * ((double)T(1, 0) - T(0, 1))};
    double A, B, C, K[9], K2[9];
    soTerms(w, A, B, C);
    hat2(w, K, K2);
    double t2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    double D = t2 < 1e-12 ? 1.0 / 12.0 + t2 / 720.0 : (1.0 - A / (2.0 * B)) / t2;
    Vector6f tw;
    for (int i = 0; i < 3; ++i) {
        double s = 0;
        for (int j = 0; j < 3; ++j) s += ((i == j ? 1.0 : 0.0) - 0.5 * K[i * 3 + j] + D * K2[i * 3 + j]) * T(j, 3);
        tw(i) = (float)s;
    }
    tw(3) = (float)w[0]; tw(4) = (float)w[1]; tw(5) = (float)w[2];
    return tw;
}

Vector6f updateTransform(const Vector6f& perturbation, const Vector6f prev_estimate) {   // ref SE3.cpp:24-26
    Matrix4x4f a = SE3Exp(perturbation), b = SE3Exp(prev_estimate), p;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += (double)a(i, k) * b(k, j);
            p(i, j) = (float)s;
        }
    return SE3Log(p);
}

// ---- SDF_Hashtable -----------------------------------------------------------------------------------
SDF_Hashtable::SDF_Hashtable() : ctx_(nullptr), stream_(nullptr), synchronous_(true) {   // ref :60-81
    vh_config cfg;
    vh_default_config(&cfg);                       // the 12 constants of common.h:39-50
    h_hashtableParams = cfg.table;
    VH_MUST(vh_create(&cfg, &ctx_));               // updateConstantHashTableParams + deviceAllocate + projection matrix
}

SDF_Hashtable::SDF_Hashtable(const vh_config& cfg) : ctx_(nullptr), stream_(nullptr), synchronous_(true) {
    h_hashtableParams = cfg.table;
    VH_MUST(vh_create(&cfg, &ctx_));
}

SDF_Hashtable::~SDF_Hashtable() { vh_destroy(ctx_); }             // ref :83-89

void SDF_Hashtable::integrate(const float4x4& viewMat, const float4* verts, const float4* normals) {   // ref :11-40
    h_hashtableParams.global_transform = viewMat;                  // :17
    h_hashtableParams.inv_global_transform = viewMat.getInverse(); // :15,18 (the device recomputes the same adjugate inverse)
    VH_MUST(vh_set_pose(ctx_, viewMat.entries, stream_));          // :21  (also the mutex reset of :24 in RefExact)
    VH_MUST(vh_fuse_frame(ctx_, verts, normals, nullptr, stream_)); // :27-36, count stays on the device
    if (synchronous_) {
        vh_stats st;
        VH_MUST(vh_get_stats(ctx_, &st, stream_));                 // one sync, where the reference has four
        h_hashtableParams.numOccupiedBlocks = (unsigned)st.numVisible;   // :32
    }
}

int SDF_Hashtable::occupiedBlockCount() {
    vh_stats st;
    VH_MUST(vh_get_stats(ctx_, &st, stream_));
    return st.numVisible;
}
const VoxelEntry* SDF_Hashtable::compactTable() const { return vh_compact_table_device(ctx_); }
const Voxel* SDF_Hashtable::voxelBlocks() const { return vh_voxel_blocks_device(ctx_); }
const int* SDF_Hashtable::compactCounter() const { return vh_compact_counter_device(ctx_); }

// ---- Solver ------------------------------------------------------------------------------------------
Solver::Solver(vh_context* ctx) : ctx_(ctx), ownsCtx_(false), stream_(nullptr), d_system_(nullptr) {   // ref :158-192
    if (!ctx_) { ctx_ = makeTrackingContext(640, 480); ownsCtx_ = true; }
    if (cudaMalloc((void**)&d_system_, sizeof(vh_icp_system)) != cudaSuccess) die("Solver: cudaMalloc");
    estimate.setZero();
    update.setZero();
}
Solver::~Solver() {
    cudaFree(d_system_);
    if (ownsCtx_) vh_destroy(ctx_);
}
void Solver::BuildLinearSystem(const float4* /*d_input*/, const float4* d_corr, const float4* d_corrN, const float* d_res,
                               int /*width*/, int /*height*/) {
    VH_MUST(vh_icp_reduce_corr(ctx_, d_corr, d_corrN, d_res, d_system_, stream_));   // :74-94 in one kernel
    VH_MUST(vh_icp_solve(ctx_, d_system_, stream_));                                  // :109-111 on the device
}
Matrix4x4f Solver::getTransform() {                              // ref Solver.h:32
    float delta[16], tw[6];
    vh_icp_system last;
    VH_MUST(vh_icp_get(ctx_, delta, tw, &last, stream_));
    for (int i = 0; i < 6; ++i) estimate(i) = tw[i];
    TotalError = last.error;
    Matrix4x4f M;
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) M(r, c) = delta[r * 4 + c];
    return M;
}
void Solver::PrintSystem() {                                     // ref :27-41
    vh_icp_system s;
    float d[16], tw[6];
    VH_MUST(vh_icp_get(ctx_, d, tw, &s, stream_));
    std::printf("\nFilled matrix system JTJ | JTr : \n");
    float full[36];
    for (int i = 0, k = 0; i < 6; ++i) for (int j = i; j < 6; ++j, ++k) full[i * 6 + j] = full[j * 6 + i] = s.JtJ[k];
    for (int i = 0; i < 6; ++i) {
        for (int j = 0; j < 6; ++j) std::printf("%g ", full[i * 6 + j]);
        std::printf("| %g\n", s.Jtr[i]);
    }
    std::printf("Calculated solution vector : \n");
    for (int i = 0; i < 6; ++i) std::printf("%g\n", tw[i]);
}
void Solver::SolveJacobianSystem(const Matrix6x6f& JTJ, const Vector6f& JTr) {   // ref :126-139 (unused there too)
    double L[36] = {0}, D[6], y[6], x[6];
    update.setZero();
    solution_exists = true;
    for (int j = 0; j < 6 && solution_exists; ++j) {             // LDL^T without pivoting
        double d = JTJ(j, j);
        for (int k = 0; k < j; ++k) d -= L[j * 6 + k] * L[j * 6 + k] * D[k];
        if (!(std::fabs(d) > 1e-12)) { solution_exists = false; break; }
        D[j] = d; L[j * 6 + j] = 1.0;
        for (int i = j + 1; i < 6; ++i) {
            double s = JTJ(i, j);
            for (int k = 0; k < j; ++k) s -= L[i * 6 + k] * L[j * 6 + k] * D[k];
            L[i * 6 + j] = s / d;
        }
    }
    if (solution_exists) {
        for (int i = 0; i < 6; ++i) { double s = -(double)JTr(i); for (int k = 0; k < i; ++k) s -= L[i * 6 + k] * y[k]; y[i] = s; }
        for (int i = 5; i >= 0; --i) { double s = y[i] / D[i]; for (int k = i + 1; k < 6; ++k) s -= L[k * 6 + i] * x[k]; x[i] = s; }
        for (int i = 0; i < 6; ++i) update(i) = (float)x[i];
    }
    estimate = updateTransform(update, estimate);
    float tw[6];
    for (int i = 0; i < 6; ++i) tw[i] = estimate(i);
    VH_MUST(vh_icp_set_delta(ctx_, tw, stream_));
}

// ---- CameraTracking -----------------------------------------------------------------------------------
CameraTracking::CameraTracking(int w, int h) : width(w), height(h), ctx_(makeTrackingContext(w, h)), ownsCtx_(true), stream_(nullptr) {}
CameraTracking::CameraTracking(int w, int h, vh_context* shared) : width(w), height(h), ctx_(shared), ownsCtx_(false), stream_(nullptr) {
    if (!ctx_) die("CameraTracking: null shared context");
}
CameraTracking::~CameraTracking() {
    if (ownsCtx_) vh_destroy(ctx_);
}
void CameraTracking::Align(float4* d_input, float4* d_inputNormals, float4* d_target, float4* d_targetNormals,
                           const uint16_t*, const uint16_t*) {
    VH_MUST(vh_icp_align(ctx_, d_input, d_inputNormals, d_target, d_targetNormals, maxIters, stream_));   // ref :35-67
}
Matrix4x4f CameraTracking::getTransform() {                      // ref CameraTracking.h:58
    float delta[16];
    VH_MUST(vh_icp_get(ctx_, delta, nullptr, nullptr, stream_));
    Matrix4x4f M;
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) M(r, c) = delta[r * 4 + c];
    return M;
}
void CameraTracking::resetEstimate() { VH_MUST(vh_icp_reset(ctx_, 1, stream_)); }
