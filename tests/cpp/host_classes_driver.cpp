// host_classes_driver.cpp -- test harness (NOT product code): drives the C++ host classes the way the reference's
// own host code does, so the parity tests can pin SURVEY.md row a23.
//
//   ./host_classes_driver <in.bin> <out.bin>
//   in.bin : two 640x480 u16 depth frames (target, then input)
//   out.bin: 3 x 16 floats (row-major 4x4): [0] CameraTracking::Align + getTransform (ref CameraTracking.cpp:26-69),
//            [1] the reference's OWN iteration loop written out -- computeCorrespondences (legacy entry point) ->
//            Solver::BuildLinearSystem -> Solver::getTransform, 20 times (ref CameraTracking.cpp:35-67,
//            Solver.cpp:48-124), [2] Solver::SolveJacobianSystem on the last system's JtJ / Jtr (host LDLT, ref
//            Solver.cpp:126-139) -> getTransform; then 6 floats: SE3Log(SE3Exp(twist)) round trip of a fixed twist.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#include "CameraTracking.h"
#include "SE3.h"
#include "Solver.h"
#include "vh/abi.h"

static void put(FILE* f, const Matrix4x4f& M) {
    float r[16];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r[i * 4 + j] = M(i, j);
    fwrite(r, sizeof(float), 16, f);
}

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    const int W = 640, H = 480, N = W * H;
    std::vector<uint16_t> d(2 * N);
    FILE* fi = fopen(argv[1], "rb");
    if (!fi || fread(d.data(), 2, 2 * N, fi) != (size_t)(2 * N)) return 3;
    fclose(fi);
    uint16_t* d_depth;
    float4 *tv, *tn, *iv, *in_, *corr, *corrN;
    float* res;
    cudaMalloc((void**)&d_depth, 2 * N * sizeof(uint16_t));
    cudaMemcpy(d_depth, d.data(), 2 * N * sizeof(uint16_t), cudaMemcpyHostToDevice);
    for (float4** p : {&tv, &tn, &iv, &in_, &corr, &corrN}) cudaMalloc((void**)p, N * sizeof(float4));
    cudaMalloc((void**)&res, N * sizeof(float));

    // the reference's start-up order: intrinsics, then the maps of both frames (Application.cpp:73-74, CameraTracking.cpp:131-134)
    const float K[9] = {517.3f, 0.f, 318.6f, 0.f, 516.5f, 255.3f, 0.f, 0.f, 1.f};
    const float Ki[9] = {1.0f / 517.3f, 0.f, -318.6f / 517.3f, 0.f, 1.0f / 516.5f, -255.3f / 516.5f, 0.f, 0.f, 1.f};
    SetCameraIntrinsic(K, Ki);
    preProcess(tv, tn, d_depth);
    preProcess(iv, in_, d_depth + N);

    FILE* fo = fopen(argv[2], "wb");
    if (!fo) return 4;
    {   // [0] the facade
        CameraTracking tracker(W, H);
        tracker.Align(iv, in_, tv, tn, d_depth + N, d_depth);
        put(fo, tracker.getTransform());
    }
    Matrix6x6f JTJ;
    Vector6f JTr;
    {   // [1] the loop of CameraTracking::Align written out with the legacy entry points and the Solver class
        Solver solver(vhLegacyContext());
        Matrix4x4f delta = Matrix4x4f::Identity();
        for (int it = 0; it < 20; ++it) {                                      // CameraTracking.cpp:35
            float4x4 dm;
            for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) dm.entries[r * 4 + c] = delta(r, c);   // :42-43 (row-major on the way down)
            computeCorrespondences(iv, tv, tn, corr, corrN, res, dm, W, H);   // :53
            solver.BuildLinearSystem(iv, corr, corrN, res, W, H);             // :62
            delta = solver.getTransform();                                     // :66
        }
        put(fo, delta);
        // the system of the last iteration, for the host-side LDLT variant
        float dl[16], tw[6];
        vh_icp_system s;
        vh_icp_get(solver.context(), dl, tw, &s, nullptr);
        for (int i = 0, k = 0; i < 6; ++i) for (int j = i; j < 6; ++j, ++k) JTJ(i, j) = JTJ(j, i) = s.JtJ[k];
        for (int i = 0; i < 6; ++i) JTr(i) = s.Jtr[i];
    }
    {   // [2] host LDLT on that system, starting from a zero estimate: exp(-(JtJ)^-1 Jtr)
        Solver solver;
        solver.SolveJacobianSystem(JTJ, JTr);
        put(fo, solver.getTransform());
    }
    Vector6f tw;
    const float t0[6] = {0.02f, -0.01f, 0.03f, 0.04f, -0.02f, 0.01f};
    for (int i = 0; i < 6; ++i) tw(i) = t0[i];
    Vector6f back = SE3Log(SE3Exp(tw));
    fwrite(back.data(), sizeof(float), 6, fo);
    fwrite(JTJ.data(), sizeof(float), 36, fo);
    fwrite(JTr.data(), sizeof(float), 6, fo);
    fclose(fo);
    deviceFree();
    return 0;
}
