// oracle/ref_harness.cu -- headless driver for the UNMODIFIED reference CUDA sources.
//
// TEST INFRASTRUCTURE, NOT PRODUCT.  Linked with VoxelUtils.o, CameraTrackingUtils.o and Solver.o
// compiled straight from /root/reference (oracle/Makefile, target `ref`) into
// oracle/_ref/libvh_ref.so.  It stands in for the host files of the reference that cannot be
// built here (SDF_Hashtable.cpp needs glm/SDL2/OpenGL; CameraTracking.cpp/Solver.cpp need Eigen):
//   * ref_integrate   = SDF_Hashtable::integrate (SDF_Hashtable.cpp:11-40) minus the GL map/unmap
//   * ref_correspond  = computeCorrespondences as called from CameraTracking.cpp:53
//   * ref_build_system= Solver::BuildLinearSystem's device half (Solver.cpp:74-94, cuBLAS)
//   * ref_align       = CameraTracking::Align's loop (CameraTracking.cpp:35-67); the Eigen half
//                       (6x6 inverse, SE(3) exp/log, Solver.cpp:109-111) is delegated to a callback
// The three buffers the reference borrows from OpenGL (compact table, voxel heap, visible
// counter; SDFRenderer.cpp:34-61) are cudaMalloc'd here and the voxel heap is ZEROED, which the
// reference never does (quirk Q13).  Used to pin the CPU oracle and as the "reference CUDA
// kernels rebuilt for the box" baseline of BASELINE.md 3.1.

#include <cstdio>
#include <cstring>
#include <cublas_v2.h>
#include <cuda_runtime.h>

#include "VoxelUtils.h"   // the reference's header: HashTableParams, PtrContainer, C prototypes
#include "common.h"       // the reference's defaults

extern PtrContainer h_ptrHldr;     // VoxelUtils.cu:26
void updateDevicePointers();       // VoxelUtils.cu:93 (C++ linkage)

extern "C" void preProcess(float4* positions, float4* normals, const uint16_t* depth);      // CameraTrackingUtils.cu:115
extern "C" bool SetCameraIntrinsic(const float* intrinsic, const float* invIntrinsic);       // :218
extern "C" float computeCorrespondences(const float4*, const float4*, const float4*, float4*, float4*, float*,
                                        const float4x4, const int, const int);               // :187
extern "C" void CalculateJacobiansAndResiduals(const float4*, const float4*, const float4*, float*);   // Solver.cu:56

namespace {
HashTableParams g_params;
bool g_ready = false;
float4* g_corr = nullptr;
float4* g_corrN = nullptr;
float* g_res = nullptr;
float* g_jac = nullptr;
float* g_jtj = nullptr;
float* g_jtr = nullptr;
cublasHandle_t g_blas = nullptr;
}  // namespace

extern "C" {

// SDF_Hashtable::SDF_Hashtable (SDF_Hashtable.cpp:60-81).  Arguments <= 0 take common.h defaults.
int ref_init(int nBuckets, int bSize, int nBlocks, float vSize, float trunc) {
    if (g_ready) return -1;    // reference state is process-global: one table per process
    memset(&g_params, 0, sizeof(g_params));
    g_params.numBuckets = nBuckets > 0 ? nBuckets : numBuckets;
    g_params.bucketSize = bSize > 0 ? bSize : bucketSize;
    g_params.attachedLinkedListSize = attachedLinkedListSize;
    g_params.numVoxelBlocks = nBlocks > 0 ? nBlocks : numVoxelBlocks;
    g_params.voxelBlockSize = voxelBlockSize;
    g_params.voxelSize = vSize > 0 ? vSize : voxelSize;
    g_params.numOccupiedBlocks = numOccupiedBlocks;
    g_params.maxIntegrationDistance = maxIntegrationDistance;
    g_params.truncScale = truncScale;
    g_params.truncation = trunc > 0 ? trunc : truncation;
    g_params.integrationWeightSample = integrationWeightSample;
    g_params.integrationWeightMax = integrationWeightMax;
    g_params.global_transform.setIdentity();
    g_params.inv_global_transform.setIdentity();

    updateConstantHashTableParams(g_params);
    deviceAllocate(g_params);
    calculateKinectProjectionMatrix();

    const size_t slots = (size_t)g_params.numBuckets * g_params.bucketSize;
    const size_t voxBytes = sizeof(Voxel) * 512 * (size_t)g_params.numVoxelBlocks;
    if (cudaMalloc((void**)&h_ptrHldr.d_compactifiedHashTable, sizeof(VoxelEntry) * slots) != cudaSuccess) return -2;
    if (cudaMalloc((void**)&h_ptrHldr.d_SDFBlocks, voxBytes) != cudaSuccess) return -2;
    if (cudaMalloc((void**)&h_ptrHldr.d_compactifiedHashCounter, sizeof(int)) != cudaSuccess) return -2;
    cudaMemset(h_ptrHldr.d_compactifiedHashTable, 0, sizeof(VoxelEntry) * slots);
    cudaMemset(h_ptrHldr.d_SDFBlocks, 0, voxBytes);                       // Q13
    cudaMemset(h_ptrHldr.d_compactifiedHashCounter, 0, sizeof(int));
    updateDevicePointers();

    const size_t px = (size_t)numCols * numRows;
    cudaMalloc((void**)&g_corr, px * sizeof(float4));                      // CameraTracking.cpp:119-126
    cudaMalloc((void**)&g_corrN, px * sizeof(float4));
    cudaMalloc((void**)&g_res, px * sizeof(float));
    cudaMalloc((void**)&g_jac, px * 6 * sizeof(float));                    // Solver.cpp:159-165
    cudaMalloc((void**)&g_jtj, 36 * sizeof(float));
    cudaMalloc((void**)&g_jtr, 6 * sizeof(float));
    cudaMemset(g_corr, 0, px * sizeof(float4));
    cudaMemset(g_corrN, 0, px * sizeof(float4));
    cudaMemset(g_res, 0, px * sizeof(float));
    cudaMemset(g_jac, 0, px * 6 * sizeof(float));
    cudaMemset(g_jtj, 0, 36 * sizeof(float));
    cudaMemset(g_jtr, 0, 6 * sizeof(float));
    cublasCreate(&g_blas);
    g_ready = cudaDeviceSynchronize() == cudaSuccess;
    return g_ready ? 0 : -3;
}

void ref_set_intrinsic(const float* K9, const float* Kinv9) { SetCameraIntrinsic(K9, Kinv9); }

void ref_preprocess(float4* verts, float4* normals, const uint16_t* depth) { preProcess(verts, normals, depth); }

// SDF_Hashtable::integrate, SDF_Hashtable.cpp:11-40.  pose: 16 floats row-major, camera->world.
int ref_integrate(const float* pose, const float4* verts, const float4* normals) {
    float4x4 viewMat(pose);
    float4x4 inv = viewMat.getInverse();                 // :15
    g_params.global_transform = viewMat;                 // :17
    g_params.inv_global_transform = inv;                 // :18
    updateConstantHashTableParams(g_params);             // :21
    resetHashTableMutexes(g_params);                     // :24
    allocBlocks(verts, normals);                         // :27
    int occupied = flattenIntoBuffer(g_params);          // :30
    g_params.numOccupiedBlocks = occupied;               // :32
    updateConstantHashTableParams(g_params);             // :33
    integrateDepthMap(g_params, verts);                  // :36
    return occupied;
}

// stage-by-stage variants for the per-stage baseline timings
void ref_stage_begin(const float* pose) {
    float4x4 viewMat(pose);
    g_params.global_transform = viewMat;
    g_params.inv_global_transform = viewMat.getInverse();
    updateConstantHashTableParams(g_params);
    resetHashTableMutexes(g_params);
}
void ref_stage_alloc(const float4* verts, const float4* normals) { allocBlocks(verts, normals); }
int ref_stage_compact(void) {
    int occupied = flattenIntoBuffer(g_params);
    g_params.numOccupiedBlocks = occupied;
    updateConstantHashTableParams(g_params);
    return occupied;
}
void ref_stage_integrate(const float4* verts) { integrateDepthMap(g_params, verts); }

int ref_num_slots(void) { return (int)(g_params.numBuckets * g_params.bucketSize); }

// whole table, reference layout, to the host (slots with ptr == -1 included)
int ref_export_table(VoxelEntry* out, int cap) {
    int n = ref_num_slots();
    if (cap < n) return -n;
    cudaMemcpy(out, h_ptrHldr.d_hashTable, sizeof(VoxelEntry) * n, cudaMemcpyDeviceToHost);
    return n;
}
int ref_export_compact(VoxelEntry* out, int cap) {
    int n = 0;
    cudaMemcpy(&n, h_ptrHldr.d_compactifiedHashCounter, sizeof(int), cudaMemcpyDeviceToHost);
    if (n > cap) n = cap;
    cudaMemcpy(out, h_ptrHldr.d_compactifiedHashTable, sizeof(VoxelEntry) * n, cudaMemcpyDeviceToHost);
    return n;
}
void ref_export_block(int ptr, Voxel* out512) {
    cudaMemcpy(out512, h_ptrHldr.d_SDFBlocks + ptr, sizeof(Voxel) * 512, cudaMemcpyDeviceToHost);
}
int ref_heap_counter(void) {
    int v = 0;
    cudaMemcpy(&v, h_ptrHldr.d_heapCounter, sizeof(int), cudaMemcpyDeviceToHost);
    return v;
}

// CameraTracking.cpp:42-53: delta is Eigen column-major in the reference, transposed into the
// row-major float4x4; here the caller passes it row-major already.
float ref_correspond(const float4* input, const float4* target, const float4* targetNormals, const float* delta,
                     float4* corrOut, float4* corrNOut, float* resOut) {
    float4x4 d(delta);
    float err = computeCorrespondences(input, target, targetNormals, g_corr, g_corrN, g_res, d, numCols, numRows);
    const size_t px = (size_t)numCols * numRows;
    if (corrOut) cudaMemcpy(corrOut, g_corr, px * sizeof(float4), cudaMemcpyDeviceToDevice);
    if (corrNOut) cudaMemcpy(corrNOut, g_corrN, px * sizeof(float4), cudaMemcpyDeviceToDevice);
    if (resOut) cudaMemcpy(resOut, g_res, px * sizeof(float), cudaMemcpyDeviceToDevice);
    return err;
}

// Solver::BuildLinearSystem, Solver.cpp:74-94.  Outputs on the host: Jtr[6]; JtJ[36] exactly as
// cuBLAS leaves it (column-major, LOWER triangle valid).  jacOut (device, 6*W*H) optional.
void ref_build_system(const float4* input, float* JtJ36, float* Jtr6, float* jacOut) {
    const float alpha = 1.0f, beta = 0.0f;
    const int n = numCols * numRows;
    CalculateJacobiansAndResiduals(input, g_corr, g_corrN, g_jac);                                    // :74
    cudaDeviceSynchronize();                                                                         // :75
    cublasSgemv(g_blas, CUBLAS_OP_N, 6, n, &alpha, g_jac, 6, g_res, 1, &beta, g_jtr, 1);             // :80
    cudaMemcpy(Jtr6, g_jtr, 6 * sizeof(float), cudaMemcpyDeviceToHost);                              // :82
    cublasSsyrk(g_blas, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, 6, n, &alpha, g_jac, 6, &beta, g_jtj, 6);   // :87
    cudaMemcpy(JtJ36, g_jtj, 36 * sizeof(float), cudaMemcpyDeviceToHost);                            // :89
    cudaDeviceSynchronize();                                                                         // :93
    if (jacOut) cudaMemcpy(jacOut, g_jac, (size_t)n * 6 * sizeof(float), cudaMemcpyDeviceToDevice);
}

// sys32: 21 upper-triangle JtJ row by row, 6 Jtr, error, count, 3 pad (the oracle's vo_icp_system)
typedef int (*ref_solve_fn)(const float* sys32, float* estimate6, float* delta16);

// CameraTracking::Align, CameraTracking.cpp:26-69.  estimate6 in/out (Solver::estimate is never
// reset, Q24); delta16 in/out row-major.  Returns the number of iterations that ran a solve.
int ref_align(const float4* input, const float4* target, const float4* targetNormals, int iters, ref_solve_fn solve,
              float* estimate6, float* delta16) {
    int done = 0;
    for (int it = 0; it < iters; ++it) {                                                              // :35
        float err = ref_correspond(input, target, targetNormals, delta16, nullptr, nullptr, nullptr);    // :53
        if (err == 0.0f) break;                                                                       // :55-58
        float JtJ[36], Jtr[6], sys[32];
        ref_build_system(input, JtJ, Jtr, nullptr);                                                   // :63
        int k = 0;
        for (int i = 0; i < 6; ++i)
            for (int j = i; j < 6; ++j) sys[k++] = JtJ[i * 6 + j];    // lower(j,i) of col-major C == C[i*6+j]
        for (int i = 0; i < 6; ++i) sys[21 + i] = Jtr[i];
        sys[27] = err; sys[28] = 0; sys[29] = sys[30] = sys[31] = 0;
        if (!solve(sys, estimate6, delta16)) break;                                                   // Solver.cpp:109-111, :64-66
        ++done;
    }
    return done;
}

}  // extern "C"
