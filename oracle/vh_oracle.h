/* oracle/vh_oracle.h -- C API of the CPU oracle.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.  Nothing under
 * voxelhashing_demo_b200/ or include/ includes, links or calls it.
 *
 * Parity status: the reference ships no tests, golden vectors or fixtures for this path
 * (SURVEY.md section 4), so the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF:
 * oracle/ref_harness.cu drives the unmodified reference .cu files (built into oracle/_ref/)
 * on the GPU box; tests/test_gpu_reference_pin.py compares them with this restatement and
 * tests/golden/ holds vectors captured from that run.  The host-side 6x6 solve and SE(3)
 * exp/log of the reference live in Eigen 3.3.7 (not vendored, not installed): that part is
 * "parity unpinned" and anchored on the mathematical definition only.
 */
#ifndef VH_ORACLE_H
#define VH_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { VO_POLICY_REF_EXACT = 0, VO_POLICY_FIXED = 1 };

typedef struct vo_config {
    int policy;
    int width, height;
    float K[9];        /* tracking-side intrinsics, row-major (common.h:14) */
    float Kinv[9];     /* its inverse as handed to SetCameraIntrinsic (CameraTracking.cpp:131-134) */
    float depthScale;  /* 5000 (CameraTrackingUtils.cu:64) */
    float depthMin, depthMax;
    unsigned int numBuckets, bucketSize, attachedLinkedListSize, numVoxelBlocks;
    unsigned int overflowSlots;
    float voxelSize, truncation, truncScale, maxIntegrationDistance;
    unsigned int integrationWeightSample;
    float integrationWeightMax;
    float icpDistThres, icpNormalThres;
    int partCount, partRank;
    float bilateralSigmaSpace, bilateralSigmaRange;   /* Fixed: 5x5 bilateral filter ahead of the maps; 0 = off */
} vo_config;

typedef struct vo_table vo_table;

typedef struct vo_alloc_report {
    int requestedPixels;      /* valid pixels that issued at least one request */
    int requestedBlocks;      /* distinct blocks requested (after frustum / ownership filter) */
    int requestedNew;         /* of those, not yet in the table at frame start */
    int inserted;             /* blocks inserted this frame */
    int bucketsTouched;       /* buckets that received a request for a NEW block */
    int bucketsContended;     /* buckets with >= 2 distinct NEW blocks requested (RefExact races here, Q4) */
    int maxNewPerBucket;
    int dropped;              /* requests that could not be stored */
} vo_alloc_report;

typedef struct vo_icp_system {
    float JtJ[21];
    float Jtr[6];
    float error;
    float count;
    float pad[3];
} vo_icp_system;

int  vo_num_threads(void);
void vo_set_num_threads(int n);

vo_table* vo_create(const vo_config* cfg);
void vo_destroy(vo_table* t);
void vo_reset(vo_table* t);

/* A.6 -- CameraTrackingUtils.cu:50-113.  verts/normals: W*H*4 floats. depthf may be NULL. */
void vo_preprocess(const vo_config* cfg, const uint16_t* depth, float* verts, float* normals, float* depthf);

/* A.3 -- VoxelUtils.cu:607-705 (+ :419-456 insert, :345-358 frustum, :251-259 hash).
 * pose: 16 floats row-major camera->world.  Pixels are processed in index order; the first
 * request for a bucket wins the per-frame try-lock (RefExact). */
void vo_alloc(vo_table* t, const float* pose, const float* verts, vo_alloc_report* rep);
/* Distinct blocks requested by the last vo_alloc that were NEW at frame start: xyz triples. */
int  vo_last_requested_new(vo_table* t, int* xyz, int cap);

/* A.4 -- VoxelUtils.cu:720-768.  Returns the visible count. */
int  vo_compact(vo_table* t, const float* pose);
/* A.5 -- VoxelUtils.cu:791-852.  verts: float4 map (z used).  Returns N_upd. */
long long vo_integrate(vo_table* t, const float* pose, const float* verts);
/* Fixed fast path: dense metric depth. */
long long vo_integrate_depthf(vo_table* t, const float* pose, const float* depthf);

/* Starvation + garbage collection (mirror of vh_garbage_collect).  scope 0 = last compaction, 1 = all.  Returns #released. */
int  vo_garbage_collect(vo_table* t, int scope, float sdfThreshold, float weightDecay);

/* Streaming (mirror of vh_stream_out / vh_stream_in).  entries5: x, y, z, 512 * record index, 0. */
int  vo_stream_out(vo_table* t, const float* center, float radius, int* entries5, float* voxelsOut, int capacity);
int  vo_stream_in(vo_table* t, const int* entries5, const float* voxels, int count);

/* Mesh of the zero level set (mirror of vh_extract_mesh): 9 floats per triangle; returns the model's triangle count. */
int  vo_extract_mesh(vo_table* t, float* tris, int capacity);

/* Table export. entries: 5 ints each (x,y,z,ptr,offset), allocated entries only. */
int  vo_num_allocated(vo_table* t);
int  vo_export_entries(vo_table* t, int* entries5, int cap);
int  vo_export_compact(vo_table* t, int* entries5, int cap);
/* 512 x {sdf, weight} of the block with key (x,y,z); returns 0 if absent. */
int  vo_get_block(vo_table* t, int x, int y, int z, float* voxels1024);
int  vo_heap_counter(vo_table* t);
unsigned int vo_hash(const vo_config* cfg, int x, int y, int z);      /* VoxelUtils.cu:251-259 (Q7) */
void vo_world2block(const vo_config* cfg, const float* p3, int* b3);  /* VoxelUtils.cu:267-309 */
int  vo_block_in_frustum(const vo_config* cfg, const float* pose, int x, int y, int z);

/* A.7 -- CameraTrackingUtils.cu:132-185.  delta: 16 floats row-major.  Outputs may be NULL.
 * Returns the residual sum accumulated in pixel order in fp32 (the reference's order is racy). */
float vo_find_correspondences(const vo_config* cfg, const float* input, const float* inputNormals,
                              const float* target, const float* targetNormals, const float* delta,
                              float* corr, float* corrNormals, float* residuals);
/* Solver.cu:25-51 -- 6 floats per pixel. */
void vo_jacobians(const vo_config* cfg, const float* corr, const float* corrNormals, float* J);
/* Normal equations accumulated in fp64 (Solver.cpp:80-94 do it in cuBLAS fp32, order unspecified). */
void vo_icp_system_build(const vo_config* cfg, const float* input, const float* inputNormals,
                         const float* target, const float* targetNormals, const float* delta,
                         int row0, int row1, vo_icp_system* out);
/* x = -(JtJ)^-1 Jtr; estimate <- log(exp(x) exp(estimate)); delta <- exp(estimate)
 * (Solver.cpp:109-111, SE3.cpp:4-19); fp64 closed forms. Returns 0 if singular. */
int  vo_icp_solve(const vo_icp_system* sys, float* estimate6, float* delta16);
/* The 20-iteration loop (CameraTracking.cpp:26-69). estimate6 is in/out (Q24). */
int  vo_icp_align(const vo_config* cfg, const float* input, const float* inputNormals,
                  const float* target, const float* targetNormals, int iterations,
                  float* estimate6, float* delta16);
void vo_se3_exp(const float* twist6, float* m16);
void vo_se3_log(const float* m16, float* twist6);
void vo_mat4_inverse(const float* m16, float* out16);   /* cuda_SimpleMatrixUtil.h:944-1069 */

/* Fixed-policy raycast through the table (no reference output exists). */
void vo_raycast(vo_table* t, const float* pose, float* verts, float* normals);

#ifdef __cplusplus
}
#endif
#endif
