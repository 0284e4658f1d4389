/* vh/abi.h -- the C ABI of libvh_b200.so.
 *
 * Two groups of entry points, both with C linkage, plain pointers and sizes only:
 *
 * (1) LEGACY names -- exactly the symbols the reference's host code binds
 *     (VoxelUtils.h:5-13, Application.cpp:22, CameraTracking.cpp:15-17, Solver.cpp:10,
 *     LinearSystem.cpp:11).  SDF_Hashtable.cpp / CameraTracking.cpp / Solver.cpp of the
 *     reference link against them unchanged.  They act on one process-global context with
 *     the RefExact arithmetic policy (the reference's as-built arithmetic, SURVEY.md
 *     Appendix A), use the legacy default stream and return only after the stream has
 *     drained, like the reference (every reference entry point ends in
 *     cudaDeviceSynchronize).  A CUDA error prints "CUDA error at file:line" and exits,
 *     like checkCudaErrors (cuda_helper/helper_cuda.h:965-981).
 *
 * (2) HANDLE API (vh_*) -- re-entrant, stream-ordered, status-returning; what the
 *     B200-native host classes (include/SDF_Hashtable.h, include/CameraTracking.h) and the
 *     Python mirror (voxelhashing_demo_b200/) are built on.  No call synchronises unless
 *     its comment says so; all device pointers are caller-owned unless stated.
 *
 * In C the reference's `const HashTableParams&` parameters appear as pointers, which is
 * what they are at the ABI level.
 */
#ifndef VH_ABI_H
#define VH_ABI_H

#include "vh/types.h"

#ifdef __cplusplus
#  define VH_REF(T) const T&
extern "C" {
#else
#  include <stdbool.h>
#  define VH_REF(T) const T*
#endif

struct cudaGraphicsResource;

/* ======================================================================================
 * (1) Legacy entry points (reference names)
 * ==================================================================================== */

/* ref VoxelUtils.h:5 / VoxelUtils.cu:87 -- copy the parameter block (incl. both poses). */
void updateConstantHashTableParams(VH_REF(HashTableParams) params);
/* ref VoxelUtils.h:6 / VoxelUtils.cu:170 -- allocate and initialise table, heap, mutexes.
 * Headless: also allocates (and ZEROES, quirk Q13) the compact list, voxel heap and
 * visible counter that the reference borrows from OpenGL. */
void deviceAllocate(VH_REF(HashTableParams) params);
/* ref VoxelUtils.h:7 / VoxelUtils.cu:214 */
void deviceFree(void);
/* ref VoxelUtils.h:8 / VoxelUtils.cu:146 -- per-frame bucket try-lock reset. */
void resetHashTableMutexes(VH_REF(HashTableParams) params);
/* ref VoxelUtils.h:9 / VoxelUtils.cu:708 -- per-pixel block request + insert. */
void allocBlocks(const float4* verts, const float4* normals);
/* ref VoxelUtils.h:10 / VoxelUtils.cu:751 -- visible-block compaction; returns the count. */
int flattenIntoBuffer(VH_REF(HashTableParams) params);
/* ref VoxelUtils.h:11 / VoxelUtils.cu:225 -- upload the fusion-side projection matrix. */
void calculateKinectProjectionMatrix(void);
/* ref VoxelUtils.h:12 / VoxelUtils.cu:844 -- TSDF integration over params.numOccupiedBlocks. */
void integrateDepthMap(VH_REF(HashTableParams) params, const float4* verts);
/* ref VoxelUtils.h:13 / VoxelUtils.cu:123 -- headless: the three resources are ignored
 * (may be NULL); the library-owned buffers stand in for the GL buffers. */
void mapGLobjectsToCUDApointers(struct cudaGraphicsResource* numBlocks_res,
                                struct cudaGraphicsResource* compactHashtable_res,
                                struct cudaGraphicsResource* sdfVolume_res);
/* ref Application.cpp:22 / CameraTrackingUtils.cu:115 -- u16 depth -> vertex + normal maps (640x480). */
void preProcess(float4* positions, float4* normals, const uint16_t* depth);
/* ref CameraTracking.cpp:17 / CameraTrackingUtils.cu:218 -- 9+9 floats, read row-major. */
bool SetCameraIntrinsic(const float* intrinsic, const float* invIntrinsic);
/* ref CameraTracking.cpp:15 / CameraTrackingUtils.cu:187 -- projective association;
 * returns the summed residual.  deltaTransform is a by-value float4x4 in the reference,
 * i.e. a hidden reference at the ABI level (see vh/types.h). */
#ifdef __cplusplus
float computeCorrespondences(const float4* d_input, const float4* d_target, const float4* d_targetNormals,
                             float4* corres, float4* corresNormals, float* residuals,
                             const float4x4 deltaTransform, const int width, const int height);
#else
float computeCorrespondences(const float4* d_input, const float4* d_target, const float4* d_targetNormals,
                             float4* corres, float4* corresNormals, float* residuals,
                             const vh_float4x4* deltaTransform, const int width, const int height);
#endif
/* ref Solver.cpp:10 / Solver.cu:56 -- 6-float Jacobian row per pixel (640x480). */
void CalculateJacobiansAndResiduals(const float4* d_src, const float4* d_targ, const float4* d_targNormals,
                                    float* d_Jac);
/* ref LinearSystem.cpp:11 / LinearSystem.cu:92 -- 300 x 27 partial normal equations
 * (the reducer the reference never builds). d_out: 300*27 floats device, h_out: same on host. */
void buildLinearSystemOnDevice(const float4* d_input, const float4* d_correspondence,
                               const float4* d_correspondenceNormals, float* d_out, float* h_out);

/* Headless additions that act on the same global context. */
const VoxelEntry* vhLegacyCompactTable(void);     /* device ptr, stands in for the GL buffer (SDFRenderer.cpp:36) */
const Voxel*      vhLegacyVoxelBlocks(void);      /* device ptr, stands in for the GL buffer (SDFRenderer.cpp:61) */
const int*        vhLegacyCompactCounter(void);   /* device ptr, stands in for the GL buffer (SDFRenderer.cpp:57) */
struct vh_context* vhLegacyContext(void);         /* the global context, for the export calls below */

/* ======================================================================================
 * (2) Handle API
 * ==================================================================================== */

typedef struct vh_context vh_context;
typedef void* vh_stream;   /* a cudaStream_t */

enum vh_status {
    VH_OK = 0,
    VH_ERR_INVALID = 1,     /* bad argument */
    VH_ERR_CUDA = 2,        /* CUDA runtime error; see vh_last_error() */
    VH_ERR_NO_DEVICE = 3,   /* no CUDA device: there is no CPU fallback */
    VH_ERR_CAPACITY = 4,    /* table / heap / overflow arena exhausted */
    VH_ERR_UNSUPPORTED = 5  /* operation not defined for this context's policy */
};

/* Arithmetic policy: same kernels, same memory traffic. */
enum vh_policy {
    VH_POLICY_REF_EXACT = 0,   /* the reference's as-built arithmetic, quirks included (Appendix A) */
    VH_POLICY_FIXED = 1        /* correct K, metric inverse pose, band allocation, overflow list */
};

typedef struct vh_config {
    HashTableParams table;       /* sizes, voxel size, truncation, weights (poses ignored) */
    int policy;                  /* enum vh_policy */
    int width, height;           /* image size; the reference bakes in 640x480 (quirk Q15) */
    float fx, fy, cx, cy;        /* pinhole intrinsics (common.h:7-10) */
    float depthScale;            /* raw units per metre; 5000 in the reference (CameraTrackingUtils.cu:64) */
    float depthMin, depthMax;    /* Fixed: valid sensor range in metres (0.1 .. maxIntegrationDistance) */
    unsigned int overflowSlots;  /* Fixed: size of the overflow arena appended to the table (0 = numBuckets) */
    float icpDistThres;          /* common.h:12 distThres = 0.08 */
    float icpNormalThres;        /* Fixed: min cos(angle) between source and target normals; <= -1 disables */
    int   icpIterations;         /* CameraTracking.h:40 maxIters = 20 */
    /* multi-GPU partition of the block-coordinate hash space: this context only inserts
     * blocks with owner(block) == partRank, owner = mix(hash) mod partCount. 1/0 = all. */
    int partCount, partRank;
    /* Fixed policy, optional (0 = off): 5x5 bilateral filter of the depth image ahead of the vertex / normal maps, as in
     * the tracking front end of KinectFusion-style pipelines ("a proper bilateral / validity
     * mask"; the reference has none, CameraTrackingUtils.cu:50-113).  sigma in pixels / metres.  Integration keeps
     * reading the RAW depth. */
    float bilateralSigmaSpace, bilateralSigmaRange;
} vh_config;

typedef struct vh_stats {
    int heapCounter;        /* free blocks - 1 (ref convention: counter starts at N-1, VoxelUtils.cu:207) */
    int numAllocated;       /* blocks handed out so far */
    int numVisible;         /* result of the last compaction */
    int overflowUsed;       /* entries of the overflow arena in use */
    int dropped;            /* requests dropped since creation (bucket/chain/heap full) */
    unsigned long long numUpdated;   /* voxels updated by the last integrate call */
    int lastInserted;       /* blocks inserted by the last allocation call */
    int lastFreed;          /* blocks released by the last vh_garbage_collect call */
    int overflowLeaked;     /* overflow-arena slots lost to append races since creation (counted in overflowUsed, never linked) */
    int exchangeTimeouts;   /* multi-GPU: Aligns stopped because a peer's ICP contribution did not arrive within ~5 s (a rank died) */
} vh_stats;

/* ICP normal equations, (v, omega) unknown order (ref Solver.cu:25-37, SE3.cpp:4-19):
 * 21 upper-triangle JtJ values row by row, 6 Jtr values, residual sum, inlier count. */
#define VH_ICP_SYSTEM_FLOATS 32
typedef struct vh_icp_system {
    float JtJ[21];
    float Jtr[6];
    float error;       /* sum of accepted residuals (ref globalError, CameraTrackingUtils.cu:175) */
    float count;       /* number of accepted correspondences */
    float pad[3];
} vh_icp_system;

const char* vh_last_error(void);
void vh_default_config(vh_config* cfg);     /* reference defaults (common.h:7-50), RefExact */
int  vh_device_count(void);

int  vh_create(const vh_config* cfg, vh_context** out);   /* allocates + zeroes everything */
void vh_destroy(vh_context* ctx);
int  vh_reset(vh_context* ctx, vh_stream s);               /* back to the freshly created state */
int  vh_get_config(const vh_context* ctx, vh_config* out);
int  vh_set_intrinsics(vh_context* ctx, float fx, float fy, float cx, float cy);
/* SetCameraIntrinsic semantics on a context: K and K^-1 used verbatim (9 floats each, row-major). */
int  vh_set_intrinsic_matrices(vh_context* ctx, const float* K9, const float* Kinv9);
unsigned long long vh_bytes_allocated(vh_context* ctx);
/* Scheduling knobs of the overlapped frame loop: CTAs of the persistent Align kernel (one per SM; default: all SMs but
 * 8) and the number of SMs the persistent integrate grid leaves free.  The two grids cannot share an SM, so tracking(k+1)
 * and fusion(k) run side by side only if each leaves the other whole SMs (large volumes on a partitioned context: a
 * small Align grid + the same number of reserved SMs).  <= 0 / < 0 keep the current value. */
int vh_set_tuning(vh_context* ctx, int align_ctas, int fusion_reserved_sms);

/* ---- pre-processing (ref CameraTrackingUtils.cu:50-120) ------------------------------ */
/* depth (u16, W*H) -> verts, normals (float4, W*H); depthf (float metres, W*H) optional. */
int vh_preprocess(vh_context* ctx, const uint16_t* d_depth, float4* d_verts, float4* d_normals,
                  float* d_depthf, vh_stream s);

/* ---- fusion (ref SDF_Hashtable.cpp:11-40) --------------------------------------------- */
/* Set the camera->world pose used by the following alloc/compact/integrate calls.
 * pose: 16 floats row-major on the HOST. */
int vh_set_pose(vh_context* ctx, const float* pose_rowmajor, vh_stream s);
/* Same, pose read from DEVICE memory at execution time (graph-capturable). */
int vh_set_pose_device(vh_context* ctx, const float* d_pose_rowmajor, vh_stream s);
int vh_alloc_blocks(vh_context* ctx, const float4* d_verts, const float4* d_normals, vh_stream s);
/* The same allocation straight from the raw u16 depth image: the back-projection of
 * vh_preprocess (ref CameraTrackingUtils.cu:50-76) runs in registers, same operations in the same order, so the
 * requested blocks are identical; 2 B per pixel are read instead of the 16 B of the vertex map.  Not available with
 * the bilateral front end (that allocates from the filtered vertex map). */
int vh_alloc_blocks_depth(vh_context* ctx, const uint16_t* d_depth, vh_stream s);
int vh_compact(vh_context* ctx, vh_stream s);             /* count stays on the device */
int vh_integrate(vh_context* ctx, const float4* d_verts, vh_stream s);
/* Fixed policy fast path: integrate from the dense metric depth image. */
int vh_integrate_depthf(vh_context* ctx, const float* d_depthf, vh_stream s);
/* alloc + compact + integrate, stream-ordered, no host sync (the reference's
 * SDF_Hashtable::integrate without its 4 syncs and 2 D2H copies). */
int vh_fuse_frame(vh_context* ctx, const float4* d_verts, const float4* d_normals,
                  const float* d_depthf_or_null, vh_stream s);
/* Synchronises the stream and copies the counters back. */
int vh_get_stats(vh_context* ctx, vh_stats* out, vh_stream s);
/* Voxel starvation + block garbage collection (Fixed policy only).  The reference's removal
 * path is dead and wrong (deleteVoxelEntry, VoxelUtils.cu:544-604; removeSingleBlockInHeap :336-341 is followed for
 * the heap push); the pass is the one of the paper the reference implements (Niessner et al. 2013, section 4.4).
 * For every block in scope -- VH_GC_VISIBLE: the list of the last vh_compact; VH_GC_ALL: every allocated block --
 *   weight <- max(weight - weight_decay, 0) for every voxel          (skipped when weight_decay <= 0)
 *   release the block when no voxel has weight > 0, or min |sdf| over the voxels with weight > 0 >= sdf_threshold
 *   (sdf_threshold <= 0 selects truncation + truncScale * depthMax).
 * A released block's voxels are zeroed, its hash slot becomes a tombstone that later allocations reclaim, and its id
 * goes back on the heap.  The visible list is emptied: call vh_compact before the next vh_integrate / vh_raycast. */
#define VH_GC_VISIBLE 0
#define VH_GC_ALL 1
int vh_garbage_collect(vh_context* ctx, int scope, float sdf_threshold, float weight_decay, vh_stream s);
/* Streaming of the model in and out of the device (Fixed policy; Niessner et al. 2013, section 4.5).
 * vh_stream_out moves every allocated block whose centre lies farther than `radius` metres from `center_xyz` (world
 * frame) into the caller's buffers and releases it like vh_garbage_collect does: entries_out[i] = {pos, ptr = 512 * i,
 * offset = 0} in the reference's 20-byte VoxelEntry layout, voxels_out[512 * i ...] its 512 voxels.  At most
 * `capacity` blocks move; the rest stay where they are.  *h_count receives the number moved (the call synchronises).
 * vh_stream_in inserts `count` such blocks again: a key that is absent gets the voxels as they are, a key that was
 * re-observed meanwhile is merged, sdf = (s1 w1 + s2 w2) / (w1 + w2), w = min(wMax, w1 + w2).  *h_count (may be
 * NULL) receives the number accepted; blocks the table or heap cannot take are counted in vh_stats.dropped.
 * The buffers must be device-ACCESSIBLE: device memory, or pinned host memory (cudaHostAlloc), in which case the
 * blocks travel straight over the host link.  Both calls empty the visible list. */
int vh_stream_out(vh_context* ctx, const float* center_xyz, float radius, VoxelEntry* entries_out, Voxel* voxels_out,
                  int capacity, int* h_count, vh_stream s);
int vh_stream_in(vh_context* ctx, const VoxelEntry* entries, const Voxel* voxels, int count, int* h_count, vh_stream s);

/* ---- tracking (ref CameraTracking.cpp:26-69, Solver.cpp:48-124) ------------------------ */
/* One fused Gauss-Newton iteration on the device: association + residual + 27-sum reduction
 * + 6x6 solve + SE(3) update of the context's delta transform.  No host round trip. */
int vh_icp_reset(vh_context* ctx, int reset_estimate, vh_stream s);
int vh_icp_iterate(vh_context* ctx, const float4* d_input, const float4* d_inputNormals,
                   const float4* d_target, const float4* d_targetNormals, vh_stream s);
/* maxIters iterations (ref Align). */
int vh_icp_align(vh_context* ctx, const float4* d_input, const float4* d_inputNormals,
                 const float4* d_target, const float4* d_targetNormals, int iterations, vh_stream s);
/* The tracking of one frame in ONE persistent launch (SURVEY.md section 8 f1 + f2): the pre-processing of the raw u16
 * frame (ref preProcess, Application.cpp:73), the whole Align against d_target / d_targetNormals (ref
 * CameraTracking.cpp:26-69) and, if the pose pointers are given, the chain d_pose_out = d_pose_in * delta (ref
 * getTransform -> integrate, Application.cpp:75-84; the two may alias).  d_verts / d_normals / d_depthf are OUTPUTS:
 * the same maps vh_preprocess writes (bit-identical) -- they are the next frame's target and the fusion's input.
 * On a context with peers (vh_set_peers) this rank reduces its share of the image rows.  Not available with the
 * bilateral front end. */
int vh_track_frame(vh_context* ctx, const uint16_t* d_depth, float4* d_verts, float4* d_normals, float* d_depthf,
                   const float4* d_target, const float4* d_targetNormals, int iterations,
                   const float* d_pose_in, float* d_pose_out, vh_stream s);
/* Reduction only (no solve): writes one vh_icp_system to d_system. Rows [row0,row1) of the
 * image only -- the multi-GPU split. */
int vh_icp_reduce(vh_context* ctx, const float4* d_input, const float4* d_inputNormals,
                  const float4* d_target, const float4* d_targetNormals,
                  int row0, int row1, vh_icp_system* d_system, vh_stream s);
/* Multi-GPU, fused form: register the peer-mapped exchange regions (one per rank, vh_peer_bytes() each,
 * zero-initialised; bufs[p] = rank p's region as mapped into this process, world <= 8), then
 * vh_icp_align_rows runs `iterations` Gauss-Newton iterations over this rank's rows with the all-reduce of
 * the 32-float system fused into each kernel's epilogue (P2P stores + sequence-numbered flags over NVLink);
 * every rank ends with the bit-identical delta. */
int vh_set_peers(vh_context* ctx, int rank, int world, void* const* bufs);
unsigned long long vh_peer_bytes(void);
int vh_icp_align_rows(vh_context* ctx, const float4* d_input, const float4* d_inputNormals,
                      const float4* d_target, const float4* d_targetNormals, int row0, int row1,
                      int iterations, vh_stream s);
/* Solve + SE(3) update from an (all-reduced) system in device memory. */
int vh_icp_solve(vh_context* ctx, const vh_icp_system* d_system, vh_stream s);
/* Synchronises; delta: 16 floats row-major (input frame -> target frame), twist: 6 floats (v, omega). */
int vh_icp_get(vh_context* ctx, float* delta_rowmajor, float* twist6, vh_icp_system* last_system, vh_stream s);
int vh_icp_set_delta(vh_context* ctx, const float* twist6, vh_stream s);
const float* vh_icp_delta_device(vh_context* ctx);    /* device ptr to the 16 floats */
/* d_pose_out = d_pose_in * delta  (camera->world chain), all on the device. */
int vh_pose_compose(vh_context* ctx, const float* d_pose_in, float* d_pose_out, vh_stream s);

/* Normal equations from stored correspondences (what Solver::BuildLinearSystem gets). */
int vh_icp_reduce_corr(vh_context* ctx, const float4* d_corr, const float4* d_corrNormals,
                       const float* d_residuals, vh_icp_system* d_system, vh_stream s);

/* Split form of an iteration (the reference's computeCorrespondences / CalculateJacobiansAndResiduals
 * on a context): correspondences, normals, residuals per pixel (zero where rejected); *d_err += sum. */
int vh_find_correspondences(vh_context* ctx, const float4* d_input, const float4* d_inputNormals,
                            const float4* d_target, const float4* d_targetNormals, const float* delta_rowmajor_host,
                            float4* d_corr, float4* d_corrNormals, float* d_residuals, float* d_err, vh_stream s);
int vh_jacobians(vh_context* ctx, const float4* d_corr, const float4* d_corrNormals, float* d_J, vh_stream s);

/* ---- raycast (replaces shaders/raycastSDF.*; Fixed policy only) ------------------------ */
/* Casts one ray per pixel from the current pose through the hash table; writes vertex and
 * normal maps in the CAMERA frame with the same conventions as vh_preprocess. */
int vh_raycast(vh_context* ctx, float4* d_verts, float4* d_normals, vh_stream s);

/* ---- export (tests, checkpoints; these synchronise) ------------------------------------ */
/* Number of allocated entries; fills up to cap entries (reference layout) on the HOST. */
int vh_export_entries(vh_context* ctx, VoxelEntry* h_entries, int cap, int* count);
/* Copies the compact (visible) list of the last vh_compact to the host. */
int vh_export_compact(vh_context* ctx, VoxelEntry* h_entries, int cap, int* count);
/* Copies the 512 voxels of the block starting at voxel index ptr to the host. */
int vh_export_block(vh_context* ctx, int ptr, Voxel* h_voxels512);
/* Device pointers of the library-owned buffers. */
const VoxelEntry* vh_compact_table_device(vh_context* ctx);
const int*        vh_compact_counter_device(vh_context* ctx);
Voxel*            vh_voxel_blocks_device(vh_context* ctx);
/* Binary checkpoint of table + heap + voxels. */
int vh_save(vh_context* ctx, const char* path);
int vh_load(vh_context* ctx, const char* path);
/* The reference's text dump (SDFRenderer.cpp:71-110) of the compact list. */
int vh_dump_text(vh_context* ctx, const char* path);

/* Triangle mesh of the zero level set (the reference has no extraction, only the text dump
 * above).  Marching tetrahedra on the Kuhn subdivision of every voxel cell whose eight corners have been observed;
 * d_triangles receives 9 floats per triangle (three world-space vertices in metres, normal towards free space).  At
 * most `capacity` triangles are written; *h_count receives the number the model HAS (call again with a larger
 * buffer if it exceeds capacity; capacity 0 with d_triangles NULL just counts).  Synchronises.  vh_save_mesh_ply writes
 * a binary little-endian PLY (vertices not shared) from a HOST copy of such a buffer. */
int vh_extract_mesh(vh_context* ctx, float* d_triangles, int capacity, int* h_count, vh_stream s);
int vh_save_mesh_ply(const char* path, const float* h_triangles, int count);

/* ---- depth images on the input side (host only, no device needed) -----------------------------------
 * The reference reads its frames with stbi_load_16("assets/T0.png", ...) (Application.cpp:28-29, vendored
 * stb_image.h): raw uint16 samples, 5000 per metre.  vh_depth_read decodes that format without third-party code
 * beyond zlib -- PNG (grey / colour, 8 / 16 bit, non-interlaced, CRC-checked; stb's conventions: 8-bit samples are
 * widened as v * 257, the first channel of a multi-channel file is returned) or binary PGM (P5) -- into a malloc'ed
 * row-major buffer the caller releases with vh_depth_free.  vh_depth_write_png writes 16-bit grey PNGs with the
 * given scanline filter (0..4). */
int  vh_depth_read(const char* path, uint16_t** out, int* width, int* height);
void vh_depth_free(uint16_t* data);
int  vh_depth_write_png(const char* path, const uint16_t* data, int width, int height, int filter);
const char* vh_depth_last_error(void);

/* ---- native frame pipeline (the host loop of Application.cpp:73-84 without host syncs) -------
 * preprocess -> ICP x iterations -> pose <- pose * delta -> alloc -> compact -> integrate
 * [-> raycast], stream-ordered, the tracked part replayed from a CUDA graph. */
typedef struct vh_pipeline vh_pipeline;
enum vh_track_mode {
    VH_TRACK_FRAME_TO_FRAME = 0,   /* ICP target = previous frame's maps (the reference's formulation) */
    VH_TRACK_FRAME_TO_MODEL = 1,   /* ICP target = raycast of the model at the previous pose (Fixed only) */
    VH_TRACK_NONE = 2              /* no ICP: every frame is fused at the pose set by vh_pipeline_reset */
};
/* useGraph is a flag word: VH_PIPE_GRAPH captures the per-frame work into CUDA graphs; VH_PIPE_OVERLAP (needs
 * VH_PIPE_GRAPH and frame-to-frame tracking, ignored otherwise) additionally runs the fusion of frame k on an internal
 * stream beside the preprocessing and tracking of frame k+1.  With VH_PIPE_OVERLAP the MODEL lags the pose: after
 * vh_pipeline_push_* the caller's stream is ordered behind the pose of the frame, not behind its fusion; call
 * vh_pipeline_flush (or vh_pipeline_pose, which includes it) before reading the table on that stream. */
#define VH_PIPE_GRAPH 1
#define VH_PIPE_OVERLAP 2
int  vh_pipeline_create(vh_context* ctx, int icpIterations, int mode, int useGraph, vh_pipeline** out);
int  vh_pipeline_flush(vh_pipeline* p, vh_stream s);
void vh_pipeline_destroy(vh_pipeline* p);
/* New sequence: frame counter, ICP state and pose (16 floats row-major on the host, NULL = identity).
 * Synchronises. The table itself is reset with vh_reset. */
int  vh_pipeline_reset(vh_pipeline* p, const float* pose_rowmajor_host, vh_stream s);
/* One frame, depth already in device memory. Returns after enqueueing. */
int  vh_pipeline_push_device(vh_pipeline* p, const uint16_t* d_depth, vh_stream s);
/* Same, for a depth image that is NOT produced by work on stream s: it is complete in device memory once `ready_event`
 * (a cudaEvent_t; NULL = complete already) has fired.  With VH_PIPE_OVERLAP the pre-processing of the frame then runs
 * BESIDE the tracking of the previous frame instead of behind it (the per-frame tracking chain is the Align kernel
 * alone).  The buffer must stay untouched until s -- which is ordered behind the frame's pose -- gets there. */
int  vh_pipeline_push_device_ready(vh_pipeline* p, const uint16_t* d_depth, void* ready_event, vh_stream s);
/* One frame end to end: H2D of the (pinned) host depth, the frame, D2H of the pose into
 * h_pose_out (pinned, 16 floats, may be NULL). Returns after enqueueing; sync the stream to read. */
int  vh_pipeline_push_host(vh_pipeline* p, const uint16_t* h_depth, float* h_pose_out, vh_stream s);
int  vh_pipeline_pose(vh_pipeline* p, float* pose_rowmajor_host, vh_stream s);   /* synchronises */
const float* vh_pipeline_pose_device(vh_pipeline* p);
int  vh_pipeline_pose_async(vh_pipeline* p, float* h_pose_rowmajor_pinned, vh_stream s);   /* stream-ordered D2H, no sync */
int  vh_pipeline_depthf(vh_pipeline* p, float** d_depthf);   /* dense metric depth of the latest frame */
/* which = 0: maps of the latest frame; 1: the ICP target the next frame will use. */
int  vh_pipeline_maps(vh_pipeline* p, int which, float4** d_verts, float4** d_normals);
long long vh_pipeline_launches(vh_pipeline* p);   /* kernels launched so far */

/* ---- multi-GPU from C / C++ (one process per GPU of one node; CUDA IPC peer memory over NVLink / NVSwitch) -----------
 * The block-coordinate hash space is partitioned with vh_config::partCount / partRank; what the ranks exchange is (1)
 * every depth frame and (2) the 32-float ICP system of every Gauss-Newton iteration.  Both go
 * through one exchange region per rank that every other rank maps:
 *     vh_dist_create      allocates this rank's region (ICP mailboxes + three frame landing buffers + flags)
 *     vh_dist_export      its 64-byte CUDA IPC handle -- the host program hands the handles of all ranks around
 *                         (MPI_Allgather, a socket, files: examples/dist_app.cpp uses files)
 *     vh_dist_connect     maps the other ranks' regions (handles: world x vh_dist_handle_bytes(), in rank order) and
 *                         registers the mailboxes with the context (vh_set_peers): from here on vh_icp_align_rows /
 *                         vh_track_frame / the pipeline carry the all-reduce INSIDE the persistent Align kernel
 *     vh_dist_broadcast_frame   rank 0 passes the frame (device memory, complete once stream s gets there), the others
 *                         NULL; rank 0's kernel stores it straight into every rank's landing buffer (no NCCL, no host
 *                         hop).  Every rank receives the address of its copy and the cudaEvent_t that fires when it is
 *                         complete -- the arguments of vh_pipeline_push_device_ready
 *     vh_dist_frame_consumed    stream s has pre-processed the oldest outstanding frame (e.g. the stream a frame was
 *                         pushed on, which is ordered behind its pose): rank 0 may overwrite that landing buffer
 * Every rank must broadcast / push the same number of frames (the fused all-reduce waits for every rank). */
typedef struct vh_dist vh_dist;
int  vh_dist_create(vh_context* ctx, int rank, int world, vh_dist** out);
unsigned long long vh_dist_handle_bytes(void);
int  vh_dist_export(vh_dist* d, void* handle_out);
int  vh_dist_connect(vh_dist* d, const void* handles_by_rank);
int  vh_dist_broadcast_frame(vh_dist* d, const uint16_t* d_depth_rank0, const uint16_t** d_frame, void** ready_event, vh_stream s);
int  vh_dist_frame_consumed(vh_dist* d, vh_stream s);
int  vh_dist_timeouts(vh_dist* d);     /* waits given up after ~5 s (a rank died): > 0 means the frames since then are not to be trusted; synchronises */
void vh_dist_destroy(vh_dist* d);

#ifdef __cplusplus
}
#endif

#endif /* VH_ABI_H */
