// vh_internal.h -- context, device views and launch declarations shared by the .cu files.
// Not installed; the public surface is include/vh/abi.h.
#ifndef VH_INTERNAL_H
#define VH_INTERNAL_H

#include <cuda_runtime.h>
#include <stdint.h>

#include "vh/abi.h"

namespace vh {

// ---- HBM layout --------------------------------------------------------------------------------
// entries   int4[S + O]       hash slots {x, y, z, ptr}: 16-byte aligned so a slot is read with one
//                             128-bit load and claimed with one 128-bit CAS (ATOMG.E.CAS.128).
//                             S = numBuckets*bucketSize in-bucket slots, then O overflow-arena slots.
// chain     int[S + O]        relative index of the next entry of the bucket's overflow chain
//                             (VoxelEntry::offset of the reference layout), 0 = none.
// mutex     int[numBuckets]   RefExact only: the per-frame try-lock of VoxelUtils.cu:444.
// heap      uint[N]           free-list of block ids; heapCounter counts down from N-1 (ref :207).
// blockInfo int4[N]           {x, y, z, slot} of the block that owns heap id i (w = -1: unowned).
//                             Compaction scans this dense array (16*N_alloc bytes) instead of the
//                             whole hash table (20*S bytes in the reference).
// voxels    Voxel[N*512]      reference layout, block = 4 KB contiguous, index z*64+y*8+x.
// compact16 int4[N]           visible list {x, y, z, ptr} consumed by integrate / raycast splat.
// compact20 VoxelEntry[N]     the same list in the reference's 20-byte layout (the buffer the
//                             reference shares with OpenGL, SDFRenderer.cpp:36).
struct Counters {
    int heapCounter;              // ref d_heapCounter
    int compactCount;             // ref d_compactifiedHashCounter[0]
    int overflowUsed;
    int dropped;
    int lastInserted;
    int icpConverged;             // set when the residual sum is exactly 0 (CameraTracking.cpp:55)
    unsigned int icpTicket;       // last-CTA-done ticket of the ICP reduction
    unsigned int icpSeq;          // sequence number of the fused cross-GPU exchange in the ICP tail
    unsigned long long numUpdated;
    int heapLow;                  // lowest heapCounter ever reached: block ids <= heapLow were never handed out
    int gcFreed;                  // blocks released by the last garbage-collection pass
    int streamCount;              // blocks moved by the last stream-out / stream-in pass
    int meshCount;                // triangles produced by the last mesh extraction
    int arenaLeaked;              // overflow-arena slots taken by an append that lost its race and were never linked (lost until vh_reset)
    int exchangeTimeouts;         // cross-GPU exchanges given up after kPeerSpinCycles (a peer died or never launched): the Align stops
};

struct FrameParams {
    float pose[16];   // camera -> world, row-major (HashTableParams::global_transform)
    float inv[16];    // world -> camera (inv_global_transform), adjugate inverse as SDF_Hashtable.cpp:15
    float proj[16];   // Fixed integration: rows 0-2 = K * inv[0:3,:] * diag(voxelSize,voxelSize,voxelSize,1), voxel INDEX -> (u*z, v*z, z)
};

struct IcpState {
    float delta[16];      // input frame -> target frame, row-major (CameraTracking::deltaTransform)
    float system[32];     // last reduced vh_icp_system
    float twist[6];       // Solver::estimate, refreshed lazily by vh_icp_get
    int iterations;
    int pad;
};

struct View {
    int4* entries;
    int* chain;
    int* mutex;
    unsigned int* heap;
    int4* blockInfo;
    Voxel* voxels;
    int4* compact16;
    VoxelEntry* compact20;
    Counters* ctr;
    const FrameParams* frame;

    unsigned int numBuckets, bucketSize, chainMax, numVoxelBlocks, numSlots, overflowSlots;
    float voxelSize, invVoxelSize, truncation, truncScale, wMax, wSample;
    float depthMin, depthMax, invDepthRange, depthScale;
    float wA, wB;                        // Fixed sample weight: w = max(wA * depth + wB, 1)
    float zFar;                          // Fixed integration: no voxel deeper than depthMax + truncation(depthMax) can be updated
    int W, H;
    float fx, fy, cx, cy;
    float K[9], Kinv[9];                 // tracking-side intrinsics (SetCameraIntrinsic)
    float wr, hb, nl, nr, nt, nb, rad;   // Fixed frustum constants (DESIGN.md "visibility")
    int partCount, partRank;
    float icpDistThres, icpNormalThres;
    // optional bilateral depth filter (Fixed): spatial weights g[|d|], range LUT over |delta depth| in raw units, scratch image
    float bilatG[3];
    const float* bilatLut;           // kBilatLut entries, nullptr = filter off
    float* depthSmooth;              // W x H filtered depth, raw units as float
    unsigned long long* tl;          // kernel timeline log (builds with -DVH_TIMELINE and VH_TIMELINE=1 in the environment), else nullptr
};

constexpr int kIcpMaxBlocks = 1024;
// exchange rows of the persistent Align kernel (k_track.cu): 2 parities x (one row per CTA + one row per group of 16 CTAs) x 32 words
constexpr size_t kIcpLLWords = (size_t)2 * kIcpMaxBlocks * 32 + (size_t)2 * (kIcpMaxBlocks / 16) * 32;
constexpr unsigned long long kTimelineCap = 1ull << 16;
constexpr int kBilatLut = 1024;      // |delta depth| >= this many raw units contributes nothing

// Fused all-reduce of the ICP normal equations over NVLink peer memory.
// buf[p] = rank p's exchange region mapped into this process: float data[2][8][32], then unsigned flag[2][8].
constexpr int kMaxPeers = 8;
// exchange region: 2 parity slots x kMaxPeers ranks x 32 values x {float value, u32 sequence}
constexpr size_t kPeerBytes = 2 * kMaxPeers * 32 * 8;
struct PeerView { int world, rank; float* buf[kMaxPeers]; int* timeouts; };
// a rank that waits longer than this for a peer's contribution gives up (about 5 s at 2 GHz) instead of hanging the GPU
constexpr long long kPeerSpinCycles = 10000000000ll;

struct Pose16f { float m[16]; };

}  // namespace vh

struct vh_context {
    vh_config cfg;
    int device;
    vh::View v;
    vh::FrameParams* frame;
    vh::IcpState* icp;
    float* icpPartials;       // kIcpMaxBlocks x 32
    unsigned long long* icpLL;   // persistent Align (k_track.cu): 2 x kIcpMaxBlocks x 32 {value, sequence} words
    int icpCtas;              // CTAs of the persistent Align kernel (0 = one per SM); VH_ICP_CTAS / vh_set_tuning
    int fusionReserveSMs;     // SMs the persistent fusion kernels leave free (for a co-resident Align grid); VH_FUSION_RESERVE_SMS / vh_set_tuning
    int* tileMin;             // raycast ray intervals, (W/8) x (H/8), float bit patterns
    int* tileMax;
    int numSMs;
    size_t bytesAllocated;
    // host-visible copy of the RefExact fusion-side projection flag etc.
    bool intrinsicsSet;
    vh::PeerView peers;       // world <= 1: single GPU
};

namespace vh {

// launchers (each is stream-ordered, no sync). policy dispatch happens inside.
cudaError_t launch_reset(vh_context* c, cudaStream_t s);
cudaError_t launch_set_frame_host(vh_context* c, const float* pose16, cudaStream_t s);
cudaError_t launch_set_frame_device(vh_context* c, const float* d_pose, const float* d_delta, float* d_poseOut, cudaStream_t s);
cudaError_t launch_alloc(vh_context* c, const float4* verts, cudaStream_t s);
cudaError_t launch_alloc_depth16(vh_context* c, const uint16_t* depth, cudaStream_t s);   // back-projects in registers (2 B / pixel)
cudaError_t launch_alloc_depthf(vh_context* c, const float* depthf, cudaStream_t s);      // from the dense metric depth (4 B / pixel)
bool alloc_depthf_ok(const vh_context* c);
cudaError_t launch_reset_mutex(vh_context* c, cudaStream_t s);
cudaError_t launch_compact(vh_context* c, cudaStream_t s);
cudaError_t launch_gc(vh_context* c, int scope, float sdfThreshold, float weightDecay, cudaStream_t s);
cudaError_t launch_stream_out(vh_context* c, const float* center, float radius, VoxelEntry* entriesOut, Voxel* voxelsOut, int capacity,
                              cudaStream_t s);
cudaError_t launch_stream_in(vh_context* c, const VoxelEntry* entries, const Voxel* voxels, int count, cudaStream_t s);
cudaError_t launch_integrate(vh_context* c, const float4* verts, const float* depthf, int countOverride, cudaStream_t s);
cudaError_t launch_preprocess(vh_context* c, const uint16_t* depth, float4* verts, float4* normals, float* depthf, cudaStream_t s);
cudaError_t launch_icp_iter(vh_context* c, const float4* in, const float4* inN, const float4* tg, const float4* tgN,
                            int row0, int row1, vh_icp_system* d_out, bool solve, cudaStream_t s);
cudaError_t launch_icp_iter_ex(vh_context* c, const float4* in, const float4* inN, const float4* tg, const float4* tgN,
                               int row0, int row1, vh_icp_system* d_out, bool solve, bool first, bool chained, cudaStream_t s);
cudaError_t launch_icp_solve(vh_context* c, const vh_icp_system* d_sys, cudaStream_t s);
// whole Align in one persistent cooperative launch (k_track.cu); peers: fuse the cross-GPU all-reduce (vh_set_peers);
// d_depth (optional): fused pre-processing -- the prologue turns the raw frame into the maps in / inN / d_depthf (which
// it WRITES) before tracking them against tg / tgN;
// d_poseOut (optional, may alias d_poseIn): 16 floats, row-major camera -> world, = d_poseIn * delta at the end of the launch
cudaError_t launch_icp_align(vh_context* c, const uint16_t* d_depth, float* d_depthf, const float4* in, const float4* inN,
                             const float4* tg, const float4* tgN, int row0, int row1, int iterations, bool peers,
                             const float* d_poseIn, float* d_poseOut, cudaStream_t s);
cudaError_t launch_icp_reset(vh_context* c, bool resetDelta, cudaStream_t s);
cudaError_t launch_icp_set_twist(vh_context* c, const float* twist6, cudaStream_t s);
cudaError_t launch_icp_twist(vh_context* c, cudaStream_t s);
cudaError_t launch_find_corr(vh_context* c, const float4* in, const float4* inN, const float4* tg, const float4* tgN,
                             const float* delta16_host, float4* corr, float4* corrN, float* res, float* d_err, cudaStream_t s);
cudaError_t launch_jacobians(vh_context* c, const float4* corr, const float4* corrN, float* J, cudaStream_t s);
cudaError_t launch_reduce_corr(vh_context* c, const float4* corr, const float4* corrN, const float* res,
                               vh_icp_system* d_out, cudaStream_t s);
cudaError_t launch_linear_system_300(vh_context* c, const float4* in, const float4* corr, const float4* corrN,
                                     float* d_out, cudaStream_t s);
cudaError_t launch_raycast(vh_context* c, float4* verts, float4* normals, cudaStream_t s);
cudaError_t launch_extract_mesh(vh_context* c, float* tris, int capacity, int* d_counter, cudaStream_t s);
cudaError_t launch_export_entries(vh_context* c, VoxelEntry* d_out, int* d_count, cudaStream_t s);

}  // namespace vh

#endif
