// k_table.cu -- table/heap initialisation, per-frame pose upload, export.
// Replaces resetHashTableKernel / resetHeapKernel / deviceAllocate (ref VoxelUtils.cu:151-211) and
// the host-side pose handling of SDF_Hashtable::integrate (ref SDF_Hashtable.cpp:15-21).
#include "vh_device.cuh"

namespace vh {

// One pass initialises every array: slots free (Q8), chains empty, heap[i] = i, owners unset.
__global__ void k_reset(View v) {
    const unsigned total = v.numSlots + v.overflowSlots;
    const unsigned stride = gridDim.x * blockDim.x;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        v.entries[i] = freeSlot();          // ref :155-157: offset 0, ptr -1, pos INT_MAX^3
        v.chain[i] = 0;
    }
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < v.numVoxelBlocks; i += stride) {
        v.heap[i] = i;                      // ref :165
        v.blockInfo[i] = make_int4(0, 0, 0, -1);
    }
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < v.numBuckets; i += stride) v.mutex[i] = 0;   // ref :199
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        Counters c{};
        c.heapCounter = (int)v.numVoxelBlocks - 1;   // ref :207
        c.heapLow = c.heapCounter;
        c.icpSeq = v.ctr->icpSeq;                    // the peers' mailboxes still hold the old sequence numbers
        *v.ctr = c;
    }
}

// 4x4 adjugate inverse, same term order as float4x4::getInverse (include/vh/types.h), which
// follows cuda_SimpleMatrixUtil.h:944-1069.  Executed by one thread.
__device__ void inverse4(const float* e, float* out) {
    float adj[16];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
            int R[3], C[3];
            for (int i = 0, k = 0; i < 4; ++i) if (i != c) R[k++] = i;
            for (int j = 0, k = 0; j < 4; ++j) if (j != r) C[k++] = j;
            float p0 = e[R[0] * 4 + C[0]] * e[R[1] * 4 + C[1]] * e[R[2] * 4 + C[2]];
            float n0 = e[R[0] * 4 + C[0]] * e[R[1] * 4 + C[2]] * e[R[2] * 4 + C[1]];
            float n1 = e[R[1] * 4 + C[0]] * e[R[0] * 4 + C[1]] * e[R[2] * 4 + C[2]];
            float p1 = e[R[1] * 4 + C[0]] * e[R[0] * 4 + C[2]] * e[R[2] * 4 + C[1]];
            float p2 = e[R[2] * 4 + C[0]] * e[R[0] * 4 + C[1]] * e[R[1] * 4 + C[2]];
            float n2 = e[R[2] * 4 + C[0]] * e[R[0] * 4 + C[2]] * e[R[1] * 4 + C[1]];
            adj[r * 4 + c] = ((r + c) & 1) ? (-p0 + n0 + n1 - p1 - p2 + n2) : (p0 - n0 - n1 + p1 + p2 - n2);
        }
    float det = e[0] * adj[0] + e[1] * adj[4] + e[2] * adj[8] + e[3] * adj[12];
    float rdet = 1.0f / det;
    for (int i = 0; i < 16; ++i) out[i] = adj[i] * rdet;
}

typedef Pose16f Pose16;

// Start of a frame: publish pose + inverse, clear the per-frame counters.
// pose = hostPose                     (d_pose == nullptr)
//      = d_pose                       (d_delta == nullptr)
//      = d_pose * d_delta             (camera->world chain: T_k = T_{k-1} * delta)
__global__ void k_set_frame(View v, FrameParams* frame, Pose16 hostPose, const float* d_pose, const float* d_delta,
                            float* d_poseOut) {
    VH_TL(TL_SET_FRAME, 0);
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float p[16];
    if (d_pose == nullptr) {
        for (int i = 0; i < 16; ++i) p[i] = hostPose.m[i];
    } else if (d_delta == nullptr) {
        for (int i = 0; i < 16; ++i) p[i] = d_pose[i];
    } else {
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c)
                p[r * 4 + c] = d_pose[r * 4 + 0] * d_delta[0 * 4 + c] + d_pose[r * 4 + 1] * d_delta[1 * 4 + c] +
                               d_pose[r * 4 + 2] * d_delta[2 * 4 + c] + d_pose[r * 4 + 3] * d_delta[3 * 4 + c];
    }
    float inv[16];
    inverse4(p, inv);
    for (int i = 0; i < 16; ++i) { frame->pose[i] = p[i]; frame->inv[i] = inv[i]; }
    // Fixed integration matrix: voxel index (i,j,k,1) -> (u*z, v*z, z) in one 3x4 product
    for (int c = 0; c < 4; ++c) {
        const float sc = c < 3 ? v.voxelSize : 1.0f;
        frame->proj[0 + c] = fmaf(v.fx, inv[0 + c], v.cx * inv[8 + c]) * sc;
        frame->proj[4 + c] = fmaf(v.fy, inv[4 + c], v.cy * inv[8 + c]) * sc;
        frame->proj[8 + c] = inv[8 + c] * sc;
        frame->proj[12 + c] = 0.0f;
    }
    if (d_poseOut) for (int i = 0; i < 16; ++i) d_poseOut[i] = p[i];
    v.ctr->compactCount = 0;     // ref flattenIntoBuffer: cudaMemset(counter, 0), VoxelUtils.cu:760
    v.ctr->numUpdated = 0ull;
    v.ctr->lastInserted = 0;
    VH_TL(TL_SET_FRAME, 1);
}

__global__ void k_reset_mutex(View v) {   // ref resetHashTableMutexes, VoxelUtils.cu:146-149
    const unsigned stride = gridDim.x * blockDim.x;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < v.numBuckets; i += stride) v.mutex[i] = 0;
}

// All allocated entries in the reference's 20-byte layout (for tests / checkpoints).
__global__ void k_export_entries(View v, VoxelEntry* out, int* count) {
    const unsigned total = v.numSlots + v.overflowSlots;
    const unsigned stride = gridDim.x * blockDim.x;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        int4 e = v.entries[i];
        if (e.w == VH_FREE_BLOCK) continue;
        int k = atomicAdd(count, 1);
        VoxelEntry o;
        o.pos = make_int3(e.x, e.y, e.z);
        o.ptr = e.w;
        o.offset = v.chain[i];
        out[k] = o;
    }
}

static int gridFor(const vh_context* c, size_t n, int threads) {
    size_t b = (n + threads - 1) / threads;
    size_t cap = (size_t)c->numSMs * 8;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

cudaError_t launch_reset(vh_context* c, cudaStream_t s) {
    size_t n = (size_t)c->v.numSlots + c->v.overflowSlots;
    if (c->v.numVoxelBlocks > n) n = c->v.numVoxelBlocks;
    k_reset<<<gridFor(c, n, 256), 256, 0, s>>>(c->v);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    // quirk Q13: the reference never clears the voxel heap; a defined result needs zeros
    return cudaMemsetAsync(c->v.voxels, 0, sizeof(Voxel) * 512 * (size_t)c->v.numVoxelBlocks, s);
}

cudaError_t launch_set_frame_host(vh_context* c, const float* pose16, cudaStream_t s) {
    Pose16 p;
    for (int i = 0; i < 16; ++i) p.m[i] = pose16[i];
    k_set_frame<<<1, 32, 0, s>>>(c->v, c->frame, p, nullptr, nullptr, nullptr);
    return cudaGetLastError();
}

cudaError_t launch_set_frame_device(vh_context* c, const float* d_pose, const float* d_delta, float* d_poseOut,
                                    cudaStream_t s) {
    Pose16 p{};
    k_set_frame<<<1, 32, 0, s>>>(c->v, c->frame, p, d_pose, d_delta, d_poseOut);
    return cudaGetLastError();
}

cudaError_t launch_reset_mutex(vh_context* c, cudaStream_t s) {
    k_reset_mutex<<<gridFor(c, c->v.numBuckets, 256), 256, 0, s>>>(c->v);
    return cudaGetLastError();
}

cudaError_t launch_export_entries(vh_context* c, VoxelEntry* d_out, int* d_count, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(d_count, 0, sizeof(int), s);
    if (e != cudaSuccess) return e;
    size_t n = (size_t)c->v.numSlots + c->v.overflowSlots;
    k_export_entries<<<gridFor(c, n, 256), 256, 0, s>>>(c->v, d_out, d_count);
    return cudaGetLastError();
}

}  // namespace vh

#ifdef VH_TIMELINE
extern "C" int vh_timeline_read(vh_context* c, unsigned long long* host, int cap, int reset) {
    if (!c || !c->v.tl) return -1;
    unsigned long long n = 0;
    cudaMemcpy(&n, c->v.tl, sizeof(n), cudaMemcpyDeviceToHost);
    if (n > vh::kTimelineCap) n = vh::kTimelineCap;
    if ((long long)n > cap) n = (unsigned long long)cap;
    cudaMemcpy(host, c->v.tl + 1, sizeof(unsigned long long) * n, cudaMemcpyDeviceToHost);
    if (reset) cudaMemset(c->v.tl, 0, sizeof(unsigned long long));
    return (int)n;
}
#endif
