// k_compact.cu -- visible-block compaction (north_star (b)).
// Replaces flattenKernel / flattenIntoBuffer (ref VoxelUtils.cu:719-768).
//
// The reference resets a 20*S-byte output buffer, scans all S hash slots, bumps a shared-memory
// counter with one atomicAdd per kept entry and reads the count back to the host.  Here the scan
// runs over the dense per-block owner array (16 bytes per ALLOCATED block, ids handed out from
// N-1 downwards, ref :207/:331), the keep-mask is a warp ballot, ranks come from popc of the
// lower lanes, each warp reserves its output range with ONE global atomicAdd, and the count
// stays in device memory for the persistent integrate kernel to read.
#include "vh_device.cuh"

namespace vh {

template <class P>
__global__ void __launch_bounds__(256) k_compact(View v) {
    __shared__ float sM[16];
    VH_TL(TL_COMPACT, 0);
    if (threadIdx.x < 16) sM[threadIdx.x] = P::fixed ? v.frame->inv[threadIdx.x] : v.frame->pose[threadIdx.x];
    __syncthreads();
    const unsigned lane = threadIdx.x & 31;
    const int N = (int)v.numVoxelBlocks;
    // ids ever handed out are (min(heapLow, heapCounter), N-1] (garbage collection pushes ids back, so the live
    // ones are no longer a dense range: blockInfo.w < 0 marks the released ones); clamp for the exhausted-heap case
    int first = min(v.ctr->heapLow, v.ctr->heapCounter) + 1;
    if (first < 0) first = 0;
    const int stride = gridDim.x * blockDim.x;
    // round the loop so whole warps stay converged for the ballot
    for (int base = first + (int)(blockIdx.x * blockDim.x + threadIdx.x - lane); base < N; base += stride) {
        int id = base + (int)lane;
        bool keep = false;
        int4 info = make_int4(0, 0, 0, -1);
        if (id < N) {
            info = __ldg(v.blockInfo + id);
            if (info.w >= 0) {
                keep = P::fixed ? fixedBlockVisible(v, sM, info.x, info.y, info.z)
                                : refBlockInFrustum(v, sM, info.x, info.y, info.z);   // ref :732
            }
        }
        unsigned ballot = __ballot_sync(0xffffffffu, keep);
        if (ballot == 0) continue;
        int warpBase = 0;
        if (lane == 0) warpBase = atomicAdd(&v.ctr->compactCount, __popc(ballot));    // ref :740, one per warp
        warpBase = __shfl_sync(0xffffffffu, warpBase, 0);
        if (keep) {
            int dst = warpBase + __popc(ballot & ((1u << lane) - 1u));
            int ptr = id * 512;
            v.compact16[dst] = make_int4(info.x, info.y, info.z, ptr);
            VoxelEntry e;
            e.pos = make_int3(info.x, info.y, info.z);
            e.ptr = ptr;
            e.offset = __ldg(v.chain + info.w);
            v.compact20[dst] = e;                                                     // ref :747
        }
    }
}

cudaError_t launch_compact(vh_context* c, cudaStream_t s) {
    // worst case N blocks; persistent-style grid.  The scan is latency-bound (load -> frustum test -> ballot -> one
    // atomic per warp -> store), so it wants every resident warp: 8 CTAs of 256 threads per SM (C4, 384 k allocated
    // blocks: 26 us with 2 CTAs per SM)
    size_t blocks = ((size_t)c->v.numVoxelBlocks + 255) / 256;
    size_t cap = (size_t)c->numSMs * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    if (c->cfg.policy == VH_POLICY_FIXED) k_compact<Fixed><<<(int)blocks, 256, 0, s>>>(c->v);
    else k_compact<RefExact><<<(int)blocks, 256, 0, s>>>(c->v);
    return cudaGetLastError();
}

}  // namespace vh
