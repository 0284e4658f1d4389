// k_gc.cu -- voxel starvation + block garbage collection.
//
// The reference has no working removal: deleteVoxelEntry (ref VoxelUtils.cu:544-604) returns false on a
// match and "deletes" on finding a free slot, and nothing calls it; removeSingleBlockInHeap (:336-341) is
// the mirror of the pop and is what the push below follows.  The pass is re-specified from the paper the
// reference implements (Niessner et al. 2013, section 4.4): for every block in scope, optionally age the
// voxel weights, reduce min |sdf| and max weight over the block, and release the block when it holds no
// observation (max weight == 0) or no voxel near a surface (min |sdf| >= threshold).
//
// One CTA of 128 threads per block, the layout of k_integrate: a thread owns one 32-byte sector (256-bit
// access), the two reductions are warp shuffles + four shared-memory words.  Releasing a block
//   * zeroes its 4 KB (the next owner of the id relies on zeros, as after vh_reset),
//   * turns its hash slot into a tombstone {key, FREE} -- chain links stay, so lookups and inserts keep
//     walking through it; k_alloc.cu reclaims tombstones,
//   * clears blockInfo[id].w (compaction skips it) and pushes the id back on the heap.
// The pass runs alone on its stream (no concurrent allocation), so distinct CTAs touch distinct slots.
// HBM-bound: 4 KB read per block in scope, + 4 KB written when weights are aged or the block is released.
#include "vh_device.cuh"

namespace vh {

struct F8g { float a[8]; };
__device__ __forceinline__ F8g ldSector(const Voxel* p) {
    F8g r;
    asm volatile("ld.global.L1::no_allocate.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(r.a[0]), "=f"(r.a[1]), "=f"(r.a[2]), "=f"(r.a[3]), "=f"(r.a[4]), "=f"(r.a[5]), "=f"(r.a[6]), "=f"(r.a[7])
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void stSector(Voxel* p, const F8g& r) {
    asm volatile("st.global.L1::no_allocate.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(r.a[0]), "f"(r.a[1]),
                 "f"(r.a[2]), "f"(r.a[3]), "f"(r.a[4]), "f"(r.a[5]), "f"(r.a[6]), "f"(r.a[7])
                 : "memory");
}

// scope 0: the compacted (visible) list of the last vh_compact; scope 1: every allocated block.
// The sector of the NEXT block of the CTA's stride is loaded before the current one is reduced (r1 ncu: the
// unpipelined form sat at 22 long-scoreboard stalls per issue, 4.0 TB/s on a read-only scan).
struct GcItem { int id; int4 info; F8g x; };

__device__ __forceinline__ void gcFetch(const View& v, int scope, int firstId, int b, int count, GcItem& it) {
    it.id = -1;
    it.info = make_int4(0, 0, 0, -1);
    if (b >= count) return;
    it.id = scope == 0 ? (__ldg(&v.compact16[b].w) >> 9) : firstId + b;    // ptr = id * 512
    it.info = v.blockInfo[it.id];                                           // plain load: this kernel rewrites it
    if (it.info.w >= 0) it.x = ldSector(v.voxels + (size_t)it.id * 512 + threadIdx.x * 4);
}

__global__ void __launch_bounds__(128, 12) k_gc(View v, int scope, float sdfThreshold, float weightDecay) {
    __shared__ float sMin[4], sMax[4];
    __shared__ int sFree;
    const int N = (int)v.numVoxelBlocks;
    // ids at or below the low-water mark were never handed out (k_gc_begin has just refreshed it; only this kernel
    // moves heapCounter now, upwards)
    const int firstId = max(v.ctr->heapLow + 1, 0);
    const int count = scope == 0 ? v.ctr->compactCount : N - firstId;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    GcItem cur, nxt;
    gcFetch(v, scope, firstId, blockIdx.x, count, cur);
    for (int b = blockIdx.x; b < count; b += gridDim.x) {
        gcFetch(v, scope, firstId, b + gridDim.x, count, nxt);
        if (cur.info.w >= 0) {                                              // CTA-uniform: released or never-used ids are skipped
            Voxel* sector = v.voxels + (size_t)cur.id * 512 + threadIdx.x * 4;
            F8g& x = cur.x;
            float mn = INFINITY, mx = 0.0f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float w = x.a[2 * k + 1];
                if (weightDecay > 0.0f) { w = fmaxf(w - weightDecay, 0.0f); x.a[2 * k + 1] = w; }
                if (w > 0.0f) { mn = fminf(mn, fabsf(x.a[2 * k])); mx = fmaxf(mx, w); }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            }
            if (lane == 0) { sMin[warp] = mn; sMax[warp] = mx; }
            __syncthreads();
            if (threadIdx.x == 0) {
                const float bmn = fminf(fminf(sMin[0], sMin[1]), fminf(sMin[2], sMin[3]));
                const float bmx = fmaxf(fmaxf(sMax[0], sMax[1]), fmaxf(sMax[2], sMax[3]));
                const bool release = bmx == 0.0f || bmn >= sdfThreshold;
                sFree = release;
                if (release) {
                    releaseBlock(v, cur.id, cur.info);
                    atomicAdd(&v.ctr->gcFreed, 1);
                }
            }
            __syncthreads();
            if (sFree) {
#pragma unroll
                for (int k = 0; k < 8; ++k) x.a[k] = 0.0f;
                stSector(sector, x);
            } else if (weightDecay > 0.0f) {
                stSector(sector, x);
            }
            // sMin / sMax / sFree are rewritten only behind the next block's first __syncthreads
        }
        cur = nxt;
    }
}

__global__ void k_gc_begin(View v) {
    Counters* c = v.ctr;
    c->heapLow = min(c->heapLow, c->heapCounter);   // pops only lower the counter between passes, pushes happen here
    c->gcFreed = 0;
}
__global__ void k_gc_end(View v) { v.ctr->compactCount = 0; }   // the visible list may name released blocks now

// ---- stream-out: blocks far from the region of interest leave the table for a caller-owned buffer ------------
// (Niessner et al. 2013, section 4.5: the active region is a sphere around the camera; blocks outside it are moved
// to host memory and come back through k_stream_in when the sphere reaches them again.)  The buffers only need to be
// device-ACCESSIBLE: pinned host memory works, the copy then goes straight over the host link.
__global__ void __launch_bounds__(128, 8) k_stream_out(View v, float cx, float cy, float cz, float radius2, VoxelEntry* entriesOut,
                                                       Voxel* voxelsOut, int capacity) {
    __shared__ int sDst;
    const int N = (int)v.numVoxelBlocks;
    const int firstId = max(v.ctr->heapLow + 1, 0);
    for (int id = firstId + blockIdx.x; id < N; id += gridDim.x) {
        // Thread 0 ALONE reads the owner record, decides and releases; the decision reaches the other warps through
        // shared memory.  (A CTA-wide load of blockInfo[id] raced with thread 0's releaseBlock(), which rewrites that
        // record: a late warp could see w = -1, skip the barriers below and leave sectors neither copied nor zeroed.)
        if (threadIdx.x == 0) {
            int dst = -1;
            const int4 info = v.blockInfo[id];
            if (info.w >= 0) {
                // centre of the block: voxel indices 8b .. 8b+7 sit at (8b + k) * voxelSize
                const float dx = ((float)(info.x * 8) + 3.5f) * v.voxelSize - cx;
                const float dy = ((float)(info.y * 8) + 3.5f) * v.voxelSize - cy;
                const float dz = ((float)(info.z * 8) + 3.5f) * v.voxelSize - cz;
                if (dx * dx + dy * dy + dz * dz > radius2) {
                    dst = atomicAdd(&v.ctr->streamCount, 1);
                    if (dst >= capacity) { atomicSub(&v.ctr->streamCount, 1); dst = -1; }   // buffer full: the block stays
                    else {
                        VoxelEntry e;
                        e.pos = make_int3(info.x, info.y, info.z);
                        e.ptr = dst * 512;
                        e.offset = 0;
                        entriesOut[dst] = e;
                        releaseBlock(v, id, info);
                    }
                }
            }
            sDst = dst;
        }
        __syncthreads();
        const int dst = sDst;
        __syncthreads();
        if (dst < 0) continue;                                               // CTA-uniform
        Voxel* sector = v.voxels + (size_t)id * 512 + threadIdx.x * 4;
        F8g x = ldSector(sector);
        float4* out = reinterpret_cast<float4*>(voxelsOut + (size_t)dst * 512 + threadIdx.x * 4);   // may be host memory: plain stores
        out[0] = make_float4(x.a[0], x.a[1], x.a[2], x.a[3]);
        out[1] = make_float4(x.a[4], x.a[5], x.a[6], x.a[7]);
#pragma unroll
        for (int k = 0; k < 8; ++k) x.a[k] = 0.0f;
        stSector(sector, x);
    }
}

__global__ void k_stream_begin(View v) {
    Counters* c = v.ctr;
    c->heapLow = min(c->heapLow, c->heapCounter);
    c->streamCount = 0;
}

cudaError_t launch_stream_out(vh_context* c, const float* center, float radius, VoxelEntry* entriesOut, Voxel* voxelsOut, int capacity,
                              cudaStream_t s) {
    k_stream_begin<<<1, 1, 0, s>>>(c->v);
    k_stream_out<<<c->numSMs * 8, 128, 0, s>>>(c->v, center[0], center[1], center[2], radius * radius, entriesOut, voxelsOut, capacity);
    k_gc_end<<<1, 1, 0, s>>>(c->v);
    return cudaGetLastError();
}

cudaError_t launch_gc(vh_context* c, int scope, float sdfThreshold, float weightDecay, cudaStream_t s) {
    k_gc_begin<<<1, 1, 0, s>>>(c->v);
    k_gc<<<c->numSMs * 12, 128, 0, s>>>(c->v, scope, sdfThreshold, weightDecay);
    k_gc_end<<<1, 1, 0, s>>>(c->v);
    return cudaGetLastError();
}

}  // namespace vh
