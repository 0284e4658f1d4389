// k_alloc.cu -- per-frame hash-block allocation (north_star (a)).
// Replaces allocBlocksKernel + insertVoxelEntry + allocSingleBlockInHeap
// (ref VoxelUtils.cu:328-334, :418-541, :606-716).
//
// The reference launches one thread per pixel and lets all ~1500 pixels that see the same block
// probe the same bucket and fight over the same mutex word.  Here a request goes through three
// filters before it touches HBM:
//   1. warp:   __match_any_sync on the block hash, one leader per distinct key;
//   2. CTA:    a 128-entry shared-memory filter updated with 64-bit atomicExch (a key is dropped
//              only if the identical key was put there by a thread that went on to stage 3);
//   3. global: read-only probe of the bucket (one LDG.128 per slot); only a miss goes on to claim.
// Fixed policy claims a free slot lock-free with ONE 128-bit CAS that publishes the key together
// with the claim ({INT_MAX^3, FREE} -> {x, y, z, LOCKED}), so a concurrent request for the same
// block recognises it without waiting; the heap pop and the ptr store follow.  Buckets that are
// full spill into the overflow arena through a lock-free chain append (CAS on the tail's link).
// RefExact keeps the reference's rule bit for bit: a per-bucket atomicExch try-lock that is never
// released inside a frame, so at most one new block per bucket per frame (quirk Q4).
#include <cooperative_groups.h>

#include "vh_device.cuh"

namespace vh {

namespace cg = cooperative_groups;

constexpr int kFilterSize = 128;
constexpr unsigned long long kFilterEmpty = ~0ull;

__device__ __forceinline__ bool packKey(int x, int y, int z, unsigned long long& out) {
    const int lim = 1 << 20;
    if (x < -lim || x >= lim || y < -lim || y >= lim || z < -lim || z >= lim) return false;
    out = ((unsigned long long)(unsigned)(x & 0x1FFFFF)) | ((unsigned long long)(unsigned)(y & 0x1FFFFF) << 21) |
          ((unsigned long long)(unsigned)(z & 0x1FFFFF) << 42);
    return true;
}

// Warp-aggregated: the lanes of a warp that reach this point together pop the free list with ONE atomicSub.  (r2 ncu, 355 k
// inserts in one launch: 2.24 L2 atomic sectors per new block -- slot CAS, heap pop, insert count -- with or without the
// aggregation: after the match / filter / probe stages the surviving lanes of a warp rarely get here in the same cycle.)
__device__ __forceinline__ bool finishInsert(const View& v, unsigned slot, int x, int y, int z) {
    const cg::coalesced_group g = cg::coalesced_threads();
    int base = 0;
    if (g.thread_rank() == 0) base = atomicSub(&v.ctr->heapCounter, (int)g.size());   // ref allocSingleBlockInHeap :331, size() pops at once
    const int addr = g.shfl(base, 0) - (int)g.thread_rank();
    if (addr < 0) {                                         // heap exhausted: the reference reads out of bounds here (Q6)
        atomicAdd(&v.ctr->heapCounter, 1);
        // ALWAYS a tombstone {key, FREE}, never back to "never used": while this slot was LOCKED a peer with another key may
        // have claimed the NEXT slot (and the last heap block) -- a never-used slot in front of a live one would end every
        // scan early and hide that block from lookups (raycast, mesh) and from later inserts of its own key.  Tombstones
        // are claimable and keep the scan going.
        v.entries[slot] = make_int4(x, y, z, VH_FREE_BLOCK);
        atomicAdd(&v.ctr->dropped, 1);
        return false;
    }
    unsigned id = v.heap[addr];                             // ref :333
    v.blockInfo[id] = make_int4(x, y, z, (int)slot);
    reinterpret_cast<volatile int*>(v.entries + slot)[3] = (int)(id * 512u);   // ref :449-451 (ptr = id*512)
    const cg::coalesced_group ok = cg::coalesced_threads();
    if (ok.thread_rank() == 0) atomicAdd(&v.ctr->lastInserted, (int)ok.size());
    return true;
}

// Slot states: never used {INT_MAX^3, FREE}; live {key, ptr}; being inserted {key, LOCKED}; tombstone {old key, FREE}
// (left by garbage collection, k_gc.cu, or by an insert that found the heap empty).  Never-used and tombstone slots
// are both claimable with one 128-bit CAS.  A key is placed in the FIRST claimable slot of its bucket-then-chain scan
// order, and only after the scan has shown the key is not stored further on -- so concurrent requests for one key
// always meet at the same slot, and a reused tombstone can never sit in front of a live copy of the same key.
// Slots fill in scan order, so a never-used slot ends the scan: nothing was ever stored behind it.
// Returns the slot the key lives in (already present, or just inserted: fresh = true), -1 when the request was dropped.
__device__ int insertFixed(const View& v, int x, int y, int z, bool& fresh) {
    fresh = false;
    const unsigned h = bucketOf(v, x, y, z);
    const unsigned base = h * v.bucketSize;
    const int4 want = make_int4(x, y, z, VH_LOCKED_BLOCK);
    for (int attempt = 0; attempt < 64; ++attempt) {
        int claim = -1;
        int4 claimSeen = freeSlot();
        bool ended = false;                                 // hit a never-used slot: end of everything stored
        for (unsigned i = 0; i < v.bucketSize && !ended; ++i) {
            const int4 e = ldSlot(v.entries + base + i);
            if (e.w != VH_FREE_BLOCK) {
                if (sameKey(e, x, y, z)) return (int)(base + i);   // present (or being inserted by a peer)
                continue;
            }
            if (claim < 0) { claim = (int)(base + i); claimSeen = e; }
            ended = e.x == VH_FREE_COORD;
        }
        unsigned cur = base + v.bucketSize - 1;
        unsigned len = 0;
        bool chainFull = false;
        if (!ended) {                                       // bucket full of live / tombstone slots: walk the overflow chain
            while (true) {
                const int off = *reinterpret_cast<volatile int*>(v.chain + cur);
                if (off == 0) break;
                cur += (unsigned)off;
                ++len;
                const int4 e = ldSlot(v.entries + cur);
                if (e.w != VH_FREE_BLOCK) {
                    if (sameKey(e, x, y, z)) return (int)cur;
                } else if (claim < 0) { claim = (int)cur; claimSeen = e; }
            }
            chainFull = len >= v.chainMax;
        }
        if (claim >= 0) {
            int4 old;
            if (casSlot(v.entries + claim, claimSeen, want, old)) {
                fresh = finishInsert(v, (unsigned)claim, x, y, z);
                return fresh ? claim : -1;
            }
            if (old.w != VH_FREE_BLOCK && sameKey(old, x, y, z)) return claim;   // a peer won the slot with the same key
            continue;                                       // someone else's key took it: rescan
        }
        // nothing claimable: extend the chain with a fresh arena slot (lock-free append, CAS on the tail's link)
        if (chainFull) { atomicAdd(&v.ctr->dropped, 1); return -1; }
        const int a = atomicAdd(&v.ctr->overflowUsed, 1);
        if (a >= (int)v.overflowSlots) { atomicSub(&v.ctr->overflowUsed, 1); atomicAdd(&v.ctr->dropped, 1); return -1; }
        const int mySlot = (int)(v.numSlots + (unsigned)a);
        v.entries[mySlot] = want;
        __threadfence();                                    // entry visible before it can be reached through the link
        while (true) {
            const int prev = atomicCAS(v.chain + cur, 0, mySlot - (int)cur);
            if (prev == 0) { fresh = finishInsert(v, (unsigned)mySlot, x, y, z); return fresh ? mySlot : -1; }
            // a peer appended first: step onto its entry and try again behind it
            cur += (unsigned)prev;
            ++len;
            const int4 e = ldSlot(v.entries + cur);
            if ((e.w != VH_FREE_BLOCK && sameKey(e, x, y, z)) || len >= v.chainMax) {
                const bool present = e.w != VH_FREE_BLOCK && sameKey(e, x, y, z);
                if (!present) atomicAdd(&v.ctr->dropped, 1);
                v.entries[mySlot] = make_int4(x, y, z, VH_FREE_BLOCK);     // never linked: invisible, leaked -- and counted
                atomicAdd(&v.ctr->arenaLeaked, 1);
                return present ? (int)cur : -1;
            }
        }
    }
    atomicAdd(&v.ctr->dropped, 1);                          // 64 lost races in a row: give up on this request
    return -1;
}

__device__ void insertRefExact(const View& v, int x, int y, int z) {
    const unsigned h = bucketOf(v, x, y, z);
    const unsigned base = h * v.bucketSize;
    for (unsigned i = 0; i < v.bucketSize; ++i) {
        unsigned idx = (base + i) % v.numSlots;                              // ref :437-438
        int4 e = ldSlot(v.entries + idx);
        if (sameKey(e, x, y, z) && e.w != VH_FREE_BLOCK) return;             // ref :440-441
        if (e.w == VH_FREE_BLOCK) {                                          // ref :442
            int prev = atomicExch(v.mutex + h, VH_LOCKED_BLOCK);             // ref :444
            if (prev != VH_LOCKED_BLOCK) {                                   // ref :445
                v.entries[idx] = make_int4(x, y, z, VH_FREE_BLOCK);          // ref :447-448 (pos, offset)
                int addr = atomicSub(&v.ctr->heapCounter, 1);                // ref :331
                if (addr < 0) { atomicAdd(&v.ctr->dropped, 1); return; }     // Q6: pos stays written, ptr stays -1
                unsigned id = v.heap[addr];
                v.blockInfo[id] = make_int4(x, y, z, (int)idx);
                reinterpret_cast<volatile int*>(v.entries + idx)[3] = (int)(id * 512u);   // ref :449-451
                atomicAdd(&v.ctr->lastInserted, 1);
                return;
            }
        }
    }
}

// Warp-cooperative request. Must be called by every lane of the warp (converged).
template <class P>
__device__ __forceinline__ void warpRequest(const View& v, const float* pose, unsigned long long* filter, bool have, int x, int y, int z) {
    const unsigned lane = threadIdx.x & 31;
    // partitioned hash space: drop what another rank owns BEFORE the warp / CTA de-duplication, so (P-1)/P of the
    // requests never reach the match, the filter or the table (the scan itself is replicated on every rank)
    if (P::fixed) have = have && ownedHere(v, x, y, z);
    const unsigned mask = __ballot_sync(0xffffffffu, have);
    if (!have) return;
    const unsigned h32 = blockHash32(x, y, z);
    const unsigned peers = __match_any_sync(mask, h32);
    const int leader = __ffs(peers) - 1;
    const int lx = __shfl_sync(peers, x, leader), ly = __shfl_sync(peers, y, leader), lz = __shfl_sync(peers, z, leader);
    const bool same = lx == x && ly == y && lz == z;
    if ((int)lane != leader && same) return;                // duplicates of the leader's key are done
    unsigned long long packed;
    if (packKey(x, y, z, packed)) {
        unsigned long long old = atomicExch(filter + ((h32 ^ (h32 >> 7)) & (kFilterSize - 1)), packed);
        if (old == packed) return;                          // already forwarded by this CTA
    }
    if (P::fixed) {
        bool fresh;
        insertFixed(v, x, y, z, fresh);
    } else {
        if (!refBlockInFrustum(v, pose, x, y, z)) return;   // ref :673; depends on (key, pose) only
        insertRefExact(v, x, y, z);
    }
}

// Where the pixel's camera-space point comes from (pre-processing fused into its consumers):
//   SRC_VERTS   the float4 vertex map (the reference's interface, allocBlocks(verts, normals)): 16 B per pixel
//   SRC_DEPTH16 the raw u16 depth image, back-projected in registers with the operations of k_preprocess in the same
//               order (metricDepth, K^-1 (x, y, 1), scale): 2 B per pixel, bit-identical points
//   SRC_DEPTHF  the dense metric depth k_preprocess leaves behind for the integration (valid as a source when the last
//               row of K^-1 is (0, 0, 1), so that depthf == the metric depth bit for bit): 4 B per pixel
enum { SRC_VERTS = 0, SRC_DEPTH16 = 1, SRC_DEPTHF = 2 };

template <class P, int SRC>
__global__ void __launch_bounds__(256) k_alloc(View v, const void* __restrict__ src) {
    __shared__ unsigned long long filter[kFilterSize];
    __shared__ float sPose[16];
    VH_TL(TL_ALLOC, 0);
    for (int i = threadIdx.x; i < kFilterSize; i += 256) filter[i] = kFilterEmpty;
    if (threadIdx.x < 16) sPose[threadIdx.x] = v.frame->pose[threadIdx.x];
    __syncthreads();

    const int px = blockIdx.x * 32 + (threadIdx.x & 31);
    const int py = blockIdx.y * 8 + (threadIdx.x >> 5);
    const bool inside = px < v.W && py < v.H;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (inside) {
        const size_t idx = (size_t)py * v.W + px;
        if (SRC == SRC_VERTS) p = __ldg(reinterpret_cast<const float4*>(src) + idx);
        else {
            float d;
            if (SRC == SRC_DEPTH16) {
                d = (float)__ldg(reinterpret_cast<const uint16_t*>(src) + idx) / v.depthScale;   // ref CameraTrackingUtils.cu:63-64
                if (P::fixed && !(d > v.depthMin && d < v.depthMax)) d = 0.0f;
            } else d = __ldg(reinterpret_cast<const float*>(src) + idx);
            const float3 k = mul3(v.Kinv, (float)px, (float)py, 1.0f);                          // ref :69-70
            p = make_float4(k.x * d, k.y * d, k.z * d, 1.0f);                                   // ref :72-73
        }
    }

    if (!P::fixed) {
        // ref :620-636: skip z == 0, transform by global_transform, one block per pixel (Q3)
        bool have = inside && !(p.z == 0.0f);
        int3 b = make_int3(0, 0, 0);
        if (have) {
            float4 w = mul4(sPose, p.x, p.y, p.z, p.w);
            b = refWorld2Block(v, w.x, w.y, w.z);
        }
        warpRequest<P>(v, sPose, filter, have, b.x, b.y, b.z);
        return;
    }

    // Fixed: voxel-block DDA over the truncation band [d - t, d + t] along the viewing ray
    // (what the reference's commented-out code at :637-703 set out to do).
    const float d = p.z;
    bool active = inside && (d > v.depthMin && d < v.depthMax);
    int cur[3] = {0, 0, 0}, end[3] = {0, 0, 0}, step[3] = {0, 0, 0};
    float tMax[3] = {0, 0, 0}, tDelta[3] = {0, 0, 0};
    if (active) {
        float tr = fmaf(v.truncScale, d, v.truncation);
        float s0 = (d - tr) / d, s1 = (d + tr) / d;
        float4 a = mul4(sPose, p.x * s0, p.y * s0, d * s0, 1.0f);
        float4 b = mul4(sPose, p.x * s1, p.y * s1, d * s1, 1.0f);
        float ga[3] = {(a.x * v.invVoxelSize + 0.5f) * 0.125f, (a.y * v.invVoxelSize + 0.5f) * 0.125f,
                       (a.z * v.invVoxelSize + 0.5f) * 0.125f};
        float gb[3] = {(b.x * v.invVoxelSize + 0.5f) * 0.125f, (b.y * v.invVoxelSize + 0.5f) * 0.125f,
                       (b.z * v.invVoxelSize + 0.5f) * 0.125f};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float fa = floorf(ga[k]), fb = floorf(gb[k]);
            cur[k] = f2i(fa); end[k] = f2i(fb);
            float dir = gb[k] - ga[k];
            if (dir > 0.0f) { step[k] = 1; tMax[k] = ((fa + 1.0f) - ga[k]) / dir; tDelta[k] = 1.0f / dir; }
            else if (dir < 0.0f) { step[k] = -1; tMax[k] = (fa - ga[k]) / dir; tDelta[k] = -1.0f / dir; }
            else { step[k] = 0; tMax[k] = INFINITY; tDelta[k] = INFINITY; }
        }
    }
    for (int iter = 0; iter < 32; ++iter) {
        if (!__any_sync(0xffffffffu, active)) break;
        warpRequest<P>(v, sPose, filter, active, cur[0], cur[1], cur[2]);
        if (active) {
            if (cur[0] == end[0] && cur[1] == end[1] && cur[2] == end[2]) active = false;
            else {
                int ax = (tMax[0] <= tMax[1] && tMax[0] <= tMax[2]) ? 0 : (tMax[1] <= tMax[2] ? 1 : 2);
                float tm = ax == 0 ? tMax[0] : (ax == 1 ? tMax[1] : tMax[2]);
                if (tm > 1.0f) active = false;
                else if (ax == 0) { cur[0] += step[0]; tMax[0] += tDelta[0]; }
                else if (ax == 1) { cur[1] += step[1]; tMax[1] += tDelta[1]; }
                else { cur[2] += step[2]; tMax[2] += tDelta[2]; }
            }
        }
    }
}

// ---- stream-in: blocks come back from a caller-owned buffer (k_gc.cu: k_stream_out) ---------------------------
// One CTA per incoming block: thread 0 inserts the key like any allocation request; a fresh block (zeroed) receives
// the 4 KB as they are, a block that was re-observed in the meantime is merged by the running-average rule
// sdf = (s1 w1 + s2 w2) / (w1 + w2), w = min(wMax, w1 + w2).
__global__ void __launch_bounds__(128, 8) k_stream_in(View v, const VoxelEntry* __restrict__ entries, const Voxel* __restrict__ voxels, int count) {
    __shared__ int sPtr, sFresh, sSrc;
    for (int b = blockIdx.x; b < count; b += gridDim.x) {
        if (threadIdx.x == 0) {
            const VoxelEntry e = entries[b];
            int ptr = -1;
            bool fresh = false;
            if (ownedHere(v, e.pos.x, e.pos.y, e.pos.z)) {
                const int slot = insertFixed(v, e.pos.x, e.pos.y, e.pos.z, fresh);
                if (slot >= 0) ptr = ldSlot(v.entries + slot).w;
            }
            if (ptr >= 0) atomicAdd(&v.ctr->streamCount, 1);
            sPtr = ptr;
            sFresh = fresh;
            sSrc = e.ptr;                                   // record index * 512 in the caller's voxel buffer
        }
        __syncthreads();
        const int ptr = sPtr;
        const bool fresh = sFresh != 0;
        const int srcIdx = sSrc;
        __syncthreads();
        if (ptr < 0) continue;
        const float4* src = reinterpret_cast<const float4*>(voxels + (size_t)srcIdx + threadIdx.x * 4);
        float4 a = src[0], c = src[1];
        float in[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
        float4* dst = reinterpret_cast<float4*>(v.voxels + (size_t)ptr + threadIdx.x * 4);
        if (!fresh) {
            float4 p = dst[0], q = dst[1];
            float cur[8] = {p.x, p.y, p.z, p.w, q.x, q.y, q.z, q.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float w1 = cur[2 * k + 1], w2 = in[2 * k + 1];
                const float wn = w1 + w2;
                if (wn > 0.0f) {
                    in[2 * k] = fmaf(cur[2 * k], w1, in[2 * k] * w2) * __frcp_rn(wn);
                    in[2 * k + 1] = fminf(v.wMax, wn);
                } else {
                    in[2 * k] = 0.0f;
                    in[2 * k + 1] = 0.0f;
                }
            }
        }
        dst[0] = make_float4(in[0], in[1], in[2], in[3]);
        dst[1] = make_float4(in[4], in[5], in[6], in[7]);
    }
}

__global__ void k_stream_in_begin(View v) { v.ctr->streamCount = 0; v.ctr->lastInserted = 0; }

cudaError_t launch_stream_in(vh_context* c, const VoxelEntry* entries, const Voxel* voxels, int count, cudaStream_t s) {
    k_stream_in_begin<<<1, 1, 0, s>>>(c->v);
    if (count <= 0) return cudaGetLastError();
    int grid = c->numSMs * 8;
    if (count < grid) grid = count;
    k_stream_in<<<grid, 128, 0, s>>>(c->v, entries, voxels, count);
    return cudaGetLastError();
}

template <int SRC>
static cudaError_t launchAllocFrom(vh_context* c, const void* src, cudaStream_t s) {
    dim3 grid((c->v.W + 31) / 32, (c->v.H + 7) / 8);
    if (c->cfg.policy == VH_POLICY_FIXED) k_alloc<Fixed, SRC><<<grid, 256, 0, s>>>(c->v, src);
    else k_alloc<RefExact, SRC><<<grid, 256, 0, s>>>(c->v, src);
    return cudaGetLastError();
}
cudaError_t launch_alloc(vh_context* c, const float4* verts, cudaStream_t s) { return launchAllocFrom<SRC_VERTS>(c, verts, s); }
cudaError_t launch_alloc_depth16(vh_context* c, const uint16_t* depth, cudaStream_t s) { return launchAllocFrom<SRC_DEPTH16>(c, depth, s); }
// depthf == metric depth only when K^-1 maps (x, y, 1) to z = 1 exactly
bool alloc_depthf_ok(const vh_context* c) { return c->v.Kinv[6] == 0.0f && c->v.Kinv[7] == 0.0f && c->v.Kinv[8] == 1.0f && c->v.bilatLut == nullptr; }
cudaError_t launch_alloc_depthf(vh_context* c, const float* depthf, cudaStream_t s) { return launchAllocFrom<SRC_DEPTHF>(c, depthf, s); }

}  // namespace vh
