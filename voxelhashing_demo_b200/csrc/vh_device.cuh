// vh_device.cuh -- device-side arithmetic shared by the kernels.
//
// Every translation unit is compiled with -fmad=false (the reference's own flag,
// CMakeLists.txt:23): a*b+c below is TWO rounded operations.  Where a fused multiply-add is
// wanted (Fixed policy only) it is written as fmaf() and is part of the policy's definition in
// DESIGN.md, mirrored by the oracle with fmaf() as well.
#ifndef VH_DEVICE_CUH
#define VH_DEVICE_CUH

#include "vh_internal.h"

namespace vh {

struct RefExact { static constexpr bool fixed = false; };
struct Fixed    { static constexpr bool fixed = true;  };

__device__ __forceinline__ int f2i(float x) { return __float2int_rz(x); }     // cvt.rzi.s32.f32 (saturates, NaN->0)
__device__ __forceinline__ int d2i(double x) { return __double2int_rz(x); }   // cvt.rzi.s32.f64

// row-major M*v, summed left to right (cuda_SimpleMatrixUtil.h:888-896)
__device__ __forceinline__ float4 mul4(const float* __restrict__ m, float x, float y, float z, float w) {
    float4 r;
    r.x = m[0] * x + m[1] * y + m[2] * z + m[3] * w;
    r.y = m[4] * x + m[5] * y + m[6] * z + m[7] * w;
    r.z = m[8] * x + m[9] * y + m[10] * z + m[11] * w;
    r.w = m[12] * x + m[13] * y + m[14] * z + m[15] * w;
    return r;
}
__device__ __forceinline__ float3 mul3(const float* __restrict__ m, float x, float y, float z) {
    float3 r;
    r.x = m[0] * x + m[1] * y + m[2] * z;
    r.y = m[3] * x + m[4] * y + m[5] * z;
    r.z = m[6] * x + m[7] * y + m[8] * z;
    return r;
}

// ---- hash / ownership ----------------------------------------------------------------------
// VoxelUtils.cu:250-259: wrapping int products, XOR, UNSIGNED modulo (quirk Q7).
__device__ __forceinline__ unsigned int blockHash32(int x, int y, int z) {
    return ((unsigned)x * 73856093u) ^ ((unsigned)y * 19349669u) ^ ((unsigned)z * 83492791u);
}
__device__ __forceinline__ unsigned int bucketOf(const View& v, int x, int y, int z) {
    return blockHash32(x, y, z) % v.numBuckets;
}
// multi-GPU owner: an independent mix so ownership and bucket index are uncorrelated
__device__ __forceinline__ unsigned int ownerMix(int x, int y, int z) {
    unsigned u = ((unsigned)x * 0x9E3779B1u) ^ ((unsigned)y * 0x85EBCA77u) ^ ((unsigned)z * 0xC2B2AE3Du);
    u ^= u >> 16; u *= 0x7FEB352Du; u ^= u >> 15; u *= 0x846CA68Bu; u ^= u >> 16;
    return u;
}
__device__ __forceinline__ bool ownedHere(const View& v, int x, int y, int z) {
    if (v.partCount <= 1) return true;
    const unsigned u = ownerMix(x, y, z), P = (unsigned)v.partCount;
    // the test runs per lane per DDA step of the allocation scan: a mask instead of a division for 2, 4, 8 ranks
    return (int)((P & (P - 1u)) == 0u ? (u & (P - 1u)) : (u % P)) == v.partRank;
}

// ---- RefExact coordinate maps (VoxelUtils.cu:266-309) -----------------------------------------
__device__ __forceinline__ int refVoxelCoord(float p, float voxelSize) {
    float q = p / voxelSize;                                 // :283
    int s = f2i(copysignf(1.0f, q));                         // :284
    float off = (float)((double)s * 0.5);                    // :285 (int * double literal -> float)
    return f2i(q + off);
}
__device__ __forceinline__ int refFloorDiv8(int a) {         // :274-277
    if (a < 0) a -= 7;
    return a / 8;
}
__device__ __forceinline__ int3 refWorld2Block(const View& v, float x, float y, float z) {
    return make_int3(refFloorDiv8(refVoxelCoord(x, v.voxelSize)), refFloorDiv8(refVoxelCoord(y, v.voxelSize)),
                     refFloorDiv8(refVoxelCoord(z, v.voxelSize)));
}
// VoxelUtils.cu:344-359 with Kt rows (fx,0,0),(0,fy,0),(cx,cy,1) (quirks Q1, Q2)
__device__ __forceinline__ bool refBlockInFrustum(const View& v, const float* __restrict__ pose, int bx, int by, int bz) {
    float wx = (float)(bx * 8) * v.voxelSize, wy = (float)(by * 8) * v.voxelSize, wz = (float)(bz * 8) * v.voxelSize;
    float4 p = mul4(pose, wx, wy, wz, 1.0f);
    float rx = v.fx * p.x + 0.0f * p.y + 0.0f * p.z;
    float ry = 0.0f * p.x + v.fy * p.y + 0.0f * p.z;
    float rz = v.cx * p.x + v.cy * p.y + 1.0f * p.z;
    int x = f2i(rx / rz), y = f2i(ry / rz);
    return x < v.W && x >= 0 && y < v.H && y >= 0;
}

// ---- Fixed visibility: bounding sphere of the block against the six frustum planes -------------
__device__ __forceinline__ bool fixedBlockVisible(const View& v, const float* __restrict__ inv, int bx, int by, int bz) {
    float cx = ((float)(bx * 8) + 3.5f) * v.voxelSize;
    float cy = ((float)(by * 8) + 3.5f) * v.voxelSize;
    float cz = ((float)(bz * 8) + 3.5f) * v.voxelSize;
    float4 p = mul4(inv, cx, cy, cz, 1.0f);
    const float r = v.rad;
    if (!(p.z + r > v.depthMin)) return false;
    if (!(p.z - r < v.depthMax)) return false;
    if (!(v.fx * p.x + v.cx * p.z > -(r * v.nl))) return false;
    if (!(v.wr * p.z - v.fx * p.x > -(r * v.nr))) return false;
    if (!(v.fy * p.y + v.cy * p.z > -(r * v.nt))) return false;
    if (!(v.hb * p.z - v.fy * p.y > -(r * v.nb))) return false;
    return true;
}

// ---- 128-bit slot primitives -------------------------------------------------------------------
__device__ __forceinline__ int4 ldSlot(const int4* p) {          // one LDG.128, bypassing L1 (slots are mutated by peers)
    int4 r;
    asm volatile("ld.global.cg.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
// atom.cas.b128 (PTX ISA 8.3, sm_90+): claim-and-publish-the-key in one atomic.
__device__ __forceinline__ bool casSlot(int4* addr, int4 expect, int4 desired, int4& old) {
    unsigned long long elo = ((unsigned long long)(unsigned)expect.y << 32) | (unsigned)expect.x;
    unsigned long long ehi = ((unsigned long long)(unsigned)expect.w << 32) | (unsigned)expect.z;
    unsigned long long dlo = ((unsigned long long)(unsigned)desired.y << 32) | (unsigned)desired.x;
    unsigned long long dhi = ((unsigned long long)(unsigned)desired.w << 32) | (unsigned)desired.z;
    unsigned long long olo, ohi;
    asm volatile(
        "{\n\t.reg .b128 c, s, o;\n\tmov.b128 c, {%2, %3};\n\tmov.b128 s, {%4, %5};\n\t"
        "atom.global.relaxed.gpu.cas.b128 o, [%6], c, s;\n\tmov.b128 {%0, %1}, o;\n\t}"
        : "=l"(olo), "=l"(ohi)
        : "l"(elo), "l"(ehi), "l"(dlo), "l"(dhi), "l"(addr)
        : "memory");
    old.x = (int)(unsigned)olo; old.y = (int)(olo >> 32); old.z = (int)(unsigned)ohi; old.w = (int)(ohi >> 32);
    return olo == elo && ohi == ehi;
}
__device__ __forceinline__ int4 freeSlot() { return make_int4(VH_FREE_COORD, VH_FREE_COORD, VH_FREE_COORD, VH_FREE_BLOCK); }
__device__ __forceinline__ bool sameKey(int4 e, int x, int y, int z) { return e.x == x && e.y == y && e.z == z; }

// read-only lookup shared by integrate-side consumers (raycast): bucket slots, then the chain.
// Mirrors getVoxelEntry4Block (VoxelUtils.cu:362-414). Returns ptr or VH_FREE_BLOCK.
__device__ __forceinline__ int lookupBlock(const View& v, int x, int y, int z) {
    unsigned h = bucketOf(v, x, y, z);
    unsigned base = h * v.bucketSize;
    for (unsigned i = 0; i < v.bucketSize; ++i) {               // most hits are in slot 0 or 1: probe in order
        int4 e = __ldg(v.entries + base + i);
        if (sameKey(e, x, y, z) && e.w != VH_FREE_BLOCK) return e.w;
        if (e.w == VH_FREE_BLOCK && e.x == VH_FREE_COORD) return VH_FREE_BLOCK;   // never-used slot: slots fill in order, so
                                                                                  // nothing lives behind it (nor in the chain)
    }
    unsigned cur = base + v.bucketSize - 1;
    for (unsigned n = 0; n < v.chainMax; ++n) {
        int off = __ldg(v.chain + cur);
        if (off == 0) break;
        cur += (unsigned)off;
        int4 e = __ldg(v.entries + cur);
        if (sameKey(e, x, y, z) && e.w != VH_FREE_BLOCK) return e.w;
    }
    return VH_FREE_BLOCK;
}

// Give a block back (garbage collection, stream-out): the hash slot becomes a tombstone {key, FREE} -- chain links
// stay, lookups and inserts keep walking through it and k_alloc.cu reclaims it --, compaction stops seeing the id,
// and the id goes back on the heap (ref removeSingleBlockInHeap, VoxelUtils.cu:336-341).  One thread per block;
// the caller zeroes the 4 KB.
__device__ __forceinline__ void releaseBlock(const View& v, int id, int4 info) {
    v.entries[info.w] = make_int4(info.x, info.y, info.z, VH_FREE_BLOCK);
    v.blockInfo[id] = make_int4(info.x, info.y, info.z, -1);
    const int addr = atomicAdd(&v.ctr->heapCounter, 1) + 1;
    v.heap[addr] = (unsigned)id;
}

// Packed fp32x2 arithmetic (FFMA2 / FMUL2 / FADD2, sm_100+): one issue slot for two IEEE-rounded lanes, so the
// four voxels of a thread are two register pairs.  Each lane rounds exactly like the scalar instruction, so
// the oracle's scalar fmaf() chain stays the definition.
// (inline PTX: nvcc/ptxas contract a packed multiply feeding a packed add into one FFMA2 even under -fmad=false
// and with explicit .rn, so the policy never relies on a separately rounded packed product.)
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    float2 r;
    asm("{\n\t.reg .b64 a, b, c, d;\n\tmov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %5};\n\tmov.b64 c, {%6, %7};\n\t"
        "fma.rn.f32x2 d, a, b, c;\n\tmov.b64 {%0, %1}, d;\n\t}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return r;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    float2 r;
    asm("{\n\t.reg .b64 a, b, d;\n\tmov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %5};\n\tmul.rn.f32x2 d, a, b;\n\tmov.b64 {%0, %1}, d;\n\t}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    float2 r;
    asm("{\n\t.reg .b64 a, b, d;\n\tmov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %5};\n\tadd.rn.f32x2 d, a, b;\n\tmov.b64 {%0, %1}, d;\n\t}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 dup(float x) { return make_float2(x, x); }

__device__ __forceinline__ float warpSum(float x) {
    x += __shfl_xor_sync(0xffffffffu, x, 16);
    x += __shfl_xor_sync(0xffffffffu, x, 8);
    x += __shfl_xor_sync(0xffffffffu, x, 4);
    x += __shfl_xor_sync(0xffffffffu, x, 2);
    x += __shfl_xor_sync(0xffffffffu, x, 1);
    return x;
}

// ---- optional kernel timeline (build with -DVH_TIMELINE; tools/timeline.py): CTA 0 of every kernel of the frame loop
// appends {kernel id, begin / end, %globaltimer} to a global log, so the interleaving of the streams can be READ
// instead of guessed.
#ifdef VH_TIMELINE
__device__ __forceinline__ void tlStamp(const View& v, int kernel, int phase) {
    if (v.tl != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        const unsigned long long i = atomicAdd(v.tl, 1ull);           // word 0 = count, then the log
        if (i < kTimelineCap) v.tl[1 + i] = (t << 8) | ((unsigned long long)kernel << 1) | (unsigned long long)phase;
    }
}
#define VH_TL(kernel, phase) tlStamp(v, kernel, phase)
#else
#define VH_TL(kernel, phase)
#endif
enum { TL_PREPROCESS = 1, TL_ALIGN = 2, TL_SET_FRAME = 3, TL_ALLOC = 4, TL_COMPACT = 5, TL_INTEGRATE = 6 };

}  // namespace vh

#endif
