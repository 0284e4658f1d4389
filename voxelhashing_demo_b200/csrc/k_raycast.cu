// k_raycast.cu -- CUDA raycaster over the hash table (north_star (e)).
// Replaces the non-working GLSL pass shaders/raycastSDF.frag:121-177,189-222 (intent in
// notes.md:9-16): it produces what the reference never did -- a vertex map and a normal map in the
// camera frame, with the conventions of preProcess (w = 1 / w = 0, invalid = zeros), so
// CameraTracking::Align can track frame-to-model.  Fixed policy only: the RefExact TSDF is not a
// usable surface (quirks Q1, Q9).
//
// Latency-bound (dependent hash probes + 8-tap gathers, all L2-resident): one thread per pixel,
// a one-entry block cache per thread (consecutive samples hit the same block), block-sized steps
// through unallocated space, TSDF-scaled steps inside the band, linear zero-crossing refinement,
// central-difference gradient for the normal.
#include "vh_device.cuh"

namespace vh {

struct BlockCache { int x, y, z; const Voxel* vox; bool valid; };

// is the block holding voxel (vx, vy, vz) allocated?  (one-entry cache: consecutive samples hit the same block)
__device__ __forceinline__ bool blockAllocated(const View& v, BlockCache& bc, int vx, int vy, int vz) {
    const int bx = vx >> 3, by = vy >> 3, bz = vz >> 3;           // floor division by 8
    if (!(bc.valid && bc.x == bx && bc.y == by && bc.z == bz)) {
        int ptr = lookupBlock(v, bx, by, bz);
        bc.x = bx; bc.y = by; bc.z = bz; bc.valid = true;
        bc.vox = ptr >= 0 ? v.voxels + ptr : nullptr;
    }
    return bc.vox != nullptr;
}

// The 2x2x2 block neighbourhood anchored at the block of the sample's base voxel: a trilinear sample whose base
// voxel sits on a block face needs 2, 4 or 8 blocks, and the samples of one ray stay in the same neighbourhood
// for several steps.  r1 profile: with a one-entry cache the 8 taps of a face-straddling sample alternated
// between two blocks and re-probed the hash table on every tap (dependent L2 round trips): 490 us per VGA frame.
struct Hood { int bx, by, bz; int ptr[8]; unsigned have; bool valid; };

__device__ __forceinline__ int hoodPtr(const View& v, Hood& h, int ci) {
    if (!(h.have & (1u << ci))) {
        h.ptr[ci] = lookupBlock(v, h.bx + (ci & 1), h.by + ((ci >> 1) & 1), h.bz + (ci >> 2));
        h.have |= 1u << ci;
    }
    return h.ptr[ci];
}

__device__ __forceinline__ bool sampleTrilinear(const View& v, Hood& h, float px, float py, float pz, float& out) {
    const float gx = px * v.invVoxelSize, gy = py * v.invVoxelSize, gz = pz * v.invVoxelSize;
    const float fx0 = floorf(gx), fy0 = floorf(gy), fz0 = floorf(gz);
    const int x0 = f2i(fx0), y0 = f2i(fy0), z0 = f2i(fz0);
    const float ax = gx - fx0, ay = gy - fy0, az = gz - fz0;
    const int bx = x0 >> 3, by = y0 >> 3, bz = z0 >> 3;           // floor division by 8
    if (!(h.valid && h.bx == bx && h.by == by && h.bz == bz)) { h.bx = bx; h.by = by; h.bz = bz; h.have = 0; h.valid = true; }
    const int lx = x0 & 7, ly = y0 & 7, lz = z0 & 7;
    const int nx = lx == 7, ny = ly == 7, nz = lz == 7;           // does the +1 tap leave the anchor block?
    // block pointers first (at most 8 probes, usually 1), then the eight voxel loads in flight together
    const Voxel* vp[8];
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int dx = k & 1, dy = (k >> 1) & 1, dz = k >> 2;
        const int ci = (dx & nx) | ((dy & ny) << 1) | ((dz & nz) << 2);
        const int ptr = hoodPtr(v, h, ci);
        ok = ok && ptr >= 0;
        vp[k] = v.voxels + (ptr >= 0 ? ptr : 0) + ((((lz + dz) & 7) * 64) + (((ly + dy) & 7) * 8) + ((lx + dx) & 7));
    }
    if (!ok) return false;
    float2 sw[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) sw[k] = __ldg(reinterpret_cast<const float2*>(vp[k]));
#pragma unroll
    for (int k = 0; k < 8; ++k) ok = ok && sw[k].y > 0.0f;
    if (!ok) return false;
    const float c00 = fmaf(ax, sw[1].x - sw[0].x, sw[0].x), c10 = fmaf(ax, sw[3].x - sw[2].x, sw[2].x);
    const float c01 = fmaf(ax, sw[5].x - sw[4].x, sw[4].x), c11 = fmaf(ax, sw[7].x - sw[6].x, sw[6].x);
    const float c0 = fmaf(ay, c10 - c00, c00), c1 = fmaf(ay, c11 - c01, c01);
    out = fmaf(az, c1 - c0, c0);
    return true;
}

// Ray intervals (what the reference's depthWrite.* pass was meant to produce, notes.md:9-16): every visible
// block splats the depth range of its eight corners into a 1/8-resolution min / max image (positive floats
// order like their bit patterns, so atomicMin / atomicMax on ints), and a ray only marches [min - vs, max + vs]
// of its tile instead of the whole sensor range.
constexpr int kTile = 8;

__global__ void __launch_bounds__(128) k_ray_interval(View v, int* __restrict__ tileMin, int* __restrict__ tileMax) {
    __shared__ float sInv[16];
    if (threadIdx.x < 16) sInv[threadIdx.x] = v.frame->inv[threadIdx.x];
    __syncthreads();
    const int count = v.ctr->compactCount;
    const int tw = (v.W + kTile - 1) / kTile;
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < count; b += gridDim.x * blockDim.x) {
        const int4 e = __ldg(v.compact16 + b);
        float umin = INFINITY, umax = -INFINITY, wmin = INFINITY, wmax = -INFINITY, zmin = INFINITY, zmax = -INFINITY;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float wx = ((float)(e.x * 8 + ((k & 1) ? 8 : 0)) - 0.5f) * v.voxelSize;
            const float wy = ((float)(e.y * 8 + ((k & 2) ? 8 : 0)) - 0.5f) * v.voxelSize;
            const float wz = ((float)(e.z * 8 + ((k & 4) ? 8 : 0)) - 0.5f) * v.voxelSize;
            const float4 p = mul4(sInv, wx, wy, wz, 1.0f);
            const float zc = fmaxf(p.z, 0.05f);
            const float u = p.x / zc * v.fx + v.cx, w = p.y / zc * v.fy + v.cy;
            umin = fminf(umin, u); umax = fmaxf(umax, u);
            wmin = fminf(wmin, w); wmax = fmaxf(wmax, w);
            zmin = fminf(zmin, p.z); zmax = fmaxf(zmax, p.z);
        }
        if (!(zmax > 0.0f)) continue;
        const int x0 = min(max(f2i(floorf(umin)) - 1, 0), v.W - 1) / kTile, x1 = min(max(f2i(floorf(umax)) + 2, 0), v.W - 1) / kTile;
        const int y0 = min(max(f2i(floorf(wmin)) - 1, 0), v.H - 1) / kTile, y1 = min(max(f2i(floorf(wmax)) + 2, 0), v.H - 1) / kTile;
        if (umax < -1.0f || wmax < -1.0f || umin > (float)v.W || wmin > (float)v.H) continue;
        const int lo = __float_as_int(fmaxf(zmin, v.depthMin)), hi = __float_as_int(zmax);
        for (int ty = y0; ty <= y1; ++ty)
            for (int tx = x0; tx <= x1; ++tx) {
                atomicMin(tileMin + ty * tw + tx, lo);
                atomicMax(tileMax + ty * tw + tx, hi);
            }
    }
}

__global__ void __launch_bounds__(128) k_raycast(View v, const int* __restrict__ tileMin, const int* __restrict__ tileMax,
                                                 float4* __restrict__ verts, float4* __restrict__ normals) {
    __shared__ float sPose[16];
    if (threadIdx.x < 16) sPose[threadIdx.x] = v.frame->pose[threadIdx.x];
    __syncthreads();
    // 16x8 pixel tile per CTA: neighbouring rays walk through the same blocks
    const int x = blockIdx.x * 16 + (threadIdx.x & 15);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 4);
    if (x >= v.W || y >= v.H) return;
    const size_t idx = (size_t)y * v.W + x;
    const float vs = v.voxelSize, coarse = 4.0f * v.voxelSize;
    const float3 rd = mul3(v.Kinv, (float)x, (float)y, 1.0f);      // camera ray with z = 1, as preProcess
    const float dwx = sPose[0] * rd.x + sPose[1] * rd.y + sPose[2] * rd.z;
    const float dwy = sPose[4] * rd.x + sPose[5] * rd.y + sPose[6] * rd.z;
    const float dwz = sPose[8] * rd.x + sPose[9] * rd.y + sPose[10] * rd.z;
    const float ox = sPose[3], oy = sPose[7], oz = sPose[11];
    BlockCache bc{0, 0, 0, nullptr, false};
    Hood hood;
    hood.valid = false;
    hood.have = 0;
    const int tile = (y / kTile) * ((v.W + kTile - 1) / kTile) + (x / kTile);
    const float tmin = __int_as_float(__ldg(tileMin + tile)), tmax = __int_as_float(__ldg(tileMax + tile));
    float z = fmaxf(v.depthMin, tmin - vs), zPrev = 0.f, sPrev = 0.f, zHit = 0.f;
    const float zEnd = (tmin <= tmax) ? fminf(v.depthMax, tmax + vs) : 0.0f;     // empty tile: no visible block on this ray
    bool havePrev = false, hit = false;
    for (int it = 0; it < 4096 && z < zEnd; ++it) {
        const float px = fmaf(z, dwx, ox), py = fmaf(z, dwy, oy), pz = fmaf(z, dwz, oz);
        float s;
        if (sampleTrilinear(v, hood, px, py, pz, s)) {
            if (havePrev && sPrev > 0.0f && s <= 0.0f) {
                zHit = zPrev + (z - zPrev) * (sPrev / (sPrev - s));
                hit = true;
                break;
            }
            zPrev = z; sPrev = s; havePrev = true;
            z += fmaxf(vs, 0.8f * s);
        } else {
            havePrev = false;
            const int vx = f2i(floorf(px * v.invVoxelSize + 0.5f)), vy = f2i(floorf(py * v.invVoxelSize + 0.5f)),
                      vz = f2i(floorf(pz * v.invVoxelSize + 0.5f));
            z += blockAllocated(v, bc, vx, vy, vz) ? vs : coarse;
        }
    }
    float4 vo = make_float4(0.f, 0.f, 0.f, 1.0f), no = make_float4(0.f, 0.f, 0.f, 0.f);
    if (hit) {
        const float hx = fmaf(zHit, dwx, ox), hy = fmaf(zHit, dwy, oy), hz = fmaf(zHit, dwz, oz);
        vo = make_float4(rd.x * zHit, rd.y * zHit, rd.z * zHit, 1.0f);
        float gxp, gxm, gyp, gym, gzp, gzm;
        const bool ok = sampleTrilinear(v, hood, hx + vs, hy, hz, gxp) && sampleTrilinear(v, hood, hx - vs, hy, hz, gxm) &&
                        sampleTrilinear(v, hood, hx, hy + vs, hz, gyp) && sampleTrilinear(v, hood, hx, hy - vs, hz, gym) &&
                        sampleTrilinear(v, hood, hx, hy, hz + vs, gzp) && sampleTrilinear(v, hood, hx, hy, hz - vs, gzm);
        if (ok) {
            const float gx = gxp - gxm, gy = gyp - gym, gz = gzp - gzm;
            const float cxn = sPose[0] * gx + sPose[4] * gy + sPose[8] * gz;     // R^T g: world -> camera
            const float cyn = sPose[1] * gx + sPose[5] * gy + sPose[9] * gz;
            const float czn = sPose[2] * gx + sPose[6] * gy + sPose[10] * gz;
            const float l = sqrtf(cxn * cxn + cyn * cyn + czn * czn);
            if (l > 0.0f) no = make_float4(cxn / l, cyn / l, czn / l, 0.0f);
        }
    }
    verts[idx] = vo;
    normals[idx] = no;
}

// Needs the visible list of the CURRENT pose (vh_raycast refreshes it; the frame pipeline has just built it).
cudaError_t launch_raycast(vh_context* c, float4* verts, float4* normals, cudaStream_t s) {
    const int tiles = ((c->v.W + kTile - 1) / kTile) * ((c->v.H + kTile - 1) / kTile);
    cudaError_t e = cudaMemsetAsync(c->tileMin, 0x7f, sizeof(int) * tiles, s);       // 0x7f7f7f7f = 3.4e38
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(c->tileMax, 0, sizeof(int) * tiles, s);
    if (e != cudaSuccess) return e;
    k_ray_interval<<<c->numSMs, 128, 0, s>>>(c->v, c->tileMin, c->tileMax);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    dim3 grid((c->v.W + 15) / 16, (c->v.H + 7) / 8);
    k_raycast<<<grid, 128, 0, s>>>(c->v, c->tileMin, c->tileMax, verts, normals);
    return cudaGetLastError();
}

}  // namespace vh
