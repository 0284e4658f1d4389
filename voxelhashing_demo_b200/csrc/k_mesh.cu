// k_mesh.cu -- triangle mesh of the zero level set of the fused TSDF ("then mesh
// extraction").  The reference has no extraction at all -- its renderer was meant to raycast (shaders/raycastSDF.frag)
// and its only export is the text dump of SDFRenderer.cpp:71-110 (vh_dump_text) -- so the operator is specified here
// and mirrored by the oracle (vo_extract_mesh) for parity:
//
//   * cells: every voxel cube whose MIN corner lies in an allocated block (so each cell belongs to exactly one
//     block; the far corners come from up to seven neighbouring blocks through the hash table); a cell is used only
//     if all eight corners have weight > 0;
//   * marching TETRAHEDRA on the Kuhn subdivision of the cube (six tetrahedra around the diagonal 0-7: the same
//     diagonal in every cell, so faces match and the surface has no cracks; no 256-case table to get wrong);
//   * inside = sdf < 0; an edge vertex is pa + t (pb - pa), t = sa / (sa - sb), endpoints ordered by corner index
//     (the same geometric order in every cell that shares the edge), world metres;
//   * orientation: the triangle normal points to the positive (free-space) side; triangles with two identical
//     vertices (a surface passing exactly through a grid point) are dropped.
// One CTA of 128 threads per block: the 9^3 corner samples are staged in shared memory (invalid = NaN), four cells
// per thread, triangles appended to the caller's buffer with one atomicAdd each (extraction is not a per-frame
// operation).  The counter keeps counting past the capacity so the caller learns the size it needs.
#include "vh_device.cuh"

namespace vh {

__device__ __forceinline__ float3 meshEdge(const float3 pa, const float3 pb, float sa, float sb) {
    const float t = sa / (sa - sb);
    return make_float3(fmaf(t, pb.x - pa.x, pa.x), fmaf(t, pb.y - pa.y, pa.y), fmaf(t, pb.z - pa.z, pa.z));
}

__device__ __forceinline__ void meshEmit(float* __restrict__ tris, int capacity, int* counter, float3 a, float3 b, float3 c, float3 dir) {
    // a surface through a grid point (sdf == 0 at a corner) collapses edges onto that corner: drop the slivers
    auto same = [](float3 p, float3 q) { return p.x == q.x && p.y == q.y && p.z == q.z; };
    if (same(a, b) || same(b, c) || same(a, c)) return;
    const float ux = b.x - a.x, uy = b.y - a.y, uz = b.z - a.z, vx = c.x - a.x, vy = c.y - a.y, vz = c.z - a.z;
    const float nx = uy * vz - uz * vy, ny = uz * vx - ux * vz, nz = ux * vy - uy * vx;
    if (nx * dir.x + ny * dir.y + nz * dir.z < 0.0f) { const float3 t = b; b = c; c = t; }
    const int i = atomicAdd(counter, 1);
    if (i >= capacity) return;
    float* o = tris + (size_t)i * 9;
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = b.x; o[4] = b.y; o[5] = b.z; o[6] = c.x; o[7] = c.y; o[8] = c.z;
}

__global__ void __launch_bounds__(128) k_extract_mesh(View v, float* __restrict__ tris, int capacity, int* counter) {
    __shared__ float sS[9 * 9 * 9];
    __shared__ int sPtr[8];
    const int N = (int)v.numVoxelBlocks;
    const int first = max(min(v.ctr->heapLow, v.ctr->heapCounter) + 1, 0);
    const int tet[6][4] = {{0, 1, 3, 7}, {0, 1, 5, 7}, {0, 2, 3, 7}, {0, 2, 6, 7}, {0, 4, 5, 7}, {0, 4, 6, 7}};
    for (int id = first + blockIdx.x; id < N; id += gridDim.x) {
        const int4 info = __ldg(v.blockInfo + id);
        if (info.w < 0) continue;                                              // CTA-uniform
        __syncthreads();                                                       // previous block's tile is done with
        if (threadIdx.x < 8) {
            const int t = threadIdx.x;
            sPtr[t] = t == 0 ? id * 512 : lookupBlock(v, info.x + (t & 1), info.y + ((t >> 1) & 1), info.z + ((t >> 2) & 1));
        }
        __syncthreads();
        for (int p = threadIdx.x; p < 729; p += 128) {
            const int x = p % 9, y = (p / 9) % 9, z = p / 81;
            const int nb = (x >> 3) | ((y >> 3) << 1) | ((z >> 3) << 2);
            const int ptr = sPtr[nb];
            float s = __int_as_float(0x7fc00000);
            if (ptr >= 0) {
                const float2 sw = __ldg(reinterpret_cast<const float2*>(v.voxels + (size_t)ptr + ((z & 7) * 64 + (y & 7) * 8 + (x & 7))));
                if (sw.y > 0.0f) s = sw.x;
            }
            sS[p] = s;
        }
        __syncthreads();
        for (int cell = threadIdx.x; cell < 512; cell += 128) {
            const int cx = cell & 7, cy = (cell >> 3) & 7, cz = cell >> 6;
            float s[8];
            float3 pos[8];
            bool valid = true, anyNeg = false, anyPos = false;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int dx = c & 1, dy = (c >> 1) & 1, dz = (c >> 2) & 1;
                s[c] = sS[(cz + dz) * 81 + (cy + dy) * 9 + (cx + dx)];
                valid = valid && (s[c] == s[c]);
                anyNeg = anyNeg || s[c] < 0.0f;
                anyPos = anyPos || !(s[c] < 0.0f);
                pos[c] = make_float3((float)(info.x * 8 + cx + dx) * v.voxelSize, (float)(info.y * 8 + cy + dy) * v.voxelSize,
                                     (float)(info.z * 8 + cz + dz) * v.voxelSize);
            }
            if (!valid || !anyNeg || !anyPos) continue;
            for (int t = 0; t < 6; ++t) {
                int in[4], out[4], ni = 0, no = 0;
                for (int k = 0; k < 4; ++k) {
                    const int c = tet[t][k];
                    if (s[c] < 0.0f) in[ni++] = c; else out[no++] = c;
                }
                if (ni == 0 || ni == 4) continue;
                // free-space direction of this tetrahedron: centroid of the outside corners minus centroid of the inside ones
                float3 ci = make_float3(0.f, 0.f, 0.f), co = ci;
                for (int k = 0; k < ni; ++k) { ci.x += pos[in[k]].x; ci.y += pos[in[k]].y; ci.z += pos[in[k]].z; }
                for (int k = 0; k < no; ++k) { co.x += pos[out[k]].x; co.y += pos[out[k]].y; co.z += pos[out[k]].z; }
                const float fi = 1.0f / (float)ni, fo = 1.0f / (float)no;
                const float3 dir = make_float3(co.x * fo - ci.x * fi, co.y * fo - ci.y * fi, co.z * fo - ci.z * fi);
                auto edge = [&](int a, int b) { return a < b ? meshEdge(pos[a], pos[b], s[a], s[b]) : meshEdge(pos[b], pos[a], s[b], s[a]); };
                if (ni == 1) {
                    meshEmit(tris, capacity, counter, edge(in[0], out[0]), edge(in[0], out[1]), edge(in[0], out[2]), dir);
                } else if (ni == 3) {
                    meshEmit(tris, capacity, counter, edge(out[0], in[0]), edge(out[0], in[1]), edge(out[0], in[2]), dir);
                } else {                                                       // 2-2: a quad, split along (in0,out1)-(in1,out0)
                    const float3 a = edge(in[0], out[0]), b = edge(in[0], out[1]), c = edge(in[1], out[1]), d = edge(in[1], out[0]);
                    meshEmit(tris, capacity, counter, a, b, c, dir);
                    meshEmit(tris, capacity, counter, a, c, d, dir);
                }
            }
        }
    }
}

cudaError_t launch_extract_mesh(vh_context* c, float* tris, int capacity, int* d_counter, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(d_counter, 0, sizeof(int), s);
    if (e != cudaSuccess) return e;
    k_extract_mesh<<<c->numSMs * 8, 128, 0, s>>>(c->v, tris, capacity, d_counter);
    return cudaGetLastError();
}

}  // namespace vh
