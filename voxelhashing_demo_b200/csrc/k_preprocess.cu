// k_preprocess.cu -- u16 depth -> vertex map, normal map, dense metric depth.
// Replaces calculateVertexPositions + calculateNormals + preProcess (ref CameraTrackingUtils.cu:50-120).
//
// The reference runs two kernels (the second re-reads five float4 vertices per pixel = 80 B).
// Here one kernel reads the 2-byte depth of the pixel and of its four neighbours (L1-resident),
// recomputes the five back-projections in registers -- the same operations in the same order, so
// the results are bit-identical to the two-pass form -- and writes 36 B per pixel.
#include "preprocess_device.cuh"

namespace vh {

// Optional front end (Fixed): 5x5 bilateral filter of the raw depth.  Spatial weight
// g[|dy|] * g[|dx|], range weight from a table indexed by the depth difference in raw units (both tables are computed
// on the host), invalid (0) pixels and differences >= kBilatLut units contribute nothing; row-major accumulation,
// sum of weights with plain adds, weighted sum with one fma per tap -- the order the oracle mirrors.
__global__ void __launch_bounds__(256) k_bilateral(View v, const uint16_t* __restrict__ depth) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= v.W || y >= v.H) return;
    const size_t idx = (size_t)y * v.W + x;
    const int d0 = (int)__ldg(depth + idx);
    float out = 0.0f;
    if (d0 != 0) {
        float sumW = 0.0f, sumD = 0.0f;
#pragma unroll
        for (int dy = -2; dy <= 2; ++dy) {
            const int yy = y + dy;
            if (yy < 0 || yy >= v.H) continue;
#pragma unroll
            for (int dx = -2; dx <= 2; ++dx) {
                const int xx = x + dx;
                if (xx < 0 || xx >= v.W) continue;
                const int dj = (int)__ldg(depth + (size_t)yy * v.W + xx);
                const int diff = abs(dj - d0);
                if (dj == 0 || diff >= kBilatLut) continue;
                const float ws = v.bilatG[dy < 0 ? -dy : dy] * v.bilatG[dx < 0 ? -dx : dx];
                const float w = ws * __ldg(v.bilatLut + diff);
                sumW = sumW + w;
                sumD = fmaf(w, (float)dj, sumD);
            }
        }
        out = sumD / sumW;                                   // the centre tap (weight g0 * g0 * lut[0] > 0) is always in
    }
    v.depthSmooth[idx] = out;
}

template <class P, bool SMOOTH>
__global__ void __launch_bounds__(256) k_preprocess(View v, const uint16_t* __restrict__ depth, float4* __restrict__ verts,
                                                    float4* __restrict__ normals, float* __restrict__ depthf) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    VH_TL(TL_PREPROCESS, 0);
    if (x >= v.W || y >= v.H) return;
    preprocessPixel<P, SMOOTH>(v, depth, x, y, verts, normals, depthf);
}

cudaError_t launch_preprocess(vh_context* c, const uint16_t* depth, float4* verts, float4* normals, float* depthf,
                              cudaStream_t s) {
    dim3 grid((c->v.W + 31) / 32, (c->v.H + 7) / 8);
    if (c->cfg.policy == VH_POLICY_FIXED && c->v.bilatLut != nullptr) {
        k_bilateral<<<grid, 256, 0, s>>>(c->v, depth);
        k_preprocess<Fixed, true><<<grid, 256, 0, s>>>(c->v, depth, verts, normals, depthf);
    } else if (c->cfg.policy == VH_POLICY_FIXED) k_preprocess<Fixed, false><<<grid, 256, 0, s>>>(c->v, depth, verts, normals, depthf);
    else k_preprocess<RefExact, false><<<grid, 256, 0, s>>>(c->v, depth, verts, normals, depthf);
    return cudaGetLastError();
}

}  // namespace vh
