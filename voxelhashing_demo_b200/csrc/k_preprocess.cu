// k_preprocess.cu -- u16 depth -> vertex map, normal map, dense metric depth.
// Replaces calculateVertexPositions + calculateNormals + preProcess (ref CameraTrackingUtils.cu:50-120).
//
// The reference runs two kernels (the second re-reads five float4 vertices per pixel = 80 B).
// Here one kernel reads the 2-byte depth of the pixel and of its four neighbours (L1-resident),
// recomputes the five back-projections in registers -- the same operations in the same order, so
// the results are bit-identical to the two-pass form -- and writes 36 B per pixel.
#include "vh_device.cuh"

namespace vh {

// Depth of pixel (x, y) in metres.  SMOOTH (Fixed, optional): from the bilateral-filtered image.
template <class P, bool SMOOTH>
__device__ __forceinline__ float metricDepth(const View& v, const uint16_t* __restrict__ depth, int x, int y) {
    const size_t i = (size_t)y * v.W + x;
    float d = (SMOOTH ? __ldg(v.depthSmooth + i) : (float)__ldg(depth + i)) / v.depthScale;     // ref :63-64
    if (P::fixed && !(d > v.depthMin && d < v.depthMax)) d = 0.0f;
    return d;
}
template <class P, bool SMOOTH>
__device__ __forceinline__ float3 backproject(const View& v, const uint16_t* __restrict__ depth, int x, int y) {
    const float d = metricDepth<P, SMOOTH>(v, depth, x, y);
    float3 k = mul3(v.Kinv, (float)x, (float)y, 1.0f);                        // ref :69-70
    return make_float3(k.x * d, k.y * d, k.z * d);
}

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
    if (x >= v.W || y >= v.H) return;
    const size_t idx = (size_t)y * v.W + x;
    const float3 CC = backproject<P, SMOOTH>(v, depth, x, y);
    verts[idx] = make_float4(CC.x, CC.y, CC.z, 1.0f);                         // ref :72-73, w = 1 always (Q27)
    if (depthf) {                                                             // integration reads the RAW depth
        float3 kz = mul3(v.Kinv, (float)x, (float)y, 1.0f);
        depthf[idx] = SMOOTH ? kz.z * metricDepth<P, false>(v, depth, x, y) : CC.z;
    }
    float4 n = make_float4(0.f, 0.f, 0.f, 0.f);                               // ref :91
    if (x > 0 && x < v.W - 1 && y > 0 && y < v.H - 1) {                       // ref :93
        const float3 PC = backproject<P, SMOOTH>(v, depth, x, y + 1);
        const float3 CP = backproject<P, SMOOTH>(v, depth, x + 1, y);
        const float3 MC = backproject<P, SMOOTH>(v, depth, x, y - 1);
        const float3 CM = backproject<P, SMOOTH>(v, depth, x - 1, y);
        bool ok;
        if (!P::fixed) {
            ok = CC.x != 0 && PC.x != 0 && CP.x != 0 && MC.x != 0 && CM.x != 0;   // ref :100 (tests .x)
        } else {
            ok = CC.z != 0 && PC.z != 0 && CP.z != 0 && MC.z != 0 && CM.z != 0;
            if (ok) {
                float lim = 0.05f * CC.z;
                ok = fabsf(PC.z - CC.z) < lim && fabsf(MC.z - CC.z) < lim && fabsf(CP.z - CC.z) < lim && fabsf(CM.z - CC.z) < lim;
            }
        }
        if (ok) {
            float ax = PC.x - MC.x, ay = PC.y - MC.y, az = PC.z - MC.z;       // ref :102
            float bx = CP.x - CM.x, by = CP.y - CM.y, bz = CP.z - CM.z;
            float nx = ay * bz - az * by, ny = az * bx - ax * bz, nz = ax * by - ay * bx;   // helper_math.h:1420
            float l = sqrtf(nx * nx + ny * ny + nz * nz);                     // helper_math.h:1291
            if (l > 0.0f) n = make_float4(nx / l, ny / l, nz / l, 0.0f);      // ref :105-109
        }
    }
    normals[idx] = n;
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
