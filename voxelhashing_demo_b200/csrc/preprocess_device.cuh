// preprocess_device.cuh -- per-pixel arithmetic of the pre-processing (u16 depth -> vertex, normal, metric depth), shared
// by k_preprocess.cu (stand-alone pass) and k_track.cu (fused into the prologue of the persistent Align kernel).
// Follows calculateVertexPositions + calculateNormals (ref CameraTrackingUtils.cu:50-113).
#ifndef VH_PREPROCESS_DEVICE_CUH
#define VH_PREPROCESS_DEVICE_CUH

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

// One pixel: the five back-projections of the 5-point stencil are recomputed from five u16 depth reads (L1-resident) --
// the same operations in the same order as the reference's two passes, so the results are bit-identical.
template <class P, bool SMOOTH>
__device__ __forceinline__ void preprocessPixel(const View& v, const uint16_t* __restrict__ depth, int x, int y,
                                                float4* __restrict__ verts, float4* __restrict__ normals, float* __restrict__ depthf) {
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

}  // namespace vh

#endif
