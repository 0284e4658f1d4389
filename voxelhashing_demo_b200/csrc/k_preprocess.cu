// k_preprocess.cu -- u16 depth -> vertex map, normal map, dense metric depth.
// Replaces calculateVertexPositions + calculateNormals + preProcess (ref CameraTrackingUtils.cu:50-120).
//
// The reference runs two kernels (the second re-reads five float4 vertices per pixel = 80 B).
// Here one kernel reads the 2-byte depth of the pixel and of its four neighbours (L1-resident),
// recomputes the five back-projections in registers -- the same operations in the same order, so
// the results are bit-identical to the two-pass form -- and writes 36 B per pixel.
#include "vh_device.cuh"

namespace vh {

template <class P>
__device__ __forceinline__ float3 backproject(const View& v, const uint16_t* __restrict__ depth, int x, int y) {
    float d = (float)__ldg(depth + (size_t)y * v.W + x) / v.depthScale;       // ref :63-64
    if (P::fixed && !(d > v.depthMin && d < v.depthMax)) d = 0.0f;
    float3 k = mul3(v.Kinv, (float)x, (float)y, 1.0f);                        // ref :69-70
    return make_float3(k.x * d, k.y * d, k.z * d);
}

template <class P>
__global__ void __launch_bounds__(256) k_preprocess(View v, const uint16_t* __restrict__ depth, float4* __restrict__ verts,
                                                    float4* __restrict__ normals, float* __restrict__ depthf) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= v.W || y >= v.H) return;
    const size_t idx = (size_t)y * v.W + x;
    const float3 CC = backproject<P>(v, depth, x, y);
    verts[idx] = make_float4(CC.x, CC.y, CC.z, 1.0f);                         // ref :72-73, w = 1 always (Q27)
    if (depthf) depthf[idx] = CC.z;
    float4 n = make_float4(0.f, 0.f, 0.f, 0.f);                               // ref :91
    if (x > 0 && x < v.W - 1 && y > 0 && y < v.H - 1) {                       // ref :93
        const float3 PC = backproject<P>(v, depth, x, y + 1);
        const float3 CP = backproject<P>(v, depth, x + 1, y);
        const float3 MC = backproject<P>(v, depth, x, y - 1);
        const float3 CM = backproject<P>(v, depth, x - 1, y);
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
    if (c->cfg.policy == VH_POLICY_FIXED) k_preprocess<Fixed><<<grid, 256, 0, s>>>(c->v, depth, verts, normals, depthf);
    else k_preprocess<RefExact><<<grid, 256, 0, s>>>(c->v, depth, verts, normals, depthf);
    return cudaGetLastError();
}

}  // namespace vh
