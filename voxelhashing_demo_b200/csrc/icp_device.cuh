// icp_device.cuh -- device-side pieces of the point-to-plane ICP shared by k_icp.cu (one launch per Gauss-Newton
// iteration, split / legacy forms) and k_track.cu (one persistent launch per Align).
// Association + residual follow FindCorrespondences (ref CameraTrackingUtils.cu:131-216), the Jacobian row
// CalculateJacAndResKernel (ref Solver.cu:39-71), the solve Solver::BuildLinearSystem (ref Solver.cpp:74-111) and
// SE3Exp / SE3Log (ref SE3.cpp:4-19).
#ifndef VH_ICP_DEVICE_CUH
#define VH_ICP_DEVICE_CUH

#include "vh_device.cuh"

namespace vh {

constexpr int kIcpThreads = 512;

struct Cand { float3 p; int tidx; };             // transformed source point and target pixel (-1: none)
struct Corr { bool ok; float3 q, n, p; float d; float qw, nw; };

// First half of FindCorrespondences: transform by delta, project with K (ref :153-162).
template <class P>
__device__ __forceinline__ Cand project(const View& v, const float* __restrict__ delta, const float4 s) {
    Cand c;
    c.tidx = -1;
    c.p = make_float3(0.f, 0.f, 0.f);
    if (!P::fixed) {
        if (!(s.z != 0)) return c;                                              // ref :153
        float4 p = mul4(delta, s.x, s.y, s.z, 1.0f);                            // ref :154-155
        float3 sp = mul3(v.K, p.x, p.y, p.z);                                   // ref :124
        int ix = d2i((double)(sp.x / sp.z) + 0.5), iy = d2i((double)(sp.y / sp.z) + 0.5);   // ref :128 (Q20)
        c.p = make_float3(p.x, p.y, p.z);
        if (!(ix > 0 && iy > 0 && ix < v.W && iy < v.H)) return c;              // ref :162
        c.tidx = iy * v.W + ix;
        return c;
    }
    // Fixed: every product-sum is a fused chain (DESIGN.md section 4 "ICP"; the oracle mirrors it with fmaf): 9 FFMA
    // for the transform instead of 9 FMUL + 9 FADD -- the loop is instruction-issue bound (r2 ncu: 210 SASS
    // instructions per pixel, 64 % issue utilisation).
    if (!(s.z > 0.0f)) return c;
    const float px = fmaf(delta[0], s.x, fmaf(delta[1], s.y, fmaf(delta[2], s.z, delta[3])));
    const float py = fmaf(delta[4], s.x, fmaf(delta[5], s.y, fmaf(delta[6], s.z, delta[7])));
    const float pz = fmaf(delta[8], s.x, fmaf(delta[9], s.y, fmaf(delta[10], s.z, delta[11])));
    if (!(pz > 1e-6f)) return c;
    // 1 / pz correctly rounded: MUFU.RCP + one Newton step, the in-range path of __frcp_rn without its exponent-window
    // branch (pz is in (1e-6, ~depthMax + |t|) here)
    float rz;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rz) : "f"(pz));
    rz = fmaf(rz, fmaf(-pz, rz, 1.0f), rz);
    const float u = fmaf(px * rz, v.fx, v.cx), w = fmaf(py * rz, v.fy, v.cy);
    // nearest pixel, ties to even, without F2I (quarter-rate pipe): u + 1.5 * 2^23 leaves the rounded value in the low
    // mantissa bits -- identical to __float2int_rn for |u| < 2^22; anything else (inf, NaN included) lands outside
    // [0, W) as an unsigned number and is rejected by the range check (as in k_integrate.cu)
    const unsigned ix = (unsigned)(__float_as_int(u + 12582912.0f) - 0x4B400000);
    const unsigned iy = (unsigned)(__float_as_int(w + 12582912.0f) - 0x4B400000);
    if (ix >= (unsigned)v.W || iy >= (unsigned)v.H) return c;
    c.p = make_float3(px, py, pz);
    c.tidx = (int)(iy * (unsigned)v.W + ix);
    return c;
}

// Second half: residual and acceptance (ref :166-178).  m = source normal (Fixed, optional).
template <class P>
__device__ __forceinline__ Corr accept(const View& v, const float* __restrict__ delta, const Cand& c, const float4 q,
                                       const float4 n, const float4 m, bool haveM) {
    Corr r;
    r.ok = false;
    const float dx = c.p.x - q.x, dy = c.p.y - q.y, dz = c.p.z - q.z;           // ref :168
    float d;
    if (!P::fixed) {
        d = dx * n.x + dy * n.y + dz * n.z;                                     // ref :169
        if (!(d < v.icpDistThres)) return r;                                    // ref :170 (signed, Q21)
    } else {
        if (!(q.z > 0.0f)) return r;
        const float nn = fmaf(n.x, n.x, fmaf(n.y, n.y, n.z * n.z));
        if (!(nn > 0.0f)) return r;
        const float e2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
        const float lim = 3.0f * v.icpDistThres;
        if (!(e2 < lim * lim)) return r;
        d = fmaf(dx, n.x, fmaf(dy, n.y, dz * n.z));
        if (!(fabsf(d) < v.icpDistThres)) return r;
        if (haveM) {
            const float rx = fmaf(delta[0], m.x, fmaf(delta[1], m.y, delta[2] * m.z));
            const float ry = fmaf(delta[4], m.x, fmaf(delta[5], m.y, delta[6] * m.z));
            const float rz = fmaf(delta[8], m.x, fmaf(delta[9], m.y, delta[10] * m.z));
            const float cosang = fmaf(rx, n.x, fmaf(ry, n.y, rz * n.z));
            if (!(cosang > v.icpNormalThres)) return r;
        }
    }
    r.ok = true; r.q = make_float3(q.x, q.y, q.z); r.n = make_float3(n.x, n.y, n.z);
    r.p = c.p; r.d = d; r.qw = q.w; r.nw = n.w;
    return r;
}

template <class P>
__device__ __forceinline__ Corr associate(const View& v, const float* __restrict__ delta, const float4* __restrict__ in,
                                          const float4* __restrict__ inN, const float4* __restrict__ tg,
                                          const float4* __restrict__ tgN, int idx) {
    Cand c = project<P>(v, delta, __ldg(in + idx));
    if (c.tidx < 0) { Corr r; r.ok = false; return r; }
    const bool haveM = P::fixed && inN != nullptr && v.icpNormalThres > -1.0f;
    float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
    if (haveM) m = __ldg(inN + idx);
    return accept<P>(v, delta, c, __ldg(tg + c.tidx), __ldg(tgN + c.tidx), m, haveM);
}

// 29 running sums: 21 upper-triangle JtJ (row by row), 6 Jtr, residual sum, count.
// Unknown order (v, omega) as in the live reference path (Solver.cu:29-34, SE3.cpp:6-9).
// Each term is ONE fused multiply-add (the file is compiled -fmad=false, so a plain `+= a * b` would be FMUL + FADD:
// 54 instructions per pixel in an issue-bound loop).  The sums have no bit-exact counterpart in the reference
// (cuBLAS Sgemv / Ssyrk, Solver.cpp:80-87, whose own accumulation order is unspecified); they are checked against
// fp64 sums to 1e-5, and the single rounding is the more accurate of the two.
__device__ __forceinline__ void accumulate(float* acc, const float* J, float r) {
    int k = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = i; j < 6; ++j) { acc[k] = fmaf(J[i], J[j], acc[k]); ++k; }
#pragma unroll
    for (int i = 0; i < 6; ++i) acc[21 + i] = fmaf(J[i], r, acc[21 + i]);
    acc[27] += r;
    acc[28] += 1.0f;
}

template <class P>
__device__ __forceinline__ void accumulateCorr(float* acc, const Corr& c) {
    const float3 a = P::fixed ? c.p : c.q;                 // ref Solver.cu:26 uses the TARGET point (Q23)
    float J[6] = {c.n.x, c.n.y, c.n.z, a.y * c.n.z - a.z * c.n.y, a.z * c.n.x - a.x * c.n.z, a.x * c.n.y - a.y * c.n.x};
    accumulate(acc, J, c.d);
}

// CTA reduction of 29 sums; result valid in warp 0 lane k (k < 29) as the return value.
// Warp stage: transpose-reduce.  A shuffle tree per value costs 29 x 5 = 145 SHFL per warp and the SHFL
// pipe (one warp-instruction per cycle per SM) was the bottleneck of the block reduce (r1b trace: 1.5 us).
// Here lane pairs exchange the HALF of the values they do not keep: 16 + 8 + 4 + 2 + 1 = 31 SHFL, after
// which lane L holds the warp total of value L.
__device__ __forceinline__ float blockReduce29(float* acc, float (*sm)[32]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float t[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) t[k] = k < 29 ? acc[k] : 0.f;
#pragma unroll
    for (int step = 16; step >= 1; step >>= 1) {
        const bool upper = (lane & step) != 0;
#pragma unroll
        for (int i = 0; i < step; ++i) {
            const float send = upper ? t[i] : t[i + step];
            const float keep = upper ? t[i + step] : t[i];
            t[i] = keep + __shfl_xor_sync(0xffffffffu, send, step);
        }
    }
    sm[warp][lane] = t[0];
    __syncthreads();
    float tot = 0.f;
    if (warp == 0 && lane < 29) {
        const int nw = blockDim.x >> 5;
        for (int w = 0; w < nw; ++w) tot += sm[w][lane];
    }
    return tot;
}

// ---- SE(3) pieces, fp64, closed form (Eigen's matrix exp/log are not available) ----
__device__ __forceinline__ void soTerms(const double* w, double& A, double& B, double& C) {
    double t2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    double th = sqrt(t2);
    if (th < 1e-6) { A = 1.0 - t2 / 6.0; B = 0.5 - t2 / 24.0; C = 1.0 / 6.0 - t2 / 120.0; }
    else { double s, c; sincos(th, &s, &c); A = s / th; B = (1.0 - c) / t2; C = (th - s) / (t2 * th); }
}
// element (i, j) of exp([[w]x v; 0 0]) (ref SE3Exp, twist = (v, omega))
__device__ __forceinline__ double se3ExpElement(const double* tw, double A, double B, double C, int i, int j) {
    const double* vv = tw; const double* w = tw + 3;
    if (i == 3) return j == 3 ? 1.0 : 0.0;
    const double K[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
    double K2row[3];
    for (int c = 0; c < 3; ++c) K2row[c] = K[i * 3] * K[c] + K[i * 3 + 1] * K[3 + c] + K[i * 3 + 2] * K[6 + c];
    if (j < 3) return ((i == j) ? 1.0 : 0.0) + A * K[i * 3 + j] + B * K2row[j];
    double t = 0;
    for (int c = 0; c < 3; ++c) t += (((i == c) ? 1.0 : 0.0) + B * K[i * 3 + c] + C * K2row[c]) * vv[c];
    return t;
}
static __device__ void se3Exp(const double* tw, double* M) {
    double A, B, C;
    soTerms(tw + 3, A, B, C);
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) M[i * 4 + j] = se3ExpElement(tw, A, B, C, i, j);
}
static __device__ void se3Log(const double* M, double* tw) {          // ref SE3Log
    double tr = M[0] + M[5] + M[10];
    double cs = fmin(1.0, fmax(-1.0, (tr - 1.0) * 0.5));
    double th = acos(cs);
    double f = (th < 1e-6) ? 0.5 + th * th / 12.0 : th / (2.0 * sin(th));
    double w[3] = {f * (M[9] - M[6]), f * (M[2] - M[8]), f * (M[4] - M[1])};
    double A, B, C;
    soTerms(w, A, B, C);
    double t2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    double D = (t2 < 1e-12) ? (1.0 / 12.0 + t2 / 720.0) : (1.0 - A / (2.0 * B)) / t2;
    double K[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
    double K2[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) K2[i * 3 + j] = K[i * 3] * K[j] + K[i * 3 + 1] * K[3 + j] + K[i * 3 + 2] * K[6 + j];
    double t[3] = {M[3], M[7], M[11]};
    for (int i = 0; i < 3; ++i) {
        double s = 0;
        for (int j = 0; j < 3; ++j) s += (((i == j) ? 1.0 : 0.0) - 0.5 * K[i * 3 + j] + D * K2[i * 3 + j]) * t[j];
        tw[i] = s;
    }
    tw[3] = w[0]; tw[4] = w[1]; tw[5] = w[2];
}

struct IcpDev {               // device-resident solver state (Solver::estimate / deltaTransform)
    double D[16];             // delta in fp64
};
__device__ __forceinline__ IcpDev* devOf(IcpState* st) { return reinterpret_cast<IcpDev*>(st + 1); }

// update = -(JtJ)^-1 Jtr; delta <- exp(update) * delta  (== exp(log(exp(update) exp(estimate))), Solver.cpp:110-111).
// Executed by ONE FULL WARP (converged).  r1b trace: the fp64 form of this step cost 3.4 us of a 13.6 us
// iteration (software fp64 division / sincos on one dependent chain), so:
//   * every lane runs Gauss-Jordan on the whole [JtJ | -Jtr] with the diagonal pivot in fp32 (JtJ is SPD when the
//     scene constrains all six degrees of freedom -- elimination without row exchanges is then as stable as
//     Cholesky; otherwise the result is not finite and the iteration stops).  Gauss-Newton is self-correcting: an
//     fp32 error in one update is removed by the next;
//   * lanes 0..15 own one element of exp(update) (fp32, series below 0.05 rad) and of the 4x4 product, which
//     is accumulated in fp64 into the fp64 state D;
//   * one Newton-Schulz step R <- 1.5 R - 0.5 R R^T R (fp64) re-orthonormalises the rotation, so the fp32
//     rounding of exp(update) cannot accumulate over the thousands of updates of a long sequence.
// The pieces, each executed by one converged warp with every lane computing the same values:
//   solveTwistWarp   [JtJ | -Jtr] -> twist (v, omega); false when the iteration must stop (too few correspondences /
//                    exactly-zero error, CameraTracking.cpp:55-58; singular system)
//   expElementWarp   element (lane >> 2 & 3, lane & 3) of exp(twist^), fp32
//   updateFp64Warp   D <- exp * D in fp64 + one Newton-Schulz step (per-launch path; dcol[k] = D[k][lane & 3])
//   updateFp32Warp   delta <- exp * delta in fp32 (persistent Align: the delta is re-orthonormalised once, at the end)
__device__ __forceinline__ bool solveTwistWarp(const float* sys, bool fixedPolicy, float (&tw)[6]) {
    if (fixedPolicy ? !(sys[28] >= 6.0f) : (sys[27] == 0.0f)) return false;     // CameraTracking.cpp:55-58
    // Symmetric Gaussian elimination (LDL^T, right-looking) on the LOWER triangle of [JtJ | -Jtr], fp32, the whole system
    // in the registers of every lane (uniform control flow: no shuffles, no divergence; r1c probe: a row-per-lane
    // shuffle form spent 3400 cycles here).  JtJ is SPD when the scene constrains all six degrees of freedom, so no
    // row exchanges are needed (as stable as Cholesky); otherwise a pivot is ~0, the result is not finite and the
    // iteration stops.  Half the multiply-adds of the Gauss-Jordan form r1 used (r2 ncu: the single warp running the
    // solve is a chain of dependent fp32 operations at ~0.25 IPC while 15 warps wait at the barrier; the critical path
    // is six dependent reciprocals, so fewer instructions = shorter wait).
    float A[6][6], bb[6], inv[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
#pragma unroll
        for (int j = i; j < 6; ++j) A[j][i] = sys[i * 6 - (i * (i - 1)) / 2 + (j - i)];   // selfadjointView, Solver.cpp:92
        bb[i] = -sys[21 + i];                                                  // update = -(JTJinv * JTr), :110
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        inv[k] = __frcp_rn(A[k][k]);                                            // == 1.0f / pivot, correctly rounded
#pragma unroll
        for (int i = k + 1; i < 6; ++i) {
            const float l = -(A[i][k] * inv[k]);
#pragma unroll
            for (int j = k + 1; j <= i; ++j) A[i][j] = fmaf(l, A[j][k], A[i][j]);   // one rounding per update (the file is -fmad=false)
            bb[i] = fmaf(l, bb[k], bb[i]);
        }
    }
    tw[5] = bb[5] * inv[5];
#pragma unroll
    for (int i = 4; i >= 0; --i) {
        float t = bb[i];
#pragma unroll
        for (int j = i + 1; j < 6; ++j) t = fmaf(-A[j][i], tw[j], t);
        tw[i] = t * inv[i];
    }
    bool ok = true;
#pragma unroll
    for (int c = 0; c < 6; ++c) ok = ok && isfinite(tw[c]);
    return ok;                                                                  // uniform: every lane holds the same values
}

// exp([[w]x v; 0 0]) element (i, j) = (lane >> 2 & 3, lane & 3), fp32 (ref SE3Exp, twist = (v, omega))
__device__ __forceinline__ float expElementWarp(const float (&tw)[6]) {
    const int lane = threadIdx.x & 31;
    const float t2 = tw[3] * tw[3] + tw[4] * tw[4] + tw[5] * tw[5];
    float A, B, C;
    if (t2 < 2.5e-3f) {
        A = 1.0f - t2 * (1.0f / 6.0f) + t2 * t2 * (1.0f / 120.0f);
        B = 0.5f - t2 * (1.0f / 24.0f) + t2 * t2 * (1.0f / 720.0f);
        C = (1.0f / 6.0f) - t2 * (1.0f / 120.0f) + t2 * t2 * (1.0f / 5040.0f);
    } else {
        const float th = sqrtf(t2);
        float sn, cs;
        sincosf(th, &sn, &cs);
        A = sn / th; B = (1.0f - cs) / t2; C = (th - sn) / (t2 * th);
    }
    // Every lane builds the whole top 3x4 block (statically indexed: registers, no local memory) and keeps its element.
    // K = [w]x, K^2 = w w^T - |w|^2 I;  R = I + A K + B K^2;  t = (I + B K + C K^2) v
    const float wx = tw[3], wy = tw[4], wz = tw[5];
    const float K[3][3] = {{0.f, -wz, wy}, {wz, 0.f, -wx}, {-wy, wx, 0.f}};
    const float w[3] = {wx, wy, wz};
    float E[3][4];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        float t = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float I = r == c ? 1.0f : 0.0f;
            const float k2 = r == c ? w[r] * w[c] - t2 : w[r] * w[c];
            E[r][c] = I + A * K[r][c] + B * k2;
            t += (I + B * K[r][c] + C * k2) * tw[c];
        }
        E[r][3] = t;
    }
    const int i = (lane >> 2) & 3, j = lane & 3;
    float e = j == 3 ? 1.0f : 0.0f;                          // row 3
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (i == r && j == c) e = E[r][c];
    return e;
}

// One Newton-Schulz step R <- 1.5 R - 0.5 R R^T R on the rotation block of the 4x4 in sP (fp64, 16 doubles of shared
// memory private to the warp, already written by lanes < 16); returns the lane's element.
__device__ __forceinline__ double newtonSchulzWarp(const double* sP, double pij) {
    const int lane = threadIdx.x & 31;
    const int i = (lane >> 2) & 3, j = lane & 3;
    if (lane < 16 && i < 3 && j < 3) {
        double rm = 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double mkj = sP[0 * 4 + k] * sP[0 * 4 + j] + sP[1 * 4 + k] * sP[1 * 4 + j] + sP[2 * 4 + k] * sP[2 * 4 + j];
            rm += sP[i * 4 + k] * mkj;
        }
        pij = 1.5 * pij - 0.5 * rm;
    }
    return pij;
}

__device__ __forceinline__ double updateFp64Warp(float uij, const double (&dcol)[4], double* sP) {
    const int lane = threadIdx.x & 31;
    const int i = (lane >> 2) & 3;
    double pij = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float uik = __shfl_sync(0xffffffffu, uij, i * 4 + k);
        pij += (double)uik * dcol[k];
    }
    if (lane < 16) sP[lane] = pij;
    __syncwarp();
    pij = newtonSchulzWarp(sP, pij);
    __syncwarp();
    return pij;
}

// delta <- exp * delta, fp32, two independent fused chains; dcol[k] = delta[k][lane & 3]
__device__ __forceinline__ float updateFp32Warp(float uij, const float (&dcol)[4]) {
    const int lane = threadIdx.x & 31;
    const int i = (lane >> 2) & 3;
    const float u0 = __shfl_sync(0xffffffffu, uij, i * 4 + 0), u1 = __shfl_sync(0xffffffffu, uij, i * 4 + 1);
    const float u2 = __shfl_sync(0xffffffffu, uij, i * 4 + 2), u3 = __shfl_sync(0xffffffffu, uij, i * 4 + 3);
    return fmaf(u0, dcol[0], u1 * dcol[1]) + fmaf(u2, dcol[2], u3 * dcol[3]);
}

// Per-launch path: dcol[k] = D[k][lane & 3] (the current fp64 delta); lane L < 16 gets element L of the new one.
__device__ __forceinline__ bool solveCoreWarp(const float* sys, const double (&dcol)[4], bool fixedPolicy, double* sP, double& pij) {
    float tw[6];
    if (!solveTwistWarp(sys, fixedPolicy, tw)) return false;
    pij = updateFp64Warp(expElementWarp(tw), dcol, sP);
    return true;
}

// ---- low-latency exchange words: a float travels WITH the sequence number of its exchange in ONE 8-byte scalar
// access (single-copy atomic under the PTX memory model), so there is no data / flag pair and no fence: the receiver
// polls the word until the sequence matches.  .gpu scope between the CTAs of one grid, .sys scope across NVLink.
__device__ __forceinline__ unsigned long long llPack(float v, unsigned seq) {
    return ((unsigned long long)seq << 32) | (unsigned long long)__float_as_uint(v);
}
__device__ __forceinline__ void llStoreGpu(unsigned long long* p, float v, unsigned seq) {
    asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(p), "l"(llPack(v, seq)) : "memory");
}
__device__ __forceinline__ unsigned long long llLoadGpu(const unsigned long long* p) {
    unsigned long long w;
    asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    return w;
}
__device__ __forceinline__ void llStoreSys(unsigned long long* p, float v, unsigned seq) {
    asm volatile("st.relaxed.sys.global.b64 [%0], %1;" ::"l"(p), "l"(llPack(v, seq)) : "memory");
}
__device__ __forceinline__ unsigned long long llLoadSys(const unsigned long long* p) {
    unsigned long long w;
    asm volatile("ld.relaxed.sys.global.b64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    return w;
}
__device__ __forceinline__ bool llReady(unsigned long long w, unsigned seq) { return (unsigned)(w >> 32) == seq; }
__device__ __forceinline__ float llValue(unsigned long long w) { return __uint_as_float((unsigned)w); }

#ifndef VH_PEER_BACKOFF
#define VH_PEER_BACKOFF 0
#endif
// ---- fused cross-GPU all-reduce (one process per GPU, peer memory over NVLink / NVSwitch) --------------
// Executed by one warp: lane L holds value L of this rank's 32-float system.  Scatter it into every rank's
// exchange region (P2P stores over NVLink), wait for all ranks' contributions, add them in RANK ORDER (so every
// rank holds the bit-identical system and solves the identical pose: no second broadcast).
// Regions are double-buffered on the sequence parity: a rank can run at most one exchange ahead of the slowest
// rank, because it cannot finish exchange k+1 without that rank's contribution to k+1.
__device__ __forceinline__ unsigned long long* peerWord(const PeerView& pv, int region, unsigned slot, int rank, int lane) {
    return reinterpret_cast<unsigned long long*>(pv.buf[region]) + (slot * kMaxPeers + (unsigned)rank) * 32u + (unsigned)lane;
}
__device__ __forceinline__ void peerScatter(const PeerView& pv, float mine, unsigned seq) {
    const int lane = threadIdx.x & 31;
    for (int p = 0; p < pv.world; ++p) llStoreSys(peerWord(pv, p, seq & 1u, pv.rank, lane), mine, seq);
}
// Bounded: a peer that died or never launched must not hang this GPU inside a kernel.  After kPeerSpinCycles the lane gives
// up, counts the time-out and returns NaN -- the solve then refuses (non-finite), every CTA stops the Align in the same
// iteration, and the host finds vh_stats::exchangeTimeouts > 0.
__device__ __forceinline__ float peerGather(const PeerView& pv, unsigned seq) {
    const int lane = threadIdx.x & 31;
    unsigned long long w[kMaxPeers];
#pragma unroll
    for (int r = 0; r < kMaxPeers; ++r) w[r] = r < pv.world ? llLoadSys(peerWord(pv, pv.rank, seq & 1u, r, lane)) : 0ull;
    float t = 0.f;
    long long t0 = 0;
    bool timing = false;
#pragma unroll
    for (int r = 0; r < kMaxPeers; ++r) {
        if (r >= pv.world) break;
        while (!llReady(w[r], seq)) {
#if VH_PEER_BACKOFF > 0
            __nanosleep(VH_PEER_BACKOFF);                    // the mailbox lines are the target of the peers' NVLink stores
#endif
            if (!timing) { t0 = clock64(); timing = true; }
            else if (clock64() - t0 > kPeerSpinCycles) {
                if (lane == 0 && pv.timeouts) atomicAdd(pv.timeouts, 1);
                return __int_as_float(0x7fc00000);
            }
            w[r] = llLoadSys(peerWord(pv, pv.rank, seq & 1u, r, lane));
        }
        t += llValue(w[r]);
    }
    return t;
}

}  // namespace vh

#endif
