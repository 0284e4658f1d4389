// k_icp.cu -- point-to-plane ICP (north_star (d)).
// Replaces FindCorrespondences (ref CameraTrackingUtils.cu:131-216), CalculateJacAndResKernel
// (ref Solver.cu:39-71), the cuBLAS Sgemv/Ssyrk pair and the Eigen solve of
// Solver::BuildLinearSystem (ref Solver.cpp:74-111), SE3Exp/SE3Log (ref SE3.cpp:4-19) and the dead
// shared-memory reducer buildLinearSystem (ref LinearSystem.cu:24-101).
//
// One Gauss-Newton iteration = ONE kernel: projective association, signed point-to-plane
// residual, the 6-float Jacobian row, 21 + 6 (+2) partial sums kept in registers, warp-shuffle
// tree, one shared-memory stage per CTA, per-CTA partials to HBM, and the LAST CTA to arrive
// (ticket) sums the partials in CTA order in fp64 (deterministic), solves the 6x6 system
// (warp-cooperative Gauss-Jordan) and left-multiplies the SE(3) update into the delta transform --
// all on the device, so a 20-iteration Align is 20 back-to-back launches with no host round trip.
// The reference moves ~268 B per pixel per iteration through a 7.4 MB Jacobian; this moves the
// algorithmic 48 B (64 B when the source normals are checked).
//
// r1a profile: at VGA everything is L2-resident and the kernel is LATENCY-bound, not bandwidth-bound
// (16 warps/SM, two dependent L2 round trips per pixel, one pixel at a time = 1.7 TB/s).  v2 keeps
// four pixels in flight per thread (4 independent source loads, then up to 12 independent gathers),
// one 512-thread CTA per SM, and a tail that touches each partial row with one 16-byte load.
#include "icp_device.cuh"

namespace vh {


#ifdef VH_ICP_TRACE
__device__ unsigned long long g_icpTrace[1024 * 8];
__device__ __forceinline__ void trace(int slot) {
    if (threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_icpTrace[blockIdx.x * 8 + slot] = t;
    }
}
#define VH_TRACE(slot) trace(slot)
#else
#define VH_TRACE(slot)
#endif

// The solve against the state in global memory (one launch per iteration, k_icp.cu).
__device__ __noinline__ void solveAndUpdateWarp(const float* sys, IcpState* st, IcpDev* dev, Counters* ctr, bool fixedPolicy) {
    __shared__ double sP[16];
    const int lane = threadIdx.x & 31;
    // the fp64 state is only needed at the very end: its loads are in flight during the elimination
    double dcol[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) dcol[k] = __ldcg(dev->D + k * 4 + (lane & 3));
    double pij;
    if (!solveCoreWarp(sys, dcol, fixedPolicy, sP, pij)) {
        if (lane == 0) ctr->icpConverged = 1;
        return;
    }
    if (lane < 16) { dev->D[lane] = pij; st->delta[lane] = (float)pij; }
    if (lane == 0) atomicAdd(&st->iterations, 1);            // RED: nothing waits for it
}

// Cross-GPU all-reduce in the epilogue of the last CTA (icp_device.cuh: peerScatter / peerGather): one warp, the
// exchange's sequence number lives in Counters::icpSeq.
__device__ void peerAllReduce(const PeerView& pv, Counters* ctr, float* sSys) {
    const int lane = threadIdx.x & 31;
    unsigned seq = 0;
    if (lane == 0) seq = ctr->icpSeq + 1;
    seq = __shfl_sync(0xffffffffu, seq, 0);
    peerScatter(pv, sSys[lane], seq);
    sSys[lane] = peerGather(pv, seq);
    if (lane == 0) ctr->icpSeq = seq;
    __syncwarp();
}

// Tail shared by the reductions: the last CTA to arrive sums the per-CTA partials and (optionally) solves.
// Each partial row is 32 floats = 8 x 16 bytes; thread t reads 16-byte column group (t & 7) of rows
// (t >> 3), (t >> 3) + R, ... (R = blockDim/8; <= 3 independent loads for 148 CTAs x 512 threads), sums
// them in fp64, and 32 threads add the R row-group sums in order.  Fixed order => deterministic.
__device__ void reduceTail(const View& v, IcpState* st, float* partials, float tot, vh_icp_system* out, bool solve,
                           bool fixedPolicy, const PeerView* pv = nullptr) {
    __shared__ bool isLast;
    __shared__ float sSys[32];
    __shared__ double sRows[64][33];
    if (threadIdx.x < 32) partials[(size_t)blockIdx.x * 32 + threadIdx.x] = threadIdx.x < 29 ? tot : 0.f;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) isLast = atomicAdd(&v.ctr->icpTicket, 1u) == gridDim.x - 1;
    __syncthreads();
    VH_TRACE(4);
    if (!isLast) return;
    // No second fence: every writer fenced before its ticket, this CTA's ticket came after all of them, and the
    // loads below bypass L1 (ld.cg), so they observe the rows in L2 (the threadFenceReduction pattern).
    {
        const unsigned c4 = threadIdx.x & 7, r0 = threadIdx.x >> 3, R = blockDim.x >> 3;
        double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll 4
        for (unsigned b = r0; b < gridDim.x; b += R) {
            const float4 p = __ldcg(reinterpret_cast<const float4*>(partials + (size_t)b * 32) + c4);
            a0 += (double)p.x; a1 += (double)p.y; a2 += (double)p.z; a3 += (double)p.w;
        }
        // the four row groups of a warp (lanes c4, c4+8, c4+16, c4+24) fold with two shuffle steps, fixed order
#pragma unroll
        for (int off = 8; off <= 16; off <<= 1) {
            a0 += __shfl_xor_sync(0xffffffffu, a0, off); a1 += __shfl_xor_sync(0xffffffffu, a1, off);
            a2 += __shfl_xor_sync(0xffffffffu, a2, off); a3 += __shfl_xor_sync(0xffffffffu, a3, off);
        }
        if ((threadIdx.x & 31) < 8) {
            const unsigned w = threadIdx.x >> 5;
            sRows[w][c4 * 4 + 0] = a0; sRows[w][c4 * 4 + 1] = a1; sRows[w][c4 * 4 + 2] = a2; sRows[w][c4 * 4 + 3] = a3;
        }
    }
    __syncthreads();
    VH_TRACE(5);
    if (threadIdx.x < 32) {
        const unsigned nw = blockDim.x >> 5;
        double t = 0;
        for (unsigned g = 0; g < nw; ++g) t += sRows[g][threadIdx.x];
        float f = (float)t;
        sSys[threadIdx.x] = f;
        if (threadIdx.x == 0) v.ctr->icpTicket = 0;
        __syncwarp();
        if (pv != nullptr && pv->world > 1) {                // the collective, fused into the epilogue
            peerAllReduce(*pv, v.ctr, sSys);
            f = sSys[threadIdx.x];
        }
        st->system[threadIdx.x] = f;
        if (out) reinterpret_cast<float*>(out)[threadIdx.x] = f;
        __syncwarp();
        VH_TRACE(6);
        if (solve) solveAndUpdateWarp(sSys, st, devOf(st), v.ctr, fixedPolicy);
        VH_TRACE(7);
#ifdef VH_ICP_TRACE_TWICE
        // experiment: is the 2.6 us solve an instruction-cache cold miss?  run it again (same code, now hot)
        if (solve) solveAndUpdateWarp(sSys, st, devOf(st), v.ctr, fixedPolicy);
        VH_TRACE(5);
#endif
    }
}

template <class P>
__global__ void __launch_bounds__(kIcpThreads, 1) k_icp_iter(View v, IcpState* st, float* partials, const float4* __restrict__ in,
                                                             const float4* __restrict__ inN, const float4* __restrict__ tg,
                                                             const float4* __restrict__ tgN, int row0, int row1,
                                                             vh_icp_system* out, int solve, int first, PeerView pv) {
    __shared__ float sDelta[16];
    __shared__ float sm[kIcpThreads / 32][32];
    // both loads issue together; the flag only changes in a tail, so the early exit is uniform over the grid
    VH_TRACE(0);
    constexpr int B = 5;                                    // pixels in flight per thread: 148 x 512 x 5 >= 640 x 480 in ONE trip
    const int begin = row0 * v.W, end = row1 * v.W;
    const int T = gridDim.x * blockDim.x;
    int i0 = begin + blockIdx.x * blockDim.x + threadIdx.x;
    // Programmatic dependent launch (iterations 2..n of an Align are launched with the PDL attribute): let the
    // NEXT iteration's CTAs be scheduled as soon as this grid's CTAs retire, and do everything that does not
    // depend on the previous iteration -- the launch itself and the first batch of source loads (the input
    // vertex map is constant during an Align) -- while the previous grid's last CTA is still reducing and
    // solving.  griddepcontrol.wait then blocks until that grid has completed and flushed (delta, flag, ticket).
    asm volatile("griddepcontrol.launch_dependents;");
    float4 s[B];
#pragma unroll
    for (int j = 0; j < B; ++j) {
        const int idx = i0 + j * T;
        s[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx < end) s[j] = __ldg(in + idx);
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // the flag and delta are independent loads: issue them together
    const int conv = first ? 0 : __ldcg(&v.ctr->icpConverged);
    float dl = 0.f;
    if (threadIdx.x < 16) dl = __ldcg(st->delta + threadIdx.x);
    if (conv) return;                                       // uniform over the grid: the flag only changes in a tail
    if (threadIdx.x < 16) sDelta[threadIdx.x] = dl;
    __syncthreads();
    VH_TRACE(1);
    float acc[29];
#pragma unroll
    for (int k = 0; k < 29; ++k) acc[k] = 0.f;
    const bool haveM = P::fixed && inN != nullptr && v.icpNormalThres > -1.0f;
    while (true) {
        Cand c[B];
#pragma unroll
        for (int j = 0; j < B; ++j) c[j] = project<P>(v, sDelta, s[j]);
        float4 q[B], n[B], m[B];
#pragma unroll
        for (int j = 0; j < B; ++j) {                       // up to 3 B independent gathers
            q[j] = n[j] = m[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c[j].tidx >= 0) {
                q[j] = __ldg(tg + c[j].tidx);
                n[j] = __ldg(tgN + c[j].tidx);
                if (haveM) m[j] = __ldg(inN + i0 + j * T);
            }
        }
        i0 += B * T;
        const bool more = i0 < end;                          // larger images: next batch of sources goes in flight now
        if (more) {
#pragma unroll
            for (int j = 0; j < B; ++j) {
                const int idx = i0 + j * T;
                s[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (idx < end) s[j] = __ldg(in + idx);
            }
        }
#pragma unroll
        for (int j = 0; j < B; ++j) {
            if (c[j].tidx < 0) continue;
            Corr r = accept<P>(v, sDelta, c[j], q[j], n[j], m[j], haveM);
            if (r.ok) accumulateCorr<P>(acc, r);
        }
        if (!more) break;
    }
    VH_TRACE(2);
    float tot = blockReduce29(acc, sm);
    VH_TRACE(3);
    if (first && threadIdx.x == 0 && blockIdx.x == 0) v.ctr->icpConverged = 0;
    reduceTail(v, st, partials, tot, out, solve != 0, P::fixed, &pv);
}

#ifdef VH_ICP_TRACE
extern "C" int vh_icp_trace_read(unsigned long long* host, int n) {
    return (int)cudaMemcpyFromSymbol(host, g_icpTrace, sizeof(unsigned long long) * n);
}
#endif

// Normal equations from stored correspondences (the arrays computeCorrespondences leaves behind):
// what Solver::BuildLinearSystem computes with Sgemv/Ssyrk (ref Solver.cpp:80-94).
__global__ void __launch_bounds__(256) k_reduce_corr(View v, IcpState* st, float* partials, const float4* __restrict__ corr,
                                                     const float4* __restrict__ corrN, const float* __restrict__ res,
                                                     vh_icp_system* out) {
    __shared__ float sm[8][32];
    float acc[29];
#pragma unroll
    for (int k = 0; k < 29; ++k) acc[k] = 0.f;
    const int n = v.W * v.H;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 d = __ldg(corr + i), m = __ldg(corrN + i);
        const float r = __ldg(res + i);
        float J[6] = {m.x, m.y, m.z, d.y * m.z - d.z * m.y, d.z * m.x - d.x * m.z, d.x * m.y - d.y * m.x};   // ref Solver.cu:25-34
        accumulate(acc, J, r);
        if (m.x == 0.f && m.y == 0.f && m.z == 0.f) acc[28] -= 1.0f;    // empty rows do not count as correspondences
    }
    float tot = blockReduce29(acc, sm);
    reduceTail(v, st, partials, tot, out, false, false);
}

// Legacy split form of one iteration's first half (ref FindCorrespondences + the three thrust::fill).
template <class P>
__global__ void __launch_bounds__(256) k_find_corr(View v, Pose16f delta, const float4* __restrict__ in, const float4* __restrict__ inN,
                                                   const float4* __restrict__ tg, const float4* __restrict__ tgN,
                                                   float4* __restrict__ corr, float4* __restrict__ corrN, float* __restrict__ res,
                                                   float* err) {
    __shared__ float sDelta[16];
    __shared__ float sErr[8];
    if (threadIdx.x < 16) sDelta[threadIdx.x] = delta.m[threadIdx.x];
    __syncthreads();
    const int n = v.W * v.H;
    float e = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        Corr c = associate<P>(v, sDelta, in, inN, tg, tgN, i);
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f), m = q;
        float r = 0.f;
        if (c.ok) { q = make_float4(c.q.x, c.q.y, c.q.z, c.qw); m = make_float4(c.n.x, c.n.y, c.n.z, c.nw); r = c.d; e += c.d; }   // whole float4s, ref :176-177
        corr[i] = q; corrN[i] = m; res[i] = r;                  // ref :176-178 (+ fills :201-203)
    }
    e = warpSum(e);
    if ((threadIdx.x & 31) == 0) sErr[threadIdx.x >> 5] = e;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sErr[w];
        atomicAdd(err, t);                                      // ref :175, one per CTA instead of one per pixel
    }
}

// ref CalculateJacAndResKernel, Solver.cu:39-51 (also covers the zero fill at :67)
__global__ void k_jacobians(View v, const float4* __restrict__ corr, const float4* __restrict__ corrN, float* __restrict__ J) {
    const int n = v.W * v.H;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 d = __ldg(corr + i), m = __ldg(corrN + i);
        float* o = J + (size_t)i * 6;
        o[0] = m.x; o[1] = m.y; o[2] = m.z;
        o[3] = d.y * m.z - d.z * m.y; o[4] = d.z * m.x - d.x * m.z; o[5] = d.x * m.y - d.y * m.x;
    }
}

// The reducer of the un-built LinearSystem.cu:24-90, same output contract: CTA b covers pixels
// [1024 b, 1024 (b+1)), writes 27 floats: 21 upper-triangle AtA, 6 Atb, with A = (s x n, n),
// b = n.d - n.s, validity n.w != -inf.  (Unknown order (omega, v) here, unlike the live path.)
__global__ void __launch_bounds__(128) k_linear_system_300(int n, const float4* __restrict__ in, const float4* __restrict__ corr,
                                                           const float4* __restrict__ corrN, float* __restrict__ out) {
    __shared__ float sm[4][32];
    float acc[27];
#pragma unroll
    for (int k = 0; k < 27; ++k) acc[k] = 0.f;
    const int base = blockIdx.x * 1024 + threadIdx.x * 8;       // WINDOW_SIZE = 8, ref :28-30
    for (int t = 0; t < 8; ++t) {
        int i = base + t;
        if (i >= n) break;
        const float4 s = __ldg(in + i), d = __ldg(corr + i), m = __ldg(corrN + i);
        if (m.w == __int_as_float(0xff800000)) continue;        // isValid, ref :12-15
        float b = (m.x * d.x + m.y * d.y + m.z * d.z) - (m.x * s.x + m.y * s.y + m.z * s.z);   // ref :18-20
        float A[6] = {s.y * m.z - s.z * m.y, s.z * m.x - s.x * m.z, s.x * m.y - s.y * m.x, m.x, m.y, m.z};   // ref :47-54
        int k = 0;
#pragma unroll
        for (int i2 = 0; i2 < 6; ++i2) {
#pragma unroll
            for (int j = i2; j < 6; ++j) acc[k++] += A[i2] * A[j];
            acc[21 + i2] += A[i2] * b;
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 27; ++k) {
        float s = warpSum(acc[k]);
        if (lane == 0) sm[warp][k] = s;
    }
    __syncthreads();
    if (threadIdx.x < 27) out[blockIdx.x * 27 + threadIdx.x] = sm[0][threadIdx.x] + sm[1][threadIdx.x] + sm[2][threadIdx.x] + sm[3][threadIdx.x];
}

__global__ void k_icp_solve(View v, IcpState* st, const vh_icp_system* sys, int fixedPolicy) {
    __shared__ float s[32];
    if (blockIdx.x != 0 || threadIdx.x >= 32) return;
    if (v.ctr->icpConverged) return;
    s[threadIdx.x] = reinterpret_cast<const float*>(sys)[threadIdx.x];
    st->system[threadIdx.x] = s[threadIdx.x];
    __syncwarp();
    solveAndUpdateWarp(s, st, devOf(st), v.ctr, fixedPolicy != 0);
}

struct Twist6 { float t[6]; };

__global__ void k_icp_reset(View v, IcpState* st, int resetDelta, int useTwist, Twist6 tw) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    v.ctr->icpConverged = 0;
    v.ctr->icpTicket = 0;
    IcpDev* dev = devOf(st);
    if (useTwist) {
        double t[6];
        for (int i = 0; i < 6; ++i) t[i] = tw.t[i];
        se3Exp(t, dev->D);
        for (int i = 0; i < 16; ++i) st->delta[i] = (float)dev->D[i];
        for (int i = 0; i < 6; ++i) st->twist[i] = tw.t[i];
    } else if (resetDelta) {
        for (int i = 0; i < 16; ++i) { dev->D[i] = (i % 5 == 0) ? 1.0 : 0.0; st->delta[i] = (i % 5 == 0) ? 1.f : 0.f; }
        for (int i = 0; i < 6; ++i) st->twist[i] = 0.f;
        st->iterations = 0;
    }
}

__global__ void k_icp_twist(IcpState* st) {                    // Solver::estimate = SE3Log(delta)
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double tw[6];
    se3Log(devOf(st)->D, tw);
    for (int i = 0; i < 6; ++i) st->twist[i] = (float)tw[i];
}

static int icpGrid(const vh_context* c, int pixels, int threads, int perSM) {
    int g = (pixels + threads - 1) / threads;
    int cap = c->numSMs * perSM;
    if (cap > kIcpMaxBlocks) cap = kIcpMaxBlocks;
    if (g > cap) g = cap;
    return g < 1 ? 1 : g;
}

// One CTA per SM (VGA = 4.1 pixels per thread, one trip of <= 5).  Iterations after the first of an Align are
// programmatic dependent launches on numSMs - 1 CTAs, so the whole next grid can become resident next to the
// previous grid's last CTA (which still owns one SM's register file while it reduces and solves).
static cudaError_t launchIcp(vh_context* c, const float4* in, const float4* inN, const float4* tg, const float4* tgN, int row0,
                             int row1, vh_icp_system* d_out, bool solve, bool first, bool chained, const PeerView& pv, cudaStream_t s) {
    // chained: the previous launch on this stream is the previous iteration of the SAME Align (same input maps)
    const bool pdl = chained && !first && c->numSMs > 2;
    int g = icpGrid(c, (row1 - row0) * c->v.W, kIcpThreads, 1);
    if (g >= c->numSMs) g = c->numSMs - 1;                  // same grid for every iteration: same summation order
    if (g < 1) g = 1;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)g);
    cfg.blockDim = dim3(kIcpThreads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    const int solveI = solve ? 1 : 0, firstI = first ? 1 : 0;
    if (c->cfg.policy == VH_POLICY_FIXED)
        return cudaLaunchKernelEx(&cfg, k_icp_iter<Fixed>, c->v, c->icp, c->icpPartials, in, inN, tg, tgN, row0, row1, d_out, solveI, firstI, pv);
    return cudaLaunchKernelEx(&cfg, k_icp_iter<RefExact>, c->v, c->icp, c->icpPartials, in, inN, tg, tgN, row0, row1, d_out, solveI, firstI, pv);
}

cudaError_t launch_icp_iter(vh_context* c, const float4* in, const float4* inN, const float4* tg, const float4* tgN,
                            int row0, int row1, vh_icp_system* d_out, bool solve, cudaStream_t s) {
    return launch_icp_iter_ex(c, in, inN, tg, tgN, row0, row1, d_out, solve, false, false, s);
}

cudaError_t launch_icp_iter_ex(vh_context* c, const float4* in, const float4* inN, const float4* tg, const float4* tgN,
                               int row0, int row1, vh_icp_system* d_out, bool solve, bool first, bool chained, cudaStream_t s) {
    PeerView none{};
    none.world = 1;
    return launchIcp(c, in, inN, tg, tgN, row0, row1, d_out, solve, first, chained, none, s);
}

// One iteration over this rank's image rows with the all-reduce fused into the kernel's epilogue.
cudaError_t launch_icp_iter_peer(vh_context* c, const float4* in, const float4* inN, const float4* tg, const float4* tgN,
                                 int row0, int row1, bool first, cudaStream_t s) {
    return launchIcp(c, in, inN, tg, tgN, row0, row1, nullptr, true, first, !first, c->peers, s);
}

cudaError_t launch_icp_solve(vh_context* c, const vh_icp_system* d_sys, cudaStream_t s) {
    k_icp_solve<<<1, 32, 0, s>>>(c->v, c->icp, d_sys, c->cfg.policy == VH_POLICY_FIXED);
    return cudaGetLastError();
}

cudaError_t launch_icp_reset(vh_context* c, bool resetDelta, cudaStream_t s) {
    Twist6 t{};
    k_icp_reset<<<1, 32, 0, s>>>(c->v, c->icp, resetDelta ? 1 : 0, 0, t);
    return cudaGetLastError();
}

cudaError_t launch_icp_set_twist(vh_context* c, const float* twist6, cudaStream_t s) {
    Twist6 t;
    for (int i = 0; i < 6; ++i) t.t[i] = twist6[i];
    k_icp_reset<<<1, 32, 0, s>>>(c->v, c->icp, 0, 1, t);
    return cudaGetLastError();
}

cudaError_t launch_icp_twist(vh_context* c, cudaStream_t s) {
    k_icp_twist<<<1, 32, 0, s>>>(c->icp);
    return cudaGetLastError();
}

cudaError_t launch_find_corr(vh_context* c, const float4* in, const float4* inN, const float4* tg, const float4* tgN,
                             const float* delta16_host, float4* corr, float4* corrN, float* res, float* d_err, cudaStream_t s) {
    Pose16f d;
    for (int i = 0; i < 16; ++i) d.m[i] = delta16_host[i];
    cudaError_t e = cudaMemsetAsync(d_err, 0, sizeof(float), s);     // ref :193
    if (e != cudaSuccess) return e;
    int g = icpGrid(c, c->v.W * c->v.H, 256, 4);
    if (c->cfg.policy == VH_POLICY_FIXED) k_find_corr<Fixed><<<g, 256, 0, s>>>(c->v, d, in, inN, tg, tgN, corr, corrN, res, d_err);
    else k_find_corr<RefExact><<<g, 256, 0, s>>>(c->v, d, in, inN, tg, tgN, corr, corrN, res, d_err);
    return cudaGetLastError();
}

cudaError_t launch_jacobians(vh_context* c, const float4* corr, const float4* corrN, float* J, cudaStream_t s) {
    k_jacobians<<<icpGrid(c, c->v.W * c->v.H, 256, 4), 256, 0, s>>>(c->v, corr, corrN, J);
    return cudaGetLastError();
}

cudaError_t launch_reduce_corr(vh_context* c, const float4* corr, const float4* corrN, const float* res, vh_icp_system* d_out,
                               cudaStream_t s) {
    k_reduce_corr<<<icpGrid(c, c->v.W * c->v.H, 256, 2), 256, 0, s>>>(c->v, c->icp, c->icpPartials, corr, corrN, res, d_out);
    return cudaGetLastError();
}

cudaError_t launch_linear_system_300(vh_context* c, const float4* in, const float4* corr, const float4* corrN, float* d_out,
                                     cudaStream_t s) {
    int n = c->v.W * c->v.H;
    k_linear_system_300<<<(n + 1023) / 1024, 128, 0, s>>>(n, in, corr, corrN, d_out);
    return cudaGetLastError();
}

}  // namespace vh
