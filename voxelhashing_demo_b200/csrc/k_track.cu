// k_track.cu -- the tracking of one frame as ONE persistent cooperative kernel: [pre-processing of the new depth frame]
// + a whole CameraTracking::Align (ref CameraTracking.cpp:26-69: maxIters x { FindCorrespondences,
// CalculateJacobiansAndResiduals, BuildLinearSystem, update }) + the camera -> world pose chain
// (ref Application.cpp:73-84: preProcess -> Align -> getTransform).
//
// r1 ran one kernel per Gauss-Newton iteration (k_icp.cu) and its own %globaltimer trace put 5.5 of the 9.7 us per
// iteration into the single-CTA tail: partial row -> fence -> ticket -> last CTA reloads 147 rows -> solve -> next
// launch reloads the delta, with 147 SMs idle.  Here every CTA stays resident for all iterations (r2: 4.9 us each):
//   * the delta lives in SHARED memory of every CTA; nothing per-iteration goes through a kernel boundary, a fence or
//     a ticket;
//   * the 32-float CTA partial is published as 32 x {value, sequence} 8-byte words (relaxed stores, no fence: the
//     sequence number IS the flag).  Two levels: the first CTA of every group of 16 sums its group's rows and
//     publishes a group row, EVERY CTA polls the <= 10 group rows -- the exchange doubles as the grid barrier -- sums
//     them in a fixed order and solves the 6x6 system itself.  All CTAs compute the bit-identical delta, so nothing is
//     broadcast.  (One level, every CTA polling all 148 rows: 5.6 MB of L2 reads per poll round, 1.2-2.1 us.)
//   * rows are double-buffered on the sequence parity (a CTA cannot publish exchange k+2 before every CTA has
//     published k+1, i.e. has finished reading k);
//   * balanced CONTIGUOUS pixel ranges per CTA: source pixels and most target gathers come out of L1 after the first
//     iteration (no fence ever invalidates it), and no CTA carries an extra pixel per thread (everybody waits for the
//     slowest CTA in the exchange);
//   * the in-loop update delta <- exp(x) * delta is fp32 (one warp, a chain of dependent operations while 15 warps
//     wait: every instruction counts); the result is re-orthonormalised ONCE, in fp64, at the end;
//   * multi-GPU (vh_set_peers): CTA 0 scatters the rank's system into every rank's mailbox over NVLink and every
//     CTA of every rank polls its own GPU's mailbox, adds the P contributions in rank order and solves -- still
//     one kernel per Align, one one-way NVLink trip per iteration;
//   * PRE: the prologue turns the raw u16 depth frame into the vertex / normal / metric-depth maps
//     (the arithmetic of k_preprocess, preprocess_device.cuh) -- the maps are still written, because they are the
//     next frame's ICP target and the fusion's input, but the stand-alone pass, its launch and its stream hand-off are
//     gone; one fence + one exchange (used as a grid barrier) separate the prologue from the first iteration.
// The grid must be co-resident (CTAs spin on each other): cooperative launch, one 512-thread CTA per SM.
#include "icp_device.cuh"
#include "preprocess_device.cuh"

namespace vh {

#ifndef VH_ALIGN_ABLATE
#define VH_ALIGN_ABLATE 0             // timing experiments (tools/align_trace.py): 1 = no solve, 2 = no exchange
#endif
#ifndef VH_ALIGN_BATCH
#define VH_ALIGN_BATCH 3
#endif
// pixels in flight per thread.  r2 sweep at VGA (whole Align, 20 iterations): 5 -> 136 us (spills at the 128-register
// cap), 4 -> 124, 3 -> 118, 2 -> 122, 1 -> 122: with the maps in L1 the loop is issue-bound, not latency-bound.
constexpr int kAlignBatch = VH_ALIGN_BATCH;

#ifdef VH_ICP_TRACE
// tools/align_trace.py: %globaltimer stamps [cta][iteration][slot]
constexpr int kTraceIters = 24, kTraceSlots = 8;
__device__ unsigned long long g_alignTrace[kIcpMaxBlocks * kTraceIters * kTraceSlots];
__device__ __forceinline__ void atrace(int it, int slot) {
    if (threadIdx.x == 0 && it < kTraceIters) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_alignTrace[((size_t)blockIdx.x * kTraceIters + it) * kTraceSlots + slot] = t;
    }
}
#define VH_ATRACE(it, slot) atrace(it, slot)
#else
#define VH_ATRACE(it, slot)
#endif

// Sequence-tagged rows of the intra-GPU exchange:
//   level 1: ll[slot][cta][32]                 -- every CTA's partial sums
//   level 2: ll[2 x kIcpMaxBlocks x 32 + ...]  -- one row per group of kGroup CTAs, written by the group's first CTA
constexpr int kGroup = kIcpThreads / 32;                       // 16: the leader reads its group's rows with one warp per row
constexpr int kMaxGroups = (kIcpMaxBlocks + kGroup - 1) / kGroup;
__device__ __forceinline__ unsigned long long* llRow(unsigned long long* ll, unsigned slot, unsigned cta) {
    return ll + ((size_t)slot * kIcpMaxBlocks + cta) * 32u;
}
__device__ __forceinline__ unsigned long long* llGroupRow(unsigned long long* ll, unsigned slot, unsigned group) {
    return ll + (size_t)2 * kIcpMaxBlocks * 32u + ((size_t)slot * kMaxGroups + group) * 32u;
}
// sum of column `lane` of a 16 x 33 table in a fixed balanced order (depth 4)
__device__ __forceinline__ float treeSum16(const float (*t)[33], int lane) {
    float a[16];
#pragma unroll
    for (int g = 0; g < 16; ++g) a[g] = t[g][lane];
#pragma unroll
    for (int w = 8; w >= 1; w >>= 1)
#pragma unroll
        for (int g = 0; g < w; ++g) a[g] = a[g] + a[g + w];
    return a[0];
}
__device__ __forceinline__ float llPoll(const unsigned long long* p, unsigned seq) {
    unsigned long long w = llLoadGpu(p);
    while (!llReady(w, seq)) w = llLoadGpu(p);
    return llValue(w);
}

// All-reduce (sum) of one float per lane of warp 0 over the CTAs of the grid = grid barrier.  Called by every thread of
// every CTA (converged); `mine` matters in warp 0 only.  On return warp 0's lane L holds the grid total of value L; the
// other warps have passed the barrier.  Sums of <= 16 rows in a FIXED balanced order (identical in every CTA ->
// bit-identical delta everywhere), fp32: the rows are fp32 sums of ~2000 pixels each, and a dependent fp64 add costs
// ~35 cycles here -- two chains of 16 were 0.55 us of the critical path of every iteration.
__device__ __forceinline__ float gridExchange(unsigned long long* ll, unsigned seq, float mine, float (*sRows)[33]) {
    constexpr int G = kIcpThreads / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#if VH_ALIGN_ABLATE == 2
    sRows[warp][lane] = warp == 0 ? mine : 0.f;
    __syncthreads();
#else
    const unsigned slot = seq & 1u;
    const unsigned nGroups = (gridDim.x + kGroup - 1) / kGroup;
    if (warp == 0) llStoreGpu(llRow(ll, slot, blockIdx.x) + lane, mine, seq);
    if (blockIdx.x % kGroup == 0) {                          // CTA-uniform: the group's leader, one warp per row of the group
        const unsigned row = blockIdx.x + (unsigned)warp;
        sRows[warp][lane] = row < gridDim.x ? llPoll(llRow(ll, slot, row) + lane, seq) : 0.f;
        __syncthreads();
        if (warp == 0) llStoreGpu(llGroupRow(ll, slot, blockIdx.x / kGroup) + lane, treeSum16(sRows, lane), seq);
        __syncthreads();
    }
    {
        float a = 0.f;
        for (unsigned g = (unsigned)warp; g < nGroups; g += G) a += llPoll(llGroupRow(ll, slot, g) + lane, seq);
        sRows[warp][lane] = a;
    }
    __syncthreads();
#endif
    return warp == 0 ? treeSum16(sRows, lane) : 0.f;
}

// source maps: read-only for the kernel's lifetime (non-coherent path) unless this kernel wrote them itself (PRE)
template <bool PRE>
__device__ __forceinline__ float4 ldMap(const float4* p) {
    if (!PRE) return __ldg(p);
    return *p;                        // plain coherent load (the pointer is neither const-restrict nor provably read-only)
}

struct PreArgs {                      // PRE: the raw frame and the maps the prologue produces (= the Align's source maps)
    const uint16_t* depth;
    float4* verts;
    float4* normals;
    float* depthf;
};

template <class P, bool PRE>
__global__ void __launch_bounds__(kIcpThreads, 1) k_icp_align(View v, IcpState* st, unsigned long long* ll, PreArgs pre,
                                                              const float4* in, const float4* inN,
                                                              const float4* __restrict__ tg, const float4* __restrict__ tgN,
                                                              int row0, int row1, int iterations, PeerView pv,
                                                              const float* poseIn, float* poseOut) {
    __shared__ float sDelta[16];
    __shared__ double sP[16];
    __shared__ float sm[kIcpThreads / 32][32];
    __shared__ float sRows[kIcpThreads / 32][33];
    __shared__ float sSys[32];
    __shared__ int sStop;
    constexpr int B = kAlignBatch;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    IcpDev* dev = devOf(st);

    // every CTA reads the sequence base before any CTA can finish (the base is only rewritten by CTA 0 after the last
    // exchange, which needs every CTA's contribution)
    VH_TL(TL_ALIGN, 0);
    unsigned seq = __ldcg(&v.ctr->icpSeq);
    const unsigned seq0 = seq;
    if (threadIdx.x < 16) sDelta[threadIdx.x] = __ldcg(st->delta + threadIdx.x);
    if (threadIdx.x == 0) sStop = 0;

    if (PRE) {
        VH_ATRACE(23, 0);
        // ---- prologue: u16 depth -> vertex / normal / metric-depth maps over the WHOLE image (every rank of a partitioned
        // run needs all of it: the target of the next frame, the input of its own fusion), balanced contiguous ranges
        const long long all = (long long)v.W * v.H;
        const int plo = (int)(all * blockIdx.x / gridDim.x), phi = (int)(all * (blockIdx.x + 1) / gridDim.x);
        for (int idx = plo + (int)threadIdx.x; idx < phi; idx += kIcpThreads) {
            const int y = idx / v.W, x = idx - y * v.W;
            preprocessPixel<P, false>(v, pre.depth, x, y, pre.verts, pre.normals, pre.depthf);
        }
        // the maps are read by OTHER CTAs below (the row split of the Align need not match the prologue's ranges):
        // stores performed (fence) -> barrier -> fence (acquire side; L1 holds nothing of them yet)
        VH_ATRACE(23, 1);
        __threadfence();
        __syncthreads();
        ++seq;
        gridExchange(ll, seq, 0.f, sRows);
        __threadfence();
        VH_ATRACE(23, 2);
    }
    __syncthreads();

    // Balanced CONTIGUOUS pixel range per CTA (strided chunks gave the first few CTAs a whole extra pixel per thread --
    // 5 instead of 4 at VGA -- and everybody waits for the slowest CTA in the exchange)
    const long long npx = (long long)(row1 - row0) * v.W;
    const int lo = row0 * v.W + (int)(npx * blockIdx.x / gridDim.x);
    const int hi = row0 * v.W + (int)(npx * (blockIdx.x + 1) / gridDim.x);
    const bool haveM = P::fixed && inN != nullptr && v.icpNormalThres > -1.0f;
    int solved = 0;
    for (int it = 0; it < iterations; ++it) {
        // ---- association + residual + Jacobian row + 29 running sums (as k_icp_iter) --------------------------
        VH_ATRACE(it, 0);
        float acc[29];
#pragma unroll
        for (int k = 0; k < 29; ++k) acc[k] = 0.f;
        int i0 = lo + (int)threadIdx.x;
        float4 s[B];
#pragma unroll
        for (int j = 0; j < B; ++j) {
            const int idx = i0 + j * kIcpThreads;
            s[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (idx < hi) s[j] = ldMap<PRE>(in + idx);
        }
        while (true) {
            Cand c[B];
#pragma unroll
            for (int j = 0; j < B; ++j) c[j] = project<P>(v, sDelta, s[j]);
            float4 q[B], n[B], m[B];
#pragma unroll
            for (int j = 0; j < B; ++j) {                   // up to 3 B independent gathers
                q[j] = n[j] = m[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c[j].tidx >= 0) {
                    q[j] = __ldg(tg + c[j].tidx);
                    n[j] = __ldg(tgN + c[j].tidx);
                    if (haveM) m[j] = ldMap<PRE>(inN + i0 + j * kIcpThreads);
                }
            }
            i0 += B * kIcpThreads;
            const bool more = i0 - (int)threadIdx.x < hi;    // CTA-uniform
            if (more) {
#pragma unroll
                for (int j = 0; j < B; ++j) {
                    const int idx = i0 + j * kIcpThreads;
                    s[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (idx < hi) s[j] = ldMap<PRE>(in + idx);
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
        VH_ATRACE(it, 1);
        const float tot = blockReduce29(acc, sm);           // warp 0, lane k < 29: this CTA's sum k
        VH_ATRACE(it, 2);

        // ---- exchange = grid barrier --------------------------------------------------------------------------
        ++seq;
        float f = gridExchange(ll, seq, lane < 29 ? tot : 0.f, sRows);
        VH_ATRACE(it, 3);
        if (warp == 0) {
            if (pv.world > 1) {                              // the cross-GPU collective, still inside the kernel
                if (blockIdx.x == 0) peerScatter(pv, f, seq);
                f = peerGather(pv, seq);
            }
            sSys[lane] = f;
            __syncwarp();
            float dcol[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) dcol[k] = sDelta[k * 4 + (lane & 3)];
            VH_ATRACE(it, 5);
            float tw[6];
#if VH_ALIGN_ABLATE == 1
            bool ok = true;
#pragma unroll
            for (int k = 0; k < 6; ++k) tw[k] = sSys[k] * 1e-30f;
#else
            const bool ok = solveTwistWarp(sSys, P::fixed, tw);
#endif
            if (ok) {
                const float d = updateFp32Warp(expElementWarp(tw), dcol);
                __syncwarp();
                if (lane < 16) sDelta[lane] = d;
            } else if (lane == 0) sStop = 1;
            VH_ATRACE(it, 6);
        }
        __syncthreads();
        VH_ATRACE(it, 4);
        if (sStop) break;                                    // uniform over the grid (and over the ranks): same sums everywhere
        ++solved;
    }

    // Publish the result (every CTA holds the same one).  The fp32 products of the loop are re-orthonormalised ONCE,
    // in fp64 (one Newton-Schulz step), so rounding cannot accumulate over the thousands of Aligns of a long
    // sequence (the delta is the warm start of the next frame); the fp64 copy is what SE3Log / the per-launch solve read.
    if (blockIdx.x == 0 && warp == 0) {
        double pij = lane < 16 ? (double)sDelta[lane] : 0.0;
        if (lane < 16) sP[lane] = pij;
        __syncwarp();
        if (solved > 0) pij = newtonSchulzWarp(sP, pij);
        else if (lane < 16) pij = __ldcg(dev->D + lane);     // nothing solved: the state stays bit for bit as it was
        if (lane < 16) { dev->D[lane] = pij; st->delta[lane] = (float)pij; }
        if (seq != seq0 + (PRE ? 1u : 0u)) st->system[lane] = sSys[lane];
        if (poseOut != nullptr) {
            // camera -> world chain T_k = T_{k-1} * delta (what getTransform() feeds integrate with, Application.cpp:75-84):
            // the same products in the same order as k_set_frame, so both routes give the same bits
            __syncwarp();
            if (lane < 16) sDelta[lane] = (float)pij;
            __syncwarp();
            if (lane < 16) {
                const int r = lane >> 2, c = lane & 3;
                const float t = poseIn[r * 4 + 0] * sDelta[0 * 4 + c] + poseIn[r * 4 + 1] * sDelta[1 * 4 + c] +
                                poseIn[r * 4 + 2] * sDelta[2 * 4 + c] + poseIn[r * 4 + 3] * sDelta[3 * 4 + c];
                __syncwarp(0x0000ffffu);                     // every lane has read the old pose (poseOut may alias poseIn)
                poseOut[lane] = t;
            }
        }
        if (lane == 0) {
            st->iterations += solved;
            v.ctr->icpConverged = sStop;
            v.ctr->icpSeq = seq;
        }
    }
    VH_TL(TL_ALIGN, 1);
}

#ifdef VH_ICP_TRACE
extern "C" int vh_align_trace_read(unsigned long long* host, int n) {
    return (int)cudaMemcpyFromSymbol(host, g_alignTrace, sizeof(unsigned long long) * n);
}
#endif

// d_depth != nullptr: fused pre-processing -- in / inN are then the maps the prologue WRITES (with d_depthf) and reads back
cudaError_t launch_icp_align(vh_context* c, const uint16_t* d_depth, float* d_depthf, const float4* in, const float4* inN,
                             const float4* tg, const float4* tgN, int row0, int row1, int iterations, bool peers,
                             const float* d_poseIn, float* d_poseOut, cudaStream_t s) {
    const bool pre = d_depth != nullptr;
    if (iterations <= 0 && !pre) return cudaSuccess;
    const long long px = pre ? (long long)c->v.W * c->v.H : (long long)(row1 - row0) * c->v.W;
    long long g = (px + kIcpThreads - 1) / kIcpThreads;
    int cap = c->icpCtas > 0 ? c->icpCtas : c->numSMs;
    if (cap > c->numSMs) cap = c->numSMs;                   // one CTA per SM: the grid must be co-resident
    if (cap > kIcpMaxBlocks) cap = kIcpMaxBlocks;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    PeerView pv{};
    pv.world = 1;
    if (peers) pv = c->peers;
    PreArgs pa{d_depth, const_cast<float4*>(in), const_cast<float4*>(inN), d_depthf};
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)g);
    cfg.blockDim = dim3(kIcpThreads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const bool fixed = c->cfg.policy == VH_POLICY_FIXED;
#define VH_LAUNCH_ALIGN(POL, PRE_) \
    cudaLaunchKernelEx(&cfg, k_icp_align<POL, PRE_>, c->v, c->icp, c->icpLL, pa, in, inN, tg, tgN, row0, row1, iterations, pv, d_poseIn, d_poseOut)
    if (pre) return fixed ? VH_LAUNCH_ALIGN(Fixed, true) : VH_LAUNCH_ALIGN(RefExact, true);
    return fixed ? VH_LAUNCH_ALIGN(Fixed, false) : VH_LAUNCH_ALIGN(RefExact, false);
#undef VH_LAUNCH_ALIGN
}

}  // namespace vh
