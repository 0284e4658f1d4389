// k_track.cu -- a whole CameraTracking::Align (ref CameraTracking.cpp:26-69: maxIters x { FindCorrespondences,
// CalculateJacobiansAndResiduals, BuildLinearSystem, update }) as ONE persistent cooperative kernel.
//
// r1 ran one kernel per Gauss-Newton iteration (k_icp.cu) and its own %globaltimer trace put 5.5 of the 9.7 us per
// iteration into the single-CTA tail: partial row -> fence -> ticket -> last CTA reloads 147 rows -> solve -> next
// launch reloads the delta, with 147 SMs idle.  Here every CTA stays resident for all iterations:
//   * the delta and its fp64 master copy live in SHARED memory of every CTA; nothing per-iteration goes through a
//     kernel boundary, a fence or a ticket;
//   * the 32-float CTA partial is published as 32 x {value, sequence} 8-byte words (relaxed stores, no fence: the
//     sequence number IS the flag) and EVERY CTA polls all rows -- the exchange doubles as the grid barrier, one L2
//     round trip -- sums them in fp64 in CTA order and solves the 6x6 system itself.  All CTAs compute the
//     bit-identical delta, so nothing is broadcast;
//   * rows are double-buffered on the sequence parity (a CTA cannot publish exchange k+2 before every CTA has
//     published k+1, i.e. has finished reading k);
//   * source vertices / normals are read-only for the whole Align: after the first iteration they come out of L1
//     (66 KB per SM at VGA), and most target gathers do too, because no fence ever invalidates L1;
//   * multi-GPU (vh_set_peers): CTA 0 scatters the rank's system into every rank's mailbox over NVLink and every
//     CTA of every rank polls its own GPU's mailbox, adds the P contributions in rank order and solves -- still
//     one kernel per Align, one one-way NVLink trip per iteration.
// The grid must be co-resident (CTAs spin on each other): cooperative launch, one 512-thread CTA per SM.
#include "icp_device.cuh"

namespace vh {

constexpr int kAlignBatch = 5;        // pixels in flight per thread (as k_icp_iter)

// sequence-tagged row of the intra-GPU exchange: ll[slot][cta][32]
__device__ __forceinline__ unsigned long long* llRow(unsigned long long* ll, unsigned slot, unsigned cta) {
    return ll + ((size_t)slot * kIcpMaxBlocks + cta) * 32u;
}

template <class P>
__global__ void __launch_bounds__(kIcpThreads, 1) k_icp_align(View v, IcpState* st, unsigned long long* ll,
                                                              const float4* __restrict__ in, const float4* __restrict__ inN,
                                                              const float4* __restrict__ tg, const float4* __restrict__ tgN,
                                                              int row0, int row1, int iterations, PeerView pv) {
    __shared__ float sDelta[16];
    __shared__ double sD[16];
    __shared__ double sP[16];
    __shared__ float sm[kIcpThreads / 32][32];
    __shared__ double sRows[kIcpThreads / 32][33];
    __shared__ float sSys[32];
    __shared__ int sStop;
    constexpr int B = kAlignBatch;
    constexpr int G = kIcpThreads / 32;                     // row groups of the exchange read: one warp per group
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int begin = row0 * v.W, end = row1 * v.W;
    const int T = gridDim.x * blockDim.x;
    IcpDev* dev = devOf(st);

    // every CTA reads the sequence base before any CTA can finish (the base is only rewritten by CTA 0 after the last
    // exchange, which needs every CTA's contribution)
    const unsigned seq0 = __ldcg(&v.ctr->icpSeq);
    if (threadIdx.x < 16) {
        sDelta[threadIdx.x] = __ldcg(st->delta + threadIdx.x);
        sD[threadIdx.x] = __ldcg(dev->D + threadIdx.x);
    }
    if (threadIdx.x == 0) sStop = 0;
    __syncthreads();

    const bool haveM = P::fixed && inN != nullptr && v.icpNormalThres > -1.0f;
    int solved = 0, exchanges = 0;
    for (int it = 0; it < iterations; ++it) {
        // ---- association + residual + Jacobian row + 29 running sums (as k_icp_iter) --------------------------
        float acc[29];
#pragma unroll
        for (int k = 0; k < 29; ++k) acc[k] = 0.f;
        int i0 = begin + blockIdx.x * blockDim.x + threadIdx.x;
        float4 s[B];
#pragma unroll
        for (int j = 0; j < B; ++j) {
            const int idx = i0 + j * T;
            s[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (idx < end) s[j] = __ldg(in + idx);
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
                    if (haveM) m[j] = __ldg(inN + i0 + j * T);
                }
            }
            i0 += B * T;
            const bool more = i0 < end;
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
        const float tot = blockReduce29(acc, sm);           // warp 0, lane k < 29: this CTA's sum k

        // ---- exchange = grid barrier: publish my row, read everybody's ---------------------------------------
        const unsigned seq = seq0 + (unsigned)it + 1u;
        const unsigned slot = seq & 1u;
        if (warp == 0) llStoreGpu(llRow(ll, slot, blockIdx.x) + lane, lane < 29 ? tot : 0.f, seq);
        {
            // warp w reads rows w, w + G, ...; all first-round loads of a thread are in flight together
            const int rows = ((int)gridDim.x - warp + G - 1) / G;       // rows this warp owns
            double a = 0.0;
            for (int r0 = 0; r0 < rows; r0 += 8) {
                unsigned long long w[8];
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (r0 + k < rows) w[k] = llLoadGpu(llRow(ll, slot, (unsigned)(warp + (r0 + k) * G)) + lane);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (r0 + k < rows) {
                        while (!llReady(w[k], seq)) w[k] = llLoadGpu(llRow(ll, slot, (unsigned)(warp + (r0 + k) * G)) + lane);
                        a += (double)llValue(w[k]);
                    }
                }
            }
            sRows[warp][lane] = a;
        }
        __syncthreads();
        ++exchanges;
        if (warp == 0) {
            double t = 0.0;
#pragma unroll
            for (int g = 0; g < G; ++g) t += sRows[g][lane];
            float f = (float)t;
            if (pv.world > 1) {                              // the cross-GPU collective, still inside the kernel
                if (blockIdx.x == 0) peerScatter(pv, f, seq);
                f = peerGather(pv, seq);
            }
            sSys[lane] = f;
            __syncwarp();
            double dcol[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) dcol[k] = sD[k * 4 + (lane & 3)];
            double pij;
            const bool ok = solveCoreWarp(sSys, dcol, P::fixed, sP, pij);
            if (ok) {
                if (lane < 16) { sD[lane] = pij; sDelta[lane] = (float)pij; }
            } else if (lane == 0) sStop = 1;
        }
        __syncthreads();
        if (sStop) break;                                    // uniform over the grid (and over the ranks): same sums everywhere
        ++solved;
    }

    if (blockIdx.x == 0 && warp == 0) {                      // publish the result; every CTA holds the same one
        if (lane < 16) { st->delta[lane] = sDelta[lane]; dev->D[lane] = sD[lane]; }
        if (exchanges > 0) st->system[lane] = sSys[lane];
        if (lane == 0) {
            st->iterations += solved;
            v.ctr->icpConverged = sStop;
            v.ctr->icpSeq = seq0 + (unsigned)exchanges;
        }
    }
}

cudaError_t launch_icp_align(vh_context* c, const float4* in, const float4* inN, const float4* tg, const float4* tgN, int row0,
                             int row1, int iterations, bool peers, cudaStream_t s) {
    if (iterations <= 0) return cudaSuccess;
    int g = ((row1 - row0) * c->v.W + kIcpThreads - 1) / kIcpThreads;
    int cap = c->icpCtas > 0 ? c->icpCtas : c->numSMs;
    if (cap > c->numSMs) cap = c->numSMs;                   // one CTA per SM: the grid must be co-resident
    if (cap > kIcpMaxBlocks) cap = kIcpMaxBlocks;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    PeerView pv{};
    pv.world = 1;
    if (peers) pv = c->peers;
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
    if (c->cfg.policy == VH_POLICY_FIXED)
        return cudaLaunchKernelEx(&cfg, k_icp_align<Fixed>, c->v, c->icp, c->icpLL, in, inN, tg, tgN, row0, row1, iterations, pv);
    return cudaLaunchKernelEx(&cfg, k_icp_align<RefExact>, c->v, c->icp, c->icpLL, in, inN, tg, tgN, row0, row1, iterations, pv);
}

}  // namespace vh
