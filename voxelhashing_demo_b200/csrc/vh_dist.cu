// vh_dist.cu -- multi-GPU plumbing behind the C ABI (vh_dist_*): one process per GPU, CUDA IPC peer memory over
// NVLink / NVSwitch, no Python and no NCCL on the data path (the reference has nothing here).
//
// Every rank owns one exchange REGION in device memory, mapped into every other rank's address space through a CUDA IPC
// handle (64 bytes, exchanged by the host program however it likes: MPI, sockets, files ...):
//     [ ICP mailboxes, kPeerBytes ]      the {value, sequence} words of the fused all-reduce (icp_device.cuh, k_track.cu)
//     [ arrived[3], consumed[8] ]        u64 flags of the frame broadcast
//     [ landing[3][W*H u16] ]            three landing buffers for the depth frames
// Frame broadcast (replaces ncclBroadcast + its stream hand-off): rank 0's kernel stores frame k straight into landing
// slot k % 3 of EVERY rank's region (P2P stores), fences, and the last CTA publishes arrived[slot] = k + 1 everywhere; a
// one-warp kernel on each rank waits for its own flag, and the event behind it is the "input ready" event of
// vh_pipeline_push_device_ready.  Flow control: after a rank has pre-processed frame k it stores consumed[rank] = k + 1
// into rank 0's region; rank 0's push of frame k + 3 (same slot) first waits for consumed[r] >= k + 1 for every r.
#include <cstdio>
#include <cstring>
#include <string>

#include "icp_device.cuh"

using namespace vh;

namespace {

constexpr int kSlots = 3;
constexpr size_t kFlagOffset = (kPeerBytes + 255) & ~(size_t)255;
constexpr size_t kFlagBytes = 256;                          // arrived[3] at +0, consumed[8] at +64, time-out count at +128
constexpr size_t kLandingOffset = kFlagOffset + kFlagBytes;

__device__ __forceinline__ unsigned long long ldAcquireSys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void stReleaseSys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

struct Regions { unsigned char* r[kMaxPeers]; };

// rank 0: frame -> landing[slot] of every rank, then arrived[slot] = seq on every rank (last CTA)
__global__ void __launch_bounds__(256) k_frame_push(Regions reg, int world, const uint4* __restrict__ src, size_t n16, size_t landingStride,
                                                    int slot, unsigned long long seq, unsigned int* ticket) {
    // the slot still holds frame seq - 1 - kSlots until every rank has consumed it
    if (threadIdx.x < (unsigned)world && seq > (unsigned long long)kSlots) {
        const unsigned long long* consumed = reinterpret_cast<const unsigned long long*>(reg.r[0] + kFlagOffset + 64) + threadIdx.x;
        const long long t0 = clock64();
        while (ldAcquireSys(consumed) + kSlots < seq)
            if (clock64() - t0 > kPeerSpinCycles) {          // a rank died: overwrite the slot rather than hang; counted, see vh_dist_timeouts
                atomicAdd(reinterpret_cast<unsigned long long*>(reg.r[0] + kFlagOffset + 128), 1ull);
                break;
            }
    }
    __syncthreads();
    // read once (the source may be pinned HOST memory: the frame then crosses the host link exactly once), store P times
    const size_t off = kLandingOffset + (size_t)slot * landingStride;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 w = __ldg(src + i);
        for (int p = 0; p < world; ++p) reinterpret_cast<uint4*>(reg.r[p] + off)[i] = w;
    }
    __threadfence_system();                                  // this thread's stores are performed system-wide
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (last) {
        if (threadIdx.x < (unsigned)world)
            stReleaseSys(reinterpret_cast<unsigned long long*>(reg.r[threadIdx.x] + kFlagOffset) + slot, seq);
        if (threadIdx.x == 0) *ticket = 0;
    }
}

// every rank: wait until frame seq has landed in my slot
__global__ void k_frame_wait(const unsigned char* region, int slot, unsigned long long seq) {
    if (threadIdx.x == 0) {
        const unsigned long long* arrived = reinterpret_cast<const unsigned long long*>(region + kFlagOffset) + slot;
        const long long t0 = clock64();
        while (ldAcquireSys(arrived) < seq)
            if (clock64() - t0 > kPeerSpinCycles) {          // rank 0 died or never pushed: give up instead of hanging the GPU
                atomicAdd(const_cast<unsigned long long*>(reinterpret_cast<const unsigned long long*>(region + kFlagOffset + 128)), 1ull);
                break;
            }
    }
}

// every rank: tell rank 0 that `count` frames have been consumed here
__global__ void k_frame_consumed(unsigned char* region0, int rank, unsigned long long count) {
    if (threadIdx.x == 0) stReleaseSys(reinterpret_cast<unsigned long long*>(region0 + kFlagOffset + 64) + rank, count);
}

thread_local std::string g_distError;
int dfail(int code, const char* what, cudaError_t e = cudaSuccess) {
    g_distError = what;
    if (e != cudaSuccess) { g_distError += ": "; g_distError += cudaGetErrorString(e); }
    fprintf(stderr, "vh_dist: %s\n", g_distError.c_str());
    return code;
}
#define DCUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return dfail(VH_ERR_CUDA, #expr, _e); } while (0)

}  // namespace

struct vh_dist {
    vh_context* ctx;
    int rank, world;
    unsigned char* region;                 // this rank's region
    Regions peers;                         // every rank's region as mapped here
    bool opened[kMaxPeers];
    bool connected;
    size_t frameBytes, landingStride, regionBytes;
    unsigned long long pushed;             // frames broadcast so far
    unsigned long long consumed;           // frames reported consumed so far
    unsigned int* ticket;
    cudaStream_t stream;                   // the broadcast runs here, beside the caller's work
    cudaEvent_t evIn, evArrived[kSlots];
};

extern "C" {

int vh_dist_create(vh_context* ctx, int rank, int world, vh_dist** out) {
    if (!ctx || !out || world < 1 || world > kMaxPeers || rank < 0 || rank >= world) return dfail(VH_ERR_INVALID, "vh_dist_create: bad argument (world <= 8)");
    vh_dist* d = new vh_dist();
    memset(static_cast<void*>(d), 0, sizeof(*d));
    d->ctx = ctx; d->rank = rank; d->world = world;
    d->frameBytes = (size_t)ctx->v.W * ctx->v.H * sizeof(uint16_t);
    d->landingStride = (d->frameBytes + 255) & ~(size_t)255;
    d->regionBytes = kLandingOffset + kSlots * d->landingStride;
    cudaError_t e = cudaMalloc((void**)&d->region, d->regionBytes);     // plain cudaMalloc: exportable through CUDA IPC
    if (e == cudaSuccess) e = cudaMemset(d->region, 0, d->regionBytes);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d->ticket, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemset(d->ticket, 0, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d->evIn, cudaEventDisableTiming);
    for (int i = 0; i < kSlots && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&d->evArrived[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { vh_dist_destroy(d); return dfail(VH_ERR_CUDA, "vh_dist_create", e); }
    d->peers.r[rank] = d->region;
    if (world == 1) d->connected = true;
    *out = d;
    return VH_OK;
}

unsigned long long vh_dist_handle_bytes(void) { return (unsigned long long)sizeof(cudaIpcMemHandle_t); }

int vh_dist_export(vh_dist* d, void* handle) {
    if (!d || !handle) return dfail(VH_ERR_INVALID, "vh_dist_export: null argument");
    cudaIpcMemHandle_t h;
    DCUDA(cudaIpcGetMemHandle(&h, d->region));
    memcpy(handle, &h, sizeof(h));
    return VH_OK;
}

int vh_dist_connect(vh_dist* d, const void* handles) {
    if (!d || (!handles && d->world > 1)) return dfail(VH_ERR_INVALID, "vh_dist_connect: null argument");
    void* bufs[kMaxPeers] = {nullptr};
    for (int p = 0; p < d->world; ++p) {
        if (p != d->rank && !d->opened[p]) {
            cudaIpcMemHandle_t h;
            memcpy(&h, static_cast<const unsigned char*>(handles) + (size_t)p * sizeof(h), sizeof(h));
            void* ptr = nullptr;
            DCUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
            d->peers.r[p] = static_cast<unsigned char*>(ptr);
            d->opened[p] = true;
        }
        bufs[p] = d->peers.r[p];                              // the mailboxes sit at the start of the region
    }
    d->connected = true;
    if (d->world > 1) return vh_set_peers(d->ctx, d->rank, d->world, bufs);
    return VH_OK;
}

// Rank 0 passes the frame (device memory or pinned host memory, complete once stream s gets here); every other rank passes NULL.  Every rank
// gets the address of its landing buffer and the event that fires when the frame is complete in it.
int vh_dist_broadcast_frame(vh_dist* d, const uint16_t* d_depth, const uint16_t** d_frame, void** ready_event, vh_stream s) {
    if (!d || !d_frame || !ready_event || (d->rank == 0 && !d_depth)) return dfail(VH_ERR_INVALID, "vh_dist_broadcast_frame: bad argument");
    if (!d->connected) return dfail(VH_ERR_INVALID, "vh_dist_broadcast_frame: vh_dist_connect first");
    if ((d->frameBytes & 15) != 0) return dfail(VH_ERR_INVALID, "vh_dist_broadcast_frame: W*H must be a multiple of 8");
    const int slot = (int)(d->pushed % kSlots);
    const unsigned long long seq = d->pushed + 1;
    if (d->rank == 0) {
        DCUDA(cudaEventRecord(d->evIn, reinterpret_cast<cudaStream_t>(s)));
        DCUDA(cudaStreamWaitEvent(d->stream, d->evIn, 0));
        k_frame_push<<<16, 256, 0, d->stream>>>(d->peers, d->world, reinterpret_cast<const uint4*>(d_depth), d->frameBytes / 16,
                                                d->landingStride, slot, seq, d->ticket);
        DCUDA(cudaGetLastError());
    }
    k_frame_wait<<<1, 32, 0, d->stream>>>(d->region, slot, seq);
    DCUDA(cudaGetLastError());
    DCUDA(cudaEventRecord(d->evArrived[slot], d->stream));
    *d_frame = reinterpret_cast<const uint16_t*>(d->region + kLandingOffset + (size_t)slot * d->landingStride);
    *ready_event = d->evArrived[slot];
    d->pushed += 1;
    return VH_OK;
}

// Stream s has consumed (pre-processed) the oldest frame not yet reported: its landing slot may be overwritten.
int vh_dist_frame_consumed(vh_dist* d, vh_stream s) {
    if (!d) return dfail(VH_ERR_INVALID, "vh_dist_frame_consumed: null argument");
    if (d->consumed >= d->pushed) return dfail(VH_ERR_INVALID, "vh_dist_frame_consumed: no frame outstanding");
    k_frame_consumed<<<1, 32, 0, reinterpret_cast<cudaStream_t>(s)>>>(d->peers.r[0], d->rank, d->consumed + 1);
    DCUDA(cudaGetLastError());
    d->consumed += 1;
    return VH_OK;
}

// Frames this rank stopped waiting for (a peer died or never pushed); synchronises.
int vh_dist_timeouts(vh_dist* d) {
    if (!d) return -1;
    unsigned long long n = 0;
    cudaDeviceSynchronize();
    cudaMemcpy(&n, d->region + kFlagOffset + 128, sizeof(n), cudaMemcpyDeviceToHost);
    return (int)n;
}

void vh_dist_destroy(vh_dist* d) {
    if (!d) return;
    cudaDeviceSynchronize();
    for (int p = 0; p < kMaxPeers; ++p) if (d->opened[p]) cudaIpcCloseMemHandle(d->peers.r[p]);
    if (d->stream) cudaStreamDestroy(d->stream);
    if (d->evIn) cudaEventDestroy(d->evIn);
    for (int i = 0; i < kSlots; ++i) if (d->evArrived[i]) cudaEventDestroy(d->evArrived[i]);
    cudaFree(d->ticket);
    cudaFree(d->region);
    delete d;
}

}  // extern "C"
