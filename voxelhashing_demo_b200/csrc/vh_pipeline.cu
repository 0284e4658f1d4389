// vh_pipeline.cu -- native per-frame runtime: the reference's host loop
// (Application.cpp:73-84: preProcess -> Align -> getTransform -> integrate) as ONE stream-ordered
// sequence with no host synchronisation, captured into CUDA graphs.
//
// Per frame:   [H2D depth] -> preprocess -> { ICP x iterations -> pose <- pose * delta -> alloc ->
//              compact -> integrate [-> raycast] } -> [D2H pose]
// The braces are a CUDA graph (one per buffer parity in frame-to-frame mode, where the previous
// frame's maps are the ICP target).  The pose lives in device memory and is chained on the device,
// so frame k+1 can be enqueued before frame k has finished.
//
// VH_PIPE_OVERLAP (frame-to-frame):
//   caller's stream:  preprocess(k) -> Align(k) [one persistent kernel; its tail also chains the pose, T_k = T_{k-1} * delta]
//   fusion stream:    [wait Align(k)] -> frame constants(k) -> { alloc -> compact -> integrate } (a graph per parity)
// The caller's stream never waits for a fusion (only for buffer reuse two frames later and for the 3-microsecond frame-
// constants kernel that reads the pose the next Align will overwrite), so its critical path per frame is preprocess +
// Align; the Align grid leaves a few SMs free (vh_context::icpCtas) and fusion(k) runs there beside Align(k+1).
// On a partitioned context with peer mailboxes (vh_set_peers) the Align kernel reduces this rank's image rows and
// carries the cross-GPU all-reduce inside.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "vh_internal.h"

using namespace vh;

constexpr int kMapSets = 3;

struct vh_pipeline {
    vh_context* ctx;
    int iterations;
    int mode;
    bool useGraph;
    long long frame;
    long long launches;
    float* d_poseBuf[2];           // camera -> world, row-major; the overlapped schedule ping-pongs so that Align(k+1) never
    int poseIdx;                   // overwrites the pose the frame-constants kernel of frame k still has to read
    float* d_pose;                 // = d_poseBuf[poseIdx]: pose of the latest frame
    uint16_t* d_depthStage[2];     // H2D landing buffers of push_host (double-buffered)
    cudaStream_t copyStream;       // H2D of frame k+1 overlaps the compute of frame k
    cudaEvent_t evCopied[2], evConsumed[2];
    long long hostFrames;
    // THREE map sets, used round robin (frame k -> set k % 3): the maps of frame k are the source of Align(k), the target
    // of Align(k+1) and the input of fusion(k); with two sets the pre-processing of frame k+2 had to wait for fusion(k) --
    // which runs beside Align(k+1) and finishes after it (r2 timeline: 13 us per frame at VGA, the whole integrate at 2 mm)
    float4* verts[kMapSets];
    float4* normals[kMapSets];
    float* depthf[kMapSets];
    float4* modelVerts;            // raycast maps (frame-to-model)
    float4* modelNormals;
    cudaGraph_t graph[kMapSets];
    cudaGraphExec_t exec[kMapSets];
    bool haveGraph[kMapSets];
    // overlapped schedule (VH_PIPE_OVERLAP): fusion of frame k runs on its own stream beside preprocess + ICP of frame k+1
    bool overlap;
    bool fusedPre;                 // VH_PIPE_FUSED_PRE=1: pre-processing inside the Align kernel (see pushFrame)
    cudaStream_t fuseStream, trackStream, prepStream;
    cudaEvent_t evIn, evPreR[kMapSets], evAlignedR[kMapSets], evFrameSetR[kMapSets];
    cudaEvent_t evFused[kMapSets];
    bool fusePending;              // a fusion has been enqueued since the last reset
    int lastFusePar;
    cudaGraph_t fuseGraph[kMapSets];
    cudaGraphExec_t fuseExec[kMapSets];
    bool haveFuseGraph[kMapSets];
};

namespace {
thread_local std::string g_pipeError;
int pfail(int code, const char* what, cudaError_t e = cudaSuccess) {
    g_pipeError = what;
    if (e != cudaSuccess) { g_pipeError += ": "; g_pipeError += cudaGetErrorString(e); }
    fprintf(stderr, "vh_pipeline: %s\n", g_pipeError.c_str());
    return code;
}
#define PCUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return pfail(VH_ERR_CUDA, #expr, _e); } while (0)

cudaError_t enqueueIcp(vh_pipeline* p, int par, const uint16_t* d_depth, const float* d_poseIn, float* d_poseOut, cudaStream_t s, int* n) {
    vh_context* c = p->ctx;
    const float4* tg = p->mode == VH_TRACK_FRAME_TO_MODEL ? p->modelVerts : p->verts[(par + kMapSets - 1) % kMapSets];
    const float4* tgN = p->mode == VH_TRACK_FRAME_TO_MODEL ? p->modelNormals : p->normals[(par + kMapSets - 1) % kMapSets];
    // Partitioned context with peer mailboxes (vh_set_peers): this rank reduces its share of the image rows and the
    // 32-float all-reduce over NVLink runs inside the kernel's epilogue -- every rank must push the same frames.
    const int world = c->peers.world > 1 ? c->peers.world : 1, rank = world > 1 ? c->peers.rank : 0;
    const int base = c->v.H / world, rem = c->v.H % world;
    const int row0 = rank * base + (rank < rem ? rank : rem), row1 = row0 + base + (rank < rem ? 1 : 0);
    // CameraTracking.cpp:35-67: the whole iteration loop is one persistent kernel (k_track.cu)
    cudaError_t e = launch_icp_align(c, d_depth, d_depth ? p->depthf[par] : nullptr, p->verts[par], p->normals[par], tg, tgN, row0, row1,
                                     p->iterations, world > 1, d_poseIn, d_poseOut, s);
    if (e != cudaSuccess) return e;
    *n = 1;
    return cudaSuccess;
}
cudaError_t enqueueFusion(vh_pipeline* p, int par, cudaStream_t s, int* n) {
    vh_context* c = p->ctx;
    cudaError_t e;
    int k = 0;
    if (c->cfg.policy == VH_POLICY_REF_EXACT) { e = launch_reset_mutex(c, s); if (e != cudaSuccess) return e; ++k; }
    // SDF_Hashtable.cpp:27 -- from the 4 B / pixel metric depth when it determines the vertex exactly, else the vertex map
    e = alloc_depthf_ok(c) ? launch_alloc_depthf(c, p->depthf[par], s) : launch_alloc(c, p->verts[par], s);
    if (e != cudaSuccess) return e;
    e = launch_compact(c, s);                                              // :30
    if (e != cudaSuccess) return e;
    e = launch_integrate(c, p->verts[par], c->cfg.policy == VH_POLICY_FIXED ? p->depthf[par] : nullptr, -1, s);   // :36
    if (e != cudaSuccess) return e;
    *n = k + 3;
    return cudaSuccess;
}

// the tracked part of a frame (everything after preprocess); returns kernels launched via *n
cudaError_t enqueueBody(vh_pipeline* p, int par, bool track, cudaStream_t s, int* n) {
    vh_context* c = p->ctx;
    cudaError_t e;
    int k = 0;
    const float4* in = p->verts[par];
    if (track) {
        e = enqueueIcp(p, par, nullptr, nullptr, nullptr, s, &k);
        if (e != cudaSuccess) return e;
        e = launch_set_frame_device(c, p->d_pose, c->icp->delta, p->d_pose, s);   // T_k = T_{k-1} * delta
    } else {
        e = launch_set_frame_device(c, p->d_pose, nullptr, nullptr, s);
    }
    if (e != cudaSuccess) return e;
    ++k;
    if (c->cfg.policy == VH_POLICY_REF_EXACT) { e = launch_reset_mutex(c, s); if (e != cudaSuccess) return e; ++k; }
    e = alloc_depthf_ok(c) ? launch_alloc_depthf(c, p->depthf[par], s) : launch_alloc(c, in, s);   // SDF_Hashtable.cpp:27
    if (e != cudaSuccess) return e;
    e = launch_compact(c, s);                                              // :30
    if (e != cudaSuccess) return e;
    e = launch_integrate(c, in, c->cfg.policy == VH_POLICY_FIXED ? p->depthf[par] : nullptr, -1, s);   // :36
    if (e != cudaSuccess) return e;
    k += 3;
    if (p->mode == VH_TRACK_FRAME_TO_MODEL) {
        e = launch_raycast(c, p->modelVerts, p->modelNormals, s);
        if (e != cudaSuccess) return e;
        ++k;
    }
    *n = k;
    return cudaSuccess;
}
}  // namespace

extern "C" {

int vh_pipeline_create(vh_context* ctx, int icpIterations, int mode, int useGraph, vh_pipeline** out) {
    if (!ctx || !out) return pfail(VH_ERR_INVALID, "vh_pipeline_create: null argument");
    if (mode == VH_TRACK_FRAME_TO_MODEL && ctx->cfg.policy != VH_POLICY_FIXED)
        return pfail(VH_ERR_INVALID, "vh_pipeline_create: frame-to-model tracking needs the Fixed policy");
    if (mode == VH_TRACK_FRAME_TO_MODEL && ctx->v.partCount > 1)
        return pfail(VH_ERR_UNSUPPORTED, "vh_pipeline_create: frame-to-model tracking raycasts the model, which a partitioned context holds only 1/P of");
    vh_pipeline* p = new vh_pipeline();
    memset(static_cast<void*>(p), 0, sizeof(*p));
    p->ctx = ctx;
    p->iterations = icpIterations > 0 ? icpIterations : ctx->cfg.icpIterations;
    p->mode = mode;
    p->useGraph = (useGraph & VH_PIPE_GRAPH) != 0;
    // the overlapped schedule needs a model-independent ICP target (frame-to-frame) and the graph path
    p->overlap = (useGraph & VH_PIPE_OVERLAP) != 0 && p->useGraph && mode == VH_TRACK_FRAME_TO_FRAME;
    { const char* env = getenv("VH_PIPE_FUSED_PRE"); p->fusedPre = env && env[0] == '1'; }
    const size_t px = (size_t)ctx->v.W * ctx->v.H;
    cudaError_t e = cudaSuccess;
    auto chk = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
    chk(cudaMalloc((void**)&p->d_poseBuf[0], 2 * 16 * sizeof(float)));
    p->d_poseBuf[1] = p->d_poseBuf[0] ? p->d_poseBuf[0] + 16 : nullptr;
    p->d_pose = p->d_poseBuf[0];
    chk(cudaStreamCreateWithFlags(&p->copyStream, cudaStreamNonBlocking));
    for (int i = 0; i < kMapSets; ++i) {
        chk(cudaMalloc((void**)&p->verts[i], px * sizeof(float4)));
        chk(cudaMalloc((void**)&p->normals[i], px * sizeof(float4)));
        chk(cudaMalloc((void**)&p->depthf[i], px * sizeof(float)));
    }
    for (int i = 0; i < 2; ++i) {
        chk(cudaMalloc((void**)&p->d_depthStage[i], px * sizeof(uint16_t)));
        chk(cudaEventCreateWithFlags(&p->evCopied[i], cudaEventDisableTiming));
        chk(cudaEventCreateWithFlags(&p->evConsumed[i], cudaEventDisableTiming));
    }
    if (p->overlap) {
        // Align runs on its own HIGH-priority stream: when Align(k) retires, Align(k+1) and the fusion of frame k become
        // runnable together, and the block scheduler must give the SMs to the cooperative Align grid first -- the fusion
        // kernels then fill the SMs the Align grid leaves free (a caller's stream cannot be given a lower priority than
        // the default, so the tracking is raised instead)
        int least = 0, greatest = 0;
        chk(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        chk(cudaStreamCreateWithPriority(&p->trackStream, cudaStreamNonBlocking, greatest));
        chk(cudaStreamCreateWithPriority(&p->prepStream, cudaStreamNonBlocking, greatest));
        chk(cudaEventCreateWithFlags(&p->evIn, cudaEventDisableTiming));
        for (int i = 0; i < kMapSets; ++i) {
            chk(cudaEventCreateWithFlags(&p->evPreR[i], cudaEventDisableTiming));
            chk(cudaEventCreateWithFlags(&p->evAlignedR[i], cudaEventDisableTiming));
            chk(cudaEventCreateWithFlags(&p->evFrameSetR[i], cudaEventDisableTiming));
        }
        chk(cudaStreamCreateWithPriority(&p->fuseStream, cudaStreamNonBlocking, least));
        for (int i = 0; i < kMapSets; ++i) chk(cudaEventCreateWithFlags(&p->evFused[i], cudaEventDisableTiming));
    }
    if (mode == VH_TRACK_FRAME_TO_MODEL) {
        chk(cudaMalloc((void**)&p->modelVerts, px * sizeof(float4)));
        chk(cudaMalloc((void**)&p->modelNormals, px * sizeof(float4)));
    }
    if (e != cudaSuccess) { vh_pipeline_destroy(p); return pfail(VH_ERR_CUDA, "vh_pipeline_create: cudaMalloc", e); }
    const float ident[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    cudaMemcpy(p->d_pose, ident, sizeof(ident), cudaMemcpyHostToDevice);
    *out = p;
    return VH_OK;
}

void vh_pipeline_destroy(vh_pipeline* p) {
    if (!p) return;
    for (int i = 0; i < kMapSets; ++i) {
        if (p->haveGraph[i]) { cudaGraphExecDestroy(p->exec[i]); cudaGraphDestroy(p->graph[i]); }
        if (p->haveFuseGraph[i]) { cudaGraphExecDestroy(p->fuseExec[i]); cudaGraphDestroy(p->fuseGraph[i]); }
        cudaFree(p->verts[i]); cudaFree(p->normals[i]); cudaFree(p->depthf[i]);
    }
    for (int i = 0; i < 2; ++i) {
        cudaFree(p->d_depthStage[i]);
        if (p->evCopied[i]) cudaEventDestroy(p->evCopied[i]);
        if (p->evConsumed[i]) cudaEventDestroy(p->evConsumed[i]);
    }
    if (p->copyStream) cudaStreamDestroy(p->copyStream);
    if (p->fuseStream) { cudaStreamSynchronize(p->fuseStream); cudaStreamDestroy(p->fuseStream); }
    if (p->trackStream) { cudaStreamSynchronize(p->trackStream); cudaStreamDestroy(p->trackStream); }
    if (p->prepStream) { cudaStreamSynchronize(p->prepStream); cudaStreamDestroy(p->prepStream); }
    if (p->evIn) cudaEventDestroy(p->evIn);
    for (int i = 0; i < kMapSets; ++i) {
        if (p->evPreR[i]) cudaEventDestroy(p->evPreR[i]);
        if (p->evAlignedR[i]) cudaEventDestroy(p->evAlignedR[i]);
        if (p->evFrameSetR[i]) cudaEventDestroy(p->evFrameSetR[i]);
    }
    for (int i = 0; i < kMapSets; ++i) if (p->evFused[i]) cudaEventDestroy(p->evFused[i]);
    cudaFree(p->modelVerts); cudaFree(p->modelNormals); cudaFree(p->d_poseBuf[0]);
    delete p;
}

// New sequence: frame counter and pose; the table itself is reset with vh_reset.
int vh_pipeline_reset(vh_pipeline* p, const float* pose16_host, vh_stream s) {
    if (!p) return pfail(VH_ERR_INVALID, "vh_pipeline_reset: null pipeline");
    const float ident[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    cudaStream_t st = reinterpret_cast<cudaStream_t>(s);
    if (p->prepStream) PCUDA(cudaStreamSynchronize(p->prepStream));
    if (p->trackStream) PCUDA(cudaStreamSynchronize(p->trackStream));
    if (p->fuseStream) PCUDA(cudaStreamSynchronize(p->fuseStream));
    p->fusePending = false;
    p->poseIdx = 0;
    p->d_pose = p->d_poseBuf[0];
    PCUDA(cudaStreamSynchronize(st));
    PCUDA(cudaMemcpy(p->d_pose, pose16_host ? pose16_host : ident, 16 * sizeof(float), cudaMemcpyHostToDevice));
    PCUDA(launch_icp_reset(p->ctx, true, st));
    p->frame = 0;
    return VH_OK;
}

// The overlapped schedule (VH_PIPE_OVERLAP).  Four streams, frame k uses map set m = k % 3:
//   prep   [input ready] [fusion(k-3), Align(k-2) done: set m is free] -> preprocess(k)
//   track  [preprocess(k)] -> Align(k) + pose chain                      (high priority, in order: Align(k-1) before)
//   fuse   [Align(k)] -> frame constants(k) -> { alloc, compact, integrate }(k)
//   caller [Align(k)]                                                     (the contract: ordered behind the pose)
// inputOnSt: the depth image is produced by work already enqueued on the caller's stream (vh_pipeline_push_device);
// otherwise it is complete once `ready` (may be null: now) has fired -- then preprocess(k+1) does not wait for the
// caller's stream, i.e. not for Align(k), and the tracking chain per frame is the Align kernel alone.
static int pushFrameOverlap(vh_pipeline* p, const uint16_t* d_depth, cudaStream_t st, bool inputOnSt, cudaEvent_t ready,
                            cudaEvent_t afterPreprocess) {
    vh_context* c = p->ctx;
    const int m = (int)(p->frame % kMapSets);
    const bool track = p->frame > 0 && p->mode != VH_TRACK_NONE;
    // Tracked frames CAN run the pre-processing as the prologue of the Align kernel (k_track.cu, vh_track_frame).
    // Measured at VGA (r2, tools/align_trace.py pre): the prologue needs 7.4 us at the persistent kernel's 16 warps per
    // SM (5 IEEE divisions + a normalisation per pixel) + 2.7 us of fence / barrier / fence, ON the tracking chain, against
    // a stand-alone pass that runs beside the previous Align: 7 820 vs 8 450 frames/s.  Opt-in: VH_PIPE_FUSED_PRE=1.
    const bool fusedPre = p->fusedPre && track && c->v.bilatLut == nullptr;
    cudaStream_t ps = fusedPre ? p->trackStream : p->prepStream;
    if (inputOnSt) {
        PCUDA(cudaEventRecord(p->evIn, st));
        PCUDA(cudaStreamWaitEvent(ps, p->evIn, 0));
    } else if (ready) {
        PCUDA(cudaStreamWaitEvent(ps, ready, 0));
    }
    if (p->frame >= kMapSets) PCUDA(cudaStreamWaitEvent(ps, p->evFused[m], 0));                 // fusion(k-3) read set m
    if (p->frame >= 2) PCUDA(cudaStreamWaitEvent(ps, p->evAlignedR[(m + 1) % kMapSets], 0));     // Align(k-2) had it as target
    if (!fusedPre) {
        PCUDA(launch_preprocess(c, d_depth, p->verts[m], p->normals[m], p->depthf[m], ps));     // Application.cpp:73
        if (afterPreprocess) PCUDA(cudaEventRecord(afterPreprocess, ps));  // the raw depth buffer may be overwritten from here on
        p->launches += (c->v.bilatLut != nullptr && c->cfg.policy == VH_POLICY_FIXED) ? 2 : 1;  // [k_bilateral +] k_preprocess
    }
    int nIcp = 0, nFuse = 0;
    if (track) {
        if (!fusedPre) {
            PCUDA(cudaEventRecord(p->evPreR[m], ps));
            PCUDA(cudaStreamWaitEvent(p->trackStream, p->evPreR[m], 0));
        }
        float* next = p->d_poseBuf[p->poseIdx ^ 1];
        // `next` still holds the pose of frame k-2: the frame-constants kernel of that frame (fusion stream) must have read it
        // (it runs two Aligns earlier in practice; the event makes it a guarantee)
        if (p->frame >= 2) PCUDA(cudaStreamWaitEvent(p->trackStream, p->evFrameSetR[(m + 1) % kMapSets], 0));
        // [Application.cpp:73] + CameraTracking.cpp:35-67 + the pose chain, one launch
        PCUDA(enqueueIcp(p, m, fusedPre ? d_depth : nullptr, p->d_pose, next, p->trackStream, &nIcp));
        p->poseIdx ^= 1;
        p->d_pose = next;
        PCUDA(cudaEventRecord(p->evAlignedR[m], p->trackStream));
        if (fusedPre && afterPreprocess) PCUDA(cudaEventRecord(afterPreprocess, p->trackStream));
    } else {
        PCUDA(cudaEventRecord(p->evAlignedR[m], ps));
    }
    PCUDA(cudaStreamWaitEvent(st, p->evAlignedR[m], 0));                   // the caller's stream is ordered behind the pose
    PCUDA(cudaStreamWaitEvent(p->fuseStream, p->evAlignedR[m], 0));
    PCUDA(launch_set_frame_device(c, p->d_pose, nullptr, nullptr, p->fuseStream));   // SDF_Hashtable.cpp:15-21
    PCUDA(cudaEventRecord(p->evFrameSetR[m], p->fuseStream));
    if (!p->haveFuseGraph[m]) {
        PCUDA(cudaStreamBeginCapture(p->fuseStream, cudaStreamCaptureModeThreadLocal));
        cudaError_t e = enqueueFusion(p, m, p->fuseStream, &nFuse);
        cudaError_t e2 = cudaStreamEndCapture(p->fuseStream, &p->fuseGraph[m]);
        if (e != cudaSuccess) return pfail(VH_ERR_CUDA, "pipeline capture", e);
        if (e2 != cudaSuccess) return pfail(VH_ERR_CUDA, "cudaStreamEndCapture", e2);
        PCUDA(cudaGraphInstantiate(&p->fuseExec[m], p->fuseGraph[m], 0));
        p->haveFuseGraph[m] = true;
    } else {
        nFuse = 3 + (c->cfg.policy == VH_POLICY_REF_EXACT ? 1 : 0);
    }
    PCUDA(cudaGraphLaunch(p->fuseExec[m], p->fuseStream));
    PCUDA(cudaEventRecord(p->evFused[m], p->fuseStream));
    p->fusePending = true;
    p->lastFusePar = m;
    p->launches += nIcp + 1 + nFuse;
    p->frame += 1;
    return VH_OK;
}

static int pushFrame(vh_pipeline* p, const uint16_t* d_depth, cudaStream_t st, bool inputOnSt, cudaEvent_t ready, cudaEvent_t afterPreprocess) {
    // (the legacy default stream is fine here: nothing is captured on the caller's stream, it only records / waits for events)
    if (p->overlap) return pushFrameOverlap(p, d_depth, st, inputOnSt, ready, afterPreprocess);
    vh_context* c = p->ctx;
    const int par = (int)(p->frame % kMapSets);            // map set of this frame
    const bool track = p->frame > 0 && p->mode != VH_TRACK_NONE;
    if (!inputOnSt && ready) PCUDA(cudaStreamWaitEvent(st, ready, 0));
    PCUDA(launch_preprocess(c, d_depth, p->verts[par], p->normals[par], p->depthf[par], st));   // Application.cpp:73
    if (afterPreprocess) PCUDA(cudaEventRecord(afterPreprocess, st));      // the raw depth buffer may be overwritten from here on
    p->launches += (c->v.bilatLut != nullptr && c->cfg.policy == VH_POLICY_FIXED) ? 2 : 1;   // [k_bilateral +] k_preprocess
    int n = 0;
    if (track && p->useGraph && st != nullptr) {
        // one graph per map set (the graph bakes in the set's addresses and those of its predecessor, the ICP target)
        if (!p->haveGraph[par]) {
            PCUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            cudaError_t e = enqueueBody(p, par, true, st, &n);
            cudaGraph_t g = nullptr;
            cudaError_t e2 = cudaStreamEndCapture(st, &g);
            if (e != cudaSuccess) return pfail(VH_ERR_CUDA, "pipeline capture", e);
            if (e2 != cudaSuccess) return pfail(VH_ERR_CUDA, "cudaStreamEndCapture", e2);
            p->graph[par] = g;
            PCUDA(cudaGraphInstantiate(&p->exec[par], g, 0));
            p->haveGraph[par] = true;
        } else {
            n = 1 + 4 + (p->mode == VH_TRACK_FRAME_TO_MODEL ? 1 : 0) + (c->cfg.policy == VH_POLICY_REF_EXACT ? 1 : 0);
        }
        PCUDA(cudaGraphLaunch(p->exec[par], st));
    } else {
        PCUDA(enqueueBody(p, par, track, st, &n));
    }
    p->launches += n;
    p->frame += 1;
    return VH_OK;
}

int vh_pipeline_push_device(vh_pipeline* p, const uint16_t* d_depth, vh_stream s) {
    if (!p || !d_depth) return pfail(VH_ERR_INVALID, "vh_pipeline_push_device: null argument");
    return pushFrame(p, d_depth, reinterpret_cast<cudaStream_t>(s), true, nullptr, nullptr);
}

// Same, for a depth image that is NOT produced on stream s: it is complete in device memory once `ready` (a cudaEvent_t,
// may be NULL = already complete) has fired.  With VH_PIPE_OVERLAP the pre-processing then runs beside the tracking of the
// previous frame instead of behind it.  The buffer must stay untouched until the frame has been pre-processed (i.e.
// until s, which is ordered behind the frame's pose, gets there).
int vh_pipeline_push_device_ready(vh_pipeline* p, const uint16_t* d_depth, void* ready_event, vh_stream s) {
    if (!p || !d_depth) return pfail(VH_ERR_INVALID, "vh_pipeline_push_device_ready: null argument");
    return pushFrame(p, d_depth, reinterpret_cast<cudaStream_t>(s), false, reinterpret_cast<cudaEvent_t>(ready_event), nullptr);
}

// e2e entry: depth in (pinned) host memory, pose back to host memory; returns after enqueueing.
// The H2D copy runs on the pipeline's own copy stream into one of two staging buffers, so the copy of frame
// k+1 overlaps the compute of frame k; events order copy -> preprocess and preprocess -> reuse of the buffer.
int vh_pipeline_push_host(vh_pipeline* p, const uint16_t* h_depth, float* h_pose_out16, vh_stream s) {
    if (!p || !h_depth) return pfail(VH_ERR_INVALID, "vh_pipeline_push_host: null argument");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(s);
    const size_t bytes = (size_t)p->ctx->v.W * p->ctx->v.H * sizeof(uint16_t);
    const int slot = (int)(p->hostFrames & 1);
    if (p->hostFrames >= 2) PCUDA(cudaStreamWaitEvent(p->copyStream, p->evConsumed[slot], 0));
    PCUDA(cudaMemcpyAsync(p->d_depthStage[slot], h_depth, bytes, cudaMemcpyHostToDevice, p->copyStream));
    PCUDA(cudaEventRecord(p->evCopied[slot], p->copyStream));
    p->hostFrames += 1;
    int rc = pushFrame(p, p->d_depthStage[slot], st, false, p->evCopied[slot], p->evConsumed[slot]);
    if (rc != VH_OK) return rc;
    if (h_pose_out16) PCUDA(cudaMemcpyAsync(h_pose_out16, p->d_pose, 16 * sizeof(float), cudaMemcpyDeviceToHost, st));
    return VH_OK;
}

// Overlapped schedule only: make stream s wait for the fusion of the latest pushed frame (a no-op otherwise).
int vh_pipeline_flush(vh_pipeline* p, vh_stream s) {
    if (!p) return pfail(VH_ERR_INVALID, "vh_pipeline_flush: null pipeline");
    if (p->fusePending) PCUDA(cudaStreamWaitEvent(reinterpret_cast<cudaStream_t>(s), p->evFused[p->lastFusePar], 0));
    return VH_OK;
}

int vh_pipeline_pose(vh_pipeline* p, float* pose16, vh_stream s) {
    if (!p || !pose16) return pfail(VH_ERR_INVALID, "vh_pipeline_pose: null argument");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(s);
    if (p->fusePending) PCUDA(cudaStreamWaitEvent(st, p->evFused[p->lastFusePar], 0));
    PCUDA(cudaMemcpyAsync(pose16, p->d_pose, 16 * sizeof(float), cudaMemcpyDeviceToHost, st));
    PCUDA(cudaStreamSynchronize(st));
    return VH_OK;
}

const float* vh_pipeline_pose_device(vh_pipeline* p) { return p ? p->d_pose : nullptr; }

// Stream-ordered D2H of the latest pose into (pinned) host memory: no synchronisation, no flush of the fusion stream.
int vh_pipeline_pose_async(vh_pipeline* p, float* h_pose16, vh_stream s) {
    if (!p || !h_pose16) return pfail(VH_ERR_INVALID, "vh_pipeline_pose_async: null argument");
    PCUDA(cudaMemcpyAsync(h_pose16, p->d_pose, 16 * sizeof(float), cudaMemcpyDeviceToHost, reinterpret_cast<cudaStream_t>(s)));
    return VH_OK;
}

// Dense metric depth (W x H floats) of the latest pushed frame, as the Fixed integration reads it.
int vh_pipeline_depthf(vh_pipeline* p, float** d_depthf) {
    if (!p || !d_depthf || p->frame == 0) return pfail(VH_ERR_INVALID, "vh_pipeline_depthf: bad argument / no frame yet");
    *d_depthf = p->depthf[(int)((p->frame - 1) % kMapSets)];
    return VH_OK;
}

// which: 0 = maps of the latest pushed frame, 1 = ICP target of the NEXT frame's tracking
int vh_pipeline_maps(vh_pipeline* p, int which, float4** verts, float4** normals) {
    if (!p || !verts || !normals || p->frame == 0) return pfail(VH_ERR_INVALID, "vh_pipeline_maps: bad argument / no frame yet");
    const int last = (int)((p->frame - 1) % kMapSets);
    if (which == 1 && p->mode == VH_TRACK_FRAME_TO_MODEL) { *verts = p->modelVerts; *normals = p->modelNormals; }
    else { *verts = p->verts[last]; *normals = p->normals[last]; }
    return VH_OK;
}

long long vh_pipeline_launches(vh_pipeline* p) { return p ? p->launches : 0; }

}  // extern "C"
