// vh_abi.cu -- the extern "C" surface of libvh_b200.so (include/vh/abi.h): context management,
// the stream-ordered handle API and the reference's legacy entry points on a global context.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>

#include "vh_internal.h"

using namespace vh;

namespace {

thread_local std::string g_lastError;

int fail(int code, const char* what, cudaError_t e = cudaSuccess) {
    g_lastError = what;
    if (e != cudaSuccess) { g_lastError += ": "; g_lastError += cudaGetErrorString(e); }
    return code;
}
#define VH_CUDA(expr)                                                        \
    do {                                                                     \
        cudaError_t _e = (expr);                                             \
        if (_e != cudaSuccess) return fail(VH_ERR_CUDA, #expr, _e);          \
    } while (0)

cudaStream_t S(vh_stream s) { return reinterpret_cast<cudaStream_t>(s); }

template <class T>
cudaError_t devAlloc(vh_context* c, T** p, size_t n) {
    size_t bytes = n * sizeof(T);
    if (bytes == 0) bytes = sizeof(T);
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), bytes);
    if (e == cudaSuccess) c->bytesAllocated += bytes;
    return e;
}

// View constants derived from the config; every derived float is computed here, once, in fp32 in
// a fixed order (DESIGN.md "derived constants"), and mirrored by the oracle.
void fillView(vh_context* c) {
    const vh_config& g = c->cfg;
    View& v = c->v;
    v.numBuckets = g.table.numBuckets;
    v.bucketSize = g.table.bucketSize;
    v.chainMax = g.policy == VH_POLICY_FIXED ? g.table.attachedLinkedListSize : 0;
    v.numVoxelBlocks = g.table.numVoxelBlocks;
    v.numSlots = g.table.numBuckets * g.table.bucketSize;
    v.overflowSlots = g.policy == VH_POLICY_FIXED ? g.overflowSlots : 0;
    v.voxelSize = g.table.voxelSize;
    v.invVoxelSize = 1.0f / g.table.voxelSize;
    v.truncation = g.table.truncation;
    v.truncScale = g.table.truncScale;
    v.wMax = g.table.integrationWeightMax;
    v.wSample = (float)g.table.integrationWeightSample;
    v.depthMin = g.depthMin;
    v.depthMax = g.depthMax;
    v.invDepthRange = 1.0f / (g.depthMax - g.depthMin);
    v.depthScale = g.depthScale;
    // Niessner's weight max(wSample * 1.5 * (1 - (d - dmin)/(dmax - dmin)), 1) as one FMA in d (ref VoxelUtils.cu:809-827)
    v.wA = -((v.wSample * 1.5f) * v.invDepthRange);
    v.wB = (v.wSample * 1.5f) * (1.0f + g.depthMin * v.invDepthRange);
    v.zFar = fmaf(v.truncScale, v.depthMax, v.truncation) + v.depthMax;
    v.W = g.width;
    v.H = g.height;
    v.fx = g.fx; v.fy = g.fy; v.cx = g.cx; v.cy = g.cy;
    const float K[9] = {g.fx, 0.f, g.cx, 0.f, g.fy, g.cy, 0.f, 0.f, 1.f};
    // closed-form inverse of the pinhole matrix (the reference gets it from Eigen, CameraTracking.cpp:131-134)
    const float Ki[9] = {1.0f / g.fx, 0.f, -g.cx / g.fx, 0.f, 1.0f / g.fy, -g.cy / g.fy, 0.f, 0.f, 1.f};
    if (!c->intrinsicsSet) { memcpy(v.K, K, sizeof(K)); memcpy(v.Kinv, Ki, sizeof(Ki)); }
    v.wr = (float)(g.width - 1) - g.cx;
    v.hb = (float)(g.height - 1) - g.cy;
    v.nl = sqrtf(g.fx * g.fx + g.cx * g.cx);
    v.nr = sqrtf(g.fx * g.fx + v.wr * v.wr);
    v.nt = sqrtf(g.fy * g.fy + g.cy * g.cy);
    v.nb = sqrtf(g.fy * g.fy + v.hb * v.hb);
    v.rad = g.table.voxelSize * 6.9282032f;
    v.partCount = g.partCount < 1 ? 1 : g.partCount;
    v.partRank = g.partRank;
    v.icpDistThres = g.icpDistThres;
    v.icpNormalThres = g.icpNormalThres;
}

}  // namespace

extern "C" {

const char* vh_last_error(void) { return g_lastError.c_str(); }

// reference defaults, common.h:7-50
void vh_default_config(vh_config* cfg) {
    memset(cfg, 0, sizeof(*cfg));
    HashTableParams& t = cfg->table;
    for (int i = 0; i < 16; ++i) {
        t.global_transform.entries[i] = (i % 5 == 0) ? 1.f : 0.f;
        t.inv_global_transform.entries[i] = (i % 5 == 0) ? 1.f : 0.f;
    }
    t.numBuckets = 5000; t.bucketSize = 5; t.attachedLinkedListSize = 4; t.numVoxelBlocks = 1000;
    t.voxelBlockSize = 8; t.voxelSize = 0.02f; t.numOccupiedBlocks = 0;
    t.maxIntegrationDistance = 4.0f; t.truncScale = 0.01f; t.truncation = 1.0f;
    t.integrationWeightSample = 10; t.integrationWeightMax = 255.f;
    cfg->policy = VH_POLICY_REF_EXACT;
    cfg->width = 640; cfg->height = 480;
    cfg->fx = 517.3f; cfg->fy = 516.5f; cfg->cx = 318.6f; cfg->cy = 255.3f;
    cfg->depthScale = 5000.f;
    cfg->depthMin = 0.1f; cfg->depthMax = 4.0f;
    cfg->overflowSlots = 0;
    cfg->icpDistThres = 0.08f; cfg->icpNormalThres = -1.0f; cfg->icpIterations = 20;
    cfg->partCount = 1; cfg->partRank = 0;
    cfg->bilateralSigmaSpace = 0.0f; cfg->bilateralSigmaRange = 0.0f;
}

int vh_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int vh_create(const vh_config* cfg, vh_context** out) {
    if (!cfg || !out) return fail(VH_ERR_INVALID, "vh_create: null argument");
    *out = nullptr;
    if (cfg->table.numBuckets == 0 || cfg->table.bucketSize == 0 || cfg->table.numVoxelBlocks == 0 ||
        cfg->table.voxelBlockSize != VH_BLOCK_SIDE || !(cfg->table.voxelSize > 0.f) || cfg->width <= 0 || cfg->height <= 0)
        return fail(VH_ERR_INVALID, "vh_create: bad table/image parameters (voxelBlockSize must be 8)");
    if ((unsigned long long)cfg->table.numVoxelBlocks * 512ull >= 0x7fffffffull)
        return fail(VH_ERR_INVALID, "vh_create: numVoxelBlocks*512 must fit the reference's int ptr");
    if (cfg->partCount > 1 && (cfg->partRank < 0 || cfg->partRank >= cfg->partCount))
        return fail(VH_ERR_INVALID, "vh_create: partRank must be in [0, partCount)");
    if (vh_device_count() == 0) return fail(VH_ERR_NO_DEVICE, "vh_create: no CUDA device (this library has no CPU fallback)");
    vh_context* c = new vh_context();
    memset(static_cast<void*>(c), 0, sizeof(*c));
    c->cfg = *cfg;
    c->peers.world = 1;
    if (c->cfg.policy == VH_POLICY_FIXED && c->cfg.overflowSlots == 0) c->cfg.overflowSlots = c->cfg.table.numBuckets;
    if (c->cfg.partCount < 1) c->cfg.partCount = 1;
    cudaGetDevice(&c->device);
    cudaDeviceGetAttribute(&c->numSMs, cudaDevAttrMultiProcessorCount, c->device);
    fillView(c);
    View& v = c->v;
    const size_t slots = (size_t)v.numSlots + v.overflowSlots;
    const size_t N = v.numVoxelBlocks;
    cudaError_t e = cudaSuccess;
    auto chk = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
    chk(devAlloc(c, &v.entries, slots));
    chk(devAlloc(c, &v.chain, slots));
    chk(devAlloc(c, &v.mutex, (size_t)v.numBuckets));
    chk(devAlloc(c, &v.heap, N));
    chk(devAlloc(c, &v.blockInfo, N));
    chk(devAlloc(c, &v.voxels, N * 512));
    chk(devAlloc(c, &v.compact16, N));
    chk(devAlloc(c, &v.compact20, N));
    chk(devAlloc(c, &v.ctr, (size_t)1));
    if (e == cudaSuccess) chk(cudaMemset(v.ctr, 0, sizeof(Counters)));   // k_reset carries icpSeq over: it must start defined
    chk(devAlloc(c, &c->frame, (size_t)1));
    if (c->cfg.policy == VH_POLICY_FIXED && c->cfg.bilateralSigmaSpace > 0.0f && c->cfg.bilateralSigmaRange > 0.0f) {
        // Host-computed tables (the oracle computes the same doubles with the same libm): spatial weights for offsets
        // 0, 1, 2 and the range weight per raw-unit depth difference.
        float* lut = nullptr;
        chk(devAlloc(c, &lut, (size_t)kBilatLut));
        chk(devAlloc(c, &v.depthSmooth, (size_t)v.W * v.H));
        std::vector<float> h(kBilatLut);
        const double sr = (double)(c->cfg.bilateralSigmaRange * c->cfg.depthScale), ss = (double)c->cfg.bilateralSigmaSpace;
        for (int i = 0; i < kBilatLut; ++i) h[i] = (float)std::exp(-((double)i * (double)i) / (2.0 * sr * sr));
        for (int i = 0; i < 3; ++i) v.bilatG[i] = (float)std::exp(-((double)i * (double)i) / (2.0 * ss * ss));
        if (e == cudaSuccess) chk(cudaMemcpy(lut, h.data(), sizeof(float) * kBilatLut, cudaMemcpyHostToDevice));
        v.bilatLut = lut;
    }
    {   // IcpState followed by the fp64 solver state
        void* p = nullptr;
        cudaError_t r = cudaMalloc(&p, sizeof(IcpState) + 16 * sizeof(double));
        chk(r);
        c->icp = static_cast<IcpState*>(p);
    }
    chk(devAlloc(c, &c->icpPartials, (size_t)kIcpMaxBlocks * 32));
    chk(devAlloc(c, &c->icpLL, kIcpLLWords));
    if (e == cudaSuccess) chk(cudaMemset(c->icpLL, 0, sizeof(unsigned long long) * kIcpLLWords));   // sequence 0 = never written
    // persistent Align grid: a few SMs stay free so that the pre-processing of the next frame and the fusion of the previous
    // one (other streams) can run beside the tracking of this one.  r2 sweep at VGA, frames/s device-resident / end to end:
    // 148: 7 770 / -, 144: 8 840 / 8 250, 140: 8 800 / 8 620, 136: 9 230 / 9 240, 132: 8 600 / 9 210, 120: 8 430 / 8 990
    c->icpCtas = c->numSMs > 32 ? c->numSMs - 12 : c->numSMs;
    if (const char* env = getenv("VH_ICP_CTAS")) c->icpCtas = atoi(env);
    if (const char* env = getenv("VH_FUSION_RESERVE_SMS")) c->fusionReserveSMs = atoi(env);
#ifdef VH_TIMELINE
    if (const char* env = getenv("VH_TIMELINE")) if (env[0] == '1') {
        chk(devAlloc(c, &v.tl, (size_t)kTimelineCap + 1));
        if (e == cudaSuccess) chk(cudaMemset(v.tl, 0, sizeof(unsigned long long) * (kTimelineCap + 1)));
    }
#endif
    {
        const size_t tiles = (size_t)((cfg->width + 7) / 8) * ((cfg->height + 7) / 8);
        chk(devAlloc(c, &c->tileMin, tiles));
        chk(devAlloc(c, &c->tileMax, tiles));
    }
    if (e != cudaSuccess) { vh_destroy(c); return fail(VH_ERR_CUDA, "vh_create: cudaMalloc", e); }
    v.frame = c->frame;
    float ident[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    chk(launch_reset(c, 0));
    chk(cudaMemsetAsync(c->icp, 0, sizeof(IcpState) + 16 * sizeof(double), 0));
    chk(launch_set_frame_host(c, ident, 0));
    chk(launch_icp_reset(c, true, 0));
    chk(cudaStreamSynchronize(0));
    if (e != cudaSuccess) { vh_destroy(c); return fail(VH_ERR_CUDA, "vh_create: initialisation", e); }
    *out = c;
    return VH_OK;
}

void vh_destroy(vh_context* c) {
    if (!c) return;
    cudaFree(c->v.entries); cudaFree(c->v.chain); cudaFree(c->v.mutex); cudaFree(c->v.heap);
    cudaFree(c->v.blockInfo); cudaFree(c->v.voxels); cudaFree(c->v.compact16); cudaFree(c->v.compact20);
    cudaFree(const_cast<float*>(c->v.bilatLut)); cudaFree(c->v.depthSmooth); cudaFree(c->v.tl);
    cudaFree(c->v.ctr); cudaFree(c->frame); cudaFree(c->icp); cudaFree(c->icpPartials); cudaFree(c->icpLL); cudaFree(c->tileMin); cudaFree(c->tileMax);
    delete c;
}

int vh_reset(vh_context* c, vh_stream s) {
    if (!c) return fail(VH_ERR_INVALID, "vh_reset: null context");
    float ident[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    VH_CUDA(launch_reset(c, S(s)));
    VH_CUDA(launch_set_frame_host(c, ident, S(s)));
    VH_CUDA(launch_icp_reset(c, true, S(s)));
    return VH_OK;
}

int vh_get_config(const vh_context* c, vh_config* out) {
    if (!c || !out) return fail(VH_ERR_INVALID, "vh_get_config: null argument");
    *out = c->cfg;
    return VH_OK;
}

// Scheduling knobs of the overlapped frame loop (vh_pipeline, VH_PIPE_OVERLAP): the Align grid (one 512-thread CTA per
// SM, the whole register file of each) and the persistent integrate grid cannot share an SM, so they overlap only if each
// leaves the other whole SMs.  align_ctas <= 0 / fusion_reserved_sms < 0 keep the current value.
int vh_set_tuning(vh_context* c, int align_ctas, int fusion_reserved_sms) {
    if (!c) return fail(VH_ERR_INVALID, "vh_set_tuning: null context");
    if (align_ctas > c->numSMs || fusion_reserved_sms >= c->numSMs) return fail(VH_ERR_INVALID, "vh_set_tuning: more SMs than the device has");
    if (align_ctas > 0) c->icpCtas = align_ctas;
    if (fusion_reserved_sms >= 0) c->fusionReserveSMs = fusion_reserved_sms;
    return VH_OK;
}

int vh_set_intrinsics(vh_context* c, float fx, float fy, float cx, float cy) {
    if (!c) return fail(VH_ERR_INVALID, "vh_set_intrinsics: null context");
    c->cfg.fx = fx; c->cfg.fy = fy; c->cfg.cx = cx; c->cfg.cy = cy;
    c->intrinsicsSet = false;
    fillView(c);
    return VH_OK;
}

// SetCameraIntrinsic semantics on a context: K and K^-1 exactly as handed in (row-major reads)
int vh_set_intrinsic_matrices(vh_context* c, const float* K9, const float* Kinv9) {
    if (!c || !K9 || !Kinv9) return fail(VH_ERR_INVALID, "vh_set_intrinsic_matrices: null argument");
    memcpy(c->v.K, K9, 9 * sizeof(float));
    memcpy(c->v.Kinv, Kinv9, 9 * sizeof(float));
    c->intrinsicsSet = true;
    return VH_OK;
}

int vh_preprocess(vh_context* c, const uint16_t* d_depth, float4* d_verts, float4* d_normals, float* d_depthf, vh_stream s) {
    if (!c || !d_depth || !d_verts || !d_normals) return fail(VH_ERR_INVALID, "vh_preprocess: null argument");
    VH_CUDA(launch_preprocess(c, d_depth, d_verts, d_normals, d_depthf, S(s)));
    return VH_OK;
}

int vh_set_pose(vh_context* c, const float* pose, vh_stream s) {
    if (!c || !pose) return fail(VH_ERR_INVALID, "vh_set_pose: null argument");
    VH_CUDA(launch_set_frame_host(c, pose, S(s)));
    return VH_OK;
}
int vh_set_pose_device(vh_context* c, const float* d_pose, vh_stream s) {
    if (!c || !d_pose) return fail(VH_ERR_INVALID, "vh_set_pose_device: null argument");
    VH_CUDA(launch_set_frame_device(c, d_pose, nullptr, nullptr, S(s)));
    return VH_OK;
}
int vh_alloc_blocks(vh_context* c, const float4* d_verts, const float4* /*d_normals*/, vh_stream s) {
    if (!c || !d_verts) return fail(VH_ERR_INVALID, "vh_alloc_blocks: null argument");
    if (c->cfg.policy == VH_POLICY_REF_EXACT) VH_CUDA(launch_reset_mutex(c, S(s)));   // SDF_Hashtable.cpp:24
    VH_CUDA(launch_alloc(c, d_verts, S(s)));
    return VH_OK;
}
// Allocation straight from the raw u16 depth image: the back-projection of vh_preprocess happens in registers
// (same operations, same order => the same blocks), 2 B per pixel instead of the 16 B of the vertex map.
int vh_alloc_blocks_depth(vh_context* c, const uint16_t* d_depth, vh_stream s) {
    if (!c || !d_depth) return fail(VH_ERR_INVALID, "vh_alloc_blocks_depth: null argument");
    if (c->v.bilatLut != nullptr) return fail(VH_ERR_UNSUPPORTED, "vh_alloc_blocks_depth: the bilateral front end allocates from the filtered vertex map");
    if (c->cfg.policy == VH_POLICY_REF_EXACT) VH_CUDA(launch_reset_mutex(c, S(s)));   // SDF_Hashtable.cpp:24
    VH_CUDA(launch_alloc_depth16(c, d_depth, S(s)));
    return VH_OK;
}
int vh_compact(vh_context* c, vh_stream s) {
    if (!c) return fail(VH_ERR_INVALID, "vh_compact: null context");
    VH_CUDA(cudaMemsetAsync(&c->v.ctr->compactCount, 0, sizeof(int), S(s)));
    VH_CUDA(launch_compact(c, S(s)));
    return VH_OK;
}
int vh_integrate(vh_context* c, const float4* d_verts, vh_stream s) {
    if (!c || !d_verts) return fail(VH_ERR_INVALID, "vh_integrate: null argument");
    VH_CUDA(cudaMemsetAsync(&c->v.ctr->numUpdated, 0, sizeof(unsigned long long), S(s)));
    VH_CUDA(launch_integrate(c, d_verts, nullptr, -1, S(s)));
    return VH_OK;
}
int vh_integrate_depthf(vh_context* c, const float* d_depthf, vh_stream s) {
    if (!c || !d_depthf) return fail(VH_ERR_INVALID, "vh_integrate_depthf: null argument");
    VH_CUDA(cudaMemsetAsync(&c->v.ctr->numUpdated, 0, sizeof(unsigned long long), S(s)));
    VH_CUDA(launch_integrate(c, nullptr, d_depthf, -1, S(s)));
    return VH_OK;
}
// set_pose (which clears the per-frame counters) must precede this call
int vh_fuse_frame(vh_context* c, const float4* d_verts, const float4* /*d_normals*/, const float* d_depthf, vh_stream s) {
    if (!c || !d_verts) return fail(VH_ERR_INVALID, "vh_fuse_frame: null argument");
    if (c->cfg.policy == VH_POLICY_REF_EXACT) VH_CUDA(launch_reset_mutex(c, S(s)));
    VH_CUDA(launch_alloc(c, d_verts, S(s)));
    VH_CUDA(launch_compact(c, S(s)));
    VH_CUDA(launch_integrate(c, d_verts, d_depthf, -1, S(s)));
    return VH_OK;
}

int vh_get_stats(vh_context* c, vh_stats* out, vh_stream s) {
    if (!c || !out) return fail(VH_ERR_INVALID, "vh_get_stats: null argument");
    Counters h;
    VH_CUDA(cudaMemcpyAsync(&h, c->v.ctr, sizeof(h), cudaMemcpyDeviceToHost, S(s)));
    VH_CUDA(cudaStreamSynchronize(S(s)));
    out->heapCounter = h.heapCounter;
    out->numAllocated = (int)c->v.numVoxelBlocks - 1 - h.heapCounter;
    out->numVisible = h.compactCount;
    out->overflowUsed = h.overflowUsed;
    out->dropped = h.dropped;
    out->numUpdated = h.numUpdated;
    out->lastInserted = h.lastInserted;
    out->lastFreed = h.gcFreed;
    out->overflowLeaked = h.arenaLeaked;
    out->exchangeTimeouts = h.exchangeTimeouts;
    return VH_OK;
}

int vh_garbage_collect(vh_context* c, int scope, float sdf_threshold, float weight_decay, vh_stream s) {
    if (!c) return fail(VH_ERR_INVALID, "vh_garbage_collect: null context");
    if (scope != VH_GC_VISIBLE && scope != VH_GC_ALL) return fail(VH_ERR_INVALID, "vh_garbage_collect: unknown scope");
    if (c->cfg.policy != VH_POLICY_FIXED)
        return fail(VH_ERR_UNSUPPORTED, "vh_garbage_collect: the RefExact table has no removal (ref deleteVoxelEntry is dead code)");
    if (!(sdf_threshold > 0.0f)) sdf_threshold = fmaf(c->v.truncScale, c->v.depthMax, c->v.truncation);
    VH_CUDA(launch_gc(c, scope, sdf_threshold, weight_decay, S(s)));
    return VH_OK;
}

static int readStreamCount(vh_context* c, int* h_count, vh_stream s) {
    Counters h;
    VH_CUDA(cudaMemcpyAsync(&h, c->v.ctr, sizeof(h), cudaMemcpyDeviceToHost, S(s)));
    VH_CUDA(cudaStreamSynchronize(S(s)));
    if (h_count) *h_count = h.streamCount;
    return VH_OK;
}
int vh_stream_out(vh_context* c, const float* center, float radius, VoxelEntry* entries_out, Voxel* voxels_out, int capacity,
                  int* h_count, vh_stream s) {
    if (!c || !center || !h_count || capacity < 0 || (capacity > 0 && (!entries_out || !voxels_out)))
        return fail(VH_ERR_INVALID, "vh_stream_out: bad argument");
    if (c->cfg.policy != VH_POLICY_FIXED) return fail(VH_ERR_UNSUPPORTED, "vh_stream_out: the RefExact table has no removal");
    if (!(radius >= 0.0f)) return fail(VH_ERR_INVALID, "vh_stream_out: radius must be >= 0");
    VH_CUDA(launch_stream_out(c, center, radius, entries_out, voxels_out, capacity, S(s)));
    return readStreamCount(c, h_count, s);
}
int vh_stream_in(vh_context* c, const VoxelEntry* entries, const Voxel* voxels, int count, int* h_count, vh_stream s) {
    if (!c || count < 0 || (count > 0 && (!entries || !voxels))) return fail(VH_ERR_INVALID, "vh_stream_in: bad argument");
    if (c->cfg.policy != VH_POLICY_FIXED) return fail(VH_ERR_UNSUPPORTED, "vh_stream_in: the RefExact table has no removal");
    VH_CUDA(launch_stream_in(c, entries, voxels, count, S(s)));
    VH_CUDA(cudaMemsetAsync(&c->v.ctr->compactCount, 0, sizeof(int), S(s)));
    return readStreamCount(c, h_count, s);
}

// ---- tracking -------------------------------------------------------------------------------------
int vh_icp_reset(vh_context* c, int reset_estimate, vh_stream s) {
    if (!c) return fail(VH_ERR_INVALID, "vh_icp_reset: null context");
    VH_CUDA(launch_icp_reset(c, reset_estimate != 0, S(s)));
    return VH_OK;
}
int vh_icp_iterate(vh_context* c, const float4* in, const float4* inN, const float4* tg, const float4* tgN, vh_stream s) {
    if (!c || !in || !tg || !tgN) return fail(VH_ERR_INVALID, "vh_icp_iterate: null argument");
    VH_CUDA(launch_icp_iter_ex(c, in, inN, tg, tgN, 0, c->v.H, nullptr, true, false, false, S(s)));
    return VH_OK;
}
int vh_icp_align(vh_context* c, const float4* in, const float4* inN, const float4* tg, const float4* tgN, int iterations,
                 vh_stream s) {
    if (!c || !in || !tg || !tgN) return fail(VH_ERR_INVALID, "vh_icp_align: null argument");
    if (iterations <= 0) iterations = c->cfg.icpIterations;
    VH_CUDA(launch_icp_align(c, nullptr, nullptr, in, inN, tg, tgN, 0, c->v.H, iterations, false, nullptr, nullptr, S(s)));   // CameraTracking.cpp:35-67, one launch
    return VH_OK;
}
// Pre-processing + Align + pose chain of one frame in ONE launch (k_track.cu).
int vh_track_frame(vh_context* c, const uint16_t* d_depth, float4* d_verts, float4* d_normals, float* d_depthf, const float4* tg,
                   const float4* tgN, int iterations, const float* d_pose_in, float* d_pose_out, vh_stream s) {
    if (!c || !d_depth || !d_verts || !d_normals || !d_depthf || !tg || !tgN) return fail(VH_ERR_INVALID, "vh_track_frame: null argument");
    if ((d_pose_in == nullptr) != (d_pose_out == nullptr)) return fail(VH_ERR_INVALID, "vh_track_frame: pose_in and pose_out go together");
    if (c->v.bilatLut != nullptr) return fail(VH_ERR_UNSUPPORTED, "vh_track_frame: the bilateral front end needs its own pass (vh_preprocess + vh_icp_align)");
    if (iterations <= 0) iterations = c->cfg.icpIterations;
    const int world = c->peers.world > 1 ? c->peers.world : 1, rank = world > 1 ? c->peers.rank : 0;
    const int base = c->v.H / world, rem = c->v.H % world;
    const int row0 = rank * base + (rank < rem ? rank : rem), row1 = row0 + base + (rank < rem ? 1 : 0);
    VH_CUDA(launch_icp_align(c, d_depth, d_depthf, d_verts, d_normals, tg, tgN, row0, row1, iterations, world > 1, d_pose_in, d_pose_out, S(s)));
    return VH_OK;
}
int vh_icp_reduce(vh_context* c, const float4* in, const float4* inN, const float4* tg, const float4* tgN, int row0, int row1,
                  vh_icp_system* d_system, vh_stream s) {
    if (!c || !in || !tg || !tgN || !d_system) return fail(VH_ERR_INVALID, "vh_icp_reduce: null argument");
    if (row0 < 0 || row1 > c->v.H || row0 > row1) return fail(VH_ERR_INVALID, "vh_icp_reduce: bad row range");
    VH_CUDA(launch_icp_iter_ex(c, in, inN, tg, tgN, row0, row1, d_system, false, true, false, S(s)));
    return VH_OK;
}
// Multi-GPU: peer-mapped exchange regions (one per rank, VH_PEER_BYTES each, zero-initialised), e.g. from
// torch.distributed._symmetric_memory or cudaIpc.  bufs[p] = rank p's region as visible from THIS process.
int vh_set_peers(vh_context* c, int rank, int world, void* const* bufs) {
    if (!c || world < 1 || world > kMaxPeers || rank < 0 || rank >= world || (world > 1 && !bufs))
        return fail(VH_ERR_INVALID, "vh_set_peers: bad argument (world <= 8)");
    c->peers.world = world;
    c->peers.rank = rank;
    for (int p = 0; p < kMaxPeers; ++p) c->peers.buf[p] = (p < world && bufs) ? static_cast<float*>(bufs[p]) : nullptr;
    c->peers.timeouts = &c->v.ctr->exchangeTimeouts;
    return VH_OK;
}
unsigned long long vh_peer_bytes(void) { return (unsigned long long)kPeerBytes; }

// Align over this rank's rows [row0,row1) with the cross-GPU all-reduce fused into every iteration's kernel.
int vh_icp_align_rows(vh_context* c, const float4* in, const float4* inN, const float4* tg, const float4* tgN, int row0, int row1,
                      int iterations, vh_stream s) {
    if (!c || !in || !tg || !tgN) return fail(VH_ERR_INVALID, "vh_icp_align_rows: null argument");
    if (row0 < 0 || row1 > c->v.H || row0 > row1) return fail(VH_ERR_INVALID, "vh_icp_align_rows: bad row range");
    if (iterations <= 0) iterations = c->cfg.icpIterations;
    VH_CUDA(launch_icp_align(c, nullptr, nullptr, in, inN, tg, tgN, row0, row1, iterations, c->peers.world > 1, nullptr, nullptr, S(s)));
    return VH_OK;
}

int vh_icp_solve(vh_context* c, const vh_icp_system* d_system, vh_stream s) {
    if (!c || !d_system) return fail(VH_ERR_INVALID, "vh_icp_solve: null argument");
    VH_CUDA(launch_icp_solve(c, d_system, S(s)));
    return VH_OK;
}
int vh_icp_get(vh_context* c, float* delta, float* twist6, vh_icp_system* last, vh_stream s) {
    if (!c) return fail(VH_ERR_INVALID, "vh_icp_get: null context");
    IcpState h;
    if (twist6) VH_CUDA(launch_icp_twist(c, S(s)));
    VH_CUDA(cudaMemcpyAsync(&h, c->icp, sizeof(h), cudaMemcpyDeviceToHost, S(s)));
    VH_CUDA(cudaStreamSynchronize(S(s)));
    if (delta) memcpy(delta, h.delta, sizeof(h.delta));
    if (twist6) memcpy(twist6, h.twist, sizeof(h.twist));
    if (last) memcpy(last, h.system, sizeof(h.system));
    return VH_OK;
}
int vh_icp_set_delta(vh_context* c, const float* twist6, vh_stream s) {
    if (!c || !twist6) return fail(VH_ERR_INVALID, "vh_icp_set_delta: null argument");
    VH_CUDA(launch_icp_set_twist(c, twist6, S(s)));
    return VH_OK;
}
const float* vh_icp_delta_device(vh_context* c) { return c ? c->icp->delta : nullptr; }
int vh_pose_compose(vh_context* c, const float* d_pose_in, float* d_pose_out, vh_stream s) {
    if (!c || !d_pose_in || !d_pose_out) return fail(VH_ERR_INVALID, "vh_pose_compose: null argument");
    VH_CUDA(launch_set_frame_device(c, d_pose_in, c->icp->delta, d_pose_out, S(s)));
    return VH_OK;
}
int vh_icp_reduce_corr(vh_context* c, const float4* corr, const float4* corrN, const float* res, vh_icp_system* d_system,
                       vh_stream s) {
    if (!c || !corr || !corrN || !res || !d_system) return fail(VH_ERR_INVALID, "vh_icp_reduce_corr: null argument");
    VH_CUDA(launch_reduce_corr(c, corr, corrN, res, d_system, S(s)));
    return VH_OK;
}
int vh_find_correspondences(vh_context* c, const float4* in, const float4* inN, const float4* tg, const float4* tgN,
                            const float* delta16, float4* corr, float4* corrN, float* res, float* d_err, vh_stream s) {
    if (!c || !in || !tg || !tgN || !delta16 || !corr || !corrN || !res || !d_err)
        return fail(VH_ERR_INVALID, "vh_find_correspondences: null argument");
    VH_CUDA(launch_find_corr(c, in, inN, tg, tgN, delta16, corr, corrN, res, d_err, S(s)));
    return VH_OK;
}
int vh_jacobians(vh_context* c, const float4* corr, const float4* corrN, float* J, vh_stream s) {
    if (!c || !corr || !corrN || !J) return fail(VH_ERR_INVALID, "vh_jacobians: null argument");
    VH_CUDA(launch_jacobians(c, corr, corrN, J, S(s)));
    return VH_OK;
}

int vh_raycast(vh_context* c, float4* d_verts, float4* d_normals, vh_stream s) {
    if (!c || !d_verts || !d_normals) return fail(VH_ERR_INVALID, "vh_raycast: null argument");
    if (c->cfg.policy != VH_POLICY_FIXED) return fail(VH_ERR_INVALID, "vh_raycast: Fixed policy only (the RefExact TSDF is not a surface)");
    if (c->v.partCount > 1)      // a rank holds ~1/P of the blocks: trilinear samples and gradients fail at every face owned elsewhere
        return fail(VH_ERR_UNSUPPORTED, "vh_raycast: not defined on a partitioned context (no halo exchange); raycast the merged model");
    VH_CUDA(cudaMemsetAsync(&c->v.ctr->compactCount, 0, sizeof(int), S(s)));
    VH_CUDA(launch_compact(c, S(s)));                    // the visible list of the CURRENT pose feeds the ray intervals
    VH_CUDA(launch_raycast(c, d_verts, d_normals, S(s)));
    return VH_OK;
}

// ---- export ---------------------------------------------------------------------------------------
int vh_export_entries(vh_context* c, VoxelEntry* h_entries, int cap, int* count) {
    if (!c || !count) return fail(VH_ERR_INVALID, "vh_export_entries: null argument");
    const size_t slots = (size_t)c->v.numSlots + c->v.overflowSlots;
    VoxelEntry* d_out = nullptr;
    int* d_count = nullptr;
    VH_CUDA(cudaMalloc((void**)&d_out, sizeof(VoxelEntry) * slots));
    VH_CUDA(cudaMalloc((void**)&d_count, sizeof(int)));
    cudaError_t e = launch_export_entries(c, d_out, d_count, 0);
    int n = 0;
    if (e == cudaSuccess) e = cudaMemcpy(&n, d_count, sizeof(int), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && h_entries && cap > 0)
        e = cudaMemcpy(h_entries, d_out, sizeof(VoxelEntry) * (size_t)(n < cap ? n : cap), cudaMemcpyDeviceToHost);
    cudaFree(d_out); cudaFree(d_count);
    if (e != cudaSuccess) return fail(VH_ERR_CUDA, "vh_export_entries", e);
    *count = n;
    return VH_OK;
}
int vh_export_compact(vh_context* c, VoxelEntry* h_entries, int cap, int* count) {
    if (!c || !count) return fail(VH_ERR_INVALID, "vh_export_compact: null argument");
    int n = 0;
    VH_CUDA(cudaDeviceSynchronize());
    VH_CUDA(cudaMemcpy(&n, &c->v.ctr->compactCount, sizeof(int), cudaMemcpyDeviceToHost));
    if (h_entries && cap > 0)
        VH_CUDA(cudaMemcpy(h_entries, c->v.compact20, sizeof(VoxelEntry) * (size_t)(n < cap ? n : cap), cudaMemcpyDeviceToHost));
    *count = n;
    return VH_OK;
}
int vh_export_block(vh_context* c, int ptr, Voxel* h512) {
    if (!c || !h512 || ptr < 0 || (unsigned)ptr / 512u >= c->v.numVoxelBlocks) return fail(VH_ERR_INVALID, "vh_export_block: bad argument");
    VH_CUDA(cudaDeviceSynchronize());
    VH_CUDA(cudaMemcpy(h512, c->v.voxels + ptr, sizeof(Voxel) * 512, cudaMemcpyDeviceToHost));
    return VH_OK;
}
const VoxelEntry* vh_compact_table_device(vh_context* c) { return c ? c->v.compact20 : nullptr; }
const int* vh_compact_counter_device(vh_context* c) { return c ? &c->v.ctr->compactCount : nullptr; }
Voxel* vh_voxel_blocks_device(vh_context* c) { return c ? c->v.voxels : nullptr; }
unsigned long long vh_bytes_allocated(vh_context* c) { return c ? (unsigned long long)c->bytesAllocated : 0ull; }

// Binary checkpoint: header, counters, entries, chain, heap, blockInfo, then the ALLOCATED voxel
// blocks only (ids above heapCounter).  SURVEY.md section 8 f4.
namespace {
struct CkptHeader { char magic[8]; unsigned version, numBuckets, bucketSize, numVoxelBlocks, overflowSlots, policy; float voxelSize; };
}
int vh_save(vh_context* c, const char* path) {
    if (!c || !path) return fail(VH_ERR_INVALID, "vh_save: null argument");
    VH_CUDA(cudaDeviceSynchronize());
    FILE* f = fopen(path, "wb");
    if (!f) return fail(VH_ERR_INVALID, "vh_save: cannot open file");
    CkptHeader h{{'V', 'H', 'B', '2', '0', '0', 0, 0}, 3, c->v.numBuckets, c->v.bucketSize, c->v.numVoxelBlocks, c->v.overflowSlots,
                 (unsigned)c->cfg.policy, c->v.voxelSize};
    Counters ctr;
    cudaMemcpy(&ctr, c->v.ctr, sizeof(ctr), cudaMemcpyDeviceToHost);
    const size_t slots = (size_t)c->v.numSlots + c->v.overflowSlots, N = c->v.numVoxelBlocks;
    std::vector<char> buf;
    auto put = [&](const void* d, size_t bytes) {
        buf.resize(bytes);
        cudaMemcpy(buf.data(), d, bytes, cudaMemcpyDeviceToHost);
        return fwrite(buf.data(), 1, bytes, f) == bytes;
    };
    bool ok = fwrite(&h, sizeof(h), 1, f) == 1 && fwrite(&ctr, sizeof(ctr), 1, f) == 1;
    ok = ok && put(c->v.entries, slots * sizeof(int4)) && put(c->v.chain, slots * sizeof(int)) &&
         put(c->v.heap, N * sizeof(unsigned)) && put(c->v.blockInfo, N * sizeof(int4));
    int first = std::min(ctr.heapLow, ctr.heapCounter) + 1;   // ids ever handed out (garbage collection leaves holes)
    if (first < 0) first = 0;
    if (ok && (size_t)first < N) ok = put(c->v.voxels + (size_t)first * 512, (N - first) * 512 * sizeof(Voxel));
    fclose(f);
    return ok ? VH_OK : fail(VH_ERR_INVALID, "vh_save: short write");
}
// The file is parsed and validated into host staging buffers first; the device is only touched once everything has
// been read, so a short or inconsistent file leaves the live table as it was.  Only the TABLE counters are restored:
// the ICP exchange sequence (icpSeq) belongs to the running context (the peers' mailboxes hold its numbers).
int vh_load(vh_context* c, const char* path) {
    if (!c || !path) return fail(VH_ERR_INVALID, "vh_load: null argument");
    FILE* f = fopen(path, "rb");
    if (!f) return fail(VH_ERR_INVALID, "vh_load: cannot open file");
    CkptHeader h;
    Counters ctr;
    bool ok = fread(&h, sizeof(h), 1, f) == 1 && fread(&ctr, sizeof(ctr), 1, f) == 1;
    if (!ok || memcmp(h.magic, "VHB200", 6) != 0 || h.version != 3 || h.numBuckets != c->v.numBuckets || h.bucketSize != c->v.bucketSize ||
        h.numVoxelBlocks != c->v.numVoxelBlocks || h.overflowSlots != c->v.overflowSlots) {
        fclose(f);
        return fail(VH_ERR_INVALID, "vh_load: checkpoint does not match this context's geometry");
    }
    if (h.policy != (unsigned)c->cfg.policy || h.voxelSize != c->v.voxelSize) {
        fclose(f);
        return fail(VH_ERR_INVALID, "vh_load: checkpoint was written with another arithmetic policy or voxel size");
    }
    const size_t slots = (size_t)c->v.numSlots + c->v.overflowSlots, N = c->v.numVoxelBlocks;
    // counters index heap[] and the overflow arena: range-check them before they reach the device
    if (ctr.heapCounter < -1 || ctr.heapCounter >= (int)N || ctr.heapLow < -1 || ctr.heapLow >= (int)N ||
        ctr.overflowUsed < 0 || (unsigned)ctr.overflowUsed > c->v.overflowSlots) {
        fclose(f);
        return fail(VH_ERR_INVALID, "vh_load: checkpoint counters out of range");
    }
    int first = std::min(ctr.heapLow, ctr.heapCounter) + 1;
    if (first < 0) first = 0;
    const size_t voxBytes = (size_t)first < N ? (N - first) * 512 * sizeof(Voxel) : 0;
    std::vector<char> entries(slots * sizeof(int4)), chain(slots * sizeof(int)), heap(N * sizeof(unsigned)), info(N * sizeof(int4)), vox;
    auto get = [&](std::vector<char>& b) { return b.empty() || fread(b.data(), 1, b.size(), f) == b.size(); };
    ok = get(entries) && get(chain) && get(heap) && get(info);
    if (ok) { vox.resize(voxBytes); ok = get(vox); }
    if (ok) ok = fgetc(f) == EOF;                            // trailing bytes: not a checkpoint of this geometry
    fclose(f);
    if (!ok) return fail(VH_ERR_INVALID, "vh_load: short or oversized file (the context was not modified)");
    const unsigned* hp = reinterpret_cast<const unsigned*>(heap.data());
    for (int i = 0; i <= ctr.heapCounter; ++i)
        if (hp[i] >= N) return fail(VH_ERR_INVALID, "vh_load: free list names a block id out of range (the context was not modified)");
    // commit
    VH_CUDA(cudaDeviceSynchronize());
    Counters live;
    VH_CUDA(cudaMemcpy(&live, c->v.ctr, sizeof(live), cudaMemcpyDeviceToHost));
    ctr.icpSeq = live.icpSeq;                                // the exchange sequence stays with the running context
    ctr.icpTicket = 0;
    ctr.icpConverged = 0;
    VH_CUDA(cudaMemcpy(c->v.entries, entries.data(), entries.size(), cudaMemcpyHostToDevice));
    VH_CUDA(cudaMemcpy(c->v.chain, chain.data(), chain.size(), cudaMemcpyHostToDevice));
    VH_CUDA(cudaMemcpy(c->v.heap, heap.data(), heap.size(), cudaMemcpyHostToDevice));
    VH_CUDA(cudaMemcpy(c->v.blockInfo, info.data(), info.size(), cudaMemcpyHostToDevice));
    VH_CUDA(cudaMemset(c->v.voxels, 0, N * 512 * sizeof(Voxel)));
    if (voxBytes) VH_CUDA(cudaMemcpy(c->v.voxels + (size_t)first * 512, vox.data(), voxBytes, cudaMemcpyHostToDevice));
    VH_CUDA(cudaMemcpy(c->v.ctr, &ctr, sizeof(ctr), cudaMemcpyHostToDevice));
    return VH_OK;
}
int vh_extract_mesh(vh_context* c, float* d_tris, int capacity, int* h_count, vh_stream s) {
    if (!c || !h_count || capacity < 0 || (capacity > 0 && !d_tris)) return fail(VH_ERR_INVALID, "vh_extract_mesh: bad argument");
    if (c->v.partCount > 1)      // cells on a block face need the neighbour block, which another rank may own
        return fail(VH_ERR_UNSUPPORTED, "vh_extract_mesh: not defined on a partitioned context (no halo exchange); extract from the merged model");
    VH_CUDA(launch_extract_mesh(c, d_tris, capacity, &c->v.ctr->meshCount, S(s)));
    VH_CUDA(cudaMemcpyAsync(h_count, &c->v.ctr->meshCount, sizeof(int), cudaMemcpyDeviceToHost, S(s)));
    VH_CUDA(cudaStreamSynchronize(S(s)));
    return VH_OK;
}
int vh_save_mesh_ply(const char* path, const float* tris, int count) {
    if (!path || count < 0 || (count > 0 && !tris)) return fail(VH_ERR_INVALID, "vh_save_mesh_ply: bad argument");
    FILE* f = fopen(path, "wb");
    if (!f) return fail(VH_ERR_INVALID, "vh_save_mesh_ply: cannot open file");
    fprintf(f, "ply\nformat binary_little_endian 1.0\ncomment libvh_b200 marching tetrahedra\nelement vertex %d\n"
               "property float x\nproperty float y\nproperty float z\nelement face %d\nproperty list uchar int vertex_indices\nend_header\n",
            count * 3, count);
    bool ok = count == 0 || fwrite(tris, sizeof(float) * 9, (size_t)count, f) == (size_t)count;
    for (int i = 0; ok && i < count; ++i) {
        const unsigned char n = 3;
        const int idx[3] = {3 * i, 3 * i + 1, 3 * i + 2};
        ok = fwrite(&n, 1, 1, f) == 1 && fwrite(idx, sizeof(int), 3, f) == 3;
    }
    fclose(f);
    return ok ? VH_OK : fail(VH_ERR_INVALID, "vh_save_mesh_ply: short write");
}
// The reference's text dump, SDFRenderer.cpp:90-108: count, then per visible entry pos/ptr/offset and 512 sdf values.
int vh_dump_text(vh_context* c, const char* path) {
    if (!c || !path) return fail(VH_ERR_INVALID, "vh_dump_text: null argument");
    int n = 0;
    int rc = vh_export_compact(c, nullptr, 0, &n);
    if (rc != VH_OK) return rc;
    std::vector<VoxelEntry> ent((size_t)n);
    if (n) { rc = vh_export_compact(c, ent.data(), n, &n); if (rc != VH_OK) return rc; }
    FILE* f = fopen(path, "w");
    if (!f) return fail(VH_ERR_INVALID, "vh_dump_text: cannot open file");
    fprintf(f, "numOccupiedBlocks from GL :%d\n", n);
    fprintf(f, "\nSDFs \n\n");
    std::vector<Voxel> vox(512);
    for (int i = 0; i < n; ++i) {
        fprintf(f, "%d) : pos : (%d, %d, %d) ptr = %d offset = %d\n", i, ent[i].pos.x, ent[i].pos.y, ent[i].pos.z, ent[i].ptr, ent[i].offset);
        cudaMemcpy(vox.data(), c->v.voxels + ent[i].ptr, sizeof(Voxel) * 512, cudaMemcpyDeviceToHost);
        for (int j = 0; j < 512; ++j) fprintf(f, "%.4f\t", vox[j].sdf);
        fprintf(f, "\n\n\n");
    }
    fclose(f);
    return VH_OK;
}

}  // extern "C"

// ====================================================================================================
// Legacy entry points (reference names) on one process-global RefExact context.
// ====================================================================================================
namespace {

vh_context* g_ctx = nullptr;
HashTableParams g_params;
bool g_haveParams = false;
float g_K[9], g_Kinv[9];
bool g_haveK = false;
float* g_err = nullptr;          // device float for computeCorrespondences
float* g_sys300 = nullptr;

[[noreturn]] void legacyDie(const char* what, const char* file, int line) {
    // behaviour of checkCudaErrors, cuda_helper/helper_cuda.h:965-981
    fprintf(stderr, "CUDA error at %s:%d \"%s\" : %s\n", file, line, what, vh_last_error());
    cudaDeviceReset();
    exit(EXIT_FAILURE);
}
#define LEGACY(expr) do { if ((expr) != VH_OK) legacyDie(#expr, __FILE__, __LINE__); } while (0)
#define LEGACY_CUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { fail(VH_ERR_CUDA, #expr, _e); legacyDie(#expr, __FILE__, __LINE__); } } while (0)

void legacyCreate(const HashTableParams& p) {
    vh_config cfg;
    vh_default_config(&cfg);
    cfg.table = p;
    cfg.policy = VH_POLICY_REF_EXACT;
    if (g_ctx) { vh_destroy(g_ctx); g_ctx = nullptr; }
    LEGACY(vh_create(&cfg, &g_ctx));
    if (g_haveK) vh_set_intrinsic_matrices(g_ctx, g_K, g_Kinv);
    if (!g_err) LEGACY_CUDA(cudaMalloc((void**)&g_err, sizeof(float)));
}

vh_context* legacyCtx() {
    if (!g_ctx) {
        vh_config cfg;
        vh_default_config(&cfg);
        legacyCreate(g_haveParams ? g_params : cfg.table);
    }
    return g_ctx;
}

void legacyPushPose(const HashTableParams& p) {
    // both matrices exactly as the host supplied them (SDF_Hashtable.cpp:17-21)
    vh_context* c = legacyCtx();
    FrameParams fp;
    memcpy(fp.pose, p.global_transform.entries, sizeof(fp.pose));
    memcpy(fp.inv, p.inv_global_transform.entries, sizeof(fp.inv));
    LEGACY_CUDA(cudaMemcpy(c->frame, &fp, sizeof(fp), cudaMemcpyHostToDevice));
}

}  // namespace

extern "C" {

void updateConstantHashTableParams(const HashTableParams& params) {
    g_params = params;
    g_haveParams = true;
    if (g_ctx) legacyPushPose(params);
}

void deviceAllocate(const HashTableParams& params) {
    g_params = params;
    g_haveParams = true;
    legacyCreate(params);
    legacyPushPose(params);
}

void deviceFree(void) {
    if (g_ctx) { vh_destroy(g_ctx); g_ctx = nullptr; }
    if (g_err) { cudaFree(g_err); g_err = nullptr; }
    if (g_sys300) { cudaFree(g_sys300); g_sys300 = nullptr; }
}

void resetHashTableMutexes(const HashTableParams& /*params*/) {
    LEGACY_CUDA(launch_reset_mutex(legacyCtx(), 0));
    LEGACY_CUDA(cudaStreamSynchronize(0));
}

void allocBlocks(const float4* verts, const float4* /*normals*/) {
    vh_context* c = legacyCtx();
    LEGACY_CUDA(launch_alloc(c, verts, 0));
    LEGACY_CUDA(cudaStreamSynchronize(0));         // ref :715
}

int flattenIntoBuffer(const HashTableParams& /*params*/) {
    vh_context* c = legacyCtx();
    LEGACY(vh_compact(c, 0));
    int n = 0;
    LEGACY_CUDA(cudaMemcpy(&n, &c->v.ctr->compactCount, sizeof(int), cudaMemcpyDeviceToHost));   // ref :765
    return n;
}

void calculateKinectProjectionMatrix(void) {
    // The fusion-side matrix is float3x3(intrinsicsTranspose) (quirk Q1); RefExact kernels build it from
    // the context's fx, fy, cx, cy, so there is nothing to upload.  Kept for link compatibility.
    legacyCtx();
}

void integrateDepthMap(const HashTableParams& params, const float4* verts) {
    vh_context* c = legacyCtx();
    if (params.numOccupiedBlocks > 0) {            // ref :848
        LEGACY_CUDA(cudaMemsetAsync(&c->v.ctr->numUpdated, 0, sizeof(unsigned long long), 0));
        LEGACY_CUDA(launch_integrate(c, verts, nullptr, (int)params.numOccupiedBlocks, 0));
        LEGACY_CUDA(cudaStreamSynchronize(0));     // ref :850
    }
}

void mapGLobjectsToCUDApointers(struct cudaGraphicsResource*, struct cudaGraphicsResource*, struct cudaGraphicsResource*) {
    legacyCtx();   // headless: library-owned buffers stand in for the three GL buffers
}

void preProcess(float4* positions, float4* normals, const uint16_t* depth) {
    vh_context* c = legacyCtx();
    LEGACY_CUDA(launch_preprocess(c, depth, positions, normals, nullptr, 0));
    LEGACY_CUDA(cudaStreamSynchronize(0));         // ref :118
}

bool SetCameraIntrinsic(const float* intrinsic, const float* invIntrinsic) {
    memcpy(g_K, intrinsic, sizeof(g_K));
    memcpy(g_Kinv, invIntrinsic, sizeof(g_Kinv));
    g_haveK = true;
    if (g_ctx) vh_set_intrinsic_matrices(g_ctx, g_K, g_Kinv);
    return true;
}

float computeCorrespondences(const float4* d_input, const float4* d_target, const float4* d_targetNormals, float4* corres,
                             float4* corresNormals, float* residuals, const float4x4 deltaTransform, const int width,
                             const int height) {
    vh_context* c = legacyCtx();
    if (width != c->v.W || height != c->v.H) { fail(VH_ERR_INVALID, "computeCorrespondences: image size differs from the context's"); legacyDie("size", __FILE__, __LINE__); }
    LEGACY_CUDA(launch_find_corr(c, d_input, nullptr, d_target, d_targetNormals, deltaTransform.entries, corres, corresNormals,
                                 residuals, g_err, 0));
    float e = 0.f;
    LEGACY_CUDA(cudaMemcpy(&e, g_err, sizeof(float), cudaMemcpyDeviceToHost));    // ref :212
    return e;
}

void CalculateJacobiansAndResiduals(const float4* /*d_src*/, const float4* d_targ, const float4* d_targNormals, float* d_Jac) {
    vh_context* c = legacyCtx();
    LEGACY_CUDA(launch_jacobians(c, d_targ, d_targNormals, d_Jac, 0));            // ref launches async too (Solver.cu:68)
}

void buildLinearSystemOnDevice(const float4* d_input, const float4* d_correspondence, const float4* d_correspondenceNormals,
                               float* d_out, float* h_out) {
    vh_context* c = legacyCtx();
    LEGACY_CUDA(launch_linear_system_300(c, d_input, d_correspondence, d_correspondenceNormals, d_out, 0));
    LEGACY_CUDA(cudaStreamSynchronize(0));                                         // ref :98
    const int blocks = (c->v.W * c->v.H + 1023) / 1024;
    if (h_out) LEGACY_CUDA(cudaMemcpy(h_out, d_out, (size_t)blocks * 27 * sizeof(float), cudaMemcpyDeviceToHost));   // ref :99
}

const VoxelEntry* vhLegacyCompactTable(void) { return legacyCtx()->v.compact20; }
const Voxel* vhLegacyVoxelBlocks(void) { return legacyCtx()->v.voxels; }
const int* vhLegacyCompactCounter(void) { return &legacyCtx()->v.ctr->compactCount; }
struct vh_context* vhLegacyContext(void) { return legacyCtx(); }

}  // extern "C"
