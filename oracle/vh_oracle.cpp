// oracle/vh_oracle.cpp -- CPU restatement of the reference's fusion-and-tracking hot path.
//
// TEST INFRASTRUCTURE, NOT PRODUCT (see vh_oracle.h).  Scalar C++ with optional OpenMP on the
// embarrassingly parallel loops; doubles as the "host transliteration" CPU baseline of
// BASELINE.md section 3.2.
//
// Build: g++ -O2 -ffp-contract=off -fno-fast-math -fopenmp  (no FMA contraction: the reference is
// compiled with -fmad=false, CMakeLists.txt:23; every fused multiply-add below is an explicit
// fmaf and only appears in the Fixed policy, whose arithmetic is defined by DESIGN.md).
//
// Policy RefExact follows SURVEY.md Appendix A line by line in meaning; each function cites
// the reference file:line it restates.  Policy Fixed is the corrected pipeline (correct K,
// metric inverse pose, truncation-band allocation, overflow chain) that the tracking loop uses.

#include "vh_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <tuple>
#include <unordered_map>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

constexpr int kFree = -1;          // VoxelUtils.cu:19
constexpr int kLockedMutex = -2;   // VoxelUtils.cu:20
constexpr int kIntMax = std::numeric_limits<int>::max();
constexpr int kIntMin = std::numeric_limits<int>::min();

// ---- device conversion semantics ----------------------------------------------------------
// cvt.rzi.s32.f32: toward zero, saturating, NaN -> 0 (SURVEY Appendix A preamble).
inline int f2i(float x) {
    if (std::isnan(x)) return 0;
    if (x >= 2147483648.0f) return kIntMax;
    if (x <= -2147483648.0f) return kIntMin;
    return (int)x;
}
// cvt.rzi.s32.f64
// cvt.rni.s32.f32: round to nearest even, saturating, NaN -> 0
inline int f2i_rn(float x) {
    if (std::isnan(x)) return 0;
    if (x >= 2147483648.0f) return kIntMax;
    if (x <= -2147483648.0f) return kIntMin;
    return (int)std::nearbyintf(x);
}
// pixel rounding of the Fixed integration: nearest, ties to even, via the 1.5*2^23 trick
// (bit-identical to nearbyintf for |u| < 2^22, far outside any image beyond that);
// the integration's own form: nearest integer of a * r with ONE rounding (the product is never rounded to float)
inline int roundPixelFma(float a, float r) {
    float t = fmaf(a, r, 12582912.0f);
    int32_t bits;
    std::memcpy(&bits, &t, sizeof(bits));
    return (int)(bits - 0x4B400000);
}
inline int d2i(double x) {
    if (std::isnan(x)) return 0;
    if (x >= 2147483648.0) return kIntMax;
    if (x <= -2147483649.0) return kIntMin;
    return (int)x;
}

struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };
struct I3 {
    int x, y, z;
    bool operator==(const I3& o) const { return x == o.x && y == o.y && z == o.z; }
    bool operator<(const I3& o) const { return std::tie(x, y, z) < std::tie(o.x, o.y, o.z); }
};
struct I3Hash {
    size_t operator()(const I3& k) const {
        uint64_t h = (uint32_t)k.x * 0x9E3779B97F4A7C15ull;
        h ^= ((uint64_t)(uint32_t)k.y + 0x7F4A7C15u) * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2);
        h ^= ((uint64_t)(uint32_t)k.z + 0x165667B1u) * 0x165667B19E3779F9ull + (h << 6) + (h >> 2);
        return (size_t)h;
    }
};

// Row-major M*v, products summed left to right (cuda_SimpleMatrixUtil.h:888-896).
inline V4 mul4(const float* m, V4 v) {
    V4 r;
    r.x = m[0] * v.x + m[1] * v.y + m[2] * v.z + m[3] * v.w;
    r.y = m[4] * v.x + m[5] * v.y + m[6] * v.z + m[7] * v.w;
    r.z = m[8] * v.x + m[9] * v.y + m[10] * v.z + m[11] * v.w;
    r.w = m[12] * v.x + m[13] * v.y + m[14] * v.z + m[15] * v.w;
    return r;
}
// cuda_SimpleMatrixUtil.h:482-488
inline V3 mul3(const float* m, V3 v) {
    V3 r;
    r.x = m[0] * v.x + m[1] * v.y + m[2] * v.z;
    r.y = m[3] * v.x + m[4] * v.y + m[5] * v.z;
    r.z = m[6] * v.x + m[7] * v.y + m[8] * v.z;
    return r;
}

// 4x4 inverse by adjugate, term order of cuda_SimpleMatrixUtil.h:944-1069.
void mat4_inverse(const float* e, float* out) {
    float inv[16];
    auto minor3 = [&](int r0, int r1, int r2, int c0, int c1, int c2, bool neg) {
        float a = e[r0 * 4 + c0] * e[r1 * 4 + c1] * e[r2 * 4 + c2];
        float b = e[r0 * 4 + c0] * e[r1 * 4 + c2] * e[r2 * 4 + c1];
        float c = e[r1 * 4 + c0] * e[r0 * 4 + c1] * e[r2 * 4 + c2];
        float d = e[r1 * 4 + c0] * e[r0 * 4 + c2] * e[r2 * 4 + c1];
        float f = e[r2 * 4 + c0] * e[r0 * 4 + c1] * e[r1 * 4 + c2];
        float g = e[r2 * 4 + c0] * e[r0 * 4 + c2] * e[r1 * 4 + c1];
        return neg ? (-a + b + c - d - f + g) : (a - b - c + d + f - g);
    };
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
            int R[3], C[3];
            for (int i = 0, k = 0; i < 4; ++i) if (i != c) R[k++] = i;
            for (int j = 0, k = 0; j < 4; ++j) if (j != r) C[k++] = j;
            inv[r * 4 + c] = minor3(R[0], R[1], R[2], C[0], C[1], C[2], ((r + c) & 1) != 0);
        }
    float det = e[0] * inv[0] + e[1] * inv[4] + e[2] * inv[8] + e[3] * inv[12];
    float rdet = 1.0f / det;
    for (int i = 0; i < 16; ++i) out[i] = inv[i] * rdet;
}

struct Entry { I3 pos; int ptr; int offset; };

}  // namespace

struct vo_table {
    vo_config cfg;
    std::vector<Entry> table;     // numBuckets*bucketSize + overflowSlots
    std::vector<int> mutex;       // numBuckets
    std::vector<unsigned> heap;   // numVoxelBlocks
    int heapCounter;
    int overflowUsed;
    int dropped;
    float* voxels;                // numVoxelBlocks*512*2, zeroed (Q13: the harness zeroes too)
    std::vector<Entry> compact;
    std::vector<I3> lastRequestedNew;
    long long lastUpdated;
};

namespace {

// ---- coordinate maps (VoxelUtils.cu:250-326) ------------------------------------------------

// VoxelUtils.cu:250-259.  int products wrap, XOR, then `% numBuckets` with numBuckets unsigned:
// the XOR is converted to uint32 and the modulo is unsigned (Q7); the `res < 0` fix-up is dead.
inline unsigned hashBlock(const vo_config& c, I3 p) {
    const uint32_t p0 = 73856093u, p1 = 19349669u, p2 = 83492791u;
    uint32_t h = ((uint32_t)p.x * p0) ^ ((uint32_t)p.y * p1) ^ ((uint32_t)p.z * p2);
    uint32_t res = h % c.numBuckets;
    int sres = (int)res;                       // `int res = ...`
    if (sres < 0) sres += (int)c.numBuckets;   // dead for numBuckets < 2^31
    return (unsigned)sres;
}

// VoxelUtils.cu:280-287: p/size, copysignf(1,.) -> int, *0.5 (double) -> float, add, trunc.
inline I3 world2Voxel(const vo_config& c, V3 p) {
    float qx = p.x / c.voxelSize, qy = p.y / c.voxelSize, qz = p.z / c.voxelSize;
    int sx = f2i(std::copysign(1.0f, qx)), sy = f2i(std::copysign(1.0f, qy)), sz = f2i(std::copysign(1.0f, qz));
    float ox = (float)(sx * 0.5), oy = (float)(sy * 0.5), oz = (float)(sz * 0.5);
    return I3{f2i(qx + ox), f2i(qy + oy), f2i(qz + oz)};
}
// VoxelUtils.cu:266-278: floor division by 8 via "subtract 7 when negative", C truncation.
inline I3 voxel2Block(I3 v) {
    const int size = 8;
    // int arithmetic wraps on the device; mirror with unsigned math
    auto fd = [&](int a) {
        if (a < 0) a = (int)((uint32_t)a - (uint32_t)(size - 1));
        return a / size;
    };
    return I3{fd(v.x), fd(v.y), fd(v.z)};
}
inline I3 world2Block(const vo_config& c, V3 p) { return voxel2Block(world2Voxel(c, p)); }
// VoxelUtils.cu:289-304
inline V3 block2World(const vo_config& c, I3 b) {
    I3 v{(int)((uint32_t)b.x * 8u), (int)((uint32_t)b.y * 8u), (int)((uint32_t)b.z * 8u)};
    return V3{(float)v.x * c.voxelSize, (float)v.y * c.voxelSize, (float)v.z * c.voxelSize};
}

// Fusion-side projection Kt = float3x3(intrinsicsTranspose) read row-major (Q1, VoxelUtils.cu:224-231):
// rows (fx,0,0) (0,fy,0) (cx,cy,1).
inline void fusionKt(const vo_config& c, float* kt) {
    const float fx = c.K[0], fy = c.K[4], cx = c.K[2], cy = c.K[5];
    const float v[9] = {fx, 0, 0, 0, fy, 0, cx, cy, 1};
    std::memcpy(kt, v, sizeof(v));
}

// VoxelUtils.cu:344-359 (Q1, Q2): min corner, camera->world transform, transposed K.
inline bool blockInFrustumRef(const vo_config& c, const float* pose, const float* kt, I3 b) {
    V3 w = block2World(c, b);
    V4 p = mul4(pose, V4{w.x, w.y, w.z, 1.0f});
    V3 r = mul3(kt, V3{p.x, p.y, p.z});
    float rx = r.x / r.z, ry = r.y / r.z;
    int x = f2i(rx), y = f2i(ry);
    return x < c.width && x >= 0 && y < c.height && y >= 0;
}

// ---- Fixed policy pieces (defined by DESIGN.md, mirrored expression for expression by the kernels) --

struct FixedFrustum { float fx, fy, cx, cy, wr, hb, nl, nr, nt, nb, rad; };
inline FixedFrustum makeFrustum(const vo_config& c) {
    FixedFrustum f;
    f.fx = c.K[0]; f.fy = c.K[4]; f.cx = c.K[2]; f.cy = c.K[5];
    f.wr = (float)(c.width - 1) - f.cx;
    f.hb = (float)(c.height - 1) - f.cy;
    f.nl = sqrtf(f.fx * f.fx + f.cx * f.cx);
    f.nr = sqrtf(f.fx * f.fx + f.wr * f.wr);
    f.nt = sqrtf(f.fy * f.fy + f.cy * f.cy);
    f.nb = sqrtf(f.fy * f.fy + f.hb * f.hb);
    f.rad = c.voxelSize * 6.9282032f;   // half diagonal of an 8^3 block: 4*sqrt(3) voxels
    return f;
}
inline bool blockVisibleFixed(const vo_config& c, const FixedFrustum& f, const float* invPose, I3 b) {
    float cx = ((float)(b.x * 8) + 3.5f) * c.voxelSize;
    float cy = ((float)(b.y * 8) + 3.5f) * c.voxelSize;
    float cz = ((float)(b.z * 8) + 3.5f) * c.voxelSize;
    V4 p = mul4(invPose, V4{cx, cy, cz, 1.0f});
    const float r = f.rad;
    if (!(p.z + r > c.depthMin)) return false;
    if (!(p.z - r < c.depthMax)) return false;
    if (!(f.fx * p.x + f.cx * p.z > -(r * f.nl))) return false;
    if (!(f.wr * p.z - f.fx * p.x > -(r * f.nr))) return false;
    if (!(f.fy * p.y + f.cy * p.z > -(r * f.nt))) return false;
    if (!(f.hb * p.z - f.fy * p.y > -(r * f.nb))) return false;
    return true;
}

inline unsigned ownerOf(I3 b, int parts) {
    uint32_t u = ((uint32_t)b.x * 0x9E3779B1u) ^ ((uint32_t)b.y * 0x85EBCA77u) ^ ((uint32_t)b.z * 0xC2B2AE3Du);
    u ^= u >> 16; u *= 0x7FEB352Du; u ^= u >> 15; u *= 0x846CA68Bu; u ^= u >> 16;
    return u % (uint32_t)parts;
}

// ---- table operations -------------------------------------------------------------------------

inline int popHeap(vo_table* t) {
    // VoxelUtils.cu:328-334: addr = atomicSub(counter,1) (old value); heap[addr].  Negative addr is an
    // out-of-bounds read in the reference (Q6); here it reports exhaustion.
    int addr = t->heapCounter;
    t->heapCounter = addr - 1;
    if (addr < 0) return -1;
    return (int)t->heap[addr];
}

// VoxelUtils.cu:418-456 (RefExact).  Returns 1 inserted, 0 otherwise.
int insertRef(vo_table* t, I3 key) {
    const vo_config& c = t->cfg;
    unsigned h = hashBlock(c, key);
    unsigned start = h * c.bucketSize;
    for (unsigned i = 0; i < c.bucketSize; ++i) {
        unsigned idx = (start + i) % (c.numBuckets * c.bucketSize);
        Entry& e = t->table[idx];
        if (e.pos == key && e.ptr != kFree) return 0;
        if (e.ptr == kFree) {
            int prev = t->mutex[h];
            t->mutex[h] = kLockedMutex;           // atomicExch, never released this frame (Q4)
            if (prev != kLockedMutex) {
                e.pos = key;
                e.offset = 0;
                int id = popHeap(t);
                if (id < 0) { t->dropped++; return 0; }   // pos stays written, ptr stays -1 (Q6)
                e.ptr = id * 512;
                return 1;
            }
        }
    }
    return 0;
}

// Fixed: bucket slots, then the overflow chain hanging off the bucket's last slot.
// Slot states as in k_alloc.cu: never used {INT_MAX^3, -1}, live, tombstone {old key, -1} (left by garbage
// collection).  The key goes into the FIRST claimable slot of the scan order, after the scan has shown it is not
// stored further on; a never-used slot ends the scan (slots fill in scan order).
// Returns 1 inserted, 0 present, -1 dropped.
int insertFixed(vo_table* t, I3 key) {
    const vo_config& c = t->cfg;
    const unsigned S = c.numBuckets * c.bucketSize;
    unsigned h = hashBlock(c, key);
    unsigned start = h * c.bucketSize;
    int claim = -1;
    bool ended = false;
    for (unsigned i = 0; i < c.bucketSize && !ended; ++i) {
        const Entry& e = t->table[start + i];
        if (e.ptr != kFree) { if (e.pos == key) return 0; continue; }
        if (claim < 0) claim = (int)(start + i);
        ended = e.pos.x == kIntMax;
    }
    unsigned cur = start + c.bucketSize - 1;
    unsigned len = 0;
    if (!ended) {
        while (t->table[cur].offset != 0) {
            cur = cur + (unsigned)t->table[cur].offset;
            ++len;
            const Entry& e = t->table[cur];
            if (e.ptr != kFree) { if (e.pos == key) return 0; }
            else if (claim < 0) claim = (int)cur;
        }
    }
    if (claim >= 0) {
        Entry& e = t->table[claim];
        int id = popHeap(t);
        if (id < 0) {                                   // heap empty: the claimed slot ALWAYS becomes a tombstone {key, FREE}, as on the device
            t->heapCounter++; t->dropped++;
            e.pos = key;
            return -1;
        }
        e.pos = key; e.ptr = id * 512;
        return 1;
    }
    if (len >= c.attachedLinkedListSize || (unsigned)t->overflowUsed >= c.overflowSlots) { t->dropped++; return -1; }
    unsigned slot = S + (unsigned)t->overflowUsed++;
    t->table[slot].pos = key; t->table[slot].ptr = kFree; t->table[slot].offset = 0;
    t->table[cur].offset = (int)(slot - cur);
    int id = popHeap(t);
    if (id < 0) { t->heapCounter++; t->dropped++; return -1; }     // linked tombstone, as on the device
    t->table[slot].ptr = id * 512;
    return 1;
}

// lookup shared by both policies (VoxelUtils.cu:362-414 getVoxelEntry4Block + chain)
const Entry* findEntry(const vo_table* t, I3 key) {
    const vo_config& c = t->cfg;
    unsigned h = hashBlock(c, key);
    unsigned start = h * c.bucketSize;
    for (unsigned i = 0; i < c.bucketSize; ++i) {
        const Entry& e = t->table[start + i];
        if (e.pos == key && e.ptr != kFree) return &e;
    }
    if (c.policy == VO_POLICY_FIXED) {
        unsigned cur = start + c.bucketSize - 1;
        while (t->table[cur].offset != 0) {
            cur = cur + (unsigned)t->table[cur].offset;
            if (t->table[cur].pos == key && t->table[cur].ptr != kFree) return &t->table[cur];
        }
    }
    return nullptr;
}

}  // namespace

extern "C" {

int vo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void vo_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

// VoxelUtils.cu:170-211 (deviceAllocate) + :151-166 (reset kernels)
void vo_reset(vo_table* t) {
    const vo_config& c = t->cfg;
    for (auto& e : t->table) { e.pos = I3{kIntMax, kIntMax, kIntMax}; e.ptr = kFree; e.offset = 0; }  // Q8
    std::fill(t->mutex.begin(), t->mutex.end(), 0);
    for (unsigned i = 0; i < c.numVoxelBlocks; ++i) t->heap[i] = i;
    t->heapCounter = (int)c.numVoxelBlocks - 1;   // VoxelUtils.cu:207
    t->overflowUsed = 0;
    t->dropped = 0;
    t->compact.clear();
    t->lastRequestedNew.clear();
    t->lastUpdated = 0;
    std::memset(t->voxels, 0, (size_t)c.numVoxelBlocks * 512 * 2 * sizeof(float));
}

vo_table* vo_create(const vo_config* cfg) {
    vo_table* t = new vo_table();
    t->cfg = *cfg;
    if (t->cfg.policy != VO_POLICY_FIXED) t->cfg.overflowSlots = 0;
    size_t S = (size_t)cfg->numBuckets * cfg->bucketSize + t->cfg.overflowSlots;
    t->table.resize(S);
    t->mutex.resize(cfg->numBuckets);
    t->heap.resize(cfg->numVoxelBlocks);
    t->voxels = (float*)std::calloc((size_t)cfg->numVoxelBlocks * 512 * 2, sizeof(float));
    if (!t->voxels) { delete t; return nullptr; }
    vo_reset(t);
    return t;
}

void vo_destroy(vo_table* t) {
    if (!t) return;
    std::free(t->voxels);
    delete t;
}

// ------------------------------------------------------------------------------------------------
// A.6 pre-processing -- CameraTrackingUtils.cu:50-113
// ------------------------------------------------------------------------------------------------
void vo_preprocess(const vo_config* cfg, const uint16_t* depth, float* verts, float* normals, float* depthf) {
    const vo_config& c = *cfg;
    const int W = c.width, H = c.height;
    const bool fixed = c.policy == VO_POLICY_FIXED;
    // Fixed, optional: 5x5 bilateral filter of the raw depth ahead of the maps (k_preprocess.cu k_bilateral).  The two
    // weight tables are computed exactly as vh_create computes them; taps in row-major order, sum of weights with
    // plain adds, weighted sum with one fma per tap.
    const bool smoothOn = fixed && c.bilateralSigmaSpace > 0.0f && c.bilateralSigmaRange > 0.0f;
    std::vector<float> smooth;
    if (smoothOn) {
        constexpr int kLut = 1024;
        std::vector<float> lut(kLut);
        float g[3];
        const double sr = (double)(c.bilateralSigmaRange * c.depthScale), ss = (double)c.bilateralSigmaSpace;
        for (int i = 0; i < kLut; ++i) lut[i] = (float)std::exp(-((double)i * (double)i) / (2.0 * sr * sr));
        for (int i = 0; i < 3; ++i) g[i] = (float)std::exp(-((double)i * (double)i) / (2.0 * ss * ss));
        smooth.assign((size_t)W * H, 0.0f);
#pragma omp parallel for schedule(static)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                const int d0 = depth[y * W + x];
                if (d0 == 0) continue;
                float sumW = 0.0f, sumD = 0.0f;
                for (int dy = -2; dy <= 2; ++dy) {
                    const int yy = y + dy;
                    if (yy < 0 || yy >= H) continue;
                    for (int dx = -2; dx <= 2; ++dx) {
                        const int xx = x + dx;
                        if (xx < 0 || xx >= W) continue;
                        const int dj = depth[yy * W + xx];
                        const int diff = std::abs(dj - d0);
                        if (dj == 0 || diff >= kLut) continue;
                        const float ws = g[std::abs(dy)] * g[std::abs(dx)];
                        const float w = ws * lut[diff];
                        sumW = sumW + w;
                        sumD = fmaf(w, (float)dj, sumD);
                    }
                }
                smooth[(size_t)y * W + x] = sumD / sumW;
            }
    }
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            int idx = y * W + x;
            float d = (smoothOn ? smooth[idx] : (float)depth[idx]) / c.depthScale;   // :63-64
            if (fixed && !(d > c.depthMin && d < c.depthMax)) d = 0.0f;   // Fixed: sensor range mask
            V3 ic{(float)x, (float)y, 1.0f};                              // :69
            V3 p = mul3(c.Kinv, ic);                                      // :70  K_inv*imageCoord
            verts[idx * 4 + 0] = p.x * d;
            verts[idx * 4 + 1] = p.y * d;
            verts[idx * 4 + 2] = p.z * d;
            verts[idx * 4 + 3] = 1.0f;                                    // w = 1 always (Q27)
            if (depthf) {                                                 // integration reads the RAW depth
                float dr = (float)depth[idx] / c.depthScale;
                if (fixed && !(dr > c.depthMin && dr < c.depthMax)) dr = 0.0f;
                depthf[idx] = p.z * dr;
            }
        }
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            int idx = y * W + x;
            float* n = normals + idx * 4;
            n[0] = n[1] = n[2] = n[3] = 0.0f;                             // :91
            if (!(x > 0 && x < W - 1 && y > 0 && y < H - 1)) continue;    // :93
            const float* CC = verts + (size_t)(y * W + x) * 4;
            const float* PC = verts + (size_t)((y + 1) * W + x) * 4;
            const float* CP = verts + (size_t)(y * W + x + 1) * 4;
            const float* MC = verts + (size_t)((y - 1) * W + x) * 4;
            const float* CM = verts + (size_t)(y * W + x - 1) * 4;
            bool ok;
            if (!fixed) {
                ok = CC[0] != 0 && PC[0] != 0 && CP[0] != 0 && MC[0] != 0 && CM[0] != 0;   // :100 (tests .x, Q27)
            } else {
                ok = CC[2] != 0 && PC[2] != 0 && CP[2] != 0 && MC[2] != 0 && CM[2] != 0;
                if (ok) {   // depth-discontinuity mask: neighbours within 5% of the centre depth
                    float lim = 0.05f * CC[2];
                    ok = fabsf(PC[2] - CC[2]) < lim && fabsf(MC[2] - CC[2]) < lim &&
                         fabsf(CP[2] - CC[2]) < lim && fabsf(CM[2] - CC[2]) < lim;
                }
            }
            if (!ok) continue;
            V3 a{PC[0] - MC[0], PC[1] - MC[1], PC[2] - MC[2]};
            V3 b{CP[0] - CM[0], CP[1] - CM[1], CP[2] - CM[2]};
            V3 nn{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};   // helper_math.h:1420
            float l = sqrtf(nn.x * nn.x + nn.y * nn.y + nn.z * nn.z);                     // helper_math.h:1291
            if (l > 0.0f) { n[0] = nn.x / l; n[1] = nn.y / l; n[2] = nn.z / l; n[3] = 0.0f; }   // :107-109
        }
}

// ------------------------------------------------------------------------------------------------
// A.3 allocation -- VoxelUtils.cu:606-716
// ------------------------------------------------------------------------------------------------
void vo_alloc(vo_table* t, const float* pose, const float* verts, vo_alloc_report* rep) {
    const vo_config& c = t->cfg;
    const int W = c.width, H = c.height;
    vo_alloc_report r{};
    std::fill(t->mutex.begin(), t->mutex.end(), 0);     // resetHashTableMutexes, VoxelUtils.cu:146-149
    t->lastRequestedNew.clear();
    std::unordered_map<I3, int, I3Hash> requested;      // key -> 1 if new at frame start
    const int heapBefore = t->heapCounter;
    const int droppedBefore = t->dropped;

    auto request = [&](I3 b) {
        auto it = requested.find(b);
        if (it == requested.end()) {
            bool isNew = findEntry(t, b) == nullptr;
            requested.emplace(b, isNew ? 1 : 0);
            if (isNew) t->lastRequestedNew.push_back(b);
        }
    };

    if (c.policy == VO_POLICY_REF_EXACT) {
        float kt[9];
        fusionKt(c, kt);
        for (int i = 0; i < W * H; ++i) {
            V4 v{verts[i * 4], verts[i * 4 + 1], verts[i * 4 + 2], verts[i * 4 + 3]};
            if (v.z == 0.0f) continue;                                   // :621
            V4 p = mul4(pose, v);                                        // :622
            I3 b = world2Block(c, V3{p.x, p.y, p.z});                    // :636
            if (!blockInFrustumRef(c, pose, kt, b)) continue;            // :673
            r.requestedPixels++;
            request(b);
            insertRef(t, b);                                             // :674
        }
    } else {
        const float invVs = 1.0f / c.voxelSize;
        for (int i = 0; i < W * H; ++i) {
            const float vx = verts[i * 4], vy = verts[i * 4 + 1], d = verts[i * 4 + 2];
            if (!(d > c.depthMin && d < c.depthMax)) continue;
            float tr = fmaf(c.truncScale, d, c.truncation);             // getTruncation, VoxelUtils.cu:261-264
            float s0 = (d - tr) / d, s1 = (d + tr) / d;
            V4 a = mul4(pose, V4{vx * s0, vy * s0, d * s0, 1.0f});
            V4 b = mul4(pose, V4{vx * s1, vy * s1, d * s1, 1.0f});
            // continuous block coordinates: block k spans [(8k-0.5)vs, (8k+7.5)vs]
            float ga[3] = {(a.x * invVs + 0.5f) * 0.125f, (a.y * invVs + 0.5f) * 0.125f, (a.z * invVs + 0.5f) * 0.125f};
            float gb[3] = {(b.x * invVs + 0.5f) * 0.125f, (b.y * invVs + 0.5f) * 0.125f, (b.z * invVs + 0.5f) * 0.125f};
            int cur[3], end[3], step[3];
            float tMax[3], tDelta[3];
            for (int k = 0; k < 3; ++k) {
                float fa = floorf(ga[k]), fb = floorf(gb[k]);
                cur[k] = f2i(fa); end[k] = f2i(fb);
                float dir = gb[k] - ga[k];
                if (dir > 0.0f) { step[k] = 1; tMax[k] = ((fa + 1.0f) - ga[k]) / dir; tDelta[k] = 1.0f / dir; }
                else if (dir < 0.0f) { step[k] = -1; tMax[k] = (fa - ga[k]) / dir; tDelta[k] = -1.0f / dir; }
                else { step[k] = 0; tMax[k] = INFINITY; tDelta[k] = INFINITY; }
            }
            r.requestedPixels++;
            for (int iter = 0; iter < 32; ++iter) {
                I3 blk{cur[0], cur[1], cur[2]};
                if (c.partCount <= 1 || (int)ownerOf(blk, c.partCount) == c.partRank) {
                    request(blk);
                    insertFixed(t, blk);
                }
                if (cur[0] == end[0] && cur[1] == end[1] && cur[2] == end[2]) break;
                int ax = (tMax[0] <= tMax[1] && tMax[0] <= tMax[2]) ? 0 : (tMax[1] <= tMax[2] ? 1 : 2);
                if (tMax[ax] > 1.0f) break;
                cur[ax] += step[ax];
                tMax[ax] += tDelta[ax];
            }
        }
    }

    // report (contention analysis of SURVEY 7.3 / 8c)
    std::map<unsigned, int> newPerBucket;
    for (auto& kv : requested) {
        r.requestedBlocks++;
        if (kv.second) { r.requestedNew++; newPerBucket[hashBlock(c, kv.first)]++; }
    }
    r.bucketsTouched = (int)newPerBucket.size();
    for (auto& kv : newPerBucket) {
        if (kv.second >= 2) r.bucketsContended++;
        r.maxNewPerBucket = std::max(r.maxNewPerBucket, kv.second);
    }
    r.dropped = t->dropped - droppedBefore;
    r.inserted = (heapBefore - t->heapCounter) - (c.policy == VO_POLICY_REF_EXACT ? r.dropped : 0);
    if (r.inserted < 0) r.inserted = 0;
    if (rep) *rep = r;
}

int vo_last_requested_new(vo_table* t, int* xyz, int cap) {
    int n = (int)t->lastRequestedNew.size();
    for (int i = 0; i < n && i < cap; ++i) {
        xyz[i * 3] = t->lastRequestedNew[i].x; xyz[i * 3 + 1] = t->lastRequestedNew[i].y; xyz[i * 3 + 2] = t->lastRequestedNew[i].z;
    }
    return n;
}

// ------------------------------------------------------------------------------------------------
// A.4 compaction -- VoxelUtils.cu:719-768
// ------------------------------------------------------------------------------------------------
int vo_compact(vo_table* t, const float* pose) {
    const vo_config& c = t->cfg;
    t->compact.clear();
    if (c.policy == VO_POLICY_REF_EXACT) {
        float kt[9];
        fusionKt(c, kt);
        for (const Entry& e : t->table)
            if (e.ptr != kFree && blockInFrustumRef(c, pose, kt, e.pos)) t->compact.push_back(e);   // :732
    } else {
        float inv[16];
        mat4_inverse(pose, inv);
        FixedFrustum f = makeFrustum(c);
        for (const Entry& e : t->table)
            if (e.ptr != kFree && blockVisibleFixed(c, f, inv, e.pos)) t->compact.push_back(e);
    }
    return (int)t->compact.size();
}

// ------------------------------------------------------------------------------------------------
// A.5 integration -- VoxelUtils.cu:770-852
// ------------------------------------------------------------------------------------------------
static long long integrateImpl(vo_table* t, const float* pose, const float* depthSrc, int stride, int zoff) {
    const vo_config& c = t->cfg;
    const int W = c.width, H = c.height;
    float inv[16];
    mat4_inverse(pose, inv);                 // SDF_Hashtable.cpp:15 (host getInverse)
    long long updated = 0;
    const int n = (int)t->compact.size();
    if (c.policy == VO_POLICY_REF_EXACT) {
        float kt[9];
        fusionKt(c, kt);
        const float T = c.truncation;
#pragma omp parallel for schedule(dynamic, 8) reduction(+ : updated)
        for (int bi = 0; bi < n; ++bi) {
            const Entry& e = t->compact[bi];
            for (int tz = 0; tz < 8; ++tz)
                for (int ty = 0; ty < 8; ++ty)
                    for (int tx = 0; tx < 8; ++tx) {
                        I3 vi{(int)((uint32_t)e.pos.x * 8u + (uint32_t)tx), (int)((uint32_t)e.pos.y * 8u + (uint32_t)ty),
                              (int)((uint32_t)e.pos.z * 8u + (uint32_t)tz)};                                 // :793-796
                        V4 vf = mul4(inv, V4{(float)vi.x, (float)vi.y, (float)vi.z, 1.0f});                   // :797-798 (Q9)
                        I3 vj{f2i(vf.x), f2i(vf.y), f2i(vf.z)};                                              // :799
                        V3 w{(float)vj.x * c.voxelSize, (float)vj.y * c.voxelSize, (float)vj.z * c.voxelSize};   // :800
                        V3 r = mul3(kt, w);                                                                  // :774 (Q1)
                        int px = f2i(r.x / r.z), py = f2i(r.y / r.z);                                        // :775-776
                        if (px < 0 || px >= W || py < 0 || py >= H) continue;                                // :803
                        float depth = depthSrc[(size_t)(py * W + px) * stride + zoff];                       // :805
                        if (depth <= 0) continue;                                                            // :806
                        float sdf = depth - w.z;                                                             // :813
                        if (sdf > -T) {                                                                      // :818
                            sdf = (sdf >= 0) ? fminf(T, sdf) : fmaxf(-T, sdf);                               // :819-824
                            float* vox = t->voxels + ((size_t)e.ptr + tz * 64 + ty * 8 + tx) * 2;            // :836
                            const float wu = 0.1f;                                                           // :829
                            float os = vox[0], ow = vox[1];
                            float ns = ((os * ow) + (sdf * wu)) / (ow + wu);                                 // :783
                            float nw = fminf(c.integrationWeightMax, ow + wu);                               // :784
                            vox[0] = ns; vox[1] = nw;
                            ++updated;
                        }
                    }
        }
    } else {
        const float fx = c.K[0], fy = c.K[4], cx = c.K[2], cy = c.K[5];
        const float invRange = 1.0f / (c.depthMax - c.depthMin);
        const float ws = (float)c.integrationWeightSample;
        // Niessner's weight max(ws * 1.5 * (1 - (d - dmin)/(dmax - dmin)), 1) as one FMA in d (ref VoxelUtils.cu:809-827)
        const float wA = -((ws * 1.5f) * invRange);
        const float wB = (ws * 1.5f) * (1.0f + c.depthMin * invRange);
        // K, the inverse pose and the voxel size as one 3x4 matrix, voxel index -> (u z, v z, z)
        float M[12];
        for (int col = 0; col < 4; ++col) {
            const float sc = col < 3 ? c.voxelSize : 1.0f;
            M[0 + col] = fmaf(fx, inv[0 + col], cx * inv[8 + col]) * sc;
            M[4 + col] = fmaf(fy, inv[4 + col], cy * inv[8 + col]) * sc;
            M[8 + col] = inv[8 + col] * sc;
        }
        const float zFar = fmaf(c.truncScale, c.depthMax, c.truncation) + c.depthMax;
#pragma omp parallel for schedule(dynamic, 8) reduction(+ : updated)
        for (int bi = 0; bi < n; ++bi) {
            const Entry& e = t->compact[bi];
            for (int tz = 0; tz < 8; ++tz)
                for (int ty = 0; ty < 8; ++ty)
                    for (int tx = 0; tx < 8; ++tx) {
                        // the kernel's thread owns voxels (tx & 4) .. (tx & 4) + 3 and steps in float
                        float X = (float)(int)((uint32_t)e.pos.x * 8u + (uint32_t)(tx & 4)) + (float)(tx & 3);
                        float Y = (float)(int)((uint32_t)e.pos.y * 8u + (uint32_t)ty);
                        float Z = (float)(int)((uint32_t)e.pos.z * 8u + (uint32_t)tz);
                        float pa = fmaf(M[0], X, fmaf(M[1], Y, fmaf(M[2], Z, M[3])));     // u * z
                        float pb = fmaf(M[4], X, fmaf(M[5], Y, fmaf(M[6], Z, M[7])));     // v * z
                        float pcz = fmaf(M[8], X, fmaf(M[9], Y, fmaf(M[10], Z, M[11])));  // z
                        if (!(pcz > 1e-6f && pcz < zFar)) continue;
                        float iz = 1.0f / pcz;
                        // nearest pixel, ties to even: product and 1.5*2^23 bias in one fma (a single rounding)
                        int px = roundPixelFma(pa, iz), py = roundPixelFma(pb, iz);
                        if ((unsigned)px >= (unsigned)W || (unsigned)py >= (unsigned)H) continue;
                        float d = depthSrc[(size_t)(py * W + px) * stride + zoff];
                        if (!(d > c.depthMin && d < c.depthMax)) continue;
                        float sdf = d - pcz;
                        float tr = fmaf(c.truncScale, d, c.truncation);
                        if (!(sdf > -tr)) continue;
                        sdf = fminf(sdf, tr);
                        float wu = fmaxf(fmaf(d, wA, wB), 1.0f);
                        float* vox = t->voxels + ((size_t)e.ptr + tz * 64 + ty * 8 + tx) * 2;
                        float os = vox[0], ow = vox[1];
                        float wn = ow + wu;
                        float ns = fmaf(os, ow, sdf * wu) * (1.0f / wn);
                        vox[0] = ns; vox[1] = fminf(c.integrationWeightMax, wn);
                        ++updated;
                    }
        }
    }
    t->lastUpdated = updated;
    return updated;
}

long long vo_integrate(vo_table* t, const float* pose, const float* verts) { return integrateImpl(t, pose, verts, 4, 2); }
long long vo_integrate_depthf(vo_table* t, const float* pose, const float* depthf) { return integrateImpl(t, pose, depthf, 1, 0); }

// ---- starvation + garbage collection (k_gc.cu; Niessner et al. 2013, section 4.4) ---------------------
// scope 0: blocks of the last vo_compact; scope 1: every allocated block.  Returns the number released.
int vo_garbage_collect(vo_table* t, int scope, float sdfThreshold, float weightDecay) {
    const vo_config& c = t->cfg;
    if (!(sdfThreshold > 0.0f)) sdfThreshold = fmaf(c.truncScale, c.depthMax, c.truncation);
    std::vector<size_t> slots;
    if (scope == 0) {
        for (const Entry& ce : t->compact) {
            const Entry* e = findEntry(t, ce.pos);
            if (e && e->ptr == ce.ptr) slots.push_back((size_t)(e - t->table.data()));
        }
    } else {
        for (size_t i = 0; i < t->table.size(); ++i) if (t->table[i].ptr != kFree) slots.push_back(i);
    }
    int freed = 0;
    for (size_t si : slots) {
        Entry& e = t->table[si];
        float* vox = t->voxels + (size_t)e.ptr * 2;
        float mn = INFINITY, mx = 0.0f;
        for (int k = 0; k < 512; ++k) {
            float w = vox[2 * k + 1];
            if (weightDecay > 0.0f) { w = fmaxf(w - weightDecay, 0.0f); vox[2 * k + 1] = w; }
            if (w > 0.0f) { mn = fminf(mn, fabsf(vox[2 * k])); mx = fmaxf(mx, w); }
        }
        if (mx == 0.0f || mn >= sdfThreshold) {
            std::memset(vox, 0, sizeof(float) * 1024);
            t->heap[++t->heapCounter] = (unsigned)(e.ptr / 512);      // ref removeSingleBlockInHeap :338-339
            e.ptr = kFree;                                           // tombstone: pos and chain link stay
            ++freed;
        }
    }
    t->compact.clear();
    return freed;
}

// ---- streaming (k_gc.cu k_stream_out / k_alloc.cu k_stream_in; Niessner et al. 2013, section 4.5) ------------
// entries5: x, y, z, ptr (= 512 * record index in voxelsOut), offset (0).  Returns the number of blocks moved.
int vo_stream_out(vo_table* t, const float* center, float radius, int* entries5, float* voxelsOut, int capacity) {
    const vo_config& c = t->cfg;
    const float r2 = radius * radius;
    int n = 0;
    for (Entry& e : t->table) {
        if (e.ptr == kFree || n >= capacity) continue;
        float dx = ((float)(e.pos.x * 8) + 3.5f) * c.voxelSize - center[0];
        float dy = ((float)(e.pos.y * 8) + 3.5f) * c.voxelSize - center[1];
        float dz = ((float)(e.pos.z * 8) + 3.5f) * c.voxelSize - center[2];
        if (!(dx * dx + dy * dy + dz * dz > r2)) continue;
        float* vox = t->voxels + (size_t)e.ptr * 2;
        std::memcpy(voxelsOut + (size_t)n * 1024, vox, sizeof(float) * 1024);
        int* o = entries5 + (size_t)n * 5;
        o[0] = e.pos.x; o[1] = e.pos.y; o[2] = e.pos.z; o[3] = n * 512; o[4] = 0;
        std::memset(vox, 0, sizeof(float) * 1024);
        t->heap[++t->heapCounter] = (unsigned)(e.ptr / 512);
        e.ptr = kFree;
        ++n;
    }
    t->compact.clear();
    return n;
}
// Returns the number of blocks accepted (inserted or merged).
int vo_stream_in(vo_table* t, const int* entries5, const float* voxels, int count) {
    const vo_config& c = t->cfg;
    int accepted = 0;
    for (int b = 0; b < count; ++b) {
        const int* in = entries5 + (size_t)b * 5;
        I3 key{in[0], in[1], in[2]};
        if (c.partCount > 1 && (int)ownerOf(key, c.partCount) != c.partRank) continue;
        int r = insertFixed(t, key);
        if (r < 0) continue;
        const Entry* e = findEntry(t, key);
        if (!e) continue;
        float* dst = t->voxels + (size_t)e->ptr * 2;
        const float* src = voxels + (size_t)in[3] * 2;
        if (r == 1) std::memcpy(dst, src, sizeof(float) * 1024);
        else {
            for (int k = 0; k < 512; ++k) {
                float w1 = dst[2 * k + 1], w2 = src[2 * k + 1], wn = w1 + w2;
                if (wn > 0.0f) {
                    dst[2 * k] = fmaf(dst[2 * k], w1, src[2 * k] * w2) * (1.0f / wn);
                    dst[2 * k + 1] = fminf(c.integrationWeightMax, wn);
                } else { dst[2 * k] = 0.0f; dst[2 * k + 1] = 0.0f; }
            }
        }
        ++accepted;
    }
    t->compact.clear();
    return accepted;
}

// ---- mesh extraction (k_mesh.cu): marching tetrahedra on the Kuhn subdivision -------------------------------
namespace {
inline V3 meshEdge(V3 pa, V3 pb, float sa, float sb) {
    const float t = sa / (sa - sb);
    return V3{fmaf(t, pb.x - pa.x, pa.x), fmaf(t, pb.y - pa.y, pa.y), fmaf(t, pb.z - pa.z, pa.z)};
}
inline void meshEmit(float* tris, int capacity, int& count, V3 a, V3 b, V3 c, V3 dir) {
    auto same = [](V3 p, V3 q) { return p.x == q.x && p.y == q.y && p.z == q.z; };
    if (same(a, b) || same(b, c) || same(a, c)) return;           // slivers of a surface through a grid point
    const float ux = b.x - a.x, uy = b.y - a.y, uz = b.z - a.z, vx = c.x - a.x, vy = c.y - a.y, vz = c.z - a.z;
    const float nx = uy * vz - uz * vy, ny = uz * vx - ux * vz, nz = ux * vy - uy * vx;
    if (nx * dir.x + ny * dir.y + nz * dir.z < 0.0f) std::swap(b, c);
    const int i = count++;
    if (i >= capacity) return;
    float* o = tris + (size_t)i * 9;
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = b.x; o[4] = b.y; o[5] = b.z; o[6] = c.x; o[7] = c.y; o[8] = c.z;
}
}  // namespace
// 9 floats per triangle; returns the number of triangles the model has (may exceed capacity).
int vo_extract_mesh(vo_table* t, float* tris, int capacity) {
    const vo_config& c = t->cfg;
    static const int tet[6][4] = {{0, 1, 3, 7}, {0, 1, 5, 7}, {0, 2, 3, 7}, {0, 2, 6, 7}, {0, 4, 5, 7}, {0, 4, 6, 7}};
    int count = 0;
    const float qnan = std::numeric_limits<float>::quiet_NaN();
    for (const Entry& e : t->table) {
        if (e.ptr == kFree) continue;
        const Entry* nb[8];
        for (int k = 0; k < 8; ++k)
            nb[k] = k == 0 ? &e : findEntry(t, I3{e.pos.x + (k & 1), e.pos.y + ((k >> 1) & 1), e.pos.z + ((k >> 2) & 1)});
        std::vector<float> S(729);
        for (int p = 0; p < 729; ++p) {
            const int x = p % 9, y = (p / 9) % 9, z = p / 81;
            const Entry* b = nb[(x >> 3) | ((y >> 3) << 1) | ((z >> 3) << 2)];
            float s = qnan;
            if (b) {
                const float* vox = t->voxels + ((size_t)b->ptr + ((z & 7) * 64 + (y & 7) * 8 + (x & 7))) * 2;
                if (vox[1] > 0.0f) s = vox[0];
            }
            S[p] = s;
        }
        for (int cell = 0; cell < 512; ++cell) {
            const int cx = cell & 7, cy = (cell >> 3) & 7, cz = cell >> 6;
            float s[8];
            V3 pos[8];
            bool valid = true, anyNeg = false, anyPos = false;
            for (int k = 0; k < 8; ++k) {
                const int dx = k & 1, dy = (k >> 1) & 1, dz = (k >> 2) & 1;
                s[k] = S[(cz + dz) * 81 + (cy + dy) * 9 + (cx + dx)];
                valid = valid && (s[k] == s[k]);
                anyNeg = anyNeg || s[k] < 0.0f;
                anyPos = anyPos || !(s[k] < 0.0f);
                pos[k] = V3{(float)(e.pos.x * 8 + cx + dx) * c.voxelSize, (float)(e.pos.y * 8 + cy + dy) * c.voxelSize,
                            (float)(e.pos.z * 8 + cz + dz) * c.voxelSize};
            }
            if (!valid || !anyNeg || !anyPos) continue;
            for (int ti = 0; ti < 6; ++ti) {
                int in[4], out[4], ni = 0, no = 0;
                for (int k = 0; k < 4; ++k) {
                    const int cc = tet[ti][k];
                    if (s[cc] < 0.0f) in[ni++] = cc; else out[no++] = cc;
                }
                if (ni == 0 || ni == 4) continue;
                V3 ci{0.f, 0.f, 0.f}, co{0.f, 0.f, 0.f};
                for (int k = 0; k < ni; ++k) { ci.x += pos[in[k]].x; ci.y += pos[in[k]].y; ci.z += pos[in[k]].z; }
                for (int k = 0; k < no; ++k) { co.x += pos[out[k]].x; co.y += pos[out[k]].y; co.z += pos[out[k]].z; }
                const float fi = 1.0f / (float)ni, fo = 1.0f / (float)no;
                const V3 dir{co.x * fo - ci.x * fi, co.y * fo - ci.y * fi, co.z * fo - ci.z * fi};
                auto edge = [&](int a, int b) { return a < b ? meshEdge(pos[a], pos[b], s[a], s[b]) : meshEdge(pos[b], pos[a], s[b], s[a]); };
                if (ni == 1) {
                    meshEmit(tris, capacity, count, edge(in[0], out[0]), edge(in[0], out[1]), edge(in[0], out[2]), dir);
                } else if (ni == 3) {
                    meshEmit(tris, capacity, count, edge(out[0], in[0]), edge(out[0], in[1]), edge(out[0], in[2]), dir);
                } else {
                    const V3 a = edge(in[0], out[0]), b = edge(in[0], out[1]), cq = edge(in[1], out[1]), d = edge(in[1], out[0]);
                    meshEmit(tris, capacity, count, a, b, cq, dir);
                    meshEmit(tris, capacity, count, a, cq, d, dir);
                }
            }
        }
    }
    return count;
}

// ---- export ---------------------------------------------------------------------------------------
int vo_num_allocated(vo_table* t) {
    int n = 0;
    for (const Entry& e : t->table) n += e.ptr != kFree;
    return n;
}
static int exportList(const std::vector<Entry>& v, bool onlyAllocated, int* out, int cap) {
    int n = 0;
    for (const Entry& e : v) {
        if (onlyAllocated && e.ptr == kFree) continue;
        if (n < cap) { out[n * 5] = e.pos.x; out[n * 5 + 1] = e.pos.y; out[n * 5 + 2] = e.pos.z; out[n * 5 + 3] = e.ptr; out[n * 5 + 4] = e.offset; }
        ++n;
    }
    return n;
}
int vo_export_entries(vo_table* t, int* entries5, int cap) { return exportList(t->table, true, entries5, cap); }
int vo_export_compact(vo_table* t, int* entries5, int cap) { return exportList(t->compact, false, entries5, cap); }
int vo_get_block(vo_table* t, int x, int y, int z, float* voxels1024) {
    const Entry* e = findEntry(t, I3{x, y, z});
    if (!e) return 0;
    std::memcpy(voxels1024, t->voxels + (size_t)e->ptr * 2, 1024 * sizeof(float));
    return 1;
}
int vo_heap_counter(vo_table* t) { return t->heapCounter; }
unsigned int vo_hash(const vo_config* cfg, int x, int y, int z) { return hashBlock(*cfg, I3{x, y, z}); }
void vo_world2block(const vo_config* cfg, const float* p3, int* b3) {
    I3 b = world2Block(*cfg, V3{p3[0], p3[1], p3[2]});
    b3[0] = b.x; b3[1] = b.y; b3[2] = b.z;
}
int vo_block_in_frustum(const vo_config* cfg, const float* pose, int x, int y, int z) {
    if (cfg->policy == VO_POLICY_REF_EXACT) {
        float kt[9];
        fusionKt(*cfg, kt);
        return blockInFrustumRef(*cfg, pose, kt, I3{x, y, z}) ? 1 : 0;
    }
    float inv[16];
    mat4_inverse(pose, inv);
    FixedFrustum f = makeFrustum(*cfg);
    return blockVisibleFixed(*cfg, f, inv, I3{x, y, z}) ? 1 : 0;
}
void vo_mat4_inverse(const float* m16, float* out16) { mat4_inverse(m16, out16); }

// ------------------------------------------------------------------------------------------------
// A.7 ICP -- CameraTrackingUtils.cu:122-185, Solver.cu:25-51, Solver.cpp:80-111, SE3.cpp:4-19
// ------------------------------------------------------------------------------------------------
namespace {

struct Corr { bool ok; V3 q, n, p; float d; float qw, nw; };

// One pixel of FindCorrespondences.  RefExact: :147-182 verbatim in meaning (Q20, Q21).
inline Corr associate(const vo_config& c, const float* input, const float* inputNormals, const float* target,
                      const float* targetNormals, const float* delta, int idx) {
    Corr r{};
    const int W = c.width, H = c.height;
    const float* s = input + (size_t)idx * 4;
    if (c.policy == VO_POLICY_REF_EXACT) {
        if (!(s[2] != 0)) return r;                                             // :153
        V4 p = mul4(delta, V4{s[0], s[1], s[2], 1.0f});                         // :154-155
        V3 sp = mul3(c.K, V3{p.x, p.y, p.z});                                   // :124
        int ix = d2i((double)(sp.x / sp.z) + 0.5), iy = d2i((double)(sp.y / sp.z) + 0.5);   // :128 (Q20)
        if (!(ix > 0 && iy > 0 && ix < W && iy < H)) return r;                  // :162
        const float* q = target + (size_t)(iy * W + ix) * 4;
        const float* n = targetNormals + (size_t)(iy * W + ix) * 4;
        V3 diff{p.x - q[0], p.y - q[1], p.z - q[2]};                            // :168
        float d = diff.x * n[0] + diff.y * n[1] + diff.z * n[2];                // :169
        if (!(d < c.icpDistThres)) return r;                                    // :170 (signed, Q21)
        r.ok = true; r.q = V3{q[0], q[1], q[2]}; r.n = V3{n[0], n[1], n[2]}; r.p = V3{p.x, p.y, p.z}; r.d = d; r.qw = q[3]; r.nw = n[3];
        return r;
    }
    // Fixed (DESIGN.md section 4 "ICP"): every product-sum is a fused chain, innermost term first
    if (!(s[2] > 0.0f)) return r;
    V4 p;
    p.x = fmaf(delta[0], s[0], fmaf(delta[1], s[1], fmaf(delta[2], s[2], delta[3])));
    p.y = fmaf(delta[4], s[0], fmaf(delta[5], s[1], fmaf(delta[6], s[2], delta[7])));
    p.z = fmaf(delta[8], s[0], fmaf(delta[9], s[1], fmaf(delta[10], s[2], delta[11])));
    p.w = 1.0f;
    if (!(p.z > 1e-6f)) return r;
    const float fx = c.K[0], fy = c.K[4], cx = c.K[2], cy = c.K[5];
    float iz = 1.0f / p.z;                                                      // correctly rounded
    float u = fmaf(p.x * iz, fx, cx), v = fmaf(p.y * iz, fy, cy);
    int ix = f2i_rn(u), iy = f2i_rn(v);                                         // nearest pixel, ties to even (as integrate)
    if ((unsigned)ix >= (unsigned)W || (unsigned)iy >= (unsigned)H) return r;
    const float* q = target + (size_t)(iy * W + ix) * 4;
    const float* n = targetNormals + (size_t)(iy * W + ix) * 4;
    if (!(q[2] > 0.0f)) return r;
    float nn = fmaf(n[0], n[0], fmaf(n[1], n[1], n[2] * n[2]));
    if (!(nn > 0.0f)) return r;
    V3 diff{p.x - q[0], p.y - q[1], p.z - q[2]};
    float e2 = fmaf(diff.x, diff.x, fmaf(diff.y, diff.y, diff.z * diff.z));
    float lim = 3.0f * c.icpDistThres;
    if (!(e2 < lim * lim)) return r;
    float d = fmaf(diff.x, n[0], fmaf(diff.y, n[1], diff.z * n[2]));
    if (!(fabsf(d) < c.icpDistThres)) return r;
    if (inputNormals && c.icpNormalThres > -1.0f) {
        const float* m = inputNormals + (size_t)idx * 4;
        float rx = fmaf(delta[0], m[0], fmaf(delta[1], m[1], delta[2] * m[2]));
        float ry = fmaf(delta[4], m[0], fmaf(delta[5], m[1], delta[6] * m[2]));
        float rz = fmaf(delta[8], m[0], fmaf(delta[9], m[1], delta[10] * m[2]));
        float cosang = fmaf(rx, n[0], fmaf(ry, n[1], rz * n[2]));
        if (!(cosang > c.icpNormalThres)) return r;
    }
    r.ok = true; r.q = V3{q[0], q[1], q[2]}; r.n = V3{n[0], n[1], n[2]}; r.p = V3{p.x, p.y, p.z}; r.d = d; r.qw = q[3]; r.nw = n[3];
    return r;
}

// Jacobian row (Solver.cu:25-37): RefExact uses the TARGET point q (Q23); Fixed the transformed source p.
inline void jacRow(const vo_config& c, const Corr& k, float* J) {
    V3 a = c.policy == VO_POLICY_REF_EXACT ? k.q : k.p;
    J[0] = k.n.x; J[1] = k.n.y; J[2] = k.n.z;
    J[3] = a.y * k.n.z - a.z * k.n.y;
    J[4] = a.z * k.n.x - a.x * k.n.z;
    J[5] = a.x * k.n.y - a.y * k.n.x;
}

void so3_hat_terms(const double w[3], double& theta, double& A, double& B, double& C) {
    double t2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    theta = std::sqrt(t2);
    if (theta < 1e-6) {   // series
        A = 1.0 - t2 / 6.0; B = 0.5 - t2 / 24.0; C = 1.0 / 6.0 - t2 / 120.0;
    } else {
        A = std::sin(theta) / theta; B = (1.0 - std::cos(theta)) / t2; C = (theta - std::sin(theta)) / (t2 * theta);
    }
}

// exp of [[w]x v; 0 0] (SE3.cpp:4-11), twist = (v, w)
void se3_exp_d(const double tw[6], double M[16]) {
    const double* v = tw; const double* w = tw + 3;
    double th, A, B, C;
    so3_hat_terms(w, th, A, B, C);
    double K[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
    double K2[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double s = 0; for (int k = 0; k < 3; ++k) s += K[i * 3 + k] * K[k * 3 + j]; K2[i * 3 + j] = s; }
    double R[9], V[9];
    for (int i = 0; i < 9; ++i) { double I = (i % 4 == 0) ? 1.0 : 0.0; R[i] = I + A * K[i] + B * K2[i]; V[i] = I + B * K[i] + C * K2[i]; }
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) M[i * 4 + j] = R[i * 3 + j];
        M[i * 4 + 3] = V[i * 3] * v[0] + V[i * 3 + 1] * v[1] + V[i * 3 + 2] * v[2];
    }
    M[12] = M[13] = M[14] = 0; M[15] = 1;
}

// log (SE3.cpp:14-19): twist = (M03, M13, M23, M21, M02, M10) of the matrix logarithm
void se3_log_d(const double M[16], double tw[6]) {
    double R[9] = {M[0], M[1], M[2], M[4], M[5], M[6], M[8], M[9], M[10]};
    double tr = R[0] + R[4] + R[8];
    double cs = std::min(1.0, std::max(-1.0, (tr - 1.0) * 0.5));
    double th = std::acos(cs);
    double w[3];
    double f;
    if (th < 1e-6) f = 0.5 + th * th / 12.0; else f = th / (2.0 * std::sin(th));
    w[0] = f * (R[7] - R[5]); w[1] = f * (R[2] - R[6]); w[2] = f * (R[3] - R[1]);
    double th2, A, B, C;
    so3_hat_terms(w, th2, A, B, C);
    double t2 = th2 * th2;
    double D = (th2 < 1e-6) ? (1.0 / 12.0 + t2 / 720.0) : (1.0 - A / (2.0 * B)) / t2;
    double K[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
    double K2[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double s = 0; for (int k = 0; k < 3; ++k) s += K[i * 3 + k] * K[k * 3 + j]; K2[i * 3 + j] = s; }
    double t[3] = {M[3], M[7], M[11]};
    for (int i = 0; i < 3; ++i) {
        double s = 0;
        for (int j = 0; j < 3; ++j) { double I = (i == j) ? 1.0 : 0.0; s += (I - 0.5 * K[i * 3 + j] + D * K2[i * 3 + j]) * t[j]; }
        tw[i] = s;
    }
    tw[3] = w[0]; tw[4] = w[1]; tw[5] = w[2];
}

// Gaussian elimination with partial pivoting, fp64 (stands in for Eigen's inverse(), Solver.cpp:109)
bool solve6(const double A[36], const double b[6], double x[6]) {
    double M[6][7];
    for (int i = 0; i < 6; ++i) { for (int j = 0; j < 6; ++j) M[i][j] = A[i * 6 + j]; M[i][6] = b[i]; }
    for (int k = 0; k < 6; ++k) {
        int p = k;
        for (int i = k + 1; i < 6; ++i) if (std::fabs(M[i][k]) > std::fabs(M[p][k])) p = i;
        if (!(std::fabs(M[p][k]) > 1e-300)) return false;
        if (p != k) for (int j = 0; j < 7; ++j) std::swap(M[p][j], M[k][j]);
        for (int i = k + 1; i < 6; ++i) {
            double f = M[i][k] / M[k][k];
            for (int j = k; j < 7; ++j) M[i][j] -= f * M[k][j];
        }
    }
    for (int i = 5; i >= 0; --i) {
        double s = M[i][6];
        for (int j = i + 1; j < 6; ++j) s -= M[i][j] * x[j];
        x[i] = s / M[i][i];
    }
    for (int i = 0; i < 6; ++i) if (!std::isfinite(x[i])) return false;
    return true;
}

}  // namespace

float vo_find_correspondences(const vo_config* cfg, const float* input, const float* inputNormals, const float* target,
                              const float* targetNormals, const float* delta, float* corr, float* corrNormals,
                              float* residuals) {
    const int N = cfg->width * cfg->height;
    if (corr) std::memset(corr, 0, (size_t)N * 4 * sizeof(float));               // :201-203 thrust::fill
    if (corrNormals) std::memset(corrNormals, 0, (size_t)N * 4 * sizeof(float));
    if (residuals) std::memset(residuals, 0, (size_t)N * sizeof(float));
    float err = 0.0f;
    for (int i = 0; i < N; ++i) {
        Corr k = associate(*cfg, input, inputNormals, target, targetNormals, delta, i);
        if (!k.ok) continue;
        err += k.d;                                                              // :175 atomicAdd (fp32)
        const int W = cfg->width;
        (void)W;
        if (corr) { corr[i * 4] = k.q.x; corr[i * 4 + 1] = k.q.y; corr[i * 4 + 2] = k.q.z; corr[i * 4 + 3] = k.qw; }   // whole float4, :176
        if (corrNormals) { corrNormals[i * 4] = k.n.x; corrNormals[i * 4 + 1] = k.n.y; corrNormals[i * 4 + 2] = k.n.z; corrNormals[i * 4 + 3] = k.nw; }
        if (residuals) residuals[i] = k.d;
    }
    return err;
}

void vo_jacobians(const vo_config* cfg, const float* corr, const float* corrNormals, float* J) {
    const int N = cfg->width * cfg->height;
    for (int i = 0; i < N; ++i) {
        const float* d = corr + (size_t)i * 4;
        const float* n = corrNormals + (size_t)i * 4;
        J[i * 6] = n[0]; J[i * 6 + 1] = n[1]; J[i * 6 + 2] = n[2];              // Solver.cu:29-31
        J[i * 6 + 3] = d[1] * n[2] - d[2] * n[1];                               // cross(d, n), Solver.cu:26
        J[i * 6 + 4] = d[2] * n[0] - d[0] * n[2];
        J[i * 6 + 5] = d[0] * n[1] - d[1] * n[0];
    }
}

void vo_icp_system_build(const vo_config* cfg, const float* input, const float* inputNormals, const float* target,
                         const float* targetNormals, const float* delta, int row0, int row1, vo_icp_system* out) {
    const int W = cfg->width;
    double acc[29];
    for (double& a : acc) a = 0;
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0, s5 = 0, s6 = 0, s7 = 0, s8 = 0, s9 = 0, s10 = 0, s11 = 0, s12 = 0, s13 = 0,
           s14 = 0, s15 = 0, s16 = 0, s17 = 0, s18 = 0, s19 = 0, s20 = 0, b0 = 0, b1 = 0, b2 = 0, b3 = 0, b4 = 0, b5 = 0, er = 0, cn = 0;
#pragma omp parallel for schedule(static) reduction(+ : s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12, s13, s14, s15, s16, s17, s18, s19, s20, b0, b1, b2, b3, b4, b5, er, cn)
    for (int i = row0 * W; i < row1 * W; ++i) {
        Corr k = associate(*cfg, input, inputNormals, target, targetNormals, delta, i);
        if (!k.ok) continue;
        float Jf[6];
        jacRow(*cfg, k, Jf);
        double J[6] = {Jf[0], Jf[1], Jf[2], Jf[3], Jf[4], Jf[5]};
        double r = k.d;
        s0 += J[0] * J[0]; s1 += J[0] * J[1]; s2 += J[0] * J[2]; s3 += J[0] * J[3]; s4 += J[0] * J[4]; s5 += J[0] * J[5];
        s6 += J[1] * J[1]; s7 += J[1] * J[2]; s8 += J[1] * J[3]; s9 += J[1] * J[4]; s10 += J[1] * J[5];
        s11 += J[2] * J[2]; s12 += J[2] * J[3]; s13 += J[2] * J[4]; s14 += J[2] * J[5];
        s15 += J[3] * J[3]; s16 += J[3] * J[4]; s17 += J[3] * J[5];
        s18 += J[4] * J[4]; s19 += J[4] * J[5];
        s20 += J[5] * J[5];
        b0 += J[0] * r; b1 += J[1] * r; b2 += J[2] * r; b3 += J[3] * r; b4 += J[4] * r; b5 += J[5] * r;
        er += r; cn += 1.0;
    }
    const double s[21] = {s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12, s13, s14, s15, s16, s17, s18, s19, s20};
    const double b[6] = {b0, b1, b2, b3, b4, b5};
    for (int i = 0; i < 21; ++i) out->JtJ[i] = (float)s[i];
    for (int i = 0; i < 6; ++i) out->Jtr[i] = (float)b[i];
    out->error = (float)er; out->count = (float)cn;
    out->pad[0] = out->pad[1] = out->pad[2] = 0;
}

int vo_icp_solve(const vo_icp_system* sys, float* estimate6, float* delta16) {
    double A[36], b[6], x[6];
    int k = 0;
    for (int i = 0; i < 6; ++i)
        for (int j = i; j < 6; ++j) { A[i * 6 + j] = A[j * 6 + i] = sys->JtJ[k++]; }   // selfadjointView, Solver.cpp:92
    for (int i = 0; i < 6; ++i) b[i] = -(double)sys->Jtr[i];                            // update = -(JTJinv*JTr), :110
    if (!solve6(A, b, x)) return 0;
    double E[16], U[16], P[16], est[6];
    for (int i = 0; i < 6; ++i) est[i] = estimate6[i];
    se3_exp_d(x, U);
    se3_exp_d(est, E);
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) { double s = 0; for (int q = 0; q < 4; ++q) s += U[i * 4 + q] * E[q * 4 + j]; P[i * 4 + j] = s; }
    se3_log_d(P, est);                                                                  // :111
    for (int i = 0; i < 6; ++i) estimate6[i] = (float)est[i];
    se3_exp_d(est, E);                                                                  // getTransform, Solver.h:32
    for (int i = 0; i < 16; ++i) delta16[i] = (float)E[i];
    return 1;
}

int vo_icp_align(const vo_config* cfg, const float* input, const float* inputNormals, const float* target,
                 const float* targetNormals, int iterations, float* estimate6, float* delta16) {
    int done = 0;
    for (int it = 0; it < iterations; ++it) {                                           // CameraTracking.cpp:35
        vo_icp_system sys;
        vo_icp_system_build(cfg, input, inputNormals, target, targetNormals, delta16, 0, cfg->height, &sys);
        if (sys.error == 0.0f) break;                                                   // :55-58 (Q22)
        if (!vo_icp_solve(&sys, estimate6, delta16)) break;
        ++done;
    }
    return done;
}

void vo_se3_exp(const float* twist6, float* m16) {
    double tw[6], M[16];
    for (int i = 0; i < 6; ++i) tw[i] = twist6[i];
    se3_exp_d(tw, M);
    for (int i = 0; i < 16; ++i) m16[i] = (float)M[i];
}
void vo_se3_log(const float* m16, float* twist6) {
    double tw[6], M[16];
    for (int i = 0; i < 16; ++i) M[i] = m16[i];
    se3_log_d(M, tw);
    for (int i = 0; i < 6; ++i) twist6[i] = (float)tw[i];
}

// ------------------------------------------------------------------------------------------------
// Fixed-policy raycast (replaces shaders/raycastSDF.frag:121-177; no reference output exists)
// ------------------------------------------------------------------------------------------------
namespace {
struct BlockCache { I3 key; const float* vox; bool valid; };

inline bool voxelAt(const vo_table* t, BlockCache& bc, int vx, int vy, int vz, float& sdf, float& w) {
    I3 b{vx >> 3, vy >> 3, vz >> 3};    // arithmetic shift = floor division by 8
    if (!(bc.valid && bc.key == b)) {
        const Entry* e = findEntry(t, b);
        bc.key = b; bc.valid = true; bc.vox = e ? t->voxels + (size_t)e->ptr * 2 : nullptr;
    }
    if (!bc.vox) return false;
    const float* v = bc.vox + (size_t)(((vz & 7) * 64) + ((vy & 7) * 8) + (vx & 7)) * 2;
    sdf = v[0]; w = v[1];
    return w > 0.0f;
}

// trilinear TSDF at world point p; false if any of the 8 taps is unobserved
inline bool sampleTrilinear(const vo_table* t, BlockCache& bc, float px, float py, float pz, float invVs, float& out) {
    float gx = px * invVs, gy = py * invVs, gz = pz * invVs;
    float fx0 = floorf(gx), fy0 = floorf(gy), fz0 = floorf(gz);
    int x0 = f2i(fx0), y0 = f2i(fy0), z0 = f2i(fz0);
    float ax = gx - fx0, ay = gy - fy0, az = gz - fz0;
    float s[8], w;
    for (int k = 0; k < 8; ++k)
        if (!voxelAt(t, bc, x0 + (k & 1), y0 + ((k >> 1) & 1), z0 + (k >> 2), s[k], w)) return false;
    float c00 = fmaf(ax, s[1] - s[0], s[0]), c10 = fmaf(ax, s[3] - s[2], s[2]);
    float c01 = fmaf(ax, s[5] - s[4], s[4]), c11 = fmaf(ax, s[7] - s[6], s[6]);
    float c0 = fmaf(ay, c10 - c00, c00), c1 = fmaf(ay, c11 - c01, c01);
    out = fmaf(az, c1 - c0, c0);
    return true;
}
}  // namespace

void vo_raycast(vo_table* t, const float* pose, float* verts, float* normals) {
    const vo_config& c = t->cfg;
    const int W = c.width, H = c.height;
    const float vs = c.voxelSize, invVs = 1.0f / vs;
    const float coarse = 4.0f * vs;
    // ray intervals from the visible blocks of THIS pose: 1/8-resolution min / max depth of the block corners
    vo_compact(t, pose);
    const int TS = 8, tw = (W + TS - 1) / TS, th = (H + TS - 1) / TS;
    std::vector<float> tileMin((size_t)tw * th, 3.3895314e38f), tileMax((size_t)tw * th, 0.0f);   // 0x7f7f7f7f, 0
    {
        float inv[16];
        mat4_inverse(pose, inv);
        const float fx = c.K[0], fy = c.K[4], cx = c.K[2], cy = c.K[5];
        for (const Entry& e : t->compact) {
            float umin = INFINITY, umax = -INFINITY, wmin = INFINITY, wmax = -INFINITY, zmin = INFINITY, zmax = -INFINITY;
            for (int k = 0; k < 8; ++k) {
                float wx = ((float)(e.pos.x * 8 + ((k & 1) ? 8 : 0)) - 0.5f) * vs;
                float wy = ((float)(e.pos.y * 8 + ((k & 2) ? 8 : 0)) - 0.5f) * vs;
                float wz = ((float)(e.pos.z * 8 + ((k & 4) ? 8 : 0)) - 0.5f) * vs;
                V4 p = mul4(inv, V4{wx, wy, wz, 1.0f});
                float zc = fmaxf(p.z, 0.05f);
                float u = p.x / zc * fx + cx, w = p.y / zc * fy + cy;
                umin = fminf(umin, u); umax = fmaxf(umax, u);
                wmin = fminf(wmin, w); wmax = fmaxf(wmax, w);
                zmin = fminf(zmin, p.z); zmax = fmaxf(zmax, p.z);
            }
            if (!(zmax > 0.0f)) continue;
            int x0 = std::min(std::max(f2i(floorf(umin)) - 1, 0), W - 1) / TS, x1 = std::min(std::max(f2i(floorf(umax)) + 2, 0), W - 1) / TS;
            int y0 = std::min(std::max(f2i(floorf(wmin)) - 1, 0), H - 1) / TS, y1 = std::min(std::max(f2i(floorf(wmax)) + 2, 0), H - 1) / TS;
            if (umax < -1.0f || wmax < -1.0f || umin > (float)W || wmin > (float)H) continue;
            float lo = fmaxf(zmin, c.depthMin);
            for (int ty = y0; ty <= y1; ++ty)
                for (int tx = x0; tx <= x1; ++tx) {
                    tileMin[(size_t)ty * tw + tx] = fminf(tileMin[(size_t)ty * tw + tx], lo);
                    tileMax[(size_t)ty * tw + tx] = fmaxf(tileMax[(size_t)ty * tw + tx], zmax);
                }
        }
    }
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            int idx = y * W + x;
            float* vo = verts + (size_t)idx * 4;
            float* no = normals + (size_t)idx * 4;
            vo[0] = vo[1] = vo[2] = 0; vo[3] = 1.0f;
            no[0] = no[1] = no[2] = no[3] = 0;
            V3 rd = mul3(c.Kinv, V3{(float)x, (float)y, 1.0f});   // camera-frame ray, z component 1
            // world ray: o + z * dw
            float dwx = pose[0] * rd.x + pose[1] * rd.y + pose[2] * rd.z;
            float dwy = pose[4] * rd.x + pose[5] * rd.y + pose[6] * rd.z;
            float dwz = pose[8] * rd.x + pose[9] * rd.y + pose[10] * rd.z;
            const float ox = pose[3], oy = pose[7], oz = pose[11];
            BlockCache bc{I3{0, 0, 0}, nullptr, false};
            const float tmin = tileMin[(size_t)(y / TS) * tw + (x / TS)], tmax = tileMax[(size_t)(y / TS) * tw + (x / TS)];
            float z = fmaxf(c.depthMin, tmin - vs), zPrev = 0, sPrev = 0;
            const float zEnd = (tmin <= tmax) ? fminf(c.depthMax, tmax + vs) : 0.0f;
            bool havePrev = false, hit = false;
            float zHit = 0;
            for (int it = 0; it < 4096 && z < zEnd; ++it) {
                float px = fmaf(z, dwx, ox), py = fmaf(z, dwy, oy), pz = fmaf(z, dwz, oz);
                float s;
                if (sampleTrilinear(t, bc, px, py, pz, invVs, s)) {
                    if (havePrev && sPrev > 0.0f && s <= 0.0f) {
                        zHit = zPrev + (z - zPrev) * (sPrev / (sPrev - s));
                        hit = true;
                        break;
                    }
                    if (havePrev && sPrev < 0.0f && s > 0.0f) { /* back face: keep marching */ }
                    zPrev = z; sPrev = s; havePrev = true;
                    z += fmaxf(vs, 0.8f * s);
                } else {
                    havePrev = false;
                    // empty: was the block at this point allocated at all?
                    float dummyS, dummyW;
                    int vx = f2i(floorf(px * invVs + 0.5f)), vy = f2i(floorf(py * invVs + 0.5f)), vz = f2i(floorf(pz * invVs + 0.5f));
                    bool near = voxelAt(t, bc, vx, vy, vz, dummyS, dummyW) || bc.vox != nullptr;
                    z += near ? vs : coarse;
                }
            }
            if (!hit) continue;
            float hx = fmaf(zHit, dwx, ox), hy = fmaf(zHit, dwy, oy), hz = fmaf(zHit, dwz, oz);
            float gxp, gxm, gyp, gym, gzp, gzm;
            bool ok = sampleTrilinear(t, bc, hx + vs, hy, hz, invVs, gxp) && sampleTrilinear(t, bc, hx - vs, hy, hz, invVs, gxm) &&
                      sampleTrilinear(t, bc, hx, hy + vs, hz, invVs, gyp) && sampleTrilinear(t, bc, hx, hy - vs, hz, invVs, gym) &&
                      sampleTrilinear(t, bc, hx, hy, hz + vs, invVs, gzp) && sampleTrilinear(t, bc, hx, hy, hz - vs, invVs, gzm);
            vo[0] = rd.x * zHit; vo[1] = rd.y * zHit; vo[2] = rd.z * zHit; vo[3] = 1.0f;
            if (!ok) continue;
            float gx = gxp - gxm, gy = gyp - gym, gz = gzp - gzm;
            // rotate the world gradient into the camera frame: R^T g
            float cxn = pose[0] * gx + pose[4] * gy + pose[8] * gz;
            float cyn = pose[1] * gx + pose[5] * gy + pose[9] * gz;
            float czn = pose[2] * gx + pose[6] * gy + pose[10] * gz;
            float l = sqrtf(cxn * cxn + cyn * cyn + czn * czn);
            if (l > 0.0f) { no[0] = cxn / l; no[1] = cyn / l; no[2] = czn / l; no[3] = 0; }
        }
}

}  // extern "C"
