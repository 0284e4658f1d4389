/* vh/types.h -- layout-compatible plain-data types of the voxel-hashing hot path.
 *
 * Every type here has the same size, alignment and field offsets as the type of the
 * same name in the reference, so the reference's host code (SDF_Hashtable.cpp,
 * CameraTracking.cpp, Solver.cpp, Application.cpp) can be compiled against this
 * header instead of its own and linked against libvh_b200.so unchanged.
 *
 *   Voxel            8 B  {float sdf; float weight;}           ref VoxelDataStructures.h:11-17
 *   VoxelEntry      20 B  {int3 pos; int ptr; int offset;}     ref VoxelDataStructures.h:19-26
 *   HashTableParams 176 B                                       ref VoxelDataStructures.h:28-52
 *   float3x3 / float4x4  row-major, M*v                         ref cuda_helper/cuda_SimpleMatrixUtil.h:311,802
 *
 * The __align__ attributes the reference writes in front of `struct` bind to nothing
 * under gcc/nvcc on Linux, hence 8/20/176 bytes with
 * 4-byte alignment; the static_asserts at the bottom pin that.
 *
 * Usable from C (plain structs), C++ and CUDA.
 */
#ifndef VH_TYPES_H
#define VH_TYPES_H

#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__)
#  include <cuda_runtime.h>
#  define VH_HD __host__ __device__ inline
#else
#  define VH_HD inline
#  if defined(__has_include)
#    if __has_include(<vector_types.h>)
#      include <vector_types.h>
#      define VH_HAVE_CUDA_VECTOR_TYPES 1
#    endif
#  endif
#  ifndef VH_HAVE_CUDA_VECTOR_TYPES
/* Stand-ins with CUDA's layout for builds that have no CUDA headers on the include path. */
#    ifndef __VECTOR_TYPES_H__
#    define __VECTOR_TYPES_H__
typedef struct int2   { int x, y; } int2;
typedef struct int3   { int x, y, z; } int3;
typedef struct float3 { float x, y, z; } float3;
#      if defined(__cplusplus)
struct alignas(16) float4 { float x, y, z, w; };
#      else
typedef struct float4 { float x, y, z, w; } __attribute__((aligned(16))) float4;
#      endif
#    endif
#  endif
#endif

/* Sentinels of VoxelEntry::ptr (ref VoxelUtils.cu:19-21). */
#define VH_FREE_BLOCK   (-1)
#define VH_LOCKED_BLOCK (-2)
#define VH_NO_OFFSET    0
/* Free slots carry pos = (INT_MAX,INT_MAX,INT_MAX): the reference stores float +inf
 * through make_int3(float...) which saturates on the device (quirk Q8, VoxelUtils.cu:155-157). */
#define VH_FREE_COORD   2147483647

#define VH_BLOCK_SIDE        8
#define VH_VOXELS_PER_BLOCK  512

#ifdef __cplusplus

/* ----------------------------------------------------------------------------------------
 * float3x3: row-major, (M*v)_i = sum_j m_ij v_j summed left to right
 * (ref cuda_SimpleMatrixUtil.h:311-320 ctor, :482-488 operator*).
 * -------------------------------------------------------------------------------------- */
class float3x3 {
public:
    union {
        struct { float m11, m12, m13, m21, m22, m23, m31, m32, m33; };
        float entries[9];
        float entries2[3][3];
    };

    VH_HD float3x3() {}
    VH_HD float3x3(const float v[9]) { for (int i = 0; i < 9; ++i) entries[i] = v[i]; }
    VH_HD float3x3(const float3x3& o) { for (int i = 0; i < 9; ++i) entries[i] = o.entries[i]; }
    VH_HD float3x3& operator=(const float3x3& o) {
        for (int i = 0; i < 9; ++i) entries[i] = o.entries[i];
        return *this;
    }
    VH_HD float& operator()(int r, int c) { return entries2[r][c]; }
    VH_HD float operator()(int r, int c) const { return entries2[r][c]; }
    VH_HD void setIdentity() {
        for (int i = 0; i < 9; ++i) entries[i] = (i % 4 == 0) ? 1.0f : 0.0f;
    }
    VH_HD void transpose() {
        for (int r = 0; r < 3; ++r)
            for (int c = r + 1; c < 3; ++c) {
                float t = entries2[r][c]; entries2[r][c] = entries2[c][r]; entries2[c][r] = t;
            }
    }
    VH_HD float3x3 getTranspose() const { float3x3 t(*this); t.transpose(); return t; }
    VH_HD float3 operator*(const float3& v) const {
        float3 r;
        r.x = m11 * v.x + m12 * v.y + m13 * v.z;
        r.y = m21 * v.x + m22 * v.y + m23 * v.z;
        r.z = m31 * v.x + m32 * v.y + m33 * v.z;
        return r;
    }
    VH_HD const float* ptr() const { return entries; }
    VH_HD float* ptr() { return entries; }
};

/* ----------------------------------------------------------------------------------------
 * float4x4: row-major rigid/projective transform, translation in m14,m24,m34
 * (ref cuda_SimpleMatrixUtil.h:802-896).  The copy constructor is user-provided ON
 * PURPOSE: it makes the class non-trivially-copyable for the Itanium C++ ABI, so a
 * by-value float4x4 argument (computeCorrespondences, ref CameraTracking.cpp:15) is passed
 * by hidden reference exactly as the reference's own class is.
 * -------------------------------------------------------------------------------------- */
class float4x4 {
public:
    union {
        struct { float m11, m12, m13, m14, m21, m22, m23, m24,
                       m31, m32, m33, m34, m41, m42, m43, m44; };
        float entries[16];
        float entries2[4][4];
    };

    VH_HD float4x4() {}
    VH_HD float4x4(const float v[16]) { for (int i = 0; i < 16; ++i) entries[i] = v[i]; }
    VH_HD float4x4(const float4x4& o) { for (int i = 0; i < 16; ++i) entries[i] = o.entries[i]; }
    VH_HD float4x4& operator=(const float4x4& o) {
        for (int i = 0; i < 16; ++i) entries[i] = o.entries[i];
        return *this;
    }
    VH_HD float& operator()(int r, int c) { return entries2[r][c]; }
    VH_HD float operator()(int r, int c) const { return entries2[r][c]; }
    VH_HD void setIdentity() {
        for (int i = 0; i < 16; ++i) entries[i] = (i % 5 == 0) ? 1.0f : 0.0f;
    }
    VH_HD void transpose() {
        for (int r = 0; r < 4; ++r)
            for (int c = r + 1; c < 4; ++c) {
                float t = entries2[r][c]; entries2[r][c] = entries2[c][r]; entries2[c][r] = t;
            }
    }
    VH_HD float4x4 getTranspose() const { float4x4 t(*this); t.transpose(); return t; }

    /* (M*v): four products per row, summed left to right (ref :888-896). */
    VH_HD float4 operator*(const float4& v) const {
        float4 r;
        r.x = m11 * v.x + m12 * v.y + m13 * v.z + m14 * v.w;
        r.y = m21 * v.x + m22 * v.y + m23 * v.z + m24 * v.w;
        r.z = m31 * v.x + m32 * v.y + m33 * v.z + m34 * v.w;
        r.w = m41 * v.x + m42 * v.y + m43 * v.z + m44 * v.w;
        return r;
    }
    VH_HD float4x4 operator*(const float4x4& o) const {
        float4x4 r;
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j)
                r.entries2[i][j] = entries2[i][0] * o.entries2[0][j] + entries2[i][1] * o.entries2[1][j] +
                                   entries2[i][2] * o.entries2[2][j] + entries2[i][3] * o.entries2[3][j];
        return r;
    }

    /* Adjugate / determinant inverse.  Cofactor C(r,c) is expanded over the remaining 3x3
     * minor in the same six-term order the reference uses (cuda_SimpleMatrixUtil.h:944-1069),
     * so for a rigid pose the result is bit-identical to viewMat.getInverse() there
     * (SDF_Hashtable.cpp:15).  Written as an index-driven expansion, not 16 literals. */
    VH_HD float4x4 getInverse() const {
        float adj[16];
        for (int r = 0; r < 4; ++r) {
            for (int c = 0; c < 4; ++c) {
                /* adj[r][c] = (-1)^(r+c) * minor(c, r): rows != c, cols != r. */
                int R[3], C[3];
                for (int i = 0, k = 0; i < 4; ++i) if (i != c) R[k++] = i;
                for (int j = 0, k = 0; j < 4; ++j) if (j != r) C[k++] = j;
#define VH_E(i, j) entries2[R[i]][C[j]]
                float pos0 = VH_E(0, 0) * VH_E(1, 1) * VH_E(2, 2);
                float neg0 = VH_E(0, 0) * VH_E(1, 2) * VH_E(2, 1);
                float neg1 = VH_E(1, 0) * VH_E(0, 1) * VH_E(2, 2);
                float pos1 = VH_E(1, 0) * VH_E(0, 2) * VH_E(2, 1);
                float pos2 = VH_E(2, 0) * VH_E(0, 1) * VH_E(1, 2);
                float neg2 = VH_E(2, 0) * VH_E(0, 2) * VH_E(1, 1);
#undef VH_E
                float minor;
                if (((r + c) & 1) == 0) minor = pos0 - neg0 - neg1 + pos1 + pos2 - neg2;
                else                    minor = -pos0 + neg0 + neg1 - pos1 - pos2 + neg2;
                adj[r * 4 + c] = minor;
            }
        }
        float det = entries[0] * adj[0] + entries[1] * adj[4] + entries[2] * adj[8] + entries[3] * adj[12];
        float rdet = 1.0f / det;
        float4x4 out;
        for (int i = 0; i < 16; ++i) out.entries[i] = adj[i] * rdet;
        return out;
    }
    VH_HD void invert() { *this = getInverse(); }
    VH_HD const float* ptr() const { return entries; }
    VH_HD float* ptr() { return entries; }

    static VH_HD float4x4 identity() { float4x4 m; m.setIdentity(); return m; }
};

typedef float4x4 vh_float4x4;
typedef float3x3 vh_float3x3;

#else  /* plain C view of the same bytes */

typedef struct vh_float3x3 { float entries[9]; }  vh_float3x3;
typedef struct vh_float4x4 { float entries[16]; } vh_float4x4;

#endif /* __cplusplus */

/* One TSDF sample (ref VoxelDataStructures.h:11-17).  512 per block, index z*64+y*8+x
 * (ref VoxelUtils.cu:312-317). */
typedef struct Voxel {
    float sdf;
    float weight;
} Voxel;

/* One hash slot (ref VoxelDataStructures.h:19-26).  ptr = voxel index of the block start
 * (heap id * 512), VH_FREE_BLOCK when unused.  offset = relative index of the next entry
 * of the bucket's overflow chain, 0 = none. */
typedef struct VoxelEntry {
    int3 pos;
    int ptr;
    int offset;
} VoxelEntry;

/* All run-time parameters of the table plus both poses (ref VoxelDataStructures.h:28-52). */
typedef struct HashTableParams {
    vh_float4x4 global_transform;      /* camera -> world of the frame being fused */
    vh_float4x4 inv_global_transform;  /* world -> camera */

    unsigned int numBuckets;
    unsigned int bucketSize;
    unsigned int attachedLinkedListSize;
    unsigned int numVoxelBlocks;

    int voxelBlockSize;
    float voxelSize;
    unsigned int numOccupiedBlocks;

    float maxIntegrationDistance;
    float truncScale;
    float truncation;

    unsigned int integrationWeightSample;
    float integrationWeightMax;
} HashTableParams;

#ifdef __cplusplus
static_assert(sizeof(Voxel) == 8 && alignof(Voxel) == 4, "Voxel layout (ref Q14)");
static_assert(sizeof(VoxelEntry) == 20 && alignof(VoxelEntry) == 4, "VoxelEntry layout (ref Q14)");
static_assert(offsetof(VoxelEntry, ptr) == 12 && offsetof(VoxelEntry, offset) == 16, "VoxelEntry offsets");
static_assert(sizeof(float4x4) == 64 && sizeof(float3x3) == 36, "matrix sizes");
static_assert(sizeof(HashTableParams) == 176, "HashTableParams size");
static_assert(offsetof(HashTableParams, numBuckets) == 128, "HashTableParams.numBuckets");
static_assert(offsetof(HashTableParams, voxelBlockSize) == 144, "HashTableParams.voxelBlockSize");
static_assert(offsetof(HashTableParams, truncation) == 164, "HashTableParams.truncation");
static_assert(offsetof(HashTableParams, integrationWeightMax) == 172, "HashTableParams.integrationWeightMax");
#endif

#endif /* VH_TYPES_H */
