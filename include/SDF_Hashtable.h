/* SDF_Hashtable.h -- headless, B200-native replacement of the reference's fusion facade.
 *
 * Same public method surface as ref SDF_Hashtable.h:24-42 (ctor, dtor, integrate,
 * registerGLtoCUDA, unmapCUDApointers) so the reference's host loop (Application.cpp:33-35,84)
 * compiles against it unchanged.  Differences, all additive:
 *   - owns a vh_context (include/vh/abi.h) instead of process-global __constant__ state, so several
 *     tables can coexist; the compact list, voxel heap and visible counter that the reference
 *     borrows from OpenGL (SDFRenderer.cpp:34-61) are library-owned device buffers;
 *   - integrate() enqueues alloc -> compact -> integrate on one stream with no intermediate host
 *     synchronisation (the reference: 4 cudaDeviceSynchronize + 2 blocking D2H) and,
 *     by default, synchronises once at the end so the call keeps the reference's blocking contract;
 *   - a second constructor takes a vh_config (Fixed policy, other image sizes, partitions).
 */
#ifndef SDF_HASHTABLE_H
#define SDF_HASHTABLE_H

#include "vh/abi.h"

class SDFRenderer;   /* OpenGL renderer of the reference; only ever named, never used, headless */

class SDF_Hashtable {
public:
    SDF_Hashtable();                              /* reference defaults (common.h:39-50), RefExact arithmetic */
    explicit SDF_Hashtable(const vh_config& cfg);
    ~SDF_Hashtable();
    SDF_Hashtable(const SDF_Hashtable&) = delete;
    SDF_Hashtable& operator=(const SDF_Hashtable&) = delete;

    /* ref SDF_Hashtable.cpp:11-40.  viewMat: camera->world, row-major; verts/normals: W*H float4 on the device. */
    void integrate(const float4x4& viewMat, const float4* verts, const float4* normals);
    void registerGLtoCUDA(SDFRenderer&) {}        /* ref :42-50 -- nothing to register headless */
    void unmapCUDApointers() {}                   /* ref :52-58 */

    /* additions */
    void setStream(vh_stream s) { stream_ = s; }
    void setSynchronous(bool on) { synchronous_ = on; }     /* default true: return after the frame is fused */
    vh_context* context() const { return ctx_; }
    int occupiedBlockCount();                     /* the value ref :30-31 prints; synchronises */
    const HashTableParams& params() const { return h_hashtableParams; }
    const VoxelEntry* compactTable() const;       /* device buffers standing in for the GL buffers */
    const Voxel* voxelBlocks() const;
    const int* compactCounter() const;

private:
    HashTableParams h_hashtableParams;
    vh_context* ctx_;
    vh_stream stream_;
    bool synchronous_;
};

#endif
