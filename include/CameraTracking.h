/* CameraTracking.h -- headless, B200-native replacement of the reference's tracking facade.
 * Same public method surface as ref CameraTracking.h:53-58: CameraTracking(int,int), ~CameraTracking,
 * Align(float4*, float4*, float4*, float4*, const uint16_t*, const uint16_t*), getTransform().
 * Align runs maxIters fused Gauss-Newton iterations on the device (association + residual + 27-sum
 * reduction + 6x6 solve + SE(3) update per launch), with no host round trip between iterations
 * (the reference: 5 syncs + 3 blocking D2H per iteration).
 */
#ifndef CAMERA_TRACKING_H
#define CAMERA_TRACKING_H

#include <stdint.h>

#include "EigenUtil.h"
#include "Solver.h"
#include "vh/abi.h"

class CameraTracking {
public:
    CameraTracking(int width, int height);                       /* reference intrinsics (common.h:7-10), RefExact */
    CameraTracking(int width, int height, vh_context* shared);   /* track with the fusion context (any policy) */
    ~CameraTracking();
    CameraTracking(const CameraTracking&) = delete;
    CameraTracking& operator=(const CameraTracking&) = delete;

    /* ref CameraTracking.cpp:26-69.  The two depth pointers are unused, as in the reference. */
    void Align(float4* d_input, float4* d_inputNormals, float4* d_target, float4* d_targetNormals,
               const uint16_t* d_depthInput, const uint16_t* d_depthTarget);
    Matrix4x4f getTransform();                                   /* input frame -> target frame; synchronises */

    /* additions */
    void setStream(vh_stream s) { stream_ = s; }
    void setMaxIters(int n) { maxIters = n; }
    void resetEstimate();                                        /* the reference never resets it (quirk Q24) */
    vh_context* context() const { return ctx_; }

private:
    int width, height;
    int maxIters = 20;                                           /* ref CameraTracking.h:40 */
    vh_context* ctx_;
    bool ownsCtx_;
    vh_stream stream_;
};

#endif
