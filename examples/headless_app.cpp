// examples/headless_app.cpp -- the reference's host loop (Application.cpp:24-103), headless.
// Same objects, same call order: CameraTracking + SDF_Hashtable, preProcess on two depth frames,
// Align, getTransform, integrate.  Inputs are two synthetic frames (plane z = 2.5 m + sphere) instead
// of assets/T0.png / T1.png, which the reference does not ship.
//   usage: vh_headless_app [dump.txt]
//          vh_headless_app --frames a.png b.png [...] [--mesh out.ply]
//                                                          16-bit depth PNG / PGM files (TUM convention, 5000 per metre),
//                                                          read like the reference's stbi_load_16 (Application.cpp:28-29)
//                                                          and pushed through the native frame pipeline
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "CameraTracking.h"
#include "SDF_Hashtable.h"

extern "C" void preProcess(float4* positions, float4* normals, const uint16_t* depth);   // Application.cpp:22

static std::vector<uint16_t> renderFrame(float camX) {
    const float fx = 517.3f, fy = 516.5f, cx = 318.6f, cy = 255.3f;
    std::vector<uint16_t> d(640 * 480);
    for (int v = 0; v < 480; ++v)
        for (int u = 0; u < 640; ++u) {
            double dx = (u - cx) / fx, dy = (v - cy) / fy, best = 2.5;   // plane z = 2.5
            double ox = camX, oz = -2.0;                                 // sphere centre (0,0,2) r 0.5, camera at (camX,0,0)
            double a = dx * dx + dy * dy + 1.0, b = 2.0 * (dx * ox + oz), c = ox * ox + oz * oz - 0.25, disc = b * b - 4 * a * c;
            if (disc >= 0) { double s = (-b - std::sqrt(disc)) / (2 * a); if (s > 0 && s < best) best = s; }
            d[v * 640 + u] = (uint16_t)std::lround(best * 5000.0);
        }
    return d;
}

// --frames: every file is one frame of a sequence; track frame-to-frame, fuse, print the camera poses.
static int runSequence(int n, char** files, const char* meshPath) {
    vh_config cfg;
    vh_default_config(&cfg);
    cfg.policy = VH_POLICY_FIXED;
    cfg.table.numBuckets = 100003; cfg.table.numVoxelBlocks = 65536; cfg.table.truncation = 0.06f; cfg.icpNormalThres = 0.8f;
    vh_context* ctx = nullptr;
    vh_pipeline* pipe = nullptr;
    uint16_t* d_depth = nullptr;
    for (int i = 0; i < n; ++i) {
        uint16_t* img = nullptr;
        int w = 0, h = 0;
        if (vh_depth_read(files[i], &img, &w, &h) != VH_OK) { std::fprintf(stderr, "%s: %s\n", files[i], vh_depth_last_error()); return 1; }
        if (!ctx) {                                       // the first frame fixes the image size (intrinsics scale with it)
            const float sx = w / 640.0f, sy = h / 480.0f;
            cfg.width = w; cfg.height = h; cfg.fx *= sx; cfg.cx *= sx; cfg.fy *= sy; cfg.cy *= sy;
            if (vh_create(&cfg, &ctx) != VH_OK || vh_pipeline_create(ctx, 20, VH_TRACK_FRAME_TO_FRAME, VH_PIPE_GRAPH | VH_PIPE_OVERLAP, &pipe) != VH_OK) {
                std::fprintf(stderr, "%s\n", vh_last_error());
                return 1;
            }
            vh_pipeline_reset(pipe, nullptr, nullptr);
            cudaMalloc(&d_depth, sizeof(uint16_t) * w * h);
        } else if (w != cfg.width || h != cfg.height) {
            std::fprintf(stderr, "%s: %dx%d, expected %dx%d\n", files[i], w, h, cfg.width, cfg.height);
            return 1;
        }
        cudaMemcpy(d_depth, img, sizeof(uint16_t) * w * h, cudaMemcpyHostToDevice);
        vh_depth_free(img);
        if (vh_pipeline_push_device(pipe, d_depth, nullptr) != VH_OK) return 1;
        float T[16];
        vh_pipeline_pose(pipe, T, nullptr);
        std::printf("frame %d  t = (% .5f % .5f % .5f)\n", i, T[3], T[7], T[11]);
    }
    vh_stats st;
    vh_get_stats(ctx, &st, nullptr);
    std::printf("allocated blocks %d, visible %d, dropped %d\n", st.numAllocated, st.numVisible, st.dropped);
    if (meshPath) {                                       // zero level set of the fused model as a binary PLY
        int nt = 0;
        vh_pipeline_flush(pipe, nullptr);
        vh_extract_mesh(ctx, nullptr, 0, &nt, nullptr);
        float* d_tris = nullptr;
        std::vector<float> h((size_t)nt * 9);
        cudaMalloc(&d_tris, sizeof(float) * 9 * (nt > 0 ? nt : 1));
        vh_extract_mesh(ctx, d_tris, nt, &nt, nullptr);
        cudaMemcpy(h.data(), d_tris, sizeof(float) * 9 * nt, cudaMemcpyDeviceToHost);
        cudaFree(d_tris);
        if (vh_save_mesh_ply(meshPath, h.data(), nt) != VH_OK) return 1;
        std::printf("mesh: %d triangles -> %s\n", nt, meshPath);
    }
    vh_pipeline_destroy(pipe);
    vh_destroy(ctx);
    cudaFree(d_depth);
    return 0;
}

int main(int argc, char** argv) {
    if (argc > 2 && std::string(argv[1]) == "--frames") {
        const char* mesh = nullptr;
        int n = argc - 2;
        if (n > 2 && std::string(argv[argc - 2]) == "--mesh") { mesh = argv[argc - 1]; n -= 2; }
        return runSequence(n, argv + 2, mesh);
    }
    const size_t px = 640 * 480;
    CameraTracking tracker(640, 480);                    // Application.cpp:32
    SDF_Hashtable fusionModule;                          // :33
    std::vector<uint16_t> img1 = renderFrame(0.0f), img2 = renderFrame(0.01f);
    uint16_t *d_depthInput, *d_depthTarget;
    float4 *d_input, *d_inputNormals, *d_target, *d_targetNormals;
    cudaMalloc(&d_depthInput, px * 2); cudaMalloc(&d_depthTarget, px * 2);
    cudaMalloc(&d_input, px * 16); cudaMalloc(&d_inputNormals, px * 16);
    cudaMalloc(&d_target, px * 16); cudaMalloc(&d_targetNormals, px * 16);
    cudaMemcpy(d_depthInput, img1.data(), px * 2, cudaMemcpyHostToDevice);      // :40-43
    cudaMemcpy(d_depthTarget, img2.data(), px * 2, cudaMemcpyHostToDevice);
    preProcess(d_input, d_inputNormals, d_depthInput);                          // :73
    preProcess(d_target, d_targetNormals, d_depthTarget);                       // :74
    tracker.Align(d_input, d_inputNormals, d_target, d_targetNormals, d_depthInput, d_depthTarget);   // :75 (commented out there)
    Matrix4x4f T = tracker.getTransform();                                      // :76
    std::printf("Final rigid transform (input -> target):\n");
    for (int r = 0; r < 4; ++r) std::printf("  % .6f % .6f % .6f % .6f\n", T(r, 0), T(r, 1), T(r, 2), T(r, 3));
    float4x4 identity;
    identity.setIdentity();
    fusionModule.integrate(identity, d_input, d_inputNormals);                  // :82-84
    std::printf("occupiedBlockCount : %d\n", fusionModule.occupiedBlockCount());   // SDF_Hashtable.cpp:31
    if (argc > 1 && vh_dump_text(fusionModule.context(), argv[1]) != VH_OK) return 1;   // :85 printSDFdata
    // the target camera sits 1 cm to the right of the input camera: input -> target translation is -1 cm in x
    bool ok = std::fabs(T(0, 3) + 0.01f) < 2e-3f && fusionModule.occupiedBlockCount() == 199;
    std::printf("%s\n", ok ? "OK" : "MISMATCH");
    return ok ? 0 : 2;
}
