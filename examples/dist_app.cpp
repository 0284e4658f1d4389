// examples/dist_app.cpp -- the multi-GPU frame loop from C++ only (no Python, no NCCL): one process per GPU, the
// block-hash space partitioned over the ranks, frames broadcast and ICP systems all-reduced over CUDA IPC peer memory
// (include/vh/abi.h, vh_dist_*).  What a C++ host built on the reference's Application.cpp would add to scale out.
//
//   vh_dist_app <frames.bin> <width> <height> <frames> <world> <out-prefix> [iterations]
//     frames.bin : <frames> raw u16 depth images (5000 units per metre), row-major
//     forks <world> processes (rank r -> GPU r); the 64-byte IPC handles travel through <out-prefix>.handle.<r> files
//     every rank writes <out-prefix>.pose.<r> (16 floats, row-major camera -> world of the last frame) and
//     <out-prefix>.stats.<r> (allocated / visible / dropped, text)
#include <sys/wait.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "vh/abi.h"

#define MUST(expr) do { if ((expr) != VH_OK) { std::fprintf(stderr, "rank %d: %s failed: %s\n", rank, #expr, vh_last_error()); return 1; } } while (0)

static bool readAll(const std::string& path, void* dst, size_t bytes) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    const bool ok = fread(dst, 1, bytes, f) == bytes;
    fclose(f);
    return ok;
}
static bool writeAll(const std::string& path, const void* src, size_t bytes) {
    const std::string tmp = path + ".tmp";
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) return false;
    const bool ok = fwrite(src, 1, bytes, f) == bytes;
    fclose(f);
    return ok && rename(tmp.c_str(), path.c_str()) == 0;      // atomic: a reader sees the whole handle or nothing
}

static int runRank(int rank, int world, const char* framesPath, int W, int H, int frames, const std::string& prefix, int iterations) {
    if (cudaSetDevice(rank) != cudaSuccess) { std::fprintf(stderr, "rank %d: no GPU %d\n", rank, rank); return 1; }
    vh_config cfg;
    vh_default_config(&cfg);
    cfg.policy = VH_POLICY_FIXED;
    cfg.width = W; cfg.height = H;
    cfg.fx *= W / 640.0f; cfg.cx *= W / 640.0f; cfg.fy *= H / 480.0f; cfg.cy *= H / 480.0f;
    cfg.table.numBuckets = 100003; cfg.table.numVoxelBlocks = 16384; cfg.table.truncation = 0.06f; cfg.overflowSlots = 8192;
    cfg.icpNormalThres = 0.8f; cfg.icpIterations = iterations;
    cfg.partCount = world; cfg.partRank = rank;
    vh_context* ctx = nullptr;
    vh_pipeline* pipe = nullptr;
    vh_dist* dist = nullptr;
    MUST(vh_create(&cfg, &ctx));
    MUST(vh_dist_create(ctx, rank, world, &dist));
    // exchange the IPC handles through files (any out-of-band channel works)
    const size_t hb = (size_t)vh_dist_handle_bytes();
    std::vector<unsigned char> mine(hb), all(hb * world);
    MUST(vh_dist_export(dist, mine.data()));
    if (!writeAll(prefix + ".handle." + std::to_string(rank), mine.data(), hb)) return 1;
    for (int p = 0; p < world; ++p)
        for (int tries = 0; !readAll(prefix + ".handle." + std::to_string(p), all.data() + hb * p, hb); ++tries) {
            if (tries > 20000) { std::fprintf(stderr, "rank %d: no handle from rank %d\n", rank, p); return 1; }
            usleep(1000);
        }
    MUST(vh_dist_connect(dist, all.data()));
    MUST(vh_pipeline_create(ctx, iterations, VH_TRACK_FRAME_TO_FRAME, VH_PIPE_GRAPH | VH_PIPE_OVERLAP, &pipe));
    cudaStream_t s;
    cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    MUST(vh_pipeline_reset(pipe, nullptr, s));

    const size_t px = (size_t)W * H;
    uint16_t* d_frames = nullptr;
    if (rank == 0) {                                          // the ingest rank holds the sequence
        std::vector<uint16_t> h(px * frames);
        if (!readAll(framesPath, h.data(), h.size() * sizeof(uint16_t))) { std::fprintf(stderr, "cannot read %s\n", framesPath); return 1; }
        cudaMalloc((void**)&d_frames, h.size() * sizeof(uint16_t));
        cudaMemcpy(d_frames, h.data(), h.size() * sizeof(uint16_t), cudaMemcpyHostToDevice);
    }
    for (int k = 0; k < frames; ++k) {
        const uint16_t* frame = nullptr;
        void* ready = nullptr;
        MUST(vh_dist_broadcast_frame(dist, rank == 0 ? d_frames + px * k : nullptr, &frame, &ready, s));
        MUST(vh_pipeline_push_device_ready(pipe, frame, ready, s));   // s is ordered behind the frame's pose -> behind its pre-processing
        MUST(vh_dist_frame_consumed(dist, s));
    }
    float pose[16];
    MUST(vh_pipeline_pose(pipe, pose, s));                    // synchronises (and includes the last fusion)
    vh_stats st;
    MUST(vh_get_stats(ctx, &st, s));
    if (!writeAll(prefix + ".pose." + std::to_string(rank), pose, sizeof(pose))) return 1;
    char line[128];
    std::snprintf(line, sizeof(line), "%d %d %d\n", st.numAllocated, st.numVisible, st.dropped);
    if (!writeAll(prefix + ".stats." + std::to_string(rank), line, std::strlen(line))) return 1;
    vh_pipeline_destroy(pipe);
    vh_dist_destroy(dist);
    vh_destroy(ctx);
    cudaFree(d_frames);
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 7) { std::fprintf(stderr, "usage: %s frames.bin width height frames world out-prefix [iterations]\n", argv[0]); return 2; }
    const int W = atoi(argv[2]), H = atoi(argv[3]), frames = atoi(argv[4]), world = atoi(argv[5]);
    const std::string prefix = argv[6];
    const int iterations = argc > 7 ? atoi(argv[7]) : 20;
    std::vector<pid_t> kids;
    for (int r = 0; r < world; ++r) {                         // fork BEFORE any CUDA call: every child initialises its own GPU
        pid_t pid = fork();
        if (pid == 0) _exit(runRank(r, world, argv[1], W, H, frames, prefix, iterations));
        kids.push_back(pid);
    }
    int rc = 0;
    for (pid_t pid : kids) {
        int status = 0;
        waitpid(pid, &status, 0);
        if (!WIFEXITED(status) || WEXITSTATUS(status) != 0) rc = 1;
    }
    return rc;
}
