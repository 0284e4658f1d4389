// GPU-box micro-benchmark of the per-iteration 6x6 solve pieces (one warp, clock64).
// nvcc -std=c++17 -O3 -fmad=false -gencode arch=compute_100a,code=sm_100a -I include -I voxelhashing_demo_b200/csrc tools/solve_micro.cu -o /tmp/solve_micro
#include <cstdio>
#include "icp_device.cuh"
using namespace vh;

__global__ void k(const float* sysIn, long long* out, float* sink) {
    __shared__ float sSys[32];
    __shared__ float sDelta[16];
    __shared__ float sRows[16][33];
    __shared__ double sP[16];
    const int lane = threadIdx.x & 31;
    sSys[lane] = sysIn[lane];
    if (lane < 16) sDelta[lane] = (lane % 5 == 0) ? 1.f : 0.f;
    for (int g = 0; g < 16; ++g) sRows[g][lane] = sysIn[lane] * (1.0f / 16.0f);
    __syncwarp();
    float acc = 0.f;
    for (int rep = 0; rep < 4; ++rep) {
        long long t0 = clock64();
        float tw[6];
        bool ok = solveTwistWarp(sSys, true, tw);
        long long t1 = clock64();
        float u = expElementWarp(tw);
        long long t2 = clock64();
        float dcol[4];
        for (int kk = 0; kk < 4; ++kk) dcol[kk] = sDelta[kk * 4 + (lane & 3)];
        float d = updateFp32Warp(u, dcol);
        __syncwarp();
        if (lane < 16) sDelta[lane] = d;
        __syncwarp();
        long long t3 = clock64();
        double dc[4];
        for (int kk = 0; kk < 4; ++kk) dc[kk] = (double)sDelta[kk * 4 + (lane & 3)];
        double p = updateFp64Warp(u, dc, sP);
        long long t4 = clock64();
        acc += (float)p + (ok ? tw[0] : 0.f);
        if (lane == 0) { out[rep * 4 + 0] = t1 - t0; out[rep * 4 + 1] = t2 - t1; out[rep * 4 + 2] = t3 - t2; out[rep * 4 + 3] = t4 - t3; }
        sSys[lane] += acc * 1e-20f;
        __syncwarp();
    }
    sink[lane] = acc;
}

int main() {
    float h[32] = {0};
    // a well-conditioned SPD system: diag-dominant JtJ (upper triangle row by row), Jtr
    int k2 = 0;
    for (int i = 0; i < 6; ++i) for (int j = i; j < 6; ++j) h[k2++] = (i == j) ? 1000.f + 10.f * i : 3.f + i - j;
    for (int i = 0; i < 6; ++i) h[21 + i] = 0.01f * (i + 1);
    h[27] = 1.f; h[28] = 100000.f;
    float *d, *sink; long long* o;
    cudaMalloc(&d, sizeof(h)); cudaMalloc(&sink, 128); cudaMalloc(&o, 16 * 8);
    cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
    k<<<1, 32>>>(d, o, sink);
    long long ho[16];
    cudaMemcpy(ho, o, sizeof(ho), cudaMemcpyDeviceToHost);
    for (int r = 0; r < 4; ++r) printf("rep %d: twist solve %lld cycles, exp element %lld, fp32 update %lld, fp64 update + Newton-Schulz %lld\n", r, ho[r*4], ho[r*4+1], ho[r*4+2], ho[r*4+3]);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
