// Microbenchmark: what does HBM3e deliver for the access pattern of a tree search - random, aligned blocks of B bytes
// out of a working set much larger than L2?  (The roofline denominator in MEASURED_PEAKS.json is a sequential copy.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/hbm_random_access tools/hbm_random_access.cu
// Results on B200 (round 1) are in profiles/README.md.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint64_t mix(uint64_t x) { x += 0x9E3779B97F4A7C15ULL; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL; x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL; return x ^ (x >> 31); }
// each lane group of `lanes` threads reads one random block of lanes*16 bytes per iteration; `rmw` also writes it back
template <int LANES, bool RMW>
__global__ void k_rand(uint4 *buf, uint64_t nblocks, int iters, uint64_t seed, uint4 *sink) {
    const uint64_t gid = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) / LANES;
    const int lane = threadIdx.x % LANES;
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (int i = 0; i < iters; ++i) {
        const uint64_t b = mix(seed ^ (gid << 20) ^ i) % nblocks;
        uint4 *p = buf + b * LANES + lane;
        uint4 v = *p;
        acc.x ^= v.x; acc.y += v.y;
        if (RMW) { v.x += 1; *p = v; }
    }
    if (acc.x == 0x12345678u) sink[0] = acc;
}
template <int LANES, bool RMW> void run(uint4 *buf, size_t bytes, uint4 *sink, const char *name, int threads_per_sm = 2048) {
    const uint64_t nblocks = bytes / (LANES * 16);
    const int threads = 148 * threads_per_sm, iters = 64;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_rand<LANES, RMW><<<threads / 256, 256>>>(buf, nblocks, 4, 1, sink);
    cudaEventRecord(e0);
    k_rand<LANES, RMW><<<threads / 256, 256>>>(buf, nblocks, iters, 2, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double blocks = (double)(threads / LANES) * iters, useful = blocks * LANES * 16 * (RMW ? 2 : 1);
    printf("%-26s set %6.0f MB  %4d thr/SM  block %4d B  %8.1f M blocks/s  %7.1f GB/s useful\n", name, bytes / 1048576.0, threads_per_sm,
           LANES * 16, blocks / ms / 1e3, useful / ms / 1e6);
}
int main() {
    const size_t maxb = 32ull << 30;
    uint4 *buf, *sink; cudaMalloc(&buf, maxb); cudaMalloc(&sink, 64); cudaMemset(buf, 1, maxb);
    for (size_t mb : {256ull, 1024ull, 4096ull, 8192ull, 32768ull}) {
        const size_t bytes = mb << 20;
        run<2, false>(buf, bytes, sink, "random read");
        run<8, false>(buf, bytes, sink, "random read");
        run<16, false>(buf, bytes, sink, "random read");
        run<2, true>(buf, bytes, sink, "random read-modify-write");
    }
    for (int t : {512, 1024}) { run<2, false>(buf, 4096ull << 20, sink, "random read", t); run<16, false>(buf, 4096ull << 20, sink, "random read", t); }
    return 0;
}
